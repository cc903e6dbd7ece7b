/* The smallest program against the library, the way INTEGRATION.md §2.2 describes it: what a caller of the
 * reference's getOverlaps() changes -- open a handle and make it current instead of running get_igdinfo /
 * get_fileinfo / fopen (src/igd_search.c:917-925,974) -- and the reference-named call itself, unchanged.
 *
 *   cc -I include examples/count_file.c -L igd_b200/lib -ligd_b200 -Wl,-rpath,$PWD/igd_b200/lib -o count_file
 *   ./count_file db.igd queries.bed [v]
 *
 * Prints `index<TAB>hits` for every file with hits and the total, or the library's message if the database cannot
 * be opened (e.g. no CUDA device: there is no CPU search path). */
#include "igd_b200.h"

#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <db.igd> <queries.bed> [v]\n", argv[0]);
        return 2;
    }
    int v = argc > 3 ? atoi(argv[3]) : 0;
    igd_b200_t *h = igd_b200_open(argv[1], 0);
    if (!h) {
        fprintf(stderr, "%s\n", igd_b200_last_error());
        return 1;
    }
    igd_b200_use(h);                                            /* the globals of src/igd.c:14-19 */
    int n = igd_b200_nfiles(h);
    int64_t *hits = calloc((size_t)n + 1, sizeof *hits);
    if (v > 0) getOverlaps_v(argv[2], hits, v);                 /* src/igd_search.h:28 */
    else getOverlaps(argv[2], hits);                            /* src/igd_search.h:27 */
    long long total = 0;
    for (int i = 0; i < n; i++) {
        if (hits[i] > 0) printf("%d\t%lld\n", i, (long long)hits[i]);
        total += hits[i];
    }
    printf("total\t%lld\n", total);
    free(hits);
    igd_b200_close(h);
    return 0;
}
