/* TEST INFRASTRUCTURE ONLY: drives the host C of igd_host.c (built with the stub device layer and the sanitizers):
 *   host_probe bed <file>...   -> per file: accepted lines, number of names, checksums of qs / qe / name ids
 *   host_probe open <db.igd>   -> igd_b200_open must fail cleanly (no device) with a message */
#include "igd_b200.h"
#include <stdio.h>
#include <string.h>

int main(int argc, char **argv)
{
    if (argc >= 3 && strcmp(argv[1], "bed") == 0) {
        for (int k = 2; k < argc; k++) {
            igd_b200_queries *q = igd_b200_read_bed(argv[k]);
            if (!q) { printf("%s\tunreadable\n", argv[k]); continue; }
            unsigned long long a = 0, b = 0, c = 0;
            for (int64_t i = 0; i < q->n; i++) {
                a = a * 1000003ull + (unsigned)q->qs[i];
                b = b * 1000003ull + (unsigned)q->qe[i];
                c = c * 1000003ull + (unsigned)q->name_id[i];
            }
            printf("%s\t%lld\t%d\t%llu\t%llu\t%llu\n", argv[k], (long long)q->n, q->n_names, a, b, c);
            igd_b200_queries_free(q);
        }
        return 0;
    }
    if (argc == 3 && strcmp(argv[1], "open") == 0) {
        igd_b200_t *h = igd_b200_open(argv[2], 0);
        printf("%s\t%s\n", h ? "opened" : "failed", igd_b200_last_error());
        if (h) igd_b200_close(h);
        return 0;
    }
    return 2;
}
