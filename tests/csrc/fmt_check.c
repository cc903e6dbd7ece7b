/* igd_fmt.h against printf itself: the CLI's -f lines must be the reference's bytes (tests/test_fmt.py) */
#include "igd_fmt.h"
#include <stdlib.h>

int main(void)
{
    char *a = NULL, *b = NULL;
    size_t na = 0, nb = 0;
    FILE *fa = open_memstream(&a, &na), *fb = open_memstream(&b, &nb);
    static char buf[256];                               /* a small buffer: many flushes, and the slow path for long names */
    fmt_sink S = {buf, buf, sizeof buf, fa};
    const int32_t edge[] = {0, 1, -1, 9, 10, 99, 100, 999999999, 1000000000, 2147483647, -2147483647 - 1, -1000000000, 16384, -42};
    const char *names[] = {"chr1", "", "a_rather_long_file_name.bed.gz", "x",
                           "0123456789012345678901234567890123456789012345678901234567890123456789012345678901234567890123456789"
                           "0123456789012345678901234567890123456789012345678901234567890123456789012345678901234567890123456789"
                           "01234567890123456789012345678901234567890123456789012345678901234567890123456789"};
    unsigned long long x = 88172645463325252ull;
    for (int it = 0; it < 200000; it++) {
        x ^= x << 13, x ^= x >> 7, x ^= x << 17;
        int32_t v[3];
        for (int k = 0; k < 3; k++) {
            x ^= x << 13, x ^= x >> 7, x ^= x << 17;
            v[k] = it < 14 * 14 * 14 ? edge[(it / (k == 0 ? 1 : k == 1 ? 14 : 196)) % 14] : (int32_t)(x >> (x & 31));
        }
        const char *nm = names[it % 5];
        if (it % 7 == 0) {
            fmt_banner(&S, nm, v[0], v[1]);
            fprintf(fb, "Query %s, %i, %i: \n", nm, v[0], v[1]);
        }
        fmt_hit(&S, v[2], v[0], v[1], nm, strlen(nm));
        fprintf(fb, "%i\t %i\t %i\t %s\n", v[2], v[0], v[1], nm);
    }
    fmt_flush(&S);
    fclose(fa), fclose(fb);
    const int bad = na != nb || memcmp(a, b, na) != 0;
    if (bad) printf("MISMATCH %zu %zu\n", na, nb);
    else printf("ok %zu bytes\n", na);
    free(a), free(b);
    return bad;
}
