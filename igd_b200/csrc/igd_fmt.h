/* igd_b200 -- formatting of the CLI's -f output without printf.  The reference prints every overlap with
 * printf("%i\t %i\t %i\t %s\n", ...) (src/igd_search.c:565,592) after a banner printf("Query %s, %i, %i: \n", ...)
 * (:545); stdout is a byte contract, so the digits below are exactly %i's -- but a search that returns 50 million
 * overlaps in 15 ms should not then spend seconds in the format-string interpreter. */
#ifndef IGD_FMT_H
#define IGD_FMT_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* decimal digits of v as %i prints them; returns the end */
static inline char *fmt_i32(char *p, int32_t v)
{
    uint32_t u = (uint32_t)v;
    if (v < 0) {
        *p++ = '-';
        u = 0u - u;                             /* INT32_MIN: 2147483648 */
    }
    char t[10];
    int n = 0;
    do {
        t[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    while (n) *p++ = t[--n];
    return p;
}

typedef struct {
    char *buf, *p;
    size_t cap;
    FILE *out;
} fmt_sink;

static inline void fmt_flush(fmt_sink *s)
{
    if (s->p > s->buf) fwrite(s->buf, 1, (size_t)(s->p - s->buf), s->out);
    s->p = s->buf;
}
static inline void fmt_room(fmt_sink *s, size_t need)
{
    if ((size_t)(s->p - s->buf) + need > s->cap) fmt_flush(s);
}
/* "Query %s, %i, %i: \n" */
static inline void fmt_banner(fmt_sink *s, const char *name, int32_t qs, int32_t qe)
{
    const size_t L = strlen(name);
    if (L + 64 > s->cap) {                      /* a name longer than the buffer: the slow way */
        fmt_flush(s);
        fprintf(s->out, "Query %s, %i, %i: \n", name, qs, qe);
        return;
    }
    fmt_room(s, L + 64);
    char *p = s->p;
    memcpy(p, "Query ", 6), p += 6;
    memcpy(p, name, L), p += L;
    *p++ = ',', *p++ = ' ';
    p = fmt_i32(p, qs);
    *p++ = ',', *p++ = ' ';
    p = fmt_i32(p, qe);
    *p++ = ':', *p++ = ' ', *p++ = '\n';
    s->p = p;
}
/* "%i\t %i\t %i\t %s\n" */
static inline void fmt_hit(fmt_sink *s, int32_t k, int32_t start, int32_t end, const char *fname, size_t flen)
{
    if (flen + 64 > s->cap) {
        fmt_flush(s);
        fprintf(s->out, "%i\t %i\t %i\t %s\n", k, start, end, fname);
        return;
    }
    fmt_room(s, flen + 64);
    char *p = s->p;
    p = fmt_i32(p, k);
    *p++ = '\t', *p++ = ' ';
    p = fmt_i32(p, start);
    *p++ = '\t', *p++ = ' ';
    p = fmt_i32(p, end);
    *p++ = '\t', *p++ = ' ';
    memcpy(p, fname, flen), p += flen;
    *p++ = '\n';
    s->p = p;
}
#endif
