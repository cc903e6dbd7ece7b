/* `igd search` front end over the GPU library: argument scan, dispatch and output formats of the
 * reference's igd_search() (src/igd_search.c:889-1079; formats collected in SURVEY.md Appendix C), plus
 * the printing entry points get_overlaps_f1 / getOverlaps_f1 (src/igd_search.c:537-620, 721-744).
 * stdout is kept byte-compatible with the reference; anything new goes to stderr. */
#define _GNU_SOURCE
#include "igd_b200.h"
#include "igd_fmt.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define IGD_PROGRAM "igd"
#define IGD_VERSION "0.1.1"

static int search_usage(void)
{
    fprintf(stderr,
            "%s, v%s\n"
            "usage:   %s search <igd database file> [options]\n"
            "         options:\n"
            "             -q <query file>\n"
            "             -r <a region: chrN start end>\n"
            "             -v <signal value 0-1000>\n"
            "             -o <output file Name>\n"
            "             -s output Seqpare similarity\n"
            "             -f output full overlaps (for -q and -r only)\n"
            "             -m hitsmap of igd datasets\n",
            IGD_PROGRAM, IGD_VERSION, IGD_PROGRAM);
    return 0;
}

static double wall_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* --metrics <file> (extension; stdout stays the reference's): one JSON object about the search just done */
static void write_metrics(const char *path, igd_b200_t *h, const char *mode, double open_s, double search_s,
                          long long overlaps)
{
    FILE *f = fopen(path, "w");
    if (!f) { fprintf(stderr, "igd: cannot write %s\n", path); return; }
    int64_t plan[4] = {0, 0, 0, 0};
    igd_b200_last_plan(h, plan);
    fprintf(f, "{\"mode\": \"%s\", \"files\": %d, \"records\": %lld, \"tiles\": %lld, \"device_bytes\": %lld, "
               "\"open_s\": %.6f, \"search_s\": %.6f, \"overlaps\": %lld, \"overlaps_per_s\": %.1f, "
               "\"kernel_launches\": %lld, \"count_plan\": {\"segments\": %lld, \"items\": %lld, \"records\": %lld, "
               "\"bytes\": %lld}}\n",
            mode, igd_b200_nfiles(h), (long long)igd_b200_nrecords(h), (long long)igd_b200_ntiles(h),
            (long long)igd_b200_device_bytes(h), open_s, search_s, overlaps,
            search_s > 0 ? (double)overlaps / search_s : 0.0, (long long)igd_b200_launch_count(h),
            (long long)plan[0], (long long)plan[1], (long long)plan[2], (long long)plan[3]);
    fclose(f);
}

/* The banner is printed once the contig is known and the first bin exists (src/igd_search.c:540-548). */
static int banner_applies(const igd_b200_t *h, int32_t ctg, int32_t qs)
{
    if (ctg < 0) return 0;
    int32_t n1 = qs / igd_b200_nbp(h);
    return n1 <= igd_b200_ntile(h, ctg) - 1;
}

/* print one batch of -f results; returns the number of overlaps printed */
static int64_t print_full_batch(igd_b200_t *h, int64_t n, char *const *names, const int32_t *name_id,
                                const int32_t *ctg, const int32_t *qs, const int32_t *qe)
{
    igd_b200_hitlist L;
    if (igd_b200_full(h, n, ctg, qs, qe, &L)) {
        fprintf(stderr, "igd: %s\n", igd_b200_last_error());
        return 0;
    }
    /* the reference's printf calls (src/igd_search.c:545,565,592), byte for byte, through one buffer (igd_fmt.h);
     * IGD_B200_CLI_PRINTF=1 keeps printf itself */
    static char *obuf = NULL;
    const size_t ocap = 4u << 20;
    if (!obuf && !getenv("IGD_B200_CLI_PRINTF")) obuf = malloc(ocap);
    if (!obuf) {
        for (int64_t q = 0; q < n; q++) {
            if (!banner_applies(h, ctg[q], qs[q])) continue;
            printf("Query %s, %i, %i: \n", names[name_id ? name_id[q] : q], qs[q], qe[q]);
            int32_t k = 0;
            for (int64_t i = L.query_off[q]; i < L.query_off[q + 1]; i++)
                printf("%i\t %i\t %i\t %s\n", k++, L.start[i], L.end[i], igd_b200_file_name(h, L.file[i]));
        }
    } else {
        const int32_t nf = igd_b200_nfiles(h);
        size_t *flen = malloc(sizeof(size_t) * (size_t)(nf > 0 ? nf : 1));
        for (int32_t f = 0; flen && f < nf; f++) flen[f] = strlen(igd_b200_file_name(h, f));
        fmt_sink S = {obuf, obuf, ocap, stdout};
        for (int64_t q = 0; q < n; q++) {
            if (!banner_applies(h, ctg[q], qs[q])) continue;
            fmt_banner(&S, names[name_id ? name_id[q] : q], qs[q], qe[q]);
            int32_t k = 0;
            for (int64_t i = L.query_off[q]; i < L.query_off[q + 1]; i++) {
                const int32_t f = L.file[i];
                const char *fn = igd_b200_file_name(h, f);
                fmt_hit(&S, k++, L.start[i], L.end[i], fn, flen && f >= 0 && f < nf ? flen[f] : strlen(fn));
            }
        }
        fmt_flush(&S);
        free(flen);
    }
    int64_t total = L.total;
    igd_b200_hitlist_free(&L);
    return total;
}

int32_t get_overlaps_f1(char *chrm, int32_t qs, int32_t qe)
{
    igd_b200_t *h = igd_b200_current();
    if (!h) return 0;
    int32_t c = igd_b200_ctg_id(h, chrm);
    char *names[1] = {chrm};
    return (int32_t)print_full_batch(h, 1, names, NULL, &c, &qs, &qe);
}

int64_t getOverlaps_f1(char *qFile)
{
    igd_b200_t *h = igd_b200_current();
    if (!h) return 0;
    igd_b200_queries *Q = igd_b200_read_bed(qFile);
    if (!Q) return 0;
    int64_t total = 0;
    if (Q->n) {
        int32_t *ctg = malloc(sizeof(int32_t) * (size_t)Q->n);
        igd_b200_resolve(h, Q, ctg);
        const int64_t B = 1 << 18;                  /* queries per device batch */
        for (int64_t a = 0; a < Q->n; a += B) {
            int64_t m = Q->n - a < B ? Q->n - a : B;
            total += print_full_batch(h, m, Q->names, Q->name_id + a, ctg + a, Q->qs + a, Q->qe + a);
        }
        free(ctg);
    }
    igd_b200_queries_free(Q);
    return total;
}

/* gType-0 twins (src/igd_search.c:103-208, 232-250): identical output over gdata0_t records */
int32_t get_overlaps_f0(char *chrm, int32_t qs, int32_t qe) { return get_overlaps_f1(chrm, qs, qe); }
int64_t getOverlaps_f0(char *qFile) { return getOverlaps_f1(qFile); }

void seqOverlaps(char *qFile, double *sm)
{
    igd_b200_t *h = igd_b200_current();
    if (!h) return;
    if (igd_b200_seqpare(h, qFile, sm)) fprintf(stderr, "igd: %s\n", igd_b200_last_error());
}

int igd_search(int argc, char **argv)
{
    if (argc < 4) return search_usage();

    char *igdName = argv[2];
    size_t L = strlen(igdName);
    if (L < 4 || strcmp(".igd", igdName + L - 4) != 0) {
        printf("%s is not an igd database", igdName);
        return 0;
    }
    FILE *fi = fopen(igdName, "rb");
    if (!fi) {
        printf("%s does not exist", igdName);
        return 0;
    }
    fclose(fi);

    /* argument scan: later mode flags override earlier ones; -s is ignored after -r
     * (src/igd_search.c:931-971) */
    int32_t v = 0, qs = 1, qe = 2;
    int mode = -1, full = 0, device = 0;
    char out[1024] = "";
    char *chrm = NULL, *qfName = "", *metrics = NULL;
    for (int i = 3; i < argc; i++) {
        if (strcmp(argv[i], "-q") == 0) {
            if (i + 1 < argc) { qfName = argv[i + 1]; mode = 1; }
            else { printf("No query file.\n"); return 0; }
        } else if (strcmp(argv[i], "-r") == 0) {
            if (i + 3 < argc) { mode = 2; chrm = argv[i + 1]; qs = atoi(argv[i + 2]); qe = atoi(argv[i + 3]); }
        } else if (strcmp(argv[i], "-v") == 0) {
            if (i + 1 < argc) v = atoi(argv[i + 1]);
        } else if (strcmp(argv[i], "-m") == 0) {
            mode = 0;
        } else if (strcmp(argv[i], "-s") == 0 && mode != 2) {
            mode = 3;
        } else if (strcmp(argv[i], "-o") == 0) {
            if (i + 1 < argc) snprintf(out, sizeof out, "%s", argv[i + 1]);
        } else if (strcmp(argv[i], "-f") == 0) {
            full = 1;
        } else if (strcmp(argv[i], "--device") == 0) {      /* extension: CUDA device ordinal */
            if (i + 1 < argc) device = atoi(argv[i + 1]);
        } else if (strcmp(argv[i], "--metrics") == 0) {     /* extension: JSON about the run, beside stdout */
            if (i + 1 < argc) metrics = argv[i + 1];
        }
    }

    double t_open = wall_s();
    igd_b200_t *h = igd_b200_open(igdName, device);
    if (!h) {
        fprintf(stderr, "igd: %s\n", igd_b200_last_error());
        return 1;
    }
    t_open = wall_s() - t_open;
    igd_b200_use(h);
    double t_search = wall_s();
    long long found = 0;
    const char *mname = "none";
    int32_t nfiles = igd_b200_nfiles(h);
    int64_t *hits = calloc((size_t)nfiles + 1, sizeof(int64_t));

    if (full) {
        if (mode == 1) {
            int64_t total = getOverlaps_f1(qfName);
            printf("Total overlaps: %lld\n", (long long)total);
            found = total, mname = "-q -f";
        } else if (mode == 2) {
            int64_t total = get_overlaps_f1(chrm, qs, qe);
            printf("Total overlaps: %lld\n", (long long)total);
            found = total, mname = "-r -f";
        } else {
            printf("Not supported -f option\n");
        }
    } else if (mode == 0) {
        uint32_t *map = calloc((size_t)nfiles * (size_t)nfiles + 1, sizeof(uint32_t));
        int64_t pairs = igd_b200_hitsmap(h, v, map);
        if (pairs < 0) fprintf(stderr, "igd: %s\n", igd_b200_last_error());
        found = pairs, mname = "-m";
        /* the reference prints a running tile counter every 1000 tiles while it works (:783-784) */
        for (int64_t m = 1000, T = igd_b200_ntiles(h); m <= T; m += 1000) printf("%i\n", (int)m);
        if (strlen(out) < 2) strcpy(out, "Hitsmap");
        FILE *fp = fopen(out, "w");
        if (!fp) printf("Can't open file %s\n", out);
        else {
            fprintf(fp, "%u\t%u\t%u\n", nfiles, nfiles, v);
            for (int i = 0; i < nfiles; i++) {
                for (int j = 0; j < nfiles; j++) fprintf(fp, "%u\t", map[(size_t)i * nfiles + j]);
                fprintf(fp, "\n");
            }
            fclose(fp);
        }
        free(map);
    } else if (mode == 1) {
        if (search(h, qfName, v, hits) < 0) fprintf(stderr, "igd: %s\n", igd_b200_last_error());
        printf("index\t number of regions\t number of hits\t File_name\n");
        int64_t total = 0;
        for (int i = 0; i < nfiles; i++) {
            if (hits[i] > 0)
                printf("%i\t%i\t%lld\t%s\n", i, igd_b200_file_nr(h, i), (long long)hits[i], igd_b200_file_name(h, i));
            total += hits[i];
        }
        printf("Total: %lld\n", (long long)total);
        found = total, mname = "-q";
    } else if (mode == 2) {
        int32_t c = igd_b200_ctg_id(h, chrm);
        if (igd_b200_count(h, 1, &c, &qs, &qe, v, hits) < 0) fprintf(stderr, "igd: %s\n", igd_b200_last_error());
        printf("index\t number of regions\t number of hits\t File_name\n");
        for (int i = 0; i < nfiles; i++) {
            printf("%i\t%i\t%lld\t%s\n", i, igd_b200_file_nr(h, i), (long long)hits[i], igd_b200_file_name(h, i));
            found += hits[i];
        }
        mname = "-r";
    } else if (mode == 3) {
        double *sm = calloc((size_t)nfiles + 1, sizeof(double));
        seqOverlaps(qfName, sm);
        mname = "-q -s";
        printf("index\t number of regions\t similarity\t dataset name\n");
        for (int i = 0; i < nfiles; i++)
            printf("%i\t%i\t%10.6f\t%s\n", i, igd_b200_file_nr(h, i), sm[i], igd_b200_file_name(h, i));
        free(sm);
    } else {
        free(hits);
        igd_b200_close(h);
        return search_usage();
    }
    if (metrics) write_metrics(metrics, h, mname, t_open, wall_s() - t_search, found);
    free(hits);
    igd_b200_close(h);
    return 0;
}
