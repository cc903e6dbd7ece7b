/* `igd create` -- index build, host-side C (north_star: "Index build (igd create) stays host-side C").
 *
 * Same command line, same accepted input, same files out as the reference's igd_create()
 * (src/igd_create.c:436-502) in its four flavours -- create_igd (:25-121, the default), create_igd_f (-f, :124-243),
 * create_igd0 (-s 0, :245-332) and create_igd_bed4 (-s 2, :334-433) -- but not its mechanism.  The reference
 * grows one malloc'd list per tile, spills every 2^28 records into one temp file per tile under <out>/data0/
 * (igd_saveT, src/igd_base.c:333-364) and re-reads, sorts and appends them tile by tile (igd_save, :396-461).
 * Here the input files are parsed in parallel (one file per thread at a time) into flat per-file arrays of
 * accepted intervals; one sequential pass over them, in file order, numbers the contigs by first appearance and
 * replays the reference's batch accounting; a counting pass sizes every tile, a second pass drops each interval
 * into the tiles it touches in arrival order (what the appends to the temp files amount to) -- both passes and
 * the tile sort run on all host cores, each thread owning a range of tiles -- and the file is written once.
 *
 * The output is meant to be byte-identical to the reference's:
 *   - header, nTile[], nCnt[][] and the record payload are; that includes the ORDER OF EQUAL STARTS inside a
 *     tile, which the reference leaves to klib's unstable radix sort (sort_tile() below follows its permutation);
 *   - the 40-byte contig names are written zero-padded, where the reference writes 40 bytes starting at a
 *     strdup'd string (src/igd_base.c:422) -- whatever the heap holds behind the terminator;
 *   - stdout (progress dots, the per-batch "nCtgs, nRegions, nTiles" lines, totals) is reproduced, batch
 *     boundaries at 2^28 records included, although nothing is spilled here.
 * Deviations, all on inputs where the reference has undefined behaviour: fewer than 10 input files (the
 * reference divides by n_files/10 == 0, :81), the uninitialised `va` of -f mode (:189; written as 0), lines
 * with fewer columns than the code reads (stale pointers there; empty fields here unless a previous line left
 * text at that column, see split_line()).
 */
#define _GNU_SOURCE
#include "igd_b200.h"

#include <glob.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

#define IGD_PROGRAM "igd"
#define IGD_VERSION "0.1.1"
#define BATCH_RECORDS 268435456LL       /* maxCount, src/igd_base.h:37: the reference spills above this */

int32_t tile_size = 16384;              /* the reference's global of the same name (src/igd.c:18) */

typedef struct { int32_t idx, start, end, value; } rec_t;               /* gdata_t, src/igd_base.h:41-46 */
typedef struct { int32_t ctg, idx, start, end, value; } raw_t;          /* one accepted input interval */

/* IGD_B200_CREATE_TIMING=1: phase times on stderr (stdout stays the reference's) */
static double now_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static double g_t0;
static void phase(const char *what)
{
    if (!getenv("IGD_B200_CREATE_TIMING")) return;
    double t = now_s();
    if (what) fprintf(stderr, "igd create: %-28s %8.3f s\n", what, t - g_t0);
    g_t0 = t;
}

/* ---- string -> small integer, first-appearance numbering (contigs; dataset names of -s 2) --------------- */
typedef struct {
    char **name;
    int32_t n, cap;
    int32_t *slot;                      /* open addressing, -1 = free */
    uint32_t nslot;
} names_t;

static uint32_t str_hash(const char *s)
{
    uint32_t h = 2166136261u;
    for (; *s; s++) h = (h ^ (uint8_t)*s) * 16777619u;
    return h;
}

static void names_init(names_t *t)
{
    t->n = 0, t->cap = 64, t->nslot = 256;
    t->name = malloc(sizeof(char *) * (size_t)t->cap);
    t->slot = malloc(sizeof(int32_t) * t->nslot);
    memset(t->slot, 0xff, sizeof(int32_t) * t->nslot);
}

static void names_free(names_t *t)
{
    for (int32_t i = 0; i < t->n; i++) free(t->name[i]);
    free(t->name), free(t->slot);
}

static int32_t names_get(names_t *t, const char *s, int add)
{
    uint32_t m = t->nslot - 1, h = str_hash(s) & m;
    for (; t->slot[h] >= 0; h = (h + 1) & m)
        if (strcmp(t->name[t->slot[h]], s) == 0) return t->slot[h];
    if (!add) return -1;
    if (t->n == t->cap) t->name = realloc(t->name, sizeof(char *) * (size_t)(t->cap *= 2));
    t->name[t->n] = strdup(s);
    t->slot[h] = t->n++;
    if ((uint32_t)t->n * 2 > t->nslot) {            /* rehash at half load */
        t->nslot *= 4;
        t->slot = realloc(t->slot, sizeof(int32_t) * t->nslot);
        memset(t->slot, 0xff, sizeof(int32_t) * t->nslot);
        m = t->nslot - 1;
        for (int32_t i = 0; i < t->n; i++) {
            for (h = str_hash(t->name[i]) & m; t->slot[h] >= 0; h = (h + 1) & m) {}
            t->slot[h] = i;
        }
    }
    return t->n - 1;
}

/* ---- the build in progress ---------------------------------------------------------------------------- */
/* what one input file contributes: its accepted intervals in line order, contigs numbered locally */
typedef struct {
    raw_t *raw;                         /* .ctg = index into `ctg` below until merge_parts() makes it global */
    int64_t nraw, cap;
    names_t ctg;
    int32_t nr;                         /* lines counted for the tsv */
    double avg;                         /* sum of end - start over those lines */
    int state_first, state_last;        /* the splitter's field limit after the first / last line (-1: no line) */
    int opened;
} part_t;

typedef struct {
    int32_t nbp, gtype;
    names_t ctg;
    int32_t *mtiles;                    /* per contig: ctg_t.mTiles = 1 + the largest (end-1)/nbp seen */
    int32_t mt_cap;
    part_t *part;
    int32_t nparts;
    int64_t batch;                      /* igd_t.total: replicated records since the last "spill" */
    int64_t nrec;                       /* replicated records in all */
} build_t;

static void part_init(part_t *p)
{
    memset(p, 0, sizeof *p);
    names_init(&p->ctg);
    p->state_first = p->state_last = -1;
}

static void part_free(part_t *p)
{
    names_free(&p->ctg);
    free(p->raw);
}

static void build_init(build_t *b, int gtype, int32_t nparts)
{
    phase(NULL);
    memset(b, 0, sizeof *b);
    b->nbp = tile_size, b->gtype = gtype;
    names_init(&b->ctg);
    b->mt_cap = 64;
    b->mtiles = calloc((size_t)b->mt_cap, sizeof(int32_t));
    b->nparts = nparts;
    b->part = malloc(sizeof(part_t) * (size_t)(nparts + 1));
    for (int32_t i = 0; i < nparts; i++) part_init(&b->part[i]);
}

static void build_free(build_t *b)
{
    for (int32_t i = 0; i < b->nparts; i++) part_free(&b->part[i]);
    names_free(&b->ctg);
    free(b->mtiles), free(b->part);
}

/* first tile of an interval; a start below -nbp would index tile -1 in the reference (out of bounds there) */
static int32_t first_tile(int32_t s, int32_t nbp) { return s < 0 ? 0 : s / nbp; }

/* igd_add / igd0_add (src/igd_base.c:118-174, 176-233), the part that depends on the line alone: empty
 * intervals vanish BEFORE the contig is looked up */
static void part_add(part_t *p, const char *chrm, int32_t s, int32_t e, int32_t v, int32_t idx)
{
    if (s >= e) return;
    if (p->nraw == p->cap) p->raw = realloc(p->raw, sizeof(raw_t) * (size_t)(p->cap = p->cap ? p->cap * 2 : 4096));
    raw_t *r = &p->raw[p->nraw++];
    r->ctg = names_get(&p->ctg, chrm, 1), r->idx = idx, r->start = s, r->end = e, r->value = v;
}

/* igd_saveT's progress line (src/igd_base.c:362); nothing is spilled here */
static void batch_line(build_t *b)
{
    int64_t nt = 0;
    for (int32_t i = 0; i < b->ctg.n; i++) nt += b->mtiles[i];
    printf("nCtgs, nRegions, nTiles: %i\t %lld\t %lld\n", b->ctg.n, (long long)b->batch, (long long)nt);
    b->batch = 0;
}

/* The rest of igd_add for one interval whose contig is already numbered: it is replicated into the tiles
 * s/nbp .. (e-1)/nbp, and a contig has as many tiles as its largest end needs */
static void account(build_t *b, const raw_t *r)
{
    if (r->ctg >= b->mt_cap) {
        int32_t cap = b->mt_cap;
        while (r->ctg >= cap) cap *= 2;
        b->mtiles = realloc(b->mtiles, sizeof(int32_t) * (size_t)cap);
        memset(b->mtiles + b->mt_cap, 0, sizeof(int32_t) * (size_t)(cap - b->mt_cap));
        b->mt_cap = cap;
    }
    int32_t n1 = first_tile(r->start, b->nbp), n2 = (r->end - 1) / b->nbp;
    if (n2 + 1 > b->mtiles[r->ctg]) b->mtiles[r->ctg] = n2 + 1;
    if (n2 >= n1) b->batch += n2 - n1 + 1, b->nrec += n2 - n1 + 1;
}

/* One part in arrival order: contigs are numbered by first appearance over all parts.  The reference checks
 * after every line whether its batch exceeds 2^28 records and then closes the file, spills and reopens
 * (src/igd_create.c:72-91); `spill` reproduces what that prints. */
static void merge_part(build_t *b, part_t *p, void (*spill)(build_t *, int32_t), int32_t ig)
{
    int32_t *global = malloc(sizeof(int32_t) * (size_t)(p->ctg.n + 1));
    for (int32_t k = 0; k < p->ctg.n; k++) global[k] = -1;
    for (int64_t i = 0; i < p->nraw; i++) {
        raw_t *r = &p->raw[i];
        if (global[r->ctg] < 0) global[r->ctg] = names_get(&b->ctg, p->ctg.name[r->ctg], 1);
        r->ctg = global[r->ctg];
        account(b, r);
        if (b->batch > BATCH_RECORDS) spill(b, ig);
    }
    free(global);
}

/* ---- tile sort ------------------------------------------------------------------------------------------
 * The reference orders a tile with klib's radix_sort (KRADIX_SORT_INIT(intv, gdata_t, .start, 4),
 * src/igd_base.c:10, macro at src/igd_base.h:196-249): runs of <= 64 records by insertion sort (stable), longer
 * ones by an in-place most-significant-byte-first distribution, one byte per level, each record carried along
 * its displacement cycle to the head of its digit's range.  That distribution is not stable, and at the last
 * byte nothing re-sorts the ranges, so the order of records with EQUAL start is an artefact of the cycles.
 * To write the same bytes the same cycles are walked here (digit of a negative start: arithmetic shift, as
 * there). */
#define RUN_BY_INSERTION 64
#define DIGIT(x, sh) (((x) >> (sh)) & 255)

static void sort_run(rec_t *a, rec_t *end)
{
    for (rec_t *p = a + 1; p < end; p++) {
        if (p->start >= p[-1].start) continue;
        rec_t x = *p, *q = p;
        for (; q > a && x.start < q[-1].start; q--) *q = q[-1];
        *q = x;
    }
}

static void sort_level(rec_t *a, rec_t *end, int sh)
{
    rec_t *head[256], *stop[256];
    size_t cnt[256];
    memset(cnt, 0, sizeof cnt);
    for (rec_t *p = a; p != end; p++) cnt[DIGIT(p->start, sh)]++;
    rec_t *at = a;
    for (int d = 0; d < 256; d++) head[d] = at, at += cnt[d], stop[d] = at;
    for (int d = 0; d < 256;) {
        if (head[d] == stop[d]) { d++; continue; }
        int t = DIGIT(head[d]->start, sh);
        if (t == d) { head[d]++; continue; }
        rec_t carried = *head[d];
        do {
            rec_t evicted = *head[t];
            *head[t]++ = carried;
            carried = evicted;
            t = DIGIT(carried.start, sh);
        } while (t != d);
        *head[d]++ = carried;
    }
    if (sh == 0) return;
    at = a;
    for (int d = 0; d < 256; at = stop[d], d++) {
        size_t n = (size_t)(stop[d] - at);
        if (n > RUN_BY_INSERTION) sort_level(at, stop[d], sh - 8);
        else if (n > 1) sort_run(at, stop[d]);
    }
}

static void sort_tile(rec_t *a, int64_t n)
{
    if (n <= RUN_BY_INSERTION) sort_run(a, a + n);
    else sort_level(a, a + n, 24);
}

/* ---- filling and sorting the tiles: every thread owns a contiguous range of global tile indices ------------ */
typedef struct {
    const build_t *b;
    const int64_t *tbase;               /* [nctg + 1] first global tile of each contig */
    int64_t *toff;                      /* [T + 1]; pass 1 leaves counts in toff[t + 1] */
    rec_t *rec;
    int64_t *cur;                       /* [T] fill cursors (pass 2) */
    int64_t t0, t1;                     /* this thread's tiles */
    int pass;
} fill_job;

static void *fill_worker(void *arg)
{
    fill_job *J = arg;
    const build_t *b = J->b;
    if (J->pass == 3) {
        for (int64_t t = J->t0; t < J->t1; t++) sort_tile(J->rec + J->toff[t], J->toff[t + 1] - J->toff[t]);
        return NULL;
    }
    for (int32_t f = 0; f < b->nparts; f++) {
        const part_t *p = &b->part[f];
        for (int64_t i = 0; i < p->nraw; i++) {
            const raw_t *r = &p->raw[i];
            const int64_t g = J->tbase[r->ctg];
            int64_t lo = g + first_tile(r->start, b->nbp), hi = g + (r->end - 1) / b->nbp;
            if (lo < J->t0) lo = J->t0;
            if (hi >= J->t1) hi = J->t1 - 1;
            if (lo > hi) continue;
            if (J->pass == 1) {
                for (int64_t t = lo; t <= hi; t++) J->toff[t + 1]++;
            } else {
                const rec_t x = { r->idx, r->start, r->end, r->value };
                for (int64_t t = lo; t <= hi; t++) J->rec[J->cur[t]++] = x;
            }
        }
    }
    return NULL;
}

static void run_fill(fill_job *jobs, int nth, int pass)
{
    pthread_t th[64];
    int started = 0;
    for (int i = 0; i < nth; i++) jobs[i].pass = pass;
    for (int i = 1; i < nth; i++)
        if (pthread_create(&th[started], NULL, fill_worker, &jobs[i]) == 0) started++;
        else fill_worker(&jobs[i]);
    fill_worker(&jobs[0]);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}

static int host_threads(void)
{
    const char *e = getenv("IGD_B200_CREATE_THREADS");
    long nc = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return (int)(nc < 1 ? 1 : nc > 64 ? 64 : nc);
}

/* ---- igd_save / igd0_save (src/igd_base.c:396-461, 463-517) --------------------------------------------- */
static int build_write(build_t *b, const char *opath, const char *name)
{
    char path[2048];
    snprintf(path, sizeof path, "%s%s.igd", opath, name);
    FILE *fp = fopen(path, "wb");
    if (!fp) { printf("Can't open file %s", path); return -1; }
    const int32_t m = b->ctg.n;
    int64_t *tbase = malloc(sizeof(int64_t) * (size_t)(m + 1));
    tbase[0] = 0;
    for (int32_t i = 0; i < m; i++) tbase[i + 1] = tbase[i] + b->mtiles[i];
    const int64_t T = tbase[m];
    int64_t *toff = calloc((size_t)T + 2, sizeof(int64_t));
    int32_t *ncnt = malloc(sizeof(int32_t) * (size_t)(T + 1));
    rec_t *rec = malloc(sizeof(rec_t) * (size_t)(b->nrec + 1));
    int64_t *cur = malloc(sizeof(int64_t) * (size_t)(T + 1));
    if (!rec || !cur || !toff || !ncnt) {
        printf("Can't alloc mem %lld\n", (long long)(b->nrec * (int64_t)sizeof(rec_t)));
        fclose(fp);
        remove(path);
        free(rec), free(ncnt), free(toff), free(tbase), free(cur);
        return -1;
    }
    int nth = host_threads();
    if (T < 4096 || b->nrec < (1 << 18)) nth = 1;
    fill_job jobs[64];
    for (int i = 0; i < nth; i++) {
        fill_job J = { b, tbase, toff, rec, cur, T * i / nth, T * (i + 1) / nth, 0 };
        jobs[i] = J;
    }
    run_fill(jobs, nth, 1);                                 /* pass 1: records per tile */
    for (int64_t t = 0; t < T; t++) ncnt[t] = (int32_t)toff[t + 1], toff[t + 1] += toff[t];
    memcpy(cur, toff, sizeof(int64_t) * (size_t)T);
    run_fill(jobs, nth, 2);                                 /* pass 2: arrival order inside every tile */
    phase("count + scatter into tiles");
    run_fill(jobs, nth, 3);                                 /* sort every tile by start */
    phase("sort tiles");
    /* header: nbp, gType, nCtg, nTile[nCtg], nCnt[][], 40-byte names */
    setvbuf(fp, NULL, _IOFBF, 1 << 22);
    fwrite(&b->nbp, 4, 1, fp);
    fwrite(&b->gtype, 4, 1, fp);
    fwrite(&m, 4, 1, fp);
    fwrite(b->mtiles, 4, (size_t)m, fp);
    fwrite(ncnt, 4, (size_t)T, fp);
    for (int32_t i = 0; i < m; i++) {
        char nm[40];
        memset(nm, 0, sizeof nm);
        size_t L = strlen(b->ctg.name[i]);          /* a name of >= 40 chars fills the field, unterminated, as there */
        memcpy(nm, b->ctg.name[i], L < sizeof nm ? L : sizeof nm);
        fwrite(nm, 40, 1, fp);
    }
    if (b->gtype == 1) fwrite(rec, sizeof(rec_t), (size_t)b->nrec, fp);
    else
        for (int64_t i = 0; i < b->nrec; i++) fwrite(&rec[i], 12, 1, fp);      /* gdata0_t = {idx, start, end} */
    int bad = ferror(fp);
    bad |= fclose(fp);
    if (bad) remove(path);                                  /* no partial database is left behind */
    phase("write");
    free(rec), free(ncnt), free(toff), free(tbase), free(cur);
    return bad ? -1 : 0;
}

/* <name>_index.tsv (src/igd_create.c:96-111): index, file name without its directory, lines, mean length */
static void write_tsv(const char *opath, const char *name, int32_t n_files, char *const *file_ids, const int32_t *nr,
                      const double *avg, int64_t *nT_out, double *lavg_out)
{
    char path[2048];
    snprintf(path, sizeof path, "%s%s_index.tsv", opath, name);
    FILE *fp = fopen(path, "w");
    double l_avg = 0.0;
    int64_t nT = 0;
    if (!fp) printf("Can't open file %s", path);
    else fprintf(fp, "Index\tFile\tNumber of regions\tAvg size\n");
    for (int32_t i = 0; i < n_files; i++) {
        const char *base = strrchr(file_ids[i], '/');
        base = base ? base + 1 : file_ids[i];
        nT += nr[i];
        l_avg += avg[i];
        if (fp) fprintf(fp, "%i\t%s\t%i\t%f\n", i, base, nr[i], avg[i] / nr[i]);
    }
    if (fp) fclose(fp);
    *nT_out = nT, *lavg_out = l_avg;
}

/* ---- line splitting ---------------------------------------------------------------------------------------
 * str_splits (src/igd_base.c:37-51) as create uses it: tabs become terminators, and the number of fields found
 * on one line is the limit (+1) for the next -- `nmax` is in/out and the callers never reset it -- so a line
 * cannot have more than one field more than its predecessor (all files share the counter).  The pointer table
 * also persists: a line with fewer fields than the caller reads leaves the older pointers in place, which then
 * look at whatever the new line put at those offsets.  Both effects are kept (same table, same buffer). */
typedef struct {
    char *field[40];                                        /* SPLIT_MAX */
    int nmax;
} splitter;

/* The reference's table has nCols + 1 = 17 entries and is overrun by wider lines (heap corruption there); here
 * the table is bounded: fields past SPLIT_MAX - 1 are not split off (create reads fields 0 .. 4 only). */
#define SPLIT_MAX 40
static void split_line(splitter *S, char *line)
{
    if (S->nmax > SPLIT_MAX - 1) S->nmax = SPLIT_MAX - 1;
    S->field[S->nmax] = NULL;
    S->field[0] = line;
    int ns = 1;
    const int lim = S->nmax + 1 < SPLIT_MAX - 1 ? S->nmax + 1 : SPLIT_MAX - 1;
    char *c = line;
    do {
        if (*c == '\t') {
            *c = '\0';
            S->field[ns++] = c + 1;
        }
        c++;
    } while (*c != '\0' && ns < lim);
    S->nmax = ns;
}

static int32_t field_int(const splitter *S, int k) { return S->field[k] ? (int32_t)atol(S->field[k]) : 0; }

/* parse_bed's acceptance rule for -f mode (src/igd_base.c:53-72) */
static char *accept_bed(char *s, int32_t *st_out, int32_t *en_out)
{
    int nf = 0;
    char *fld = s, *ctg = NULL;
    int32_t st = -1, en = -1;
    for (char *p = s;; p++) {
        char ch = *p;
        if (ch != '\t' && ch != '\0') continue;
        *p = '\0';
        if (nf == 0) ctg = fld;
        else if (nf == 1) st = (int32_t)atol(fld);
        else if (nf == 2) en = (int32_t)atol(fld);
        nf++;
        fld = p + 1;
        if (ch == '\0') break;
    }
    *st_out = st, *en_out = en;
    return nf >= 3 && strncmp(ctg, "chr", 3) == 0 && strlen(ctg) < 40 && en > 0 ? ctg : NULL;
}

/* ---- the file loop shared by create_igd, create_igd_f and create_igd0 ---------------------------------------
 * Files are parsed independently, several at a time.  The one thing a file inherits from its predecessor in the
 * reference is the splitter's field limit; a file is parsed as if it inherited the initial limit and parsed again,
 * during the ordered merge, in the rare case where the limit it really inherits would have cut its first line
 * shorter (parse_file() records the limit after the first line for that check). */
enum { MODE_DIR, MODE_LIST, MODE_G0 };

static void parse_file(part_t *p, const char *path, int mode, int state_in, int32_t ig)
{
    const int linecap = mode == MODE_G0 ? 256 : 1024;       /* gzgets buffer sizes, :46 and :268 */
    char buffer[1024];
    splitter S;
    memset(&S, 0, sizeof S);                                /* no field pointers survive from another file */
    S.nmax = state_in;
    gzFile fp = gzopen(path, "r");
    p->opened = fp != NULL;
    if (!fp) return;
    gzbuffer(fp, 1 << 18);
    while (gzgets(fp, buffer, linecap) != NULL) {
        if (mode == MODE_LIST) {
            int32_t st, en;
            char *ctg = accept_bed(buffer, &st, &en);
            if (ctg && st >= 0 && en < 321000000) {         /* :188 */
                part_add(p, ctg, st, en, 0, ig);
                p->nr++;
                p->avg += en - st;
            }
            continue;
        }
        split_line(&S, buffer);
        if (p->state_first < 0) p->state_first = S.nmax;
        p->state_last = S.nmax;
        int32_t st = field_int(&S, 1), en = field_int(&S, 2), va = 0;
        if (mode == MODE_DIR && S.nmax > 4) va = field_int(&S, 4);
        part_add(p, S.field[0], st, en, va, ig);
        p->nr++;
        p->avg += en - st;
    }
    gzclose(fp);
}

typedef struct {
    build_t *b;
    char **file_ids;
    int mode;
    int32_t next;
    pthread_mutex_t mu;
} parse_job;

static void *parse_worker(void *arg)
{
    parse_job *J = arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int32_t ig = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (ig >= J->b->nparts) return NULL;
        parse_file(&J->b->part[ig], J->file_ids[ig], J->mode, 16, ig);
    }
}

static int32_t g_nf10;
static void spill_trace(build_t *b, int32_t ig)             /* the file is closed, the batch spilled (:77-85) */
{
    if (g_nf10 && ig % g_nf10 == 0) printf(".");
    batch_line(b);
}
static void spill_quiet(build_t *b, int32_t ig) { (void)ig; b->batch = 0; }    /* igd0_saveT prints nothing */

/* returns -1 when a file of a directory build cannot be opened: the reference returns from create_igd at that
 * point (src/igd_create.c:57-58) -- no _index.tsv, no .igd -- and so do the callers here */
static int read_files(build_t *b, int mode, int32_t n_files, char **file_ids, int32_t *nr, double *avg)
{
    parse_job J = { b, file_ids, mode, 0, PTHREAD_MUTEX_INITIALIZER };
    int nth = host_threads();
    if (nth > n_files) nth = n_files;
    pthread_t th[64];
    int started = 0;
    for (int i = 1; i < nth; i++)
        if (pthread_create(&th[started], NULL, parse_worker, &J) == 0) started++;
    parse_worker(&J);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    phase("read + parse");
    g_nf10 = n_files / 10;
    int state = 16;                                         /* nCols, :45 */
    for (int32_t ig = 0; ig < n_files; ig++) {
        part_t *p = &b->part[ig];
        if (!p->opened) {
            if (mode == MODE_LIST) continue;
            return -1;                                      /* :57-58: the reference gives up here */
        }
        if (mode != MODE_LIST && p->state_first > state + 1) {      /* would have been cut at state + 1 fields */
            part_free(p);
            part_init(p);
            parse_file(p, file_ids[ig], mode, state, ig);
        }
        if (p->state_last >= 0) state = p->state_last;
        nr[ig] = p->nr, avg[ig] = p->avg;
        merge_part(b, p, mode == MODE_G0 ? spill_quiet : spill_trace, ig);
        if (mode != MODE_G0 && g_nf10 && (ig + 1) % g_nf10 == 0) printf(".");
    }
    if (mode != MODE_G0) batch_line(b);
    phase("merge (contig numbering)");
    return 0;
}

static void closing_lines(int64_t nT, double l_avg)
{
    printf("Total intervals, l_avg:  %lld %12.3f\n", (long long)nT, l_avg / nT);
}

/* create_igd (src/igd_create.c:25-121): every file matched by the glob, in glob order = file index */
void create_igd(char *iPath, char *oPath, char *igdName)
{
    printf("Create igd from %s: \n", iPath);
    glob_t g;
    if (glob(iPath, 0, NULL, &g) != 0) { printf("wrong dir path: %s\n", iPath); return; }
    int32_t n_files = (int32_t)g.gl_pathc;
    if (n_files < 1) printf("Too few files (add to path /*): %i\n", n_files);
    int32_t *nr = calloc((size_t)n_files + 1, sizeof(int32_t));
    double *avg = calloc((size_t)n_files + 1, sizeof(double));
    build_t b;
    build_init(&b, 1, n_files);
    if (read_files(&b, MODE_DIR, n_files, g.gl_pathv, nr, avg)) {          /* an input file does not open */
        build_free(&b);
        globfree(&g);
        free(nr), free(avg);
        return;
    }
    printf("\n");
    int64_t nT;
    double l_avg;
    write_tsv(oPath, igdName, n_files, g.gl_pathv, nr, avg, &nT, &l_avg);
    if (build_write(&b, oPath, igdName) == 0) printf("Save igd database to %s%s.igd\n", oPath, igdName);
    closing_lines(nT, l_avg);
    build_free(&b);
    globfree(&g);
    free(nr), free(avg);
}

/* create_igd_f (src/igd_create.c:124-243): iPath lists the source files, one per line; a file is taken if it
 * opens and its first line passes parse_bed */
void create_igd_f(char *iPath, char *oPath, char *igdName)
{
    printf("Create igd from %s: \n", iPath);
    FILE *fl = fopen(iPath, "r");
    if (!fl) { printf("Can't open file %s", iPath); return; }
    char buf[1024], first[1024];
    int32_t cap = 64, n_files = 0, listed = 0;
    char **file_ids = malloc(sizeof(char *) * (size_t)cap);
    while (fgets(buf, sizeof buf, fl)) {
        listed++;
        buf[strcspn(buf, "\n")] = 0;
        gzFile fp = gzopen(buf, "r");
        if (!fp) continue;
        int32_t st, en;
        first[0] = 0;
        if (gzgets(fp, first, sizeof first) && accept_bed(first, &st, &en)) {
            if (n_files == cap) file_ids = realloc(file_ids, sizeof(char *) * (size_t)(cap *= 2));
            file_ids[n_files++] = strdup(buf);
        }
        gzclose(fp);
    }
    fclose(fl);
    if (listed < 1) { printf("Too few files (add to path /*): %i\n", listed); free(file_ids); return; }
    int32_t *nr = calloc((size_t)n_files + 1, sizeof(int32_t));
    double *avg = calloc((size_t)n_files + 1, sizeof(double));
    build_t b;
    build_init(&b, 1, n_files);
    read_files(&b, MODE_LIST, n_files, file_ids, nr, avg);
    printf("\n");
    int64_t nT;
    double l_avg;
    write_tsv(oPath, igdName, n_files, file_ids, nr, avg, &nT, &l_avg);
    if (build_write(&b, oPath, igdName) == 0) printf("Save igd database to %s%s.igd\n", oPath, igdName);
    closing_lines(nT, l_avg);
    build_free(&b);
    for (int32_t i = 0; i < n_files; i++) free(file_ids[i]);
    free(file_ids), free(nr), free(avg);
}

/* create_igd0 (src/igd_create.c:245-332): gdata0_t records, no value column */
void create_igd0(char *iPath, char *oPath, char *igdName)
{
    printf("igd_create 0\n");
    glob_t g;
    if (glob(iPath, 0, NULL, &g) != 0) { printf("wrong dir path: %s", iPath); return; }
    int32_t n_files = (int32_t)g.gl_pathc;
    if (n_files < 1) printf("Too few files (add to path /*): %i\n", n_files);
    int32_t *nr = calloc((size_t)n_files + 1, sizeof(int32_t));
    double *avg = calloc((size_t)n_files + 1, sizeof(double));
    printf("igd_create 1: %i\n", n_files);
    build_t b;
    build_init(&b, 0, n_files);
    if (read_files(&b, MODE_G0, n_files, g.gl_pathv, nr, avg)) {           /* an input file does not open */
        build_free(&b);
        globfree(&g);
        free(nr), free(avg);
        return;
    }
    int64_t nT;
    double l_avg;
    write_tsv(oPath, igdName, n_files, g.gl_pathv, nr, avg, &nT, &l_avg);
    printf("igd_create 3\n");
    build_write(&b, oPath, igdName);
    printf("igd_create 4\n");
    closing_lines(nT, l_avg);
    build_free(&b);
    globfree(&g);
    free(nr), free(avg);
}

/* create_igd_bed4 (src/igd_create.c:334-433): ONE bed file whose 4th column names the dataset of each line;
 * datasets are numbered by first appearance, the 5th column is the value */
void create_igd_bed4(char *iPath, char *oPath, char *igdName)
{
    printf("igd_create 1\n");
    build_t b;
    build_init(&b, 1, 1);
    part_t *P = &b.part[0];             /* one input file: its contig numbering is the global one */
    names_t sets;
    names_init(&sets);
    int32_t cap = 1024;
    int32_t *nr = calloc((size_t)cap, sizeof(int32_t));
    double *avg = calloc((size_t)cap, sizeof(double));
    char buffer[1024];
    splitter S;
    memset(&S, 0, sizeof S);
    S.nmax = 32;
    int64_t nL = 0, seen = 0;
    int spills = 0;
    gzFile fp = gzopen(iPath, "r");
    while (fp && gzgets(fp, buffer, sizeof buffer) != NULL) {
        split_line(&S, buffer);
        int32_t idx = names_get(&sets, S.field[3] ? S.field[3] : "", 1);
        if (idx >= cap) {
            nr = realloc(nr, sizeof(int32_t) * (size_t)cap * 2);
            avg = realloc(avg, sizeof(double) * (size_t)cap * 2);
            memset(nr + cap, 0, sizeof(int32_t) * (size_t)cap);
            memset(avg + cap, 0, sizeof(double) * (size_t)cap);
            cap *= 2;
        }
        int32_t st = field_int(&S, 1), en = field_int(&S, 2);
        part_add(P, S.field[0], st, en, field_int(&S, 4), idx);
        nr[idx]++;
        avg[idx] += en - st;
        nL++;
        if (P->nraw > seen) {           /* the line added an interval: account for it in step with the reading */
            const raw_t *r = &P->raw[seen++];
            if (r->ctg == b.ctg.n) names_get(&b.ctg, P->ctg.name[r->ctg], 1);
            account(&b, r);
        }
        if (b.batch >= BATCH_RECORDS) {
            printf("--igd_saveT1--%i, %lld\n", ++spills, (long long)b.batch);
            batch_line(&b);
            nL = 0;
        }
    }
    if (fp) gzclose(fp);
    if (nL > 0) batch_line(&b);
    printf("igd_create 2\n");
    int64_t nT;
    double l_avg;
    write_tsv(oPath, igdName, sets.n, sets.name, nr, avg, &nT, &l_avg);
    printf("igd_create 3\n");
    build_write(&b, oPath, igdName);
    printf("igd_create 4\n");
    closing_lines(nT, l_avg);
    build_free(&b);
    names_free(&sets);
    free(nr), free(avg);
}

/* create_iGD (src_py/igd_create.c:19-49): what igd_py.create calls -- the default flavour, bin width in bp */
int create_iGD(const char *iPath, const char *oPath, const char *igdName, int tile_size_bp)
{
    char ipath[1024], opath[1024], ftmp[2200], name[512];
    snprintf(ipath, sizeof ipath - 2, "%s", iPath);
    snprintf(opath, sizeof opath - 1, "%s", oPath);
    snprintf(name, sizeof name, "%s", igdName);
    size_t lo = strlen(opath), li = strlen(ipath);
    if (lo == 0 || opath[lo - 1] != '/') strcat(opath, "/");
    if (li > 0 && ipath[li - 1] == '/') strcat(ipath, "*");
    else if (li > 0 && ipath[li - 1] != '*') strcat(ipath, "/*");
    struct stat st;
    snprintf(ftmp, sizeof ftmp, "%s%s.igd", opath, name);
    if (stat(ftmp, &st) == 0) {
        printf("The igd database file %s exists!\n", ftmp);
        return 1;
    }
    if (stat(opath, &st) == -1) mkdir(opath, 0777);
    snprintf(ftmp, sizeof ftmp, "%sdata0", opath);
    if (stat(ftmp, &st) == -1) mkdir(ftmp, 0777);
    tile_size = tile_size_bp > 0 ? tile_size_bp : 16384;
    create_igd(ipath, opath, name);
    fflush(stdout);
    return 0;
}

static int create_usage(void)
{
    fprintf(stderr,
            "%s, v%s\n"
            "usage:   %s create <input dir> <output dir> <output igd name> [options] \n"
            "             -s  <Type of data structure> \n"
            "                   0 for [index, start, end]\n"
            "                   1 for [index, start, end, value], default\n"
            "             -f  (iPath is a file that lists paths of data src files) \n"
            "             -b  <Tile size in power of 2 (default 14)> \n",
            IGD_PROGRAM, IGD_VERSION, IGD_PROGRAM);
    return 0;
}

/* igd_create (src/igd_create.c:436-502): argument scan and dispatch */
int igd_create(int argc, char **argv)
{
    if (argc < 5) return create_usage();
    char ipath[1024], opath[1024], ftmp[2200];
    snprintf(ipath, sizeof ipath - 2, "%s", argv[2]);
    snprintf(opath, sizeof opath - 1, "%s", argv[3]);
    char *dbname = argv[4];
    int dtype = 1, ftype = 0;
    tile_size = 16384;
    for (int i = 5; i < argc; i++) {
        if (strcmp(argv[i], "-s") == 0 && i + 1 < argc) dtype = atoi(argv[i + 1]);
        if (strcmp(argv[i], "-b") == 0 && i + 1 < argc) {
            int n = atoi(argv[i + 1]);
            if (n > 10 && n < 20) tile_size = 1 << n;
        }
        if (strcmp(argv[i], "-f") == 0) ftype = 1;
    }
    size_t lo = strlen(opath), li = strlen(ipath);
    if (lo == 0 || opath[lo - 1] != '/') strcat(opath, "/");
    if (ftype == 0 && dtype != 2 && li > 0) {
        if (ipath[li - 1] == '/') strcat(ipath, "*");
        else if (ipath[li - 1] != '*') strcat(ipath, "/*");
    }
    struct stat st;
    snprintf(ftmp, sizeof ftmp, "%s%s.igd", opath, dbname);
    if (stat(ftmp, &st) == 0) {
        printf("The igd database file %s exists!\n", ftmp);
        return 0;
    }
    if (stat(opath, &st) == -1) mkdir(opath, 0777);
    snprintf(ftmp, sizeof ftmp, "%sdata0", opath);      /* the reference's spill directory; left empty, as there */
    if (stat(ftmp, &st) == -1) mkdir(ftmp, 0777);
    if (dtype == 0) create_igd0(ipath, opath, dbname);
    else if (dtype == 2) create_igd_bed4(ipath, opath, dbname);
    else if (ftype == 1) create_igd_f(ipath, opath, dbname);
    else create_igd(ipath, opath, dbname);
    return 0;
}
