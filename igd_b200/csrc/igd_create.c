/* `igd create` -- index build, host-side C (north_star: "Index build (igd create) stays host-side C").
 *
 * Same command line, same accepted input, same files out as the reference's igd_create()
 * (src/igd_create.c:436-502) in its four flavours -- create_igd (:25-121, the default), create_igd_f (-f, :124-243),
 * create_igd0 (-s 0, :245-332) and create_igd_bed4 (-s 2, :334-433) -- but not its mechanism.  The reference
 * grows one malloc'd list per tile, spills every 2^28 records into one temp file per tile under <out>/data0/
 * (igd_saveT, src/igd_base.c:333-364) and re-reads, sorts and appends them tile by tile (igd_save, :396-461).
 * Here the accepted intervals are kept as one flat array; a counting pass sizes every tile, a second pass drops
 * each interval into the tiles it touches in arrival order (what the appends to the temp files amount to), the
 * tiles are sorted in parallel on all host cores and the file is written once.
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
#include <unistd.h>
#include <zlib.h>

#define IGD_PROGRAM "igd"
#define IGD_VERSION "0.1.1"
#define BATCH_RECORDS 268435456LL       /* maxCount, src/igd_base.h:37: the reference spills above this */

int32_t tile_size = 16384;              /* the reference's global of the same name (src/igd.c:18) */

typedef struct { int32_t idx, start, end, value; } rec_t;               /* gdata_t, src/igd_base.h:41-46 */
typedef struct { int32_t ctg, idx, start, end, value; } raw_t;          /* one accepted input interval */

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
typedef struct {
    int32_t nbp, gtype;
    names_t ctg;
    int32_t *mtiles;                    /* per contig: ctg_t.mTiles = 1 + the largest (end-1)/nbp seen */
    int32_t mt_cap;
    raw_t *raw;
    int64_t nraw, raw_cap;
    int64_t batch;                      /* igd_t.total: replicated records since the last "spill" */
    int64_t nrec;                       /* replicated records in all */
} build_t;

static void build_init(build_t *b, int gtype)
{
    memset(b, 0, sizeof *b);
    b->nbp = tile_size, b->gtype = gtype;
    names_init(&b->ctg);
    b->mt_cap = 64;
    b->mtiles = calloc((size_t)b->mt_cap, sizeof(int32_t));
    b->raw_cap = 1 << 20;
    b->raw = malloc(sizeof(raw_t) * (size_t)b->raw_cap);
}

static void build_free(build_t *b)
{
    names_free(&b->ctg);
    free(b->mtiles), free(b->raw);
}

/* first tile of an interval; a start below -nbp would index tile -1 in the reference (out of bounds there) */
static int32_t first_tile(int32_t s, int32_t nbp) { return s < 0 ? 0 : s / nbp; }

/* igd_add / igd0_add (src/igd_base.c:118-174, 176-233): empty intervals vanish BEFORE the contig is looked
 * up; the interval is replicated into the tiles s/nbp .. (e-1)/nbp. */
static void build_add(build_t *b, const char *chrm, int32_t s, int32_t e, int32_t v, int32_t idx)
{
    if (s >= e) return;
    int32_t c = names_get(&b->ctg, chrm, 1);
    if (c >= b->mt_cap) {
        b->mtiles = realloc(b->mtiles, sizeof(int32_t) * (size_t)b->mt_cap * 2);
        memset(b->mtiles + b->mt_cap, 0, sizeof(int32_t) * (size_t)b->mt_cap);
        b->mt_cap *= 2;
    }
    int32_t n1 = first_tile(s, b->nbp), n2 = (e - 1) / b->nbp;
    if (n2 + 1 > b->mtiles[c]) b->mtiles[c] = n2 + 1;
    if (b->nraw == b->raw_cap) b->raw = realloc(b->raw, sizeof(raw_t) * (size_t)(b->raw_cap *= 2));
    raw_t *r = &b->raw[b->nraw++];
    r->ctg = c, r->idx = idx, r->start = s, r->end = e, r->value = v;
    if (n2 >= n1) b->batch += n2 - n1 + 1, b->nrec += n2 - n1 + 1;
}

/* igd_saveT's progress line (src/igd_base.c:362); nothing is spilled here */
static void batch_line(build_t *b)
{
    int64_t nt = 0;
    for (int32_t i = 0; i < b->ctg.n; i++) nt += b->mtiles[i];
    printf("nCtgs, nRegions, nTiles: %i\t %lld\t %lld\n", b->ctg.n, (long long)b->batch, (long long)nt);
    b->batch = 0;
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

typedef struct {
    rec_t *rec;
    const int64_t *toff;
    int64_t ntiles;
    int64_t next;                       /* shared cursor, claimed in blocks */
    pthread_mutex_t mu;
} sort_job;

static void *sort_worker(void *arg)
{
    sort_job *J = arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int64_t t0 = J->next, t1 = t0 + 256 < J->ntiles ? t0 + 256 : J->ntiles;
        J->next = t1;
        pthread_mutex_unlock(&J->mu);
        if (t0 >= t1) return NULL;
        for (int64_t t = t0; t < t1; t++) sort_tile(J->rec + J->toff[t], J->toff[t + 1] - J->toff[t]);
    }
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
    /* pass 1: records per tile */
    for (int64_t i = 0; i < b->nraw; i++) {
        const raw_t *r = &b->raw[i];
        int64_t g = tbase[r->ctg];
        for (int32_t j = first_tile(r->start, b->nbp), n2 = (r->end - 1) / b->nbp; j <= n2; j++) toff[g + j + 1]++;
    }
    int32_t *ncnt = malloc(sizeof(int32_t) * (size_t)(T + 1));
    for (int64_t t = 0; t < T; t++) ncnt[t] = (int32_t)toff[t + 1], toff[t + 1] += toff[t];
    /* pass 2: arrival order inside every tile */
    rec_t *rec = malloc(sizeof(rec_t) * (size_t)(b->nrec + 1));
    int64_t *cur = malloc(sizeof(int64_t) * (size_t)(T + 1));
    if (!rec || !cur) { printf("Can't alloc mem %lld\n", (long long)(b->nrec * (int64_t)sizeof(rec_t))); return -1; }
    memcpy(cur, toff, sizeof(int64_t) * (size_t)(T + 1));
    for (int64_t i = 0; i < b->nraw; i++) {
        const raw_t *r = &b->raw[i];
        int64_t g = tbase[r->ctg];
        rec_t x = { r->idx, r->start, r->end, r->value };
        for (int32_t j = first_tile(r->start, b->nbp), n2 = (r->end - 1) / b->nbp; j <= n2; j++) rec[cur[g + j]++] = x;
    }
    free(cur);
    free(b->raw), b->raw = NULL;
    /* sort every tile by start, all cores */
    {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        int nth = (int)(nc < 1 ? 1 : nc > 64 ? 64 : nc);
        if (T < 512) nth = 1;
        sort_job J = { rec, toff, T, 0, PTHREAD_MUTEX_INITIALIZER };
        pthread_t th[64];
        int started = 0;
        for (int i = 1; i < nth; i++)
            if (pthread_create(&th[started], NULL, sort_worker, &J) == 0) started++;
        sort_worker(&J);
        for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    }
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
    free(rec), free(ncnt), free(toff), free(tbase);
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
    char *field[40];
    int nmax;
} splitter;

static void split_line(splitter *S, char *line)
{
    S->field[S->nmax] = NULL;
    S->field[0] = line;
    int ns = 1;
    char *c = line;
    do {
        if (*c == '\t') {
            *c = '\0';
            S->field[ns++] = c + 1;
        }
        c++;
    } while (*c != '\0' && ns < S->nmax + 1);
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
 * The reference reads until igd->total exceeds 2^28, spills, and resumes in the same file at the next line
 * (src/igd_create.c:51-91).  Reading straight through adds the same records in the same order; what is left of
 * the batches is their trace on stdout, reproduced by batch_line() and the dots. */
enum { MODE_DIR, MODE_LIST, MODE_G0 };

static void read_files(build_t *b, int mode, int32_t n_files, char **file_ids, int32_t *nr, double *avg)
{
    const int linecap = mode == MODE_G0 ? 256 : 1024;       /* gzgets buffer sizes, :46 and :268 */
    const int32_t nf10 = n_files / 10;
    char buffer[1024];
    splitter S;
    memset(&S, 0, sizeof S);
    S.nmax = 16;
    for (int32_t ig = 0; ig < n_files; ig++) {
        gzFile fp = gzopen(file_ids[ig], "r");
        if (!fp) {
            if (mode == MODE_LIST) continue;
            return;                                         /* :57-58 */
        }
        gzbuffer(fp, 1 << 18);
        while (gzgets(fp, buffer, linecap) != NULL) {
            if (mode == MODE_LIST) {
                int32_t st, en;
                char *ctg = accept_bed(buffer, &st, &en);
                if (ctg && st >= 0 && en < 321000000) {     /* :188 */
                    build_add(b, ctg, st, en, 0, ig);
                    nr[ig]++;
                    avg[ig] += en - st;
                }
            } else {
                split_line(&S, buffer);
                int32_t st = field_int(&S, 1), en = field_int(&S, 2), va = 0;
                if (mode == MODE_DIR && S.nmax > 4) va = field_int(&S, 4);
                build_add(b, S.field[0], st, en, va, ig);
                nr[ig]++;
                avg[ig] += en - st;
            }
            if (b->batch > BATCH_RECORDS && mode != MODE_G0) {      /* the reference closes the file, spills, reopens */
                if (nf10 && ig % nf10 == 0) printf(".");
                batch_line(b);
            } else if (b->batch > BATCH_RECORDS) b->batch = 0;      /* igd0_saveT prints nothing */
        }
        gzclose(fp);
        if (mode != MODE_G0 && nf10 && (ig + 1) % nf10 == 0) printf(".");
    }
    if (mode != MODE_G0) batch_line(b);
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
    build_init(&b, 1);
    read_files(&b, MODE_DIR, n_files, g.gl_pathv, nr, avg);
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
    build_init(&b, 1);
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
    build_init(&b, 0);
    read_files(&b, MODE_G0, n_files, g.gl_pathv, nr, avg);
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
    build_init(&b, 1);
    names_t sets;
    names_init(&sets);
    int32_t cap = 1024;
    int32_t *nr = calloc((size_t)cap, sizeof(int32_t));
    double *avg = calloc((size_t)cap, sizeof(double));
    char buffer[1024];
    splitter S;
    memset(&S, 0, sizeof S);
    S.nmax = 32;
    int64_t nL = 0;
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
        build_add(&b, S.field[0], st, en, field_int(&S, 4), idx);
        nr[idx]++;
        avg[idx] += en - st;
        nL++;
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
