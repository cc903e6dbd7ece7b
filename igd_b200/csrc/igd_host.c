/* igd_b200 host side (C): .igd / _index.tsv parsing, BED query ingest, the handle-based C-ABI of
 * include/igd_b200.h and the reference-named global-state entry points.  All searching is done by the
 * CUDA layer (igd_dev_*.cu) through igd_dev.h; there is no CPU search path in this library. */
#define _GNU_SOURCE
#include "igd_b200.h"
#include "igd_dev.h"

#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

struct igd_b200 {
    int32_t nbp, gtype, nctg, nfiles;
    int32_t *ntile;                 /* [nctg]            iGD_t.nTile */
    int64_t *tbase;                 /* [nctg+1]          first global tile of each contig */
    int32_t *ncnt;                  /* [sum ntile]       iGD_t.nCnt, flattened contig-major */
    char (*cname)[40];              /* iGD_t.cName */
    char **fname;                   /* info_t.fileName */
    int32_t *fnr;                   /* info_t.nr */
    int64_t nrec, ntiles;
    igdb_dev *dev;
    float t_ms[3];
    uint64_t *delta;                /* [nfiles] per-call counts */
};

static __thread char g_err[512] = "";
static igd_b200_t *g_current = NULL;

const char *igd_b200_last_error(void) { return g_err[0] ? g_err : igdb_dev_error(); }
int igd_b200_device_count(void) { return igdb_device_count(); }

static void set_err(const char *fmt, const char *arg)
{
    snprintf(g_err, sizeof g_err, fmt, arg);
}

/* ---- <stem>_index.tsv: written by src/igd_create.c:96-111, read by get_fileinfo src/igd_base.c:235-267.
 * The header line is skipped; column 2 is the file name, column 3 the number of regions. -------------- */
static int read_index_tsv(igd_b200_t *h, const char *igd_path)
{
    size_t L = strlen(igd_path);
    char *tsv = malloc(L + 16);
    memcpy(tsv, igd_path, L + 1);
    char *dot = strrchr(tsv, '.');
    if (!dot) { free(tsv); return -1; }
    strcpy(dot, "_index.tsv");
    FILE *f = fopen(tsv, "r");
    if (!f) { set_err("file not found:%s", tsv); free(tsv); return -1; }
    free(tsv);
    char buf[1024];
    int cap = 64, n = 0;
    h->fname = malloc(sizeof(char *) * cap);
    h->fnr = malloc(sizeof(int32_t) * cap);
    if (!h->fname || !h->fnr) { fclose(f); set_err("%s: out of memory reading the file table", igd_path); return -1; }
    if (!fgets(buf, sizeof buf, f)) { fclose(f); h->nfiles = 0; return 0; }
    while (fgets(buf, sizeof buf, f)) {
        if (n == cap) {
            cap *= 2;
            char **fn = realloc(h->fname, sizeof(char *) * cap);
            int32_t *fr = realloc(h->fnr, sizeof(int32_t) * cap);
            if (fn) h->fname = fn;
            if (fr) h->fnr = fr;
            if (!fn || !fr) { fclose(f); h->nfiles = n; set_err("%s: out of memory reading the file table", igd_path); return -1; }
        }
        char *save = NULL;
        strtok_r(buf, "\t", &save);
        char *name = strtok_r(NULL, "\t", &save);
        char *nr = strtok_r(NULL, "\t", &save);
        h->fname[n] = strdup(name ? name : "");
        h->fnr[n] = nr ? (int32_t)atol(nr) : 0;
        n++;
    }
    fclose(f);
    h->nfiles = n;
    return 0;
}

/* the payload is read by several threads at once (pread of disjoint slices): one thread copies out of the page
 * cache at 3-4 GB/s, which is what bounded the open of a 17 GB database */
typedef struct { int fd; char *dst; size_t len; off_t off; int ok; } slice_t;
static void *read_slice(void *arg)
{
    slice_t *s = arg;
    size_t done = 0;
    while (done < s->len) {
        ssize_t r = pread(s->fd, s->dst + done, s->len - done, s->off + (off_t)done);
        if (r <= 0) { s->ok = 0; return NULL; }
        done += (size_t)r;
    }
    s->ok = 1;
    return NULL;
}
static int read_parallel(int fd, char *dst, size_t len, off_t off)
{
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    int nt = (int)(nc < 1 ? 1 : nc > 8 ? 8 : nc);
    if (len < ((size_t)4 << 20)) nt = 1;
    slice_t sl[8];
    pthread_t th[8];
    int started = 0;
    for (int i = 0; i < nt; i++) {
        size_t a = len * (size_t)i / (size_t)nt, b = len * (size_t)(i + 1) / (size_t)nt;
        sl[i].fd = fd, sl[i].dst = dst + a, sl[i].len = b - a, sl[i].off = off + (off_t)a, sl[i].ok = 0;
    }
    for (int i = 1; i < nt; i++) {
        if (pthread_create(&th[started], NULL, read_slice, &sl[i]) == 0) started++;
        else read_slice(&sl[i]);
    }
    read_slice(&sl[0]);
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    for (int i = 0; i < nt; i++)
        if (!sl[i].ok) return -1;
    return 0;
}

/* ---- .igd: header tables as in get_igdinfo (src/igd_base.c:269-323); payload streamed to the device.
 * Layout (src/igd_base.c:408-458): nbp, gType, nCtg, nTile[nCtg], nCnt[sum nTile], 40-byte names, then
 * every tile's gdata_t records, contig-major. ---------------------------------------------------------- */
igd_b200_t *igd_b200_open(const char *igd_path, int device)
{
    g_err[0] = 0;
    FILE *f = fopen(igd_path, "rb");
    if (!f) { set_err("Can't open file %s", igd_path); return NULL; }
    igd_b200_t *h = calloc(1, sizeof *h);
    int32_t hd[3];
    if (fread(hd, 4, 3, f) != 3) { set_err("%s: truncated header", igd_path); goto fail; }
    h->nbp = hd[0], h->gtype = hd[1], h->nctg = hd[2];
    /* gType 1 = gdata_t records (16 B), gType 0 = gdata0_t (12 B, no value; `igd create -s 0`,
     * src/igd_create.c:245-332, record size chosen as in get_igdinfo src/igd_base.c:279-283) */
    if (h->gtype != 0 && h->gtype != 1) { set_err("%s: unknown gType (0 and 1 are supported)", igd_path); goto fail; }
    if (h->nbp <= 0 || h->nctg < 0) { set_err("%s: corrupt header", igd_path); goto fail; }
    int32_t m = h->nctg;
    /* the table sizes come from the file: a corrupt header must end in an error, not in a NULL dereference */
    h->ntile = malloc(sizeof(int32_t) * ((size_t)m + 1));
    h->tbase = malloc(sizeof(int64_t) * ((size_t)m + 1));
    h->cname = malloc(40 * ((size_t)m + 1));
    if (!h->ntile || !h->tbase || !h->cname) { set_err("%s: corrupt header", igd_path); goto fail; }
    if (m && fread(h->ntile, 4, m, f) != (size_t)m) { set_err("%s: truncated header", igd_path); goto fail; }
    h->tbase[0] = 0;
    for (int i = 0; i < m; i++) {
        if (h->ntile[i] < 0) { set_err("%s: corrupt header", igd_path); goto fail; }
        h->tbase[i + 1] = h->tbase[i] + h->ntile[i];
    }
    h->ntiles = h->tbase[m];
    if (h->ntiles >= 0x7fffffff) { set_err("%s: too many tiles", igd_path); goto fail; }
    h->ncnt = malloc(sizeof(int32_t) * (size_t)(h->ntiles + 1));
    if (!h->ncnt) { set_err("%s: corrupt header", igd_path); goto fail; }
    if (h->ntiles && fread(h->ncnt, 4, h->ntiles, f) != (size_t)h->ntiles) {
        set_err("%s: truncated header", igd_path);
        goto fail;
    }
    if (m && fread(h->cname, 40, m, f) != (size_t)m) { set_err("%s: truncated header", igd_path); goto fail; }
    for (int i = 0; i < m; i++) h->cname[i][39] = 0;
    for (int64_t t = 0; t < h->ntiles; t++) {
        if (h->ncnt[t] < 0) { set_err("%s: corrupt header", igd_path); goto fail; }
        h->nrec += h->ncnt[t];
    }
    if (read_index_tsv(h, igd_path)) goto fail;
    if (igdb_device_count() <= 0) {
        set_err("%s: no CUDA device available (igd_b200 has no CPU search path)", igd_path);
        goto fail;
    }
    h->dev = igdb_dev_create(device, h->nbp, h->nctg, h->ntile, h->ncnt, h->nfiles);
    if (!h->dev) goto fail;
    h->delta = igdb_host_alloc(((size_t)h->nfiles + 1) * sizeof(uint64_t));    /* pinned: hits[] D2H target */
    if (!h->delta) { set_err("%s: cannot allocate pinned host memory", igd_path); goto fail; }
    if (igdb_dev_set_contigs(h->dev, (const char *)h->cname)) goto fail;
    {
        const int64_t CH = 4 << 20;                     /* records per chunk: 64 MiB */
        const int W = h->gtype == 0 ? 3 : 4;            /* 32-bit words per record */
        int64_t chunk = h->nrec < CH ? h->nrec : CH;
        /* two halves of one pinned buffer: chunk k + 1 is read from the file while chunk k is copied and
         * transposed on the device.  File indices index hits[]: the transpose flags any outside
         * [0, nFiles) -- records the reference would write out of bounds with -- and finalize fails. */
        char *buf = chunk ? igdb_dev_pinned(h->dev, (size_t)chunk * 16 * 2) : NULL;
        if (chunk && !buf) { set_err("%s: cannot allocate the pinned upload buffer", igd_path); goto fail; }
        const off_t payload = ftello(f);
        const int fd = fileno(f);
        int k = 0;
        for (int64_t r = 0; r < h->nrec; r += chunk, k++) {
            int64_t m = h->nrec - r < chunk ? h->nrec - r : chunk;
            char *half = buf + (size_t)(k & 1) * (size_t)chunk * 16;
            if (igdb_dev_upload_wait(h->dev, k & 1)) goto fail;
            if (read_parallel(fd, half, (size_t)m * 4 * (size_t)W, payload + (off_t)r * 4 * W)) {
                set_err("%s: truncated tile data", igd_path);
                goto fail;
            }
            if (igdb_dev_upload_slot(h->dev, k & 1, r, m, half, W)) goto fail;
        }
        if (igdb_dev_finalize(h->dev)) {
            if (strstr(igdb_dev_error(), "file index")) set_err("%s: record with file index outside _index.tsv", igd_path);
            goto fail;
        }
    }
    fclose(f);
    return h;
fail:
    fclose(f);
    igd_b200_close(h);
    return NULL;
}

void igd_b200_close(igd_b200_t *h)
{
    if (!h) return;
    if (g_current == h) g_current = NULL;
    if (h->dev) igdb_dev_destroy(h->dev);
    if (h->fname) for (int i = 0; i < h->nfiles; i++) free(h->fname[i]);
    free(h->fname); free(h->fnr); free(h->ntile); free(h->tbase); free(h->ncnt); free(h->cname);
    igdb_host_free(h->delta); free(h);
}

int32_t igd_b200_nfiles(const igd_b200_t *h) { return h->nfiles; }
int32_t igd_b200_nctg(const igd_b200_t *h) { return h->nctg; }
int32_t igd_b200_nbp(const igd_b200_t *h) { return h->nbp; }
int32_t igd_b200_gtype(const igd_b200_t *h) { return h->gtype; }
int64_t igd_b200_ntiles(const igd_b200_t *h) { return h->ntiles; }
int64_t igd_b200_nrecords(const igd_b200_t *h) { return h->nrec; }
int32_t igd_b200_ntile(const igd_b200_t *h, int32_t c) { return c >= 0 && c < h->nctg ? h->ntile[c] : 0; }
int32_t igd_b200_ncnt(const igd_b200_t *h, int32_t c, int32_t t)
{
    if (c < 0 || c >= h->nctg || t < 0 || t >= h->ntile[c]) return 0;
    return h->ncnt[h->tbase[c] + t];
}
int32_t igd_b200_file_nr(const igd_b200_t *h, int32_t i) { return i >= 0 && i < h->nfiles ? h->fnr[i] : 0; }
const char *igd_b200_file_name(const igd_b200_t *h, int32_t i) { return i >= 0 && i < h->nfiles ? h->fname[i] : ""; }
const char *igd_b200_ctg_name(const igd_b200_t *h, int32_t c) { return c >= 0 && c < h->nctg ? h->cname[c] : ""; }
int64_t igd_b200_device_bytes(const igd_b200_t *h) { return igdb_dev_bytes(h->dev); }
int64_t igd_b200_launch_count(const igd_b200_t *h) { return igdb_dev_launches(h->dev); }
void igd_b200_last_plan(const igd_b200_t *h, int64_t out[4]) { igdb_dev_last_plan(h->dev, out); }
float igd_b200_last_search_kernel_ms(igd_b200_t *h) { return igdb_dev_last_search_ms(h->dev); }
void igd_b200_stage_timing(igd_b200_t *h, int on) { igdb_dev_stage_timing(h->dev, on); }
int igd_b200_last_stage_ms(igd_b200_t *h, float t[4]) { return igdb_dev_last_stages(h->dev, t); }
void igd_b200_last_timing(const igd_b200_t *h, float t[3]) { t[0] = h->t_ms[0], t[1] = h->t_ms[1], t[2] = h->t_ms[2]; }

/* get_id (src/igd_base.c:325-331): exact string match on the contig name */
int32_t igd_b200_ctg_id(const igd_b200_t *h, const char *chrm)
{
    for (int i = 0; i < h->nctg; i++)
        if (strcmp(h->cname[i], chrm) == 0) return i;
    return -1;
}

/* ---- count -------------------------------------------------------------------------------------- */
static int64_t count_mode(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                          int mode_v, int32_t v, int64_t *hits)
{
    g_err[0] = 0;
    if (n <= 0) return 0;
    if (igdb_dev_count_host(h->dev, n, ctg, qs, qe, mode_v, v, h->delta, h->t_ms)) return -1;
    int64_t total = 0;
    for (int i = 0; i < h->nfiles; i++) {
        hits[i] += (int64_t)h->delta[i];
        total += (int64_t)h->delta[i];
    }
    return total;
}

/* the CLI's dispatch: the signal filter (and get_overlaps_v's control flow) only when v > 0
 * (src/igd_search.c:1027,1045), and never on a gType-0 database, whose records have no value: the reference
 * calls getOverlaps0 / get_overlaps0 there whatever -v says (:1024-1025, :1042-1043) */
static int filter_on(const igd_b200_t *h, int32_t v) { return v > 0 && h->gtype != 0; }

int64_t igd_b200_count(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                       int32_t v, int64_t *hits)
{
    return count_mode(h, n, ctg, qs, qe, filter_on(h, v), v, hits);
}

int igd_b200_count_submit(igd_b200_t *h, int slot, int64_t n, const int32_t *ctg, const int32_t *qs,
                          const int32_t *qe, int32_t v)
{
    g_err[0] = 0;
    return igdb_dev_count_submit(h->dev, slot, n, ctg, qs, qe, filter_on(h, v), v);
}

int64_t igd_b200_count_wait(igd_b200_t *h, int slot, int64_t *hits)
{
    g_err[0] = 0;
    if (igdb_dev_count_wait(h->dev, slot, h->delta)) return -1;
    int64_t total = 0;
    for (int i = 0; i < h->nfiles; i++) {
        hits[i] += (int64_t)h->delta[i];
        total += (int64_t)h->delta[i];
    }
    return total;
}

int igd_b200_count_device(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs,
                          const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    return igdb_dev_count_device(h->dev, n, d_ctg, d_qs, d_qe, filter_on(h, v), v, d_hits, stream);
}

int igd_b200_count_device_tiles(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs,
                                const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream, int32_t tile_lo,
                                int32_t tile_hi)
{
    g_err[0] = 0;
    return igdb_dev_count_device_tiles(h->dev, n, d_ctg, d_qs, d_qe, filter_on(h, v), v, d_hits, stream, tile_lo, tile_hi);
}

/* K8: the shard plan of a multi-GPU search, on the device (igd_dev_shard.cu) */
int igd_b200_shard_cost(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        uint64_t *d_tile_cost, void *stream)
{
    g_err[0] = 0;
    return igdb_dev_shard_cost(h->dev, n, d_ctg, d_qs, d_qe, d_tile_cost, stream);
}

int igd_b200_shard_cuts(igd_b200_t *h, const uint64_t *d_tile_cost, int32_t n_parts, int32_t *d_cuts, void *stream)
{
    g_err[0] = 0;
    return igdb_dev_shard_cuts(h->dev, d_tile_cost, n_parts, d_cuts, stream);
}

int igd_b200_shard_partition(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             const int32_t *d_cuts, int32_t n_parts, int32_t *d_out_ctg, int32_t *d_out_qs,
                             int32_t *d_out_qe, int64_t *d_part_off, int32_t *d_reach, void *stream)
{
    g_err[0] = 0;
    return igdb_dev_shard_partition(h->dev, n, d_ctg, d_qs, d_qe, d_cuts, n_parts, d_out_ctg, d_out_qs, d_out_qe,
                                    d_part_off, d_reach, stream);
}

int igd_b200_count_device_shaped(igd_b200_t *h, int64_t n_max, const uint64_t *d_shape, const int32_t *d_ctg,
                                 const int32_t *d_qs, const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    return igdb_dev_count_device_dyn(h->dev, n_max, d_shape, d_ctg, d_qs, d_qe, filter_on(h, v), v, d_hits, stream);
}

int igd_b200_count_device_finish(igd_b200_t *h)
{
    g_err[0] = 0;
    return igdb_dev_count_finish(h->dev);
}

int64_t get_overlaps_n(igd_b200_t *h, int64_t n, const char *const *chrm, const int32_t *qs,
                       const int32_t *qe, int32_t v, int64_t *hits)
{
    if (n <= 0) return 0;
    int32_t *ctg = malloc(sizeof(int32_t) * (size_t)n);
    int32_t last = -1;
    for (int64_t i = 0; i < n; i++) {
        if (last >= 0 && strcmp(h->cname[last], chrm[i]) == 0) ctg[i] = last;
        else ctg[i] = last = igd_b200_ctg_id(h, chrm[i]);
    }
    int64_t r = igd_b200_count(h, n, ctg, qs, qe, v, hits);
    free(ctg);
    return r;
}

int64_t get_overlaps_n32(igd_b200_t *h, int64_t n, const char *const *chrm, const int32_t *qs,
                         const int32_t *qe, int32_t *hits)
{
    int64_t *wide = calloc((size_t)h->nfiles + 1, sizeof(int64_t));
    int64_t r = get_overlaps_n(h, n, chrm, qs, qe, 0, wide);
    if (r >= 0)
        for (int i = 0; i < h->nfiles; i++) hits[i] = (int32_t)((uint32_t)hits[i] + (uint32_t)wide[i]);
    free(wide);
    return r;
}

int32_t igd_b200_bin_data(igd_b200_t *h, int32_t ctg, int32_t bin, int32_t cap, int32_t *idx, int32_t *start,
                          int32_t *end, int32_t *value)
{
    g_err[0] = 0;
    if (ctg < 0 || ctg >= h->nctg || bin < 0 || bin >= h->ntile[ctg]) return 0;
    int64_t t = h->tbase[ctg] + bin;
    if (cap > 0 && igdb_dev_tile_read(h->dev, t, cap, idx, start, end, value) < 0) return -1;
    return h->ncnt[t];
}

int igd_b200_full(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                  igd_b200_hitlist *out)
{
    g_err[0] = 0;
    memset(out, 0, sizeof *out);
    out->n_queries = n;
    out->total = igdb_dev_full_host(h->dev, n, ctg, qs, qe, &out->query_off, &out->start, &out->end, &out->file,
                                    h->t_ms);
    if (out->total < 0) { igd_b200_hitlist_free(out); return -1; }
    return 0;
}

int64_t igd_b200_full_device(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             void *stream, int64_t **d_query_off, int32_t **d_lists)
{
    g_err[0] = 0;
    return igdb_dev_full_device(h->dev, n, d_ctg, d_qs, d_qe, stream, d_query_off, d_lists);
}

/* the arrays belong to the handle's pinned arena: nothing to release, the struct is just cleared */
void igd_b200_hitlist_free(igd_b200_hitlist *l)
{
    if (!l) return;
    memset(l, 0, sizeof *l);
}

int igd_b200_full_restore(igd_b200_t *h, int64_t n, const int64_t *d_order, const int64_t *d_cnt,
                          const int32_t *d_lists_sorted, int64_t total, int64_t *d_query_off, int32_t *d_lists_out,
                          void *stream)
{
    g_err[0] = 0;
    return igdb_dev_full_restore(h->dev, n, d_order, d_cnt, d_lists_sorted, total, d_query_off, d_lists_out, stream);
}

int64_t igd_b200_hitsmap(igd_b200_t *h, int32_t v, uint32_t *hitmap)
{
    g_err[0] = 0;
    return igdb_dev_hitsmap_host(h->dev, v > 0, v, 0, h->ntiles, hitmap);
}

int64_t igd_b200_hitsmap_tiles(igd_b200_t *h, int32_t v, int64_t tile_lo, int64_t tile_hi, uint32_t *hitmap)
{
    g_err[0] = 0;
    return igdb_dev_hitsmap_host(h->dev, v > 0, v, tile_lo, tile_hi, hitmap);
}

int igd_b200_hitsmap_tiles_device(igd_b200_t *h, int32_t v, int64_t tile_lo, int64_t tile_hi, uint32_t *d_hitmap,
                                  void *stream)
{
    g_err[0] = 0;
    return igdb_dev_hitsmap_device(h->dev, v > 0, v, tile_lo, tile_hi, d_hitmap, stream);
}

/* estimated scan bytes per query: 4 * sum of nCnt over the bins the query touches */
int igd_b200_query_cost(const igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs,
                        const int32_t *qe, int64_t *cost)
{
    for (int64_t i = 0; i < n; i++) {
        int64_t c = 16;
        int32_t k = ctg[i];
        if (k >= 0 && k < h->nctg) {
            int32_t n1 = qs[i] / h->nbp, n2 = (qe[i] - 1) / h->nbp, mT = h->ntile[k] - 1;
            if (n1 >= 0 && n1 <= mT) {
                if (n2 > mT) n2 = mT;
                const int32_t *nc = h->ncnt + h->tbase[k];
                for (int32_t j = n1; j <= (n2 > n1 ? n2 : n1); j++) c += 8 + 4 * (int64_t)nc[j];
            }
        }
        cost[i] = c;
    }
    return 0;
}

/* ---- query ingest ----------------------------------------------------------------------------------
 * parse_bed's acceptance rule (src/igd_base.c:53-72): split on tabs, contig = field 0, start/end =
 * atol of fields 1/2, accept iff >= 3 fields, contig begins "chr" and is shorter than 40 chars, end > 0.
 * Lines end at '\n' with one trailing '\r' dropped (src/kseq.h:127); gzopen reads plain text too. */
static int accept_query_line(char *s, char **ctg_out, int32_t *st_out, int32_t *en_out)
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
    *ctg_out = ctg, *st_out = st, *en_out = en;
    return nf >= 3 && strncmp(ctg, "chr", 3) == 0 && strlen(ctg) < 40 && en > 0;
}

static void queries_push(igd_b200_queries *Q, int64_t *cap, int *ncap, int *last, char *ctg, int32_t st, int32_t en)
{
    int id = -1;
    if (*last >= 0 && strcmp(Q->names[*last], ctg) == 0) id = *last;
    else
        for (int k = 0; k < Q->n_names; k++)
            if (strcmp(Q->names[k], ctg) == 0) { id = k; break; }
    if (id < 0) {
        if (Q->n_names == *ncap) Q->names = realloc(Q->names, sizeof(char *) * (size_t)(*ncap *= 2));
        id = Q->n_names++;
        Q->names[id] = strdup(ctg);
    }
    *last = id;
    if (Q->n == *cap) {
        *cap *= 2;
        Q->name_id = realloc(Q->name_id, sizeof(int32_t) * (size_t)*cap);
        Q->qs = realloc(Q->qs, sizeof(int32_t) * (size_t)*cap);
        Q->qe = realloc(Q->qe, sizeof(int32_t) * (size_t)*cap);
    }
    Q->name_id[Q->n] = id, Q->qs[Q->n] = st, Q->qe[Q->n] = en;
    Q->n++;
}

igd_b200_queries *igd_b200_read_bed(const char *path)
{
    gzFile fp = gzopen(path, "r");
    if (!fp) return NULL;
    gzbuffer(fp, 1 << 20);
    igd_b200_queries *Q = calloc(1, sizeof *Q);
    int64_t cap = 4096;
    int ncap = 64, last = -1;
    Q->name_id = malloc(sizeof(int32_t) * (size_t)cap);
    Q->qs = malloc(sizeof(int32_t) * (size_t)cap);
    Q->qe = malloc(sizeof(int32_t) * (size_t)cap);
    Q->names = malloc(sizeof(char *) * (size_t)ncap);
    size_t lcap = 1 << 12, llen = 0;
    char *line = malloc(lcap);
    static const int BUFSZ = 1 << 20;
    char *buf = malloc(BUFSZ);
    for (;;) {
        int got = gzread(fp, buf, BUFSZ);
        int eof = got <= 0;
        if (eof) got = 0;
        char *p = buf, *end = buf + got;
        while (p < end || (eof && llen)) {
            char *nl = p < end ? memchr(p, '\n', (size_t)(end - p)) : NULL;
            size_t take = nl ? (size_t)(nl - p) : (size_t)(end - p);
            if (llen + take + 2 > lcap) {
                while (llen + take + 2 > lcap) lcap *= 2;
                line = realloc(line, lcap);
            }
            memcpy(line + llen, p, take);
            llen += take;
            p += take;
            if (!nl && !eof) break;                 /* partial line: wait for more input */
            if (nl) p++;
            if (llen && line[llen - 1] == '\r') llen--;
            line[llen] = '\0';
            char *ctg;
            int32_t st, en;
            if (accept_query_line(line, &ctg, &st, &en)) queries_push(Q, &cap, &ncap, &last, ctg, st, en);
            llen = 0;
            if (!nl) break;
        }
        if (eof) break;
    }
    free(buf);
    free(line);
    gzclose(fp);
    return Q;
}

void igd_b200_queries_free(igd_b200_queries *q)
{
    if (!q) return;
    for (int i = 0; i < q->n_names; i++) free(q->names[i]);
    free(q->names); free(q->name_id); free(q->qs); free(q->qe); free(q);
}

void igd_b200_resolve(const igd_b200_t *h, const igd_b200_queries *q, int32_t *out)
{
    int32_t *map = malloc(sizeof(int32_t) * (size_t)(q->n_names + 1));
    for (int k = 0; k < q->n_names; k++) map[k] = igd_b200_ctg_id(h, q->names[k]);
    for (int64_t i = 0; i < q->n; i++) out[i] = map[q->name_id[i]];
    free(map);
}

/* Whole-file count search -- getOverlaps / getOverlaps_v (src/igd_search.c:696-719, 746-769).  The query file is
 * streamed through two halves of a pinned buffer in chunks of whole lines; each chunk goes to the GPU as TEXT
 * and is parsed there (k_parse_bed: parse_bed's acceptance rule and get_id), so the host touches every byte
 * once (gzread) and the search of chunk k overlaps the reading of chunk k + 1. */
static double mono_s(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* A query file is read through zlib like the reference does (gzopen reads plain text too) -- unless it is a
 * regular file that does not start with the gzip magic: then its chunks are pread by several threads straight
 * into the pinned buffer (one thread through gzread copies at ~6 GB/s, which was the whole cost of search()). */
typedef struct { gzFile gz; int fd; off_t off, size; } qsrc_t;
static int qsrc_open(qsrc_t *q, const char *path)
{
    q->gz = NULL, q->fd = -1, q->off = 0, q->size = 0;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    unsigned char magic[2] = {0, 0};
    if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && !(pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b)) {
        q->fd = fd, q->size = st.st_size;
        return 0;
    }
    close(fd);
    q->gz = gzopen(path, "r");
    if (!q->gz) return -1;
    gzbuffer(q->gz, 1 << 20);
    return 0;
}
static int qsrc_read(qsrc_t *q, char *dst, size_t cap)      /* bytes read; <= 0 at the end */
{
    if (q->gz) return gzread(q->gz, dst, (unsigned)cap);
    off_t left = q->size - q->off;
    size_t len = left < (off_t)cap ? (size_t)(left > 0 ? left : 0) : cap;
    if (len == 0 || read_parallel(q->fd, dst, len, q->off)) return 0;
    q->off += (off_t)len;
    return (int)len;
}
static void qsrc_close(qsrc_t *q)
{
    if (q->gz) gzclose(q->gz);
    if (q->fd >= 0) close(q->fd);
}

static int64_t search_mode(igd_b200_t *h, const char *qFile, int mode_v, int32_t v, int64_t *hits)
{
    g_err[0] = 0;
    qsrc_t src;
    if (qsrc_open(&src, qFile)) return 0;           /* unopenable query file -> 0 (src/igd_search.c:701-702) */
    const size_t CH = (size_t)16 << 20;
    char *base = igdb_dev_pinned(h->dev, 2 * CH + 16);
    int64_t r = -1;
    if (!base) { set_err("%s: cannot allocate the pinned query buffer", qFile); goto out; }
    if (igdb_dev_hits_begin(h->dev)) goto out;
    size_t carry = 0;
    int skipping = 0;                               /* inside a line longer than a chunk: drop up to its '\n' */
    const int dbg = getenv("IGD_B200_DEBUG") != NULL;
    double t_read = 0, t_dev = 0, t0 = 0;
    for (int k = 0;; k++) {
        char *buf = base + (size_t)(k & 1) * (CH + 8);
        if (dbg) t0 = mono_s();
        int got = qsrc_read(&src, buf + carry, CH - carry);
        if (dbg) t_read += mono_s() - t0;
        const int eof = got <= 0;
        if (eof) got = 0;
        size_t have = carry + (size_t)got, from = 0;
        if (skipping) {
            char *nl = memchr(buf, '\n', have);
            if (!nl) { carry = 0; if (eof) break; continue; }
            from = (size_t)(nl - buf) + 1, skipping = 0;
        }
        size_t use = have;                          /* end of the last whole line */
        if (!eof) {
            while (use > from && buf[use - 1] != '\n') use--;
            /* no newline in a full chunk: the chunk is the HEAD of one line and stands for it, the rest is skipped up
             * to its newline.  Same hits as the reference as long as the line's first three fields end inside these
             * 16 MiB (the reference itself sees lines through kseq's growing buffer and has no such limit; a contig
             * name of < 40 characters and two integers do). */
            if (use == from && have - from == CH) {
                use = have;
                skipping = 1;
            }
        }
        if (dbg) t0 = mono_s();
        if (use > from && igdb_dev_count_text(h->dev, buf + from, use - from, mode_v, v, NULL)) goto out;
        if (dbg) t_dev += mono_s() - t0;
        carry = have - use;
        if (carry) memcpy(base + (size_t)((k + 1) & 1) * (CH + 8), buf + use, carry);
        if (eof) break;
    }
    if (dbg) t0 = mono_s();
    if (igdb_dev_hits_end(h->dev, h->delta)) goto out;
    if (dbg)
        fprintf(stderr, "[igd_b200] search(%s): read %.3f ms, count_text calls %.3f ms, hits_end %.3f ms\n", qFile,
                t_read * 1e3, t_dev * 1e3, (mono_s() - t0) * 1e3);
    r = 0;
    for (int i = 0; i < h->nfiles; i++) {
        hits[i] += (int64_t)h->delta[i];
        r += (int64_t)h->delta[i];
    }
out:
    qsrc_close(&src);
    return r;
}

int64_t search(igd_b200_t *h, const char *qFile, int32_t v, int64_t *hits)
{
    return search_mode(h, qFile, filter_on(h, v), v, hits);
}

/* ---- the multi-GPU exchange step (igd_dev_p2p.cu) ---------------------------------------------------------- */
/* the partition fused with its all-to-all over peer memory (igd_dev_shard.cu) */
igd_b200_dispatcher *igd_b200_dispatcher_create(igd_b200_t *h, int rank, int world, int64_t cap)
{
    g_err[0] = 0;
    return igdb_dispatcher_create(h->dev, rank, world, cap);
}
int igd_b200_dispatcher_handle(igd_b200_dispatcher *x, void *out64) { g_err[0] = 0; return igdb_dispatcher_handle(x, out64); }
int igd_b200_dispatcher_connect(igd_b200_dispatcher *x, const void *handles) { g_err[0] = 0; return igdb_dispatcher_connect(x, handles); }
int igd_b200_dispatch(igd_b200_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                      const int32_t *d_cuts, void *stream, int32_t **d_out_ctg, int32_t **d_out_qs, int32_t **d_out_qe,
                      int64_t *m, int32_t *tile_lo, int32_t *tile_hi)
{
    g_err[0] = 0;
    return igdb_dispatch(x, n, d_ctg, d_qs, d_qe, d_cuts, stream, d_out_ctg, d_out_qs, d_out_qe, m, tile_lo, tile_hi);
}
int igd_b200_dispatch_async(igd_b200_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                            const int32_t *d_cuts, void *stream, int32_t **d_out_ctg, int32_t **d_out_qs,
                            int32_t **d_out_qe, const uint64_t **d_shape)
{
    g_err[0] = 0;
    return igdb_dispatch_async(x, n, d_ctg, d_qs, d_qe, d_cuts, stream, d_out_ctg, d_out_qs, d_out_qe, d_shape);
}
int igd_b200_dispatcher_status(igd_b200_dispatcher *x) { g_err[0] = 0; return igdb_dispatcher_status(x); }
void igd_b200_dispatcher_destroy(igd_b200_dispatcher *x) { igdb_dispatcher_destroy(x); }

/* ---- one multi-GPU count step as ONE call (no counterpart in the reference, which is one process): everything
 * between "this rank's 1/N of the queries is in its HBM" and "hits[] of the whole query set is in its HBM" queued on
 * one stream with no host round trip and no library collective -- per-tile cost, its sum over the ranks (the
 * reducer's kernel on nTiles + 1 words), cuts, the query exchange (igd_b200_dispatch_async), the search on the shard
 * (igd_b200_count_device_shaped), the sum of hits[] over the ranks. ------------------------------------------- */
struct igd_b200_sharded {
    igd_b200_t *h;
    int rank, world;
    int64_t cap;
    igdb_dispatcher *disp;
    igdb_reducer *red_hits, *red_cost;
    uint64_t *d_cost, *d_part;      /* [ntiles + 1], [nfiles] */
    int32_t *d_cuts;                /* [world + 1] */
};

void igd_b200_sharded_destroy(igd_b200_sharded *s)
{
    if (!s) return;
    if (s->disp) igdb_dispatcher_destroy(s->disp);
    if (s->red_hits) igdb_reducer_destroy(s->red_hits);
    if (s->red_cost) igdb_reducer_destroy(s->red_cost);
    igdb_dev_free(s->h->dev, s->d_cost);
    igdb_dev_free(s->h->dev, s->d_part);
    igdb_dev_free(s->h->dev, s->d_cuts);
    free(s);
}

igd_b200_sharded *igd_b200_sharded_create(igd_b200_t *h, int device, int rank, int world, int64_t cap)
{
    g_err[0] = 0;
    if (h->ntiles + 1 > 0x7fffffff) { set_err("%s", "sharded: too many tiles"); return NULL; }
    if (cap < 1 || cap > ((int64_t)1 << 26)) { set_err("%s", "sharded: receive buffers of 1 .. 2^26 queries"); return NULL; }
    igd_b200_sharded *s = calloc(1, sizeof *s);
    if (!s) return NULL;
    s->h = h, s->rank = rank, s->world = world, s->cap = cap;
    s->disp = igdb_dispatcher_create(h->dev, rank, world, cap);
    s->red_hits = igdb_reducer_create(device, rank, world, h->nfiles);
    s->red_cost = igdb_reducer_create(device, rank, world, (int32_t)(h->ntiles + 1));
    s->d_cost = igdb_dev_malloc(h->dev, (size_t)(h->ntiles + 1) * 8);
    s->d_part = igdb_dev_malloc(h->dev, ((size_t)h->nfiles + 1) * 8);
    s->d_cuts = igdb_dev_malloc(h->dev, ((size_t)world + 1) * 4);
    if (!s->disp || !s->red_hits || !s->red_cost || !s->d_cost || !s->d_part || !s->d_cuts) {
        if (!g_err[0] && !igdb_dev_error()[0]) set_err("%s", "sharded: out of device memory");
        igd_b200_sharded_destroy(s);
        return NULL;
    }
    return s;
}

/* 3 x 64 bytes: the CUDA IPC handles of this rank's dispatcher block, hits[] mailbox and cost mailbox */
int igd_b200_sharded_handles(igd_b200_sharded *s, void *out192)
{
    g_err[0] = 0;
    char *o = out192;
    return igdb_dispatcher_handle(s->disp, o) || igdb_reducer_handle(s->red_hits, o + 64) ||
           igdb_reducer_handle(s->red_cost, o + 128) ? -1 : 0;
}

/* handles: world x 192 bytes, rank-major */
int igd_b200_sharded_connect(igd_b200_sharded *s, const void *handles)
{
    g_err[0] = 0;
    if (s->world == 1) return 0;
    char *t = malloc((size_t)s->world * 64);
    if (!t) return -1;
    int rc = 0;
    for (int k = 0; k < 3 && !rc; k++) {
        for (int p = 0; p < s->world; p++) memcpy(t + 64 * (size_t)p, (const char *)handles + 192 * (size_t)p + 64 * k, 64);
        rc = k == 0 ? igdb_dispatcher_connect(s->disp, t)
                    : igdb_reducer_connect(k == 1 ? s->red_hits : s->red_cost, t);
    }
    free(t);
    return rc ? -1 : 0;
}

/* d_hits[nFiles] (device, uint64) = hits[] of ALL ranks' queries (overwritten, not incremented); the n queries are this
 * rank's part of the set, device arrays in any order.  Enqueued on `stream`; returns once the search has been planned
 * (igd_b200_count_device_finish's rule).  Every rank must call it the same number of times. */
int igd_b200_sharded_count(igd_b200_sharded *s, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                           int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    igdb_dev *dev = s->h->dev;
    int32_t *c, *qs, *qe;
    const uint64_t *shape;
    if (n > s->cap) { set_err("%s", "sharded_count: more queries than the receive buffers hold"); return -1; }
    if (!stream) stream = igdb_dev_stream(dev);             /* one stream for every launch of the step */
    if (igdb_dev_memset_async(dev, s->d_cost, (size_t)(s->h->ntiles + 1) * 8, stream)) return -1;
    if (igdb_dev_shard_cost(dev, n, d_ctg, d_qs, d_qe, s->d_cost, stream)) return -1;
    if (igdb_reducer_allreduce(s->red_cost, (unsigned long long *)s->d_cost, stream)) return -1;
    if (igdb_dev_shard_cuts(dev, s->d_cost, s->world, s->d_cuts, stream)) return -1;
    if (igdb_dispatch_async(s->disp, n, d_ctg, d_qs, d_qe, s->d_cuts, stream, &c, &qs, &qe, &shape)) return -1;
    if (igdb_dev_count_device_dyn(dev, s->cap, shape, c, qs, qe, filter_on(s->h, v), v, s->d_part, stream)) return -1;
    if (igdb_dev_count_finish(dev)) return -1;
    return igdb_reducer_allreduce_into(s->red_hits, (unsigned long long *)s->d_part, (unsigned long long *)d_hits, stream);
}

/* waits for the device; 0 = every step so far met all its peers and fitted its buffers */
int igd_b200_sharded_status(igd_b200_sharded *s)
{
    g_err[0] = 0;
    if (igdb_dispatcher_status(s->disp)) return -1;
    int a = igdb_reducer_status(s->red_hits), b = igdb_reducer_status(s->red_cost);
    if (a || b) { set_err("%s", "sharded: an exchange of hits[] or of the tile costs missed a peer"); return -1; }
    return 0;
}

igd_b200_reducer *igd_b200_reducer_create(int device, int rank, int world, int32_t nfiles)
{
    g_err[0] = 0;
    return igdb_reducer_create(device, rank, world, nfiles);
}
int igd_b200_reducer_handle(igd_b200_reducer *r, void *out64) { g_err[0] = 0; return igdb_reducer_handle(r, out64); }
int igd_b200_reducer_connect(igd_b200_reducer *r, const void *handles) { g_err[0] = 0; return igdb_reducer_connect(r, handles); }
int igd_b200_reducer_allreduce(igd_b200_reducer *r, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    return igdb_reducer_allreduce(r, (unsigned long long *)d_hits, stream);
}
int igd_b200_reducer_allreduce_into(igd_b200_reducer *r, uint64_t *d_in, uint64_t *d_out, void *stream)
{
    g_err[0] = 0;
    return igdb_reducer_allreduce_into(r, (unsigned long long *)d_in, (unsigned long long *)d_out, stream);
}
int igd_b200_reducer_status(igd_b200_reducer *r) { g_err[0] = 0; return igdb_reducer_status(r); }
void igd_b200_reducer_destroy(igd_b200_reducer *r) { igdb_reducer_destroy(r); }

/* ---- (2) reference-named entry points over the current handle ---------------------------------------- */
void igd_b200_use(igd_b200_t *h) { g_current = h; }
igd_b200_t *igd_b200_current(void) { return g_current; }

/* get_overlaps always returns 0: its counter is never incremented (src/igd_search.c:461,533) */
int32_t get_overlaps(char *chrm, int32_t qs, int32_t qe, int64_t *hits)
{
    if (!g_current) return 0;
    int32_t c = igd_b200_ctg_id(g_current, chrm);
    igd_b200_count(g_current, 1, &c, &qs, &qe, 0, hits);
    return 0;
}

/* get_overlaps_v returns the number of hits (src/igd_search.c:653,683,693).  Called directly it applies
 * its own control flow and value >= v for ANY v (the v > 0 dispatch lives in igd_search, :1027,:1045). */
int32_t get_overlaps_v(char *chrm, int32_t qs, int32_t qe, int32_t v, int64_t *hits)
{
    if (!g_current) return 0;
    int32_t c = igd_b200_ctg_id(g_current, chrm);
    int64_t r = count_mode(g_current, 1, &c, &qs, &qe, 1, v, hits);
    return r < 0 ? 0 : (int32_t)r;
}

int64_t getOverlaps(char *qFile, int64_t *hits)
{
    if (!g_current) return 0;
    search(g_current, qFile, 0, hits);
    return 0;                                       /* sum of get_overlaps' return values */
}

int64_t getOverlaps_v(char *qFile, int64_t *hits, int32_t v)
{
    if (!g_current) return 0;
    int64_t r = search_mode(g_current, qFile, 1, v, hits);
    return r < 0 ? 0 : r;
}

/* gType-0 twins (src/igd_search.c:30-101, 210-230): the same control flow over gdata0_t records.  The handle
 * knows its record type, so they are the same searches; like the reference's they return 0. */
int32_t get_overlaps0(char *chrm, int32_t qs, int32_t qe, int64_t *hits) { return get_overlaps(chrm, qs, qe, hits); }
int64_t getOverlaps0(char *qFile, int64_t *hits) { return getOverlaps(qFile, hits); }

static int64_t get_map(uint32_t **hitmap, int filter, int32_t v);
int64_t getMap(uint32_t **hitmap) { return get_map(hitmap, 0, 0); }
/* getMap_v keeps the pairs whose two values exceed v for ANY v (src/igd_search.c:856,872); it is the CLI that calls
 * it only when v > 0 (:1000) */
int64_t getMap_v(uint32_t **hitmap, int32_t v) { return get_map(hitmap, 1, v); }

static int64_t get_map(uint32_t **hitmap, int filter, int32_t v)
{
    if (!g_current) return 0;
    int32_t n = g_current->nfiles;
    uint32_t *dense = malloc(sizeof(uint32_t) * (size_t)n * (size_t)n + 4);
    for (int i = 0; i < n; i++) memcpy(dense + (size_t)i * n, hitmap[i], sizeof(uint32_t) * (size_t)n);
    g_err[0] = 0;
    int64_t r = igdb_dev_hitsmap_host(g_current->dev, filter, v, 0, g_current->ntiles, dense);
    if (r >= 0)
        for (int i = 0; i < n; i++) memcpy(hitmap[i], dense + (size_t)i * n, sizeof(uint32_t) * (size_t)n);
    free(dense);
    return r < 0 ? 0 : r;
}

/* ---- seq_overlaps (src/igd_search.c:253-353) ------------------------------------------------------------ */
int64_t igd_b200_seq_overlaps(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                              int64_t **query_off, igd_b200_overlap **list)
{
    g_err[0] = 0;
    void *p = NULL;
    int64_t total = igdb_dev_seq_pairs_host(h->dev, n, ctg, qs, qe, query_off, &p);
    *list = p;
    return total;
}

/* one query, appended to the caller's growable list (EXPAND doubles the capacity, 2 if empty) */
void seq_overlaps(char *chrm, int32_t qs, int32_t qe, igd_b200_overlaps *olp)
{
    if (!g_current || !olp) return;
    int32_t c = igd_b200_ctg_id(g_current, chrm);
    int64_t *qoff;
    igd_b200_overlap *list;
    int64_t total = igd_b200_seq_overlaps(g_current, 1, &c, &qs, &qe, &qoff, &list);
    if (total <= 0) return;
    if ((int64_t)olp->nn + total > olp->mm) {
        int32_t mm = olp->mm;
        while ((int64_t)olp->nn + total > mm) mm = mm ? mm << 1 : 2;
        olp->olist = realloc(olp->olist, sizeof(igd_b200_overlap) * (size_t)mm);
        olp->mm = mm;
    }
    memcpy(olp->olist + olp->nn, list, sizeof(igd_b200_overlap) * (size_t)total);
    olp->nn += (int32_t)total;
}

/* ---- -s: seqOverlaps (src/igd_search.c:355-451) ------------------------------------------------------
 * Host part = the reference's query container: readBED (src/igd_base.c:628-649) keeps the lines parse_bed
 * accepts, ailist_add (:601-626) drops s > e (as unsigned) and groups by chrom NAME in first-appearance order
 * of the surviving lines; each group is then sorted by start (qsort + compare_qstart, :373, :23-28; glibc's
 * qsort is a stable merge sort, reproduced here with the line ordinal as tie-break).  The device does the
 * rest; the final ratio of :445-448 is applied here. */
typedef struct { int32_t start, end; int64_t ord; } seq_q;
static int seq_q_cmp(const void *a, const void *b)
{
    const seq_q *x = a, *y = b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    return x->ord < y->ord ? -1 : x->ord > y->ord;
}

int igd_b200_seqpare(igd_b200_t *h, const char *qFile, double *sm)
{
    g_err[0] = 0;
    igd_b200_queries *Q = igd_b200_read_bed(qFile);
    if (!Q) { set_err("Can't open file %s", qFile); return -1; }
    int rc = 0;
    int32_t *gid = malloc(sizeof(int32_t) * (size_t)(Q->n_names + 1));      /* name id -> group */
    int32_t *gname = malloc(sizeof(int32_t) * (size_t)(Q->n_names + 1));    /* group -> name id */
    int64_t *gcnt = calloc((size_t)Q->n_names + 1, sizeof(int64_t));
    int32_t ng = 0;
    int64_t Nq = 0;
    for (int k = 0; k < Q->n_names; k++) gid[k] = -1;
    for (int64_t i = 0; i < Q->n; i++) {
        if ((uint32_t)Q->qs[i] > (uint32_t)Q->qe[i]) continue;
        int k = Q->name_id[i];
        if (gid[k] < 0) { gid[k] = ng; gname[ng++] = k; }
        gcnt[gid[k]]++;
        Nq++;
    }
    int64_t *goff = malloc(sizeof(int64_t) * (size_t)(ng + 1));
    goff[0] = 0;
    for (int g = 0; g < ng; g++) goff[g + 1] = goff[g] + gcnt[g];
    seq_q *L = malloc(sizeof(seq_q) * (size_t)(Nq + 1));
    int64_t *fill = calloc((size_t)ng + 1, sizeof(int64_t));
    for (int64_t i = 0; i < Q->n; i++) {
        if ((uint32_t)Q->qs[i] > (uint32_t)Q->qe[i]) continue;
        int g = gid[Q->name_id[i]];
        seq_q *p = &L[goff[g] + fill[g]];
        p->start = Q->qs[i], p->end = Q->qe[i], p->ord = fill[g]++;
    }
    int32_t *ctg = malloc(sizeof(int32_t) * (size_t)(Nq + 1)), *qs = malloc(sizeof(int32_t) * (size_t)(Nq + 1));
    int32_t *qe = malloc(sizeof(int32_t) * (size_t)(Nq + 1)), *grp = malloc(sizeof(int32_t) * (size_t)(Nq + 1));
    int32_t *gnq = malloc(sizeof(int32_t) * (size_t)(ng + 1));
    for (int g = 0; g < ng; g++) {
        qsort(L + goff[g], (size_t)gcnt[g], sizeof(seq_q), seq_q_cmp);
        int32_t c = igd_b200_ctg_id(h, Q->names[gname[g]]);
        gnq[g] = (int32_t)gcnt[g];
        for (int64_t i = goff[g]; i < goff[g + 1]; i++) ctg[i] = c, qs[i] = L[i].start, qe[i] = L[i].end, grp[i] = g;
    }
    double *acc = calloc((size_t)h->nfiles + 1, sizeof(double));
    if (igdb_dev_seqpare(h->dev, Nq, ctg, qs, qe, grp, ng, gnq, acc)) rc = -1;
    else
        for (int m = 0; m < h->nfiles; m++) sm[m] = acc[m] / ((int)Nq + h->fnr[m] - acc[m]);
    free(acc); free(gnq); free(grp); free(qe); free(qs); free(ctg); free(fill); free(L); free(goff);
    free(gcnt); free(gname); free(gid);
    igd_b200_queries_free(Q);
    return rc;
}
