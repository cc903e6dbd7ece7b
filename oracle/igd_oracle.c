/* TEST INFRASTRUCTURE ONLY -- see igd_oracle.h.  Not linked into, imported by or executed from the
 * product path.  Each function cites the reference code (under /root/reference/) whose behaviour it
 * restates, quirks included. */
#define _GNU_SOURCE
#include "igd_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct { int32_t idx, start, end, value; } rec_t;       /* src/igd_base.h:41-46 */

struct orc_db {
    int32_t nbp, gtype, nctg, nfiles;
    int32_t *ntile;         /* [nctg] */
    int64_t *tbase;         /* [nctg+1] global tile index of each contig's tile 0 */
    int32_t *ncnt;          /* [total tiles] */
    int64_t *toff;          /* [total tiles + 1] record offset of each tile */
    char (*cname)[40];
    rec_t *rec;
    int64_t nrec;
    char **fname;
    int32_t *fnr;
};

/* ---- .igd header + payload: writer src/igd_base.c:396-461, reader :269-323 (Appendix A) ---------- */
orc_db *orc_open(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    orc_db *db = calloc(1, sizeof *db);
    int32_t h[3];
    if (fread(h, 4, 3, f) != 3) goto fail;
    db->nbp = h[0], db->gtype = h[1], db->nctg = h[2];
    /* gType 1 = gdata_t (16 B), gType 0 = gdata0_t (12 B, src/igd_base.h:48-52); record size as in
     * get_igdinfo, src/igd_base.c:279-283 */
    if ((db->gtype != 0 && db->gtype != 1) || db->nctg <= 0 || db->nbp <= 0) goto fail;
    int32_t m = db->nctg;
    db->ntile = malloc(4 * (size_t)m);
    if (fread(db->ntile, 4, m, f) != (size_t)m) goto fail;
    db->tbase = malloc(8 * (size_t)(m + 1));
    db->tbase[0] = 0;
    for (int i = 0; i < m; i++) db->tbase[i + 1] = db->tbase[i] + db->ntile[i];
    int64_t T = db->tbase[m];
    db->ncnt = malloc(4 * (size_t)T + 4);
    if (T && fread(db->ncnt, 4, T, f) != (size_t)T) goto fail;
    db->cname = malloc(40 * (size_t)m);
    if (fread(db->cname, 40, m, f) != (size_t)m) goto fail;
    db->toff = malloc(8 * (size_t)(T + 1));
    db->toff[0] = 0;
    for (int64_t t = 0; t < T; t++) db->toff[t + 1] = db->toff[t] + db->ncnt[t];
    db->nrec = db->toff[T];
    db->rec = malloc(sizeof(rec_t) * (size_t)db->nrec + 16);
    if (db->gtype == 1) {
        if (db->nrec && fread(db->rec, sizeof(rec_t), db->nrec, f) != (size_t)db->nrec) goto fail;
    } else {                                /* gdata0_t {idx,start,end}: kept as rec_t with value 0 */
        for (int64_t i = 0; i < db->nrec; i++) {
            int32_t r3[3];
            if (fread(r3, 4, 3, f) != 3) goto fail;
            db->rec[i].idx = r3[0], db->rec[i].start = r3[1], db->rec[i].end = r3[2], db->rec[i].value = 0;
        }
    }
    fclose(f);
    f = NULL;

    /* <stem>_index.tsv: src/igd_base.c:235-267 -- header line skipped, columns 2 and 3 used */
    {
        size_t L = strlen(path);
        char *tsv = malloc(L + 16);
        memcpy(tsv, path, L + 1);
        char *dot = strrchr(tsv, '.');
        if (!dot) { free(tsv); goto fail; }
        strcpy(dot, "_index.tsv");
        FILE *g = fopen(tsv, "r");
        free(tsv);
        if (!g) goto fail;
        char buf[1024];
        int cap = 16, n = 0;
        db->fname = malloc(sizeof(char *) * cap);
        db->fnr = malloc(4 * cap);
        if (!fgets(buf, sizeof buf, g)) { fclose(g); goto fail; }
        while (fgets(buf, sizeof buf, g)) {
            if (n == cap) {
                cap *= 2;
                db->fname = realloc(db->fname, sizeof(char *) * cap);
                db->fnr = realloc(db->fnr, 4 * cap);
            }
            char *save = NULL;
            strtok_r(buf, "\t", &save);
            char *s1 = strtok_r(NULL, "\t", &save);
            char *s2 = strtok_r(NULL, "\t", &save);
            db->fname[n] = strdup(s1 ? s1 : "");
            db->fnr[n] = s2 ? (int32_t)atol(s2) : 0;
            n++;
        }
        fclose(g);
        db->nfiles = n;
    }
    return db;
fail:
    if (f) fclose(f);
    orc_close(db);
    return NULL;
}

void orc_close(orc_db *db)
{
    if (!db) return;
    for (int i = 0; i < db->nfiles; i++) free(db->fname[i]);
    free(db->fname); free(db->fnr); free(db->ntile); free(db->tbase); free(db->ncnt);
    free(db->toff); free(db->cname); free(db->rec); free(db);
}

int32_t orc_nfiles(const orc_db *db) { return db->nfiles; }
int32_t orc_nctg(const orc_db *db) { return db->nctg; }
int32_t orc_nbp(const orc_db *db) { return db->nbp; }
int32_t orc_gtype(const orc_db *db) { return db->gtype; }
int32_t orc_ntile(const orc_db *db, int32_t c) { return db->ntile[c]; }
int64_t orc_nrecords(const orc_db *db) { return db->nrec; }
int32_t orc_file_nr(const orc_db *db, int32_t i) { return db->fnr[i]; }
const char *orc_file_name(const orc_db *db, int32_t i) { return db->fname[i]; }
const char *orc_ctg_name(const orc_db *db, int32_t c) { return db->cname[c]; }

/* get_id, src/igd_base.c:325-331 (the hash is an implementation detail; exact string match) */
int32_t orc_ctg_id(const orc_db *db, const char *chrm)
{
    for (int i = 0; i < db->nctg; i++)
        if (strcmp(db->cname[i], chrm) == 0) return i;
    return -1;
}

/* ---- query ingest: parse_bed src/igd_base.c:53-72, line reader src/kseq.h:83-130 -------------------- */
static int accept_line(char *s, char **ctg_, int32_t *st_, int32_t *en_)
{
    char *p = s, *q = s, *ctg = NULL;
    int32_t st = -1, en = -1;
    int i = 0;
    for (;; ++q) {
        if (*q == '\t' || *q == '\0') {
            int c = *q;
            *q = 0;
            if (i == 0) ctg = p;
            else if (i == 1) st = (int32_t)atol(p);
            else if (i == 2) en = (int32_t)atol(p);
            ++i, p = q + 1;
            if (c == '\0') break;
        }
    }
    *st_ = st, *en_ = en, *ctg_ = ctg;
    return i >= 3 && ctg[0] == 'c' && ctg[1] == 'h' && ctg[2] == 'r' && strlen(ctg) < 40 && en > 0;
}

orc_queries *orc_read_bed(const char *path)
{
    gzFile fp = gzopen(path, "r");
    if (!fp) return NULL;
    orc_queries *Q = calloc(1, sizeof *Q);
    int64_t cap = 1024;
    int ncap = 32;
    Q->name_id = malloc(4 * cap), Q->qs = malloc(4 * cap), Q->qe = malloc(4 * cap);
    Q->names = malloc(sizeof(char *) * ncap);
    size_t lcap = 1 << 16, llen = 0;
    char *line = malloc(lcap);
    char buf[1 << 16];
    int got, eof = 0, last = -1;
    while (!eof) {
        got = gzread(fp, buf, sizeof buf);
        if (got <= 0) { eof = 1; got = 0; }
        for (int b = 0; b <= got; b++) {
            int endl;
            if (b == got) { if (!eof || llen == 0) break; endl = 1; }    /* final line without '\n' */
            else endl = buf[b] == '\n';
            if (!endl) {
                if (llen + 2 > lcap) line = realloc(line, lcap *= 2);
                line[llen++] = buf[b];
                continue;
            }
            if (llen && line[llen - 1] == '\r') llen--;                   /* kseq.h:127 */
            line[llen] = 0;
            char *ctg; int32_t st, en;
            if (accept_line(line, &ctg, &st, &en)) {
                int id = -1;
                if (last >= 0 && strcmp(Q->names[last], ctg) == 0) id = last;
                else for (int k = 0; k < Q->n_names; k++) if (strcmp(Q->names[k], ctg) == 0) { id = k; break; }
                if (id < 0) {
                    if (Q->n_names == ncap) Q->names = realloc(Q->names, sizeof(char *) * (ncap *= 2));
                    Q->names[id = Q->n_names++] = strdup(ctg);
                }
                last = id;
                if (Q->n == cap) {
                    cap *= 2;
                    Q->name_id = realloc(Q->name_id, 4 * cap);
                    Q->qs = realloc(Q->qs, 4 * cap), Q->qe = realloc(Q->qe, 4 * cap);
                }
                Q->name_id[Q->n] = id, Q->qs[Q->n] = st, Q->qe[Q->n] = en;
                Q->n++;
            }
            llen = 0;
        }
    }
    free(line);
    gzclose(fp);
    return Q;
}

void orc_free_queries(orc_queries *q)
{
    if (!q) return;
    for (int i = 0; i < q->n_names; i++) free(q->names[i]);
    free(q->names); free(q->name_id); free(q->qs); free(q->qe); free(q);
}

void orc_resolve(const orc_db *db, const orc_queries *q, int32_t *out)
{
    int32_t *map = malloc(4 * (size_t)(q->n_names + 1));
    for (int k = 0; k < q->n_names; k++) map[k] = orc_ctg_id(db, q->names[k]);
    for (int64_t i = 0; i < q->n; i++) out[i] = map[q->name_id[i]];
    free(map);
}

/* ---- the per-query walk shared by count / -v / -f / -s -------------------------------------------- */
/* "last index with start < qe", the result of every binary-search variant in the reference
 * (inline loop src/igd_search.c:479-487, bSearch src/igd_base.c:74-94, linear form :645-647); -1 if none */
static int32_t last_start_lt(const rec_t *r, int32_t n, int32_t qe)
{
    int32_t lo = 0, hi = n;                 /* first index with start >= qe */
    while (lo < hi) {
        int32_t mid = lo + (hi - lo) / 2;
        if (r[mid].start < qe) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

static int ceil_log2(int64_t x)             /* ceil(log2(x)), x >= 1 */
{
    int b = 0;
    while (((int64_t)1 << b) < x) b++;
    return b;
}

typedef void (*hit_fn)(void *ctx, int64_t q, int32_t n1, int32_t i, const rec_t *r);

/* Visits, in the reference's order, every record the reference would report for query (c,qs,qe).
 * mode_v = 0: get_overlaps / get_overlaps_f1 / seq_overlaps control flow (src/igd_search.c:454-534,
 *             537-620, 253-353) -- later bins are only reached when bin n1 is non-empty (the quirk).
 * mode_v = 1: get_overlaps_v control flow (src/igd_search.c:623-694) -- no such nesting; value >= v. */
static int64_t walk(const orc_db *db, int64_t q, int32_t c, int32_t qs, int32_t qe, int mode_v, int32_t v,
                    hit_fn fn, void *ctx, orc_stats *st, uint8_t *seen)
{
    int64_t nh = 0;
    if (st) { st->queries++; st->bytes += 16; }
    if (c < 0) return 0;                                            /* unknown chrom :456-458 */
    int32_t nbp = db->nbp;
    int32_t n1 = qs / nbp, n2 = (qe - 1) / nbp, mT = db->ntile[c] - 1;    /* C truncating division :459 */
    if (n1 > mT || n1 < 0) return 0;        /* :462; n1 < 0 is an out-of-bounds read in the reference */
    if (n2 > mT) n2 = mT;
    const int32_t *ncnt = db->ncnt + db->tbase[c];
    const int64_t *toff = db->toff + db->tbase[c];
    if (!mode_v && ncnt[n1] == 0) return 0;                         /* the first-bin-empty quirk :468 */
    for (int32_t j = n1; j <= (n2 > n1 ? n2 : n1); j++) {
        int32_t n = ncnt[j];
        if (n == 0) continue;
        const rec_t *r = db->rec + toff[j];
        int32_t bd = nbp * j;
        int later = j > n1;
        if (st) {
            st->visited++;
            st->bytes += 8;
            if (seen && !seen[db->tbase[c] + j]) {
                seen[db->tbase[c] + j] = 1;
                st->tile_bytes += (int64_t)n * (mode_v ? 16 : 12);
            }
        }
        if (!(qe > r[0].start)) {                                   /* gate :477,:509 */
            if (st) { st->gated++; st->bytes += 4; }
            continue;
        }
        int32_t tS = 0;
        if (later) while (tS < n && r[tS].start < bd) tS++;         /* dedupe rule :510-511 */
        int32_t tL = last_start_lt(r, n, qe);
        int64_t h0 = nh;
        for (int32_t i = tL; i >= tS; i--) {
            if (r[i].end > qs && (!mode_v || r[i].value >= v)) {
                nh++;
                fn(ctx, q, n1, i, &r[i]);
            }
        }
        if (st) {
            int pr = ceil_log2((int64_t)n + 1) * (1 + later);
            int64_t sc = tL >= tS ? tL - tS + 1 : 0;
            st->probes += pr;
            st->scanned += sc;
            st->hits += nh - h0;
            st->bytes += 4 * (int64_t)pr + 4 * sc * (1 + mode_v) + 4 * (nh - h0);
        }
    }
    return nh;
}

/* ---- count / -v : hits[idx]++ (src/igd_search.c:489-493, 522-526, 650-655, 680-685) ------------------ */
static void hit_count(void *ctx, int64_t q, int32_t n1, int32_t i, const rec_t *r)
{
    (void)q; (void)n1; (void)i;
    ((int64_t *)ctx)[r->idx]++;
}

int64_t orc_count(const orc_db *db, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                  int32_t v, int64_t *hits, orc_stats *st)
{
    /* igd_search dispatches on v>0 (src/igd_search.c:1027,1045) -- and on a gType-0 database on nothing:
     * getOverlaps0 / get_overlaps0 (:30-101, same control flow as get_overlaps) whatever -v says (:1024,:1042) */
    int mode_v = v > 0 && db->gtype != 0;
    uint8_t *seen = st ? calloc(db->tbase[db->nctg] + 1, 1) : NULL;
    int64_t ret = 0;
    for (int64_t q = 0; q < n; q++) {
        int64_t h = walk(db, q, ctg[q], qs[q], qe[q], mode_v, v, hit_count, hits, st, seen);
        if (mode_v) ret += h;               /* get_overlaps returns nols == 0 always (:461,:533) */
    }
    free(seen);
    return ret;
}

/* ---- -f : full overlap list (src/igd_search.c:537-620) ---------------------------------------------- */
typedef struct { int64_t cap, n; int32_t *q, *s, *e, *f; } full_ctx;
static void hit_full(void *ctx_, int64_t q, int32_t n1, int32_t i, const rec_t *r)
{
    full_ctx *c = ctx_;
    (void)n1; (void)i;
    if (c->n < c->cap) {
        c->q[c->n] = (int32_t)q, c->s[c->n] = r->start, c->e[c->n] = r->end, c->f[c->n] = r->idx;
    }
    c->n++;
}

int64_t orc_full(const orc_db *db, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                 int64_t cap, int32_t *out_q, int32_t *out_s, int32_t *out_e, int32_t *out_f,
                 int64_t *per_query)
{
    full_ctx c = { cap, 0, out_q, out_s, out_e, out_f };
    for (int64_t q = 0; q < n; q++) {
        int64_t h = walk(db, q, ctg[q], qs[q], qe[q], 0, 0, hit_full, &c, NULL, NULL);
        if (per_query) per_query[q] = h;
    }
    return c.n;
}

/* ---- -s : Seqpare (seq_overlaps src/igd_search.c:253-353, seqOverlaps :355-451) ---------------------- */
typedef struct { int32_t idx_t, idx_g, idx_f; float sm; int32_t ord; } cand_t;   /* overlap_t igd_base.h:109-114 */
typedef struct { cand_t *a; int64_t n, cap; int32_t qs, qe; } cand_ctx;

static void hit_seq(void *ctx_, int64_t q, int32_t n1, int32_t i, const rec_t *r)
{
    cand_ctx *c = ctx_;
    (void)q;
    if (c->n == c->cap) c->a = realloc(c->a, sizeof(cand_t) * (c->cap = c->cap ? c->cap * 2 : 1024));
    /* fp32 throughout, in this order (:255, :291-299): st/(qlen+rlen-st) */
    float qlen = c->qe - c->qs, st, rlen;
    int32_t re = r->end, rs = r->start;
    st = (re < c->qe ? re : c->qe) - (rs > c->qs ? rs : c->qs);
    rlen = re - rs;
    cand_t *p = &c->a[c->n];
    p->idx_g = i, p->idx_f = r->idx, p->idx_t = n1;                 /* idx_t = the QUERY's n1 (:298,:342) */
    p->sm = st / (qlen + rlen - st);
    p->ord = (int32_t)c->n;
    c->n++;
}

/* seq_overlaps for one query as a list: up to cap records {idx_t, idx_g, idx_f, sm} (16 B each) in scan order;
 * returns how many there are */
int64_t orc_seq_overlaps(const orc_db *db, int32_t ctg, int32_t qs, int32_t qe, int64_t cap, void *out16)
{
    cand_ctx cc = { NULL, 0, 0, qs, qe };
    walk(db, 0, ctg, qs, qe, 0, 0, hit_seq, &cc, NULL, NULL);
    struct { int32_t t, g, f; float sm; } *o = out16;
    for (int64_t i = 0; i < cc.n && i < cap; i++)
        o[i].t = cc.a[i].idx_t, o[i].g = cc.a[i].idx_g, o[i].f = cc.a[i].idx_f, o[i].sm = cc.a[i].sm;
    free(cc.a);
    return cc.n;
}

typedef struct { int32_t start, end; int64_t ord; } qrec_t;
static int cmp_qstart(const void *a, const void *b)                /* compare_qstart igd_base.c:23-28, stable */
{
    const qrec_t *x = a, *y = b;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    return x->ord < y->ord ? -1 : x->ord > y->ord;
}
static int cmp_fidx(const void *a, const void *b)                  /* compare_fidx igd_base.c:30-35, stable */
{
    const cand_t *x = a, *y = b;
    if (x->idx_f != y->idx_f) return x->idx_f < y->idx_f ? -1 : 1;
    return x->ord < y->ord ? -1 : x->ord > y->ord;
}

void orc_seqpare(const orc_db *db, const orc_queries *Q, double *sm)
{
    int nf = db->nfiles;
    int64_t Nq = 0;
    for (int m = 0; m < nf; m++) sm[m] = 0.0;
    /* readBED/ailist_add (igd_base.c:601-649): per-chrom lists in first-appearance order; a chrom is only
     * created by a line that survives the unsigned `s > e` drop (:603) */
    int32_t *grp = malloc(4 * (size_t)(Q->n_names + 1));
    int32_t *order = malloc(4 * (size_t)(Q->n_names + 1));
    int ng = 0;
    for (int k = 0; k < Q->n_names; k++) grp[k] = -1;
    int64_t *cnt = calloc(Q->n_names + 1, 8);
    for (int64_t i = 0; i < Q->n; i++) {
        if ((uint32_t)Q->qs[i] > (uint32_t)Q->qe[i]) continue;
        int k = Q->name_id[i];
        if (grp[k] < 0) { grp[k] = ng; order[ng++] = k; }
        cnt[grp[k]]++;
    }
    for (int g = 0; g < ng; g++) {
        int k = order[g];
        int64_t nq = cnt[g], a = 0;
        Nq += nq;
        qrec_t *L = malloc(sizeof(qrec_t) * (size_t)(nq + 1));
        for (int64_t i = 0; i < Q->n; i++)
            if (Q->name_id[i] == k && !((uint32_t)Q->qs[i] > (uint32_t)Q->qe[i])) {
                L[a].start = Q->qs[i], L[a].end = Q->qe[i], L[a].ord = a;
                a++;
            }
        qsort(L, nq, sizeof(qrec_t), cmp_qstart);                   /* :373 */
        int32_t c = orc_ctg_id(db, Q->names[k]);
        cand_t **ol = calloc(nq + 1, sizeof(cand_t *));
        int64_t *nh = calloc(nq + 1, 8);
        for (int64_t j = 0; j < nq; j++) {                           /* :380-389 */
            cand_ctx cc = { NULL, 0, 0, L[j].start, L[j].end };
            walk(db, j, c, L[j].start, L[j].end, 0, 0, hit_seq, &cc, NULL, NULL);
            if (cc.n) qsort(cc.a, cc.n, sizeof(cand_t), cmp_fidx);
            ol[j] = cc.a, nh[j] = cc.n;
        }
        int64_t *kst0 = calloc(nq + 1, 8), *kst = calloc(nq + 1, 8), *nst0 = calloc(nq + 1, 8);
        for (int m = 0; m < nf; m++) {
            float maxf = 0.0f;
            int64_t maxk = 0, maxj = 0;
            for (int64_t j = 0; j < nq; j++) {                       /* phase 1 :399-415 */
                int64_t kk = kst[j];
                while (kk < nh[j] && ol[j][kk].idx_f < m) kk++;
                kst0[j] = kk;
                while (kk < nh[j] && ol[j][kk].idx_f == m) {
                    if (ol[j][kk].sm > maxf) { maxf = ol[j][kk].sm; maxk = kk; maxj = j; }
                    kk++;
                }
                kst[j] = kk;
                nst0[j] = kk - kst0[j];
            }
            while (maxf > 0.0f) {                                    /* phase 2 :418-437 */
                sm[m] += maxf;
                nst0[maxj] = 0;
                int32_t it = ol[maxj][maxk].idx_t, ig = ol[maxj][maxk].idx_g;
                maxf = 0.0f;
                for (int64_t j = 0; j < nq; j++) {
                    if (nst0[j] <= 0) continue;
                    for (int64_t kk = kst0[j]; kk < kst0[j] + nst0[j]; kk++) {
                        if (ol[j][kk].idx_g == ig && ol[j][kk].idx_t == it) ol[j][kk].sm = 0.0f;
                        else if (ol[j][kk].sm > maxf) { maxf = ol[j][kk].sm; maxk = kk; maxj = j; }
                    }
                }
            }
        }
        for (int64_t j = 0; j < nq; j++) free(ol[j]);
        free(ol); free(nh); free(kst0); free(kst); free(nst0); free(L);
    }
    for (int m = 0; m < nf; m++) sm[m] = sm[m] / (Nq + db->fnr[m] - sm[m]);       /* :445-448 */
    free(grp); free(order); free(cnt);
}

/* ---- -m : hitsmap (getMap src/igd_search.c:772-826, getMap_v :829-886) ------------------------------- */
int64_t orc_hitsmap(const orc_db *db, int32_t v, uint32_t *hitmap)
{
    return orc_hitsmap_tiles(db, v, 0, INT64_MAX, hitmap);
}

/* the same loops restricted to the global tile indices [t_lo, t_hi): getMap's tiles are independent
 * (src/igd_search.c:780-782 iterates them one by one and no state survives a tile) */
int64_t orc_hitsmap_tiles(const orc_db *db, int32_t v, int64_t t_lo, int64_t t_hi, uint32_t *hitmap)
{
    int use_v = v > 0;                      /* dispatch src/igd_search.c:1000-1003 */
    int64_t nols = 0, nf = db->nfiles;
    for (int c = 0; c < db->nctg; c++) {
        for (int32_t n1 = 0; n1 < db->ntile[c]; n1++) {
            int64_t t = db->tbase[c] + n1;
            if (t < t_lo || t >= t_hi) continue;
            int32_t n = db->ncnt[t], bd = db->nbp * n1;
            if (n == 0) continue;
            const rec_t *r = db->rec + db->toff[t];
            int32_t tSb = 0;
            while (tSb < n && r[tSb].start < bd) tSb++;
            for (int32_t j = 0; j < n; j++) {
                if (use_v && !(r[j].value > v)) continue;            /* strict > in -m (:856) */
                int32_t qe = r[j].end, qs = r[j].start;
                if (!(qe > r[0].start)) continue;
                int32_t tS = qs < bd ? tSb : 0;                       /* :802-804 */
                int32_t i = last_start_lt(r, n, qe);
                /* the running-max-end early exit (:812) never changes the result: maxE[i] <= qs implies
                 * no record in [0,i] has end > qs */
                for (; i >= tS; i--)
                    if (r[i].end > qs && (!use_v || r[i].value > v)) {
                        nols++;
                        hitmap[(int64_t)r[j].idx * nf + r[i].idx]++;
                    }
            }
        }
    }
    return nols;
}
