/* TEST INFRASTRUCTURE ONLY.  A device layer that has no device: every igdb_dev_* entry point of
 * igd_b200/csrc/igd_dev.h fails.  Linked with igd_host.c under -fsanitize=address,undefined so that the host C
 * (query ingest, .igd / _index.tsv header parsing, error paths) can run on a box without a GPU or the CUDA runtime.
 * Never part of the product: libigd_b200.so links the real CUDA layer. */
#include "igd_dev.h"
#include <stdlib.h>

int igdb_device_count(void) { return 0; }
const char *igdb_dev_error(void) { return "stub device layer"; }
igdb_dev *igdb_dev_create(int a, int32_t b, int32_t c, const int32_t *d, const int32_t *e, int32_t f) { (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; return NULL; }
int igdb_dev_upload(igdb_dev *d, int64_t a, int64_t b, const void *c, int w) { (void)d; (void)a; (void)b; (void)c; (void)w; return -1; }
int igdb_dev_upload_slot(igdb_dev *d, int s, int64_t a, int64_t b, const void *c, int w) { (void)d; (void)s; (void)a; (void)b; (void)c; (void)w; return -1; }
int igdb_dev_upload_wait(igdb_dev *d, int s) { (void)d; (void)s; return -1; }
int igdb_dev_finalize(igdb_dev *d) { (void)d; return -1; }
int igdb_dev_tile_read(igdb_dev *d, int64_t t, int32_t c, int32_t *a, int32_t *b, int32_t *e, int32_t *v) { (void)d; (void)t; (void)c; (void)a; (void)b; (void)e; (void)v; return -1; }
void igdb_dev_destroy(igdb_dev *d) { (void)d; }
int64_t igdb_dev_bytes(const igdb_dev *d) { (void)d; return 0; }
int64_t igdb_dev_launches(const igdb_dev *d) { (void)d; return 0; }
void igdb_dev_last_plan(const igdb_dev *d, int64_t out[4]) { (void)d; out[0] = out[1] = out[2] = out[3] = 0; }
int igdb_dev_packed(const igdb_dev *d) { (void)d; return 0; }
float igdb_dev_last_search_ms(igdb_dev *d) { (void)d; return -1.f; }
void igdb_dev_stage_timing(igdb_dev *d, int on) { (void)d; (void)on; }
int igdb_dev_last_stages(igdb_dev *d, float out[4]) { (void)d; (void)out; return -1; }
void *igdb_dev_pinned(igdb_dev *d, size_t bytes) { (void)d; (void)bytes; return NULL; }
void *igdb_host_alloc(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
void igdb_host_free(void *p) { free(p); }
void *igdb_dev_stream(igdb_dev *d) { (void)d; return NULL; }
void *igdb_dev_malloc(igdb_dev *d, size_t b) { (void)d; (void)b; return NULL; }
void igdb_dev_free(igdb_dev *d, void *p) { (void)d; (void)p; }
int igdb_dev_memset_async(igdb_dev *d, void *p, size_t b, void *s) { (void)d; (void)p; (void)b; (void)s; return -1; }
int igdb_dev_count_host(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int m, int32_t v, uint64_t *h, float *t) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)m; (void)v; (void)h; (void)t; return -1; }
int igdb_dev_count_device(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int m, int32_t v, uint64_t *h, void *s) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)m; (void)v; (void)h; (void)s; return -1; }
int igdb_dev_count_finish(igdb_dev *d) { (void)d; return -1; }
int igdb_dev_count_device_dyn(igdb_dev *d, int64_t n, const uint64_t *y, const int32_t *a, const int32_t *b, const int32_t *c, int m, int32_t v, uint64_t *h, void *s) { (void)d; (void)n; (void)y; (void)a; (void)b; (void)c; (void)m; (void)v; (void)h; (void)s; return -1; }
int igdb_dev_count_submit(igdb_dev *d, int s, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int m, int32_t v) { (void)d; (void)s; (void)n; (void)a; (void)b; (void)c; (void)m; (void)v; return -1; }
int igdb_dev_count_wait(igdb_dev *d, int s, uint64_t *h) { (void)d; (void)s; (void)h; return -1; }
int igdb_dev_set_contigs(igdb_dev *d, const char *n) { (void)d; (void)n; return -1; }
int igdb_dev_hits_begin(igdb_dev *d) { (void)d; return -1; }
int igdb_dev_count_text(igdb_dev *d, const char *t, size_t l, int m, int32_t v, int64_t *n) { (void)d; (void)t; (void)l; (void)m; (void)v; (void)n; return -1; }
int igdb_dev_hits_end(igdb_dev *d, uint64_t *h) { (void)d; (void)h; return -1; }
int64_t igdb_dev_full_host(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int64_t **q, int32_t **s, int32_t **e, int32_t **f, float *t) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)q; (void)s; (void)e; (void)f; (void)t; return -1; }
int64_t igdb_dev_full_device(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, void *st, int64_t **q, int32_t **l) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)st; (void)q; (void)l; return -1; }
int64_t igdb_dev_seq_pairs_host(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int64_t **q, void **p) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)q; (void)p; return -1; }
int igdb_dev_seqpare(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, const int32_t *g, int32_t ng, const int32_t *gn, double *acc) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)g; (void)ng; (void)gn; (void)acc; return -1; }
int64_t igdb_dev_hitsmap_host(igdb_dev *d, int f, int32_t v, int64_t t0, int64_t t1, uint32_t *h) { (void)d; (void)f; (void)v; (void)t0; (void)t1; (void)h; return -1; }
int igdb_dev_hitsmap_device(igdb_dev *d, int f, int32_t v, int64_t t0, int64_t t1, uint32_t *h, void *s) { (void)d; (void)f; (void)v; (void)t0; (void)t1; (void)h; (void)s; return -1; }
int igdb_dev_count_device_tiles(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, int m, int32_t v, uint64_t *h, void *s, int32_t lo, int32_t hi) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)m; (void)v; (void)h; (void)s; (void)lo; (void)hi; return -1; }
int igdb_dev_shard_cost(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, uint64_t *t, void *s) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)t; (void)s; return -1; }
int igdb_dev_shard_cuts(igdb_dev *d, const uint64_t *t, int32_t n, int32_t *c, void *s) { (void)d; (void)t; (void)n; (void)c; (void)s; return -1; }
int igdb_dev_shard_partition(igdb_dev *d, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, const int32_t *cu, int32_t np, int32_t *oa, int32_t *ob, int32_t *oc, int64_t *off, int32_t *r, void *s) { (void)d; (void)n; (void)a; (void)b; (void)c; (void)cu; (void)np; (void)oa; (void)ob; (void)oc; (void)off; (void)r; (void)s; return -1; }
int igdb_dev_full_restore(igdb_dev *d, int64_t n, const int64_t *o, const int64_t *c, const int32_t *l, int64_t t, int64_t *q, int32_t *lo, void *s) { (void)d; (void)n; (void)o; (void)c; (void)l; (void)t; (void)q; (void)lo; (void)s; return -1; }
igdb_dispatcher *igdb_dispatcher_create(igdb_dev *d, int r, int w, int64_t c) { (void)d; (void)r; (void)w; (void)c; return NULL; }
int igdb_dispatcher_handle(igdb_dispatcher *x, void *o) { (void)x; (void)o; return -1; }
int igdb_dispatcher_connect(igdb_dispatcher *x, const void *h) { (void)x; (void)h; return -1; }
int igdb_dispatch(igdb_dispatcher *x, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, const int32_t *k, void *s, int32_t **oa, int32_t **ob, int32_t **oc, int64_t *m, int32_t *lo, int32_t *hi) { (void)x; (void)n; (void)a; (void)b; (void)c; (void)k; (void)s; (void)oa; (void)ob; (void)oc; (void)m; (void)lo; (void)hi; return -1; }
int igdb_dispatch_async(igdb_dispatcher *x, int64_t n, const int32_t *a, const int32_t *b, const int32_t *c, const int32_t *k, void *s, int32_t **oa, int32_t **ob, int32_t **oc, const uint64_t **dy) { (void)x; (void)n; (void)a; (void)b; (void)c; (void)k; (void)s; (void)oa; (void)ob; (void)oc; (void)dy; return -1; }
int igdb_dispatcher_status(igdb_dispatcher *x) { (void)x; return -1; }
void igdb_dispatcher_destroy(igdb_dispatcher *x) { (void)x; }
igdb_reducer *igdb_reducer_create(int d, int r, int w, int32_t n) { (void)d; (void)r; (void)w; (void)n; return NULL; }
int igdb_reducer_handle(igdb_reducer *r, void *o) { (void)r; (void)o; return -1; }
int igdb_reducer_connect(igdb_reducer *r, const void *h) { (void)r; (void)h; return -1; }
int igdb_reducer_allreduce(igdb_reducer *r, unsigned long long *h, void *s) { (void)r; (void)h; (void)s; return -1; }
int igdb_reducer_allreduce_into(igdb_reducer *r, unsigned long long *i, unsigned long long *o, void *s) { (void)r; (void)i; (void)o; (void)s; return -1; }
void igdb_reducer_destroy(igdb_reducer *r) { (void)r; }
int igdb_reducer_status(igdb_reducer *r) { (void)r; return -1; }
