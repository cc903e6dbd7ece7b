/* Thin C-ABI between the host C code (igd_host.c) and the CUDA layer (igd_dev_*.cu; shared definitions in igd_dev_impl.cuh).
 * Internal to libigd_b200.so; the public boundary is include/igd_b200.h. */
#ifndef IGD_DEV_H
#define IGD_DEV_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct igdb_dev igdb_dev;

int igdb_device_count(void);
const char *igdb_dev_error(void);       /* thread-local message of the last CUDA-layer failure */

/* Device image of one database.  ntile[nctg], ncnt[sum ntile] are the header tables of the .igd file
 * (reference layout: src/igd_base.c:408-419). */
igdb_dev *igdb_dev_create(int device, int32_t nbp, int32_t nctg, const int32_t *ntile,
                          const int32_t *ncnt, int32_t nfiles);
/* stream a run of on-disk records [rec0, rec0+nrec) (AoS) into the SoA image: rec_words = 4 for gdata_t
 * (16 B, gType 1), 3 for gdata0_t (12 B, gType 0; the image's value column reads 0) */
int igdb_dev_upload(igdb_dev *d, int64_t rec0, int64_t nrec, const void *aos, int rec_words);
/* the same without waiting: slot 0 or 1 names the staging buffer; the host buffer may be reused after
 * igdb_dev_upload_wait(slot).  Records whose file index is outside [0, nfiles) make igdb_dev_finalize fail. */
int igdb_dev_upload_slot(igdb_dev *d, int slot, int64_t rec0, int64_t nrec, const void *aos, int rec_words);
int igdb_dev_upload_wait(igdb_dev *d, int slot);
/* after the last upload: per-tile dedupe start (tS) and first start */
int igdb_dev_finalize(igdb_dev *d);
/* copies up to cap records of one tile (global tile index) out of the image; returns how many (< 0: error) */
int igdb_dev_tile_read(igdb_dev *d, int64_t tile, int32_t cap, int32_t *idx, int32_t *start, int32_t *end,
                       int32_t *value);
void igdb_dev_destroy(igdb_dev *d);
int64_t igdb_dev_bytes(const igdb_dev *d);
int64_t igdb_dev_launches(const igdb_dev *d);
/* plan of the last settled count batch: segments, work items, records read by the items, and the bytes
 * the search kernel had to read for them */
void igdb_dev_last_plan(const igdb_dev *d, int64_t out[4]);
/* 1 if count mode runs on the packed 16-bit layout (nbp <= 16384 and nFiles <= 65536) */
int igdb_dev_packed(const igdb_dev *d);
/* device time of the last count-mode search kernel (CUDA events on its stream); < 0 if none */
float igdb_dev_last_search_ms(igdb_dev *d);
/* diagnostics: CUDA events around the four launches of a count batch (segment count, plan, scatter, search) */
void igdb_dev_stage_timing(igdb_dev *d, int on);
int igdb_dev_last_stages(igdb_dev *d, float out[4]);
void *igdb_dev_pinned(igdb_dev *d, size_t bytes);   /* reusable pinned staging buffer (upload) */
void *igdb_host_alloc(size_t bytes);                /* zeroed page-locked host memory (small result buffers) */
void igdb_host_free(void *p);
void *igdb_dev_stream(igdb_dev *d);                 /* the handle's own cudaStream_t (what a NULL stream argument means) */
void *igdb_dev_malloc(igdb_dev *d, size_t bytes);   /* zeroed device memory on the handle's device (NULL on failure) */
void igdb_dev_free(igdb_dev *d, void *p);
int igdb_dev_memset_async(igdb_dev *d, void *p, size_t bytes, void *stream);    /* zeroes; NULL stream = the handle's */

/* count (mode_v = 0: get_overlaps control flow) / -v (mode_v = 1: get_overlaps_v control flow and the
 * value >= v predicate).  Host buffers; hits_delta[nfiles] receives the counts of THIS call. t_ms[3] = H2D,
 * kernels, D2H (may be NULL). */
int igdb_dev_count_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                        int mode_v, int32_t v, uint64_t *hits_delta, float *t_ms);
/* the same for a stream of batches, two in flight (slot 0 / 1): submit enqueues H2D (on a copy stream) + search +
 * D2H of the counts; wait settles the batch and hands out its counts.  Waits come in submission order. */
int igdb_dev_count_submit(igdb_dev *d, int slot, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                          int mode_v, int32_t v);
int igdb_dev_count_wait(igdb_dev *d, int slot, uint64_t *hits_delta);
/* device buffers, enqueued on stream; d_hits[nfiles] incremented.  Pair with igdb_dev_count_finish. */
int igdb_dev_count_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs,
                          const int32_t *d_qe, int mode_v, int32_t v, uint64_t *d_hits, void *stream);
/* the same when no query touches a tile outside [tile_lo, tile_hi) (global tile indices): plan and descriptors
 * cover that range only */
int igdb_dev_count_device_tiles(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                                int mode_v, int32_t v, uint64_t *d_hits, void *stream, int32_t tile_lo, int32_t tile_hi);
/* the same for a batch whose shape is in device memory: d_dyn = {queries, tile_lo, tile_hi} (uint64), written earlier on
 * the stream (igdb_dispatch_async); n_max bounds the number of queries */
int igdb_dev_count_device_dyn(igdb_dev *d, int64_t n_max, const uint64_t *d_dyn, const int32_t *d_ctg, const int32_t *d_qs,
                              const int32_t *d_qe, int mode_v, int32_t v, uint64_t *d_hits, void *stream);
/* waits until the batch enqueued by igdb_dev_count_device has been planned; if its plan did not fit the
 * scratch buffers (nothing was added to d_hits then) grows them and runs it again.  0 = d_hits is valid
 * in stream order. */
int igdb_dev_count_finish(igdb_dev *d);

/* Query ingest on the device.  set_contigs: the database's contig names (nctg x 40 bytes, NUL-terminated) for
 * get_id on the device.  Between hits_begin and hits_end, count_text takes chunks of BED text made of whole
 * lines: H2D, parse (parse_bed's rule), count; n_queries = accepted lines on known contigs.  hits_end waits and
 * returns the counts of all chunks in hits_delta[nfiles]. */
int igdb_dev_set_contigs(igdb_dev *d, const char *names40);
int igdb_dev_hits_begin(igdb_dev *d);
int igdb_dev_count_text(igdb_dev *d, const char *text, size_t len, int mode_v, int32_t v, int64_t *n_queries);
int igdb_dev_hits_end(igdb_dev *d, uint64_t *hits_delta);

/* -f: returns total (>= 0) and host arrays in the handle's pinned arena (valid until the next -f call on the
 * handle); query_off has n+1 entries.  t_ms[3] = H2D, kernels, D2H (may be NULL). */
int64_t igdb_dev_full_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                           int64_t **query_off, int32_t **start, int32_t **end, int32_t **file, float *t_ms);

/* -f, device to device: queries in device buffers, results left in HBM -- *d_query_off = int64[n+1],
 * *d_lists = int32 start[total] | end[total] | file[total]; both owned by the handle until the next call. */
int64_t igdb_dev_full_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             void *stream, int64_t **d_query_off, int32_t **d_lists);

/* multi-GPU -f on the gathering rank: the lists of the concatenated shards (d_cnt[i] overlaps of their i-th query,
 * which is query d_order[i] of the caller's array) rearranged into the caller's query order */
int igdb_dev_full_restore(igdb_dev *d, int64_t n, const int64_t *d_order, const int64_t *d_cnt, const int32_t *d_lists_sorted,
                          int64_t total, int64_t *d_query_off, int32_t *d_lists_out, void *stream);

/* -s: n queries grouped by query-file chrom (grp[] ascending) and start-sorted inside a group; grp_nq[g] =
 * queries of group g; acc[nfiles] = sum of matched similarities per file */
int igdb_dev_seqpare(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                     const int32_t *grp, int32_t ngroups, const int32_t *grp_nq, double *acc);

/* seq_overlaps (src/igd_search.c:253-353) for n queries: overlap_t records {idx_t, idx_g, idx_f, float sm} (16 B) in
 * the reference's scan order, query q owning [query_off[q], query_off[q+1]); arrays in the handle's pinned arena */
int64_t igdb_dev_seq_pairs_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                                int64_t **query_off, void **pairs);

/* -m restricted to the tiles [t0, t1) (global tile indices, contig-major): hitmap[nfiles*nfiles] is incremented.
 * filter != 0: only pairs whose two values both exceed v (getMap_v, src/igd_search.c:856,872, for any v). */
int64_t igdb_dev_hitsmap_host(igdb_dev *d, int filter, int32_t v, int64_t t0, int64_t t1, uint32_t *hitmap);
/* the same into a DEVICE matrix, enqueued on stream (what a rank of a multi-GPU -m runs before the all-reduce) */
int igdb_dev_hitsmap_device(igdb_dev *d, int filter, int32_t v, int64_t t0, int64_t t1, uint32_t *d_map, void *stream);

/* Sharding a query set over n_parts ranks on the device (igd_dev_shard.cu): cost adds every query's estimated scan
 * bytes to its first tile's slot of d_tile_cost[nTiles + 1]; cuts turns the (all-reduced) slots into n_parts + 1
 * tile boundaries; partition reorders the queries by destination rank (d_part_off[n_parts + 1], device) and
 * reports in d_reach[n_parts] (may be NULL) one past the last tile each destination's queries touch. */
int igdb_dev_shard_cost(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        uint64_t *d_tile_cost, void *stream);
int igdb_dev_shard_cuts(igdb_dev *d, const uint64_t *d_tile_cost, int32_t n_parts, int32_t *d_cuts, void *stream);
int igdb_dev_shard_partition(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             const int32_t *d_cuts, int32_t n_parts, int32_t *o_ctg, int32_t *o_qs, int32_t *o_qe,
                             int64_t *d_part_off, int32_t *d_reach, void *stream);

/* The partition fused with its all-to-all over peer memory (igd_dev_shard.cu): a rank's queries go straight into the
 * receive buffers (cap queries each) of the ranks that own their first tile.  create -> handle (64-byte CUDA IPC
 * handle of this rank's block) -> connect (all ranks' handles, rank-major) -> dispatch any number of times, the same
 * number on every rank, each followed on the same stream by the search of the shard it returned. */
typedef struct igdb_dispatcher igdb_dispatcher;
igdb_dispatcher *igdb_dispatcher_create(igdb_dev *d, int rank, int world, int64_t cap);
int igdb_dispatcher_handle(igdb_dispatcher *x, void *out64);
int igdb_dispatcher_connect(igdb_dispatcher *x, const void *handles);
int igdb_dispatch(igdb_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                  const int32_t *d_cuts, void *stream, int32_t **o_ctg, int32_t **o_qs, int32_t **o_qe, int64_t *m,
                  int32_t *tile_lo, int32_t *tile_hi);
/* the same without waiting: the shard's shape -- {queries, tile_lo, tile_hi}, uint64 -- stays in device memory at *d_dyn
 * for igdb_dev_count_device_dyn; the pointers into the receive buffer are fixed for the life of the dispatcher */
int igdb_dispatch_async(igdb_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        const int32_t *d_cuts, void *stream, int32_t **o_ctg, int32_t **o_qs, int32_t **o_qe,
                        const uint64_t **d_dyn);
/* waits for the device; 0 = every dispatch so far completed, < 0 = a peer did not arrive or a receive buffer was too small */
int igdb_dispatcher_status(igdb_dispatcher *x);
void igdb_dispatcher_destroy(igdb_dispatcher *x);

/* The exchange step of the multi-GPU count path (igd_dev_p2p.cu): hits[] summed over the ranks by one kernel per
 * rank over peer memory.  create -> handle (64-byte CUDA IPC handle of this rank's mailbox) -> connect (all ranks'
 * handles, rank-major) -> allreduce any number of times, the same number on every rank. */
typedef struct igdb_reducer igdb_reducer;
igdb_reducer *igdb_reducer_create(int device, int rank, int world, int32_t nfiles);
int igdb_reducer_handle(igdb_reducer *r, void *out64);
int igdb_reducer_connect(igdb_reducer *r, const void *handles);
int igdb_reducer_allreduce(igdb_reducer *r, unsigned long long *d_hits, void *stream);
/* out of place: d_out = the sum over the ranks of their d_in, and d_in is left zeroed */
int igdb_reducer_allreduce_into(igdb_reducer *r, unsigned long long *d_in, unsigned long long *d_out, void *stream);
int igdb_reducer_status(igdb_reducer *r);       /* 0 ok; k > 0: a wait for rank k - 1 timed out; < 0: CUDA error */
void igdb_reducer_destroy(igdb_reducer *r);

#ifdef __cplusplus
}
#endif
#endif
