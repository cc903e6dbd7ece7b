/* igd_b200 -- C-ABI of the B200-native IGD batch overlap search.
 *
 * Plain pointers and sizes only; no CUDA or torch types.  Two groups of entry points:
 *
 *  (1) a handle-based, batched interface (what a new binding should call), and
 *  (2) the reference's own global-state entry points of src/igd_search.h, with the same names,
 *      argument meaning, return values and (absence of) error reporting, implemented over (1).
 *
 * Every declaration cites the reference interface it replaces (paths under databio/IGD).
 * All searches run on the GPU; there is no CPU fallback -- a call that cannot reach a CUDA device
 * fails (NULL / negative return, message in igd_b200_last_error()).
 */
#ifndef IGD_B200_H
#define IGD_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * (1) handle-based interface
 * ---------------------------------------------------------------------------------------------- */
typedef struct igd_b200 igd_b200_t;          /* replaces iGD_t*, IGDr/src/igd_base.h:71-83 */

/* open_iGD (IGDr/src/igd_base.h:97) + get_igdinfo/get_fileinfo (src/igd_base.c:269-323, 235-267):
 * parses <db>.igd and <db>_index.tsv and uploads every tile into the HBM of CUDA device `device`
 * (CSR of tile offsets + SoA start/end/idx/value).  Both record types are read: gType 1 (gdata_t, 16 B) and
 * gType 0 (gdata0_t, 12 B: no value -- -v is ignored in count mode exactly as igd_search does,
 * src/igd_search.c:1024,1042, and the value column reads 0 everywhere else).  NULL on failure. */
igd_b200_t *igd_b200_open(const char *igd_path, int device);
/* close_iGD (IGDr/src/igd_base.h:124) */
void igd_b200_close(igd_b200_t *h);
/* message of the last failure on this thread ("" if none) */
const char *igd_b200_last_error(void);
/* number of visible CUDA devices (<= 0: none) */
int igd_b200_device_count(void);

/* get_nFiles / get_nCtgs / get_nbp (IGDr/src/igd_search.c:357-393) and iGD_t fields */
int32_t igd_b200_nfiles(const igd_b200_t *h);
int32_t igd_b200_nctg(const igd_b200_t *h);
int32_t igd_b200_nbp(const igd_b200_t *h);
int32_t igd_b200_gtype(const igd_b200_t *h);
int64_t igd_b200_ntiles(const igd_b200_t *h);
int64_t igd_b200_nrecords(const igd_b200_t *h);
int32_t igd_b200_ntile(const igd_b200_t *h, int32_t ctg);           /* iGD_t.nTile[ctg] */
int32_t igd_b200_ncnt(const igd_b200_t *h, int32_t ctg, int32_t tile); /* iGD_t.nCnt[ctg][tile] */
int32_t igd_b200_file_nr(const igd_b200_t *h, int32_t i);           /* info_t.nr   src/igd_base.h:54-58 */
const char *igd_b200_file_name(const igd_b200_t *h, int32_t i);     /* info_t.fileName */
const char *igd_b200_ctg_name(const igd_b200_t *h, int32_t ctg);    /* iGD_t.cName[ctg] */
/* get_id (src/igd_base.c:325-331): contig index, -1 if absent */
int32_t igd_b200_ctg_id(const igd_b200_t *h, const char *chrm);
/* get_binData (IGDr/src/igd_search.c:357-393): the records of tile `bin` of contig `ctg`, in file order, read back
 * from the HBM image.  Copies up to cap entries into each non-NULL array; returns nCnt of the tile (0 if the
 * tile does not exist), negative on a CUDA error. */
int32_t igd_b200_bin_data(igd_b200_t *h, int32_t ctg, int32_t bin, int32_t cap, int32_t *idx, int32_t *start,
                          int32_t *end, int32_t *value);
/* bytes of HBM held by the database image */
int64_t igd_b200_device_bytes(const igd_b200_t *h);

/* Batched count search -- the n-query array form of search_nr (IGDr/src/igd_search.c:340-355) with the
 * semantics of get_overlaps (v <= 0; src/igd_search.c:454-534) or get_overlaps_v (v > 0; :623-694) per
 * query.  hits[nFiles] is caller-allocated, caller-zeroed and INCREMENTED (src/igd_search.c:925,491).
 * Returns the number of overlaps found in this call (>= 0) or a negative error code.
 * (The reference's own return value -- 0 for plain, the count for -v -- is what group (2) returns.) */
int64_t get_overlaps_n(igd_b200_t *h, int64_t n, const char *const *chrm, const int32_t *qs,
                       const int32_t *qe, int32_t v, int64_t *hits);
/* search_nr as the R package calls it (IGDr/src/igd_search.c:340-355): n x get_overlaps32 (:105-186), 32-bit
 * counters, incremented (they wrap like the reference's int).  Returns the overlaps found or a negative error code. */
int64_t get_overlaps_n32(igd_b200_t *h, int64_t n, const char *const *chrm, const int32_t *qs,
                         const int32_t *qe, int32_t *hits);
/* same, contig ids already resolved with igd_b200_ctg_id (-1 = unknown chrom: contributes nothing) */
int64_t igd_b200_count(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs,
                       const int32_t *qe, int32_t v, int64_t *hits);
/* The same for a STREAM of batches, two in flight (no counterpart in the reference, whose search_nr is one
 * synchronous loop, IGDr/src/igd_search.c:340-355).  End to end a batch is PCIe first and HBM second -- 12 B per query
 * cross the bus before a search that needs the whole batch for its plan -- so one call cannot overlap the two, but a
 * caller with several batches can:
 *     igd_b200_count_submit(h, 0, n0, ...);
 *     igd_b200_count_submit(h, 1, n1, ...);  igd_b200_count_wait(h, 0, hits);     -- batch 1 copies while batch 0 searches
 *     igd_b200_count_submit(h, 0, n2, ...);  igd_b200_count_wait(h, 1, hits);  ...
 * submit: slot is 0 or 1 and must be free; n <= 2^26; the host arrays (pinned memory for full PCIe speed) must stay
 * untouched until the wait for that slot returns.  Returns 0 once copy and search are enqueued.
 * wait: batches are waited for in submission order; hits[nFiles] is INCREMENTED like igd_b200_count's; returns the
 * overlaps of the batch or a negative error code.  The one-call entry points fail while a slot is busy. */
int igd_b200_count_submit(igd_b200_t *h, int slot, int64_t n, const int32_t *ctg, const int32_t *qs,
                          const int32_t *qe, int32_t v);
int64_t igd_b200_count_wait(igd_b200_t *h, int slot, int64_t *hits);
/* same, but every buffer is a DEVICE pointer on the handle's device and the work is enqueued on
 * `stream` (a cudaStream_t, may be NULL) without any host round trip.  d_hits is uint64[nFiles],
 * incremented.  Returns 0 once the work is enqueued.  This is the entry point the multi-GPU driver uses,
 * with an NCCL allreduce of d_hits behind it.
 * igd_b200_count_device_finish must be called before d_hits is consumed (read back, or reduced across
 * ranks): the search is enqueued speculatively against the handle's scratch buffers, and a batch whose
 * (query, tile) expansion did not fit them adds NOTHING to d_hits; finish detects that, grows the buffers
 * and runs the batch again on the same stream (only ever on the first batch of a larger size).  It waits
 * for the batch's plan, not for the search kernel. */
int igd_b200_count_device(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs,
                          const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream);
int igd_b200_count_device_finish(igd_b200_t *h);
/* The same when the caller knows that no query touches a tile outside [tile_lo, tile_hi) (global tile indices,
 * contig-major: tile j of contig c is (sum of nTile of the contigs before c) + j) -- a shard of a multi-GPU search,
 * below: the batch's plan then covers that range only. */
int igd_b200_count_device_tiles(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs,
                                const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream, int32_t tile_lo,
                                int32_t tile_hi);

/* igd_b200_count_device_tiles for a batch whose shape is only known on the device: d_shape = {queries, tile_lo, tile_hi}
 * (uint64[3], device), written by earlier work on `stream` (igd_b200_dispatch_async); n_max >= queries sizes the scratch
 * buffers and the grids (<= 2^26, and <= the capacity of the arrays).  Pair with igd_b200_count_device_finish. */
int igd_b200_count_device_shaped(igd_b200_t *h, int64_t n_max, const uint64_t *d_shape, const int32_t *d_ctg,
                                 const int32_t *d_qs, const int32_t *d_qe, int32_t v, uint64_t *d_hits, void *stream);

/* Sharding a query set over the ranks of a multi-GPU search, on the device (the reference is one process and has
 * no counterpart; it implements north_star's "the sorted query set is sharded into contiguous ranges balanced by a
 * prefix sum of estimated per-query scan bytes").  Rank r takes the queries whose FIRST bin is a tile of
 * [cuts[r], cuts[r+1]): contiguous ranges of the (contig, start) order at tile granularity.  All pointers are
 * device pointers; the calls are enqueued on `stream`.
 *   shard_cost       d_tile_cost[nTiles + 1] += the estimated scan bytes of every query (the figure of
 *                    igd_b200_query_cost), at its first tile's slot; slot nTiles takes the queries that reach no tile
 *   (sum the slots over the ranks, e.g. one NCCL all-reduce)
 *   shard_cuts       d_cuts[n_parts + 1]: cuts[0] = 0, cuts[n_parts] = nTiles, cuts[p] = the first tile at which the
 *                    running sum reaches p / n_parts of the total
 *   shard_partition  the n queries reordered by destination rank into d_out_*; d_part_off[n_parts + 1] (int64) = where
 *                    each destination's run starts; d_reach[n_parts] (int32, may be NULL) = one past the last tile the
 *                    queries of a destination touch (the tile_hi of igd_b200_count_device_tiles)
 * n_parts <= 64.  0 on success. */
int igd_b200_shard_cost(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        uint64_t *d_tile_cost, void *stream);
int igd_b200_shard_cuts(igd_b200_t *h, const uint64_t *d_tile_cost, int32_t n_parts, int32_t *d_cuts, void *stream);
int igd_b200_shard_partition(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             const int32_t *d_cuts, int32_t n_parts, int32_t *d_out_ctg, int32_t *d_out_qs,
                             int32_t *d_out_qe, int64_t *d_part_off, int32_t *d_reach, void *stream);

/* shard_partition FUSED with the all-to-all that follows it, over NVLink peer memory: the scatter pass stores every
 * query straight into the receive buffer (cap queries) of the rank that will search it, so the query exchange of a
 * multi-GPU search is three launches per rank and no library collective (igd_dev_shard.cu: per-destination counts
 * pushed to every rank -> every rank knows the whole source x destination matrix, hence every slot range -> stores
 * into the peers' buffers -> flags).  One process per GPU:
 *   x = igd_b200_dispatcher_create(h, rank, world, cap);       -- cap >= the largest shard (the whole query set is safe)
 *   igd_b200_dispatcher_handle(x, h64);  ... exchange the 64-byte CUDA IPC handles ...  igd_b200_dispatcher_connect(x, all);
 *   per batch: shard_cost, (sum over the ranks), shard_cuts, then
 *   igd_b200_dispatch(x, n, d_ctg, d_qs, d_qe, d_cuts, stream, &ctg, &qs, &qe, &m, &lo, &hi);
 *   igd_b200_count_device_tiles(h, m, ctg, qs, qe, v, d_hits, stream, lo, hi);   -- SAME stream, before the next dispatch
 * dispatch returns when this rank's shard (m queries, device pointers into its receive buffer) is complete; every rank
 * must call it the same number of times.  0 on success; a peer that does not arrive within 20 s is an error. */
typedef struct igdb_dispatcher igd_b200_dispatcher;
igd_b200_dispatcher *igd_b200_dispatcher_create(igd_b200_t *h, int rank, int world, int64_t cap);
int igd_b200_dispatcher_handle(igd_b200_dispatcher *x, void *out64);
int igd_b200_dispatcher_connect(igd_b200_dispatcher *x, const void *handles);
int igd_b200_dispatch(igd_b200_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                      const int32_t *d_cuts, void *stream, int32_t **d_out_ctg, int32_t **d_out_qs, int32_t **d_out_qe,
                      int64_t *m, int32_t *tile_lo, int32_t *tile_hi);
/* The same without the host in the loop: nothing is read back, the shard's shape -- {queries, tile_lo, tile_hi}, uint64
 * -- stays in device memory at *d_shape for igd_b200_count_device_shaped (same stream).  The output pointers are the
 * same for the life of the dispatcher.  igd_b200_dispatcher_status (waits for the device) tells afterwards whether every
 * dispatch completed: 0, or negative when a peer did not arrive or a receive buffer was too small. */
int igd_b200_dispatch_async(igd_b200_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                            const int32_t *d_cuts, void *stream, int32_t **d_out_ctg, int32_t **d_out_qs,
                            int32_t **d_out_qe, const uint64_t **d_shape);
int igd_b200_dispatcher_status(igd_b200_dispatcher *x);
void igd_b200_dispatcher_destroy(igd_b200_dispatcher *x);

/* One multi-GPU count step as ONE call: from "this rank's part of the query set is in its HBM, any order" to "hits[] of
 * the whole set is in its HBM", queued on one stream with no host round trip and no library collective -- shard_cost,
 * its sum over the ranks (the reducer's kernel over nTiles + 1 words), shard_cuts, igd_b200_dispatch_async,
 * igd_b200_count_device_shaped, igd_b200_reducer_allreduce_into.  One process per GPU:
 *   s = igd_b200_sharded_create(h, device, rank, world, cap);   -- cap >= queries of the whole set (receive buffers)
 *   igd_b200_sharded_handles(s, h192);  ... exchange the 192 bytes between the ranks ...  igd_b200_sharded_connect(s, all);
 *   igd_b200_sharded_count(s, n, d_ctg, d_qs, d_qe, v, d_hits, stream);    -- d_hits[nFiles] (uint64) is OVERWRITTEN
 * Every rank must call sharded_count the same number of times.  igd_b200_sharded_status waits for the device and
 * returns 0 when every step met all its peers and fitted its buffers. */
typedef struct igd_b200_sharded igd_b200_sharded;
igd_b200_sharded *igd_b200_sharded_create(igd_b200_t *h, int device, int rank, int world, int64_t cap);
int igd_b200_sharded_handles(igd_b200_sharded *s, void *out192);
int igd_b200_sharded_connect(igd_b200_sharded *s, const void *handles);
int igd_b200_sharded_count(igd_b200_sharded *s, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                           int32_t v, uint64_t *d_hits, void *stream);
int igd_b200_sharded_status(igd_b200_sharded *s);
void igd_b200_sharded_destroy(igd_b200_sharded *s);

/* -f: the overlaps themselves, in the reference's order (get_overlaps_f1, src/igd_search.c:537-620:
 * per query, bin n1 from the last start < qe downwards, then later bins ascending).
 * The arrays live in a pinned host arena owned by the handle (so that the lists come back at PCIe speed): they
 * stay valid until the next igd_b200_full / -f call on the same handle or igd_b200_close, and must not be
 * freed by the caller; igd_b200_hitlist_free only clears the struct.  igd_b200_last_timing reports the H2D /
 * kernel / D2H split of the call. */
typedef struct {
    int64_t total;          /* number of overlaps */
    int64_t n_queries;
    int64_t *query_off;     /* [n_queries + 1] : overlaps of query q are [query_off[q], query_off[q+1]) */
    int32_t *start;         /* [total] gdata_t.start */
    int32_t *end;           /* [total] gdata_t.end */
    int32_t *file;          /* [total] gdata_t.idx */
} igd_b200_hitlist;
int igd_b200_full(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                  igd_b200_hitlist *out);
void igd_b200_hitlist_free(igd_b200_hitlist *l);
/* the same with every buffer on the handle's device: the queries are read from device memory and the result stays
 * there -- *d_query_off = int64[n + 1], *d_lists = int32 start[total] | end[total] | file[total] (one allocation),
 * both owned by the handle until the next igd_b200_full_device call or igd_b200_close.  The work is enqueued on
 * `stream` (a cudaStream_t; NULL = the handle's own) and has completed when the call returns.  Returns total or a
 * negative error code.  This is what a multi-GPU -f hands to its gather (NCCL reads the lists where they are). */
int64_t igd_b200_full_device(igd_b200_t *h, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             void *stream, int64_t **d_query_off, int32_t **d_lists);

/* The last step of a multi-GPU -f, on the rank that gathers (device pointers, enqueued on `stream`, returns when
 * done): d_lists_sorted = start[total] | end[total] | file[total] of the queries in the order the shards were
 * gathered, d_cnt[i] = overlaps of the i-th of them, d_order[i] = its index in the caller's query array.  Writes the
 * lists in the CALLER's query order -- each query's run unchanged, i.e. in the reference's order -- into d_lists_out
 * (same layout) and d_query_off[n + 1]: what a single-GPU igd_b200_full_device returns for the unsharded input. */
int igd_b200_full_restore(igd_b200_t *h, int64_t n, const int64_t *d_order, const int64_t *d_cnt,
                          const int32_t *d_lists_sorted, int64_t total, int64_t *d_query_off, int32_t *d_lists_out,
                          void *stream);

/* -m: getMap / getMap_v (src/igd_search.c:772-886) as one dense row-major uint32[nFiles*nFiles],
 * incremented (wraps mod 2^32 like the reference).  v > 0 selects the strict `value > v` filter.
 * Returns the number of pairs counted or a negative error code. */
int64_t igd_b200_hitsmap(igd_b200_t *h, int32_t v, uint32_t *hitmap);
/* the same over the tiles [tile_lo, tile_hi) only (global tile index = tiles of the contigs before + bin;
 * getMap's outer loops, src/igd_search.c:780-782, cut into a range).  Tiles are independent, so the maps of
 * disjoint ranges add up to the whole map: this is what one rank of a multi-GPU -m computes before the
 * uint32 allreduce. */
int64_t igd_b200_hitsmap_tiles(igd_b200_t *h, int32_t v, int64_t tile_lo, int64_t tile_hi, uint32_t *hitmap);
/* the same into a matrix in DEVICE memory (uint32[nFiles*nFiles], incremented), enqueued on `stream`: the matrix of a
 * rank stays in HBM through the all-reduce of a multi-GPU -m and comes to the host once.  0 or a negative error code. */
int igd_b200_hitsmap_tiles_device(igd_b200_t *h, int32_t v, int64_t tile_lo, int64_t tile_hi, uint32_t *d_hitmap,
                                  void *stream);

/* Query ingest: the accepted lines of a BED(.gz) file by parse_bed's rule (src/igd_base.c:53-72: >= 3
 * tab-separated fields, contig starts with "chr", shorter than 40 chars, end > 0), read like
 * getOverlaps does (src/igd_search.c:696-719). */
typedef struct {
    int64_t n;
    int32_t *name_id;       /* [n] index into names[] (first-appearance order) */
    int32_t *qs, *qe;       /* [n] */
    int32_t n_names;
    char **names;
} igd_b200_queries;
igd_b200_queries *igd_b200_read_bed(const char *path);      /* NULL if the file cannot be opened */
void igd_b200_queries_free(igd_b200_queries *q);
/* contig id per query (-1 = not in the database) */
void igd_b200_resolve(const igd_b200_t *h, const igd_b200_queries *q, int32_t *ctg_out);

/* Whole-file search -- search_n of src_py/igd_py.pyx:42-44 / getOverlaps(iGD, qFile, hits)
 * (src_py/igd_search.c:104-128): parse qFile, count (v <= 0) or count with the signal filter (v > 0).
 * Returns the number of overlaps, or negative on error. */
int64_t search(igd_b200_t *h, const char *qFile, int32_t v, int64_t *hits);
/* -s: seqOverlaps (src/igd_search.c:355-451); sm[nFiles] is overwritten.  0 on success. */
int igd_b200_seqpare(igd_b200_t *h, const char *qFile, double *sm);

/* seq_overlaps (src/igd_search.c:253-353), batched: for every query the records it overlaps as the reference's
 * overlap_t (src/igd_base.h:109-114) -- idx_t = the QUERY's first bin, idx_g = index inside the tile the record
 * was found in, idx_f = file, sm = fp32 pair similarity st / (qlen + rlen - st) -- in the reference's scan order.
 * query_off[n+1] and list[] live in the handle's pinned arena (valid until the next list-returning call on the
 * handle).  Returns the number of records or a negative error code. */
typedef struct { int32_t idx_t, idx_g, idx_f; float sm; } igd_b200_overlap;             /* overlap_t */
typedef struct { int32_t nn, mm; igd_b200_overlap *olist; } igd_b200_overlaps;          /* overlaps_t, igd_base.h:116-119 */
int64_t igd_b200_seq_overlaps(igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                              int64_t **query_off, igd_b200_overlap **list);

/* Multi-GPU planning: cost[i] = estimated scan bytes of query i (sum over its bins of
 * 4*nCnt); used to cut a (ctg,start)-sorted query set into contiguous, cost-balanced shards. */
int igd_b200_query_cost(const igd_b200_t *h, int64_t n, const int32_t *ctg, const int32_t *qs,
                        const int32_t *qe, int64_t *cost);

/* The one exchange of the multi-GPU count path -- hits[] summed over the ranks ("one allreduce of
 * hits[nFiles]") -- as ONE small kernel per rank over NVLink peer memory instead of a library collective: every
 * rank stores its counts into a mailbox in every peer's HBM, raises a flag there, waits for the flags in its own
 * mailbox and adds the slots up in rank order (igd_dev_p2p.cu).  One process per GPU:
 *   r = igd_b200_reducer_create(device, rank, world, nFiles);
 *   igd_b200_reducer_handle(r, h64);            -- 64-byte CUDA IPC handle of this rank's mailbox
 *   ... exchange the handles between the processes (any transport) ...
 *   igd_b200_reducer_connect(r, all_handles);   -- world x 64 bytes, rank-major
 *   igd_b200_count_device(..., d_hits, stream); igd_b200_count_device_finish(h);
 *   igd_b200_reducer_allreduce(r, d_hits, stream);    -- d_hits (uint64[nFiles], device) becomes the sum; enqueued
 * Every rank must call allreduce the same number of times.  0 on success, negative on error. */
typedef struct igdb_reducer igd_b200_reducer;
igd_b200_reducer *igd_b200_reducer_create(int device, int rank, int world, int32_t nfiles);
int igd_b200_reducer_handle(igd_b200_reducer *r, void *out64);
int igd_b200_reducer_connect(igd_b200_reducer *r, const void *handles);
int igd_b200_reducer_allreduce(igd_b200_reducer *r, uint64_t *d_hits, void *stream);
/* out of place: d_out = the sum over the ranks of their d_in, and d_in is left ZEROED -- ready to take the next batch's
 * counts, so a search step needs no separate launch to clear its counters */
int igd_b200_reducer_allreduce_into(igd_b200_reducer *r, uint64_t *d_in, uint64_t *d_out, void *stream);
/* waits for the device; 0 = every exchange so far met all its peers, k > 0 = the wait for rank k - 1 gave up after
 * 20 s (a rank died or skipped a call: the sums are then invalid), negative = CUDA error */
int igd_b200_reducer_status(igd_b200_reducer *r);
void igd_b200_reducer_destroy(igd_b200_reducer *r);

/* timing of the last igd_b200_count / get_overlaps_n call on this handle, milliseconds (CUDA events):
 * t[0] H2D, t[1] kernels, t[2] D2H */
void igd_b200_last_timing(const igd_b200_t *h, float t[3]);
/* device time (ms, CUDA events on the launching stream) of the search kernel of the last count call;
 * blocks until that kernel has finished; < 0 if there was none */
float igd_b200_last_search_kernel_ms(igd_b200_t *h);
/* diagnostics: with stage timing on, every count batch records CUDA events around its four launches;
 * t[0..3] = segment count, plan, segment scatter, search kernel of the last batch, ms (< 0 return: none) */
void igd_b200_stage_timing(igd_b200_t *h, int on);
int igd_b200_last_stage_ms(igd_b200_t *h, float t[4]);
/* plan of the last settled count batch: out[0] (query, tile) segments, out[1] work items, out[2] records the
 * items read (each touched tile once per 64 of its segments), out[3] the bytes the search kernel had to read
 * for them: 16*items + 6*records (8 with -v) + 4*segments on the packed 16-bit layout,
 * 16*items + 12*records (16 with -v) + 8*segments on the int32 layout */
void igd_b200_last_plan(const igd_b200_t *h, int64_t out[4]);
/* kernel launches issued by this handle since open (for bench.py's gpu_launches) */
int64_t igd_b200_launch_count(const igd_b200_t *h);

/* ------------------------------------------------------------------------------------------------
 * (2) the reference's global-state entry points (src/igd_search.h), operating on the handle made
 *     current with igd_b200_use() -- the counterpart of the reference's process globals
 *     IGD / fP / gData (src/igd.c:14-19) that igd_search() sets up (src/igd_search.c:917-925,974).
 * ---------------------------------------------------------------------------------------------- */
void igd_b200_use(igd_b200_t *h);
igd_b200_t *igd_b200_current(void);

int32_t get_overlaps(char *chrm, int32_t qs, int32_t qe, int64_t *hits);              /* igd_search.h:23 */
int32_t get_overlaps_v(char *chrm, int32_t qs, int32_t qe, int32_t v, int64_t *hits); /* igd_search.h:24 */
int64_t getOverlaps(char *qFile, int64_t *hits);                                      /* igd_search.h:27 */
int64_t getOverlaps_v(char *qFile, int64_t *hits, int32_t v);                         /* igd_search.h:28 */
int64_t getMap(uint32_t **hitmap);                                                    /* igd_search.h:31 */
int64_t getMap_v(uint32_t **hitmap, int32_t v);                                       /* igd_search.h:32 */
int32_t get_overlaps_f1(char *chrm, int32_t qs, int32_t qe);                          /* igd_search.h:36 */
int64_t getOverlaps_f1(char *qFile);                                                  /* igd_search.h:38 */
void seqOverlaps(char *qFile, double *sm);                                            /* igd_search.h:20 */
/* appends to olp->olist (malloc'ed by the caller or NULL with mm == 0; grown by doubling like EXPAND) */
void seq_overlaps(char *chrm, int32_t qs, int32_t qe, igd_b200_overlaps *olp);        /* igd_search.h:19 */
/* the gType-0 (gdata0_t, `igd create -s 0`) twins: same searches, the handle knows its record type */
int32_t get_overlaps0(char *chrm, int32_t qs, int32_t qe, int64_t *hits);             /* igd_search.h:15 */
int64_t getOverlaps0(char *qFile, int64_t *hits);                                     /* igd_search.h:16 */
int32_t get_overlaps_f0(char *chrm, int32_t qs, int32_t qe);                          /* igd_search.h:35 */
int64_t getOverlaps_f0(char *qFile);                                                  /* igd_search.h:37 */
int igd_search(int argc, char **argv);                                                /* igd_search.h:41 */

/* ------------------------------------------------------------------------------------------------
 * (3) index build -- src/igd_create.h:11-17.  Host-side C (no GPU needed): reads the same inputs and writes the
 *     same <out>/<name>.igd and <name>_index.tsv as the reference, records in the same order.
 * ---------------------------------------------------------------------------------------------- */
extern int32_t tile_size;                                           /* src/igd.c:18; bin width of the next build */
void create_igd(char *iPath, char *oPath, char *igdName);           /* igd_create.h:11: glob of BED(.gz) files */
void create_igd0(char *iPath, char *oPath, char *igdName);          /* igd_create.h:12: -s 0, gdata0_t records */
void create_igd_f(char *iPath, char *oPath, char *igdName);         /* igd_create.h:13: -f, list of files */
void create_igd_bed4(char *iPath, char *oPath, char *igdName);      /* igd_create.h:14: -s 2, one file, column 4 = dataset */
int igd_create(int argc, char **argv);                              /* igd_create.h:17 */
/* create_iGD (src_py/igd_create.c:19-49, igd_py.create of src_py/igd_py.pyx:33-34): the default build with the
 * bin width given in bp; 0 on success, 1 if <oPath>/<igdName>.igd exists already (nothing is written then) */
int create_iGD(const char *iPath, const char *oPath, const char *igdName, int tile_size_bp);

#ifdef __cplusplus
}
#endif
#endif
