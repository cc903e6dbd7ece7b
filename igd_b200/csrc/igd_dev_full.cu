// igd_b200 CUDA layer -- -f: the overlaps themselves, in the reference's order.
#include "igd_dev_impl.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// query -> segments.  Restates the control flow of get_overlaps (src/igd_search.c:454-534; V=false:
// nothing at all when bin n1 is empty) and get_overlaps_v (:623-694; V=true: later bins are visited
// regardless).  f(tile, later, k, bin boundary) is called for each tile the reference would scan.
// ------------------------------------------------------------------------------------------------
template <bool V, class F>
__device__ __forceinline__ int for_each_segment(const DbView &db, int c, int qs, int qe, F f)
{
    if (c < 0 || c >= db.nctg) return 0;                    // unknown chrom (:456-458)
    const int nbp = db.nbp;
    int n1 = qs / nbp, n2 = (qe - 1) / nbp;                 // C truncating division (:459)
    const int tb = __ldg(&db.tbase[c]);
    const int mT = __ldg(&db.tbase[c + 1]) - tb - 1;
    if (n1 > mT || n1 < 0) return 0;                        // :462 (n1 < 0 is out of bounds in the reference)
    n2 = min(n2, mT);
    if (!V && db.tiles[tb + n1].n == 0) return 0;           // first-bin-empty quirk (:468)
    int k = 0;
    for (int j = n1; j <= max(n1, n2); ++j) {
        const int4 m = __ldg(reinterpret_cast<const int4 *>(&db.tiles[tb + j]));
        if (m.y > 0 && qe > m.w) {                           // nCnt > 0 and the qe > start[0] gate
            f(tb + j, j > n1, k, j * nbp);
            ++k;
        }
    }
    return k;
}

template <bool V>
__global__ void k_seg_count(DbView db, int64_t n, const int32_t *__restrict__ ctg, const int32_t *__restrict__ qs,
                            const int32_t *__restrict__ qe, uint32_t *__restrict__ tile_cnt,
                            uint32_t *__restrict__ nseg)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int k = for_each_segment<V>(db, ctg[q], qs[q], qe[q],
                                [&](int tile, bool, int, int) { atomicAdd(&tile_cnt[tile], 1u); });
    if (nseg) nseg[q] = k;
}

// tile_cnt is counted back down to zero here, so it is clean for the next batch.
template <bool V>
__global__ void k_seg_scatter(DbView db, int64_t n, const int32_t *__restrict__ ctg, const int32_t *__restrict__ qs,
                              const int32_t *__restrict__ qe, uint32_t *__restrict__ tile_cnt,
                              const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ qseg_off,
                              Seg *__restrict__ segs)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int s = qs[q], e = qe[q];
    const uint32_t sid0 = qseg_off ? qseg_off[q] : 0u;
    for_each_segment<V>(db, ctg[q], s, e, [&](int tile, bool later, int k, int) {
        uint32_t pos = tile_off[tile] + atomicSub(&tile_cnt[tile], 1u) - 1u;
        Seg sg;
        sg.qs = s, sg.qe = e, sg.tile = (uint32_t)tile;
        sg.w = (qseg_off ? sid0 + (uint32_t)k : 0u) | (later ? 0x80000000u : 0u);
        segs[pos] = sg;
    });
}


// ------------------------------------------------------------------------------------------------
// -f: one warp per segment, in the reference's visiting order (get_overlaps_f1, src/igd_search.c:537-620).
//   1. 32-ary ballot search for cnt = #{start < qe} in the tile (tL = cnt-1): every round the 32 lanes
//      probe 32 evenly spaced records and one ballot narrows the range 32x -- 2 dependent loads for a
//      1 k-record tile instead of 10.
//   2. scan end[tS..tL] (tS = 0 in the query's first bin, the tile's dedupe start in later bins);
//      predicate end > qs.
//   3. OUT_COUNT: per-segment hit count (first pass)
//      OUT_FILL : records written in the reference's order (descending index) at the segment's offset
// ------------------------------------------------------------------------------------------------
enum { OUT_COUNT = 1, OUT_FILL = 2 };

struct SearchArgs {
    DbView db;
    const Seg *segs;
    const uint32_t *nsegs;          // device scalar
    uint32_t *seg_hits;             // OUT_COUNT  [sid]
    const unsigned long long *seg_off;   // OUT_FILL [sid]
    int32_t *out_s, *out_e, *out_f, *out_g;     // OUT_FILL (out_g: index within the tile, may be NULL)
};

template <int OUT>
__global__ void __launch_bounds__(512) k_search(const SearchArgs a)
{
    const DbView &db = a.db;
    const int lane = threadIdx.x & 31;
    const uint32_t nsegs = __ldg(a.nsegs);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nsegs; s += nwarps) {
        const int4 sg = __ldg(reinterpret_cast<const int4 *>(&a.segs[s]));
        const int qs = sg.x, qe = sg.y;
        const int4 tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[sg.z]));
        const int64_t off = (int64_t)(uint32_t)tm.x * 8;
        const int tS = ((uint32_t)sg.w >> 31) ? tm.z : 0;
        const uint32_t sid = (uint32_t)sg.w & 0x7fffffffu;
        const int hi = warp_count_lt(db.start + off, tm.y, qe, lane);        // tL + 1
        const int32_t *__restrict__ E = db.end + off;
        if (OUT == OUT_COUNT) {
            int cnt = 0;
            for (int b = (tS & ~3) + lane * 4; b < hi; b += 128) {
                const int4 e = ldg4(E + b);
                const int ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = b + k;
                    cnt += (i >= tS) & (i < hi) & (ev[k] > qs);
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
            if (lane == 0) a.seg_hits[sid] = (uint32_t)cnt;
        } else {
            unsigned long long pos = __ldg(&a.seg_off[sid]);
            for (int top = hi - 1; top >= tS; top -= 32) {
                const int i = top - lane;
                int e = 0;
                bool h = false;
                if (i >= tS) {
                    e = __ldg(&E[i]);
                    h = e > qs;
                }
                const unsigned b = __ballot_sync(FULL, h);
                if (h) {
                    const unsigned long long p = pos + __popc(b & ((1u << lane) - 1u));
                    a.out_s[p] = __ldg(&db.start[off + i]);
                    a.out_e[p] = e;
                    a.out_f[p] = __ldg(&db.idx[off + i]);
                    if (a.out_g) a.out_g[p] = i;
                }
                pos += __popc(b);
            }
        }
    }
}

// query_off[q] = seg_out_off[qseg_off[q]], q in [0, n]
__global__ void k_query_off(int64_t n, const uint32_t *__restrict__ qseg_off,
                            const unsigned long long *__restrict__ seg_out_off, int64_t *__restrict__ query_off)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= n) query_off[q] = (int64_t)seg_out_off[qseg_off[q]];
}


// -f kernels: persistent-style grid, a whole number of CTAs per SM
template <int OUT>
int launch_search(igdb_dev *d, const SearchArgs &a, cudaStream_t st)
{
    const int threads = 512;
    int per_sm = 0;
    auto kern = k_search<OUT>;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    per_sm = std::max(per_sm, 1);
    kern<<<d->sm_count * per_sm, threads, 0, st>>>(a);
    d->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// -f: queries (device) -> tile-sorted segments.  Leaves the segment count in tile_off[T] and the per-query
// segment prefix in qseg_off.  One stream sync (to size the segment buffer).
int build_segments(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                   cudaStream_t st, uint32_t *nsegs_out)
{
    const DbView db = d->view();
    const int thr = 256;
    const int nb = (int)((n + thr - 1) / thr);
    k_seg_count<false><<<nb, thr, 0, st>>>(db, n, ctg, qs, qe, d->tile_cnt, d->nseg);
    d->launches += 1;
    CK(cudaGetLastError());
    if (exclusive_scan<uint32_t, uint32_t>(d, d->tile_cnt, d->tile_off, d->T, st)) return -1;
    if (exclusive_scan<uint32_t, uint32_t>(d, d->nseg, d->qseg_off, n, st)) return -1;
    CK(cudaMemcpyAsync(d->h_scalar, d->tile_off + d->T, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t nsegs = *d->h_scalar;
    if (nsegs >= 0x7fffffffu) {
        snprintf(g_err, sizeof g_err, "query batch expands to too many (query, tile) segments");
        return -1;
    }
    if (grow(&d->segs, &d->seg_cap, (int64_t)nsegs)) return -1;
    k_seg_scatter<false><<<nb, thr, 0, st>>>(db, n, ctg, qs, qe, d->tile_cnt, d->tile_off, d->qseg_off, d->segs);
    d->launches += 1;
    CK(cudaGetLastError());
    *nsegs_out = nsegs;
    return 0;
}


// ---- multi-GPU -f: the lists of the (sorted) shards back into the order of the caller's query array ----------
// order[i] = index, in the caller's array, of the i-th query of the concatenated shards; cnt[i] = its overlaps.
__global__ void k_restore_counts(int64_t n, const long long *__restrict__ order, const long long *__restrict__ cnt,
                                 unsigned long long *__restrict__ cnt_orig)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt_orig[order[i]] = (unsigned long long)cnt[i];
}

// one warp per query: its run of overlaps from where the shards left it to where the caller's order wants it
__global__ void __launch_bounds__(256) k_restore_copy(int64_t n, const long long *__restrict__ order,
                                                       const unsigned long long *__restrict__ soff,
                                                       const unsigned long long *__restrict__ qoff, int64_t total,
                                                       const int32_t *__restrict__ in, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nw) {
        const unsigned long long src = soff[i], c = soff[i + 1] - src, dst = qoff[order[i]];
        for (unsigned long long k = lane; k < c; k += 32) {
            out[dst + k] = in[src + k];
            out[total + dst + k] = in[total + src + k];
            out[2 * total + dst + k] = in[2 * total + src + k];
        }
    }
}

}  // namespace

// -f core: the queries are in d->q_ctg/q_qs/q_qe (device).  Leaves query_off[n+1] (overlaps of query q are
// [query_off[q], query_off[q+1])) in qoff and the overlaps, SoA, in out: start[total] | end[total] | file[total]
// (| index within the tile [total] if with_g), in the reference's order.  Returns total.
// qoff and out come from two buffers that the handle keeps and grows (a cudaMalloc / cudaFree of the lists per call
// cost as much as the kernels): they are valid until the next -f / -s call on the handle.
static int pool_take(void **pool, size_t *cap, size_t bytes, DevBuf &b)
{
    if (bytes > *cap) {
        if (*pool) CK(cudaFree(*pool));
        *pool = nullptr, *cap = 0;
        const size_t want = bytes + (bytes >> 3) + 4096;
        if (cudaMalloc(pool, want) != cudaSuccess) {
            cudaGetLastError();
            *pool = nullptr;
            return -1;
        }
        *cap = want;
    }
    b.p = *pool, b.owned = false;
    return 0;
}

int64_t igdb::full_device(igdb_dev *d, int64_t n, cudaStream_t st, bool with_g, DevBuf &qoff, DevBuf &out)
{
    if (pool_take(&d->fp_qoff, &d->fp_qoff_cap, ((size_t)n + 1) * 8, qoff)) {
        snprintf(g_err, sizeof g_err, "-f: query offsets do not fit in device memory");
        return -1;
    }
    uint32_t nsegs = 0;
    if (build_segments(d, n, d->q_ctg, d->q_qs, d->q_qe, st, &nsegs)) return -1;
    if (nsegs == 0) {
        CK(cudaMemsetAsync(qoff.p, 0, ((size_t)n + 1) * 8, st));
        return 0;
    }
    {
        int64_t c = d->segx_cap;
        if (grow(&d->seg_hits, &c, (int64_t)nsegs + 1)) return -1;
        c = d->segx_cap;
        if (grow(&d->seg_out_off, &c, (int64_t)nsegs + 1)) return -1;
        d->segx_cap = c;
    }
    SearchArgs a{};
    a.db = d->view(), a.segs = d->segs, a.nsegs = d->tile_off + d->T, a.seg_hits = d->seg_hits;
    if (launch_search<OUT_COUNT>(d, a, st)) return -1;
    if (exclusive_scan<uint32_t, unsigned long long>(d, d->seg_hits, d->seg_out_off, nsegs, st)) return -1;
    unsigned long long *h_total = (unsigned long long *)(d->h_scalar + 2);
    CK(cudaMemcpyAsync(h_total, d->seg_out_off + nsegs, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t total = (int64_t)*h_total;
    k_query_off<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(n, d->qseg_off, d->seg_out_off, qoff.as<int64_t>());
    d->launches += 1;
    if (total > 0) {
        if (pool_take(&d->fp_out, &d->fp_out_cap, (size_t)total * (with_g ? 16 : 12), out)) {
            snprintf(g_err, sizeof g_err, "%lld overlaps do not fit in device memory", (long long)total);
            return -1;
        }
        int32_t *o = out.as<int32_t>();
        a.seg_off = d->seg_out_off;
        a.out_s = o, a.out_e = o + total, a.out_f = o + 2 * total, a.out_g = with_g ? o + 3 * total : nullptr;
        if (launch_search<OUT_FILL>(d, a, st)) return -1;
    }
    CK(cudaGetLastError());
    return total;
}


extern "C" {

// The result arrays live in a pinned arena owned by the handle (grown to the high-water mark, reused by every
// call): the lists come back at PCIe speed instead of being paged into freshly malloc'ed memory.  They stay
// valid until the next -f call on this handle.
int64_t igdb_dev_full_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                           int64_t **query_off, int32_t **out_start, int32_t **out_end, int32_t **out_file,
                           float *t_ms)
{
    g_err[0] = 0;
    *query_off = nullptr, *out_start = *out_end = *out_file = nullptr;
    if (t_ms) t_ms[0] = t_ms[1] = t_ms[2] = 0.f;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    cudaStream_t st = d->stream;
    if (n >= (1 << 26)) {
        snprintf(g_err, sizeof g_err, "igdb_dev_full_host: batch of %lld queries is too large", (long long)n);
        return -1;
    }
    auto arena = [&](size_t bytes) -> int {
        if (bytes <= d->arena_cap) return 0;
        if (d->arena) CK(cudaFreeHost(d->arena));
        d->arena = nullptr, d->arena_cap = 0;
        const size_t want = bytes + (bytes >> 3) + 4096;
        if (cudaHostAlloc(&d->arena, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            snprintf(g_err, sizeof g_err, "cannot pin %zu bytes of host memory for the overlap lists", want);
            return -1;
        }
        d->arena_cap = want;
        return 0;
    };
    const size_t qbytes = (((size_t)n + 1) * 8 + 255) & ~(size_t)255;
    if (n == 0 || d->T == 0) {
        if (arena(qbytes)) return -1;
        memset(d->arena, 0, qbytes);
        *query_off = (int64_t *)d->arena;
        return 0;
    }
    CK(cudaEventRecord(d->ev[0], st));
    if (upload_queries(d, n, ctg, qs, qe, st)) return -1;
    CK(cudaEventRecord(d->ev[1], st));
    DevBuf qoff, out;
    const int64_t total = full_device(d, n, st, false, qoff, out);
    if (total < 0) return -1;
    CK(cudaEventRecord(d->ev[2], st));
    if (arena(qbytes + (size_t)total * 12)) return -1;
    int64_t *h_qoff = (int64_t *)d->arena;
    int32_t *hs = (int32_t *)((char *)d->arena + qbytes);
    CK(cudaMemcpyAsync(h_qoff, qoff.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total > 0) CK(cudaMemcpyAsync(hs, out.p, (size_t)total * 12, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(d->ev[3], st));
    CK(cudaStreamSynchronize(st));
    *query_off = h_qoff;
    if (total > 0) *out_start = hs, *out_end = hs + total, *out_file = hs + 2 * total;
    if (t_ms) {
        CK(cudaEventElapsedTime(&t_ms[0], d->ev[0], d->ev[1]));
        CK(cudaEventElapsedTime(&t_ms[1], d->ev[1], d->ev[2]));
        CK(cudaEventElapsedTime(&t_ms[2], d->ev[2], d->ev[3]));
    }
    return total;
}

// -f with everything left on the device: queries come from device buffers, query_off[n+1] (int64) and the lists
// start[total] | end[total] | file[total] (int32, one allocation) stay in HBM, owned by the handle until the next
// call of this function (or destroy).  What a multi-GPU driver hands to its gather without crossing the host.
int64_t igdb_dev_full_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             void *stream, int64_t **d_query_off, int32_t **d_lists)
{
    g_err[0] = 0;
    *d_query_off = nullptr, *d_lists = nullptr;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    if (n < 0 || n >= (1 << 26)) {
        snprintf(g_err, sizeof g_err, "igdb_dev_full_device: batch of %lld queries is too large", (long long)n);
        return -1;
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    d->dres_qoff = d->dres_out = nullptr;
    DevBuf qoff, out;
    int64_t total = 0;
    if (n == 0 || d->T == 0) {
        if (pool_take(&d->fp_qoff, &d->fp_qoff_cap, ((size_t)n + 1) * 8, qoff)) return -1;
        CK(cudaMemsetAsync(qoff.p, 0, ((size_t)n + 1) * 8, st));
    } else {
        if (ensure_query_cap(d, n)) return -1;
        CK(cudaMemcpyAsync(d->q_ctg, d_ctg, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(d->q_qs, d_qs, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(d->q_qe, d_qe, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        total = full_device(d, n, st, false, qoff, out);
        if (total < 0) return -1;
    }
    CK(cudaStreamSynchronize(st));
    d->dres_qoff = qoff.p, d->dres_out = out.p;                 // (memory of the handle's pools)
    *d_query_off = (int64_t *)d->dres_qoff;
    *d_lists = (int32_t *)d->dres_out;
    return total;
}

// multi-GPU -f, on the gathering rank: d_lists_sorted = start | end | file [total] of the queries in SHARD order,
// d_cnt[i] = overlaps of the i-th of them, d_order[i] = its index in the caller's query array.  Writes the caller's
// order: d_query_off[n + 1] and d_lists_out (same layout).  Enqueued on stream; temporaries are stream-ordered.
int igdb_dev_full_restore(igdb_dev *d, int64_t n, const int64_t *d_order, const int64_t *d_cnt, const int32_t *d_lists_sorted,
                          int64_t total, int64_t *d_query_off, int32_t *d_lists_out, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    if (n <= 0) {
        CK(cudaMemsetAsync(d_query_off, 0, 8, st));
        return 0;
    }
    DevBuf corig, soff;
    if (corig.alloc(((size_t)n + 1) * 8) || soff.alloc(((size_t)n + 1) * 8)) return -1;
    const unsigned nb = (unsigned)((n + 255) / 256);
    k_restore_counts<<<nb, 256, 0, st>>>(n, (const long long *)d_order, (const long long *)d_cnt,
                                        corig.as<unsigned long long>());
    if (exclusive_scan<unsigned long long, unsigned long long>(d, corig.as<unsigned long long>(),
                                                                (unsigned long long *)d_query_off, n, st))
        return -1;
    if (exclusive_scan<unsigned long long, unsigned long long>(d, (const unsigned long long *)d_cnt,
                                                                soff.as<unsigned long long>(), n, st))
        return -1;
    if (total > 0)
        k_restore_copy<<<d->sm_count * 8, 256, 0, st>>>(n, (const long long *)d_order, soff.as<unsigned long long>(),
                                                        (const unsigned long long *)d_query_off, total, d_lists_sorted,
                                                        d_lists_out);
    d->launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));                              // the temporaries go out of scope
    return 0;
}

}  // extern "C"
