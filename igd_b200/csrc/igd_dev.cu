// igd_b200 CUDA layer: HBM image of an .igd database and the search kernels (sm_100a).
//
// Data layout in HBM (built once per igd_b200_open):
//   tiles[T]          16 B per (contig, bin) tile: {off4, n, ts, s0}
//                       off4 : first record of the tile in the SoA arrays, in units of 4 records
//                              (every tile is padded to a multiple of 4 records => 16 B aligned int4 loads)
//                       n    : nCnt of the tile                       (iGD_t.nCnt, src/igd_base.h:103)
//                       ts   : #records with start < bin boundary     (the per-query linear scan of
//                              src/igd_search.c:510-511 is a per-tile constant; computed once here)
//                       s0   : start of the first record              (the `qe > gData[0].start` gate, :477)
//   start/end/idx/value[P]  SoA int32 copies of gdata_t's fields (src/igd_base.h:41-46), tile-major in
//                       the file's order (each tile sorted by start).  Pad records: start=0x7f7f7f7f,
//                       end=0x80808080 (never overlap anything).
//   tbase[nctg+1]     global tile index of each contig's tile 0.
//
// Query batches are turned into (query, tile) SEGMENTS -- one per non-empty, non-gated tile the reference
// would visit for that query, first-bin-empty quirk included -- counting-sorted by tile on the device, so
// that concurrently running warps work on neighbouring tiles whatever the order of the input.
#include "igd_dev.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

thread_local char g_err[512] = "";

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(g_err, sizeof g_err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                     __FILE__, __LINE__);                                                          \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

struct TileMeta {
    uint32_t off4;
    int32_t n, ts, s0;
};
static_assert(sizeof(TileMeta) == 16, "TileMeta must be one int4");

struct DbView {
    const TileMeta *tiles;
    const int32_t *start, *end, *idx, *value;
    const int32_t *tbase;
    int32_t nctg, nbp, nfiles, ntiles;
};

// segment: one (query, tile) unit of work. w = sid | later<<31
struct __align__(16) Seg {
    int32_t qs, qe;
    uint32_t tile, w;
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int4 ldg4(const int32_t *p) { return __ldg(reinterpret_cast<const int4 *>(p)); }

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
// AoS gdata_t {idx,start,end,value} -> padded SoA.  toff[T+1] = unpadded record prefix (file order).
__global__ void k_transpose(const int4 *__restrict__ aos, int64_t rec0, int64_t nrec,
                            const int64_t *__restrict__ toff, int32_t T, const TileMeta *__restrict__ tiles,
                            int32_t *__restrict__ start, int32_t *__restrict__ end, int32_t *__restrict__ idx,
                            int32_t *__restrict__ value)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    int64_t g = rec0 + i;
    int32_t lo = 0, hi = T;                 // last t with toff[t] <= g
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(&toff[mid]) <= g) lo = mid; else hi = mid;
    }
    int4 r = __ldg(&aos[i]);
    int64_t dst = (int64_t)tiles[lo].off4 * 4 + (g - __ldg(&toff[lo]));
    idx[dst] = r.x, start[dst] = r.y, end[dst] = r.z, value[dst] = r.w;
}

// per tile: ts = #records with start < bd (bd arrives in .ts), s0 = first start
__global__ void k_tile_meta(TileMeta *tiles, int32_t T, const int32_t *__restrict__ start)
{
    int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    TileMeta m = tiles[t];
    const int32_t *s = start + (int64_t)m.off4 * 4;
    int32_t bd = m.ts, lo = 0, hi = m.n;
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (s[mid] < bd) lo = mid + 1; else hi = mid;
    }
    m.ts = lo;
    m.s0 = m.n > 0 ? s[0] : 0x7fffffff;
    tiles[t] = m;
}

// ------------------------------------------------------------------------------------------------
// query -> segments.  Restates the control flow of get_overlaps (src/igd_search.c:454-534; V=false:
// nothing at all when bin n1 is empty) and get_overlaps_v (:623-694; V=true: later bins are visited
// regardless).  f(tile, later, k) is called for each tile the reference would scan.
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
            f(tb + j, j > n1, k);
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
                                [&](int tile, bool, int) { atomicAdd(&tile_cnt[tile], 1u); });
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
    for_each_segment<V>(db, ctg[q], s, e, [&](int tile, bool later, int k) {
        uint32_t pos = tile_off[tile] + atomicSub(&tile_cnt[tile], 1u) - 1u;
        Seg sg;
        sg.qs = s, sg.qe = e, sg.tile = (uint32_t)tile;
        sg.w = (qseg_off ? sid0 + (uint32_t)k : 0u) | (later ? 0x80000000u : 0u);
        segs[pos] = sg;
    });
}

// ------------------------------------------------------------------------------------------------
// exclusive scan (three small kernels; n up to ~1e8).  out has n+1 entries, out[n] = total.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 512, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

template <class Tout>
__device__ __forceinline__ Tout block_exclusive(Tout x, Tout *total)
{
    __shared__ Tout wsum[SCAN_T / 32];
    __shared__ Tout btot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Tout inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Tout y = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        Tout s = lane < SCAN_T / 32 ? wsum[lane] : 0, si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Tout y = __shfl_up_sync(FULL, si, d);
            if (lane >= d) si += y;
        }
        if (lane < SCAN_T / 32) wsum[lane] = si - s;
        if (lane == SCAN_T / 32 - 1) btot = si;
    }
    __syncthreads();
    *total = btot;
    Tout r = wsum[w] + inc - x;
    __syncthreads();
    return r;
}

template <class Tin, class Tout>
__global__ void __launch_bounds__(SCAN_T) k_scan1(const Tin *__restrict__ in, Tout *__restrict__ out, int64_t n,
                                                    Tout *__restrict__ bsum)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    Tout v[SCAN_I], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        v[k] = base + k < n ? (Tout)in[base + k] : (Tout)0;
        s += v[k];
    }
    Tout tot, ex = block_exclusive<Tout>(s, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

template <class Tout>
__global__ void __launch_bounds__(SCAN_T) k_scan2(Tout *bsum, int nb)
{
    Tout carry = 0;
    for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
        int i = b0 + threadIdx.x;
        Tout x = i < nb ? bsum[i] : (Tout)0, tot;
        Tout ex = block_exclusive<Tout>(x, &tot);
        if (i < nb) bsum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}

template <class Tout>
__global__ void __launch_bounds__(SCAN_T) k_scan3(Tout *__restrict__ out, int64_t n, const Tout *__restrict__ bsum,
                                                    int nb)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_I;
    const Tout add = bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_I; ++k)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

// ------------------------------------------------------------------------------------------------
// the search kernel: one warp per segment.
//   1. 32-ary ballot search for cnt = #{start < qe} in the tile (tL = cnt-1): every round the 32 lanes
//      probe 32 evenly spaced records and one ballot narrows the range 32x -- 2 dependent loads for a
//      1 k-record tile instead of 10.
//   2. scan end[tS..tL] (tS = 0 in the query's first bin, the tile's dedupe start in later bins) with
//      128-bit loads; predicate end > qs (and value >= v).
//   3. OUT_HIST : hits[idx]++ in a block-private shared-memory histogram, flushed once per block
//      OUT_COUNT: per-segment hit count (first pass of -f)
//      OUT_FILL : records written in the reference's order (descending index) at the segment's offset
// ------------------------------------------------------------------------------------------------
enum { OUT_HIST = 0, OUT_COUNT = 1, OUT_FILL = 2 };

__device__ __forceinline__ int warp_count_lt(const int32_t *__restrict__ s, int n, int key, int lane)
{
    int lo = 0, hi = n;                                     // answer (count of start < key) in [lo, hi]
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5;
        const int p = min(lo + (lane + 1) * step - 1, hi - 1);
        const unsigned b = __ballot_sync(FULL, __ldg(&s[p]) < key);
        const int k = __popc(b);                            // sorted => true lanes form a prefix
        const int nlo = min(lo + k * step, hi);
        hi = k == 32 ? hi : min(lo + (k + 1) * step - 1, hi);
        lo = nlo;
    }
    const int i = lo + lane;
    const unsigned b = __ballot_sync(FULL, i < hi && __ldg(&s[i]) < key);
    return lo + __popc(b);
}

struct SearchArgs {
    DbView db;
    const Seg *segs;
    const uint32_t *nsegs;          // device scalar
    int32_t v;
    unsigned long long *hits;       // OUT_HIST
    uint32_t *seg_hits;             // OUT_COUNT  [sid]
    const unsigned long long *seg_off;   // OUT_FILL [sid]
    int32_t *out_s, *out_e, *out_f;      // OUT_FILL
};

template <bool V, int OUT, bool SMEM>
__global__ void __launch_bounds__(512) k_search(const SearchArgs a)
{
    extern __shared__ uint32_t hist[];
    const DbView &db = a.db;
    if (OUT == OUT_HIST && SMEM) {
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) hist[f] = 0;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const uint32_t nsegs = __ldg(a.nsegs);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nsegs; s += nwarps) {
        const int4 sg = __ldg(reinterpret_cast<const int4 *>(&a.segs[s]));
        const int qs = sg.x, qe = sg.y;
        const int4 tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[sg.z]));
        const int64_t off = (int64_t)(uint32_t)tm.x * 4;
        const int tS = ((uint32_t)sg.w >> 31) ? tm.z : 0;
        const uint32_t sid = (uint32_t)sg.w & 0x7fffffffu;
        const int hi = warp_count_lt(db.start + off, tm.y, qe, lane);        // tL + 1
        const int32_t *__restrict__ E = db.end + off;
        const int32_t *__restrict__ Vv = db.value + off;
        const int32_t *__restrict__ I = db.idx + off;
        if (OUT != OUT_FILL) {
            int cnt = 0;
            for (int b = (tS & ~3) + lane * 4; b < hi; b += 128) {
                const int4 e = ldg4(E + b);
                int4 val = make_int4(0, 0, 0, 0);
                if (V) val = ldg4(Vv + b);
                const int ev[4] = {e.x, e.y, e.z, e.w};
                const int vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = b + k;
                    if (i >= tS && i < hi && ev[k] > qs && (!V || vv[k] >= a.v)) {
                        if (OUT == OUT_HIST) {
                            const int f = __ldg(&I[i]);
                            if (SMEM) atomicAdd(&hist[f], 1u);
                            else atomicAdd(&a.hits[f], 1ull);
                        } else {
                            ++cnt;
                        }
                    }
                }
            }
            if (OUT == OUT_COUNT) {
#pragma unroll
                for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(FULL, cnt, d);
                if (lane == 0) a.seg_hits[sid] = (uint32_t)cnt;
            }
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
                    a.out_f[p] = __ldg(&I[i]);
                }
                pos += __popc(b);
            }
        }
    }
    if (OUT == OUT_HIST && SMEM) {
        __syncthreads();
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) {
            const uint32_t c = hist[f];
            if (c) atomicAdd(&a.hits[f], (unsigned long long)c);
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

// ------------------------------------------------------------------------------------------------
// -m hitsmap: getMap / getMap_v (src/igd_search.c:772-886).  One warp per (tile, record j): the record
// is a query against its own tile; partners i in [tS, last start < end_j] with end_i > start_j, where
// tS is the tile's dedupe start only when start_j < bin boundary (:802-804).  -v is the STRICT `> v`.
// work items = padded record slots, so a flat grid-stride loop over P records needs no extra index.
// ------------------------------------------------------------------------------------------------
template <bool V>
__global__ void __launch_bounds__(256) k_hitsmap(DbView db, const int32_t *__restrict__ rec_tile, int64_t P,
                                                  int32_t v, uint32_t *__restrict__ hitmap,
                                                  unsigned long long *__restrict__ nols)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    unsigned long long mine = 0;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < P; r += nw) {
        const int t = __ldg(&rec_tile[r]);
        if (t < 0) continue;                                 // pad slot
        const int4 tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
        const int64_t off = (int64_t)(uint32_t)tm.x * 4;
        const int j = (int)(r - off);
        const int qs = __ldg(&db.start[r]), qe = __ldg(&db.end[r]);
        if (V && !(__ldg(&db.value[r]) > v)) continue;
        if (!(qe > tm.w)) continue;
        // bin boundary of this tile: ts>0 records start before it; start_j < bd <=> j < ts (sorted by start)
        const int tS = j < tm.z ? tm.z : 0;
        const int hi = warp_count_lt(db.start + off, tm.y, qe, lane);
        const int jj = __ldg(&db.idx[r]);
        uint32_t *row = hitmap + (int64_t)jj * db.nfiles;
        for (int b = (tS & ~3) + lane * 4; b < hi; b += 128) {
            const int4 e = ldg4(db.end + off + b);
            int4 val = make_int4(0, 0, 0, 0);
            if (V) val = ldg4(db.value + off + b);
            const int ev[4] = {e.x, e.y, e.z, e.w};
            const int vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = b + k;
                if (i >= tS && i < hi && ev[k] > qs && (!V || vv[k] > v)) {
                    atomicAdd(&row[__ldg(&db.idx[off + i])], 1u);
                    ++mine;
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(FULL, mine, d);
    if (lane == 0 && mine) atomicAdd(nols, mine);
}

// rec_tile[slot] = tile of the record in that padded slot, -1 for pad slots
__global__ void k_rec_tile(const TileMeta *__restrict__ tiles, int32_t T, int32_t *__restrict__ rec_tile)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += nw) {
        const TileMeta m = tiles[t];
        const int64_t off = (int64_t)m.off4 * 4;
        const int np = (m.n + 3) & ~3;
        for (int i = lane; i < np; i += 32) rec_tile[off + i] = i < m.n ? (int32_t)t : -1;
    }
}

}  // namespace

// ================================================================================================
// host side of the CUDA layer
// ================================================================================================
struct igdb_dev {
    int device = 0;
    int sm_count = 0;
    int32_t nbp = 0, nctg = 0, nfiles = 0, T = 0;
    int64_t nrec = 0, P = 0;                // records, padded records
    std::vector<int64_t> toff;              // [T+1] unpadded prefix
    TileMeta *tiles = nullptr;
    int64_t *d_toff = nullptr;
    int32_t *start = nullptr, *end = nullptr, *idx = nullptr, *value = nullptr, *tbase = nullptr;
    int64_t bytes = 0;
    int64_t launches = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[2] = {nullptr, nullptr};               // around the last count-mode search kernel
    bool kev_valid = false;
    // staging / scratch (grown on demand)
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    void *d_stage = nullptr;
    size_t d_stage_cap = 0;
    int32_t *q_ctg = nullptr, *q_qs = nullptr, *q_qe = nullptr;
    int64_t q_cap = 0;
    uint32_t *tile_cnt = nullptr, *tile_off = nullptr;     // [T], [T+1]
    uint32_t *nseg = nullptr, *qseg_off = nullptr;          // [q_cap], [q_cap+1]
    void *bsum = nullptr;                                   // scan block sums
    size_t bsum_cap = 0;
    Seg *segs = nullptr;
    int64_t seg_cap = 0;
    uint32_t *seg_hits = nullptr;
    unsigned long long *seg_out_off = nullptr;
    int64_t segx_cap = 0;
    unsigned long long *d_hits = nullptr;                   // [nfiles]
    uint32_t *h_scalar = nullptr;                           // pinned readback
    DbView view() const
    {
        DbView v;
        v.tiles = tiles, v.start = start, v.end = end, v.idx = idx, v.value = value, v.tbase = tbase;
        v.nctg = nctg, v.nbp = nbp, v.nfiles = nfiles, v.ntiles = T;
        return v;
    }
};

namespace {

template <class T>
int dev_alloc(igdb_dev *d, T **p, size_t count)
{
    CK(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
    d->bytes += (int64_t)(count * sizeof(T));
    return 0;
}

template <class T>
int grow(T **p, int64_t *cap, int64_t need, int64_t slack_num = 5, int64_t slack_den = 4)
{
    if (need <= *cap) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    int64_t c = need * slack_num / slack_den + 1024;
    CK(cudaMalloc((void **)p, (size_t)c * sizeof(T)));
    *cap = c;
    return 0;
}

template <class Tin, class Tout>
int exclusive_scan(igdb_dev *d, const Tin *in, Tout *out, int64_t n, cudaStream_t st)
{
    const int nb = (int)((n + SCAN_B - 1) / SCAN_B);
    const size_t need = (size_t)(nb + 1) * sizeof(unsigned long long);
    if (need > d->bsum_cap) {
        if (d->bsum) CK(cudaFree(d->bsum));
        d->bsum = nullptr;
        CK(cudaMalloc(&d->bsum, need * 2));
        d->bsum_cap = need * 2;
    }
    Tout *bs = (Tout *)d->bsum;
    if (nb == 0) {
        CK(cudaMemsetAsync(out, 0, sizeof(Tout), st));
        return 0;
    }
    k_scan1<Tin, Tout><<<nb, SCAN_T, 0, st>>>(in, out, n, bs);
    k_scan2<Tout><<<1, SCAN_T, 0, st>>>(bs, nb);
    k_scan3<Tout><<<nb, SCAN_T, 0, st>>>(out, n, bs, nb);
    d->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

int ensure_query_cap(igdb_dev *d, int64_t n)
{
    if (n <= d->q_cap) return 0;
    int64_t c = 0;
    int64_t cap0 = d->q_cap;
    c = cap0; if (grow(&d->q_ctg, &c, n)) return -1;
    c = cap0; if (grow(&d->q_qs, &c, n)) return -1;
    c = cap0; if (grow(&d->q_qe, &c, n)) return -1;
    c = cap0; if (grow(&d->nseg, &c, n)) return -1;
    c = cap0; if (grow(&d->qseg_off, &c, n + 1)) return -1;
    d->q_cap = n * 5 / 4 + 1024 - 1;        // the smallest of the capacities above
    return 0;
}

// search kernel launch geometry: persistent-style grid, a whole number of CTAs per SM
template <bool V, int OUT>
int launch_search(igdb_dev *d, const SearchArgs &a, cudaStream_t st)
{
    const int threads = 512;
    const size_t hist_bytes = (size_t)d->nfiles * sizeof(uint32_t);
    const bool smem = OUT == OUT_HIST && hist_bytes <= 200 * 1024;
    const size_t shm = smem ? hist_bytes : 0;
    int per_sm = 0;
    if (OUT == OUT_HIST) CK(cudaEventRecord(d->kev[0], st));
    if (smem) {
        auto kern = k_search<V, OUT, true>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, shm));
        per_sm = std::max(per_sm, 1);
        kern<<<d->sm_count * per_sm, threads, shm, st>>>(a);
    } else {
        auto kern = k_search<V, OUT, false>;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
        per_sm = std::max(per_sm, 1);
        kern<<<d->sm_count * per_sm, threads, 0, st>>>(a);
    }
    if (OUT == OUT_HIST) {
        CK(cudaEventRecord(d->kev[1], st));
        d->kev_valid = true;
    }
    d->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// queries (device) -> tile-sorted segments.  Leaves the segment count in tile_off[T] and, if with_sid,
// the per-query segment prefix in qseg_off.  One stream sync (to size the segment buffer).
template <bool V>
int build_segments(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                   bool with_sid, cudaStream_t st, uint32_t *nsegs_out)
{
    const DbView db = d->view();
    const int thr = 256;
    const int nb = (int)((n + thr - 1) / thr);
    k_seg_count<V><<<nb, thr, 0, st>>>(db, n, ctg, qs, qe, d->tile_cnt, with_sid ? d->nseg : nullptr);
    d->launches += 1;
    CK(cudaGetLastError());
    if (exclusive_scan<uint32_t, uint32_t>(d, d->tile_cnt, d->tile_off, d->T, st)) return -1;
    if (with_sid && exclusive_scan<uint32_t, uint32_t>(d, d->nseg, d->qseg_off, n, st)) return -1;
    CK(cudaMemcpyAsync(d->h_scalar, d->tile_off + d->T, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t nsegs = *d->h_scalar;
    if (nsegs >= 0x7fffffffu) {
        snprintf(g_err, sizeof g_err, "query batch expands to too many (query, tile) segments");
        return -1;
    }
    if (grow(&d->segs, &d->seg_cap, (int64_t)nsegs)) return -1;
    k_seg_scatter<V><<<nb, thr, 0, st>>>(db, n, ctg, qs, qe, d->tile_cnt, d->tile_off,
                                          with_sid ? d->qseg_off : nullptr, d->segs);
    d->launches += 1;
    CK(cudaGetLastError());
    *nsegs_out = nsegs;
    return 0;
}

template <bool V>
int count_device_impl(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe, int32_t v,
                      unsigned long long *d_hits, cudaStream_t st)
{
    uint32_t nsegs = 0;
    if (build_segments<V>(d, n, ctg, qs, qe, false, st, &nsegs)) return -1;
    if (nsegs == 0) return 0;
    SearchArgs a{};
    a.db = d->view(), a.segs = d->segs, a.nsegs = d->tile_off + d->T, a.v = v, a.hits = d_hits;
    return launch_search<V, OUT_HIST>(d, a, st);
}

}  // namespace

extern "C" {

const char *igdb_dev_error(void) { return g_err; }

int igdb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void igdb_dev_destroy(igdb_dev *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    void *ptrs[] = {d->tiles, d->d_toff, d->start, d->end, d->idx, d->value, d->tbase, d->d_stage, d->q_ctg,
                    d->q_qs, d->q_qe, d->tile_cnt, d->tile_off, d->nseg, d->qseg_off, d->bsum, d->segs,
                    d->seg_hits, d->seg_out_off, d->d_hits};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (d->pinned) cudaFreeHost(d->pinned);
    if (d->h_scalar) cudaFreeHost(d->h_scalar);
    for (auto &e : d->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : d->kev)
        if (e) cudaEventDestroy(e);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

static int create_impl(igdb_dev *d, const int32_t *ntile, const int32_t *ncnt)
{
    CK(cudaSetDevice(d->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, d->device));
    d->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    for (auto &e : d->ev) CK(cudaEventCreate(&e));
    for (auto &e : d->kev) CK(cudaEventCreate(&e));
    CK(cudaMallocHost((void **)&d->h_scalar, 64));

    std::vector<int32_t> tbase(d->nctg + 1, 0);
    for (int c = 0; c < d->nctg; ++c) tbase[c + 1] = tbase[c] + ntile[c];
    d->T = tbase[d->nctg];
    std::vector<TileMeta> meta((size_t)d->T);
    d->toff.assign((size_t)d->T + 1, 0);
    int64_t p4 = 0;
    for (int c = 0, t = 0; c < d->nctg; ++c)
        for (int j = 0; j < ntile[c]; ++j, ++t) {
            const int32_t n = ncnt[t];
            meta[t].off4 = (uint32_t)p4;
            meta[t].n = n;
            meta[t].ts = d->nbp * j;            // bin boundary; k_tile_meta turns it into tS
            meta[t].s0 = 0x7fffffff;
            d->toff[t + 1] = d->toff[t] + n;
            p4 += (n + 3) / 4;
        }
    if (p4 >= (int64_t)0xffffffff) {
        snprintf(g_err, sizeof g_err, "database too large for 32-bit tile offsets");
        return -1;
    }
    d->nrec = d->toff[d->T];
    d->P = p4 * 4;
    if (dev_alloc(d, &d->tiles, (size_t)d->T) || dev_alloc(d, &d->d_toff, (size_t)d->T + 1) ||
        dev_alloc(d, &d->tbase, (size_t)d->nctg + 1) || dev_alloc(d, &d->start, (size_t)d->P + 4) ||
        dev_alloc(d, &d->end, (size_t)d->P + 4) || dev_alloc(d, &d->idx, (size_t)d->P + 4) ||
        dev_alloc(d, &d->value, (size_t)d->P + 4) || dev_alloc(d, &d->tile_cnt, (size_t)d->T + 1) ||
        dev_alloc(d, &d->tile_off, (size_t)d->T + 1) || dev_alloc(d, &d->d_hits, (size_t)d->nfiles))
        return -1;
    CK(cudaMemcpy(d->tiles, meta.data(), meta.size() * sizeof(TileMeta), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_toff, d->toff.data(), d->toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->tbase, tbase.data(), tbase.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemset(d->start, 0x7f, ((size_t)d->P + 4) * 4));
    CK(cudaMemset(d->end, 0x80, ((size_t)d->P + 4) * 4));
    CK(cudaMemset(d->idx, 0, ((size_t)d->P + 4) * 4));
    CK(cudaMemset(d->value, 0x80, ((size_t)d->P + 4) * 4));
    CK(cudaMemset(d->tile_cnt, 0, ((size_t)d->T + 1) * 4));
    return 0;
}

igdb_dev *igdb_dev_create(int device, int32_t nbp, int32_t nctg, const int32_t *ntile, const int32_t *ncnt,
                          int32_t nfiles)
{
    g_err[0] = 0;
    igdb_dev *d = new igdb_dev();
    d->device = device, d->nbp = nbp, d->nctg = nctg, d->nfiles = nfiles;
    if (create_impl(d, ntile, ncnt)) {
        igdb_dev_destroy(d);
        return nullptr;
    }
    return d;
}

void *igdb_dev_pinned(igdb_dev *d, size_t bytes)
{
    if (bytes > d->pinned_cap) {
        if (d->pinned) cudaFreeHost(d->pinned);
        d->pinned = nullptr, d->pinned_cap = 0;
        if (cudaMallocHost(&d->pinned, bytes) != cudaSuccess) return nullptr;
        d->pinned_cap = bytes;
    }
    return d->pinned;
}

int igdb_dev_upload(igdb_dev *d, int64_t rec0, int64_t nrec, const void *aos)
{
    if (nrec <= 0) return 0;
    CK(cudaSetDevice(d->device));
    const size_t bytes = (size_t)nrec * 16;
    if (bytes > d->d_stage_cap) {
        if (d->d_stage) CK(cudaFree(d->d_stage));
        d->d_stage = nullptr;
        CK(cudaMalloc(&d->d_stage, bytes));
        d->d_stage_cap = bytes;
    }
    CK(cudaMemcpyAsync(d->d_stage, aos, bytes, cudaMemcpyHostToDevice, d->stream));
    const int thr = 256;
    k_transpose<<<(unsigned)((nrec + thr - 1) / thr), thr, 0, d->stream>>>(
        (const int4 *)d->d_stage, rec0, nrec, d->d_toff, d->T, d->tiles, d->start, d->end, d->idx, d->value);
    d->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d->stream));       // the caller reuses its host buffer
    return 0;
}

int igdb_dev_finalize(igdb_dev *d)
{
    CK(cudaSetDevice(d->device));
    if (d->T > 0) {
        k_tile_meta<<<(d->T + 255) / 256, 256, 0, d->stream>>>(d->tiles, d->T, d->start);
        d->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(d->stream));
    if (d->d_stage) {
        CK(cudaFree(d->d_stage));
        d->d_stage = nullptr, d->d_stage_cap = 0;
    }
    return 0;
}

float igdb_dev_last_search_ms(igdb_dev *d)
{
    float ms = -1.f;
    if (!d->kev_valid) return ms;
    if (cudaEventSynchronize(d->kev[1]) != cudaSuccess) return ms;
    if (cudaEventElapsedTime(&ms, d->kev[0], d->kev[1]) != cudaSuccess) return -1.f;
    return ms;
}

int64_t igdb_dev_bytes(const igdb_dev *d) { return d->bytes; }
int64_t igdb_dev_launches(const igdb_dev *d) { return d->launches; }

int igdb_dev_count_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                          int mode_v, int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    if (n <= 0) return 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    const int64_t BATCH = 1 << 26;
    for (int64_t a = 0; a < n; a += BATCH) {
        const int64_t m = std::min(BATCH, n - a);
        int rc = mode_v ? count_device_impl<true>(d, m, d_ctg + a, d_qs + a, d_qe + a, v,
                                                 (unsigned long long *)d_hits, st)
                       : count_device_impl<false>(d, m, d_ctg + a, d_qs + a, d_qe + a, v,
                                                  (unsigned long long *)d_hits, st);
        if (rc) return rc;
    }
    return 0;
}

int igdb_dev_count_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe, int mode_v,
                        int32_t v, uint64_t *hits_delta, float *t_ms)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    CK(cudaMemsetAsync(d->d_hits, 0, (size_t)d->nfiles * 8, st));
    if (t_ms) t_ms[0] = t_ms[1] = t_ms[2] = 0.f;
    const int64_t BATCH = 1 << 26;
    for (int64_t a = 0; a < n; a += BATCH) {
        const int64_t m = std::min(BATCH, n - a);
        if (ensure_query_cap(d, m)) return -1;
        CK(cudaEventRecord(d->ev[0], st));
        CK(cudaMemcpyAsync(d->q_ctg, ctg + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->q_qs, qs + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->q_qe, qe + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(d->ev[1], st));
        int rc = mode_v ? count_device_impl<true>(d, m, d->q_ctg, d->q_qs, d->q_qe, v, d->d_hits, st)
                       : count_device_impl<false>(d, m, d->q_ctg, d->q_qs, d->q_qe, v, d->d_hits, st);
        if (rc) return rc;
        CK(cudaEventRecord(d->ev[2], st));
        if (t_ms) {
            CK(cudaEventSynchronize(d->ev[2]));
            float x = 0, y = 0;
            CK(cudaEventElapsedTime(&x, d->ev[0], d->ev[1]));
            CK(cudaEventElapsedTime(&y, d->ev[1], d->ev[2]));
            t_ms[0] += x, t_ms[1] += y;
        }
    }
    CK(cudaEventRecord(d->ev[2], st));
    CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(d->ev[3], st));
    CK(cudaStreamSynchronize(st));
    if (t_ms) CK(cudaEventElapsedTime(&t_ms[2], d->ev[2], d->ev[3]));
    return 0;
}

int64_t igdb_dev_full_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                           int64_t **query_off, int32_t **out_start, int32_t **out_end, int32_t **out_file)
{
    g_err[0] = 0;
    *query_off = nullptr, *out_start = *out_end = *out_file = nullptr;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    if (n >= (1 << 26)) {
        snprintf(g_err, sizeof g_err, "igdb_dev_full_host: batch of %lld queries is too large", (long long)n);
        return -1;
    }
    if (ensure_query_cap(d, std::max<int64_t>(n, 1))) return -1;
    int64_t *h_qoff = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    if (!h_qoff) return -1;
    *query_off = h_qoff;
    if (n == 0) return 0;
    CK(cudaMemcpyAsync(d->q_ctg, ctg, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qs, qs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qe, qe, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    uint32_t nsegs = 0;
    if (build_segments<false>(d, n, d->q_ctg, d->q_qs, d->q_qe, true, st, &nsegs)) return -1;
    if (nsegs == 0) return 0;
    {
        int64_t c = d->segx_cap;
        if (grow(&d->seg_hits, &c, (int64_t)nsegs + 1)) return -1;
        c = d->segx_cap;
        if (grow(&d->seg_out_off, &c, (int64_t)nsegs + 1)) return -1;
        d->segx_cap = c;
    }
    SearchArgs a{};
    a.db = d->view(), a.segs = d->segs, a.nsegs = d->tile_off + d->T, a.seg_hits = d->seg_hits;
    if (launch_search<false, OUT_COUNT>(d, a, st)) return -1;
    if (exclusive_scan<uint32_t, unsigned long long>(d, d->seg_hits, d->seg_out_off, nsegs, st)) return -1;
    unsigned long long *h_total = (unsigned long long *)(d->h_scalar + 2);
    CK(cudaMemcpyAsync(h_total, d->seg_out_off + nsegs, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t total = (int64_t)*h_total;
    int32_t *d_out = nullptr;
    int64_t *d_qoff = nullptr;
    CK(cudaMalloc((void **)&d_qoff, ((size_t)n + 1) * 8));
    k_query_off<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(n, d->qseg_off, d->seg_out_off, d_qoff);
    d->launches += 1;
    CK(cudaMemcpyAsync(h_qoff, d_qoff, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total > 0) {
        cudaError_t e = cudaMalloc((void **)&d_out, (size_t)total * 12);
        if (e != cudaSuccess) {
            cudaFree(d_qoff);
            snprintf(g_err, sizeof g_err, "-f result of %lld overlaps does not fit in device memory",
                     (long long)total);
            return -1;
        }
        a.seg_off = d->seg_out_off;
        a.out_s = d_out, a.out_e = d_out + total, a.out_f = d_out + 2 * total;
        if (launch_search<false, OUT_FILL>(d, a, st)) return -1;
        int32_t *hs = (int32_t *)malloc((size_t)total * 4), *he = (int32_t *)malloc((size_t)total * 4),
                *hf = (int32_t *)malloc((size_t)total * 4);
        if (!hs || !he || !hf) {
            snprintf(g_err, sizeof g_err, "out of host memory for %lld overlaps", (long long)total);
            return -1;
        }
        CK(cudaMemcpyAsync(hs, a.out_s, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(he, a.out_e, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hf, a.out_f, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        *out_start = hs, *out_end = he, *out_file = hf;
    }
    CK(cudaStreamSynchronize(st));
    if (d_out) CK(cudaFree(d_out));
    CK(cudaFree(d_qoff));
    return total;
}

int64_t igdb_dev_hitsmap_host(igdb_dev *d, int32_t v, uint32_t *hitmap)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    const size_t cells = (size_t)d->nfiles * d->nfiles;
    uint32_t *d_map = nullptr;
    int32_t *d_rt = nullptr;
    unsigned long long *d_n = nullptr;
    CK(cudaMalloc((void **)&d_map, std::max<size_t>(cells, 1) * 4));
    CK(cudaMalloc((void **)&d_rt, ((size_t)d->P + 4) * 4));
    CK(cudaMalloc((void **)&d_n, 8));
    CK(cudaMemcpyAsync(d_map, hitmap, cells * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_n, 0, 8, st));
    const int grid = d->sm_count * 8;
    k_rec_tile<<<grid, 256, 0, st>>>(d->tiles, d->T, d_rt);
    if (v > 0) k_hitsmap<true><<<grid, 256, 0, st>>>(d->view(), d_rt, d->P, v, d_map, d_n);
    else k_hitsmap<false><<<grid, 256, 0, st>>>(d->view(), d_rt, d->P, v, d_map, d_n);
    d->launches += 2;
    CK(cudaGetLastError());
    unsigned long long h_n = 0;
    CK(cudaMemcpyAsync(hitmap, d_map, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_n, d_n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_map));
    CK(cudaFree(d_rt));
    CK(cudaFree(d_n));
    return (int64_t)h_n;
}

}  // extern "C"
