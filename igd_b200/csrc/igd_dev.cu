// igd_b200 CUDA layer: HBM image of an .igd database and the search kernels (sm_100a).
//
// Data layout in HBM (built once per igd_b200_open):
//   tiles[T]          16 B per (contig, bin) tile: {off8, n, ts, s0}
//                       off8 : first record of the tile in the SoA arrays, in units of 8 records
//                              (every tile is padded to a multiple of 8 records => 16 B aligned vector loads
//                              of the int32 arrays and of the packed 16-bit arrays below)
//                       n    : nCnt of the tile                       (iGD_t.nCnt, src/igd_base.h:103)
//                       ts   : #records with start < bin boundary     (the per-query linear scan of
//                              src/igd_search.c:510-511 is a per-tile constant; computed once here)
//                       s0   : start of the first record              (the `qe > gData[0].start` gate, :477)
//   start/end/idx/value[P]  SoA int32 copies of gdata_t's fields (src/igd_base.h:41-46), tile-major in
//                       the file's order (each tile sorted by start).  Pad records: start=INT_MAX,
//                       end=INT_MIN (overlap no query).
//   ps/pe/pf[P]       uint16 copies for count mode (when nbp <= 16384 and nFiles <= 65536), 6 B per record
//                       instead of 12: coordinates RELATIVE to the tile's bin boundary bd, clamped to the
//                       bin -- x' = clamp(x - bd + 1, 0, nbp + 1), exact for every comparison a query that
//                       visits this tile can make (see k_count_packed) -- and stored pre-biased,
//                       ps = start' + 1, pe = end' + 0x8000, so that two records are compared with one
//                       32-bit subtraction per predicate (bit 15 of each half is the answer).
//                       pf = file index.  Pads: ps = 0x8000 (starts after every query end), pe = 0x8000.
//   tbase[nctg+1]     global tile index of each contig's tile 0.
//
// Query batches are turned into (query, tile) SEGMENTS -- one per non-empty, non-gated tile the reference
// would visit for that query, first-bin-empty quirk included -- counting-sorted by tile on the device, so
// that the work is organised by tile whatever the order of the input.
//
// Count mode (-q, -v) is TILE-CENTRIC and RECORD-STATIONARY: a warp task ("item") is 128 consecutive
// records of one tile against up to 256 of the tile's segments.  Every lane keeps 4 records in registers
// (one 128-bit load each of start/end/idx[/value] -- the tile is read from HBM once however many queries
// hit it), the segments are streamed through the warp 32 at a time, pruned against the warp's
// [min start, max end] window with one ballot, and broadcast with shuffles; a lane adds the number of
// segments overlapping each of its records and issues ONE shared-memory atomic per record with a non-zero
// count (hits[idx] += count), not one per overlap.  -f keeps the query-centric warp-per-segment kernel
// because it needs the reference's per-query output order.
#include "igd_dev.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

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
    uint32_t off8;
    int32_t n, ts, s0;
};
static_assert(sizeof(TileMeta) == 16, "TileMeta must be one int4");

struct DbView {
    const TileMeta *tiles;
    const int32_t *start, *end, *idx, *value;
    const uint16_t *ps, *pe, *pf;           // packed tile-relative copies for count mode (may be NULL)
    const int32_t *tbase;
    int32_t nctg, nbp, nfiles, ntiles;
};

// segment: one (query, tile) unit of work. w = sid | later<<31
struct __align__(16) Seg {
    int32_t qs, qe;
    uint32_t tile, w;
};

constexpr unsigned FULL = 0xffffffffu;

// Programmatic dependent launch (sm_90+): the four kernels of a count batch are launched with the
// programmatic-stream-serialization attribute, so that the launch of kernel k + 1 overlaps the execution of
// kernel k.  pdl_wait() must come before the first access to anything the previous kernel wrote;
// pdl_launch_dependents() lets the next kernel's CTAs be scheduled as soon as every CTA of this one has started.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

__device__ __forceinline__ int4 ldg4(const int32_t *p) { return __ldg(reinterpret_cast<const int4 *>(p)); }

// tile-relative coordinates of the packed count-mode layout: x' = clamp(x - bd + 1, 0, nbp + 1).
//   a START (record start, always < bd + nbp; or a query start): 0 = before the bin, else 1 .. nbp
//   an END (record end, always > bd; or a query end > bd): 2 .. nbp + 1, nbp + 1 = at or past the bin's end
// For a record r of tile j and a query q that the reference lets visit tile j (0 <= qs < qe):
//   start_r < qe  <=>  start'_r < qe'      and      end_r > qs  <=>  end'_r > qs'
// (a clamped value only ever stands for "outside the bin on that side", where the comparison is decided).
__host__ __device__ __forceinline__ int32_t pk_rel(int32_t x, int32_t bd, int32_t nbp)
{
    return x < bd ? 0 : min(x - bd, nbp) + 1;       // x - bd cannot overflow: both >= 0 here
}
__host__ __device__ __forceinline__ int32_t pk_rel_end(int32_t x, int32_t bd, int32_t nbp)
{
    return x <= bd ? 1 : min(x - bd, nbp) + 1;
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
// AoS gdata_t {idx,start,end,value} -> padded SoA.  toff[T+1] = unpadded record prefix (file order).
__global__ void k_transpose(const int4 *__restrict__ aos, int64_t rec0, int64_t nrec,
                            const int64_t *__restrict__ toff, int32_t T, const TileMeta *__restrict__ tiles,
                            int32_t *__restrict__ start, int32_t *__restrict__ end, int32_t *__restrict__ idx,
                            int32_t *__restrict__ value, int32_t nbp, uint16_t *__restrict__ ps,
                            uint16_t *__restrict__ pe, uint16_t *__restrict__ pf)
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
    int64_t dst = (int64_t)tiles[lo].off8 * 8 + (g - __ldg(&toff[lo]));
    idx[dst] = r.x, start[dst] = r.y, end[dst] = r.z, value[dst] = r.w;
    if (ps) {                                   // tiles[].ts still holds the bin boundary here (k_tile_meta)
        const int32_t bd = tiles[lo].ts;
        ps[dst] = (uint16_t)(pk_rel(r.y, bd, nbp) + 1);
        pe[dst] = (uint16_t)(pk_rel_end(r.z, bd, nbp) + 0x8000);
        pf[dst] = (uint16_t)r.x;
    }
}

__global__ void k_fill16(uint16_t *__restrict__ p, int64_t n, uint16_t v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

__global__ void k_fill(int32_t *__restrict__ p, int64_t n, int32_t v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// per tile: ts = #records with start < bd (bd arrives in .ts), s0 = first start
__global__ void k_tile_meta(TileMeta *tiles, int32_t T, const int32_t *__restrict__ start)
{
    int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    TileMeta m = tiles[t];
    const int32_t *s = start + (int64_t)m.off8 * 8;
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
// exclusive scan (three small kernels; n up to ~1e8).  out has n+1 entries, out[n] = total.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 512, SCAN_I = 2, SCAN_B = SCAN_T * SCAN_I;

// two running sums scanned together: segments and work items per tile
struct Plan {
    unsigned long long segs, items;
};
__device__ __forceinline__ Plan operator+(Plan a, Plan b) { return Plan{a.segs + b.segs, a.items + b.items}; }
__device__ __forceinline__ Plan operator-(Plan a, Plan b) { return Plan{a.segs - b.segs, a.items - b.items}; }
__device__ __forceinline__ Plan &operator+=(Plan &a, Plan b) { a = a + b; return a; }

template <class T>
__device__ __forceinline__ T shfl_up(T v, int d) { return __shfl_up_sync(FULL, v, d); }
template <>
__device__ __forceinline__ Plan shfl_up<Plan>(Plan v, int d)
{
    return Plan{__shfl_up_sync(FULL, v.segs, d), __shfl_up_sync(FULL, v.items, d)};
}

template <class Tout>
__device__ __forceinline__ Tout block_exclusive(Tout x, Tout *total)
{
    __shared__ Tout wsum[SCAN_T / 32];
    __shared__ Tout btot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Tout inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Tout y = shfl_up<Tout>(inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        Tout s = lane < SCAN_T / 32 ? wsum[lane] : Tout{}, si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Tout y = shfl_up<Tout>(si, d);
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
    Tout carry = Tout{};
    for (int b0 = 0; b0 < nb; b0 += SCAN_T) {
        int i = b0 + threadIdx.x;
        Tout x = i < nb ? bsum[i] : Tout{}, tot;
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
// count mode, planning.  Segments are keyed by (tile, later): cnt2[2*tile + later].  One scan over the tiles
// yields, per tile, where its segments start (seg_off), where its later-bin segments start (seg_mid) and where
// its work items start (item_off); a tile with s segments and n records has ceil(n/GROUP) * ceil(s/SMAX) items.
// ------------------------------------------------------------------------------------------------
constexpr int SUB = 2;          // 128-record sub-groups per item
constexpr int GROUP = 128 * SUB; // records per item: a warp evaluates them 128 (4 per lane) at a time
constexpr int SMAX = 64;        // segments per item: what one pipeline stage holds (hot tiles get many items)

struct __align__(8) CSeg {       // count-mode segment: the query's interval, nothing else
    int32_t qs, qe;
};

// Which queries a count pipeline takes, and in which layout it searches:
//   M_ALL32  every query, int32 records (k_count_items)           -- databases the packed layout cannot hold
//   M_PACKED the queries with 0 <= qs < qe, packed 16-bit records (k_count_packed); the others are only counted
//            (cnt2[2T]) and, if there are any, a second pipeline in mode M_ABN32 takes exactly those
//   M_ABN32  the queries with qs < 0 or qs >= qe (accepted by parse_bed, src/igd_base.c:69: only en > 0 is
//            required), int32 records
enum { M_ALL32 = 0, M_PACKED = 1, M_ABN32 = 2 };

template <int MODE>
__device__ __forceinline__ bool mode_takes(int qs, int qe)
{
    const bool normal = qs >= 0 && qs < qe;
    return MODE == M_ALL32 ? true : (MODE == M_PACKED ? normal : !normal);
}

// Count pass and scatter pass share one walker that keeps CSEG_Q queries per thread in flight: their loads of the
// query, of the first-bin tile and their atomics are issued back to back (four independent dependency chains per
// thread instead of one), the later bins -- which few queries have -- follow in a per-query loop.  Restates the
// control flow of get_overlaps / get_overlaps_v exactly as for_each_segment does (see there).
constexpr int CSEG_Q = 4, CSEG_T = 256;
constexpr int CSEG_MAXCTG = 1024;           // contig table kept in shared memory up to this many contigs

template <bool V, int MODE, class F1, class FL>
__device__ __forceinline__ void cseg_walk(const DbView &db, int64_t n, const int32_t *__restrict__ ctg,
                                          const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                          uint32_t *__restrict__ abn, int abn_sign, F1 first, FL later)
{
    __shared__ int32_t tb_s[CSEG_MAXCTG + 1];
    const bool in_smem = db.nctg <= CSEG_MAXCTG;
    pdl_launch_dependents();
    if (in_smem) {
        for (int i = threadIdx.x; i <= db.nctg; i += blockDim.x) tb_s[i] = __ldg(&db.tbase[i]);
        __syncthreads();
    }
    const int nbp = db.nbp;
    const int64_t q0 = (int64_t)blockIdx.x * (CSEG_T * CSEG_Q) + threadIdx.x;
    pdl_wait();                                             // the queries / counters of the previous launch
    int c[CSEG_Q], s[CSEG_Q], e[CSEG_Q], tile[CSEG_Q], nlast[CSEG_Q], bd[CSEG_Q];
    bool ok[CSEG_Q];
    int nabn = 0;
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        const int64_t q = q0 + (int64_t)k * CSEG_T;
        ok[k] = q < n;
        c[k] = ok[k] ? ctg[q] : -1, s[k] = ok[k] ? qs[q] : 0, e[k] = ok[k] ? qe[q] : 0;
    }
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        if (ok[k] && !mode_takes<MODE>(s[k], e[k])) ok[k] = false, nabn += MODE == M_PACKED;
        ok[k] = ok[k] && c[k] >= 0 && c[k] < db.nctg;                       // unknown chrom (:456-458)
        tile[k] = nlast[k] = bd[k] = 0;
        if (ok[k]) {
            const int n1 = s[k] / nbp, n2 = (e[k] - 1) / nbp;               // C truncating division (:459)
            const int tb = in_smem ? tb_s[c[k]] : __ldg(&db.tbase[c[k]]);
            const int mT = (in_smem ? tb_s[c[k] + 1] : __ldg(&db.tbase[c[k] + 1])) - tb - 1;
            ok[k] = n1 >= 0 && n1 <= mT;                                    // :462
            tile[k] = tb + n1, nlast[k] = tb + max(n1, min(n2, mT)), bd[k] = n1 * nbp;
        }
    }
    int4 m[CSEG_Q];
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k)
        m[k] = ok[k] ? __ldg(reinterpret_cast<const int4 *>(&db.tiles[tile[k]])) : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        if (!V && m[k].y == 0) ok[k] = false;                               // first-bin-empty quirk (:468)
        if (ok[k] && m[k].y > 0 && e[k] > m[k].w) first(tile[k], s[k], e[k], bd[k]);   // nCnt > 0 and qe > start[0]
    }
    if (MODE == M_PACKED && nabn) atomicAdd(abn, (uint32_t)(abn_sign * nabn));   // left to the M_ABN32 pipeline
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        if (!ok[k]) continue;
        for (int t = tile[k] + 1; t <= nlast[k]; ++t) {                     // later bins (:500-531)
            const int4 mm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
            if (mm.y > 0 && e[k] > mm.w) later(t, s[k], e[k], bd[k] + (t - tile[k]) * nbp);
        }
    }
}

template <bool V, int MODE>
__global__ void __launch_bounds__(CSEG_T) k_cseg_count(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                        const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                        uint32_t *__restrict__ cnt2)
{
    cseg_walk<V, MODE>(
        db, n, ctg, qs, qe, &cnt2[2 * db.ntiles], +1, [&](int tile, int, int, int) { atomicAdd(&cnt2[2 * tile], 1u); },
        [&](int tile, int, int, int) { atomicAdd(&cnt2[2 * tile + 1], 1u); });
}

__device__ __forceinline__ Plan plan_of_tile(const DbView &db, const uint32_t *__restrict__ cnt2, int32_t t)
{
    const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cnt2) + t);
    const unsigned long long s = (unsigned long long)c.x + c.y;
    if (s == 0) return Plan{0, 0};
    const unsigned long long g = ((unsigned long long)(uint32_t)db.tiles[t].n + GROUP - 1) / GROUP;
    // a tile's segments start at an even index: 8-byte segments, 16-byte aligned bulk copies
    return Plan{(s + 1) & ~1ull, g * ((s + SMAX - 1) / SMAX)};
}

// The whole plan in ONE launch: a chained scan over the tiles fused with the item fill.  The grid has at most four
// 512-thread CTAs per SM (all co-resident), CTA b owns a contiguous range of tiles:
//   1. block-reduce the plan of its range and publish it (agg[b], then flag[b] = epoch);
//   2. wait for every CTA's aggregate: the sum of the earlier ones is this CTA's prefix, the sum of all is the
//      batch total -- if that does not fit the scratch buffers nothing more is written (the host grows them
//      and runs the batch again);
//   3. walk the range again in chunks of SCAN_T tiles: block scan + running carry give each tile its first
//      segment (seg_off), its first later-bin segment (seg_mid) and its first work item, and the tile's items are
//      written on the spot.
// Work items of tile t, group-major.  An item is self-contained, so the search kernel goes from the item to
// its records and segments without touching the tile table:
//   x = position of the group's first record in the SoA arrays, in units of 8 records
//   y = first segment
//   z = records in the group (bits 0-8, 1..GROUP) | first record of the group that starts inside this bin
//       (bits 9-17, 0..GROUP: the dedupe rule of later bins) | number of first-bin segments (bits 18-24, 0..64;
//       the later-bin segments follow them) | number of segments (bits 25-31, 1..64)
//   w = 0
struct PlanAgg {
    unsigned long long segs, items, recs;
};

__global__ void __launch_bounds__(SCAN_T) k_plan(DbView db, const uint32_t *__restrict__ cnt2, uint32_t epoch,
                                                  volatile PlanAgg *agg, volatile uint32_t *flag,
                                                  uint32_t *__restrict__ seg_off, uint32_t *__restrict__ seg_mid,
                                                  unsigned long long *__restrict__ tot, unsigned long long seg_cap,
                                                  unsigned long long item_cap, int4 *__restrict__ items)
{
    __shared__ unsigned long long red[3][SCAN_T / 32], red5[5][SCAN_T / 32];
    __shared__ unsigned long long sh_prefix[2], sh_total[3];
    const int32_t T = db.ntiles;
    const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int32_t per = (T + G - 1) / G, t0 = min(b * per, T), t1 = min(t0 + per, T);
    pdl_launch_dependents();
    pdl_wait();                                             // the segment counters

    // 1. aggregate of this CTA's range
    unsigned long long segs = 0, its = 0, recs = 0;
    for (int32_t t = t0 + threadIdx.x; t < t1; t += SCAN_T) {
        const Plan p = plan_of_tile(db, cnt2, t);
        segs += p.segs, its += p.items;
        if (p.items) {
            const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cnt2) + t);
            recs += (unsigned long long)(uint32_t)db.tiles[t].n * ((c.x + c.y + SMAX - 1) / SMAX);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        segs += __shfl_xor_sync(FULL, segs, d), its += __shfl_xor_sync(FULL, its, d);
        recs += __shfl_xor_sync(FULL, recs, d);
    }
    if (lane == 0) red[0][wid] = segs, red[1][wid] = its, red[2][wid] = recs;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long x = 0, y = 0, z = 0;
        for (int w = 0; w < SCAN_T / 32; ++w) x += red[0][w], y += red[1][w], z += red[2][w];
        agg[b].segs = x, agg[b].items = y, agg[b].recs = z;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&flag[b]), "r"(epoch) : "memory");
    }
    // 2. everybody's aggregates (every thread waits for its share of the CTAs)
    {
        unsigned long long v5[5] = {0, 0, 0, 0, 0};         // prefix segs / items, total segs / items / records
        for (int i = threadIdx.x; i < G; i += SCAN_T) {
            for (;;) {                                      // acquire: the aggregate is visible once the flag is
                uint32_t f;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(&flag[i]) : "memory");
                if (f == epoch) break;
                __nanosleep(20);
            }
            const unsigned long long x = agg[i].segs, y = agg[i].items, z = agg[i].recs;
            v5[2] += x, v5[3] += y, v5[4] += z;
            if (i < b) v5[0] += x, v5[1] += y;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
            for (int d = 16; d; d >>= 1) v5[k] += __shfl_xor_sync(FULL, v5[k], d);
        }
        __syncthreads();                                    // red[] was read by thread 0 above
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) red5[k][wid] = v5[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t5[5] = {0, 0, 0, 0, 0};
            for (int w = 0; w < SCAN_T / 32; ++w)
                for (int k = 0; k < 5; ++k) t5[k] += red5[k][w];
            sh_prefix[0] = t5[0], sh_prefix[1] = t5[1], sh_total[0] = t5[2], sh_total[1] = t5[3], sh_total[2] = t5[4];
            if (b == G - 1) {
                tot[0] = t5[2], tot[1] = t5[3], tot[2] = t5[4];
                tot[3] = cnt2[2 * T];                      // queries left to the M_ABN32 pipeline (M_PACKED only)
            }
        }
    }
    __syncthreads();
    const unsigned long long nseg = sh_total[0], nit = sh_total[1];
    if (nseg > seg_cap || nit > item_cap) return;
    if (b == G - 1 && threadIdx.x < 32) {                   // empty items up to a multiple of 32
        const unsigned long long k = nit + threadIdx.x;
        if (k < ((nit + 31) & ~31ull)) items[k] = make_int4(0, 0, 0, 0);
        if (threadIdx.x == 0) seg_off[T] = (uint32_t)nseg;
    }
    // 3. per-tile offsets and items
    Plan carry{sh_prefix[0], sh_prefix[1]};
    for (int32_t c0 = t0; c0 < t1; c0 += SCAN_T) {
        const int32_t t = c0 + threadIdx.x;
        const bool in = t < t1;
        const Plan mine = in ? plan_of_tile(db, cnt2, t) : Plan{0, 0};
        Plan ctot;
        const Plan ex = carry + block_exclusive<Plan>(mine, &ctot);
        carry += ctot;
        uint32_t i0 = (uint32_t)ex.items, ni = (uint32_t)mine.items, sa0 = (uint32_t)ex.segs, mid = 0, sb0 = 0;
        int4 tm = make_int4(0, 0, 0, 0);
        if (in) {
            const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cnt2) + t);
            mid = sa0 + c.x, sb0 = mid + c.y;
            seg_off[t] = sa0, seg_mid[t] = mid;
            if (ni) tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
        }
        auto write = [&](uint32_t k, uint32_t i0_, uint32_t a, uint32_t m, uint32_t e, int4 tm_) {
            const uint32_t ch = (e - a + SMAX - 1) / SMAX;
            const uint32_t g = k / ch, c = k - g * ch;
            const uint32_t sa = a + c * SMAX, sb = min(sa + (uint32_t)SMAX, e);
            const int r0 = (int)(g * GROUP);
            const uint32_t nv = (uint32_t)min(tm_.y - r0, GROUP);
            const uint32_t tsr = (uint32_t)min(max(tm_.z - r0, 0), GROUP);
            const uint32_t nfirst = min(max(m, sa), sb) - sa;
            items[i0_ + k] = make_int4((int)((uint32_t)tm_.x + g * (GROUP / 8)), (int)sa,
                                       (int)(nv | tsr << 9 | nfirst << 18 | (sb - sa) << 25), 0);
        };
        // a thread writes up to 32 items of its tile itself; tiles with more are then done by the whole warp
        if (ni <= 32)
            for (uint32_t k = 0; k < ni; ++k) write(k, i0, sa0, mid, sb0, tm);
        unsigned heavy = __ballot_sync(FULL, ni > 32);
        while (heavy) {
            const int src = __ffs(heavy) - 1;
            heavy &= heavy - 1;
            const uint32_t i0_ = __shfl_sync(FULL, i0, src), ni_ = __shfl_sync(FULL, ni, src);
            const uint32_t a_ = __shfl_sync(FULL, sa0, src), m_ = __shfl_sync(FULL, mid, src), e_ = __shfl_sync(FULL, sb0, src);
            const int4 tm_ = make_int4(__shfl_sync(FULL, tm.x, src), __shfl_sync(FULL, tm.y, src),
                                       __shfl_sync(FULL, tm.z, src), 0);
            for (uint32_t k = lane; k < ni_; k += 32) write(k, i0_, a_, m_, e_, tm_);
        }
    }
}

// cnt2 is counted back down to zero here, so it is clean for the next batch.  Nothing is written when the
// plan does not fit the scratch buffers: the host then grows them and runs the batch again.
// M_PACKED writes 4-byte segments in the tile's packed coordinates, pre-biased like the records:
//   bits 0-15 = qs' + 1, bits 16-31 = qe' + 0x8000.
template <bool V, int MODE>
__global__ void __launch_bounds__(CSEG_T) k_cseg_scatter(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                          const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                          uint32_t *__restrict__ cnt2, const uint32_t *__restrict__ seg_off,
                                                          const uint32_t *__restrict__ seg_mid,
                                                          const unsigned long long *__restrict__ tot,
                                                          unsigned long long seg_cap, unsigned long long item_cap,
                                                          void *__restrict__ segs)
{
    pdl_wait();                                             // the plan
    if (tot[0] > seg_cap || tot[1] > item_cap) return;
    const int nbp = db.nbp;
    auto put = [&](uint32_t pos, int s, int e, int bd) {
        if (MODE == M_PACKED)
            static_cast<uint32_t *>(segs)[pos] =
                (uint32_t)(pk_rel(s, bd, nbp) + 1) | (uint32_t)(pk_rel_end(e, bd, nbp) + 0x8000) << 16;
        else
            static_cast<CSeg *>(segs)[pos] = CSeg{s, e};
    };
    // (the queries left to the M_ABN32 pipeline are counted back down too)
    cseg_walk<V, MODE>(
        db, n, ctg, qs, qe, &cnt2[2 * db.ntiles], -1,
        [&](int tile, int s, int e, int bd) {
            put(__ldg(&seg_off[tile]) + atomicSub(&cnt2[2 * tile], 1u) - 1u, s, e, bd);
        },
        [&](int tile, int s, int e, int bd) {
            put(__ldg(&seg_mid[tile]) + atomicSub(&cnt2[2 * tile + 1], 1u) - 1u, s, e, bd);
        });
}

// ------------------------------------------------------------------------------------------------
// count mode, the search kernel (see the file header).  Restates the per-record predicate of get_overlaps
// (src/igd_search.c:489-493, 522-526: start < qe from the binary search, end > qs from the scan, and in later
// bins i >= tS, i.e. start >= bin boundary) and get_overlaps_v (:650-653, 680-683: ... && value >= v); which
// tiles a query reaches at all was decided when its segments were made (for_each_segment).
// ------------------------------------------------------------------------------------------------
struct CountArgs {
    DbView db;
    const int4 *items;
    const CSeg *segs;
    const unsigned long long *tot;      // device: {segments, items} of this batch
    unsigned long long seg_cap, item_cap;
    int32_t v;
    unsigned long long *hits;
};

constexpr int IMAX = 0x7fffffff, IMIN = -0x7fffffff - 1;
constexpr int COUNT_MAXW = 24;          // most warps per CTA of k_count_items (register budget: 65536 / 768 = 85)
constexpr int COUNT_K = 3;              // pipeline stages per warp

// ---- asynchronous copies (cp.async -> LDGSTS), shared::cta addresses as 32-bit ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ int4 lds128(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 lds64(uint32_t addr)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// c += (s < qe && e > qs)
__device__ __forceinline__ void count_if(int &c, int s, int e, int qs, int qe)
{
    asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, %4;\n\tsetp.gt.and.s32 p, %2, %3, p;\n\t@p add.s32 %0, %0, 1;\n\t}"
        : "+r"(c)
        : "r"(s), "r"(e), "r"(qs), "r"(qe));
}

// the segments selected by mask m (lanes of q) against this lane's 4 records
__device__ __forceinline__ void count_segments(unsigned m, const int2 &q, const int (&s)[4], const int (&e)[4],
                                               int (&c)[4])
{
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int qs = __shfl_sync(FULL, q.x, src), qe = __shfl_sync(FULL, q.y, src);
#pragma unroll
        for (int k = 0; k < 4; ++k) count_if(c[k], s[k], e[k], qs, qe);
    }
}

// Shared memory of a CTA: [histogram, nfiles x u32, rounded to 128 B] then per warp K stages of
//   { start[GROUP] | end[GROUP] | idx[GROUP] | (value[GROUP]) | 64 segments x 8 B }  -- the records and segments of
// one item.  Inside an array, sub-group h (records 128 h .. 128 h + 127) is the 512-byte run at 512 h.
template <bool V>
struct CountSmem {
    static constexpr uint32_t ARR = GROUP * 4;                          // one SoA array of a stage
    static constexpr uint32_t QOFF = (V ? 4 : 3) * ARR;
    static constexpr uint32_t STAGE = QOFF + SMAX * 8;
    static constexpr uint32_t WARP = COUNT_K * STAGE;
    static size_t hist_bytes(int nfiles) { return ((size_t)nfiles * 4 + 127) / 128 * 128; }
    // warps of the (single) CTA of an SM: as many pipelines as fit beside the histogram
    static int warps(int nfiles, bool smem_hist)
    {
        const size_t total = 226 * 1024, hist = smem_hist ? hist_bytes(nfiles) : 0;
        if (hist >= total) return 0;                        // the histogram alone does not fit
        return (int)std::min<size_t>((total - hist) / WARP, COUNT_MAXW);
    }
    static size_t cta_bytes(int nfiles, bool smem_hist, int nwarps)
    {
        return (smem_hist ? hist_bytes(nfiles) : 0) + (size_t)nwarps * WARP;
    }
};

// The count-mode search kernel: ONE CTA per SM, as many warps as shared memory has room for pipelines.  Each
// warp owns blocks of 32 consecutive items (block b of warp w is items [32 (b NW + w), +32); lane j keeps
// descriptor j of the current and of the next block in registers) and runs a private K-stage asynchronous-copy
// ring over them: K - 1 items ahead of the one being evaluated, every lane cp.async's ITS OWN records of each
// SoA array (16 B = 4 records per sub-group) and its own one or two segments into its own slots of the stage,
// so no lane ever reads another lane's bytes and the ring needs no barrier at all -- a lane's
// cp.async.wait_group covers exactly the bytes it is about to read.
template <bool V, bool SMEM>
__global__ void __launch_bounds__(COUNT_MAXW * 32, 1) k_count_items(const CountArgs a)
{
    using L = CountSmem<V>;
    extern __shared__ __align__(128) unsigned char smem[];
    const DbView &db = a.db;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem);
    const uint32_t hist_bytes = SMEM ? ((uint32_t)db.nfiles * 4 + 127) / 128 * 128 : 0;
    if (SMEM) {
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) hist[f] = 0;
        __syncthreads();
    }
    const uint32_t hist_s = smem_u32(smem);
    const int lane = threadIdx.x & 31, W = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const uint32_t wbase = hist_s + hist_bytes + (uint32_t)wib * L::WARP;
    const unsigned long long nseg = __ldg(&a.tot[0]), nit = __ldg(&a.tot[1]);
    const uint32_t n_items = (nseg > a.seg_cap || nit > a.item_cap) ? 0u : (uint32_t)nit;
    const uint32_t nblk = (n_items + 31) >> 5, NW = gridDim.x * W, w = blockIdx.x * W + wib;
    const uint32_t nb = w < nblk ? (nblk - w + NW - 1) / NW : 0;       // blocks of this warp
    const uint32_t N = nb * 32;                                         // item slots of this warp

    const int4 zero = make_int4(0, 0, 0, 0);
    // descriptor j of block 0 / block 1 of this warp (the item array is padded to a multiple of 32)
    int4 dcur = nb > 0 ? __ldg(&a.items[(size_t)w * 32 + lane]) : zero;
    int4 dnxt = nb > 1 ? __ldg(&a.items[(size_t)(NW + w) * 32 + lane]) : zero;
    // this lane's column of the SoA arrays and of the segment array
    const int32_t *const pS = db.start + lane * 4, *const pE = db.end + lane * 4, *const pI = db.idx + lane * 4;
    const int32_t *const pV = db.value + lane * 4;
    const CSeg *const pQ = a.segs + lane;
    const uint32_t lslot = wbase + lane * 16, qslot = wbase + L::QOFF + lane * 8;

    // issue the copies of item slot n (descriptor in lane n % 32 of dd) into stage n % K
    auto issue = [&](uint32_t n, const int4 &dd) {
        const int j = n & 31;
        const uint32_t x = __shfl_sync(FULL, dd.x, j), y = __shfl_sync(FULL, dd.y, j);
        const uint32_t z = __shfl_sync(FULL, dd.z, j);
        const int nv = z & 0x1ff, ns = z >> 25;
        const uint32_t so = (n % COUNT_K) * L::STAGE;
        const size_t p = (size_t)x * 8;
#pragma unroll
        for (int h = 0; h < SUB; ++h)
            if (h * 128 + lane * 4 < nv) {
                cp_async16(lslot + so + h * 512, pS + p + h * 128);
                cp_async16(lslot + so + L::ARR + h * 512, pE + p + h * 128);
                cp_async16(lslot + so + 2 * L::ARR + h * 512, pI + p + h * 128);
                if (V) cp_async16(lslot + so + 3 * L::ARR + h * 512, pV + p + h * 128);
            }
        if (lane < ns) cp_async8(qslot + so, pQ + y);
        if (lane + 32 < ns) cp_async8(qslot + so + 256, pQ + y + 32);
    };

#pragma unroll
    for (int n = 0; n < COUNT_K - 1; ++n) {                             // prologue: slots 0 .. K-2 (block 0)
        if ((uint32_t)n < N) issue(n, dcur);
        cp_async_commit();
    }

    for (uint32_t n = 0; n < N; ++n) {
        const int j = n & 31;
        // slot n + K - 1 goes into the stage that slot n - 1 has released
        {
            const uint32_t m = n + COUNT_K - 1;
            if (m < N) {
                if ((m >> 5) == (n >> 5)) issue(m, dcur);
                else issue(m, dnxt);
            }
            cp_async_commit();
        }
        cp_async_wait<COUNT_K - 1>();                                   // slot n has landed (this lane's bytes)
        const uint32_t dz = __shfl_sync(FULL, dcur.z, j);
        const uint32_t so = (n % COUNT_K) * L::STAGE;
        const int nv = dz & 0x1ff, tsr = (dz >> 9) & 0x1ff, nf = (dz >> 18) & 0x7f, ns = dz >> 25;
        int2 q0 = make_int2(IMAX, IMIN), q1 = make_int2(IMAX, IMIN);    // segments lane and 32 + lane of the item
        if (lane < ns) q0 = lds64(qslot + so);
        if (lane + 32 < ns) q1 = lds64(qslot + so + 256);
#pragma unroll
        for (int h = 0; h < SUB; ++h) {
            if (h * 128 >= nv) break;
            // this lane's 4 records of the sub-group; a slot past the group's end gets (IMAX, IMIN) and can
            // match no query, and so do the pad records inside the tile's last block of 4 (written as
            // (IMAX, IMIN) at upload)
            const int r0 = h * 128 + lane * 4;
            const bool have = r0 < nv;
            int4 S = lds128(lslot + so + h * 512), E = lds128(lslot + so + L::ARR + h * 512);
            const int4 I = lds128(lslot + so + 2 * L::ARR + h * 512);
            if (V) {
                const int4 VV = lds128(lslot + so + 3 * L::ARR + h * 512);
                if (VV.x < a.v) S.x = IMAX;                             // filtered out by -v: matches no query
                if (VV.y < a.v) S.y = IMAX;
                if (VV.z < a.v) S.z = IMAX;
                if (VV.w < a.v) S.w = IMAX;
            }
            const int s[4] = {have ? S.x : IMAX, have ? S.y : IMAX, have ? S.z : IMAX, have ? S.w : IMAX};
            const int e[4] = {have ? E.x : IMIN, have ? E.y : IMIN, have ? E.z : IMIN, have ? E.w : IMIN};
            const int f[4] = {I.x, I.y, I.z, I.w};
            int c[4] = {0, 0, 0, 0};
            const int smin = min(min(s[0], s[1]), min(s[2], s[3]));
            const int wsmin = V ? __reduce_min_sync(FULL, smin) : __shfl_sync(FULL, s[0], 0);     // sorted by start
            const int emax = __reduce_max_sync(FULL, max(max(e[0], e[1]), max(e[2], e[3])));
            // first-bin segments [0, nf): every record of the tile is a candidate
            if (nf > 0) {
                count_segments(__ballot_sync(FULL, lane < nf && q0.y > wsmin && q0.x < emax), q0, s, e, c);
                if (nf > 32) count_segments(__ballot_sync(FULL, lane + 32 < nf && q1.y > wsmin && q1.x < emax), q1, s, e, c);
            }
            // later-bin segments [nf, ns): only records that start inside this bin (index >= tS,
            // src/igd_search.c:510-511,522)
            if (ns > nf) {
                int sl[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) sl[k] = r0 + k >= tsr ? s[k] : IMAX;
                const int wslmin = __reduce_min_sync(FULL, min(min(sl[0], sl[1]), min(sl[2], sl[3])));
                if (wslmin != IMAX) {
                    if (nf < 32)
                        count_segments(__ballot_sync(FULL, lane >= nf && q0.y > wslmin && q0.x < emax), q0, sl, e, c);
                    if (ns > 32)
                        count_segments(__ballot_sync(FULL, lane + 32 >= nf && q1.y > wslmin && q1.x < emax), q1, sl, e, c);
                }
            }
            // hits[idx] += count: one atomic per record that overlaps anything
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (SMEM) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                                 :
                                 : "r"(hist_s + 4u * (uint32_t)f[k]), "r"(c[k])
                                 : "memory");
                } else if (c[k]) {
                    atomicAdd(&a.hits[f[k]], (unsigned long long)c[k]);
                }
            }
        }
        if (j == 31) {                                                  // next block of descriptors
            dcur = dnxt;
            const uint32_t b2 = (n >> 5) + 2;
            dnxt = b2 < nb ? __ldg(&a.items[(size_t)(b2 * NW + w) * 32 + lane]) : zero;
        }
    }
    cp_async_wait<0>();
    if (SMEM) {
        __syncthreads();
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) {
            const uint32_t c = hist[f];
            if (c) atomicAdd(&a.hits[f], (unsigned long long)c);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// count mode on the packed layout (the default path): the same tile-centric, record-stationary scheme as
// k_count_items, on 16-bit tile-relative coordinates -- half the bytes per record and two records per
// 32-bit operation.
//   records  S = start' + 1 (1 .. nbp + 1; 0x8000 = pad / filtered out by -v), E = end' + 0x8000
//   segments QS = qs' + 1, QE = qe' + 0x8000, each replicated into both halves of a register
//   A = QE2 - S2 : bit 15 of a half is set  <=>  qe' > start'   (the halves are 0x8000 + qe' - S >= 0: no borrow)
//   B = E2 - QS2 : bit 15 of a half is set  <=>  end' > qs'     (0x8000 + end' - QS >= 0x8000 - nbp - 1 > 0)
//   counts (two 16-bit counters per register, <= 64 per item) += (A & B & 0x80008000) >> 15
// i.e. 2 subtractions, 1 LOP3 and 1 shift-add per PAIR of (record, query) tests; the int32 kernel spends 3
// ALU-pipe instructions per test.  An item is 256 records (8 per lane: one 16-byte load per array) x <= 64
// segments.  Later-bin segments see S replaced by 0x8000 for the records that start before the bin (the
// reference's i >= tS, src/igd_search.c:510-511,522).
// Scheduling: one CTA per SM; a warp owns every NW-th BLOCK of 8 consecutive items (static, fine-grained
// interleave: the spread between warps is at most one block), four blocks' descriptors are fetched with one
// coalesced load.  Per warp a private K-stage cp.async ring as in k_count_items.
// ------------------------------------------------------------------------------------------------
constexpr int PK_K = 4;                 // pipeline stages per warp
constexpr int PK_MAXW = 24;             // most warps per CTA (register budget 65536 / 768 = 85)
constexpr int PK_IB = 8;                // items per block

template <bool V>
struct PackedSmem {
    static constexpr uint32_t ARR = GROUP * 2;                          // one 16-bit array of a stage
    static constexpr uint32_t VOFF = 3 * ARR;                           // int32 value[GROUP] (-v only)
    static constexpr uint32_t QOFF = VOFF + (V ? GROUP * 4 : 0);
    static constexpr uint32_t STAGE = QOFF + SMAX * 4;
    static constexpr uint32_t LIST = PK_K * STAGE;                      // survivor list: SMAX x {QS2, QE2}
    static constexpr uint32_t WARP = LIST + SMAX * 8;
    static size_t hist_bytes(int nfiles) { return ((size_t)nfiles * 4 + 127) / 128 * 128; }
    static int warps(int nfiles, bool smem_hist)
    {
        const size_t total = 226 * 1024, hist = smem_hist ? hist_bytes(nfiles) : 0;
        if (hist >= total) return 0;                        // the histogram alone does not fit
        return (int)std::min<size_t>((total - hist) / WARP, PK_MAXW);
    }
    static size_t cta_bytes(int nfiles, bool smem_hist, int nwarps)
    {
        return (smem_hist ? hist_bytes(nfiles) : 0) + (size_t)nwarps * WARP;
    }
};

__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
// segments [i0, i1) of the warp's survivor list against this lane's 4 record pairs
__device__ __forceinline__ void count_pairs(uint32_t list, int i0, int i1, const uint32_t (&S)[4],
                                            const uint32_t (&E)[4], uint32_t (&c)[4])
{
    auto eval = [&](const int2 q) {
        const uint32_t s2 = (uint32_t)q.x, e2 = (uint32_t)q.y;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t t;                                     // (A & B & 0x80008000): one 3-input LOP3
            asm("lop3.b32 %0, %1, %2, 0x80008000, 0x80;" : "=r"(t) : "r"(e2 - S[k]), "r"(E[k] - s2));
            c[k] += t >> 15;                                // LEA.HI
        }
    };
    uint32_t p = list + 8u * (uint32_t)i0;
    const uint32_t end = list + 8u * (uint32_t)i1;
#pragma unroll 1
    for (; p + 16 <= end; p += 16) {                        // same address in every lane: broadcast loads
        const int2 q0 = lds64(p), q1 = lds64(p + 8);
        eval(q0);
        eval(q1);
    }
    if (p < end) eval(lds64(p));
}

template <bool V, bool SMEM>
__global__ void __launch_bounds__(PK_MAXW * 32, 1) k_count_packed(const CountArgs a)
{
    using L = PackedSmem<V>;
    extern __shared__ __align__(128) unsigned char smem[];
    const DbView &db = a.db;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem);
    const uint32_t hist_bytes = SMEM ? ((uint32_t)db.nfiles * 4 + 127) / 128 * 128 : 0;
    if (SMEM) {
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) hist[f] = 0;
        __syncthreads();
    }
    const uint32_t hist_s = smem_u32(smem);
    const int lane = threadIdx.x & 31, W = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const uint32_t wbase = hist_s + hist_bytes + (uint32_t)wib * L::WARP;
    pdl_wait();                                             // the plan, the items, the segments
    const unsigned long long nseg = a.tot[0], nit = a.tot[1];
    const uint32_t n_items = (nseg > a.seg_cap || nit > a.item_cap) ? 0u : (uint32_t)nit;
    const uint32_t nblk = (n_items + PK_IB - 1) / PK_IB, NW = gridDim.x * W, w = blockIdx.x * W + wib;
    const uint32_t nb = w < nblk ? (nblk - w + NW - 1) / NW : 0;       // blocks of this warp
    const uint32_t N = nb * PK_IB;                                      // item slots of this warp

    // descriptor registers: lane j of quad-block g holds item (j % 8) of this warp's block 4 g + j / 8
    const int4 zero = make_int4(0, 0, 0, 0);
    auto load_desc = [&](uint32_t g) {
        const uint32_t b = 4 * g + (lane >> 3);
        const size_t it = ((size_t)b * NW + w) * PK_IB + (lane & 7);
        return (b < nb && it < n_items) ? __ldg(&a.items[it]) : zero;
    };
    int4 dcur = load_desc(0), dnxt = load_desc(1);
    const uint16_t *const pS = db.ps + lane * 8, *const pE = db.pe + lane * 8, *const pF = db.pf + lane * 8;
    const int32_t *const pV = db.value + lane * 8;
    const uint32_t *const pQ = static_cast<const uint32_t *>(static_cast<const void *>(a.segs)) + lane;
    const uint32_t lslot = wbase + lane * 16, vslot = wbase + L::VOFF + lane * 32, qslot = wbase + L::QOFF + lane * 4;

    // The stages start out as pads: a lane without records in an item reads what an earlier item (or this fill)
    // left in its slots -- its counts are discarded, and the prune window only gets wider.
    for (uint32_t o = lane * 16; o < L::LIST; o += 512)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(wbase + o), "r"(0x80008000u) : "memory");
    __syncwarp();

    // issue the copies of item slot n (descriptor in lane n % 32 of dd) into stage n % K
    auto issue = [&](uint32_t n, const int4 &dd) {
        const int j = n & 31;
        const uint32_t x = __shfl_sync(FULL, dd.x, j), y = __shfl_sync(FULL, dd.y, j), z = __shfl_sync(FULL, dd.z, j);
        const int nv = z & 0x1ff, ns = z >> 25;
        const uint32_t so = (n % PK_K) * L::STAGE;
        const size_t p = (size_t)x * 8;
        if (lane * 8 < nv) {
            cp_async16(lslot + so, pS + p);
            cp_async16(lslot + so + L::ARR, pE + p);
            cp_async16(lslot + so + 2 * L::ARR, pF + p);
            if (V) {
                cp_async16(vslot + so, pV + p);
                cp_async16(vslot + so + 16, pV + p + 4);
            }
        }
        if (lane < ns) cp_async4(qslot + so, pQ + y);
        if (lane + 32 < ns) cp_async4(qslot + so + 128, pQ + y + 32);
    };

#pragma unroll
    for (int n = 0; n < PK_K - 1; ++n) {                                // prologue: slots 0 .. K-2 (quad-block 0)
        if ((uint32_t)n < N) issue(n, dcur);
        cp_async_commit();
    }

    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t list = wbase + L::LIST;
    for (uint32_t n = 0; n < N; ++n) {
        const int j = n & 31;
        {
            const uint32_t m = n + PK_K - 1;                            // goes into the stage slot n - 1 released
            if (m < N) {
                if ((m >> 5) == (n >> 5)) issue(m, dcur);
                else issue(m, dnxt);
            }
            cp_async_commit();
        }
        cp_async_wait<PK_K - 1>();                                      // slot n has landed (this lane's bytes)
        const uint32_t dz = __shfl_sync(FULL, dcur.z, j);
        const uint32_t so = (n % PK_K) * L::STAGE;
        const int nv = dz & 0x1ff, tsr = (dz >> 9) & 0x1ff, nf = (dz >> 18) & 0x7f, ns = dz >> 25;
        if (nv > 0) {
            const bool have = lane * 8 < nv;
            const int4 s_ = lds128(lslot + so), e_ = lds128(lslot + so + L::ARR), f_ = lds128(lslot + so + 2 * L::ARR);
            uint32_t S[4] = {(uint32_t)s_.x, (uint32_t)s_.y, (uint32_t)s_.z, (uint32_t)s_.w};
            const uint32_t E[4] = {(uint32_t)e_.x, (uint32_t)e_.y, (uint32_t)e_.z, (uint32_t)e_.w};
            const uint32_t F[4] = {(uint32_t)f_.x, (uint32_t)f_.y, (uint32_t)f_.z, (uint32_t)f_.w};
            if (V && have) {                                            // value < v: the record matches no query
                const int4 v0 = lds128(vslot + so), v1 = lds128(vslot + so + 16);
                const int vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (vv[2 * k] < a.v) S[k] = (S[k] & 0xffff0000u) | 0x8000u;
                    if (vv[2 * k + 1] < a.v) S[k] = (S[k] & 0x0000ffffu) | 0x80000000u;
                }
            }
            uint32_t c[4] = {0, 0, 0, 0};
            // prune window of the group: smallest S (records are sorted by start: the first one; with -v the
            // smallest surviving one) and largest E
            uint32_t smin;
            if (V) {
                const uint32_t m2 = __vminu2(__vminu2(S[0], S[1]), __vminu2(S[2], S[3]));
                smin = __reduce_min_sync(FULL, have ? min(m2 & 0xffffu, m2 >> 16) : 0xffffu);
            } else {
                smin = __shfl_sync(FULL, S[0], 0) & 0xffffu;
            }
            const uint32_t x2 = __vmaxu2(__vmaxu2(E[0], E[1]), __vmaxu2(E[2], E[3]));
            const uint32_t emax = __reduce_max_sync(FULL, max(x2 & 0xffffu, x2 >> 16));
            // a segment can touch the group iff qe' > smallest start' and qs' < largest end':
            //   QE - smin >= 0x8000  and  emax - QS >= 0x8000
            const uint32_t qe_lo = smin + 0x8000u, qs_hi = emax - 0x8000u;     // QE >= qe_lo, QS <= qs_hi
            // segments lane and 32 + lane of the item; an absent one has qe' = 0 and survives no prune
            const uint32_t w0 = lane < ns ? lds32(qslot + so) : 0x80000000u;
            const uint32_t qs2_0 = __byte_perm(w0, 0, 0x1010), qe2_0 = __byte_perm(w0, 0, 0x3232);
            uint32_t w1 = 0x80000000u, qs2_1 = 0, qe2_1 = 0;
            if (ns > 32) {
                if (lane + 32 < ns) w1 = lds32(qslot + so + 128);
                qs2_1 = __byte_perm(w1, 0, 0x1010), qe2_1 = __byte_perm(w1, 0, 0x3232);
            }
            // The surviving segments are compacted into the warp's list (first-bin ones, then later-bin ones)
            // and streamed back with broadcast loads: 1 load + 16 ALU/FMA instructions per segment x 256 records.
            __syncwarp();                                               // the previous item's list has been read
            // first-bin segments [0, nf): every record of the tile is a candidate
            int nF = 0;
            if (nf > 0) {
                const unsigned f0 = __ballot_sync(FULL, (w0 >> 16) >= qe_lo && (w0 & 0xffffu) <= qs_hi && lane < nf);
                if (f0 >> lane & 1) sts64(list + 8u * __popc(f0 & lt), qs2_0, qe2_0);
                nF = __popc(f0);
                if (nf > 32) {
                    const unsigned f1 = __ballot_sync(FULL, (w1 >> 16) >= qe_lo && (w1 & 0xffffu) <= qs_hi && lane + 32 < nf);
                    if (f1 >> lane & 1) sts64(list + 8u * (nF + __popc(f1 & lt)), qs2_1, qe2_1);
                    nF += __popc(f1);
                }
            }
            // later-bin segments [nf, ns): only the records that start inside this bin count (index >= tS,
            // src/igd_search.c:510-511,522), i.e. start' >= 1.  Such a segment has qs' = 0 (QS = 1), so its
            // end test is vacuous; it is given X = S + 0x7fff in place of E, which turns that test into
            // bit 15 of (S + 0x7fff - 1), set  <=>  S >= 2  <=>  start' >= 1.
            int nL = 0;
            if (ns > nf && tsr < nv) {
                // prune with the first record that starts inside the bin (records are sorted by start)
                uint32_t slmin = smin;
                if (tsr > 0) {
                    if (V) {
                        slmin = max(smin, 2u);
                    } else {
                        const int nb4 = min(max(tsr - lane * 8, 0), 8);     // this lane's records before the bin
                        const uint32_t wd = nb4 < 2 ? S[0] : (nb4 < 4 ? S[1] : (nb4 < 6 ? S[2] : S[3]));
                        const uint32_t mine = (nb4 >= 8 || !have) ? 0xffffu : ((nb4 & 1) ? wd >> 16 : wd & 0xffffu);
                        slmin = __reduce_min_sync(FULL, mine);
                    }
                }
                const uint32_t qel = slmin + 0x8000u;
                if (nf < 32) {
                    const unsigned l0 = __ballot_sync(FULL, (w0 >> 16) >= qel && (w0 & 0xffffu) <= qs_hi && lane >= nf);
                    if (l0 >> lane & 1) sts64(list + 8u * (nF + __popc(l0 & lt)), qs2_0, qe2_0);
                    nL = __popc(l0);
                }
                if (ns > 32) {
                    const unsigned l1 = __ballot_sync(FULL, (w1 >> 16) >= qel && (w1 & 0xffffu) <= qs_hi && lane + 32 >= nf);
                    if (l1 >> lane & 1) sts64(list + 8u * (nF + nL + __popc(l1 & lt)), qs2_1, qe2_1);
                    nL += __popc(l1);
                }
            }
            __syncwarp();
            count_pairs(list, 0, nF, S, E, c);
            if (nL > 0) {
                const uint32_t X[4] = {S[0] + 0x7fff7fffu, S[1] + 0x7fff7fffu, S[2] + 0x7fff7fffu, S[3] + 0x7fff7fffu};
                count_pairs(list, nF, nF + nL, S, X, c);
            }
            // hits[file] += count.  One test per PAIR of records (neighbours in start order are mostly hit
            // together or not at all), then one shared-memory atomic per record.
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (have && c[k]) {
                    const uint32_t c0 = c[k] & 0xffffu, c1 = c[k] >> 16;
                    if (SMEM) {
                        const uint32_t a0 = hist_s + ((F[k] << 2) & 0x3fffcu), a1 = hist_s + ((F[k] >> 14) & 0x3fffcu);
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a0), "r"(c0) : "memory");
                        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a1), "r"(c1) : "memory");
                    } else {
                        const uint32_t f0 = F[k] & 0xffffu, f1 = F[k] >> 16;
                        if (c0) atomicAdd(&a.hits[f0], (unsigned long long)c0);
                        if (c1) atomicAdd(&a.hits[f1], (unsigned long long)c1);
                    }
                }
            }
        }
        if (j == 31) {                                                  // next quad-block of descriptors
            dcur = dnxt;
            dnxt = load_desc((n >> 5) + 2);
        }
    }
    cp_async_wait<0>();
    if (SMEM) {
        __syncthreads();
        for (int f = threadIdx.x; f < db.nfiles; f += blockDim.x) {
            const uint32_t c = hist[f];
            if (c) atomicAdd(&a.hits[f], (unsigned long long)c);
        }
    }
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

// ------------------------------------------------------------------------------------------------
// -m hitsmap: getMap / getMap_v (src/igd_search.c:772-886).  One warp per (tile, record j): the record
// is a query against its own tile; partners i in [tS, last start < end_j] with end_i > start_j, where
// tS is the tile's dedupe start only when start_j < bin boundary (:802-804).  -v is the STRICT `> v`.
// work items = padded record slots, so a flat grid-stride loop over P records needs no extra index.
// ------------------------------------------------------------------------------------------------
template <bool V>
__global__ void __launch_bounds__(256) k_hitsmap(DbView db, const int32_t *__restrict__ rec_tile, int64_t r0,
                                                  int64_t P, int32_t v, uint32_t *__restrict__ hitmap,
                                                  unsigned long long *__restrict__ nols)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    unsigned long long mine = 0;
    for (int64_t r = r0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < P; r += nw) {
        const int t = __ldg(&rec_tile[r]);
        if (t < 0) continue;                                 // pad slot
        const int4 tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
        const int64_t off = (int64_t)(uint32_t)tm.x * 8;
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
        const int64_t off = (int64_t)m.off8 * 8;
        const int np = (m.n + 7) & ~7;
        for (int i = lane; i < np; i += 32) rec_tile[off + i] = i < m.n ? (int32_t)t : -1;
    }
}

// ------------------------------------------------------------------------------------------------
// -s Seqpare (seq_overlaps + seqOverlaps, src/igd_search.c:253-451).
//   K4  k_seq_candidates: one thread per overlap of the -f hit list (same hit set and order as seq_overlaps:
//       :275-350).  Pair similarity in IEEE fp32 exactly as the reference computes it (:255, :291-299:
//       st / ((qlen + rlen) - st), round-to-nearest adds/subs/div, no FMA).  The candidate's sort key is
//       (problem = query-chrom group * nFiles + file) << 32 | ~bits(sm): ascending key = per (chrom, file)
//       problem, similarity descending.  The input order is (query rank, scan order), so a STABLE sort by that
//       key alone gives the reference's tie-break: smallest query rank, then earliest in scan order.
//   K5  k_seq_sweep: one thread per (chrom, file) problem walks its candidates in that order and accepts one
//       iff neither its row (query) nor its column is taken -- equivalent to the reference's repeated arg-max
//       with row retirement and column zeroing (:397-438).  A column is identified the way
//       the reference does it, quirk included: (n1 OF THE QUERY, index inside whichever tile the record was
//       found in) (:298, :342).  Taken rows/columns live in a per-problem open-addressing table.
//       The accepted similarities are written out in pick order; k_seq_sum adds them per file in the
//       reference's order (chrom groups ascending, picks in order) so the fp64 sums are bit-identical.
// ------------------------------------------------------------------------------------------------
__global__ void k_seq_candidates(DbView db, int64_t n, const int64_t *__restrict__ qoff,
                                 const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                 const int32_t *__restrict__ grp, int64_t total, const int32_t *__restrict__ os,
                                 const int32_t *__restrict__ oe, const int32_t *__restrict__ of,
                                 const int32_t *__restrict__ og, unsigned long long *__restrict__ key,
                                 uint32_t *__restrict__ val, uint32_t *__restrict__ row,
                                 unsigned long long *__restrict__ col)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int64_t lo = 0, hi = n;                                 // last q with qoff[q] <= p
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&qoff[mid]) <= p) lo = mid; else hi = mid;
    }
    const int32_t s = __ldg(&qs[lo]), e = __ldg(&qe[lo]);
    const int32_t rs = os[p], re = oe[p];
    const float qlen = __int2float_rn(e - s), rlen = __int2float_rn(re - rs);
    const float st = __int2float_rn(min(e, re) - max(s, rs));
    const float sm = __fdiv_rn(st, __fsub_rn(__fadd_rn(qlen, rlen), st));
    const unsigned long long prob = (unsigned long long)__ldg(&grp[lo]) * (unsigned)db.nfiles + (unsigned)of[p];
    key[p] = prob << 32 | (uint32_t)~__float_as_uint(sm);
    val[p] = (uint32_t)p;
    row[p] = (uint32_t)lo;
    col[p] = 1ull << 62 | (unsigned long long)(uint32_t)(s / db.nbp) << 32 | (uint32_t)og[p];
}

__device__ __forceinline__ int64_t lower_bound_key(const unsigned long long *__restrict__ k, int64_t n,
                                                   unsigned long long x)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&k[mid]) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// per problem: where its candidates start in the sorted array, and the size of its table
__global__ void k_seq_ranges(const unsigned long long *__restrict__ key, int64_t total, int64_t nprob, int32_t nfiles,
                             const int32_t *__restrict__ grp_nq, int64_t *__restrict__ pstart,
                             uint32_t *__restrict__ tsize)
{
    const int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid > nprob) return;
    const int64_t a = lower_bound_key(key, total, (unsigned long long)pid << 32);
    pstart[pid] = a;
    if (pid == nprob) return;
    const int64_t b = lower_bound_key(key, total, (unsigned long long)(pid + 1) << 32);
    // at most one row and one column are inserted per accepted candidate, and every accepted candidate retires
    // a distinct query of the group
    const int64_t picks = min(b - a, (int64_t)grp_nq[pid / nfiles]);
    uint32_t sz = 0;
    if (picks > 0) {
        sz = 16;
        while ((int64_t)sz < 4 * picks) sz <<= 1;
    }
    tsize[pid] = sz;
}

__device__ __forceinline__ uint32_t seq_hash(unsigned long long x)
{
    x ^= x >> 33, x *= 0xff51afd7ed558ccdULL, x ^= x >> 33, x *= 0xc4ceb9fe1a85ec53ULL, x ^= x >> 33;
    return (uint32_t)x;
}
// true if x was already in the table; inserts it otherwise when `insert`
__device__ __forceinline__ bool seq_probe(unsigned long long *t, uint32_t mask, unsigned long long x, bool insert)
{
    for (uint32_t h = seq_hash(x) & mask;; h = (h + 1) & mask) {
        const unsigned long long y = t[h];
        if (y == x) return true;
        if (y == 0) {
            if (insert) t[h] = x;
            return false;
        }
    }
}

__global__ void k_seq_sweep(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ val,
                            const uint32_t *__restrict__ row, const unsigned long long *__restrict__ col,
                            int64_t nprob, const int64_t *__restrict__ pstart, const uint32_t *__restrict__ tsize,
                            const unsigned long long *__restrict__ toff, unsigned long long *__restrict__ tables,
                            float *__restrict__ pick, uint32_t *__restrict__ npick)
{
    const int64_t pid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= nprob) return;
    const int64_t a = pstart[pid], b = pstart[pid + 1];
    uint32_t np = 0;
    if (b > a) {
        unsigned long long *t = tables + toff[pid];
        const uint32_t mask = tsize[pid] - 1;
        for (int64_t i = a; i < b; ++i) {
            const uint32_t bits = ~(uint32_t)key[i];
            if ((int32_t)bits <= 0) break;                  // similarity 0: the reference stops at maxf == 0
            const uint32_t c = val[i];
            const unsigned long long r = 1ull << 63 | row[c], cc = col[c];
            if (seq_probe(t, mask, r, false) || seq_probe(t, mask, cc, false)) continue;
            seq_probe(t, mask, r, true);
            seq_probe(t, mask, cc, true);
            pick[a + np++] = __uint_as_float(bits);
        }
    }
    npick[pid] = np;
}

// sm[m] += maxf in the reference's order (src/igd_search.c:367,397,419): chrom groups ascending, picks in order
__global__ void k_seq_sum(int32_t nfiles, int32_t ngroups, const int64_t *__restrict__ pstart,
                          const float *__restrict__ pick, const uint32_t *__restrict__ npick, double *__restrict__ acc)
{
    const int32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfiles) return;
    double s = 0.0;
    for (int32_t g = 0; g < ngroups; ++g) {
        const int64_t pid = (int64_t)g * nfiles + f, a = pstart[pid];
        const uint32_t np = npick[pid];
        for (uint32_t k = 0; k < np; ++k) s += (double)pick[a + k];
    }
    acc[f] = s;
}

// ------------------------------------------------------------------------------------------------
// Query ingest on the device: the BED text of a query file goes to HBM as it is and is parsed there.
// Restates the line loop of getOverlaps (src/igd_search.c:696-719: one line per '\n') and parse_bed
// (src/igd_base.c:53-72): fields split on '\t' (a NUL ends the line early), contig = field 0, start / end =
// atol of fields 1 / 2 narrowed to int32, accepted iff >= 3 fields, contig begins "chr" and is shorter than 40
// chars, end > 0; then get_id (src/igd_base.c:325-331): exact match of the contig name, -1 if absent.
// One thread per BYTE looks whether a line starts there (offset 0 or the byte after a '\n'); the thread that
// finds one parses it and appends (ctg, qs, qe) at an atomic cursor -- the order of the queries does not matter
// to hits[].  Rejected lines and unknown contigs produce nothing.
// ------------------------------------------------------------------------------------------------
struct CtgTable {
    const char *names;          // [nctg][40]
    const int32_t *slots;       // open addressing on FNV-1a of the name, -1 = empty
    uint32_t mask;
};

__host__ __device__ __forceinline__ uint32_t fnv1a(const char *s, int n)
{
    uint32_t h = 2166136261u;
    for (int i = 0; i < n; ++i) h = (h ^ (uint8_t)s[i]) * 16777619u;
    return h;
}

// glibc atol = strtol(base 10) on a field that ends at '\t', '\n', NUL or the end of the text:
// skip isspace, optional sign, digits, saturate at LONG_MIN / LONG_MAX; the caller narrows to int32.
__device__ __forceinline__ int32_t dev_atol(const char *__restrict__ t, int64_t p, int64_t end)
{
    auto at = [&](int64_t i) -> int { return i < end ? (int)(uint8_t)t[i] : 0; };
    int c = at(p);
    while (c == ' ' || c == '\v' || c == '\f' || c == '\r') c = at(++p);     // '\t' and '\n' end the field
    bool neg = false;
    if (c == '-' || c == '+') neg = c == '-', c = at(++p);
    unsigned long long v = 0;
    const unsigned long long lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    bool sat = false;
    while (c >= '0' && c <= '9') {
        const unsigned d = (unsigned)(c - '0');
        if (v > (lim - d) / 10) sat = true;
        v = v * 10 + d;
        c = at(++p);
    }
    if (sat) v = lim;
    const long long r = neg ? (long long)(0ull - v) : (long long)v;
    return (int32_t)r;
}

__global__ void __launch_bounds__(256) k_parse_bed(const char *__restrict__ text, int64_t len, CtgTable tab,
                                                    int32_t *__restrict__ ctg, int32_t *__restrict__ qs,
                                                    int32_t *__restrict__ qe, unsigned long long *__restrict__ n_out,
                                                    unsigned long long cap)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        if (i > 0 && text[i - 1] != '\n') continue;
        // field 0: the contig name
        int64_t p = i;
        int c;
        auto at = [&](int64_t k) -> int { return k < len ? (int)(uint8_t)text[k] : '\n'; };
        while ((c = at(p)) != '\t' && c != '\n' && c != 0) ++p;
        const int64_t nlen = p - i;
        if (c != '\t' || nlen < 3 || nlen >= 40) continue;                  // < 3 fields, or strlen(ctg) >= 40
        if (text[i] != 'c' || text[i + 1] != 'h' || text[i + 2] != 'r') continue;
        const int64_t f1 = p + 1;
        p = f1;
        while ((c = at(p)) != '\t' && c != '\n' && c != 0) ++p;
        if (c != '\t') continue;                                            // only 2 fields
        const int64_t f2 = p + 1;
        const int32_t st = dev_atol(text, f1, len), en = dev_atol(text, f2, len);
        if (!(en > 0)) continue;
        // get_id
        int32_t id = -1;
        for (uint32_t h = fnv1a(text + i, (int)nlen) & tab.mask;; h = (h + 1) & tab.mask) {
            const int32_t k = tab.slots[h];
            if (k < 0) break;
            const char *nm = tab.names + 40 * k;
            bool same = nm[nlen] == 0;
            for (int j = 0; same && j < (int)nlen; ++j) same = nm[j] == text[i + j];
            if (same) {
                id = k;
                break;
            }
        }
        if (id < 0) continue;                                               // unknown contig: contributes nothing
        const unsigned long long slot = atomicAdd(n_out, 1ull);
        if (slot < cap) ctg[slot] = id, qs[slot] = st, qe[slot] = en;
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
    std::vector<uint32_t> off8;             // [T] first padded record slot of a tile / 8 (host copy)
    TileMeta *tiles = nullptr;
    int64_t *d_toff = nullptr;
    int32_t *start = nullptr, *end = nullptr, *idx = nullptr, *value = nullptr, *tbase = nullptr;
    uint16_t *ps = nullptr, *pe = nullptr, *pf = nullptr;  // packed count-mode copies (NULL: M_ALL32 only)
    char *ctg_names = nullptr;              // [nctg][40] contig names (device), for the text ingest
    int32_t *ctg_slots = nullptr;           // hash table over them
    uint32_t ctg_mask = 0;
    void *arena = nullptr;                  // pinned host arena of the -f result lists
    size_t arena_cap = 0;
    char *d_text = nullptr;                 // query text staging
    size_t d_text_cap = 0;
    bool packed = false;
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
    // count mode (tile-centric): plan scratch, segments, work items
    uint32_t *cnt2 = nullptr, *cseg_off = nullptr, *cseg_mid = nullptr;   // [2T+2],[T+1],[T+1]
    PlanAgg *plan_agg = nullptr;                            // [plan_grid] per-CTA aggregates of k_plan
    uint32_t *plan_flag = nullptr, plan_epoch = 0;          // [plan_grid] = epoch of the batch that published
    int plan_grid = 1;
    unsigned long long *plan_tot = nullptr;                 // device {segments, items} of the batch in flight
    volatile unsigned long long *h_plan = nullptr;          // pinned copy of plan_tot
    cudaEvent_t plan_ev = nullptr, plan_done = nullptr;
    cudaEvent_t stage_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // around the 4 launches (diagnostics)
    bool stage_timing = false, stage_valid = false;
    cudaStream_t side = nullptr;                            // carries the plan readback
    CSeg *csegs = nullptr;
    int64_t cseg_cap = 0;
    int4 *items = nullptr;
    int64_t item_cap = 0;
    int64_t db_groups = 0;                                  // sum over tiles of ceil(n / GROUP)
    int count_grid[2][2] = {{0, 0}, {0, 0}};                // CTAs of the search kernel [packed][V] (0 = not set up)
    int64_t last_plan[4] = {0, 0, 0, 0};                    // segments, items, records, bytes of the last batch
    struct Pending {                                        // count batch enqueued but not yet settled
        bool active = false, mode_v = false;
        int mode = M_ALL32;
        int64_t n = 0;
        const int32_t *ctg = nullptr, *qs = nullptr, *qe = nullptr;
        int32_t v = 0;
        unsigned long long *hits = nullptr;
        cudaStream_t st = nullptr;
    } pending;
    DbView view() const
    {
        DbView v;
        v.tiles = tiles, v.start = start, v.end = end, v.idx = idx, v.value = value, v.tbase = tbase;
        v.ps = ps, v.pe = pe, v.pf = pf;
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

// ---- count mode ----------------------------------------------------------------------------------
// One batch = 7 launches on one stream with no host round trip in between: the kernels that consume the
// plan check on the device that it fits the scratch buffers and do nothing otherwise; the host looks at
// the plan's totals when it settles the batch and, in the rare case they did not fit (first batch of a
// larger size), grows the buffers and runs the batch again -- or splits it when the plan itself is too big.
// launch with programmatic stream serialization (see pdl_wait): only for kernels that call pdl_wait()
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*k)(KArgs...), unsigned grid, unsigned block, size_t shm, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = shm, cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, KArgs(std::forward<Args>(args))...);
}

template <bool V, bool PK>
int count_geometry(igdb_dev *d, int *grid, int *threads, size_t *shm, bool *smem)
{
    using LP = PackedSmem<V>;
    using L32 = CountSmem<V>;
    // the shared histogram is used while it leaves room for at least 8 warps' pipelines
    *smem = (PK ? LP::warps(d->nfiles, true) : L32::warps(d->nfiles, true)) >= 8;
    const int nw = PK ? LP::warps(d->nfiles, *smem) : L32::warps(d->nfiles, *smem);
    *shm = PK ? LP::cta_bytes(d->nfiles, *smem, nw) : L32::cta_bytes(d->nfiles, *smem, nw);
    *threads = nw * 32;
    if (d->count_grid[PK][V] == 0) {
        const void *k = PK ? (*smem ? (const void *)k_count_packed<V, true> : (const void *)k_count_packed<V, false>)
                           : (*smem ? (const void *)k_count_items<V, true> : (const void *)k_count_items<V, false>);
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*shm));
        d->count_grid[PK][V] = d->sm_count;
    }
    *grid = d->count_grid[PK][V];
    return 0;
}

// a CTA's shared histogram is uint32: an item adds at most GROUP*SMAX <= 2^15, a CTA takes at most
// items/grid + 32*COUNT_MAXW items, so this many items per launch cannot overflow a counter
int64_t item_limit(const igdb_dev *d, int grid) { return ((int64_t)(1 << 17) - 1024) * grid; }
// the kernels may touch up to 32 descriptors / 1 segment past the plan's totals (padding)
constexpr int64_t CAP_SLACK = 64;

template <bool V, int MODE>
int enqueue_count(igdb_dev *d, const igdb_dev::Pending &p)
{
    constexpr bool PK = MODE == M_PACKED;
    cudaStream_t st = p.st;
    const DbView db = d->view();
    int grid = 0, threads = 0;
    size_t shm = 0;
    bool smem = false;
    if (count_geometry<V, PK>(d, &grid, &threads, &shm, &smem)) return -1;
    if (grow(&d->csegs, &d->cseg_cap, std::max<int64_t>(2 * p.n + 2 * d->T, 4096), 1, 1)) return -1;
    if (grow(&d->items, &d->item_cap, std::min<int64_t>(d->db_groups + p.n / 64, item_limit(d, grid)) + 1024, 1, 1))
        return -1;
    const unsigned long long seg_cap = (unsigned long long)std::min<int64_t>(d->cseg_cap - CAP_SLACK, 0x7fffffff);
    const unsigned long long item_cap =
        (unsigned long long)std::min<int64_t>(d->item_cap - CAP_SLACK, item_limit(d, grid));
    const int thr = CSEG_T;
    const int nb = (int)((p.n + CSEG_T * CSEG_Q - 1) / (CSEG_T * CSEG_Q));
    const bool stages = d->stage_timing && MODE != M_ABN32;
    if (stages) CK(cudaEventRecord(d->stage_ev[0], st));
    CK(launch_pdl(k_cseg_count<V, MODE>, nb, thr, 0, st, db, p.n, p.ctg, p.qs, p.qe, d->cnt2));
    if (stages) CK(cudaEventRecord(d->stage_ev[1], st));
    // (the plan reads the later-bin counters that the scatter counts back down to zero)
    CK(launch_pdl(k_plan, d->plan_grid, SCAN_T, 0, st, db, d->cnt2, ++d->plan_epoch, d->plan_agg, d->plan_flag,
                  d->cseg_off, d->cseg_mid, d->plan_tot, seg_cap, item_cap, d->items));
    if (stages) CK(cudaEventRecord(d->stage_ev[2], st));
    CK(launch_pdl(k_cseg_scatter<V, MODE>, nb, thr, 0, st, db, p.n, p.ctg, p.qs, p.qe, d->cnt2, d->cseg_off,
                  d->cseg_mid, d->plan_tot, seg_cap, item_cap, (void *)d->csegs));
    // the plan's totals go back to the host on a side stream, so that the copy does not sit between the kernels
    CK(cudaEventRecord(d->plan_done, st));
    CK(cudaStreamWaitEvent(d->side, d->plan_done, 0));
    CK(cudaMemcpyAsync((void *)d->h_plan, d->plan_tot, 32, cudaMemcpyDeviceToHost, d->side));
    CK(cudaEventRecord(d->plan_ev, d->side));
    CountArgs a{};
    a.db = db, a.items = d->items, a.segs = d->csegs, a.tot = d->plan_tot;
    a.seg_cap = seg_cap, a.item_cap = item_cap, a.v = p.v, a.hits = p.hits;
    if (stages) CK(cudaEventRecord(d->stage_ev[3], st));
    if (MODE != M_ABN32) CK(cudaEventRecord(d->kev[0], st));
    if (PK) {
        if (smem) CK(launch_pdl(k_count_packed<V, true>, grid, threads, shm, st, a));
        else CK(launch_pdl(k_count_packed<V, false>, grid, threads, shm, st, a));
    } else {
        if (smem) k_count_items<V, true><<<grid, threads, shm, st>>>(a);
        else k_count_items<V, false><<<grid, threads, shm, st>>>(a);
    }
    if (MODE != M_ABN32) {
        CK(cudaEventRecord(d->kev[1], st));
        d->kev_valid = true;
    }
    if (stages) {
        CK(cudaEventRecord(d->stage_ev[4], st));
        d->stage_valid = true;
    }
    d->launches += 4;
    CK(cudaGetLastError());
    return 0;
}

int enqueue_count(igdb_dev *d, const igdb_dev::Pending &p)
{
    if (p.n <= 0) return 0;
    switch (p.mode) {
    case M_PACKED: return p.mode_v ? enqueue_count<true, M_PACKED>(d, p) : enqueue_count<false, M_PACKED>(d, p);
    case M_ABN32: return p.mode_v ? enqueue_count<true, M_ABN32>(d, p) : enqueue_count<false, M_ABN32>(d, p);
    default: return p.mode_v ? enqueue_count<true, M_ALL32>(d, p) : enqueue_count<false, M_ALL32>(d, p);
    }
}

// wait for the plan of the enqueued batch p; make sure it ran (see above)
int settle_count(igdb_dev *d, const igdb_dev::Pending &p)
{
    if (p.n <= 0) return 0;
    const bool pk = p.mode == M_PACKED;
    for (;;) {
        CK(cudaEventSynchronize(d->plan_ev));
        const unsigned long long nsegs = d->h_plan[0], nitems = d->h_plan[1], nrec = d->h_plan[2], nabn = d->h_plan[3];
        if (getenv("IGD_B200_DEBUG"))
            fprintf(stderr, "[igd_b200] count batch (mode %d): n=%lld segments=%llu items=%llu left=%llu (caps %lld / %lld)\n",
                    p.mode, (long long)p.n, nsegs, nitems, nabn, (long long)d->cseg_cap, (long long)d->item_cap);
        const int64_t lim = item_limit(d, d->count_grid[pk][p.mode_v]);
        if (nsegs <= (unsigned long long)std::min<int64_t>(d->cseg_cap - CAP_SLACK, 0x7fffffff) &&
            nitems <= (unsigned long long)std::min<int64_t>(d->item_cap - CAP_SLACK, lim)) {
            if (p.mode != M_ABN32) {
                // what the search kernel of this batch reads: descriptors, records, segments
                const int64_t rb = pk ? (p.mode_v ? 10 : 6) : (p.mode_v ? 16 : 12), sgb = pk ? 4 : 8;
                d->last_plan[0] = (int64_t)nsegs, d->last_plan[1] = (int64_t)nitems, d->last_plan[2] = (int64_t)nrec;
                d->last_plan[3] = 16 * (int64_t)nitems + rb * (int64_t)nrec + sgb * (int64_t)nsegs;
            }
            if (pk && nabn > 0) {           // the queries the packed layout cannot express: second pipeline
                igdb_dev::Pending q = p;
                q.mode = M_ABN32;
                if (enqueue_count(d, q) || settle_count(d, q)) return -1;
            }
            return 0;
        }
        // none of this batch's kernels wrote anything, but its segment counters were not counted back down
        CK(cudaStreamSynchronize(p.st));
        CK(cudaMemsetAsync(d->cnt2, 0, ((size_t)2 * d->T + 2) * 4, p.st));
        if (nsegs >= 0x7fffffffull || nitems > (unsigned long long)lim) {
            if (p.n <= 1) {
                snprintf(g_err, sizeof g_err, "a single query expands to too many (query, tile) segments");
                return -1;
            }
            igdb_dev::Pending a = p, b = p;
            a.n = p.n / 2, b.n = p.n - a.n;
            b.ctg += a.n, b.qs += a.n, b.qe += a.n;
            if (enqueue_count(d, a) || settle_count(d, a) || enqueue_count(d, b) || settle_count(d, b)) return -1;
            return 0;
        }
        if (grow(&d->csegs, &d->cseg_cap, (int64_t)nsegs + CAP_SLACK) ||
            grow(&d->items, &d->item_cap, (int64_t)nitems + CAP_SLACK))
            return -1;
        if (enqueue_count(d, p)) return -1;
    }
}

int finish_pending(igdb_dev *d)
{
    if (!d->pending.active) return 0;
    d->pending.active = false;
    return settle_count(d, d->pending);
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
                    d->seg_hits, d->seg_out_off, d->d_hits, d->cnt2, d->cseg_off, d->cseg_mid,
                    d->plan_agg, d->plan_flag, d->plan_tot, d->csegs, d->items, d->ps, d->pe, d->pf,
                    d->ctg_names, d->ctg_slots, d->d_text};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (d->pinned) cudaFreeHost(d->pinned);
    if (d->arena) cudaFreeHost(d->arena);
    if (d->h_scalar) cudaFreeHost(d->h_scalar);
    if (d->plan_ev) cudaEventDestroy(d->plan_ev);
    if (d->plan_done) cudaEventDestroy(d->plan_done);
    if (d->side) cudaStreamDestroy(d->side);
    for (auto &e : d->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : d->kev)
        if (e) cudaEventDestroy(e);
    for (auto &e : d->stage_ev)
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
    for (auto &e : d->stage_ev) CK(cudaEventCreate(&e));
    CK(cudaMallocHost((void **)&d->h_scalar, 64));
    d->h_plan = reinterpret_cast<volatile unsigned long long *>(d->h_scalar + 8);
    CK(cudaEventCreateWithFlags(&d->plan_ev, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->plan_done, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking));

    std::vector<int32_t> tbase(d->nctg + 1, 0);
    for (int c = 0; c < d->nctg; ++c) tbase[c + 1] = tbase[c] + ntile[c];
    d->T = tbase[d->nctg];
    std::vector<TileMeta> meta((size_t)d->T);
    d->toff.assign((size_t)d->T + 1, 0);
    int64_t p8 = 0;
    for (int c = 0, t = 0; c < d->nctg; ++c)
        for (int j = 0; j < ntile[c]; ++j, ++t) {
            const int32_t n = ncnt[t];
            meta[t].off8 = (uint32_t)p8;
            d->off8.push_back((uint32_t)p8);
            meta[t].n = n;
            meta[t].ts = d->nbp * j;            // bin boundary; k_tile_meta turns it into tS
            meta[t].s0 = 0x7fffffff;
            d->toff[t + 1] = d->toff[t] + n;
            d->db_groups += (n + GROUP - 1) / GROUP;
            p8 += (n + 7) / 8;
        }
    if (p8 >= (int64_t)0xffffffff) {
        snprintf(g_err, sizeof g_err, "database too large for 32-bit tile offsets");
        return -1;
    }
    d->nrec = d->toff[d->T];
    d->P = p8 * 8;
    if (dev_alloc(d, &d->tiles, (size_t)d->T) || dev_alloc(d, &d->d_toff, (size_t)d->T + 1) ||
        dev_alloc(d, &d->tbase, (size_t)d->nctg + 1) || dev_alloc(d, &d->start, (size_t)d->P + 8) ||
        dev_alloc(d, &d->end, (size_t)d->P + 8) || dev_alloc(d, &d->idx, (size_t)d->P + 8) ||
        dev_alloc(d, &d->value, (size_t)d->P + 8) || dev_alloc(d, &d->tile_cnt, (size_t)d->T + 1) ||
        dev_alloc(d, &d->tile_off, (size_t)d->T + 1) || dev_alloc(d, &d->d_hits, (size_t)d->nfiles) ||
        dev_alloc(d, &d->cnt2, (size_t)2 * d->T + 2) || dev_alloc(d, &d->cseg_off, (size_t)d->T + 1) ||
        dev_alloc(d, &d->cseg_mid, (size_t)d->T + 1) ||
        dev_alloc(d, &d->plan_agg, (size_t)4 * d->sm_count) || dev_alloc(d, &d->plan_flag, (size_t)4 * d->sm_count) ||
        dev_alloc(d, &d->plan_tot, 8))
        return -1;
    // packed count-mode copies: tile-relative coordinates need nbp + 1 < 2^15, file indices 16 bits
    d->packed = d->nbp <= 16384 && d->nfiles <= 65536 && !getenv("IGD_B200_NO_PACKED");
    if (d->packed && (dev_alloc(d, &d->ps, (size_t)d->P + 8) || dev_alloc(d, &d->pe, (size_t)d->P + 8) ||
                      dev_alloc(d, &d->pf, (size_t)d->P + 8)))
        return -1;
    CK(cudaMemcpy(d->tiles, meta.data(), meta.size() * sizeof(TileMeta), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_toff, d->toff.data(), d->toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->tbase, tbase.data(), tbase.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    // pad records (tiles are padded to a multiple of 4): start = INT_MAX, end = INT_MIN -- overlap no query
    const unsigned fb = (unsigned)std::min<int64_t>(((int64_t)d->P + 8 + 255) / 256, 148 * 32);
    k_fill<<<fb, 256>>>(d->start, (int64_t)d->P + 8, IMAX);
    k_fill<<<fb, 256>>>(d->end, (int64_t)d->P + 8, IMIN);
    k_fill<<<fb, 256>>>(d->value, (int64_t)d->P + 8, IMIN);
    if (d->packed) {
        k_fill16<<<fb, 256>>>(d->ps, (int64_t)d->P + 8, (uint16_t)0x8000);
        k_fill16<<<fb, 256>>>(d->pe, (int64_t)d->P + 8, (uint16_t)0x8000);
        CK(cudaMemset(d->pf, 0, ((size_t)d->P + 8) * 2));
    }
    CK(cudaGetLastError());
    CK(cudaMemset(d->idx, 0, ((size_t)d->P + 8) * 4));
    CK(cudaMemset(d->tile_cnt, 0, ((size_t)d->T + 1) * 4));
    CK(cudaMemset(d->cnt2, 0, ((size_t)2 * d->T + 2) * 4));
    CK(cudaMemset(d->plan_flag, 0, (size_t)4 * d->sm_count * 4));
    // k_plan's CTAs wait for each other, so all of them must be resident at once: never more than the
    // occupancy calculator allows (and at most 4 per SM, the size of the flag arrays)
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_plan, SCAN_T, 0));
    per_sm = std::max(1, std::min(per_sm, 4));
    d->plan_grid = std::max(1, std::min(per_sm * d->sm_count, (d->T + SCAN_T - 1) / SCAN_T));
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

void *igdb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    memset(p, 0, bytes);
    return p;
}
void igdb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

void *igdb_dev_pinned(igdb_dev *d, size_t bytes)
{
    if (bytes > d->pinned_cap) {
        if (d->pinned) cudaFreeHost(d->pinned);
    if (d->arena) cudaFreeHost(d->arena);
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
        (const int4 *)d->d_stage, rec0, nrec, d->d_toff, d->T, d->tiles, d->start, d->end, d->idx, d->value, d->nbp,
        d->ps, d->pe, d->pf);
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

void igdb_dev_stage_timing(igdb_dev *d, int on) { d->stage_timing = on != 0, d->stage_valid = false; }
int igdb_dev_last_stages(igdb_dev *d, float out[4])
{
    if (!d->stage_valid || cudaEventSynchronize(d->stage_ev[4]) != cudaSuccess) return -1;
    for (int k = 0; k < 4; ++k)
        if (cudaEventElapsedTime(&out[k], d->stage_ev[k], d->stage_ev[k + 1]) != cudaSuccess) return -1;
    return 0;
}

int64_t igdb_dev_bytes(const igdb_dev *d) { return d->bytes; }
void igdb_dev_last_plan(const igdb_dev *d, int64_t out[4])
{
    for (int k = 0; k < 4; ++k) out[k] = d->last_plan[k];
}
int igdb_dev_packed(const igdb_dev *d) { return d->packed ? 1 : 0; }
int64_t igdb_dev_launches(const igdb_dev *d) { return d->launches; }

int igdb_dev_count_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                          int mode_v, int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    if (n <= 0) return 0;
    if (d->T == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    const int64_t BATCH = 1 << 26;
    for (int64_t a = 0; a < n; a += BATCH) {
        igdb_dev::Pending p;
        p.mode_v = mode_v != 0, p.n = std::min(BATCH, n - a);
        p.ctg = d_ctg + a, p.qs = d_qs + a, p.qe = d_qe + a, p.v = v;
        p.hits = (unsigned long long *)d_hits, p.st = st;
        p.mode = d->packed ? M_PACKED : M_ALL32;
        if (enqueue_count(d, p)) return -1;
        if (a + BATCH < n) {                        // the scratch buffers are reused by the next chunk
            if (settle_count(d, p)) return -1;
        } else {
            p.active = true;
            d->pending = p;
        }
    }
    return 0;
}

int igdb_dev_count_finish(igdb_dev *d)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    return finish_pending(d);
}

int igdb_dev_count_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe, int mode_v,
                        int32_t v, uint64_t *hits_delta, float *t_ms)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    cudaStream_t st = d->stream;
    CK(cudaMemsetAsync(d->d_hits, 0, (size_t)d->nfiles * 8, st));
    if (t_ms) t_ms[0] = t_ms[1] = t_ms[2] = 0.f;
    const int64_t BATCH = 1 << 26;
    for (int64_t a = 0; a < n && d->T > 0; a += BATCH) {
        const int64_t m = std::min(BATCH, n - a);
        if (ensure_query_cap(d, m)) return -1;
        CK(cudaEventRecord(d->ev[0], st));
        CK(cudaMemcpyAsync(d->q_ctg, ctg + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->q_qs, qs + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->q_qe, qe + a, (size_t)m * 4, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(d->ev[1], st));
        igdb_dev::Pending p;
        p.mode_v = mode_v != 0, p.n = m, p.ctg = d->q_ctg, p.qs = d->q_qs, p.qe = d->q_qe, p.v = v;
        p.hits = d->d_hits, p.st = st;
        p.mode = d->packed ? M_PACKED : M_ALL32;
        if (enqueue_count(d, p)) return -1;
        const bool last = a + BATCH >= n;
        if (last) {                                 // the result copy rides behind the kernels; no sync yet
            CK(cudaEventRecord(d->ev[2], st));
            CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(d->ev[3], st));
        }
        // in the rare case the batch had to be run again, the copy is repeated below
        const int64_t launches0 = d->launches;
        if (settle_count(d, p)) return -1;
        if (last && d->launches != launches0) {
            CK(cudaEventRecord(d->ev[2], st));
            CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(d->ev[3], st));
        }
        if (t_ms || !last) CK(cudaStreamSynchronize(st));
        if (t_ms) {
            float x = 0, y = 0;
            CK(cudaEventElapsedTime(&x, d->ev[0], d->ev[1]));
            t_ms[0] += x;
            if (last) {
                CK(cudaEventElapsedTime(&y, d->ev[1], d->ev[2]));
                t_ms[1] += y;
            }
        }
    }
    if (n <= 0 || d->T == 0) {
        CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    CK(cudaStreamSynchronize(st));
    if (t_ms) CK(cudaEventElapsedTime(&t_ms[2], d->ev[2], d->ev[3]));
    return 0;
}

int igdb_dev_set_contigs(igdb_dev *d, const char *names40)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    uint32_t sz = 16;
    while (sz < 4u * (uint32_t)std::max(d->nctg, 1)) sz <<= 1;
    std::vector<int32_t> slots(sz, -1);
    for (int k = 0; k < d->nctg; ++k) {
        const char *nm = names40 + 40 * (size_t)k;
        uint32_t h = fnv1a(nm, (int)strnlen(nm, 39)) & (sz - 1);
        bool dup = false;                           // get_id returns the first contig of that name
        while (slots[h] >= 0) {
            if (strncmp(names40 + 40 * (size_t)slots[h], nm, 40) == 0) { dup = true; break; }
            h = (h + 1) & (sz - 1);
        }
        if (!dup) slots[h] = k;
    }
    if (dev_alloc(d, &d->ctg_names, 40 * (size_t)std::max(d->nctg, 1)) || dev_alloc(d, &d->ctg_slots, sz)) return -1;
    CK(cudaMemcpy(d->ctg_names, names40, 40 * (size_t)d->nctg, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->ctg_slots, slots.data(), sz * sizeof(int32_t), cudaMemcpyHostToDevice));
    d->ctg_mask = sz - 1;
    return 0;
}

int igdb_dev_hits_begin(igdb_dev *d)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    CK(cudaMemsetAsync(d->d_hits, 0, (size_t)d->nfiles * 8, d->stream));
    return 0;
}

int igdb_dev_hits_end(igdb_dev *d, uint64_t *hits_delta)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    return 0;
}

// One chunk of query text (whole lines): H2D, parse on the device, count into d_hits (between hits_begin and
// hits_end).  The search of this chunk is still running when the call returns: the caller reads the next chunk
// of the file meanwhile.  `text` must stay untouched until the next call on this handle has returned.
int igdb_dev_count_text(igdb_dev *d, const char *text, size_t len, int mode_v, int32_t v, int64_t *n_queries)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    if (n_queries) *n_queries = 0;
    if (len == 0 || d->T == 0) return 0;
    cudaStream_t st = d->stream;
    if (len > d->d_text_cap) {
        if (d->d_text) CK(cudaFree(d->d_text));
        d->d_text = nullptr;
        CK(cudaMalloc((void **)&d->d_text, len + (len >> 2) + 64));
        d->d_text_cap = len + (len >> 2);
    }
    // an accepted line has at least 8 characters ("chr" + tab + digit + tab + digit + newline)
    const int64_t cap = (int64_t)(len / 8 + 1);
    if (ensure_query_cap(d, cap)) return -1;
    CK(cudaMemcpyAsync(d->d_text, text, len, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d->plan_tot + 4, 0, 8, st));
    const CtgTable tab{d->ctg_names, d->ctg_slots, d->ctg_mask};
    const unsigned grid = (unsigned)std::min<size_t>((len + 255) / 256, (size_t)d->sm_count * 64);
    k_parse_bed<<<grid, 256, 0, st>>>(d->d_text, (int64_t)len, tab, d->q_ctg, d->q_qs, d->q_qe, d->plan_tot + 4,
                                      (unsigned long long)cap);
    d->launches += 1;
    CK(cudaGetLastError());
    unsigned long long *h_n = (unsigned long long *)(d->h_scalar + 4);
    CK(cudaMemcpyAsync(h_n, d->plan_tot + 4, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t n = (int64_t)std::min<unsigned long long>(*h_n, (unsigned long long)cap);
    if (n_queries) *n_queries = n;
    if (n == 0) return 0;
    igdb_dev::Pending p;
    p.mode_v = mode_v != 0, p.n = n, p.ctg = d->q_ctg, p.qs = d->q_qs, p.qe = d->q_qe, p.v = v;
    p.hits = d->d_hits, p.st = st;
    p.mode = d->packed ? M_PACKED : M_ALL32;
    if (enqueue_count(d, p)) return -1;
    p.active = true;
    d->pending = p;
    return 0;
}

}  // extern "C"

namespace {

struct DevBuf {                             // cudaFree on scope exit
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T>
    T *as() const { return (T *)p; }
    int alloc(size_t bytes)
    {
        CK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        return 0;
    }
};

// -f core: the queries are in d->q_ctg/q_qs/q_qe (device).  Leaves query_off[n+1] (overlaps of query q are
// [query_off[q], query_off[q+1])) in qoff and the overlaps, SoA, in out: start[total] | end[total] | file[total]
// (| index within the tile [total] if with_g), in the reference's order.  Returns total.
int64_t full_device(igdb_dev *d, int64_t n, cudaStream_t st, bool with_g, DevBuf &qoff, DevBuf &out)
{
    if (qoff.alloc(((size_t)n + 1) * 8)) return -1;
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
        if (cudaMalloc(&out.p, (size_t)total * (with_g ? 16 : 12)) != cudaSuccess) {
            out.p = nullptr;
            cudaGetLastError();
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

int upload_queries(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe, cudaStream_t st)
{
    if (ensure_query_cap(d, std::max<int64_t>(n, 1))) return -1;
    CK(cudaMemcpyAsync(d->q_ctg, ctg, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qs, qs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qe, qe, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace

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

// -s: queries grouped by query-file chrom (grp[q] ascending, 0..ngroups-1) and sorted by start inside a group
// (the host did readBED + qsort, src/igd_base.c:628-649, src/igd_search.c:373).  acc[m] receives the sum of the
// matched similarities of file m, i.e. sm[m] before the final ratio of src/igd_search.c:445-448.
int igdb_dev_seqpare(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                     const int32_t *grp, int32_t ngroups, const int32_t *grp_nq, double *acc)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    cudaStream_t st = d->stream;
    for (int32_t m = 0; m < d->nfiles; ++m) acc[m] = 0.0;
    if (n == 0 || d->T == 0 || d->nfiles == 0) return 0;
    const int64_t nprob = (int64_t)ngroups * d->nfiles;
    if (n >= (1 << 26) || nprob >= (1ll << 31)) {
        snprintf(g_err, sizeof g_err, "Seqpare: query set too large (%lld queries, %lld chrom x file problems)",
                 (long long)n, (long long)nprob);
        return -1;
    }
    if (upload_queries(d, n, ctg, qs, qe, st)) return -1;
    DevBuf qoff, out;
    const int64_t total = full_device(d, n, st, true, qoff, out);
    if (total < 0) return -1;
    if (total == 0) return 0;
    if (total >= (1ll << 31)) {
        snprintf(g_err, sizeof g_err, "Seqpare: %lld candidate pairs exceed one pass (2^31)", (long long)total);
        return -1;
    }
    DevBuf b_grp, b_gnq, b_key, b_key2, b_val, b_val2, b_row, b_col, b_tmp;
    if (b_grp.alloc((size_t)n * 4) || b_gnq.alloc((size_t)ngroups * 4) || b_key.alloc((size_t)total * 8) ||
        b_key2.alloc((size_t)total * 8) || b_val.alloc((size_t)total * 4) || b_val2.alloc((size_t)total * 4) ||
        b_row.alloc((size_t)total * 4) || b_col.alloc((size_t)total * 8))
        return -1;
    CK(cudaMemcpyAsync(b_grp.p, grp, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_gnq.p, grp_nq, (size_t)ngroups * 4, cudaMemcpyHostToDevice, st));
    const int32_t *o = out.as<int32_t>();
    k_seq_candidates<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        d->view(), n, qoff.as<int64_t>(), d->q_qs, d->q_qe, b_grp.as<int32_t>(), total, o, o + total, o + 2 * total,
        o + 3 * total, b_key.as<unsigned long long>(), b_val.as<uint32_t>(), b_row.as<uint32_t>(),
        b_col.as<unsigned long long>());
    CK(cudaGetLastError());
    // stable LSD radix sort (CUB, a library primitive) on the used key bits only
    int end_bit = 32;
    while (end_bit < 64 && (nprob >> (end_bit - 32)) != 0) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, b_key.as<unsigned long long>(),
                                       b_key2.as<unsigned long long>(), b_val.as<uint32_t>(), b_val2.as<uint32_t>(),
                                       (int)total, 0, end_bit, st));
    if (b_tmp.alloc(tmp_bytes)) return -1;
    CK(cub::DeviceRadixSort::SortPairs(b_tmp.p, tmp_bytes, b_key.as<unsigned long long>(),
                                       b_key2.as<unsigned long long>(), b_val.as<uint32_t>(), b_val2.as<uint32_t>(),
                                       (int)total, 0, end_bit, st));
    DevBuf b_pstart, b_tsize, b_toff, b_tab, b_pick, b_npick, b_acc;
    if (b_pstart.alloc((size_t)(nprob + 1) * 8) || b_tsize.alloc((size_t)(nprob + 1) * 4) ||
        b_toff.alloc((size_t)(nprob + 1) * 8) || b_pick.alloc((size_t)total * 4) ||
        b_npick.alloc((size_t)nprob * 4) || b_acc.alloc((size_t)d->nfiles * 8))
        return -1;
    k_seq_ranges<<<(unsigned)((nprob + 1 + 255) / 256), 256, 0, st>>>(
        b_key2.as<unsigned long long>(), total, nprob, d->nfiles, b_gnq.as<int32_t>(), b_pstart.as<int64_t>(),
        b_tsize.as<uint32_t>());
    CK(cudaGetLastError());
    if (exclusive_scan<uint32_t, unsigned long long>(d, b_tsize.as<uint32_t>(), b_toff.as<unsigned long long>(), nprob,
                                                     st))
        return -1;
    unsigned long long *h_total = (unsigned long long *)(d->h_scalar + 2);
    CK(cudaMemcpyAsync(h_total, b_toff.as<unsigned long long>() + nprob, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const size_t slots = (size_t)*h_total;
    if (b_tab.alloc(slots * 8)) {
        snprintf(g_err, sizeof g_err, "Seqpare: matching tables (%zu slots) do not fit in device memory", slots);
        return -1;
    }
    CK(cudaMemsetAsync(b_tab.p, 0, std::max<size_t>(slots * 8, 16), st));
    k_seq_sweep<<<(unsigned)((nprob + 127) / 128), 128, 0, st>>>(
        b_key2.as<unsigned long long>(), b_val2.as<uint32_t>(), b_row.as<uint32_t>(), b_col.as<unsigned long long>(),
        nprob, b_pstart.as<int64_t>(), b_tsize.as<uint32_t>(), b_toff.as<unsigned long long>(),
        b_tab.as<unsigned long long>(), b_pick.as<float>(), b_npick.as<uint32_t>());
    k_seq_sum<<<(d->nfiles + 127) / 128, 128, 0, st>>>(d->nfiles, ngroups, b_pstart.as<int64_t>(), b_pick.as<float>(),
                                                       b_npick.as<uint32_t>(), b_acc.as<double>());
    d->launches += 5;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(acc, b_acc.p, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

int64_t igdb_dev_hitsmap_host(igdb_dev *d, int32_t v, int64_t t0, int64_t t1, uint32_t *hitmap)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    t0 = std::max<int64_t>(t0, 0), t1 = std::min<int64_t>(t1, d->T);
    if (t0 >= t1) return 0;
    // padded record slots of the tile range [t0, t1)
    const int64_t r0 = (int64_t)d->off8[t0] * 8, r1 = t1 < d->T ? (int64_t)d->off8[t1] * 8 : d->P;
    const size_t cells = (size_t)d->nfiles * d->nfiles;
    uint32_t *d_map = nullptr;
    int32_t *d_rt = nullptr;
    unsigned long long *d_n = nullptr;
    CK(cudaMalloc((void **)&d_map, std::max<size_t>(cells, 1) * 4));
    CK(cudaMalloc((void **)&d_rt, ((size_t)d->P + 8) * 4));
    CK(cudaMalloc((void **)&d_n, 8));
    CK(cudaMemcpyAsync(d_map, hitmap, cells * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_n, 0, 8, st));
    const int grid = d->sm_count * 8;
    k_rec_tile<<<grid, 256, 0, st>>>(d->tiles, d->T, d_rt);
    if (v > 0) k_hitsmap<true><<<grid, 256, 0, st>>>(d->view(), d_rt, r0, r1, v, d_map, d_n);
    else k_hitsmap<false><<<grid, 256, 0, st>>>(d->view(), d_rt, r0, r1, v, d_map, d_n);
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
