// igd_b200 CUDA layer, shared definitions (internal to libigd_b200.so): HBM image of an .igd database and the search kernels (sm_100a).
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
//   rb[4(P+8)], rv[2(P+8)]  the count-mode copy, uint16 (when nbp <= 16384 and nFiles <= 65535): every record
//                       appears TWICE, as a START entry and as an END entry, each {key, file} (+ value in rv):
//                         START half: key = start' + 1, the tile in START order (the file's)
//                         END half (from entry P+8 on): key = end', the tile sorted by END
//                       with coordinates RELATIVE to the tile's bin boundary bd and clamped to the bin,
//                       x' = clamp(x - bd + 1, 0, nbp + 1) -- exact for every comparison a query that visits the
//                       tile can make.  The overlap count of a record against a tile's segments separates into a
//                       function of its start alone minus a function of its end alone (see k_count_rank), so the
//                       two halves are searched independently, each a merge of two sorted sequences.
//                       A tile's entries are stored in GROUPS of 256 (the last one: what is left, rounded up to 8),
//                       each one block of keys followed by its file indices -- one contiguous read per work item.
//                       A group is a contiguous slice of the tile's sorted sequence, but INSIDE a group the entries
//                       are ordered by (file mod 32), the shared-memory bank of their histogram counter: the k-th
//                       entries of the 32 lanes then fall on (nearly) distinct banks (k_group_banksort; the
//                       group's smallest and largest key are kept in gk[]).
//                       rv = clamp(value, 0, 32767) + 0x8000 per entry (exact for the -v thresholds 1 .. 32767;
//                       any other threshold runs on the int32 arrays).
//                       Pads: START key 0x8000, END key 0x7fff (past every real key), file = nFiles (a trash slot).
//                       With nFiles + 1 <= 16383 the file fields hold 4 * file, the byte offset of hist[file] (fshift).
//   tbase[nctg+1]     global tile index of each contig's tile 0.
//
// Query batches are turned into (query, tile) SEGMENTS -- one per non-empty, non-gated tile the reference
// would visit for that query, first-bin-empty quirk included -- counting-sorted by tile on the device, so
// that the work is organised by tile whatever the order of the input.
//
// Count mode (-q, -v) is TILE-CENTRIC and RECORD-STATIONARY: a warp task ("item") is up to 256 consecutive
// entries of one tile against up to 64 of the tile's segments.  Every lane keeps 8 entries in registers (one
// 128-bit load per array -- the tile is read from HBM once however many queries hit it) and adds, per entry, the
// number of segments on the far side of its key; ONE shared-memory atomic per entry (hits[file] += count), not
// one per overlap.  -f keeps the query-centric warp-per-segment kernel because it needs the reference's
// per-query output order.
#ifndef IGD_DEV_IMPL_CUH
#define IGD_DEV_IMPL_CUH
#include "igd_dev.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

// The CUDA layer is split by function: igd_dev_image.cu (HBM image of the database: create / upload / destroy),
// igd_dev_count.cu (count mode: the hot path), igd_dev_full.cu (-f), igd_dev_seqpare.cu (-s), igd_dev_hitsmap.cu
// (-m), igd_dev_ingest.cu (query text parsed on the device).  Kernels live in anonymous namespaces of their
// files; what more than one file needs is here.
namespace igdb {

extern thread_local char g_err[512];           // message of the last failure on this thread (igd_dev_image.cu)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(igdb::g_err, sizeof igdb::g_err, "%s failed: %s (%s:%d)", #call,              \
                     cudaGetErrorString(e_), __FILE__, __LINE__);                                  \
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
    const uint16_t *rb, *rv;                // count-mode copy: {keys, files} blocks and values; START half then END half (may be NULL)
    uint32_t eoff8;                         // first entry of the END half, in units of 8 entries
    const uint32_t *gbase, *gk;             // first group of each tile; per group smallest | largest << 16 real key (END half: + ngroups)
    uint32_t ngroups;
    int32_t fshift;                         // rb's file fields are file << fshift
    const int32_t *tbase;
    int32_t nctg, nbp, nfiles, ntiles;
};

// segment: one (query, tile) unit of work. w = sid | later<<31
struct __align__(16) Seg {
    int32_t qs, qe;
    uint32_t tile, w;
};

constexpr unsigned FULL = 0xffffffffu;
constexpr int IMAX = 0x7fffffff, IMIN = -0x7fffffff - 1;

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
//   M_PACKED the queries with 0 <= qs < qe, 16-bit START / END entries (k_count_rank); the others are only counted
//            (cnt2[2T]) and, if there are any, a second pipeline in mode M_ABN32 takes exactly those
//   M_ABN32  the queries with qs < 0 or qs >= qe (accepted by parse_bed, src/igd_base.c:69: only en > 0 is
//            required), int32 records
enum { M_ALL32 = 0, M_PACKED = 1, M_ABN32 = 2 };

// per-CTA aggregate of the plan scan (k_plan)
struct PlanAgg {
    unsigned long long segs, items, recs;
};

// 32-ary ballot search shared by -f and -m: count of start < key in a sorted tile
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


}  // namespace igdb

using namespace igdb;

// ================================================================================================
// the device image of one database and its scratch state
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
    uint16_t *rb = nullptr, *rv = nullptr;  // count-mode copy, START half + END half (NULL: M_ALL32 only)
    uint32_t *gbase = nullptr, *gk = nullptr;   // [T] first group of a tile, [2 db_groups] key range of each group
    char *ctg_names = nullptr;              // [nctg][40] contig names (device), for the text ingest
    int32_t *ctg_slots = nullptr;           // hash table over them
    uint32_t ctg_mask = 0;
    void *arena = nullptr;                  // pinned host arena of the -f result lists
    size_t arena_cap = 0;
    void *dres_qoff = nullptr, *dres_out = nullptr;     // device-resident -f result of the last igdb_dev_full_device (in the pools)
    void *fp_qoff = nullptr, *fp_out = nullptr;         // -f: pooled query_off / list buffers (grown to the high-water mark)
    size_t fp_qoff_cap = 0, fp_out_cap = 0;
    char *d_text = nullptr;                 // query text staging
    size_t d_text_cap = 0;
    uint32_t *d_starts = nullptr;                           // ingest: offsets of the candidate lines of a chunk
    bool packed = false;
    bool bulk = false;                      // the search kernel stages an item's entries by one bulk copy (TMA) instead of cp.async
    int fshift = 0;                         // the count-mode copy stores file << fshift (2: the byte offset of hist[file])
    int64_t bytes = 0;
    int64_t launches = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[2] = {nullptr, nullptr};               // around the last count-mode search kernel
    bool kev_valid = false;
    // staging / scratch (grown on demand)
    void *pinned = nullptr;
    size_t pinned_cap = 0;
    void *up_stage[2] = {nullptr, nullptr};                 // upload: two device staging slots, filled alternately
    size_t up_cap[2] = {0, 0};
    cudaEvent_t up_ev[2] = {nullptr, nullptr};              // slot k's transpose has read its host and staging buffer
    bool up_busy[2] = {false, false};
    uint32_t *bad_idx = nullptr;                            // device flag: a record's file index was outside [0, nFiles)
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
    int plan_grid = 0, plan_grid_max = 0;
    unsigned long long *plan_tot = nullptr;                 // device {segments, items} of the batch in flight
    volatile unsigned long long *h_plan = nullptr;          // pinned copies of plan_tot, 4 words per slot (Pending::slot)
    cudaEvent_t plan_ev[2] = {nullptr, nullptr}, plan_done = nullptr;   // plan_ev[slot]: that slot's copy has landed
    cudaEvent_t last_plan_ev = nullptr;                     // the plan readback enqueued last (the next k_plan waits for it)
    cudaEvent_t stage_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // around the 4 launches (diagnostics)
    bool stage_timing = false, stage_valid = false;
    cudaStream_t side = nullptr;                            // carries the plan readback
    CSeg *csegs = nullptr;
    int64_t cseg_cap = 0;
    int4 *items = nullptr;
    int64_t item_cap = 0;
    int4 *items2 = nullptr;                                 // k_count_rank's descriptors (2 x int4 per item), k_windows
    int32_t *hm_rt = nullptr, *hm_pmax = nullptr;           // -m: per record slot its tile / the prefix maximum of end (made once)
    uint32_t *hm_half = nullptr, *hm_map = nullptr;         // -m: half matrix (zero between calls), result matrix of the host entry
    size_t hm_map_cells = 0;
    uint32_t *hm_frec = nullptr;                            // -m: the record slots grouped by file
    unsigned long long *hm_fstart = nullptr;                // -m: [nfiles + 1] where each file's slots start in hm_frec
    bool hm_rows = false;                                   // -m: a file's row of counters fits in shared memory
    unsigned long long *hm_n = nullptr;                     // -m: pairs counted by the last call
    void *seq_pool = nullptr, *seq_tab = nullptr;           // -s: the temporaries of a call, the matching tables
    size_t seq_pool_cap = 0, seq_tab_cap = 0;
    unsigned long long *shard_scratch = nullptr;            // igd_dev_shard.cu: per-destination counters
    uint32_t *qrank = nullptr;                              // [n] per query: rank of its first-bin segment (k_cseg_count)
    int64_t qrank_cap = 0;
    int4 *tplan = nullptr;                                  // [2T] per tile {first segment, first later-bin one, end, first item}, {off8, n, ts, first group}
    int64_t item2_cap = 0;
    int64_t db_groups = 0;                                  // sum over tiles of ceil(n / GROUP)
    int64_t max_tile_groups = 0;                            // the most groups any tile has
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
        int32_t tile_lo = 0, tile_hi = 0x7fffffff;          // no query touches a tile outside [tile_lo, tile_hi)
        int slot = 0;                                       // which h_plan / plan_ev carries this batch's plan back
        const unsigned long long *dyn = nullptr;            // device {queries, tile_lo, tile_hi}: n and the tile range above are upper bounds then
    } pending;
    // The pipelined host entry (igdb_dev_count_submit / _wait): two batches in flight.  A slot owns what must outlive
    // the next batch's copy -- its queries in HBM, its hits[] on both sides, its events; the plan scratch is shared,
    // because the batches run one after the other on d->stream and only the H2D of the next one (d->copy) overlaps.
    struct CountSlot {
        bool active = false;
        uint64_t seq = 0;                                   // order of submission
        Pending p;
        int32_t *q = nullptr;                               // ctg[cap] | qs[cap] | qe[cap]
        int64_t q_cap = 0;
        unsigned long long *d_hits = nullptr, *h_hits = nullptr;    // [nfiles] device / pinned host
        cudaEvent_t h2d = nullptr, done = nullptr;
    } slot[2];
    cudaStream_t copy = nullptr;                            // carries the H2D of submitted batches
    uint64_t submit_seq = 0, misfits = 0;                   // misfits: batches whose plan did not fit (settle_count)
    DbView view() const
    {
        DbView v;
        v.tiles = tiles, v.start = start, v.end = end, v.idx = idx, v.value = value, v.tbase = tbase;
        v.rb = rb, v.rv = rv, v.eoff8 = (uint32_t)((P + 8) / 8), v.gbase = gbase, v.gk = gk, v.ngroups = (uint32_t)db_groups, v.fshift = fshift;
        v.nctg = nctg, v.nbp = nbp, v.nfiles = nfiles, v.ntiles = T;
        return v;
    }
};

namespace igdb {

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

struct DevBuf {                             // cudaFree on scope exit (unless the memory belongs to a pool of the handle)
    void *p = nullptr;
    bool owned = true;
    ~DevBuf() { if (p && owned) cudaFree(p); }
    template <class T>
    T *as() const { return (T *)p; }
    int alloc(size_t bytes)
    {
        CK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        return 0;
    }
};


// defined in one file, used by several
int ensure_query_cap(igdb_dev *d, int64_t n);                                      // igd_dev_image.cu
int upload_queries(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                   cudaStream_t st);                                               // igd_dev_image.cu
int enqueue_count(igdb_dev *d, const igdb_dev::Pending &p);                        // igd_dev_count.cu
int finish_pending(igdb_dev *d);                                                   // igd_dev_count.cu
int slots_idle(const igdb_dev *d, const char *who);                                // igd_dev_count.cu
int count_mode_for(const igdb_dev *d, bool mode_v, int32_t v);                     // igd_dev_count.cu
int64_t full_device(igdb_dev *d, int64_t n, cudaStream_t st, bool with_g, DevBuf &qoff, DevBuf &out);   // igd_dev_full.cu

}  // namespace igdb
#endif
