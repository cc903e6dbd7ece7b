// igd_b200 CUDA layer -- sharding a query set over the ranks of a multi-GPU search, on the device.
//
// The database is replicated and the work is cut by QUERY: rank r takes the queries whose first bin lies in a
// contiguous range of tiles [cut[r], cut[r+1]) -- contiguous ranges of the (contig, start)-sorted query set, at
// tile granularity -- with the ranges balanced by the prefix sum of the estimated scan bytes of their queries
// (16 + sum over the bins a query touches of (8 + 4 nCnt), the figure of igd_b200_query_cost).  Inside a shard
// the order of the queries does not matter: the count pipeline sorts its segments by tile itself.
//   k_shard_cost       every query adds its cost to its first tile's slot (queries that reach no tile: slot T)
//   (the caller sums the slots over the ranks: one all-reduce of nTiles + 1 words)
//   k_shard_cuts       one CTA: prefix sum over the tiles, cut[p] = first tile at which the sum reaches p/n of it
//   k_shard_count / k_shard_scatter   counting sort of the queries by destination rank
#include "igd_dev_impl.cuh"

#include <cstring>

namespace {

constexpr int SH_T = 256, SH_MAXP = 64;

__device__ __forceinline__ int32_t first_tile(const DbView &db, int32_t c, int32_t s, int32_t *mT_out, int32_t *tb_out)
{
    if (c < 0 || c >= db.nctg) return -1;
    const int32_t tb = __ldg(&db.tbase[c]), mT = __ldg(&db.tbase[c + 1]) - tb - 1;
    const int32_t n1 = s / db.nbp;                          // C truncating division (src/igd_search.c:459)
    if (n1 < 0 || n1 > mT) return -1;
    *mT_out = mT, *tb_out = tb;
    return tb + n1;
}

__global__ void __launch_bounds__(SH_T) k_shard_cost(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                      const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                      unsigned long long *__restrict__ cost)
{
    for (int64_t q = (int64_t)blockIdx.x * SH_T + threadIdx.x; q < n; q += (int64_t)gridDim.x * SH_T) {
        const int32_t s = qs[q], e = qe[q];
        int32_t mT = 0, tb = 0;
        const int32_t t = first_tile(db, ctg[q], s, &mT, &tb);
        unsigned long long c = 16;
        if (t >= 0) {
            const int32_t n1 = t - tb, n2 = max(n1, min((e - 1) / db.nbp, mT));
            for (int32_t j = n1; j <= n2; ++j) c += 8ull + 4ull * (unsigned long long)(uint32_t)__ldg(&db.tiles[tb + j].n);
        }
        atomicAdd(&cost[t >= 0 ? t : db.ntiles], c);
    }
}

// cuts[0] = 0, cuts[n_parts] = T; cuts[p] = the first tile t with (cost of tiles [0, t]) >= total * p / n_parts --
// what numpy.searchsorted(cumsum(cost), total * p / n_parts) gives.  (Slot T, the queries that reach no tile, is not
// part of the sum: they go to the last rank and cost it nothing.)
__global__ void __launch_bounds__(SCAN_T) k_shard_cuts(const unsigned long long *__restrict__ cost, int32_t T,
                                                        int32_t n_parts, int32_t *__restrict__ cuts)
{
    __shared__ int32_t s_cut[SH_MAXP + 1];
    const int32_t per = (T + SCAN_T - 1) / SCAN_T, t0 = min((int32_t)threadIdx.x * per, T), t1 = min(t0 + per, T);
    unsigned long long sum = 0;
    for (int32_t t = t0; t < t1; ++t) sum += cost[t];
    for (int p = threadIdx.x; p <= n_parts; p += SCAN_T) s_cut[p] = T;
    unsigned long long total;
    unsigned long long run = block_exclusive<unsigned long long>(sum, &total);
    for (int32_t t = t0; t < t1; ++t) {
        const unsigned long long before = run;
        run += cost[t];
        // tile t is cut p's place iff before < total p / n <= run  (in exact arithmetic: compare x n_parts)
        for (int p = 1; p < n_parts; ++p) {
            const unsigned long long target = total * (unsigned long long)p;
            if (before * (unsigned long long)n_parts < target && target <= run * (unsigned long long)n_parts) atomicMin(&s_cut[p], t);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_cut[0] = 0;
        for (int p = 1; p <= n_parts; ++p) s_cut[p] = max(s_cut[p], s_cut[p - 1]);     // monotone (total == 0: all T)
        s_cut[n_parts] = T;
    }
    __syncthreads();
    for (int p = threadIdx.x; p <= n_parts; p += SCAN_T) cuts[p] = s_cut[p];
}

__device__ __forceinline__ int dest_of(const int32_t *cut_s, int n_parts, int32_t t)
{
    if (t < 0) return n_parts - 1;
    int p = 0;
    while (p + 1 < n_parts && t >= cut_s[p + 1]) ++p;
    return p;
}


// the lanes of a warp that go to the same destination count themselves with ONE shared-memory atomic (a few
// destinations, millions of queries: one atomic per query is 256 threads queueing on two or eight words); the reach
// is a running maximum, raised only when a lane sees that it would move it (a stale read costs a spare atomic)
__device__ __forceinline__ void warp_count_reach(unsigned int *cnt_s, int32_t *reach_s, int p, int32_t far)
{
    const unsigned act = __activemask();
    const unsigned same = __match_any_sync(act, p);
    if ((int)(threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&cnt_s[p], (unsigned)__popc(same));
    if (far > *(volatile int32_t *)&reach_s[p]) atomicMax(&reach_s[p], far);
}

// the same for the scatter pass: the lane's rank among the warp's lanes of the same destination + the run the leader
// reserved in the CTA's counter
__device__ __forceinline__ unsigned warp_rank_in(unsigned int *cnt_s, int p)
{
    const unsigned act = __activemask();
    const unsigned same = __match_any_sync(act, p);
    const int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(&cnt_s[p], (unsigned)__popc(same));
    base = __shfl_sync(act, base, leader);
    return base + __popc(same & ((1u << lane) - 1));
}

// pass 1: queries per destination; also the last tile any query of a destination reaches (+ 1)
__global__ void __launch_bounds__(SH_T) k_shard_count(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                       const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                       const int32_t *__restrict__ cuts, int n_parts,
                                                       unsigned long long *__restrict__ count, int32_t *__restrict__ reach)
{
    __shared__ int32_t cut_s[SH_MAXP + 1];
    __shared__ unsigned int cnt_s[SH_MAXP];
    __shared__ int32_t reach_s[SH_MAXP];
    for (int p = threadIdx.x; p <= n_parts; p += SH_T) cut_s[p] = cuts[p];
    for (int p = threadIdx.x; p < n_parts; p += SH_T) cnt_s[p] = 0, reach_s[p] = 0;
    __syncthreads();
    for (int64_t q = (int64_t)blockIdx.x * SH_T + threadIdx.x; q < n; q += (int64_t)gridDim.x * SH_T) {
        int32_t mT = 0, tb = 0;
        const int32_t t = first_tile(db, ctg[q], qs[q], &mT, &tb);
        const int p = dest_of(cut_s, n_parts, t);
        const int32_t far = t >= 0 ? tb + max(t - tb, min((qe[q] - 1) / db.nbp, mT)) + 1 : 0;
        warp_count_reach(cnt_s, reach_s, p, far);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_parts; p += SH_T) {
        if (cnt_s[p]) atomicAdd(&count[p], (unsigned long long)cnt_s[p]);
        if (reach_s[p]) atomicMax(&reach[p], reach_s[p]);
    }
}

// off[p] = first output slot of destination p (exclusive prefix of count; off[n_parts] = n); cursor = copy
__global__ void k_shard_offsets(const unsigned long long *__restrict__ count, int n_parts, long long *__restrict__ off,
                                unsigned long long *__restrict__ cursor)
{
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int p = 0; p < n_parts; ++p) {
            off[p] = (long long)run, cursor[p] = run;
            run += count[p];
        }
        off[n_parts] = (long long)run;
    }
}

// pass 2: a CTA reserves, per destination, one run of slots for its queries and writes them there
__global__ void __launch_bounds__(SH_T) k_shard_scatter(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                         const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                         const int32_t *__restrict__ cuts, int n_parts,
                                                         unsigned long long *__restrict__ cursor,
                                                         int32_t *__restrict__ o_ctg, int32_t *__restrict__ o_qs,
                                                         int32_t *__restrict__ o_qe)
{
    __shared__ int32_t cut_s[SH_MAXP + 1];
    __shared__ unsigned int cnt_s[SH_MAXP];
    __shared__ unsigned long long base_s[SH_MAXP];
    for (int p = threadIdx.x; p <= n_parts; p += SH_T) cut_s[p] = cuts[p];
    __syncthreads();
    constexpr int PER = 8;                                  // queries per thread and turn
    for (int64_t q0 = (int64_t)blockIdx.x * (SH_T * PER); q0 < n; q0 += (int64_t)gridDim.x * (SH_T * PER)) {
        for (int p = threadIdx.x; p < n_parts; p += SH_T) cnt_s[p] = 0;
        __syncthreads();
        int32_t c[PER], s[PER], e[PER];
        int dst[PER];
        unsigned int rk[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t q = q0 + (int64_t)k * SH_T + threadIdx.x;
            dst[k] = -1;
            if (q < n) {
                c[k] = ctg[q], s[k] = qs[q], e[k] = qe[q];
                int32_t mT = 0, tb = 0;
                dst[k] = dest_of(cut_s, n_parts, first_tile(db, c[k], s[k], &mT, &tb));
                rk[k] = warp_rank_in(cnt_s, dst[k]);
            }
        }
        __syncthreads();
        for (int p = threadIdx.x; p < n_parts; p += SH_T)
            base_s[p] = cnt_s[p] ? atomicAdd(&cursor[p], (unsigned long long)cnt_s[p]) : 0ull;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (dst[k] >= 0) {
                const unsigned long long o = base_s[dst[k]] + rk[k];
                o_ctg[o] = c[k], o_qs[o] = s[k], o_qe[o] = e[k];
            }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// The partition FUSED with its all-to-all over NVLink peer memory (igd_b200_dispatcher_*): instead of regrouping the
// queries by destination in local memory and handing the runs to a library all-to-all, the scatter pass stores every
// query straight into the receive buffer of the rank that will search it.  One process per GPU; every rank owns one
// block of HBM, exported to the others as a CUDA IPC handle:
//     ctrl:  cnt[2][W][W]   cnt[s][src][dst] = queries rank src sends to rank dst      (slot s = epoch & 1)
//            reach[2][W][W] one past the last tile those queries touch
//            flagA[2][W], flagB[2][W], status
//     recv:  ctg[cap] | qs[cap] | qe[cap]
// A dispatch of epoch e is three launches per rank, queued on the caller's stream:
//   k_dispatch_count    per-destination counts and reach of this rank's queries; the last CTA to finish writes the row
//                       [rank][*] into EVERY rank's cnt / reach and raises flagA[rank] = e there;
//   k_dispatch_scatter  waits for all W flagA of its own block -- it then holds the whole W x W matrix, so the slot
//                       range of (src -> dst) in dst's buffer is known without another exchange: it starts at the sum
//                       of cnt[src'][dst] over src' < src -- and stores its queries into the destinations' buffers
//                       (its own included), a CTA reserving one run of slots per destination at a time; every CTA
//                       fences (system scope), the last one raises flagB[rank] = e in every rank's block;
//   k_dispatch_wait     waits for all W flagB of its own block: its shard is complete in recv, rank-major by source.
// ONE receive buffer is enough as long as the search that consumes it is queued on the same stream before the next
// dispatch: a peer stores into it only after this rank has raised flagA for that epoch, which k_dispatch_count does,
// in stream order, after the search.  The matrix has two slots by epoch parity for the same reason as the reducer's
// mailbox (igd_dev_p2p.cu).  Waits are bounded (20 s) and leave a mark in the status word.
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long DISPATCH_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

struct DispatchView {
    unsigned long long *const *peer;        // every rank's block, as mapped here
    int rank, world;
    unsigned long long epoch;
    long long cap;                          // queries a receive buffer holds
    size_t ctrl_words;                      // recv starts this many 8-byte words into a block
};

__device__ __forceinline__ size_t dc_cnt(const DispatchView &v, int s, int src, int dst) { return ((size_t)s * v.world + src) * v.world + dst; }
__device__ __forceinline__ size_t dc_reach(const DispatchView &v, int s, int src, int dst) { return (size_t)2 * v.world * v.world + dc_cnt(v, s, src, dst); }
__device__ __forceinline__ size_t dc_flag(const DispatchView &v, int which, int s, int q) { return (size_t)4 * v.world * v.world + ((size_t)which * 2 + s) * v.world + q; }
__device__ __forceinline__ size_t dc_status(const DispatchView &v) { return (size_t)4 * v.world * v.world + (size_t)4 * v.world; }

__device__ __forceinline__ void st_release_sys64(unsigned long long *p, unsigned long long x)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(x) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys64(const unsigned long long *p)
{
    unsigned long long x;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(p) : "memory");
    return x;
}
// thread q < world waits until flag[which][s][q] of this rank's own block shows the epoch
__device__ __forceinline__ void dispatch_wait_flags(const DispatchView &v, int which)
{
    if ((int)threadIdx.x < v.world) {
        unsigned long long *mine = v.peer[v.rank];
        const unsigned long long *fl = mine + dc_flag(v, which, (int)(v.epoch & 1), threadIdx.x);
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys64(fl) != v.epoch) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > DISPATCH_TIMEOUT_NS) {
                mine[dc_status(v)] = 1ull + threadIdx.x;
                break;
            }
        }
    }
    __syncthreads();
}

// local scratch of a dispatch: lcnt[SH_MAXP] | lreach[SH_MAXP] | cursor[SH_MAXP] | ticket[2] | meta[4]
__global__ void __launch_bounds__(SH_T) k_dispatch_count(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                          const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                          const int32_t *__restrict__ cuts, DispatchView v,
                                                          unsigned long long *__restrict__ scratch)
{
    __shared__ int32_t cut_s[SH_MAXP + 1];
    __shared__ unsigned int cnt_s[SH_MAXP];
    __shared__ int32_t reach_s[SH_MAXP];
    __shared__ bool last;
    const int W = v.world;
    unsigned long long *lcnt = scratch, *lreach = scratch + SH_MAXP, *ticket = scratch + 3 * SH_MAXP;
    for (int p = threadIdx.x; p <= W; p += SH_T) cut_s[p] = cuts[p];
    for (int p = threadIdx.x; p < W; p += SH_T) cnt_s[p] = 0, reach_s[p] = 0;
    __syncthreads();
    for (int64_t q = (int64_t)blockIdx.x * SH_T + threadIdx.x; q < n; q += (int64_t)gridDim.x * SH_T) {
        int32_t mT = 0, tb = 0;
        const int32_t t = first_tile(db, ctg[q], qs[q], &mT, &tb);
        const int p = dest_of(cut_s, W, t);
        const int32_t far = t >= 0 ? tb + max(t - tb, min((qe[q] - 1) / db.nbp, mT)) + 1 : 0;
        warp_count_reach(cnt_s, reach_s, p, far);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < W; p += SH_T) {
        if (cnt_s[p]) atomicAdd(&lcnt[p], (unsigned long long)cnt_s[p]);
        if (reach_s[p]) atomicMax(&lreach[p], (unsigned long long)reach_s[p]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&ticket[0], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    // the last CTA: this rank's row of the matrix into every rank's block, then the flag
    __threadfence();
    const int s = (int)(v.epoch & 1);
    for (int i = threadIdx.x; i < W * W; i += SH_T) {
        const int p = i / W, d = i - p * W;
        const unsigned long long c = *(volatile unsigned long long *)&lcnt[d], r = *(volatile unsigned long long *)&lreach[d];
        v.peer[p][dc_cnt(v, s, v.rank, d)] = c;
        v.peer[p][dc_reach(v, s, v.rank, d)] = r;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W) st_release_sys64(v.peer[threadIdx.x] + dc_flag(v, 0, s, v.rank), v.epoch);
    if (threadIdx.x == 0) ticket[0] = 0;
}

__global__ void __launch_bounds__(SH_T) k_dispatch_scatter(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                            const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                            const int32_t *__restrict__ cuts, DispatchView v,
                                                            unsigned long long *__restrict__ scratch)
{
    __shared__ int32_t cut_s[SH_MAXP + 1];
    __shared__ unsigned int cnt_s[SH_MAXP];
    __shared__ unsigned long long base_s[SH_MAXP], first_s[SH_MAXP];
    __shared__ int32_t *recv_s[SH_MAXP];
    __shared__ bool last;
    const int W = v.world, s = (int)(v.epoch & 1);
    unsigned long long *cursor = scratch + 2 * SH_MAXP, *ticket = scratch + 3 * SH_MAXP, *meta = scratch + 3 * SH_MAXP + 2;
    unsigned long long *mine = v.peer[v.rank];
    for (int p = threadIdx.x; p <= W; p += SH_T) cut_s[p] = cuts[p];
    dispatch_wait_flags(v, 0);                              // the whole W x W matrix is here
    for (int d = threadIdx.x; d < W; d += SH_T) {
        unsigned long long before = 0;                      // slots of dst d taken by the sources before this rank
        for (int src = 0; src < v.rank; ++src) before += *(volatile unsigned long long *)&mine[dc_cnt(v, s, src, d)];
        first_s[d] = before;
        recv_s[d] = reinterpret_cast<int32_t *>(v.peer[d] + v.ctrl_words);
    }
    __syncthreads();
    constexpr int PER = 8;                                  // queries per thread and turn
    for (int64_t q0 = (int64_t)blockIdx.x * (SH_T * PER); q0 < n; q0 += (int64_t)gridDim.x * (SH_T * PER)) {
        for (int p = threadIdx.x; p < W; p += SH_T) cnt_s[p] = 0;
        __syncthreads();
        int32_t c[PER], b[PER], e[PER];
        int dst[PER];
        unsigned int rk[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int64_t q = q0 + (int64_t)k * SH_T + threadIdx.x;
            dst[k] = -1;
            if (q < n) {
                c[k] = ctg[q], b[k] = qs[q], e[k] = qe[q];
                int32_t mT = 0, tb = 0;
                dst[k] = dest_of(cut_s, W, first_tile(db, c[k], b[k], &mT, &tb));
                rk[k] = warp_rank_in(cnt_s, dst[k]);
            }
        }
        __syncthreads();
        for (int p = threadIdx.x; p < W; p += SH_T)
            base_s[p] = first_s[p] + (cnt_s[p] ? atomicAdd(&cursor[p], (unsigned long long)cnt_s[p]) : 0ull);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; ++k)
            if (dst[k] >= 0) {
                const unsigned long long o = base_s[dst[k]] + rk[k];
                if (o < (unsigned long long)v.cap) {
                    int32_t *r = recv_s[dst[k]];
                    r[o] = c[k], r[v.cap + o] = b[k], r[2 * v.cap + o] = e[k];
                } else {
                    mine[dc_status(v)] = 0x100;             // a receive buffer is too small (the host reports it)
                }
            }
        __syncthreads();
    }
    __threadfence_system();                                 // this CTA's stores, wherever they went
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&ticket[1], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if ((int)threadIdx.x < W) st_release_sys64(v.peer[threadIdx.x] + dc_flag(v, 1, s, v.rank), v.epoch);
    if (threadIdx.x == 0) {
        unsigned long long m = 0, far = 0;                  // what this rank receives, how far its shard reaches
        for (int src = 0; src < W; ++src) {
            m += *(volatile unsigned long long *)&mine[dc_cnt(v, s, src, v.rank)];
            far = max(far, *(volatile unsigned long long *)&mine[dc_reach(v, s, src, v.rank)]);
        }
        meta[0] = m, meta[1] = (unsigned long long)cut_s[v.rank], meta[2] = far;     // {queries, tile_lo, tile_hi}: igdb_dev_count_device_dyn's d_dyn
        ticket[1] = 0;
    }
}

__global__ void __launch_bounds__(64) k_dispatch_wait(DispatchView v, unsigned long long *__restrict__ scratch)
{
    dispatch_wait_flags(v, 1);
    if (threadIdx.x == 0) scratch[3 * SH_MAXP + 2 + 3] = *(volatile unsigned long long *)&v.peer[v.rank][dc_status(v)];
}

}  // namespace

extern "C" {

int igdb_dev_shard_cost(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        uint64_t *d_tile_cost, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (n <= 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    const unsigned grid = (unsigned)std::min<int64_t>((n + SH_T - 1) / SH_T, (int64_t)d->sm_count * 16);
    k_shard_cost<<<grid, SH_T, 0, st>>>(d->view(), n, d_ctg, d_qs, d_qe, (unsigned long long *)d_tile_cost);
    d->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int igdb_dev_shard_cuts(igdb_dev *d, const uint64_t *d_tile_cost, int32_t n_parts, int32_t *d_cuts, void *stream)
{
    g_err[0] = 0;
    if (n_parts < 1 || n_parts > SH_MAXP) {
        snprintf(g_err, sizeof g_err, "shard_cuts: 1 .. %d parts", SH_MAXP);
        return -1;
    }
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    k_shard_cuts<<<1, SCAN_T, 0, st>>>((const unsigned long long *)d_tile_cost, d->T, n_parts, d_cuts);
    d->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// d_part_off[n_parts + 1] (int64, device): where each destination's queries start in the output arrays;
// d_reach[n_parts] (int32, device, may be NULL): one past the last tile a destination's queries touch
int igdb_dev_shard_partition(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                             const int32_t *d_cuts, int32_t n_parts, int32_t *o_ctg, int32_t *o_qs, int32_t *o_qe,
                             int64_t *d_part_off, int32_t *d_reach, void *stream)
{
    g_err[0] = 0;
    if (n_parts < 1 || n_parts > SH_MAXP) {
        snprintf(g_err, sizeof g_err, "shard_partition: 1 .. %d parts", SH_MAXP);
        return -1;
    }
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    if (!d->shard_scratch) CK(cudaMalloc((void **)&d->shard_scratch, (size_t)(3 * SH_MAXP) * 8));
    unsigned long long *count = d->shard_scratch, *cursor = d->shard_scratch + SH_MAXP;
    int32_t *reach = d_reach ? d_reach : reinterpret_cast<int32_t *>(d->shard_scratch + 2 * SH_MAXP);
    CK(cudaMemsetAsync(count, 0, (size_t)SH_MAXP * 8, st));
    CK(cudaMemsetAsync(reach, 0, (size_t)n_parts * 4, st));
    const DbView db = d->view();
    if (n > 0) {
        const unsigned g1 = (unsigned)std::min<int64_t>((n + SH_T - 1) / SH_T, (int64_t)d->sm_count * 16);
        k_shard_count<<<g1, SH_T, 0, st>>>(db, n, d_ctg, d_qs, d_qe, d_cuts, n_parts, count, reach);
    }
    k_shard_offsets<<<1, 32, 0, st>>>(count, n_parts, (long long *)d_part_off, cursor);
    if (n > 0) {
        const unsigned g2 = (unsigned)std::min<int64_t>((n + SH_T * 8 - 1) / (SH_T * 8), (int64_t)d->sm_count * 8);
        k_shard_scatter<<<g2, SH_T, 0, st>>>(db, n, d_ctg, d_qs, d_qe, d_cuts, n_parts, cursor, o_ctg, o_qs, o_qe);
    }
    d->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

// ---- the dispatcher object -------------------------------------------------------------------------------
struct igdb_dispatcher {
    igdb_dev *dev = nullptr;
    int rank = 0, world = 1;
    int64_t cap = 0;
    size_t ctrl_words = 0, block_bytes = 0;
    unsigned long long *block = nullptr;            // own block: ctrl | recv
    unsigned long long *peer[SH_MAXP] = {};
    unsigned long long **d_peer = nullptr;
    unsigned long long *scratch = nullptr;          // lcnt | lreach | cursor | ticket[2] | meta[4]
    unsigned long long *h_meta = nullptr;           // pinned copy of meta
    unsigned long long epoch = 0;
    bool connected = false;
};

static size_t dispatch_ctrl_words(int world)
{
    const size_t w = (size_t)4 * world * world + (size_t)4 * world + 1;
    return (w + 31) / 32 * 32;                      // recv starts 256-byte aligned
}

igdb_dispatcher *igdb_dispatcher_create(igdb_dev *d, int rank, int world, int64_t cap)
{
    g_err[0] = 0;
    if (world < 1 || world > SH_MAXP || rank < 0 || rank >= world || cap < 1) {
        snprintf(g_err, sizeof g_err, "dispatcher: rank %d of %d, capacity %lld", rank, world, (long long)cap);
        return nullptr;
    }
    igdb_dispatcher *x = new igdb_dispatcher();
    x->dev = d, x->rank = rank, x->world = world, x->cap = cap;
    x->ctrl_words = dispatch_ctrl_words(world);
    x->block_bytes = x->ctrl_words * 8 + (size_t)3 * cap * 4;
    const size_t scratch_words = 3 * SH_MAXP + 2 + 4;
    if (cudaSetDevice(d->device) != cudaSuccess || cudaMalloc(&x->block, x->block_bytes) != cudaSuccess ||
        cudaMemset(x->block, 0, x->ctrl_words * 8) != cudaSuccess ||
        cudaMalloc(&x->d_peer, SH_MAXP * sizeof(void *)) != cudaSuccess ||
        cudaMalloc(&x->scratch, scratch_words * 8) != cudaSuccess ||
        cudaMemset(x->scratch, 0, scratch_words * 8) != cudaSuccess ||
        cudaMallocHost(&x->h_meta, 4 * 8) != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "dispatcher: %s", cudaGetErrorString(cudaGetLastError()));
        igdb_dispatcher_destroy(x);
        return nullptr;
    }
    x->peer[rank] = x->block;
    if (world == 1) {
        if (cudaMemcpy(x->d_peer, x->peer, SH_MAXP * sizeof(void *), cudaMemcpyHostToDevice) != cudaSuccess) {
            snprintf(g_err, sizeof g_err, "dispatcher: %s", cudaGetErrorString(cudaGetLastError()));
            igdb_dispatcher_destroy(x);
            return nullptr;
        }
        x->connected = true;
    }
    return x;
}

int igdb_dispatcher_handle(igdb_dispatcher *x, void *out64)
{
    g_err[0] = 0;
    cudaIpcMemHandle_t h;
    CK(cudaSetDevice(x->dev->device));
    CK(cudaIpcGetMemHandle(&h, x->block));
    memcpy(out64, &h, 64);
    return 0;
}

int igdb_dispatcher_connect(igdb_dispatcher *x, const void *handles)
{
    g_err[0] = 0;
    CK(cudaSetDevice(x->dev->device));
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)p, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[p] = (unsigned long long *)ptr;
    }
    CK(cudaMemcpy(x->d_peer, x->peer, SH_MAXP * sizeof(void *), cudaMemcpyHostToDevice));
    x->connected = true;
    return 0;
}

// The n queries of this rank (device arrays, any order) go to the ranks that own their first tile under d_cuts
// (igdb_dev_shard_cuts); on return this rank's shard -- *m queries, rank-major by source -- is in its receive buffer
// (*o_ctg / *o_qs / *o_qe point into it) and touches the tiles [*tile_lo, *tile_hi) only.  Enqueued on `stream`; the
// call waits for the three launches (the host needs m).  Every rank must call it the same number of times, and the
// search of the shard must be queued on the same stream before the next dispatch.
static int dispatch_enqueue(igdb_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                            const int32_t *d_cuts, cudaStream_t st)
{
    igdb_dev *d = x->dev;
    if (!x->connected) {
        snprintf(g_err, sizeof g_err, "dispatcher: not connected");
        return -1;
    }
    ++x->epoch;
    DispatchView v;
    v.peer = x->d_peer, v.rank = x->rank, v.world = x->world, v.epoch = x->epoch, v.cap = x->cap, v.ctrl_words = x->ctrl_words;
    const DbView db = d->view();
    n = std::max<int64_t>(n, 0);
    CK(cudaMemsetAsync(x->scratch, 0, (size_t)3 * SH_MAXP * 8, st));        // lcnt, lreach, cursor
    const unsigned g1 = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + SH_T - 1) / SH_T, (int64_t)d->sm_count * 8));
    k_dispatch_count<<<g1, SH_T, 0, st>>>(db, n, d_ctg, d_qs, d_qe, d_cuts, v, x->scratch);
    const unsigned g2 = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + SH_T * 8 - 1) / (SH_T * 8), (int64_t)d->sm_count * 8));
    k_dispatch_scatter<<<g2, SH_T, 0, st>>>(db, n, d_ctg, d_qs, d_qe, d_cuts, v, x->scratch);
    k_dispatch_wait<<<1, 64, 0, st>>>(v, x->scratch);
    d->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

static int dispatch_check(igdb_dispatcher *x)
{
    if (x->h_meta[3] == 0) return 0;
    if (x->h_meta[3] == 0x100)
        snprintf(g_err, sizeof g_err, "dispatcher: a receive buffer (%lld queries) is too small for its shard", (long long)x->cap);
    else
        snprintf(g_err, sizeof g_err, "dispatcher: the wait for rank %llu timed out", x->h_meta[3] - 1);
    return -1;
}

int igdb_dispatch(igdb_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                  const int32_t *d_cuts, void *stream, int32_t **o_ctg, int32_t **o_qs, int32_t **o_qe, int64_t *m,
                  int32_t *tile_lo, int32_t *tile_hi)
{
    g_err[0] = 0;
    igdb_dev *d = x->dev;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    if (dispatch_enqueue(x, n, d_ctg, d_qs, d_qe, d_cuts, st)) return -1;
    CK(cudaMemcpyAsync(x->h_meta, x->scratch + 3 * SH_MAXP + 2, 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dispatch_check(x)) return -1;
    int32_t *recv = reinterpret_cast<int32_t *>(x->block + x->ctrl_words);
    *o_ctg = recv, *o_qs = recv + x->cap, *o_qe = recv + 2 * x->cap;
    *m = (int64_t)x->h_meta[0];
    *tile_lo = (int32_t)x->h_meta[1];
    *tile_hi = std::max(*tile_lo, (int32_t)x->h_meta[2]);
    if (*m > x->cap) {
        snprintf(g_err, sizeof g_err, "dispatcher: shard of %lld queries, receive buffer of %lld", (long long)*m, (long long)x->cap);
        return -1;
    }
    return 0;
}

int igdb_dispatch_async(igdb_dispatcher *x, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                        const int32_t *d_cuts, void *stream, int32_t **o_ctg, int32_t **o_qs, int32_t **o_qe,
                        const uint64_t **d_dyn)
{
    g_err[0] = 0;
    igdb_dev *d = x->dev;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    if (dispatch_enqueue(x, n, d_ctg, d_qs, d_qe, d_cuts, st)) return -1;
    int32_t *recv = reinterpret_cast<int32_t *>(x->block + x->ctrl_words);
    *o_ctg = recv, *o_qs = recv + x->cap, *o_qe = recv + 2 * x->cap;
    *d_dyn = reinterpret_cast<const uint64_t *>(x->scratch + 3 * SH_MAXP + 2);
    return 0;
}

int igdb_dispatcher_status(igdb_dispatcher *x)
{
    g_err[0] = 0;
    CK(cudaSetDevice(x->dev->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(x->h_meta, x->scratch + 3 * SH_MAXP + 2, 32, cudaMemcpyDeviceToHost));
    return dispatch_check(x);
}

void igdb_dispatcher_destroy(igdb_dispatcher *x)
{
    if (!x) return;
    cudaSetDevice(x->dev->device);
    for (int p = 0; p < x->world; ++p)
        if (p != x->rank && x->peer[p]) cudaIpcCloseMemHandle(x->peer[p]);
    if (x->block) cudaFree(x->block);
    if (x->d_peer) cudaFree(x->d_peer);
    if (x->scratch) cudaFree(x->scratch);
    if (x->h_meta) cudaFreeHost(x->h_meta);
    delete x;
}

}  // extern "C"
