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
        atomicAdd(&cnt_s[p], 1u);
        if (t >= 0) atomicMax(&reach_s[p], tb + max(t - tb, min((qe[q] - 1) / db.nbp, mT)) + 1);
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
                rk[k] = atomicAdd(&cnt_s[dst[k]], 1u);
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

}  // extern "C"
