// igd_b200 CUDA layer -- a stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass, on the key
// bits [0, end_bit).  Seqpare orders its candidate pairs with it (igd_dev_seqpare.cu: (chrom group, file) major,
// similarity descending, ties in the reference's scan order -- which is why the sort has to be stable).
//
// A pass: k_rs_hist counts the digit of every key per tile of RS_TILE keys (shared-memory histogram), an exclusive
// scan over the (digit, tile) counts gives every tile its first output slot per digit, and k_rs_scatter places the
// keys: a warp owns a contiguous 1/RS_WARPS of the tile and keeps its RS_IPT x 32 keys in registers; the lanes of
// one load that share a digit find each other with match.any, so an item's rank among the equal digits of its tile
// is (equal digits in the warps before) + (in the earlier loads of its warp) + (in the lower lanes of its load) --
// the order of the input.  The tile is first laid out by digit in shared memory and then written out run by run,
// so that consecutive threads write consecutive slots (writing every key straight to its slot -- 32 different
// runs per store -- took 12 ms per pass for 3e8 pairs where this takes 2-3).
#ifndef IGD_DEV_SORT_CUH
#define IGD_DEV_SORT_CUH
#include "igd_dev_impl.cuh"

namespace igdb {

constexpr int RS_T = 512, RS_WARPS = RS_T / 32, RS_IPT = 8, RS_TILE = RS_T * RS_IPT;      // 4096 keys per tile
constexpr size_t RS_SMEM = (size_t)RS_TILE * 12 + (size_t)RS_WARPS * 256 * 4;

__global__ void __launch_bounds__(RS_T) k_rs_hist(const unsigned long long *__restrict__ keys, int64_t n, int shift,
                                                   uint32_t ntiles, uint32_t *__restrict__ gh)
{
    __shared__ uint32_t h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_IPT; ++k) {
        const int64_t i = base + (int64_t)k * RS_T + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) gh[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];     // digit-major: one scan
}

__global__ void __launch_bounds__(RS_T) k_rs_scatter(const unsigned long long *__restrict__ keys,
                                                      const uint32_t *__restrict__ vals, int64_t n, int shift,
                                                      uint32_t ntiles, const uint32_t *__restrict__ goff,
                                                      unsigned long long *__restrict__ okeys, uint32_t *__restrict__ ovals)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(rs_smem);             // [RS_TILE] the tile, by digit
    uint32_t *sval = reinterpret_cast<uint32_t *>(rs_smem + (size_t)RS_TILE * 8);           // [RS_TILE]
    uint32_t(*wcnt)[256] = reinterpret_cast<uint32_t(*)[256]>(rs_smem + (size_t)RS_TILE * 12);   // [RS_WARPS][256]
    __shared__ uint32_t tstart[256], gbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_T) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    // the warp's keys, in input order: load k covers [wbase + 32 k, + 32)
    const int64_t tbase = (int64_t)blockIdx.x * RS_TILE, wbase = tbase + (int64_t)w * (RS_IPT * 32);
    unsigned long long key[RS_IPT];
    uint32_t rank[RS_IPT];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const int64_t i = wbase + (int64_t)k * 32 + lane;
        const bool in = i < n;
        key[k] = in ? keys[i] : ~0ull;
        const uint32_t dg = (uint32_t)(key[k] >> shift) & 255u;
        // lanes of this load with the same digit (the ones past the end of the input form their own group)
        const unsigned peers = __match_any_sync(FULL, in ? dg : 256u);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (in && lane == leader) {
            before = wcnt[w][dg];
            wcnt[w][dg] = before + __popc(peers);
        }
        rank[k] = __shfl_sync(FULL, before, leader) + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // per digit: where it starts inside the tile (digits ascending, the warps one after the other) and in the output
    uint32_t tot = 0;
    if (threadIdx.x < 256)
#pragma unroll
        for (int q = 0; q < RS_WARPS; ++q) tot += wcnt[q][threadIdx.x];
    uint32_t all;
    uint32_t run = block_exclusive<uint32_t>(tot, &all);
    if (threadIdx.x < 256) {
        tstart[threadIdx.x] = run;
        gbase[threadIdx.x] = goff[(size_t)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int q = 0; q < RS_WARPS; ++q) {
            const uint32_t c = wcnt[q][threadIdx.x];
            wcnt[q][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    // the tile into shared memory in digit order (stable), then out in runs: consecutive threads, consecutive slots
#pragma unroll
    for (int k = 0; k < RS_IPT; ++k) {
        const int64_t i = wbase + (int64_t)k * 32 + lane;
        if (i < n) {
            const uint32_t lp = wcnt[w][(uint32_t)(key[k] >> shift) & 255u] + rank[k];
            skey[lp] = key[k];
            sval[lp] = vals[i];
        }
    }
    __syncthreads();
    const uint32_t tn = (uint32_t)min((int64_t)RS_TILE, n - tbase);
    for (uint32_t i = threadIdx.x; i < tn; i += RS_T) {
        const unsigned long long kk = skey[i];
        const uint32_t dg = (uint32_t)(kk >> shift) & 255u, o = gbase[dg] + (i - tstart[dg]);
        okeys[o] = kk;
        ovals[o] = sval[i];
    }
}

// sorts (keys, vals) by key bits [0, end_bit), stable; the result is in (keys_out, vals_out) or back in the inputs,
// *in_out says which (0: the inputs, 1: the second pair).  tmp: (256 * ntiles + 1) x 2 uint32 words.
inline size_t radix_sort_tmp_bytes(int64_t n)
{
    const size_t ntiles = (size_t)((n + RS_TILE - 1) / RS_TILE);
    return (256 * ntiles + 2) * 2 * sizeof(uint32_t);
}

inline int radix_sort_pairs(igdb_dev *d, unsigned long long *keys, unsigned long long *keys2, uint32_t *vals,
                            uint32_t *vals2, int64_t n, int end_bit, void *tmp, cudaStream_t st, int *in_out)
{
    *in_out = 0;
    if (n <= 1) return 0;
    if (n >= (1ll << 32) - RS_TILE) {
        snprintf(g_err, sizeof g_err, "radix sort: %lld items exceed 32-bit positions", (long long)n);
        return -1;
    }
    const uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t *gh = (uint32_t *)tmp, *goff = gh + (256 * (size_t)ntiles + 2);
    static_assert(RS_T == SCAN_T, "k_rs_scatter uses block_exclusive");
    CK(cudaFuncSetAttribute(k_rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
    int cur = 0;
    for (int shift = 0; shift < end_bit; shift += 8) {
        const unsigned long long *ki = cur ? keys2 : keys;
        const uint32_t *vi = cur ? vals2 : vals;
        k_rs_hist<<<ntiles, RS_T, 0, st>>>(ki, n, shift, ntiles, gh);
        if (exclusive_scan<uint32_t, uint32_t>(d, gh, goff, 256 * (int64_t)ntiles, st)) return -1;
        k_rs_scatter<<<ntiles, RS_T, RS_SMEM, st>>>(ki, vi, n, shift, ntiles, goff, cur ? keys : keys2, cur ? vals : vals2);
        d->launches += 2;
        CK(cudaGetLastError());
        cur ^= 1;
    }
    *in_out = cur;
    return 0;
}

}  // namespace igdb
#endif
