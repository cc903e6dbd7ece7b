// igd_b200 CUDA layer -- -m: the file x file hitsmap (database self-join).
#include "igd_dev_impl.cuh"

namespace {

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


}  // namespace

extern "C" {

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
