// igd_b200 CUDA layer -- -m: the file x file hitsmap (database self-join).
#include "igd_dev_impl.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// -m hitsmap: getMap / getMap_v (src/igd_search.c:772-886).  The reference takes every record j of a tile as a
// query against the tile: partners i in [tS, last start < end_j] with end_i > start_j, where tS is the tile's
// dedupe start only when start_j < bin boundary (:802-804); -v keeps the pairs whose two values both exceed v
// (STRICT `>`, :856,872).  Worked through, that rule counts a pair (j, i) in a tile iff both records are in it,
// they overlap, and at least one of them starts inside the bin -- a condition that is symmetric in j and i.  So
// (j, i) is counted exactly where (i, j) is, and only HALF of the pairs need to be found: one warp per record j
// that starts inside the bin (index >= ts) scans the records BEFORE it in start order (every one of them has
// start_i <= start_j < end_j: no search needed) for end_i > start_j and adds each find to half[file_j][file_i];
// the record itself goes to the diagonal of the result.  k_hitsmap_mirror then adds half[a][b] + half[b][a] to
// hitmap[a][b] (32-bit wrap-around is preserved: addition mod 2^32 twice).  Half the partner scans, half the
// atomics, no binary search.  (tests/test_brute_model.py states the same closed form on the CPU and checks it
// against the reference's committed hitsmaps.)
// work items = padded record slots, so a flat grid-stride loop over P records needs no extra index.
// ------------------------------------------------------------------------------------------------
template <bool V>
__global__ void __launch_bounds__(256) k_hitsmap(DbView db, const int32_t *__restrict__ rec_tile, int64_t r0,
                                                  int64_t P, int32_t v, uint32_t *__restrict__ half,
                                                  uint32_t *__restrict__ hitmap, unsigned long long *__restrict__ nols)
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
        if (j < tm.z) continue;                              // starts before the bin: its pairs belong to the partner
        if (V && !(__ldg(&db.value[r]) > v)) continue;
        const int qs = __ldg(&db.start[r]);
        const int jj = __ldg(&db.idx[r]);
        uint32_t *row = half + (int64_t)jj * db.nfiles;
        if (lane == 0) {                                     // (j, j)
            atomicAdd(&hitmap[(int64_t)jj * db.nfiles + jj], 1u);
            ++mine;
        }
        for (int b = lane * 4; b < j; b += 128) {
            const int4 e = ldg4(db.end + off + b);
            int4 val = make_int4(0, 0, 0, 0);
            if (V) val = ldg4(db.value + off + b);
            const int ev[4] = {e.x, e.y, e.z, e.w};
            const int vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = b + k;
                if (i < j && ev[k] > qs && (!V || vv[k] > v)) {
                    atomicAdd(&row[__ldg(&db.idx[off + i])], 1u);
                    mine += 2;                               // (j, i) and (i, j)
                }
            }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(FULL, mine, d);
    if (lane == 0 && mine) atomicAdd(nols, mine);
}

// hitmap[a][b] += half[a][b] + half[b][a]
__global__ void __launch_bounds__(256) k_hitsmap_mirror(const uint32_t *__restrict__ half, int32_t nfiles,
                                                         uint32_t *__restrict__ hitmap)
{
    const int64_t cells = (int64_t)nfiles * nfiles;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = c / nfiles, b = c - a * nfiles;
        const uint32_t x = half[c] + half[b * nfiles + a];
        if (x) hitmap[c] += x;
    }
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
    uint32_t *d_map = nullptr, *d_half = nullptr;
    int32_t *d_rt = nullptr;
    unsigned long long *d_n = nullptr;
    CK(cudaMalloc((void **)&d_map, std::max<size_t>(cells, 1) * 4));
    CK(cudaMalloc((void **)&d_half, std::max<size_t>(cells, 1) * 4));
    CK(cudaMemsetAsync(d_half, 0, std::max<size_t>(cells, 1) * 4, st));
    CK(cudaMalloc((void **)&d_rt, ((size_t)d->P + 8) * 4));
    CK(cudaMalloc((void **)&d_n, 8));
    CK(cudaMemcpyAsync(d_map, hitmap, cells * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_n, 0, 8, st));
    const int grid = d->sm_count * 8;
    k_rec_tile<<<grid, 256, 0, st>>>(d->tiles, d->T, d_rt);
    if (v > 0) k_hitsmap<true><<<grid, 256, 0, st>>>(d->view(), d_rt, r0, r1, v, d_half, d_map, d_n);
    else k_hitsmap<false><<<grid, 256, 0, st>>>(d->view(), d_rt, r0, r1, v, d_half, d_map, d_n);
    k_hitsmap_mirror<<<grid, 256, 0, st>>>(d_half, d->nfiles, d_map);
    d->launches += 3;
    CK(cudaGetLastError());
    unsigned long long h_n = 0;
    CK(cudaMemcpyAsync(hitmap, d_map, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_n, d_n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_map));
    CK(cudaFree(d_half));
    CK(cudaFree(d_rt));
    CK(cudaFree(d_n));
    return (int64_t)h_n;
}

}  // extern "C"
