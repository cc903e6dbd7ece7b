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
// Eight lanes per record, four records per warp: with the early exit a record looks at 32-64 predecessors, not at
// hundreds (a whole warp per record spent 204 warp instructions per record, most of its lanes idle).
template <bool V>
__global__ void __launch_bounds__(256) k_hitsmap(DbView db, const int32_t *__restrict__ rec_tile,
                                                  const int32_t *__restrict__ pmax, int64_t r0, int64_t P, int32_t v,
                                                  uint32_t *__restrict__ half, uint32_t *__restrict__ hitmap,
                                                  unsigned long long *__restrict__ nols)
{
    const int lane = threadIdx.x & 31, l8 = lane & 7;
    const int64_t nsub = (int64_t)gridDim.x * (blockDim.x >> 3);
    unsigned long long mine = 0;
    for (int64_t r = r0 + (int64_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); r < P; r += nsub) {
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
        if (l8 == 0) {                                       // (j, j)
            atomicAdd(&hitmap[(int64_t)jj * db.nfiles + jj], 1u);
            ++mine;
        }
        // the records before j in start order, nearest first, 32 at a time; pmax[i] = the largest end among
        // records 0 .. i of the tile (k_hm_prepare), so once pmax[b1 - 1] <= start_j nobody in [0, b1) reaches j --
        // the reference's own early exit (src/igd_search.c:797-800), which leaves the result as it is
        for (int b1 = j; b1 > 0;) {
            if (__ldg(&pmax[off + b1 - 1]) <= qs) break;
            const int b0 = (b1 - 1) & ~31, b = b0 + l8 * 4;
            if (b < b1) {
                const int4 e = ldg4(db.end + off + b);
                int4 val = make_int4(0, 0, 0, 0);
                if (V) val = ldg4(db.value + off + b);
                const int ev[4] = {e.x, e.y, e.z, e.w};
                const int vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = b + k;
                    if (i < b1 && ev[k] > qs && (!V || vv[k] > v)) {
                        atomicAdd(&row[__ldg(&db.idx[off + i])], 1u);
                        mine += 2;                           // (j, i) and (i, j)
                    }
                }
            }
            b1 = b0;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(FULL, mine, d);
    if (lane == 0 && mine) atomicAdd(nols, mine);
}

// hitmap[a][b] += half[a][b] + half[b][a]; half is left zeroed for the next call
__global__ void __launch_bounds__(256) k_hitsmap_mirror(uint32_t *__restrict__ half, int32_t nfiles,
                                                         uint32_t *__restrict__ hitmap)
{
    const int64_t cells = (int64_t)nfiles * nfiles;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = c / nfiles, b = c - a * nfiles;
        if (b < a) continue;                                 // the cell pair (a, b), (b, a) is done by its upper one
        const uint32_t x = half[c] + (a == b ? 0u : half[b * nfiles + a]);
        if (x) {
            hitmap[c] += x + (a == b ? x : 0u);              // (a, a): half[a][a] counts each unordered pair once
            if (a != b) hitmap[b * nfiles + a] += x;
        }
        half[c] = 0;
        if (a != b) half[b * nfiles + a] = 0;
    }
}

// The same pairs, found FILE BY FILE: the record slots are listed by file (k_hm_by_file, once per handle), a CTA takes
// a contiguous stretch of that list, i.e. the records of a few files one file after the other, and counts the partners
// of a file's records in a row of nfiles counters in SHARED memory -- one shared-memory atomic per pair where the
// tile-by-tile kernel above issues one L2 atomic per pair into a 14 MB matrix (2.0 G of them at BASELINE config 2).
// MEASURED SLOWER (48 ms against 31 ms) and not the default: see hm_prepare.  A finished row is added to hitmap[f][*] and to hitmap[*][f] (the mirror image), the
// pairs inside the file and the records themselves to the diagonal.  The price is locality: consecutive records of a
// file lie in different tiles, so the partner scans read L2, not L1.
template <bool V>
__global__ void __launch_bounds__(256) k_hitsmap_rows(DbView db, const int32_t *__restrict__ rec_tile,
                                                       const int32_t *__restrict__ pmax,
                                                       const uint32_t *__restrict__ frec,
                                                       const unsigned long long *__restrict__ fstart, int64_t r0,
                                                       int64_t r1, int32_t v, uint32_t *__restrict__ hitmap,
                                                       unsigned long long *__restrict__ nols)
{
    extern __shared__ uint32_t row[];                       // [nfiles] + {records of the file taken here}
    __shared__ int32_t s_f;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nf = db.nfiles;
    const unsigned long long ntot = fstart[nf];
    const unsigned long long per = (ntot + gridDim.x - 1) / gridDim.x;
    const unsigned long long p0 = min((unsigned long long)blockIdx.x * per, ntot), p1 = min(p0 + per, ntot);
    if (p0 >= p1) return;
    if (threadIdx.x == 0) {                                 // the file of position p0: last f with fstart[f] <= p0
        int lo = 0, hi = nf;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (fstart[mid] <= p0) lo = mid; else hi = mid;
        }
        s_f = lo;
    }
    __syncthreads();
    unsigned long long mine = 0;
    for (int f = s_f; f < nf && fstart[f] < p1; ++f) {
        const unsigned long long a = max(p0, fstart[f]), b = min(p1, fstart[f + 1]);
        if (a >= b) continue;
        for (int i = threadIdx.x; i <= nf; i += blockDim.x) row[i] = 0;
        __syncthreads();
        for (unsigned long long p = a + w; p < b; p += 8) {
            const int64_t r = (int64_t)frec[p];
            if (r < r0 || r >= r1) continue;                 // outside the tile range of this call
            const int t = __ldg(&rec_tile[r]);
            const int4 tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
            const int64_t off = (int64_t)(uint32_t)tm.x * 8;
            const int j = (int)(r - off);
            if (j < tm.z) continue;                          // starts before the bin: its pairs belong to the partner
            if (V && !(__ldg(&db.value[r]) > v)) continue;
            const int qs = __ldg(&db.start[r]);
            if (lane == 0) atomicAdd(&row[nf], 1u);          // (j, j)
            for (int b1 = j; b1 > 0;) {                      // nearest first; pmax: see k_hitsmap
                if (__ldg(&pmax[off + b1 - 1]) <= qs) break;
                const int b0 = (b1 - 1) & ~127, bb = b0 + lane * 4;
                if (bb < b1) {
                    const int4 e = ldg4(db.end + off + bb);
                    int4 val = make_int4(0, 0, 0, 0);
                    if (V) val = ldg4(db.value + off + bb);
                    const int ev[4] = {e.x, e.y, e.z, e.w};
                    const int vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = bb + k;
                        if (i < b1 && ev[k] > qs && (!V || vv[k] > v)) atomicAdd(&row[__ldg(&db.idx[off + i])], 1u);
                    }
                }
                b1 = b0;
            }
        }
        __syncthreads();
        uint32_t *hrow = hitmap + (int64_t)f * nf;
        for (int i = threadIdx.x; i < nf; i += blockDim.x) {
            const uint32_t c = row[i];
            if (i == f) {
                const uint32_t d2 = 2u * c + row[nf];
                if (d2) atomicAdd(&hrow[f], d2), mine += d2;
            } else if (c) {
                atomicAdd(&hrow[i], c);
                atomicAdd(&hitmap[(int64_t)i * nf + f], c);
                mine += 2ull * c;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(FULL, mine, d);
    if (lane == 0 && mine) atomicAdd(nols, mine);
}

// once per handle: the real record slots grouped by file (any order inside a file): count, scan, scatter
__global__ void k_hm_file_count(const int32_t *__restrict__ rec_tile, const int32_t *__restrict__ idx, int64_t P,
                                uint32_t *__restrict__ cnt)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P; r += (int64_t)gridDim.x * blockDim.x)
        if (rec_tile[r] >= 0) atomicAdd(&cnt[idx[r]], 1u);
}
__global__ void k_hm_file_scatter(const int32_t *__restrict__ rec_tile, const int32_t *__restrict__ idx, int64_t P,
                                  unsigned long long *__restrict__ cursor, uint32_t *__restrict__ frec)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < P; r += (int64_t)gridDim.x * blockDim.x)
        if (rec_tile[r] >= 0) frec[atomicAdd(&cursor[idx[r]], 1ull)] = (uint32_t)r;
}

// once per handle: rec_tile[slot] = tile of the record in that padded slot (-1 for pad slots), pmax[slot] = the
// largest end among the tile's records up to and including that one.  One warp per tile.
__global__ void k_hm_prepare(const TileMeta *__restrict__ tiles, int32_t T, const int32_t *__restrict__ end,
                             int32_t *__restrict__ rec_tile, int32_t *__restrict__ pmax)
{
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < T; t += nw) {
        const TileMeta m = tiles[t];
        const int64_t off = (int64_t)m.off8 * 8;
        const int np = (m.n + 7) & ~7;
        int carry = IMIN;
        for (int i0 = 0; i0 < np; i0 += 32) {
            const int i = i0 + lane;
            int x = i < m.n ? end[off + i] : IMIN;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(FULL, x, d);
                if (lane >= d) x = max(x, y);
            }
            x = max(x, carry);
            if (i < np) rec_tile[off + i] = i < m.n ? (int32_t)t : -1, pmax[off + i] = x;
            carry = __shfl_sync(FULL, x, 31);
        }
    }
}


}  // namespace

namespace {

// scratch of -m, kept by the handle: the two per-slot tables, the half matrix (zero between calls), the pair counter
int hm_prepare(igdb_dev *d, cudaStream_t st)
{
    const size_t cells = std::max<size_t>((size_t)d->nfiles * d->nfiles, 1);
    if (!d->hm_rt || !d->hm_pmax) {
        if (!d->hm_rt) CK(cudaMalloc((void **)&d->hm_rt, ((size_t)d->P + 8) * 4));
        if (!d->hm_pmax) CK(cudaMalloc((void **)&d->hm_pmax, ((size_t)d->P + 8) * 4));
        k_hm_prepare<<<d->sm_count * 8, 256, 0, st>>>(d->tiles, d->T, d->end, d->hm_rt, d->hm_pmax);
        d->launches += 1;
        CK(cudaGetLastError());
    }
    // the file-by-file kernel is the slower one on B200 (48 ms against 31 ms at BASELINE config 2: its partner scans
    // miss L1 and L2) and stays switched off unless asked for
    d->hm_rows = ((size_t)d->nfiles + 1) * 4 <= 160 * 1024 && d->P < (1ll << 32) && getenv("IGD_B200_HM_ROWS") != nullptr;
    if (d->hm_rows && !d->hm_frec) {
        // the record slots by file: count per file, exclusive scan, scatter (the cursors end as the next file's start)
        DevBuf cnt, cur;
        if (cnt.alloc(((size_t)d->nfiles + 1) * 4) || cur.alloc(((size_t)d->nfiles + 1) * 8)) return -1;
        CK(cudaMalloc((void **)&d->hm_frec, (size_t)std::max<int64_t>(d->nrec, 1) * 4));
        CK(cudaMalloc((void **)&d->hm_fstart, ((size_t)d->nfiles + 2) * 8));
        CK(cudaMemsetAsync(cnt.p, 0, ((size_t)d->nfiles + 1) * 4, st));
        k_hm_file_count<<<d->sm_count * 8, 256, 0, st>>>(d->hm_rt, d->idx, d->P, cnt.as<uint32_t>());
        if (exclusive_scan<uint32_t, unsigned long long>(d, cnt.as<uint32_t>(), d->hm_fstart, d->nfiles, st)) return -1;
        CK(cudaMemcpyAsync(cur.p, d->hm_fstart, ((size_t)d->nfiles + 1) * 8, cudaMemcpyDeviceToDevice, st));
        k_hm_file_scatter<<<d->sm_count * 8, 256, 0, st>>>(d->hm_rt, d->idx, d->P, cur.as<unsigned long long>(), d->hm_frec);
        d->launches += 2;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));                      // cnt / cur go out of scope
    }
    if (!d->hm_rows && !d->hm_half) {
        CK(cudaMalloc((void **)&d->hm_half, cells * 4));
        CK(cudaMemsetAsync(d->hm_half, 0, cells * 4, st));
    }
    if (!d->hm_n) CK(cudaMalloc((void **)&d->hm_n, 8));
    return 0;
}

}  // namespace

extern "C" {

// the tiles [t0, t1) into d_map (device, uint32[nfiles * nfiles], incremented); enqueued on stream.  filter != 0
// keeps the pairs whose two values both exceed v (getMap_v's rule for ANY v; the CLI asks for it when v > 0).
int igdb_dev_hitsmap_device(igdb_dev *d, int filter, int32_t v, int64_t t0, int64_t t1, uint32_t *d_map, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : d->stream;
    t0 = std::max<int64_t>(t0, 0), t1 = std::min<int64_t>(t1, d->T);
    if (hm_prepare(d, st)) return -1;
    CK(cudaMemsetAsync(d->hm_n, 0, 8, st));
    if (t0 >= t1) return 0;
    // padded record slots of the tile range [t0, t1)
    const int64_t r0 = (int64_t)d->off8[t0] * 8, r1 = t1 < d->T ? (int64_t)d->off8[t1] * 8 : d->P;
    const int grid = d->sm_count * 8;
    if (d->hm_rows) {
        const size_t shm = ((size_t)d->nfiles + 1) * 4;
        auto kern = filter ? k_hitsmap_rows<true> : k_hitsmap_rows<false>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        kern<<<grid, 256, shm, st>>>(d->view(), d->hm_rt, d->hm_pmax, d->hm_frec, d->hm_fstart, r0, r1, v, d_map, d->hm_n);
        d->launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    if (filter) k_hitsmap<true><<<grid, 256, 0, st>>>(d->view(), d->hm_rt, d->hm_pmax, r0, r1, v, d->hm_half, d_map, d->hm_n);
    else k_hitsmap<false><<<grid, 256, 0, st>>>(d->view(), d->hm_rt, d->hm_pmax, r0, r1, v, d->hm_half, d_map, d->hm_n);
    k_hitsmap_mirror<<<grid, 256, 0, st>>>(d->hm_half, d->nfiles, d_map);
    d->launches += 2;
    CK(cudaGetLastError());
    return 0;
}

int64_t igdb_dev_hitsmap_host(igdb_dev *d, int filter, int32_t v, int64_t t0, int64_t t1, uint32_t *hitmap)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    cudaStream_t st = d->stream;
    const size_t cells = (size_t)d->nfiles * d->nfiles;
    if (cells > d->hm_map_cells) {
        if (d->hm_map) CK(cudaFree(d->hm_map));
        d->hm_map = nullptr, d->hm_map_cells = 0;
        CK(cudaMalloc((void **)&d->hm_map, std::max<size_t>(cells, 1) * 4));
        d->hm_map_cells = cells;
    }
    CK(cudaMemcpyAsync(d->hm_map, hitmap, cells * 4, cudaMemcpyHostToDevice, st));
    if (igdb_dev_hitsmap_device(d, filter, v, t0, t1, d->hm_map, st)) return -1;
    unsigned long long h_n = 0;
    CK(cudaMemcpyAsync(hitmap, d->hm_map, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h_n, d->hm_n, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return (int64_t)h_n;
}

}  // extern "C"
