// igd_b200 CUDA layer -- the HBM image of a database: create, upload (AoS -> SoA + packed copy), finalize, destroy.
#include "igd_dev_impl.cuh"

namespace igdb {
thread_local char g_err[512] = "";
}

namespace {

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
// AoS gdata_t {idx,start,end,value} (W = 4, src/igd_base.h:41-46) or gdata0_t {idx,start,end} (W = 3, :48-52:
// gType-0 databases, value reads as 0) -> padded SoA.  toff[T+1] = unpadded record prefix (file order).
template <int W>
__global__ void k_transpose(const int32_t *__restrict__ aos, int64_t rec0, int64_t nrec,
                            const int64_t *__restrict__ toff, int32_t T, const TileMeta *__restrict__ tiles,
                            int32_t *__restrict__ start, int32_t *__restrict__ end, int32_t *__restrict__ idx,
                            int32_t *__restrict__ value, int32_t nbp, uint32_t nfiles, uint32_t *__restrict__ bad)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrec) return;
    int64_t g = rec0 + i;
    int32_t lo = 0, hi = T;                 // last t with toff[t] <= g
    while (hi - lo > 1) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(&toff[mid]) <= g) lo = mid; else hi = mid;
    }
    int4 r;
    if (W == 4) r = __ldg(reinterpret_cast<const int4 *>(aos) + i);
    else r = make_int4(__ldg(&aos[3 * i]), __ldg(&aos[3 * i + 1]), __ldg(&aos[3 * i + 2]), 0);
    if ((uint32_t)r.x >= nfiles) {              // would index hits[] out of bounds: the open fails (igdb_dev_finalize)
        atomicOr(bad, 1u);
        r.x = 0;
    }
    int64_t dst = (int64_t)tiles[lo].off8 * 8 + (g - __ldg(&toff[lo]));
    idx[dst] = r.x, start[dst] = r.y, end[dst] = r.z, value[dst] = r.w;
}

// The count-mode copy of every tile (rb / rv: 16-bit, relative to the tile's bin -- see pk_rel), written from the
// SoA arrays once they are complete.  Every record becomes a START entry and an END entry (igd_dev_impl.cuh).
// Entry e of a tile lies in group e / 256; a group of nvp entries (256, or what is left of the tile rounded up to 8)
// is one block of 4 nvp bytes: nvp keys, then nvp file indices.
__device__ __forceinline__ int64_t rb_key_index(int64_t base16, int32_t n, int32_t e, int32_t *nvp)
{
    const int32_t g = e >> 8, n8 = (n + 7) & ~7;
    *nvp = min(n8 - (g << 8), GROUP);
    return base16 + (int64_t)g * (2 * GROUP) + (e & (GROUP - 1));
}

// START half: the tile in the file's order, which is start order (src/igd_base.c:426-458 sorts every tile by start
// before it is written; a clamped start is monotone in the start).  One warp per tile; pads included (key 0x8000,
// file = nfiles: the trash slot of the histogram).  tiles[].ts still holds the bin boundary here (k_tile_meta runs
// afterwards).
__global__ void __launch_bounds__(256) k_pack_start(const TileMeta *__restrict__ tiles, int32_t T,
                                                     const int32_t *__restrict__ start, const int32_t *__restrict__ idx,
                                                     const int32_t *__restrict__ value, int32_t nbp, int32_t nfiles,
                                                     int fshift, uint16_t *__restrict__ rb, uint16_t *__restrict__ rv)
{
    const int32_t t = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const TileMeta m = tiles[t];
    const int64_t base = (int64_t)m.off8 * 8;
    const int32_t n = m.n, bd = m.ts, n8 = (n + 7) & ~7;
    for (int32_t i = lane; i < n8; i += 32) {
        int32_t nvp;
        const int64_t k = rb_key_index(base * 2, n, i, &nvp);
        const bool real = i < n;
        rb[k] = real ? (uint16_t)(pk_rel(start[base + i], bd, nbp) + 1) : (uint16_t)0x8000;
        rb[k + nvp] = (uint16_t)((real ? idx[base + i] : nfiles) << fshift);
        rv[base + i] = real ? (uint16_t)(min(max(value[base + i], 0), 32767) + 0x8000) : (uint16_t)0x8000;
    }
}

// END half: the tile's records ordered by their clamped end -- a counting sort over the nbp + 2 possible keys in
// shared memory, one CTA per tile at a time (any tile size; the order among equal ends is irrelevant, count mode
// adds them up).  eo = first entry of the END half.  Pads: key 0x7fff, file = nfiles.
__global__ void __launch_bounds__(SCAN_T) k_pack_end(const TileMeta *__restrict__ tiles, int32_t T,
                                                      const int32_t *__restrict__ end, const int32_t *__restrict__ idx,
                                                      const int32_t *__restrict__ value, int32_t nbp, int32_t nfiles,
                                                      int fshift, int64_t eo, uint16_t *__restrict__ rb,
                                                      uint16_t *__restrict__ rv)
{
    extern __shared__ uint32_t bucket[];                        // [nbp + 2]
    const int32_t nb = nbp + 2, per = (nb + SCAN_T - 1) / SCAN_T;
    for (int32_t t = blockIdx.x; t < T; t += gridDim.x) {
        const TileMeta m = tiles[t];
        const int32_t n = m.n, bd = m.ts, n8 = (n + 7) & ~7;
        if (n <= 0) continue;
        const int64_t base = (int64_t)m.off8 * 8;
        for (int32_t k = threadIdx.x; k < nb; k += SCAN_T) bucket[k] = 0;
        __syncthreads();
        for (int32_t i = threadIdx.x; i < n; i += SCAN_T) atomicAdd(&bucket[pk_rel_end(end[base + i], bd, nbp)], 1u);
        __syncthreads();
        // exclusive scan of the buckets: every thread owns `per` consecutive ones
        const int32_t k0 = min((int32_t)threadIdx.x * per, nb), k1 = min(k0 + per, nb);
        uint32_t sum = 0;
        for (int32_t k = k0; k < k1; ++k) sum += bucket[k];
        uint32_t tot, ex = block_exclusive<uint32_t>(sum, &tot);
        for (int32_t k = k0; k < k1; ++k) {
            const uint32_t c = bucket[k];
            bucket[k] = ex;
            ex += c;
        }
        __syncthreads();
        for (int32_t i = threadIdx.x; i < n8; i += SCAN_T) {
            const bool real = i < n;
            const int32_t e = real ? pk_rel_end(end[base + i], bd, nbp) : 0x7fff;
            const int32_t pos = real ? (int32_t)atomicAdd(&bucket[e], 1u) : i;
            int32_t nvp;
            const int64_t k = rb_key_index((eo + base) * 2, n, pos, &nvp);
            rb[k] = (uint16_t)e;
            rb[k + nvp] = (uint16_t)((real ? idx[base + i] : nfiles) << fshift);
            rv[eo + base + pos] = real ? (uint16_t)(min(max(value[base + i], 0), 32767) + 0x8000) : (uint16_t)0x8000;
        }
        __syncthreads();
    }
}

// Inside every group of 256 entries (both halves), order the real entries by (file mod 32) -- the bank of hist[file]
// in shared memory -- and record the group's key range.  k_count_rank gives lane l entries 8 l .. 8 l + 7 and issues
// one shared-memory atomic per k for the 32 entries {8 l + k}: 32 positions 8 apart in bank order, i.e. about one
// per bank (measured on B200: 3.6 -> ~2 wavefronts per ATOMS).  One warp per tile.
__global__ void __launch_bounds__(256) k_group_banksort(const TileMeta *__restrict__ tiles, int32_t T, int64_t eo,
                                                         const uint32_t *__restrict__ gbase, uint32_t ngroups,
                                                         int fshift, uint16_t *__restrict__ rb, uint16_t *__restrict__ rv,
                                                         uint32_t *__restrict__ gk)
{
    __shared__ uint16_t sk[8][GROUP], sf[8][GROUP], sv[8][GROUP];
    __shared__ uint32_t cnt[8][32];
    const int32_t t = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (t >= T) return;
    const TileMeta m = tiles[t];
    const int32_t n = m.n, n8 = (n + 7) & ~7;
    if (n <= 0) return;
    for (int half = 0; half < 2; ++half) {
        const int64_t base = (int64_t)m.off8 * 8 + (half ? eo : 0);
        for (int32_t g = 0; g * GROUP < n; ++g) {
            const int32_t nvp = min(n8 - g * GROUP, GROUP), nv = min(n - g * GROUP, GROUP);
            uint16_t *kb = rb + base * 2 + (int64_t)g * (2 * GROUP), *fb = kb + nvp, *vb = rv + base + (int64_t)g * GROUP;
            cnt[w][lane] = 0;
            __syncwarp();
            uint32_t kmin = 0xffffu, kmax = 0;
            for (int32_t i = lane; i < nv; i += 32) {
                const uint32_t k = kb[i];
                kmin = min(kmin, k), kmax = max(kmax, k);
                atomicAdd(&cnt[w][(fb[i] >> fshift) & 31], 1u);
            }
            kmin = __reduce_min_sync(FULL, kmin), kmax = __reduce_max_sync(FULL, kmax);
            __syncwarp();
            uint32_t c = cnt[w][lane], inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += o;
            }
            __syncwarp();
            cnt[w][lane] = inc - c;                             // first position of bank `lane`
            __syncwarp();
            for (int32_t i = lane; i < nv; i += 32) {
                const uint16_t f = fb[i];
                const uint32_t pos = atomicAdd(&cnt[w][(f >> fshift) & 31], 1u);
                sk[w][pos] = kb[i], sf[w][pos] = f, sv[w][pos] = vb[i];
            }
            __syncwarp();
            for (int32_t i = lane; i < nv; i += 32) kb[i] = sk[w][i], fb[i] = sf[w][i], vb[i] = sv[w][i];
            if (lane == 0) gk[gbase[t] + (uint32_t)g + (half ? ngroups : 0u)] = kmin | kmax << 16;
            __syncwarp();
        }
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


}  // namespace

namespace igdb {

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

int upload_queries(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe, cudaStream_t st)
{
    if (ensure_query_cap(d, std::max<int64_t>(n, 1))) return -1;
    CK(cudaMemcpyAsync(d->q_ctg, ctg, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qs, qs, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d->q_qe, qe, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace igdb

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
    void *ptrs[] = {d->tiles, d->d_toff, d->start, d->end, d->idx, d->value, d->tbase, d->q_ctg,
                    d->q_qs, d->q_qe, d->tile_cnt, d->tile_off, d->nseg, d->qseg_off, d->bsum, d->segs,
                    d->seg_hits, d->seg_out_off, d->d_hits, d->cnt2, d->cseg_off, d->cseg_mid,
                    d->plan_agg, d->plan_flag, d->plan_tot, d->csegs, d->items, d->items2, d->tplan, d->qrank, d->shard_scratch, d->hm_rt, d->hm_pmax, d->hm_half, d->hm_map, d->hm_n, d->hm_frec, d->hm_fstart, d->seq_pool, d->seq_tab, d->rb, d->rv, d->gbase, d->gk,
                    d->ctg_names, d->ctg_slots, d->d_text, d->d_starts};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int k = 0; k < 2; ++k) {
        if (d->up_stage[k]) cudaFree(d->up_stage[k]);
        if (d->up_ev[k]) cudaEventDestroy(d->up_ev[k]);
    }
    if (d->bad_idx) cudaFree(d->bad_idx);
    if (d->fp_qoff) cudaFree(d->fp_qoff);
    if (d->fp_out) cudaFree(d->fp_out);
    if (d->pinned) cudaFreeHost(d->pinned);
    if (d->arena) cudaFreeHost(d->arena);
    if (d->h_scalar) cudaFreeHost(d->h_scalar);
    for (auto &e : d->plan_ev)
        if (e) cudaEventDestroy(e);
    if (d->plan_done) cudaEventDestroy(d->plan_done);
    if (d->side) cudaStreamDestroy(d->side);
    for (auto &s : d->slot) {
        if (s.q) cudaFree(s.q);
        if (s.d_hits) cudaFree(s.d_hits);
        if (s.h_hits) cudaFreeHost(s.h_hits);
        if (s.h2d) cudaEventDestroy(s.h2d);
        if (s.done) cudaEventDestroy(s.done);
    }
    if (d->copy) cudaStreamDestroy(d->copy);
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
    CK(cudaMallocHost((void **)&d->h_scalar, 128));
    d->h_plan = reinterpret_cast<volatile unsigned long long *>(d->h_scalar + 8);      // 2 slots x 4 words
    for (auto &e : d->plan_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->plan_done, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking));

    std::vector<int32_t> tbase(d->nctg + 1, 0);
    for (int c = 0; c < d->nctg; ++c) tbase[c + 1] = tbase[c] + ntile[c];
    d->T = tbase[d->nctg];
    std::vector<TileMeta> meta((size_t)d->T);
    std::vector<uint32_t> gbase((size_t)d->T + 1, 0);
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
            gbase[t] = (uint32_t)d->db_groups;
            d->db_groups += (n + GROUP - 1) / GROUP;
            d->max_tile_groups = std::max<int64_t>(d->max_tile_groups, (n + GROUP - 1) / GROUP);
            p8 += (n + 7) / 8;
        }
    if (p8 >= (int64_t)0x7ffffff0) {             // the END half of the count-mode copy starts at p8 + 1
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
    d->packed = d->nbp <= 16384 && d->nfiles <= 65535 && !getenv("IGD_B200_NO_PACKED");
    // with few enough files the count-mode copy keeps 4 * file -- the byte offset of the file's counter in the shared
    // histogram -- so the search kernel spends one instruction per entry on the address instead of two
    d->bulk = getenv("IGD_B200_BULK") ? atoi(getenv("IGD_B200_BULK")) != 0 : false;
    d->fshift = d->nfiles + 1 <= 16383 && !getenv("IGD_B200_NO_PRESCALE") ? 2 : 0;
    const size_t half = (size_t)d->P + 8;       // START half, END half (igd_dev_impl.cuh)
    if (d->db_groups >= (int64_t)0x7fffffff) {
        snprintf(g_err, sizeof g_err, "database too large for 32-bit group numbers");
        return -1;
    }
    if (d->packed && (dev_alloc(d, &d->rb, 4 * half) || dev_alloc(d, &d->rv, 2 * half) ||
                      dev_alloc(d, &d->gbase, (size_t)d->T + 1) || dev_alloc(d, &d->gk, (size_t)2 * d->db_groups + 1)))
        return -1;
    if (d->packed) CK(cudaMemcpy(d->gbase, gbase.data(), gbase.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->tiles, meta.data(), meta.size() * sizeof(TileMeta), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->d_toff, d->toff.data(), d->toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->tbase, tbase.data(), tbase.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    // pad records (tiles are padded to a multiple of 4): start = INT_MAX, end = INT_MIN -- overlap no query
    const unsigned fb = (unsigned)std::min<int64_t>(((int64_t)d->P + 8 + 255) / 256, 148 * 32);
    k_fill<<<fb, 256>>>(d->start, (int64_t)d->P + 8, IMAX);
    k_fill<<<fb, 256>>>(d->end, (int64_t)d->P + 8, IMIN);
    k_fill<<<fb, 256>>>(d->value, (int64_t)d->P + 8, IMIN);
    if (d->packed) {
        // (every entry of a tile, pads included, is written by k_pack_start / k_pack_end; this covers the slack)
        k_fill16<<<fb, 256>>>(d->rb, (int64_t)(4 * half), (uint16_t)0x8000);
        k_fill16<<<fb, 256>>>(d->rv, (int64_t)(2 * half), (uint16_t)0x8000);
    }
    CK(cudaGetLastError());
    CK(cudaMemset(d->idx, 0, ((size_t)d->P + 8) * 4));
    CK(cudaMemset(d->tile_cnt, 0, ((size_t)d->T + 1) * 4));
    CK(cudaMemset(d->cnt2, 0, ((size_t)2 * d->T + 2) * 4));
    CK(cudaMemset(d->plan_flag, 0, (size_t)4 * d->sm_count * 4));
    d->plan_grid = d->plan_grid_max = 0;        // chosen at the first count batch (igd_dev_count.cu)
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
void *igdb_dev_stream(igdb_dev *d) { return (void *)d->stream; }

// plain device memory for the host C layer (the sharded step's cost / cuts / partial-count buffers)
void *igdb_dev_malloc(igdb_dev *d, size_t bytes)
{
    void *p = nullptr;
    if (cudaSetDevice(d->device) != cudaSuccess || cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, bytes ? bytes : 1) != cudaSuccess) {
        cudaFree(p);
        return nullptr;
    }
    return p;
}
void igdb_dev_free(igdb_dev *d, void *p)
{
    if (!p) return;
    cudaSetDevice(d->device);
    cudaFree(p);
}
int igdb_dev_memset_async(igdb_dev *d, void *p, size_t bytes, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    CK(cudaMemsetAsync(p, 0, bytes, stream ? (cudaStream_t)stream : d->stream));
    return 0;
}

void igdb_host_free(void *p)
{
    if (p) cudaFreeHost(p);
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

// One chunk of the file's record payload: H2D into staging slot `slot` and the transpose, both asynchronous.  The
// host buffer and the slot are free again after igdb_dev_upload_wait(slot); the caller alternates two slots so that
// reading chunk k + 1 from the file overlaps the copy and transpose of chunk k.
int igdb_dev_upload_slot(igdb_dev *d, int slot, int64_t rec0, int64_t nrec, const void *aos, int rec_words)
{
    if (nrec <= 0) return 0;
    if ((rec_words != 3 && rec_words != 4) || slot < 0 || slot > 1) {
        snprintf(g_err, sizeof g_err, "upload: records of %d words (gdata_t = 4, gdata0_t = 3), slot %d", rec_words, slot);
        return -1;
    }
    CK(cudaSetDevice(d->device));
    const size_t bytes = (size_t)nrec * 4 * (size_t)rec_words;
    if (!d->up_ev[slot]) CK(cudaEventCreateWithFlags(&d->up_ev[slot], cudaEventDisableTiming));
    if (!d->bad_idx) {
        CK(cudaMalloc(&d->bad_idx, 4));
        CK(cudaMemsetAsync(d->bad_idx, 0, 4, d->stream));
    }
    if (bytes > d->up_cap[slot]) {
        if (d->up_busy[slot]) CK(cudaEventSynchronize(d->up_ev[slot]));
        if (d->up_stage[slot]) CK(cudaFree(d->up_stage[slot]));
        d->up_stage[slot] = nullptr, d->up_cap[slot] = 0;
        CK(cudaMalloc(&d->up_stage[slot], bytes));
        d->up_cap[slot] = bytes;
    }
    CK(cudaMemcpyAsync(d->up_stage[slot], aos, bytes, cudaMemcpyHostToDevice, d->stream));
    const int thr = 256;
    auto kern = rec_words == 4 ? k_transpose<4> : k_transpose<3>;
    kern<<<(unsigned)((nrec + thr - 1) / thr), thr, 0, d->stream>>>(
        (const int32_t *)d->up_stage[slot], rec0, nrec, d->d_toff, d->T, d->tiles, d->start, d->end, d->idx, d->value,
        d->nbp, (uint32_t)d->nfiles, d->bad_idx);
    d->launches += 1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(d->up_ev[slot], d->stream));
    d->up_busy[slot] = true;
    return 0;
}

int igdb_dev_upload_wait(igdb_dev *d, int slot)
{
    if (slot < 0 || slot > 1 || !d->up_busy[slot]) return 0;
    CK(cudaEventSynchronize(d->up_ev[slot]));
    d->up_busy[slot] = false;
    return 0;
}

int igdb_dev_upload(igdb_dev *d, int64_t rec0, int64_t nrec, const void *aos, int rec_words)
{
    if (igdb_dev_upload_slot(d, 0, rec0, nrec, aos, rec_words)) return -1;
    return igdb_dev_upload_wait(d, 0);          // the caller reuses its host buffer
}

int igdb_dev_finalize(igdb_dev *d)
{
    CK(cudaSetDevice(d->device));
    if (d->T > 0) {
        if (d->packed) {
            k_pack_start<<<(unsigned)(((int64_t)d->T * 32 + 255) / 256), 256, 0, d->stream>>>(
                d->tiles, d->T, d->start, d->idx, d->value, d->nbp, d->nfiles, d->fshift, d->rb, d->rv);
            const size_t shm = ((size_t)d->nbp + 2) * 4;
            CK(cudaFuncSetAttribute(k_pack_end, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
            k_pack_end<<<(unsigned)std::min<int64_t>(d->T, (int64_t)d->sm_count * 16), SCAN_T, shm, d->stream>>>(
                d->tiles, d->T, d->end, d->idx, d->value, d->nbp, d->nfiles, d->fshift, (int64_t)d->P + 8, d->rb, d->rv);
            k_group_banksort<<<(unsigned)(((int64_t)d->T * 32 + 255) / 256), 256, 0, d->stream>>>(
                d->tiles, d->T, (int64_t)d->P + 8, d->gbase, (uint32_t)d->db_groups, d->fshift, d->rb, d->rv, d->gk);
            d->launches += 3;
        }
        k_tile_meta<<<(d->T + 255) / 256, 256, 0, d->stream>>>(d->tiles, d->T, d->start);
        d->launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(d->stream));
    for (int k = 0; k < 2; ++k) {
        if (d->up_stage[k]) CK(cudaFree(d->up_stage[k]));
        d->up_stage[k] = nullptr, d->up_cap[k] = 0, d->up_busy[k] = false;
    }
    if (d->bad_idx) {
        uint32_t bad = 0;
        CK(cudaMemcpy(&bad, d->bad_idx, 4, cudaMemcpyDeviceToHost));
        CK(cudaFree(d->bad_idx));
        d->bad_idx = nullptr;
        if (bad) {
            snprintf(g_err, sizeof g_err, "record with file index outside _index.tsv");
            return -1;
        }
    }
    return 0;
}

// the records of one tile back on the host (get_binData, IGDr/src/igd_search.c:357-393): up to cap of each column
int igdb_dev_tile_read(igdb_dev *d, int64_t tile, int32_t cap, int32_t *idx, int32_t *start, int32_t *end,
                       int32_t *value)
{
    g_err[0] = 0;
    if (tile < 0 || tile >= d->T) return 0;
    CK(cudaSetDevice(d->device));
    const int64_t n = std::min<int64_t>(d->toff[tile + 1] - d->toff[tile], cap);
    if (n <= 0) return 0;
    const size_t o = (size_t)d->off8[tile] * 8, b = (size_t)n * 4;
    CK(cudaStreamSynchronize(d->stream));
    if (idx) CK(cudaMemcpy(idx, d->idx + o, b, cudaMemcpyDeviceToHost));
    if (start) CK(cudaMemcpy(start, d->start + o, b, cudaMemcpyDeviceToHost));
    if (end) CK(cudaMemcpy(end, d->end + o, b, cudaMemcpyDeviceToHost));
    if (value) CK(cudaMemcpy(value, d->value + o, b, cudaMemcpyDeviceToHost));
    return (int)n;
}

int64_t igdb_dev_bytes(const igdb_dev *d) { return d->bytes; }
int igdb_dev_packed(const igdb_dev *d) { return d->packed ? 1 : 0; }
int64_t igdb_dev_launches(const igdb_dev *d) { return d->launches; }


}  // extern "C"
