// igd_b200 CUDA layer -- count mode (-q, -v): segments, plan, the search kernels, the host pipeline.  The hot path.
#include "igd_dev_impl.cuh"

namespace {

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
// The count pass (SCATTER = false) leaves, per query, what the scatter pass needs to place the first-bin segment
// without another atomic or another look at the tile: qrank[q] = the segment's rank among its tile's first-bin
// segments (what the counting atomic returned), Q_NOSEG = no first-bin segment but later bins may have one,
// Q_NONE = nothing at all.
constexpr int CSEG_Q = 4, CSEG_T = 256;
constexpr int CSEG_MAXCTG = 1024;           // contig table kept in shared memory up to this many contigs
constexpr uint32_t Q_NOSEG = 0xffffffffu, Q_NONE = 0xfffffffeu;

template <bool V, int MODE, bool SCATTER, class F1, class FL>
__device__ __forceinline__ void cseg_walk(const DbView &db, int64_t n, const int32_t *__restrict__ ctg,
                                          const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                          uint32_t *__restrict__ qrank, uint32_t *__restrict__ abn, int abn_sign,
                                          const unsigned long long *__restrict__ dyn, F1 first, FL later)
{
    __shared__ int32_t tb_s[CSEG_MAXCTG + 1];
    const bool in_smem = db.nctg <= CSEG_MAXCTG;
    pdl_launch_dependents();
    if (in_smem) {
        for (int i = threadIdx.x; i <= db.nctg; i += blockDim.x) tb_s[i] = __ldg(&db.tbase[i]);
        __syncthreads();
    }
    const int nbp = db.nbp;
    pdl_wait();                                             // the queries / counters of the previous launch
    if (dyn) n = min(n, (int64_t)dyn[0]);                   // the batch's size lives on the device (a dispatched shard)
    int nabn = 0;
    // (grid-stride: the launch may bring fewer CTAs than blocks of CSEG_T * CSEG_Q queries)
    for (int64_t q0 = (int64_t)blockIdx.x * (CSEG_T * CSEG_Q) + threadIdx.x; q0 - threadIdx.x < n;
         q0 += (int64_t)gridDim.x * (CSEG_T * CSEG_Q)) {
    int c[CSEG_Q], s[CSEG_Q], e[CSEG_Q], tile[CSEG_Q], nlast[CSEG_Q], bd[CSEG_Q];
    uint32_t r[CSEG_Q];
    bool ok[CSEG_Q];
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        const int64_t q = q0 + (int64_t)k * CSEG_T;
        ok[k] = q < n;
        c[k] = ok[k] ? ctg[q] : -1, s[k] = ok[k] ? qs[q] : 0, e[k] = ok[k] ? qe[q] : 0;
        r[k] = SCATTER && ok[k] ? qrank[q] : Q_NONE;
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
    if (!SCATTER) {
        // A query that stays inside ONE bin (most do) needs no look at its tile here: if the tile is empty, or the
        // query ends before the tile's first start (the gate of :477), its segment finds nothing -- an empty tile has
        // no work items, a gated query is past no start -- so it is simply counted; the first-bin-empty rule (:468)
        // only matters to a query that would go on to later bins.
        int4 m[CSEG_Q];
        bool one[CSEG_Q];
#pragma unroll
        for (int k = 0; k < CSEG_Q; ++k) {
            one[k] = MODE == M_PACKED && nlast[k] == tile[k];
            m[k] = ok[k] && !one[k] ? __ldg(reinterpret_cast<const int4 *>(&db.tiles[tile[k]])) : make_int4(0, 1, 0, IMIN);
        }
#pragma unroll
        for (int k = 0; k < CSEG_Q; ++k) {
            if (!V && m[k].y == 0) ok[k] = false;                           // first-bin-empty quirk (:468)
            r[k] = ok[k] ? Q_NOSEG : Q_NONE;
            if (ok[k] && m[k].y > 0 && e[k] > m[k].w) r[k] = first(tile[k], s[k], e[k], bd[k], 0u);   // nCnt > 0 and qe > start[0]
        }
#pragma unroll
        for (int k = 0; k < CSEG_Q; ++k) {
            const int64_t q = q0 + (int64_t)k * CSEG_T;
            if (q < n) qrank[q] = r[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < CSEG_Q; ++k) {
            ok[k] = ok[k] && r[k] != Q_NONE;
            if (ok[k] && r[k] != Q_NOSEG) first(tile[k], s[k], e[k], bd[k], r[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < CSEG_Q; ++k) {
        if (!ok[k]) continue;
        for (int t = tile[k] + 1; t <= nlast[k]; ++t) {                     // later bins (:500-531)
            const int4 mm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
            if (mm.y > 0 && e[k] > mm.w) later(t, s[k], e[k], bd[k] + (t - tile[k]) * nbp);
        }
    }
    }
    if (MODE == M_PACKED && nabn) atomicAdd(abn, (uint32_t)(abn_sign * nabn));   // left to the M_ABN32 pipeline
}

template <bool V, int MODE>
__global__ void __launch_bounds__(CSEG_T) k_cseg_count(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                        const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                        uint32_t *__restrict__ cnt2, uint32_t *__restrict__ qrank,
                                                        const unsigned long long *__restrict__ dyn)
{
    cseg_walk<V, MODE, false>(
        db, n, ctg, qs, qe, qrank, &cnt2[2 * db.ntiles], +1, dyn,
        [&](int tile, int, int, int, uint32_t) { return atomicAdd(&cnt2[2 * tile], 1u); },
        [&](int tile, int, int, int) { atomicAdd(&cnt2[2 * tile + 1], 1u); });
}

// the tile range of a batch whose shape lives on the device: dyn = {queries, tile_lo, tile_hi}
__device__ __forceinline__ void dyn_range(const unsigned long long *__restrict__ dyn, int32_t T, int32_t &tlo, int32_t &thi)
{
    if (!dyn) return;
    tlo = (int32_t)min(dyn[1], (unsigned long long)T);
    thi = max(tlo, (int32_t)min(dyn[2], (unsigned long long)T));
}

constexpr uint32_t RK_END = 0x1ff;      // an END item of k_count_rank carries this in the tsr field of its descriptor

__device__ __forceinline__ Plan plan_of_tile(const DbView &db, const uint32_t *__restrict__ cnt2, int32_t t, int rank,
                                             unsigned long long *chunks = nullptr)
{
    const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cnt2) + t);
    const unsigned long long s = (unsigned long long)c.x + c.y;
    if (chunks) *chunks = 0;
    if (s == 0) return Plan{0, 0};
    const unsigned long long g = ((unsigned long long)(uint32_t)db.tiles[t].n + GROUP - 1) / GROUP;
    const unsigned long long ch = (s + SMAX - 1) / SMAX + (rank ? ((unsigned long long)c.x + SMAX - 1) / SMAX : 0);
    if (chunks) *chunks = ch;
    // a tile's segments start at an even index: 8-byte segments, 16-byte aligned bulk copies
    return Plan{(s + 1) & ~1ull, g * ch};
}

// The whole plan in ONE launch: a scan over the tiles fused with the item fill.  The grid is small enough that all
// of its 512-thread CTAs are resident at once; CTA b owns a contiguous range of tiles:
//   1. block-reduce the plan of its range, publish it (agg[b]) and take a ticket;
//   2. the CTA that draws the last ticket turns the aggregates into exclusive prefixes, writes the batch totals
//      and raises ONE flag (= epoch of this launch) that every CTA waits on with one thread -- if the totals do
//      not fit the scratch buffers nothing more is written (the host grows them and runs the batch again);
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
//   w = 0 (rank != 0: the group's number, see gk[])
// An END item of k_count_rank (rank != 0) has x inside the END half of the count-mode copy, RK_END in the second
// field of z, and only the first-bin segments of its chunk.

__global__ void __launch_bounds__(SCAN_T) k_plan(DbView db, const uint32_t *cnt2, uint32_t *cnt2w, uint32_t epoch,
                                                  volatile PlanAgg *agg, volatile uint32_t *flag,
                                                  uint32_t *ticket, uint32_t *__restrict__ seg_off, uint32_t *__restrict__ seg_mid,
                                                  unsigned long long *__restrict__ tot, unsigned long long seg_cap,
                                                  unsigned long long item_cap, int4 *__restrict__ items, int rank,
                                                  int4 *__restrict__ tplan, int32_t tlo, int32_t thi,
                                                  const unsigned long long *__restrict__ dyn)
{
    __shared__ unsigned long long red[3][SCAN_T / 32];
    __shared__ unsigned long long sh_prefix[2], sh_total[3];
    __shared__ bool sh_last;
    const int32_t T = db.ntiles;
    const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();                                             // the segment counters
    // only the tiles [tlo, thi) can have segments (a shard of a multi-GPU search: igd_dev_shard.cu; else all of them)
    dyn_range(dyn, T, tlo, thi);
    const int32_t per = (thi - tlo + G - 1) / G, t0 = min(tlo + b * per, thi), t1 = min(t0 + per, thi);

    // 1. aggregate of this CTA's range
    unsigned long long segs = 0, its = 0, recs = 0;
    for (int32_t t = t0 + threadIdx.x; t < t1; t += SCAN_T) {
        unsigned long long ch;
        const Plan p = plan_of_tile(db, cnt2, t, rank, &ch);
        segs += p.segs, its += p.items;
        if (p.items) recs += (unsigned long long)(uint32_t)db.tiles[t].n * ch;
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
        __threadfence();
        sh_last = atomicAdd(ticket, 1u) == (uint32_t)G - 1;         // the last CTA to get here scans for everybody
    }
    __syncthreads();
    // 2. the last CTA turns the aggregates into exclusive prefixes (G <= SCAN_T: one per thread), writes the batch
    //    totals and raises the flag; every CTA waits on that ONE flag with ONE thread
    if (sh_last) {
        __threadfence();
        const bool in = (int)threadIdx.x < G;
        const Plan mine = in ? Plan{agg[threadIdx.x].segs, agg[threadIdx.x].items} : Plan{0, 0};
        unsigned long long r = in ? agg[threadIdx.x].recs : 0ull;
        Plan all;
        const Plan ex = block_exclusive<Plan>(mine, &all);
        if (in) agg[threadIdx.x].segs = ex.segs, agg[threadIdx.x].items = ex.items;     // now: prefixes
#pragma unroll
        for (int d = 16; d; d >>= 1) r += __shfl_xor_sync(FULL, r, d);
        if (lane == 0) red[2][wid] = r;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long z = 0;
            for (int w = 0; w < SCAN_T / 32; ++w) z += red[2][w];
            tot[0] = all.segs, tot[1] = all.items, tot[2] = z;
            tot[3] = cnt2[2 * T];                          // queries left to the M_ABN32 pipeline (M_PACKED only)
            *ticket = 0;                                    // ready for the next launch
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        }
    }
    if (threadIdx.x == 0) {
        for (;;) {                                          // acquire: prefixes and totals are visible with the flag
            uint32_t f;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
            if (f == epoch) break;
            __nanosleep(40);
        }
        sh_prefix[0] = agg[b].segs, sh_prefix[1] = agg[b].items;
        const volatile unsigned long long *vt = tot;
        sh_total[0] = vt[0], sh_total[1] = vt[1], sh_total[2] = vt[2];
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
        const Plan mine = in ? plan_of_tile(db, cnt2, t, rank) : Plan{0, 0};
        Plan ctot;
        const Plan ex = carry + block_exclusive<Plan>(mine, &ctot);
        carry += ctot;
        uint32_t i0 = (uint32_t)ex.items, ni = (uint32_t)mine.items, sa0 = (uint32_t)ex.segs, mid = 0, sb0 = 0;   // (ni: items this thread writes)
        int4 tm = make_int4(0, 0, 0, 0);
        if (in) {
            const uint2 c = __ldg(reinterpret_cast<const uint2 *>(cnt2) + t);
            if (c.x) cnt2w[2 * t] = 0;                      // (the scatter pass counts only the later-bin ones back down)
            mid = sa0 + c.x, sb0 = mid + c.y;
            seg_off[t] = sa0, seg_mid[t] = mid;
            if (rank) {                                     // k_windows makes the descriptors of this tile's items
                tplan[2 * t] = make_int4((int)sa0, (int)mid, (int)sb0, (int)i0);
                if (sb0 != sa0) {                           // everything else it needs about the tile, in one place
                    const int4 tq = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
                    tplan[2 * t + 1] = make_int4(tq.x, tq.y, tq.z, (int)__ldg(&db.gbase[t]));
                }
                ni = 0;
            }
            if (ni) {
                tm = __ldg(reinterpret_cast<const int4 *>(&db.tiles[t]));
                tm.w = rank ? (int)__ldg(&db.gbase[t]) : 0;     // (s0 is not needed here)
            }
        }
        auto write = [&](uint32_t k, uint32_t i0_, uint32_t a, uint32_t m, uint32_t e, int4 tm_) {   // tm_.w = first group
            const uint32_t ch = (e - a + SMAX - 1) / SMAX;
            const uint32_t ng = ((uint32_t)tm_.y + GROUP - 1) / GROUP;
            const bool endk = k >= ng * ch;                 // the END items follow the START items (rank only)
            const uint32_t kk = endk ? k - ng * ch : k, chx = endk ? (m - a + SMAX - 1) / SMAX : ch;
            const uint32_t g = kk / chx, c = kk - g * chx;
            const uint32_t sa = a + c * SMAX, sb = min(sa + (uint32_t)SMAX, endk ? m : e);
            const int r0 = (int)(g * GROUP);
            const uint32_t nv = (uint32_t)min(tm_.y - r0, GROUP);
            const uint32_t tsr = endk ? RK_END : (uint32_t)min(max(tm_.z - r0, 0), GROUP);
            const uint32_t nfirst = min(max(m, sa), sb) - sa;
            items[i0_ + k] = make_int4((int)((uint32_t)tm_.x + g * (GROUP / 8) + (endk ? db.eoff8 : 0u)), (int)sa,
                                       (int)(nv | tsr << 9 | nfirst << 18 | (sb - sa) << 25),
                                       (int)((uint32_t)tm_.w + g + (endk ? db.ngroups : 0u)));
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
                                       __shfl_sync(FULL, tm.z, src), __shfl_sync(FULL, tm.w, src));
            for (uint32_t k = lane; k < ni_; k += 32) write(k, i0_, a_, m_, e_, tm_);
        }
    }
}

// The later-bin counters of cnt2 are counted back down to zero here (the first-bin ones were reset by k_plan), so
// cnt2 is clean for the next batch.  Nothing is written when the plan does not fit the scratch buffers: the host
// then grows them and runs the batch again.  A first-bin segment goes to its tile's first segment + the rank the
// count pass kept.
// M_PACKED writes 4-byte segments in the tile's clamped coordinates, biased for k_count_rank's subtractions:
//   bits 0-15 = qs' + 0x8000, bits 16-31 = qe' + 0x8000.
template <bool V, int MODE>
__global__ void __launch_bounds__(CSEG_T) k_cseg_scatter(DbView db, int64_t n, const int32_t *__restrict__ ctg,
                                                          const int32_t *__restrict__ qs, const int32_t *__restrict__ qe,
                                                          uint32_t *__restrict__ cnt2, uint32_t *__restrict__ qrank,
                                                          const uint32_t *__restrict__ seg_off,
                                                          const uint32_t *__restrict__ seg_mid,
                                                          const unsigned long long *__restrict__ tot,
                                                          unsigned long long seg_cap, unsigned long long item_cap,
                                                          void *__restrict__ segs, const unsigned long long *__restrict__ dyn)
{
    pdl_wait();                                             // the plan
    if (tot[0] > seg_cap || tot[1] > item_cap) return;
    const int nbp = db.nbp;
    auto put = [&](uint32_t pos, int s, int e, int bd) {
        if (MODE == M_PACKED)
            static_cast<uint32_t *>(segs)[pos] =
                (uint32_t)(pk_rel(s, bd, nbp) + 0x8000) | (uint32_t)(pk_rel_end(e, bd, nbp) + 0x8000) << 16;
        else
            static_cast<CSeg *>(segs)[pos] = CSeg{s, e};
    };
    // (the queries left to the M_ABN32 pipeline are counted back down too)
    cseg_walk<V, MODE, true>(
        db, n, ctg, qs, qe, qrank, &cnt2[2 * db.ntiles], -1, dyn,
        [&](int tile, int s, int e, int bd, uint32_t rank) {
            put(__ldg(&seg_off[tile]) + rank, s, e, bd);
            return 0u;
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
    const int4 *desc;                   // k_count_rank: 2 x int4 per item (k_windows_warp / k_windows_small / k_windows<32>)
    const CSeg *segs;
    const unsigned long long *tot;      // device: {segments, items} of this batch
    unsigned long long seg_cap, item_cap;
    int32_t v;
    unsigned long long *hits;
};

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
// count mode on the 16-bit START / END entries (the default path).
//
// For a record r of a tile and the tile's first-bin segments F (queries whose first bin this is) and later-bin
// segments L (queries that began in an earlier bin; src/igd_search.c:500-531 lets them count only the records that
// start inside this bin, index >= tS), in the tile's clamped coordinates (pk_rel):
//
//   hits(r) = #{F : start'_r < qe' and end'_r > qs'} + [start'_r >= 1] #{L : start'_r < qe'}
//           = #{F : qe' > start'_r} - #{F : qs' >= end'_r} + [start'_r >= 1] #{L : qe' > start'_r}
//
// (end'_r <= qs' implies start'_r < end'_r <= qs' < qe', so the second set lies inside the first).  The first and
// third term depend on the record's START only, the second on its END only: a record is searched as two
// independent entries, hits[file] += H(start') and hits[file] -= K(end'), with
//   H(s) = #{F u L : qe' > s} for s >= 1,  H(0) = |F|;      K(e) = #{F : qs' >= e}
// -- per-CTA sums are taken modulo 2^32 and flushed sign-extended.  Both halves of the count-mode copy are sorted
// by their key inside a tile (igd_dev_impl.cuh), so an item's 256 keys lie in [klo, khi] = [first, last] and a
// segment is on the far side of ALL of them (it goes into a base count), of NONE, or INSIDE THE WINDOW -- only
// those are compared entry by entry: with ~56 segments per 5 500-record tile (BASELINE config 3) 2-3 of them per
// item, where a pair test of records against segments meets 14-20.
//
//   entries  START: S = start' + 1 (1 .. nbp + 1; pads 0x8000)        END: E = end' (2 .. nbp + 1; pads 0x7fff)
//   segments 4 B: xe = qs' + 0x8000 | (xs = qe' + 0x8000) << 16;      x = xs (START items), xe (END items)
//   x - key >= 0x8000  <=>  qe' > start'  resp.  qs' >= end';   no 16-bit half ever borrows (tests/test_packed_math.py)
//   counts (two 16-bit counters per register, <= 64 per item) += ((x2 - key2) & 0x80008000) >> 15
// i.e. 1 subtraction, 1 LOP3 and 1 shift-add per PAIR of entries and segment inside the window.
//
// Which segments are in an item's window is the same for all 32 lanes, so it is not decided in the search kernel,
// where a warp-uniform instruction costs 32 lanes: the descriptor kernels below (k_windows_warp for dense tiles,
// k_windows_small for tiles of a few segments and groups, k_windows<32> for tiles of 128 groups or more) leave the
// base count and the window -- the x values themselves when there are <= RK_INL of them, else a 64-bit mask of
// the chunk's segments -- in the item's descriptor.
// ------------------------------------------------------------------------------------------------
constexpr int RK_MAXW = 32;             // most warps per CTA
constexpr int RK_IB = 8;                // items per block
constexpr uint32_t RK_INL = 4;          // segments inside an item's window that travel in its descriptor

// descriptor of k_count_rank, 2 x int4:
//   A.x = first entry of the item in the count-mode copy, in units of 8 entries     A.y = first segment
//   A.z = entries (bits 0-8) | segments (9-15) | base count (16-22) | later-bin segments (23-29) | END item (30)
//         | fix (31: a START item with entries before the bin AND later-bin segments -- see k_count_rank)
//   A.w = window mask of segments 0-31                                                B.x = of segments 32-63
//   B.w = number of segments inside the window; when there are <= RK_INL of them their x values themselves are in
//         B.y (first | second << 16) and B.z, and the search kernel touches neither the masks nor the segments
// tiles that k_windows_small takes (see there)
constexpr uint32_t WS_S = 16, WS_G = 8;
__device__ __forceinline__ bool windows_small_tile(uint32_t nseg, uint32_t ng) { return nseg <= WS_S && ng <= WS_G; }

// LPT lanes per tile (grid-stride).  Per (chunk of <= 64 segments,
// START / END): the groups' key ranges ascend, so a segment x is past ALL keys of the groups before its HOME group
// (the first one whose largest key it is not past), possibly inside the window of the home group, and past NO key
// of the groups after it.  The lanes take the segments (a binary search each over the tile's gk[]), leave one
// count per home group and one window bit in shared memory, and then take the groups: an item's base count is the
// number of segments whose home lies behind its group (a suffix sum), its window the bits left at its group.
// O(segments * log(groups) + groups) per chunk, where testing every segment against every item is their product.
template <int LPT>
__global__ void __launch_bounds__(256, 4) k_windows(DbView db, const int4 *__restrict__ tplan,
                                                  const uint32_t *__restrict__ segs,
                                                  const unsigned long long *__restrict__ tot,
                                                  unsigned long long seg_cap, unsigned long long item_cap,
                                                  int4 *__restrict__ out, int small_done, uint32_t tlo, uint32_t thi,
                                                  uint32_t min_groups, const unsigned long long *__restrict__ dyn)
{
    constexpr int NSUB = 256 / LPT, NGT = 4 * LPT;          // NGT: groups of a tile whose key ranges are staged in shared memory
    __shared__ uint32_t s_cnt[NSUB][LPT + 1], s_m0[NSUB][LPT], s_m1[NSUB][LPT], s_x[NSUB][SMAX], s_gk[NSUB][2][NGT];
    pdl_launch_dependents();
    pdl_wait();                                             // the plan, the segments
    if (tot[0] > seg_cap || tot[1] > item_cap) return;
    {
        int32_t lo_ = (int32_t)tlo, hi_ = (int32_t)thi;
        dyn_range(dyn, db.ntiles, lo_, hi_);
        tlo = (uint32_t)lo_, thi = (uint32_t)hi_;
    }
    static_assert(LPT == 32, "a warp per tile: the tiles are taken 32 at a time, one per lane, to see which need work");
    const uint32_t l = threadIdx.x % LPT, sub = threadIdx.x / LPT;
    const unsigned lmask = FULL;
    uint32_t *cnt = s_cnt[sub], *m0 = s_m0[sub], *m1 = s_m1[sub], *sx = s_x[sub];
    const uint32_t T = thi;
    // per warp `tpw` tiles at a time (32 when k_windows_small has taken most of them, else 1): lane i looks at tile
    // tb + i (coalesced), then the warp works through the ones that have items to describe
    const uint32_t tpw = small_done ? 32u : 1u;
    for (uint32_t tb = tlo + (blockIdx.x * NSUB + sub) * tpw; tb < T; tb += gridDim.x * NSUB * tpw) {
      int4 tp_l = make_int4(0, 0, 0, 0), tm_l = make_int4(0, 0, 0, 0);
      if (l < tpw && tb + l < T) {
          tp_l = __ldg(&tplan[2 * (tb + l)]);
          if (tp_l.z != tp_l.x) tm_l = __ldg(&tplan[2 * (tb + l) + 1]);
      }
      bool need = tp_l.z != tp_l.x && ((uint32_t)tm_l.y + GROUP - 1) / GROUP >= min_groups;   // (fewer: k_windows_warp's)
      if (need && small_done && windows_small_tile((uint32_t)(tp_l.z - tp_l.x), ((uint32_t)tm_l.y + GROUP - 1) / GROUP))
          need = false;                                     // k_windows_small has done this tile
      unsigned todo = __ballot_sync(FULL, need);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int4 tp = make_int4(__shfl_sync(FULL, tp_l.x, src), __shfl_sync(FULL, tp_l.y, src),
                                  __shfl_sync(FULL, tp_l.z, src), __shfl_sync(FULL, tp_l.w, src));
        const int4 tm = make_int4(__shfl_sync(FULL, tm_l.x, src), __shfl_sync(FULL, tm_l.y, src),
                                  __shfl_sync(FULL, tm_l.z, src), __shfl_sync(FULL, tm_l.w, src));   // off8, n, ts, first group
        const uint32_t a = (uint32_t)tp.x, m = (uint32_t)tp.y, e = (uint32_t)tp.z, i0 = (uint32_t)tp.w;
        const uint32_t gb = (uint32_t)tm.w;
        const uint32_t n = (uint32_t)tm.y, ng = (n + GROUP - 1) / GROUP;
        const uint32_t ch = (e - a + SMAX - 1) / SMAX, chF = (m - a + SMAX - 1) / SMAX;
        const bool staged = ng <= (uint32_t)NGT;            // the key ranges of the tile's groups, START and END
        if (staged) {
            for (uint32_t i = l; i < ng; i += LPT)
                s_gk[sub][0][i] = __ldg(&db.gk[gb + i]), s_gk[sub][1][i] = __ldg(&db.gk[gb + db.ngroups + i]);
            __syncwarp(lmask);
        }
        for (uint32_t c = 0; c < ch; ++c) {
            const uint32_t sa = a + c * SMAX, ns_all = min(sa + (uint32_t)SMAX, e) - sa;
            const uint32_t nf = min(max(m, sa), sa + ns_all) - sa;         // first-bin segments of the chunk
            for (uint32_t kind = 0; kind < (c < chF ? 2u : 1u); ++kind) {
                const uint32_t ns = kind ? nf : ns_all, sh = kind ? 0u : 16u;
                const uint32_t *gk = staged ? s_gk[sub][kind] : db.gk + gb + (kind ? db.ngroups : 0u);
                for (uint32_t g0 = 0; g0 < ng; g0 += LPT) {
                    cnt[l] = 0, m0[l] = 0, m1[l] = 0;
                    if (l == 0) cnt[LPT] = 0;
                    __syncwarp(lmask);
                    for (uint32_t j = l; j < ns; j += LPT) {
                        const uint32_t x = (__ldg(&segs[sa + j]) >> sh) & 0xffffu;
                        if (g0 == 0) sx[j] = x;
                        uint32_t lo = 0, hi = ng;           // home = #groups whose largest key x is past
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (x >= (gk[mid] >> 16) + 0x8000u) lo = mid + 1;
                            else hi = mid;
                        }
                        if (lo >= g0 + LPT) atomicAdd(&cnt[LPT], 1u);
                        else if (lo >= g0) {
                            atomicAdd(&cnt[lo - g0], 1u);
                            if (lo < ng && x >= (gk[lo] & 0xffffu) + 0x8000u)      // past the home group's smallest key
                                atomicOr(j < 32 ? &m0[lo - g0] : &m1[lo - g0], 1u << (j & 31));
                        }
                    }
                    __syncwarp(lmask);
                    // lane l <-> group g0 + l: base = segments whose home lies behind it
                    uint32_t suf = cnt[l + 1];
#pragma unroll
                    for (int d = 1; d < LPT; d <<= 1) {
                        const uint32_t o = __shfl_down_sync(lmask, suf, d, LPT);
                        if (l + d < LPT) suf += o;
                    }
                    const uint32_t g = g0 + l;
                    if (g < ng) {
                        const uint32_t mm0 = m0[l], mm1 = m1[l], nin = __popc(mm0) + __popc(mm1);
                        uint32_t v01 = 0, v23 = 0;
                        if (nin <= RK_INL) {
                            uint32_t q0 = mm0, q1 = mm1;
                            for (uint32_t r = 0; r < nin; ++r) {
                                uint32_t j;
                                if (q0) j = __ffs(q0) - 1, q0 &= q0 - 1;
                                else j = 32 + __ffs(q1) - 1, q1 &= q1 - 1;
                                if (r < 2) v01 |= sx[j] << (16 * r);
                                else v23 |= sx[j] << (16 * (r - 2));
                            }
                        }
                        const int r0 = (int)(g * GROUP);
                        const uint32_t nv = (uint32_t)min((int)n - r0, GROUP), tsr = (uint32_t)min(max(tm.z - r0, 0), GROUP);
                        const uint32_t nl = kind ? 0u : ns_all - nf, fix = !kind && tsr > 0 && nl > 0;
                        const size_t k = (size_t)i0 + (kind ? ng * ch + g * chF + c : g * ch + c);
                        out[2 * k] = make_int4((int)((uint32_t)tm.x + g * (GROUP / 8) + (kind ? db.eoff8 : 0u)), (int)sa,
                                               (int)(nv | ns << 9 | suf << 16 | nl << 23 | kind << 30 | fix << 31), (int)mm0);
                        out[2 * k + 1] = make_int4((int)mm1, (int)v01, (int)v23, (int)nin);
                    }
                    __syncwarp(lmask);
                }
            }
        }
        __syncwarp(lmask);
      }
    }
}

// The descriptors of the tiles of a dense database (BASELINE config 3: ~22 groups and ~56 segments per tile), one
// WARP per tile, tiles of fewer than W2_NG groups (the others: k_windows<32> above, launched only for a database
// that has such tiles).  Same result as k_windows<32> in two thirds of its instructions (138 M against 211 M warp
// instructions at BASELINE config 3) and 0.165 ms instead of 0.296:
//   * the groups' thresholds (largest key + 0x8000 padded with +inf to a power of two P, smallest key + 0x8000) are
//     staged once per tile, so a segment's HOME group -- the number of groups whose largest key it is past -- is a
//     branch-free search of log2(P) steps of 4 instructions;
//   * a lane searches for segments `lane` and `lane + 32` of the chunk at once (two independent chains), and both
//     kinds (START: x = qe', END: x = qs') use the same two loaded words;
//   * ONE shared-memory atomic per segment leaves both counts at its home group -- segments (low half) and segments
//     inside the window (high half); its return value is the slot of the segment's x among the <= RK_INL values
//     that travel in the descriptor, so the group phase reads them back with one load instead of walking a mask.
constexpr uint32_t W2_NG = 128;

__global__ void __launch_bounds__(256) k_windows_warp(DbView db, const int4 *__restrict__ tplan,
                                                       const uint32_t *__restrict__ segs,
                                                       const unsigned long long *__restrict__ tot,
                                                       unsigned long long seg_cap, unsigned long long item_cap,
                                                       int4 *__restrict__ out, int small_done, uint32_t tpw, uint32_t tlo,
                                                       uint32_t thi, const unsigned long long *__restrict__ dyn)
{
    __shared__ uint32_t s_hi[8][2][W2_NG], s_lo[8][2][W2_NG];
    __shared__ uint32_t s_cw[8][W2_NG], s_m0[8][W2_NG], s_m1[8][W2_NG];
    __shared__ uint2 s_wv[8][W2_NG];
    pdl_launch_dependents();
    pdl_wait();                                             // the plan, the segments
    if (tot[0] > seg_cap || tot[1] > item_cap) return;
    {
        int32_t lo_ = (int32_t)tlo, hi_ = (int32_t)thi;
        dyn_range(dyn, db.ntiles, lo_, hi_);
        tlo = (uint32_t)lo_, thi = (uint32_t)hi_;
    }
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *cw = s_cw[wib], *m0 = s_m0[wib], *m1 = s_m1[wib];
    uint16_t *wv = reinterpret_cast<uint16_t *>(s_wv[wib]);
    // `tpw` tiles per warp and turn (32 in a database of small tiles, most of which k_windows_small has done; else 8):
    // lane i looks at tile tb + i, then the warp works through the ones that are its to describe
    for (uint32_t tb = tlo + (blockIdx.x * 8 + wib) * tpw; tb < thi; tb += gridDim.x * 8 * tpw) {
      bool need = false;
      if (lane < tpw && tb + lane < thi) {
          const int4 tp_l = __ldg(&tplan[2 * (tb + lane)]);
          if (tp_l.z != tp_l.x) {                           // (else: no query reached this tile)
              const uint32_t ng_l = ((uint32_t)__ldg(&tplan[2 * (tb + lane) + 1]).y + GROUP - 1) / GROUP;
              need = ng_l < W2_NG &&                        // (else: k_windows<32>)
                     !(small_done && windows_small_tile((uint32_t)(tp_l.z - tp_l.x), ng_l));   // (else: k_windows_small)
          }
      }
      unsigned todo = __ballot_sync(FULL, need);
      while (todo) {
        const uint32_t t = tb + (uint32_t)__ffs(todo) - 1;
        todo &= todo - 1;
        const int4 tp = __ldg(&tplan[2 * t]), tm = __ldg(&tplan[2 * t + 1]);      // {segments .. first item}, {off8, n, ts, first group}
        const uint32_t a = (uint32_t)tp.x, m = (uint32_t)tp.y, e = (uint32_t)tp.z, i0 = (uint32_t)tp.w;
        const uint32_t gb = (uint32_t)tm.w, n = (uint32_t)tm.y, ng = (n + GROUP - 1) / GROUP;
        const uint32_t P = ng < 32 ? 32u : (ng < 64 ? 64u : 128u);     // > ng: the search can answer "past all groups"
        const uint32_t ch = (e - a + SMAX - 1) / SMAX, chF = (m - a + SMAX - 1) / SMAX;
        __syncwarp();
        for (uint32_t g = lane; g < P; g += 32) {
            const uint32_t ks = g < ng ? __ldg(&db.gk[gb + g]) : 0u, ke = g < ng ? __ldg(&db.gk[gb + db.ngroups + g]) : 0u;
            s_hi[wib][0][g] = g < ng ? (ks >> 16) + 0x8000u : 0xffffffffu, s_lo[wib][0][g] = (ks & 0xffffu) + 0x8000u;
            s_hi[wib][1][g] = g < ng ? (ke >> 16) + 0x8000u : 0xffffffffu, s_lo[wib][1][g] = (ke & 0xffffu) + 0x8000u;
        }
        for (uint32_t c = 0; c < ch; ++c) {
            const uint32_t sa = a + c * SMAX, ns_all = min(sa + (uint32_t)SMAX, e) - sa;
            const uint32_t nf = min(max(m, sa), sa + ns_all) - sa;         // first-bin segments of the chunk
            const uint32_t w0 = lane < ns_all ? __ldg(&segs[sa + lane]) : 0u;
            const uint32_t w1 = lane + 32 < ns_all ? __ldg(&segs[sa + 32 + lane]) : 0u;
            for (uint32_t kind = 0; kind < (c < chF ? 2u : 1u); ++kind) {
                const uint32_t ns = kind ? nf : ns_all, sh = kind ? 0u : 16u;
                const uint32_t *hi = s_hi[wib][kind], *lo = s_lo[wib][kind];
                for (uint32_t g = lane; g < P; g += 32) cw[g] = 0, m0[g] = 0, m1[g] = 0;
                __syncwarp();
                const uint32_t x0 = (w0 >> sh) & 0xffffu, x1 = (w1 >> sh) & 0xffffu;
                uint32_t h0 = 0, h1 = 0;                    // home = #groups whose largest key x is past
                if (P == 32) {                              // (the common depth, unrolled: 4 instructions per step)
#pragma unroll
                    for (uint32_t step = 16; step; step >>= 1) {
                        if (x0 >= hi[h0 + step - 1]) h0 += step;
                        if (x1 >= hi[h1 + step - 1]) h1 += step;
                    }
                } else {
                    for (uint32_t step = P >> 1; step; step >>= 1) {
                        if (x0 >= hi[h0 + step - 1]) h0 += step;
                        if (x1 >= hi[h1 + step - 1]) h1 += step;
                    }
                }
                if (lane < ns) {
                    const bool in = h0 < ng && x0 >= lo[h0];        // past the home group's smallest key: in its window
                    const uint32_t slot = atomicAdd(&cw[h0], in ? 0x10001u : 1u) >> 16;
                    if (in) {
                        atomicOr(&m0[h0], 1u << lane);
                        if (slot < RK_INL) wv[h0 * 4 + slot] = (uint16_t)x0;
                    }
                }
                if (lane + 32 < ns) {
                    const bool in = h1 < ng && x1 >= lo[h1];
                    const uint32_t slot = atomicAdd(&cw[h1], in ? 0x10001u : 1u) >> 16;
                    if (in) {
                        atomicOr(&m1[h1], 1u << lane);
                        if (slot < RK_INL) wv[h1 * 4 + slot] = (uint16_t)x1;
                    }
                }
                __syncwarp();
                // lane <-> group g0 + lane, from the window that holds slot ng (the segments past every group) down:
                // an item's base count = the segments whose home lies behind its group
                uint32_t carry = 0;
                for (int g0 = (int)(ng & ~31u); g0 >= 0; g0 -= 32) {
                    const uint32_t g = (uint32_t)g0 + lane;
                    const uint32_t cwg = g < P ? cw[g] : 0u, cnt = cwg & 0xffffu, nin = cwg >> 16;
                    uint32_t suf = cnt;                     // inclusive suffix sum over the window
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t o = __shfl_down_sync(FULL, suf, d);
                        if (lane + d < 32) suf += o;
                    }
                    const uint32_t base = suf - cnt + carry;
                    carry += __shfl_sync(FULL, suf, 0);
                    if (g < ng) {
                        const uint2 v = s_wv[wib][g];
                        const int r0 = (int)(g * GROUP);
                        const uint32_t nv = (uint32_t)min((int)n - r0, GROUP), tsr = (uint32_t)min(max(tm.z - r0, 0), GROUP);
                        const uint32_t nl = kind ? 0u : ns_all - nf, fix = !kind && tsr > 0 && nl > 0;
                        const size_t k = (size_t)i0 + (kind ? ng * ch + g * chF + c : g * ch + c);
                        out[2 * k] = make_int4((int)((uint32_t)tm.x + g * (GROUP / 8) + (kind ? db.eoff8 : 0u)), (int)sa,
                                               (int)(nv | ns << 9 | base << 16 | nl << 23 | kind << 30 | fix << 31), (int)m0[g]);
                        out[2 * k + 1] = make_int4((int)m1[g], (int)v.x, (int)v.y, (int)nin);
                    }
                }
                __syncwarp();
            }
        }
      }
    }
}

// The same for SMALL tiles -- at most WS_S segments in all and at most WS_G groups, what a sparse database
// (BASELINE config 2: ~6 segments and 3 groups per tile) consists of: one THREAD per tile, segments and key ranges
// in registers, every (group, segment) pair tested.  k_windows<LPT> skips these tiles when `small_done` is set.

__global__ void __launch_bounds__(128) k_windows_small(DbView db, const int4 *__restrict__ tplan,
                                                        const uint32_t *__restrict__ segs,
                                                        const unsigned long long *__restrict__ tot,
                                                        unsigned long long seg_cap, unsigned long long item_cap,
                                                        int4 *__restrict__ out, uint32_t tlo, uint32_t thi,
                                                        const unsigned long long *__restrict__ dyn)
{
    pdl_launch_dependents();
    pdl_wait();                                             // the plan, the segments
    if (tot[0] > seg_cap || tot[1] > item_cap) return;
    {
        int32_t lo_ = (int32_t)tlo, hi_ = (int32_t)thi;
        dyn_range(dyn, db.ntiles, lo_, hi_);
        tlo = (uint32_t)lo_, thi = (uint32_t)hi_;
    }
    const uint32_t t = tlo + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= thi) return;
    const int4 tp = __ldg(&tplan[2 * t]);
    const uint32_t a = (uint32_t)tp.x, m = (uint32_t)tp.y, e = (uint32_t)tp.z, i0 = (uint32_t)tp.w;
    if (e == a) return;                                     // no query reached this tile
    const int4 tm = __ldg(&tplan[2 * t + 1]);               // the tile's off8, n, ts, first group
    const uint32_t n = (uint32_t)tm.y, ng = (n + GROUP - 1) / GROUP, ns_all = e - a, nf = m - a;
    if (!windows_small_tile(ns_all, ng)) return;
    uint32_t w[WS_S], ks[WS_G], ke[WS_G];
#pragma unroll
    for (uint32_t j = 0; j < WS_S; ++j) w[j] = j < ns_all ? __ldg(&segs[a + j]) : 0u;
#pragma unroll
    for (uint32_t g = 0; g < WS_G; ++g) {
        ks[g] = g < ng ? __ldg(&db.gk[(uint32_t)tm.w + g]) : 0u;
        ke[g] = g < ng && nf > 0 ? __ldg(&db.gk[(uint32_t)tm.w + db.ngroups + g]) : 0u;
    }
#pragma unroll
    for (uint32_t kind = 0; kind < 2; ++kind) {
        if (kind && nf == 0) break;
        const uint32_t ns = kind ? nf : ns_all, sh = kind ? 0u : 16u;
#pragma unroll
        for (uint32_t g = 0; g < WS_G; ++g) {
            if (g >= ng) break;
            const uint32_t kr = kind ? ke[g] : ks[g];
            const uint32_t lo_t = (kr & 0xffffu) + 0x8000u, hi_t = (kr >> 16) + 0x8000u;
            uint32_t base = 0, mk = 0, nin = 0, v01 = 0, v23 = 0;
#pragma unroll
            for (uint32_t j = 0; j < WS_S; ++j) {
                const uint32_t x = (w[j] >> sh) & 0xffffu;
                const bool on = j < ns;
                base += on && x >= hi_t;                    // past every key of the item
                if (on && x >= lo_t && x < hi_t) {          // past some: compared entry by entry
                    mk |= 1u << j;
                    if (nin < 2) v01 |= x << (16 * nin);
                    else if (nin < 4) v23 |= x << (16 * (nin - 2));
                    ++nin;
                }
            }
            const int r0 = (int)(g * GROUP);
            const uint32_t nv = (uint32_t)min((int)n - r0, GROUP), tsr = (uint32_t)min(max(tm.z - r0, 0), GROUP);
            const uint32_t nl = kind ? 0u : ns_all - nf, fix = !kind && tsr > 0 && nl > 0;
            const size_t k = (size_t)i0 + (kind ? ng + g : g);      // one chunk: ch = chF = 1
            out[2 * k] = make_int4((int)((uint32_t)tm.x + g * (GROUP / 8) + (kind ? db.eoff8 : 0u)), (int)a,
                                   (int)(nv | ns << 9 | base << 16 | nl << 23 | kind << 30 | fix << 31), (int)mk);
            out[2 * k + 1] = make_int4(0, (int)v01, (int)v23, (int)nin);
        }
    }
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const int4 v)
{
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- bulk asynchronous copies (cp.async.bulk -> UBLKCP, the TMA unit) completing on an mbarrier --------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

constexpr int RK_K = 3;                 // stages of a warp's entry ring

template <bool V>
struct RankSmem {
    static constexpr uint32_t DESC = 2 * 32 * 32;                       // a warp's ring of 64 descriptors
    static constexpr uint32_t ARR = GROUP * 2;                          // one 16-bit array of a stage
    static constexpr uint32_t STAGE = (V ? 3 : 2) * ARR;                // keys | files | (values)
    static constexpr uint32_t BARS = 32;                                // RK_K mbarriers of 8 bytes (bulk-copy variant)
    static constexpr uint32_t WARP = DESC + RK_K * STAGE + BARS;
    static size_t hist_bytes(int nfiles) { return ((size_t)(nfiles + 1) * 4 + 127) / 128 * 128; }   // + the pads' slot
    static int warps(int nfiles, bool smem_hist)
    {
        const size_t total = 226 * 1024, hist = smem_hist ? hist_bytes(nfiles) : 0;
        if (hist >= total) return 0;                        // the histogram alone does not fit
        return (int)std::min<size_t>((total - hist) / WARP, RK_MAXW);
    }
    static size_t cta_bytes(int nfiles, bool smem_hist, int nwarps)
    {
        return (smem_hist ? hist_bytes(nfiles) : 0) + (size_t)nwarps * WARP;
    }
};

// The count-mode search kernel: ONE CTA per SM, up to 32 warps.  A warp owns every NW-th BLOCK of 8 consecutive
// items (static, fine-grained interleave); the descriptors of 32 items at a time go through a warp-private ring in
// shared memory (read back with broadcast loads); the entries go through a private K-stage cp.async ring: K - 1
// items ahead of the one being evaluated every lane copies ITS OWN 16 bytes of keys and of file indices into its
// own slots, so no lane reads another lane's bytes and the ring needs no barrier -- a lane's cp.async.wait_group
// covers exactly the bytes it is about to read.  (Loading straight into registers two items ahead was tried: the
// hardware's six scoreboards per warp are shared between the slots, and a use of the oldest slot then waits for
// the loads issued just before it.)
// PRE4: the file fields hold 4 * file (DbView::fshift == 2).  BULK: an item's entries arrive by ONE bulk copy (the TMA
// unit, issued by lane 0, completing on the stage's mbarrier) instead of two 16-byte cp.async per lane -- the kernel's
// busiest unit is the shared-memory / L1 pipe, and a bulk copy's writes do not pass through it.
template <bool V, bool SMEM, bool PRE4 = false, bool BULK = false>
__global__ void __launch_bounds__(RK_MAXW * 32, 1) k_count_rank(const CountArgs a)
{
    using L = RankSmem<V>;
    extern __shared__ __align__(128) unsigned char smem[];
    const DbView &db = a.db;
    uint32_t *hist = reinterpret_cast<uint32_t *>(smem);
    const uint32_t hist_bytes = SMEM ? ((uint32_t)(db.nfiles + 1) * 4 + 127) / 128 * 128 : 0;
    if (SMEM) {
        for (int f = threadIdx.x; f <= db.nfiles; f += blockDim.x) hist[f] = 0;
        __syncthreads();
    }
    const uint32_t hist_s = smem_u32(smem);
    const int lane = threadIdx.x & 31, W = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const uint32_t ring = hist_s + hist_bytes + (uint32_t)wib * L::WARP;       // descriptors
    const uint32_t lslot = ring + L::DESC + lane * 16;                          // this lane's column of the stages
    const uint32_t bars = ring + L::DESC + RK_K * L::STAGE;                     // BULK: one mbarrier per stage
    uint32_t par = 0;                                                           // BULK: the phase of each stage's barrier
    if (BULK) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < RK_K; ++k) mbar_init(bars + 8 * k, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    pdl_wait();                                             // the plan, the descriptors, the segments
    const unsigned long long nseg = a.tot[0], nit = a.tot[1];
    const uint32_t n_items = (nseg > a.seg_cap || nit > a.item_cap) ? 0u : (uint32_t)nit;
    const uint32_t nblk = (n_items + RK_IB - 1) / RK_IB, NW = gridDim.x * W, w = blockIdx.x * W + wib;
    const uint32_t nb = w < nblk ? (nblk - w + NW - 1) / NW : 0;       // blocks of this warp
    const uint32_t N = nb * RK_IB;                                      // item slots of this warp

    // descriptors: lane j of quad-block g holds item (j % 8) of this warp's block 4 g + j / 8; quad-block g lives in
    // half g % 2 of the ring, 32 bytes per item
    const int4 zero = make_int4(0, 0, 0, 0);
    auto load_desc = [&](uint32_t g, int4 &A, int4 &B) {
        const uint32_t b = 4 * g + (lane >> 3);
        const size_t it = ((size_t)b * NW + w) * RK_IB + (lane & 7);
        const bool in = b < nb && it < n_items;
        A = in ? __ldg(&a.desc[2 * it]) : zero;
        B = in ? __ldg(&a.desc[2 * it + 1]) : zero;
    };
    auto store_desc = [&](uint32_t g, const int4 &A, const int4 &B) {
        const uint32_t p = ring + (g & 1) * 1024 + lane * 32;
        sts128(p, A), sts128(p + 16, B);
    };
    int4 nA, nB;
    load_desc(0, nA, nB);
    store_desc(0, nA, nB);
    load_desc(1, nA, nB);
    store_desc(1, nA, nB);
    __syncwarp();
    load_desc(2, nA, nB);

    const uint16_t *const pB = db.rb + lane * 8;
    const uint16_t *const pV = db.rv + lane * 8;

    // issue the copies of item slot n into stage `stg` = n % K (kept by the caller: a modulo by 3 per item is five
    // instructions of a kernel that is bound by its issue slots)
    auto issue = [&](uint32_t n, uint32_t stg) {
        if (BULK) {
            // (every lane has taken the entries of the slot that used this stage out of shared memory: __syncwarp
            // in the loop below)
            if (lane == 0 && n < N) {
                const uint32_t nv = lds32(ring + (n & 63) * 32 + 8) & 0x1ffu;
                if (nv) {
                    const uint32_t x = lds32(ring + (n & 63) * 32), nvp = (nv + 7) & ~7u;
                    const uint32_t dst = ring + L::DESC + stg * L::STAGE, bar = bars + 8 * stg;
                    mbar_expect_tx(bar, nvp * (V ? 6u : 4u));
                    bulk_g2s(dst, db.rb + (size_t)x * 16, nvp * 4, bar);        // keys[nvp] | files[nvp]
                    if (V) bulk_g2s(dst + 2 * L::ARR, db.rv + (size_t)x * 8, nvp * 2, bar);
                }
            }
            return;
        }
        if (n < N) {
            const int2 xz = lds64(ring + (n & 63) * 32);                // A.x, A.y
            const uint32_t z = lds32(ring + (n & 63) * 32 + 8), nv = z & 0x1ffu;
            if ((uint32_t)lane * 8 < nv) {
                const uint32_t so = lslot + stg * L::STAGE, nvp = (nv + 7) & ~7u;
                const uint16_t *p = pB + (size_t)(uint32_t)xz.x * 16;
                cp_async16(so, p);
                cp_async16(so + L::ARR, p + nvp);
                if (V) cp_async16(so + 2 * L::ARR, pV + (size_t)(uint32_t)xz.x * 8);
            }
        }
        cp_async_commit();
    };

    // evaluate item slot n (in stage `stg` = n % K)
    auto eval = [&](uint32_t n, uint32_t stg) {
        const int4 A = lds128(ring + (n & 63) * 32);
        const uint32_t z = (uint32_t)A.z, nv = z & 0x1ffu;
        if (nv == 0) return;
        const int4 B = lds128(ring + (n & 63) * 32 + 16);
        const bool endk = z >> 30 & 1;
        const uint32_t so = lslot + stg * L::STAGE;
        if (BULK) {
            mbar_wait(bars + 8 * stg, (par >> stg) & 1u);
            par ^= 1u << stg;
        }
        // (BULK: the files follow the nvp keys of the group directly)
        const int4 k_ = lds128(so), f_ = lds128(so + (BULK ? ((nv + 7) & ~7u) * 2 : L::ARR));
        const uint32_t K[4] = {(uint32_t)k_.x, (uint32_t)k_.y, (uint32_t)k_.z, (uint32_t)k_.w};
        const uint32_t F[4] = {(uint32_t)f_.x, (uint32_t)f_.y, (uint32_t)f_.z, (uint32_t)f_.w};
        const uint32_t base2 = ((z >> 16) & 0x7fu) * 0x10001u;
        uint32_t c[4] = {base2, base2, base2, base2};
        auto step = [&](const uint32_t x2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t t;                                             // (x2 - K) & 0x80008000
                asm("lop3.b32 %0, %1, 0x80008000, 0, 0xc0;" : "=r"(t) : "r"(x2 - K[k]));
                c[k] += t >> 15;                                        // LEA.HI
            }
        };
        const uint32_t nin = (uint32_t)B.w;                             // the segments inside the window
        if (nin <= RK_INL) {
            if (nin > 0) step(__byte_perm((uint32_t)B.y, 0, 0x1010));
            if (nin > 1) step(__byte_perm((uint32_t)B.y, 0, 0x3232));
            if (nin > 2) step(__byte_perm((uint32_t)B.z, 0, 0x1010));
            if (nin > 3) step(__byte_perm((uint32_t)B.z, 0, 0x3232));
        } else {
            // more than the descriptor holds (few items): lane i fetches segment i / 32 + i if it is in the window
            const uint32_t sel = endk ? 0x1010u : 0x3232u;              // x = xe (END items) or xs, in both halves
            const uint32_t *pQ = static_cast<const uint32_t *>(static_cast<const void *>(a.segs)) + (uint32_t)A.y + lane;
            uint32_t m0 = (uint32_t)A.w, m1 = (uint32_t)B.x;
            const uint32_t w0 = m0 >> lane & 1 ? __ldg(pQ) : 0u, w1 = m1 >> lane & 1 ? __ldg(pQ + 32) : 0u;
            while (m0) {
                const int i = __ffs(m0) - 1;
                m0 &= m0 - 1;
                step(__byte_perm(__shfl_sync(FULL, w0, i), 0, sel));
            }
            while (m1) {
                const int i = __ffs(m1) - 1;
                m1 &= m1 - 1;
                step(__byte_perm(__shfl_sync(FULL, w1, i), 0, sel));
            }
        }
        if (z >> 31) {
            // a START item with entries that start before the bin (S == 1) and later-bin segments, which count only
            // for the entries inside the bin (src/igd_search.c:510-511,522): every xs is past S == 1, so those
            // entries have counted all of them -- taken back here
            const uint32_t nl = (z >> 23) & 0x7fu;
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] -= (((0x80018001u - K[k]) & 0x80008000u) >> 15) * nl;
        }
        if (V) {                                                        // value < v: the record counts for no query
            // rv = value' + 0x8000 with value' = clamp(value, 0, 32767), and 1 <= v <= 32767 on this path:
            // bit 15 of a half of (rv2 - v2) is set  <=>  value' >= v  <=>  value >= v  (no half borrows)
            const int4 v_ = lds128(so + 2 * L::ARR);
            const uint32_t pv2[4] = {(uint32_t)v_.x, (uint32_t)v_.y, (uint32_t)v_.z, (uint32_t)v_.w};
            const uint32_t v2 = (uint32_t)a.v * 0x10001u;
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] &= (((pv2[k] - v2) & 0x80008000u) >> 15) * 0xffffu;
        }
        // hits[file] += count (START) or -= count (END): one shared-memory atomic per entry; the pads of a tile's
        // last group carry file = nfiles, the histogram's trash slot
        if ((uint32_t)lane * 8 < nv) {
            // byte offsets of the two counters of a pair of entries
            auto off0 = [&](uint32_t f2) { return PRE4 ? f2 & 0xffffu : (f2 << 2) & 0x3fffcu; };
            auto off1 = [&](uint32_t f2) { return PRE4 ? f2 >> 16 : (f2 >> 14) & 0x3fffcu; };
            if (SMEM && !endk) {                                        // (warp-uniform: START items add as they are)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s + off0(F[k])), "r"(c[k] & 0xffffu) : "memory");
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s + off1(F[k])), "r"(c[k] >> 16) : "memory");
                }
                return;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t c0 = c[k] & 0xffffu, c1 = c[k] >> 16;
                if (endk) c0 = 0u - c0, c1 = 0u - c1;
                if (SMEM) {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s + off0(F[k])), "r"(c0) : "memory");
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hist_s + off1(F[k])), "r"(c1) : "memory");
                } else {
                    const uint32_t f0 = (F[k] & 0xffffu) >> db.fshift, f1 = (F[k] >> 16) >> db.fshift;
                    if (c0 && f0 < (uint32_t)db.nfiles) atomicAdd(&a.hits[f0], (unsigned long long)(long long)(int32_t)c0);
                    if (c1 && f1 < (uint32_t)db.nfiles) atomicAdd(&a.hits[f1], (unsigned long long)(long long)(int32_t)c1);
                }
            }
        }
    };

#pragma unroll
    for (int n = 0; n < RK_K - 1; ++n) issue(n, n);                     // prologue: slots 0 .. K-2
    uint32_t stg = 0;                                                   // n % K
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t prev = stg == 0 ? RK_K - 1 : stg - 1;            // (n + K - 1) % K: the stage slot n - 1 released
        if (BULK) __syncwarp();                                         // ... by every lane
        issue(n + RK_K - 1, prev);
        if (!BULK) cp_async_wait<RK_K - 1>();                           // slot n has landed (this lane's bytes)
        eval(n, stg);
        stg = stg == RK_K - 1 ? 0 : stg + 1;
        // every 32 items the descriptors of the quad-block after next move into the half of the ring that the one
        // just finished has left (items n + 1 .. n + K - 1 -- already issued -- are in the other half)
        if ((n & 31) == 31) {
            const uint32_t g = (n >> 5) + 2;
            __syncwarp();
            store_desc(g, nA, nB);
            __syncwarp();
            load_desc(g + 1, nA, nB);
        }
    }
    if (!BULK) cp_async_wait<0>();
    if (SMEM) {
        __syncthreads();
        // every CTA starts at its own place of the file range: the CTAs finish together, and the same counter from
        // all of them at once is one L2 atomic unit serving 148 requests while the others idle
        const int off = (int)(((long long)blockIdx.x * db.nfiles) / gridDim.x);
        for (int i = threadIdx.x; i < db.nfiles; i += blockDim.x) {
            const int f = i + off < db.nfiles ? i + off : i + off - db.nfiles;
            const uint32_t c = hist[f];                                 // a sum modulo 2^32 of +H and -K: signed
            if (c) atomicAdd(&a.hits[f], (unsigned long long)(long long)(int32_t)c);
        }
    }
}


// ---- count mode ----------------------------------------------------------------------------------
// One batch = 5 or 6 launches on one stream (segment count, plan, segment scatter, one or two descriptor kernels,
// search) with no host round trip in between: the kernels that consume the
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
    using LP = RankSmem<V>;
    using L32 = CountSmem<V>;
    // the shared histogram is used while it leaves room for at least 8 warps' pipelines
    *smem = (PK ? LP::warps(d->nfiles, true) : L32::warps(d->nfiles, true)) >= 8;
    const int nw = PK ? LP::warps(d->nfiles, *smem) : L32::warps(d->nfiles, *smem);
    *shm = PK ? LP::cta_bytes(d->nfiles, *smem, nw) : L32::cta_bytes(d->nfiles, *smem, nw);
    *threads = nw * 32;
    if (d->count_grid[PK][V] == 0) {
        const void *k = PK ? (*smem ? (d->fshift == 2 ? (d->bulk ? (const void *)k_count_rank<V, true, true, true> : (const void *)k_count_rank<V, true, true>)
                                                      : (d->bulk ? (const void *)k_count_rank<V, true, false, true> : (const void *)k_count_rank<V, true>))
                                    : (const void *)k_count_rank<V, false>)
                           : (*smem ? (const void *)k_count_items<V, true> : (const void *)k_count_items<V, false>);
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*shm));
        d->count_grid[PK][V] = d->sm_count;
    }
    *grid = d->count_grid[PK][V];
    return 0;
}

// a CTA's shared histogram is 32 bits wide (k_count_rank: a SIGNED sum of +H and -K): an item moves a counter by at
// most GROUP*SMAX = 2^14, a CTA takes at most items/grid + 32*COUNT_MAXW items, so with this many items per launch
// no counter leaves (-2^31, 2^31)
int64_t item_limit(const igdb_dev *d, int grid) { return ((int64_t)(1 << 17) - 1024) * grid; }
// the kernels may touch up to 32 descriptors / 1 segment past the plan's totals (padding)
constexpr int64_t CAP_SLACK = 64;

template <bool V, int MODE>
int enqueue_count_t(igdb_dev *d, const igdb_dev::Pending &p)
{
    constexpr bool PK = MODE == M_PACKED;
    cudaStream_t st = p.st;
    const DbView db = d->view();
    int grid = 0, threads = 0;
    size_t shm = 0;
    bool smem = false;
    if (count_geometry<V, PK>(d, &grid, &threads, &shm, &smem)) return -1;
    if (grow(&d->csegs, &d->cseg_cap, std::max<int64_t>(2 * p.n + 2 * d->T, 4096), 1, 1)) return -1;
    if (grow(&d->items, &d->item_cap, std::min<int64_t>((PK ? 2 : 1) * (d->db_groups + p.n / 64), item_limit(d, grid)) + 1024, 1, 1))
        return -1;
    if (PK && grow(&d->items2, &d->item2_cap, 2 * d->item_cap, 1, 1)) return -1;
    if (PK && !d->tplan && dev_alloc(d, &d->tplan, 2 * (size_t)d->T + 2)) return -1;
    const unsigned long long seg_cap = (unsigned long long)std::min<int64_t>(d->cseg_cap - CAP_SLACK, 0x7fffffff);
    const unsigned long long item_cap =
        (unsigned long long)std::min<int64_t>(d->item_cap - CAP_SLACK, item_limit(d, grid));
    const int thr = CSEG_T;
    // (one block of queries per CTA: a resident wave that strides over the batch was measured slower at 10 M queries --
    // 0.121 + 0.18 ms against 0.113 + 0.135 -- the hardware's CTA scheduler balances the tail better than a fixed stride)
    const int nb = (int)((p.n + CSEG_T * CSEG_Q - 1) / (CSEG_T * CSEG_Q));
    const int32_t tlo = std::max(0, std::min(p.tile_lo, d->T)), thi = std::max(tlo, std::min(p.tile_hi, d->T));
    if (d->plan_grid_max == 0) {
        // k_plan's CTAs wait for each other, so all of them must be resident at once: never more than the
        // occupancy calculator allows, and at most SCAN_T (the last CTA scans one aggregate per thread)
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_plan, SCAN_T, 0));
        per_sm = std::max(1, std::min(per_sm, 4));
        d->plan_grid_max = std::max(1, std::min(per_sm * d->sm_count, SCAN_T));
    }
    d->plan_grid = std::max(1, std::min(d->plan_grid_max, (thi - tlo + SCAN_T - 1) / SCAN_T));
    const bool stages = d->stage_timing && MODE != M_ABN32;
    if (grow(&d->qrank, &d->qrank_cap, p.n)) return -1;
    // plan_tot is about to be rewritten: the readback of the previous batch's totals (side stream) must have read it
    if (d->last_plan_ev) CK(cudaStreamWaitEvent(st, d->last_plan_ev, 0));
    if (stages) CK(cudaEventRecord(d->stage_ev[0], st));
    CK(launch_pdl(k_cseg_count<V, MODE>, nb, thr, 0, st, db, p.n, p.ctg, p.qs, p.qe, d->cnt2, d->qrank, p.dyn));
    if (stages) CK(cudaEventRecord(d->stage_ev[1], st));
    // (the plan reads the later-bin counters that the scatter counts back down to zero)
    CK(launch_pdl(k_plan, d->plan_grid, SCAN_T, 0, st, db, d->cnt2, d->cnt2, ++d->plan_epoch, d->plan_agg, d->plan_flag,
                  d->plan_flag + 1, d->cseg_off, d->cseg_mid, d->plan_tot, seg_cap, item_cap, d->items, PK ? 1 : 0, d->tplan,
                  tlo, thi, p.dyn));
    if (stages) CK(cudaEventRecord(d->stage_ev[2], st));
    CK(launch_pdl(k_cseg_scatter<V, MODE>, nb, thr, 0, st, db, p.n, p.ctg, p.qs, p.qe, d->cnt2, d->qrank, d->cseg_off,
                  d->cseg_mid, d->plan_tot, seg_cap, item_cap, (void *)d->csegs, p.dyn));
    // the plan's totals go back to the host on a side stream, so that the copy does not sit between the kernels
    CK(cudaEventRecord(d->plan_done, st));
    CK(cudaStreamWaitEvent(d->side, d->plan_done, 0));
    CK(cudaMemcpyAsync((void *)(d->h_plan + 4 * p.slot), d->plan_tot, 32, cudaMemcpyDeviceToHost, d->side));
    CK(cudaEventRecord(d->plan_ev[p.slot], d->side));
    d->last_plan_ev = d->plan_ev[p.slot];
    CountArgs a{};
    a.db = db, a.items = d->items, a.desc = d->items2, a.segs = d->csegs, a.tot = d->plan_tot;
    a.seg_cap = seg_cap, a.item_cap = item_cap, a.v = p.v, a.hits = p.hits;
    if (PK) {
        // the descriptors of k_count_rank (with the windows of the items): one thread per item, 8 or 32 per tile
        // a database of small tiles: most of them by k_windows_small, the others (many segments) by the general kernel
        const bool sparse = d->db_groups < 6 * (int64_t)d->T;       // groups per tile on average
        const int64_t nt = std::max<int64_t>(thi - tlo, 1);
        if (sparse) {
            CK(launch_pdl(k_windows_small, (unsigned)((nt + 127) / 128), 128, 0, st, db, (const int4 *)d->tplan,
                          (const uint32_t *)(const void *)d->csegs, (const unsigned long long *)d->plan_tot, seg_cap,
                          item_cap, d->items2, (uint32_t)tlo, (uint32_t)thi, p.dyn));
            d->launches += 1;
        }
        // the others: a warp per tile; tiles of W2_NG groups or more (none in the BASELINE configurations) by the
        // general kernel
        // tiles per warp and turn: few enough that a small tile range (a shard of a multi-GPU search) still fills the GPU
        const int64_t tpw = sparse ? 32 : (p.dyn ? 2 : std::max<int64_t>(1, std::min<int64_t>(8, nt / ((int64_t)d->sm_count * 40))));
        const unsigned wgrid = (unsigned)std::min<int64_t>((nt + 8 * tpw - 1) / (8 * tpw), (int64_t)d->sm_count * 48);
        CK(launch_pdl(k_windows_warp, wgrid, 256, 0, st, db, (const int4 *)d->tplan,
                      (const uint32_t *)(const void *)d->csegs, (const unsigned long long *)d->plan_tot, seg_cap, item_cap,
                      d->items2, sparse ? 1 : 0, (uint32_t)tpw, (uint32_t)tlo, (uint32_t)thi, p.dyn));
        d->launches += 1;
        if (d->max_tile_groups >= (int64_t)W2_NG) {
            const unsigned ggrid = (unsigned)std::min<int64_t>((nt + 255) / 256, (int64_t)d->sm_count * 64);
            CK(launch_pdl(k_windows<32>, ggrid, 256, 0, st, db, (const int4 *)d->tplan,
                          (const uint32_t *)(const void *)d->csegs, (const unsigned long long *)d->plan_tot, seg_cap,
                          item_cap, d->items2, 1, (uint32_t)tlo, (uint32_t)thi, W2_NG, p.dyn));
            d->launches += 1;
        }
    }
    if (stages) CK(cudaEventRecord(d->stage_ev[3], st));
    if (MODE != M_ABN32) CK(cudaEventRecord(d->kev[0], st));
    if (PK) {
        if (smem && d->fshift == 2 && d->bulk) CK(launch_pdl(k_count_rank<V, true, true, true>, grid, threads, shm, st, a));
        else if (smem && d->fshift == 2) CK(launch_pdl(k_count_rank<V, true, true>, grid, threads, shm, st, a));
        else if (smem && d->bulk) CK(launch_pdl(k_count_rank<V, true, false, true>, grid, threads, shm, st, a));
        else if (smem) CK(launch_pdl(k_count_rank<V, true>, grid, threads, shm, st, a));
        else CK(launch_pdl(k_count_rank<V, false>, grid, threads, shm, st, a));
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

}  // namespace

int igdb::enqueue_count(igdb_dev *d, const igdb_dev::Pending &p)
{
    if (p.n <= 0) return 0;
    switch (p.mode) {
    case M_PACKED: return p.mode_v ? enqueue_count_t<true, M_PACKED>(d, p) : enqueue_count_t<false, M_PACKED>(d, p);
    case M_ABN32: return p.mode_v ? enqueue_count_t<true, M_ABN32>(d, p) : enqueue_count_t<false, M_ABN32>(d, p);
    default: return p.mode_v ? enqueue_count_t<true, M_ALL32>(d, p) : enqueue_count_t<false, M_ALL32>(d, p);
    }
}

// the one-call entry points and the pipelined one (igdb_dev_count_submit) do not mix on a handle
int igdb::slots_idle(const igdb_dev *d, const char *who)
{
    for (const auto &s : d->slot)
        if (s.active) {
            snprintf(g_err, sizeof g_err, "%s: submitted count batches are still in flight (igd_b200_count_wait them first)", who);
            return -1;
        }
    return 0;
}

namespace {

// wait for the plan of the enqueued batch p; make sure it ran (see above)
int settle_count(igdb_dev *d, const igdb_dev::Pending &p)
{
    if (p.n <= 0) return 0;
    const bool pk = p.mode == M_PACKED;
    for (;;) {
        CK(cudaEventSynchronize(d->plan_ev[p.slot]));
        const volatile unsigned long long *hp = d->h_plan + 4 * p.slot;
        const unsigned long long nsegs = hp[0], nitems = hp[1], nrec = hp[2], nabn = hp[3];
        if (getenv("IGD_B200_DEBUG"))
            fprintf(stderr, "[igd_b200] count batch (mode %d): n=%lld segments=%llu items=%llu left=%llu (caps %lld / %lld)\n",
                    p.mode, (long long)p.n, nsegs, nitems, nabn, (long long)d->cseg_cap, (long long)d->item_cap);
        const int64_t lim = item_limit(d, d->count_grid[pk][p.mode_v]);
        if (nsegs <= (unsigned long long)std::min<int64_t>(d->cseg_cap - CAP_SLACK, 0x7fffffff) &&
            nitems <= (unsigned long long)std::min<int64_t>(d->item_cap - CAP_SLACK, lim)) {
            if (p.mode != M_ABN32) {
                // what the search kernel of this batch reads: descriptors, records, segments
                const int64_t rb = pk ? (p.mode_v ? 6 : 4) : (p.mode_v ? 16 : 12), sgb = pk ? 4 : 8;
                d->last_plan[0] = (int64_t)nsegs, d->last_plan[1] = (int64_t)nitems, d->last_plan[2] = (int64_t)nrec;
                d->last_plan[3] = (pk ? 32 : 16) * (int64_t)nitems + rb * (int64_t)nrec + sgb * (int64_t)nsegs;
            }
            if (pk && nabn > 0) {           // the queries the packed layout cannot express: second pipeline
                igdb_dev::Pending q = p;
                q.mode = M_ABN32;
                if (enqueue_count(d, q) || settle_count(d, q)) return -1;
            }
            return 0;
        }
        // none of this batch's kernels wrote anything, but its segment counters were not counted back down
        ++d->misfits;
        CK(cudaStreamSynchronize(p.st));
        CK(cudaMemsetAsync(d->cnt2, 0, ((size_t)2 * d->T + 2) * 4, p.st));
        if (nsegs >= 0x7fffffffull || nitems > (unsigned long long)lim) {
            if (p.n <= 1 || p.dyn) {
                snprintf(g_err, sizeof g_err, p.dyn ? "a dispatched shard expands to too many (query, tile) segments for one batch"
                                                    : "a single query expands to too many (query, tile) segments");
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

}  // namespace

int igdb::count_mode_for(const igdb_dev *d, bool mode_v, int32_t v)
{
    // the packed layout holds value clamped to [0, 32767]: exact for the thresholds 1 .. 32767 only
    return d->packed && (!mode_v || (v >= 1 && v <= 32767)) ? M_PACKED : M_ALL32;
}

int igdb::finish_pending(igdb_dev *d)
{
    if (!d->pending.active) return 0;
    d->pending.active = false;
    return settle_count(d, d->pending);
}


extern "C" {

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


void igdb_dev_last_plan(const igdb_dev *d, int64_t out[4])
{
    for (int k = 0; k < 4; ++k) out[k] = d->last_plan[k];
}

int igdb_dev_count_device_tiles(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                                int mode_v, int32_t v, uint64_t *d_hits, void *stream, int32_t tile_lo, int32_t tile_hi);

int igdb_dev_count_device(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                          int mode_v, int32_t v, uint64_t *d_hits, void *stream)
{
    return igdb_dev_count_device_tiles(d, n, d_ctg, d_qs, d_qe, mode_v, v, d_hits, stream, 0, 0x7fffffff);
}

// the same when the caller knows that no query touches a tile outside [tile_lo, tile_hi) (a shard of a multi-GPU
// search): the plan and the descriptors then cover that range only
int igdb_dev_count_device_tiles(igdb_dev *d, int64_t n, const int32_t *d_ctg, const int32_t *d_qs, const int32_t *d_qe,
                                int mode_v, int32_t v, uint64_t *d_hits, void *stream, int32_t tile_lo, int32_t tile_hi)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (slots_idle(d, "count_device")) return -1;
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
        p.tile_lo = tile_lo, p.tile_hi = tile_hi;
        p.mode = count_mode_for(d, p.mode_v, p.v);
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

// the same for a batch whose shape lives in device memory -- d_dyn = {queries, tile_lo, tile_hi} (uint64), written by
// an earlier kernel on the stream (igdb_dispatch_async): n_max bounds the number of queries (it sizes the scratch and
// the grids), nothing is read back before the launches, so the host never waits between the query exchange of a
// multi-GPU step and its search
int igdb_dev_count_device_dyn(igdb_dev *d, int64_t n_max, const uint64_t *d_dyn, const int32_t *d_ctg, const int32_t *d_qs,
                              const int32_t *d_qe, int mode_v, int32_t v, uint64_t *d_hits, void *stream)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (slots_idle(d, "count_device")) return -1;
    if (finish_pending(d)) return -1;
    if (n_max <= 0 || d->T == 0) return 0;
    if (n_max > ((int64_t)1 << 26)) {
        snprintf(g_err, sizeof g_err, "count_device_dyn: at most 2^26 queries per batch");
        return -1;
    }
    igdb_dev::Pending p;
    p.mode_v = mode_v != 0, p.n = n_max, p.ctg = d_ctg, p.qs = d_qs, p.qe = d_qe, p.v = v;
    p.hits = (unsigned long long *)d_hits, p.st = stream ? (cudaStream_t)stream : d->stream;
    p.dyn = (const unsigned long long *)d_dyn;
    p.mode = count_mode_for(d, p.mode_v, p.v);
    if (enqueue_count(d, p)) return -1;
    p.active = true;
    d->pending = p;
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
    if (slots_idle(d, "count")) return -1;
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
        p.mode = count_mode_for(d, p.mode_v, p.v);
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

}  // extern "C"

// ---- the pipelined host entry: a STREAM of batches, two in flight -----------------------------------------
// End to end a count call is PCIe first and HBM second: 12 B per query cross the bus before a search that needs the
// whole batch for its plan (BASELINE config 3: 2.2 ms of H2D in front of 2.0 ms of kernels).  One call cannot
// overlap the two -- cutting a batch into sub-batches would read the database once per sub-batch -- but a caller
// with more than one batch can: submit(k + 1) copies on d->copy while batch k searches on d->stream.
namespace {

// everything behind the search of slot s: the counts to the pinned result, the event wait() sleeps on
int slot_tail(igdb_dev *d, igdb_dev::CountSlot &s)
{
    CK(cudaMemcpyAsync(s.h_hits, s.d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaEventRecord(s.done, d->stream));
    return 0;
}

}  // namespace

extern "C" {

int igdb_dev_count_submit(igdb_dev *d, int slot, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                          int mode_v, int32_t v)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (slot < 0 || slot > 1) {
        snprintf(g_err, sizeof g_err, "count_submit: slot must be 0 or 1");
        return -1;
    }
    igdb_dev::CountSlot &s = d->slot[slot];
    if (s.active) {
        snprintf(g_err, sizeof g_err, "count_submit: slot %d holds a batch that has not been waited for", slot);
        return -1;
    }
    if (n > ((int64_t)1 << 26)) {
        snprintf(g_err, sizeof g_err, "count_submit: at most 2^26 queries per batch");
        return -1;
    }
    if (finish_pending(d)) return -1;
    if (!d->copy) CK(cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking));
    // (each piece on its own: a failure half way leaves the slot usable for the next attempt)
    if (!s.d_hits) CK(cudaMalloc((void **)&s.d_hits, std::max<size_t>((size_t)d->nfiles, 1) * 8));
    if (!s.h_hits) CK(cudaMallocHost((void **)&s.h_hits, std::max<size_t>((size_t)d->nfiles, 1) * 8));
    if (!s.h2d) CK(cudaEventCreateWithFlags(&s.h2d, cudaEventDisableTiming));
    if (!s.done) CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    n = std::max<int64_t>(n, 0);
    if (d->T == 0) n = 0;
    if (grow(&s.q, &s.q_cap, 3 * std::max<int64_t>(n, 1))) return -1;
    const int64_t cap = s.q_cap / 3;
    int32_t *d_ctg = s.q, *d_qs = s.q + cap, *d_qe = s.q + 2 * cap;
    if (n > 0) {
        CK(cudaMemcpyAsync(d_ctg, ctg, (size_t)n * 4, cudaMemcpyHostToDevice, d->copy));
        CK(cudaMemcpyAsync(d_qs, qs, (size_t)n * 4, cudaMemcpyHostToDevice, d->copy));
        CK(cudaMemcpyAsync(d_qe, qe, (size_t)n * 4, cudaMemcpyHostToDevice, d->copy));
    }
    CK(cudaEventRecord(s.h2d, d->copy));
    CK(cudaStreamWaitEvent(d->stream, s.h2d, 0));
    CK(cudaMemsetAsync(s.d_hits, 0, (size_t)d->nfiles * 8, d->stream));
    igdb_dev::Pending p;
    p.mode_v = mode_v != 0, p.n = n, p.ctg = d_ctg, p.qs = d_qs, p.qe = d_qe, p.v = v;
    p.hits = s.d_hits, p.st = d->stream, p.slot = slot;
    p.mode = count_mode_for(d, p.mode_v, p.v);
    if (enqueue_count(d, p)) return -1;
    if (slot_tail(d, s)) return -1;
    s.p = p, s.active = true, s.seq = ++d->submit_seq;
    return 0;
}

// hits_delta[nfiles] receives the counts of the batch in `slot`.  Batches are waited for in the order they were
// submitted.  A batch whose plan did not fit the scratch buffers is run again here (settle_count) -- and so is the
// batch submitted behind it, which planned on segment counters the first one had left behind.
int igdb_dev_count_wait(igdb_dev *d, int slot, uint64_t *hits_delta)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (slot < 0 || slot > 1 || !d->slot[slot].active) {
        snprintf(g_err, sizeof g_err, "count_wait: no batch was submitted in slot %d", slot);
        return -1;
    }
    igdb_dev::CountSlot &s = d->slot[slot], &o = d->slot[slot ^ 1];
    if (o.active && o.seq < s.seq) {
        snprintf(g_err, sizeof g_err, "count_wait: the batch in slot %d was submitted first; wait for it first", slot ^ 1);
        return -1;
    }
    s.active = false;
    const int64_t launches0 = d->launches;
    const uint64_t misfits0 = d->misfits;
    if (settle_count(d, s.p)) return -1;
    if (d->launches != launches0 && slot_tail(d, s)) return -1;        // it ran again (or its second pipeline ran)
    if (d->misfits != misfits0 && o.active) {
        CK(cudaMemsetAsync(o.d_hits, 0, (size_t)d->nfiles * 8, d->stream));
        if (enqueue_count(d, o.p) || slot_tail(d, o)) return -1;
    }
    CK(cudaEventSynchronize(s.done));
    memcpy(hits_delta, s.h_hits, (size_t)d->nfiles * 8);
    return 0;
}

}  // extern "C"
