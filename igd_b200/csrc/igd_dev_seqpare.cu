// igd_b200 CUDA layer -- -s: Seqpare similarity.
#include "igd_dev_impl.cuh"

#include "igd_dev_sort.cuh"

namespace {
struct Slice {                              // a piece of the handle's pooled temporaries
    void *p = nullptr;
    template <class T>
    T *as() const { return (T *)p; }
};
}  // namespace

namespace {

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

// seq_overlaps as a list (src/igd_search.c:253-353): the same pairs, as the reference's overlap_t records
// {idx_t = n1 of the QUERY, idx_g = index in the tile the record was found in, idx_f, sm} in its scan order
__global__ void k_seq_pairs(int32_t nbp, int64_t n, const int64_t *__restrict__ qoff, const int32_t *__restrict__ qs,
                            const int32_t *__restrict__ qe, int64_t total, const int32_t *__restrict__ os,
                            const int32_t *__restrict__ oe, const int32_t *__restrict__ of,
                            const int32_t *__restrict__ og, int4 *__restrict__ out)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= total) return;
    int64_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&qoff[mid]) <= p) lo = mid; else hi = mid;
    }
    const int32_t s = __ldg(&qs[lo]), e = __ldg(&qe[lo]);
    const int32_t rs = os[p], re = oe[p];
    const float qlen = __int2float_rn(e - s), rlen = __int2float_rn(re - rs);
    const float st = __int2float_rn(min(e, re) - max(s, rs));
    const float sm = __fdiv_rn(st, __fsub_rn(__fadd_rn(qlen, rlen), st));
    out[p] = make_int4(s / nbp, og[p], of[p], __float_as_int(sm));
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

// The same sweep, one WARP per problem.  A candidate whose row or column is taken stays rejected for good (the
// tables only grow), so the lanes test 32 candidates at a time against the tables -- their key / value / row /
// column gathers in flight together instead of one dependent chain per candidate -- and only the few that pass
// are replayed in order by lane 0, which tests them again (an earlier one of the same batch may have taken their
// row or column) and inserts.  Picks and their order are exactly those of k_seq_sweep.
__global__ void k_seq_sweep_warp(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ val,
                                 const uint32_t *__restrict__ row, const unsigned long long *__restrict__ col,
                                 int64_t nprob, const int64_t *__restrict__ pstart, const uint32_t *__restrict__ tsize,
                                 const unsigned long long *__restrict__ toff, unsigned long long *__restrict__ tables,
                                 float *__restrict__ pick, uint32_t *__restrict__ npick)
{
    const int64_t pid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pid >= nprob) return;
    const int64_t a = pstart[pid], b = pstart[pid + 1];
    uint32_t np = 0;
    if (b > a) {
        unsigned long long *t = tables + toff[pid];
        const uint32_t mask = tsize[pid] - 1;
        for (int64_t base = a; base < b; base += 32) {
            const int64_t i = base + lane;
            const bool valid = i < b;
            const uint32_t bits = valid ? ~(uint32_t)key[i] : 0u;
            const bool alive = valid && (int32_t)bits > 0;              // similarity 0: the reference stops (maxf == 0)
            unsigned long long r = 0, cc = 0;
            bool open = false;
            if (alive) {
                const uint32_t c = val[i];
                r = 1ull << 63 | row[c], cc = col[c];
                open = !seq_probe(t, mask, r, false) && !seq_probe(t, mask, cc, false);
            }
            unsigned m = __ballot_sync(FULL, open);
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const unsigned long long rj = __shfl_sync(FULL, r, j), cj = __shfl_sync(FULL, cc, j);
                const uint32_t bj = __shfl_sync(FULL, bits, j);
                int took = 0;
                if (lane == 0 && !seq_probe(t, mask, rj, false) && !seq_probe(t, mask, cj, false)) {
                    seq_probe(t, mask, rj, true);
                    seq_probe(t, mask, cj, true);
                    pick[a + np] = __uint_as_float(bj);
                    took = 1;
                }
                __syncwarp();                                           // lane 0's inserts, for everybody's next probes
                np += __shfl_sync(FULL, took, 0);
            }
            if (__any_sync(FULL, valid && !alive)) break;
        }
    }
    if (lane == 0) npick[pid] = np;
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


}  // namespace

extern "C" {

// seq_overlaps for a batch of queries: query_off[n+1] and the overlap_t records (16 B each) in the handle's pinned
// arena (valid until the next list-returning call on the handle).  Returns the number of records.
int64_t igdb_dev_seq_pairs_host(igdb_dev *d, int64_t n, const int32_t *ctg, const int32_t *qs, const int32_t *qe,
                                int64_t **query_off, void **pairs)
{
    g_err[0] = 0;
    *query_off = nullptr, *pairs = nullptr;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    cudaStream_t st = d->stream;
    if (n >= (1 << 26)) {
        snprintf(g_err, sizeof g_err, "seq_overlaps: batch of %lld queries is too large", (long long)n);
        return -1;
    }
    const size_t qbytes = (((size_t)std::max<int64_t>(n, 0) + 1) * 8 + 255) & ~(size_t)255;
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
    if (n <= 0 || d->T == 0) {
        if (arena(qbytes)) return -1;
        memset(d->arena, 0, qbytes);
        *query_off = (int64_t *)d->arena;
        return 0;
    }
    if (upload_queries(d, n, ctg, qs, qe, st)) return -1;
    DevBuf qoff, out, rec;
    const int64_t total = full_device(d, n, st, true, qoff, out);
    if (total < 0) return -1;
    if (arena(qbytes + (size_t)total * 16)) return -1;
    CK(cudaMemcpyAsync(d->arena, qoff.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (total > 0) {
        if (rec.alloc((size_t)total * 16)) return -1;
        const int32_t *o = out.as<int32_t>();
        k_seq_pairs<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d->nbp, n, qoff.as<int64_t>(), d->q_qs, d->q_qe,
                                                                    total, o, o + total, o + 2 * total, o + 3 * total,
                                                                    rec.as<int4>());
        d->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync((char *)d->arena + qbytes, rec.p, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    *query_off = (int64_t *)d->arena;
    if (total > 0) *pairs = (char *)d->arena + qbytes;
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
    // IGD_B200_DEBUG: stage times on stderr (events on the stream)
    const bool dbg = getenv("IGD_B200_DEBUG") != nullptr;
    cudaEvent_t tev[6] = {};
    auto mark = [&](int k) {
        if (!dbg) return;
        if (!tev[k]) cudaEventCreate(&tev[k]);
        cudaEventRecord(tev[k], st);
    };
    mark(0);
    if (upload_queries(d, n, ctg, qs, qe, st)) return -1;
    DevBuf qoff, out;
    const int64_t total = full_device(d, n, st, true, qoff, out);
    if (total < 0) return -1;
    if (total == 0) return 0;
    mark(1);
    if (total >= (1ll << 31)) {
        snprintf(g_err, sizeof g_err, "Seqpare: %lld candidate pairs exceed one pass (2^31)", (long long)total);
        return -1;
    }
    // the temporaries of a call come out of ONE buffer that the handle keeps (grown to the high-water mark): a
    // cudaMalloc / cudaFree pair per multi-GB array and call was a sizeable part of the call
    Slice b_grp, b_gnq, b_key, b_key2, b_val, b_val2, b_row, b_col, b_tmp, b_pstart, b_tsize, b_toff, b_pick, b_npick, b_acc;
    {
        struct Want { Slice *s; size_t bytes; };
        const Want want[] = {{&b_grp, (size_t)n * 4}, {&b_gnq, (size_t)ngroups * 4}, {&b_key, (size_t)total * 8},
                             {&b_key2, (size_t)total * 8}, {&b_val, (size_t)total * 4}, {&b_val2, (size_t)total * 4},
                             {&b_row, (size_t)total * 4}, {&b_col, (size_t)total * 8},
                             {&b_tmp, radix_sort_tmp_bytes(total)}, {&b_pstart, (size_t)(nprob + 1) * 8},
                             {&b_tsize, (size_t)(nprob + 1) * 4}, {&b_toff, (size_t)(nprob + 1) * 8},
                             {&b_pick, (size_t)total * 4}, {&b_npick, (size_t)nprob * 4}, {&b_acc, (size_t)d->nfiles * 8}};
        size_t need = 0;
        for (const Want &w : want) need += (w.bytes + 255) & ~(size_t)255;
        if (need > d->seq_pool_cap) {
            if (d->seq_pool) CK(cudaFree(d->seq_pool));
            d->seq_pool = nullptr, d->seq_pool_cap = 0;
            if (cudaMalloc(&d->seq_pool, need + (need >> 3)) != cudaSuccess) {
                cudaGetLastError();
                snprintf(g_err, sizeof g_err, "Seqpare: %zu bytes of temporaries do not fit in device memory", need);
                return -1;
            }
            d->seq_pool_cap = need + (need >> 3);
        }
        char *p = (char *)d->seq_pool;
        for (const Want &w : want) w.s->p = p, p += (w.bytes + 255) & ~(size_t)255;
    }
    CK(cudaMemcpyAsync(b_grp.p, grp, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b_gnq.p, grp_nq, (size_t)ngroups * 4, cudaMemcpyHostToDevice, st));
    const int32_t *o = out.as<int32_t>();
    k_seq_candidates<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        d->view(), n, qoff.as<int64_t>(), d->q_qs, d->q_qe, b_grp.as<int32_t>(), total, o, o + total, o + 2 * total,
        o + 3 * total, b_key.as<unsigned long long>(), b_val.as<uint32_t>(), b_row.as<uint32_t>(),
        b_col.as<unsigned long long>());
    CK(cudaGetLastError());
    mark(2);
    // stable LSD radix sort (igd_dev_sort.cuh) on the used key bits only
    int end_bit = 32;
    while (end_bit < 64 && (nprob >> (end_bit - 32)) != 0) ++end_bit;
    int where = 0;
    if (radix_sort_pairs(d, b_key.as<unsigned long long>(), b_key2.as<unsigned long long>(), b_val.as<uint32_t>(),
                         b_val2.as<uint32_t>(), total, end_bit, b_tmp.p, st, &where))
        return -1;
    if (where == 0) std::swap(b_key.p, b_key2.p), std::swap(b_val.p, b_val2.p);     // sorted pairs: (b_key2, b_val2)
    mark(3);
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
    Slice b_tab;
    if (slots * 8 + 16 > d->seq_tab_cap) {
        if (d->seq_tab) CK(cudaFree(d->seq_tab));
        d->seq_tab = nullptr, d->seq_tab_cap = 0;
        if (cudaMalloc(&d->seq_tab, slots * 8 + (slots >> 1) + 16) != cudaSuccess) {
            cudaGetLastError();
            snprintf(g_err, sizeof g_err, "Seqpare: matching tables (%zu slots) do not fit in device memory", slots);
            return -1;
        }
        d->seq_tab_cap = slots * 8 + (slots >> 1) + 16;
    }
    b_tab.p = d->seq_tab;
    CK(cudaMemsetAsync(b_tab.p, 0, std::max<size_t>(slots * 8, 16), st));
    static const bool thread_sweep = getenv("IGD_B200_SEQ_THREAD_SWEEP") != nullptr;    // the one-thread-per-problem form
    if (thread_sweep)
        k_seq_sweep<<<(unsigned)((nprob + 127) / 128), 128, 0, st>>>(
            b_key2.as<unsigned long long>(), b_val2.as<uint32_t>(), b_row.as<uint32_t>(), b_col.as<unsigned long long>(),
            nprob, b_pstart.as<int64_t>(), b_tsize.as<uint32_t>(), b_toff.as<unsigned long long>(),
            b_tab.as<unsigned long long>(), b_pick.as<float>(), b_npick.as<uint32_t>());
    else
        k_seq_sweep_warp<<<(unsigned)((nprob * 32 + 255) / 256), 256, 0, st>>>(
            b_key2.as<unsigned long long>(), b_val2.as<uint32_t>(), b_row.as<uint32_t>(), b_col.as<unsigned long long>(),
            nprob, b_pstart.as<int64_t>(), b_tsize.as<uint32_t>(), b_toff.as<unsigned long long>(),
            b_tab.as<unsigned long long>(), b_pick.as<float>(), b_npick.as<uint32_t>());
    mark(4);
    k_seq_sum<<<(d->nfiles + 127) / 128, 128, 0, st>>>(d->nfiles, ngroups, b_pstart.as<int64_t>(), b_pick.as<float>(),
                                                       b_npick.as<uint32_t>(), b_acc.as<double>());
    mark(5);
    d->launches += 5;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(acc, b_acc.p, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dbg) {
        float ms[5] = {};
        for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&ms[k], tev[k], tev[k + 1]);
        fprintf(stderr, "[igd_b200] seqpare: %lld pairs, %lld problems; -f lists %.2f ms, candidates %.2f, sort %.2f, "
                        "ranges+tables+sweep %.2f, sum %.2f\n", (long long)total, (long long)nprob, ms[0], ms[1], ms[2],
                ms[3], ms[4]);
        for (auto &e : tev) if (e) cudaEventDestroy(e);
    }
    return 0;
}

}  // extern "C"
