// igd_b200 CUDA layer -- query ingest: BED text parsed on the device, then counted.
#include "igd_dev_impl.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Query ingest on the device: the BED text of a query file goes to HBM as it is and is parsed there.
// Restates the line loop of getOverlaps (src/igd_search.c:696-719: one line per '\n') and parse_bed
// (src/igd_base.c:53-72): fields split on '\t' (a NUL ends the line early), contig = field 0, start / end =
// atol of fields 1 / 2 narrowed to int32, accepted iff >= 3 fields, contig begins "chr" and is shorter than 40
// chars, end > 0; then get_id (src/igd_base.c:325-331): exact match of the contig name, -1 if absent.
// Two launches: k_line_starts streams the text 16 bytes per thread and compacts the offsets where a line that
// could be accepted starts (offset 0 or the byte after a '\n', first character 'c'); k_parse_lines then gives
// every such line its own thread, which parses it and appends (ctg, qs, qe) at an atomic cursor -- the order of
// the queries does not matter to hits[].  Rejected lines and unknown contigs produce nothing.  (The first
// version looked for a line start with one thread per BYTE and let the finder parse: 1 thread in 30 working,
// ~1.2 ms per 16 MiB chunk -- most of what search() cost.)
// ------------------------------------------------------------------------------------------------
struct CtgTable {
    const char *names;          // [nctg][40]
    const int32_t *slots;       // open addressing on FNV-1a of the name, -1 = empty
    uint32_t mask;
};

__host__ __device__ __forceinline__ uint32_t fnv1a(const char *s, int n)
{
    uint32_t h = 2166136261u;
    for (int i = 0; i < n; ++i) h = (h ^ (uint8_t)s[i]) * 16777619u;
    return h;
}

// glibc atol = strtol(base 10) on a field that ends at '\t', '\n', NUL or the end of the text:
// skip isspace, optional sign, digits, saturate at LONG_MIN / LONG_MAX; the caller narrows to int32.
__device__ __forceinline__ int32_t dev_atol(const char *__restrict__ t, int64_t p, int64_t end)
{
    auto at = [&](int64_t i) -> int { return i < end ? (int)(uint8_t)t[i] : 0; };
    int c = at(p);
    while (c == ' ' || c == '\v' || c == '\f' || c == '\r') c = at(++p);     // '\t' and '\n' end the field
    bool neg = false;
    if (c == '-' || c == '+') neg = c == '-', c = at(++p);
    unsigned long long v = 0;
    const unsigned long long lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    bool sat = false;
    while (c >= '0' && c <= '9') {
        const unsigned d = (unsigned)(c - '0');
        if (v > (lim - d) / 10) sat = true;
        v = v * 10 + d;
        c = at(++p);
    }
    if (sat) v = lim;
    const long long r = neg ? (long long)(0ull - v) : (long long)v;
    return (int32_t)r;
}

// offsets of the lines that start with 'c' (every accepted line does: "chr..."), in no particular order.
// At most len / 2 + 1 of them: such a line is at least one character and its newline.
__global__ void __launch_bounds__(256) k_line_starts(const char *__restrict__ text, int64_t len,
                                                      uint32_t *__restrict__ starts, unsigned int *__restrict__ n_starts)
{
    __shared__ uint32_t wsum[8], base_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t nvec = (len + 15) / 16;
    for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x; v0 < nvec; v0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = v0 + threadIdx.x, b0 = v * 16;
        uint32_t mask = 0;                                      // bit j: a candidate line starts at b0 + j
        if (v < nvec) {
            // bytes b0 - 1 .. b0 + 15; the text buffer is padded, reads past len see what the tail check rejects
            const uint4 w = __ldg(reinterpret_cast<const uint4 *>(text) + v);
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            uint32_t prev = b0 == 0 ? '\n' : (uint8_t)text[b0 - 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t c = (ws[j >> 2] >> (8 * (j & 3))) & 0xffu;
                if (prev == '\n' && c == 'c' && b0 + j < len) mask |= 1u << j;
                prev = c;
            }
        }
        const uint32_t cnt = __popc(mask);
        uint32_t inc = cnt;                                     // inclusive scan inside the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; ++k) {
                const uint32_t x = wsum[k];
                wsum[k] = tot;
                tot += x;
            }
            base_s = tot ? atomicAdd(n_starts, tot) : 0u;
        }
        __syncthreads();
        uint32_t at = base_s + wsum[wid] + inc - cnt;
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            starts[at++] = (uint32_t)(b0 + j);
        }
        __syncthreads();                                        // wsum / base_s are reused by the next round
    }
}

__global__ void __launch_bounds__(256) k_parse_lines(const char *__restrict__ text, int64_t len, CtgTable tab,
                                                      const uint32_t *__restrict__ starts,
                                                      const unsigned int *__restrict__ n_starts,
                                                      int32_t *__restrict__ ctg, int32_t *__restrict__ qs,
                                                      int32_t *__restrict__ qe, unsigned long long *__restrict__ n_out,
                                                      unsigned long long cap)
{
    const uint32_t nl = *n_starts;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < nl; k += gridDim.x * blockDim.x) {
        const int64_t i = starts[k];
        // field 0: the contig name
        int64_t p = i;
        int c;
        auto at = [&](int64_t k_) -> int { return k_ < len ? (int)(uint8_t)text[k_] : '\n'; };
        while ((c = at(p)) != '\t' && c != '\n' && c != 0) ++p;
        const int64_t nlen = p - i;
        if (c != '\t' || nlen < 3 || nlen >= 40) continue;                  // < 3 fields, or strlen(ctg) >= 40
        if (text[i] != 'c' || text[i + 1] != 'h' || text[i + 2] != 'r') continue;
        const int64_t f1 = p + 1;
        p = f1;
        while ((c = at(p)) != '\t' && c != '\n' && c != 0) ++p;
        if (c != '\t') continue;                                            // only 2 fields
        const int64_t f2 = p + 1;
        const int32_t st = dev_atol(text, f1, len), en = dev_atol(text, f2, len);
        if (!(en > 0)) continue;
        // get_id
        int32_t id = -1;
        for (uint32_t h = fnv1a(text + i, (int)nlen) & tab.mask;; h = (h + 1) & tab.mask) {
            const int32_t kk = tab.slots[h];
            if (kk < 0) break;
            const char *nm = tab.names + 40 * kk;
            bool same = nm[nlen] == 0;
            for (int j = 0; same && j < (int)nlen; ++j) same = nm[j] == text[i + j];
            if (same) {
                id = kk;
                break;
            }
        }
        if (id < 0) continue;                                               // unknown contig: contributes nothing
        const unsigned long long slot = atomicAdd(n_out, 1ull);
        if (slot < cap) ctg[slot] = id, qs[slot] = st, qe[slot] = en;
    }
}


}  // namespace

extern "C" {

int igdb_dev_set_contigs(igdb_dev *d, const char *names40)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    uint32_t sz = 16;
    while (sz < 4u * (uint32_t)std::max(d->nctg, 1)) sz <<= 1;
    std::vector<int32_t> slots(sz, -1);
    for (int k = 0; k < d->nctg; ++k) {
        const char *nm = names40 + 40 * (size_t)k;
        uint32_t h = fnv1a(nm, (int)strnlen(nm, 39)) & (sz - 1);
        bool dup = false;                           // get_id returns the first contig of that name
        while (slots[h] >= 0) {
            if (strncmp(names40 + 40 * (size_t)slots[h], nm, 40) == 0) { dup = true; break; }
            h = (h + 1) & (sz - 1);
        }
        if (!dup) slots[h] = k;
    }
    if (dev_alloc(d, &d->ctg_names, 40 * (size_t)std::max(d->nctg, 1)) || dev_alloc(d, &d->ctg_slots, sz)) return -1;
    CK(cudaMemcpy(d->ctg_names, names40, 40 * (size_t)d->nctg, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d->ctg_slots, slots.data(), sz * sizeof(int32_t), cudaMemcpyHostToDevice));
    d->ctg_mask = sz - 1;
    return 0;
}

int igdb_dev_hits_begin(igdb_dev *d)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (slots_idle(d, "search") || finish_pending(d)) return -1;
    CK(cudaMemsetAsync(d->d_hits, 0, (size_t)d->nfiles * 8, d->stream));
    return 0;
}

int igdb_dev_hits_end(igdb_dev *d, uint64_t *hits_delta)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    CK(cudaMemcpyAsync(hits_delta, d->d_hits, (size_t)d->nfiles * 8, cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    return 0;
}

// One chunk of query text (whole lines): H2D, parse on the device, count into d_hits (between hits_begin and
// hits_end).  The search of this chunk is still running when the call returns: the caller reads the next chunk
// of the file meanwhile.  `text` must stay untouched until the next call on this handle has returned.
int igdb_dev_count_text(igdb_dev *d, const char *text, size_t len, int mode_v, int32_t v, int64_t *n_queries)
{
    g_err[0] = 0;
    CK(cudaSetDevice(d->device));
    if (finish_pending(d)) return -1;
    if (n_queries) *n_queries = 0;
    if (len == 0 || d->T == 0) return 0;
    cudaStream_t st = d->stream;
    if (len >= ((size_t)1 << 32)) {
        snprintf(g_err, sizeof g_err, "count_text: a chunk of %zu bytes (line offsets are 32-bit)", len);
        return -1;
    }
    if (len > d->d_text_cap) {
        if (d->d_text) CK(cudaFree(d->d_text));
        if (d->d_starts) CK(cudaFree(d->d_starts));
        d->d_text = nullptr, d->d_starts = nullptr;
        const size_t tcap = len + (len >> 2);
        CK(cudaMalloc((void **)&d->d_text, tcap + 64));                     // + padding for the 16-byte loads
        CK(cudaMalloc((void **)&d->d_starts, (tcap / 2 + 2) * sizeof(uint32_t)));
        d->d_text_cap = tcap;
    }
    // an accepted line has at least 6 characters before its newline: "chr", tab, an EMPTY start field (atol("") = 0 is
    // accepted by parse_bed), tab, one digit -- and the last line of a chunk may come without a newline
    const int64_t cap = (int64_t)(len / 6 + 2);
    if (ensure_query_cap(d, cap)) return -1;
    CK(cudaMemcpyAsync(d->d_text, text, len, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d->plan_tot + 4, 0, 8, st));
    const CtgTable tab{d->ctg_names, d->ctg_slots, d->ctg_mask};
    // plan_tot[4] = accepted queries (64-bit), plan_tot[5] = candidate line starts (its low word)
    unsigned int *n_starts = reinterpret_cast<unsigned int *>(d->plan_tot + 5);
    CK(cudaMemsetAsync(d->plan_tot + 5, 0, 8, st));
    const unsigned g1 = (unsigned)std::min<size_t>((len / 16 + 255) / 256 + 1, (size_t)d->sm_count * 16);
    k_line_starts<<<g1, 256, 0, st>>>(d->d_text, (int64_t)len, d->d_starts, n_starts);
    const unsigned g2 = (unsigned)std::min<size_t>((len / 8 + 255) / 256 + 1, (size_t)d->sm_count * 16);
    k_parse_lines<<<g2, 256, 0, st>>>(d->d_text, (int64_t)len, tab, d->d_starts, n_starts, d->q_ctg, d->q_qs, d->q_qe,
                                      d->plan_tot + 4, (unsigned long long)cap);
    d->launches += 2;
    CK(cudaGetLastError());
    unsigned long long *h_n = (unsigned long long *)(d->h_scalar + 4);
    CK(cudaMemcpyAsync(h_n, d->plan_tot + 4, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (*h_n > (unsigned long long)cap) {                   // cannot happen (see cap); never drop queries silently
        snprintf(g_err, sizeof g_err, "count_text: %llu accepted lines in %zu bytes", *h_n, len);
        return -1;
    }
    const int64_t n = (int64_t)*h_n;
    if (n_queries) *n_queries = n;
    if (n == 0) return 0;
    igdb_dev::Pending p;
    p.mode_v = mode_v != 0, p.n = n, p.ctg = d->q_ctg, p.qs = d->q_qs, p.qe = d->q_qe, p.v = v;
    p.hits = d->d_hits, p.st = st;
    p.mode = count_mode_for(d, p.mode_v, p.v);
    if (enqueue_count(d, p)) return -1;
    p.active = true;
    d->pending = p;
    return 0;
}

}  // extern "C"
