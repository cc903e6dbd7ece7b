// igd_b200 CUDA layer -- the one exchange step of the multi-GPU count path: the sum of the per-rank hits[] over
// all ranks, done by ONE small kernel over NVLink peer memory instead of an NCCL all-reduce.
//
// One process per GPU.  Every rank owns a mailbox in its HBM -- recv[2][world][nfiles] (uint64) and
// flag[2][world][slices] -- exported to the other processes as a CUDA IPC handle.  An all-reduce of epoch e is one
// launch of k_hits_exchange per rank, the file range cut into slices of 512 files, one CTA each; per slice:
//   1. push: the rank stores its hits[] into slot [e & 1][rank] of EVERY rank's mailbox (its own included) --
//      plain coalesced stores, a few KB per peer, travelling over NVLink / NVSwitch -- then fences (system scope)
//      and raises flag[e & 1][rank] = e in each mailbox with a release store;
//   2. wait: it spins (acquire loads) until all `world` flags of its own mailbox show e;
//   3. sum: hits[f] = sum over ranks q = 0 .. world-1 of recv[e & 1][q][f], in rank order: integer addition, the
//      same result on every rank.
// Two slots by epoch parity are enough: a rank can only finish epoch e + 1 after every peer has pushed e + 1, which
// a peer does only after it has finished reading slot e & 1 in epoch e; so nobody overwrites a slot that is still
// being read.  hits[nFiles] is 15 KB at 1 905 files, 80 KB at 10 000: the exchange is latency, not bandwidth --
// which is why it is a single launch queued right behind the search kernel rather than a library collective.
#include "igd_dev_impl.cuh"

#include <cstring>

struct igdb_reducer {
    int device = 0, rank = 0, world = 1;
    int32_t nfiles = 0;
    unsigned long long *mailbox = nullptr;          // own: recv[2][world][nfiles], flag[2][world], one status word
    unsigned long long *peer[64] = {};              // every rank's mailbox as mapped here (peer[rank] == mailbox)
    unsigned long long **d_peer = nullptr;          // device copy of peer[]
    unsigned long long epoch = 0;
    bool connected = false;
};

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

constexpr unsigned long long EXCHANGE_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;
constexpr int EX_MAXCTA = 32;               // the file range is cut into at most this many slices, one CTA each

// CTA c of every rank handles slice c of the files -- push, flag[.][rank][c], wait for flag[.][q][c] of all ranks q,
// sum -- so the slices are exchanged side by side and no CTA waits for another CTA of its own GPU.
// in == out: the classic in-place all-reduce.  in != out: out = the sum and `in` is left ZEROED, ready to take the
// next batch's counts (what a search step otherwise spends one more launch on).
__global__ void __launch_bounds__(512) k_hits_exchange(const unsigned long long *in, unsigned long long *out, int zero_in,
                                                        int32_t nfiles, int rank, int world, unsigned long long epoch,
                                                        unsigned long long *const *__restrict__ peer)
{
    const int G = gridDim.x, c = blockIdx.x;
    const int per = (nfiles + G - 1) / G, f0 = min(c * per, nfiles), f1 = min(f0 + per, nfiles);
    const size_t slot = (size_t)(epoch & 1) * (size_t)world * (size_t)nfiles;
    const size_t flags = (size_t)2 * (size_t)world * (size_t)nfiles + (size_t)(epoch & 1) * (size_t)world * EX_MAXCTA;
    // 1. push
    for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        const unsigned long long v = in[f];
        for (int p = 0; p < world; ++p) peer[p][slot + (size_t)rank * (size_t)nfiles + f] = v;
        if (zero_in) const_cast<unsigned long long *>(in)[f] = 0;
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(peer[threadIdx.x] + flags + (size_t)rank * EX_MAXCTA + c, epoch);
    // 2. wait for everybody's push of this slice into MY mailbox (a peer that never arrives must not hang the GPU:
    //    after EXCHANGE_TIMEOUT_NS the wait gives up and leaves a mark that igdb_reducer_status reports)
    if ((int)threadIdx.x < world) {
        const unsigned long long *fl = peer[rank] + flags + (size_t)threadIdx.x * EX_MAXCTA + c;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(fl) != epoch) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > EXCHANGE_TIMEOUT_NS) {
                peer[rank][(size_t)2 * world * nfiles + (size_t)2 * world * EX_MAXCTA] = 1ull + threadIdx.x;    // status word
                break;
            }
        }
    }
    __syncthreads();
    // 3. sum in rank order
    const unsigned long long *mine = peer[rank] + slot;
    for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        unsigned long long s = 0;
        for (int q = 0; q < world; ++q) s += mine[(size_t)q * (size_t)nfiles + f];
        out[f] = s;
    }
}

// recv[2][world][nfiles], flag[2][world][EX_MAXCTA], one status word
size_t mailbox_words(int world, int32_t nfiles) { return (size_t)2 * world * nfiles + (size_t)2 * world * EX_MAXCTA + 1; }

}  // namespace

extern "C" {

igdb_reducer *igdb_reducer_create(int device, int rank, int world, int32_t nfiles)
{
    g_err[0] = 0;
    if (world < 1 || world > 64 || rank < 0 || rank >= world || nfiles < 0) {
        snprintf(g_err, sizeof g_err, "reducer: rank %d of %d", rank, world);
        return nullptr;
    }
    igdb_reducer *r = new igdb_reducer();
    r->device = device, r->rank = rank, r->world = world, r->nfiles = nfiles;
    const size_t bytes = mailbox_words(world, nfiles) * 8;
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&r->mailbox, bytes) != cudaSuccess ||
        cudaMemset(r->mailbox, 0, bytes) != cudaSuccess || cudaMalloc(&r->d_peer, 64 * sizeof(void *)) != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "reducer: %s", cudaGetErrorString(cudaGetLastError()));
        igdb_reducer_destroy(r);
        return nullptr;
    }
    r->peer[rank] = r->mailbox;
    return r;
}

// the CUDA IPC handle (64 bytes) of this rank's mailbox, to be handed to every other rank
int igdb_reducer_handle(igdb_reducer *r, void *out64)
{
    g_err[0] = 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaSetDevice(r->device));
    CK(cudaIpcGetMemHandle(&h, r->mailbox));
    memcpy(out64, &h, 64);
    return 0;
}

// handles: world x 64 bytes, rank-major (this rank's own entry is ignored)
int igdb_reducer_connect(igdb_reducer *r, const void *handles)
{
    g_err[0] = 0;
    CK(cudaSetDevice(r->device));
    for (int p = 0; p < r->world; ++p) {
        if (p == r->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)p, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        r->peer[p] = (unsigned long long *)ptr;
    }
    CK(cudaMemcpy(r->d_peer, r->peer, 64 * sizeof(void *), cudaMemcpyHostToDevice));
    r->connected = true;
    return 0;
}

// d_out[nfiles] (uint64, on this rank's device) = the sum over all ranks of their d_in[]; enqueued on `stream`.
// d_in == d_out: in place.  Otherwise d_in is left zeroed.  Every rank must call it the same number of times.
static int reducer_launch(igdb_reducer *r, unsigned long long *d_in, unsigned long long *d_out, void *stream)
{
    g_err[0] = 0;
    if (!r->connected && r->world > 1) {
        snprintf(g_err, sizeof g_err, "reducer: not connected");
        return -1;
    }
    if (r->nfiles == 0) return 0;
    CK(cudaSetDevice(r->device));
    if (r->world == 1) {
        if (d_in != d_out) {
            CK(cudaMemcpyAsync(d_out, d_in, (size_t)r->nfiles * 8, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
            CK(cudaMemsetAsync(d_in, 0, (size_t)r->nfiles * 8, (cudaStream_t)stream));
        }
        return 0;
    }
    ++r->epoch;
    const int grid = std::max(1, std::min(EX_MAXCTA, (r->nfiles + 511) / 512));
    k_hits_exchange<<<grid, 512, 0, (cudaStream_t)stream>>>(d_in, d_out, d_in != d_out, r->nfiles, r->rank, r->world,
                                                            r->epoch, r->d_peer);
    CK(cudaGetLastError());
    return 0;
}

int igdb_reducer_allreduce(igdb_reducer *r, unsigned long long *d_hits, void *stream)
{
    return reducer_launch(r, d_hits, d_hits, stream);
}

int igdb_reducer_allreduce_into(igdb_reducer *r, unsigned long long *d_in, unsigned long long *d_out, void *stream)
{
    return reducer_launch(r, d_in, d_out, stream);
}

// 0 = every exchange so far met all its peers; k > 0 = a wait for rank k - 1 timed out (results are then invalid).
// Synchronises the device.
int igdb_reducer_status(igdb_reducer *r)
{
    g_err[0] = 0;
    if (r->world == 1) return 0;
    CK(cudaSetDevice(r->device));
    unsigned long long st = 0;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&st, r->mailbox + mailbox_words(r->world, r->nfiles) - 1, 8, cudaMemcpyDeviceToHost));
    return (int)st;
}

void igdb_reducer_destroy(igdb_reducer *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    for (int p = 0; p < r->world; ++p)
        if (p != r->rank && r->peer[p]) cudaIpcCloseMemHandle(r->peer[p]);
    if (r->mailbox) cudaFree(r->mailbox);
    if (r->d_peer) cudaFree(r->d_peer);
    delete r;
}

}  // extern "C"
