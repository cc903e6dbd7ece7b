// igd_b200 CUDA layer -- the one exchange step of the multi-GPU count path: the sum of the per-rank hits[] over
// all ranks, done by ONE small kernel over NVLink peer memory instead of an NCCL all-reduce.
//
// One process per GPU.  Every rank owns a mailbox in its HBM -- recv[2][world][nfiles] (uint64) and flag[2][world]
// -- exported to the other processes as a CUDA IPC handle.  An all-reduce of epoch e is one launch of
// k_hits_exchange per rank:
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

__global__ void __launch_bounds__(1024) k_hits_exchange(unsigned long long *__restrict__ hits, int32_t nfiles, int rank,
                                                         int world, unsigned long long epoch,
                                                         unsigned long long *const *__restrict__ peer)
{
    const size_t slot = (size_t)(epoch & 1) * (size_t)world * (size_t)nfiles;
    const size_t flags = (size_t)2 * (size_t)world * (size_t)nfiles + (size_t)(epoch & 1) * (size_t)world;
    // 1. push
    for (int p = 0; p < world; ++p) {
        unsigned long long *dst = peer[p] + slot + (size_t)rank * (size_t)nfiles;
        for (int f = threadIdx.x; f < nfiles; f += blockDim.x) dst[f] = hits[f];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) st_release_sys(peer[threadIdx.x] + flags + rank, epoch);
    // 2. wait for everybody's push into MY mailbox (a peer that never arrives must not hang the GPU: after
    //    EXCHANGE_TIMEOUT_NS the wait gives up and leaves a mark that igdb_reducer_status reports)
    if ((int)threadIdx.x < world) {
        const unsigned long long *fl = peer[rank] + flags + threadIdx.x;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(fl) != epoch) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > EXCHANGE_TIMEOUT_NS) {
                peer[rank][(size_t)2 * world * nfiles + (size_t)2 * world] = 1ull + threadIdx.x;    // status word
                break;
            }
        }
    }
    __syncthreads();
    // 3. sum in rank order
    const unsigned long long *mine = peer[rank] + slot;
    for (int f = threadIdx.x; f < nfiles; f += blockDim.x) {
        unsigned long long s = 0;
        for (int q = 0; q < world; ++q) s += mine[(size_t)q * (size_t)nfiles + f];
        hits[f] = s;
    }
}

// recv[2][world][nfiles], flag[2][world], one status word
size_t mailbox_words(int world, int32_t nfiles) { return (size_t)2 * world * nfiles + (size_t)2 * world + 1; }

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

// d_hits[nfiles] (uint64, on this rank's device) becomes the sum over all ranks; enqueued on `stream`.
// Every rank must call it the same number of times.
int igdb_reducer_allreduce(igdb_reducer *r, unsigned long long *d_hits, void *stream)
{
    g_err[0] = 0;
    if (!r->connected && r->world > 1) {
        snprintf(g_err, sizeof g_err, "reducer: not connected");
        return -1;
    }
    if (r->world == 1 || r->nfiles == 0) return 0;
    CK(cudaSetDevice(r->device));
    ++r->epoch;
    const int threads = std::max(64, std::min(1024, (r->nfiles + 31) / 32 * 32));
    k_hits_exchange<<<1, threads, 0, (cudaStream_t)stream>>>(d_hits, r->nfiles, r->rank, r->world, r->epoch, r->d_peer);
    CK(cudaGetLastError());
    return 0;
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
