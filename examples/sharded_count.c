/* A multi-GPU count search driven from plain C: one process per rank, no Python, no torch, no NCCL -- the whole step
 * is igd_b200_sharded_count, whose exchanges are the library's own kernels over peer memory.
 * The reference has no counterpart (it is one process, src/igd.c:21-38); what it replaces is the loop of
 * getOverlaps (src/igd_search.c:696-719) over ONE query file, here cut into `world` slices of lines.
 *
 *   cc -I include -I /usr/local/cuda/include examples/sharded_count.c -L igd_b200/lib -ligd_b200 \
 *      -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/igd_b200/lib -o sharded_count
 *   ./sharded_count db.igd queries.bed <world> [v]
 *
 * Forks `world` processes; rank r uses GPU r % (number of GPUs) -- with fewer GPUs than ranks they share a device,
 * which is slow but correct.  The 192-byte IPC handles travel through a shared anonymous mapping.  Rank 0 prints
 * `index<TAB>hits` for every file with hits and the total, like examples/count_file.c. */
#include "igd_b200.h"

#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

typedef struct {
    volatile int arrived[4];                    /* one counter per barrier */
    volatile int failed;
    unsigned char handles[64][192];
} shared_t;

static int barrier(shared_t *sh, int k, int world)
{
    __sync_fetch_and_add(&sh->arrived[k], 1);
    for (long spins = 0; sh->arrived[k] < world && !sh->failed; spins++) {
        usleep(200);
        if (spins > 5 * 60 * 5000) return -1;   /* five minutes: a rank died */
    }
    return sh->failed ? -1 : 0;
}

static int rank_main(int rank, int world, const char *db, const char *qfile, int v, shared_t *sh)
{
#define FAIL(msg)                                                        \
    do {                                                                 \
        fprintf(stderr, "rank %d: %s\n", rank, msg);                     \
        sh->failed = 1;                                                  \
        return 1;                                                        \
    } while (0)
    int ndev = igd_b200_device_count();
    if (ndev <= 0) FAIL("no CUDA device available (igd_b200 has no CPU search path)");
    const int dev = rank % ndev;
    igd_b200_t *h = igd_b200_open(db, dev);
    if (!h) FAIL(igd_b200_last_error());
    igd_b200_queries *Q = igd_b200_read_bed(qfile);
    if (!Q) FAIL("cannot read the query file");
    int32_t *ctg = malloc(sizeof(int32_t) * (size_t)(Q->n + 1));
    igd_b200_resolve(h, Q, ctg);
    const int64_t a = Q->n * rank / world, b = Q->n * (rank + 1) / world, n = b - a;

    igd_b200_sharded *s = igd_b200_sharded_create(h, dev, rank, world, Q->n > 0 ? Q->n : 1);
    if (!s || igd_b200_sharded_handles(s, sh->handles[rank])) FAIL(igd_b200_last_error());
    if (barrier(sh, 0, world)) return 1;
    if (igd_b200_sharded_connect(s, sh->handles)) FAIL(igd_b200_last_error());
    if (barrier(sh, 1, world)) return 1;        /* every block is mapped everywhere before the first store */

    const int nf = igd_b200_nfiles(h);
    int32_t *d_q = NULL;
    uint64_t *d_hits = NULL;
    cudaStream_t st;
    if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreate(&st) != cudaSuccess ||
        cudaMalloc((void **)&d_q, sizeof(int32_t) * 3 * (size_t)(n + 1)) != cudaSuccess ||
        cudaMalloc((void **)&d_hits, sizeof(uint64_t) * (size_t)(nf + 1)) != cudaSuccess)
        FAIL("CUDA set-up failed");
    cudaMemcpyAsync(d_q, ctg + a, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_q + n, Q->qs + a, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_q + 2 * n, Q->qe + a, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st);
    if (igd_b200_sharded_count(s, n, d_q, d_q + n, d_q + 2 * n, v, d_hits, st)) FAIL(igd_b200_last_error());
    uint64_t *hits = calloc((size_t)nf + 1, sizeof *hits);
    cudaMemcpyAsync(hits, d_hits, sizeof(uint64_t) * (size_t)nf, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) FAIL("the step failed on the device");
    if (igd_b200_sharded_status(s)) FAIL(igd_b200_last_error());
    if (rank == 0) {
        unsigned long long total = 0;
        for (int i = 0; i < nf; i++) {
            if (hits[i]) printf("%d\t%llu\n", i, (unsigned long long)hits[i]);
            total += hits[i];
        }
        printf("total\t%llu\n", total);
        fflush(stdout);
    }
    if (barrier(sh, 2, world)) return 1;        /* nobody unmaps a block a peer may still store into */
    free(hits), free(ctg);
    cudaFree(d_q), cudaFree(d_hits), cudaStreamDestroy(st);
    igd_b200_sharded_destroy(s);
    igd_b200_queries_free(Q);
    igd_b200_close(h);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s <db.igd> <queries.bed> <world> [v]\n", argv[0]);
        return 2;
    }
    const int world = atoi(argv[3]), v = argc > 4 ? atoi(argv[4]) : 0;
    if (world < 1 || world > 64) {
        fprintf(stderr, "world: 1 .. 64\n");
        return 2;
    }
    shared_t *sh = mmap(NULL, sizeof *sh, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (sh == MAP_FAILED) return 1;
    memset(sh, 0, sizeof *sh);
    pid_t pid[64];
    for (int r = 0; r < world; r++) {           /* fork BEFORE anything touches CUDA */
        pid[r] = fork();
        if (pid[r] == 0) _exit(rank_main(r, world, argv[1], argv[2], v, sh));
        if (pid[r] < 0) return 1;
    }
    int rc = 0;
    for (int r = 0; r < world; r++) {
        int stt = 0;
        waitpid(pid[r], &stt, 0);
        if (!WIFEXITED(stt) || WEXITSTATUS(stt) != 0) rc = 1;
    }
    return rc;
}
