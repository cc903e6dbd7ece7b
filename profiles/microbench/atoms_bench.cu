// Microbenchmark: cost of one warp-wide shared-memory `red.shared.add.u32` on a B200 SM as a function of how the
// 32 lanes' addresses fall on the 32 banks -- the per-CTA hits[] histogram of the count kernel issues one such
// instruction per 32 records.  One 768-thread CTA per SM (the count kernel's geometry), nbins-word histogram.
//   R   random bins                         (what a tile's records give: file indices are unrelated)
//   P   one lane per bank, random rows      (conflict-free by construction)
//   T2  two lanes per bank                  (conflict degree 2)
//   S   all lanes the same word
//   H   random bins, every other lane predicated off
//   L   the same addresses read with ld.shared (no atomic), for the crossbar floor
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_bench atoms_bench.cu ; run: ./atoms_bench [nbins]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(768, 1) k(unsigned *gout, int nbins, int iters, long long *cyc)
{
    extern __shared__ unsigned hist[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned rows = (unsigned)nbins / 32u;
    unsigned r = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned wr = (blockIdx.x * 24 + (threadIdx.x >> 5)) * 40503u + 977u;      // warp-uniform stream
    const unsigned base = (unsigned)__cvta_generic_to_shared(hist);
    unsigned acc = 0;
    const unsigned val = (r >> 27) + 1u;        // a run-time count: `add 1` would compile to ATOMS.POPC.INC
    // 16 addresses per thread, made before the clock starts: the timed loop is the shared-memory instruction alone
    unsigned a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        r = r * 1664525u + 1013904223u;
        wr = wr * 22695477u + 1u;
        unsigned b = 0;
        if (MODE == 0 || MODE == 4 || MODE == 5) b = (r >> 8) % (unsigned)nbins;
        if (MODE == 1) b = ((r >> 8) % rows) * 32u + ((lane + (wr >> 20)) & 31u);
        if (MODE == 2) b = ((r >> 8) % rows) * 32u + (((lane >> 1) + (wr >> 20)) & 31u);
        if (MODE == 3) b = (wr >> 8) % (unsigned)nbins;
        a[j] = base + 4u * b;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 5) {
                unsigned v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a[j]));
                acc += v;
            } else if (MODE == 4) {
                if (lane & 1) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[j]), "r"(val) : "memory");
            } else {
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a[j]), "r"(val) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    unsigned s = acc;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s += hist[i];
    if (s == 0xdeadbeef) gout[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int nbins)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 768, blocks = sms, iters = 8192;
    unsigned *gout;
    long long *cyc, h[1024];
    cudaMalloc(&gout, 4);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nbins * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<blocks, threads, nbins * 4>>>(gout, nbins, iters, cyc);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < blocks; ++i) c += (double)h[i];
    c /= blocks;
    // 24 warps x iters warp-instructions per SM
    printf("%-46s nbins=%6d  %8.3f ms  %7.2f cycles per warp instruction per SM (CTA clock %0.f)\n", name, nbins, ms,
           c / (24.0 * iters), c);
    cudaFree(gout), cudaFree(cyc);
}

int main(int argc, char **argv)
{
    int nbins = argc > 1 ? atoi(argv[1]) : 10016;
    nbins = nbins / 32 * 32;
    run<0>("R  red.shared random bins", nbins);
    run<1>("P  red.shared one lane per bank", nbins);
    run<2>("T2 red.shared two lanes per bank", nbins);
    run<3>("S  red.shared all lanes one word", nbins);
    run<4>("H  red.shared random bins, 16 lanes", nbins);
    run<5>("L  ld.shared random bins", nbins);
    return 0;
}
