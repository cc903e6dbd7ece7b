// Microbenchmark behind DESIGN.md's histogram decision: cost per hits[idx]++ on one B200 SM partition for
//   A  shared-memory atomicAdd, all 32 lanes active, random bins
//   B  shared-memory atomicAdd, ~25 % of the lanes active (what a scan with sparse hits issues)
//   C  non-atomic shared load/add/store of the same bins (valid only when a warp's bins are distinct)
//   D  global red.add to a CTA-private histogram in global memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_bench hist_bench.cu ; run: ./hist_bench [nbins]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned *gout, unsigned *gpriv, int nbins, int iters)
{
    extern __shared__ unsigned hist[];
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned r = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned *priv = gpriv + (size_t)blockIdx.x * nbins;
    for (int it = 0; it < iters; ++it) {
        r = r * 1664525u + 1013904223u;
        const unsigned b = (r >> 8) % (unsigned)nbins;
        if (MODE == 0) atomicAdd(&hist[b], 1u);
        if (MODE == 1) { if ((r >> 29) < 2u) atomicAdd(&hist[b], 1u); }
        if (MODE == 2) { volatile unsigned *h = hist; h[b] = h[b] + 1u; }
        if (MODE == 3) atomicAdd(&priv[b], 1u);
        if (MODE == 4) { if ((r >> 29) < 2u) { volatile unsigned *h = hist; h[b] = h[b] + 1u; } }
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) s += hist[i];
    if (s == 0xdeadbeef) gout[0] = s;
}

template <int MODE>
void run(const char *name, int nbins, double active_frac)
{
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 512, blocks = sms * 4, iters = 4096;
    unsigned *gout, *gpriv;
    cudaMalloc(&gout, 4);
    cudaMalloc(&gpriv, (size_t)blocks * nbins * 4);
    cudaMemset(gpriv, 0, (size_t)blocks * nbins * 4);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, nbins * 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<MODE><<<blocks, threads, nbins * 4>>>(gout, gpriv, nbins, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double updates = (double)blocks * threads * iters * active_frac;
    printf("%-44s nbins=%6d  %8.3f ms  %8.2f G updates/s  (%.3f ns per update per SM)\n", name, nbins, ms,
           updates / ms / 1e6, ms * 1e6 / (updates / sms));
    cudaFree(gout), cudaFree(gpriv);
}

int main(int argc, char **argv)
{
    int nbins = argc > 1 ? atoi(argv[1]) : 1905;
    run<0>("A smem atomicAdd, 32 lanes", nbins, 1.0);
    run<1>("B smem atomicAdd, 25% lanes", nbins, 0.25);
    run<2>("C smem ld/add/st (non-atomic), 32 lanes", nbins, 1.0);
    run<4>("E smem ld/add/st (non-atomic), 25% lanes", nbins, 0.25);
    run<3>("D global red.add, CTA-private hist", nbins, 1.0);
    return 0;
}
