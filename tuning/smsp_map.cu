// Which warps of a block share an SM sub-partition (fp64 pipe)?  Pairwise probe: warp 0 and warp w run a saturating
// DFMA loop; if they share a sub-partition the pair takes ~2x the single-warp time.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned mask, int iters, double* sink, long long* cycles) {
    const int wid = threadIdx.x >> 5;
    __shared__ long long s_t[32];
    double a[16];
    for (int j = 0; j < 16; j++) a[j] = threadIdx.x * 1e-3 + j;
    __syncthreads();
    long long t0 = clock64();
    if ((mask >> wid) & 1u) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = fma(a[j], 1.0000001, 1e-9);
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) s_t[wid] = ((mask >> wid) & 1u) ? t1 - t0 : 0;
    __syncthreads();
    if (threadIdx.x == 0) { long long m = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = s_t[w] > m ? s_t[w] : m; cycles[blockIdx.x] = m; }
    double s = 0; for (int j = 0; j < 16; j++) s += a[j];
    if (s == 12345.678) sink[0] = s;
}
int main() {
    double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8 * 148);
    const int iters = 20000;
    for (int nthreads : {512, 448}) {
        int W = nthreads / 32;
        auto run = [&](unsigned mask) { probe<<<1, nthreads>>>(mask, iters, sink, cyc); long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); return c; };
        run(1u);
        long long single = run(1u);
        printf("threads %d single-warp cycles %lld; pair(0,w)/single:", nthreads, single);
        for (int w = 1; w < W; w++) printf(" %d:%.2f", w, (double)run(1u | (1u << w)) / single);
        printf("\n  pair(1,w):");
        for (int w = 2; w < W; w++) printf(" %d:%.2f", w, (double)run(2u | (1u << w)) / single);
        printf("\n  all: %.2f  first-half: %.2f\n", (double)run((1u << W) - 1) / single, (double)run((1u << (W / 2)) - 1) / single);
    }
    return 0;
}
