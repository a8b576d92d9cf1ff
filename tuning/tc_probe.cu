// tc_probe.cu -- one-CTA probe of the tcgen05 operand forms the in-loop MLP kernel needs:
//  T1  member GEMM   D1[m][n] = sum_k A[m][k] W[n][k]      A, W K-major (known-good form of mlp_umma.cuh), M=128 N=64 K=64
//  T2  gradient GEMM D2[f][g] = sum_m A[m][f] B[m][g]      the SAME shared-memory tiles read as MN-major operands, M=128 N=80 K=128
//  T3  gradient GEMM D3[f][c] = sum_m A[m][f] C[m][c]      N=16
// Each MN-major test is run with both assignments of the two descriptor strides so that one run settles the semantics.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nLAB_WAIT%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE%=;\nbra LAB_WAIT%=;\nLAB_DONE%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int amn, int bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

constexpr int AF = 128, BF = 80, CF = 16;       // feature widths of the tiles
// out layout: [5 tests][128 rows][80 cols]
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, const __nv_bfloat16* C, const __nv_bfloat16* W, float* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* sA = sm;                       // [128][128]: (m/8)*2048 + (f/8)*128 + (m%8)*16 + (f%8)*2
    unsigned char* sB = sA + 128 * AF * 2;        // [128][80]:  (m/8)*1280 + ...
    unsigned char* sC = sB + 128 * BF * 2;        // [128][16]:  (m/8)*256 + ...
    unsigned char* sW = sC + 128 * CF * 2;        // [64][64]:   (n/8)*1024 + (k/8)*128 + (n%8)*16 + (k%8)*2
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int t = threadIdx.x, warp = t >> 5;
    for (int x = t; x < 128 * AF; x += 128) { int m = x / AF, f = x % AF; *(__nv_bfloat16*)(sA + (m / 8) * (AF / 8) * 128 + (f / 8) * 128 + (m % 8) * 16 + (f % 8) * 2) = A[x]; }
    for (int x = t; x < 128 * BF; x += 128) { int m = x / BF, f = x % BF; *(__nv_bfloat16*)(sB + (m / 8) * (BF / 8) * 128 + (f / 8) * 128 + (m % 8) * 16 + (f % 8) * 2) = B[x]; }
    for (int x = t; x < 128 * CF; x += 128) { int m = x / CF, f = x % CF; *(__nv_bfloat16*)(sC + (m / 8) * (CF / 8) * 128 + (f / 8) * 128 + (m % 8) * 16 + (f % 8) * 2) = C[x]; }
    for (int x = t; x < 64 * 64; x += 128) { int n = x / 64, k = x % 64; *(__nv_bfloat16*)(sW + (n / 8) * 1024 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = W[x]; }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_s;
    // column bases: T1 0 (64), T2a 64 (80), T2b 144 (80), T3a 224 (16), T3b 240 (16)
    if (t == 0) {
        // T1: K-major A (feats 0..63 of sA), K-major W
        for (int s = 0; s < 4; s++) mma(tm + 0, desc(smem_u32(sA) + s * 256, 128, (AF / 8) * 128), desc(smem_u32(sW) + s * 256, 128, 1024), idesc(128, 64, 0, 0), s > 0);
        // T2a: MN-major, LBO field = K-block stride, SBO field = MN-block stride (cute make_umma_desc<Major::MN>, SWIZZLE_NONE)
        for (int s = 0; s < 8; s++) mma(tm + 64, desc(smem_u32(sA) + s * 2 * (AF / 8) * 128, (AF / 8) * 128, 128), desc(smem_u32(sB) + s * 2 * (BF / 8) * 128, (BF / 8) * 128, 128), idesc(128, 80, 1, 1), s > 0);
        // T2b: the two strides swapped
        for (int s = 0; s < 8; s++) mma(tm + 144, desc(smem_u32(sA) + s * 2 * (AF / 8) * 128, 128, (AF / 8) * 128), desc(smem_u32(sB) + s * 2 * (BF / 8) * 128, 128, (BF / 8) * 128), idesc(128, 80, 1, 1), s > 0);
        // T3a / T3b: N = 16
        for (int s = 0; s < 8; s++) mma(tm + 224, desc(smem_u32(sA) + s * 2 * (AF / 8) * 128, (AF / 8) * 128, 128), desc(smem_u32(sC) + s * 2 * (CF / 8) * 128, (CF / 8) * 128, 128), idesc(128, 16, 1, 1), s > 0);
        for (int s = 0; s < 8; s++) mma(tm + 240, desc(smem_u32(sA) + s * 2 * (AF / 8) * 128, 128, (AF / 8) * 128), desc(smem_u32(sC) + s * 2 * (CF / 8) * 128, 128, (CF / 8) * 128), idesc(128, 16, 1, 1), s > 0);
        commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int colbase[5] = {0, 64, 144, 224, 240}, ncol[5] = {64, 80, 80, 16, 16};
    for (int tt = 0; tt < 5; tt++)
        for (int cb = 0; cb < ncol[tt]; cb += 16) {
            uint32_t r[16];
            ld16(tm + ((uint32_t)(warp * 32) << 16) + colbase[tt] + cb, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int q = 0; q < 16; q++) out[((size_t)tt * 128 + t) * 80 + cb + q] = __uint_as_float(r[q]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> A(128 * AF), B(128 * BF), C(128 * CF), W(64 * 64);
    std::vector<float> Af(128 * AF), Bf(128 * BF), Cf(128 * CF), Wf(64 * 64);
    srand(1);
    auto fill = [](std::vector<__nv_bfloat16>& v, std::vector<float>& f) { for (size_t i = 0; i < v.size(); i++) { float x = (rand() % 2001 - 1000) / 1000.0f; v[i] = __float2bfloat16(x); f[i] = __bfloat162float(v[i]); } };
    fill(A, Af); fill(B, Bf); fill(C, Cf); fill(W, Wf);
    __nv_bfloat16 *dA, *dB, *dC, *dW; float* dout;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dC, C.size() * 2); cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dout, 5 * 128 * 80 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C.data(), C.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 5 * 128 * 80 * 4);
    const int smem = 128 * AF * 2 + 128 * BF * 2 + 128 * CF * 2 + 64 * 64 * 2 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(dA, dB, dC, dW, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> out(5 * 128 * 80);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    // references
    double e1 = 0, e2a = 0, e2b = 0, e3a = 0, e3b = 0;
    for (int m = 0; m < 128; m++) for (int n = 0; n < 64; n++) { double s = 0; for (int k = 0; k < 64; k++) s += (double)Af[m * AF + k] * Wf[n * 64 + k]; e1 = fmax(e1, fabs(s - out[(0 * 128 + m) * 80 + n])); }
    for (int f = 0; f < 128; f++) for (int g = 0; g < 80; g++) { double s = 0; for (int m = 0; m < 128; m++) s += (double)Af[m * AF + f] * Bf[m * BF + g];
        e2a = fmax(e2a, fabs(s - out[(1 * 128 + f) * 80 + g])); e2b = fmax(e2b, fabs(s - out[(2 * 128 + f) * 80 + g])); }
    for (int f = 0; f < 128; f++) for (int c = 0; c < 16; c++) { double s = 0; for (int m = 0; m < 128; m++) s += (double)Af[m * AF + f] * Cf[m * CF + c];
        e3a = fmax(e3a, fabs(s - out[(3 * 128 + f) * 80 + c])); e3b = fmax(e3b, fabs(s - out[(4 * 128 + f) * 80 + c])); }
    printf("T1 K-major member GEMM           max err %.3e\n", e1);
    printf("T2a MN-major (LBO=Kblk,SBO=MNblk) max err %.3e\n", e2a);
    printf("T2b MN-major (swapped)            max err %.3e\n", e2b);
    printf("T3a N=16 (LBO=Kblk,SBO=MNblk)     max err %.3e\n", e3a);
    printf("T3b N=16 (swapped)                max err %.3e\n", e3b);
    return 0;
}
