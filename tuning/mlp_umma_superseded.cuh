// mlp_umma.cuh -- bf16 tensor-core path for the neural-ODE parameter VJP (BASELINE config C4, dtype BF16_F32ACC).
//
// The weight gradient of the hidden layer,  dW2 = sum over (reverse step, RK stage, member) of  (h b_j Delta2) H1',
// is a [64 x K] x [K x 64] contraction with K = 6 S N (737 280 for C4): the one GEMM-shaped piece of the hot path
// ("over a batch the outer products become GEMMs").  In bf16 mode the reverse kernel writes the two
// operands as K-major bf16 "tapes" and this kernel contracts them on the 5th-generation tensor cores:
//   tcgen05.mma.cta_group::1.kind::f16, M = 128 (rows 64..127 of A are a zero tile), N = 64, K = 16 per instruction,
//   operands in shared memory in the canonical no-swizzle K-major layout (8 x 16 B core matrices), fp32 accumulator
//   in TMEM (64 columns), completion through tcgen05.commit -> mbarrier, epilogue tcgen05.ld -> registers -> global.
// SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (commit).
// Each CTA owns a contiguous K range; the per-CTA partial 64 x 64 tiles are summed in CTA order (deterministic).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "ode_tsit5.cuh"

namespace b200adj {

constexpr int UM_BK = 64;                  // K elements per smem stage
constexpr int UM_M = 128, UM_N = 64;
constexpr uint32_t UM_TMEM_COLS = 64;

struct UmmaArgs {
    const __nv_bfloat16* TA;               // [64][Ktot]  (h b_j Delta2), K-major
    const __nv_bfloat16* TB;               // [64][Ktot]  H1, K-major
    float* partials;                       // [gridDim][64*64]  row-major [i][j]
    int64_t Ktot;                          // multiple of UM_BK
};

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start_address[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type=0 (no swizzle) [61,64)
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t UM_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(UM_N >> 3) << 17) | ((uint32_t)(UM_M >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128) mlp_dw2_umma_kernel(const __grid_constant__ UmmaArgs a) {
    __shared__ __align__(128) unsigned char sA[UM_M * UM_BK * 2];     // 16 KB, core-matrix layout
    __shared__ __align__(128) unsigned char sB[UM_N * UM_BK * 2];     // 8 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // rows 64..127 of A: zero tile, written once
    for (int x = tid; x < (UM_M / 2) * UM_BK * 2 / 16; x += 128) reinterpret_cast<uint4*>(sA + (UM_M / 2) * UM_BK * 2)[x] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(UM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    const int64_t nchunks = a.Ktot / UM_BK;
    const int64_t c0 = nchunks * blockIdx.x / gridDim.x, c1 = nchunks * (blockIdx.x + 1) / gridDim.x;
    uint32_t phase = 0;
    for (int64_t ch = c0; ch < c1; ch++) {
        const int64_t k0 = ch * UM_BK;
        // global (K-major rows of 64 bf16 = 128 B) -> shared core matrices: element (r, k) at ((r/8)*(BK/8) + k/8)*128 + (r%8)*16 + (k%8)*2
        for (int x = tid; x < 2 * 64 * (UM_BK / 8); x += 128) {
            const int isB = x >= 64 * (UM_BK / 8), y = x - isB * 64 * (UM_BK / 8), r = y / (UM_BK / 8), kc = y % (UM_BK / 8);
            const uint4 v = __ldg(reinterpret_cast<const uint4*>((isB ? a.TB : a.TA) + (int64_t)r * a.Ktot + k0 + kc * 8));
            *reinterpret_cast<uint4*>((isB ? sB : sA) + ((r / 8) * (UM_BK / 8) + kc) * 128 + (r % 8) * 16) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int s = 0; s < UM_BK / 16; s++) {
                const uint64_t ad = umma_smem_desc(smem_u32(sA) + s * 256, 128, (UM_BK / 8) * 128);
                const uint64_t bd = umma_smem_desc(smem_u32(sB) + s * 256, 128, (UM_BK / 8) * 128);
                umma_f16(tmem_d, ad, bd, UM_IDESC, (ch > c0 || s > 0) ? 1u : 0u);
            }
            umma_commit(&bar);              // arrives when the MMAs above have finished reading shared memory
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: TMEM lane = row i of dW2 (warps 0, 1 hold rows 0..63), column = j
    if (warp < 2 && c1 > c0) {
        float* out = a.partials + (int64_t)blockIdx.x * 64 * 64 + (int64_t)tid * 64;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 16) {
            uint32_t r[16];
            tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + cb, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q++) out[cb + q] = __uint_as_float(r[q]);
        }
    } else if (warp < 2) {
        float* out = a.partials + (int64_t)blockIdx.x * 64 * 64 + (int64_t)tid * 64;
        for (int q = 0; q < 64; q++) out[q] = 0.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(UM_TMEM_COLS) : "memory");
}

// dp[OW2 + j*64 + i] = sum over CTAs (in CTA order) of partials[cta][i][j]
__global__ void mlp_dw2_reduce_kernel(const float* partials, float* dp_w2, int nctas) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= 64 * 64) return;
    const int i = x / 64, j = x % 64;
    double acc = 0;
    for (int c = 0; c < nctas; c++) acc += (double)partials[(int64_t)c * 4096 + x];
    dp_w2[j * 64 + i] = (float)acc;
}

}  // namespace b200adj
