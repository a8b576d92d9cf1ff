// mlp_tc.cuh -- the neural-ODE family (BASELINE config C4, dtype BF16_F32ACC) with EVERY GEMM-shaped piece of the time loop on
// the 5th-generation tensor cores: bf16 operands in shared memory, fp32 accumulators in TMEM, tcgen05.mma issued by one
// thread, no operand tapes and no reduction over members on the CUDA cores.
//
// Round-2 layout (the round-1 kernel put 128 members on one CTA, one thread per member: at the BASELINE size N = 4096 that is
// 32 CTAs on 148 SMs, each a serial chain of 128 x 128 MUFU.TANH + MMA round trips per stage):
//   one CTA = 32 ensemble members x 4 warps.  Thread (w, m) = warp w, lane m owns member m's feature QUARTER [16 w, 16 w + 16):
//   16 tanh per layer instead of 64, 128 CTAs at N = 4096, two CTAs per SM at large N.  The M = 128 rows of every member MMA
//   hold FOUR COPIES of the 32 member rows (row 32 c + m = member m), because a warp can only read the 32 TMEM lanes it owns:
//   warp w reads its accumulator quarter from copy w (lane 32 w + m, columns 16 w ..).  The member state (u, lam, RK stages,
//   d = 2) is kept redundantly by the 4 threads of a member; the two 2-vectors a stage returns (f, J'lam) are summed over the
//   quarters through shared memory.
//
//   member GEMMs (M = 128 = 4 x 32 members, N = 64, K = 64; K-major operands):
//     forward   Z2 = H1 W2'          A = tile TB (H1, 4 copies),      B = W2   (n = out, k = in)
//     backward  dH1 = dZ2 W2         A = tile TA (wt dZ2, 4 copies),  B = W2^T (n = in,  k = out)
//   gradient GEMMs (K = the 32 members of copy 0; the SAME tiles read as MN-major operands -- element (m, f) of a member tile
//   sits at (m/8) ROW + (f/8) 128 + (m%8) 16 + (f%8) 2, which is at once the canonical K-major layout of [member x feature] and
//   the canonical MN-major layout of [feature x member]; probed by tuning/tc_probe.cu), accumulated in TMEM over the whole
//   reverse pass:
//     G1 [128 x 80] += [wt dZ2 | wt dZ1]' [H1 | y0 y1 1 | 0]   ->  dW2, db2 (rows 0..63), dW1, db1 (rows 64..127)
//     G2 [128 x 16] += [H2 | 1 | 0]' [wt L0, wt L1 | 0]        ->  dW3' (rows 0..63), db3 (row 64)
//   wt = h b_j is folded into the cotangent side before the bf16 rounding; the backward member GEMM is linear, so the
//   vector-Jacobian product is recovered by dividing by wt.
//
// Per adjoint stage: 2 member GEMMs (4 MMAs each) + 2 gradient GEMMs (2 MMAs each); per forward stage: 1 member GEMM.
// tanh is the hardware tanh.approx.f32 (relative error 2^-11, below the bf16 rounding of the operands).
// Reference functions replaced: as mlp.cuh (sense functor, split_states, vecjacobian!, ReverseLossCallback).
// SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "mlp.cuh"
#include "umma.cuh"

namespace b200adj {

constexpr int TC_MEM = 32;                                  // members per CTA
constexpr int TC_M = 128;                                   // threads per CTA = rows of the member MMAs (4 copies x 32 members)
constexpr int TC_TA_F = 128, TC_TB_F = 80, TC_TC_F = 16;      // feature widths of the tiles
constexpr uint32_t TC_TMEM_COLS = 256;                      // D: 0..63, G1: 64..143, G2: 144..159
constexpr uint32_t TC_COL_D = 0, TC_COL_G1 = 64, TC_COL_G2 = 144;

struct TcSmem {
    alignas(128) unsigned char TA[TC_M * TC_TA_F * 2];      // [row][wt dZ2 (64) | wt dZ1 (64)]; dZ2 in all 4 copies, dZ1 in copy 0
    alignas(128) unsigned char TH[TC_MEM * TC_TA_F * 2];    // [member][H2 (64) | 1 | 0 ...]
    alignas(128) unsigned char TB[TC_M * TC_TB_F * 2];      // [row][H1 (64) | y0 y1 1 | 0 ...]; H1 in all 4 copies
    alignas(128) unsigned char TC[TC_MEM * TC_TC_F * 2];    // [member][wt L0, wt L1 | 0 ...]
    alignas(128) unsigned char W2[64 * 64 * 2];             // (n = out i, k = in j)  = W2[i][j]
    alignas(128) unsigned char W2T[64 * 64 * 2];            // (n = in j,  k = out i) = W2[i][j]
    float W1a[64], W1b[64], b1[64], b2[64], W3a[64], W3b[64], b3[2];
    float red[2][4][TC_MEM][2];                             // per-quarter partial sums of the 2-vectors a stage returns (double-buffered)
    alignas(8) uint64_t barM, barG;
    uint32_t tmem;
};

__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ constexpr uint32_t tc_idesc(int M, int N, int amn, int bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 16-byte chunk kc (features 8 kc .. 8 kc + 7) of row r in a tile of F features
template <int F> __device__ __forceinline__ uint4* tc_chunk(unsigned char* tile, int r, int kc) {
    return reinterpret_cast<uint4*>(tile + (r >> 3) * (F / 8) * 128 + kc * 128 + (r & 7) * 16);
}

struct TcState {               // per-thread pipeline bookkeeping (identical in all threads)
    uint32_t phM = 0, phG = 0, rb = 0;
    bool gpend = false, gfirst = true;
};

__device__ __forceinline__ void tc_setup(TcSmem& s, const float* p) {
    const int t = threadIdx.x, m = t & 31, w = t >> 5;
    for (int x = t; x < 64 * 64; x += TC_M) {
        const int j = x / 64, i = x % 64;                                     // p[OW2 + j*64 + i] = W2[i][j]
        const __nv_bfloat16 wv = __float2bfloat16(p[MLP_OW2 + x]);
        *reinterpret_cast<__nv_bfloat16*>(s.W2 + (i >> 3) * 1024 + (j >> 3) * 128 + (i & 7) * 16 + (j & 7) * 2) = wv;
        *reinterpret_cast<__nv_bfloat16*>(s.W2T + (j >> 3) * 1024 + (i >> 3) * 128 + (j & 7) * 16 + (i & 7) * 2) = wv;
    }
    if (t < 64) {
        s.W1a[t] = p[MLP_OW1 + t]; s.W1b[t] = p[MLP_OW1 + 64 + t]; s.b1[t] = p[MLP_OB1 + t]; s.b2[t] = p[MLP_OB2 + t];
        s.W3a[t] = p[MLP_OW3 + t * 2]; s.W3b[t] = p[MLP_OW3 + t * 2 + 1];
    }
    if (t < 2) s.b3[t] = p[MLP_OB3 + t];
    // constant parts of the tiles: TH features 64.. = [1, 0, ...], TB features 64..79 = 0 (y, 1 written per stage), TC = 0,
    // TA features 64..127 of the copies 1..3 = 0 (never read by a gradient GEMM, kept finite)
    const uint4 zero = make_uint4(0, 0, 0, 0);
    *tc_chunk<TC_TA_F>(s.TH, m, 8 + 2 * w) = w == 0 ? make_uint4(0x00003F80u, 0, 0, 0) : zero;       // bf16(1.0) = 0x3F80
    *tc_chunk<TC_TA_F>(s.TH, m, 9 + 2 * w) = zero;
    *tc_chunk<TC_TB_F>(s.TB, t, 8) = zero;
    *tc_chunk<TC_TB_F>(s.TB, t, 9) = zero;
#pragma unroll
    for (int kc = 8; kc < 16; kc++) *tc_chunk<TC_TA_F>(s.TA, t, kc) = zero;
    if (w == 0) { *tc_chunk<TC_TC_F>(s.TC, m, 0) = zero; *tc_chunk<TC_TC_F>(s.TC, m, 1) = zero; }
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 0) { mbar_init(&s.barM, 1); mbar_init(&s.barG, 1); mbar_fence_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_wait_grad(TcSmem& s, TcState& st) {
    if (st.gpend) { mbar_wait(&s.barG, st.phG); st.phG ^= 1; st.gpend = false; asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
}

// F = f(y) for this thread's member; leaves H1 (bf16, 4 row copies) in TB and this thread's H2 quarter in registers
template <bool GRAD>
__device__ __forceinline__ void tc_forward(TcSmem& s, TcState& st, float y0, float y1, float* F, float* H2q) {
    const int t = threadIdx.x, m = t & 31, w = t >> 5;
    uint4 row[2];                                        // this thread's H1 quarter, bf16, computed while the previous stage's
#pragma unroll                                           // gradient GEMMs may still be reading the tiles
    for (int c = 0; c < 2; c++) {
        float h[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { const int j = 16 * w + 8 * c + q; h[q] = tanh_fast(fmaf(s.W1a[j], y0, fmaf(s.W1b[j], y1, s.b1[j]))); }
        row[c] = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    }
    if (GRAD) tc_wait_grad(s, st);
#pragma unroll
    for (int cp = 0; cp < 4; cp++) {
        *tc_chunk<TC_TB_F>(s.TB, 32 * cp + m, 2 * w) = row[0];
        *tc_chunk<TC_TB_F>(s.TB, 32 * cp + m, 2 * w + 1) = row[1];
    }
    if (GRAD && w == 0) *tc_chunk<TC_TB_F>(s.TB, m, 8) = make_uint4(pack_bf16(y0, y1), pack_bf16(1.0f, 0.0f), 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 4; k++)
            umma_f16(s.tmem + TC_COL_D, umma_smem_desc(smem_u32(s.TB) + k * 256, 128, (TC_TB_F / 8) * 128),
                     umma_smem_desc(smem_u32(s.W2) + k * 256, 128, 1024), tc_idesc(128, 64, 0, 0), k > 0);
        umma_commit(&s.barM);
    }
    mbar_wait(&s.barM, st.phM); st.phM ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float f0 = 0.0f, f1 = 0.0f;
    {
        uint32_t r[16];
        tmem_ld16(s.tmem + ((uint32_t)(w * 32) << 16) + TC_COL_D + 16 * w, r);       // copy w of member m, columns 16 w ..
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int n = 16 * w + q;
            const float h2 = tanh_fast(__uint_as_float(r[q]) + s.b2[n]);
            H2q[q] = h2;
            f0 = fmaf(s.W3a[n], h2, f0); f1 = fmaf(s.W3b[n], h2, f1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // sum over the four feature quarters (fixed order => all four threads of a member hold the same bits)
    float (*red)[TC_MEM][2] = s.red[st.rb]; st.rb ^= 1;
    red[w][m][0] = f0; red[w][m][1] = f1;
    __syncthreads();
    F[0] = s.b3[0] + ((red[0][m][0] + red[1][m][0]) + (red[2][m][0] + red[3][m][0]));
    F[1] = s.b3[1] + ((red[0][m][1] + red[1][m][1]) + (red[2][m][1] + red[3][m][1]));
}

// J = (df/dy)' L at the point of the last tc_forward; issues the gradient GEMMs with weight wt (members with valid = false
// contribute nothing)
template <bool GRAD = true>
__device__ __forceinline__ void tc_backward(TcSmem& s, TcState& st, float wt, float L0, float L1, bool valid, const float* H2q, float* J) {
    const int t = threadIdx.x, m = t & 31, w = t >> 5;
    const float wv = valid ? wt : 0.0f;
    uint4 dzc[2], hhc[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        float dz[8], hh[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int n = 16 * w + 8 * c + q;
            hh[q] = H2q[8 * c + q];
            dz[q] = wv * fmaf(s.W3a[n], L0, s.W3b[n] * L1) * (1.0f - hh[q] * hh[q]);
        }
        dzc[c] = make_uint4(pack_bf16(dz[0], dz[1]), pack_bf16(dz[2], dz[3]), pack_bf16(dz[4], dz[5]), pack_bf16(dz[6], dz[7]));
        hhc[c] = make_uint4(pack_bf16(hh[0], hh[1]), pack_bf16(hh[2], hh[3]), pack_bf16(hh[4], hh[5]), pack_bf16(hh[6], hh[7]));
    }
#pragma unroll
    for (int cp = 0; cp < 4; cp++) {
        *tc_chunk<TC_TA_F>(s.TA, 32 * cp + m, 2 * w) = dzc[0];
        *tc_chunk<TC_TA_F>(s.TA, 32 * cp + m, 2 * w + 1) = dzc[1];
    }
    *tc_chunk<TC_TA_F>(s.TH, m, 2 * w) = hhc[0];
    *tc_chunk<TC_TA_F>(s.TH, m, 2 * w + 1) = hhc[1];
    if (w == 0) *tc_chunk<TC_TC_F>(s.TC, m, 0) = make_uint4(pack_bf16(wv * L0, wv * L1), 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 4; k++)
            umma_f16(s.tmem + TC_COL_D, umma_smem_desc(smem_u32(s.TA) + k * 256, 128, (TC_TA_F / 8) * 128),
                     umma_smem_desc(smem_u32(s.W2T) + k * 256, 128, 1024), tc_idesc(128, 64, 0, 0), k > 0);
        umma_commit(&s.barM);
    }
    mbar_wait(&s.barM, st.phM); st.phM ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float j0 = 0.0f, j1 = 0.0f;
    {
        uint32_t r[16];
        tmem_ld16(s.tmem + ((uint32_t)(w * 32) << 16) + TC_COL_D + 16 * w, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float dz1[16];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const uint4 hv = *tc_chunk<TC_TB_F>(s.TB, m, 2 * w + c);      // this member's H1 (bf16), features 16 w + 8 c ..
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = 16 * w + 8 * c + q;
                const float h1 = __uint_as_float((q & 1) ? (hw[q >> 1] & 0xFFFF0000u) : (hw[q >> 1] << 16));
                const float d = __uint_as_float(r[8 * c + q]) * (1.0f - h1 * h1);      // wt dZ1
                dz1[8 * c + q] = d;
                j0 = fmaf(s.W1a[j], d, j0); j1 = fmaf(s.W1b[j], d, j1);
            }
        }
        *tc_chunk<TC_TA_F>(s.TA, m, 8 + 2 * w) = make_uint4(pack_bf16(dz1[0], dz1[1]), pack_bf16(dz1[2], dz1[3]), pack_bf16(dz1[4], dz1[5]), pack_bf16(dz1[6], dz1[7]));
        *tc_chunk<TC_TA_F>(s.TA, m, 9 + 2 * w) = make_uint4(pack_bf16(dz1[8], dz1[9]), pack_bf16(dz1[10], dz1[11]), pack_bf16(dz1[12], dz1[13]), pack_bf16(dz1[14], dz1[15]));
    }
    float (*red)[TC_MEM][2] = s.red[st.rb]; st.rb ^= 1;
    red[w][m][0] = j0; red[w][m][1] = j1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const float inv = 1.0f / wt;
    J[0] = ((red[0][m][0] + red[1][m][0]) + (red[2][m][0] + red[3][m][0])) * inv;
    J[1] = ((red[0][m][1] + red[1][m][1]) + (red[2][m][1] + red[3][m][1])) * inv;
    if (GRAD && t == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc0 = st.gfirst ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < TC_MEM / 16; k++)        // K = 16 members per MMA = 2 member groups of 8; the 32 members of copy 0
            umma_f16(s.tmem + TC_COL_G1, umma_smem_desc(smem_u32(s.TA) + k * 2 * (TC_TA_F / 8) * 128, (TC_TA_F / 8) * 128, 128),
                     umma_smem_desc(smem_u32(s.TB) + k * 2 * (TC_TB_F / 8) * 128, (TC_TB_F / 8) * 128, 128), tc_idesc(128, TC_TB_F, 1, 1), (k > 0) ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < TC_MEM / 16; k++)
            umma_f16(s.tmem + TC_COL_G2, umma_smem_desc(smem_u32(s.TH) + k * 2 * (TC_TA_F / 8) * 128, (TC_TA_F / 8) * 128, 128),
                     umma_smem_desc(smem_u32(s.TC) + k * 2 * (TC_TC_F / 8) * 128, (TC_TC_F / 8) * 128, 128), tc_idesc(128, TC_TC_F, 1, 1), (k > 0) ? 1u : acc0);
        umma_commit(&s.barG);
    }
    if (GRAD) { st.gfirst = false; st.gpend = true; }
}

__device__ __forceinline__ void tc_teardown(TcSmem& s) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tmem), "r"(TC_TMEM_COLS) : "memory");
}

// ---- forward ensemble solve (fixed-step Tsit5) ----
template <int UNUSED = 0>
__global__ void __launch_bounds__(TC_M) mlp_tc_forward_kernel(const __grid_constant__ MlpArgs<float> a) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(tc_smem_raw);
    const int64_t N = a.N, base = (int64_t)blockIdx.x * TC_MEM;
    const int t = threadIdx.x, m = t & 31;
    const bool live = base + m < N, writer = live && (t >> 5) == 0;        // one of the four threads of a member stores
    const int64_t col = live ? base + m : N - 1;
    tc_setup(s, a.p);
    TcState st;
    float u[2], kf[7][2], H2q[16], F[2];
    u[0] = a.u0[col]; u[1] = a.u0[N + col];
    if (writer) {
        a.ckpt[col] = u[0]; a.ckpt[N + col] = u[1];
        if (a.saved) { const int ks = a.save_of_step[0]; if (ks >= 0) { a.saved[((int64_t)ks * 2) * N + col] = u[0]; a.saved[((int64_t)ks * 2 + 1) * N + col] = u[1]; } }
    }
    tc_forward<false>(s, st, u[0], u[1], kf[0], H2q);
    for (int n = 0; n < a.S; n++) {
        float y[2];
#pragma unroll 1
        for (int sg = 1; sg <= 6; sg++) {
#pragma unroll
            for (int c = 0; c < 2; c++) {
                double acc = (double)u[c];
                for (int j = 0; j < sg; j++) acc = fma(a.tb.hA[sg][j], (double)kf[j][c], acc);
                y[c] = (float)acc;
            }
            tc_forward<false>(s, st, y[0], y[1], F, H2q);
            if (sg < 6) { kf[sg][0] = F[0]; kf[sg][1] = F[1]; }
        }
        if (writer && a.kst) {
            // the dense forward solution of this step (k1..k6, k7 = f(u_{n+1})): the reverse pass reads it back instead of
            // repeating the six stage evaluations (6 of its 18 tensor-core round trips per step), 56 B per member-step
            float* ks_ = a.kst + ((int64_t)n * 14) * N + col;
#pragma unroll
            for (int j = 0; j < 6; j++) { ks_[(int64_t)(2 * j) * N] = kf[j][0]; ks_[(int64_t)(2 * j + 1) * N] = kf[j][1]; }
            ks_[(int64_t)12 * N] = F[0]; ks_[(int64_t)13 * N] = F[1];
        }
        u[0] = y[0]; u[1] = y[1]; kf[0][0] = F[0]; kf[0][1] = F[1];         // FSAL: f(u_{n+1})
        if (writer) {
            a.ckpt[((int64_t)(n + 1) * 2) * N + col] = u[0]; a.ckpt[((int64_t)(n + 1) * 2 + 1) * N + col] = u[1];
            if (a.saved) { const int ks = a.save_of_step[n + 1]; if (ks >= 0) { a.saved[((int64_t)ks * 2) * N + col] = u[0]; a.saved[((int64_t)ks * 2 + 1) * N + col] = u[1]; } }
        }
    }
    if (writer && a.status) a.status[col] = (isfinite(u[0]) && isfinite(u[1])) ? 0 : 1;
    tc_teardown(s);
}

// ---- fused reverse pass (same stage sequence as mlp_reverse_kernel): InterpolatingAdjoint, or GaussAdjoint (GAUSS: seven
// adjoint stages without gradient GEMMs, then the gradient GEMMs at the three Gauss-Legendre nodes of the step, weight (h/2) w_g,
// accumulated in the same TMEM tiles) ----
template <int COST, bool GAUSS = false>
__global__ void __launch_bounds__(TC_M) mlp_tc_reverse_kernel(const __grid_constant__ MlpArgs<float> a) {
    extern __shared__ __align__(128) unsigned char tc_smem_raw[];
    TcSmem& s = *reinterpret_cast<TcSmem*>(tc_smem_raw);
    const int64_t N = a.N, base = (int64_t)blockIdx.x * TC_MEM;
    const int t = threadIdx.x, m = t & 31;
    const bool live = base + m < N, writer = live && (t >> 5) == 0;
    const int64_t col = live ? base + m : N - 1;
    const Tsit5Tables& tb = a.tb;
    tc_setup(s, a.p);
    TcState st;
    float lam[2] = {0.0f, 0.0f}, uhi[2], ulo[2], kf[7][2], ka[7][2], H2q[16], F[2], J[2];
    auto cotangent = [&](int ks, const float* yy) {
        if (COST == COST_EXPLICIT) { lam[0] += a.dLdu[((int64_t)ks * 2) * N + col]; lam[1] += a.dLdu[((int64_t)ks * 2 + 1) * N + col]; }
        else { lam[0] += (float)(a.cost_a[0] * (double)yy[0] + a.cost_b[0]); lam[1] += (float)(a.cost_a[1] * (double)yy[1] + a.cost_b[1]); }
    };
    uhi[0] = a.ckpt[((int64_t)a.S * 2) * N + col]; uhi[1] = a.ckpt[((int64_t)a.S * 2 + 1) * N + col];
    { const int ks = a.save_of_step[a.S]; if (ks >= 0) cotangent(ks, uhi); }
    if (!a.kst) tc_forward<true>(s, st, uhi[0], uhi[1], kf[6], H2q);    // f(u_S) = forward k7 of the last step
    for (int n = a.S - 1; n >= 0; n--) {
        ulo[0] = a.ckpt[((int64_t)n * 2) * N + col]; ulo[1] = a.ckpt[((int64_t)n * 2 + 1) * N + col];
        // ---- forward stages k1..k7 of [t_n, t_{n+1}]: read back from the forward pass (a.kst) or recomputed ----
        if (a.kst) {
            const float* ks_ = a.kst + ((int64_t)n * 14) * N + col;
#pragma unroll
            for (int j = 0; j < 7; j++) { kf[j][0] = ks_[(int64_t)(2 * j) * N]; kf[j][1] = ks_[(int64_t)(2 * j + 1) * N]; }
        } else {
            tc_forward<true>(s, st, ulo[0], ulo[1], kf[0], H2q);
#pragma unroll 1
            for (int sg = 1; sg <= 5; sg++) {
                float y[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double acc = (double)ulo[c];
                    for (int j = 0; j < sg; j++) acc = fma(tb.hA[sg][j], (double)kf[j][c], acc);
                    y[c] = (float)acc;
                }
                tc_forward<true>(s, st, y[0], y[1], kf[sg], H2q);
            }
        }
        // ---- adjoint stages 0..5 (GaussAdjoint: 0..6, the 7th derivative feeds the dense output of the adjoint step) ----
#pragma unroll 1
        for (int sg = 0; sg <= (GAUSS ? 6 : 5); sg++) {
            float L[2], y[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                double l = (double)lam[c];
                for (int j = 0; j < sg && j < 6; j++) l = fma(tb.hA[sg][j], (double)ka[j][c], l);
                L[c] = (float)l;
                double yv;
                if (sg == 0) yv = (double)uhi[c];
                else if (sg >= 5) yv = (double)ulo[c];
                else { yv = (double)ulo[c]; for (int j = 0; j < 7; j++) yv = fma(tb.hBst[sg - 1][j], (double)kf[j][c], yv); }
                y[c] = (float)yv;
            }
            tc_forward<true>(s, st, y[0], y[1], F, H2q);
            if (GAUSS) tc_backward<false>(s, st, 1.0f, L[0], L[1], live, H2q, J);
            else tc_backward<true>(s, st, (float)tb.hA[6][sg], L[0], L[1], live, H2q, J);
            ka[sg][0] = J[0]; ka[sg][1] = J[1];
        }
        if (GAUSS) {
#pragma unroll 1
            for (int gq = 0; gq < 3; gq++) {
                float L[2], y[2];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    double l = (double)lam[c], yv = (double)ulo[c];
                    for (int j = 0; j < 7; j++) { l = fma(tb.hBq[gq][j], (double)ka[j][c], l); yv = fma(tb.hBq[2 - gq][j], (double)kf[j][c], yv); }
                    L[c] = (float)l; y[c] = (float)yv;
                }
                tc_forward<true>(s, st, y[0], y[1], F, H2q);
                tc_backward<true>(s, st, (float)tb.hGW[gq], L[0], L[1], live, H2q, J);
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            double l = (double)lam[c];
            for (int j = 0; j < 6; j++) l = fma(tb.hA[6][j], (double)ka[j][c], l);
            lam[c] = (float)l;
        }
        { const int ks = a.save_of_step[n]; if (ks >= 0 && !((a.flags & 1u) && n == 0)) cotangent(ks, ulo); }
        uhi[0] = ulo[0]; uhi[1] = ulo[1];
        if (!a.kst) { kf[6][0] = kf[0][0]; kf[6][1] = kf[0][1]; }
    }
    if (writer) { a.du0[col] = lam[0]; a.du0[N + col] = lam[1]; }
    // ---- parameter gradient of this CTA out of TMEM: thread t = row t of G1 / G2 ----
    tc_wait_grad(s, st);
    float* out = a.partials + (int64_t)blockIdx.x * MLP_P;
    const uint32_t lane_base = s.tmem + ((uint32_t)((t >> 5) * 32) << 16);
    const bool any = a.S > 0;
#pragma unroll 1
    for (int cb = 0; cb < TC_TB_F; cb += 16) {
        uint32_t r[16];
        tmem_ld16(lane_base + TC_COL_G1 + cb, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int g = cb + q;
            const float v = any ? __uint_as_float(r[q]) : 0.0f;
            if (t < 64) {                    // row i = t: dW2[i][g] (g < 64), db2[i] (g = 66)
                if (g < 64) out[MLP_OW2 + g * 64 + t] = v;
                else if (g == 66) out[MLP_OB2 + t] = v;
            } else {                         // row 64 + j: dW1[j][c] (g = 64 + c), db1[j] (g = 66)
                const int j = t - 64;
                if (g == 64) out[MLP_OW1 + j] = v;
                else if (g == 65) out[MLP_OW1 + 64 + j] = v;
                else if (g == 66) out[MLP_OB1 + j] = v;
            }
        }
    }
    {
        uint32_t r[16];
        tmem_ld16(lane_base + TC_COL_G2, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float v0 = any ? __uint_as_float(r[0]) : 0.0f, v1 = any ? __uint_as_float(r[1]) : 0.0f;
        if (t < 64) { out[MLP_OW3 + t * 2] = v0; out[MLP_OW3 + t * 2 + 1] = v1; }      // dW3[c][n = t]
        else if (t == 64) { out[MLP_OB3] = v0; out[MLP_OB3 + 1] = v1; }
    }
    tc_teardown(s);
}

}  // namespace b200adj
