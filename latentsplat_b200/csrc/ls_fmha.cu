// ls_fmha.cu -- multi-head self-attention core on tcgen05 (flash attention, TF32 operands, fp32 accumulation).
//
//   O = softmax(Q K^T * scale) V      per (batch, head), Q/K/V token-major fp32 matrices read in place by TMA
//
// Forward: one CTA per 128-query block and (batch, head); 128 threads, thread t owns query row t (= TMEM lane t).
// Per 64-key block:   S = Q K^T          tcgen05.mma kind::tf32, A = Q tile (K-major), B = K tile (K-major)  -> TMEM (64 cols)
//                     online softmax     tcgen05.ld -> registers: running max / sum in the log2 domain, P = exp2(S*c - m)
//                     P -> shared memory in the K-major 128B-swizzled operand layout, fence.proxy.async
//                     PV = P V           A = P (shared), B = V tile read MN-major (32-B swizzle atoms) -> TMEM (D cols)
//                     O  = O * alpha + PV in registers (one row of D floats per thread)
// The score matrix never leaves the SM.  K and V tiles are single-buffered, but the next K tile is requested as soon as
// S is complete and the next V tile as soon as PV is complete; with D = 64 a CTA needs 96 KB of shared memory and 128 TMEM
// columns, so two CTAs share an SM and one CTA's softmax overlaps the other's tensor work.
//
// Backward (flash-attention recomputation, two kernels, no atomics):
//   k_fmha_bwd_dq   per query block: for every key block S, P, dP = dO V^T, dS = P (dP - delta) scale; dQ += dS K in TMEM
//   k_fmha_bwd_dkv  per key block, transposed formulation so that every operand produced by the threads is K-major:
//                   S^T = K Q^T, P^T, dP^T = V dO^T, dS^T;  dV += P^T dO,  dK += dS^T Q  accumulate in TMEM over the query blocks
// delta_i = sum_d dO_i O_i is a small pre-pass.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_fmha.h"
#include "ls_host.h"
#include "ls_tc.cuh"

namespace lsf {
using namespace lstc;

constexpr int BM = 128;      // query rows per CTA (forward, dq) / key rows per CTA (dkv)
constexpr int BN = 64;       // rows of the other side per iteration
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// address of the 16-byte unit holding elements [4u, 4u+4) of row `row` in a K-major 128B-swizzled [rows x 32 floats] chunk
__device__ __forceinline__ uint32_t swz(uint32_t chunk_base, int row, int unit) {
    return chunk_base + (uint32_t)row * 128u + (uint32_t)((unit ^ (row & 7)) << 4);
}

// instruction descriptor: D = f32, A = B = tf32, M = 128, N = n; b_mn: B operand MN-major
__host__ __device__ constexpr uint32_t idesc(int n, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int D>
struct Fwd {
    static constexpr int DC = D / 32;
    static constexpr int Q_BYTES = DC * BM * 128;
    static constexpr int K_BYTES = DC * BN * 128;
    static constexpr int V_BYTES = DC * BN * 128;
    static constexpr int P_BYTES = (BN / 32) * BM * 128;
    static constexpr int SMEM = Q_BYTES + K_BYTES + V_BYTES + P_BYTES + 1024 + 64;
    static constexpr int TMEM_COLS = D == 64 ? 128 : 256;        // S: 64 columns, PV: D columns
};

template <int D>
__global__ void __launch_bounds__(128, D == 64 ? 2 : 1)
k_fmha_fwd(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
           const __grid_constant__ CUtensorMap map_v, const LsFmha a) {
    using C = Fwd<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem), sK = sQ + C::Q_BYTES, sV = sK + C::K_BYTES, sP = sV + C::V_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::Q_BYTES + C::K_BYTES + C::V_BYTES + C::P_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const uint32_t b_q = smem_u32(bars), b_k = b_q + 8, b_v = b_q + 16, b_s = b_q + 24, b_o = b_q + 32;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
    const int q0 = blockIdx.x * BM;
    const int tok0 = b * a.L;                       // first token row of this batch element
    const int col0 = h * D;
    const int nblk = (a.L + BN - 1) / BN;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        for (int i = 0; i < 5; ++i) mbar_init(b_q + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tm_s = tmem, tm_o = tmem + BN;

    auto load_k = [&](int j) {
        mbar_expect_tx(b_k, C::K_BYTES);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) tma_load_2d(sK + c * (BN * 128), &map_k, b_k, col0 + 32 * c, tok0 + j * BN);
    };
    auto load_v = [&](int j) {
        mbar_expect_tx(b_v, C::V_BYTES);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) tma_load_2d(sV + c * (BN * 128), &map_v, b_v, col0 + 32 * c, tok0 + j * BN);
    };
    if (tid == 0) {
        mbar_expect_tx(b_q, C::Q_BYTES);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) tma_load_2d(sQ + c * (BM * 128), &map_q, b_q, col0 + 32 * c, tok0 + q0);
        load_k(0);
        load_v(0);
    }

    const float sl2 = a.scale * kLog2e;
    float m = -INFINITY, l = 0.f;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;        // this warp's TMEM lane quarter

    for (int j = 0; j < nblk; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        if (tid == 0) {
            if (j == 0) mbar_wait(b_q, 0);
            mbar_wait(b_k, ph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < C::DC; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_tf32(tm_s, make_desc(sQ + c * (BM * 128) + k * 32, 16, 1024, 2), make_desc(sK + c * (BN * 128) + k * 32, 16, 1024, 2),
                                idesc(BN, false), (c | k) != 0 ? 1u : 0u);
            tc_commit(b_s);
        }
        mbar_wait(b_s, ph);
        tc_fence_after();
        if (tid == 0 && j + 1 < nblk) load_k(j + 1);               // S is complete: the K tile is free

        // ---- online softmax of this thread's row ----
        uint32_t s0[32], s1[32];
        tc_ld32_nowait(tm_s + lane_base, s0);
        tc_ld32_nowait(tm_s + lane_base + 32, s1);
        tc_wait_ld();
        const int nvalid = a.L - j * BN;                           // keys of this block inside the sequence (>= 1)
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x0 = i < nvalid ? __uint_as_float(s0[i]) * sl2 : -INFINITY;
            const float x1 = i + 32 < nvalid ? __uint_as_float(s1[i]) * sl2 : -INFINITY;
            s0[i] = __float_as_uint(x0);
            s1[i] = __float_as_uint(x1);
            mx = fmaxf(mx, fmaxf(x0, x1));
        }
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2(m - m_new);                        // 0 on the first block (m = -inf)
        float rs = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {                              // 16-byte units of key chunk 0 and 1
            float p[4], r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                p[e] = ex2(__uint_as_float(s0[4 * u + e]) - m_new);
                r[e] = ex2(__uint_as_float(s1[4 * u + e]) - m_new);
                rs += p[e] + r[e];
            }
            sts128(swz(sP, tid, u), p[0], p[1], p[2], p[3]);
            sts128(swz(sP + BM * 128, tid, u), r[0], r[1], r[2], r[3]);
        }
        l = fmaf(l, alpha, rs);
        m = m_new;
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] *= alpha;
        fence_async_smem();                                        // P (generic-proxy stores) -> visible to the tensor core
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(b_v, ph);
#pragma unroll
            for (int kk = 0; kk < BN / 8; ++kk)                    // 8 keys per MMA: P chunk kk/4, K step kk%4; V: two 4-row atoms
                tc_mma_tf32(tm_o, make_desc(sP + (kk >> 2) * (BM * 128) + (kk & 3) * 32, 16, 1024, 2),
                            make_desc(sV + kk * 1024, BN * 128, 512, 1), idesc(D, true), kk != 0 ? 1u : 0u);
            tc_commit(b_o);
        }
        mbar_wait(b_o, ph);
        tc_fence_after();
        if (tid == 0 && j + 1 < nblk) load_v(j + 1);               // PV is complete: the V tile (and P) are free
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t pv[32];
            tc_ld32(tm_o + lane_base + 32 * c, pv);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[32 * c + i] += __uint_as_float(pv[i]);
        }
        tc_fence_before();                                         // the next QK / PV overwrite S / PV after the next barrier
    }

    const int q = q0 + tid;
    if (q < a.L) {
        const float inv = 1.f / l;
        float* dst = a.o + (long long)(tok0 + q) * a.ld_o + col0;
#pragma unroll
        for (int d = 0; d < D; d += 4)
            *reinterpret_cast<float4*>(dst + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
        a.lse[(long long)bh * a.L + q] = m + log2f(l);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
    }
}


// =================================================================================================================
// backward
// =================================================================================================================
// delta[b, h, q] = sum_d dO[q, h, d] * O[q, h, d]: one warp per (token, head)
__global__ void __launch_bounds__(256) k_fmha_delta(const float* __restrict__ o, const float* __restrict__ d_o, float* __restrict__ delta,
                                                    int B, int H, int L, int D, long long ld_o) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)B * L * H) return;
    const int h = (int)(w % H);
    const long long tok = w / H;
    const float* po = o + tok * ld_o + h * D;
    const float* pg = d_o + tok * ld_o + h * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(po[d], pg[d], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        const int b = (int)(tok / L), q = (int)(tok - (long long)b * L);
        delta[((long long)b * H + h) * L + q] = acc;
    }
}

template <int D, int BNB>
struct Bwd {
    static constexpr int DC = D / 32;
    static constexpr int BIG = DC * BM * 128;          // a [128 x D] K-major tile
    static constexpr int SMALL = DC * BNB * 128;       // a [BNB x D] tile (K-major or MN-major)
    static constexpr int PT = (BNB / 32) * BM * 128;   // a [128 x BNB] K-major tile written by the threads
    static constexpr int SMEM_DQ = 2 * BIG + 3 * SMALL + PT + 1024 + 128;
    static constexpr int SMEM_DKV = 2 * BIG + 4 * SMALL + 2 * PT + 2 * BNB * 4 + 1024 + 128;
    static constexpr int TMEM_DQ = 256;                // S, dP (BNB each), dQ (D)
    static constexpr int TMEM_DKV = D == 64 ? 256 : 512;   // S^T, dP^T (BNB each), dV, dK (D each)
};

// ---- dQ: one CTA per 128-query block, loop over BNB-key blocks ----------------------------------------------------
template <int D, int BNB>
__global__ void __launch_bounds__(128, 1)
k_fmha_bwd_dq(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
              const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
              const __grid_constant__ CUtensorMap map_kmn, const LsFmha a, const float* __restrict__ delta, float* __restrict__ dq) {
    using C = Bwd<D, BNB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem), sdO = sQ + C::BIG, sK = sdO + C::BIG, sV = sK + C::SMALL, sKmn = sV + C::SMALL,
                   sdS = sKmn + C::SMALL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * C::BIG + 3 * C::SMALL + C::PT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const uint32_t b_q = smem_u32(bars), b_kv = b_q + 8, b_kmn = b_q + 16, b_s = b_q + 24, b_o = b_q + 32;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
    const int q0 = blockIdx.x * BM, tok0 = b * a.L, col0 = h * D;
    const int nblk = (a.L + BNB - 1) / BNB;

    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(b_q + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_DQ) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tm_s = tmem, tm_dp = tmem + BNB, tm_dq = tmem + 2 * BNB;

    auto load_kv = [&](int j) {
        mbar_expect_tx(b_kv, 2 * C::SMALL);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) {
            tma_load_2d(sK + c * (BNB * 128), &map_k, b_kv, col0 + 32 * c, tok0 + j * BNB);
            tma_load_2d(sV + c * (BNB * 128), &map_v, b_kv, col0 + 32 * c, tok0 + j * BNB);
        }
    };
    auto load_kmn = [&](int j) {
        mbar_expect_tx(b_kmn, C::SMALL);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) tma_load_2d(sKmn + c * (BNB * 128), &map_kmn, b_kmn, col0 + 32 * c, tok0 + j * BNB);
    };
    if (tid == 0) {
        mbar_expect_tx(b_q, 2 * C::BIG);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) {
            tma_load_2d(sQ + c * (BM * 128), &map_q, b_q, col0 + 32 * c, tok0 + q0);
            tma_load_2d(sdO + c * (BM * 128), &map_do, b_q, col0 + 32 * c, tok0 + q0);
        }
        load_kv(0);
        load_kmn(0);
    }
    const int q = q0 + tid;
    const bool q_ok = q < a.L;
    const float lse2 = q_ok ? a.lse[(long long)bh * a.L + q] : 0.f;
    const float dlt = q_ok ? delta[(long long)bh * a.L + q] : 0.f;
    const float sl2 = a.scale * kLog2e;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    for (int j = 0; j < nblk; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        if (tid == 0) {
            if (j == 0) mbar_wait(b_q, 0);
            mbar_wait(b_kv, ph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < C::DC; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t kd = make_desc(sK + c * (BNB * 128) + k * 32, 16, 1024, 2), vd = make_desc(sV + c * (BNB * 128) + k * 32, 16, 1024, 2);
                    tc_mma_tf32(tm_s, make_desc(sQ + c * (BM * 128) + k * 32, 16, 1024, 2), kd, idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                    tc_mma_tf32(tm_dp, make_desc(sdO + c * (BM * 128) + k * 32, 16, 1024, 2), vd, idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                }
            tc_commit(b_s);
        }
        mbar_wait(b_s, ph);
        tc_fence_after();
        if (tid == 0 && j + 1 < nblk) load_kv(j + 1);
        const int nvalid = a.L - j * BNB;
#pragma unroll
        for (int c = 0; c < BNB / 32; ++c) {
            uint32_t s[32], dp[32];
            tc_ld32_nowait(tm_s + lane_base + 32 * c, s);
            tc_ld32_nowait(tm_dp + lane_base + 32 * c, dp);
            tc_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * u + e;
                    const float p = (32 * c + i < nvalid) ? ex2(fmaf(__uint_as_float(s[i]), sl2, -lse2)) : 0.f;
                    ds[e] = p * (__uint_as_float(dp[i]) - dlt) * a.scale;
                }
                sts128(swz(sdS + c * (BM * 128), tid, u), ds[0], ds[1], ds[2], ds[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(b_kmn, ph);
#pragma unroll
            for (int kk = 0; kk < BNB / 8; ++kk)
                tc_mma_tf32(tm_dq, make_desc(sdS + (kk >> 2) * (BM * 128) + (kk & 3) * 32, 16, 1024, 2),
                            make_desc(sKmn + kk * 1024, BNB * 128, 512, 1), idesc(D, true), (j | kk) != 0 ? 1u : 0u);
            tc_commit(b_o);
        }
        mbar_wait(b_o, ph);                                        // dS and the MN-major K tile are free again
        tc_fence_after();
        if (tid == 0 && j + 1 < nblk) load_kmn(j + 1);
        tc_fence_before();
    }
    {
        float* dst = dq + (long long)(tok0 + q) * a.ld_q + col0;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t r[32];
            tc_ld32(tm_dq + lane_base + 32 * c, r);                // warp-collective: every lane loads, valid rows store
            if (q_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dst + 32 * c + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                               __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_DQ) : "memory");
    }
}

// ---- dK, dV: one CTA per 128-key block, loop over BNB-query blocks (transposed formulation) -----------------------
template <int D, int BNB>
__global__ void __launch_bounds__(128, 1)
k_fmha_bwd_dkv(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
               const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
               const __grid_constant__ CUtensorMap map_qmn, const __grid_constant__ CUtensorMap map_domn, const LsFmha a,
               const float* __restrict__ delta, float* __restrict__ dk, float* __restrict__ dv) {
    using C = Bwd<D, BNB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sK = smem_u32(smem), sV = sK + C::BIG, sQ = sV + C::BIG, sdO = sQ + C::SMALL, sQmn = sdO + C::SMALL,
                   sdOmn = sQmn + C::SMALL, sPT = sdOmn + C::SMALL, sdST = sPT + C::PT;
    float* s_lse = reinterpret_cast<float*>(smem + 2 * C::BIG + 4 * C::SMALL + 2 * C::PT);
    float* s_dlt = s_lse + BNB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dlt + BNB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    const uint32_t b_kv = smem_u32(bars), b_q = b_kv + 8, b_mn = b_kv + 16, b_s = b_kv + 24, b_o = b_kv + 32;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
    const int k0 = blockIdx.x * BM, tok0 = b * a.L, col0 = h * D;
    const int nblk = (a.L + BNB - 1) / BNB;

    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(b_kv + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C::TMEM_DKV) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tm_s = tmem, tm_dp = tmem + BNB, tm_dv = tmem + 2 * BNB, tm_dk = tmem + 2 * BNB + D;

    auto load_q = [&](int i) {
        mbar_expect_tx(b_q, 2 * C::SMALL);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) {
            tma_load_2d(sQ + c * (BNB * 128), &map_q, b_q, col0 + 32 * c, tok0 + i * BNB);
            tma_load_2d(sdO + c * (BNB * 128), &map_do, b_q, col0 + 32 * c, tok0 + i * BNB);
        }
    };
    auto load_mn = [&](int i) {
        mbar_expect_tx(b_mn, 2 * C::SMALL);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) {
            tma_load_2d(sQmn + c * (BNB * 128), &map_qmn, b_mn, col0 + 32 * c, tok0 + i * BNB);
            tma_load_2d(sdOmn + c * (BNB * 128), &map_domn, b_mn, col0 + 32 * c, tok0 + i * BNB);
        }
    };
    if (tid == 0) {
        mbar_expect_tx(b_kv, 2 * C::BIG);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) {
            tma_load_2d(sK + c * (BM * 128), &map_k, b_kv, col0 + 32 * c, tok0 + k0);
            tma_load_2d(sV + c * (BM * 128), &map_v, b_kv, col0 + 32 * c, tok0 + k0);
        }
        load_q(0);
        load_mn(0);
    }
    const float sl2 = a.scale * kLog2e;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    for (int i = 0; i < nblk; ++i) {
        const uint32_t ph = (uint32_t)i & 1u;
        if (tid < BNB) {                                           // softmax statistics of this block's queries
            const int q = i * BNB + tid;
            s_lse[tid] = q < a.L ? a.lse[(long long)bh * a.L + q] : INFINITY;      // +inf -> p = 0 for queries outside the sequence
            s_dlt[tid] = q < a.L ? delta[(long long)bh * a.L + q] : 0.f;
        }
        if (tid == 0) {
            if (i == 0) mbar_wait(b_kv, 0);
            mbar_wait(b_q, ph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < C::DC; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t qd = make_desc(sQ + c * (BNB * 128) + k * 32, 16, 1024, 2), od = make_desc(sdO + c * (BNB * 128) + k * 32, 16, 1024, 2);
                    tc_mma_tf32(tm_s, make_desc(sK + c * (BM * 128) + k * 32, 16, 1024, 2), qd, idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                    tc_mma_tf32(tm_dp, make_desc(sV + c * (BM * 128) + k * 32, 16, 1024, 2), od, idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                }
            tc_commit(b_s);
        }
        __syncthreads();                                           // s_lse / s_dlt visible
        mbar_wait(b_s, ph);
        tc_fence_after();
        if (tid == 0 && i + 1 < nblk) load_q(i + 1);
#pragma unroll
        for (int c = 0; c < BNB / 32; ++c) {
            uint32_t s[32], dp[32];
            tc_ld32_nowait(tm_s + lane_base + 32 * c, s);
            tc_ld32_nowait(tm_dp + lane_base + 32 * c, dp);
            tc_wait_ld();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float p[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int jq = 32 * c + 4 * u + e;
                    p[e] = ex2(fmaf(__uint_as_float(s[4 * u + e]), sl2, -s_lse[jq]));
                    ds[e] = p[e] * (__uint_as_float(dp[4 * u + e]) - s_dlt[jq]) * a.scale;
                }
                sts128(swz(sPT + c * (BM * 128), tid, u), p[0], p[1], p[2], p[3]);
                sts128(swz(sdST + c * (BM * 128), tid, u), ds[0], ds[1], ds[2], ds[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(b_mn, ph);
#pragma unroll
            for (int kk = 0; kk < BNB / 8; ++kk) {
                const uint32_t aoff = (kk >> 2) * (BM * 128) + (kk & 3) * 32;
                tc_mma_tf32(tm_dv, make_desc(sPT + aoff, 16, 1024, 2), make_desc(sdOmn + kk * 1024, BNB * 128, 512, 1), idesc(D, true),
                            (i | kk) != 0 ? 1u : 0u);
                tc_mma_tf32(tm_dk, make_desc(sdST + aoff, 16, 1024, 2), make_desc(sQmn + kk * 1024, BNB * 128, 512, 1), idesc(D, true),
                            (i | kk) != 0 ? 1u : 0u);
            }
            tc_commit(b_o);
        }
        mbar_wait(b_o, ph);
        tc_fence_after();
        if (tid == 0 && i + 1 < nblk) load_mn(i + 1);
        tc_fence_before();
        __syncthreads();                                           // s_lse / s_dlt are rewritten at the top of the next round
    }
    const int key = k0 + tid;
    const bool ok = key < a.L;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
        uint32_t rv[32], rk[32];
        tc_ld32_nowait(tm_dv + lane_base + 32 * c, rv);
        tc_ld32_nowait(tm_dk + lane_base + 32 * c, rk);
        tc_wait_ld();
        if (ok) {
            float* pv = dv + (long long)(tok0 + key) * a.ld_v + col0 + 32 * c;
            float* pk = dk + (long long)(tok0 + key) * a.ld_k + col0 + 32 * c;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                *reinterpret_cast<float4*>(pv + i) = make_float4(__uint_as_float(rv[i]), __uint_as_float(rv[i + 1]), __uint_as_float(rv[i + 2]), __uint_as_float(rv[i + 3]));
                *reinterpret_cast<float4*>(pk + i) = make_float4(__uint_as_float(rk[i]), __uint_as_float(rk[i + 1]), __uint_as_float(rk[i + 2]), __uint_as_float(rk[i + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_DKV) : "memory");
    }
}


// =================================================================================================================
// row softmax (wide single-head attention: scores are materialised by ls_gemm_tf32, e.g. the VAE mid block, D = 512)
// =================================================================================================================
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, off);
        v = is_max ? fmaxf(v, o) : v + o;
    }
    __syncthreads();                                               // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = is_max ? fmaxf(v, red[w]) : v + red[w];
    return v;
}

// x[r, :] <- softmax(scale * x[r, :]) in place; one 256-thread block per row, the row stays in registers (cols <= 4096)
__global__ void __launch_bounds__(256) k_softmax_rows_fwd(float* __restrict__ x, int cols, long long ld, float scale) {
    __shared__ float red[8];
    float* row = x + (long long)blockIdx.x * ld;
    const float sl2 = scale * kLog2e;
    float4 v[4];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * (threadIdx.x + 256 * i);
        if (c < cols) {
            v[i] = *reinterpret_cast<const float4*>(row + c);
            m = fmaxf(fmaxf(m, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
        }
    }
    m = block_reduce(m, red, true) * sl2;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * (threadIdx.x + 256 * i);
        if (c < cols) {
            v[i].x = ex2(fmaf(v[i].x, sl2, -m)); v[i].y = ex2(fmaf(v[i].y, sl2, -m));
            v[i].z = ex2(fmaf(v[i].z, sl2, -m)); v[i].w = ex2(fmaf(v[i].w, sl2, -m));
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float inv = 1.f / block_reduce(sum, red, false);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * (threadIdx.x + 256 * i);
        if (c < cols) *reinterpret_cast<float4*>(row + c) = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    }
}

// dp[r, :] <- scale * p[r, :] * (dp[r, :] - sum_j p[r, j] dp[r, j]) in place
__global__ void __launch_bounds__(256) k_softmax_rows_bwd(const float* __restrict__ p, float* __restrict__ dp, int cols, long long ld,
                                                          float scale) {
    __shared__ float red[8];
    const float* prow = p + (long long)blockIdx.x * ld;
    float* grow = dp + (long long)blockIdx.x * ld;
    float4 pv[4], gv[4];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * (threadIdx.x + 256 * i);
        if (c < cols) {
            pv[i] = *reinterpret_cast<const float4*>(prow + c);
            gv[i] = *reinterpret_cast<const float4*>(grow + c);
            dot += (pv[i].x * gv[i].x + pv[i].y * gv[i].y) + (pv[i].z * gv[i].z + pv[i].w * gv[i].w);
        }
    }
    dot = block_reduce(dot, red, false);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = 4 * (threadIdx.x + 256 * i);
        if (c < cols)
            *reinterpret_cast<float4*>(grow + c) = make_float4(scale * pv[i].x * (gv[i].x - dot), scale * pv[i].y * (gv[i].y - dot),
                                                               scale * pv[i].z * (gv[i].z - dot), scale * pv[i].w * (gv[i].w - dot));
    }
}

}  // namespace lsf

using namespace lsf;

namespace {
int make_map2d(CUtensorMap* map, const float* ptr, long long cols, long long rows, long long ld, int box_rows, CUtensorMapSwizzle sw) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ls_fail("cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ls_fail("fmha: cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld ld=%lld", (int)r, cols, rows, ld);
    return 0;
}

int check(const LsFmha* a) {
    if (!a) return ls_fail("fmha: args is NULL");
    bind_context();
    if (a->B <= 0 || a->H <= 0 || a->L <= 0) return ls_fail("fmha: bad sizes B=%d H=%d L=%d", a->B, a->H, a->L);
    if (a->D != 64 && a->D != 128) return ls_fail("fmha: head dim %d (64 and 128 are built)", a->D);
    if (!a->q || !a->k || !a->v || !a->o || !a->lse) return ls_fail("fmha: NULL pointer");
    if ((a->ld_q | a->ld_k | a->ld_v | a->ld_o) % 4) return ls_fail("fmha: row strides must be multiples of 4 floats");
    if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
         reinterpret_cast<uintptr_t>(a->o)) & 15)
        return ls_fail("fmha: q / k / v / o must be 16-byte aligned");
    if ((long long)a->B * a->H > 65535) return ls_fail("fmha: B*H exceeds the grid limit");
    return 0;
}

template <int D>
int launch_fwd(const LsFmha* a, cudaStream_t stream) {
    const long long cols = (long long)a->H * D, rows = (long long)a->B * a->L;
    CUtensorMap mq, mk, mv;
    if (make_map2d(&mq, a->q, cols, rows, a->ld_q, BM, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (make_map2d(&mk, a->k, cols, rows, a->ld_k, BN, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (make_map2d(&mv, a->v, cols, rows, a->ld_v, BN, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    static PerDeviceOnce once;
    if (once.ensure_smem(k_fmha_fwd<D>, Fwd<D>::SMEM) != cudaSuccess) return ls_check_cuda("fmha smem attribute");
    const dim3 grid((a->L + BM - 1) / BM, a->B * a->H);
    k_fmha_fwd<D><<<grid, 128, Fwd<D>::SMEM, stream>>>(mq, mk, mv, *a);
    return ls_check_cuda("k_fmha_fwd");
}
}  // namespace

extern "C" int ls_fmha_forward(const LsFmha* a, void* stream) {
    if (check(a)) return -1;
    return a->D == 64 ? launch_fwd<64>(a, (cudaStream_t)stream) : launch_fwd<128>(a, (cudaStream_t)stream);
}

namespace {
template <int D, int BNB>
int launch_bwd(const LsFmha* a, const float* d_o, float* dq, float* dk, float* dv, float* delta, cudaStream_t stream) {
    const long long cols = (long long)a->H * D, rows = (long long)a->B * a->L;
    const long long warps = rows * a->H;
    k_fmha_delta<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, stream>>>(a->o, d_o, delta, a->B, a->H, a->L, D, a->ld_o);
    if (ls_check_cuda("k_fmha_delta")) return -1;
    const CUtensorMapSwizzle KM = CU_TENSOR_MAP_SWIZZLE_128B, MN = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    CUtensorMap q_big, do_big, k_small, v_small, k_mn, k_big, v_big, q_small, do_small, q_mn, do_mn;
    if (make_map2d(&q_big, a->q, cols, rows, a->ld_q, BM, KM) || make_map2d(&do_big, d_o, cols, rows, a->ld_o, BM, KM) ||
        make_map2d(&k_small, a->k, cols, rows, a->ld_k, BNB, KM) || make_map2d(&v_small, a->v, cols, rows, a->ld_v, BNB, KM) ||
        make_map2d(&k_mn, a->k, cols, rows, a->ld_k, BNB, MN) || make_map2d(&k_big, a->k, cols, rows, a->ld_k, BM, KM) ||
        make_map2d(&v_big, a->v, cols, rows, a->ld_v, BM, KM) || make_map2d(&q_small, a->q, cols, rows, a->ld_q, BNB, KM) ||
        make_map2d(&do_small, d_o, cols, rows, a->ld_o, BNB, KM) || make_map2d(&q_mn, a->q, cols, rows, a->ld_q, BNB, MN) ||
        make_map2d(&do_mn, d_o, cols, rows, a->ld_o, BNB, MN))
        return -1;
    static PerDeviceOnce once_q, once_kv;
    if (once_q.ensure_smem(k_fmha_bwd_dq<D, BNB>, Bwd<D, BNB>::SMEM_DQ) != cudaSuccess) return ls_check_cuda("fmha dq smem attribute");
    if (once_kv.ensure_smem(k_fmha_bwd_dkv<D, BNB>, Bwd<D, BNB>::SMEM_DKV) != cudaSuccess) return ls_check_cuda("fmha dkv smem attribute");
    const dim3 grid((a->L + BM - 1) / BM, a->B * a->H);
    k_fmha_bwd_dq<D, BNB><<<grid, 128, Bwd<D, BNB>::SMEM_DQ, stream>>>(q_big, do_big, k_small, v_small, k_mn, *a, delta, dq);
    if (ls_check_cuda("k_fmha_bwd_dq")) return -1;
    k_fmha_bwd_dkv<D, BNB><<<grid, 128, Bwd<D, BNB>::SMEM_DKV, stream>>>(k_big, v_big, q_small, do_small, q_mn, do_mn, *a, delta, dk, dv);
    return ls_check_cuda("k_fmha_bwd_dkv");
}
}  // namespace

extern "C" int ls_fmha_backward(const LsFmha* a, const float* d_o, float* dq, float* dk, float* dv, float* delta, void* stream) {
    if (check(a)) return -1;
    if (!d_o || !dq || !dk || !dv || !delta) return ls_fail("fmha backward: NULL pointer");
    if ((reinterpret_cast<uintptr_t>(d_o) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15)
        return ls_fail("fmha backward: gradient pointers must be 16-byte aligned");
    return a->D == 64 ? launch_bwd<64, 64>(a, d_o, dq, dk, dv, delta, (cudaStream_t)stream)
                      : launch_bwd<128, 32>(a, d_o, dq, dk, dv, delta, (cudaStream_t)stream);
}

namespace {
int check_rows(const void* x, long long rows, int cols, long long ld) {
    if (!x) return ls_fail("softmax_rows: NULL pointer");
    if (rows <= 0 || rows > 2147483647LL) return ls_fail("softmax_rows: bad row count %lld", rows);
    if (cols <= 0 || cols > 4096 || cols % 4 || ld % 4 || ld < cols) return ls_fail("softmax_rows: cols=%d ld=%lld (cols <= 4096, both %% 4)", cols, ld);
    if (reinterpret_cast<uintptr_t>(x) & 15) return ls_fail("softmax_rows: pointer must be 16-byte aligned");
    return 0;
}
}  // namespace

extern "C" int ls_softmax_rows_forward(float* x, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream) {
    if (check_rows(x, rows, cols, ld)) return -1;
    k_softmax_rows_fwd<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, cols, ld, scale);
    return ls_check_cuda("k_softmax_rows_fwd");
}

extern "C" int ls_softmax_rows_backward(const float* p, float* dp, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream) {
    if (check_rows(p, rows, cols, ld) || check_rows(dp, rows, cols, ld)) return -1;
    k_softmax_rows_bwd<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(p, dp, cols, ld, scale);
    return ls_check_cuda("k_softmax_rows_bwd");
}
