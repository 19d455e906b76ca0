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

extern "C" int ls_fmha_backward(const LsFmha* a, const float* d_o, float* dq, float* dk, float* dv, float* delta, void* stream) {
    (void)a; (void)d_o; (void)dq; (void)dk; (void)dv; (void)delta; (void)stream;
    return ls_fail("fmha backward: not built yet");
}
