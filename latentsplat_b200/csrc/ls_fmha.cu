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

// The probability / score-gradient tiles never leave the tensor core's own memory: the softmax warps read S (and dP) from TMEM
// with tcgen05.ld, write P (dS) back IN PLACE with tcgen05.st, and the next MMA takes that tile as its A operand straight from
// TMEM (tcgen05.mma with [a_tmem]): no shared-memory round trip, no proxy fence, and 32 KB less shared memory per CTA.
// MMAs execute in issue order, so the MMA that overwrites a score buffer is simply issued behind the one that consumed it.
// Shared-memory matrix descriptors split into their 32-bit halves: the high word is a compile-time constant per layout and the
// low word is (address >> 4) | (LBO >> 4) << 16, so stepping through a tile is ONE 32-bit add of a constant per MMA instead of a
// mask / shift / or chain (the single issuing thread is the bottleneck of these kernels: N = 64 MMAs last only 32 cycles).
constexpr uint32_t kHiK = (1024u >> 4) | (1u << 14) | (2u << 29);                 // K-major, SBO 1024, SWIZZLE_128B
constexpr uint32_t kHiMN = (512u >> 4) | (1u << 14) | (1u << 29);                 // MN-major, SBO 512, SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint32_t lo_k(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }                  // LBO 16
__device__ __forceinline__ uint32_t lo_mn(uint32_t saddr, uint32_t lbo) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint64_t dsc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <int D>
struct Fwd {
    static constexpr int DC = D / 32;
    static constexpr int Q_BYTES = DC * BM * 128;
    static constexpr int K_BYTES = DC * BN * 128;
    static constexpr int V_BYTES = DC * BN * 128;
    static constexpr int SMEM = Q_BYTES + K_BYTES + V_BYTES + 1024 + 64;
    static constexpr int TMEM_COLS = D == 64 ? 128 : 256;        // S / P: 64 columns, O: D columns
    static constexpr int CTAS = D == 64 ? 3 : 1;                 // 65 KB and 128 TMEM columns each
};

// ---- forward: one CTA per 128-query block, loop over 64-key blocks --------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128, Fwd<D>::CTAS)
k_fmha_fwd(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
           const __grid_constant__ CUtensorMap map_v, const LsFmha a) {
    using C = Fwd<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem), sK = sQ + C::Q_BYTES, sV = sK + C::K_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::Q_BYTES + C::K_BYTES + C::V_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t b_q = smem_u32(bars), b_k = b_q + 8, b_v = b_q + 16, b_s = b_q + 24;

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
        for (int i = 0; i < 4; ++i) mbar_init(b_q + 8 * i, 1);
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
    const uint32_t q_lo = lo_k(sQ), k_lo = lo_k(sK), v_lo = lo_mn(sV, BN * 128);
    auto issue_qk = [&]() {                                        // S = Q K^T (K = D in steps of 8)
#pragma unroll
        for (int c = 0; c < C::DC; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                tc_mma_tf32(tm_s, dsc(q_lo + ((c * (BM * 128) + k * 32) >> 4), kHiK), dsc(k_lo + ((c * (BN * 128) + k * 32) >> 4), kHiK),
                            idesc(BN, false), (c | k) != 0 ? 1u : 0u);
    };
    if (warp == 0 && elect_one()) {                                // one elected lane issues every TMA load and MMA
        mbar_expect_tx(b_q, C::Q_BYTES);
#pragma unroll
        for (int c = 0; c < C::DC; ++c) tma_load_2d(sQ + c * (BM * 128), &map_q, b_q, col0 + 32 * c, tok0 + q0);
        load_k(0);
        load_v(0);
        mbar_wait(b_q, 0);
        mbar_wait(b_k, 0);
        tc_fence_after();
        issue_qk();
        tc_commit(b_s);
    }

    const float sl2 = a.scale * kLog2e;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;        // this warp's TMEM lane quarter
    float m_ref = -INFINITY, l = 0.f;                              // reference maximum of the exponentials (lazy), row sum

    for (int j = 0; j < nblk; ++j) {
        mbar_wait(b_s, (uint32_t)j & 1u);                          // S(j) is complete -- and so is P V (j - 1)
        tc_fence_after();
        if (warp == 0 && elect_one()) {
            if (j + 1 < nblk) load_k(j + 1);                       // the K tile is free
            if (j > 0) load_v(j);                                  // the V tile is free
        }
        uint32_t s0[32], s1[32];
        tc_ld32_nowait(tm_s + lane_base, s0);
        tc_ld32_nowait(tm_s + lane_base + 32, s1);
        tc_wait_ld();
        const int nvalid = a.L - j * BN;                           // keys of this block inside the sequence (>= 1)
        float mx = -INFINITY;
        if (nvalid >= BN) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (i >= nvalid) s0[i] = 0xff800000u;              // -inf: exp2 -> 0
                if (i + 32 >= nvalid) s1[i] = 0xff800000u;
                mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
            }
        }
        mx *= sl2;                                                 // scale > 0
        // Lazy rescaling: the exponentials keep their reference maximum until a row's maximum outgrows it by 2^8; only then is
        // the running O (in TMEM, quiescent right now) multiplied down.  After the first blocks this is rare.
        const bool grow = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, grow)) {
            float alpha = 1.f;
            if (grow) {
                alpha = ex2(m_ref - mx);                           // 0 on the first block
                m_ref = mx;
                l *= alpha;
            }
            if (j > 0) {
#pragma unroll
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t o[32];
                    tc_ld32(tm_o + lane_base + 32 * c, o);
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tc_st32(tm_o + lane_base + 32 * c, o);
                }
            }
        }
        float rs = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float p0 = ex2(fmaf(__uint_as_float(s0[i]), sl2, -m_ref));
            const float p1 = ex2(fmaf(__uint_as_float(s1[i]), sl2, -m_ref));
            rs += p0 + p1;
            s0[i] = __float_as_uint(p0);
            s1[i] = __float_as_uint(p1);
        }
        l += rs;
        tc_st32(tm_s + lane_base, s0);                             // P over S, in place
        tc_st32(tm_s + lane_base + 32, s1);
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            mbar_wait(b_v, (uint32_t)j & 1u);
#pragma unroll
            for (int kk = 0; kk < BN / 8; ++kk)                    // O += P V, 8 keys per MMA: A = P columns 8 kk.. from TMEM
                tc_mma_tf32_ts(tm_o, tm_s + 8 * kk, dsc(v_lo + ((kk * 1024) >> 4), kHiMN), idesc(D, true), (j | kk) != 0 ? 1u : 0u);
            if (j + 1 < nblk) {
                mbar_wait(b_k, (uint32_t)(j + 1) & 1u);
                issue_qk();                                        // overwrites P only after the MMAs above have consumed it
            }
            tc_commit(b_s);
        }
    }
    mbar_wait(b_s, (uint32_t)nblk & 1u);
    tc_fence_after();
    const int q = q0 + tid;
    const float inv = 1.f / l;
    float* dst = a.o + (long long)(tok0 + q) * a.ld_o + col0;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
        uint32_t o[32];
        tc_ld32(tm_o + lane_base + 32 * c, o);                     // warp-collective: every lane loads, valid rows store
        if (q < a.L) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dst + 32 * c + i) = make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv,
                                                                           __uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
        }
    }
    if (q < a.L) a.lse[(long long)bh * a.L + q] = m_ref + log2f(l);
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
// delta[b, h, q] = sum_d dO[q, h, d] * O[q, h, d]: one thread per float4 of a token row (coalesced 16-byte loads), the D / 4
// lanes of a head reduce with shuffles
__global__ void __launch_bounds__(256) k_fmha_delta(const float* __restrict__ o, const float* __restrict__ d_o, float* __restrict__ delta,
                                                    int B, int H, int L, int D, long long ld_o) {
    const int per_head = D >> 2;                                   // 16 or 32 lanes
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x, per_row = (long long)H * per_head;
    const long long tok = t / per_row;
    const int rem = (int)(t - tok * per_row);
    float acc = 0.f;
    const bool ok = tok < (long long)B * L;
    if (ok) {
        const float4 x = *reinterpret_cast<const float4*>(o + tok * ld_o + 4 * rem);
        const float4 g = *reinterpret_cast<const float4*>(d_o + tok * ld_o + 4 * rem);
        acc = (x.x * g.x + x.y * g.y) + (x.z * g.z + x.w * g.w);
    }
    for (int off = per_head >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (ok && (rem & (per_head - 1)) == 0) {
        const int b = (int)(tok / L), q = (int)(tok - (long long)b * L), h = rem / per_head;
        delta[((long long)b * H + h) * L + q] = acc;
    }
}

// Backward kernels: warp-specialised.  Warps 0..7 do the per-element math, two per TMEM lane quarter (warp w: rows
// 32 (w & 3) .. +31, column half w >> 2); warp 8 issues every TMA load and every tcgen05.mma.  The score tiles S / dP are
// double-buffered in TMEM and the streamed operand tiles multi-buffered in shared memory: the tensor core computes the scores of
// block j + 1 while the math warps turn block j into dS (written over dP in TMEM), and the gradient MMA of block j -- A operand
// straight from TMEM -- queues behind.  The math warps only ever wait for "scores ready"; nothing waits for them but the issuer.
constexpr int kMathWarps = 8;
constexpr int kBwdThreads = 32 * (kMathWarps + 1);

template <int D, int BNB>
struct Bwd {
    static constexpr int DC = D / 32;
    static constexpr int BIG = DC * BM * 128;          // a [128 x D] K-major tile
    static constexpr int SMALL = DC * BNB * 128;       // a [BNB x D] tile (K-major or MN-major)
    static constexpr int CH = BNB / 2;                 // score columns per math thread
    // (S, dP) score buffers in TMEM: the scores of block j + NB are issued as soon as block j has been consumed, NB - 1 blocks ahead
    // of the math warps.  The streamed shared-memory tiles need one stage more than that (TMA latency is about one block of work).
    static constexpr int NB_DQ = 2, NB_DKV = 2;             // (3 buffers + 4 K/V stages measured no faster: 122 vs 116 us on the DINO shape)
    // dq: Q, dO resident; (K, V) K-major x KS stages; K MN-major x MS stages
    static constexpr int DQ_KS = D == 64 ? 3 : 2, DQ_MS = D == 64 ? 2 : 1;
    static constexpr int SMEM_DQ = 2 * BIG + DQ_KS * 2 * SMALL + DQ_MS * SMALL + 1024 + 256;
    static constexpr int TMEM_DQ = D == 64 ? 512 : 256;          // NB x (S, dP) of BNB columns + dQ (D)
    // dkv: K, V resident; (Q, dO) K-major x QS stages; (Q, dO) MN-major x MS stages; lse / delta of the block's queries
    static constexpr int DKV_QS = D == 64 ? 3 : 2, DKV_MS = D == 64 ? 2 : 1;
    static constexpr int SMEM_DKV = 2 * BIG + DKV_QS * 2 * SMALL + DKV_MS * 2 * SMALL + 4 * BNB * 4 + 1024 + 256;
    static constexpr int TMEM_DKV = 512;                         // NB x (S^T, dP^T) + dV + dK
};

__device__ __forceinline__ void math_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ---- dQ: one CTA per 128-query block, loop over BNB-key blocks ----------------------------------------------------
template <int D, int BNB>
__global__ void __launch_bounds__(kBwdThreads, 1)
k_fmha_bwd_dq(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
              const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
              const __grid_constant__ CUtensorMap map_kmn, const LsFmha a, const float* __restrict__ delta, float* __restrict__ dq) {
    using C = Bwd<D, BNB>;
    constexpr int KS = C::DQ_KS, MS = C::DQ_MS, CH = C::CH, NB = C::NB_DQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sQ = smem_u32(smem), sdO = sQ + C::BIG, sK = sdO + C::BIG, sV = sK + KS * C::SMALL, sKmn = sV + KS * C::SMALL;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * C::BIG + 2 * KS * C::SMALL + MS * C::SMALL);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t b_q = b0, b_kv = b0 + 8 /* [4] */, b_kmn = b0 + 40 /* [2] */, b_sfull = b0 + 56 /* [3] */, b_pready = b0 + 80 /* [3] */,
                   b_odone = b0 + 104, b_final = b0 + 112;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
    const int q0 = blockIdx.x * BM, tok0 = b * a.L, col0 = h * D;
    const int nblk = (a.L + BNB - 1) / BNB;

    if (tid == 0) {
        mbar_init(b_q, 1);
        for (int i = 0; i < 4; ++i) mbar_init(b_kv + 8 * i, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(b_sfull + 8 * i, 1); mbar_init(b_pready + 8 * i, kMathWarps); }
        for (int i = 0; i < 2; ++i) mbar_init(b_kmn + 8 * i, 1);
        mbar_init(b_odone, 1);
        mbar_init(b_final, 1);
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
    const uint32_t tm_dq = tmem + NB * 2 * BNB;                    // S[b] at tmem + 2 b BNB, dP[b] (later dS[b]) right behind it

    if (warp == kMathWarps) {
        // ================= issuer warp: TMA + MMA, one lane =================
        if (elect_one()) {
            const uint32_t q_lo = lo_k(sQ), do_lo = lo_k(sdO), k_lo = lo_k(sK), v_lo = lo_k(sV), kmn_lo = lo_mn(sKmn, BNB * 128);
            auto load_kv = [&](int j) {
                const uint32_t st = (uint32_t)(j % KS), bar = b_kv + 8 * st;
                mbar_expect_tx(bar, 2 * C::SMALL);
#pragma unroll
                for (int c = 0; c < C::DC; ++c) {
                    tma_load_2d(sK + st * C::SMALL + c * (BNB * 128), &map_k, bar, col0 + 32 * c, tok0 + j * BNB);
                    tma_load_2d(sV + st * C::SMALL + c * (BNB * 128), &map_v, bar, col0 + 32 * c, tok0 + j * BNB);
                }
            };
            auto load_kmn = [&](int j) {
                const uint32_t st = (uint32_t)(j % MS), bar = b_kmn + 8 * st;
                mbar_expect_tx(bar, C::SMALL);
#pragma unroll
                for (int c = 0; c < C::DC; ++c) tma_load_2d(sKmn + st * C::SMALL + c * (BNB * 128), &map_kmn, bar, col0 + 32 * c, tok0 + j * BNB);
            };
            auto issue_s = [&](int j) {                            // S(j) = Q K^T, dP(j) = dO V^T into TMEM buffer j % NB
                const uint32_t st = (uint32_t)(j % KS), tb = (uint32_t)(j % NB);
                mbar_wait(b_kv + 8 * st, (uint32_t)(j / KS) & 1u);
                tc_fence_after();
                const uint32_t tm_s = tmem + tb * 2 * BNB, tm_dp = tm_s + BNB;
                const uint32_t kl = k_lo + st * (C::SMALL >> 4), vl = v_lo + st * (C::SMALL >> 4);
#pragma unroll
                for (int c = 0; c < C::DC; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t oa = (c * (BM * 128) + k * 32) >> 4, ob = (c * (BNB * 128) + k * 32) >> 4;
                        tc_mma_tf32(tm_s, dsc(q_lo + oa, kHiK), dsc(kl + ob, kHiK), idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                        tc_mma_tf32(tm_dp, dsc(do_lo + oa, kHiK), dsc(vl + ob, kHiK), idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                    }
                tc_commit(b_sfull + 8 * tb);
            };
            mbar_expect_tx(b_q, 2 * C::BIG);
#pragma unroll
            for (int c = 0; c < C::DC; ++c) {
                tma_load_2d(sQ + c * (BM * 128), &map_q, b_q, col0 + 32 * c, tok0 + q0);
                tma_load_2d(sdO + c * (BM * 128), &map_do, b_q, col0 + 32 * c, tok0 + q0);
            }
            for (int j = 0; j < KS && j < nblk; ++j) load_kv(j);
            for (int j = 0; j < MS && j < nblk; ++j) load_kmn(j);
            mbar_wait(b_q, 0);
            for (int j = 0; j < NB && j < nblk; ++j) issue_s(j);
            for (int j = 0; j < nblk; ++j) {
                // S(j), dP(j) are complete (long ago): their K / V stage takes block j + KS
                if (j + KS < nblk) {
                    mbar_wait(b_sfull + 8 * (j % NB), (uint32_t)(j / NB) & 1u);
                    load_kv(j + KS);
                }
                if (j >= 1 && j - 1 + MS < nblk) {                 // dQ(j - 1) is complete: its MN-major K stage takes block j - 1 + MS
                    mbar_wait(b_odone, (uint32_t)(j - 1) & 1u);
                    load_kmn(j - 1 + MS);
                }
                mbar_wait(b_pready + 8 * (j % NB), (uint32_t)(j / NB) & 1u);   // dS(j) sits in TMEM over dP(j)
                mbar_wait(b_kmn + 8 * (j % MS), (uint32_t)(j / MS) & 1u);
                tc_fence_after();
                const uint32_t tm_ds = tmem + (uint32_t)(j % NB) * 2 * BNB + BNB, ml = kmn_lo + (uint32_t)(j % MS) * (C::SMALL >> 4);
#pragma unroll
                for (int kk = 0; kk < BNB / 8; ++kk)               // dQ += dS K, A = dS columns 8 kk.. from TMEM
                    tc_mma_tf32_ts(tm_dq, tm_ds + 8 * kk, dsc(ml + ((kk * 1024) >> 4), kHiMN), idesc(D, true), (j | kk) != 0 ? 1u : 0u);
                tc_commit(b_odone);
                if (j + NB < nblk) issue_s(j + NB);                // overwrites buffer j % NB behind the MMAs that just consumed it
            }
            tc_commit(b_final);                                    // its own barrier: the math warps never followed b_odone's phases
        }
    } else {
        // ================= math warps =================
        const int row = 32 * (warp & 3) + lane, half = warp >> 2;
        const int q = q0 + row;
        const bool q_ok = q < a.L;
        const float lse2 = q_ok ? a.lse[(long long)bh * a.L + q] : 0.f;
        const float dlt = q_ok ? delta[(long long)bh * a.L + q] : 0.f;
        const float sl2 = a.scale * kLog2e;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int j = 0; j < nblk; ++j) {
            const uint32_t tb = (uint32_t)(j % NB);
            mbar_wait(b_sfull + 8 * tb, (uint32_t)(j / NB) & 1u);
            tc_fence_after();
            uint32_t s[CH], dp[CH];
            const uint32_t t_dp = tmem + tb * 2 * BNB + BNB + lane_base + half * CH;
            tc_ld_nowait(tmem + tb * 2 * BNB + lane_base + half * CH, s);
            tc_ld_nowait(t_dp, dp);
            tc_wait_ld();
            const int nvalid = a.L - j * BNB - half * CH;          // columns of this thread inside the sequence
            if (nvalid >= CH) {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    dp[i] = __float_as_uint(ex2(fmaf(__uint_as_float(s[i]), sl2, -lse2)) * (__uint_as_float(dp[i]) - dlt));
            } else {
#pragma unroll
                for (int i = 0; i < CH; ++i)
                    dp[i] = i < nvalid ? __float_as_uint(ex2(fmaf(__uint_as_float(s[i]), sl2, -lse2)) * (__uint_as_float(dp[i]) - dlt)) : 0u;
            }
            tc_st(t_dp, dp);                                       // dS over dP, in place (the 1/sqrt(d) factor is applied to dQ at the end)
            tc_wait_st();
            tc_fence_before();
            warp_arrive(b_pready + 8 * tb, lane);
        }
        mbar_wait(b_final, 0);
        tc_fence_after();
        float* dst = dq + (long long)(tok0 + q) * a.ld_q + col0 + half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t r[32];
            tc_ld32(tm_dq + lane_base + half * (D / 2) + 32 * c, r);           // warp-collective: every lane loads, valid rows store
            if (q_ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(dst + 32 * c + i) =
                        make_float4(a.scale * __uint_as_float(r[i]), a.scale * __uint_as_float(r[i + 1]),
                                    a.scale * __uint_as_float(r[i + 2]), a.scale * __uint_as_float(r[i + 3]));
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
__global__ void __launch_bounds__(kBwdThreads, 1)
k_fmha_bwd_dkv(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
               const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
               const __grid_constant__ CUtensorMap map_qmn, const __grid_constant__ CUtensorMap map_domn, const LsFmha a,
               const float* __restrict__ delta, float* __restrict__ dk, float* __restrict__ dv) {
    using C = Bwd<D, BNB>;
    constexpr int QS = C::DKV_QS, MS = C::DKV_MS, CH = C::CH, NB = C::NB_DKV;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sK = smem_u32(smem), sV = sK + C::BIG, sQ = sV + C::BIG, sdO = sQ + QS * C::SMALL, sQmn = sdO + QS * C::SMALL,
                   sdOmn = sQmn + MS * C::SMALL;
    constexpr int kStatOff = 2 * C::BIG + 2 * QS * C::SMALL + 2 * MS * C::SMALL;
    float* s_stat = reinterpret_cast<float*>(smem + kStatOff);                  // [2][2][BNB]: lse, delta
    const uint32_t s_stat_u32 = sK + kStatOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 4 * BNB);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
    const uint32_t b0 = smem_u32(bars);
    const uint32_t b_kv = b0, b_q = b0 + 8 /* [3] */, b_mn = b0 + 32 /* [2] */, b_sfull = b0 + 48 /* [3] */, b_pready = b0 + 72 /* [3] */,
                   b_odone = b0 + 96, b_final = b0 + 104;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.y, b = bh / a.H, h = bh - b * a.H;
    const int k0 = blockIdx.x * BM, tok0 = b * a.L, col0 = h * D;
    const int nblk = (a.L + BNB - 1) / BNB;

    if (tid == 0) {
        mbar_init(b_kv, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(b_q + 8 * i, 1); mbar_init(b_sfull + 8 * i, 1); mbar_init(b_pready + 8 * i, kMathWarps); }
        for (int i = 0; i < 2; ++i) mbar_init(b_mn + 8 * i, 1);
        mbar_init(b_odone, 1);
        mbar_init(b_final, 1);
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
    const uint32_t tm_dv = tmem + NB * 2 * BNB, tm_dk = tm_dv + D;

    if (warp == kMathWarps) {
        if (elect_one()) {
            const uint32_t k_lo = lo_k(sK), v_lo = lo_k(sV), q_lo = lo_k(sQ), do_lo = lo_k(sdO), qmn_lo = lo_mn(sQmn, BNB * 128),
                           domn_lo = lo_mn(sdOmn, BNB * 128);
            auto load_q = [&](int i) {
                const uint32_t st = (uint32_t)(i % QS), bar = b_q + 8 * st;
                mbar_expect_tx(bar, 2 * C::SMALL);
#pragma unroll
                for (int c = 0; c < C::DC; ++c) {
                    tma_load_2d(sQ + st * C::SMALL + c * (BNB * 128), &map_q, bar, col0 + 32 * c, tok0 + i * BNB);
                    tma_load_2d(sdO + st * C::SMALL + c * (BNB * 128), &map_do, bar, col0 + 32 * c, tok0 + i * BNB);
                }
            };
            auto load_mn = [&](int i) {
                const uint32_t st = (uint32_t)(i % MS), bar = b_mn + 8 * st;
                mbar_expect_tx(bar, 2 * C::SMALL);
#pragma unroll
                for (int c = 0; c < C::DC; ++c) {
                    tma_load_2d(sQmn + st * C::SMALL + c * (BNB * 128), &map_qmn, bar, col0 + 32 * c, tok0 + i * BNB);
                    tma_load_2d(sdOmn + st * C::SMALL + c * (BNB * 128), &map_domn, bar, col0 + 32 * c, tok0 + i * BNB);
                }
            };
            auto issue_s = [&](int i) {                            // S^T(i) = K Q^T, dP^T(i) = V dO^T into TMEM buffer i % NB
                const uint32_t st = (uint32_t)(i % QS), tb = (uint32_t)(i % NB);
                mbar_wait(b_q + 8 * st, (uint32_t)(i / QS) & 1u);
                tc_fence_after();
                const uint32_t tm_s = tmem + tb * 2 * BNB, tm_dp = tm_s + BNB;
                const uint32_t ql = q_lo + st * (C::SMALL >> 4), ol = do_lo + st * (C::SMALL >> 4);
#pragma unroll
                for (int c = 0; c < C::DC; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t oa = (c * (BM * 128) + k * 32) >> 4, ob = (c * (BNB * 128) + k * 32) >> 4;
                        tc_mma_tf32(tm_s, dsc(k_lo + oa, kHiK), dsc(ql + ob, kHiK), idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                        tc_mma_tf32(tm_dp, dsc(v_lo + oa, kHiK), dsc(ol + ob, kHiK), idesc(BNB, false), (c | k) != 0 ? 1u : 0u);
                    }
                tc_commit(b_sfull + 8 * tb);
            };
            mbar_expect_tx(b_kv, 2 * C::BIG);
#pragma unroll
            for (int c = 0; c < C::DC; ++c) {
                tma_load_2d(sK + c * (BM * 128), &map_k, b_kv, col0 + 32 * c, tok0 + k0);
                tma_load_2d(sV + c * (BM * 128), &map_v, b_kv, col0 + 32 * c, tok0 + k0);
            }
            for (int i = 0; i < QS && i < nblk; ++i) load_q(i);
            for (int i = 0; i < MS && i < nblk; ++i) load_mn(i);
            mbar_wait(b_kv, 0);
            for (int i = 0; i < NB && i < nblk; ++i) issue_s(i);
            for (int i = 0; i < nblk; ++i) {
                if (i + QS < nblk) {                               // S^T(i), dP^T(i) complete: their Q / dO stage takes block i + QS
                    mbar_wait(b_sfull + 8 * (i % NB), (uint32_t)(i / NB) & 1u);
                    load_q(i + QS);
                }
                if (i >= 1 && i - 1 + MS < nblk) {                 // dV / dK (i - 1) complete: their MN-major stage takes block i - 1 + MS
                    mbar_wait(b_odone, (uint32_t)(i - 1) & 1u);
                    load_mn(i - 1 + MS);
                }
                mbar_wait(b_pready + 8 * (i % NB), (uint32_t)(i / NB) & 1u);   // P^T(i), dS^T(i) sit in TMEM over S^T(i), dP^T(i)
                mbar_wait(b_mn + 8 * (i % MS), (uint32_t)(i / MS) & 1u);
                tc_fence_after();
                const uint32_t tm_p = tmem + (uint32_t)(i % NB) * 2 * BNB, tm_ds = tm_p + BNB, mo = (uint32_t)(i % MS) * (C::SMALL >> 4);
#pragma unroll
                for (int kk = 0; kk < BNB / 8; ++kk) {
                    tc_mma_tf32_ts(tm_dv, tm_p + 8 * kk, dsc(domn_lo + mo + ((kk * 1024) >> 4), kHiMN), idesc(D, true), (i | kk) != 0 ? 1u : 0u);
                    tc_mma_tf32_ts(tm_dk, tm_ds + 8 * kk, dsc(qmn_lo + mo + ((kk * 1024) >> 4), kHiMN), idesc(D, true), (i | kk) != 0 ? 1u : 0u);
                }
                tc_commit(b_odone);
                if (i + NB < nblk) issue_s(i + NB);
            }
            tc_commit(b_final);
        }
    } else {
        const int half = warp >> 2;
        const float sl2 = a.scale * kLog2e;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int i = 0; i < nblk; ++i) {
            if (tid < 2 * BNB) {                                   // softmax statistics of this block's queries
                const int q = i * BNB + (tid < BNB ? tid : tid - BNB);
                // lse = +inf makes p = 0 for queries outside the sequence
                s_stat[(i & 1) * 2 * BNB + tid] =
                    tid < BNB ? (q < a.L ? a.lse[(long long)bh * a.L + q] : INFINITY) : (q < a.L ? delta[(long long)bh * a.L + q] : 0.f);
            }
            math_bar_sync();
            const uint32_t tb = (uint32_t)(i % NB);
            mbar_wait(b_sfull + 8 * tb, (uint32_t)(i / NB) & 1u);
            tc_fence_after();
            uint32_t s[CH], dp[CH];
            const uint32_t t_s = tmem + tb * 2 * BNB + lane_base + half * CH, t_dp = t_s + BNB;
            tc_ld_nowait(t_s, s);
            tc_ld_nowait(t_dp, dp);
            tc_wait_ld();
            const uint32_t st_u32 = s_stat_u32 + (uint32_t)((i & 1) * 2 * BNB + half * CH) * 4u;
#pragma unroll
            for (int u = 0; u < CH / 4; ++u) {
                const float4 l4 = lds128(st_u32 + 16 * u), d4 = lds128(st_u32 + BNB * 4 + 16 * u);
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p = ex2(fmaf(__uint_as_float(s[4 * u + e]), sl2, -ls[e]));
                    s[4 * u + e] = __float_as_uint(p);
                    dp[4 * u + e] = __float_as_uint(p * (__uint_as_float(dp[4 * u + e]) - dl[e]));
                }
            }
            tc_st(t_s, s);                                         // P^T over S^T, dS^T over dP^T (1/sqrt(d) goes on dK at the end)
            tc_st(t_dp, dp);
            tc_wait_st();
            tc_fence_before();
            warp_arrive(b_pready + 8 * tb, lane);
        }
        mbar_wait(b_final, 0);
        tc_fence_after();
        const int key = k0 + 32 * (warp & 3) + lane;
        const bool ok = key < a.L;
        float* pv = dv + (long long)(tok0 + key) * a.ld_v + col0 + half * (D / 2);
        float* pk = dk + (long long)(tok0 + key) * a.ld_k + col0 + half * (D / 2);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            uint32_t rv[32], rk[32];
            tc_ld32_nowait(tm_dv + lane_base + half * (D / 2) + 32 * c, rv);
            tc_ld32_nowait(tm_dk + lane_base + half * (D / 2) + 32 * c, rk);
            tc_wait_ld();
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    *reinterpret_cast<float4*>(pv + 32 * c + i) = make_float4(__uint_as_float(rv[i]), __uint_as_float(rv[i + 1]),
                                                                              __uint_as_float(rv[i + 2]), __uint_as_float(rv[i + 3]));
                    *reinterpret_cast<float4*>(pk + 32 * c + i) =
                        make_float4(a.scale * __uint_as_float(rk[i]), a.scale * __uint_as_float(rk[i + 1]),
                                    a.scale * __uint_as_float(rk[i + 2]), a.scale * __uint_as_float(rk[i + 3]));
                }
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
    const long long lanes = rows * a->H * (D / 4);
    k_fmha_delta<<<(unsigned)((lanes + 255) / 256), 256, 0, stream>>>(a->o, d_o, delta, a->B, a->H, a->L, D, a->ld_o);
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
    k_fmha_bwd_dq<D, BNB><<<grid, kBwdThreads, Bwd<D, BNB>::SMEM_DQ, stream>>>(q_big, do_big, k_small, v_small, k_mn, *a, delta, dq);
    if (ls_check_cuda("k_fmha_bwd_dq")) return -1;
    k_fmha_bwd_dkv<D, BNB><<<grid, kBwdThreads, Bwd<D, BNB>::SMEM_DKV, stream>>>(k_big, v_big, q_small, do_small, q_mn, do_mn, *a, delta, dk, dv);
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
