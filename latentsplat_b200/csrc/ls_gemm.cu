// ls_gemm.cu -- TF32 tensor-core GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
//   C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias) -> activation            (see include/ls_gemm.h)
//
// Persistent kernel, one CTA per SM, each walking a list of 128 x 128 output tiles (optionally K-slices of them):
//   warp 0, one lane : TMA producer -- 128B-swizzled boxes of fp32 into a 6-stage ring (32 KB per stage), running
//                                      ahead across tile boundaries
//   warp 1, one lane : MMA issuer   -- 4 x tcgen05.mma.cta_group::1.kind::tf32 (M128 N128 K8) per stage,
//                                      tcgen05.commit releases the stage / publishes the accumulator
//   warp 2           : TMEM allocator (2 x 128 fp32 columns x 128 lanes: double-buffered accumulator)
//   warps 4..7       : epilogue      -- tcgen05.ld 32x32b.x32, smem transpose, bias + ReLU/GELU, coalesced 16-byte
//                                      stores or red.add; overlaps the next tile's main loop
// fp32 operands go from HBM to the tensor core untouched: kind::tf32 reads 32-bit elements from shared memory,
// so there is no cast pass and no bf16 copy of activations or weights.  Operands may be K-major (row = M/N index,
// K contiguous) or MN-major (row = K index, M/N contiguous), which covers forward (X W^T), dgrad (dY W) and
// wgrad (dY^T X) without transposed copies:
//   K-major  tile: one TMA box {32 K-floats, 128 rows}; canonical layout ((8,n),2):((8,SBO=1024B),1)
//   MN-major tile: four TMA boxes {32 MN-floats, 32 K-rows} in the 128B-swizzle-with-32B-atom mode, the only
//                  MN-major layout tcgen05 accepts for 32-bit operands (CUTLASS sm100_common.inl: "for mn-major
//                  tf32 operands, SW128_32B is the only available smem layout"): canonical
//                  ((8,n),(4,k)):((1,LBO=4096B),(4,SBO=512B)), descriptor layout type SWIZZLE_128B_BASE32B
// (layout algebra: CUTLASS cute/atom/mma_traits_sm100.hpp "make_umma_desc"; descriptor bit fields:
//  cute/arch/mma_sm100_desc.hpp).  Every mbarrier wait is bounded and traps instead of hanging.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ls_gemm.h"
#include "ls_host.h"
#include "ls_tc.cuh"

namespace lsg {

constexpr int BM = 128, BK = 32;              // BK fp32 = 128 B = one swizzle row; BN (128 | 256) is a template parameter
constexpr int UMMA_K = 8;                    // 32 B of tf32 per instruction
constexpr int A_BYTES = BM * BK * 4;
__host__ __device__ constexpr int stage_bytes(int bn) { return A_BYTES + bn * BK * 4; }
constexpr int kThreads = 384;                               // warps 0-3: producer, MMA issuer, TMEM allocator, (idle); 4-11: epilogue
constexpr int kEpiWarps = 8;
constexpr int PATCH_BYTES = kEpiWarps * 32 * 32 * 4;         // one XOR-swizzled 32 x 32 transpose patch (4 KB) per epilogue warp
constexpr int smem_bytes(int stages, int bn) { return stages * stage_bytes(bn) + PATCH_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/; }

using namespace lstc;   // mbarrier / TMA / tcgen05 wrappers, make_desc, sts128 / lds128 (ls_tc.cuh)

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == LS_ACT_RELU) return fmaxf(v, 0.f);
    if (act == LS_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    if (act == LS_ACT_SILU) return v / (1.f + __expf(-v));
    if (act == LS_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    return v;
}

// One epilogue warp's share of one output tile.  The accumulator arrives "one row per lane" (tcgen05.ld 32x32b); writing it
// out like that would touch 32 different rows per store instruction, so the warp transposes each 32 x 32 chunk through a
// private 4 KB shared-memory patch (16-byte units XOR-swizzled with the row: conflict-free both ways) and stores 4 rows x 128
// contiguous bytes per instruction.  EIGHT epilogue warps share a tile: warp w reads TMEM lanes (w % 4) * 32.. and takes the
// column chunks of parity (w - 4) / 4 -- with four warps the erf of a fused GELU (+54 % on the DINO fc1 shape) and above all
// the skip-connection read (3x: a dependent DRAM round trip per row and chunk) serialised behind the tensor core.  The
// residual rows of a chunk are therefore requested BEFORE the accumulator is fetched and transposed.
template <int BN>
__device__ __forceinline__ void epilogue_chunks(const uint32_t tmem_acc, const uint32_t patch_s, const int lane, const int half,
                                                const int row_base, const int col_base, const int M, const int N,
                                                float* __restrict__ C, const long long ldc, const float* __restrict__ bias,
                                                const int act, const int atomic, const float* __restrict__ residual,
                                                const long long ldr, float* __restrict__ pre_out, const bool vec_ok,
                                                const bool res_vec) {
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    const bool fast = !atomic && vec_ok;
    const bool pre_res = fast && residual != nullptr && res_vec;
#pragma unroll 1
    for (int c0 = half * 32; c0 < BN; c0 += 64) {
        const int col0 = col_base + c0;
        if (col0 >= N) break;                                   // warp-uniform
        const int col = col0 + sub_c;
        float4 res4[8];
        if (pre_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = row_base + 4 * i + sub_r;
                res4[i] = (row < M && col + 3 < N) ? __ldg(reinterpret_cast<const float4*>(residual + (long long)row * ldr + col))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        uint32_t r[32];
        tc_ld32(tmem_acc + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            sts128(patch_s + lane * 128 + ((((j >> 2) ^ (lane & 7))) << 4), __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias != nullptr) {
#pragma unroll
            for (int t = 0; t < 4; ++t) if (col + t < N) b4[t] = bias[col + t];
        }
        float4 rows4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + sub_r;
            rows4[i] = lds128(patch_s + rr * 128 + ((((lane & 7) ^ (rr & 7))) << 4));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = row_base + 4 * i + sub_r;
            const float4 v4 = rows4[i];
            const float pre[4] = {v4.x + b4[0], v4.y + b4[1], v4.z + b4[2], v4.w + b4[3]};
            float v[4] = {apply_act(pre[0], act), apply_act(pre[1], act), apply_act(pre[2], act), apply_act(pre[3], act)};
            if (row < M) {
                float* dst = C + (long long)row * ldc + col;
                if (fast && col + 3 < N) {
                    if (pre_res) {
                        v[0] += res4[i].x; v[1] += res4[i].y; v[2] += res4[i].z; v[3] += res4[i].w;
                    } else if (residual != nullptr) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) v[t] += residual[(long long)row * ldr + col + t];
                    }
                    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    if (pre_out != nullptr)
                        *reinterpret_cast<float4*>(pre_out + (long long)row * ldc + col) = make_float4(pre[0], pre[1], pre[2], pre[3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (col + t < N) {
                            if (atomic) {
                                atomicAdd(dst + t, v[t]);
                            } else {
                                dst[t] = residual != nullptr ? v[t] + residual[(long long)row * ldr + col + t] : v[t];
                                if (pre_out != nullptr) pre_out[(long long)row * ldc + col + t] = pre[t];
                            }
                        }
                }
            }
        }
        __syncwarp();
    }
}

// Persistent: each CTA walks the work list  item = blockIdx.x + i * gridDim.x  (item -> n-tile fastest, then m-tile,
// then K-split, so that concurrently running CTAs share A row panels in L2).  The TMA ring runs ahead across tiles,
// the accumulator is double buffered in TMEM (2 x 128 columns) so that the epilogue of tile i overlaps the main loop
// of tile i+1.
template <bool A_MN, bool B_MN, int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
            const float* __restrict__ bias, int M, int N, int K, long long ldc, int act, int kb_per_split, int atomic,
            int tiles_m, int tiles_n, int splits, const float* __restrict__ residual, long long ldr, float* __restrict__ pre_out) {
    constexpr int STAGE_BYTES = stage_bytes(BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* patches = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);                    // 8 x 4 KB swizzled patches
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + PATCH_BYTES);    // full[S] empty[S] tfull[2] tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb_total = (K + BK - 1) / BK;
    const int n_items = tiles_m * tiles_n * splits;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    auto decode = [&](int item, int& tile_m, int& tile_n, int& kb0, int& nkb) {
        tile_n = item % tiles_n;
        const int rest = item / tiles_n;
        tile_m = rest % tiles_m;
        kb0 = (rest / tiles_m) * kb_per_split;
        nkb = min(nkb_total, kb0 + kb_per_split) - kb0;      // >= 1 by construction of `splits`
    };

    if (warp == 0 && elect_one()) {
        // ------------------------------ TMA producer ------------------------------
        uint32_t it = 0;                                       // global k-block counter -> stage / phase
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                const int k0 = (kb0 + kb) * BK;
                const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
                if (!A_MN) {
                    tma_load_2d(a_dst, &map_a, full0 + 8 * s, k0, tile_m * BM);
                } else {
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) tma_load_2d(a_dst + j * 4096, &map_a, full0 + 8 * s, tile_m * BM + 32 * j, k0);
                }
                if (!B_MN) {
                    tma_load_2d(b_dst, &map_b, full0 + 8 * s, k0, tile_n * BN);
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_dst + j * 4096, &map_b, full0 + 8 * s, tile_n * BN + 32 * j, k0);
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        // ------------------------------ MMA issuer --------------------------------
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        uint32_t it = 0, local = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_src = smem_base + s * STAGE_BYTES, b_src = a_src + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // K-major: step 32 B inside the 128-B swizzle row.  MN-major: step 8 K-rows = two 4-row (512 B) atoms.
                    const uint64_t adesc = A_MN ? make_desc(a_src + k * 1024, 4096, 512, 1) : make_desc(a_src + k * 32, 16, 1024, 2);
                    const uint64_t bdesc = B_MN ? make_desc(b_src + k * 1024, 4096, 512, 1) : make_desc(b_src + k * 32, 16, 1024, 2);
                    tc_mma_tf32(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(empty0 + 8 * s);          // frees the stage once these MMAs have read it
            }
            tc_commit(tfull0 + 8 * acc);            // accumulator complete
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue (8 warps) ------------------------
        const int q = warp & 3, half = (warp - 4) >> 2;          // TMEM lane quarter == warp % 4; even / odd column chunks
        const uint32_t patch_s = smem_u32(patches) + (uint32_t)(warp - 4) * 4096u;
        const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                            (pre_out == nullptr || (reinterpret_cast<uintptr_t>(pre_out) & 15) == 0);
        const bool res_vec = (ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0);
        uint32_t local = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tc_fence_after();
            epilogue_chunks<BN>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN), patch_s, lane, half,
                                tile_m * BM + q * 32, tile_n * BN, M, N, C, ldc, (bias != nullptr && kb0 == 0) ? bias : nullptr, act,
                                atomic, residual, ldr, pre_out, vec_ok, res_vec);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);           // 8 arrivals release the accumulator
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// =================================================================================================================
// cta_group::2 variant (default for M > 128; LS_GEMM_2CTA=0 selects the 1-CTA kernel) -- a cluster of two CTAs (one TPC) owns a 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows); the leader's tcgen05.mma.cta_group::2 reads
// A and B from both CTAs' shared memory, so per MMA (135 tensor cycles at M256 N256 K8) a CTA streams 4 KB of A +
// 4 KB of B instead of 4 + 8 KB -- the shared-memory bandwidth that caps the 1-CTA kernel at 57-60 % tensor pipe.
// Protocol (CUTLASS sm100 PipelineTmaUmmaAsync / Allocator2Sm, restated):
//   full[s]   lives in the leader; armed by the leader's producer with the bytes of BOTH CTAs; each CTA's TMA
//             (cp.async.bulk.tensor...cta_group::2) signals it through the peer-bit-cleared barrier address
//   empty[s]  one per CTA, released by the leader's tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11)
//   tfull[a]  one per CTA, same multicast commit -> each CTA's epilogue drains its own 128 TMEM lanes
//   tempty[a] lives in the leader, 8 arrivals (4 epilogue warps x 2 CTAs; the peer arrives through shared::cluster)
// Status: passes tests/test_gemm_gpu.py (all four operand layouts, tails, split-K, bias / activation / residual epilogues, Linear
// forward/backward) and the whole GPU suite; 4-13 % faster than the 1-CTA kernel on the DINO shapes (scripts/gemm_ab.py:
// 8200x3072x768 506 -> 525 TF/s, 8200x768x3072 429 -> 484 TF/s); whole-step A/B on one box (round 2): GEMM family 13.58 -> 12.78 ms,
// step 76.35 -> 75.12 ms -- hence the default.
// =================================================================================================================
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;        // shared::cluster address of the even (leader) CTA's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

template <bool A_MN, bool B_MN, int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_gemm_tf32_2cta(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
                 const float* __restrict__ bias, int M, int N, int K, long long ldc, int act, int kb_per_split, int atomic,
                 int tiles_m, int tiles_n, int splits, const float* __restrict__ residual, long long ldr, float* __restrict__ pre_out) {
    constexpr int BH = BN / 2;                                   // rows of B staged by each CTA
    constexpr int STAGE_BYTES = A_BYTES + BH * BK * 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* patches = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + PATCH_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nkb_total = (K + BK - 1) / BK;
    const int n_items = tiles_m * tiles_n * splits;               // tiles_m counts 256-row tiles here

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                          // barriers of both CTAs initialised, TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    auto decode = [&](int item, int& tile_m, int& tile_n, int& kb0, int& nkb) {
        tile_n = item % tiles_n;
        const int rest = item / tiles_n;
        tile_m = rest % tiles_m;
        kb0 = (rest / tiles_m) * kb_per_split;
        nkb = min(nkb_total, kb0 + kb_per_split) - kb0;
    };

    if (warp == 0 && elect_one()) {
        // ------------------------------ TMA producer (both CTAs) ------------------------------
        uint32_t it = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            const int m0 = tile_m * 2 * BM + (int)rank * BM, n0 = tile_n * BN + (int)rank * BH;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty0 + 8 * s, ph ^ 1);                                   // own stage is free
                if (rank == 0) mbar_expect_tx(full0 + 8 * s, 2 * STAGE_BYTES);       // bytes of both CTAs
                const uint32_t leader_full = (full0 + 8 * s) & kPeerBitMask;
                const int k0 = (kb0 + kb) * BK;
                const uint32_t a_dst = smem_base + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
                if (!A_MN) {
                    tma_load_2d_2sm(a_dst, &map_a, leader_full, k0, m0);
                } else {
#pragma unroll
                    for (int j = 0; j < BM / 32; ++j) tma_load_2d_2sm(a_dst + j * 4096, &map_a, leader_full, m0 + 32 * j, k0);
                }
                if (!B_MN) {
                    tma_load_2d_2sm(b_dst, &map_b, leader_full, k0, n0);
                } else {
#pragma unroll
                    for (int j = 0; j < BH / 32; ++j) tma_load_2d_2sm(b_dst + j * 4096, &map_b, leader_full, n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 1 && rank == 0 && elect_one()) {
        // ------------------------------ MMA issuer (leader CTA only) --------------------------
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
        uint32_t it = 0, local = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++local) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1);            // both CTAs' epilogues have drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_src = smem_base + s * STAGE_BYTES, b_src = a_src + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adesc = A_MN ? make_desc(a_src + k * 1024, 4096, 512, 1) : make_desc(a_src + k * 32, 16, 1024, 2);
                    const uint64_t bdesc = B_MN ? make_desc(b_src + k * 1024, 4096, 512, 1) : make_desc(b_src + k * 32, 16, 1024, 2);
                    tc_mma_tf32_2sm(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit_2sm(empty0 + 8 * s);       // releases the stage in BOTH CTAs
            }
            tc_commit_2sm(tfull0 + 8 * acc);         // accumulator complete, both CTAs
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue (both CTAs, own 128 rows, 8 warps each) ------
        const int q = warp & 3, half = (warp - 4) >> 2;
        const uint32_t patch_s = smem_u32(patches) + (uint32_t)(warp - 4) * 4096u;
        const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                            (pre_out == nullptr || (reinterpret_cast<uintptr_t>(pre_out) & 15) == 0);
        const bool res_vec = (ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(residual) & 15) == 0);
        uint32_t local = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++local) {
            int tile_m, tile_n, kb0, nkb;
            decode(item, tile_m, tile_n, kb0, nkb);
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tc_fence_after();
            epilogue_chunks<BN>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN), patch_s, lane, half,
                                tile_m * 2 * BM + (int)rank * BM + q * 32, tile_n * BN, M, N, C, ldc,
                                (bias != nullptr && kb0 == 0) ? bias : nullptr, act, atomic, residual, ldr, pre_out, vec_ok, res_vec);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty0 + 8 * acc);    // 16 arrivals (2 CTAs x 8 warps) release it
        }
    }
    tc_fence_before();
    cluster_sync_all();                                              // nobody frees TMEM / exits while the peer still uses it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

}  // namespace lsg

using namespace lsg;

namespace {
// 2-D fp32 tensor [outer][inner] with row pitch ld (elements), 128B-swizzled boxes {32, box_rows}, zero fill out of bounds
int make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_rows,
             CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ls_fail("cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ls_fail("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                          (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
    return 0;
}

template <bool A_MN, bool B_MN, int BN, int STAGES>
int launch_s(const CUtensorMap& ma, const CUtensorMap& mb, const LsGemmArgs* a, dim3 grid, int kb_per_split, int atomic,
             cudaStream_t stream) {
    static PerDeviceOnce once;                    // per instantiation, per device (ls_tc.cuh)
    constexpr int smem = smem_bytes(STAGES, BN);
    if (once.ensure_smem(k_gemm_tf32<A_MN, B_MN, BN, STAGES>, smem) != cudaSuccess) return ls_check_cuda("gemm smem attribute");
    const int n_items = (int)(grid.x * grid.y * grid.z);
    const int num_sms = current_sm_count();
    const int ctas = n_items < num_sms ? n_items : num_sms;
    k_gemm_tf32<A_MN, B_MN, BN, STAGES><<<ctas, kThreads, smem, stream>>>(ma, mb, a->C, a->bias, a->M, a->N, a->K,
                                                                       (long long)a->ldc, a->act, kb_per_split, atomic,
                                                                       (int)grid.x, (int)grid.y, (int)grid.z, a->residual,
                                                                       (long long)a->ldr, a->pre_out);
    return ls_check_cuda("k_gemm_tf32");
}

// 2-CTA launch: grid.x counts 256-row tiles, one cluster of two CTAs per work item slot.
constexpr int smem_bytes_2cta(int stages, int bn) { return stages * (A_BYTES + (bn / 2) * BK * 4) + PATCH_BYTES + 1024 + 256; }

template <bool A_MN, bool B_MN, int BN, int STAGES>
int launch_2cta_s(const CUtensorMap& ma, const CUtensorMap& mb, const LsGemmArgs* a, dim3 grid, int kb_per_split, int atomic,
                  cudaStream_t stream) {
    static PerDeviceOnce once;
    constexpr int smem = smem_bytes_2cta(STAGES, BN);
    if (once.ensure_smem(k_gemm_tf32_2cta<A_MN, B_MN, BN, STAGES>, smem) != cudaSuccess) return ls_check_cuda("gemm 2cta smem attribute");
    const int n_items = (int)(grid.x * grid.y * grid.z);
    const int num_sms = current_sm_count();
    const int clusters = n_items < num_sms / 2 ? n_items : num_sms / 2;
    k_gemm_tf32_2cta<A_MN, B_MN, BN, STAGES><<<2 * clusters, kThreads, smem, stream>>>(
        ma, mb, a->C, a->bias, a->M, a->N, a->K, (long long)a->ldc, a->act, kb_per_split, atomic, (int)grid.x, (int)grid.y,
        (int)grid.z, a->residual, (long long)a->ldr, a->pre_out);
    return ls_check_cuda("k_gemm_tf32_2cta");
}

template <bool A_MN, bool B_MN>
int launch_2cta(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const LsGemmArgs* a, dim3 grid, int kb_per_split, int atomic,
                cudaStream_t stream) {
    // per CTA and stage: 16 KB of A + BN/2 rows of B (16 KB at BN=256, 8 KB at BN=128); 6 stages + patches <= 211 KB
    return bn == 256 ? launch_2cta_s<A_MN, B_MN, 256, 6>(ma, mb, a, grid, kb_per_split, atomic, stream)
                     : launch_2cta_s<A_MN, B_MN, 128, 6>(ma, mb, a, grid, kb_per_split, atomic, stream);
}

bool use_2cta() {
    static int flag = -1;
    if (flag < 0) {
        const char* e = getenv("LS_GEMM_2CTA");
        flag = (e && atoi(e) == 0) ? 0 : 1;                 // on by default
    }
    return flag == 1;
}

template <bool A_MN, bool B_MN>
int launch(int bn, const CUtensorMap& ma, const CUtensorMap& mb, const LsGemmArgs* a, dim3 grid, int kb_per_split, int atomic,
           cudaStream_t stream) {
    // 128 x 128 tiles: 6 x 32 KB ring; 128 x 256 tiles: 4 x 48 KB ring (+ 18 KB transpose patches) ~ 211 KB
    return bn == 256 ? launch_s<A_MN, B_MN, 256, 4>(ma, mb, a, grid, kb_per_split, atomic, stream)
                     : launch_s<A_MN, B_MN, 128, 6>(ma, mb, a, grid, kb_per_split, atomic, stream);
}

// fp32 operands make the 128 x 128 tile shared-memory bound (32 KB written by TMA and 32 KB read by the tensor core
// per 256 MMA cycles = 256 B/cycle against 128 B/cycle of smem bandwidth, ncu: tensor pipe 40 %); a 128 x 256 tile
// moves 48 KB per 512 cycles.  Narrow outputs keep the square tile.
int pick_bn(int N) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("LS_GEMM_BN");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 128 || forced == 256) return forced;
    return N > 128 ? 256 : 128;
}
}  // namespace

extern "C" int ls_gemm_tf32(const LsGemmArgs* a, void* stream_) {
    if (!a) return ls_fail("gemm args is NULL");
    bind_context();
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) return ls_fail("gemm: bad sizes M=%d N=%d K=%d", a->M, a->N, a->K);
    if (!a->A || !a->B || !a->C) return ls_fail("gemm: NULL operand");
    if ((a->lda % 4) || (a->ldb % 4)) return ls_fail("gemm: lda/ldb must be multiples of 4 elements (TMA 16-byte stride rule)");
    if ((reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) return ls_fail("gemm: A/B must be 16-byte aligned");
    int split = a->split_k < 1 ? 1 : a->split_k;
    const int nkb = (a->K + BK - 1) / BK;
    if (split > nkb) split = nkb;
    const int kb_per_split = (nkb + split - 1) / split;
    split = (nkb + kb_per_split - 1) / kb_per_split;
    const int atomic = (split > 1 || a->accumulate) ? 1 : 0;
    if (atomic && a->act != LS_ACT_NONE) return ls_fail("gemm: activation cannot be fused with split-K / accumulate");
    if (atomic && (a->residual || a->pre_out)) return ls_fail("gemm: residual / pre_out cannot be combined with split-K / accumulate");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (split > 1 && !a->accumulate) {
        if (cudaMemset2DAsync(a->C, a->ldc * sizeof(float), 0, (size_t)a->N * sizeof(float), (size_t)a->M, stream) != cudaSuccess)
            return ls_check_cuda("gemm memset");
    }
    const int BN = pick_bn(a->N);
    CUtensorMap ma, mb;
    const CUtensorMapSwizzle sw_k = CU_TENSOR_MAP_SWIZZLE_128B, sw_mn = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (use_2cta() && a->M > BM) {                     // cta_group::2 path: 256-row tiles, B box = BN/2 rows
        if (a->a_mn_major ? make_map(&ma, a->A, a->M, a->K, a->lda, 32, sw_mn) : make_map(&ma, a->A, a->K, a->M, a->lda, BM, sw_k)) return -1;
        if (a->b_mn_major ? make_map(&mb, a->B, a->N, a->K, a->ldb, 32, sw_mn) : make_map(&mb, a->B, a->K, a->N, a->ldb, (uint32_t)(BN / 2), sw_k)) return -1;
        dim3 grid2((a->M + 2 * BM - 1) / (2 * BM), (a->N + BN - 1) / BN, split);
        if (a->a_mn_major) return a->b_mn_major ? launch_2cta<true, true>(BN, ma, mb, a, grid2, kb_per_split, atomic, stream)
                                                : launch_2cta<true, false>(BN, ma, mb, a, grid2, kb_per_split, atomic, stream);
        return a->b_mn_major ? launch_2cta<false, true>(BN, ma, mb, a, grid2, kb_per_split, atomic, stream)
                             : launch_2cta<false, false>(BN, ma, mb, a, grid2, kb_per_split, atomic, stream);
    }
    if (a->a_mn_major ? make_map(&ma, a->A, a->M, a->K, a->lda, 32, sw_mn) : make_map(&ma, a->A, a->K, a->M, a->lda, BM, sw_k)) return -1;
    if (a->b_mn_major ? make_map(&mb, a->B, a->N, a->K, a->ldb, 32, sw_mn) : make_map(&mb, a->B, a->K, a->N, a->ldb, (uint32_t)BN, sw_k)) return -1;
    dim3 grid((a->M + BM - 1) / BM, (a->N + BN - 1) / BN, split);
    if (a->a_mn_major) return a->b_mn_major ? launch<true, true>(BN, ma, mb, a, grid, kb_per_split, atomic, stream)
                                            : launch<true, false>(BN, ma, mb, a, grid, kb_per_split, atomic, stream);
    return a->b_mn_major ? launch<false, true>(BN, ma, mb, a, grid, kb_per_split, atomic, stream)
                         : launch<false, false>(BN, ma, mb, a, grid, kb_per_split, atomic, stream);
}
