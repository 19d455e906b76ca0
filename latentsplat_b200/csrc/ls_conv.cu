// ls_conv.cu -- implicit-GEMM convolutions for sm_100a on the tcgen05 pipeline of ls_gemm.cu (include/ls_conv.h).
//
// Activations are NHWC, so for one filter tap (r, s) the im2col matrix of a convolution is simply the activation tensor
// shifted by (r - pad, s - pad): a 4-D TMA box {32 channels, BW pixels, BH rows, 1 image} of it IS a K-major GEMM operand
// tile, with TMA's out-of-bounds zero fill playing the padding and its element strides playing the convolution stride.
// Nothing is unfolded or transposed in memory.  Two kernels share one warp-specialised pipeline (TMA producer lane / MMA
// issuer lane / TMEM allocator warp / epilogue warps, persistent over a work list); both exist for cta_group::1 and for
// cta_group::2 (a cluster of two CTAs on one TPC computes a 256-row tile; each CTA stages its own 128 rows of A and HALF of
// B, which cuts the L2 -> shared-memory traffic per flop by a third -- with fp32 operands that traffic, ~61 B/clk/SM measured,
// is what bounds the 1-CTA kernels at ~70 % of the tensor pipe):
//
//   k_conv_t  D[channel, pixel] = sum_taps sum_k W[channel, tap, k] * X[pixel + tap, k]          (forward, dgrad, transposed conv)
//             A = 128 (x2) output channels of the weight matrix (K-major box, or MN-major boxes for dgrad / transposed-conv
//             forward: dX = dY * W without a transposed weight copy), B = 256 pixels of the activation tensor in ONE 4-D TMA
//             box per K block.  The accumulator arrives with one output channel per TMEM lane, so a warp's tcgen05.ld hands
//             lane l the values of channel c0+l for 32 consecutive pixels and every store instruction writes 32 consecutive
//             channels of one NHWC pixel (128 contiguous bytes): no shared-memory transpose; bias + ReLU / GELU / SiLU /
//             LeakyReLU fused, optionally storing the pre-activation for the backward pass; an output stride/offset serves the
//             stride classes of a strided dgrad and the transposed convolution.
//   k_conv_w  dW[m, n] = sum_pixels P[pixel + tap(m), m] * Q[pixel + tap(n), n]                   (weight gradient)
//             both operands are activation tensors read MN-major in 64-pixel K blocks (SWIZZLE_128B with 32-B atoms, the only
//             MN-major layout tcgen05 takes for 32-bit operands); each 32-channel box of an operand carries its own filter tap,
//             so one tile may span several taps.  The pixel loop is cut into chunks and an item is (chunk, tile) with the tile
//             index fastest: the CTAs running at any moment work on the same few pixel chunks, so both activation tensors are
//             read from DRAM once (a tile-major split re-read them once per tile: 1.3 GB for 268 MB of operands in ncu);
//             items leave through red.global.add.v4.  64-pixel K blocks halve the number of 4-D boxes per byte (ncu: the 12
//             boxes of a 32-pixel block cost ~40 cycles each in the TMA unit on top of the data time).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ls_conv.h"
#include "ls_host.h"
#include "ls_tc.cuh"

namespace lsc {
using namespace lstc;

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int BKW = 64;                                     // K block (pixels) of the weight-gradient kernel
constexpr int A_BYTES = BM * BK * 4;
constexpr int PITCH = 36;                                   // floats; 144-B rows: conflict-free float4 access
constexpr int PATCH_BYTES = 4 * 32 * PITCH * 4;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;              // shared::cluster address of the even (leader) CTA's copy

// how the 32-channel boxes of an activation operand map to TMA coordinates
struct ActOp {
    int cp;            // k_conv_w: MN-index pitch of one filter tap (multiple of 32); one tap only: >= the MN extent
    int S;             // k_conv_w: taps per filter row
    int sx, sy;        // coordinate multipliers (convolution stride) applied to the iteration-grid position
    int dx0, dy0;      // offset of tap (0, 0)
    int dxs, dys;      // offset step per tap (+1 forward, -1 flipped for dgrad)
};

struct ConvParams {
    int P, Q, n_img;          // iteration grid (rows, columns) per image, images
    int bw_log2;              // T: one CTA's pixel box = (1 << bw_log2) x (NSUB >> bw_log2);  W: K block = (1 << bw_log2) x (64 >> bw_log2)
    int sub_dx, sub_dy;       // T, cta_group::2: pixel offset of the second CTA's box inside the pair's tile
    int tiles_x, tiles_y;     // T: pixel tiles per image;  W: K blocks per image
    int tiles_m, tiles_n, splits, kb_per_split, nkb_total;
    int R, S, chunks;         // T: taps iterated and 32-channel chunks per tap
    ActOp a, b;
    int wp, wr0, wrs, ws0, wss, wS;   // T: weight k/column offset of tap (r, s) = wp * ((wr0 + r*wrs) * wS + ws0 + s*wss)
    long long o_sn, o_sh, o_sw;       // T: output element strides (image, row, pixel)
    int oys, oyo, oxs, oxo;           // T: output pixel = (iy*oys + oyo, ix*oxs + oxo)
    int n_out, act, vec_ok;
    float* out;
    float* pre_out;
    const float* bias;
    int cp_r, c_r, t_r, cp_c, c_c, t_c;   // W: tile row/column -> (tap, channel): tap = idx / cp, channel = idx % cp (< c), tap < t
    long long rs, cs;                     // W: output address = (tap_r*c_r + ch_r)*rs + (tap_c*c_c + ch_c)*cs
    int atomic;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case LS_ACT_RELU: return fmaxf(v, 0.f);
        case LS_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        case LS_ACT_SILU: return v / (1.f + __expf(-v));
        case LS_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
        default: return v;
    }
}

// ---- cta_group-dependent PTX (CG = 1: one CTA; CG = 2: a CTA pair, protocol of ls_gemm.cu's k_gemm_tf32_2cta) ----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void cg_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if constexpr (CG == 1) {
        tma_load_2d(dst, map, bar, c0, c1);
    } else {        // signals the LEADER's barrier (peer bit cleared)
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    if constexpr (CG == 1) {
        tma_load_4d(dst, map, bar, c0, c1, c2, c3);
    } else {
        asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_commit(uint32_t bar) {
    if constexpr (CG == 1) {
        tc_commit(bar);
    } else {        // arrives on the barrier at this offset in BOTH CTAs
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_arrive_leader(uint32_t bar) {      // accumulator-drained signal to the MMA issuer
    if constexpr (CG == 1) {
        mbar_arrive(bar);
    } else {
        asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_tmem_alloc(uint32_t slot, uint32_t cols) {
    if constexpr (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void cg_tmem_dealloc(uint32_t base, uint32_t cols) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void cg_sync() {        // all threads of the CTA (pair)
    if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
}

// =================================================================================================================
// k_conv_t
// =================================================================================================================
constexpr int kThreadsT = 384;      // warps 0-3: TMA producer, MMA issuer, TMEM allocator, (idle); warps 4-11: epilogue
constexpr int smem_bytes_t(int stages, int bn, int cg) { return stages * (A_BYTES + (bn / cg) * BK * 4) + 1024 + 256; }

template <bool AMN, int BN, int STAGES, int CG>
__global__ void __launch_bounds__(kThreadsT, 1)
k_conv_t(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
         const __grid_constant__ ConvParams p) {
    constexpr int NSUB = BN / CG;                               // pixels staged by one CTA
    constexpr int B_BYTES = NSUB * BK * 4;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int TMEM_COLS = 2 * BN;                           // 256 / 512: double-buffered accumulator
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);      // full[S] empty[S] tfull[2] tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CG == 1 ? 0 : (int)cluster_ctarank();
    const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;     // a worker = one CTA or one CTA pair
    const int n_items = p.tiles_m * p.tiles_n;                  // tiles_m: channel tiles (128*CG), tiles_n: pixel tiles (BN)

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8 * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) cg_tmem_alloc<CG>(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    cg_sync<CG>();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    // item -> channel tile (fastest: workers running together share the pixel box in L2) and pixel tile -> image, first pixel
    auto decode = [&](int item, int& tile_m, int& n, int& y0, int& x0) {
        tile_m = item % p.tiles_m;
        const int tp = item / p.tiles_m;
        const int tx = tp % p.tiles_x;
        const int t = tp / p.tiles_x;
        n = t / p.tiles_y;
        // the pair's tile = the CTA box doubled along y (sub_dy > 0) or x (sub_dx > 0)
        y0 = (t - n * p.tiles_y) * ((NSUB >> p.bw_log2) + (CG == 2 ? p.sub_dy : 0));
        x0 = tx * ((1 << p.bw_log2) + (CG == 2 ? p.sub_dx : 0));
    };

    if (warp == 0 && elect_one()) {
        // ------------------------------ TMA producer (every CTA) ------------------
        uint32_t it = 0;
        for (int item = worker; item < n_items; item += n_workers) {
            int tile_m, n, y0, x0;
            decode(item, tile_m, n, y0, x0);
            const int m0 = tile_m * BM * CG + rank * BM;
            const int ys = y0 + rank * p.sub_dy, xs = x0 + rank * p.sub_dx;       // this CTA's pixel box
            for (int r = 0; r < p.R; ++r) {
                const int by = ys * p.a.sy + p.a.dy0 + r * p.a.dys;
                for (int s = 0; s < p.S; ++s) {
                    const int bx = xs * p.a.sx + p.a.dx0 + s * p.a.dxs;
                    const int wk = p.wp * ((p.wr0 + r * p.wrs) * p.wS + p.ws0 + s * p.wss);
                    for (int ch = 0; ch < p.chunks; ++ch, ++it) {
                        const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(empty0 + 8 * st, ph ^ 1);
                        if (rank == 0) mbar_expect_tx(full0 + 8 * st, CG * STAGE_BYTES);        // bytes of both CTAs
                        const uint32_t a_dst = smem_base + st * STAGE_BYTES, b_dst = a_dst + A_BYTES;
                        if (!AMN) {
                            cg_tma_2d<CG>(a_dst, &map_w, full0 + 8 * st, wk + ch * BK, m0);
                        } else {
#pragma unroll
                            for (int j = 0; j < BM / 32; ++j)
                                cg_tma_2d<CG>(a_dst + j * 4096, &map_w, full0 + 8 * st, wk + m0 + 32 * j, ch * BK);
                        }
                        cg_tma_4d<CG>(b_dst, &map_x, full0 + 8 * st, ch * BK, bx, by, n);
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0 && elect_one()) {
        // ------------------------------ MMA issuer (leader CTA) -------------------
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((AMN ? 1u : 0u) << 15) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
        uint32_t it = 0, local = 0;
        for (int item = worker; item < n_items; item += n_workers, ++local) {
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < p.nkb_total; ++kb, ++it) {
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_src = smem_base + st * STAGE_BYTES, b_src = a_src + A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t adesc = AMN ? make_desc(a_src + k * 1024, 4096, 512, 1) : make_desc(a_src + k * 32, 16, 1024, 2);
                    const uint64_t bdesc = make_desc(b_src + k * 32, 16, 1024, 2);
                    cg_mma<CG>(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                cg_commit<CG>(empty0 + 8 * st);
            }
            cg_commit<CG>(tfull0 + 8 * acc);
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue (8 warps per CTA) ----------------
        // warp w reads TMEM lanes (w % 4) * 32 .. +31 (its output channels); warps 4-7 take the even 32-pixel column chunks,
        // warps 8-11 the odd ones.  A chunk is 32 consecutive pixels of ONE grid row (the host guarantees box width >= 32), so
        // the output address is base + j * step with a warp-uniform base: per value one compare, one add, one store.
        const int q = warp & 3, half = (warp - 4) >> 2;
        constexpr int kChunks = BN / 64;                            // chunks per warp
        uint32_t local = 0;
        for (int item = worker; item < n_items; item += n_workers, ++local) {
            int tile_m, n, y0, x0;
            decode(item, tile_m, n, y0, x0);
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            const int c_base = tile_m * BM * CG + rank * BM + q * 32;
            const int co = c_base + lane;                           // this lane's output channel
            const bool live = c_base < p.n_out;                     // warp-uniform: the warp owns live channels
            const bool co_ok = co < p.n_out;
            const float bias = (p.bias != nullptr && co_ok) ? p.bias[co] : 0.f;
            const long long img = (long long)n * p.o_sn + co;
            const long long step = (long long)p.oxs * p.o_sw;
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tc_fence_after();
            if (live) {
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
                uint32_t ra[32], rb[32];
                tc_ld32_nowait(t0 + (uint32_t)(half * 32), ra);
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
                    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
                    tc_wait_ld();
                    if (c + 1 < kChunks) tc_ld32_nowait(t0 + (uint32_t)((2 * (c + 1) + half) * 32), nxt);   // overlaps the stores below
                    const int col = (2 * c + half) * 32;            // first pixel column of the chunk
                    const int sub = col / NSUB, cc = col - sub * NSUB;                   // which CTA's pixel box, position in it
                    const int iy = y0 + sub * p.sub_dy + (cc >> p.bw_log2);
                    const int ix0 = x0 + sub * p.sub_dx + (cc & ((1 << p.bw_log2) - 1));
                    int nvalid = (iy < p.P && co_ok) ? p.Q - ix0 : 0;                   // pixels of this chunk inside the grid
                    nvalid = nvalid > 32 ? 32 : nvalid;
                    const long long off = img + (long long)(iy * p.oys + p.oyo) * p.o_sh + (long long)(ix0 * p.oxs + p.oxo) * p.o_sw;
                    float* __restrict__ o = p.out + off;
                    if (p.act == LS_ACT_NONE && p.pre_out == nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) o[j * step] = __uint_as_float(cur[j]) + bias;
                    } else {
                        float* __restrict__ pr = p.pre_out ? p.pre_out + off : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                const float v = __uint_as_float(cur[j]) + bias;
                                o[j * step] = apply_act(v, p.act);
                                if (pr) pr[j * step] = v;
                            }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) cg_arrive_leader<CG>(tempty0 + 8 * acc);
        }
    }
    tc_fence_before();
    cg_sync<CG>();                                                   // nobody frees TMEM / exits while the peer still uses it
    if (warp == 2) {
        tc_fence_after();
        cg_tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
}

// =================================================================================================================
// k_conv_w
// =================================================================================================================
constexpr int kThreadsW = 256;      // warps 0-3: producer, MMA issuer, TMEM allocator, (idle); warps 4-7: epilogue
constexpr int smem_bytes_w(int stages, int bn, int cg) { return stages * (BM + bn / cg) * BKW * 4 + PATCH_BYTES + 1024 + 256; }

template <int BN, int STAGES, int CG>
__global__ void __launch_bounds__(kThreadsW, 1)
k_conv_w(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
         const __grid_constant__ ConvParams p) {
    constexpr int NSUB = BN / CG;                               // B columns staged by one CTA
    constexpr int AW_BYTES = BM * BKW * 4, BW_BYTES = NSUB * BKW * 4;
    constexpr int STAGE_BYTES = AW_BYTES + BW_BYTES;
    constexpr int CHUNK = BKW * 128;                            // one 32-channel x 64-pixel box: 8 KB
    constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* patches = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + PATCH_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CG == 1 ? 0 : (int)cluster_ctarank();
    const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;

    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull0 = smem_u32(bars + 2 * STAGES), tempty0 = smem_u32(bars + 2 * STAGES + 2);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4 * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) cg_tmem_alloc<CG>(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    cg_sync<CG>();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    // item = (pixel chunk, tile), tile fastest
    const long long w_tiles = (long long)p.tiles_m * p.tiles_n;
    const long long w_end = w_tiles * p.splits;
    auto next_work = [&](long long& cur, int& tile_m, int& tile_n, int& kb0, int& nkb) -> bool {
        if (cur >= w_end) return false;
        const int chunk = (int)(cur / w_tiles), tile = (int)(cur - (long long)chunk * w_tiles);
        kb0 = chunk * p.kb_per_split;
        nkb = min(p.nkb_total - kb0, p.kb_per_split);
        tile_n = tile % p.tiles_n;
        tile_m = tile / p.tiles_n;
        cur += n_workers;
        return true;
    };

    if (warp == 0 && elect_one()) {
        // ------------------------------ TMA producer (every CTA) ------------------
        uint32_t it = 0;
        long long cur = worker;
        int tile_m, tile_n, kb0, nkb;
        while (next_work(cur, tile_m, tile_n, kb0, nkb)) {
            // per 32-channel box: channel offset and tap shift (constant over the K loop)
            int ac[BM / 32], ax[BM / 32], ay[BM / 32], bc[NSUB / 32], bx[NSUB / 32], by[NSUB / 32];
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) {
                const int m = tile_m * BM * CG + rank * BM + 32 * j, tap = m / p.a.cp, r = tap / p.a.S, s = tap - r * p.a.S;
                ac[j] = m - tap * p.a.cp; ax[j] = p.a.dx0 + s * p.a.dxs; ay[j] = p.a.dy0 + r * p.a.dys;
            }
#pragma unroll
            for (int j = 0; j < NSUB / 32; ++j) {
                const int m = tile_n * BN + rank * NSUB + 32 * j, tap = m / p.b.cp, r = tap / p.b.S, s = tap - r * p.b.S;
                bc[j] = m - tap * p.b.cp; bx[j] = p.b.dx0 + s * p.b.dxs; by[j] = p.b.dy0 + r * p.b.dys;
            }
            const int kw_log2 = p.bw_log2, kh = BKW >> kw_log2;
            int xb = kb0 % p.tiles_x, t = kb0 / p.tiles_x;
            int n = t / p.tiles_y, yb = t - n * p.tiles_y;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty0 + 8 * st, ph ^ 1);
                if (rank == 0) mbar_expect_tx(full0 + 8 * st, CG * STAGE_BYTES);
                const uint32_t a_dst = smem_base + st * STAGE_BYTES, b_dst = a_dst + AW_BYTES;
                const int x0 = xb << kw_log2, y0 = yb * kh;
#pragma unroll
                for (int j = 0; j < BM / 32; ++j)
                    cg_tma_4d<CG>(a_dst + j * CHUNK, &map_a, full0 + 8 * st, ac[j], x0 * p.a.sx + ax[j], y0 * p.a.sy + ay[j], n);
#pragma unroll
                for (int j = 0; j < NSUB / 32; ++j)
                    cg_tma_4d<CG>(b_dst + j * CHUNK, &map_b, full0 + 8 * st, bc[j], x0 * p.b.sx + bx[j], y0 * p.b.sy + by[j], n);
                if (++xb == p.tiles_x) { xb = 0; if (++yb == p.tiles_y) { yb = 0; ++n; } }
            }
        }
    } else if (warp == 1 && rank == 0 && elect_one()) {
        // ------------------------------ MMA issuer (leader CTA) -------------------
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
        uint32_t it = 0, local = 0;
        long long cur = worker;
        int tile_m, tile_n, kb0, nkb;
        for (; next_work(cur, tile_m, tile_n, kb0, nkb); ++local) {
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t st = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * st, ph);
                tc_fence_after();
                const uint32_t a_src = smem_base + st * STAGE_BYTES, b_src = a_src + AW_BYTES;
#pragma unroll
                for (int k = 0; k < BKW / UMMA_K; ++k) {        // 8 K rows = two 4-row (512 B) atoms per MMA; 32-channel chunks CHUNK apart
                    const uint64_t adesc = make_desc(a_src + k * 1024, CHUNK, 512, 1);
                    const uint64_t bdesc = make_desc(b_src + k * 1024, CHUNK, 512, 1);
                    cg_mma<CG>(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                cg_commit<CG>(empty0 + 8 * st);
            }
            cg_commit<CG>(tfull0 + 8 * acc);
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue (4 warps per CTA) ----------------
        const int q = warp - 4;
        const uint32_t patch_s = smem_u32(patches + q * 32 * PITCH);
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
        uint32_t local = 0;
        long long cur = worker;
        int tile_m, tile_n, kb0, nkb;
        for (; next_work(cur, tile_m, tile_n, kb0, nkb); ++local) {
            const uint32_t acc = local & 1, acc_ph = (local >> 1) & 1;
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tc_fence_after();
            // rows of this warp: 8 per lane (4*i + sub_r), resolved to output offsets once per tile (one tap per 32 rows)
            long long row_off[8];
            bool row_ok[8];
            {
                const int row0 = tile_m * BM * CG + rank * BM + q * 32, tap = row0 / p.cp_r, ch0 = row0 - tap * p.cp_r;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int ch = ch0 + 4 * i + sub_r;
                    row_ok[i] = tap < p.t_r && ch < p.c_r;
                    row_off[i] = (long long)(tap * p.c_r + ch) * p.rs;
                }
            }
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                const int col0 = tile_n * BN + c0;
                const int tapc = col0 / p.cp_c, chc0 = col0 - tapc * p.cp_c;
                if (tapc >= p.t_c) break;                                // warp-uniform
                if (p.c_c - chc0 <= 0) continue;                         // padded channels of this tap
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts128(patch_s + (lane * PITCH + j) * 4, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                __syncwarp();
                float4 rows4[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) rows4[i] = lds128(patch_s + ((4 * i + sub_r) * PITCH + sub_c) * 4);
                const int ch = chc0 + sub_c;
                const long long col_off = (long long)(tapc * p.c_c + ch) * p.cs;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (!row_ok[i]) continue;
                    const float4 v4 = rows4[i];
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                    float* dst = p.out + row_off[i] + col_off;
                    if (p.vec_ok && ch + 3 < p.c_c) {
                        if (p.atomic) red_add_v4(dst, v[0], v[1], v[2], v[3]);
                        else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (ch + t < p.c_c) {
                                if (p.atomic) atomicAdd(dst + t * p.cs, v[t]); else dst[t * p.cs] = v[t];
                            }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) cg_arrive_leader<CG>(tempty0 + 8 * acc);
        }
    }
    tc_fence_before();
    cg_sync<CG>();
    if (warp == 2) {
        tc_fence_after();
        cg_tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
}

// elementwise backward of the fused epilogue activations: dx = dy * act'(pre)
__global__ void __launch_bounds__(256) k_act_bwd(const float4* __restrict__ dy, const float4* __restrict__ pre,
                                                 float4* __restrict__ dx, long long n4, int act) {
    auto d = [act](float g, float x) {
        switch (act) {
            case LS_ACT_RELU: return x > 0.f ? g : 0.f;
            case LS_ACT_GELU: {
                const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
                return g * (cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x));
            }
            case LS_ACT_SILU: {
                const float s = 1.f / (1.f + __expf(-x));
                return g * s * (1.f + x * (1.f - s));
            }
            case LS_ACT_LRELU: return x > 0.f ? g : 0.2f * g;
            default: return g;
        }
    };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g = dy[i], x = pre[i];
        dx[i] = make_float4(d(g.x, x.x), d(g.y, x.y), d(g.z, x.z), d(g.w, x.w));
    }
}

// ---- nearest-2x up-sampling folded into the following 3x3 convolution ------------------------------------------------
// y = conv3x3(upsample2x(x)): output pixel (2Y+py, 2X+px) reads up-sampled rows 2Y+py-1 .. 2Y+py+1, i.e. input rows
// {Y-1, Y, Y} (py = 0) or {Y, Y, Y+1} (py = 1): per output parity class the 3 x 3 filter collapses to 2 x 2 taps with summed
// weights (rows {0 | 1+2} resp. {0+1 | 2}, same for columns).  Four 2x2 convolutions on the LOW-resolution tensor replace one
// 3x3 convolution on the 4x larger one: 16 instead of 36 tap-MACs per input pixel (2.25x fewer flops) and the up-sampled
// tensor is never written or read.  The transposed relations give dgrad as ONE 4x4 stride-2 convolution over dy.
//   w4[cls = py*2+px][co][a][b][ci]   folded forward weights            (4, Cout, 2, 2, Cin)
//   wd[co][ty][tx][ci]                the same values as a 4x4 filter    (Cout, 4, 4, Cin), ty <-> (py, a): 0:(1,1) 1:(0,1) 2:(1,0) 3:(0,0)
__device__ __forceinline__ void up_rows(int py, int a, int& r0, int& r1) {      // filter rows summed into tap a of class py
    if (py == 0) { r0 = a == 0 ? 0 : 1; r1 = a == 0 ? 0 : 2; }
    else         { r0 = a == 0 ? 0 : 2; r1 = a == 0 ? 1 : 2; }
}
__global__ void __launch_bounds__(256) k_upconv_fold(const float* __restrict__ w, float* __restrict__ w4, float* __restrict__ wd,
                                                     int Cout, int Cin) {
    const long long total = 4ll * Cout * 4 * Cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        long long t = i / Cin;
        const int b = (int)(t & 1), a = (int)((t >> 1) & 1);
        t >>= 2;
        const int co = (int)(t % Cout), cls = (int)(t / Cout), py = cls >> 1, px = cls & 1;
        int r0, r1, s0, s1;
        up_rows(py, a, r0, r1);
        up_rows(px, b, s0, s1);
        float v = 0.f;
        for (int r = r0; r <= r1; ++r)
            for (int sc = s0; sc <= s1; ++sc) v += w[(((long long)co * 3 + r) * 3 + sc) * Cin + ci];
        w4[i] = v;
        const int ty = py == 1 ? (a == 1 ? 0 : 2) : (a == 1 ? 1 : 3), tx = px == 1 ? (b == 1 ? 0 : 2) : (b == 1 ? 1 : 3);
        wd[(((long long)co * 4 + ty) * 4 + tx) * Cin + ci] = v;
    }
}
// dw[co][r][s][ci] = sum of the folded-weight gradients every (class, tap) that contains filter tap (r, s)
__global__ void __launch_bounds__(256) k_upconv_unfold(const float* __restrict__ dw4, float* __restrict__ dw, int Cout, int Cin) {
    const long long total = (long long)Cout * 9 * Cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        long long t = i / Cin;
        const int sc = (int)(t % 3), r = (int)((t / 3) % 3), co = (int)(t / 9);
        float v = 0.f;
        for (int py = 0; py < 2; ++py)
            for (int a = 0; a < 2; ++a) {
                int r0, r1;
                up_rows(py, a, r0, r1);
                if (r < r0 || r > r1) continue;
                for (int px = 0; px < 2; ++px)
                    for (int b = 0; b < 2; ++b) {
                        int s0, s1;
                        up_rows(px, b, s0, s1);
                        if (sc < s0 || sc > s1) continue;
                        v += dw4[((((long long)(py * 2 + px) * Cout + co) * 2 + a) * 2 + b) * Cin + ci];
                    }
            }
        dw[i] = v;
    }
}

}  // namespace lsc

using namespace lsc;

namespace {

struct Act4 {                 // an NHWC activation tensor
    const float* ptr;
    int N, H, W, C;
};

int pow2_ceil_log2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

// 4-D map over an NHWC tensor: box {32 channels, bx pixels, by rows, 1 image} traversed with strides (sx, sy)
int make_act_map(CUtensorMap* map, const Act4& t, int bx, int by, int sx, int sy, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ls_fail("cuTensorMapEncodeTiled entry point not available");
    if (t.C % 4) return ls_fail("conv: activation channels (%d) must be a multiple of 4 (TMA 16-byte stride rule)", t.C);
    if (reinterpret_cast<uintptr_t>(t.ptr) & 15) return ls_fail("conv: activation pointer must be 16-byte aligned");
    if (bx * sx > 256 || by * sy > 256) return ls_fail("conv: TMA box %dx%d with strides %dx%d exceeds 256", bx, by, sx, sy);
    const cuuint64_t dims[4] = {(cuuint64_t)t.C, (cuuint64_t)t.W, (cuuint64_t)t.H, (cuuint64_t)t.N};
    const cuuint64_t strides[3] = {(cuuint64_t)t.C * 4, (cuuint64_t)t.W * t.C * 4, (cuuint64_t)t.H * t.W * t.C * 4};
    const cuuint32_t box[4] = {32u, (cuuint32_t)(bx * sx), (cuuint32_t)(by * sy), 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)sx, (cuuint32_t)sy, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(t.ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return ls_fail("cuTensorMapEncodeTiled(4d) failed (%d): N=%d H=%d W=%d C=%d box=%dx%d stride=%dx%d", (int)r, t.N, t.H, t.W, t.C,
                       bx, by, sx, sy);
    return 0;
}

// 2-D map over a weight matrix [rows][cols] (row pitch = cols)
int make_w_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, int box_rows, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ls_fail("cuTensorMapEncodeTiled entry point not available");
    if (cols % 4) return ls_fail("conv: weight row length (%lld) must be a multiple of 4", cols);
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return ls_fail("conv: weight pointer must be 16-byte aligned");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ls_fail("cuTensorMapEncodeTiled(2d weights) failed (%d): rows=%lld cols=%lld", (int)r, rows, cols);
    return 0;
}

// cta_group::2 switch: LS_CONV_2CTA=0 keeps every convolution on the one-CTA kernels (A/B measurements), =2 takes the pair
// kernels whenever the shape allows it, however little work there is (tests at small sizes); default 1 = heuristic
int mode_2cta() {
    static int flag = -1;
    if (flag < 0) {
        const char* e = getenv("LS_CONV_2CTA");
        flag = e ? atoi(e) : 1;
        if (flag < 0 || flag > 2) flag = 1;
    }
    return flag;
}
bool allow_2cta() { return mode_2cta() != 0; }

// launch `workers` CTAs (CG = 1) or CTA pairs (CG = 2: cluster dimension 2)
template <int CG, class Kernel>
int launch_workers(Kernel kernel, int workers, int threads, int smem, cudaStream_t stream, const CUtensorMap& m0, const CUtensorMap& m1,
                   const ConvParams& p, const char* what) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(workers * CG));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kernel, m0, m1, p) != cudaSuccess) return ls_check_cuda(what);
    return ls_check_cuda(what);
}

template <bool AMN, int BN, int STAGES, int CG>
int launch_t(const CUtensorMap& mw, const CUtensorMap& mx, const ConvParams& p, cudaStream_t stream) {
    static PerDeviceOnce once;
    constexpr int smem = smem_bytes_t(STAGES, BN, CG);
    if (once.ensure_smem(k_conv_t<AMN, BN, STAGES, CG>, smem) != cudaSuccess) return ls_check_cuda("conv smem attribute");
    const long long n_items = (long long)p.tiles_m * p.tiles_n;
    if (n_items <= 0 || n_items > 0x7fffffffLL) return ls_fail("conv: bad work-list size %lld", n_items);
    const int slots = current_sm_count() / CG;
    const int workers = n_items < slots ? (int)n_items : slots;
    return launch_workers<CG>(k_conv_t<AMN, BN, STAGES, CG>, workers, kThreadsT, smem, stream, mw, mx, p, "k_conv_t");
}

// ------------------------------------------------------------------------------------------------------------------
// k_conv_t problem description (host)
// ------------------------------------------------------------------------------------------------------------------
struct FProblem {
    Act4 in;                  // activation operand
    int P, Q;                 // iteration grid
    int R, S;                 // taps iterated
    int sy, sx, dy0, dx0, dys, dxs;
    const float* w; long long w_rows, w_cols; int w_mn;      // weights [w_rows][w_cols]; w_mn: rows = k (input channel), cols = tap*wp + output column
    int wp, wr0, wrs, ws0, wss, wS;
    float* out; float* pre; const float* bias;
    long long o_sn, o_sh, o_sw;
    int oys, oyo, oxs, oxo;
    int n_out, act;
};

int run_t(const FProblem& f, cudaStream_t stream) {
    ConvParams p{};
    p.P = f.P; p.Q = f.Q; p.n_img = f.in.N;
    const int num_sms = current_sm_count();
    const long long px256 = (long long)f.in.N * ((f.P * (long long)f.Q + 255) / 256);     // 256-pixel tiles, roughly
    // cta_group::2 (256 channels x 256 pixels per CTA pair) when the channel count fills pairs and there is work for all pairs
    const bool pair = allow_2cta() && f.n_out >= 256 && f.n_out % 256 == 0 && (mode_2cta() == 2 || px256 * (f.n_out / 256) >= num_sms / 2);
    const int CG = pair ? 2 : 1;
    const int tiles_m = (f.n_out + BM * CG - 1) / (BM * CG);
    // 256-pixel tiles unless they leave most SMs idle (small feature maps): then 128-pixel tiles double the CTA count
    int BN = 256;
    if (!pair && px256 * tiles_m * 5 < num_sms * 3) BN = 128;      // < 60 % of the SMs busy with 256-pixel tiles
    const int nsub = BN / CG;                                // pixels per CTA box
    int bw = pow2_ceil_log2(f.Q);
    const int nsub_log2 = nsub == 256 ? 8 : 7;
    if (bw > nsub_log2) bw = nsub_log2;
    if (bw < 5) bw = 5;                                      // the epilogue's 32-pixel chunks must not straddle grid rows
    while ((1 << bw) * f.sx > 256) --bw;
    if (bw < 5 || (nsub >> bw) * f.sy > 256) return ls_fail("conv: no TMA box for a %d-pixel tile at stride %d", nsub, f.sx);
    p.bw_log2 = bw;
    const int BW = 1 << bw, BH = nsub >> bw;
    p.sub_dx = p.sub_dy = 0;
    if (pair) {                                              // the pair's tile: two boxes side by side, else stacked
        if (f.Q >= 2 * BW) p.sub_dx = BW; else p.sub_dy = BH;
    }
    p.tiles_x = (f.Q + BW + p.sub_dx - 1) / (BW + p.sub_dx);
    p.tiles_y = (f.P + BH + p.sub_dy - 1) / (BH + p.sub_dy);
    p.tiles_m = tiles_m;
    p.tiles_n = f.in.N * p.tiles_x * p.tiles_y;
    p.splits = 1;
    p.R = f.R; p.S = f.S; p.chunks = (f.in.C + BK - 1) / BK;
    p.nkb_total = p.kb_per_split = p.R * p.S * p.chunks;
    p.a = ActOp{1 << 30, 1, f.sx, f.sy, f.dx0, f.dy0, f.dxs, f.dys};
    p.b = p.a;
    p.wp = f.wp; p.wr0 = f.wr0; p.wrs = f.wrs; p.ws0 = f.ws0; p.wss = f.wss; p.wS = f.wS;
    p.o_sn = f.o_sn; p.o_sh = f.o_sh; p.o_sw = f.o_sw;
    p.oys = f.oys; p.oyo = f.oyo; p.oxs = f.oxs; p.oxo = f.oxo;
    p.n_out = f.n_out; p.act = f.act;
    p.out = f.out; p.pre_out = f.pre; p.bias = f.bias;
    if (p.nkb_total <= 0 || p.tiles_n <= 0) return ls_fail("conv: empty problem");
    CUtensorMap mw, mx;
    if (make_act_map(&mx, f.in, BW, BH, f.sx, f.sy, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (f.w_mn) {
        if (make_w_map(&mw, f.w, f.w_rows, f.w_cols, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
        if (pair) return launch_t<true, 256, 6, 2>(mw, mx, p, stream);
        return BN == 256 ? launch_t<true, 256, 4, 1>(mw, mx, p, stream) : launch_t<true, 128, 6, 1>(mw, mx, p, stream);
    }
    if (make_w_map(&mw, f.w, f.w_rows, f.w_cols, BM, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (pair) return launch_t<false, 256, 6, 2>(mw, mx, p, stream);
    return BN == 256 ? launch_t<false, 256, 4, 1>(mw, mx, p, stream) : launch_t<false, 128, 6, 1>(mw, mx, p, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// k_conv_w problem description (host): out[(tap_r, ch_r), (tap_c, ch_c)] = sum_pixels A[pixel + tap_r][ch_r] * B[pixel + tap_c][ch_c]
// ------------------------------------------------------------------------------------------------------------------
struct WOperand {
    Act4 t;
    int C;               // channels used (== t.C)
    int taps, S;         // filter taps carried by this operand (1 = unshifted)
    int sx, sy, dx0, dy0;
};

template <int BN, int STAGES, int CG>
int launch_w(const CUtensorMap& ma, const CUtensorMap& mb, const ConvParams& p, cudaStream_t stream) {
    static PerDeviceOnce once;
    constexpr int smem = smem_bytes_w(STAGES, BN, CG);
    if (once.ensure_smem(k_conv_w<BN, STAGES, CG>, smem) != cudaSuccess) return ls_check_cuda("conv smem attribute");
    const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
    if (total <= 0 || total > 0x7fffffffLL) return ls_fail("conv: bad work-list size %lld", total);
    const int slots = current_sm_count() / CG;
    const int workers = total < slots ? (int)total : slots;
    return launch_workers<CG>(k_conv_w<BN, STAGES, CG>, workers, kThreadsW, smem, stream, ma, mb, p, "k_conv_w");
}

int run_w(const WOperand& A, const WOperand& B, int P, int Q, float* out, long long rs, long long cs, cudaStream_t stream) {
    ConvParams p{};
    p.P = P; p.Q = Q; p.n_img = A.t.N;
    int kw = pow2_ceil_log2(Q);
    if (kw > 6) kw = 6;
    p.bw_log2 = kw;
    const int KW = 1 << kw, KH = BKW >> kw;
    p.tiles_x = (Q + KW - 1) / KW;
    p.tiles_y = (P + KH - 1) / KH;
    p.nkb_total = A.t.N * p.tiles_x * p.tiles_y;
    const int cpa = (A.C + 31) & ~31, cpb = (B.C + 31) & ~31;
    const int M = A.taps * cpa, N = B.taps * cpb;
    const int BN = N <= 32 ? 32 : (N <= 128 ? 128 : 256);
    const bool pair = allow_2cta() && BN == 256 && M >= 256 && M % 256 == 0;
    const int CG = pair ? 2 : 1;
    p.tiles_m = (M + BM * CG - 1) / (BM * CG);
    p.tiles_n = (N + BN - 1) / BN;
    // pixel chunks: equal K-block counts, >= 16 K blocks (1024 pixels) each (amortises the red.add epilogue), ~8 items per worker
    {
        const long long tiles = (long long)p.tiles_m * p.tiles_n;
        long long want = (8ll * (current_sm_count() / CG) + tiles - 1) / tiles;
        const long long max_chunks = p.nkb_total / 16 > 0 ? p.nkb_total / 16 : 1;
        if (want > max_chunks) want = max_chunks;
        if (want < 1) want = 1;
        p.kb_per_split = (int)((p.nkb_total + want - 1) / want);
        p.splits = (p.nkb_total + p.kb_per_split - 1) / p.kb_per_split;
    }
    p.a = ActOp{A.taps > 1 ? cpa : (1 << 30), A.S, A.sx, A.sy, A.dx0, A.dy0, 1, 1};
    p.b = ActOp{B.taps > 1 ? cpb : (1 << 30), B.S, B.sx, B.sy, B.dx0, B.dy0, 1, 1};
    p.cp_r = A.taps > 1 ? cpa : (1 << 30); p.c_r = A.C; p.t_r = A.taps;
    p.cp_c = B.taps > 1 ? cpb : (1 << 30); p.c_c = B.C; p.t_c = B.taps;
    p.rs = rs; p.cs = cs;
    p.out = out;
    p.atomic = p.splits > 1 ? 1 : 0;                // chunks of one tile come from several CTAs; dw is zero-filled by the caller
    p.vec_ok = (cs == 1 && rs % 4 == 0 && B.C % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    if (p.nkb_total <= 0) return ls_fail("conv wgrad: empty problem");
    CUtensorMap ma, mb;
    if (make_act_map(&ma, A.t, KW, KH, A.sx, A.sy, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    if (make_act_map(&mb, B.t, KW, KH, B.sx, B.sy, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return -1;
    // stage = (128 + BN/CG) channels x 64 pixels x 4 B: 96 KB x 2 (BN 256), 64 KB x 3 (BN 256 pair / BN 128), 40 KB x 5 (BN 32)
    if (pair) return launch_w<256, 3, 2>(ma, mb, p, stream);
    if (BN == 256) return launch_w<256, 2, 1>(ma, mb, p, stream);
    if (BN == 128) return launch_w<128, 3, 1>(ma, mb, p, stream);
    return launch_w<32, 5, 1>(ma, mb, p, stream);
}

int validate(const LsConv2d* c) {
    if (!c) return ls_fail("conv: descriptor is NULL");
    bind_context();
    if (c->N <= 0 || c->H <= 0 || c->W <= 0 || c->Cin <= 0 || c->Cout <= 0 || c->R <= 0 || c->S <= 0)
        return ls_fail("conv: bad sizes N=%d H=%d W=%d Cin=%d Cout=%d R=%d S=%d", c->N, c->H, c->W, c->Cin, c->Cout, c->R, c->S);
    if (c->Cin % 4 || c->Cout % 4) return ls_fail("conv: Cin (%d) and Cout (%d) must be multiples of 4 (pad the channels)", c->Cin, c->Cout);
    if (c->stride < 1 || c->stride > 8) return ls_fail("conv: stride %d not in 1..8 (TMA element-stride limit)", c->stride);
    if (c->pad < 0) return ls_fail("conv: negative padding");
    if (c->transposed && (c->R != c->stride || c->S != c->stride || c->pad != 0))
        return ls_fail("conv: transposed convolutions are supported for kernel == stride, pad == 0 only");
    return 0;
}

void out_size(const LsConv2d* c, int& OH, int& OW) {
    if (c->transposed) { OH = c->H * c->stride; OW = c->W * c->stride; }
    else { OH = (c->H + 2 * c->pad - c->R) / c->stride + 1; OW = (c->W + 2 * c->pad - c->S) / c->stride + 1; }
}

}  // namespace

extern "C" int ls_conv2d_out_size(const LsConv2d* c, int32_t* out_h, int32_t* out_w) {
    if (validate(c)) return -1;
    int OH, OW;
    out_size(c, OH, OW);
    if (OH <= 0 || OW <= 0) return ls_fail("conv: empty output");
    if (out_h) *out_h = OH;
    if (out_w) *out_w = OW;
    return 0;
}

extern "C" int ls_conv2d_forward(const LsConv2d* c, const float* x, const float* w, const float* bias, float* y, float* y_pre,
                                 int32_t act, void* stream_) {
    if (validate(c)) return -1;
    if (!x || !w || !y) return ls_fail("conv forward: NULL pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    int OH, OW;
    out_size(c, OH, OW);
    if (OH <= 0 || OW <= 0) return ls_fail("conv: empty output");
    FProblem f{};
    f.in = Act4{x, c->N, c->H, c->W, c->Cin};
    f.w = w; f.out = y; f.pre = y_pre; f.bias = bias; f.act = act;
    f.n_out = c->Cout;
    f.o_sw = c->Cout; f.o_sh = (long long)OW * c->Cout; f.o_sn = (long long)OH * OW * c->Cout;
    if (!c->transposed) {
        // y[n,oy,ox,:] = sum_{r,s} x[n, oy*st + r - pad, ox*st + s - pad, :] W[:, r, s, :]^T ; W = [Cout][R*S*Cin] K-major
        f.P = OH; f.Q = OW; f.R = c->R; f.S = c->S;
        f.sy = f.sx = c->stride; f.dy0 = f.dx0 = -c->pad; f.dys = f.dxs = 1;
        f.w_rows = c->Cout; f.w_cols = (long long)c->R * c->S * c->Cin; f.w_mn = 0;
        f.wp = c->Cin; f.wr0 = 0; f.wrs = 1; f.ws0 = 0; f.wss = 1; f.wS = c->S;
        f.oys = f.oxs = 1; f.oyo = f.oxo = 0;
        return run_t(f, stream);
    }
    // transposed, kernel == stride: y[n, k*Y + r, k*X + s, :] = x[n, Y, X, :] W[:, r, s, :] ; one 1-tap GEMM per (r, s) class,
    // W = [Cin][R*S*Cout] read MN-major (rows = input channel = k, columns = tap*Cout + output channel)
    f.P = c->H; f.Q = c->W; f.R = 1; f.S = 1;
    f.sy = f.sx = 1; f.dy0 = f.dx0 = 0; f.dys = f.dxs = 1;
    f.w_rows = c->Cin; f.w_cols = (long long)c->R * c->S * c->Cout; f.w_mn = 1;
    f.wp = c->Cout; f.wrs = 1; f.wss = 1; f.wS = c->S;
    f.oys = f.oxs = c->stride;
    for (int r = 0; r < c->R; ++r)
        for (int s = 0; s < c->S; ++s) {
            f.wr0 = r; f.ws0 = s; f.oyo = r; f.oxo = s;
            if (run_t(f, stream)) return -1;
        }
    return 0;
}

extern "C" int ls_conv2d_dgrad(const LsConv2d* c, const float* dy, const float* w, float* dx, void* stream_) {
    if (validate(c)) return -1;
    if (!dy || !w || !dx) return ls_fail("conv dgrad: NULL pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    int OH, OW;
    out_size(c, OH, OW);
    if (OH <= 0 || OW <= 0) return ls_fail("conv: empty output");
    FProblem f{};
    f.in = Act4{dy, c->N, OH, OW, c->Cout};
    f.w = w; f.out = dx; f.pre = nullptr; f.bias = nullptr; f.act = LS_ACT_NONE;
    f.n_out = c->Cin;
    f.o_sw = c->Cin; f.o_sh = (long long)c->W * c->Cin; f.o_sn = (long long)c->H * c->W * c->Cin;
    if (c->transposed) {
        // dx[n,Y,X,ci] = sum_{r,s,co} dy[n, k*Y + r, k*X + s, co] W[ci, r, s, co] : a strided convolution, W K-major [Cin][R*S*Cout]
        f.P = c->H; f.Q = c->W; f.R = c->R; f.S = c->S;
        f.sy = f.sx = c->stride; f.dy0 = f.dx0 = 0; f.dys = f.dxs = 1;
        f.w_rows = c->Cin; f.w_cols = (long long)c->R * c->S * c->Cout; f.w_mn = 0;
        f.wp = c->Cout; f.wr0 = 0; f.wrs = 1; f.ws0 = 0; f.wss = 1; f.wS = c->S;
        f.oys = f.oxs = 1; f.oyo = f.oxo = 0;
        return run_t(f, stream);
    }
    // dx[n,y,x,ci] = sum_{r,s,co} dy[n, (y + pad - r)/st, (x + pad - s)/st, co] W[co, r, s, ci]  (terms with a remainder vanish):
    // one stride-1 gather per residue class (y % st, x % st), taps r = ra + st*i with ra = (py + pad) % st.
    // W = [Cout][R*S*Cin] read MN-major (rows = output channel = k, columns = tap*Cin + input channel).
    const int st = c->stride;
    f.w_rows = c->Cout; f.w_cols = (long long)c->R * c->S * c->Cin; f.w_mn = 1;
    f.wp = c->Cin; f.wS = c->S;
    f.sy = f.sx = 1; f.dys = f.dxs = -1;
    f.oys = f.oxs = st;
    // residue classes no filter tap reaches (only when stride > kernel) keep a zero gradient: clear dx once up front
    bool uncovered = false;
    for (int q = 0; q < st; ++q) uncovered |= ((q + c->pad) % st >= c->R) || ((q + c->pad) % st >= c->S);
    if (uncovered && cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)c->N * c->H * c->W * c->Cin, stream) != cudaSuccess)
        return ls_check_cuda("conv dgrad memset");
    for (int py = 0; py < st; ++py)
        for (int px = 0; px < st; ++px) {
            const int ra = (py + c->pad) % st, sa = (px + c->pad) % st;
            const int nr = ra < c->R ? (c->R - ra + st - 1) / st : 0, ns = sa < c->S ? (c->S - sa + st - 1) / st : 0;
            const int P = py < c->H ? (c->H - py + st - 1) / st : 0, Q = px < c->W ? (c->W - px + st - 1) / st : 0;
            if (P == 0 || Q == 0 || nr == 0 || ns == 0) continue;
            f.P = P; f.Q = Q; f.R = nr; f.S = ns;
            f.dy0 = (py + c->pad - ra) / st; f.dx0 = (px + c->pad - sa) / st;
            f.wr0 = ra; f.wrs = st; f.ws0 = sa; f.wss = st;
            f.oyo = py; f.oxo = px;
            if (run_t(f, stream)) return -1;
        }
    return 0;
}

extern "C" int ls_conv2d_wgrad(const LsConv2d* c, const float* dy, const float* x, float* dw, void* stream_) {
    if (validate(c)) return -1;
    if (!dy || !x || !dw) return ls_fail("conv wgrad: NULL pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    int OH, OW;
    out_size(c, OH, OW);
    if (OH <= 0 || OW <= 0) return ls_fail("conv: empty output");
    const size_t n_w = (size_t)c->Cout * c->R * c->S * c->Cin;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * n_w, stream) != cudaSuccess) return ls_check_cuda("conv wgrad memset");
    const int taps = c->R * c->S;
    if (!c->transposed) {
        // dw[co, r, s, ci] = sum_{n,oy,ox} dy[n,oy,ox,co] x[n, oy*st + r - pad, ox*st + s - pad, ci]; pixel loop over the OUTPUT grid
        WOperand g{Act4{dy, c->N, OH, OW, c->Cout}, c->Cout, 1, 1, 1, 1, 0, 0};
        WOperand xs{Act4{x, c->N, c->H, c->W, c->Cin}, c->Cin, taps, c->S, c->stride, c->stride, -c->pad, -c->pad};
        const long long ldw = (long long)taps * c->Cin;
        if (c->Cout <= 32 && taps * ((c->Cin + 31) & ~31) >= 128)       // few output channels: put the taps on the M side
            return run_w(xs, g, OH, OW, dw, 1, ldw, stream);
        return run_w(g, xs, OH, OW, dw, ldw, 1, stream);
    }
    // transposed: dw[ci, r, s, co] = sum_{n,Y,X} x[n,Y,X,ci] dy[n, k*Y + r, k*X + s, co]; pixel loop over the INPUT grid
    WOperand xi{Act4{x, c->N, c->H, c->W, c->Cin}, c->Cin, 1, 1, 1, 1, 0, 0};
    WOperand gs{Act4{dy, c->N, OH, OW, c->Cout}, c->Cout, taps, c->S, c->stride, c->stride, 0, 0};
    return run_w(xi, gs, c->H, c->W, dw, (long long)taps * c->Cout, 1, stream);
}

extern "C" int ls_act_backward(const float* dy, const float* pre, float* dx, int64_t n, int32_t act, void* stream_) {
    if (!dy || !pre || !dx) return ls_fail("act backward: NULL pointer");
    if (n <= 0 || n % 4) return ls_fail("act backward: element count %lld must be a positive multiple of 4", (long long)n);
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(dx)) & 15)
        return ls_fail("act backward: pointers must be 16-byte aligned");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)current_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k_act_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(pre),
                                                                  reinterpret_cast<float4*>(dx), n4, act);
    return ls_check_cuda("k_act_bwd");
}

// ---- nearest-2x up-sampling + 3x3 convolution (see k_upconv_fold) ---------------------------------------------------
namespace {
int validate_up(const LsConv2d* c) {
    if (validate(c)) return -1;
    if (c->transposed || c->R != 3 || c->S != 3 || c->stride != 1 || c->pad != 1)
        return ls_fail("upconv2x: the fused up-sampling needs a 3x3 / stride 1 / pad 1 convolution");
    return 0;
}
int grid_for(long long n) {
    long long b = (n + 255) / 256;
    const long long cap = (long long)current_sm_count() * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}
}  // namespace

extern "C" int ls_upconv2x_workspace(const LsConv2d* c, int64_t* floats) {
    if (validate_up(c)) return -1;
    if (!floats) return ls_fail("upconv2x: NULL size pointer");
    *floats = 2ll * 16 * c->Cout * c->Cin;        // w4 (4 classes x 2x2 taps) followed by wd (4x4 taps)
    return 0;
}

extern "C" int ls_upconv2x_forward(const LsConv2d* c, const float* x, const float* w, const float* bias, float* y, float* wk,
                                   void* stream_) {
    if (validate_up(c)) return -1;
    if (!x || !w || !y || !wk) return ls_fail("upconv2x forward: NULL pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n4 = 16ll * c->Cout * c->Cin;
    float* w4 = wk;
    float* wd = wk + n4;
    k_upconv_fold<<<grid_for(n4), 256, 0, stream>>>(w, w4, wd, c->Cout, c->Cin);
    if (ls_check_cuda("k_upconv_fold")) return -1;
    FProblem f{};
    f.in = Act4{x, c->N, c->H, c->W, c->Cin};
    f.P = c->H; f.Q = c->W; f.R = 2; f.S = 2;
    f.sy = f.sx = 1; f.dys = f.dxs = 1;
    f.w_rows = c->Cout; f.w_cols = 4ll * c->Cin; f.w_mn = 0;
    f.wp = c->Cin; f.wr0 = 0; f.wrs = 1; f.ws0 = 0; f.wss = 1; f.wS = 2;
    f.out = y; f.pre = nullptr; f.bias = bias; f.act = LS_ACT_NONE; f.n_out = c->Cout;
    f.o_sw = c->Cout; f.o_sh = 2ll * c->W * c->Cout; f.o_sn = 4ll * c->H * c->W * c->Cout;
    f.oys = f.oxs = 2;
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            f.w = w4 + (long long)(py * 2 + px) * c->Cout * 4 * c->Cin;
            f.dy0 = py - 1; f.dx0 = px - 1; f.oyo = py; f.oxo = px;
            if (run_t(f, stream)) return -1;
        }
    return 0;
}

extern "C" int ls_upconv2x_dgrad(const LsConv2d* c, const float* dy, const float* wk, float* dx, void* stream_) {
    if (validate_up(c)) return -1;
    if (!dy || !wk || !dx) return ls_fail("upconv2x dgrad: NULL pointer");
    // dx[n,Y,X,ci] = sum_{ty,tx,co} dy[n, 2Y - 1 + ty, 2X - 1 + tx, co] wd[co, ty, tx, ci]: one 4x4 stride-2 gather over dy
    FProblem f{};
    f.in = Act4{dy, c->N, 2 * c->H, 2 * c->W, c->Cout};
    f.P = c->H; f.Q = c->W; f.R = 4; f.S = 4;
    f.sy = f.sx = 2; f.dy0 = f.dx0 = -1; f.dys = f.dxs = 1;
    f.w = wk + 16ll * c->Cout * c->Cin;
    f.w_rows = c->Cout; f.w_cols = 16ll * c->Cin; f.w_mn = 1;
    f.wp = c->Cin; f.wr0 = 0; f.wrs = 1; f.ws0 = 0; f.wss = 1; f.wS = 4;
    f.out = dx; f.pre = nullptr; f.bias = nullptr; f.act = LS_ACT_NONE; f.n_out = c->Cin;
    f.o_sw = c->Cin; f.o_sh = (long long)c->W * c->Cin; f.o_sn = (long long)c->H * c->W * c->Cin;
    f.oys = f.oxs = 1; f.oyo = f.oxo = 0;
    return run_t(f, (cudaStream_t)stream_);
}

extern "C" int ls_upconv2x_wgrad(const LsConv2d* c, const float* dy, const float* x, float* dw, float* scratch, void* stream_) {
    if (validate_up(c)) return -1;
    if (!dy || !x || !dw || !scratch) return ls_fail("upconv2x wgrad: NULL pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long n4 = 16ll * c->Cout * c->Cin;
    if (cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)n4, stream) != cudaSuccess) return ls_check_cuda("upconv2x wgrad memset");
    // dw4[cls][co][a][b][ci] = sum_{n,Y,X} dy[n, 2Y+py, 2X+px, co] x[n, Y+py-1+a, X+px-1+b, ci]; pixel loop over the low-res grid
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            WOperand g{Act4{dy, c->N, 2 * c->H, 2 * c->W, c->Cout}, c->Cout, 1, 1, 2, 2, px, py};
            WOperand xs{Act4{x, c->N, c->H, c->W, c->Cin}, c->Cin, 4, 2, 1, 1, px - 1, py - 1};
            if (run_w(g, xs, c->H, c->W, scratch + (long long)(py * 2 + px) * c->Cout * 4 * c->Cin, 4ll * c->Cin, 1, stream)) return -1;
        }
    k_upconv_unfold<<<grid_for((long long)c->Cout * 9 * c->Cin), 256, 0, stream>>>(scratch, dw, c->Cout, c->Cin);
    return ls_check_cuda("k_upconv_unfold");
}
