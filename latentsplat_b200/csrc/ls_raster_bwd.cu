// ls_raster_bwd.cu -- backward pipeline of the sm_100a Gaussian rasterizer.
//
// Replaces the autograd backward of `GaussianRasterizer(settings)(...)`
// (/root/reference/src/model/decoder/cuda_splatting.py:150-158, driven by
// model_wrapper.py:440).  Algorithm = back-to-front replay of each tile list and the
// per-Gaussian chain rule, restated in oracle/raster_oracle.c (oracle_backward).
// B200-first differences from the lineage kernel:
//   * per-pixel state is a scalar: with g = dL/dpixel, the lineage's per-channel
//     `accum_rec[ch]` recurrence is linear, so A = sum_ch accum_rec[ch]*g[ch] obeys the
//     same recurrence with sum_ch c[ch]*g[ch]; registers drop from 3*NC to NC+4;
//   * the 7+NC per-Gaussian partial sums of a warp are reduced with a transposing
//     shuffle butterfly and leave the SM as 16-byte red.global.add.v4.f32 into a packed
//     64-byte gradient record, instead of (9+NC) atomics per pixel-Gaussian pair;
//   * the tile queues arrive through 1-D bulk TMA copies of the permuted-after-sort
//     records, and each warp compacts the queue against its 16x2 strip before the
//     replay (see ls_raster_fwd.cu);
//   * depth and alpha (mask) are differentiable channels; all views in one launch;
//   * gradients of inputs shared by the views of a scene are accumulated in-kernel.
#include "ls_common.cuh"
#include "ls_host.h"
#include "ls_tc.cuh"

namespace ls {

// =========================================================================================
// R.7 blend backward: one CTA per (view, tile), one pixel per thread, back to front
// =========================================================================================
template <int NC>
__global__ void __launch_bounds__(kTilePixels) k_blend_bwd(const LsRasterScene sc, const LsRasterState st,
                                                           const LsRasterGrads gr, const int ncol) {
    constexpr int CS = (NC + 3) & ~3;
    constexpr int CV = CS / 4;
    constexpr int RV = 2 + CV;         // float4 per queue record
    constexpr int K = 7 + NC;          // values reduced per Gaussian
    extern __shared__ __align__(128) uint8_t s_dyn[];
    float4* s_rec = reinterpret_cast<float4*>(s_dyn);                                   // [2][kBatch][RV]
    float4* s_cull = s_rec + 2 * kBatch * RV;                                           // [2][kBatch]
    uint8_t* s_list = reinterpret_cast<uint8_t*>(s_cull + 2 * kBatch);                  // [8 warps][kBatch]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_max;

    const int gx = (sc.W + kTile - 1) / kTile, gy = (sc.H + kTile - 1) / kTile;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = tx * kTile + (tid & 15), py = ty * kTile + (tid >> 4);
    const bool inside = px < sc.W && py < sc.H;
    const float fxp = (float)px, fyp = (float)py;
    const size_t hw = (size_t)sc.H * sc.W;
    const size_t pid = (size_t)py * sc.W + px;

    const size_t t = (size_t)v * gx * gy + tile;
    const long long s = st.tile_offsets[t];

    // per-pixel constants
    float g[NC];
    float g_d = 0.f, g_a = 0.f, T_final = 1.f, bg_dot = 0.f;
    uint32_t last = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) g[c] = 0.f;
    if (inside) {
        T_final = st.final_T[(size_t)v * hw + pid];
        last = st.n_contrib[(size_t)v * hw + pid];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (c < ncol) {
                if (gr.dL_dcolor) g[c] = gr.dL_dcolor[((size_t)v * 3 + c) * hw + pid];
            } else if (gr.dL_dfeature) {
                g[c] = gr.dL_dfeature[((size_t)v * (NC - ncol) + (c - ncol)) * hw + pid];
            }
        }
        if (gr.dL_ddepth) g_d = gr.dL_ddepth[(size_t)v * hw + pid];
        if (gr.dL_dalpha) g_a = gr.dL_dalpha[(size_t)v * hw + pid];
        if (ncol) {
            const float* bg = sc.bg + 3 * v;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < NC) bg_dot = fmaf(bg[c], g[c], bg_dot);
        }
    }
    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) { s_max = 0; mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_fence_init(); }
    __syncthreads();
    {
        uint32_t m = last;
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
        if (lane == 0) atomicMax(&s_max, m);
    }
    __syncthreads();
    const int n = (int)s_max;  // list entries any pixel of the tile used
    if (n == 0) return;
    const int nb = (n + kBatch - 1) / kBatch;

    float* __restrict__ rec_out = gr.dL_drecord + (size_t)v * sc.G * gr.grad_stride;

    // batches are consumed back to front; `it` counts consumed batches (buffer = it & 1, barrier phase = (it >> 1) & 1)
    auto stage = [&](int b, int buf) {
        const int cnt = min(kBatch, n - b * kBatch);
        const uint32_t bar = bar0 + 8 * buf;
        const long long first = s + (long long)b * kBatch;
        mbar_expect_tx(bar, (uint32_t)cnt * (16u + 16u * RV));
        bulk_load(smem_u32(s_rec + buf * kBatch * RV), reinterpret_cast<const float4*>(st.sorted_rec) + first * RV, (uint32_t)cnt * 16u * RV, bar);
        bulk_load(smem_u32(s_cull + buf * kBatch), reinterpret_cast<const float4*>(st.sorted_cull) + first, (uint32_t)cnt * 16u, bar);
    };

    const float wx = (float)(tx * kTile) + 7.5f;                       // centre of the warp's 16x2 strip
    const float wy = (float)(ty * kTile + 2 * warp) + 0.5f;
    float T = T_final;
    float A = 0.f, last_alpha = 0.f, last_cg = 0.f;
    const float half_w = 0.5f * (float)sc.W, half_h = 0.5f * (float)sc.H;
    uint8_t* my_list = s_list + warp * kBatch;

    int it = 0;
    if (tid == 0) stage(nb - 1, 0);
    for (int b = nb - 1; b >= 0; --b, ++it) {
        const int buf = it & 1;
        if (b > 0 && tid == 0) stage(b - 1, buf ^ 1);                  // released by the barrier that ended the previous round
        mbar_wait(bar0 + 8 * buf, (uint32_t)(it >> 1) & 1u);
        const int cnt = min(kBatch, n - b * kBatch);
        const float4* cull = s_cull + buf * kBatch;
        const float4* rec = s_rec + buf * kBatch * RV;
        // ---- phase A: compaction (ascending list order), as in the forward kernel
        int nsurv = 0;
#pragma unroll 1
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int j = k0 + lane;
            bool keep = false;
            if (j < cnt) {
                const float4 c = cull[j];
                const float2 ext = unpack_extent(c.z);
                keep = fabsf(c.y - wy) <= ext.y + 0.5f && fabsf(c.x - wx) <= ext.x + 7.5f;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) my_list[nsurv + __popc(m & ((1u << lane) - 1u))] = (uint8_t)j;
            nsurv += __popc(m);
        }
        __syncwarp();
        // ---- phase B: back-to-front replay over the survivors
        for (int i = nsurv - 1; i >= 0; --i) {
            const int j = my_list[i];
            const uint32_t pos = (uint32_t)(b * kBatch + j);
            const float4 g0 = rec[j * RV], g1 = rec[j * RV + 1];
            const float dx = g0.x - fxp, dy = g0.y - fyp;
            const float p2 = fmaf(g0.z * dx, dx, fmaf(g1.x * dy, dy, g0.w * dx * dy));
            const float Gv = ex2_approx(p2);
            const float alpha = fminf(kAlphaMax, g1.y * Gv);
            const bool hit = pos < last && p2 <= 0.f && alpha >= kAlphaMin;
            if (!__any_sync(0xffffffffu, hit)) continue;

            float val[K];
#pragma unroll
            for (int q = 0; q < K; ++q) val[q] = 0.f;
            if (hit) {
                const float one_m = 1.f - alpha;
                T = __fdividef(T, one_m);
                const float wgt = alpha * T;
                float cg = fmaf(g1.z, g_d, g_a);  // depth * g_d + 1 * g_a
#pragma unroll
                for (int q = 0; q < CV; ++q) {
                    const float4 cq = rec[j * RV + 2 + q];
                    if (4 * q + 0 < NC) { cg = fmaf(cq.x, g[4 * q + 0], cg); val[7 + 4 * q + 0] = wgt * g[4 * q + 0]; }
                    if (4 * q + 1 < NC) { cg = fmaf(cq.y, g[4 * q + 1], cg); val[7 + 4 * q + 1] = wgt * g[4 * q + 1]; }
                    if (4 * q + 2 < NC) { cg = fmaf(cq.z, g[4 * q + 2], cg); val[7 + 4 * q + 2] = wgt * g[4 * q + 2]; }
                    if (4 * q + 3 < NC) { cg = fmaf(cq.w, g[4 * q + 3], cg); val[7 + 4 * q + 3] = wgt * g[4 * q + 3]; }
                }
                A = fmaf(last_alpha, last_cg, (1.f - last_alpha) * A);
                last_cg = cg;
                last_alpha = alpha;
                float dL_dalpha = (cg - A) * T;
                dL_dalpha = fmaf(-__fdividef(T_final, one_m), bg_dot, dL_dalpha);
                const float dL_dG = g1.y * dL_dalpha;
                const float gdx = Gv * dx, gdy = Gv * dy;
                // d power / d dx = -(cxx dx + cxy dy) = (2 A' dx + B' dy) ln2 with the stored scaled conic
                const float dG_ddelx = kLn2 * fmaf(2.f * g0.z, gdx, g0.w * gdy);
                const float dG_ddely = kLn2 * fmaf(2.f * g1.x, gdy, g0.w * gdx);
                val[0] = dL_dG * dG_ddelx * half_w;
                val[1] = dL_dG * dG_ddely * half_h;
                val[2] = -0.5f * gdx * dx * dL_dG;
                val[3] = -0.5f * gdx * dy * dL_dG;
                val[4] = -0.5f * gdy * dy * dL_dG;
                val[5] = Gv * dL_dalpha;
                val[6] = wgt * g_d;
            }
            // Warp reduction of K values in ~K shuffles instead of 5K: a transposing butterfly.  At each step a lane keeps
            // one half of its values and ships the other half to its partner, so the value count halves while the lane
            // distance halves; after the halving steps lane L holds the partial sum of value L >> kShift over its lane
            // group, the remaining xor-adds complete it.  Three more shuffles gather 4 consecutive values into every
            // (4 << kShift)-th lane, which issues ONE 16-byte red.global.add.v4.f32 into the packed gradient record
            // (K scalar atomics before: 2.4 M red requests per step in ncu r01).
            constexpr int KP = K <= 8 ? 8 : (K <= 16 ? 16 : 32);
            float w_[KP];
#pragma unroll
            for (int q = 0; q < KP; ++q) w_[q] = q < K ? val[q] : 0.f;
#pragma unroll
            for (int half = KP / 2, off = 16; half >= 1; half >>= 1, off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int q = 0; q < half; ++q) {
                    const float send = up ? w_[q] : w_[q + half];
                    const float keep = up ? w_[q + half] : w_[q];
                    w_[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            float tot = w_[0];
            if (KP <= 16) tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            if (KP <= 8) tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            constexpr int kShift = KP == 32 ? 0 : (KP == 16 ? 1 : 2);          // value index = lane >> kShift
            const float v1 = __shfl_down_sync(0xffffffffu, tot, 1 << kShift);
            const float v2 = __shfl_down_sync(0xffffffffu, tot, 2 << kShift);
            const float v3 = __shfl_down_sync(0xffffffffu, tot, 3 << kShift);
            const int vidx = lane >> kShift;
            if ((lane & ((4 << kShift) - 1)) == 0 && vidx < gr.grad_stride) {   // grad_stride = round_up4(K): whole float4 groups
                const uint32_t id = __float_as_uint(cull[j].w);
                lstc::red_add_v4(rec_out + (size_t)id * gr.grad_stride + vidx, tot, v1, v2, v3);
            }
        }
        __syncthreads();  // every warp is done with buffer `buf`
    }
}

template <int NC>
static int launch_blend_bwd(const LsRasterScene& sc, const LsRasterState& st, const LsRasterGrads& gr, int ncol,
                            dim3 grid, cudaStream_t stream) {
    constexpr int RV = 2 + ((NC + 3) & ~3) / 4;
    constexpr int smem = blend_smem_bytes(RV);
    static lstc::PerDeviceOnce once;
    if (smem > 48 * 1024 && once.ensure_smem(k_blend_bwd<NC>, smem) != cudaSuccess) return ls_check_cuda("blend bwd smem attribute");
    k_blend_bwd<NC><<<grid, kTilePixels, smem, stream>>>(sc, st, gr, ncol);
    return 0;
}

// =========================================================================================
// R.8 per-Gaussian backward: conic -> cov2D -> (cov3D, mean), screen mean -> mean,
// depth -> mean, colour/feature SH -> (coefficients, mean).  One thread per (view, Gaussian).
// =========================================================================================
__device__ __forceinline__ void accum(float* dst, float v, bool atomic) {
    if (atomic) atomicAdd(dst, v); else *dst = v;
}

// FC = 0: any configuration (SH rows staged 24 floats at a time through a small per-warp buffer).
// FC = 4 | 8: the product's configuration -- colour SH degree 4 in the in-tree basis, FC feature channels of SH degree 2, dense
// gradient rows, G % 4 == 0.  Each warp's 32 consecutive coefficient rows (32 x 300 B colour, 32 x FC x 36 B features) are ONE
// contiguous piece of memory: a single cp.async.bulk brings them to shared memory while the warp does its geometry backward, the
// gradients overwrite them in place (every index a compile-time constant: basis and partial sums live in registers, no local
// frame) and one bulk store -- or cp.reduce add when a scene has several views -- takes them back.  No LSU traffic to HBM.
template <int FC>
__global__ void __launch_bounds__(256) k_preprocess_bwd(const LsRasterScene sc, const LsRasterState st,
                                                        const LsRasterGrads gr) {
    __shared__ float s_stage[FC == 0 ? 8 : 1][32 * kStagePitch];
    extern __shared__ __align__(16) unsigned char s_pre[];
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float* stage = s_stage[FC == 0 ? (threadIdx.x >> 5) : 0];
    float* mine = stage + lane * kStagePitch;
    const int i0 = i - lane, nrows = min(32, sc.G - i0);       // this warp's 32 consecutive Gaussians
    if (nrows <= 0) return;                                     // warp-uniform: the grid's tail beyond G
    const size_t vi = (size_t)v * sc.G + i;
    // lanes never return early: the SH rows and their gradients move warp-cooperatively (stage_load / stage_store)
    const bool alive = i < sc.G && st.radii[vi] > 0;
    const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
    const int s_idx = v / sc.views_per_scene;
    const size_t si = (size_t)s_idx * sc.G + i;
    const size_t si0 = (size_t)s_idx * sc.G + i0;
    // several views per scene: gradients of the shared inputs accumulate with atomics into buffers the host zero-filled.
    // one view per scene: every row is written exactly once -- culled Gaussians write their zeros here, no memset passes.
    const bool at = sc.views_per_scene > 1;
    if (at && alive_mask == 0u) return;                         // warp-uniform
    const uint32_t row_mask = at ? alive_mask : (nrows >= 32 ? 0xffffffffu : ((1u << nrows) - 1u));
    constexpr int kCRow = 75, kFRow = FC * 9;                      // floats per colour / feature coefficient row (fast path)
    float* wbuf = nullptr;
    uint32_t wbar = 0;
    if constexpr (FC > 0) {
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        wbuf = reinterpret_cast<float*>(s_pre) + (size_t)warp * 32 * (kCRow + kFRow);
        wbar = smem_u32(reinterpret_cast<uint64_t*>(s_pre + (size_t)nwarps * 32 * (kCRow + kFRow) * 4) + warp);
        if (lane == 0) {
            mbar_init(wbar, 1);
            mbar_fence_init();
            mbar_expect_tx(wbar, (uint32_t)nrows * (kCRow + kFRow) * 4u);
            bulk_load(smem_u32(wbuf), sc.color + si0 * kCRow, (uint32_t)nrows * kCRow * 4u, wbar);
            bulk_load(smem_u32(wbuf + 32 * kCRow), sc.feature + si0 * kFRow, (uint32_t)nrows * kFRow * 4u, wbar);
        }
        __syncwarp();
    }
    if (!at && i < sc.G && !alive) {
#pragma unroll
        for (int k = 0; k < 3; ++k) gr.dL_dmeans3D[3 * si + k] = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) gr.dL_dcov3D[6 * si + k] = 0.f;
        gr.dL_dopacity[si] = 0.f;
        if (gr.dL_dmeans2D) { gr.dL_dmeans2D[3 * vi] = 0.f; gr.dL_dmeans2D[3 * vi + 1] = 0.f; gr.dL_dmeans2D[3 * vi + 2] = 0.f; }
        if (sc.color_mode == LS_COLOR_PRECOMP)
            for (int ch = 0; ch < 3; ++ch) gr.dL_dcolor_in[3 * si + ch] = 0.f;
        if (sc.feature_mode == LS_FEATURE_PRECOMP)
            for (int ch = 0; ch < sc.C; ++ch) gr.dL_dfeature_in[si * sc.C + ch] = 0.f;
    }

    float p[3] = {0.f, 0.f, 0.f}, dmean[3] = {0.f, 0.f, 0.f};
    const float scale = sc.scene_scale ? sc.scene_scale[v] : 1.0f;
    const float* __restrict__ r = gr.dL_drecord + vi * gr.grad_stride;       // dereferenced by alive lanes only
    if (alive) {
    const float* __restrict__ vm = sc.viewmatrix + 16 * v;
    const float* __restrict__ pm = sc.projmatrix + 16 * v;
    const float scale2 = scale * scale;
    const float tanx = sc.tanfov[2 * v], tany = sc.tanfov[2 * v + 1];
    const float fx = div_((float)sc.W, mul_(2.0f, tanx)), fy = div_((float)sc.H, mul_(2.0f, tany));
    p[0] = mul_(sc.means3D[3 * si], scale); p[1] = mul_(sc.means3D[3 * si + 1], scale); p[2] = mul_(sc.means3D[3 * si + 2], scale);
    float cv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = mul_(sc.cov3D[6 * si + k], scale * scale);
    Cov2D q;
    cov2d(p, fx, fy, tanx, tany, cv, vm, q);

    const float g2x = r[0], g2y = r[1], gcx = r[2], gcy = r[3], gcz = r[4], gop = r[5], gdep = r[6];

    // conic = inverse(cov2D)
    const float a = q.a, b = q.b, c = q.c;
    const float denom = a * c - b * b;
    const float d2inv = 1.0f / (denom * denom + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    if (d2inv != 0.f) {
        dL_da = d2inv * (-c * c * gcx + 2.f * b * c * gcy + (denom - a * c) * gcz);
        dL_dc = d2inv * (-a * a * gcz + 2.f * a * b * gcy + (denom - a * c) * gcx);
        dL_db = d2inv * 2.f * (b * c * gcx - (denom + 2.f * b * b) * gcy + a * b * gcz);
        const float(*M)[3] = q.M;
        float dcov[6];
        dcov[0] = M[0][0] * M[0][0] * dL_da + M[0][0] * M[1][0] * dL_db + M[1][0] * M[1][0] * dL_dc;
        dcov[3] = M[0][1] * M[0][1] * dL_da + M[0][1] * M[1][1] * dL_db + M[1][1] * M[1][1] * dL_dc;
        dcov[5] = M[0][2] * M[0][2] * dL_da + M[0][2] * M[1][2] * dL_db + M[1][2] * M[1][2] * dL_dc;
        dcov[1] = 2.f * M[0][0] * M[0][1] * dL_da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * dL_db + 2.f * M[1][0] * M[1][1] * dL_dc;
        dcov[2] = 2.f * M[0][0] * M[0][2] * dL_da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * dL_db + 2.f * M[1][0] * M[1][2] * dL_dc;
        dcov[4] = 2.f * M[0][2] * M[0][1] * dL_da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * dL_db + 2.f * M[1][1] * M[1][2] * dL_dc;
#pragma unroll
        for (int k = 0; k < 6; ++k) accum(gr.dL_dcov3D + 6 * si + k, dcov[k] * scale2, at);
    }
    float dM0[3], dM1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dM0[k] = 2.f * q.v0[k] * dL_da + q.v1[k] * dL_db;
        dM1[k] = 2.f * q.v1[k] * dL_dc + q.v0[k] * dL_db;
    }
    float dJ00 = 0.f, dJ02 = 0.f, dJ11 = 0.f, dJ12 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dJ00 += dM0[j] * vm[4 * j + 0];
        dJ02 += dM0[j] * vm[4 * j + 2];
        dJ11 += dM1[j] * vm[4 * j + 1];
        dJ12 += dM1[j] * vm[4 * j + 2];
    }
    const float tz = 1.f / q.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
    float dt[3];
    dt[0] = q.xmul * -fx * tz2 * dJ02;
    dt[1] = q.ymul * -fy * tz2 * dJ12;
    dt[2] = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * q.t[0]) * tz3 * dJ02 + (2.f * fy * q.t[1]) * tz3 * dJ12;
    dt[2] += gdep;  // depth channel = row 2 of the view transform
#pragma unroll
    for (int j = 0; j < 3; ++j) dmean[j] = vm[4 * j + 0] * dt[0] + vm[4 * j + 1] * dt[1] + vm[4 * j + 2] * dt[2];

    // screen-space mean (dL_dmean2D already carries the NDC->pixel factor)
    const float hx = xform_row(pm, 0, p[0], p[1], p[2]);
    const float hy = xform_row(pm, 1, p[0], p[1], p[2]);
    const float hw = xform_row(pm, 3, p[0], p[1], p[2]);
    const float mw = 1.f / (hw + 0.0000001f);
    const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
    dmean[0] += (pm[0] * mw - pm[3] * mul1) * g2x + (pm[1] * mw - pm[3] * mul2) * g2y;
    dmean[1] += (pm[4] * mw - pm[7] * mul1) * g2x + (pm[5] * mw - pm[7] * mul2) * g2y;
    dmean[2] += (pm[8] * mw - pm[11] * mul1) * g2x + (pm[9] * mw - pm[11] * mul2) * g2y;

    accum(gr.dL_dopacity + si, gop, at);
    if (gr.dL_dmeans2D) {
        gr.dL_dmeans2D[3 * vi + 0] = g2x;
        gr.dL_dmeans2D[3 * vi + 1] = g2y;
        gr.dL_dmeans2D[3 * vi + 2] = 0.f;
    }

    }   // alive (geometry)

    // colour / feature inputs
    const int ncol = n_color(sc.color_mode);
    const bool color_sh = color_is_sh(sc.color_mode), permuted = sc.color_mode == LS_COLOR_SH_3DGS;
    const bool need_dir = color_sh || sc.feature_mode == LS_FEATURE_SH;
    float u[3] = {0.f, 0.f, 0.f}, inv = 0.f;
    float ddir[3] = {0.f, 0.f, 0.f};
    float basis[25];
    float dbasis[FC == 0 ? 25 : 1][3];
    if (need_dir && alive) {
        const float* cp = sc.campos + 3 * v;
        const float d0 = p[0] - cp[0], d1 = p[1] - cp[1], d2 = p[2] - cp[2];
        inv = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        u[0] = d0 * inv; u[1] = d1 * inv; u[2] = d2 * inv;
        if constexpr (FC > 0) {
            sh_basis<false>(4, u[0], u[1], u[2], basis, nullptr);
        } else if (permuted) {     // [EXT] 3DGS order: in-tree polynomials at (y, z, x), k = 14 patched; ddir is un-permuted after the colour loop
            sh_basis<true>(sc.sh_degree, u[1], u[2], u[0], basis, dbasis);
            sh_patch_3dgs<true>(sc.sh_degree, u[1], u[2], u[0], basis, dbasis);
        } else {
            const int deg = max(color_sh ? sc.sh_degree : 0, sc.feature_mode == LS_FEATURE_SH ? sc.feature_sh_degree : 0);
            sh_basis<true>(deg, u[0], u[1], u[2], basis, dbasis);
        }
    }
    if constexpr (FC > 0) {
        // ---- specialised coefficient backward: rows in shared memory, every index a compile-time constant ----
        float* crow = wbuf + lane * kCRow;
        const uint32_t frow = smem_u32(wbuf + 32 * kCRow) + (uint32_t)lane * kFRow * 4u;
        mbar_wait(wbar, 0);                                                        // the warp's rows have landed
        if (alive) {
            const uint8_t cl = st.clamped[vi];
            const float g0 = (cl & 1) ? 0.f : r[7], g1 = (cl & 2) ? 0.f : r[8], g2 = (cl & 4) ? 0.f : r[9];
            float sg[25];
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                const float s0 = crow[3 * k], s1 = crow[3 * k + 1], s2 = crow[3 * k + 2];
                sg[k] = fmaf(s0, g0, fmaf(s1, g1, s2 * g2));
                crow[3 * k] = basis[k] * g0; crow[3 * k + 1] = basis[k] * g1; crow[3 * k + 2] = basis[k] * g2;
            }
            sh_basis_vjp<4>(u[0], u[1], u[2], sg, ddir);
            float gf[FC], sf[9];
#pragma unroll
            for (int ch = 0; ch < FC; ++ch) gf[ch] = r[10 + ch];
#pragma unroll
            for (int k = 0; k < 9; ++k) sf[k] = 0.f;
#pragma unroll
            for (int j = 0; j < kFRow / 4; ++j) {                                  // flat (channel, coefficient) index, one float4 at a time
                const float4 x4 = lstc::lds128(frow + 16 * j);
                const float x[4] = {x4.x, x4.y, x4.z, x4.w};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = 4 * j + e, ch = f / 9, k = f - 9 * ch;           // compile-time after unrolling
                    sf[k] = fmaf(x[e], gf[ch], sf[k]);
                    o[e] = basis[k] * gf[ch];
                }
                lstc::sts128(frow + 16 * j, o[0], o[1], o[2], o[3]);
            }
            sh_basis_vjp<2>(u[0], u[1], u[2], sf, ddir);
        } else {                                                                   // culled here: its rows carry zeros
#pragma unroll
            for (int k = 0; k < kCRow; ++k) crow[k] = 0.f;
#pragma unroll
            for (int j = 0; j < kFRow / 4; ++j) lstc::sts128(frow + 16 * j, 0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();                                                  // generic-proxy writes -> visible to the bulk copy engine
        __syncwarp();
        if (lane == 0) {
            float* dc = gr.dL_dcolor_in + si0 * kCRow;
            float* df = gr.dL_dfeature_in + si0 * kFRow;
            if (at) {
                bulk_reduce_add_f32(dc, smem_u32(wbuf), (uint32_t)nrows * kCRow * 4u);
                bulk_reduce_add_f32(df, smem_u32(wbuf + 32 * kCRow), (uint32_t)nrows * kFRow * 4u);
            } else {
                bulk_store(dc, smem_u32(wbuf), (uint32_t)nrows * kCRow * 4u);
                bulk_store(df, smem_u32(wbuf + 32 * kCRow), (uint32_t)nrows * kFRow * 4u);
            }
            bulk_commit();
            bulk_wait_read_all();                                                  // shared memory must outlive the engine's reads
        }
    } else {
    if (sc.color_mode == LS_COLOR_PRECOMP) {
        if (alive) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) accum(gr.dL_dcolor_in + 3 * si + ch, r[7 + ch], at);
        }
    } else if (color_sh) {
        const int n = (sc.sh_degree + 1) * (sc.sh_degree + 1), row = 3 * n;
        const int pitch = gr.color_grad_pitch > 0 ? gr.color_grad_pitch : row;
        const int row_out = gr.color_grad_pitch > 0 ? min(pitch, (row + 7) & ~7) : row;   // pad columns get zeros: whole sectors
        float gc[3] = {0.f, 0.f, 0.f};
        if (alive) {
            const uint8_t cl = st.clamped[vi];
            gc[0] = (cl & 1) ? 0.f : r[7]; gc[1] = (cl & 2) ? 0.f : r[8]; gc[2] = (cl & 4) ? 0.f : r[9];
        }
        const float* __restrict__ sh0 = sc.color + si0 * (size_t)row;              // the warp's 32 rows
        float* dsh0 = gr.dL_dcolor_in + si0 * (size_t)pitch;
        for (int c0 = 0; c0 < row_out; c0 += kChunk) {                             // 8 coefficients (24 floats = 3 sectors) per pass
            const int len_in = max(0, min(kChunk, row - c0)), len_out = min(kChunk, row_out - c0);
            if (len_in > 0) stage_load(stage, sh0, row, c0, len_in, nrows, lane);
            const int k0 = c0 / 3;
            float out[kChunk];
#pragma unroll
            for (int e = 0; e < kChunk; ++e) {
                float o = 0.f;
                if (alive && e < len_in) {
                    const int k = k0 + e / 3;
                    const float gch = gc[e % 3];
                    o = basis[k] * gch;
                    const float sg = mine[e] * gch;
                    ddir[0] = fmaf(dbasis[k][0], sg, ddir[0]);
                    ddir[1] = fmaf(dbasis[k][1], sg, ddir[1]);
                    ddir[2] = fmaf(dbasis[k][2], sg, ddir[2]);
                }
                out[e] = o;
            }
            __syncwarp();                                                          // everyone has read its sh chunk
#pragma unroll
            for (int e = 0; e < kChunk; ++e)
                if (e < len_out) mine[e] = out[e];
            stage_store(stage, dsh0, pitch, c0, len_out, row_mask, at, lane);
        }
    }
    if (permuted) {                                            // gradients w.r.t. (y, z, x) back to (x, y, z)
        const float gx_ = ddir[2], gy_ = ddir[0], gz_ = ddir[1];
        ddir[0] = gx_; ddir[1] = gy_; ddir[2] = gz_;
        if (sc.feature_mode == LS_FEATURE_SH && alive) sh_basis<true>(sc.feature_sh_degree, u[0], u[1], u[2], basis, dbasis);
    }
    if (sc.feature_mode == LS_FEATURE_PRECOMP) {
        if (alive)
            for (int ch = 0; ch < sc.C; ++ch) accum(gr.dL_dfeature_in + si * sc.C + ch, r[7 + ncol + ch], at);
    } else if (sc.feature_mode == LS_FEATURE_SH) {
        const int n = (sc.feature_sh_degree + 1) * (sc.feature_sh_degree + 1), row = sc.C * n;
        const int pitch = gr.feature_grad_pitch > 0 ? gr.feature_grad_pitch : row;
        const int row_out = gr.feature_grad_pitch > 0 ? min(pitch, (row + 7) & ~7) : row;
        const uint32_t magic = (65536u + (uint32_t)n - 1u) / (uint32_t)n;          // f / n for f < 800, n <= 25
        const float* __restrict__ fs0 = sc.feature + si0 * (size_t)row;
        float* dfs0 = gr.dL_dfeature_in + si0 * (size_t)pitch;
        for (int c0 = 0; c0 < row_out; c0 += kChunk) {                             // flat (channel, coefficient) index, 24 per pass
            const int len_in = max(0, min(kChunk, row - c0)), len_out = min(kChunk, row_out - c0);
            if (len_in > 0) stage_load(stage, fs0, row, c0, len_in, nrows, lane);
            float out[kChunk];
#pragma unroll
            for (int e = 0; e < kChunk; ++e) {
                float o = 0.f;
                if (alive && e < len_in) {
                    const int f = c0 + e, ch = (int)(((uint32_t)f * magic) >> 16), k = f - ch * n;
                    const float gf = r[7 + ncol + ch];
                    o = basis[k] * gf;
                    const float sg = mine[e] * gf;
                    ddir[0] = fmaf(dbasis[k][0], sg, ddir[0]);
                    ddir[1] = fmaf(dbasis[k][1], sg, ddir[1]);
                    ddir[2] = fmaf(dbasis[k][2], sg, ddir[2]);
                }
                out[e] = o;
            }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < kChunk; ++e)
                if (e < len_out) mine[e] = out[e];
            stage_store(stage, dfs0, pitch, c0, len_out, row_mask, at, lane);
        }
    }
    }   // generic coefficient path
    if (!alive) return;                                        // no warp-level operation below this line
    if (need_dir) {
        const float dot = u[0] * ddir[0] + u[1] * ddir[1] + u[2] * ddir[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) dmean[j] += (ddir[j] - u[j] * dot) * inv;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) accum(gr.dL_dmeans3D + 3 * si + j, dmean[j] * scale, at);
}

}  // namespace ls

using namespace ls;

// 1 when the specialised coefficient backward applies to this scene configuration: the caller then allocates DENSE gradient rows
// (pitch = row length; the generic kernel prefers rows padded to 8 floats, see LsRasterGrads.color_grad_pitch)
extern "C" int ls_raster_dense_sh_grads(const LsRasterScene* sc) {
    return sc && sc->color_mode == LS_COLOR_SH && sc->sh_degree == 4 && sc->feature_mode == LS_FEATURE_SH && sc->feature_sh_degree == 2 &&
           (sc->C == 4 || sc->C == 8) && sc->G % 4 == 0;
}

extern "C" int ls_raster_backward(const LsRasterScene* sc, const LsRasterState* st, const LsRasterGrads* gr,
                                  int32_t stages, void* stream_) {
    if (ls_validate_scene(sc)) return -1;
    if (!st || !gr) return ls_fail("state/grads is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ncol = n_color(sc->color_mode);
    const int nc = ncol + sc->C;
    if (gr->grad_stride != round_up4(7 + nc)) return ls_fail("grad_stride %d != %d", gr->grad_stride, round_up4(7 + nc));
    if (sc->G == 0) return 0;  // nothing to differentiate; zero-sized outputs have NULL pointers
    if (!gr->dL_drecord || !gr->dL_dmeans3D || !gr->dL_dcov3D || !gr->dL_dopacity) return ls_fail("a required gradient pointer is NULL");
    if (ncol && !gr->dL_dcolor_in) return ls_fail("dL_dcolor_in is NULL");
    if (sc->C && !gr->dL_dfeature_in) return ls_fail("dL_dfeature_in is NULL");
    const int gx = (sc->W + kTile - 1) / kTile, gy = (sc->H + kTile - 1) / kTile;
    const size_t S = (size_t)(sc->n_views / sc->views_per_scene);
    const size_t SG = S * (size_t)sc->G, VG = (size_t)sc->n_views * sc->G;

    if (stages & LS_BWD_BLEND) {
        cudaMemsetAsync(gr->dL_drecord, 0, sizeof(float) * VG * gr->grad_stride, stream);
        dim3 grid(gx * gy, sc->n_views);
        switch (nc) {
#define LS_CASE(N) case N: if (launch_blend_bwd<N>(*sc, *st, *gr, ncol, grid, stream)) return -1; break;
            LS_CASE(1) LS_CASE(2) LS_CASE(3) LS_CASE(4) LS_CASE(5) LS_CASE(6) LS_CASE(7) LS_CASE(8)
            LS_CASE(9) LS_CASE(10) LS_CASE(11) LS_CASE(12) LS_CASE(13) LS_CASE(14) LS_CASE(15) LS_CASE(16)
#undef LS_CASE
            default: return ls_fail("unsupported channel count %d", nc);
        }
        if (ls_check_cuda("backward blend")) return -1;
    }
    if (!(stages & LS_BWD_GEOMETRY)) return 0;
    const size_t cpitch = gr->color_grad_pitch > 0 ? (size_t)gr->color_grad_pitch : (size_t)3 * (sc->sh_degree + 1) * (sc->sh_degree + 1);
    const size_t fpitch = gr->feature_grad_pitch > 0 ? (size_t)gr->feature_grad_pitch
                                                     : (size_t)sc->C * (sc->feature_sh_degree + 1) * (sc->feature_sh_degree + 1);
    if (color_is_sh(sc->color_mode) && gr->color_grad_pitch > 0 && gr->color_grad_pitch < 3 * (sc->sh_degree + 1) * (sc->sh_degree + 1))
        return ls_fail("color_grad_pitch %d is shorter than a coefficient row", gr->color_grad_pitch);
    if (sc->feature_mode == LS_FEATURE_SH && gr->feature_grad_pitch > 0 &&
        gr->feature_grad_pitch < sc->C * (sc->feature_sh_degree + 1) * (sc->feature_sh_degree + 1))
        return ls_fail("feature_grad_pitch %d is shorter than a coefficient row", gr->feature_grad_pitch);
    if (sc->views_per_scene > 1) {          // shared inputs accumulate with atomics; one view per scene writes every row itself
        cudaMemsetAsync(gr->dL_dmeans3D, 0, sizeof(float) * SG * 3, stream);
        cudaMemsetAsync(gr->dL_dcov3D, 0, sizeof(float) * SG * 6, stream);
        cudaMemsetAsync(gr->dL_dopacity, 0, sizeof(float) * SG, stream);
        if (gr->dL_dmeans2D) cudaMemsetAsync(gr->dL_dmeans2D, 0, sizeof(float) * VG * 3, stream);
        if (sc->color_mode == LS_COLOR_PRECOMP) cudaMemsetAsync(gr->dL_dcolor_in, 0, sizeof(float) * SG * 3, stream);
        if (color_is_sh(sc->color_mode)) cudaMemsetAsync(gr->dL_dcolor_in, 0, sizeof(float) * SG * cpitch, stream);
        if (sc->feature_mode == LS_FEATURE_PRECOMP) cudaMemsetAsync(gr->dL_dfeature_in, 0, sizeof(float) * SG * sc->C, stream);
        if (sc->feature_mode == LS_FEATURE_SH) cudaMemsetAsync(gr->dL_dfeature_in, 0, sizeof(float) * SG * fpitch, stream);
    }
    const int crow = 3 * (sc->sh_degree + 1) * (sc->sh_degree + 1), frow = sc->C * (sc->feature_sh_degree + 1) * (sc->feature_sh_degree + 1);
    const bool dense = (gr->color_grad_pitch == 0 || gr->color_grad_pitch == crow) && (gr->feature_grad_pitch == 0 || gr->feature_grad_pitch == frow);
    const bool aligned = ((reinterpret_cast<uintptr_t>(sc->color) | reinterpret_cast<uintptr_t>(sc->feature) |
                           reinterpret_cast<uintptr_t>(gr->dL_dcolor_in) | reinterpret_cast<uintptr_t>(gr->dL_dfeature_in)) & 15) == 0;
    if (ls_raster_dense_sh_grads(sc) && dense && aligned) {
        constexpr int kWarps = 4;
        dim3 gridf((sc->G + 32 * kWarps - 1) / (32 * kWarps), sc->n_views);
        const int smem = kWarps * 32 * (75 + sc->C * 9) * 4 + kWarps * 8;
        static lstc::PerDeviceOnce once4, once8;
        if (sc->C == 4) {
            if (smem > 48 * 1024 && once4.ensure_smem(k_preprocess_bwd<4>, smem) != cudaSuccess) return ls_check_cuda("preprocess bwd smem attribute");
            k_preprocess_bwd<4><<<gridf, 32 * kWarps, smem, stream>>>(*sc, *st, *gr);
        } else {
            if (smem > 48 * 1024 && once8.ensure_smem(k_preprocess_bwd<8>, smem) != cudaSuccess) return ls_check_cuda("preprocess bwd smem attribute");
            k_preprocess_bwd<8><<<gridf, 32 * kWarps, smem, stream>>>(*sc, *st, *gr);
        }
        return ls_check_cuda("backward");
    }
    dim3 grid2((sc->G + 255) / 256, sc->n_views);
    k_preprocess_bwd<0><<<grid2, 256, 0, stream>>>(*sc, *st, *gr);
    return ls_check_cuda("backward");
}
