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
//   * the 7+NC per-Gaussian partial sums of a warp are reduced with shuffles and
//     leave the SM as ONE red.global.add instruction of 7+NC lanes into a packed
//     64-byte gradient record, instead of (9+NC) atomics per pixel-Gaussian pair;
//   * warp-uniform skip of Gaussians that contribute to none of the warp's pixels;
//   * depth and alpha (mask) are differentiable channels; all views in one launch;
//   * gradients of inputs shared by the views of a scene are accumulated in-kernel.
#include "ls_common.cuh"
#include "ls_host.h"

namespace ls {

// =========================================================================================
// R.7 blend backward: one CTA per (view, tile), one pixel per thread, back to front
// =========================================================================================
template <int NC>
__global__ void __launch_bounds__(kTilePixels) k_blend_bwd(const LsRasterScene sc, const LsRasterState st,
                                                           const LsRasterGrads gr, const int ncol) {
    constexpr int CS = (NC + 3) & ~3;
    constexpr int CV = CS / 4;
    constexpr int K = 7 + NC;  // values reduced per Gaussian
    extern __shared__ __align__(16) float4 s_dyn[];
    float4(*s_geom)[kTilePixels][2] = reinterpret_cast<float4(*)[kTilePixels][2]>(s_dyn);
    float4(*s_chan)[kTilePixels][CV] = reinterpret_cast<float4(*)[kTilePixels][CV]>(s_dyn + 2 * kTilePixels * 2);
    uint32_t(*s_id)[kTilePixels] = reinterpret_cast<uint32_t(*)[kTilePixels]>(s_dyn + 2 * kTilePixels * (2 + CV));
    __shared__ uint32_t s_max;

    const int gx = (sc.W + kTile - 1) / kTile, gy = (sc.H + kTile - 1) / kTile;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31;
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
    if (tid == 0) s_max = 0;
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
    const int nb = (n + kTilePixels - 1) / kTilePixels;

    const uint64_t* __restrict__ keys = st.keys + s;
    const float* __restrict__ geom = st.geom + (size_t)v * sc.G * LS_GEOM_STRIDE;
    const float* __restrict__ chan = st.chan + (size_t)v * sc.G * CS;
    float* __restrict__ rec = gr.dL_drecord + (size_t)v * sc.G * gr.grad_stride;

    auto prefetch = [&](int b, int buf) {
        const int j = b * kTilePixels + tid;
        if (j < n) {
            const uint32_t id = (uint32_t)keys[j];
            s_id[buf][tid] = id;
            const float* gsrc = geom + (size_t)id * LS_GEOM_STRIDE;
            cp_async16(&s_geom[buf][tid][0], gsrc);
            cp_async16(&s_geom[buf][tid][1], gsrc + 4);
            const float* csrc = chan + (size_t)id * CS;
#pragma unroll
            for (int q = 0; q < CV; ++q) cp_async16(&s_chan[buf][tid][q], csrc + 4 * q);
        }
        cp_async_commit();
    };

    const float wx = (float)(tx * kTile) + 7.5f;                       // centre of the warp's 16x2 strip
    const float wy = (float)(ty * kTile + 2 * (tid >> 5)) + 0.5f;
    float T = T_final;
    float A = 0.f, last_alpha = 0.f, last_cg = 0.f;
    const float half_w = 0.5f * (float)sc.W, half_h = 0.5f * (float)sc.H;

    int it = 0;
    prefetch(nb - 1, 0);
    for (int b = nb - 1; b >= 0; --b, ++it) {
        const int buf = it & 1;
        if (b > 0) { prefetch(b - 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int cnt = min(kTilePixels, n - b * kTilePixels);
        const float4(*sg)[2] = s_geom[buf];
        const float4(*sc4)[CV] = s_chan[buf];
        for (int j = cnt - 1; j >= 0; --j) {
            const uint32_t pos = (uint32_t)(b * kTilePixels + j);
            const float4 g0 = sg[j][0];
            const float4 g1 = sg[j][1];
            const float2 ext = unpack_extent(g1.w);
            if (fabsf(g0.y - wy) > ext.y + 0.5f || fabsf(g0.x - wx) > ext.x + 7.5f) continue;  // warp-uniform
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
                    const float4 cq = sc4[j][q];
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
            // Warp reduction of K values in ~K shuffles instead of 5K: a transposing butterfly.  At each step
            // a lane keeps one half of its values and ships the other half to its partner, so the value
            // count halves while the lane distance halves; after 4 steps lane L holds the partial sum of
            // value L>>1 over its 16-lane group, a last xor-1 add completes it.  Even lanes then issue ONE
            // red.global.add over the packed gradient record (<= 64 B, one L2 line).
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
            // remaining lane bits that were not consumed by the halving (KP = 16 -> bit 0; KP = 8 -> bits 1,0)
            float tot = w_[0];
            if (KP <= 16) tot += __shfl_xor_sync(0xffffffffu, tot, 1);
            if (KP <= 8) tot += __shfl_xor_sync(0xffffffffu, tot, 2);
            constexpr int kShift = KP == 32 ? 0 : (KP == 16 ? 1 : 2);
            const int vidx = lane >> kShift;
            if ((lane & ((1 << kShift) - 1)) == 0 && vidx < K)
                atomicAdd(rec + (size_t)s_id[buf][j] * gr.grad_stride + vidx, tot);
        }
        __syncthreads();
    }
}

template <int NC>
static void launch_blend_bwd(const LsRasterScene& sc, const LsRasterState& st, const LsRasterGrads& gr, int ncol,
                             dim3 grid, cudaStream_t stream) {
    constexpr int CV = ((NC + 3) & ~3) / 4;
    constexpr size_t smem = (size_t)2 * kTilePixels * ((2 + CV) * sizeof(float4) + sizeof(uint32_t));
    if (smem > 48 * 1024) {
        static bool configured = false;  // per instantiation; attribute is idempotent
        if (!configured) {
            cudaFuncSetAttribute(k_blend_bwd<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured = true;
        }
    }
    k_blend_bwd<NC><<<grid, kTilePixels, smem, stream>>>(sc, st, gr, ncol);
}

// =========================================================================================
// R.8 per-Gaussian backward: conic -> cov2D -> (cov3D, mean), screen mean -> mean,
// depth -> mean, colour/feature SH -> (coefficients, mean).  One thread per (view, Gaussian).
// =========================================================================================
__device__ __forceinline__ void accum(float* dst, float v, bool atomic) {
    if (atomic) atomicAdd(dst, v); else *dst = v;
}

__global__ void __launch_bounds__(256) k_preprocess_bwd(const LsRasterScene sc, const LsRasterState st,
                                                        const LsRasterGrads gr) {
    __shared__ float s_stage[8][32 * kStagePitch];
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float* stage = s_stage[threadIdx.x >> 5];
    float* mine = stage + lane * kStagePitch;
    const int i0 = i - lane, nrows = min(32, sc.G - i0);       // this warp's 32 consecutive Gaussians
    const size_t vi = (size_t)v * sc.G + i;
    // lanes never return early: the SH rows and their gradients move warp-cooperatively (stage_load / stage_store)
    const bool alive = i < sc.G && st.radii[vi] > 0;
    const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
    if (alive_mask == 0u) return;                               // warp-uniform
    const int s_idx = v / sc.views_per_scene;
    const size_t si = (size_t)s_idx * sc.G + i;
    const size_t si0 = (size_t)s_idx * sc.G + i0;
    const bool at = sc.views_per_scene > 1;

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
    const bool need_dir = sc.color_mode == LS_COLOR_SH || sc.feature_mode == LS_FEATURE_SH;
    float u[3] = {0.f, 0.f, 0.f}, inv = 0.f;
    float ddir[3] = {0.f, 0.f, 0.f};
    float basis[25];
    float dbasis[25][3];
    if (need_dir && alive) {
        const float* cp = sc.campos + 3 * v;
        const float d0 = p[0] - cp[0], d1 = p[1] - cp[1], d2 = p[2] - cp[2];
        inv = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        u[0] = d0 * inv; u[1] = d1 * inv; u[2] = d2 * inv;
        const int deg = max(sc.color_mode == LS_COLOR_SH ? sc.sh_degree : 0,
                            sc.feature_mode == LS_FEATURE_SH ? sc.feature_sh_degree : 0);
        sh_basis<true>(deg, u[0], u[1], u[2], basis, dbasis);
    }
    if (sc.color_mode == LS_COLOR_PRECOMP) {
        if (alive) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) accum(gr.dL_dcolor_in + 3 * si + ch, r[7 + ch], at);
        }
    } else if (sc.color_mode == LS_COLOR_SH) {
        const int n = (sc.sh_degree + 1) * (sc.sh_degree + 1);
        float gc[3] = {0.f, 0.f, 0.f};
        if (alive) {
            const uint8_t cl = st.clamped[vi];
            gc[0] = (cl & 1) ? 0.f : r[7]; gc[1] = (cl & 2) ? 0.f : r[8]; gc[2] = (cl & 4) ? 0.f : r[9];
        }
        const float* __restrict__ sh0 = sc.color + si0 * (size_t)(n * 3);          // the warp's 32 rows
        float* dsh0 = gr.dL_dcolor_in + si0 * (size_t)(n * 3);
        for (int k0 = 0; k0 < n; k0 += 5) {                                        // 5 coefficients (15 floats) per pass
            const int cnt = min(5, n - k0);
            stage_load(stage, sh0, n * 3, k0 * 3, cnt * 3, nrows, lane);
            float out[15];
            if (alive) {
#pragma unroll
                for (int kk = 0; kk < 5; ++kk) {
                    if (kk < cnt) {
                        const int k = k0 + kk;
                        float sg = 0.f;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            out[3 * kk + ch] = basis[k] * gc[ch];
                            sg = fmaf(mine[3 * kk + ch], gc[ch], sg);
                        }
                        ddir[0] = fmaf(dbasis[k][0], sg, ddir[0]);
                        ddir[1] = fmaf(dbasis[k][1], sg, ddir[1]);
                        ddir[2] = fmaf(dbasis[k][2], sg, ddir[2]);
                    }
                }
            }
            __syncwarp();                                                          // everyone has read its sh chunk
            if (alive) {
#pragma unroll
                for (int e = 0; e < 15; ++e)
                    if (e < 3 * cnt) mine[e] = out[e];
            }
            stage_store(stage, dsh0, n * 3, k0 * 3, cnt * 3, alive_mask, at, lane);
        }
    }
    if (sc.feature_mode == LS_FEATURE_PRECOMP) {
        if (alive)
            for (int ch = 0; ch < sc.C; ++ch) accum(gr.dL_dfeature_in + si * sc.C + ch, r[7 + ncol + ch], at);
    } else if (sc.feature_mode == LS_FEATURE_SH) {
        const int n = (sc.feature_sh_degree + 1) * (sc.feature_sh_degree + 1);
        const float* __restrict__ fs0 = sc.feature + si0 * (size_t)(sc.C * n);
        float* dfs0 = gr.dL_dfeature_in + si0 * (size_t)(sc.C * n);
        const int cpc = max(1, kStagePitch / n);                                   // whole channels per pass
        for (int c0 = 0; c0 < sc.C; c0 += cpc) {
            const int nch = min(cpc, sc.C - c0);
            stage_load(stage, fs0, sc.C * n, c0 * n, nch * n, nrows, lane);
            if (alive) {
                for (int cc = 0; cc < nch; ++cc) {
                    const float gf = r[7 + ncol + c0 + cc];
                    for (int k = 0; k < n; ++k) {
                        const float sg = mine[cc * n + k] * gf;                    // read the coefficient ...
                        mine[cc * n + k] = basis[k] * gf;                          // ... then overwrite its slot with its gradient
                        ddir[0] = fmaf(dbasis[k][0], sg, ddir[0]);
                        ddir[1] = fmaf(dbasis[k][1], sg, ddir[1]);
                        ddir[2] = fmaf(dbasis[k][2], sg, ddir[2]);
                    }
                }
            }
            stage_store(stage, dfs0, sc.C * n, c0 * n, nch * n, alive_mask, at, lane);
        }
    }
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
#define LS_CASE(N) case N: launch_blend_bwd<N>(*sc, *st, *gr, ncol, grid, stream); break;
            LS_CASE(1) LS_CASE(2) LS_CASE(3) LS_CASE(4) LS_CASE(5) LS_CASE(6) LS_CASE(7) LS_CASE(8)
            LS_CASE(9) LS_CASE(10) LS_CASE(11) LS_CASE(12) LS_CASE(13) LS_CASE(14) LS_CASE(15) LS_CASE(16)
#undef LS_CASE
            default: return ls_fail("unsupported channel count %d", nc);
        }
        if (ls_check_cuda("backward blend")) return -1;
    }
    if (!(stages & LS_BWD_GEOMETRY)) return 0;
    cudaMemsetAsync(gr->dL_dmeans3D, 0, sizeof(float) * SG * 3, stream);
    cudaMemsetAsync(gr->dL_dcov3D, 0, sizeof(float) * SG * 6, stream);
    cudaMemsetAsync(gr->dL_dopacity, 0, sizeof(float) * SG, stream);
    if (gr->dL_dmeans2D) cudaMemsetAsync(gr->dL_dmeans2D, 0, sizeof(float) * VG * 3, stream);
    if (sc->color_mode == LS_COLOR_PRECOMP) cudaMemsetAsync(gr->dL_dcolor_in, 0, sizeof(float) * SG * 3, stream);
    if (sc->color_mode == LS_COLOR_SH)
        cudaMemsetAsync(gr->dL_dcolor_in, 0, sizeof(float) * SG * 3 * (sc->sh_degree + 1) * (sc->sh_degree + 1), stream);
    if (sc->feature_mode == LS_FEATURE_PRECOMP) cudaMemsetAsync(gr->dL_dfeature_in, 0, sizeof(float) * SG * sc->C, stream);
    if (sc->feature_mode == LS_FEATURE_SH)
        cudaMemsetAsync(gr->dL_dfeature_in, 0, sizeof(float) * SG * sc->C * (sc->feature_sh_degree + 1) * (sc->feature_sh_degree + 1), stream);
    dim3 grid2((sc->G + 255) / 256, sc->n_views);
    k_preprocess_bwd<<<grid2, 256, 0, stream>>>(*sc, *st, *gr);
    return ls_check_cuda("backward");
}
