// ls_raster_fwd.cu -- forward pipeline of the sm_100a Gaussian rasterizer.
//
// Replaces the forward of `GaussianRasterizer(settings)(...)` as called at
// /root/reference/src/model/decoder/cuda_splatting.py:146-158 (algorithm: the
// 3DGS tile rasterizer lineage, see oracle/raster_oracle.c for the restatement
// and its [EXT] markers).  B200-first differences from that lineage:
//   * all V views of a step in one launch sequence (grid.y / grid.x = view), no
//     per-view host loop, tan-fov read from device memory (no .item() syncs);
//   * binning = per-tile counting + exclusive scan + scatter into tile segments,
//     then ONE per-tile shared-memory radix sort of (depth_bits, id) keys instead
//     of 5 global radix passes over 64-bit (tile|depth) keys; the resulting order
//     is identical to the stable global sort (ties fall back to ascending id);
//   * the sort stage also PERMUTES: it writes every tile's queue in list order (a
//     16-byte cull record x, y, extents, id and the packed 32 B geometry + 16*k B
//     channel record per entry), so the blend kernels stage 256 queue entries with
//     ONE 1-D bulk TMA copy per array (cp.async.bulk + mbarrier, double buffered)
//     instead of 256 threads x (key load + 4 cp.async gathers);
//   * warp-level compaction: each warp tests 32 DIFFERENT queue entries per
//     instruction against the extent of its 16x2 pixel strip, ballots, and writes
//     the survivors' positions to a per-warp list; the alpha math then runs over
//     survivors only (before: every lane walked the whole queue and rejected per
//     entry -- 80 % issue-active, mostly on rejected entries);
//   * feature SH (0.5 + eval_sh, cuda_splatting.py:94-101) evaluated in the
//     preprocess kernel instead of ~30 eager torch kernels.
#include <stdarg.h>
#include <stdio.h>

#include "ls_common.cuh"
#include "ls_host.h"
#include "ls_tc.cuh"

namespace ls {

// =========================================================================================
// R.1 preprocess: one thread per (view, Gaussian)
// =========================================================================================
// FC = 0: any configuration.  FC = 4 | 8 (colour SH degree 4 in the in-tree basis, FC feature channels of SH degree 2, G % 4 == 0):
// the warp's 32 coefficient rows arrive in shared memory by one cp.async.bulk each while the threads cull and project; survivors
// evaluate their SH from there with compile-time indices (see k_preprocess_bwd<FC>).
template <int FC>
__device__ __forceinline__ void preprocess_one(const LsRasterScene& sc, const LsRasterState& st, const int v, const int i,
                                               const float* wbuf, const uint32_t wbar) {
    const int s_idx = v / sc.views_per_scene;
    const size_t vi = (size_t)v * sc.G + i;
    const size_t si = (size_t)s_idx * sc.G + i;
    const int gx = (sc.W + kTile - 1) / kTile, gy = (sc.H + kTile - 1) / kTile;

    float4* grec = reinterpret_cast<float4*>(st.geom + vi * LS_GEOM_STRIDE);
    st.radii[vi] = 0;
    st.tiles_touched[vi] = 0;
    st.clamped[vi] = 0;
    // a culled Gaussian leaves a zero geometry record (written on the way out: survivors write their record once, not twice)
    auto cull = [&]() { grec[0] = make_float4(0.f, 0.f, 0.f, 0.f); grec[1] = make_float4(0.f, 0.f, 0.f, 0.f); };

    const float* __restrict__ vm = sc.viewmatrix + 16 * v;
    const float* __restrict__ pm = sc.projmatrix + 16 * v;
    const float scale = sc.scene_scale ? sc.scene_scale[v] : 1.0f;
    const float scale2 = mul_(scale, scale);
    const float tanx = sc.tanfov[2 * v], tany = sc.tanfov[2 * v + 1];
    const float fx = div_((float)sc.W, mul_(2.0f, tanx)), fy = div_((float)sc.H, mul_(2.0f, tany));

    const float p[3] = {mul_(sc.means3D[3 * si], scale), mul_(sc.means3D[3 * si + 1], scale),
                        mul_(sc.means3D[3 * si + 2], scale)};
    const float zv = xform_row(vm, 2, p[0], p[1], p[2]);
    if (zv <= 0.2f) { cull(); return; }  // [EXT] near cull
    const float hx = xform_row(pm, 0, p[0], p[1], p[2]);
    const float hy = xform_row(pm, 1, p[0], p[1], p[2]);
    const float hw = xform_row(pm, 3, p[0], p[1], p[2]);
    const float pw = div_(1.0f, add_(hw, 0.0000001f));
    const float ppx = mul_(hx, pw), ppy = mul_(hy, pw);

    float cv[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) cv[k] = mul_(sc.cov3D[6 * si + k], scale2);
    Cov2D q;
    cov2d(p, fx, fy, tanx, tany, cv, vm, q);
    const float det = sub_(mul_(q.a, q.c), mul_(q.b, q.b));
    if (det == 0.0f) { cull(); return; }
    const float det_inv = div_(1.0f, det);
    const float cxx = mul_(q.c, det_inv), cxy = mul_(-q.b, det_inv), cyy = mul_(q.a, det_inv);
    const float mid = mul_(0.5f, add_(q.a, q.c));
    const float disc = __fsqrt_rn(fmaxf(0.1f, sub_(mul_(mid, mid), det)));
    const float l1 = add_(mid, disc), l2 = sub_(mid, disc);
    const int radius = (int)ceilf(mul_(3.0f, __fsqrt_rn(fmaxf(l1, l2))));
    const float px = mul_(sub_(mul_(add_(ppx, 1.0f), (float)sc.W), 1.0f), 0.5f);
    const float py = mul_(sub_(mul_(add_(ppy, 1.0f), (float)sc.H), 1.0f), 0.5f);
    int rmin[2], rmax[2];
    get_rect(px, py, radius, gx, gy, rmin, rmax);
    const int ntiles = (rmax[0] - rmin[0]) * (rmax[1] - rmin[1]);
    if (ntiles == 0) { cull(); return; }

    // ---- colour / feature values of this Gaussian for this view -----------------------
    const int ncol = n_color(sc.color_mode);
    float* crec = st.chan + vi * st.chan_stride;
    const bool color_sh = color_is_sh(sc.color_mode), permuted = sc.color_mode == LS_COLOR_SH_3DGS;
    const bool need_dir = color_sh || sc.feature_mode == LS_FEATURE_SH;
    float col3[3] = {0.f, 0.f, 0.f};                               // colour of the specialised path (stored with the features)
    float basis[25];
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (need_dir) {
        const float* cp = sc.campos + 3 * v;
        d0 = p[0] - cp[0]; d1 = p[1] - cp[1]; d2 = p[2] - cp[2];
        const float inv = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        d0 *= inv; d1 *= inv; d2 *= inv;
        if (permuted) {            // [EXT] 3DGS order: in-tree polynomials at (y, z, x), k = 14 patched; features keep the in-tree basis
            sh_basis<false>(sc.sh_degree, d1, d2, d0, basis, nullptr);
            sh_patch_3dgs<false>(sc.sh_degree, d1, d2, d0, basis, nullptr);
        } else {
            const int deg = max(color_sh ? sc.sh_degree : 0, sc.feature_mode == LS_FEATURE_SH ? sc.feature_sh_degree : 0);
            sh_basis<false>(deg, d0, d1, d2, basis, nullptr);
        }
    }
    if (color_sh) {
        float r = 0.f, g = 0.f, b = 0.f;
        if constexpr (FC > 0) {
            mbar_wait(wbar, 0);                                    // the warp's rows have landed
            const float* __restrict__ sh = wbuf + (threadIdx.x & 31) * 75;       // row pitch 75: conflict-free
#pragma unroll
            for (int k = 0; k < 25; ++k) {
                r = fmaf(basis[k], sh[3 * k + 0], r);
                g = fmaf(basis[k], sh[3 * k + 1], g);
                b = fmaf(basis[k], sh[3 * k + 2], b);
            }
        } else {
            const int n = (sc.sh_degree + 1) * (sc.sh_degree + 1);
            const float* __restrict__ sh = sc.color + si * (size_t)(n * 3);
            for (int k = 0; k < n; ++k) {
                r = fmaf(basis[k], sh[3 * k + 0], r);
                g = fmaf(basis[k], sh[3 * k + 1], g);
                b = fmaf(basis[k], sh[3 * k + 2], b);
            }
        }
        r += 0.5f; g += 0.5f; b += 0.5f;
        st.clamped[vi] = (uint8_t)((r < 0.f ? 1 : 0) | (g < 0.f ? 2 : 0) | (b < 0.f ? 4 : 0));
        if constexpr (FC > 0) { col3[0] = fmaxf(r, 0.f); col3[1] = fmaxf(g, 0.f); col3[2] = fmaxf(b, 0.f); }
        else { crec[0] = fmaxf(r, 0.f); crec[1] = fmaxf(g, 0.f); crec[2] = fmaxf(b, 0.f); }
    } else if (sc.color_mode == LS_COLOR_PRECOMP) {
        crec[0] = sc.color[3 * si]; crec[1] = sc.color[3 * si + 1]; crec[2] = sc.color[3 * si + 2];
    }
    if (sc.feature_mode == LS_FEATURE_PRECOMP) {
        for (int c = 0; c < sc.C; ++c) crec[ncol + c] = sc.feature[si * sc.C + c];
    } else if (sc.feature_mode == LS_FEATURE_SH) {
        if constexpr (FC > 0) {
            const uint32_t frow = smem_u32(wbuf + 32 * 75) + (uint32_t)(threadIdx.x & 31) * FC * 36u;
            float acc[FC];
#pragma unroll
            for (int c = 0; c < FC; ++c) acc[c] = 0.f;
#pragma unroll
            for (int j = 0; j < FC * 9 / 4; ++j) {                 // flat (channel, coefficient) index, one float4 at a time
                const float4 x4 = lstc::lds128(frow + 16 * j);
                const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = 4 * j + e, c = f / 9, k = f - 9 * c;           // compile-time after unrolling
                    acc[c] = fmaf(basis[k], x[e], acc[c]);
                }
            }
            // the whole channel record (colour + features, padded to float4s) leaves in 16-byte stores
            constexpr int kVals = (3 + FC + 3) / 4 * 4;
            float vals[kVals];
#pragma unroll
            for (int q = 0; q < kVals; ++q) vals[q] = q < 3 ? col3[q] : (q < 3 + FC ? 0.5f + acc[q - 3] : 0.f);
#pragma unroll
            for (int q = 0; q < kVals; q += 4)
                *reinterpret_cast<float4*>(crec + q) = make_float4(vals[q], vals[q + 1], vals[q + 2], vals[q + 3]);
        } else {
            if (permuted) sh_basis<false>(sc.feature_sh_degree, d0, d1, d2, basis, nullptr);
            const int n = (sc.feature_sh_degree + 1) * (sc.feature_sh_degree + 1);
            const float* __restrict__ fs = sc.feature + si * (size_t)(sc.C * n);
            for (int c = 0; c < sc.C; ++c) {
                float r = 0.f;
                for (int k = 0; k < n; ++k) r = fmaf(basis[k], fs[c * n + k], r);
                crec[ncol + c] = 0.5f + r;  // cuda_splatting.py:97
            }
        }
    }

    // ---- geometry record ------------------------------------------------------------------
    // Conservative half-extents (pixels) of the region where alpha can reach 1/255:
    //   o * exp(-q) >= 1/255  <=>  q <= tau = ln(255 o),  and  min_dx q(dx,dy) = dy^2 / (2 Sigma_yy)
    // so |dy| <= sqrt(2 tau Sigma_yy), |dx| <= sqrt(2 tau Sigma_xx) with Sigma = cov2D (a, b, c).  The blend
    // kernels use them for a warp-uniform reject before touching the per-pixel math; +0.2 % and fp16
    // round-up keep the test strictly conservative w.r.t. the fp32 evaluation.  tau < 0: never visible.
    const float opac = sc.opacity[si];
    const float tau = logf(255.0f * opac);
    float ex = -1.0f, ey = -1.0f;
    if (tau >= 0.0f) {
        ex = sqrtf(2.0f * tau * q.a) * 1.002f + 0.01f;
        ey = sqrtf(2.0f * tau * q.c) * 1.002f + 0.01f;
    }
    const __half2 ext = __halves2half2(__float2half_ru(ex), __float2half_ru(ey));
    grec[0] = make_float4(px, py, -0.5f * kLog2e * cxx, -kLog2e * cxy);
    grec[1] = make_float4(-0.5f * kLog2e * cyy, opac, zv, __uint_as_float(*reinterpret_cast<const uint32_t*>(&ext)));
    st.radii[vi] = radius;
    st.tiles_touched[vi] = (uint32_t)ntiles;

    // ---- tile counting --------------------------------------------------------------------
    uint32_t* cnt = st.tile_count + (size_t)v * gx * gy;
    for (int y = rmin[1]; y < rmax[1]; ++y)
        for (int x = rmin[0]; x < rmax[0]; ++x) atomicAdd(cnt + y * gx + x, 1u);
}

template <int FC>
__global__ void __launch_bounds__(256) k_preprocess(const LsRasterScene sc, const LsRasterState st) {
    extern __shared__ __align__(16) unsigned char s_pre[];
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* wbuf = nullptr;
    uint32_t wbar = 0;
    if constexpr (FC > 0) {
        constexpr int kRow = 75 + FC * 9;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const int i0 = i - lane, nrows = min(32, sc.G - i0);
        if (nrows <= 0) return;                                    // warp-uniform
        float* buf = reinterpret_cast<float*>(s_pre) + (size_t)warp * 32 * kRow;
        wbuf = buf;
        wbar = smem_u32(reinterpret_cast<uint64_t*>(s_pre + (size_t)nwarps * 32 * kRow * 4) + warp);
        if (lane == 0) {
            const size_t si0 = (size_t)(v / sc.views_per_scene) * sc.G + i0;
            mbar_init(wbar, 1);
            mbar_fence_init();
            mbar_expect_tx(wbar, (uint32_t)nrows * kRow * 4u);
            bulk_load(smem_u32(buf), sc.color + si0 * 75, (uint32_t)nrows * 300u, wbar);
            bulk_load(smem_u32(buf + 32 * 75), sc.feature + si0 * (FC * 9), (uint32_t)nrows * FC * 36u, wbar);
        }
        __syncwarp();
        if (i < sc.G) preprocess_one<FC>(sc, st, v, i, wbuf, wbar);
        if (lane == 0) mbar_wait(wbar, 0);                         // never leave with the copy still in flight
    } else {
        if (i < sc.G) preprocess_one<0>(sc, st, v, i, nullptr, 0u);
    }
}

// =========================================================================================
// R.2 exclusive scan of the per-tile counts (one CTA; N = V*T is a few thousand)
// =========================================================================================
__global__ void __launch_bounds__(1024) k_scan_tiles(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ off,
                                                     uint32_t* __restrict__ stats, int n, long long capacity) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_carry, s_max;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { s_carry = 0; s_max = 0; }
    __syncthreads();
    uint32_t mx = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const uint32_t c = i < n ? cnt[i] : 0u;
        mx = max(mx, c);
        uint32_t x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t s = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += y;
            }
            wsum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = x - c + (w ? wsum[w - 1] : 0u) + carry;
        if (i < n) off[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = carry + wsum[31];
        __syncthreads();
    }
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    if (lane == 0) atomicMax(&s_max, mx);
    __syncthreads();
    if (tid == 0) {
        off[n] = s_carry;
        stats[0] = s_carry;
        stats[1] = s_max;
        stats[2] = (capacity >= 0 && (long long)s_carry > capacity) ? 1u : 0u;
        stats[3] = 0u;
    }
}

// =========================================================================================
// R.3 scatter (depth_bits << 32 | id) into the tile segments
// =========================================================================================
__global__ void __launch_bounds__(256) k_scatter_keys(const LsRasterScene sc, const LsRasterState st) {
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sc.G) return;
    const size_t vi = (size_t)v * sc.G + i;
    const int radius = st.radii[vi];
    if (radius <= 0) return;
    const int gx = (sc.W + kTile - 1) / kTile, gy = (sc.H + kTile - 1) / kTile;
    const float4 g0 = *reinterpret_cast<const float4*>(st.geom + vi * LS_GEOM_STRIDE);
    const float depth = st.geom[vi * LS_GEOM_STRIDE + 6];
    int rmin[2], rmax[2];
    get_rect(g0.x, g0.y, radius, gx, gy, rmin, rmax);
    const uint64_t key = ((uint64_t)__float_as_uint(depth) << 32) | (uint32_t)i;
    const size_t tbase = (size_t)v * gx * gy;
    for (int y = rmin[1]; y < rmax[1]; ++y)
        for (int x = rmin[0]; x < rmax[0]; ++x) {
            const size_t t = tbase + y * gx + x;
            const uint32_t pos = atomicSub(st.tile_count + t, 1u) - 1u;  // unique slot in [0, count)
            const long long dst = (long long)st.tile_offsets[t] + pos;
            if (dst < st.capacity) st.keys[dst] = key;
        }
}

// =========================================================================================
// R.4 per-tile LSD radix sort of 64-bit keys (8-bit digits), one CTA per (view, tile).
// Tiles that fit `smem_keys` are sorted entirely in shared memory (one HBM read + one
// write of the segment); longer ones ping-pong between keys and keys_tmp (L2-resident).
// Digit positions on which all keys of the tile agree are skipped -- the id half of the
// key only matters among equal depths and its high bytes are constant.
// =========================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

// write-out of one sorted queue entry: key, cull record (x, y, extents, id), packed geometry + channel record
__device__ __forceinline__ void write_entry(uint64_t k, long long dst, uint64_t* __restrict__ keys, const float* __restrict__ geom,
                                            const float* __restrict__ chan, float* __restrict__ cull, float* __restrict__ rec,
                                            int cv, bool write_key) {
    if (write_key) keys[dst] = k;
    const uint32_t id = (uint32_t)k;
    const float4* g = reinterpret_cast<const float4*>(geom + (size_t)id * LS_GEOM_STRIDE);
    const float4 g0 = __ldg(g), g1 = __ldg(g + 1);
    reinterpret_cast<float4*>(cull)[dst] = make_float4(g0.x, g0.y, g1.w, __uint_as_float(id));
    float4* r = reinterpret_cast<float4*>(rec) + dst * (2 + cv);
    r[0] = g0;
    r[1] = g1;
    const float4* c = reinterpret_cast<const float4*>(chan) + (size_t)id * cv;
    for (int q = 0; q < cv; ++q) r[2 + q] = __ldg(c + q);
}

__global__ void __launch_bounds__(kSortThreads) k_tile_sort(uint64_t* __restrict__ keys, uint64_t* __restrict__ tmp,
                                                            const uint32_t* __restrict__ offsets, int smem_keys,
                                                            long long capacity, const float* __restrict__ geom_all,
                                                            const float* __restrict__ chan_all, float* __restrict__ cull,
                                                            float* __restrict__ rec, int G, int tiles_per_view, int cv) {
    extern __shared__ __align__(16) uint64_t sbuf[];
    __shared__ uint32_t hist[8][256];
    __shared__ uint32_t base[256];
    __shared__ uint32_t wcnt[kSortWarps][256];
    __shared__ uint32_t wsum[kSortWarps];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long s = offsets[blockIdx.x];
    long long e = offsets[blockIdx.x + 1];
    if (e > capacity) e = capacity;
    const int n = (int)(e - s);
    if (n <= 0) return;
    const int view = blockIdx.x / tiles_per_view;
    const float* __restrict__ geom = geom_all + (size_t)view * G * LS_GEOM_STRIDE;
    const float* __restrict__ chan = chan_all + (size_t)view * G * (4 * cv);
    if (n == 1) {                                     // nothing to sort, but the queue entry must exist
        if (tid == 0) write_entry(keys[s], s, keys, geom, chan, cull, rec, cv, false);
        return;
    }

    const bool in_smem = n <= smem_keys;
    uint64_t* a = in_smem ? sbuf : keys + s;
    uint64_t* b = in_smem ? sbuf + smem_keys : tmp + s;

    for (int d = 0; d < 8; ++d) hist[d][tid] = 0;
    for (int ww = 0; ww < kSortWarps; ++ww) wcnt[ww][tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kSortThreads) {
        const uint64_t k = keys[s + i];
        if (in_smem) a[i] = k;
#pragma unroll
        for (int d = 0; d < 8; ++d) atomicAdd(&hist[d][(uint32_t)(k >> (8 * d)) & 255u], 1u);
    }
    __syncthreads();

    for (int d = 0; d < 8; ++d) {
        const uint32_t c = hist[d][tid];
        if (__syncthreads_or(c == (uint32_t)n)) continue;  // all keys share this digit
        // exclusive scan of the 256 bins
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t pre = 0;
#pragma unroll
        for (int ww = 0; ww < kSortWarps; ++ww) pre += (ww < w) ? wsum[ww] : 0u;
        base[tid] = x - c + pre;
        __syncthreads();

        const int shift = 8 * d;
        for (int c0 = 0; c0 < n; c0 += kSortThreads) {
            const int i = c0 + tid;
            const bool valid = i < n;
            const uint64_t k = valid ? a[i] : 0ull;
            const uint32_t dg = valid ? ((uint32_t)(k >> shift) & 255u) : 0xffffffffu;
            const uint32_t m = __match_any_sync(0xffffffffu, dg);
            const uint32_t lt = m & ((1u << lane) - 1u);
            const uint32_t r = __popc(lt), cnt = __popc(m);
            const bool leader = valid && lt == 0u;
            if (leader) wcnt[w][dg] = cnt;
            __syncthreads();
            uint32_t before = 0, tot = 0;
            if (valid) {
#pragma unroll
                for (int ww = 0; ww < kSortWarps; ++ww) {
                    const uint32_t cc = wcnt[ww][dg];
                    tot += cc;
                    before += (ww < w) ? cc : 0u;
                }
                b[base[dg] + before + r] = k;
            }
            __syncthreads();
            if (leader) {
                wcnt[w][dg] = 0;
                if (before + cnt == tot) base[dg] += tot;  // highest warp holding this digit
            }
            __syncwarp();  // next iteration's first barrier orders these updates for the other warps
        }
        __syncthreads();
        uint64_t* t = a; a = b; b = t;
    }
    // write-out: sorted keys (unless they already sit in place) + the permuted queue records
    const bool write_key = a != keys + s;
    for (int i = tid; i < n; i += kSortThreads) write_entry(a[i], s + i, keys, geom, chan, cull, rec, cv, write_key);
}

// =========================================================================================
// R.6 blend: one CTA per (view, tile), one pixel per thread, front to back.
// NC = blended value channels (colour then features); depth and alpha ride along.
// =========================================================================================
template <int NC>
__global__ void __launch_bounds__(kTilePixels) k_blend_fwd(const LsRasterScene sc, const LsRasterState st,
                                                           const LsRasterImages im, const int ncol) {
    constexpr int CS = (NC + 3) & ~3;  // chan_stride
    constexpr int CV = CS / 4;
    constexpr int RV = 2 + CV;         // float4 per queue record
    extern __shared__ __align__(128) uint8_t s_dyn[];
    float4* s_rec = reinterpret_cast<float4*>(s_dyn);                                   // [2][kBatch][RV]
    float4* s_cull = s_rec + 2 * kBatch * RV;                                           // [2][kBatch]
    uint8_t* s_list = reinterpret_cast<uint8_t*>(s_cull + 2 * kBatch);                  // [8 warps][kBatch]
    __shared__ __align__(8) uint64_t s_bar[2];

    const int gx = (sc.W + kTile - 1) / kTile, gy = (sc.H + kTile - 1) / kTile;
    const int tile = blockIdx.x, v = blockIdx.y;
    const int tx = tile % gx, ty = tile / gx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = tx * kTile + (tid & 15), py = ty * kTile + (tid >> 4);
    const bool inside = px < sc.W && py < sc.H;
    const float fxp = (float)px, fyp = (float)py;

    const size_t t = (size_t)v * gx * gy + tile;
    const long long s = st.tile_offsets[t];
    long long e = st.tile_offsets[t + 1];
    if (e > st.capacity) e = st.capacity;
    const int n = (int)(e > s ? e - s : 0);
    const int nb = (n + kBatch - 1) / kBatch;

    const uint32_t bar0 = smem_u32(&s_bar[0]);
    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_fence_init(); }
    __syncthreads();
    // one elected thread stages batch b of the tile's queue: two 1-D bulk copies signalling the buffer's mbarrier
    auto stage = [&](int b) {
        const int buf = b & 1, cnt = min(kBatch, n - b * kBatch);
        const uint32_t bar = bar0 + 8 * buf;
        const long long first = s + (long long)b * kBatch;
        mbar_expect_tx(bar, (uint32_t)cnt * (16u + 16u * RV));
        bulk_load(smem_u32(s_rec + buf * kBatch * RV), reinterpret_cast<const float4*>(st.sorted_rec) + first * RV, (uint32_t)cnt * 16u * RV, bar);
        bulk_load(smem_u32(s_cull + buf * kBatch), reinterpret_cast<const float4*>(st.sorted_cull) + first, (uint32_t)cnt * 16u, bar);
    };

    const float wx = (float)(tx * kTile) + 7.5f;                       // centre of the warp's 16x2 strip
    const float wy = (float)(ty * kTile + 2 * warp) + 0.5f;
    float T = 1.0f;
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.f;
    float acc_d = 0.f, acc_a = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    uint8_t* my_list = s_list + warp * kBatch;

    int staged = 0;                                                    // batches whose copy has been issued
    if (tid == 0 && nb > 0) { stage(0); }
    staged = nb > 0 ? 1 : 0;
    int b = 0;
    for (; b < nb; ++b) {
        const int buf = b & 1;
        if (b + 1 < nb) {                                              // buffer (b+1)&1 was released by the barrier that ended batch b-1
            if (tid == 0) stage(b + 1);
            staged = b + 2;
        }
        mbar_wait(bar0 + 8 * buf, (uint32_t)(b >> 1) & 1u);
        if (__syncthreads_count(done) == kTilePixels) { ++b; break; }
        const int cnt = min(kBatch, n - b * kBatch);
        const float4* cull = s_cull + buf * kBatch;
        const float4* rec = s_rec + buf * kBatch * RV;
        // ---- phase A: compaction.  32 different queue entries per instruction against the strip's extent
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
        // ---- phase B: alpha math over the survivors, in list order
        for (int i = 0; i < nsurv; ++i) {
            const int j = my_list[i];
            const float4 g0 = rec[j * RV], g1 = rec[j * RV + 1];
            const float dx = g0.x - fxp, dy = g0.y - fyp;
            const float p2 = fmaf(g0.z * dx, dx, fmaf(g1.x * dy, dy, g0.w * dx * dy));  // log2 domain, <= 0
            const float alpha = fminf(kAlphaMax, g1.y * ex2_approx(p2));
            const bool hit = !done && p2 <= 0.f && alpha >= kAlphaMin;
            if (!__any_sync(0xffffffffu, hit)) continue;
            if (hit) {
                const float test_T = T * (1.f - alpha);
                if (test_T < kTMin) {
                    done = true;
                } else {
                    const float wgt = alpha * T;
#pragma unroll
                    for (int q = 0; q < CV; ++q) {
                        const float4 cq = rec[j * RV + 2 + q];
                        if (4 * q + 0 < NC) acc[4 * q + 0] = fmaf(cq.x, wgt, acc[4 * q + 0]);
                        if (4 * q + 1 < NC) acc[4 * q + 1] = fmaf(cq.y, wgt, acc[4 * q + 1]);
                        if (4 * q + 2 < NC) acc[4 * q + 2] = fmaf(cq.z, wgt, acc[4 * q + 2]);
                        if (4 * q + 3 < NC) acc[4 * q + 3] = fmaf(cq.w, wgt, acc[4 * q + 3]);
                    }
                    acc_d = fmaf(g1.z, wgt, acc_d);
                    acc_a += wgt;
                    T = test_T;
                    last = (uint32_t)(b * kBatch + j + 1);
                }
            }
        }
        __syncthreads();  // every warp is done with buffer `buf`: batch b+2 may be staged into it
    }
    // a copy issued but not consumed (early exit) must land before the CTA gives its shared memory back
    if (b < staged) mbar_wait(bar0 + 8 * (b & 1), (uint32_t)(b >> 1) & 1u);

    if (inside) {
        const size_t hw = (size_t)sc.H * sc.W;
        const size_t pid = (size_t)py * sc.W + px;
        st.final_T[(size_t)v * hw + pid] = T;
        st.n_contrib[(size_t)v * hw + pid] = last;
        if (ncol) {
            const float* bg = sc.bg + 3 * v;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (c < NC) im.color[((size_t)v * 3 + c) * hw + pid] = fmaf(T, bg[c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c >= ncol) im.feature[((size_t)v * (NC - ncol) + (c - ncol)) * hw + pid] = acc[c];
        im.alpha[(size_t)v * hw + pid] = acc_a;
        im.depth[(size_t)v * hw + pid] = acc_d;
    }
}

template <int NC>
static int launch_blend_fwd(const LsRasterScene& sc, const LsRasterState& st, const LsRasterImages& im, int ncol,
                            dim3 grid, cudaStream_t stream) {
    constexpr int RV = 2 + ((NC + 3) & ~3) / 4;
    constexpr int smem = blend_smem_bytes(RV);
    static lstc::PerDeviceOnce once;
    if (smem > 48 * 1024 && once.ensure_smem(k_blend_fwd<NC>, smem) != cudaSuccess) return ls_check_cuda("blend smem attribute");
    k_blend_fwd<NC><<<grid, kTilePixels, smem, stream>>>(sc, st, im, ncol);
    return 0;
}

}  // namespace ls

using namespace ls;

// =========================================================================================
// host side
// =========================================================================================
namespace {
thread_local char g_err[512] = "";
}
int ls_fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}
int ls_check_cuda(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ls_fail("%s: %s", what, cudaGetErrorString(e));
    return 0;
}

extern "C" const char* ls_last_error(void) { return g_err; }
extern "C" int ls_raster_abi_version(void) { return LS_RASTER_ABI_VERSION; }

int ls_validate_scene(const LsRasterScene* sc) {
    if (!sc) return ls_fail("scene is NULL");
    if (sc->n_views <= 0 || sc->G < 0 || sc->H <= 0 || sc->W <= 0) return ls_fail("bad sizes V=%d G=%d H=%d W=%d", sc->n_views, sc->G, sc->H, sc->W);
    if (sc->views_per_scene <= 0 || sc->n_views % sc->views_per_scene) return ls_fail("n_views %d not a multiple of views_per_scene %d", sc->n_views, sc->views_per_scene);
    if (sc->color_mode < 0 || sc->color_mode > LS_COLOR_SH_3DGS) return ls_fail("bad color_mode %d", sc->color_mode);
    if (sc->feature_mode < 0 || sc->feature_mode > 2) return ls_fail("bad feature_mode %d", sc->feature_mode);
    if (sc->C < 0 || (sc->feature_mode == LS_FEATURE_NONE) != (sc->C == 0)) return ls_fail("feature_mode %d inconsistent with C=%d", sc->feature_mode, sc->C);
    if (sc->sh_degree < 0 || sc->sh_degree > 4 || sc->feature_sh_degree < 0 || sc->feature_sh_degree > 4) return ls_fail("SH degree out of range (0..4)");
    const int nc = n_color(sc->color_mode) + sc->C;
    if (nc < 1) return ls_fail("nothing to render: no colour and no features (cuda_splatting.py:71)");
    if (nc > LS_MAX_VALUE_CHANNELS) return ls_fail("colour+feature channels %d exceed LS_MAX_VALUE_CHANNELS=%d", nc, LS_MAX_VALUE_CHANNELS);
    if (!sc->viewmatrix || !sc->projmatrix || !sc->campos || !sc->tanfov) return ls_fail("a required camera pointer is NULL");
    if (sc->color_mode != LS_COLOR_NONE && !sc->bg) return ls_fail("bg pointer is NULL");
    if (sc->G > 0) {  // zero-sized arrays legitimately have NULL data pointers
        if (!sc->means3D || !sc->cov3D || !sc->opacity) return ls_fail("a required scene pointer is NULL");
        if (sc->color_mode != LS_COLOR_NONE && !sc->color) return ls_fail("color pointer is NULL");
        if (sc->feature_mode != LS_FEATURE_NONE && !sc->feature) return ls_fail("feature pointer is NULL");
    }
    return 0;
}

extern "C" int ls_raster_sizes(const LsRasterScene* sc, LsRasterSizes* out) {
    if (!out) return ls_fail("sizes out is NULL");
    if (!sc || sc->n_views <= 0 || sc->views_per_scene <= 0) return ls_fail("bad scene");
    const int ncol = n_color(sc->color_mode);
    const int nc = ncol + sc->C;
    const int64_t gx = (sc->W + kTile - 1) / kTile, gy = (sc->H + kTile - 1) / kTile;
    out->n_scenes = sc->n_views / sc->views_per_scene;
    out->tiles_per_view = gx * gy;
    out->chan_stride = round_up4(nc < 1 ? 1 : nc);
    out->grad_stride = round_up4(7 + nc);
    out->n_color = ncol;
    out->n_value_channels = nc;
    out->per_view_gaussian = (int64_t)sc->n_views * sc->G;
    out->geom = out->per_view_gaussian * LS_GEOM_STRIDE;
    out->chan = out->per_view_gaussian * out->chan_stride;
    out->tile_slots = (int64_t)sc->n_views * gx * gy;
    out->pixels = (int64_t)sc->n_views * sc->H * sc->W;
    out->grad_record = out->per_view_gaussian * out->grad_stride;
    out->rec_stride = LS_GEOM_STRIDE + out->chan_stride;
    out->reserved0 = 0;
    return 0;
}

extern "C" int ls_raster_forward(const LsRasterScene* sc, const LsRasterState* st, const LsRasterImages* im,
                                 int32_t stages, void* stream_) {
    if (ls_validate_scene(sc)) return -1;
    if (!st) return ls_fail("state is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ncol = n_color(sc->color_mode);
    const int nc = ncol + sc->C;
    if (st->chan_stride != round_up4(nc)) return ls_fail("chan_stride %d != %d", st->chan_stride, round_up4(nc));
    const int gx = (sc->W + kTile - 1) / kTile, gy = (sc->H + kTile - 1) / kTile;
    const int n_slots = sc->n_views * gx * gy;
    if (!st->tile_count || !st->tile_offsets || !st->stats) return ls_fail("a required state pointer is NULL");
    if (sc->G > 0 && (!st->geom || !st->chan || !st->radii || !st->tiles_touched || !st->clamped))
        return ls_fail("a required per-Gaussian state pointer is NULL");

    if (stages & LS_STAGE_GEOMETRY) {
        cudaMemsetAsync(st->tile_count, 0, sizeof(uint32_t) * (size_t)n_slots, stream);
        if (sc->G > 0) {
            const bool aligned = ((reinterpret_cast<uintptr_t>(sc->color) | reinterpret_cast<uintptr_t>(sc->feature)) & 15) == 0;
            if (ls_raster_dense_sh_grads(sc) && aligned) {         // the specialised configuration (also used by the backward)
                constexpr int kWarps = 4;
                dim3 gridf((sc->G + 32 * kWarps - 1) / (32 * kWarps), sc->n_views);
                const int smem = kWarps * 32 * (75 + sc->C * 9) * 4 + kWarps * 8;
                static lstc::PerDeviceOnce once4, once8;
                if (sc->C == 4) {
                    if (smem > 48 * 1024 && once4.ensure_smem(k_preprocess<4>, smem) != cudaSuccess) return ls_check_cuda("preprocess smem attribute");
                    k_preprocess<4><<<gridf, 32 * kWarps, smem, stream>>>(*sc, *st);
                } else {
                    if (smem > 48 * 1024 && once8.ensure_smem(k_preprocess<8>, smem) != cudaSuccess) return ls_check_cuda("preprocess smem attribute");
                    k_preprocess<8><<<gridf, 32 * kWarps, smem, stream>>>(*sc, *st);
                }
            } else {
                dim3 grid((sc->G + 255) / 256, sc->n_views);
                k_preprocess<0><<<grid, 256, 0, stream>>>(*sc, *st);
            }
        }
        // geometry launched on its own = exact sizing: the caller reads stats[0] and allocates, nothing can overflow
        const long long cap_check = (stages & LS_STAGE_RENDER) ? (long long)st->capacity : -1LL;
        k_scan_tiles<<<1, 1024, 0, stream>>>(st->tile_count, st->tile_offsets, st->stats, n_slots, cap_check);
        if (ls_check_cuda("geometry stage")) return -1;
    }
    if ((stages & LS_STAGE_RENDER) && st->capacity > 0 && (!st->keys || !st->keys_tmp))
        return ls_fail("keys/keys_tmp is NULL with capacity %lld", (long long)st->capacity);
    if ((stages & LS_STAGE_SCATTER) && sc->G > 0 && st->capacity > 0) {
        dim3 grid((sc->G + 255) / 256, sc->n_views);
        k_scatter_keys<<<grid, 256, 0, stream>>>(*sc, *st);
        if (ls_check_cuda("scatter stage")) return -1;
    }
    if ((stages & LS_STAGE_SORT) && sc->G > 0 && st->capacity > 0) {
        int smem_keys = st->sort_smem_keys > 0 ? st->sort_smem_keys : 4096;
        if (smem_keys > 12288) smem_keys = 12288;
        const size_t smem_bytes = (size_t)smem_keys * 2 * sizeof(uint64_t);
        // the attribute is per device and the request may grow: setting it is idempotent and cheap, so no cache
        if (smem_bytes > 48 * 1024 &&
            cudaFuncSetAttribute(k_tile_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess)
            return ls_check_cuda("sort smem attribute");
        if (!st->sorted_cull || !st->sorted_rec) return ls_fail("sorted_cull / sorted_rec is NULL");
        if (st->rec_stride != LS_GEOM_STRIDE + st->chan_stride) return ls_fail("rec_stride %d != %d", st->rec_stride, LS_GEOM_STRIDE + st->chan_stride);
        k_tile_sort<<<n_slots, kSortThreads, smem_bytes, stream>>>(st->keys, st->keys_tmp, st->tile_offsets, smem_keys,
                                                                    (long long)st->capacity, st->geom, st->chan, st->sorted_cull,
                                                                    st->sorted_rec, sc->G, gx * gy, st->chan_stride / 4);
        if (ls_check_cuda("sort stage")) return -1;
    }
    if (stages & LS_STAGE_BLEND) {
        if (!im || !im->alpha || !im->depth || !st->final_T || !st->n_contrib) return ls_fail("image/state output pointer is NULL");
        if (ncol && !im->color) return ls_fail("color image pointer is NULL");
        if (sc->C && !im->feature) return ls_fail("feature image pointer is NULL");
        dim3 grid(gx * gy, sc->n_views);
        switch (nc) {
#define LS_CASE(N) case N: if (launch_blend_fwd<N>(*sc, *st, *im, ncol, grid, stream)) return -1; break;
            LS_CASE(1) LS_CASE(2) LS_CASE(3) LS_CASE(4) LS_CASE(5) LS_CASE(6) LS_CASE(7) LS_CASE(8)
            LS_CASE(9) LS_CASE(10) LS_CASE(11) LS_CASE(12) LS_CASE(13) LS_CASE(14) LS_CASE(15) LS_CASE(16)
#undef LS_CASE
            default: return ls_fail("unsupported channel count %d", nc);
        }
        if (ls_check_cuda("blend stage")) return -1;
    }
    return 0;
}
