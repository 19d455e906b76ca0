// ls_ghead.cu -- the encoder's per-ray tail as one kernel each way: depth-bucket sampling + Gaussian adapter.
//
// What it replaces in the reference (Chrixtar/latentsplat), per context-view pixel ("ray"):
//   DepthPredictorMonocular.forward   src/model/encoder/epipolar/depth_predictor_monocular.py:37-81   softmax over 32 depth
//       buckets, sigmoid offsets, sample_discrete_distribution (cumsum + rand + searchsorted, src/misc/
//       discrete_probability_distribution.py:7-20) or top-1, relative_disparity_to_depth (conversions.py:5-14)
//   EncoderEpipolar.forward :176-242  xy offset sigmoid, map_pdf_to_opacity (:113-126) / gaussians_per_pixel
//   GaussianAdapter.forward           src/model/encoder/common/gaussian_adapter.py:63-114   scale from depth and the pixel
//       footprint, quaternion -> rotation, covariance C R S S^T R^T C^T (gaussians.py:8-44), mean = origin + direction*depth
//       (get_world_rays, src/geometry/projection.py:98-121), broadcast of the SH coefficient rows over the samples of a ray
// -- in torch ~60 elementwise / reduction kernels forward and ~100 backward over (rays x samples) tensors plus three
// materialising copies of the SH rows.  Here: one warp per ray, lane l owns depth bucket l; 888 B read and 1.9 KB written
// per ray forward (HBM-bound, coalesced SH row copies), the mirror image backward.  The uniform samples `u` come from
// torch.rand (same shape and order as the reference's draw, so seeded runs consume the same stream).
//
// raw row layout (gaussian_adapter.py:138-139 preceded by the 2 xy offsets of encoder_epipolar.py:183):
//   [ox oy | s0 s1 s2 | qx qy qz qw | colour SH (3 x 25) | feature SH (C x 9) ], SH blocks already masked and rotated to world
//   space (the encoder folds both linear maps into the to_gaussians weights).
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_ghead.h"
#include "ls_host.h"

namespace lsh {

constexpr int kMaxSamples = 4;
constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }

struct Camera {          // per view, rebuilt by every warp from 28 floats (L1-resident)
    float C[3][3];       // camera-to-world rotation
    float o[3];          // camera centre
    float Ki[3][3];      // inverse intrinsics (closed-form adjugate: no cuSOLVER round trip for a 3x3)
    float dn, df;        // 1/(near+eps), 1/(far+eps)
    float mult;          // get_scale_multiplier: 0.1 * sum((K[:2,:2])^-1 (1/w, 1/h))
};

__device__ __forceinline__ Camera load_camera(const LsGaussianHead& a, int view) {
    Camera c;
    const float* E = a.extrinsics + 16 * view;
    const float* K = a.intrinsics + 9 * view;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) c.C[i][j] = E[4 * i + j];
        c.o[i] = E[4 * i + 3];
    }
    const float k00 = K[0], k01 = K[1], k02 = K[2], k10 = K[3], k11 = K[4], k12 = K[5], k20 = K[6], k21 = K[7], k22 = K[8];
    const float c00 = k11 * k22 - k12 * k21, c01 = k02 * k21 - k01 * k22, c02 = k01 * k12 - k02 * k11;
    const float c10 = k12 * k20 - k10 * k22, c11 = k00 * k22 - k02 * k20, c12 = k02 * k10 - k00 * k12;
    const float c20 = k10 * k21 - k11 * k20, c21 = k01 * k20 - k00 * k21, c22 = k00 * k11 - k01 * k10;
    const float inv_det = 1.f / (k00 * c00 + k01 * c10 + k02 * c20);
    c.Ki[0][0] = c00 * inv_det; c.Ki[0][1] = c01 * inv_det; c.Ki[0][2] = c02 * inv_det;
    c.Ki[1][0] = c10 * inv_det; c.Ki[1][1] = c11 * inv_det; c.Ki[1][2] = c12 * inv_det;
    c.Ki[2][0] = c20 * inv_det; c.Ki[2][1] = c21 * inv_det; c.Ki[2][2] = c22 * inv_det;
    const float eps = 1e-10f;
    c.dn = 1.f / (a.near[view] + eps);
    c.df = 1.f / (a.far[view] + eps);
    const float d2 = 1.f / (k00 * k11 - k01 * k10);                 // inverse of the upper-left 2x2
    const float pw = 1.f / (float)a.width, ph = 1.f / (float)a.height;
    c.mult = 0.1f * ((k11 * pw - k01 * ph) * d2 + (-k10 * pw + k00 * ph) * d2);
    return c;
}

// quaternion (x, y, z, w) normalised with eps as the adapter does, then gaussians.py:quaternion_to_matrix (its own 2/(q.q+eps))
__device__ __forceinline__ void quat_to_rot(const float q[4], float qn[4], float& two_s, float R[3][3]) {
    const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float inv = 1.f / (n + 1e-8f);
#pragma unroll
    for (int i = 0; i < 4; ++i) qn[i] = q[i] * inv;
    const float i_ = qn[0], j = qn[1], k = qn[2], r = qn[3];
    two_s = 2.f / (i_ * i_ + j * j + k * k + r * r + 1e-8f);
    R[0][0] = 1.f - two_s * (j * j + k * k); R[0][1] = two_s * (i_ * j - k * r); R[0][2] = two_s * (i_ * k + j * r);
    R[1][0] = two_s * (i_ * j + k * r); R[1][1] = 1.f - two_s * (i_ * i_ + k * k); R[1][2] = two_s * (j * k - i_ * r);
    R[2][0] = two_s * (i_ * k - j * r); R[2][1] = two_s * (j * k + i_ * r); R[2][2] = 1.f - two_s * (i_ * i_ + j * j);
}

__device__ __forceinline__ float opacity_map(float p, float e) {     // encoder_epipolar.py:113-126
    if (e == 1.f) return p;
    return 0.5f * (1.f - __powf(1.f - p, e) + __powf(p, 1.f / e));
}
__device__ __forceinline__ float opacity_map_grad(float p, float e) {
    if (e == 1.f) return 1.f;
    return 0.5f * (e * __powf(1.f - p, e - 1.f) + (1.f / e) * __powf(p, 1.f / e - 1.f));
}

// everything of a ray that does not depend on the sample: shared by forward and backward
struct RayState {
    float pdf, S, D;             // this lane's softmax probability, its warp sum, eps + S
    float off;                   // this lane's sigmoid offset
    float sx, sy;                // sigmoid of the xy offsets
    float x, y;                  // image coordinates of the ray
    float u[3], un, dc[3], dw[3];// K^-1 (x, y, 1), its norm, unit camera / world direction
    float sg[3], sb[3];          // sigmoid of the raw scales, scale_min + (max - min) * sigmoid
    float qn[4], two_s, R[3][3];
    float CR[3][3];              // C * R
};

__device__ __forceinline__ void ray_state(const LsGaussianHead& a, const Camera& cam, const float* __restrict__ dl,
                                          const float* __restrict__ raw, int pixel, int lane, RayState& s) {
    const float2 lg = *reinterpret_cast<const float2*>(dl + 2 * lane);       // (pdf logit, offset logit) of bucket `lane`
    const float m = warp_max(lg.x);
    const float ex = __expf(lg.x - m);
    const float sum = warp_sum(ex);
    s.pdf = ex / sum;
    s.S = warp_sum(s.pdf);
    s.D = 1.1920929e-07f + s.S;                                              // torch.finfo(float32).eps + sum
    s.off = sigmoidf(lg.y);
    s.sx = sigmoidf(raw[0]);
    s.sy = sigmoidf(raw[1]);
    const int px = pixel % a.width, py = pixel / a.width;
    s.x = ((float)px + 0.5f) / (float)a.width + (s.sx - 0.5f) / (float)a.width;
    s.y = ((float)py + 0.5f) / (float)a.height + (s.sy - 0.5f) / (float)a.height;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.u[i] = cam.Ki[i][0] * s.x + cam.Ki[i][1] * s.y + cam.Ki[i][2];
    s.un = sqrtf(s.u[0] * s.u[0] + s.u[1] * s.u[1] + s.u[2] * s.u[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) s.dc[i] = s.u[i] / s.un;
#pragma unroll
    for (int i = 0; i < 3; ++i) s.dw[i] = cam.C[i][0] * s.dc[0] + cam.C[i][1] * s.dc[1] + cam.C[i][2] * s.dc[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        s.sg[i] = sigmoidf(raw[2 + i]);
        s.sb[i] = a.scale_min + (a.scale_max - a.scale_min) * s.sg[i];
    }
    const float q[4] = {raw[5], raw[6], raw[7], raw[8]};
    quat_to_rot(q, s.qn, s.two_s, s.R);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) s.CR[i][j] = cam.C[i][0] * s.R[0][j] + cam.C[i][1] * s.R[1][j] + cam.C[i][2] * s.R[2][j];
}

// ---- forward -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_ghead_fwd(const LsGaussianHead a, const LsGaussianHeadOut o) {
    const int lane = threadIdx.x & 31;
    const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.rays) return;                                               // warp-uniform
    const int view = (int)(ray / a.rays_per_view), pixel = (int)(ray - (long long)view * a.rays_per_view);
    const int row = 9 + a.d_color + a.d_feature;
    const float* __restrict__ dl = a.dlog + ray * 64;
    const float* __restrict__ raw = a.raw + ray * row;
    const Camera cam = load_camera(a, view);
    RayState s;
    ray_state(a, cam, dl, raw, pixel, lane, s);

    // inclusive prefix sum of the normalised pdf over the 32 buckets (lane = bucket)
    const float pn = s.pdf / s.D;
    float cdf = pn;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, cdf, d);
        if (lane >= d) cdf += t;
    }
    const float pmax = warp_max(s.pdf);
    const uint32_t argmax_mask = __ballot_sync(0xffffffffu, s.pdf == pmax);

    for (int k = 0; k < a.samples; ++k) {
        int idx;
        if (a.deterministic) {
            idx = __ffs(argmax_mask) - 1;                                    // top-1 (lowest index among equal maxima)
        } else {
            const float uk = a.u[ray * a.samples + k];
            idx = __popc(__ballot_sync(0xffffffffu, cdf <= uk));             // searchsorted(right=True)
            idx = idx > 31 ? 31 : idx;                                       // .clip(max = buckets - 1)
        }
        const float p_i = __shfl_sync(0xffffffffu, pn, idx);
        const float off_i = __shfl_sync(0xffffffffu, s.off, idx);
        const float rd = ((float)idx + off_i) * (1.f / 32.f);
        const float depth = 1.f / ((1.f - rd) * (cam.dn - cam.df) + cam.df + 1e-10f);
        const long long g = ray * a.samples + k;
        if (lane == 0) {
            o.index[g] = idx;
            o.opacity[g] = opacity_map(p_i, a.opacity_exponent) * a.inv_gpp;
        }
        // the 3 + 9 small outputs: every lane computes them (registers only), lane e stores entry e
        float sc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) sc[c] = s.sb[c] * depth * cam.mult;
        float v12[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) v12[i] = cam.o[i] + s.dw[i] * depth;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {                                    // covariance M M^T, M = C R diag(scale)
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) acc = fmaf(s.CR[i][c] * sc[c], s.CR[j][c] * sc[c], acc);
                v12[3 + 3 * i + j] = acc;
            }
        float sel = v12[0];
#pragma unroll
        for (int e = 1; e < 12; ++e) sel = lane == e ? v12[e] : sel;
        if (lane < 3) o.means[3 * g + lane] = sel;
        else if (lane < 12) o.covariances[9 * g + lane - 3] = sel;
        float* __restrict__ oc = o.color_sh + g * a.d_color;
        for (int e = lane; e < a.d_color; e += 32) oc[e] = raw[9 + e];
        float* __restrict__ of = o.feature_sh + g * a.d_feature;
        for (int e = lane; e < a.d_feature; e += 32) of[e] = raw[9 + a.d_color + e];
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_ghead_bwd(const LsGaussianHead a, const LsGaussianHeadGrad g) {
    const int lane = threadIdx.x & 31;
    const long long ray = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray >= a.rays) return;
    const int view = (int)(ray / a.rays_per_view), pixel = (int)(ray - (long long)view * a.rays_per_view);
    const int row = 9 + a.d_color + a.d_feature;
    const float* __restrict__ dl = a.dlog + ray * 64;
    const float* __restrict__ raw = a.raw + ray * row;
    const Camera cam = load_camera(a, view);
    RayState s;
    ray_state(a, cam, dl, raw, pixel, lane, s);
    const float A = cam.dn - cam.df;

    float g_sb[3] = {0.f, 0.f, 0.f};            // d loss / d (scale_min + range * sigmoid), summed over the samples
    float g_R[3][3] = {};                       // d loss / d R
    float g_dw[3] = {0.f, 0.f, 0.f};            // d loss / d world direction
    float g_pdf = 0.f, g_off = 0.f;             // this lane's bucket: d loss / d pdf, d loss / d offset (pre-sigmoid below)
    float gp_sum = 0.f;                         // sum_k gp_k * pdf[idx_k] / D^2

    for (int k = 0; k < a.samples; ++k) {
        const long long gi = ray * a.samples + k;
        const int idx = g.index[gi];
        const float pn_i = __shfl_sync(0xffffffffu, s.pdf, idx) / s.D;
        const float off_i = __shfl_sync(0xffffffffu, s.off, idx);
        const float rd = ((float)idx + off_i) * (1.f / 32.f);
        const float depth = 1.f / ((1.f - rd) * A + cam.df + 1e-10f);
        // upstream gradients of this sample (broadcast loads: every lane reads the same 13 floats)
        float gm[3], gc[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) gm[i] = g.d_means[3 * gi + i];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) gc[i][j] = g.d_covariances[9 * gi + 3 * i + j];
        const float go = g.d_opacity[gi];
        // covariance = M M^T, M = CR diag(sc):  dM = (G + G^T) M;  d sc_c = sum_i dM[i][c] CR[i][c];  d CR[i][c] = dM[i][c] sc_c
        float sc[3], g_depth = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) sc[c] = s.sb[c] * depth * cam.mult;
        float dCR[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float g_sc = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float dM = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) dM = fmaf(gc[i][j] + gc[j][i], s.CR[j][c] * sc[c], dM);
                g_sc = fmaf(dM, s.CR[i][c], g_sc);
                dCR[i][c] = dM * sc[c];
            }
            g_sb[c] = fmaf(g_sc, depth * cam.mult, g_sb[c]);
            g_depth = fmaf(g_sc, s.sb[c] * cam.mult, g_depth);
        }
        // CR = C R  ->  dR = C^T dCR
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                g_R[i][j] += cam.C[0][i] * dCR[0][j] + cam.C[1][i] * dCR[1][j] + cam.C[2][i] * dCR[2][j];
        // mean = origin + dw * depth
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g_depth = fmaf(gm[i], s.dw[i], g_depth);
            g_dw[i] = fmaf(gm[i], depth, g_dw[i]);
        }
        // depth -> relative disparity -> sigmoid offset of bucket idx
        const float g_rd = g_depth * depth * depth * A;
        if (lane == idx) g_off += g_rd * (1.f / 32.f);
        // opacity = map(pdf_norm[idx]) / gpp; pdf_norm = pdf / (eps + sum pdf)
        const float gp = go * a.inv_gpp * opacity_map_grad(pn_i, a.opacity_exponent);
        if (lane == idx) g_pdf += gp / s.D;
        gp_sum = fmaf(gp, pn_i / s.D, gp_sum);
    }
    // softmax backward: d pdf_j = g_pdf_j - gp_sum;  d logit_j = pdf_j (d pdf_j - sum_i pdf_i d pdf_i)
    const float dpdf = g_pdf - gp_sum;
    const float dot = warp_sum(s.pdf * dpdf);
    const float d_logit = s.pdf * (dpdf - dot);
    const float d_offlogit = g_off * s.off * (1.f - s.off);
    *reinterpret_cast<float2*>(g.d_dlog + ray * 64 + 2 * lane) = make_float2(d_logit, d_offlogit);

    // the nine leading entries of the raw row (identical in every lane; lane e writes entry e)
    float out9[9];
    {
        // world direction -> camera direction -> K^-1 (x, y, 1) -> (x, y) -> xy offsets
        float g_dc[3], g_u[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) g_dc[i] = cam.C[0][i] * g_dw[0] + cam.C[1][i] * g_dw[1] + cam.C[2][i] * g_dw[2];
        const float dd = s.dc[0] * g_dc[0] + s.dc[1] * g_dc[1] + s.dc[2] * g_dc[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) g_u[i] = (g_dc[i] - s.dc[i] * dd) / s.un;
        const float g_x = g_u[0] * cam.Ki[0][0] + g_u[1] * cam.Ki[1][0] + g_u[2] * cam.Ki[2][0];
        const float g_y = g_u[0] * cam.Ki[0][1] + g_u[1] * cam.Ki[1][1] + g_u[2] * cam.Ki[2][1];
        out9[0] = g_x / (float)a.width * s.sx * (1.f - s.sx);
        out9[1] = g_y / (float)a.height * s.sy * (1.f - s.sy);
#pragma unroll
        for (int c = 0; c < 3; ++c) out9[2 + c] = g_sb[c] * (a.scale_max - a.scale_min) * s.sg[c] * (1.f - s.sg[c]);
        // rotation matrix -> normalised quaternion (with the 2/(q.q+eps) factor) -> raw quaternion
        const float i_ = s.qn[0], j = s.qn[1], k = s.qn[2], r = s.qn[3], t = s.two_s;
        // R = I + t * P(q) with P the quadratic forms below; d two_s via P, d q via t * dP
        const float P[3][3] = {{-(j * j + k * k), i_ * j - k * r, i_ * k + j * r},
                               {i_ * j + k * r, -(i_ * i_ + k * k), j * k - i_ * r},
                               {i_ * k - j * r, j * k + i_ * r, -(i_ * i_ + j * j)}};
        float g_t = 0.f;
#pragma unroll
        for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y) g_t = fmaf(g_R[x][y], P[x][y], g_t);
        float gq[4];
        gq[0] = t * (g_R[0][1] * j + g_R[0][2] * k + g_R[1][0] * j - 2.f * g_R[1][1] * i_ - g_R[1][2] * r + g_R[2][0] * k + g_R[2][1] * r - 2.f * g_R[2][2] * i_);
        gq[1] = t * (-2.f * g_R[0][0] * j + g_R[0][1] * i_ + g_R[0][2] * r + g_R[1][0] * i_ + g_R[1][2] * k - g_R[2][0] * r + g_R[2][1] * k - 2.f * g_R[2][2] * j);
        gq[2] = t * (-2.f * g_R[0][0] * k - g_R[0][1] * r + g_R[0][2] * i_ + g_R[1][0] * r - 2.f * g_R[1][1] * k + g_R[1][2] * j + g_R[2][0] * i_ + g_R[2][1] * j);
        gq[3] = t * (-g_R[0][1] * k + g_R[0][2] * j + g_R[1][0] * k - g_R[1][2] * i_ - g_R[2][0] * j + g_R[2][1] * i_);
        // two_s = 2 / (q.q + eps): d two_s / d q = -two_s^2 * q
#pragma unroll
        for (int c = 0; c < 4; ++c) gq[c] -= g_t * t * t * s.qn[c];
        // qn = q / (|q| + 1e-8)
        const float q[4] = {raw[5], raw[6], raw[7], raw[8]};
        const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        const float inv = 1.f / (n + 1e-8f);
        const float qg = q[0] * gq[0] + q[1] * gq[1] + q[2] * gq[2] + q[3] * gq[3];
        const float coef = n > 0.f ? qg * inv * inv / n : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) out9[5 + c] = gq[c] * inv - q[c] * coef;
    }
    float* __restrict__ draw = g.d_raw + ray * row;
    if (lane < 9) {
        float v = out9[0];
#pragma unroll
        for (int e = 1; e < 9; ++e) v = lane == e ? out9[e] : v;
        draw[lane] = v;
    }
    // SH rows: the samples of a ray share them -> sum of the per-sample gradients
    for (int e = lane; e < a.d_color; e += 32) {
        float acc = 0.f;
        for (int k = 0; k < a.samples; ++k) acc += g.d_color_sh[(ray * a.samples + k) * a.d_color + e];
        draw[9 + e] = acc;
    }
    for (int e = lane; e < a.d_feature; e += 32) {
        float acc = 0.f;
        for (int k = 0; k < a.samples; ++k) acc += g.d_feature_sh[(ray * a.samples + k) * a.d_feature + e];
        draw[9 + a.d_color + e] = acc;
    }
}

static int check(const LsGaussianHead* a) {
    if (!a) return ls_fail("gaussian head: args is NULL");
    if (a->rays <= 0 || a->rays_per_view <= 0 || a->rays % a->rays_per_view) return ls_fail("gaussian head: bad ray counts %lld / %d", (long long)a->rays, a->rays_per_view);
    if (a->width <= 0 || a->height <= 0 || a->width * a->height != a->rays_per_view) return ls_fail("gaussian head: width*height != rays_per_view");
    if (a->samples < 1 || a->samples > kMaxSamples) return ls_fail("gaussian head: samples %d not in 1..%d", a->samples, kMaxSamples);
    if (a->buckets != 32) return ls_fail("gaussian head: %d depth buckets (the kernel maps one bucket to one lane: 32)", a->buckets);
    if (a->d_color < 0 || a->d_feature < 0) return ls_fail("gaussian head: negative SH widths");
    if (!a->dlog || !a->raw || !a->extrinsics || !a->intrinsics || !a->near || !a->far) return ls_fail("gaussian head: NULL input");
    if (!a->deterministic && !a->u) return ls_fail("gaussian head: stochastic sampling needs the uniform samples u");
    if (a->deterministic && a->samples != 1) return ls_fail("gaussian head: deterministic (top-1) mode has one sample per ray");
    if (reinterpret_cast<uintptr_t>(a->dlog) & 7) return ls_fail("gaussian head: dlog must be 8-byte aligned");
    return 0;
}

// ---- reparameterised sample of a diagonal Gaussian stored as params = [mean | logvar] per row -------------------------------
// out[r, j] = mean + exp(0.5 * clamp(logvar, lo, hi)) * eps,   params row r = (mean[0..half), logvar[0..half)), half % 4 == 0
__global__ void __launch_bounds__(256) k_reparam_fwd(const float* __restrict__ params, const float* __restrict__ eps,
                                                     float* __restrict__ out, long long rows, int half, float lo, float hi) {
    const int q = half >> 2;                                               // float4s per half row
    const long long n4 = rows * q;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / q;
        const int j = (int)(t - r * q) << 2;
        const float4 m = *reinterpret_cast<const float4*>(params + r * 2 * half + j);
        const float4 lv = *reinterpret_cast<const float4*>(params + r * 2 * half + half + j);
        const float4 e = *reinterpret_cast<const float4*>(eps + r * half + j);
        float4 o;
        o.x = fmaf(expf(0.5f * fminf(fmaxf(lv.x, lo), hi)), e.x, m.x);
        o.y = fmaf(expf(0.5f * fminf(fmaxf(lv.y, lo), hi)), e.y, m.y);
        o.z = fmaf(expf(0.5f * fminf(fmaxf(lv.z, lo), hi)), e.z, m.z);
        o.w = fmaf(expf(0.5f * fminf(fmaxf(lv.w, lo), hi)), e.w, m.w);
        *reinterpret_cast<float4*>(out + r * half + j) = o;
    }
}

// d_params[r] = (g, g * eps * 0.5 * std * [lo <= logvar <= hi])   (torch.clamp passes the gradient on the closed interval)
__global__ void __launch_bounds__(256) k_reparam_bwd(const float* __restrict__ params, const float* __restrict__ eps,
                                                     const float* __restrict__ g, float* __restrict__ d_params, long long rows, int half,
                                                     float lo, float hi) {
    const int q = half >> 2;
    const long long n4 = rows * q;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / q;
        const int j = (int)(t - r * q) << 2;
        const float4 lv = *reinterpret_cast<const float4*>(params + r * 2 * half + half + j);
        const float4 e = *reinterpret_cast<const float4*>(eps + r * half + j);
        const float4 go = *reinterpret_cast<const float4*>(g + r * half + j);
        auto dl = [&](float l, float ee, float gg) { return (l >= lo && l <= hi) ? gg * ee * 0.5f * expf(0.5f * l) : 0.f; };
        *reinterpret_cast<float4*>(d_params + r * 2 * half + j) = go;
        *reinterpret_cast<float4*>(d_params + r * 2 * half + half + j) = make_float4(dl(lv.x, e.x, go.x), dl(lv.y, e.y, go.y), dl(lv.z, e.z, go.z), dl(lv.w, e.w, go.w));
    }
}

static int check_reparam(const void* a, const void* b, const void* c, long long rows, int half) {
    if (!a || !b || !c) return ls_fail("reparam: NULL pointer");
    if (rows <= 0 || half <= 0 || half % 4) return ls_fail("reparam: rows=%lld half=%d (half must be a positive multiple of 4)", rows, half);
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) return ls_fail("reparam: pointers must be 16-byte aligned");
    return 0;
}

}  // namespace lsh

using namespace lsh;

extern "C" int ls_reparam_forward(const float* params, const float* eps, float* out, int64_t rows, int32_t half, float lo, float hi, void* stream) {
    if (check_reparam(params, eps, out, rows, half)) return -1;
    const long long n4 = rows * (half / 4);
    const unsigned blocks = (unsigned)(n4 + 255) / 256 < 148u * 16u ? (unsigned)((n4 + 255) / 256) : 148u * 16u;
    k_reparam_fwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, eps, out, rows, half, lo, hi);
    return ls_check_cuda("k_reparam_fwd");
}

extern "C" int ls_reparam_backward(const float* params, const float* eps, const float* g, float* d_params, int64_t rows, int32_t half,
                                   float lo, float hi, void* stream) {
    if (check_reparam(params, eps, g, rows, half) || check_reparam(d_params, eps, g, rows, half)) return -1;
    const long long n4 = rows * (half / 4);
    const unsigned blocks = (unsigned)(n4 + 255) / 256 < 148u * 16u ? (unsigned)((n4 + 255) / 256) : 148u * 16u;
    k_reparam_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, eps, g, d_params, rows, half, lo, hi);
    return ls_check_cuda("k_reparam_bwd");
}

extern "C" int ls_gaussian_head_forward(const LsGaussianHead* a, const LsGaussianHeadOut* o, void* stream) {
    if (check(a)) return -1;
    if (!o || !o->means || !o->covariances || !o->opacity || !o->index || (a->d_color && !o->color_sh) || (a->d_feature && !o->feature_sh))
        return ls_fail("gaussian head forward: NULL output");
    const long long threads = (long long)a->rays * 32;
    k_ghead_fwd<<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(*a, *o);
    return ls_check_cuda("k_ghead_fwd");
}

extern "C" int ls_gaussian_head_backward(const LsGaussianHead* a, const LsGaussianHeadGrad* g, void* stream) {
    if (check(a)) return -1;
    if (!g || !g->index || !g->d_means || !g->d_covariances || !g->d_opacity || !g->d_dlog || !g->d_raw ||
        (a->d_color && !g->d_color_sh) || (a->d_feature && !g->d_feature_sh))
        return ls_fail("gaussian head backward: NULL pointer");
    if (reinterpret_cast<uintptr_t>(g->d_dlog) & 7) return ls_fail("gaussian head backward: d_dlog must be 8-byte aligned");
    const long long threads = (long long)a->rays * 32;
    k_ghead_bwd<<<(unsigned)((threads + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(*a, *g);
    return ls_check_cuda("k_ghead_bwd");
}
