// ls_norm_nhwc.cu -- GroupNorm (+ fused SiLU) for NHWC fp32 activations, forward and backward (include/ls_norm.h).
//
// The implicit-GEMM convolutions (ls_conv.cu) keep the VAE decoder's activations NHWC end to end, so the
// `nonlinearity(norm(x))` pairs of the decoder (/root/reference/src/model/autoencoder/autoencoder_kl.py:93-124 -> diffusers
// ResnetBlock2D / Decoder.conv_norm_out: GroupNorm(32, C, eps=1e-6) + SiLU; the mid-block attention's GroupNorm without
// activation) normalise over (pixels x C/G channels) of an (N, H*W, C) tensor.
//
//   pass 1  per-(image, channel) sums over the pixels: forward {sum x, sum x^2}, backward {sum ds, sum ds*xhat} with
//           ds = dy * silu'(u).  A thread owns 4 consecutive channels (one float4 lane of a pixel) and walks pixel rows, so a
//           warp reads 512 contiguous bytes per instruction; partials meet in shared memory, then fp64 global atomics.
//   pass 2  elementwise apply with per-thread constants (the thread's 4 channels never change):
//           forward  y  = silu(a_c x + b_c),  backward  dx = rstd (gamma ds - P_g/L - xhat Q_g/L).
// Group statistics are assembled from the per-channel sums by each thread for its own channels (C/G <= 16 values).
// HBM-bound: 12 B/element forward (x twice, y once), 20 B/element backward (x, dy twice; dx once).
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_host.h"
#include "ls_norm.h"

namespace lsnh {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;        // pixels per thread and loop trip (loads issued back to back)

__device__ __forceinline__ float silu(float u) { return u / (1.f + __expf(-u)); }
__device__ __forceinline__ float dsilu(float u, float g) {
    const float sg = 1.f / (1.f + __expf(-u));
    return g * sg * fmaf(u, 1.f - sg, 1.f);
}

// group statistics of channel c from the per-channel {sum, sum of squares}
__device__ __forceinline__ void group_mean_rstd(const double* __restrict__ sums_n, int c, int cpg, double inv_len, float eps,
                                                float& mean, float& rstd) {
    const int c0 = (c / cpg) * cpg;
    double s = 0.0, ss = 0.0;
    for (int k = 0; k < cpg; ++k) { s += sums_n[2 * (c0 + k)]; ss += sums_n[2 * (c0 + k) + 1]; }
    const double m = s * inv_len;
    double var = ss * inv_len - m * m;
    var = var > 0.0 ? var : 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// pass 1.  grid (splits, N); block = (C/4 lanes) x (kThreads / (C/4) pixel rows); dynamic smem 2*C floats.
// BWD = 0: out[n][c] += {sum x, sum x^2};  BWD = 1: out[n][c] += {sum ds, sum ds * xhat}
template <int BWD, int ACT>
__global__ void __launch_bounds__(kThreads) k_gn_nhwc_sums(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const double* __restrict__ stats, double* __restrict__ out, int C, int G,
                                                           long long HW, long long chunk, float eps) {
    extern __shared__ float s_acc[];          // [C][2]
    const int lanes = C >> 2, rows = kThreads / lanes;
    const int lane = threadIdx.x % lanes, prow = threadIdx.x / lanes;
    const int n = blockIdx.y, c = 4 * lane;
    for (int i = threadIdx.x; i < 2 * C; i += kThreads) s_acc[i] = 0.f;
    __syncthreads();
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f}, gm[4], bt[4];
    if (BWD && prow < rows) {
        const int cpg = C / G;
        const double inv_len = 1.0 / ((double)cpg * (double)HW);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            group_mean_rstd(stats + 2ll * n * C, c + k, cpg, inv_len, eps, mean[k], rstd[k]);
            gm[k] = gamma[c + k];
            bt[k] = beta[c + k];
        }
    }
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > HW) hi = HW;
    if (prow < rows) {
        const float* xp = x + ((long long)n * HW) * C + c;
        const float* gp = BWD ? dy + ((long long)n * HW) * C + c : nullptr;
        // kUnroll pixels per thread and trip: all loads of a trip are issued before the first use (bytes in flight, not
        // occupancy, is what keeps an HBM-bound streaming kernel fed: 2 x 4 x 16 B per thread here)
        for (long long p0 = lo + prow; p0 < hi; p0 += (long long)kUnroll * rows) {
            float4 v[kUnroll], d[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long p = p0 + (long long)u * rows;
                const bool ok = p < hi;
                v[u] = ok ? __ldg(reinterpret_cast<const float4*>(xp + p * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (BWD) d[u] = ok ? __ldg(reinterpret_cast<const float4*>(gp + p * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (p0 + (long long)u * rows >= hi) break;
                const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                if (!BWD) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { a0[k] += xs[k]; a1[k] = fmaf(xs[k], xs[k], a1[k]); }
                } else {
                    const float ds_[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float xh = (xs[k] - mean[k]) * rstd[k];
                        const float ds = ACT ? dsilu(fmaf(xh, gm[k], bt[k]), ds_[k]) : ds_[k];
                        a0[k] += ds;
                        a1[k] = fmaf(ds, xh, a1[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { atomicAdd(&s_acc[2 * (c + k)], a0[k]); atomicAdd(&s_acc[2 * (c + k) + 1], a1[k]); }
    }
    __syncthreads();
    if (hi > lo)
        for (int i = threadIdx.x; i < 2 * C; i += kThreads) atomicAdd(&out[2ll * n * C + i], (double)s_acc[i]);
}

// pass 2 forward.  grid (splits, N)
template <int ACT>
__global__ void __launch_bounds__(kThreads) k_gn_nhwc_apply(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const double* __restrict__ stats,
                                                            float* __restrict__ y, int C, int G, long long HW, long long chunk,
                                                            float eps) {
    const int lanes = C >> 2, rows = kThreads / lanes;
    const int lane = threadIdx.x % lanes, prow = threadIdx.x / lanes;
    if (prow >= rows) return;
    const int n = blockIdx.y, c = 4 * lane, cpg = C / G;
    const double inv_len = 1.0 / ((double)cpg * (double)HW);
    float a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float mean, rstd;
        group_mean_rstd(stats + 2ll * n * C, c + k, cpg, inv_len, eps, mean, rstd);
        a[k] = rstd * gamma[c + k];
        b[k] = beta[c + k] - mean * a[k];
    }
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > HW) hi = HW;
    const float* xp = x + ((long long)n * HW) * C + c;
    float* yp = y + ((long long)n * HW) * C + c;
    for (long long p0 = lo + prow; p0 < hi; p0 += (long long)kUnroll * rows) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long p = p0 + (long long)u * rows;
            if (p < hi) v[u] = __ldg(reinterpret_cast<const float4*>(xp + p * C));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long p = p0 + (long long)u * rows;
            if (p >= hi) break;
            float4 w = v[u];
            w.x = fmaf(a[0], w.x, b[0]); w.y = fmaf(a[1], w.y, b[1]); w.z = fmaf(a[2], w.z, b[2]); w.w = fmaf(a[3], w.w, b[3]);
            if (ACT) { w.x = silu(w.x); w.y = silu(w.y); w.z = silu(w.z); w.w = silu(w.w); }
            *reinterpret_cast<float4*>(yp + p * C) = w;
        }
    }
}

// pass 2 backward.  sums = per-(image, channel) {sum ds, sum ds*xhat} of pass 1
template <int ACT>
__global__ void __launch_bounds__(kThreads) k_gn_nhwc_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const double* __restrict__ stats, const double* __restrict__ sums,
                                                                float* __restrict__ dx, int C, int G, long long HW, long long chunk,
                                                                float eps) {
    const int lanes = C >> 2, rows = kThreads / lanes;
    const int lane = threadIdx.x % lanes, prow = threadIdx.x / lanes;
    if (prow >= rows) return;
    const int n = blockIdx.y, c = 4 * lane, cpg = C / G;
    const double inv_len = 1.0 / ((double)cpg * (double)HW);
    float mean[4], rstd[4], gm[4], bt[4], pm[4], qm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        group_mean_rstd(stats + 2ll * n * C, c + k, cpg, inv_len, eps, mean[k], rstd[k]);
        gm[k] = gamma[c + k];
        bt[k] = beta[c + k];
        const int c0 = ((c + k) / cpg) * cpg;
        double P = 0.0, Q = 0.0;
        for (int j = 0; j < cpg; ++j) {
            const double gj = (double)gamma[c0 + j];
            P += gj * sums[2 * ((long long)n * C + c0 + j)];
            Q += gj * sums[2 * ((long long)n * C + c0 + j) + 1];
        }
        pm[k] = (float)(P * inv_len);
        qm[k] = (float)(Q * inv_len);
    }
    const long long lo = (long long)blockIdx.x * chunk;
    long long hi = lo + chunk;
    if (hi > HW) hi = HW;
    const float* xp = x + ((long long)n * HW) * C + c;
    const float* gp = dy + ((long long)n * HW) * C + c;
    float* op = dx + ((long long)n * HW) * C + c;
    for (long long p0 = lo + prow; p0 < hi; p0 += (long long)kUnroll * rows) {
        float4 v[kUnroll], d[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long p = p0 + (long long)u * rows;
            if (p < hi) {
                v[u] = __ldg(reinterpret_cast<const float4*>(xp + p * C));
                d[u] = __ldg(reinterpret_cast<const float4*>(gp + p * C));
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long p = p0 + (long long)u * rows;
            if (p >= hi) break;
            const float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, ds_[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xh = (xs[k] - mean[k]) * rstd[k];
                const float ds = ACT ? dsilu(fmaf(xh, gm[k], bt[k]), ds_[k]) : ds_[k];
                r[k] = rstd[k] * (fmaf(gm[k], ds, -pm[k]) - xh * qm[k]);
            }
            *reinterpret_cast<float4*>(op + p * C) = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

static int check_args(const LsGroupNorm* a) {
    if (!a) return ls_fail("groupnorm args is NULL");
    if (a->N <= 0 || a->C <= 0 || a->G <= 0 || a->C % a->G || a->HW <= 0) return ls_fail("groupnorm: bad sizes N=%d C=%d G=%d HW=%lld", a->N, a->C, a->G, (long long)a->HW);
    if (a->C % 4 || a->C > 4 * kThreads) return ls_fail("groupnorm (NHWC): C=%d must be a multiple of 4 and <= %d", a->C, 4 * kThreads);
    if (a->N > 65535) return ls_fail("groupnorm (NHWC): N=%d exceeds the grid limit", a->N);
    if (!a->x || !a->gamma || !a->beta || !a->stats) return ls_fail("groupnorm: NULL pointer");
    if (reinterpret_cast<uintptr_t>(a->x) & 15) return ls_fail("groupnorm: x must be 16-byte aligned");
    return 0;
}

// pixels per block chunk: enough blocks to fill the GPU ~4x over, at least 64 pixel rows per block
static void split(const LsGroupNorm* a, int& splits, long long& chunk) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (4ll * sms + a->N - 1) / a->N;
    const long long max_splits = (a->HW + 63) / 64;
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    chunk = (a->HW + want - 1) / want;
    splits = (int)((a->HW + chunk - 1) / chunk);
}

}  // namespace lsnh

using namespace lsnh;

extern "C" int ls_groupnorm_nhwc_forward(const LsGroupNorm* a, float* y, void* stream_) {
    if (check_args(a)) return -1;
    if (!y || (reinterpret_cast<uintptr_t>(y) & 15)) return ls_fail("groupnorm: y must be a 16-byte aligned pointer");
    cudaStream_t stream = (cudaStream_t)stream_;
    int splits;
    long long chunk;
    split(a, splits, chunk);
    cudaMemsetAsync(a->stats, 0, sizeof(double) * 2 * (size_t)a->N * a->C, stream);
    const dim3 grid(splits, a->N);
    const size_t smem = sizeof(float) * 2 * a->C;
    k_gn_nhwc_sums<0, 0><<<grid, kThreads, smem, stream>>>(a->x, nullptr, a->gamma, a->beta, nullptr, a->stats, a->C, a->G, a->HW, chunk, a->eps);
    if (a->act) k_gn_nhwc_apply<1><<<grid, kThreads, 0, stream>>>(a->x, a->gamma, a->beta, a->stats, y, a->C, a->G, a->HW, chunk, a->eps);
    else k_gn_nhwc_apply<0><<<grid, kThreads, 0, stream>>>(a->x, a->gamma, a->beta, a->stats, y, a->C, a->G, a->HW, chunk, a->eps);
    return ls_check_cuda("groupnorm nhwc forward");
}

extern "C" int ls_groupnorm_nhwc_backward(const LsGroupNorm* a, const float* dy, float* dx, double* sums, void* stream_) {
    if (check_args(a)) return -1;
    if (!dy || !dx || !sums) return ls_fail("groupnorm backward: NULL pointer");
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) return ls_fail("groupnorm backward: dy/dx must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    int splits;
    long long chunk;
    split(a, splits, chunk);
    cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)a->N * a->C, stream);
    const dim3 grid(splits, a->N);
    const size_t smem = sizeof(float) * 2 * a->C;
    if (a->act) {
        k_gn_nhwc_sums<1, 1><<<grid, kThreads, smem, stream>>>(a->x, dy, a->gamma, a->beta, a->stats, sums, a->C, a->G, a->HW, chunk, a->eps);
        k_gn_nhwc_bwd_apply<1><<<grid, kThreads, 0, stream>>>(a->x, dy, a->gamma, a->beta, a->stats, sums, dx, a->C, a->G, a->HW, chunk, a->eps);
    } else {
        k_gn_nhwc_sums<1, 0><<<grid, kThreads, smem, stream>>>(a->x, dy, a->gamma, a->beta, a->stats, sums, a->C, a->G, a->HW, chunk, a->eps);
        k_gn_nhwc_bwd_apply<0><<<grid, kThreads, 0, stream>>>(a->x, dy, a->gamma, a->beta, a->stats, sums, dx, a->C, a->G, a->HW, chunk, a->eps);
    }
    return ls_check_cuda("groupnorm nhwc backward");
}
