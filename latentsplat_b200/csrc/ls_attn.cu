// ls_attn.cu -- fused single-query multi-head attention (epipolar cross-attention), forward and backward.
//
// Replaces, for the epipolar transformer's cross-attention (one query token per ray, S = 32 key/value tokens sampled
// on the epipolar line; /root/reference/src/model/transformer/attention.py:54-70 called from
// epipolar_transformer.py:127-135), the eager sequence  chunk -> rearrange (2 transposed copies of the 4.3 GB kv
// tensor) -> bmm(q, k^T) -> softmax -> bmm(attn, v) -> rearrange.  Here each warp owns one (ray, head): it streams the
// head's K rows once (coalesced float4 per lane), reduces the 32 scores with shuffles, soft-maxes them in registers and
// streams the V rows once.  HBM traffic = the kv tensor read exactly once per pass (+ written once in backward).
//
// Layouts (all fp32, row-major):  q (R, H*D);  kv (R, S, 2*H*D) with K = columns [0, H*D), V = [H*D, 2*H*D), head h at
// [h*D, (h+1)*D) -- exactly what `to_kv(z).chunk(2, -1)` + "b n (h d) -> b h n d" index;  out (R, H*D);
// p (R, H, S) softmax probabilities saved for backward.   D must be 128 (4 floats per lane), S <= 32.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_gemm.h"
#include "ls_host.h"

namespace lsa {

constexpr int D = 128;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(256) k_sq_attn_fwd(const float* __restrict__ q, const float* __restrict__ kv,
                                                     float* __restrict__ out, float* __restrict__ p, int R, int H, int S,
                                                     float scale) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (ray, head)
    if (w >= (long long)R * H) return;
    const int r = (int)(w / H), h = (int)(w % H);
    const int HD = H * D;
    const float4 q4 = *reinterpret_cast<const float4*>(q + (size_t)r * HD + h * D + 4 * lane);
    const float* kbase = kv + (size_t)r * S * 2 * HD + h * D + 4 * lane;
    float my_score = -INFINITY;                       // lane j keeps score j
    for (int j = 0; j < S; ++j) {
        const float4 k4 = *reinterpret_cast<const float4*>(kbase + (size_t)j * 2 * HD);
        const float s = warp_sum(dot4(q4, k4)) * scale;
        if (lane == j) my_score = s;
    }
    const float m = warp_max(my_score);
    const float e = lane < S ? __expf(my_score - m) : 0.f;
    const float prob = e / warp_sum(e);
    if (lane < S) p[((size_t)r * H + h) * S + lane] = prob;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* vbase = kbase + HD;
    for (int j = 0; j < S; ++j) {
        const float pj = __shfl_sync(0xffffffffu, prob, j);
        const float4 v4 = *reinterpret_cast<const float4*>(vbase + (size_t)j * 2 * HD);
        acc.x = fmaf(pj, v4.x, acc.x); acc.y = fmaf(pj, v4.y, acc.y); acc.z = fmaf(pj, v4.z, acc.z); acc.w = fmaf(pj, v4.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + (size_t)r * HD + h * D + 4 * lane) = acc;
}

__global__ void __launch_bounds__(256) k_sq_attn_bwd(const float* __restrict__ q, const float* __restrict__ kv,
                                                     const float* __restrict__ p, const float* __restrict__ dout,
                                                     float* __restrict__ dq, float* __restrict__ dkv, int R, int H, int S,
                                                     float scale) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)R * H) return;
    const int r = (int)(w / H), h = (int)(w % H);
    const int HD = H * D;
    const size_t qoff = (size_t)r * HD + h * D + 4 * lane;
    const float4 q4 = *reinterpret_cast<const float4*>(q + qoff);
    const float4 g4 = *reinterpret_cast<const float4*>(dout + qoff);
    const size_t kvoff = (size_t)r * S * 2 * HD + h * D + 4 * lane;
    const float prob = lane < S ? p[((size_t)r * H + h) * S + lane] : 0.f;
    // dp_j = dout . v_j ; dv_j = p_j dout
    float my_dp = 0.f;
    for (int j = 0; j < S; ++j) {
        const float4 v4 = *reinterpret_cast<const float4*>(kv + kvoff + HD + (size_t)j * 2 * HD);
        const float dp = warp_sum(dot4(g4, v4));
        if (lane == j) my_dp = dp;
        const float pj = __shfl_sync(0xffffffffu, prob, j);
        *reinterpret_cast<float4*>(dkv + kvoff + HD + (size_t)j * 2 * HD) = make_float4(pj * g4.x, pj * g4.y, pj * g4.z, pj * g4.w);
    }
    const float ds = prob * (my_dp - warp_sum(prob * my_dp)) * scale;     // d score_j (scale folded in)
    float4 dq4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < S; ++j) {
        const float dsj = __shfl_sync(0xffffffffu, ds, j);
        const float4 k4 = *reinterpret_cast<const float4*>(kv + kvoff + (size_t)j * 2 * HD);
        dq4.x = fmaf(dsj, k4.x, dq4.x); dq4.y = fmaf(dsj, k4.y, dq4.y); dq4.z = fmaf(dsj, k4.z, dq4.z); dq4.w = fmaf(dsj, k4.w, dq4.w);
        *reinterpret_cast<float4*>(dkv + kvoff + (size_t)j * 2 * HD) = make_float4(dsj * q4.x, dsj * q4.y, dsj * q4.z, dsj * q4.w);
    }
    *reinterpret_cast<float4*>(dq + qoff) = dq4;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight-absorbed variant.  With one query per ray and bias-free projections,
//     score_{h,j} = q_h . (W_k,h z_j) = (W_k,h^T q_h) . z_j          out_h = sum_j p_{h,j} W_v,h z_j = W_v,h (sum_j p_{h,j} z_j)
// so the 1 048 576 x 128 -> 1024 `to_kv` GEMM (137 GMAC and a 4.3 GB kv tensor per layer, epipolar_transformer.py:127-135
// + attention.py:60-61) is never formed: the kernels below attend over the raw 128-wide samples z with absorbed queries
// qt_h = W_k,h^T q_h and return zbar_h = sum_j p_{h,j} z_j; the (tiny) per-head GEMMs around them run on ls_gemm_tf32.
// One warp per ray; all heads share each z row, which is read from HBM exactly once per pass.
constexpr int kMaxHeads = 8;

// kMaxHeads is a template parameter: the per-head register arrays (absorbed query, gradient, accumulators: 3 float4 + 3
// scalars per head in backward) set the occupancy -- sized for 8 heads the backward kernel needs 136 registers (one block
// per SM, 12 % warps active, 1.09 ms); sized for the 4 heads the model uses it fits three.
template <int kMaxHeads>
__global__ void __launch_bounds__(256) k_absorbed_attn_fwd(const float* __restrict__ qt, const float* __restrict__ z,
                                                           float* __restrict__ zbar, float* __restrict__ p, int R, int H,
                                                           int S, float scale) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    float4 q4[kMaxHeads];
    float score[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
        score[h] = -INFINITY;
        if (h < H) q4[h] = *reinterpret_cast<const float4*>(qt + ((size_t)r * H + h) * D + 4 * lane);
    }
    const float* zr = z + (size_t)r * S * D + 4 * lane;
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
        const float4 z4 = *reinterpret_cast<const float4*>(zr + (size_t)j * D);
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
            if (h < H) {
                const float s = warp_sum(dot4(q4[h], z4)) * scale;
                if (lane == j) score[h] = s;
            }
    }
    float prob[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
        if (h < H) {
            const float m = warp_max(score[h]);
            const float e = lane < S ? __expf(score[h] - m) : 0.f;
            prob[h] = e / warp_sum(e);
            if (lane < S) p[((size_t)r * H + h) * S + lane] = prob[h];
        }
    float4 acc[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < S; ++j) {
        const float4 z4 = *reinterpret_cast<const float4*>(zr + (size_t)j * D);
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
            if (h < H) {
                const float pj = __shfl_sync(0xffffffffu, prob[h], j);
                acc[h].x = fmaf(pj, z4.x, acc[h].x); acc[h].y = fmaf(pj, z4.y, acc[h].y);
                acc[h].z = fmaf(pj, z4.z, acc[h].z); acc[h].w = fmaf(pj, z4.w, acc[h].w);
            }
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
        if (h < H) *reinterpret_cast<float4*>(zbar + ((size_t)r * H + h) * D + 4 * lane) = acc[h];
}

template <int kMaxHeads>
__global__ void __launch_bounds__(256, kMaxHeads <= 4 ? 3 : 1) k_absorbed_attn_bwd(const float* __restrict__ qt, const float* __restrict__ z,
                                                           const float* __restrict__ p, const float* __restrict__ dzbar,
                                                           float* __restrict__ dqt, float* __restrict__ dz, int R, int H,
                                                           int S, float scale) {
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    float4 q4[kMaxHeads], g4[kMaxHeads], dq4[kMaxHeads];
    float prob[kMaxHeads], dp[kMaxHeads], ds[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
        dp[h] = 0.f; prob[h] = 0.f; dq4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < H) {
            q4[h] = *reinterpret_cast<const float4*>(qt + ((size_t)r * H + h) * D + 4 * lane);
            g4[h] = *reinterpret_cast<const float4*>(dzbar + ((size_t)r * H + h) * D + 4 * lane);
            if (lane < S) prob[h] = p[((size_t)r * H + h) * S + lane];
        }
    }
    const float* zr = z + (size_t)r * S * D + 4 * lane;
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
        const float4 z4 = *reinterpret_cast<const float4*>(zr + (size_t)j * D);
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
            if (h < H) {
                const float v = warp_sum(dot4(g4[h], z4));
                if (lane == j) dp[h] = v;
            }
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
        if (h < H) ds[h] = prob[h] * (dp[h] - warp_sum(prob[h] * dp[h])) * scale;
    float* dzr = dz + (size_t)r * S * D + 4 * lane;
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
        const float4 z4 = *reinterpret_cast<const float4*>(zr + (size_t)j * D);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
            if (h < H) {
                const float pj = __shfl_sync(0xffffffffu, prob[h], j);
                const float dsj = __shfl_sync(0xffffffffu, ds[h], j);
                o.x += pj * g4[h].x + dsj * q4[h].x; o.y += pj * g4[h].y + dsj * q4[h].y;
                o.z += pj * g4[h].z + dsj * q4[h].z; o.w += pj * g4[h].w + dsj * q4[h].w;
                dq4[h].x = fmaf(dsj, z4.x, dq4[h].x); dq4[h].y = fmaf(dsj, z4.y, dq4[h].y);
                dq4[h].z = fmaf(dsj, z4.z, dq4[h].z); dq4[h].w = fmaf(dsj, z4.w, dq4[h].w);
            }
        *reinterpret_cast<float4*>(dzr + (size_t)j * D) = o;
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
        if (h < H) *reinterpret_cast<float4*>(dqt + ((size_t)r * H + h) * D + 4 * lane) = dq4[h];
}

}  // namespace lsa

extern "C" LS_API int ls_absorbed_attention_forward(const float* qt, const float* z, float* zbar, float* p, int32_t R, int32_t H,
                                                    int32_t S, int32_t Dz, float scale, void* stream) {
    if (Dz != lsa::D) return ls_fail("absorbed attention: sample width %d != 128", Dz);
    if (S < 1 || S > 32 || H < 1 || H > lsa::kMaxHeads) return ls_fail("absorbed attention: S=%d (1..32) H=%d (1..8)", S, H);
    if (R <= 0) return 0;
    if (H <= 4) lsa::k_absorbed_attn_fwd<4><<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(qt, z, zbar, p, R, H, S, scale);
    else lsa::k_absorbed_attn_fwd<8><<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(qt, z, zbar, p, R, H, S, scale);
    return ls_check_cuda("k_absorbed_attn_fwd");
}

extern "C" LS_API int ls_absorbed_attention_backward(const float* qt, const float* z, const float* p, const float* dzbar,
                                                     float* dqt, float* dz, int32_t R, int32_t H, int32_t S, int32_t Dz,
                                                     float scale, void* stream) {
    if (Dz != lsa::D) return ls_fail("absorbed attention: sample width %d != 128", Dz);
    if (S < 1 || S > 32 || H < 1 || H > lsa::kMaxHeads) return ls_fail("absorbed attention: S=%d (1..32) H=%d (1..8)", S, H);
    if (R <= 0) return 0;
    if (H <= 4) lsa::k_absorbed_attn_bwd<4><<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(qt, z, p, dzbar, dqt, dz, R, H, S, scale);
    else lsa::k_absorbed_attn_bwd<8><<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(qt, z, p, dzbar, dqt, dz, R, H, S, scale);
    return ls_check_cuda("k_absorbed_attn_bwd");
}

extern "C" LS_API int ls_sq_attention_forward(const float* q, const float* kv, float* out, float* p, int32_t R, int32_t H,
                                              int32_t S, int32_t Dh, float scale, void* stream) {
    if (Dh != lsa::D) return ls_fail("single-query attention: head dim %d != 128", Dh);
    if (S < 1 || S > 32) return ls_fail("single-query attention: S=%d not in 1..32", S);
    if (R <= 0 || H <= 0) return 0;
    const long long warps = (long long)R * H;
    lsa::k_sq_attn_fwd<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(q, kv, out, p, R, H, S, scale);
    return ls_check_cuda("k_sq_attn_fwd");
}

extern "C" LS_API int ls_sq_attention_backward(const float* q, const float* kv, const float* p, const float* dout, float* dq,
                                               float* dkv, int32_t R, int32_t H, int32_t S, int32_t Dh, float scale,
                                               void* stream) {
    if (Dh != lsa::D) return ls_fail("single-query attention: head dim %d != 128", Dh);
    if (S < 1 || S > 32) return ls_fail("single-query attention: S=%d not in 1..32", S);
    if (R <= 0 || H <= 0) return 0;
    const long long warps = (long long)R * H;
    lsa::k_sq_attn_bwd<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(q, kv, p, dout, dq, dkv, R, H, S, scale);
    return ls_check_cuda("k_sq_attn_bwd");
}
