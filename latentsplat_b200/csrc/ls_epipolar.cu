// ls_epipolar.cu -- fused epipolar-line feature gather (+ depth encoding), forward and backward.
//
// What it replaces in the reference's encoder (one pass instead of ~10 over a 537 MB tensor):
//   /root/reference/src/model/encoder/epipolar/epipolar_sampler.py:96-112   transpose -> F.grid_sample(bilinear, zeros,
//       align_corners=False) on NCHW features -> rearrange -> transpose back -> multiply by the overlap mask
//   /root/reference/src/model/encoder/epipolar/epipolar_transformer.py:121-122   + Linear(PositionalEncoding(depth))
// i.e.   z[row, s, :] = valid[row] * bilinear(feat[image[row]], xy[row, s]) + (W_e pe(d[row, s]) + b_e)
// with pe(d)[2o + t] = sin(d * 2 pi 2^o + t * pi/2)   (positional_encoding.py:8-36).
//
// B200 mapping: features are read channels-last (I, Hf, Wf, 128): one bilinear corner = one 512-B line, a float4 per lane,
// and the whole (I * Hf * Wf * 512 B = 16.8 MB) map stays L2 resident, so HBM traffic is the z tensor written once
// (forward) / read once (backward).  One warp per sample, rows strided over a persistent grid.  Backward scatters with
// 16-B vector reductions (red.global.add.v4.f32) into the L2-resident feature gradient and keeps the 128 x P encoding
// weight gradient in registers across the warp's whole sample range (one flush per block at the end).
//
// The sample positions / depths are pure camera geometry (no parameters), so no gradient is produced for them -- the
// reference's autograd does not reach them either.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_epipolar.h"
#include "ls_host.h"

namespace lse {

constexpr int C = 128;       // feature width: 4 floats per lane
constexpr int kMaxP = 32;    // encoding width (2 * octaves) supported
constexpr int kThreads = 256;

struct Corner {
    int off[4];      // element offsets of the 4 corners' channel rows inside the image, -1 when outside (zero padding)
    float w[4];
};

// grid_sample(align_corners=False): pixel = ((2 xy - 1) + 1) * size / 2 - 0.5  (GridSamplerKernel unnormalize)
__device__ __forceinline__ Corner corners(float x, float y, int Hf, int Wf, float valid) {
    const float gx = 2.f * x - 1.f, gy = 2.f * y - 1.f;
    const float ix = ((gx + 1.f) * Wf - 1.f) * 0.5f, iy = ((gy + 1.f) * Hf - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const float tx = ix - fx, ty = iy - fy;
    // floorf of +-inf / nan: treat as outside
    const bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
    const int x0 = finite ? (int)fx : -2, y0 = finite ? (int)fy : -2;
    Corner c;
    const float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        const bool in = xx >= 0 && xx < Wf && yy >= 0 && yy < Hf;
        c.off[k] = in ? (yy * Wf + xx) * C : -1;
        c.w[k] = in ? wx[k & 1] * wy[k >> 1] * valid : 0.f;
    }
    return c;
}

__device__ __forceinline__ float pe_value(float d, int k) {
    // sin(d * (2 pi 2^octave) + phase): multiply then add, as torch does (positional_encoding.py:31-33)
    const float freq = 6.28318530717958647692f * (float)(1u << (k >> 1));
    return sinf(__fadd_rn(__fmul_rn(d, freq), (k & 1) ? 1.57079632679489661923f : 0.f));
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int P>
__global__ void __launch_bounds__(kThreads) k_gather_fwd(const float* __restrict__ feat, const float* __restrict__ xy,
                                                         const float* __restrict__ depth, const int32_t* __restrict__ image,
                                                         const float* __restrict__ valid, const float* __restrict__ We,
                                                         const float* __restrict__ be, float* __restrict__ z, int rows, int S,
                                                         int Hf, int Wf) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float w[4][P > 0 ? P : 1];
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (P > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < P; ++k) w[i][k] = We[(4 * lane + i) * P + k];
        bias = *reinterpret_cast<const float4*>(be + 4 * lane);
    }
    const size_t image_stride = (size_t)Hf * Wf * C;
    for (int row = warp; row < rows; row += warps) {
        const float* img = feat + (size_t)image[row] * image_stride + 4 * lane;
        const float v = valid[row];
        // lane j owns sample j's geometry
        float sx = 0.f, sy = 0.f, sd = 0.f;
        if (lane < S) {
            const float2 p = *reinterpret_cast<const float2*>(xy + ((size_t)row * S + lane) * 2);
            sx = p.x; sy = p.y;
            if (P > 0) sd = depth[(size_t)row * S + lane];
        }
        const Corner mine = corners(sx, sy, Hf, Wf, v);
        float* zr = z + (size_t)row * S * C + 4 * lane;
        for (int j = 0; j < S; ++j) {
            float4 acc = bias;
            if (P > 0) {
                const float d = __shfl_sync(0xffffffffu, sd, j);
                const float pe = lane < P ? pe_value(d, lane) : 0.f;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float e = __shfl_sync(0xffffffffu, pe, k);
                    acc.x = fmaf(w[0][k], e, acc.x); acc.y = fmaf(w[1][k], e, acc.y);
                    acc.z = fmaf(w[2][k], e, acc.z); acc.w = fmaf(w[3][k], e, acc.w);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int off = __shfl_sync(0xffffffffu, mine.off[k], j);
                const float wk = __shfl_sync(0xffffffffu, mine.w[k], j);
                if (off >= 0 && wk != 0.f) {
                    const float4 f = *reinterpret_cast<const float4*>(img + off);
                    acc.x = fmaf(wk, f.x, acc.x); acc.y = fmaf(wk, f.y, acc.y);
                    acc.z = fmaf(wk, f.z, acc.z); acc.w = fmaf(wk, f.w, acc.w);
                }
            }
            __stcs(reinterpret_cast<float4*>(zr + (size_t)j * C), acc);      // streamed: keep the feature map in L2
        }
    }
}

template <int P>
__global__ void __launch_bounds__(kThreads) k_gather_bwd(const float* __restrict__ dz, const float* __restrict__ xy,
                                                         const float* __restrict__ depth, const int32_t* __restrict__ image,
                                                         const float* __restrict__ valid, float* __restrict__ dfeat,
                                                         float* __restrict__ dWe, float* __restrict__ dbe, int rows, int S,
                                                         int Hf, int Wf) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float gw[4][P > 0 ? P : 1];
    float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < (P > 0 ? P : 1); ++k) gw[i][k] = 0.f;
    const size_t image_stride = (size_t)Hf * Wf * C;
    for (int row = warp; row < rows; row += warps) {
        float* img = dfeat + (size_t)image[row] * image_stride + 4 * lane;
        const float v = valid[row];
        float sx = 0.f, sy = 0.f, sd = 0.f;
        if (lane < S) {
            const float2 p = *reinterpret_cast<const float2*>(xy + ((size_t)row * S + lane) * 2);
            sx = p.x; sy = p.y;
            if (P > 0) sd = depth[(size_t)row * S + lane];
        }
        const Corner mine = corners(sx, sy, Hf, Wf, v);
        const float* gr = dz + (size_t)row * S * C + 4 * lane;
        for (int j = 0; j < S; ++j) {
            const float4 g = __ldcs(reinterpret_cast<const float4*>(gr + (size_t)j * C));
            if (P > 0) {
                const float d = __shfl_sync(0xffffffffu, sd, j);
                const float pe = lane < P ? pe_value(d, lane) : 0.f;
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const float e = __shfl_sync(0xffffffffu, pe, k);
                    gw[0][k] = fmaf(g.x, e, gw[0][k]); gw[1][k] = fmaf(g.y, e, gw[1][k]);
                    gw[2][k] = fmaf(g.z, e, gw[2][k]); gw[3][k] = fmaf(g.w, e, gw[3][k]);
                }
                gb.x += g.x; gb.y += g.y; gb.z += g.z; gb.w += g.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int off = __shfl_sync(0xffffffffu, mine.off[k], j);
                const float wk = __shfl_sync(0xffffffffu, mine.w[k], j);
                if (off >= 0 && wk != 0.f) red_add_v4(img + off, make_float4(wk * g.x, wk * g.y, wk * g.z, wk * g.w));
            }
        }
    }
    if (P > 0) {
        // block-level combine of the encoding gradients: shared-memory float atomics, then one global atomic per entry
        __shared__ float s_w[C * (P > 0 ? P : 1)];
        __shared__ float s_b[C];
        for (int i = threadIdx.x; i < C * P; i += blockDim.x) s_w[i] = 0.f;
        if (threadIdx.x < C) s_b[threadIdx.x] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < P; ++k) atomicAdd(&s_w[(4 * lane + i) * P + k], gw[i][k]);
        atomicAdd(&s_b[4 * lane + 0], gb.x); atomicAdd(&s_b[4 * lane + 1], gb.y);
        atomicAdd(&s_b[4 * lane + 2], gb.z); atomicAdd(&s_b[4 * lane + 3], gb.w);
        __syncthreads();
        for (int i = threadIdx.x; i < C * P; i += blockDim.x) atomicAdd(&dWe[i], s_w[i]);
        if (threadIdx.x < C) atomicAdd(&dbe[threadIdx.x], s_b[threadIdx.x]);
    }
}

static int check_args(const LsEpipolarGather* a) {
    if (!a) return ls_fail("epipolar gather: args is NULL");
    if (a->channels != C) return ls_fail("epipolar gather: channels=%d, kernel is specialised for 128", a->channels);
    if (a->samples < 1 || a->samples > 32) return ls_fail("epipolar gather: samples=%d not in 1..32", a->samples);
    if (a->rows < 0 || a->images < 1 || a->height < 1 || a->width < 1) return ls_fail("epipolar gather: bad sizes");
    if (a->encoding_width != 0 && a->encoding_width != 20)
        return ls_fail("epipolar gather: encoding_width=%d (0 = no depth encoding, 20 = 10 octaves as shipped)", a->encoding_width);
    if (a->rows > 0 && (!a->xy || !a->image || !a->valid)) return ls_fail("epipolar gather: NULL geometry");
    if (a->encoding_width && a->rows > 0 && !a->depth) return ls_fail("epipolar gather: depth is NULL but encoding_width > 0");
    return 0;
}

static int grid_for(int rows) {
    const int warps_per_block = kThreads / 32;
    const int want = (rows + warps_per_block - 1) / warps_per_block;
    const int cap = 148 * 8;                     // persistent: 8 blocks of 256 threads per SM at most
    return want < cap ? (want < 1 ? 1 : want) : cap;
}

}  // namespace lse

extern "C" LS_API int ls_epipolar_gather_forward(const LsEpipolarGather* a, const float* feat, const float* We, const float* be,
                                                 float* z, void* stream) {
    if (int e = lse::check_args(a)) return e;
    if (a->rows == 0) return 0;
    if (!feat || !z) return ls_fail("epipolar gather forward: NULL feat / z");
    if (a->encoding_width && (!We || !be)) return ls_fail("epipolar gather forward: NULL encoding weight / bias");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = lse::grid_for(a->rows);
    if (a->encoding_width)
        lse::k_gather_fwd<20><<<grid, lse::kThreads, 0, s>>>(feat, a->xy, a->depth, a->image, a->valid, We, be, z, a->rows, a->samples,
                                                             a->height, a->width);
    else
        lse::k_gather_fwd<0><<<grid, lse::kThreads, 0, s>>>(feat, a->xy, a->depth, a->image, a->valid, We, be, z, a->rows, a->samples,
                                                            a->height, a->width);
    return ls_check_cuda("k_gather_fwd");
}

extern "C" LS_API int ls_epipolar_gather_backward(const LsEpipolarGather* a, const float* dz, float* dfeat, float* dWe, float* dbe,
                                                  void* stream) {
    if (int e = lse::check_args(a)) return e;
    if (a->rows == 0) return 0;
    if (!dz || !dfeat) return ls_fail("epipolar gather backward: NULL dz / dfeat");
    if (a->encoding_width && (!dWe || !dbe)) return ls_fail("epipolar gather backward: NULL encoding gradients");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = lse::grid_for(a->rows) / 2 > 148 ? lse::grid_for(a->rows) / 2 : lse::grid_for(a->rows);
    if (a->encoding_width)
        lse::k_gather_bwd<20><<<grid, lse::kThreads, 0, s>>>(dz, a->xy, a->depth, a->image, a->valid, dfeat, dWe, dbe, a->rows,
                                                             a->samples, a->height, a->width);
    else
        lse::k_gather_bwd<0><<<grid, lse::kThreads, 0, s>>>(dz, a->xy, a->depth, a->image, a->valid, dfeat, dWe, dbe, a->rows,
                                                            a->samples, a->height, a->width);
    return ls_check_cuda("k_gather_bwd");
}
