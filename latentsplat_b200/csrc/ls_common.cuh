// ls_common.cuh -- shared device helpers of the sm_100a rasterizer.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ls_raster.h"

namespace ls {

constexpr int kTile = LS_TILE;
constexpr int kTilePixels = kTile * kTile;  // 256 = one CTA
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kAlphaMin = 1.0f / 255.0f;  // [EXT] skip threshold
constexpr float kAlphaMax = 0.99f;          // [EXT] clamp
constexpr float kTMin = 0.0001f;            // [EXT] early stop

__host__ __device__ inline int round_up4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int n_color(int color_mode) { return color_mode == LS_COLOR_NONE ? 0 : 3; }
__host__ __device__ inline bool color_is_sh(int color_mode) { return color_mode == LS_COLOR_SH || color_mode == LS_COLOR_SH_3DGS; }

// ---- exact-order arithmetic (mirrors oracle/raster_oracle.c FMA/MUL/ADD) ----------------
// Everything that decides sort keys, radii and tile rectangles uses these, so the
// compiler can neither contract nor reorder it; the CPU oracle performs the same
// operations in the same order, which is what makes the key lists bit-exact.
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

// transformPoint4x3/4x4 [EXT]: m[k]*x + m[4+k]*y + m[8+k]*z + m[12+k]
__device__ __forceinline__ float xform_row(const float* __restrict__ m, int k, float x, float y, float z) {
    return add_(fma_(m[8 + k], z, fma_(m[k], x, mul_(m[4 + k], y))), m[12 + k]);
}

struct Cov2D {
    float t[3];
    float xmul, ymul;
    float M[2][3];
    float a, b, c;
    float v0[3], v1[3];
};

// EWA projection J R Sigma R^T J^T + 0.3 I   ([EXT] computeCov2D; order = oracle cov2d()).
__device__ __forceinline__ void cov2d(const float p[3], float fx, float fy, float tanx, float tany,
                                      const float cv[6], const float* __restrict__ vm, Cov2D& o) {
    float t0 = xform_row(vm, 0, p[0], p[1], p[2]);
    float t1 = xform_row(vm, 1, p[0], p[1], p[2]);
    const float t2 = xform_row(vm, 2, p[0], p[1], p[2]);
    const float limx = mul_(1.3f, tanx), limy = mul_(1.3f, tany);
    const float txtz = div_(t0, t2), tytz = div_(t1, t2);
    o.xmul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    o.ymul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    t0 = mul_(fminf(limx, fmaxf(-limx, txtz)), t2);
    t1 = mul_(fminf(limy, fmaxf(-limy, tytz)), t2);
    o.t[0] = t0; o.t[1] = t1; o.t[2] = t2;
    const float tz2 = mul_(t2, t2);
    const float J00 = div_(fx, t2), J02 = div_(-mul_(fx, t0), tz2);
    const float J11 = div_(fy, t2), J12 = div_(-mul_(fy, t1), tz2);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        o.M[0][j] = fma_(J02, vm[4 * j + 2], mul_(J00, vm[4 * j + 0]));
        o.M[1][j] = fma_(J12, vm[4 * j + 2], mul_(J11, vm[4 * j + 1]));
    }
    const float S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.v0[k] = fma_(S[k][2], o.M[0][2], fma_(S[k][1], o.M[0][1], mul_(S[k][0], o.M[0][0])));
        o.v1[k] = fma_(S[k][2], o.M[1][2], fma_(S[k][1], o.M[1][1], mul_(S[k][0], o.M[1][0])));
    }
    const float a = fma_(o.M[0][2], o.v0[2], fma_(o.M[0][1], o.v0[1], mul_(o.M[0][0], o.v0[0])));
    const float b = fma_(o.M[1][2], o.v0[2], fma_(o.M[1][1], o.v0[1], mul_(o.M[1][0], o.v0[0])));
    const float c = fma_(o.M[1][2], o.v1[2], fma_(o.M[1][1], o.v1[1], mul_(o.M[1][0], o.v1[0])));
    o.a = add_(a, 0.3f);
    o.b = b;
    o.c = add_(c, 0.3f);
}

// [EXT] getRect, tile rectangle [rmin, rmax) clamped to the grid.
__device__ __forceinline__ void get_rect(float x, float y, int radius, int gx, int gy, int rmin[2], int rmax[2]) {
    const float r = (float)radius;
    rmin[0] = min(gx, max(0, (int)div_(sub_(x, r), (float)kTile)));
    rmin[1] = min(gy, max(0, (int)div_(sub_(y, r), (float)kTile)));
    rmax[0] = min(gx, max(0, (int)div_(add_(add_(x, r), (float)(kTile - 1)), (float)kTile)));
    rmax[1] = min(gy, max(0, (int)div_(add_(add_(y, r), (float)(kTile - 1)), (float)kTile)));
}

// ---- real spherical harmonics, basis of /root/reference/src/misc/sh_utils.py:42-97 -------
// b[k] multiplies coefficient k.  kGrad additionally returns d b[k] / d (x,y,z).
template <bool kGrad>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* __restrict__ b,
                                         float (*__restrict__ db)[3]) {
    b[0] = 0.28209479177387814f;
    if (kGrad) { db[0][0] = db[0][1] = db[0][2] = 0.f; }
    if (deg < 1) return;
    const float c1 = 0.4886025119029199f;
    b[1] = -c1 * x; b[2] = c1 * y; b[3] = -c1 * z;
    if (kGrad) {
        db[1][0] = -c1; db[1][1] = 0.f; db[1][2] = 0.f;
        db[2][0] = 0.f; db[2][1] = c1;  db[2][2] = 0.f;
        db[3][0] = 0.f; db[3][1] = 0.f; db[3][2] = -c1;
    }
    if (deg < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float c20 = 1.0925484305920792f, c22 = 0.31539156525252005f, c24 = 0.5462742152960396f;
    b[4] = c20 * xz; b[5] = -c20 * xy; b[6] = c22 * (2.f * yy - zz - xx); b[7] = -c20 * yz; b[8] = c24 * (zz - xx);
    if (kGrad) {
        db[4][0] = c20 * z;        db[4][1] = 0.f;           db[4][2] = c20 * x;
        db[5][0] = -c20 * y;       db[5][1] = -c20 * x;      db[5][2] = 0.f;
        db[6][0] = -2.f * c22 * x; db[6][1] = 4.f * c22 * y; db[6][2] = -2.f * c22 * z;
        db[7][0] = 0.f;            db[7][1] = -c20 * z;      db[7][2] = -c20 * y;
        db[8][0] = -2.f * c24 * x; db[8][1] = 0.f;           db[8][2] = 2.f * c24 * z;
    }
    if (deg < 3) return;
    const float c30 = -0.5900435899266435f, c31 = 2.890611442640554f, c32 = -0.4570457994644658f,
                c33 = 0.3731763325901154f, c35 = 1.445305721320277f;
    b[9] = c30 * x * (3.f * zz - xx);
    b[10] = c31 * xz * y;
    b[11] = c32 * x * (4.f * yy - zz - xx);
    b[12] = c33 * y * (2.f * yy - 3.f * zz - 3.f * xx);
    b[13] = c32 * z * (4.f * yy - zz - xx);
    b[14] = c35 * z * (zz - xx);
    b[15] = c30 * z * (zz - 3.f * xx);
    if (kGrad) {
        db[9][0] = 3.f * c30 * (zz - xx);            db[9][1] = 0.f;              db[9][2] = 6.f * c30 * xz;
        db[10][0] = c31 * yz;                        db[10][1] = c31 * xz;        db[10][2] = c31 * xy;
        db[11][0] = c32 * (4.f * yy - zz - 3.f * xx); db[11][1] = 8.f * c32 * xy;  db[11][2] = -2.f * c32 * xz;
        db[12][0] = -6.f * c33 * xy; db[12][1] = c33 * (6.f * yy - 3.f * zz - 3.f * xx); db[12][2] = -6.f * c33 * yz;
        db[13][0] = -2.f * c32 * xz; db[13][1] = 8.f * c32 * yz; db[13][2] = c32 * (4.f * yy - 3.f * zz - xx);
        db[14][0] = -2.f * c35 * xz; db[14][1] = 0.f;            db[14][2] = c35 * (3.f * zz - xx);
        db[15][0] = -6.f * c30 * xz; db[15][1] = 0.f;            db[15][2] = 3.f * c30 * (zz - xx);
    }
    if (deg < 4) return;
    const float c40 = 2.5033429417967046f, c41 = -1.7701307697799304f, c42 = 0.9461746957575601f,
                c43 = -0.6690465435572892f, c44 = 0.10578554691520431f, c46 = 0.47308734787878004f,
                c48 = 0.6258357354491761f;
    const float s7y1 = 7.f * yy - 1.f, s7y3 = 7.f * yy - 3.f;
    b[16] = c40 * xz * (zz - xx);
    b[17] = c41 * xy * (3.f * zz - xx);
    b[18] = c42 * xz * s7y1;
    b[19] = c43 * xy * s7y3;
    b[20] = c44 * (yy * (35.f * yy - 30.f) + 3.f);
    b[21] = c43 * yz * s7y3;
    b[22] = c46 * (zz - xx) * s7y1;
    b[23] = c41 * yz * (zz - 3.f * xx);
    b[24] = c48 * (zz * (zz - 3.f * xx) - xx * (3.f * zz - xx));
    if (kGrad) {
        db[16][0] = c40 * z * (zz - 3.f * xx); db[16][1] = 0.f; db[16][2] = c40 * x * (3.f * zz - xx);
        db[17][0] = 3.f * c41 * y * (zz - xx); db[17][1] = c41 * x * (3.f * zz - xx); db[17][2] = 6.f * c41 * xy * z;
        db[18][0] = c42 * z * s7y1; db[18][1] = 14.f * c42 * xz * y; db[18][2] = c42 * x * s7y1;
        db[19][0] = c43 * y * s7y3; db[19][1] = c43 * x * (21.f * yy - 3.f); db[19][2] = 0.f;
        db[20][0] = 0.f; db[20][1] = c44 * y * (140.f * yy - 60.f); db[20][2] = 0.f;
        db[21][0] = 0.f; db[21][1] = c43 * z * (21.f * yy - 3.f); db[21][2] = c43 * y * s7y3;
        db[22][0] = -2.f * c46 * x * s7y1; db[22][1] = 14.f * c46 * (zz - xx) * y; db[22][2] = 2.f * c46 * z * s7y1;
        db[23][0] = -6.f * c41 * xy * z; db[23][1] = c41 * z * (zz - 3.f * xx); db[23][2] = 3.f * c41 * y * (zz - xx);
        db[24][0] = c48 * 4.f * x * (xx - 3.f * zz); db[24][1] = 0.f; db[24][2] = c48 * 4.f * z * (zz - 3.f * xx);
    }
}

// [EXT] stock-3DGS colour SH (graphdeco computeColorFromSH, degrees 0..3) in terms of the basis above: coefficient k of 3DGS at
// direction (x, y, z) is the in-tree polynomial k at (X, Y, Z) = (y, z, x) -- for every k except 14, where the in-tree file has
// z (zz - xx) (sh_utils.py:83) and 3DGS has z (xx - yy) = Y (ZZ - XX).  Call after sh_basis(deg, y, z, x, ...).
template <bool kGrad>
__device__ __forceinline__ void sh_patch_3dgs(int deg, float X, float Y, float Z, float* __restrict__ b, float (*__restrict__ db)[3]) {
    if (deg < 3) return;
    const float c35 = 1.445305721320277f;
    b[14] = c35 * Y * (Z * Z - X * X);
    if (kGrad) { db[14][0] = -2.f * c35 * X * Y; db[14][1] = c35 * (Z * Z - X * X); db[14][2] = 2.f * c35 * Y * Z; }
}

// d/d(x,y,z) of sum_k sg[k] * basis_k(x, y, z): the derivative table of sh_basis contracted on the fly (no 25 x 3 array),
// for callers that hold sg[k] = sum_c sh[k][c] * dL/dvalue[c] in registers.  Generated from sh_basis above (same expressions).
template <int DEG>
__device__ __forceinline__ void sh_basis_vjp(float x, float y, float z, const float* __restrict__ sg, float* __restrict__ ddir) {
    if (DEG < 1) return;
    const float c1 = 0.4886025119029199f;
    ddir[0] = fmaf(-c1, sg[1], ddir[0]);
    ddir[1] = fmaf(c1, sg[2], ddir[1]);
    ddir[2] = fmaf(-c1, sg[3], ddir[2]);
    if (DEG < 2) return;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float c20 = 1.0925484305920792f, c22 = 0.31539156525252005f, c24 = 0.5462742152960396f;
    ddir[0] = fmaf(c20 * z, sg[4], ddir[0]);
    ddir[2] = fmaf(c20 * x, sg[4], ddir[2]);
    ddir[0] = fmaf(-c20 * y, sg[5], ddir[0]);
    ddir[1] = fmaf(-c20 * x, sg[5], ddir[1]);
    ddir[0] = fmaf(-2.f * c22 * x, sg[6], ddir[0]);
    ddir[1] = fmaf(4.f * c22 * y, sg[6], ddir[1]);
    ddir[2] = fmaf(-2.f * c22 * z, sg[6], ddir[2]);
    ddir[1] = fmaf(-c20 * z, sg[7], ddir[1]);
    ddir[2] = fmaf(-c20 * y, sg[7], ddir[2]);
    ddir[0] = fmaf(-2.f * c24 * x, sg[8], ddir[0]);
    ddir[2] = fmaf(2.f * c24 * z, sg[8], ddir[2]);
    if (DEG < 3) return;
    const float c30 = -0.5900435899266435f, c31 = 2.890611442640554f, c32 = -0.4570457994644658f,
                c33 = 0.3731763325901154f, c35 = 1.445305721320277f;
    ddir[0] = fmaf(3.f * c30 * (zz - xx), sg[9], ddir[0]);
    ddir[2] = fmaf(6.f * c30 * xz, sg[9], ddir[2]);
    ddir[0] = fmaf(c31 * yz, sg[10], ddir[0]);
    ddir[1] = fmaf(c31 * xz, sg[10], ddir[1]);
    ddir[2] = fmaf(c31 * xy, sg[10], ddir[2]);
    ddir[0] = fmaf(c32 * (4.f * yy - zz - 3.f * xx), sg[11], ddir[0]);
    ddir[1] = fmaf(8.f * c32 * xy, sg[11], ddir[1]);
    ddir[2] = fmaf(-2.f * c32 * xz, sg[11], ddir[2]);
    ddir[0] = fmaf(-6.f * c33 * xy, sg[12], ddir[0]);
    ddir[1] = fmaf(c33 * (6.f * yy - 3.f * zz - 3.f * xx), sg[12], ddir[1]);
    ddir[2] = fmaf(-6.f * c33 * yz, sg[12], ddir[2]);
    ddir[0] = fmaf(-2.f * c32 * xz, sg[13], ddir[0]);
    ddir[1] = fmaf(8.f * c32 * yz, sg[13], ddir[1]);
    ddir[2] = fmaf(c32 * (4.f * yy - 3.f * zz - xx), sg[13], ddir[2]);
    ddir[0] = fmaf(-2.f * c35 * xz, sg[14], ddir[0]);
    ddir[2] = fmaf(c35 * (3.f * zz - xx), sg[14], ddir[2]);
    ddir[0] = fmaf(-6.f * c30 * xz, sg[15], ddir[0]);
    ddir[2] = fmaf(3.f * c30 * (zz - xx), sg[15], ddir[2]);
    if (DEG < 4) return;
    const float c40 = 2.5033429417967046f, c41 = -1.7701307697799304f, c42 = 0.9461746957575601f,
                c43 = -0.6690465435572892f, c44 = 0.10578554691520431f, c46 = 0.47308734787878004f,
                c48 = 0.6258357354491761f;
    const float s7y1 = 7.f * yy - 1.f, s7y3 = 7.f * yy - 3.f;
    ddir[0] = fmaf(c40 * z * (zz - 3.f * xx), sg[16], ddir[0]);
    ddir[2] = fmaf(c40 * x * (3.f * zz - xx), sg[16], ddir[2]);
    ddir[0] = fmaf(3.f * c41 * y * (zz - xx), sg[17], ddir[0]);
    ddir[1] = fmaf(c41 * x * (3.f * zz - xx), sg[17], ddir[1]);
    ddir[2] = fmaf(6.f * c41 * xy * z, sg[17], ddir[2]);
    ddir[0] = fmaf(c42 * z * s7y1, sg[18], ddir[0]);
    ddir[1] = fmaf(14.f * c42 * xz * y, sg[18], ddir[1]);
    ddir[2] = fmaf(c42 * x * s7y1, sg[18], ddir[2]);
    ddir[0] = fmaf(c43 * y * s7y3, sg[19], ddir[0]);
    ddir[1] = fmaf(c43 * x * (21.f * yy - 3.f), sg[19], ddir[1]);
    ddir[1] = fmaf(c44 * y * (140.f * yy - 60.f), sg[20], ddir[1]);
    ddir[1] = fmaf(c43 * z * (21.f * yy - 3.f), sg[21], ddir[1]);
    ddir[2] = fmaf(c43 * y * s7y3, sg[21], ddir[2]);
    ddir[0] = fmaf(-2.f * c46 * x * s7y1, sg[22], ddir[0]);
    ddir[1] = fmaf(14.f * c46 * (zz - xx) * y, sg[22], ddir[1]);
    ddir[2] = fmaf(2.f * c46 * z * s7y1, sg[22], ddir[2]);
    ddir[0] = fmaf(-6.f * c41 * xy * z, sg[23], ddir[0]);
    ddir[1] = fmaf(c41 * z * (zz - 3.f * xx), sg[23], ddir[1]);
    ddir[2] = fmaf(3.f * c41 * y * (zz - xx), sg[23], ddir[2]);
    ddir[0] = fmaf(c48 * 4.f * x * (xx - 3.f * zz), sg[24], ddir[0]);
    ddir[2] = fmaf(c48 * 4.f * z * (zz - 3.f * xx), sg[24], ddir[2]);
}

// geometry record slot 7: two fp16 half-extents (x, y) of the alpha >= 1/255 region, rounded up
__device__ __forceinline__ float2 unpack_extent(float packed) {
    const uint32_t u = __float_as_uint(packed);
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}

// ---- small PTX wrappers ------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier + 1-D bulk TMA copies (cp.async.bulk): staging of the sorted per-tile queues ----------------------------
constexpr int kBatch = 256;                         // queue entries staged per round by the blend kernels
// shared memory of a blend kernel: 2 x kBatch records of `rv` float4 + 2 x kBatch cull records + 8 per-warp survivor lists
__host__ __device__ constexpr int blend_smem_bytes(int rv) { return 2 * kBatch * rv * 16 + 2 * kBatch * 16 + 8 * kBatch; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();           // a protocol bug must fail loudly, never hang the device
    }
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global, plain store or fp32 add (cp.reduce): `bytes` a multiple of 16; completion through the bulk async-group
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp-cooperative staging of per-Gaussian coefficient rows ---------------------------------------------------
// A thread per Gaussian reading its own 300-B SH row makes every warp-level load touch 32 different sectors for 128
// useful bytes (the preprocess kernels then sit at 12-35 % of HBM bandwidth, L2->L1 sector traffic bound).  Instead the
// warp moves a chunk of `len` <= kStagePitch consecutive floats of each of its 32 consecutive rows with lanes running
// over the FLAT (row, element) index -- consecutive lanes touch consecutive addresses inside a row piece -- through a
// small per-warp shared buffer (32 x 25 floats; pitch 25 is odd, so the per-lane row reads are conflict-free).
constexpr int kStagePitch = 25;
constexpr int kChunk = 24;       // floats per row and pass of the backward SH staging: 8 colour coefficients = 3 whole sectors

__device__ __forceinline__ void stage_load(float* __restrict__ buf, const float* __restrict__ base, int row_len, int c0,
                                           int len, int nrows, int lane) {
    const uint32_t magic = (65536u + (uint32_t)len - 1u) / (uint32_t)len;     // f / len for f < 800, len <= 25
    const int total = 32 * len;
    for (int f = lane; f < total; f += 32) {
        const int g = (int)(((uint32_t)f * magic) >> 16), e = f - g * len;
        if (g < nrows) buf[g * kStagePitch + e] = __ldg(base + (size_t)g * row_len + c0 + e);
    }
    __syncwarp();
}

// rows whose bit is clear in row_mask are left untouched (their destination was zero-filled by the caller)
__device__ __forceinline__ void stage_store(const float* __restrict__ buf, float* __restrict__ base, int row_len, int c0,
                                            int len, uint32_t row_mask, bool atomic, int lane) {
    __syncwarp();
    const uint32_t magic = (65536u + (uint32_t)len - 1u) / (uint32_t)len;
    const int total = 32 * len;
    for (int f = lane; f < total; f += 32) {
        const int g = (int)(((uint32_t)f * magic) >> 16), e = f - g * len;
        if ((row_mask >> g) & 1u) {
            float* dst = base + (size_t)g * row_len + c0 + e;
            const float v = buf[g * kStagePitch + e];
            if (atomic) atomicAdd(dst, v); else *dst = v;
        }
    }
    __syncwarp();
}

}  // namespace ls
