/*
 * raster_oracle.c -- CPU restatement of the tile-based differentiable Gaussian
 * rasterizer that latentSplat calls at
 *     /root/reference/src/model/decoder/cuda_splatting.py:132-158
 * (`diff_gaussian_rasterization.GaussianRasterizer`).
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product
 * (latentsplat_b200/) never imports, links or falls back to anything here.
 *
 * PARITY UNPINNED.  The rasterizer's source is a third-party, un-pinned pip
 * dependency (git+https://github.com/Chrixtar/latent-gaussian-rasterization,
 * requirements.txt:34) that is not vendored in /root/reference and is not on
 * this machine; the reference holds no tests or golden vectors for it
 * (SURVEY.md section 8c).  What is restated here is therefore
 *   (1) the call contract visible in the reference:
 *         cuda_splatting.py:115-118  transposed (column-major) view / full
 *                                    projection matrices,
 *         cuda_splatting.py:19-46    projection maps z to [0,1], w = z_view,
 *         cuda_splatting.py:148,157  cov3D_precomp = upper triangle
 *                                    (00,01,02,11,12,22),
 *         cuda_splatting.py:91       shs laid out (G, n_coeff, 3),
 *         cuda_splatting.py:97-101   features arrive pre-evaluated (0.5+SH),
 *         cuda_splatting.py:150-166  outputs (image|None, feature_map|None,
 *                                    mask(1,H,W), depth(1,H,W), _),
 *         decoder_splatting_cuda.py:45-47  feature background is 0 and the
 *                                    mask is the accumulated alpha;
 *   (2) the SH polynomials of the in-tree src/misc/sh_utils.py:42-97 (degree <= 4), pinned by
 *       tests/golden/sh_eval.npz generated from that file.  The reference applies that function to the
 *       FEATURES only (cuda_splatting.py:97); colour `shs` go to the fork's CUDA, so WHICH basis the colours
 *       use inside the rasterizer is [EXT]: in-tree order by default, stock-3DGS order (the same polynomials
 *       at (y, z, x), k = 14 patched) through oracle_set_color_sh_basis(1) -- see g_color_sh_3dgs below;
 *   (3) the published algorithm of the lineage the fork descends from
 *       (graphdeco-inria/diff-gaussian-rasterization: 16x16 tiles, near cull
 *       z_view <= 0.2, tan-fov clamp 1.3, +0.3 px^2 dilation, 3-sigma radius
 *       from max eigenvalue with max(0.1, .) guard, alpha = min(0.99, o*G),
 *       skip alpha < 1/255, stop before T < 1e-4, 64-bit (tile | depth-bits)
 *       keys radix-sorted stably, backward = back-to-front replay).
 *   Items under (3) are marked [EXT] below: they cannot be verified here.
 *
 * Floating point: every operation of the per-Gaussian preprocess that feeds
 * the sort keys / tile rectangles is written with an explicit order and
 * explicit fused multiply-adds (FMA/MUL/ADD macros); the CUDA kernels use the
 * same order with __fmaf_rn/__fmul_rn/__fadd_rn, so keys, radii, rectangles
 * and sorted lists are bit-exact by construction.  Compile with
 * -ffp-contract=off so the compiler adds no contractions of its own.
 * The order chosen imitates nvcc's default contraction of the upstream
 * expressions ( a*b + c*d -> fma(a,b,c*d) ) -- [EXT], unverifiable.
 *
 * Build: see oracle/Makefile.   -DORACLE_DOUBLE builds a float64 twin used
 * only for finite-difference gradient checks.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_DOUBLE
typedef double real;
#define FMA(a, b, c) fma((a), (b), (c))
#define R(x) x
#define SQRT sqrt
#define EXP exp
#define CEIL ceil
#define FN(name) name##_f64
#else
typedef float real;
#define FMA(a, b, c) fmaf((a), (b), (c))
#define R(x) x##f
#define SQRT sqrtf
#define EXP expf
#define CEIL ceilf
#define FN(name) name##_f32
#endif
#define MUL(a, b) ((real)((a) * (b)))
#define ADD(a, b) ((real)((a) + (b)))
#define SUB(a, b) ((real)((a) - (b)))

#define TILE 16 /* [EXT] BLOCK_X = BLOCK_Y = 16 */

/* ------------------------------------------------------------------ */
/* SH basis: restates /root/reference/src/misc/sh_utils.py:9-39 (constants)
 * and :42-97 (polynomials).  basis[k] multiplies sh[..., k].            */
static const double SH_C0 = 0.28209479177387814;
static const double SH_C1 = 0.4886025119029199;
static const double SH_C2[5] = {1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                                -1.0925484305920792, 0.5462742152960396};
static const double SH_C3[7] = {-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
                                0.3731763325901154,  -0.4570457994644658, 1.445305721320277,
                                -0.5900435899266435};
static const double SH_C4[9] = {2.5033429417967046,  -1.7701307697799304, 0.9461746957575601,
                                -0.6690465435572892, 0.10578554691520431, -0.6690465435572892,
                                0.47308734787878004, -1.7701307697799304, 0.6258357354491761};

/* basis values b[0..n) and, if db != NULL, d basis / d (x,y,z) in db[k][3]. */
static void sh_basis(int deg, real x, real y, real z, real *b, real (*db)[3]) {
    b[0] = (real)SH_C0;
    if (db) { db[0][0] = db[0][1] = db[0][2] = 0; }
    if (deg < 1) return;
    const real c1 = (real)SH_C1;
    b[1] = -c1 * x; b[2] = c1 * y; b[3] = -c1 * z;          /* sh_utils.py:62-65 */
    if (db) {
        db[1][0] = -c1; db[1][1] = 0;  db[1][2] = 0;
        db[2][0] = 0;   db[2][1] = c1; db[2][2] = 0;
        db[3][0] = 0;   db[3][1] = 0;  db[3][2] = -c1;
    }
    if (deg < 2) return;
    const real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const real c20 = (real)SH_C2[0], c21 = (real)SH_C2[1], c22 = (real)SH_C2[2],
               c23 = (real)SH_C2[3], c24 = (real)SH_C2[4];
    b[4] = c20 * xz; b[5] = c21 * xy; b[6] = c22 * (R(2.0) * yy - zz - xx);   /* :70-75 */
    b[7] = c23 * yz; b[8] = c24 * (zz - xx);
    if (db) {
        db[4][0] = c20 * z;  db[4][1] = 0;        db[4][2] = c20 * x;
        db[5][0] = c21 * y;  db[5][1] = c21 * x;  db[5][2] = 0;
        db[6][0] = c22 * R(-2.0) * x; db[6][1] = c22 * R(4.0) * y; db[6][2] = c22 * R(-2.0) * z;
        db[7][0] = 0;        db[7][1] = c23 * z;  db[7][2] = c23 * y;
        db[8][0] = c24 * R(-2.0) * x; db[8][1] = 0; db[8][2] = c24 * R(2.0) * z;
    }
    if (deg < 3) return;
    const real c30 = (real)SH_C3[0], c31 = (real)SH_C3[1], c32 = (real)SH_C3[2], c33 = (real)SH_C3[3],
               c34 = (real)SH_C3[4], c35 = (real)SH_C3[5], c36 = (real)SH_C3[6];
    b[9]  = c30 * x * (R(3.0) * zz - xx);                                      /* :78-85 */
    b[10] = c31 * xz * y;
    b[11] = c32 * x * (R(4.0) * yy - zz - xx);
    b[12] = c33 * y * (R(2.0) * yy - R(3.0) * zz - R(3.0) * xx);
    b[13] = c34 * z * (R(4.0) * yy - zz - xx);
    b[14] = c35 * z * (zz - xx);
    b[15] = c36 * z * (zz - R(3.0) * xx);
    if (db) {
        db[9][0]  = c30 * (R(3.0) * zz - R(3.0) * xx); db[9][1] = 0; db[9][2] = c30 * R(6.0) * xz;
        db[10][0] = c31 * yz; db[10][1] = c31 * xz; db[10][2] = c31 * xy;
        db[11][0] = c32 * (R(4.0) * yy - zz - R(3.0) * xx); db[11][1] = c32 * R(8.0) * xy;
        db[11][2] = c32 * R(-2.0) * xz;
        db[12][0] = c33 * R(-6.0) * xy; db[12][1] = c33 * (R(6.0) * yy - R(3.0) * zz - R(3.0) * xx);
        db[12][2] = c33 * R(-6.0) * yz;
        db[13][0] = c34 * R(-2.0) * xz; db[13][1] = c34 * R(8.0) * yz;
        db[13][2] = c34 * (R(4.0) * yy - R(3.0) * zz - xx);
        db[14][0] = c35 * R(-2.0) * xz; db[14][1] = 0; db[14][2] = c35 * (R(3.0) * zz - xx);
        db[15][0] = c36 * R(-6.0) * xz; db[15][1] = 0; db[15][2] = c36 * (R(3.0) * zz - R(3.0) * xx);
    }
    if (deg < 4) return;
    const real c40 = (real)SH_C4[0], c41 = (real)SH_C4[1], c42 = (real)SH_C4[2], c43 = (real)SH_C4[3],
               c44 = (real)SH_C4[4], c45 = (real)SH_C4[5], c46 = (real)SH_C4[6], c47 = (real)SH_C4[7],
               c48 = (real)SH_C4[8];
    b[16] = c40 * xz * (zz - xx);                                              /* :88-96 */
    b[17] = c41 * xy * (R(3.0) * zz - xx);
    b[18] = c42 * xz * (R(7.0) * yy - R(1.0));
    b[19] = c43 * xy * (R(7.0) * yy - R(3.0));
    b[20] = c44 * (yy * (R(35.0) * yy - R(30.0)) + R(3.0));
    b[21] = c45 * yz * (R(7.0) * yy - R(3.0));
    b[22] = c46 * (zz - xx) * (R(7.0) * yy - R(1.0));
    b[23] = c47 * yz * (zz - R(3.0) * xx);
    b[24] = c48 * (zz * (zz - R(3.0) * xx) - xx * (R(3.0) * zz - xx));
    if (db) {
        db[16][0] = c40 * z * (zz - R(3.0) * xx); db[16][1] = 0; db[16][2] = c40 * x * (R(3.0) * zz - xx);
        db[17][0] = c41 * y * (R(3.0) * zz - R(3.0) * xx); db[17][1] = c41 * x * (R(3.0) * zz - xx);
        db[17][2] = c41 * R(6.0) * xy * z;
        db[18][0] = c42 * z * (R(7.0) * yy - R(1.0)); db[18][1] = c42 * R(14.0) * xz * y;
        db[18][2] = c42 * x * (R(7.0) * yy - R(1.0));
        db[19][0] = c43 * y * (R(7.0) * yy - R(3.0)); db[19][1] = c43 * x * (R(21.0) * yy - R(3.0));
        db[19][2] = 0;
        db[20][0] = 0; db[20][1] = c44 * y * (R(140.0) * yy - R(60.0)); db[20][2] = 0;
        db[21][0] = 0; db[21][1] = c45 * z * (R(21.0) * yy - R(3.0)); db[21][2] = c45 * y * (R(7.0) * yy - R(3.0));
        db[22][0] = c46 * R(-2.0) * x * (R(7.0) * yy - R(1.0)); db[22][1] = c46 * (zz - xx) * R(14.0) * y;
        db[22][2] = c46 * R(2.0) * z * (R(7.0) * yy - R(1.0));
        db[23][0] = c47 * R(-6.0) * xy * z; db[23][1] = c47 * z * (zz - R(3.0) * xx);
        db[23][2] = c47 * y * (R(3.0) * zz - R(3.0) * xx);
        /* b24 = c48 * (z^4 - 6 x^2 z^2 + x^4) */
        db[24][0] = c48 * (R(-12.0) * x * zz + R(4.0) * x * xx); db[24][1] = 0;
        db[24][2] = c48 * (R(4.0) * z * zz - R(12.0) * xx * z);
    }
}

/* [EXT] Which convention the colour `shs` use inside the rasterizer: 0 = the in-tree basis above (default), 1 = the
 * coefficient order of graphdeco-inria/diff-gaussian-rasterization (b1..3 = -C1 y, +C1 z, -C1 x; ...), see sh_patch_3dgs.
 * The fork's CUDA is not in /root/reference, so neither can be confirmed; the product
 * exposes the same switch (LS_COLOR_SH / LS_COLOR_SH_3DGS, env LS_SH_BASIS).  Features always use the in-tree basis
 * (cuda_splatting.py:97 calls the in-tree eval_sh).                                                                     */
static int g_color_sh_3dgs = 0;
/* 3DGS coefficient k at (x, y, z) = in-tree polynomial k at (X, Y, Z) = (y, z, x) for every k of degrees 0..3 except k = 14:
 * sh_utils.py:83 has z (zz - xx) there, 3DGS z (xx - yy) = Y (ZZ - XX).  Degree 4 is not defined by 3DGS; the permuted
 * in-tree polynomials are used (an extension).  Apply after sh_basis(deg, y, z, x, ...).                                */
static void sh_patch_3dgs(int deg, real X, real Y, real Z, real *b, real (*db)[3]) {
    if (deg < 3) return;
    const real c35 = (real)SH_C3[5];
    b[14] = c35 * Y * (Z * Z - X * X);
    if (db) { db[14][0] = R(-2.0) * c35 * X * Y; db[14][1] = c35 * (Z * Z - X * X); db[14][2] = R(2.0) * c35 * Y * Z; }
}
void FN(oracle_set_color_sh_basis)(int use_3dgs) { g_color_sh_3dgs = use_3dgs != 0; }

/* Evaluate n_ch channels of SH (layout sh[k*n_ch + c], i.e. (n_coeff, n_ch)) at
 * unit direction d; result[c] = sum_k basis_k * sh[k][c].  Test helper for the
 * golden SH vectors; also used by the colour path below.                     */
void FN(oracle_sh_eval)(int deg, int n_ch, const real *sh, const real *dir, real *out) {
    real b[25];
    sh_basis(deg, dir[0], dir[1], dir[2], b, NULL);
    const int n = (deg + 1) * (deg + 1);
    for (int c = 0; c < n_ch; ++c) {
        real r = 0;
        for (int k = 0; k < n; ++k) r += b[k] * sh[k * n_ch + c];
        out[c] = r;
    }
}

/* Same as oracle_sh_eval in the 3DGS coefficient order ([EXT] switch above); dir is the true (x, y, z). */
void FN(oracle_sh_eval_3dgs)(int deg, int n_ch, const real *sh, const real *dir, real *out) {
    real b[25];
    sh_basis(deg, dir[1], dir[2], dir[0], b, NULL);
    sh_patch_3dgs(deg, dir[1], dir[2], dir[0], b, NULL);
    const int n = (deg + 1) * (deg + 1);
    for (int c = 0; c < n_ch; ++c) {
        real r = 0;
        for (int k = 0; k < n; ++k) r += b[k] * sh[k * n_ch + c];
        out[c] = r;
    }
}

/* ------------------------------------------------------------------ */
typedef struct {
    /* sizes */
    int G, H, W, C;           /* C = feature channels (0 if features == NULL)       */
    int sh_degree;            /* colour SH degree, used when shs != NULL             */
    /* per-Gaussian inputs (one view)                                              */
    const real *means3D;      /* (G,3)                                              */
    const real *cov3D;        /* (G,6) upper triangle 00 01 02 11 12 22             */
    const real *opacity;      /* (G)                                                */
    const real *shs;          /* (G, n_coeff, 3) or NULL                            */
    const real *colors_precomp; /* (G,3) or NULL                                    */
    const real *features;     /* (G,C) or NULL                                      */
    /* camera                                                                      */
    const real *viewmatrix;   /* 16, transposed world->view as handed over by
                                 cuda_splatting.py:116 (m[4*c + r] = V[r][c])       */
    const real *projmatrix;   /* 16, transposed full projection (:117)              */
    const real *campos;       /* 3                                                  */
    real tanfovx, tanfovy;
    const real *bg;           /* 3                                                  */
    real scene_scale;         /* means *= s, cov *= s*s (cuda_splatting.py:75-82);
                                 1 for the plain GaussianRasterizer call           */
} OracleIn;

typedef struct {
    /* per-Gaussian state (caller allocates G of each)                             */
    real *depths;             /* (G)   view-space z                                 */
    real *xy;                 /* (G,2) pixel-space mean                             */
    real *conic_opacity;      /* (G,4)                                              */
    int32_t *radii;           /* (G)                                                */
    uint32_t *tiles_touched;  /* (G)                                                */
    real *rgb;                /* (G,3) colour used for blending                     */
    uint8_t *clamped;         /* (G,3)                                              */
    /* binning: allocated by the oracle (malloc), freed with oracle_free            */
    int64_t num_rendered;
    uint64_t *keys_sorted;    /* (num_rendered) tile<<32 | depth bits               */
    uint32_t *point_list;     /* (num_rendered) Gaussian ids in blend order         */
    uint32_t *ranges;         /* (tiles,2) [start,end)                              */
    /* images (caller allocates)                                                    */
    real *out_color;          /* (3,H,W) or NULL                                    */
    real *out_feature;        /* (C,H,W) or NULL                                    */
    real *out_alpha;          /* (H,W)                                              */
    real *out_depth;          /* (H,W)                                              */
    real *final_T;            /* (H,W)                                              */
    uint32_t *n_contrib;      /* (H,W)                                              */
    /* decision margins for the parity tests: per pixel, the largest blend term
       whose keep/skip decision (power>0, alpha<1/255, T<1e-4) lies within
       `margin_eps` relative of its threshold.  A GPU exp that differs in the last
       ulps may flip exactly those terms.                                          */
    real *flip_bound;         /* (H,W) or NULL                                      */
    real margin_eps;
    uint8_t *marginal;        /* (G) or NULL: 1 where a marginal decision concerned
                                 this Gaussian at some pixel                        */
} OracleOut;

static inline void xform4x3(const real *m, const real *p, real *o) {
    /* [EXT] transformPoint4x3: m[k]*x + m[4+k]*y + m[8+k]*z + m[12+k] */
    for (int k = 0; k < 3; ++k)
        o[k] = ADD(FMA(m[8 + k], p[2], FMA(m[k], p[0], MUL(m[4 + k], p[1]))), m[12 + k]);
}
static inline void xform4x4(const real *m, const real *p, real *o) {
    for (int k = 0; k < 4; ++k)
        o[k] = ADD(FMA(m[8 + k], p[2], FMA(m[k], p[0], MUL(m[4 + k], p[1]))), m[12 + k]);
}

/* EWA projection of the 3D covariance.  Returns M (2x3) = J*R and Sigma*M^T
 * products for the backward.  [EXT] computeCov2D.                            */
typedef struct {
    real t[3];          /* view-space mean with clamped x,y                       */
    real xmul, ymul;    /* 0 when the tan-fov clamp is active (backward)          */
    real M[2][3];
    real a, b, c;       /* cov2D incl. +0.3 dilation                              */
    real v0[3], v1[3];  /* Sigma * M0^T, Sigma * M1^T                             */
} Cov2D;

static void cov2d(const real *p, real fx, real fy, real tanx, real tany, const real *cv,
                  const real *vm, Cov2D *o) {
    real t[3];
    xform4x3(vm, p, t);
    const real limx = MUL(R(1.3), tanx), limy = MUL(R(1.3), tany);
    const real txtz = t[0] / t[2], tytz = t[1] / t[2];
    o->xmul = (txtz < -limx || txtz > limx) ? R(0.0) : R(1.0);
    o->ymul = (tytz < -limy || tytz > limy) ? R(0.0) : R(1.0);
    t[0] = MUL(fmin(limx, fmax(-limx, txtz)), t[2]);
    t[1] = MUL(fmin(limy, fmax(-limy, tytz)), t[2]);
    o->t[0] = t[0]; o->t[1] = t[1]; o->t[2] = t[2];
    const real tz2 = MUL(t[2], t[2]);
    const real J00 = fx / t[2], J02 = -MUL(fx, t[0]) / tz2;
    const real J11 = fy / t[2], J12 = -MUL(fy, t[1]) / tz2;
    /* R[r][c] = vm[4*c + r];  M = J * R */
    for (int j = 0; j < 3; ++j) {
        o->M[0][j] = FMA(J02, vm[4 * j + 2], MUL(J00, vm[4 * j + 0]));
        o->M[1][j] = FMA(J12, vm[4 * j + 2], MUL(J11, vm[4 * j + 1]));
    }
    const real S[3][3] = {{cv[0], cv[1], cv[2]}, {cv[1], cv[3], cv[4]}, {cv[2], cv[4], cv[5]}};
    for (int k = 0; k < 3; ++k) {
        o->v0[k] = FMA(S[k][2], o->M[0][2], FMA(S[k][1], o->M[0][1], MUL(S[k][0], o->M[0][0])));
        o->v1[k] = FMA(S[k][2], o->M[1][2], FMA(S[k][1], o->M[1][1], MUL(S[k][0], o->M[1][0])));
    }
    real a = FMA(o->M[0][2], o->v0[2], FMA(o->M[0][1], o->v0[1], MUL(o->M[0][0], o->v0[0])));
    real b = FMA(o->M[1][2], o->v0[2], FMA(o->M[1][1], o->v0[1], MUL(o->M[1][0], o->v0[0])));
    real c = FMA(o->M[1][2], o->v1[2], FMA(o->M[1][1], o->v1[1], MUL(o->M[1][0], o->v1[0])));
    o->a = ADD(a, R(0.3)); /* [EXT] low-pass dilation */
    o->b = b;
    o->c = ADD(c, R(0.3));
}

static inline void get_rect(const real *xy, int radius, int gx, int gy, int *rmin, int *rmax) {
    /* [EXT] getRect */
    const real r = (real)radius;
    int v;
    v = (int)(SUB(xy[0], r) / (real)TILE);                       rmin[0] = v < 0 ? 0 : (v > gx ? gx : v);
    v = (int)(SUB(xy[1], r) / (real)TILE);                       rmin[1] = v < 0 ? 0 : (v > gy ? gy : v);
    v = (int)(ADD(ADD(xy[0], r), (real)(TILE - 1)) / (real)TILE); rmax[0] = v < 0 ? 0 : (v > gx ? gx : v);
    v = (int)(ADD(ADD(xy[1], r), (real)(TILE - 1)) / (real)TILE); rmax[1] = v < 0 ? 0 : (v > gy ? gy : v);
}

static inline uint32_t depth_bits(real d) {
    float f = (float)d; /* keys always carry fp32 bit patterns */
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

typedef struct { uint64_t key; uint32_t id; } KeyId;
static int cmp_keyid(const void *a, const void *b) {
    const KeyId *x = (const KeyId *)a, *y = (const KeyId *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    /* stable radix sort of keys emitted in ascending id order == ties by id */
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}

/* ---------------- forward ---------------- */
int FN(oracle_forward)(const OracleIn *in, OracleOut *out, int n_threads) {
    const int G = in->G, H = in->H, W = in->W, C = in->features ? in->C : 0;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const real fx = (real)W / MUL(R(2.0), in->tanfovx), fy = (real)H / MUL(R(2.0), in->tanfovy);
    const int has_color = in->shs || in->colors_precomp;
    const real s = in->scene_scale, s2 = MUL(s, s);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif

    /* R.1 preprocess */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        out->radii[i] = 0;
        out->tiles_touched[i] = 0;
        out->depths[i] = 0; out->xy[2 * i] = out->xy[2 * i + 1] = 0;
        for (int k = 0; k < 4; ++k) out->conic_opacity[4 * i + k] = 0;
        for (int k = 0; k < 3; ++k) { out->rgb[3 * i + k] = 0; out->clamped[3 * i + k] = 0; }
        const real p[3] = {MUL(in->means3D[3 * i], s), MUL(in->means3D[3 * i + 1], s), MUL(in->means3D[3 * i + 2], s)};
        real pv[3];
        xform4x3(in->viewmatrix, p, pv);
        if (pv[2] <= R(0.2)) continue; /* [EXT] near cull */
        real ph[4];
        xform4x4(in->projmatrix, p, ph);
        const real pw = R(1.0) / ADD(ph[3], R(0.0000001));
        const real pp[2] = {MUL(ph[0], pw), MUL(ph[1], pw)};
        real cv[6];
        for (int k = 0; k < 6; ++k) cv[k] = MUL(in->cov3D[6 * i + k], s2);
        Cov2D q;
        cov2d(p, fx, fy, in->tanfovx, in->tanfovy, cv, in->viewmatrix, &q);
        const real det = SUB(MUL(q.a, q.c), MUL(q.b, q.b));
        if (det == R(0.0)) continue;
        const real det_inv = R(1.0) / det;
        const real conic[3] = {MUL(q.c, det_inv), MUL(-q.b, det_inv), MUL(q.a, det_inv)};
        const real mid = MUL(R(0.5), ADD(q.a, q.c));
        const real disc = SQRT(fmax(R(0.1), SUB(MUL(mid, mid), det)));
        const real l1 = ADD(mid, disc), l2 = SUB(mid, disc);
        const int radius = (int)CEIL(MUL(R(3.0), SQRT(fmax(l1, l2))));
        /* ndc2Pix: ((v + 1) * S - 1) * 0.5 */
        const real px = MUL(SUB(MUL(ADD(pp[0], R(1.0)), (real)W), R(1.0)), R(0.5));
        const real py = MUL(SUB(MUL(ADD(pp[1], R(1.0)), (real)H), R(1.0)), R(0.5));
        const real pxy[2] = {px, py};
        int rmin[2], rmax[2];
        get_rect(pxy, radius, gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (in->shs) {
            /* colour from SH, +0.5, clamp at 0 ([EXT] computeColorFromSH with the
               in-tree basis).  direction = normalize(p - campos).                 */
            real d[3] = {p[0] - in->campos[0], p[1] - in->campos[1], p[2] - in->campos[2]};
            const real inv = R(1.0) / SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            d[0] *= inv; d[1] *= inv; d[2] *= inv;
            const int n = (in->sh_degree + 1) * (in->sh_degree + 1);
            real col[3];
            if (g_color_sh_3dgs) {
                real bas[25];
                sh_basis(in->sh_degree, d[1], d[2], d[0], bas, NULL);
                sh_patch_3dgs(in->sh_degree, d[1], d[2], d[0], bas, NULL);
                for (int c = 0; c < 3; ++c) {
                    real acc = 0;
                    for (int k = 0; k < n; ++k) acc += bas[k] * in->shs[((size_t)i * n + k) * 3 + c];
                    col[c] = acc;
                }
            } else {
                FN(oracle_sh_eval)(in->sh_degree, 3, in->shs + (size_t)i * n * 3, d, col);
            }
            for (int k = 0; k < 3; ++k) {
                col[k] += R(0.5);
                out->clamped[3 * i + k] = col[k] < 0;
                out->rgb[3 * i + k] = col[k] < 0 ? 0 : col[k];
            }
        } else if (in->colors_precomp) {
            for (int k = 0; k < 3; ++k) out->rgb[3 * i + k] = in->colors_precomp[3 * i + k];
        }
        out->depths[i] = pv[2];
        out->radii[i] = radius;
        out->xy[2 * i] = px; out->xy[2 * i + 1] = py;
        out->conic_opacity[4 * i + 0] = conic[0];
        out->conic_opacity[4 * i + 1] = conic[1];
        out->conic_opacity[4 * i + 2] = conic[2];
        out->conic_opacity[4 * i + 3] = in->opacity[i];
        out->tiles_touched[i] = (uint32_t)((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]));
    }

    /* R.2-R.5 duplicate with keys, sort, tile ranges */
    int64_t n = 0;
    for (int i = 0; i < G; ++i) n += out->tiles_touched[i];
    out->num_rendered = n;
    KeyId *kv = (KeyId *)malloc(sizeof(KeyId) * (size_t)(n > 0 ? n : 1));
    int64_t off = 0;
    for (int i = 0; i < G; ++i) {
        if (out->radii[i] <= 0) continue;
        int rmin[2], rmax[2];
        get_rect(out->xy + 2 * i, out->radii[i], gx, gy, rmin, rmax);
        for (int y = rmin[1]; y < rmax[1]; ++y)
            for (int x = rmin[0]; x < rmax[0]; ++x) {
                kv[off].key = ((uint64_t)(y * gx + x) << 32) | depth_bits(out->depths[i]);
                kv[off].id = (uint32_t)i;
                ++off;
            }
    }
    qsort(kv, (size_t)n, sizeof(KeyId), cmp_keyid);
    out->keys_sorted = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    out->point_list = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    out->ranges = (uint32_t *)calloc((size_t)gx * gy * 2, sizeof(uint32_t));
    for (int64_t k = 0; k < n; ++k) {
        out->keys_sorted[k] = kv[k].key;
        out->point_list[k] = kv[k].id;
        const uint32_t tile = (uint32_t)(kv[k].key >> 32);
        if (k == 0 || (uint32_t)(kv[k - 1].key >> 32) != tile) out->ranges[2 * tile] = (uint32_t)k;
        if (k == n - 1 || (uint32_t)(kv[k + 1].key >> 32) != tile) out->ranges[2 * tile + 1] = (uint32_t)(k + 1);
    }
    free(kv);

    /* R.6 blend, front to back */
    const real eps = out->margin_eps;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = out->ranges[2 * (ty * gx + tx)], r1 = out->ranges[2 * (ty * gx + tx) + 1];
            for (int py = ty * TILE; py < ty * TILE + TILE && py < H; ++py)
                for (int px = tx * TILE; px < tx * TILE + TILE && px < W; ++px) {
                    real T = R(1.0), acc_c[3] = {0, 0, 0}, acc_d = 0, acc_a = 0, fb = 0;
                    real acc_f[64];
                    for (int c = 0; c < C; ++c) acc_f[c] = 0;
                    uint32_t contributor = 0, last = 0;
                    for (uint32_t k = r0; k < r1; ++k) {
                        ++contributor;
                        const uint32_t g = out->point_list[k];
                        const real dx = out->xy[2 * g] - (real)px, dy = out->xy[2 * g + 1] - (real)py;
                        const real *co = out->conic_opacity + 4 * g;
                        const real power = R(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        const real w_full = fmin(R(0.99), co[3] * EXP(fmin(power, R(0.0)))) * T;
                        if (out->flip_bound && fabs(power) <= eps) {
                            if (w_full > fb) fb = w_full;
                            if (out->marginal) out->marginal[g] = 1;
                        }
                        if (power > R(0.0)) continue;
                        const real alpha = fmin(R(0.99), co[3] * EXP(power));
                        if (out->flip_bound && fabs(alpha - R(1.0) / R(255.0)) <= eps * (R(1.0) / R(255.0))) {
                            if (alpha * T > fb) fb = alpha * T;
                            if (out->marginal) out->marginal[g] = 1;
                        }
                        if (alpha < R(1.0) / R(255.0)) continue;
                        const real test_T = T * (R(1.0) - alpha);
                        if (out->flip_bound && fabs(test_T - R(0.0001)) <= eps * R(0.0001)) {
                            if (T > fb) fb = T;
                            if (out->marginal) out->marginal[g] = 1;
                        }
                        if (test_T < R(0.0001)) break; /* [EXT] done = true */
                        const real w = alpha * T;
                        if (has_color) for (int c = 0; c < 3; ++c) acc_c[c] += out->rgb[3 * g + c] * w;
                        for (int c = 0; c < C; ++c) acc_f[c] += in->features[(size_t)g * C + c] * w;
                        acc_d += out->depths[g] * w;
                        acc_a += w;
                        T = test_T;
                        last = contributor;
                    }
                    const size_t pid = (size_t)py * W + px;
                    out->final_T[pid] = T;
                    out->n_contrib[pid] = last;
                    if (has_color && out->out_color)
                        for (int c = 0; c < 3; ++c) out->out_color[(size_t)c * H * W + pid] = acc_c[c] + T * in->bg[c];
                    if (C && out->out_feature)
                        for (int c = 0; c < C; ++c) out->out_feature[(size_t)c * H * W + pid] = acc_f[c];
                    out->out_alpha[pid] = acc_a;
                    out->out_depth[pid] = acc_d;
                    if (out->flip_bound) out->flip_bound[pid] = fb;
                }
        }
    return 0;
}

void FN(oracle_free)(OracleOut *out) {
    free(out->keys_sorted); free(out->point_list); free(out->ranges);
    out->keys_sorted = NULL; out->point_list = NULL; out->ranges = NULL;
}

/* ---------------- backward ---------------- */
typedef struct {
    const real *dL_dcolor;   /* (3,H,W) or NULL */
    const real *dL_dfeature; /* (C,H,W) or NULL */
    const real *dL_dalpha;   /* (H,W) or NULL   */
    const real *dL_ddepth;   /* (H,W) or NULL   */
    /* outputs, caller allocates, oracle zero-fills */
    real *dL_dmeans3D;       /* (G,3) */
    real *dL_dmeans2D;       /* (G,2) screen-space sink (NDC-scaled, [EXT]) */
    real *dL_dshs;           /* (G,n,3) or NULL */
    real *dL_dcolors;        /* (G,3): grad of the blended colour (after clamp mask) */
    real *dL_dfeatures;      /* (G,C) or NULL */
    real *dL_dopacity;       /* (G) */
    real *dL_dcov3D;         /* (G,6) */
    real *dL_dconic;         /* (G,3) scratch/inspection: conic xx, xy, yy */
    real *dL_ddepths;        /* (G) scratch/inspection */
} OracleGrad;

int FN(oracle_backward)(const OracleIn *in, const OracleOut *out, OracleGrad *gr, int n_threads) {
    const int G = in->G, H = in->H, W = in->W, C = in->features ? in->C : 0;
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const real fx = (real)W / MUL(R(2.0), in->tanfovx), fy = (real)H / MUL(R(2.0), in->tanfovy);
    const int has_color = in->shs || in->colors_precomp;
    const real s = in->scene_scale, s2 = MUL(s, s);
    const int nsh = (in->sh_degree + 1) * (in->sh_degree + 1);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    memset(gr->dL_dmeans3D, 0, sizeof(real) * 3 * G);
    memset(gr->dL_dmeans2D, 0, sizeof(real) * 2 * G);
    memset(gr->dL_dcolors, 0, sizeof(real) * 3 * G);
    if (gr->dL_dshs) memset(gr->dL_dshs, 0, sizeof(real) * 3 * nsh * G);
    if (gr->dL_dfeatures) memset(gr->dL_dfeatures, 0, sizeof(real) * (size_t)C * G);
    memset(gr->dL_dopacity, 0, sizeof(real) * G);
    memset(gr->dL_dcov3D, 0, sizeof(real) * 6 * G);
    memset(gr->dL_dconic, 0, sizeof(real) * 3 * G);
    memset(gr->dL_ddepths, 0, sizeof(real) * G);

    /* R.7 blend backward: per pixel, back to front ([EXT] renderCUDA backward).
       The min(0.99, .) clamp is treated as identity for the gradient, as upstream. */
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int ty = 0; ty < gy; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t r0 = out->ranges[2 * (ty * gx + tx)];
            for (int py = ty * TILE; py < ty * TILE + TILE && py < H; ++py)
                for (int px = tx * TILE; px < tx * TILE + TILE && px < W; ++px) {
                    const size_t pid = (size_t)py * W + px;
                    const real T_final = out->final_T[pid];
                    real T = T_final;
                    const uint32_t last = out->n_contrib[pid];
                    /* channel vector: colour(3) | features(C) | depth | alpha */
                    real g[3 + 64 + 2], accum[3 + 64 + 2], lastc[3 + 64 + 2];
                    const int nch = 3 + C + 2;
                    for (int c = 0; c < 3; ++c) g[c] = (has_color && gr->dL_dcolor) ? gr->dL_dcolor[(size_t)c * H * W + pid] : 0;
                    for (int c = 0; c < C; ++c) g[3 + c] = gr->dL_dfeature ? gr->dL_dfeature[(size_t)c * H * W + pid] : 0;
                    g[3 + C] = gr->dL_ddepth ? gr->dL_ddepth[pid] : 0;
                    g[3 + C + 1] = gr->dL_dalpha ? gr->dL_dalpha[pid] : 0;
                    for (int c = 0; c < nch; ++c) accum[c] = lastc[c] = 0;
                    real last_alpha = 0;
                    real bg_dot = 0;
                    for (int c = 0; c < 3; ++c) bg_dot += in->bg[c] * g[c];
                    for (uint32_t k = last; k-- > 0;) {
                        const uint32_t gi = out->point_list[r0 + k];
                        const real dx = out->xy[2 * gi] - (real)px, dy = out->xy[2 * gi + 1] - (real)py;
                        const real *co = out->conic_opacity + 4 * gi;
                        const real power = R(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (power > R(0.0)) continue;
                        const real Gv = EXP(power);
                        const real alpha = fmin(R(0.99), co[3] * Gv);
                        if (alpha < R(1.0) / R(255.0)) continue;
                        T = T / (R(1.0) - alpha);
                        const real w = alpha * T;
                        real cval[3 + 64 + 2];
                        for (int c = 0; c < 3; ++c) cval[c] = has_color ? out->rgb[3 * gi + c] : 0;
                        for (int c = 0; c < C; ++c) cval[3 + c] = in->features[(size_t)gi * C + c];
                        cval[3 + C] = out->depths[gi];
                        cval[3 + C + 1] = R(1.0);
                        real dL_dalpha = 0;
                        for (int c = 0; c < nch; ++c) {
                            accum[c] = last_alpha * lastc[c] + (R(1.0) - last_alpha) * accum[c];
                            lastc[c] = cval[c];
                            dL_dalpha += (cval[c] - accum[c]) * g[c];
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (R(1.0) - alpha)) * bg_dot;
                        const real dL_dG = co[3] * dL_dalpha;
                        const real gdx = Gv * dx, gdy = Gv * dy;
                        const real dG_ddelx = -gdx * co[0] - gdy * co[1];
                        const real dG_ddely = -gdy * co[2] - gdx * co[1];
                        real add[16 + 64];
                        int na = 0;
                        add[na++] = dL_dG * dG_ddelx * (R(0.5) * W);
                        add[na++] = dL_dG * dG_ddely * (R(0.5) * H);
                        add[na++] = R(-0.5) * gdx * dx * dL_dG;
                        add[na++] = R(-0.5) * gdx * dy * dL_dG;
                        add[na++] = R(-0.5) * gdy * dy * dL_dG;
                        add[na++] = Gv * dL_dalpha;
                        real *dst[6] = {&gr->dL_dmeans2D[2 * gi], &gr->dL_dmeans2D[2 * gi + 1], &gr->dL_dconic[3 * gi],
                                        &gr->dL_dconic[3 * gi + 1], &gr->dL_dconic[3 * gi + 2], &gr->dL_dopacity[gi]};
                        for (int q = 0; q < 6; ++q) {
#pragma omp atomic
                            *dst[q] += add[q];
                        }
                        if (has_color)
                            for (int c = 0; c < 3; ++c) {
#pragma omp atomic
                                gr->dL_dcolors[3 * gi + c] += w * g[c];
                            }
                        if (gr->dL_dfeatures)
                            for (int c = 0; c < C; ++c) {
#pragma omp atomic
                                gr->dL_dfeatures[(size_t)gi * C + c] += w * g[3 + c];
                            }
#pragma omp atomic
                        gr->dL_ddepths[gi] += w * g[3 + C];
                    }
                }
        }

    /* R.8 per-Gaussian backward ([EXT] computeCov2DCUDA + preprocessCUDA backward) */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        if (out->radii[i] <= 0) continue;
        const real p[3] = {MUL(in->means3D[3 * i], s), MUL(in->means3D[3 * i + 1], s), MUL(in->means3D[3 * i + 2], s)};
        real cv[6];
        for (int k = 0; k < 6; ++k) cv[k] = MUL(in->cov3D[6 * i + k], s2);
        Cov2D q;
        cov2d(p, fx, fy, in->tanfovx, in->tanfovy, cv, in->viewmatrix, &q);
        const real a = q.a, b = q.b, c = q.c;
        const real denom = a * c - b * b;
        const real d2inv = R(1.0) / (denom * denom + R(0.0000001));
        const real gcx = gr->dL_dconic[3 * i], gcy = gr->dL_dconic[3 * i + 1], gcz = gr->dL_dconic[3 * i + 2];
        real dL_da = 0, dL_db = 0, dL_dc = 0;
        if (d2inv != 0) {
            dL_da = d2inv * (-c * c * gcx + R(2.0) * b * c * gcy + (denom - a * c) * gcz);
            dL_dc = d2inv * (-a * a * gcz + R(2.0) * a * b * gcy + (denom - a * c) * gcx);
            dL_db = d2inv * R(2.0) * (b * c * gcx - (denom + R(2.0) * b * b) * gcy + a * b * gcz);
            const real(*M)[3] = q.M;
            real dcov[6];
            dcov[0] = M[0][0] * M[0][0] * dL_da + M[0][0] * M[1][0] * dL_db + M[1][0] * M[1][0] * dL_dc;
            dcov[3] = M[0][1] * M[0][1] * dL_da + M[0][1] * M[1][1] * dL_db + M[1][1] * M[1][1] * dL_dc;
            dcov[5] = M[0][2] * M[0][2] * dL_da + M[0][2] * M[1][2] * dL_db + M[1][2] * M[1][2] * dL_dc;
            dcov[1] = R(2.0) * M[0][0] * M[0][1] * dL_da + (M[0][0] * M[1][1] + M[0][1] * M[1][0]) * dL_db +
                      R(2.0) * M[1][0] * M[1][1] * dL_dc;
            dcov[2] = R(2.0) * M[0][0] * M[0][2] * dL_da + (M[0][0] * M[1][2] + M[0][2] * M[1][0]) * dL_db +
                      R(2.0) * M[1][0] * M[1][2] * dL_dc;
            dcov[4] = R(2.0) * M[0][2] * M[0][1] * dL_da + (M[0][1] * M[1][2] + M[0][2] * M[1][1]) * dL_db +
                      R(2.0) * M[1][1] * M[1][2] * dL_dc;
            /* chain through cov *= s^2 */
            for (int k = 0; k < 6; ++k) gr->dL_dcov3D[6 * i + k] = dcov[k] * s2;
        }
        /* dL/dM rows */
        real dM0[3], dM1[3];
        for (int k = 0; k < 3; ++k) {
            dM0[k] = R(2.0) * q.v0[k] * dL_da + q.v1[k] * dL_db;
            dM1[k] = R(2.0) * q.v1[k] * dL_dc + q.v0[k] * dL_db;
        }
        const real *vm = in->viewmatrix; /* R[r][c] = vm[4c + r] */
        real dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int j = 0; j < 3; ++j) {
            dJ00 += dM0[j] * vm[4 * j + 0];
            dJ02 += dM0[j] * vm[4 * j + 2];
            dJ11 += dM1[j] * vm[4 * j + 1];
            dJ12 += dM1[j] * vm[4 * j + 2];
        }
        const real tz = R(1.0) / q.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        real dt[3];
        dt[0] = q.xmul * -fx * tz2 * dJ02;
        dt[1] = q.ymul * -fy * tz2 * dJ12;
        dt[2] = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (R(2.0) * fx * q.t[0]) * tz3 * dJ02 + (R(2.0) * fy * q.t[1]) * tz3 * dJ12;
        /* depth channel: depth = row 2 of the view transform */
        dt[2] += gr->dL_ddepths[i];
        real dmean[3];
        for (int j = 0; j < 3; ++j) dmean[j] = vm[4 * j + 0] * dt[0] + vm[4 * j + 1] * dt[1] + vm[4 * j + 2] * dt[2];
        /* screen-space mean */
        const real *pm = in->projmatrix;
        real ph[4];
        xform4x4(pm, p, ph);
        const real mw = R(1.0) / (ph[3] + R(0.0000001));
        const real mul1 = ph[0] * mw * mw, mul2 = ph[1] * mw * mw;
        const real g2x = gr->dL_dmeans2D[2 * i], g2y = gr->dL_dmeans2D[2 * i + 1];
        dmean[0] += (pm[0] * mw - pm[3] * mul1) * g2x + (pm[1] * mw - pm[3] * mul2) * g2y;
        dmean[1] += (pm[4] * mw - pm[7] * mul1) * g2x + (pm[5] * mw - pm[7] * mul2) * g2y;
        dmean[2] += (pm[8] * mw - pm[11] * mul1) * g2x + (pm[9] * mw - pm[11] * mul2) * g2y;
        /* colour SH */
        if (in->shs) {
            real d[3] = {p[0] - in->campos[0], p[1] - in->campos[1], p[2] - in->campos[2]};
            const real len2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            const real inv = R(1.0) / SQRT(len2);
            const real u[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
            real bas[25], dbas[25][3];
            if (g_color_sh_3dgs) {
                sh_basis(in->sh_degree, u[1], u[2], u[0], bas, dbas);
                sh_patch_3dgs(in->sh_degree, u[1], u[2], u[0], bas, dbas);
            } else {
                sh_basis(in->sh_degree, u[0], u[1], u[2], bas, dbas);
            }
            real gc[3];
            for (int k = 0; k < 3; ++k) gc[k] = out->clamped[3 * i + k] ? 0 : gr->dL_dcolors[3 * i + k];
            real ddir[3] = {0, 0, 0};
            const real *sh = in->shs + (size_t)i * nsh * 3;
            for (int k = 0; k < nsh; ++k) {
                real sg = 0;
                for (int ch = 0; ch < 3; ++ch) {
                    if (gr->dL_dshs) gr->dL_dshs[((size_t)i * nsh + k) * 3 + ch] = bas[k] * gc[ch];
                    sg += sh[k * 3 + ch] * gc[ch];
                }
                for (int a3 = 0; a3 < 3; ++a3) ddir[a3] += dbas[k][a3] * sg;
            }
            if (g_color_sh_3dgs) {   /* gradients w.r.t. (y, z, x) back to (x, y, z) */
                const real gx_ = ddir[2], gy_ = ddir[0], gz_ = ddir[1];
                ddir[0] = gx_; ddir[1] = gy_; ddir[2] = gz_;
            }
            /* through the normalisation u = d/|d| */
            const real dot = u[0] * ddir[0] + u[1] * ddir[1] + u[2] * ddir[2];
            for (int a3 = 0; a3 < 3; ++a3) dmean[a3] += (ddir[a3] - u[a3] * dot) * inv;
        }
        /* chain through means *= s */
        for (int j = 0; j < 3; ++j) gr->dL_dmeans3D[3 * i + j] = dmean[j] * s;
    }
    return 0;
}
