/*
 * surfel_oracle.c -- CPU restatement of the reference 2D-Gaussian (surfel) rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (g4splat_b200/) may import, link
 * or call this file.  It is used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py as the checker / reported CPU baseline.
 *
 * Parity pin: the reference ships no golden vectors or tests for this path (SURVEY.md 4,
 * 8c).  The pin is the reference implementation itself (oracle/_ref, built from the sources
 * under /root/reference by oracle/build_ref.sh) run on a B200; its outputs on seeded scenes
 * are committed under tests/golden/ and this file is checked against them.
 *
 * Every function cites the reference file:line it restates.  Paths:
 *   CR/  = 2d-gaussian-splatting/submodules/diff-surfel-rasterization/cuda_rasterizer/
 *
 * Build: -DORC_REAL=float (default, fp32 like the reference) or -DORC_REAL=double
 * (ground truth used to judge which fp32 implementation is closer).  -ffp-contract=off so
 * the fp32 build is plain IEEE without FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORC_REAL
#define ORC_REAL float
#endif
typedef ORC_REAL real;

#define BLOCK_X 16 /* CR/config.h:15 */
#define BLOCK_Y 16 /* CR/config.h:16 */

/* CR/auxiliary.h:37-39 */
static const real near_n = (real)0.2f;
static const real far_n = (real)100.0f;
static const real FilterInvSquare = (real)2.0f;

/* CR/auxiliary.h:42-59 */
static const real SH_C0 = (real)0.28209479177387814f;
static const real SH_C1 = (real)0.4886025119029199f;
static const real SH_C2[5] = {(real)1.0925484305920792f, (real)-1.0925484305920792f,
                              (real)0.31539156525252005f, (real)-1.0925484305920792f,
                              (real)0.5462742152960396f};
static const real SH_C3[7] = {(real)-0.5900435899266435f, (real)2.890611442640554f,
                              (real)-0.4570457994644658f, (real)0.3731763325901154f,
                              (real)-0.4570457994644658f, (real)1.445305721320277f,
                              (real)-0.5900435899266435f};

static inline real r_sqrt(real x) { return (real)sqrt((double)x) ; }
static inline real r_exp(real x) {
    if (sizeof(real) == 4) return (real)expf((float)x);
    return (real)exp((double)x);
}
static inline real r_max(real a, real b) { return a > b ? a : b; } /* fmaxf-like for non-NaN */
static inline real r_min(real a, real b) { return a < b ? a : b; }

/* CUDA float->int conversion (cvt.rzi.s32.f32): truncates, saturates, NaN -> 0. */
static inline int f2i(real v) {
    if (v != v) return 0;
    if (v >= (real)2147483647.0) return 2147483647;
    if (v <= (real)-2147483648.0) return (int)(-2147483647 - 1);
    return (int)v;
}
/* CUDA float->uint conversion (cvt.rzi.u32.f32): negative and NaN -> 0. */
static inline uint32_t f2u(real v) {
    if (!(v > (real)0)) return 0u;
    if (v >= (real)4294967295.0) return 4294967295u;
    return (uint32_t)v;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------ */
/* CR/auxiliary.h:66-76  getRect                                                         */
static void get_rect(real px, real py, int max_radius, int gx, int gy, uint32_t* rmin, uint32_t* rmax) {
    rmin[0] = (uint32_t)imin(gx, imax(0, f2i((px - (real)max_radius) / (real)BLOCK_X)));
    rmin[1] = (uint32_t)imin(gy, imax(0, f2i((py - (real)max_radius) / (real)BLOCK_Y)));
    rmax[0] = (uint32_t)imin(gx, imax(0, f2i((px + (real)max_radius + (real)(BLOCK_X - 1)) / (real)BLOCK_X)));
    rmax[1] = (uint32_t)imin(gy, imax(0, f2i((py + (real)max_radius + (real)(BLOCK_Y - 1)) / (real)BLOCK_Y)));
}

/* CR/auxiliary.h:212-234  quat_to_rotmat; R[c][r] column-major like glm.  quat = (w,x,y,z)
 * stored in glm::vec4 fields (x,y,z,w). */
static void quat_to_rotmat(const real* q, real R[3][3]) {
    /* reference sums quat.w^2 + quat.x^2 + quat.y^2 + quat.z^2 with glm fields, i.e.
     * q[3]^2 + q[0]^2 + q[1]^2 + q[2]^2 (CR/auxiliary.h:214-216); rsqrtf there. */
    real s = (real)1 / r_sqrt(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    real w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    R[0][0] = (real)1 - (real)2 * (y * y + z * z);
    R[0][1] = (real)2 * (x * y + w * z);
    R[0][2] = (real)2 * (x * z - w * y);
    R[1][0] = (real)2 * (x * y - w * z);
    R[1][1] = (real)1 - (real)2 * (x * x + z * z);
    R[1][2] = (real)2 * (y * z + w * x);
    R[2][0] = (real)2 * (x * z + w * y);
    R[2][1] = (real)2 * (y * z - w * x);
    R[2][2] = (real)1 - (real)2 * (x * x + y * y);
}

/* CR/auxiliary.h:237-281  quat_to_rotmat_vjp; vR[c][r] column-major. */
static void quat_to_rotmat_vjp(const real* q, real vR[3][3], real* vq) {
    real s = (real)1 / r_sqrt(q[3] * q[3] + q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    real w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    vq[0] = (real)2 * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
    vq[1] = (real)2 * ((real)-2 * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) +
                       z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
    vq[2] = (real)2 * (x * (vR[0][1] + vR[1][0]) - (real)2 * y * (vR[0][0] + vR[2][2]) +
                       z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
    vq[3] = (real)2 * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) -
                       (real)2 * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
}

/* Shared by forward (CR/forward.cu:75-115) and backward (CR/backward.cu:476-506).
 * Builds X = M^T * world2ndc (clip coords of the two scaled tangent axes and the centre) and
 * T = X * ndc2pix with glm's left-to-right sums.  T[c][r]: T[0]=Tu, T[1]=Tv, T[2]=Tw.
 * Pm[c][k] = (world2ndc * ndc2pix)[c][k] is returned for the backward (mat3x4 P). */
static void build_T(const real* p, real sx, real sy, real R[3][3], const real* pm, int W, int H,
                    real T[3][3], real Pm[3][4]) {
    /* L = R * S, S = diag(sx, sy, 1): L[c] = R[c] * s_c (CR/forward.cu:88-90, glm mat3*mat3
     * with zero off-diagonals adds exact zeros). */
    real S[3][4]; /* splat2world columns as rows: S[r][j] */
    for (int j = 0; j < 3; j++) { S[0][j] = R[0][j] * sx; S[1][j] = R[1][j] * sy; S[2][j] = p[j]; }
    S[0][3] = 0; S[1][3] = 0; S[2][3] = 1;
    real nd[3][4] = {{(real)W / (real)2, 0, 0, (real)(W - 1) / (real)2},
                     {0, (real)H / (real)2, 0, (real)(H - 1) / (real)2},
                     {0, 0, 0, 1}};
    real X[4][3]; /* X[c][r] */
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++)
            X[c][r] = S[r][0] * pm[c] + S[r][1] * pm[c + 4] + S[r][2] * pm[c + 8] + S[r][3] * pm[c + 12];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            T[c][r] = X[0][r] * nd[c][0] + X[1][r] * nd[c][1] + X[2][r] * nd[c][2] + X[3][r] * nd[c][3];
    if (Pm) {
        /* Backward association (CR/backward.cu:490-504): P = world2ndc(mat4) * ndc2pix(mat3x4),
         * P[c][k] = sum_j w2n[j][k] * nd[c][j] with w2n[j][k] = pm[j + 4k]; then
         * T = transpose(M) * P, T[c][r] = sum_j S[r][j] * P[c][j]. */
        for (int c = 0; c < 3; c++)
            for (int k = 0; k < 4; k++)
                Pm[c][k] = pm[0 + 4 * k] * nd[c][0] + pm[1 + 4 * k] * nd[c][1] + pm[2 + 4 * k] * nd[c][2] +
                           pm[3 + 4 * k] * nd[c][3];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++)
                T[c][r] = S[r][0] * Pm[c][0] + S[r][1] * Pm[c][1] + S[r][2] * Pm[c][2] + S[r][3] * Pm[c][3];
    }
}

/* CR/forward.cu:20-71  computeColorFromSH (forward) */
static void sh_to_rgb(int idx, int deg, int M, const real* means, const real* campos, const real* shs,
                      uint8_t* clamped, real* out) {
    const real* pos = means + 3 * idx;
    real d[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    real len = r_sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    real x = d[0] / len, y = d[1] / len, z = d[2] / len;
    const real* sh = shs + (size_t)idx * M * 3;
    for (int c = 0; c < 3; c++) {
#define SH(k) sh[3 * (k) + c]
        real r = SH_C0 * SH(0);
        if (deg > 0) {
            r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) +
                    SH_C2[2] * ((real)2 * zz - xx - yy) * SH(6) + SH_C2[3] * xz * SH(7) +
                    SH_C2[4] * (xx - yy) * SH(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * ((real)3 * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
                        SH_C3[2] * y * ((real)4 * zz - xx - yy) * SH(11) +
                        SH_C3[3] * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy) * SH(12) +
                        SH_C3[4] * x * ((real)4 * zz - xx - yy) * SH(13) + SH_C3[5] * z * (xx - yy) * SH(14) +
                        SH_C3[6] * x * (xx - (real)3 * yy) * SH(15);
                }
            }
        }
#undef SH
        r += (real)0.5;
        clamped[3 * idx + c] = (r < 0);
        out[c] = r_max(r, 0);
    }
}

/* ------------------------------------------------------------------------------------ */
/* CR/forward.cu:150-253  preprocessCUDA (forward), incl. in_frustum (CR/auxiliary.h:184-209),
 * compute_transmat (CR/forward.cu:75-115), compute_aabb (CR/forward.cu:119-147).
 * Outputs are only written for Gaussians that survive every cull, exactly like the
 * reference; callers pre-fill them (the reference leaves them uninitialised).
 * Returns the number of prefiltered-contract violations (reference: printf + __trap). */
int orc_project(int P, int D, int M, const real* means3D, const real* scales, real scale_modifier,
                const real* rotations, const real* opacities, const real* shs, const real* transMat_precomp,
                const real* colors_precomp, const real* view, const real* proj, const real* campos, int W,
                int H, int prefiltered, int* radii, real* means2D, real* depths, real* transMats, real* rgb,
                real* normal_opacity, uint8_t* clamped, uint32_t* tiles_touched) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    int violations = 0;
#pragma omp parallel for schedule(static) reduction(+ : violations)
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        const real* p = means3D + 3 * idx;
        /* in_frustum: p_view = transformPoint4x3 */
        real pv[3];
        for (int i = 0; i < 3; i++) pv[i] = view[i] * p[0] + view[4 + i] * p[1] + view[8 + i] * p[2] + view[12 + i];
        if (pv[2] <= (real)0.2f) {
            if (prefiltered) violations++;
            continue;
        }
        real T[3][3];
        real normal[3];
        if (transMat_precomp == NULL) {
            real R[3][3];
            quat_to_rotmat(rotations + 4 * idx, R);
            build_T(p, scale_modifier * scales[2 * idx], scale_modifier * scales[2 * idx + 1], R, proj, W, H, T,
                    NULL);
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++) transMats[9 * idx + 3 * c + r] = T[c][r];
            /* normal = transformVec4x3(L[2] = R[2], view) */
            for (int i = 0; i < 3; i++) normal[i] = view[i] * R[2][0] + view[4 + i] * R[2][1] + view[8 + i] * R[2][2];
        } else {
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++) T[c][r] = transMat_precomp[9 * idx + 3 * c + r];
            normal[0] = 0; normal[1] = 0; normal[2] = 1;
        }
        /* DUAL_VISIABLE, CR/forward.cu:211-216 */
        real cosv = -((pv[0] * normal[0] + pv[1] * normal[1]) + pv[2] * normal[2]);
        if (cosv == 0) continue;
        real mult = cosv > 0 ? (real)1 : (real)-1;
        for (int i = 0; i < 3; i++) normal[i] = mult * normal[i];

        /* compute_aabb, cutoff = 3 */
        const real cutoff = (real)3;
        real t[3] = {cutoff * cutoff, cutoff * cutoff, (real)-1};
        real dist = ((T[2][0] * T[2][0]) * t[0] + (T[2][1] * T[2][1]) * t[1]) + (T[2][2] * T[2][2]) * t[2];
        real f[3] = {((real)1 / dist) * t[0], ((real)1 / dist) * t[1], ((real)1 / dist) * t[2]};
        if (dist == 0) continue;
        real pix[2], tmp[2];
        for (int a = 0; a < 2; a++) {
            pix[a] = ((f[0] * T[a][0]) * T[2][0] + (f[1] * T[a][1]) * T[2][1]) + (f[2] * T[a][2]) * T[2][2];
            tmp[a] = ((f[0] * T[a][0]) * T[a][0] + (f[1] * T[a][1]) * T[a][1]) + (f[2] * T[a][2]) * T[a][2];
        }
        real ex = r_sqrt(r_max((real)1e-4f, pix[0] * pix[0] - tmp[0]));
        real ey = r_sqrt(r_max((real)1e-4f, pix[1] * pix[1] - tmp[1]));
        real radius = (real)ceil((double)r_max(ex, ey));

        uint32_t rmin[2], rmax[2];
        get_rect(pix[0], pix[1], f2i(radius), gx, gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;

        if (colors_precomp == NULL) sh_to_rgb(idx, D, M, means3D, campos, shs, clamped, rgb + 3 * idx);

        depths[idx] = pv[2];
        radii[idx] = f2i(radius);
        means2D[2 * idx] = pix[0];
        means2D[2 * idx + 1] = pix[1];
        normal_opacity[4 * idx + 0] = normal[0];
        normal_opacity[4 * idx + 1] = normal[1];
        normal_opacity[4 * idx + 2] = normal[2];
        normal_opacity[4 * idx + 3] = opacities[idx];
        tiles_touched[idx] = (rmax[1] - rmin[1]) * (rmax[0] - rmin[0]);
    }
    return violations;
}

/* CR/rasterizer_impl.cu:54-66,141-153  markVisible / checkFrustum */
void orc_mark_visible(int P, const real* means3D, const real* view, const real* proj, uint8_t* present) {
    (void)proj;
    for (int idx = 0; idx < P; idx++) {
        const real* p = means3D + 3 * idx;
        real z = view[2] * p[0] + view[6] * p[1] + view[10] * p[2] + view[14];
        present[idx] = !(z <= (real)0.2f);
    }
}

/* ------------------------------------------------------------------------------------ */
/* Binning: CR/rasterizer_impl.cu:276-319 (InclusiveSum, duplicateWithKeys :70-111, SortPairs,
 * identifyTileRanges :116-138).  The reference sorts 64-bit keys tile<<32 | float_bits(depth)
 * with a stable LSD radix sort after emitting in Gaussian-index order, so the order is
 * (tile, depth bits as uint32, Gaussian index).  The fp64 build orders by the double depth. */
typedef struct { uint32_t tile; uint32_t idx; uint64_t dkey; } orc_inst_t;

static int inst_cmp(const void* a, const void* b) {
    const orc_inst_t* x = (const orc_inst_t*)a;
    const orc_inst_t* y = (const orc_inst_t*)b;
    if (x->tile != y->tile) return x->tile < y->tile ? -1 : 1;
    if (x->dkey != y->dkey) return x->dkey < y->dkey ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

long long orc_count_instances(int P, const uint32_t* tiles_touched) {
    long long R = 0;
    for (int i = 0; i < P; i++) R += tiles_touched[i];
    return R;
}

/* point_list[R], ranges[2*T] (zero for untouched tiles, like the cudaMemset at :311). */
int orc_bin(int P, int W, int H, const real* means2D, const real* depths, const int* radii, long long R,
            uint32_t* point_list, uint32_t* ranges) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R == 0) return 0;
    orc_inst_t* inst = (orc_inst_t*)malloc(sizeof(orc_inst_t) * (size_t)R);
    if (!inst) return -1;
    long long off = 0;
    for (int idx = 0; idx < P; idx++) {
        if (radii[idx] > 0) {
            uint32_t rmin[2], rmax[2];
            get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, rmin, rmax);
            uint64_t dkey;
            if (sizeof(real) == 4) {
                float d = (float)depths[idx];
                uint32_t b;
                memcpy(&b, &d, 4);
                dkey = b;
            } else {
                double d = (double)depths[idx];
                memcpy(&dkey, &d, 8); /* positive doubles order like their bit patterns */
            }
            for (uint32_t y = rmin[1]; y < rmax[1]; y++)
                for (uint32_t x = rmin[0]; x < rmax[0]; x++) {
                    if (off >= R) { free(inst); return -2; }
                    inst[off].tile = y * (uint32_t)gx + x;
                    inst[off].idx = (uint32_t)idx;
                    inst[off].dkey = dkey;
                    off++;
                }
        }
    }
    if (off != R) { free(inst); return -3; }
    qsort(inst, (size_t)R, sizeof(orc_inst_t), inst_cmp);
    for (long long i = 0; i < R; i++) {
        point_list[i] = inst[i].idx;
        uint32_t cur = inst[i].tile;
        if (i == 0) ranges[2 * cur] = 0;
        else if (inst[i - 1].tile != cur) {
            ranges[2 * inst[i - 1].tile + 1] = (uint32_t)i;
            ranges[2 * cur] = (uint32_t)i;
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    free(inst);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* CR/forward.cu:258-443  renderCUDA (forward).  One pixel at a time; the block-level
 * "all 256 threads done" early exit (:327-329) only skips work whose results are discarded.
 * features = colors_precomp or the rgb from orc_project ([P,3]).                          */
/* test instrumentation (no reference counterpart): when set, pixel p's number of blended
 * (pixel, surfel) pairs is stored at pairs_per_pixel[p] by the next orc_blend_forward calls. */
static uint32_t* g_pairs_per_pixel = 0;
void orc_set_pair_counter(uint32_t* pairs_per_pixel) { g_pairs_per_pixel = pairs_per_pixel; }

void orc_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* means2D,
                       const real* features, const real* transMats, const real* normal_opacity, const real* bg,
                       real* out_color, real* out_others, real* final_T, uint32_t* n_contrib) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const size_t pix_id = (size_t)W * pyi + pxi;
                const real pxf = (real)pxi, pyf = (real)pyi;
                real T = 1;
                uint32_t contributor = 0, last_contributor = 0, blended = 0;
                real C[3] = {0, 0, 0}, N[3] = {0, 0, 0};
                real Dd = 0, M1 = 0, M2 = 0, distortion = 0, median_depth = 0;
                real median_contributor = -1;
                for (uint32_t e = r0; e < r1; e++) {
                    contributor++;
                    const uint32_t g = point_list[e];
                    const real* Tu = transMats + 9 * (size_t)g;
                    const real* Tv = Tu + 3;
                    const real* Tw = Tu + 6;
                    real k[3], l[3];
                    for (int i = 0; i < 3; i++) { k[i] = pxf * Tw[i] - Tu[i]; l[i] = pyf * Tw[i] - Tv[i]; }
                    real p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0) continue;
                    real sx = p[0] / p[2], sy = p[1] / p[2];
                    real rho3d = sx * sx + sy * sy;
                    real dx = means2D[2 * (size_t)g] - pxf, dy = means2D[2 * (size_t)g + 1] - pyf;
                    real rho2d = FilterInvSquare * (dx * dx + dy * dy);
                    real rho = r_min(rho3d, rho2d);
                    real depth = (rho3d <= rho2d) ? (sx * Tw[0] + sy * Tw[1]) + Tw[2] : Tw[2];
                    if (depth < near_n) continue;
                    const real* no = normal_opacity + 4 * (size_t)g;
                    real opa = no[3];
                    real power = (real)-0.5 * rho;
                    if (power > 0) continue;
                    real alpha = r_min((real)0.99f, opa * r_exp(power));
                    if (alpha < (real)(1.0f / 255.0f)) continue;
                    real test_T = T * ((real)1 - alpha);
                    if (test_T < (real)0.0001f) break; /* done = true (:385-389) */
                    real w = alpha * T;
                    real A = (real)1 - T;
                    real m = far_n / (far_n - near_n) * ((real)1 - near_n / depth);
                    distortion += (m * m * A + M2 - (real)2 * m * M1) * w;
                    Dd += depth * w;
                    M1 += m * w;
                    M2 += m * m * w;
                    if (T > (real)0.5) { median_depth = depth; median_contributor = (real)contributor; }
                    for (int ch = 0; ch < 3; ch++) N[ch] += no[ch] * w;
                    for (int ch = 0; ch < 3; ch++) C[ch] += features[3 * (size_t)g + ch] * w;
                    T = test_T;
                    last_contributor = contributor;
                    blended++;
                }
                if (g_pairs_per_pixel) g_pairs_per_pixel[pix_id] = blended;
                final_T[pix_id] = T;
                n_contrib[pix_id] = last_contributor;
                for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg[ch];
                n_contrib[pix_id + HW] = f2u(median_contributor); /* float -1 saturates to 0 (:432) */
                final_T[pix_id + HW] = M1;
                final_T[pix_id + 2 * HW] = M2;
                out_others[pix_id + 0 * HW] = Dd;
                out_others[pix_id + 1 * HW] = (real)1 - T;
                for (int ch = 0; ch < 3; ch++) out_others[pix_id + (2 + ch) * HW] = N[ch];
                out_others[pix_id + 5 * HW] = median_depth;
                out_others[pix_id + 6 * HW] = distortion;
            }
    }
}

/* ------------------------------------------------------------------------------------ */
/* CR/backward.cu:143-440  renderCUDA (backward).  Per pixel, back-to-front.  The per-Gaussian
 * accumulators replace the reference's global atomicAdd's; with one thread the order is
 * deterministic (tile-major, pixel row-major), with OpenMP the adds are atomic.             */
static inline void acc_add(real* dst, real v) {
#pragma omp atomic
    *dst += v;
}

void orc_blend_backward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const real* bg,
                        const real* means2D, const real* normal_opacity, const real* transMats,
                        const real* colors, const real* final_Ts, const uint32_t* n_contrib,
                        const real* dL_dpixels, const real* dL_depths, real* dL_dtransMat, real* dL_dmean2D,
                        real* dL_dnormal3D, real* dL_dopacity, real* dL_dcolors) {
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int ty = tile / gx, tx = tile % gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        const int toDo = (int)(r1 - r0);
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const size_t pix_id = (size_t)W * pyi + pxi;
                const real pxf = (real)pxi, pyf = (real)pyi;
                const real T_final = final_Ts[pix_id];
                real T = T_final;
                uint32_t contributor = (uint32_t)toDo;
                const int last_contributor = (int)n_contrib[pix_id];
                real accum_rec[3] = {0, 0, 0}, dL_dpixel[3];
                const int median_contributor = (int)n_contrib[pix_id + HW];
                const real dL_ddepth = dL_depths[0 * HW + pix_id];
                const real dL_daccum = dL_depths[1 * HW + pix_id];
                const real dL_dreg = dL_depths[6 * HW + pix_id];
                real dL_dnormal2D[3];
                for (int i = 0; i < 3; i++) dL_dnormal2D[i] = dL_depths[(2 + i) * HW + pix_id];
                const real dL_dmedian_depth = dL_depths[5 * HW + pix_id];
                real last_depth = 0, last_normal[3] = {0, 0, 0};
                real accum_depth_rec = 0, accum_alpha_rec = 0, accum_normal_rec[3] = {0, 0, 0};
                const real final_D = final_Ts[pix_id + HW];
                const real final_D2 = final_Ts[pix_id + 2 * HW];
                const real final_A = (real)1 - T_final;
                real last_dL_dT = 0;
                for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * HW + pix_id];
                real last_alpha = 0, last_color[3] = {0, 0, 0};

                for (int j = 0; j < toDo; j++) {
                    contributor--;
                    if (contributor >= (uint32_t)last_contributor) continue;
                    const uint32_t g = point_list[r1 - 1 - (uint32_t)j];
                    const real* Tu = transMats + 9 * (size_t)g;
                    const real* Tv = Tu + 3;
                    const real* Tw = Tu + 6;
                    real k[3], l[3];
                    for (int i = 0; i < 3; i++) { k[i] = pxf * Tw[i] - Tu[i]; l[i] = pyf * Tw[i] - Tv[i]; }
                    real p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0) continue;
                    real sx = p[0] / p[2], sy = p[1] / p[2];
                    real rho3d = sx * sx + sy * sy;
                    real dx = means2D[2 * (size_t)g] - pxf, dy = means2D[2 * (size_t)g + 1] - pyf;
                    real rho2d = FilterInvSquare * (dx * dx + dy * dy);
                    real rho = r_min(rho3d, rho2d);
                    real c_d = (rho3d <= rho2d) ? (sx * Tw[0] + sy * Tw[1]) + Tw[2] : Tw[2];
                    if (c_d < near_n) continue;
                    const real* no = normal_opacity + 4 * (size_t)g;
                    real opa = no[3];
                    real power = (real)-0.5 * rho;
                    if (power > 0) continue;
                    const real G = r_exp(power);
                    const real alpha = r_min((real)0.99f, opa * G);
                    if (alpha < (real)(1.0f / 255.0f)) continue;

                    T = T / ((real)1 - alpha);
                    const real dchannel_dcolor = alpha * T;
                    real dL_dalpha = 0;
                    for (int ch = 0; ch < 3; ch++) {
                        const real c = colors[3 * (size_t)g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + ((real)1 - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                        acc_add(&dL_dcolors[3 * (size_t)g + ch], dchannel_dcolor * dL_dpixel[ch]);
                    }
                    real dL_dz = 0, dL_dweight = 0;
                    const real m_d = far_n / (far_n - near_n) * ((real)1 - near_n / c_d);
                    const real dmd_dd = (far_n * near_n) / ((far_n - near_n) * c_d * c_d);
                    if (contributor == (uint32_t)(median_contributor - 1)) dL_dz += dL_dmedian_depth;
                    dL_dweight += (final_D2 + m_d * m_d * final_A - (real)2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + ((real)1 - alpha) * last_dL_dT;
                    const real dL_dmd = (real)2 * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;

                    accum_depth_rec = last_alpha * last_depth + ((real)1 - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                    accum_alpha_rec = last_alpha * (real)1 + ((real)1 - last_alpha) * accum_alpha_rec;
                    dL_dalpha += ((real)1 - accum_alpha_rec) * dL_daccum;
                    for (int ch = 0; ch < 3; ch++) {
                        accum_normal_rec[ch] = last_alpha * last_normal[ch] + ((real)1 - last_alpha) * accum_normal_rec[ch];
                        last_normal[ch] = no[ch];
                        dL_dalpha += (no[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
                        acc_add(&dL_dnormal3D[3 * (size_t)g + ch], alpha * T * dL_dnormal2D[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    real bg_dot_dpixel = 0;
                    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg[i] * dL_dpixel[i];
                    dL_dalpha += (-T_final / ((real)1 - alpha)) * bg_dot_dpixel;

                    const real dL_dG = opa * dL_dalpha;
                    dL_dz += alpha * T * dL_ddepth;

                    if (rho3d <= rho2d) {
                        const real dL_dsx = dL_dG * -G * sx + dL_dz * Tw[0];
                        const real dL_dsy = dL_dG * -G * sy + dL_dz * Tw[1];
                        const real dz_dTw[3] = {sx, sy, (real)1};
                        const real dsx_pz = dL_dsx / p[2];
                        const real dsy_pz = dL_dsy / p[2];
                        const real dL_dp[3] = {dsx_pz, dsy_pz, -(dsx_pz * sx + dsy_pz * sy)};
                        /* dL_dk = cross(l, dL_dp); dL_dl = cross(dL_dp, k) */
                        const real dL_dk[3] = {l[1] * dL_dp[2] - l[2] * dL_dp[1], l[2] * dL_dp[0] - l[0] * dL_dp[2],
                                               l[0] * dL_dp[1] - l[1] * dL_dp[0]};
                        const real dL_dl[3] = {dL_dp[1] * k[2] - dL_dp[2] * k[1], dL_dp[2] * k[0] - dL_dp[0] * k[2],
                                               dL_dp[0] * k[1] - dL_dp[1] * k[0]};
                        for (int i = 0; i < 3; i++) {
                            acc_add(&dL_dtransMat[9 * (size_t)g + i], -dL_dk[i]);
                            acc_add(&dL_dtransMat[9 * (size_t)g + 3 + i], -dL_dl[i]);
                            acc_add(&dL_dtransMat[9 * (size_t)g + 6 + i],
                                    pxf * dL_dk[i] + pyf * dL_dl[i] + dL_dz * dz_dTw[i]);
                        }
                    } else {
                        const real dG_ddelx = -G * FilterInvSquare * dx;
                        const real dG_ddely = -G * FilterInvSquare * dy;
                        acc_add(&dL_dmean2D[3 * (size_t)g + 0], dL_dG * dG_ddelx);
                        acc_add(&dL_dmean2D[3 * (size_t)g + 1], dL_dG * dG_ddely);
                        acc_add(&dL_dtransMat[9 * (size_t)g + 8], dL_dz);
                    }
                    acc_add(&dL_dopacity[g], G * dL_dalpha);
                }
            }
    }
}

/* ------------------------------------------------------------------------------------ */
/* CR/backward.cu:20-139  computeColorFromSH (backward) */
static void sh_backward(int idx, int deg, int M, const real* means, const real* campos, const real* shs,
                        const uint8_t* clamped, const real* dL_dcolor, real* dL_dmeans, real* dL_dshs) {
    const real* pos = means + 3 * idx;
    real d0[3] = {pos[0] - campos[0], pos[1] - campos[1], pos[2] - campos[2]};
    real len = r_sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
    real x = d0[0] / len, y = d0[1] / len, z = d0[2] / len;
    const real* sh = shs + (size_t)idx * M * 3;
    real* dsh = dL_dshs + (size_t)idx * M * 3;
    real dRGB[3];
    for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[3 * idx + c] * (clamped[3 * idx + c] ? (real)0 : (real)1);
    real dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
#define SH(k) sh[3 * (k) + c]
#define DSH(k, v) for (int c = 0; c < 3; c++) dsh[3 * (k) + c] = (v) * dRGB[c]
    DSH(0, SH_C0);
    if (deg > 0) {
        DSH(1, -SH_C1 * y);
        DSH(2, SH_C1 * z);
        DSH(3, -SH_C1 * x);
        for (int c = 0; c < 3; c++) {
            dRGBdx[c] = -SH_C1 * SH(3);
            dRGBdy[c] = -SH_C1 * SH(1);
            dRGBdz[c] = SH_C1 * SH(2);
        }
        if (deg > 1) {
            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            DSH(4, SH_C2[0] * xy);
            DSH(5, SH_C2[1] * yz);
            DSH(6, SH_C2[2] * ((real)2 * zz - xx - yy));
            DSH(7, SH_C2[3] * xz);
            DSH(8, SH_C2[4] * (xx - yy));
            for (int c = 0; c < 3; c++) {
                dRGBdx[c] += SH_C2[0] * y * SH(4) + SH_C2[2] * (real)2 * -x * SH(6) + SH_C2[3] * z * SH(7) +
                             SH_C2[4] * (real)2 * x * SH(8);
                dRGBdy[c] += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * (real)2 * -y * SH(6) +
                             SH_C2[4] * (real)2 * -y * SH(8);
                dRGBdz[c] += SH_C2[1] * y * SH(5) + SH_C2[2] * (real)2 * (real)2 * z * SH(6) + SH_C2[3] * x * SH(7);
            }
            if (deg > 2) {
                DSH(9, SH_C3[0] * y * ((real)3 * xx - yy));
                DSH(10, SH_C3[1] * xy * z);
                DSH(11, SH_C3[2] * y * ((real)4 * zz - xx - yy));
                DSH(12, SH_C3[3] * z * ((real)2 * zz - (real)3 * xx - (real)3 * yy));
                DSH(13, SH_C3[4] * x * ((real)4 * zz - xx - yy));
                DSH(14, SH_C3[5] * z * (xx - yy));
                DSH(15, SH_C3[6] * x * (xx - (real)3 * yy));
                for (int c = 0; c < 3; c++) {
                    dRGBdx[c] += (SH_C3[0] * SH(9) * (real)3 * (real)2 * xy + SH_C3[1] * SH(10) * yz +
                                  SH_C3[2] * SH(11) * (real)-2 * xy + SH_C3[3] * SH(12) * (real)-3 * (real)2 * xz +
                                  SH_C3[4] * SH(13) * ((real)-3 * xx + (real)4 * zz - yy) +
                                  SH_C3[5] * SH(14) * (real)2 * xz + SH_C3[6] * SH(15) * (real)3 * (xx - yy));
                    dRGBdy[c] += (SH_C3[0] * SH(9) * (real)3 * (xx - yy) + SH_C3[1] * SH(10) * xz +
                                  SH_C3[2] * SH(11) * ((real)-3 * yy + (real)4 * zz - xx) +
                                  SH_C3[3] * SH(12) * (real)-3 * (real)2 * yz + SH_C3[4] * SH(13) * (real)-2 * xy +
                                  SH_C3[5] * SH(14) * (real)-2 * yz + SH_C3[6] * SH(15) * (real)-3 * (real)2 * xy);
                    dRGBdz[c] += (SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * (real)4 * (real)2 * yz +
                                  SH_C3[3] * SH(12) * (real)3 * ((real)2 * zz - xx - yy) +
                                  SH_C3[4] * SH(13) * (real)4 * (real)2 * xz + SH_C3[5] * SH(14) * (xx - yy));
                }
            }
        }
    }
#undef SH
#undef DSH
    real dL_ddir[3] = {(dRGBdx[0] * dRGB[0] + dRGBdx[1] * dRGB[1]) + dRGBdx[2] * dRGB[2],
                       (dRGBdy[0] * dRGB[0] + dRGBdy[1] * dRGB[1]) + dRGBdy[2] * dRGB[2],
                       (dRGBdz[0] * dRGB[0] + dRGBdz[1] * dRGB[1]) + dRGBdz[2] * dRGB[2]};
    /* dnormvdv, CR/auxiliary.h:127-137 */
    real sum2 = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
    real inv = (real)1 / r_sqrt(sum2 * sum2 * sum2);
    real g0 = ((+sum2 - d0[0] * d0[0]) * dL_ddir[0] - d0[1] * d0[0] * dL_ddir[1] - d0[2] * d0[0] * dL_ddir[2]) * inv;
    real g1 = (-d0[0] * d0[1] * dL_ddir[0] + (sum2 - d0[1] * d0[1]) * dL_ddir[1] - d0[2] * d0[1] * dL_ddir[2]) * inv;
    real g2 = (-d0[0] * d0[2] * dL_ddir[0] - d0[1] * d0[2] * dL_ddir[1] + (sum2 - d0[2] * d0[2]) * dL_ddir[2]) * inv;
    dL_dmeans[3 * idx + 0] += g0;
    dL_dmeans[3 * idx + 1] += g1;
    dL_dmeans[3 * idx + 2] += g2;
}

/* CR/backward.cu:586-641  preprocessCUDA (backward) incl. compute_transmat_aabb (:443-584).
 * img_w/img_h/tanfov reproduce W = int(focal_x * tan_fovx * 2) in fp32 (quirk 1).
 * transMats = the matrices the blend used (computed or precomputed).  scales == NULL selects
 * the precomputed-transMat path (:620).  dL_dtransMats / dL_dmean2Ds are in/out.          */
void orc_project_backward(int P, int D, int M, const real* means3D, const real* transMats, const int* radii,
                          const real* shs, const uint8_t* clamped, const real* scales, const real* rotations,
                          const real* view, const real* proj, int img_w, int img_h, float tan_fovx,
                          float tan_fovy, const real* campos, real* dL_dtransMats, const real* dL_dnormal3Ds,
                          const real* dL_dcolors, real* dL_dshs, real* dL_dmean2Ds, real* dL_dmean3Ds,
                          real* dL_dscales, real* dL_drots) {
    /* CR/rasterizer_impl.cu:387-388 + CR/backward.cu:618-619, evaluated in fp32 on purpose */
    volatile float focal_y = (float)img_h / (2.0f * tan_fovy);
    volatile float focal_x = (float)img_w / (2.0f * tan_fovx);
    volatile float wx = focal_x * tan_fovx;
    volatile float wy = focal_y * tan_fovy;
    const int W = (int)(wx * 2.0f);
    const int H = (int)(wy * 2.0f);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        if (!(radii[idx] > 0)) continue;
        real T[3][3], Pm[3][4], R[3][3];
        real normal[3] = {0, 0, 0};
        const int precomp = (scales == NULL);
        real sx = 0, sy = 0;
        const real* p = means3D + 3 * idx;
        if (precomp) {
            for (int c = 0; c < 3; c++)
                for (int r = 0; r < 3; r++) T[c][r] = transMats[9 * idx + 3 * c + r];
        } else {
            sx = scales[2 * idx];
            sy = scales[2 * idx + 1];
            quat_to_rotmat(rotations + 4 * idx, R);
            build_T(p, sx, sy, R, proj, W, H, T, Pm); /* scale_modifier ignored: quirk 2 */
            for (int i = 0; i < 3; i++) normal[i] = view[i] * R[2][0] + view[4 + i] * R[2][1] + view[8 + i] * R[2][2];
        }
        real dT[3][3];
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) dT[c][r] = dL_dtransMats[9 * idx + 3 * c + r];
        const real m2x = dL_dmean2Ds[3 * idx], m2y = dL_dmean2Ds[3 * idx + 1];
        if (m2x != 0 || m2y != 0) {
            const real distance = T[2][0] * T[2][0] + T[2][1] * T[2][1] - T[2][2] * T[2][2];
            const real f = (real)1 / distance;
            const real dpx_dT00 = f * T[2][0], dpx_dT01 = f * T[2][1], dpx_dT02 = -f * T[2][2];
            const real dpy_dT10 = f * T[2][0], dpy_dT11 = f * T[2][1], dpy_dT12 = -f * T[2][2];
            const real dpx_dT30 = T[0][0] * (f - (real)2 * f * f * T[2][0] * T[2][0]);
            const real dpx_dT31 = T[0][1] * (f - (real)2 * f * f * T[2][1] * T[2][1]);
            const real dpx_dT32 = -T[0][2] * (f + (real)2 * f * f * T[2][2] * T[2][2]);
            const real dpy_dT30 = T[1][0] * (f - (real)2 * f * f * T[2][0] * T[2][0]);
            const real dpy_dT31 = T[1][1] * (f - (real)2 * f * f * T[2][1] * T[2][1]);
            const real dpy_dT32 = -T[1][2] * (f + (real)2 * f * f * T[2][2] * T[2][2]);
            dT[0][0] += m2x * dpx_dT00; dT[0][1] += m2x * dpx_dT01; dT[0][2] += m2x * dpx_dT02;
            dT[1][0] += m2y * dpy_dT10; dT[1][1] += m2y * dpy_dT11; dT[1][2] += m2y * dpy_dT12;
            dT[2][0] += m2x * dpx_dT30 + m2y * dpy_dT30;
            dT[2][1] += m2x * dpx_dT31 + m2y * dpy_dT31;
            dT[2][2] += m2x * dpx_dT32 + m2y * dpy_dT32;
            if (precomp)
                for (int c = 0; c < 3; c++)
                    for (int r = 0; r < 3; r++) dL_dtransMats[9 * idx + 3 * c + r] = dT[c][r];
        }
        if (!precomp) {
            /* dL_dM = P(mat3x4) * transpose(dL_dT)(mat3): dL_dM[c][k] = sum_j P[j][k] * dT[j][c] */
            real dM[3][4];
            for (int c = 0; c < 3; c++)
                for (int k = 0; k < 4; k++) dM[c][k] = Pm[0][k] * dT[0][c] + Pm[1][k] * dT[1][c] + Pm[2][k] * dT[2][c];
            const real* dn = dL_dnormal3Ds + 3 * idx;
            real dtn[3];
            for (int i = 0; i < 3; i++) dtn[i] = view[4 * i] * dn[0] + view[4 * i + 1] * dn[1] + view[4 * i + 2] * dn[2];
            real pv[3];
            for (int i = 0; i < 3; i++) pv[i] = view[i] * p[0] + view[4 + i] * p[1] + view[8 + i] * p[2] + view[12 + i];
            real cosv = -((pv[0] * normal[0] + pv[1] * normal[1]) + pv[2] * normal[2]);
            real mult = cosv > 0 ? (real)1 : (real)-1;
            for (int i = 0; i < 3; i++) dtn[i] = mult * dtn[i];
            real dRS[3][3];
            for (int i = 0; i < 3; i++) { dRS[0][i] = dM[0][i]; dRS[1][i] = dM[1][i]; dRS[2][i] = dtn[i]; }
            real dR[3][3];
            for (int i = 0; i < 3; i++) { dR[0][i] = dRS[0][i] * sx; dR[1][i] = dRS[1][i] * sy; dR[2][i] = dRS[2][i]; }
            quat_to_rotmat_vjp(rotations + 4 * idx, dR, dL_drots + 4 * idx);
            dL_dscales[2 * idx] = (dRS[0][0] * R[0][0] + dRS[0][1] * R[0][1]) + dRS[0][2] * R[0][2];
            dL_dscales[2 * idx + 1] = (dRS[1][0] * R[1][0] + dRS[1][1] * R[1][1]) + dRS[1][2] * R[1][2];
            for (int i = 0; i < 3; i++) dL_dmean3Ds[3 * idx + i] = dM[2][i];
        }
        if (shs) sh_backward(idx, D, M, means3D, campos, shs, clamped, dL_dcolors, dL_dmean3Ds, dL_dshs);
        /* densification proxy, quirk 5 (CR/backward.cu:637-640) */
        const real depth = transMats[9 * idx + 8];
        dL_dmean2Ds[3 * idx + 0] = dL_dtransMats[9 * idx + 2] * depth * (real)0.5 * (real)W;
        dL_dmean2Ds[3 * idx + 1] = dL_dtransMats[9 * idx + 5] * depth * (real)0.5 * (real)H;
    }
}

int orc_real_size(void) { return (int)sizeof(real); }

/* Number of OpenMP threads the parallel loops use (1 => deterministic accumulation order). */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

