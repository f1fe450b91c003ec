// common.cuh -- shared constants, buffer layouts and small device helpers of the B200 surfel
// rasterizer.  Semantics follow the reference (RAST = 2d-gaussian-splatting/submodules/
// diff-surfel-rasterization, CR = RAST/cuda_rasterizer); code and data layout are our own.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace g4s {

// ---- constants of the spec (CR/config.h:14-16, CR/auxiliary.h:21-39) -------------------------
constexpr int TILE = 16;                 // binning tile edge in pixels (BLOCK_X = BLOCK_Y = 16)
constexpr int TILE_PIX = TILE * TILE;    // 256 pixels = 8 warps
constexpr int REGION_W = 8;              // a warp owns an 8x4 pixel region of the tile
constexpr int REGION_H = 4;
constexpr float NEAR_N = 0.2f;           // near_n
constexpr float FAR_N = 100.0f;          // far_n
constexpr float FILTER_INV_SQUARE = 2.0f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_MIN = 0.0001f;

// ---- projected record: 7 x float4 = 112 B per Gaussian ----------------------------------------
// q0 = contribution bbox in pixels (xmin, ymin, xmax, ymax): pixels outside can never reach
//      alpha >= 1/255 for this Gaussian (exact culling, see project.cu)
// q1 = (Tu.x, Tu.y, Tu.z, Tv.x)   q2 = (Tv.y, Tv.z, Tw.x, Tw.y)   q3 = (Tw.z, cx, cy, opacity)
// q4 = (nx, ny, nz, r)            q5 = (g, b, By, rr2)
// q6 = (Axx, Axy, Ayy, Bx): with By, the contribution ellipse in pixel offsets from (cx, cy),
//      Axx x^2 + 2 Axy xy + Ayy y^2 + 2 Bx x + 2 By y - 1 <= 0 (all zero: no ellipse test possible);
//      rr2 = squared radius of the low-pass disc around (cx, cy).  Only the forward reads q6.
constexpr int REC_F4 = 7;
constexpr int REC_FLOATS = REC_F4 * 4;

// ---- blend-stage gradient accumulator: 6 x float4 = 96 B per Gaussian (21 floats used) ----------------
// The nine dT sums are carried as MOMENTS of q = dL/dp (CR/backward.cu:396-426) about a per-Gaussian origin (ox, oy),
// which costs the blend kernel 17 scalar operations per lane instead of two cross products and three dot products
// per pixel; project_bwd turns them into dT once per Gaussian:
//   [0..2]  Q0 = sum q            [3..5]  Qx = sum (px - ox) q       [6..8]  Qy = sum (py - oy) q
//   [9..11] Z  = sum (zs sx, zs sy, dL_dz)
//   [12,13] dmean2D   [14] dopacity   [15..17] dcolor   [18..20] dnormal   [21..23] unused
// with k_o = ox Tw - Tu, l_o = oy Tw - Tv (k, l of the origin pixel):
//   dTu = Qy x Tw + Q0 x l_o      dTv = Tw x Qx + k_o x Q0      dTw = Z - (ox dTu + oy dTv + Qx x l_o + k_o x Qy)
// (the dx dy terms cancel exactly).  The origin is the splat centre clamped into the image (moment_origin): every
// pixel offset then stays below the image size, so the sums have the magnitudes of the reference's per-pixel
// terms even for splats whose centre projects thousands of pixels off screen.
constexpr int ACC_F4 = 6;
constexpr int ACC_FLOATS = ACC_F4 * 4;
constexpr int ACC_USED = 21;

__host__ __device__ inline float moment_origin(float centre, float max_coord) { return fminf(fmaxf(centre, 0.0f), max_coord); }

// ---- counters (device int32[8] inside the image buffer) ---------------------------------------
enum { CNT_RENDERED = 0, CNT_MAXLEN = 1, CNT_VISIBLE = 2, CNT_PREFILTER_VIOLATION = 3, CNT_LONG_TILES = 4, CNT_N = 8 };

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Geometry buffer (per Gaussian).  All sections 256-byte aligned.
struct GeomView {
    float4* rec;        // [P][REC_F4]
    float* depth;       // [P]
    uint32_t* ntiles;   // [P] tiles touched after exact culling
    ushort4* rect;      // [P] culled tile rectangle (x0, y0, x1, y1), x1/y1 exclusive
    uint8_t* clamped;   // [P] bit c set when SH colour channel c was clamped at 0
    uint32_t* visible_list;   // [P] ids of the Gaussians with radii > 0 (counters[CNT_VISIBLE] of them), in no particular order
};
__host__ __device__ inline size_t geom_layout(int P, char* base, GeomView* v) {
    size_t off = 0;
    size_t n = (size_t)(P > 0 ? P : 1);
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_rec = take(n * REC_FLOATS * sizeof(float));
    size_t o_depth = take(n * sizeof(float));
    size_t o_nt = take(n * sizeof(uint32_t));
    size_t o_rect = take(n * sizeof(ushort4));
    size_t o_cl = take(n);
    size_t o_vl = take(n * sizeof(uint32_t));
    if (v) {
        v->rec = (float4*)(base + o_rec);
        v->depth = (float*)(base + o_depth);
        v->ntiles = (uint32_t*)(base + o_nt);
        v->rect = (ushort4*)(base + o_rect);
        v->clamped = (uint8_t*)(base + o_cl);
        v->visible_list = (uint32_t*)(base + o_vl);
    }
    return off;
}

// Image buffer (per pixel + per tile).
struct ImageView {
    float* final_T;         // [3][N]: T, M1, M2
    uint32_t* n_contrib;    // [2][N]: last contributor, median contributor (1-based list pos)
    uint32_t* tile_count;   // [T]  instances per tile (plan); reused as scatter cursor
    uint32_t* tile_offset;  // [T+1] exclusive scan
    uint32_t* tile_order;   // [T]  tile ids, longest list first (blend launch order)
    int32_t* counters;      // [CNT_N]
};
__host__ __device__ inline size_t image_layout(int W, int H, char* base, ImageView* v) {
    size_t N = (size_t)W * H;
    size_t T = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_ft = take(3 * N * sizeof(float));
    size_t o_nc = take(2 * N * sizeof(uint32_t));
    size_t o_tc = take(T * sizeof(uint32_t));
    size_t o_to = take((T + 1) * sizeof(uint32_t));
    size_t o_ord = take(T * sizeof(uint32_t));
    size_t o_cnt = take(CNT_N * sizeof(int32_t));
    if (v) {
        v->final_T = (float*)(base + o_ft);
        v->n_contrib = (uint32_t*)(base + o_nc);
        v->tile_count = (uint32_t*)(base + o_tc);
        v->tile_offset = (uint32_t*)(base + o_to);
        v->tile_order = (uint32_t*)(base + o_ord);
        v->counters = (int32_t*)(base + o_cnt);
    }
    return off;
}

// Binning buffer (per instance).
struct BinView {
    uint32_t* list;            // [cap]     gaussian ids, per tile sorted by (depth, id)
    uint32_t* masks;           // [cap * 8] per tile [8 regions][n entries]: lanes of the region that blended the entry (forward)
    unsigned long long* keys;  // [cap]     depth_bits << 32 | gaussian, bucketed by tile, unsorted
};
__host__ __device__ inline size_t bin_layout(int64_t cap, char* base, BinView* v) {
    size_t n = (size_t)(cap > 0 ? cap : 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_l = take(n * sizeof(uint32_t));
    size_t o_m = take(n * 8 * sizeof(uint32_t));
    size_t o_k = take(n * sizeof(unsigned long long));
    if (v) {
        v->list = (uint32_t*)(base + o_l);
        v->masks = (uint32_t*)(base + o_m);
        v->keys = (unsigned long long*)(base + o_k);
    }
    return off;
}

// ---- tiny vector helpers (componentwise, evaluation order spelled out) ------------------------
struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 mul3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ f3 scale3(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float sum3(f3 a) { return a.x + a.y + a.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// Can a pixel rectangle [rx0,rx1] x [ry0,ry1] (inclusive pixel centres) hold a pixel that reaches
// alpha >= 1/255 for this splat?  Conservative (never false for a contributing rectangle):
// the rectangle meets the low-pass disc, or the minimum of the ellipse form over it is <= 0.
// The form is convex (Axx > 0, det > 0), so the minimum is the unconstrained minimiser when that
// lies inside, else it lies on one of the four edges (a clamped 1-D parabola each).
__device__ __forceinline__ bool rect_may_contribute(float cx, float cy, float4 cn, float By, float rr2,
                                                    float rx0, float ry0, float rx1, float ry1) {
    const float x0 = rx0 - cx, x1 = rx1 - cx, y0 = ry0 - cy, y1 = ry1 - cy;
    const float nx = fminf(fmaxf(0.f, x0), x1), ny = fminf(fmaxf(0.f, y0), y1);
    if (!(nx * nx + ny * ny > rr2)) return true;
    const float Axx = cn.x, Axy = cn.y, Ayy = cn.z, Bx = cn.w;
    const float det = Axx * Ayy - Axy * Axy;
    if (!(Axx > 0.f) || !(Ayy > 0.f) || !(det > 0.f)) return true;   // no usable ellipse: keep
    const float inv = 1.0f / det;
    const float xm = (Axy * By - Ayy * Bx) * inv, ym = (Axy * Bx - Axx * By) * inv;
    if (xm >= x0 && xm <= x1 && ym >= y0 && ym <= y1) return true;
    auto form = [&](float x, float y) {
        return x * (Axx * x + 2.f * (Axy * y + Bx)) + y * (Ayy * y + 2.f * By) - 1.f;
    };
    const float iyy = 1.0f / Ayy, ixx = 1.0f / Axx;
    auto cl = [](float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); };
    float fmin = form(x0, cl(-(Axy * x0 + By) * iyy, y0, y1));
    fmin = fminf(fmin, form(x1, cl(-(Axy * x1 + By) * iyy, y0, y1)));
    fmin = fminf(fmin, form(cl(-(Axy * y0 + Bx) * ixx, x0, x1), y0));
    fmin = fminf(fmin, form(cl(-(Axy * y1 + Bx) * ixx, x0, x1), y1));
    return !(fmin > 0.f);
}

// ---- rounding-pinned arithmetic -----------------------------------------------------------------
// Everything that feeds a hard decision of the reference (alpha >= 1/255, T < 1e-4, T > 0.5,
// ceil(radius), tile rectangles) is written with explicit _rn intrinsics in the order nvcc's FMA
// contraction gives the reference's expressions, so that the result does not depend on how this
// code happens to be inlined or split into basic blocks:
// The sequences below were read off the SASS / PTX of the reference kernels compiled with the same
// toolchain (nvcc 12.9, sm_100a): which product of a sum is rounded and which is fused is NOT a
// fixed rule (it depends on use counts and on ptxas' own mul+add fusion), so each expression names
// the helper that reproduces its actual sequence:
//   dot2_rn(a,b,c,d)         fma(a, b, round(c*d))
//   dot3_rn(a,b,c,d,e,f)     fma(e, f, fma(a, b, round(c*d)))      (second product rounded)
//   dot3_first_rn(a,...,f)   fma(e, f, fma(c, d, round(a*b)))      (first product rounded)
//   diff2_rn(a,b,c,d)        fma(a, b, -round(c*d))
__device__ __forceinline__ float dot2_rn(float a, float b, float c, float d) { return __fmaf_rn(a, b, __fmul_rn(c, d)); }
__device__ __forceinline__ float dot3_rn(float a, float b, float c, float d, float e, float f) {
    return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
__device__ __forceinline__ float dot3_first_rn(float a, float b, float c, float d, float e, float f) {
    return __fmaf_rn(e, f, __fmaf_rn(c, d, __fmul_rn(a, b)));
}
__device__ __forceinline__ float diff2_rn(float a, float b, float c, float d) { return __fmaf_rn(a, b, -__fmul_rn(c, d)); }

// view/projection matrices are 16 floats, memory = column-major of the column-vector matrix
// (CR/auxiliary.h:78-121)
__device__ __forceinline__ f3 xform_point_4x3(f3 p, const float* m) {
    return mk3(__fadd_rn(dot3_rn(m[0], p.x, m[4], p.y, m[8], p.z), m[12]),
               __fadd_rn(dot3_rn(m[1], p.x, m[5], p.y, m[9], p.z), m[13]),
               __fadd_rn(dot3_rn(m[2], p.x, m[6], p.y, m[10], p.z), m[14]));
}
__device__ __forceinline__ f3 xform_vec_4x3(f3 p, const float* m) {
    return mk3(dot3_rn(m[0], p.x, m[4], p.y, m[8], p.z),
               dot3_rn(m[1], p.x, m[5], p.y, m[9], p.z),
               dot3_rn(m[2], p.x, m[6], p.y, m[10], p.z));
}
__device__ __forceinline__ f3 xform_vec_4x3_T(f3 p, const float* m) {
    return mk3(m[0] * p.x + m[1] * p.y + m[2] * p.z,
               m[4] * p.x + m[5] * p.y + m[6] * p.z,
               m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// Rotation matrix of a (w,x,y,z) quaternion, normalised with rsqrtf like the reference
// (CR/auxiliary.h:212-234).  R[c] = column c.
__device__ __forceinline__ void quat_to_R(float4 q /* x=w y=x z=y w=z */, f3 R[3]) {
    // reference SASS: n2 = fma(q2,q2, fma(q1,q1, fma(q0,q0, round(q3*q3))));  the products with w are
    // rounded and the other product of each pair fused; y*y + z*z is a plain add of rounded squares,
    // x*x is fused into x*x + z*z and x*x + y*y; 2*() is an exact doubling, 1 - () a plain subtraction.
    const float n2 = __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmaf_rn(q.x, q.x, __fmul_rn(q.w, q.w))));
    const float s = rsqrtf(n2);
    const float w = __fmul_rn(q.x, s), x = __fmul_rn(q.y, s), y = __fmul_rn(q.z, s), z = __fmul_rn(q.w, s);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    auto twice = [](float v) { return __fadd_rn(v, v); };
    const float wz = __fmul_rn(w, z), wy = __fmul_rn(w, y), wx = __fmul_rn(w, x);
    R[0] = mk3(__fsub_rn(1.f, twice(__fadd_rn(yy, zz))), twice(__fmaf_rn(x, y, wz)), twice(__fmaf_rn(x, z, -wy)));
    R[1] = mk3(twice(__fmaf_rn(x, y, -wz)), __fsub_rn(1.f, twice(__fmaf_rn(x, x, zz))), twice(__fmaf_rn(y, z, wx)));
    R[2] = mk3(twice(__fmaf_rn(x, z, wy)), twice(__fmaf_rn(y, z, -wx)), __fsub_rn(1.f, twice(__fmaf_rn(x, x, yy))));
}

// Tangent-plane -> pixel homography T = (Tu, Tv, Tw) (CR/forward.cu:75-115, glm evaluation order):
//   X[c][r] = sum_j S[r][j] * proj[c + 4j]  with S[0] = (sx*R0, 0), S[1] = (sy*R1, 0), S[2] = (p, 1)
//   Tu[r] = X[0][r]*W/2 + X[3][r]*(W-1)/2 ; Tv[r] = X[1][r]*H/2 + X[3][r]*(H-1)/2 ; Tw[r] = X[3][r]
// The zero entries of splat2world / ndc2pix contribute exact zeros and are left out.
__device__ __forceinline__ void build_T(f3 p, float sx, float sy, const f3 R[3], const float* pm,
                                        int W, int H, f3& Tu, f3& Tv, f3& Tw) {
    f3 L0 = mk3(__fmul_rn(R[0].x, sx), __fmul_rn(R[0].y, sx), __fmul_rn(R[0].z, sx));
    f3 L1 = mk3(__fmul_rn(R[1].x, sy), __fmul_rn(R[1].y, sy), __fmul_rn(R[1].z, sy));
    float X[4][3];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        X[c][0] = dot3_rn(L0.x, pm[c], L0.y, pm[c + 4], L0.z, pm[c + 8]);
        X[c][1] = dot3_rn(L1.x, pm[c], L1.y, pm[c + 4], L1.z, pm[c + 8]);
        X[c][2] = __fadd_rn(dot3_rn(p.x, pm[c], p.y, pm[c + 4], p.z, pm[c + 8]), pm[c + 12]);
    }
    const float hw = float(W) / 2.0f, hw1 = float(W - 1) / 2.0f;
    const float hh = float(H) / 2.0f, hh1 = float(H - 1) / 2.0f;
    // The reference sums X0*hw + X1*0 + X2*0 + X3*hw1 left to right; nvcc contracts that to
    // fma(X3, hw1, round(X0*hw)) (the zero terms are exact no-ops).  Spelled with intrinsics so
    // that the rounding points are the reference's: the ray-splat solve is ill-conditioned
    // (pix*Tw - Tu cancels ~3 digits), one ulp in T moves alpha by 1e-5 and flips thresholds.
    Tu = mk3(__fmaf_rn(X[3][0], hw1, __fmul_rn(X[0][0], hw)), __fmaf_rn(X[3][1], hw1, __fmul_rn(X[0][1], hw)),
             __fmaf_rn(X[3][2], hw1, __fmul_rn(X[0][2], hw)));
    Tv = mk3(__fmaf_rn(X[3][0], hh1, __fmul_rn(X[1][0], hh)), __fmaf_rn(X[3][1], hh1, __fmul_rn(X[1][1], hh)),
             __fmaf_rn(X[3][2], hh1, __fmul_rn(X[1][2], hh)));
    Tw = mk3(X[3][0], X[3][1], X[3][2]);
}

}  // namespace g4s
