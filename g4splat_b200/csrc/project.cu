// project.cu -- per-Gaussian stages: forward projection (+ exact tile culling + per-tile counts),
// instance scatter, backward projection, markVisible.
//
// Reference semantics restated from (CR = RAST/cuda_rasterizer):
//   forward   CR/forward.cu:150-253 preprocessCUDA, :75-115 compute_transmat, :119-147 compute_aabb,
//             :20-71 computeColorFromSH, CR/auxiliary.h:66-76 getRect, :184-209 in_frustum
//   scatter   CR/rasterizer_impl.cu:70-111 duplicateWithKeys
//   backward  CR/backward.cu:586-641 preprocessCUDA, :443-584 compute_transmat_aabb, :20-139 SH
//   visible   CR/rasterizer_impl.cu:54-66 checkFrustum
// What is different by design: the tile set of a Gaussian is the reference rectangle intersected
// with the bounding box of the pixels that can reach alpha >= 1/255 ("contribution bbox"), so
// the per-tile lists only hold instances that can change a pixel; per-tile counts are
// accumulated here, so no prefix sum over Gaussians and no global 64-bit sort is needed.
#include "kernels.cuh"

namespace g4s {

// SH basis constants (CR/auxiliary.h:42-59)
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                  -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                  0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                  -0.5900435899266435f};

// Degree-D real SH -> RGB for one Gaussian, + 0.5, clamp at 0 (CR/forward.cu:20-71).  Rounding pinned
// to the reference's SASS: every term is accumulated with one fma(coef, sh, r) in ascending coefficient
// order; the coefficients are rounded products, with xx*3 - yy, zz*4 - xx, 2zz - 3xx - 3yy and xx - 3yy fused
// as fma(.., +-3|4, ..).  sh_basis() produces the 15 coefficients kf[1..15] (kf[0] is SH_C0 itself).
__device__ __forceinline__ void sh_basis(int deg, f3 dir, float (&kf)[16]) {
    if (deg > 0) {
        const float x = dir.x, y = dir.y, z = dir.z;
        kf[1] = -__fmul_rn(SH_C1, y);
        kf[2] = __fmul_rn(SH_C1, z);
        kf[3] = -__fmul_rn(SH_C1, x);
        if (deg > 1) {
            const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
            const float xy = __fmul_rn(x, y), yz = __fmul_rn(y, z), xz = __fmul_rn(x, z);
            const float zz2 = __fadd_rn(zz, zz), xx_yy = __fsub_rn(xx, yy);
            kf[4] = __fmul_rn(SH_C2[0], xy);
            kf[5] = __fmul_rn(SH_C2[1], yz);
            kf[6] = __fmul_rn(SH_C2[2], __fsub_rn(__fsub_rn(zz2, xx), yy));
            kf[7] = __fmul_rn(SH_C2[3], xz);
            kf[8] = __fmul_rn(SH_C2[4], xx_yy);
            if (deg > 2) {
                const float zz4_xx_yy = __fsub_rn(__fmaf_rn(zz, 4.0f, -xx), yy);
                kf[9] = __fmul_rn(__fmul_rn(SH_C3[0], y), __fmaf_rn(xx, 3.0f, -yy));
                kf[10] = __fmul_rn(__fmul_rn(SH_C3[1], xy), z);
                kf[11] = __fmul_rn(__fmul_rn(SH_C3[2], y), zz4_xx_yy);
                kf[12] = __fmul_rn(__fmul_rn(SH_C3[3], z), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2)));
                kf[13] = __fmul_rn(__fmul_rn(SH_C3[4], x), zz4_xx_yy);
                kf[14] = __fmul_rn(__fmul_rn(SH_C3[5], z), xx_yy);
                kf[15] = __fmul_rn(__fmul_rn(SH_C3[6], x), __fmaf_rn(yy, -3.0f, xx));
            }
        }
    }
}
__device__ __forceinline__ f3 sh_finish(f3 r, uint8_t& clamp_mask) {
    r = mk3(__fadd_rn(r.x, 0.5f), __fadd_rn(r.y, 0.5f), __fadd_rn(r.z, 0.5f));
    clamp_mask = (uint8_t)((r.x < 0 ? 1 : 0) | (r.y < 0 ? 2 : 0) | (r.z < 0 ? 4 : 0));
    return mk3(fmaxf(r.x, 0.0f), fmaxf(r.y, 0.0f), fmaxf(r.z, 0.0f));
}
// Coefficient 0 is read from `sh0`, coefficients i >= 1 from `shr[3 (i - 1) ..]`: one [M][3] array
// (shr = sh0 + 3) for the operator API, the trainer's separate _features_dc / _features_rest tensors
// for the raw-parameter entry points.
__device__ __forceinline__ f3 sh_to_rgb(int deg, const float* __restrict__ sh0, const float* __restrict__ shr, f3 dir,
                                        uint8_t& clamp_mask) {
    float kf[16];
    sh_basis(deg, dir, kf);
    f3 r = mk3(__fmul_rn(SH_C0, sh0[0]), __fmul_rn(SH_C0, sh0[1]), __fmul_rn(SH_C0, sh0[2]));
    const int n = (deg + 1) * (deg + 1);
#pragma unroll
    for (int i = 1; i < 16; i++) {
        if (i < n) {
            r.x = __fmaf_rn(kf[i], shr[3 * i - 3], r.x); r.y = __fmaf_rn(kf[i], shr[3 * i - 2], r.y); r.z = __fmaf_rn(kf[i], shr[3 * i - 1], r.z);
        }
    }
    return sh_finish(r, clamp_mask);
}
// The same evaluation for a 16-byte aligned [16][3] row (the usual case: 192-byte rows of a torch tensor):
// twelve 128-bit loads instead of 48 scalar ones, four coefficients per three loads, same order.
__device__ __forceinline__ f3 sh_to_rgb_row16(int deg, const float4* __restrict__ s4, f3 dir, uint8_t& clamp_mask) {
    float kf[16];
    sh_basis(deg, dir, kf);
    auto acc = [](f3& r, float k, float s0, float s1, float s2) {
        r.x = __fmaf_rn(k, s0, r.x); r.y = __fmaf_rn(k, s1, r.y); r.z = __fmaf_rn(k, s2, r.z);
    };
    float4 a = s4[0];
    f3 r = mk3(__fmul_rn(SH_C0, a.x), __fmul_rn(SH_C0, a.y), __fmul_rn(SH_C0, a.z));
    if (deg > 0) {
        float4 b = s4[1], c = s4[2];
        acc(r, kf[1], a.w, b.x, b.y); acc(r, kf[2], b.z, b.w, c.x); acc(r, kf[3], c.y, c.z, c.w);
        if (deg > 1) {
            a = s4[3]; b = s4[4]; c = s4[5];
            acc(r, kf[4], a.x, a.y, a.z); acc(r, kf[5], a.w, b.x, b.y); acc(r, kf[6], b.z, b.w, c.x); acc(r, kf[7], c.y, c.z, c.w);
            a = s4[6];
            acc(r, kf[8], a.x, a.y, a.z);
            if (deg > 2) {
                b = s4[7]; c = s4[8];
                acc(r, kf[9], a.w, b.x, b.y); acc(r, kf[10], b.z, b.w, c.x); acc(r, kf[11], c.y, c.z, c.w);
                a = s4[9]; b = s4[10]; c = s4[11];
                acc(r, kf[12], a.x, a.y, a.z); acc(r, kf[13], a.w, b.x, b.y); acc(r, kf[14], b.z, b.w, c.x); acc(r, kf[15], c.y, c.z, c.w);
            }
        }
    }
    return sh_finish(r, clamp_mask);
}

// Bounding box (in pixels, inclusive, already padded) of every pixel for which this Gaussian can
// reach alpha >= 1/255, i.e. min(rho3d, rho2d) <= rho_cut with rho_cut = 2 ln(255 opacity) padded.
//   rho2d <= rho_cut : disc of radius sqrt(rho_cut / 2) around the low-pass centre (cx, cy)
//   rho3d <= rho_cut : image of the tangent-plane disc u^2 + v^2 <= rho_cut under the homography
//                      T; its axis-aligned extent has the closed form of CR/forward.cu:119-147
//                      with cutoff^2 = rho_cut.  Evaluated in coordinates shifted to (cx, cy) so
//                      that the centre^2 - second-moment cancellation stays small.
// Returns false when the Gaussian can never contribute (opacity < 1/255).
__device__ __forceinline__ bool contribution_bbox(f3 Tu, f3 Tv, f3 Tw, float cx, float cy, float opacity,
                                                  float4& bb, float4& conic, float& conic_By, float& rr2) {
    conic = make_float4(0.f, 0.f, 0.f, 0.f);
    conic_By = 0.f;
    rr2 = 1e30f;
    if (!(opacity >= ALPHA_MIN)) return false;  // alpha <= opacity < 1/255 for every pixel
    const float rho_cut = 2.0f * logf(255.0f * opacity) * 1.001f + 1e-3f;
    const float big = 1e30f;
    // low-pass disc
    const float rr = sqrtf(0.5f * rho_cut);
    rr2 = 0.5f * rho_cut * 1.001f + 1e-3f;
    float x0 = -rr, x1 = rr, y0 = -rr, y1 = rr;
    // projected tangent disc, shifted frame: Tu' = Tu - cx Tw, Tv' = Tv - cy Tw
    const f3 U = mk3(Tu.x - cx * Tw.x, Tu.y - cx * Tw.y, Tu.z - cx * Tw.z);
    const f3 V = mk3(Tv.x - cy * Tw.x, Tv.y - cy * Tw.y, Tv.z - cy * Tw.z);
    const float d = rho_cut * (Tw.x * Tw.x + Tw.y * Tw.y) - Tw.z * Tw.z;
    bool bounded = false;
    if (d < 0.0f) {  // the disc does not reach the camera plane: its image is an ellipse
        const float fxy = rho_cut / d, fz = -1.0f / d;
        const float mx = fxy * (U.x * Tw.x + U.y * Tw.y) + fz * U.z * Tw.z;
        const float my = fxy * (V.x * Tw.x + V.y * Tw.y) + fz * V.z * Tw.z;
        const float hx = mx * mx - (fxy * (U.x * U.x + U.y * U.y) + fz * U.z * U.z);
        const float hy = my * my - (fxy * (V.x * V.x + V.y * V.y) + fz * V.z * V.z);
        if (hx >= 0.0f && hy >= 0.0f && hx < 1e12f && hy < 1e12f) {
            const float ex = sqrtf(hx) * 1.0005f, ey = sqrtf(hy) * 1.0005f;
            x0 = fminf(x0, mx - ex); x1 = fmaxf(x1, mx + ex);
            y0 = fminf(y0, my - ey); y1 = fmaxf(y1, my + ey);
            bounded = true;
            // the same ellipse as a quadratic form in pixel offsets from (cx, cy):
            // p = x m0 + y m1 + m2 (the homography's adjugate columns), p.x^2 + p.y^2 - rho_cut p.z^2 <= 0
            const f3 m0 = cross3(V, Tw), m1 = cross3(Tw, U), m2 = cross3(U, V);
            const float C0 = m2.x * m2.x + m2.y * m2.y - rho_cut * m2.z * m2.z;
            if (C0 < 0.0f) {
                const float sc = -1.0f / C0;
                const float Axx = (m0.x * m0.x + m0.y * m0.y - rho_cut * m0.z * m0.z) * sc;
                const float Ayy = (m1.x * m1.x + m1.y * m1.y - rho_cut * m1.z * m1.z) * sc;
                const float Axy = (m0.x * m1.x + m0.y * m1.y - rho_cut * m0.z * m1.z) * sc;
                const float Bx = (m0.x * m2.x + m0.y * m2.y - rho_cut * m0.z * m2.z) * sc;
                const float By = (m1.x * m2.x + m1.y * m2.y - rho_cut * m1.z * m2.z) * sc;
                const bool finite = fabsf(Axx) < 1e30f && fabsf(Ayy) < 1e30f && fabsf(Axy) < 1e30f &&
                                    fabsf(Bx) < 1e30f && fabsf(By) < 1e30f;
                if (finite && Axx > 0.0f && Ayy > 0.0f && Axx * Ayy - Axy * Axy > 0.0f) {
                    conic = make_float4(Axx, Axy, Ayy, Bx);
                    conic_By = By;
                }
            }
        }
    }
    if (!bounded) { bb = make_float4(-big, -big, big, big); return true; }
    const float pad = 0.02f + 2e-6f * (fabsf(cx) + fabsf(cy));
    bb = make_float4(cx + x0 - pad, cy + y0 - pad, cx + x1 + pad, cy + y1 + pad);
    if (!(bb.x == bb.x && bb.y == bb.y && bb.z == bb.z && bb.w == bb.w)) bb = make_float4(-big, -big, big, big);
    return true;
}

// Parameter activations of the trainer (2DGS/scene/gaussian_model.py:158-192), applied in registers by
// the raw-parameter entry points instead of ~10 torch kernels per call:
//   scaling  = exp(_scaling)                       [+ mip filter: sqrt(scaling^2 + filter^2)]
//   rotation = _rotation / max(|_rotation|, 1e-12) (torch.nn.functional.normalize)
//   opacity  = sigmoid(_opacity)                   [+ mip filter: * sqrt(det1 / det2),
//              det1 = prod exp(_scaling)^2, det2 = prod (exp(_scaling)^2 + filter^2)]
// Each operation is rounded where torch rounds it (one kernel per operator there).
struct Activated {
    float2 scale;      // activated scaling
    float2 e2;         // exp(_scaling)^2
    float4 rot;        // normalised rotation
    float inv_norm;    // 1 / max(|_rotation|, eps)
    float opacity, sig, coef, f2;   // activated opacity = sig * coef; f2 = filter^2 (0 without mip filter)
};
__device__ __forceinline__ Activated activate(float2 s_raw, float4 r_raw, float o_raw, const float* mip_filter, int idx) {
    Activated a;
    const float e0 = expf(s_raw.x), e1 = expf(s_raw.y);
    a.e2 = make_float2(__fmul_rn(e0, e0), __fmul_rn(e1, e1));
    a.sig = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-o_raw)));
    if (mip_filter != nullptr) {
        const float f = mip_filter[idx];
        a.f2 = __fmul_rn(f, f);
        const float d0 = __fadd_rn(a.e2.x, a.f2), d1 = __fadd_rn(a.e2.y, a.f2);
        a.scale = make_float2(__fsqrt_rn(d0), __fsqrt_rn(d1));
        a.coef = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.e2.x, a.e2.y), __fmul_rn(d0, d1)));
        a.opacity = __fmul_rn(a.sig, a.coef);
    } else {
        a.f2 = 0.0f;
        a.scale = make_float2(e0, e1);
        a.coef = 1.0f;
        a.opacity = a.sig;
    }
    const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r_raw.x, r_raw.x), __fmul_rn(r_raw.y, r_raw.y)),
                                         __fmul_rn(r_raw.z, r_raw.z)), __fmul_rn(r_raw.w, r_raw.w));
    const float n = fmaxf(__fsqrt_rn(n2), 1e-12f);
    a.inv_norm = __fdiv_rn(1.0f, n);
    a.rot = make_float4(__fdiv_rn(r_raw.x, n), __fdiv_rn(r_raw.y, n), __fdiv_rn(r_raw.z, n), __fdiv_rn(r_raw.w, n));
    return a;
}

// Per-thread part of the forward projection.  Returns false when the Gaussian is culled (radii 0).
// ---- forward projection in two phases ---------------------------------------------------------------
// Phase 1 (every Gaussian): view transform, frustum cull, T matrix, normal flip, 3-sigma radius, reference
// tile rectangle.  Phase 2 (survivors only, ~1 in 4 at c2): colour from SH, contribution bounding box and
// ellipse, culled tile rectangle, record.  Between the two the survivors of the CTA are compacted through
// shared memory, so phase 2 -- two thirds of the instructions and all of the SH traffic -- runs on dense
// warps instead of on 8 warps with 7 of 32 lanes alive.
struct ProjGeo {                     // what phase 1 hands to phase 2 (PG_WORDS 32-bit words)
    f3 Tu, Tv, Tw, normal;
    float cx, cy, depth, opacity;    // opacity: activated opacity in raw mode, unused otherwise
    int radius, rx0, ry0, rx1, ry1;  // reference tile rectangle (non-empty)
};
constexpr int PG_WORDS = 20;

struct ProjOut {
    f3 rgb;
    float opacity;
    float4 bb, conic;
    float conic_By, rr2;
    int rx0, ry0, rx1, ry1;          // culled tile rectangle: reference rect x contribution bbox (may be empty)
    uint8_t clamp_mask;
};

template <bool RAW>
__device__ __forceinline__ bool project_geometry(const ProjectArgs& a, int idx, ProjGeo& o) {
    const f3 p = mk3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
    // scale and rotation are fetched together with the mean, before the frustum test decides whether they
    // are needed: one memory round trip per Gaussian instead of two (24 wasted bytes for a culled one)
    float2 sc = make_float2(0.f, 0.f);
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (a.transMat_precomp == nullptr) {
        sc = ((const float2*)a.scales)[idx];
        q = ((const float4*)a.rotations)[idx];
    }
    const f3 pv = xform_point_4x3(p, a.view);
    if (pv.z <= 0.2f) {  // in_frustum (CR/auxiliary.h:199)
        if (a.prefiltered) atomicAdd(&a.counters[CNT_PREFILTER_VIOLATION], 1);
        return false;
    }
    f3 Tu, Tv, Tw, normal;
    o.opacity = 0.0f;
    if (a.transMat_precomp == nullptr) {
        if (RAW) {
            const Activated act = activate(sc, q, a.opacities[idx], a.mip_filter, idx);
            sc = act.scale; q = act.rot; o.opacity = act.opacity;
        }
        f3 R[3];
        quat_to_R(q, R);
        build_T(p, __fmul_rn(a.scale_modifier, sc.x), __fmul_rn(a.scale_modifier, sc.y), R, a.proj, a.W, a.H, Tu, Tv, Tw);
        normal = xform_vec_4x3(R[2], a.view);
    } else {
        const float* t = a.transMat_precomp + 9 * (size_t)idx;
        Tu = mk3(t[0], t[1], t[2]); Tv = mk3(t[3], t[4], t[5]); Tw = mk3(t[6], t[7], t[8]);
        normal = mk3(0.0f, 0.0f, 1.0f);
    }
    // dual-visible flip (CR/forward.cu:211-216)
    const float cosv = -dot3_rn(pv.x, normal.x, pv.y, normal.y, pv.z, normal.z);
    if (cosv == 0) return false;
    const float mult = cosv > 0 ? 1.0f : -1.0f;
    normal = mk3(__fmul_rn(mult, normal.x), __fmul_rn(mult, normal.y), __fmul_rn(mult, normal.z));

    // 3-sigma AABB: centre + radius exactly as the reference (CR/forward.cu:119-147, :222-233).
    // radius = ceil() of a cancelling difference: rounding points pinned to the reference's SASS
    // (dist = fma(-Tw.z, Tw.z, fma(Tw.x^2, 9, round(Tw.y^2 * 9))); sums round their FIRST product).
    const float tw2x = __fmul_rn(Tw.x, Tw.x), tw2y = __fmul_rn(Tw.y, Tw.y);
    const float dist = __fmaf_rn(-Tw.z, Tw.z, __fmaf_rn(tw2x, 9.0f, __fmul_rn(tw2y, 9.0f)));
    const float inv_dist = __frcp_rn(dist);
    if (dist == 0.0f) return false;
    const float f9 = __fmul_rn(inv_dist, 9.0f);
    const f3 fu = mk3(__fmul_rn(f9, Tu.x), __fmul_rn(f9, Tu.y), __fmul_rn(inv_dist, -Tu.z));
    const f3 fv = mk3(__fmul_rn(f9, Tv.x), __fmul_rn(f9, Tv.y), __fmul_rn(inv_dist, -Tv.z));
    const float cx = dot3_first_rn(fu.x, Tw.x, fu.y, Tw.y, fu.z, Tw.z);
    const float cy = dot3_first_rn(fv.x, Tw.x, fv.y, Tw.y, fv.z, Tw.z);
    const float t0 = dot3_first_rn(fu.x, Tu.x, fu.y, Tu.y, fu.z, Tu.z);
    const float t1 = dot3_first_rn(fv.x, Tv.x, fv.y, Tv.y, fv.z, Tv.z);
    const float ex = sqrtf(fmaxf(1e-4f, __fmaf_rn(cx, cx, -t0)));
    const float ey = sqrtf(fmaxf(1e-4f, __fmaf_rn(cy, cy, -t1)));
    const float radius = ceilf(fmaxf(ex, ey));

    // getRect (CR/auxiliary.h:66-76)
    const int max_radius = (int)radius;
    const int gx = a.grid_x, gy = a.grid_y;
    const int rx0 = min(gx, max(0, (int)((cx - max_radius) / TILE)));
    const int ry0 = min(gy, max(0, (int)((cy - max_radius) / TILE)));
    const int rx1 = min(gx, max(0, (int)((cx + max_radius + TILE - 1) / TILE)));
    const int ry1 = min(gy, max(0, (int)((cy + max_radius + TILE - 1) / TILE)));
    if ((rx1 - rx0) * (ry1 - ry0) == 0) return false;
    o.Tu = Tu; o.Tv = Tv; o.Tw = Tw; o.normal = normal;
    o.cx = cx; o.cy = cy; o.depth = pv.z; o.radius = max_radius;
    o.rx0 = rx0; o.ry0 = ry0; o.rx1 = rx1; o.ry1 = ry1;
    return true;
}

template <bool RAW>
__device__ __forceinline__ void project_appearance(const ProjectArgs& a, int idx, const ProjGeo& g, ProjOut& o) {
    // colour
    o.clamp_mask = 0;
    if (a.colors_precomp == nullptr) {
        const f3 p = mk3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
        f3 dir = sub3(p, mk3(a.campos[0], a.campos[1], a.campos[2]));
        const float len = __fsqrt_rn(dot3_rn(dir.x, dir.x, dir.y, dir.y, dir.z, dir.z));
        dir = mk3(__fdiv_rn(dir.x, len), __fdiv_rn(dir.y, len), __fdiv_rn(dir.z, len));
        if (RAW) {
            o.rgb = sh_to_rgb(a.D, a.shs + (size_t)idx * 3, a.sh_rest + (size_t)idx * (a.M - 1) * 3, dir, o.clamp_mask);
        } else {
            const float* sh = a.shs + (size_t)idx * a.M * 3;
            if (a.M == 16 && (reinterpret_cast<uintptr_t>(a.shs) & 15) == 0)
                o.rgb = sh_to_rgb_row16(a.D, reinterpret_cast<const float4*>(sh), dir, o.clamp_mask);
            else
                o.rgb = sh_to_rgb(a.D, sh, sh + 3, dir, o.clamp_mask);
        }
    } else {
        o.rgb = mk3(a.colors_precomp[3 * idx], a.colors_precomp[3 * idx + 1], a.colors_precomp[3 * idx + 2]);
    }
    o.opacity = RAW ? g.opacity : a.opacities[idx];

    // exact culling, step 1: tiles of the reference rectangle that hold a pixel of the contribution bbox
    int rx0 = g.rx0, ry0 = g.ry0, rx1 = g.rx1, ry1 = g.ry1;
    const bool can_contribute = contribution_bbox(g.Tu, g.Tv, g.Tw, g.cx, g.cy, o.opacity, o.bb, o.conic, o.conic_By, o.rr2);
    if (can_contribute) {
        // tile t covers pixels [16 t, 16 t + 15]
        const float fx0 = fmaxf(ceilf((o.bb.x - (TILE - 1)) / TILE), (float)rx0);
        const float fy0 = fmaxf(ceilf((o.bb.y - (TILE - 1)) / TILE), (float)ry0);
        const float fx1 = fminf(floorf(o.bb.z / TILE) + 1.0f, (float)rx1);
        const float fy1 = fminf(floorf(o.bb.w / TILE) + 1.0f, (float)ry1);
        rx0 = (int)fx0; ry0 = (int)fy0; rx1 = (int)fx1; ry1 = (int)fy1;
    }
    if (!can_contribute || rx1 <= rx0 || ry1 <= ry0) { rx0 = ry0 = rx1 = ry1 = 0; }
    o.rx0 = rx0; o.ry0 = ry0; o.rx1 = rx1; o.ry1 = ry1;
}

// One thread per Gaussian for the arithmetic; the per-tile counting of phase 2 is warp-cooperative: for
// each Gaussian of the warp with a long rectangle in turn, the 32 lanes take 32 tiles of it, so close-up
// splats that cover thousands of tiles do not serialise on one thread.
constexpr int PF_THREADS = 256;
template <bool RAW>
__global__ void __launch_bounds__(PF_THREADS, 4) project_fwd_kernel(ProjectArgs a) {
    __shared__ uint32_t s_geo[PG_WORDS][PF_THREADS];   // SoA: slot-major reads and writes are conflict-free
    __shared__ int s_wcount[PF_THREADS / 32];
    const int idx = blockIdx.x * PF_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- phase 1 ------------------------------------------------------------------------------
    ProjGeo g;
    const bool visible = idx < a.P && project_geometry<RAW>(a, idx, g);
    if (idx < a.P) {
        a.radii[idx] = visible ? g.radius : 0;
        if (!visible) a.geom.ntiles[idx] = 0u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, visible);
    if (lane == 0) s_wcount[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PF_THREADS / 32; w++) {
        const int c = s_wcount[w];
        if (w < warp) base += c;
        total += c;
    }
    if (visible) {
        const int slot = base + __popc(bal & ((1u << lane) - 1u));
        const float f[15] = {g.Tu.x, g.Tu.y, g.Tu.z, g.Tv.x, g.Tv.y, g.Tv.z, g.Tw.x, g.Tw.y, g.Tw.z,
                             g.normal.x, g.normal.y, g.normal.z, g.cx, g.cy, g.depth};
#pragma unroll
        for (int k = 0; k < 15; k++) s_geo[k][slot] = __float_as_uint(f[k]);
        s_geo[15][slot] = __float_as_uint(g.opacity);
        s_geo[16][slot] = (uint32_t)g.radius;
        s_geo[17][slot] = (uint32_t)g.rx0 | ((uint32_t)g.ry0 << 16);
        s_geo[18][slot] = (uint32_t)g.rx1 | ((uint32_t)g.ry1 << 16);
        s_geo[19][slot] = (uint32_t)idx;
    }
    __syncthreads();
    // the survivors also go into the view's list of visible Gaussians (the backward projection walks it): every warp
    // reserves room for its share of the compacted survivors
    const int warp_share = min(32, max(0, total - warp * 32));
    int list_base = 0;
    if (lane == 0 && warp_share > 0) list_base = atomicAdd(&a.counters[CNT_VISIBLE], warp_share);
    list_base = __shfl_sync(0xffffffffu, list_base, 0);

    // ---- phase 2: thread t takes the t-th survivor of the CTA ---------------------------------------
    const int t = threadIdx.x;
    const bool active = t < total;
    if (__ballot_sync(0xffffffffu, active) == 0u) return;   // whole warp idle
    ProjOut o;
    o.rx0 = o.ry0 = o.rx1 = o.ry1 = 0;
    int id2 = 0;
    if (active) {
        float f[15];
#pragma unroll
        for (int k = 0; k < 15; k++) f[k] = __uint_as_float(s_geo[k][t]);
        g.Tu = mk3(f[0], f[1], f[2]); g.Tv = mk3(f[3], f[4], f[5]); g.Tw = mk3(f[6], f[7], f[8]);
        g.normal = mk3(f[9], f[10], f[11]); g.cx = f[12]; g.cy = f[13]; g.depth = f[14];
        g.opacity = __uint_as_float(s_geo[15][t]);
        g.radius = (int)s_geo[16][t];
        const uint32_t r0 = s_geo[17][t], r1 = s_geo[18][t];
        g.rx0 = (int)(r0 & 0xffffu); g.ry0 = (int)(r0 >> 16); g.rx1 = (int)(r1 & 0xffffu); g.ry1 = (int)(r1 >> 16);
        id2 = (int)s_geo[19][t];
        a.geom.visible_list[list_base + lane] = (uint32_t)id2;
        project_appearance<RAW>(a, id2, g, o);
    }
    const int gx = a.grid_x;
    const int my_w = o.rx1 - o.rx0, my_total = active ? my_w * (o.ry1 - o.ry0) : 0;
    // per-tile counters.  Rectangles of up to 32 tiles are counted by their own lane (fire-and-forget
    // reductions); longer ones (close-up splats cover thousands of tiles) by the whole warp.
    // (An exact ellipse-vs-tile test here was measured: it removes ~12 % of the instances of this
    // workload but costs more in this kernel than it saves downstream; the blend kernel applies the
    // same test per 8x4 region, where it is nearly free.)
    constexpr int SERIAL_TILES = 32;
    if (my_total > 0 && my_total <= SERIAL_TILES) {
        for (int y = o.ry0; y < o.ry1; y++)
            for (int x = o.rx0; x < o.rx1; x++) atomicAdd(&a.tile_count[y * gx + x], 1u);
    }
    unsigned pending = __ballot_sync(0xffffffffu, my_total > SERIAL_TILES);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int rx0 = __shfl_sync(0xffffffffu, o.rx0, src), ry0 = __shfl_sync(0xffffffffu, o.ry0, src);
        const int w = __shfl_sync(0xffffffffu, my_w, src), tot = __shfl_sync(0xffffffffu, my_total, src);
        for (int i = lane; i < tot; i += 32) {
            const int iy = i / w, ix = i - iy * w;
            atomicAdd(&a.tile_count[(ry0 + iy) * gx + rx0 + ix], 1u);
        }
    }
    if (!active) return;
    a.geom.ntiles[id2] = (uint32_t)my_total;
    if (my_total == 0) { o.rx0 = o.ry0 = o.rx1 = o.ry1 = 0; o.bb = make_float4(1e30f, 1e30f, -1e30f, -1e30f); }
    float4* rec = a.geom.rec + (size_t)id2 * REC_F4;
    rec[0] = o.bb;
    rec[1] = make_float4(g.Tu.x, g.Tu.y, g.Tu.z, g.Tv.x);
    rec[2] = make_float4(g.Tv.y, g.Tv.z, g.Tw.x, g.Tw.y);
    rec[3] = make_float4(g.Tw.z, g.cx, g.cy, o.opacity);
    rec[4] = make_float4(g.normal.x, g.normal.y, g.normal.z, o.rgb.x);
    rec[5] = make_float4(o.rgb.y, o.rgb.z, o.conic_By, o.rr2);
    rec[6] = o.conic;
    a.geom.depth[id2] = g.depth;
    a.geom.clamped[id2] = o.clamp_mask;
    a.geom.rect[id2] = make_ushort4((unsigned short)o.rx0, (unsigned short)o.ry0, (unsigned short)o.rx1, (unsigned short)o.ry1);
}

// One (Gaussian, tile) instance per surviving tile: key = depth bits << 32 | Gaussian id, written
// to an arbitrary free slot of the tile's bucket; the per-tile sort orders the bucket afterwards.
// The instances of a warp's 32 Gaussians are spread evenly over its lanes: an exclusive scan of the tile
// counts numbers them, lane l takes instances l, l + 32, ... and finds the owning Gaussian by searching
// the scan, so all cursor atomics of the warp (they return a value: ~1 us each) are in flight together
// instead of one Gaussian after the other.
__global__ void __launch_bounds__(256) scatter_kernel(ScatterArgs a) {
    if ((int64_t)a.counters[CNT_RENDERED] > a.capacity) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = idx < a.P && a.geom.ntiles[idx] != 0;
    ushort4 r = make_ushort4(0, 0, 0, 0);
    unsigned long long key = 0ull;
    if (live) {
        r = a.geom.rect[idx];
        key = (((unsigned long long)__float_as_uint(a.geom.depth[idx])) << 32) | (unsigned long long)(uint32_t)idx;
    }
    if (__ballot_sync(0xffffffffu, live) == 0u) return;
    const int my_w = r.z - r.x, my_total = my_w * (r.w - r.y);
    // inclusive scan of the instance counts over the warp
    int incl = my_total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - my_total;
    for (int i = lane; i - lane < warp_total; i += 32) {   // uniform trip count: the shuffles below need every lane
        // owner = the last lane whose exclusive offset is <= i (binary search over the scan by shuffles)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int probe = lo + step;
            const int e = __shfl_sync(0xffffffffu, excl, probe & 31);
            if (probe < 32 && e <= i) lo = probe;
        }
        // (lanes with my_total == 0 share their offset with the next live lane: the search lands on the
        //  last of them, i.e. on the live one, because it takes the LAST lane with excl <= i)
        const int o_excl = __shfl_sync(0xffffffffu, excl, lo);
        const int o_w = __shfl_sync(0xffffffffu, my_w, lo);
        const int o_x0 = __shfl_sync(0xffffffffu, (int)r.x, lo), o_y0 = __shfl_sync(0xffffffffu, (int)r.y, lo);
        const unsigned long long k = __shfl_sync(0xffffffffu, key, lo);
        if (i < warp_total) {
            const int j = i - o_excl;
            const int iy = j / o_w, ix = j - iy * o_w;
            const int t = (o_y0 + iy) * a.grid_x + o_x0 + ix;
            const uint32_t slot = a.tile_offset[t] + atomicAdd(&a.tile_cursor[t], 1u);
            a.keys[slot] = k;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward of the SH colour (CR/backward.cu:20-139): writes dL_dsh[idx][k] for k < (D+1)^2 and
// zeros above, returns the view-direction term to add to dL_dmean3D.
// `row4` (may be null): the Gaussian's whole [16][3] coefficient row when it is 16-byte aligned -- the usual case, 192-byte
// rows of a torch tensor -- read with twelve 128-bit loads instead of 45 scalar ones (a scalar load of a warp touches
// one cache line per lane: the loads of this function were a third of the kernel's L1 wavefronts).
__device__ __forceinline__ f3 sh_backward(int deg, int M, const float* __restrict__ shr /* coefficients 1.. */,
                                          const float4* __restrict__ row4, f3 dir_orig,
                                          uint8_t clamp_mask, f3 dL_dcolor, float* __restrict__ dsh) {
    float row[48];   // row[3 k + c] = coefficient k, channel c; filled per degree block, compile-time indices only
    auto load_block = [&](int q0, int q1) {
        if (row4 != nullptr) {
#pragma unroll
            for (int q = 0; q < 12; q++) {
                if (q >= q0 && q < q1) {
                    const float4 v = row4[q];
                    row[4 * q] = v.x; row[4 * q + 1] = v.y; row[4 * q + 2] = v.z; row[4 * q + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int j = 3; j < 48; j++)
                if (j >= 4 * q0 && j < 4 * q1 && j < 3 * M) row[j] = shr[j - 3];
        }
    };
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    f3 g = mk3(dL_dcolor.x * ((clamp_mask & 1) ? 0.f : 1.f), dL_dcolor.y * ((clamp_mask & 2) ? 0.f : 1.f),
               dL_dcolor.z * ((clamp_mask & 4) ? 0.f : 1.f));
    auto c = [&](int k) { return mk3(row[3 * k], row[3 * k + 1], row[3 * k + 2]); };   // k >= 1 only
    auto put = [&](int k, float v) { dsh[3 * k] = v * g.x; dsh[3 * k + 1] = v * g.y; dsh[3 * k + 2] = v * g.z; };
    auto axpy = [&](f3& acc, float s, f3 v) { acc.x += s * v.x; acc.y += s * v.y; acc.z += s * v.z; };
    f3 dx = mk3(0, 0, 0), dy = mk3(0, 0, 0), dz = mk3(0, 0, 0);
    put(0, SH_C0);
    int used = 1;
    if (deg > 0) {
        load_block(0, 3);
        put(1, -SH_C1 * y); put(2, SH_C1 * z); put(3, -SH_C1 * x);
        dx = scale3(-SH_C1, c(3)); dy = scale3(-SH_C1, c(1)); dz = scale3(SH_C1, c(2));
        used = 4;
        if (deg > 1) {
            load_block(3, 7);
            const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            put(4, SH_C2[0] * xy); put(5, SH_C2[1] * yz); put(6, SH_C2[2] * (2.f * zz - xx - yy));
            put(7, SH_C2[3] * xz); put(8, SH_C2[4] * (xx - yy));
            axpy(dx, SH_C2[0] * y, c(4)); axpy(dx, SH_C2[2] * 2.f * -x, c(6)); axpy(dx, SH_C2[3] * z, c(7)); axpy(dx, SH_C2[4] * 2.f * x, c(8));
            axpy(dy, SH_C2[0] * x, c(4)); axpy(dy, SH_C2[1] * z, c(5)); axpy(dy, SH_C2[2] * 2.f * -y, c(6)); axpy(dy, SH_C2[4] * 2.f * -y, c(8));
            axpy(dz, SH_C2[1] * y, c(5)); axpy(dz, SH_C2[2] * 2.f * 2.f * z, c(6)); axpy(dz, SH_C2[3] * x, c(7));
            used = 9;
            if (deg > 2) {
                load_block(7, 12);
                put(9, SH_C3[0] * y * (3.f * xx - yy)); put(10, SH_C3[1] * xy * z);
                put(11, SH_C3[2] * y * (4.f * zz - xx - yy)); put(12, SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy));
                put(13, SH_C3[4] * x * (4.f * zz - xx - yy)); put(14, SH_C3[5] * z * (xx - yy));
                put(15, SH_C3[6] * x * (xx - 3.f * yy));
                axpy(dx, SH_C3[0] * 3.f * 2.f * xy, c(9)); axpy(dx, SH_C3[1] * yz, c(10)); axpy(dx, SH_C3[2] * -2.f * xy, c(11));
                axpy(dx, SH_C3[3] * -3.f * 2.f * xz, c(12)); axpy(dx, SH_C3[4] * (-3.f * xx + 4.f * zz - yy), c(13));
                axpy(dx, SH_C3[5] * 2.f * xz, c(14)); axpy(dx, SH_C3[6] * 3.f * (xx - yy), c(15));
                axpy(dy, SH_C3[0] * 3.f * (xx - yy), c(9)); axpy(dy, SH_C3[1] * xz, c(10)); axpy(dy, SH_C3[2] * (-3.f * yy + 4.f * zz - xx), c(11));
                axpy(dy, SH_C3[3] * -3.f * 2.f * yz, c(12)); axpy(dy, SH_C3[4] * -2.f * xy, c(13));
                axpy(dy, SH_C3[5] * -2.f * yz, c(14)); axpy(dy, SH_C3[6] * -3.f * 2.f * xy, c(15));
                axpy(dz, SH_C3[1] * xy, c(10)); axpy(dz, SH_C3[2] * 4.f * 2.f * yz, c(11)); axpy(dz, SH_C3[3] * 3.f * (2.f * zz - xx - yy), c(12));
                axpy(dz, SH_C3[4] * 4.f * 2.f * xz, c(13)); axpy(dz, SH_C3[5] * (xx - yy), c(14));
                used = 16;
            }
        }
    }
    for (int k = used; k < M; k++) { dsh[3 * k] = 0.f; dsh[3 * k + 1] = 0.f; dsh[3 * k + 2] = 0.f; }
    const f3 ddir = mk3(dx.x * g.x + dx.y * g.y + dx.z * g.z, dy.x * g.x + dy.y * g.y + dy.z * g.z,
                        dz.x * g.x + dz.y * g.y + dz.z * g.z);
    // dnormvdv (CR/auxiliary.h:127-137)
    const f3 v = dir_orig;
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    return mk3(((+sum2 - v.x * v.x) * ddir.x - v.y * v.x * ddir.y - v.z * v.x * ddir.z) * inv,
               (-v.x * v.y * ddir.x + (sum2 - v.y * v.y) * ddir.y - v.z * v.y * ddir.z) * inv,
               (-v.x * v.z * ddir.x - v.y * v.z * ddir.y + (sum2 - v.z * v.z) * ddir.z) * inv);
}

// Backward projection: one thread per Gaussian.  EVERY output element is written (zeros when the
// Gaussian was not visible) so the caller never has to clear the gradient tensors.  A thread's
// 73 output floats are staged in shared memory in exactly the global layout of its warp's 32
// Gaussians, then the warp streams them out with fully coalesced stores (128-bit for the SH
// block, which is 2/3 of the bytes): a thread-per-Gaussian store would scatter every 4-byte
// word of a warp over 32 different cache lines.
// Reductions into a running multi-view / multi-rank sum.  MC == false: the destination is ordinary device
// memory (RED).  MC == true: the destination is an NVSwitch MULTICAST address that maps the same buffer of
// every rank (view_parallel, transport "multimem"): one multimem.red adds the value into all replicas inside
// the switch, so the per-step gradient all-reduce disappears into the kernel that produces the gradient.
__device__ __forceinline__ void red_add(float* p, float v, bool mc) {
    if (mc) asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else atomicAdd(p, v);   // result unused: fire-and-forget RED
}
__device__ __forceinline__ void red_add4(float* p, float4 v, bool mc) {
    if (mc) asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    else atomicAdd(reinterpret_cast<float4*>(p), v);
}
__device__ __forceinline__ void red_max(int* p, int v, bool mc) {
    if (mc) asm volatile("multimem.red.relaxed.sys.global.max.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else atomicMax(p, v);
}

constexpr int PB_THREADS = 128;
constexpr int PB_WARPS = PB_THREADS / 32;
constexpr int PB_SMALL = 25;      // means3D 3, means2D 3, colors 3, opacity 1, scales 2, rots 4, transMat 9
constexpr int PB_MAX_M = 16;
constexpr int PB_SH_STRIDE = PB_MAX_M * 3 + 1;   // odd row stride: conflict-free per-thread writes

// Everything between the blend-stage accumulators of ONE visible Gaussian and its 73 gradient values, staged into
// `my` (slots: 0-2 means3D, 3-5 means2D, 6-8 colors, 9 opacity, 10-11 scales, 12-15 rots, 16-24 transMat; zeroed by the
// caller) and `my_sh` (the SH row).
template <bool RAW>
__device__ __forceinline__ void project_bwd_one(const ProjectBwdArgs& a, int idx, float* __restrict__ my, float* __restrict__ my_sh) {
    const int M = a.M;
    // blend-stage accumulators
    const float4* acc = a.acc + (size_t)idx * ACC_F4;
    const float4 a0 = acc[0], a1 = acc[1], a2 = acc[2], a3 = acc[3], a4 = acc[4], a5 = acc[5];
    const float4* rec0 = a.geom.rec + (size_t)idx * REC_F4;
    float dT[3][3];   // dT[j] = d/d(Tu,Tv,Tw)[j]
    {
        // moments of q about the Gaussian's moment origin -> dT (common.cuh, accumulator layout)
        const float4 r1 = rec0[1], r2 = rec0[2], r3 = rec0[3];
        const f3 rTu = mk3(r1.x, r1.y, r1.z), rTv = mk3(r1.w, r2.x, r2.y), rTw = mk3(r2.z, r2.w, r3.x);
        const float ccx = moment_origin(r3.y, (float)(a.W - 1)), ccy = moment_origin(r3.z, (float)(a.H - 1));
        const f3 kc = sub3(scale3(ccx, rTw), rTu), lc = sub3(scale3(ccy, rTw), rTv);
        const f3 Q0 = mk3(a0.x, a0.y, a0.z), Qx = mk3(a0.w, a1.x, a1.y), Qy = mk3(a1.z, a1.w, a2.x), Z = mk3(a2.y, a2.z, a2.w);
        const f3 c1 = cross3(Qy, rTw), c2 = cross3(Q0, lc), c3 = cross3(rTw, Qx), c4 = cross3(kc, Q0);
        const f3 c5 = cross3(Qx, lc), c6 = cross3(kc, Qy);
        const f3 dTu = mk3(c1.x + c2.x, c1.y + c2.y, c1.z + c2.z), dTv = mk3(c3.x + c4.x, c3.y + c4.y, c3.z + c4.z);
        dT[0][0] = dTu.x; dT[0][1] = dTu.y; dT[0][2] = dTu.z;
        dT[1][0] = dTv.x; dT[1][1] = dTv.y; dT[1][2] = dTv.z;
        dT[2][0] = Z.x - (ccx * dTu.x + ccy * dTv.x + c5.x + c6.x);
        dT[2][1] = Z.y - (ccx * dTu.y + ccy * dTv.y + c5.y + c6.y);
        dT[2][2] = Z.z - (ccx * dTu.z + ccy * dTv.z + c5.z + c6.z);
    }
    const float raw_dT[9] = {dT[0][0], dT[0][1], dT[0][2], dT[1][0], dT[1][1], dT[1][2], dT[2][0], dT[2][1], dT[2][2]};
    const float m2x = a3.x, m2y = a3.y;
    const f3 dcol = mk3(a3.w, a4.x, a4.y);
    const f3 dnrm = mk3(a4.z, a4.w, a5.x);
    my[9] = a3.z;
    my[6] = dcol.x; my[7] = dcol.y; my[8] = dcol.z;

    // W, H as the reference rebuilds them in fp32 (CR/backward.cu:618-619)
    const int Wq = int(a.focal_x * a.tan_fovx * 2);
    const int Hq = int(a.focal_y * a.tan_fovy * 2);

    const float4* rec = a.geom.rec + (size_t)idx * REC_F4;
    const float4 q1 = rec[1], q2 = rec[2], q3 = rec[3];
    const bool precomp = (a.scales == nullptr);
    const f3 p = mk3(a.means3D[3 * idx], a.means3D[3 * idx + 1], a.means3D[3 * idx + 2]);
    f3 Tu, Tv, Tw, R[3], normal = mk3(0, 0, 0);
    float sx = 0, sy = 0;
    float Pm[3][4];
    float4 quat = make_float4(1, 0, 0, 0);
    Activated act;
    act.inv_norm = 1.0f;
    if (precomp) {
        Tu = mk3(q1.x, q1.y, q1.z); Tv = mk3(q1.w, q2.x, q2.y); Tw = mk3(q2.z, q2.w, q3.x);
    } else {
        float2 sc = ((const float2*)a.scales)[idx];
        quat = ((const float4*)a.rotations)[idx];
        if (RAW) {
            act = activate(sc, quat, a.opacities[idx], a.mip_filter, idx);
            sc = act.scale; quat = act.rot;
        }
        sx = sc.x; sy = sc.y;  // scale_modifier ignored on purpose (quirk 2, CR/backward.cu:481)
        quat_to_R(quat, R);
        // P = world2ndc * ndc2pix (mat3x4), T = transpose(M) * P (CR/backward.cu:490-504)
        const float nd[3][4] = {{float(Wq) / 2.0f, 0, 0, float(Wq - 1) / 2.0f},
                                {0, float(Hq) / 2.0f, 0, float(Hq - 1) / 2.0f},
                                {0, 0, 0, 1}};
        const float* pm = a.proj;
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < 4; k++)
                Pm[c][k] = pm[0 + 4 * k] * nd[c][0] + pm[1 + 4 * k] * nd[c][1] + pm[2 + 4 * k] * nd[c][2] + pm[3 + 4 * k] * nd[c][3];
        const f3 L0 = mk3(R[0].x * sx, R[0].y * sx, R[0].z * sx), L1 = mk3(R[1].x * sy, R[1].y * sy, R[1].z * sy);
        float Tm[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            Tm[c][0] = L0.x * Pm[c][0] + L0.y * Pm[c][1] + L0.z * Pm[c][2];
            Tm[c][1] = L1.x * Pm[c][0] + L1.y * Pm[c][1] + L1.z * Pm[c][2];
            Tm[c][2] = p.x * Pm[c][0] + p.y * Pm[c][1] + p.z * Pm[c][2] + Pm[c][3];
        }
        Tu = mk3(Tm[0][0], Tm[0][1], Tm[0][2]); Tv = mk3(Tm[1][0], Tm[1][1], Tm[1][2]); Tw = mk3(Tm[2][0], Tm[2][1], Tm[2][2]);
        normal = xform_vec_4x3(R[2], a.view);
    }
    // low-pass centre gradient -> dT (CR/backward.cu:515-541; cutoff-1 approximation, quirk 4)
    if (m2x != 0 || m2y != 0) {
        const float distance = Tw.x * Tw.x + Tw.y * Tw.y - Tw.z * Tw.z;
        const float f = 1 / distance;
        const float dpx_dT00 = f * Tw.x, dpx_dT01 = f * Tw.y, dpx_dT02 = -f * Tw.z;
        const float dpx_dT30 = Tu.x * (f - 2 * f * f * Tw.x * Tw.x);
        const float dpx_dT31 = Tu.y * (f - 2 * f * f * Tw.y * Tw.y);
        const float dpx_dT32 = -Tu.z * (f + 2 * f * f * Tw.z * Tw.z);
        const float dpy_dT30 = Tv.x * (f - 2 * f * f * Tw.x * Tw.x);
        const float dpy_dT31 = Tv.y * (f - 2 * f * f * Tw.y * Tw.y);
        const float dpy_dT32 = -Tv.z * (f + 2 * f * f * Tw.z * Tw.z);
        dT[0][0] += m2x * dpx_dT00; dT[0][1] += m2x * dpx_dT01; dT[0][2] += m2x * dpx_dT02;
        dT[1][0] += m2y * dpx_dT00; dT[1][1] += m2y * dpx_dT01; dT[1][2] += m2y * dpx_dT02;
        dT[2][0] += m2x * dpx_dT30 + m2y * dpy_dT30;
        dT[2][1] += m2x * dpx_dT31 + m2y * dpy_dT31;
        dT[2][2] += m2x * dpx_dT32 + m2y * dpy_dT32;
    }
    float proxy2, proxy5;
    if (precomp) {
        // dL_dtransMat is the gradient of the precomputed input and feeds the proxy after the
        // update above (CR/backward.cu:542-553)
        for (int j = 0; j < 3; j++) for (int c = 0; c < 3; c++) my[16 + 3 * j + c] = dT[j][c];
        proxy2 = dT[0][2]; proxy5 = dT[1][2];
    } else {
        // the reference returns the raw blend-stage accumulator here (dT before the low-pass update above)
#pragma unroll
        for (int j = 0; j < 9; j++) my[16 + j] = raw_dT[j];
        proxy2 = raw_dT[2]; proxy5 = raw_dT[5];
        // dL_dM[c][k] = sum_j P[j][k] * dT[j][c]
        float dM[3][3];
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int k = 0; k < 3; k++) dM[c][k] = Pm[0][k] * dT[0][c] + Pm[1][k] * dT[1][c] + Pm[2][k] * dT[2][c];
        f3 dtn = xform_vec_4x3_T(dnrm, a.view);
        const f3 pv = xform_point_4x3(p, a.view);
        const float cosv = -sum3(mul3(pv, normal));
        const float mult = cosv > 0 ? 1.0f : -1.0f;
        dtn = scale3(mult, dtn);
        // dL_dR columns: dRS0 * sx, dRS1 * sy, dtn ; quat_to_rotmat_vjp (CR/auxiliary.h:237-281)
        const f3 v0 = mk3(dM[0][0] * sx, dM[0][1] * sx, dM[0][2] * sx);
        const f3 v1 = mk3(dM[1][0] * sy, dM[1][1] * sy, dM[1][2] * sy);
        const f3 v2 = dtn;
        const float s = rsqrtf(quat.w * quat.w + quat.x * quat.x + quat.y * quat.y + quat.z * quat.z);
        const float w = quat.x * s, x = quat.y * s, y = quat.z * s, z = quat.w * s;
        const float vR00 = v0.x, vR01 = v0.y, vR02 = v0.z, vR10 = v1.x, vR11 = v1.y, vR12 = v1.z, vR20 = v2.x, vR21 = v2.y, vR22 = v2.z;
        my[12] = 2.f * (x * (vR12 - vR21) + y * (vR20 - vR02) + z * (vR01 - vR10));
        my[13] = 2.f * (-2.f * x * (vR11 + vR22) + y * (vR01 + vR10) + z * (vR02 + vR20) + w * (vR12 - vR21));
        my[14] = 2.f * (x * (vR01 + vR10) - 2.f * y * (vR00 + vR22) + z * (vR12 + vR21) + w * (vR20 - vR02));
        my[15] = 2.f * (x * (vR02 + vR20) + y * (vR12 + vR21) - 2.f * z * (vR00 + vR11) + w * (vR01 - vR10));
        my[10] = dM[0][0] * R[0].x + dM[0][1] * R[0].y + dM[0][2] * R[0].z;
        my[11] = dM[1][0] * R[1].x + dM[1][1] * R[1].y + dM[1][2] * R[1].z;
        my[0] = dM[2][0]; my[1] = dM[2][1]; my[2] = dM[2][2];
    }
    if (a.shs) {
        const f3 dir = sub3(p, mk3(a.campos[0], a.campos[1], a.campos[2]));
        const float* shr = RAW ? a.sh_rest + (size_t)idx * (M - 1) * 3 : a.shs + (size_t)idx * M * 3 + 3;
        // whole-row vector loads when the [M][3] rows are 192 bytes on a 16-byte aligned base
        const float4* row4 = (!RAW && M == 16 && (reinterpret_cast<uintptr_t>(a.shs) & 15u) == 0)
                                 ? reinterpret_cast<const float4*>(a.shs + (size_t)idx * 48) : nullptr;
        const f3 dmean = sh_backward(a.D, M, shr, row4, dir, a.geom.clamped[idx], dcol, my_sh);
        my[0] += dmean.x; my[1] += dmean.y; my[2] += dmean.z;
    }
    if (RAW && !precomp) {
        // chain rule through the activations (what autograd does after the reference operator):
        //   exp:        d/d_scaling = g e            [mip: scale = sqrt(e^2 + f^2): g e^2 / scale,
        //               plus the opacity's dependence on the scales: g_o opacity f^2 / (e^2 + f^2)]
        //   sigmoid:    d/d_opacity = g_o coef sig (1 - sig)
        //   normalize:  d/d_rotation = (g - q (q . g)) / max(|r|, eps)
        const float g_o = my[9];
        if (a.mip_filter != nullptr) {
            my[10] = my[10] * act.e2.x / act.scale.x + g_o * act.opacity * act.f2 / (act.e2.x + act.f2);
            my[11] = my[11] * act.e2.y / act.scale.y + g_o * act.opacity * act.f2 / (act.e2.y + act.f2);
        } else {
            my[10] *= act.scale.x;
            my[11] *= act.scale.y;
        }
        my[9] = g_o * act.coef * act.sig * (1.0f - act.sig);
        const float qg = quat.x * my[12] + quat.y * my[13] + quat.z * my[14] + quat.w * my[15];
        my[12] = (my[12] - quat.x * qg) * act.inv_norm; my[13] = (my[13] - quat.y * qg) * act.inv_norm;
        my[14] = (my[14] - quat.z * qg) * act.inv_norm; my[15] = (my[15] - quat.w * qg) * act.inv_norm;
    }
    // densification proxy (CR/backward.cu:637-640, quirk 5): depth = forward T[8]
    const float depth = q3.x;
    my[3] = proxy2 * depth * 0.5f * float(Wq);
    my[4] = proxy5 * depth * 0.5f * float(Hq);
}

template <bool RAW>
__global__ void __launch_bounds__(PB_THREADS, 6) project_bwd_kernel(ProjectBwdArgs a) {
    __shared__ float s_sh[PB_WARPS][32 * PB_SH_STRIDE];
    __shared__ float s_small[PB_WARPS][32 * PB_SMALL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_base = (blockIdx.x * PB_WARPS + warp) * 32;   // first Gaussian of this warp
    if (warp_base >= a.P) return;
    const int idx = warp_base + lane;
    const int M = a.M;
    const bool has_sh_out = a.dL_dsh != nullptr;
    float* my_sh = &s_sh[warp][lane * PB_SH_STRIDE];
    float* my = &s_small[warp][lane * PB_SMALL];
    // slots of `my`: 0-2 means3D, 3-5 means2D, 6-8 colors, 9 opacity, 10-11 scales, 12-15 rots, 16-24 transMat
#pragma unroll
    for (int i = 0; i < PB_SMALL; i++) my[i] = 0.f;
    const bool visible = idx < a.P && a.radii[idx] > 0;
    // Gaussians without an SH gradient are written as zeros straight from registers below
    const unsigned sh_rows = __ballot_sync(0xffffffffu, visible && a.shs != nullptr);
    if (visible) project_bwd_one<RAW>(a, idx, my, my_sh);
    __syncwarp();
    // ---- coalesced write-out of the warp's 32 Gaussians ---------------------------------------
    // accumulate bit set: the destination is a running sum over views (view_parallel's flat gradient
    // buffer): rows of visible Gaussians are added to, rows of invisible ones are not touched at all.
    const int nvalid = min(32, a.P - warp_base);
    const unsigned vis_rows = __ballot_sync(0xffffffffu, visible);
    const bool mc = (a.accumulate & ACC_MULTIMEM) != 0;
    auto stream_out = [&](float* dst, int width, int slot, bool accumulate) {
        float* g = dst + (size_t)warp_base * width;
        for (int e = lane; e < nvalid * width; e += 32) {
            const int row = e / width;
            const float v = s_small[warp][row * PB_SMALL + slot + (e - row * width)];
            if (!accumulate) g[e] = v;
            else if ((vis_rows >> row) & 1u) red_add(&g[e], v, mc);
        }
    };
    stream_out(a.dL_dmeans3D, 3, 0, a.accumulate & ACC_MEANS3D);
    stream_out(a.dL_dmeans2D, 3, 3, false);
    if (a.dL_dcolors) stream_out(a.dL_dcolors, 3, 6, false);       // null: colours came from SH, nobody reads this
    stream_out(a.dL_dopacity, 1, 9, a.accumulate & ACC_OPACITY);
    stream_out(a.dL_dscales, 2, 10, a.accumulate & ACC_SCALES);
    stream_out(a.dL_drots, 4, 12, a.accumulate & ACC_ROTATIONS);
    if (a.dL_dtransMat) stream_out(a.dL_dtransMat, 9, 16, false);  // null: T came from scales / rotations
    if (has_sh_out) {
        // `row` floats per Gaussian taken from staging offset `src_off` of its SH row
        auto write_sh = [&](float* dst, int row, int src_off) {
            const int total = nvalid * row;
            float* g = dst + (size_t)warp_base * row;
            int i = lane / row, c = lane - i * row;
            const bool acc_sh = a.accumulate & ACC_SH;
            if (!acc_sh) {
                for (int e = lane; e < total; e += 32) {      // every store instruction covers 128 contiguous bytes
                    g[e] = ((sh_rows >> i) & 1u) ? s_sh[warp][i * PB_SH_STRIDE + src_off + c] : 0.f;
                    c += 32;
                    while (c >= row) { c -= row; i++; }
                }
            } else if ((row & 3) == 0 && src_off == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                // running sum over views: only rows with a gradient are touched, 16 bytes per reduction
                // (the row belongs to this warp alone; RED is used because it does not wait for the load)
                const int quads = row >> 2;
                const int nlive = __popc(sh_rows);
                for (int f = lane; f < nlive * quads; f += 32) {
                    const int k = f / quads, c4 = (f - k * quads) * 4;
                    const int r = __fns(sh_rows, 0, k + 1);   // k-th row with a gradient
                    const float* src = &s_sh[warp][r * PB_SH_STRIDE + c4];
                    red_add4(g + (size_t)r * row + c4, make_float4(src[0], src[1], src[2], src[3]), mc);
                }
            } else {
                for (int e = lane; e < total; e += 32) {
                    if ((sh_rows >> i) & 1u) red_add(&g[e], s_sh[warp][i * PB_SH_STRIDE + src_off + c], mc);
                    c += 32;
                    while (c >= row) { c -= row; i++; }
                }
            }
        };
        if (RAW) {   // the trainer's two SH tensors: _features_dc [P,1,3] and _features_rest [P,M-1,3]
            write_sh(a.dL_dsh, 3, 0);
            if (M > 1) write_sh(a.dL_dsh_rest, (M - 1) * 3, 3);
        } else {
            write_sh(a.dL_dsh, M * 3, 0);
        }
    }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ view, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const f3 p = mk3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const f3 pv = xform_point_4x3(p, view);
    present[idx] = !(pv.z <= 0.2f);
}

// Densification bookkeeping the trainer runs after every backward (2DGS/scene/gaussian_model.py:
// 649-651 add_densification_stats + train_with_refine_depth.py:583 max_radii2D), fused in one pass:
//   visible = radii > 0;  accum += |dL_dmean2D.xy| on visible;  denom += visible;  max_radii = max(.)
// multimem != 0: accum / denom / max_radii are multicast addresses (see red_add): every rank's replica is updated.
__global__ void __launch_bounds__(256) densify_stats_kernel(int P, const float* __restrict__ dL_dmeans2D,
                                                            const int* __restrict__ radii, float* accum,
                                                            float* denom, int* max_radii, int multimem) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int r = radii[idx];
    if (r > 0) {
        const float gx = dL_dmeans2D[3 * idx], gy = dL_dmeans2D[3 * idx + 1];
        if (multimem) {
            red_add(&accum[idx], sqrtf(gx * gx + gy * gy), true);
            red_add(&denom[idx], 1.0f, true);
            red_max(&max_radii[idx], r, true);
        } else {
            accum[idx] += sqrtf(gx * gx + gy * gy);
            denom[idx] += 1.0f;
            if (r > max_radii[idx]) max_radii[idx] = r;
        }
    }
}

// ---- NVLS all-reduce of the view-sharded gradient buffer (view_parallel, transport "multimem") -----------------
// Every rank holds its own partial sums in a buffer that all ranks map at the same offset of one NVSwitch multicast
// object.  Rank r owns the r-th slice of the buffer: one multimem.ld_reduce pulls the slice's 16 bytes from EVERY
// rank and adds them inside the switch, one multimem.st pushes the sum back into every rank's replica -- each byte
// crosses each link once per direction (a ring all-reduce moves 2 (N-1)/N of the buffer in N-1 dependent hops).
// The int32 block behind the floats (max_radii2D) is combined with max instead of add.  The caller brackets the
// kernel with barriers (all partial sums complete before, all slices written after).
__global__ void __launch_bounds__(512) multimem_allreduce_kernel(float* mc, size_t q_begin, size_t q_end, int* mc_max, size_t m_begin,
                                                                size_t m_end) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent 16-byte reductions in flight per thread: the round trip through the switch is long
    constexpr int U = 4;
    size_t q = q_begin + first;
    for (; q + (U - 1) * stride < q_end; q += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + 4 * (q + u * stride)) : "memory");
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"l"(mc + 4 * (q + u * stride)), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
    }
    for (; q < q_end; q += stride) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc + 4 * q) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(mc + 4 * q), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    for (size_t i = m_begin + first; i < m_end; i += stride) {
        int v;
        int* p = mc_max + i;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    }
}
void launch_multimem_allreduce(float* mc, size_t n_floats, int* mc_max, size_t n_ints, int rank, int world, cudaStream_t s) {
    const size_t quads = n_floats / 4;
    const size_t qb = quads * rank / world, qe = quads * (rank + 1) / world;
    const size_t mb = n_ints * rank / world, me = n_ints * (rank + 1) / world;
    if (qe == qb && me == mb) return;
    multimem_allreduce_kernel<<<148 * 4, 512, 0, s>>>(mc, qb, qe, mc_max, mb, me);
    count_launch();
}

// The same backward for callers that ACCUMULATE every parameter gradient (view_parallel's gradient sink: rows of
// invisible Gaussians are not touched at all).  A view sees a fraction of the Gaussians (a quarter at c2), and the kernel
// above spends ~1900 instructions per warp of 32 consecutive rows whatever the number of live lanes; this one walks the
// forward's list of visible Gaussians instead, 32 of them per warp, and adds each staged row to wherever its Gaussian
// lives.  Outputs that are not accumulated (dL_dmeans2D, dL_dcolors, dL_dtransMat) are cleared by the launcher and the
// visible rows stored.  Warps are independent: no block barrier.
template <bool RAW>
__global__ void __launch_bounds__(PB_THREADS, 6) project_bwd_compact_kernel(ProjectBwdArgs a) {
    __shared__ float s_sh[PB_WARPS][32 * PB_SH_STRIDE];
    __shared__ float s_small[PB_WARPS][32 * PB_SMALL];
    __shared__ int s_gid[PB_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int V = a.counters[CNT_VISIBLE];
    const int M = a.M;
    const bool mc = (a.accumulate & ACC_MULTIMEM) != 0;
    float* my_sh = &s_sh[warp][lane * PB_SH_STRIDE];
    float* my = &s_small[warp][lane * PB_SMALL];
    {
        const int base = (blockIdx.x * PB_WARPS + warp) * 32;
        if (base >= V) return;
        const int k = base + lane;
        int idx = k < V ? (int)a.geom.visible_list[k] : -1;
        if (idx >= 0 && a.radii[idx] <= 0) idx = -1;          // the dense kernel's definition of "visible"
#pragma unroll
        for (int i = 0; i < PB_SMALL; i++) my[i] = 0.f;
        s_gid[warp][lane] = idx;
        if (idx >= 0) project_bwd_one<RAW>(a, idx, my, my_sh);
        __syncwarp();
        // small outputs: element e of the warp's [32][width] block -> row e / width of the staging
        auto scatter_out = [&](float* dst, int width, int slot, bool accumulate) {
            for (int e = lane; e < 32 * width; e += 32) {
                const int row = e / width, c = e - row * width;
                const int g = s_gid[warp][row];
                if (g < 0) continue;
                const float v = s_small[warp][row * PB_SMALL + slot + c];
                float* d = dst + (size_t)g * width + c;
                if (accumulate) red_add(d, v, mc);
                else *d = v;
            }
        };
        scatter_out(a.dL_dmeans3D, 3, 0, true);
        scatter_out(a.dL_dmeans2D, 3, 3, false);
        if (a.dL_dcolors) scatter_out(a.dL_dcolors, 3, 6, false);
        scatter_out(a.dL_dopacity, 1, 9, true);
        scatter_out(a.dL_dscales, 2, 10, true);
        scatter_out(a.dL_drots, 4, 12, true);
        if (a.dL_dtransMat) scatter_out(a.dL_dtransMat, 9, 16, false);
        if (a.dL_dsh != nullptr) {
            auto scatter_sh = [&](float* dst, int rowlen, int src_off) {
                if ((rowlen & 3) == 0 && src_off == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                    const int quads = rowlen >> 2;                       // 16 bytes per reduction
                    for (int f = lane; f < 32 * quads; f += 32) {
                        const int row = f / quads, c4 = (f - row * quads) * 4;
                        const int g = s_gid[warp][row];
                        if (g < 0) continue;
                        const float* src = &s_sh[warp][row * PB_SH_STRIDE + c4];
                        red_add4(dst + (size_t)g * rowlen + c4, make_float4(src[0], src[1], src[2], src[3]), mc);
                    }
                } else {
                    for (int e = lane; e < 32 * rowlen; e += 32) {
                        const int row = e / rowlen, c = e - row * rowlen;
                        const int g = s_gid[warp][row];
                        if (g < 0) continue;
                        red_add(dst + (size_t)g * rowlen + c, s_sh[warp][row * PB_SH_STRIDE + src_off + c], mc);
                    }
                }
            };
            if (RAW) {
                scatter_sh(a.dL_dsh, 3, 0);
                if (M > 1) scatter_sh(a.dL_dsh_rest, (M - 1) * 3, 3);
            } else {
                scatter_sh(a.dL_dsh, M * 3, 0);
            }
        }
    }
}

// Zero the blend-stage accumulator rows of the visible Gaussians only (the others are never read):
// 4 B read per Gaussian + 80 B written per visible one, instead of an 80 P byte memset.
__global__ void __launch_bounds__(256) acc_clear_kernel(int P, const int* __restrict__ radii, float4* __restrict__ acc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P || radii[idx] <= 0) return;
    float4* row = acc + (size_t)idx * ACC_F4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < ACC_F4; q++) row[q] = z;
}
void launch_acc_clear(int P, const int* radii, float4* acc, cudaStream_t s) {
    if (P <= 0) return;
    acc_clear_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, radii, acc);
    count_launch();
}

// ---- launchers --------------------------------------------------------------------------------
void launch_densify_stats(int P, const float* dL_dmeans2D, const int* radii, float* accum, float* denom,
                          int* max_radii, int multimem, cudaStream_t s) {
    if (P <= 0) return;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, dL_dmeans2D, radii, accum, denom, max_radii, multimem);
    count_launch();
}
void launch_project_fwd(const ProjectArgs& a, cudaStream_t s) {
    if (a.P <= 0) return;
    const int blocks = (a.P + PF_THREADS - 1) / PF_THREADS;
    if (a.raw) project_fwd_kernel<true><<<blocks, PF_THREADS, 0, s>>>(a);
    else project_fwd_kernel<false><<<blocks, PF_THREADS, 0, s>>>(a);
    count_launch();
}
void launch_scatter(const ScatterArgs& a, cudaStream_t s) {
    if (a.P <= 0) return;
    scatter_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(a);
    count_launch();
}
void launch_project_bwd(const ProjectBwdArgs& a, cudaStream_t s) {
    if (a.P <= 0) return;
    // every parameter gradient accumulated (the gradient sink of view_parallel): walk the visible list
    constexpr int ACC_ALL = ACC_MEANS3D | ACC_SH | ACC_OPACITY | ACC_SCALES | ACC_ROTATIONS;
    if ((a.accumulate & ACC_ALL) == ACC_ALL && a.counters != nullptr) {
        cudaMemsetAsync(a.dL_dmeans2D, 0, sizeof(float) * 3 * (size_t)a.P, s);
        if (a.dL_dcolors) cudaMemsetAsync(a.dL_dcolors, 0, sizeof(float) * 3 * (size_t)a.P, s);
        if (a.dL_dtransMat) cudaMemsetAsync(a.dL_dtransMat, 0, sizeof(float) * 9 * (size_t)a.P, s);
        // the number of visible Gaussians is known on the device only: one CTA per 128 Gaussians, the CTAs past the
        // end of the list read one counter and leave
        const int grid = (a.P + PB_THREADS - 1) / PB_THREADS;
        if (a.raw) project_bwd_compact_kernel<true><<<grid, PB_THREADS, 0, s>>>(a);
        else project_bwd_compact_kernel<false><<<grid, PB_THREADS, 0, s>>>(a);
        count_launch();
        return;
    }
    if (a.raw) project_bwd_kernel<true><<<(a.P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, 0, s>>>(a);
    else project_bwd_kernel<false><<<(a.P + PB_THREADS - 1) / PB_THREADS, PB_THREADS, 0, s>>>(a);
    count_launch();
}
void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s) {
    if (P <= 0) return;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
    count_launch();
}

}  // namespace g4s
