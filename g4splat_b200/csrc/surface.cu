// surface.cu -- SURVEY.md 8(f) row 1: what render() does to the rasterizer's 7-channel `allmap` after
// every call (2d-gaussian-splatting/gaussian_renderer/__init__.py:118-164 and depth_to_normal /
// depths_to_points, utils/point_utils.py:9-37), forward and backward, one kernel each.
//
// The reference spends ~15 torch kernels, two matrix inversions (with their host synchronisation) and
// an [N,3] ray grid per call on it; autograd then walks the same graph backwards.  Here the camera
// constants are derived in fp64 by one thread per CTA and every pixel is visited once per direction.
//
//   rend_alpha      = allmap[1]                      rend_dist = allmap[6]
//   rend_normal     = allmap[2:5] rotated to world   rend_normal_cam = allmap[2:5]
//   rend_depth      = nan_to_num(allmap[0] / allmap[1], 0, 0)
//   surf_depth      = rend_depth * (1 - ratio) + ratio * nan_to_num(allmap[5], 0, 0)
//   surf_normal     = normalize(cross(P[y+1,x] - P[y-1,x], P[y,x+1] - P[y,x-1])) * alpha.detach(),
//                     P = surf_depth * ray + origin, zero on the one-pixel border
//   surf_normal_cam = surf_normal rotated back to the camera
//
// Finite differences of P are formed without the cancellation of the reference's fp32 point grid:
// with ray(x, y) = M0 x + M1 y + M2,
//   P[y+1,x] - P[y-1,x] = (d_dn - d_up) ray(x, y) + (d_dn + d_up) M1      (and likewise along x with M0),
// which is the same function of the depths (and has the same derivative), evaluated more accurately.
#include "kernels.cuh"

namespace g4s {

constexpr int SF_TW = 32, SF_TH = 8;             // pixels per CTA tile: one 128-byte row segment per warp
constexpr int SF_THREADS = SF_TW * SF_TH;
constexpr float NORMALIZE_EPS = 1e-12f;          // torch.nn.functional.normalize default

struct SurfCam {
    float M[3][3];   // ray(x, y)_i = M[i][0] x + M[i][1] y + M[i][2]   (= c2w[:3,:3] intrins^-1)
    float Rv[3][3];  // world_view_transform[:3,:3]
};

// Gauss-Jordan with partial pivoting on an n x n system (n <= 4), fp64, one thread.
__device__ void invert_small(const double* a_in, int n, double* inv) {
    double a[4][8];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) { a[i][j] = a_in[i * n + j]; a[i][n + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 2 * n; j++) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        const double d = 1.0 / a[c][c];
        for (int j = 0; j < 2 * n; j++) a[c][j] *= d;
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const double f = a[r][c];
            for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = a[i][n + j];
}

// utils/point_utils.py:10-21 as matrix algebra, V = world_view_transform, FP = full_proj_transform
// (both row-major [4][4], row-vector convention):
//   c2w = (V^T)^-1;  PM = c2w^T FP;  intrins = ((PM ndc2pix)[:3,:3])^T;  rays = pts intrins^-T c2w[:3,:3]^T
__device__ void surface_camera(const float* V, const float* FP, int W, int H, SurfCam* out) {
    double A[16], c2w[16], PM[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) A[i * 4 + j] = (double)V[j * 4 + i];
    invert_small(A, 4, c2w);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += c2w[k * 4 + i] * (double)FP[k * 4 + j];
            PM[i * 4 + j] = s;
        }
    // ndc2pix (4x3): rows (W/2, 0, 0), (0, H/2, 0), (0, 0, 0), (W/2, H/2, 1)
    const double hw = 0.5 * W, hh = 0.5 * H;
    double I3[9], Kinv[9];
    for (int i = 0; i < 3; i++) {   // Q[i][:] = PM[i][:] @ ndc2pix;  intrins[j][i] = Q[i][j]
        const double q0 = PM[i * 4 + 0] * hw + PM[i * 4 + 3] * hw;
        const double q1 = PM[i * 4 + 1] * hh + PM[i * 4 + 3] * hh;
        const double q2 = PM[i * 4 + 3];
        I3[0 * 3 + i] = q0; I3[1 * 3 + i] = q1; I3[2 * 3 + i] = q2;
    }
    invert_small(I3, 3, Kinv);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += c2w[i * 4 + k] * Kinv[k * 3 + j];
            out->M[i][j] = (float)s;
            out->Rv[i][j] = V[i * 4 + j];
        }
}

// torch.nan_to_num(x, 0, 0): nan -> 0, +inf -> 0, -inf -> lowest finite
__device__ __forceinline__ float nan_to_num00(float x) {
    if (isnan(x)) return 0.0f;
    if (isinf(x)) return x > 0 ? 0.0f : -3.4028234663852886e38f;
    return x;
}
__device__ __forceinline__ bool finite_f(float x) { return !(isnan(x) || isinf(x)); }

// surf_depth of one pixel (:128-141); *expected receives rend_depth
__device__ __forceinline__ float surf_depth_at(const float* __restrict__ allmap, size_t N, size_t pix, float r0, float r1,
                                               float* expected = nullptr) {
    const float D = allmap[pix], a = allmap[N + pix], med = allmap[5 * N + pix];
    const float e = nan_to_num00(__fdiv_rn(D, a));
    if (expected) *expected = e;
    return __fadd_rn(__fmul_rn(e, r0), __fmul_rn(r1, nan_to_num00(med)));
}

struct f3s { float x, y, z; };
__device__ __forceinline__ f3s ray_at(const SurfCam& c, float x, float y) {
    f3s r;
    r.x = fmaf(c.M[0][0], x, fmaf(c.M[0][1], y, c.M[0][2]));
    r.y = fmaf(c.M[1][0], x, fmaf(c.M[1][1], y, c.M[1][2]));
    r.z = fmaf(c.M[2][0], x, fmaf(c.M[2][1], y, c.M[2][2]));
    return r;
}
__device__ __forceinline__ f3s cross_s(f3s a, f3s b) {
    f3s r;
    r.x = a.y * b.z - a.z * b.y;
    r.y = a.z * b.x - a.x * b.z;
    r.z = a.x * b.y - a.y * b.x;
    return r;
}
// the two finite differences at centre (x, y) from its four neighbours' depths
__device__ __forceinline__ void finite_differences(const SurfCam& c, float x, float y, float d_up, float d_dn, float d_l,
                                                   float d_r, f3s& dxv, f3s& dyv) {
    const f3s rc = ray_at(c, x, y);
    const float dv = d_dn - d_up, sv = d_dn + d_up, dh = d_r - d_l, sh = d_r + d_l;
    dxv.x = fmaf(dv, rc.x, sv * c.M[0][1]); dxv.y = fmaf(dv, rc.y, sv * c.M[1][1]); dxv.z = fmaf(dv, rc.z, sv * c.M[2][1]);
    dyv.x = fmaf(dh, rc.x, sh * c.M[0][0]); dyv.y = fmaf(dh, rc.y, sh * c.M[1][0]); dyv.z = fmaf(dh, rc.z, sh * c.M[2][0]);
}

// ---------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(SF_THREADS, 4) surface_fwd_kernel(SurfaceFwdArgs a) {
    __shared__ SurfCam cam;
    if (threadIdx.x == 0) surface_camera(a.view, a.proj, a.W, a.H, &cam);
    __syncthreads();
    const int W = a.W, H = a.H;
    const size_t N = (size_t)W * H;
    const int tiles_x = (W + SF_TW - 1) / SF_TW, tiles_y = (H + SF_TH - 1) / SF_TH;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x = tx * SF_TW + (threadIdx.x & (SF_TW - 1)), y = ty * SF_TH + (threadIdx.x / SF_TW);
        if (x >= W || y >= H) continue;
        const size_t pix = (size_t)W * y + x;
        const float alpha = a.allmap[N + pix];
        const float n0 = a.allmap[2 * N + pix], n1 = a.allmap[3 * N + pix], n2 = a.allmap[4 * N + pix];
        float expected;
        const float sd = surf_depth_at(a.allmap, N, pix, a.r0, a.r1, &expected);
        a.rend_alpha[pix] = alpha;
        a.rend_dist[pix] = a.allmap[6 * N + pix];
        a.rend_depth[pix] = expected;
        a.surf_depth[pix] = sd;
        // rend_normal = n_cam @ V[:3,:3].T (:123)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            a.rend_normal[j * N + pix] = fmaf(n2, cam.Rv[j][2], fmaf(n1, cam.Rv[j][1], n0 * cam.Rv[j][0]));
        }
        a.rend_normal_cam[pix] = n0; a.rend_normal_cam[N + pix] = n1; a.rend_normal_cam[2 * N + pix] = n2;
        f3s sn = {0.f, 0.f, 0.f};
        if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) {
            const float d_up = surf_depth_at(a.allmap, N, pix - W, a.r0, a.r1);
            const float d_dn = surf_depth_at(a.allmap, N, pix + W, a.r0, a.r1);
            const float d_l = surf_depth_at(a.allmap, N, pix - 1, a.r0, a.r1);
            const float d_r = surf_depth_at(a.allmap, N, pix + 1, a.r0, a.r1);
            f3s dxv, dyv;
            finite_differences(cam, (float)x, (float)y, d_up, d_dn, d_l, d_r, dxv, dyv);
            const f3s c = cross_s(dxv, dyv);
            const float nrm = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
            const float s = alpha / fmaxf(nrm, NORMALIZE_EPS);   // (:148) times the detached alpha
            sn.x = c.x * s; sn.y = c.y * s; sn.z = c.z * s;
        }
        a.surf_normal[pix] = sn.x; a.surf_normal[N + pix] = sn.y; a.surf_normal[2 * N + pix] = sn.z;
        // surf_normal_cam = surf_normal @ V[:3,:3] (:152)
#pragma unroll
        for (int i = 0; i < 3; i++)
            a.surf_normal_cam[i * N + pix] = fmaf(sn.z, cam.Rv[2][i], fmaf(sn.y, cam.Rv[1][i], sn.x * cam.Rv[0][i]));
    }
}

// --------------------------------------------------------------------------------------- backward
// Per tile: (1) surf_depth of the tile + a two-pixel halo into shared memory, (2) for every centre in
// the tile + one-pixel halo the gradients of its two finite differences, (3) every pixel gathers the
// four centres it is a neighbour of and writes its seven allmap gradients once.
constexpr int SB_DW = SF_TW + 4, SB_DH = SF_TH + 4;   // depth tile
constexpr int SB_GW = SF_TW + 2, SB_GH = SF_TH + 2;   // centre tile

__device__ __forceinline__ float ld0(const float* p, size_t i) { return p ? p[i] : 0.0f; }

__global__ void __launch_bounds__(SF_THREADS, 4) surface_bwd_kernel(SurfaceBwdArgs a) {
    __shared__ SurfCam cam;
    __shared__ float s_depth[SB_DH][SB_DW];
    __shared__ float s_G[6][SB_GH * SB_GW];
    if (threadIdx.x == 0) surface_camera(a.view, a.proj, a.W, a.H, &cam);
    __syncthreads();
    const int W = a.W, H = a.H;
    const size_t N = (size_t)W * H;
    const int tiles_x = (W + SF_TW - 1) / SF_TW, tiles_y = (H + SF_TH - 1) / SF_TH;
    const bool normals_live = a.g_surf_normal != nullptr || a.g_surf_normal_cam != nullptr;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x0 = tx * SF_TW, y0 = ty * SF_TH;
        if (normals_live) {
            for (int i = threadIdx.x; i < SB_DH * SB_DW; i += SF_THREADS) {
                const int ly = i / SB_DW, lx = i - ly * SB_DW;
                const int gy = y0 - 2 + ly, gx = x0 - 2 + lx;
                float d = 0.0f;
                if (gx >= 0 && gx < W && gy >= 0 && gy < H) d = surf_depth_at(a.allmap, N, (size_t)W * gy + gx, a.r0, a.r1);
                s_depth[ly][lx] = d;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < SB_GH * SB_GW; i += SF_THREADS) {
                const int ly = i / SB_GW, lx = i - ly * SB_GW;
                const int cy = y0 - 1 + ly, cx = x0 - 1 + lx;
                f3s gdx = {0.f, 0.f, 0.f}, gdy = {0.f, 0.f, 0.f};
                if (cx >= 1 && cx <= W - 2 && cy >= 1 && cy <= H - 2) {
                    const size_t cp = (size_t)W * cy + cx;
                    // upstream gradient of the unit normal: alpha (g_surf_normal + Rv g_surf_normal_cam)
                    const float alpha = a.allmap[N + cp];
                    const float c0 = ld0(a.g_surf_normal_cam, cp), c1 = ld0(a.g_surf_normal_cam, N + cp),
                                c2 = ld0(a.g_surf_normal_cam, 2 * N + cp);
                    f3s gn;
                    gn.x = alpha * (ld0(a.g_surf_normal, cp) + fmaf(cam.Rv[0][2], c2, fmaf(cam.Rv[0][1], c1, cam.Rv[0][0] * c0)));
                    gn.y = alpha * (ld0(a.g_surf_normal, N + cp) + fmaf(cam.Rv[1][2], c2, fmaf(cam.Rv[1][1], c1, cam.Rv[1][0] * c0)));
                    gn.z = alpha * (ld0(a.g_surf_normal, 2 * N + cp) + fmaf(cam.Rv[2][2], c2, fmaf(cam.Rv[2][1], c1, cam.Rv[2][0] * c0)));
                    f3s dxv, dyv;
                    finite_differences(cam, (float)cx, (float)cy, s_depth[ly][lx + 1], s_depth[ly + 2][lx + 1],
                                       s_depth[ly + 1][lx], s_depth[ly + 1][lx + 2], dxv, dyv);
                    const f3s c = cross_s(dxv, dyv);
                    const float nrm = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
                    f3s gc;
                    if (nrm >= NORMALIZE_EPS) {   // v / |v|: (g - n (n.g)) / |v|
                        const float inv = 1.0f / nrm;
                        const f3s n = {c.x * inv, c.y * inv, c.z * inv};
                        const float ng = n.x * gn.x + n.y * gn.y + n.z * gn.z;
                        gc.x = (gn.x - n.x * ng) * inv; gc.y = (gn.y - n.y * ng) * inv; gc.z = (gn.z - n.z * ng) * inv;
                    } else {                      // clamp_min(eps) is active: v / eps
                        gc.x = gn.x / NORMALIZE_EPS; gc.y = gn.y / NORMALIZE_EPS; gc.z = gn.z / NORMALIZE_EPS;
                    }
                    gdx = cross_s(dyv, gc);       // c = dx x dy
                    gdy = cross_s(gc, dxv);
                }
                s_G[0][i] = gdx.x; s_G[1][i] = gdx.y; s_G[2][i] = gdx.z;
                s_G[3][i] = gdy.x; s_G[4][i] = gdy.y; s_G[5][i] = gdy.z;
            }
            __syncthreads();
        }
        const int lx = threadIdx.x & (SF_TW - 1), ly = threadIdx.x / SF_TW;
        const int x = x0 + lx, y = y0 + ly;
        if (x < W && y < H) {
            const size_t pix = (size_t)W * y + x;
            float g_sd = ld0(a.g_surf_depth, pix);
            if (normals_live) {
                const int up = ly * SB_GW + (lx + 1), dn = (ly + 2) * SB_GW + (lx + 1);
                const int lf = (ly + 1) * SB_GW + lx, rt = (ly + 1) * SB_GW + (lx + 2);
                const float px_ = s_G[0][up] - s_G[0][dn] + s_G[3][lf] - s_G[3][rt];
                const float py_ = s_G[1][up] - s_G[1][dn] + s_G[4][lf] - s_G[4][rt];
                const float pz_ = s_G[2][up] - s_G[2][dn] + s_G[5][lf] - s_G[5][rt];
                const f3s r = ray_at(cam, (float)x, (float)y);
                g_sd += px_ * r.x + py_ * r.y + pz_ * r.z;
            }
            const float D = a.allmap[pix], alpha = a.allmap[N + pix], med = a.allmap[5 * N + pix];
            // rend_depth = nan_to_num(D / alpha): the gradient passes where the quotient is finite.  Where
            // alpha == 0 torch's division backward turns the masked 0 into 0/0 = NaN; such a pixel has no
            // contributor and the rasterizer ignores its gradients, so zeros are written instead.
            const float quot = __fdiv_rn(D, alpha);
            const float g_e = finite_f(quot) ? ld0(a.g_rend_depth, pix) + a.r0 * g_sd : 0.0f;
            const float inv_a = alpha != 0.0f ? 1.0f / alpha : 0.0f;
            const float g_D = g_e * inv_a;
            const float g_alpha = ld0(a.g_rend_alpha, pix) - (g_e != 0.0f ? g_e * quot * inv_a : 0.0f);
            const float g_med = finite_f(med) ? a.r1 * g_sd : 0.0f;
            const float w0 = ld0(a.g_rend_normal, pix), w1 = ld0(a.g_rend_normal, N + pix), w2 = ld0(a.g_rend_normal, 2 * N + pix);
            a.g_allmap[pix] = g_D;
            a.g_allmap[N + pix] = g_alpha;
#pragma unroll
            for (int i = 0; i < 3; i++)   // rend_normal_j = sum_i n_i V[j][i]
                a.g_allmap[(2 + i) * N + pix] = ld0(a.g_rend_normal_cam, i * N + pix) +
                                                fmaf(w2, cam.Rv[2][i], fmaf(w1, cam.Rv[1][i], w0 * cam.Rv[0][i]));
            a.g_allmap[5 * N + pix] = g_med;
            a.g_allmap[6 * N + pix] = ld0(a.g_rend_dist, pix);
        }
        __syncthreads();   // shared tiles are rewritten by the next iteration
    }
}

static int surface_grid(int W, int H) {
    const int tiles = ((W + SF_TW - 1) / SF_TW) * ((H + SF_TH - 1) / SF_TH);
    return tiles < 148 * 8 ? tiles : 148 * 8;   // persistent CTAs: the fp64 camera prologue runs once per CTA
}

void launch_surface_fwd(const SurfaceFwdArgs& a, cudaStream_t s) {
    count_launch();
    surface_fwd_kernel<<<surface_grid(a.W, a.H), SF_THREADS, 0, s>>>(a);
}
void launch_surface_bwd(const SurfaceBwdArgs& a, cudaStream_t s) {
    count_launch();
    surface_bwd_kernel<<<surface_grid(a.W, a.H), SF_THREADS, 0, s>>>(a);
}

}  // namespace g4s
