// regularizers.cu -- the two image-space regularisers the trainer evaluates on the rasterizer's outputs every
// iteration (the part the photometric kernel of loss.cu does not cover):
//
//   normal2curv            matcha/dm_utils/rendering.py:392-406, called at 2DGS/train_with_refine_depth.py:415
//                          (8 slicing / padding / multiply kernels + permutes and their autograd in the reference)
//   compute_depth_order_loss  matcha/dm_regularization/depth.py:142-222, called at train_with_refine_depth.py:465
//                          (meshgrid, randint, two advanced-indexing gathers, ~10 elementwise kernels, a reduction)
//
// One kernel per direction each.  HBM-bound elementwise work: every input element is read once per use from a
// cache-resident neighbourhood, every output written once; nothing here is GEMM-shaped.
#include "kernels.cuh"

namespace g4s {

constexpr int RG_THREADS = 256;

// ---- normal2curv ---------------------------------------------------------------------------------------------
// curv(y, x) = sum_ch | m_c * sum_{nb in up, left, bottom, right} (n[nb, ch] - n[c, ch] m_c) m_nb |
// with replicate padding of both the normal map and the mask (a neighbour outside the image is the border pixel
// itself).  sg (optional, [3][N]) receives sign(.) * m_c per channel: the backward's only dependence on values.
__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

__global__ void __launch_bounds__(RG_THREADS) normal2curv_fwd_kernel(int W, int H, const float* __restrict__ normal,
                                                                     const float* __restrict__ mask, float* __restrict__ curv,
                                                                     float* __restrict__ sg) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t N = (size_t)W * H;
    const int yu = max(y - 1, 0), yb = min(y + 1, H - 1), xl = max(x - 1, 0), xr = min(x + 1, W - 1);
    const size_t c = (size_t)y * W + x, u = (size_t)yu * W + x, b = (size_t)yb * W + x, l = (size_t)y * W + xl, r = (size_t)y * W + xr;
    auto m_at = [&](size_t i) { return mask ? ((mask[i] != 0.f) ? 1.f : 0.f) : 1.f; };   // the reference converts the padded mask to bool
    const float mc = m_at(c), mu = m_at(u), mb = m_at(b), ml = m_at(l), mr = m_at(r);
    const float mc_out = mask ? mask[c] : 1.f;                                            // the final `* mask` uses the caller's values
    float total = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float* n = normal + ch * N;
        const float nc = n[c] * mc;
        // (n_u + n_l + n_b + n_r), the reference's order
        float s = (n[u] - nc) * mu;
        s += (n[l] - nc) * ml;
        s += (n[b] - nc) * mb;
        s += (n[r] - nc) * mr;
        const float z = s * mc_out;
        total += fabsf(z);
        if (sg) sg[ch * N + c] = sgnf(z) * mc_out;
    }
    curv[c] = total;
}

// dL/dn[q, ch] = sum over the pixels p that use q as a neighbour (clamping makes a border pixel its own neighbour)
//                of ds_p[ch] * m_q   -   ds_q[ch] * (sum_nb m_nb(q)) * m_q,       ds_p[ch] = g_p * sg_p[ch]
__global__ void __launch_bounds__(RG_THREADS) normal2curv_bwd_kernel(int W, int H, const float* __restrict__ mask,
                                                                     const float* __restrict__ sg, const float* __restrict__ g_curv,
                                                                     float* __restrict__ g_normal) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t N = (size_t)W * H;
    const int yu = max(y - 1, 0), yb = min(y + 1, H - 1), xl = max(x - 1, 0), xr = min(x + 1, W - 1);
    const size_t c = (size_t)y * W + x, u = (size_t)yu * W + x, b = (size_t)yb * W + x, l = (size_t)y * W + xl, r = (size_t)y * W + xr;
    auto m_at = [&](size_t i) { return mask ? ((mask[i] != 0.f) ? 1.f : 0.f) : 1.f; };
    const float mq = m_at(c);
    const float msum = m_at(u) + m_at(l) + m_at(b) + m_at(r);
    const float gc = g_curv[c];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float* s = sg + ch * N;
        const float dsq = gc * s[c];
        float acc = 0.f;
        // q is the UP neighbour of the pixel below it (and of itself on the top row), and so on
        if (y + 1 <= H - 1) acc += g_curv[b] * s[b];
        if (y == 0) acc += dsq;
        if (y - 1 >= 0) acc += g_curv[u] * s[u];      // q is the BOTTOM neighbour of the pixel above it
        if (y == H - 1) acc += dsq;
        if (x + 1 <= W - 1) acc += g_curv[r] * s[r];  // q is the LEFT neighbour of the pixel to its right
        if (x == 0) acc += dsq;
        if (x - 1 >= 0) acc += g_curv[l] * s[l];      // q is the RIGHT neighbour of the pixel to its left
        if (x == W - 1) acc += dsq;
        g_normal[ch * N + c] = (acc - dsq * msum) * mq;
    }
}

// ---- depth-order loss ------------------------------------------------------------------------------------------
// per pixel i with shifted partner j = clamp(i + shift_i):
//   diff = (d_i - d_j) * inv_extent ; prior = (p_i - p_j) * inv_extent ; normalize: prior /= max(|prior|, 1e-8)
//   l = -min(diff * prior, 0) ; log space: l = log(1 + log_scale * l)
__device__ __forceinline__ void order_pair(int i, int W, int H, const long long* __restrict__ shifts, const float* __restrict__ depth,
                                           const float* __restrict__ prior, float inv_extent, int normalize, int& j,
                                           float& diff, float& pr) {
    const int y = i / W, x = i - y * W;
    const int sy = min(max(y + (int)shifts[2 * (size_t)i], 0), H - 1);
    const int sx = min(max(x + (int)shifts[2 * (size_t)i + 1], 0), W - 1);
    j = sy * W + sx;
    diff = (depth[i] - depth[j]) * inv_extent;
    pr = (prior[i] - prior[j]) * inv_extent;
    if (normalize) pr = pr / fmaxf(fabsf(pr), 1e-8f);
}

__global__ void __launch_bounds__(RG_THREADS) depth_order_fwd_kernel(int W, int H, const float* __restrict__ depth,
                                                                     const float* __restrict__ prior, const long long* __restrict__ shifts,
                                                                     float inv_extent, int normalize, int log_space, float log_scale,
                                                                     float* __restrict__ per_pixel, double* __restrict__ sum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = W * H;
    float l = 0.f;
    if (i < N) {
        int j; float diff, pr;
        order_pair(i, W, H, shifts, depth, prior, inv_extent, normalize, j, diff, pr);
        l = -fminf(diff * pr, 0.f);
        if (log_space) l = logf(1.f + log_scale * l);
        if (per_pixel) per_pixel[i] = l;
    }
    if (sum) {
        __shared__ double s_part[RG_THREADS / 32];
        double v = (double)l;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < RG_THREADS / 32; k++) t += s_part[k];
            atomicAdd(sum, t);
        }
    }
}

// g: per-pixel upstream gradient (reduction "none") or null with `g_scalar` (device float[1], null = 1) * scale
__global__ void __launch_bounds__(RG_THREADS) depth_order_bwd_kernel(int W, int H, const float* __restrict__ depth,
                                                                     const float* __restrict__ prior, const long long* __restrict__ shifts,
                                                                     float inv_extent, int normalize, int log_space, float log_scale,
                                                                     const float* __restrict__ g, const float* __restrict__ g_scalar, float scale,
                                                                     float* __restrict__ g_depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * H) return;
    int j; float diff, pr;
    order_pair(i, W, H, shifts, depth, prior, inv_extent, normalize, j, diff, pr);
    const float prod = diff * pr;
    if (!(prod <= 0.f)) return;                 // clamp(max=0) passes its gradient where the product is <= 0
    float up = g ? g[i] : (g_scalar ? g_scalar[0] : 1.f) * scale;
    if (log_space) up *= log_scale / (1.f + log_scale * (-prod));
    const float gd = -up * pr * inv_extent;     // d l / d d_i ; the partner gets the opposite
    atomicAdd(&g_depth[i], gd);
    atomicAdd(&g_depth[j], -gd);
}

void launch_normal2curv_fwd(int W, int H, const float* normal, const float* mask, float* curv, float* sg, cudaStream_t s) {
    dim3 grid((W + RG_THREADS - 1) / RG_THREADS, H);
    normal2curv_fwd_kernel<<<grid, RG_THREADS, 0, s>>>(W, H, normal, mask, curv, sg);
    count_launch();
}
void launch_normal2curv_bwd(int W, int H, const float* mask, const float* sg, const float* g_curv, float* g_normal, cudaStream_t s) {
    dim3 grid((W + RG_THREADS - 1) / RG_THREADS, H);
    normal2curv_bwd_kernel<<<grid, RG_THREADS, 0, s>>>(W, H, mask, sg, g_curv, g_normal);
    count_launch();
}
void launch_depth_order_fwd(int W, int H, const float* depth, const float* prior, const long long* shifts, float inv_extent,
                            int normalize, int log_space, float log_scale, float* per_pixel, double* sum, cudaStream_t s) {
    depth_order_fwd_kernel<<<(W * H + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, s>>>(W, H, depth, prior, shifts, inv_extent, normalize,
                                                                                       log_space, log_scale, per_pixel, sum);
    count_launch();
}
void launch_depth_order_bwd(int W, int H, const float* depth, const float* prior, const long long* shifts, float inv_extent,
                            int normalize, int log_space, float log_scale, const float* g, const float* g_scalar, float scale,
                            float* g_depth, cudaStream_t s) {
    depth_order_bwd_kernel<<<(W * H + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, s>>>(W, H, depth, prior, shifts, inv_extent, normalize,
                                                                                       log_space, log_scale, g, g_scalar, scale, g_depth);
    count_launch();
}

}  // namespace g4s
