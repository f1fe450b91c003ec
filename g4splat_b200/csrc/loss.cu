// loss.cu -- SURVEY.md 8(f) row 4: the photometric loss the trainer forms on the rasterizer's image,
//   loss = (1 - lambda) * l1_loss(image, gt) + lambda * (1 - ssim(image, gt))
// (2d-gaussian-splatting/train_with_refine_depth.py:382-383; utils/loss_utils.py:17-18 l1_loss,
// :29-31 gaussian, :44-48 create_window, :49-80 ssim / _ssim), forward and backward, fused.
//
// The reference spends five depthwise 11x11 conv2d launches, ~15 elementwise kernels and their
// autograd on it per iteration.  Here: one forward kernel (separable 11-tap Gaussian over a
// shared-memory tile, the five moments of a pixel stay in registers, |x - y| is summed on the way)
// and one backward kernel.  Per pixel (zero padding, window w = g g^T):
//   m1 = w*x, m2 = w*y, e11 = w*x^2, e22 = w*y^2, e12 = w*xy
//   a = 2 m1 m2 + C1, b = 2 (e12 - m1 m2) + C2, c = m1^2 + m2^2 + C1, d = (e11 - m1^2) + (e22 - m2^2) + C2
//   ssim = a b / (c d)
// The forward also stores the three partial derivatives the backward needs (with respect to the raw
// window sums that involve x):
//   dA = d ssim / d m1 = 2 m2 (b - a) / (c d) - 2 m1 ssim (1/c - 1/d)
//   dB = d ssim / d e11 = -ssim / d          dC = d ssim / d e12 = 2 a / (c d)
// so that   d mean(ssim) / d x(q) = (1/N) [ (w*dA)(q) + 2 x(q) (w*dB)(q) + y(q) (w*dC)(q) ].
#include "kernels.cuh"

namespace g4s {

constexpr int LS_TILE = 32;                 // output pixels per CTA: 32 x 32 of one channel
constexpr int LS_HALO = 5;                  // window_size // 2
constexpr int LS_IN = LS_TILE + 2 * LS_HALO;  // 42
constexpr int LS_THREADS = 256;
constexpr float SSIM_C1 = 0.01f * 0.01f;
constexpr float SSIM_C2 = 0.03f * 0.03f;

struct Window11 { float g[11]; };

// stage a LS_IN x LS_IN window of one channel plane (zero outside the image)
__device__ __forceinline__ void stage_plane(float* __restrict__ dst, const float* __restrict__ plane, int W, int H,
                                            int x0, int y0) {
    for (int i = threadIdx.x; i < LS_IN * LS_IN; i += LS_THREADS) {
        const int r = i / LS_IN, c = i - r * LS_IN;
        const int x = x0 + c - LS_HALO, y = y0 + r - LS_HALO;
        dst[i] = (x >= 0 && x < W && y >= 0 && y < H) ? plane[(size_t)y * W + x] : 0.0f;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid: (ceil(W/32), ceil(H/32), C).  sums[0] += sum |x - y|, sums[1] += sum ssim.
__global__ void __launch_bounds__(LS_THREADS) photometric_fwd_kernel(int W, int H, const float* __restrict__ img,
                                                                     const float* __restrict__ gt, Window11 win,
                                                                     double* __restrict__ sums,
                                                                     float* __restrict__ dmaps /* [3][C][H][W] or null */) {
    __shared__ float s_x[LS_IN * LS_IN], s_y[LS_IN * LS_IN];
    __shared__ float s_h[5][LS_IN][LS_TILE + 1];   // horizontally filtered moments
    __shared__ double s_part[2][LS_THREADS / 32];
    const int x0 = blockIdx.x * LS_TILE, y0 = blockIdx.y * LS_TILE, ch = blockIdx.z;
    const size_t plane = (size_t)W * H;
    const float* px = img + ch * plane;
    const float* py = gt + ch * plane;
    stage_plane(s_x, px, W, H, x0, y0);
    stage_plane(s_y, py, W, H, x0, y0);
    __syncthreads();

    // horizontal pass: a work item is (input row r, 8 consecutive output columns)
    for (int item = threadIdx.x; item < LS_IN * (LS_TILE / 8); item += LS_THREADS) {
        const int r = item / (LS_TILE / 8), c0 = (item - r * (LS_TILE / 8)) * 8;
        float vx[18], vy[18];
#pragma unroll
        for (int k = 0; k < 18; k++) { vx[k] = s_x[r * LS_IN + c0 + k]; vy[k] = s_y[r * LS_IN + c0 + k]; }
#pragma unroll
        for (int o = 0; o < 8; o++) {
            float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) {
                const float g = win.g[k], a = vx[o + k], b = vy[o + k];
                const float ga = g * a, gb = g * b;
                m1 += ga; m2 += gb; e11 += ga * a; e22 += gb * b; e12 += ga * b;
            }
            s_h[0][r][c0 + o] = m1; s_h[1][r][c0 + o] = m2; s_h[2][r][c0 + o] = e11; s_h[3][r][c0 + o] = e22; s_h[4][r][c0 + o] = e12;
        }
    }
    __syncthreads();

    // vertical pass: a thread owns (column c, 4 consecutive output rows)
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float acc[5][4];
#pragma unroll
    for (int q = 0; q < 5; q++)
#pragma unroll
        for (int o = 0; o < 4; o++) acc[q][o] = 0.f;
#pragma unroll
    for (int k = 0; k < 14; k++) {
        float v[5];
#pragma unroll
        for (int q = 0; q < 5; q++) v[q] = s_h[q][r0 + k][c];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int tap = k - o;
            if (tap >= 0 && tap < 11) {
#pragma unroll
                for (int q = 0; q < 5; q++) acc[q][o] += win.g[tap] * v[q];
            }
        }
    }
    double l1 = 0.0, ss = 0.0;
    const int x = x0 + c;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const int y = y0 + r0 + o;
        if (x < W && y < H) {
            const float m1 = acc[0][o], m2 = acc[1][o], e11 = acc[2][o], e22 = acc[3][o], e12 = acc[4][o];
            const float m1m2 = m1 * m2, m1s = m1 * m1, m2s = m2 * m2;
            const float a = 2.f * m1m2 + SSIM_C1;
            const float b = 2.f * (e12 - m1m2) + SSIM_C2;
            const float cc = m1s + m2s + SSIM_C1;
            const float d = (e11 - m1s) + (e22 - m2s) + SSIM_C2;
            const float icd = 1.0f / (cc * d);
            const float ssim = a * b * icd;
            ss += (double)ssim;
            const float xv = s_x[(r0 + o + LS_HALO) * LS_IN + c + LS_HALO], yv = s_y[(r0 + o + LS_HALO) * LS_IN + c + LS_HALO];
            l1 += (double)fabsf(xv - yv);
            if (dmaps) {
                const size_t at = ch * plane + (size_t)y * W + x;
                const size_t stride = (size_t)gridDim.z * plane;
                dmaps[at] = 2.f * m2 * (b - a) * icd - 2.f * m1 * ssim * (1.0f / cc - 1.0f / d);
                dmaps[at + stride] = -ssim / d;
                dmaps[at + 2 * stride] = 2.f * a * icd;
            }
        }
    }
    l1 = warp_sum(l1); ss = warp_sum(ss);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_part[0][warp] = l1; s_part[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < LS_THREADS / 32; w++) t += s_part[threadIdx.x][w];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

// out[0] = loss = (1 - lambda) * L1 + lambda * (1 - SSIM), out[1] = L1 = mean |x - y|, out[2] = SSIM = mean ssim
__global__ void photometric_finish_kernel(const double* __restrict__ sums, double count, float lambda, float* __restrict__ out) {
    const float l1 = (float)(sums[0] / count), ss = (float)(sums[1] / count);
    out[0] = (1.0f - lambda) * l1 + lambda * (1.0f - ss);
    out[1] = l1;
    out[2] = ss;
}

// dL/dimg = upstream * [ w_l1 * sign(x - y) + w_ssim * ((w*dA) + 2 x (w*dB) + y (w*dC)) ]
// with w_l1 = (1 - lambda) / count, w_ssim = -lambda / count (the loss holds 1 - ssim), upstream = *dL_dloss.
__global__ void __launch_bounds__(LS_THREADS) photometric_bwd_kernel(int W, int H, const float* __restrict__ img,
                                                                     const float* __restrict__ gt, Window11 win,
                                                                     const float* __restrict__ dmaps, float w_l1, float w_ssim,
                                                                     const float* __restrict__ dL_dloss, float* __restrict__ dL_dimg) {
    __shared__ float s_in[3][LS_IN * LS_IN];
    __shared__ float s_h[3][LS_IN][LS_TILE + 1];
    const int x0 = blockIdx.x * LS_TILE, y0 = blockIdx.y * LS_TILE, ch = blockIdx.z;
    const size_t plane = (size_t)W * H, stride = (size_t)gridDim.z * plane;
#pragma unroll
    for (int q = 0; q < 3; q++) stage_plane(s_in[q], dmaps + q * stride + ch * plane, W, H, x0, y0);
    __syncthreads();
    for (int item = threadIdx.x; item < LS_IN * (LS_TILE / 8); item += LS_THREADS) {
        const int r = item / (LS_TILE / 8), c0 = (item - r * (LS_TILE / 8)) * 8;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            float v[18];
#pragma unroll
            for (int k = 0; k < 18; k++) v[k] = s_in[q][r * LS_IN + c0 + k];
#pragma unroll
            for (int o = 0; o < 8; o++) {
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 11; k++) s += win.g[k] * v[o + k];
                s_h[q][r][c0 + o] = s;
            }
        }
    }
    __syncthreads();
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float acc[3][4];
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
        for (int o = 0; o < 4; o++) acc[q][o] = 0.f;
#pragma unroll
    for (int k = 0; k < 14; k++) {
        float v[3];
#pragma unroll
        for (int q = 0; q < 3; q++) v[q] = s_h[q][r0 + k][c];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int tap = k - o;
            if (tap >= 0 && tap < 11) {
#pragma unroll
                for (int q = 0; q < 3; q++) acc[q][o] += win.g[tap] * v[q];
            }
        }
    }
    const float up = dL_dloss ? *dL_dloss : 1.0f;
    const int x = x0 + c;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const int y = y0 + r0 + o;
        if (x < W && y < H) {
            const size_t at = ch * plane + (size_t)y * W + x;
            const float xv = img[at], yv = gt[at];
            const float diff = xv - yv;
            const float sgn = diff > 0.f ? 1.0f : (diff < 0.f ? -1.0f : 0.0f);   // torch.abs backward: sign(), 0 at 0
            dL_dimg[at] = up * (w_l1 * sgn + w_ssim * (acc[0][o] + 2.f * xv * acc[1][o] + yv * acc[2][o]));
        }
    }
}

void launch_photometric_fwd(int W, int H, int C, const float* img, const float* gt, const float* window11, float lambda,
                            double* sums, float* dmaps, float* out3, cudaStream_t s) {
    Window11 win;
    for (int i = 0; i < 11; i++) win.g[i] = window11[i];
    dim3 grid((W + LS_TILE - 1) / LS_TILE, (H + LS_TILE - 1) / LS_TILE, C);
    photometric_fwd_kernel<<<grid, LS_THREADS, 0, s>>>(W, H, img, gt, win, sums, dmaps);
    count_launch();
    photometric_finish_kernel<<<1, 1, 0, s>>>(sums, (double)W * H * C, lambda, out3);
    count_launch();
}

void launch_photometric_bwd(int W, int H, int C, const float* img, const float* gt, const float* window11, float lambda,
                            const float* dmaps, const float* dL_dloss, float* dL_dimg, cudaStream_t s) {
    Window11 win;
    for (int i = 0; i < 11; i++) win.g[i] = window11[i];
    const double count = (double)W * H * C;
    dim3 grid((W + LS_TILE - 1) / LS_TILE, (H + LS_TILE - 1) / LS_TILE, C);
    photometric_bwd_kernel<<<grid, LS_THREADS, 0, s>>>(W, H, img, gt, win, dmaps, (float)((1.0 - lambda) / count),
                                                       (float)(-(double)lambda / count), dL_dloss, dL_dimg);
    count_launch();
}

}  // namespace g4s
