// gaussian_model.cu -- SURVEY.md 8(f) row 3: per-Gaussian bookkeeping the trainer runs around the
// rasterizer on P-sized tensors.
//
// compute_mip_filter (2d-gaussian-splatting/scene/gaussian_model.py:388-434).  The reference loops
// over ALL training cameras with ~14 torch kernels per camera on [P]-sized tensors (a [P,3]x[3,3]
// matmul, norm, clamp, two divisions, four comparisons, three logical ops, two masked scatters...),
// i.e. ~14 C launches and ~100 C P bytes of traffic; here every Gaussian is read once, the C camera
// records (20 floats each) are staged in shared memory and walked in registers, and a second tiny
// pass fills the Gaussians no camera sees with the largest distance found:
//   pass 1: dist[i] = min over cameras c with (z_c > znear and the projection lies inside the image
//           enlarged by 15 %) of max(z_c, 0.001), started at 100000;  seen[i] = any such camera;
//           max_seen = max over seen Gaussians of dist  (fp32 bits, atomicMax: all values > 0)
//   pass 2: filter[i] = (seen[i] ? dist[i] : max_seen) / focal_length * sqrt(filter_variance)
#include "kernels.cuh"

namespace g4s {

constexpr int MIP_THREADS = 256;
constexpr int MIP_CAM_FLOATS = 20;
constexpr int MIP_CAM_CHUNK = 256;   // camera records staged per round: 20 KB of shared memory

// camera record (MIP_CAM_FLOATS floats, built by the host in double and rounded once, as torch
// rounds the reference's Python scalars): R[0..8] row-major as the reference stores camera.R (points
// are row vectors: xyz_cam = xyz @ R + T), T[9..11], focal_x, focal_y, W/2, H/2,
// -0.15 W, 1.15 W, -0.15 H, 1.15 H
__global__ void __launch_bounds__(MIP_THREADS) mip_distance_kernel(int P, const float* __restrict__ xyz, int C,
                                                                   const float* __restrict__ cams, float znear,
                                                                   float* __restrict__ dist, unsigned int* __restrict__ max_bits) {
    __shared__ float s_cam[MIP_CAM_CHUNK * MIP_CAM_FLOATS];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (idx < P) { x = xyz[3 * idx]; y = xyz[3 * idx + 1]; z = xyz[3 * idx + 2]; }
    float best = 100000.0f;
    bool seen = false;
    for (int c0 = 0; c0 < C; c0 += MIP_CAM_CHUNK) {
        const int nc = min(MIP_CAM_CHUNK, C - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * MIP_CAM_FLOATS; i += MIP_THREADS) s_cam[i] = cams[(size_t)c0 * MIP_CAM_FLOATS + i];
        __syncthreads();
        for (int c = 0; c < nc; c++) {
            const float* k = &s_cam[c * MIP_CAM_FLOATS];
            // xyz @ R + T, summed left to right like a row of the matrix product
            const float xc = __fadd_rn(__fmaf_rn(z, k[6], __fmaf_rn(y, k[3], __fmul_rn(x, k[0]))), k[9]);
            const float yc = __fadd_rn(__fmaf_rn(z, k[7], __fmaf_rn(y, k[4], __fmul_rn(x, k[1]))), k[10]);
            const float zc = __fadd_rn(__fmaf_rn(z, k[8], __fmaf_rn(y, k[5], __fmul_rn(x, k[2]))), k[11]);
            const float zz = fmaxf(zc, 0.001f);
            const float u = __fadd_rn(__fmul_rn(__fdiv_rn(xc, zz), k[12]), k[14]);
            const float v = __fadd_rn(__fmul_rn(__fdiv_rn(yc, zz), k[13]), k[15]);
            const bool ok = zc > znear && u >= k[16] && u <= k[17] && v >= k[18] && v <= k[19];
            if (ok) { best = fminf(best, zz); seen = true; }
        }
    }
    // encode "no camera sees it" as a negative distance for pass 2
    if (idx < P) dist[idx] = seen ? best : -1.0f;
    float m = seen && idx < P ? best : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_bits, __float_as_uint(m));
}

__global__ void __launch_bounds__(MIP_THREADS) mip_finish_kernel(int P, float* __restrict__ filter,
                                                                 const unsigned int* __restrict__ max_bits,
                                                                 float focal_length, float sqrt_variance) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float d = filter[idx];
    const float dist = d < 0.0f ? __uint_as_float(*max_bits) : d;
    filter[idx] = __fmul_rn(__fdiv_rn(dist, focal_length), sqrt_variance);
}

void launch_mip_filter(int P, const float* xyz, int C, const float* cams, float znear, float focal_length,
                       float sqrt_variance, float* filter, unsigned int* max_bits, cudaStream_t s) {
    if (P <= 0) return;
    const int blocks = (P + MIP_THREADS - 1) / MIP_THREADS;
    mip_distance_kernel<<<blocks, MIP_THREADS, 0, s>>>(P, xyz, C, cams, znear, filter, max_bits);
    count_launch();
    mip_finish_kernel<<<blocks, MIP_THREADS, 0, s>>>(P, filter, max_bits, focal_length, sqrt_variance);
    count_launch();
}

}  // namespace g4s
