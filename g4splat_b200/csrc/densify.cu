// densify.cu -- adaptive density control of the trainer, fused.
//
// Reference: GaussianModel.densify_and_prune and what it calls (2DGS/scene/gaussian_model.py:528-647:
// densify_and_clone, densify_and_split, prune_points, cat_tensors_to_optimizer, _prune_optimizer), driven from
// train_with_refine_depth.py:582-599.  There it is ~60 torch kernels: boolean masks, six masked gathers per step,
// six torch.cat per step, the same again for both Adam moments, three rounds of it (clone, split, prune).  Here:
//   densify_classify   one pass over the per-Gaussian statistics: which Gaussians are cloned, split, pruned
//   densify_gather     one pass that writes the final parameter tensors AND both Adam moments of every survivor,
//                      clone and split child, in the reference's final row order
// The host turns the flags into index lists (the reference synchronises for the same sizes) and draws the split
// samples with the reference's own torch.normal call, so a seeded run creates the same children.
// HBM-bound row copies; nothing GEMM-shaped.
#include "kernels.cuh"

namespace g4s {

enum { DF_CLONE = 1, DF_SPLIT = 2, DF_PRUNE_SELF = 4, DF_PRUNE_CHILD = 8 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256) densify_classify_kernel(int P, const float* __restrict__ accum, const float* __restrict__ denom,
                                                               const float* __restrict__ scaling, const float* __restrict__ opacity,
                                                               float grad_threshold, float dense_extent, float min_opacity,
                                                               float big_ws, float child_scale, uint8_t* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float g = accum[i] / denom[i];                 // gaussian_model.py:626-627: grads = accum / denom; NaN -> 0
    if (g != g) g = 0.0f;
    const float s0 = expf(scaling[2 * i]), s1 = expf(scaling[2 * i + 1]);
    const float maxs = fmaxf(s0, s1);
    const bool hot = fabsf(g) >= grad_threshold;   // torch.norm(grads, dim=-1) of a [P,1] tensor
    uint8_t f = 0;
    if (hot && maxs <= dense_extent) f |= DF_CLONE;   // :603-606
    if (hot && maxs > dense_extent) f |= DF_SPLIT;    // :574-577
    // final prune (:632-636) on the tensors as they are after clone + split.  max_radii2D was reset to zero by
    // densification_postfix (:566), so the screen-size test never fires; big_ws < 0 means max_screen_size is unset.
    const float op = sigmoidf_(opacity[i]);
    if (op < min_opacity || (big_ws >= 0.0f && maxs > big_ws)) f |= DF_PRUNE_SELF;
    // children carry scaling = log(exp(s) / (0.8 N)) (:584): what get_scaling returns for them
    const float c0 = expf(logf(s0 * child_scale)), c1 = expf(logf(s1 * child_scale));
    if (op < min_opacity || (big_ws >= 0.0f && fmaxf(c0, c1) > big_ws)) f |= DF_PRUNE_CHILD;
    flags[i] = f;
}

struct DensifyTensors {
    // xyz, features_dc, features_rest, opacity, scaling, rotation: parameter, exp_avg, exp_avg_sq (null when the
    // optimizer holds no state for it)
    const float* src[18];
    float* dst[18];
};

// one warp per output row
__global__ void __launch_bounds__(256) densify_gather_kernel(int P_new, int rest_w, const int* __restrict__ src_row,
                                                             const uint8_t* __restrict__ kind, const int* __restrict__ sample_row,
                                                             const float* __restrict__ samples, float child_scale, DensifyTensors t) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= P_new) return;
    const int s = src_row[row];
    const int k = kind[row];                       // 0 survivor, 1 clone, 2 split child
    const int widths[6] = {3, 3, rest_w, 1, 2, 4};
#pragma unroll
    for (int ti = 0; ti < 6; ti++) {
        const int w = widths[ti];
        const float* ps = t.src[3 * ti];
        float* pd = t.dst[3 * ti];
        for (int c = lane; c < w; c += 32) {
            float v = ps[(size_t)s * w + c];
            if (k == 2 && ti == 0) {
                // new_xyz = R(normalize(q)) (sample_x, sample_y, 0) + xyz   (:580-583, utils/general_utils.py build_rotation)
                const float* q = t.src[15] + 4 * (size_t)s;
                const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
                const float r = q[0] / nrm, x = q[1] / nrm, y = q[2] / nrm, z = q[3] / nrm;
                const float* sm = samples + 3 * (size_t)sample_row[row];
                const float R0 = (c == 0) ? 1 - 2 * (y * y + z * z) : (c == 1) ? 2 * (x * y + r * z) : 2 * (x * z - r * y);
                const float R1 = (c == 0) ? 2 * (x * y - r * z) : (c == 1) ? 1 - 2 * (x * x + z * z) : 2 * (y * z + r * x);
                const float R2 = (c == 0) ? 2 * (x * z + r * y) : (c == 1) ? 2 * (y * z - r * x) : 1 - 2 * (x * x + y * y);
                v = (R0 * sm[0] + R1 * sm[1] + R2 * sm[2]) + v;
            }
            if (k == 2 && ti == 4) v = logf(expf(v) * child_scale);
            pd[(size_t)row * w + c] = v;
            // Adam moments: carried for survivors, zero for everything new (cat_tensors_to_optimizer :554-555)
            if (t.dst[3 * ti + 1]) t.dst[3 * ti + 1][(size_t)row * w + c] = (k == 0 && t.src[3 * ti + 1]) ? t.src[3 * ti + 1][(size_t)s * w + c] : 0.0f;
            if (t.dst[3 * ti + 2]) t.dst[3 * ti + 2][(size_t)row * w + c] = (k == 0 && t.src[3 * ti + 2]) ? t.src[3 * ti + 2][(size_t)s * w + c] : 0.0f;
        }
    }
}

void launch_densify_classify(int P, const float* accum, const float* denom, const float* scaling, const float* opacity,
                             float grad_threshold, float dense_extent, float min_opacity, float big_ws, float child_scale,
                             uint8_t* flags, cudaStream_t s) {
    if (P <= 0) return;
    densify_classify_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, accum, denom, scaling, opacity, grad_threshold, dense_extent,
                                                          min_opacity, big_ws, child_scale, flags);
    count_launch();
}
void launch_densify_gather(int P_new, int rest_w, const int* src_row, const uint8_t* kind, const int* sample_row, const float* samples,
                           float child_scale, const float* const* src, float* const* dst, cudaStream_t s) {
    if (P_new <= 0) return;
    DensifyTensors t;
    for (int i = 0; i < 18; i++) { t.src[i] = src[i]; t.dst[i] = dst[i]; }
    const long long threads = (long long)P_new * 32;
    densify_gather_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(P_new, rest_w, src_row, kind, sample_row, samples, child_scale, t);
    count_launch();
}

}  // namespace g4s
