/*
 * g4s_rasterizer.h -- C ABI of the B200-native 2D-Gaussian (surfel) rasterizer.
 *
 * This is the drop-in boundary for the hot path of DaLi-Jack/G4Splat: it replaces the three
 * entry points the reference binds with pybind11/libtorch
 *   RAST = 2d-gaussian-splatting/submodules/diff-surfel-rasterization
 *   RAST/ext.cpp:15-19                      PYBIND11_MODULE(_C): rasterize_gaussians,
 *                                           rasterize_gaussians_backward, mark_visible
 *   RAST/rasterize_points.h:17-67           their C++ signatures (torch::Tensor in / out)
 *   RAST/cuda_rasterizer/rasterizer.h:24-86 CudaRasterizer::Rasterizer::{forward,backward,markVisible}
 * with plain-pointer functions: no torch types, no C++ types, every buffer owned by the caller.
 * All pointers are DEVICE pointers unless a parameter says "host".  Every launch goes to the
 * `stream` argument (a cudaStream_t passed as void*; NULL = legacy default stream).  No call
 * synchronises the host unless `debug` is non-zero (reference: CHECK_CUDA, auxiliary.h:295-302).
 *
 * Optional inputs follow the reference convention (RAST/diff_surfel_rasterization/__init__.py:
 * 198-208 -> empty tensor -> null data pointer -> kernel branch): pass NULL.
 *   exactly one of  shs | colors_precomp          must be non-NULL
 *   exactly one of  (scales and rotations) | transMat_precomp   must be non-NULL
 *
 * Return value: 0 on success, a negative G4S_E* code otherwise; g4s_last_error() returns a
 * thread-local human-readable string for the last failing call.
 *
 * Scratch buffers (the reference's geomBuffer / binningBuffer / imgBuffer byte tensors,
 * RAST/rasterize_points.cu:89-96) are opaque byte blobs sized by the g4s_*_bytes() functions and
 * must be kept alive, unmodified, from g4s_forward_* to the matching g4s_backward.
 *
 * Forward is split in two calls so that the host never has to block in the middle of the
 * pipeline to size the per-tile instance lists (the reference does a blocking cudaMemcpy of
 * num_rendered, rasterizer_impl.cu:282):
 *   g4s_forward_plan    project + exact tile culling + per-tile counts + scan; copies
 *                       {num_rendered, max tile list length} to `host_counts` asynchronously.
 *   g4s_forward_render  scatter + per-tile depth sort + blend, for a binning buffer of
 *                       `capacity` instances.  Every kernel in it checks num_rendered <= capacity
 *                       on the device and turns into a no-op otherwise, so the caller may launch
 *                       it speculatively with a guessed capacity, wait only for the plan event,
 *                       and re-issue it with a larger buffer in the rare overflow case.
 */
#ifndef G4S_RASTERIZER_H_
#define G4S_RASTERIZER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G4S_VERSION 100 /* 0.1.0 */

enum {
    G4S_OK = 0,
    G4S_EINVAL = -1,   /* bad argument combination / sizes */
    G4S_ECUDA = -2,    /* a CUDA runtime call or (debug) a kernel failed */
    G4S_ECAPACITY = -3 /* debug only: num_rendered > capacity */
};

int g4s_version(void);
const char* g4s_last_error(void);

/* Arithmetic of the forward blend (process-wide; returns the previous setting).
 *   0 (default)  IEEE division and expf in the reference's rounding sequence: colour, the seven allmap
 *                channels and radii are bit-identical to the reference (CR/forward.cu:356-419).
 *   1            rcp.approx / ex2.approx (2 ulp) for the ray-splat solve, the Gaussian weight and the NDC
 *                depth: values within 1e-5; a pixel whose alpha, T < 1e-4 or T > 0.5 decision sits within an
 *                ulp of its threshold may flip (north_star's 1e-4 parity with a counted flip budget,
 *                tests/test_parity_gpu.py).  The backward is value-only and identical in both modes. */
int g4s_set_fast_math(int on);
int g4s_get_fast_math(void);

/* ---- scratch sizes ------------------------------------------------------------------------- */
/* per-Gaussian state (projected records, clamp masks, depths): reference GeometryState,
 * rasterizer_impl.cu:155-170 */
size_t g4s_geom_bytes(int P);
/* per-pixel + per-tile state (final T / M1 / M2, last + median contributor, tile counts,
 * offsets): reference ImageState, rasterizer_impl.cu:172-179 */
size_t g4s_image_bytes(int W, int H);
/* per-instance state for `capacity` (Gaussian, tile) instances (sorted lists, per-warp
 * contribution masks, unsorted 64-bit keys): reference BinningState, rasterizer_impl.cu:181-194 */
size_t g4s_binning_bytes(int64_t capacity);
/* backward scratch: per-Gaussian gradient accumulators of the blend stage (the reference's
 * dL_dtransMat / dL_dnormal / dL_dcolors / dL_dmean2D temporaries, rasterize_points.cu:187-195) */
size_t g4s_backward_scratch_bytes(int P);

/* ---- forward ------------------------------------------------------------------------------- */
/* Replaces the first half of CudaRasterizer::Rasterizer::forward (rasterizer_impl.cu:198-283).
 * host_counts: PINNED host int32[4], filled asynchronously on `stream`:
 *   [0] num_rendered (instances after exact tile culling)   [1] longest tile list
 *   [2] number of visible Gaussians (radii > 0)              [3] reserved
 * radii[P] (int32, output) is fully written here. */
int g4s_forward_plan(int P, int D, int M, int W, int H,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                     float tan_fovx, float tan_fovy, int prefiltered,
                     int* radii, void* geom_buffer, void* img_buffer,
                     int32_t* host_counts, void* stream, int debug);

/* Replaces the second half of Rasterizer::forward (rasterizer_impl.cu:285-341).
 * out_color[3,H,W], out_others[7,H,W] are fully written (callers need not zero them). */
int g4s_forward_render(int P, int W, int H, const float* background,
                       const void* geom_buffer, void* img_buffer,
                       void* binning_buffer, int64_t capacity,
                       float* out_color, float* out_others, void* stream, int debug);
/* The two halves of g4s_forward_render, for callers that run the front end of a view (plan + bin: every kernel
 * before the blend, all latency-bound) on a side stream while the previous view's blend kernels occupy the main
 * one (diff_surfel_rasterization.view_batch): g4s_forward_bin = scatter + per-tile sort, g4s_forward_blend = the
 * blend.  g4s_forward_render(...) == g4s_forward_bin(...) followed by g4s_forward_blend(...) on the same stream. */
int g4s_forward_bin(int P, int W, int H, const void* geom_buffer, void* img_buffer, void* binning_buffer,
                    int64_t capacity, void* stream, int debug);
int g4s_forward_blend(int P, int W, int H, const float* background, const void* geom_buffer, void* img_buffer,
                      void* binning_buffer, int64_t capacity, float* out_color, float* out_others, void* stream, int debug);

/* ---- backward ------------------------------------------------------------------------------ */
/* Replaces CudaRasterizer::Rasterizer::backward (rasterizer_impl.cu:346-448) and the gradient
 * allocation of RasterizeGaussiansBackwardCUDA (rasterize_points.cu:187-195): every output
 * element is written (zeros for Gaussians that were not visible), callers need not zero them.
 * Outputs (device): dL_dmeans3D[P,3] dL_dmeans2D[P,3] dL_dsh[P,M,3] (may be NULL when M == 0)
 *   dL_dcolors[P,3] dL_dopacity[P] dL_dscales[P,2] dL_drotations[P,4] dL_dtransMat[P,9].
 *   dL_dcolors may be NULL when colors_precomp is NULL, dL_dtransMat when transMat_precomp is NULL
 *   (the reference returns dense tensors autograd then drops; 48 B per Gaussian of writes saved).
 * scratch: g4s_backward_scratch_bytes(P) bytes.  capacity: the value g4s_forward_render was given
 * for this binning_buffer.
 * accumulate_mask (0 = reference behaviour): G4S_ACC_* bits select outputs that are running sums
 * over several views (multi-view steps, view-sharded training): the kernel ADDS the rows of
 * visible Gaussians into them and does not touch the other rows, which replaces autograd's
 * dense `grad += new` pass (3 x 232 B per Gaussian at SH degree 3) by a sparse read-modify-write. */
#define G4S_ACC_MEANS3D 1
#define G4S_ACC_SH 2
#define G4S_ACC_OPACITY 4
#define G4S_ACC_SCALES 8
#define G4S_ACC_ROTATIONS 16
/* The accumulated outputs are NVSwitch multicast addresses (one buffer per rank mapped at the same offset of
 * a CUDA multicast object): the kernel adds with multimem.red, so every rank's replica receives this view's
 * gradient and the per-step all-reduce of the view-sharded trainer needs no data movement of its own
 * (fused producer + collective).  Requires NVLS-capable hardware; the caller owns the barrier
 * that orders these reductions before any rank reads its replica. */
#define G4S_ACC_MULTIMEM 32
int g4s_backward(int P, int D, int M, int W, int H, const float* background,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* transMat_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* cam_pos, float tan_fovx, float tan_fovy, const int* radii,
                 const void* geom_buffer, const void* binning_buffer, int64_t capacity,
                 const void* img_buffer,
                 const float* dL_dout_color, const float* dL_dout_others,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors,
                 float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dtransMat,
                 int accumulate_mask, void* scratch, void* stream, int debug);

/* ---- raw-parameter entry points ---------------------------------------------
 * The trainer calls the operator with ACTIVATED parameters: every render() runs exp / sigmoid /
 * normalize / cat (and, with the mip filter, sqrt(s^2 + f^2) and the opacity compensation) as ~10
 * torch kernels plus their autograd (2DGS/scene/gaussian_model.py:158-192 get_scaling, get_rotation,
 * get_features, get_opacity; gaussian_renderer/__init__.py:55-106).  These two calls take the
 * optimiser's own leaves instead, apply the activations in registers inside the projection kernels
 * and return gradients with respect to the raw leaves:
 *   xyz[P,3]  features_dc[P,1,3]  features_rest[P,M-1,3] (NULL when M == 1)  opacity_raw[P,1]
 *   scaling_raw[P,2]  rotation_raw[P,4]  mip_filter[P] or NULL (use_mip_filter off)
 * M counts all SH coefficients (1 + rest).  Everything else as in g4s_forward_plan / g4s_backward;
 * g4s_forward_render is shared.  scratch for g4s_backward_raw: g4s_backward_scratch_bytes_raw(P). */
size_t g4s_backward_scratch_bytes_raw(int P);
int g4s_forward_plan_raw(int P, int D, int M, int W, int H,
                         const float* xyz, const float* features_dc, const float* features_rest,
                         const float* opacity_raw, const float* scaling_raw, float scale_modifier,
                         const float* rotation_raw, const float* mip_filter,
                         const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                         float tan_fovx, float tan_fovy, int prefiltered,
                         int* radii, void* geom_buffer, void* img_buffer,
                         int32_t* host_counts, void* stream, int debug);
int g4s_backward_raw(int P, int D, int M, int W, int H, const float* background,
                     const float* xyz, const float* features_dc, const float* features_rest,
                     const float* opacity_raw, const float* scaling_raw, float scale_modifier,
                     const float* rotation_raw, const float* mip_filter,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                     float tan_fovx, float tan_fovy, const int* radii,
                     const void* geom_buffer, const void* binning_buffer, int64_t capacity,
                     const void* img_buffer,
                     const float* dL_dout_color, const float* dL_dout_others,
                     float* dL_dxyz, float* dL_dmeans2D, float* dL_dfeatures_dc, float* dL_dfeatures_rest,
                     float* dL_dopacity_raw, float* dL_dscaling_raw, float* dL_drotation_raw,
                     int accumulate_mask, void* scratch, void* stream, int debug);

/* ---- markVisible --------------------------------------------------------------------------- */
/* Replaces Rasterizer::markVisible / checkFrustum (rasterizer_impl.cu:54-66,141-153):
 * present[i] = (viewmatrix * means3D[i]).z > 0.2.  present is uint8[P] (bool). */
int g4s_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* ---- densification statistics (caller-side row "next") ---------------------- */
/* One fused pass over what the trainer does with the operator's outputs after every backward
 * (2DGS/scene/gaussian_model.py:649-651 add_densification_stats, train_with_refine_depth.py:583):
 * for radii[i] > 0:  accum[i] += |dL_dmeans2D[i].xy|, denom[i] += 1, max_radii[i] = max(., radii[i]). */
int g4s_densify_stats(int P, const float* dL_dmeans2D, const int* radii, float* accum, float* denom,
                      int* max_radii, void* stream);
/* Same pass with accum / denom / max_radii given as NVSwitch MULTICAST addresses of buffers every rank holds
 * at the same offset (CUDA multicast object, e.g. torch symmetric memory's multicast_ptr): the update of a
 * visible Gaussian is one multimem.red per value and lands in every rank's replica (G4S_ACC_MULTIMEM above). */
int g4s_densify_stats_multimem(int P, const float* dL_dmeans2D, const int* radii, float* accum_mc, float* denom_mc,
                               int* max_radii_mc, void* stream);

/* ---- adaptive density control --------------------------------------------------------------------
 * Replaces GaussianModel.densify_and_prune and its helpers (2DGS/scene/gaussian_model.py:528-647; driven from
 * train_with_refine_depth.py:582-599): ~60 torch kernels (masks, masked gathers, cats, for six parameters and both
 * Adam moments, three times) become two.
 * g4s_densify_classify: per Gaussian, from accum / denom (xyz_gradient_accum, denom), the raw scaling [P,2] and raw
 *   opacity [P]: flags[i] bit 0 = cloned (|grad| >= grad_threshold, max scale <= dense_extent = percent_dense * extent),
 *   bit 1 = split (same gradient test, max scale > dense_extent), bit 2 = the Gaussian (and its clone) is pruned at the
 *   end (sigmoid(opacity) < min_opacity, or max scale > big_world_size when big_world_size >= 0), bit 3 = its split
 *   children (scale / (0.8 n_split)) are pruned.
 * g4s_densify_gather: writes the P_new final rows.  src_row[j] = source Gaussian of output row j, kind[j] = 0 survivor /
 *   1 clone / 2 split child, sample_row[j] = row of `samples` ([n,3], the caller's torch.normal draw) for children.
 *   src_tensors / dst_tensors: HOST arrays of 18 device pointers = {xyz, features_dc, features_rest, opacity, scaling,
 *   rotation} x {parameter, exp_avg, exp_avg_sq}; moment pointers may be NULL (no optimizer state).  Children get
 *   xyz + R(q) sample and log(exp(scaling) / (0.8 n_split)); new rows get zero moments. */
int g4s_densify_classify(int P, const float* accum, const float* denom, const float* scaling_raw, const float* opacity_raw,
                         float grad_threshold, float dense_extent, float min_opacity, float big_world_size, int n_split,
                         uint8_t* flags, void* stream);
int g4s_densify_gather(int P_new, int rest_width, const int* src_row, const uint8_t* kind, const int* sample_row,
                       const float* samples, int n_split, const float* const* src_tensors, float* const* dst_tensors, void* stream);

/* ---- gradient all-reduce of the view-sharded trainer over NVSwitch multicast ---------------------
 * sum_mc / max_mc: MULTICAST addresses (CUDA multicast object; torch symmetric memory's multicast_ptr) of an fp32
 * block of n_floats (multiple of 4, 16-byte aligned) and an int32 block of n_ints that every rank holds at the same
 * offset.  Rank `rank` of `world` combines its slice of both blocks across all ranks with multimem.ld_reduce (add /
 * max, inside the switch) and writes the result to every replica with multimem.st.  The caller orders it with a
 * barrier on either side (all ranks' partial sums complete / all slices written). */
int g4s_multimem_allreduce(float* sum_mc, int64_t n_floats, int* max_mc, int64_t n_ints, int rank, int world, void* stream);

/* ---- image-space regularisers (the part beside the photometric loss) -------------
 * normal2curv: matcha/dm_utils/rendering.py:392-406 (2DGS/train_with_refine_depth.py:415):
 *   curv[1,H,W] = L1 norm over channels of  mask * sum_{up,left,bottom,right} (n_nb - n_c m_c) m_nb, replicate padding.
 *   normal [3,H,W]; mask [1,H,W] fp32 or NULL (= ones; the trainer passes ones).  sign_map [3,H,W] (optional
 *   output, required by the backward): sign of each channel's sum times the mask.
 * depth-order loss: matcha/dm_regularization/depth.py:142-222 (train_with_refine_depth.py:465): every pixel is
 *   paired with the pixel `pixel_shifts[i]` away (int64 [N,2] = (dy, dx), drawn by the caller exactly as the
 *   reference draws them, clamped to the image here); loss_i = -min(diff * prior_diff, 0) with both differences
 *   divided by scene_extent, prior_diff optionally reduced to its sign, optionally log(1 + log_scale * loss_i).
 *   per_pixel [N] and / or sum (device double[1]) are written.  Backward: dL_dper_pixel [N], or NULL with the
 *   scalar upstream dL_dloss (device float[1], NULL = 1) times `scale` (1/N for the mean); dL_ddepth [N] is
 *   fully written. */
int g4s_normal2curv_forward(int W, int H, const float* normal, const float* mask, float* curv, float* sign_map, void* stream);
int g4s_normal2curv_backward(int W, int H, const float* mask, const float* sign_map, const float* dL_dcurv, float* dL_dnormal,
                             void* stream);
int g4s_depth_order_forward(int W, int H, const float* depth, const float* prior_depth, const int64_t* pixel_shifts,
                            float scene_extent, int normalize_loss, int log_space, float log_scale, float* per_pixel,
                            double* sum, void* stream);
int g4s_depth_order_backward(int W, int H, const float* depth, const float* prior_depth, const int64_t* pixel_shifts,
                             float scene_extent, int normalize_loss, int log_space, float log_scale, const float* dL_dper_pixel,
                             const float* dL_dloss, float scale, float* dL_ddepth, void* stream);

/* ---- photometric loss ---------------------------------------------------------
 * Replaces what the trainer does with the rendered image every iteration
 * (2DGS/train_with_refine_depth.py:382-383 with utils/loss_utils.py:17-18 l1_loss and :49-80 ssim):
 *   loss = (1 - lambda_dssim) * mean|image - gt| + lambda_dssim * (1 - mean(ssim_map(image, gt)))
 * -- five depthwise 11x11 conv2d launches, ~15 elementwise kernels and their autograd in the
 * reference; one kernel per direction here (+ a one-thread finishing kernel).
 *   image, gt: [C,H,W] fp32 (device).  window11: HOST float[11], the 1-D Gaussian of
 *   loss_utils.gaussian(11, 1.5) (the reference's 2-D window is its outer product).
 *   sums: DEVICE double[2] scratch (sum |x-y|, sum ssim).  out3: DEVICE float[3] = {loss, l1, ssim}.
 *   dmaps: DEVICE float[3][C][H][W] written by the forward and read by the backward (the three
 *   partial derivatives of the per-pixel ssim with respect to its window sums); NULL = forward only.
 * Backward: dL_dloss = DEVICE float[1] upstream gradient of `loss` (NULL = 1); dL_dimage[C,H,W] is
 * fully written.  Zero padding at the image border, like conv2d(padding=5). */
int g4s_photometric_forward(int W, int H, int C, const float* image, const float* gt,
                            const float* window11, float lambda_dssim, double* sums, float* dmaps,
                            float* out3, void* stream);
int g4s_photometric_backward(int W, int H, int C, const float* image, const float* gt,
                             const float* window11, float lambda_dssim, const float* dmaps,
                             const float* dL_dloss, float* dL_dimage, void* stream);

/* ---- compute_mip_filter ------------------------------------------------------
 * Replaces GaussianModel.compute_mip_filter (2DGS/scene/gaussian_model.py:388-434): a Python loop
 * over ALL cameras with ~14 torch kernels each on [P]-sized tensors.  Two launches here.
 *   cameras: DEVICE float[num_cameras][20] = { R[9] (camera.R row-major; xyz_cam = xyz @ R + T),
 *            T[3], focal_x, focal_y, W/2, H/2, -0.15 W, 1.15 W, -0.15 H, 1.15 H }
 *   focal_length = max over cameras of focal_x (the reference takes it on the host, :428-429)
 *   sqrt_filter_variance = filter_variance ** 0.5
 *   mip_filter[P] (output) = distance / focal_length * sqrt_filter_variance, distance = the smallest
 *            clamped camera-space depth over the cameras that see the point (z > znear, projection
 *            inside the image enlarged by 15 %), or the largest such distance of any point for
 *            points no camera sees (:431).
 *   max_distance_bits: DEVICE uint32 scratch; afterwards the fp32 bits of that largest distance,
 *            0 when NO point is seen by any camera (the reference raises there: max() of an empty
 *            tensor), mip_filter is then 0. */
int g4s_mip_filter(int P, const float* xyz, int num_cameras, const float* cameras, float znear,
                   float focal_length, float sqrt_filter_variance, float* mip_filter,
                   uint32_t* max_distance_bits, void* stream);

/* ---- render() post-processing --------------------------------------------
 * Replaces the ~15 torch kernels (and their autograd) that follow every rasterizer call in
 * 2d-gaussian-splatting/gaussian_renderer/__init__.py:118-164 and utils/point_utils.py:9-37
 * (depths_to_points, depth_to_normal): one kernel per direction, no host synchronisation (the
 * reference inverts two matrices per call, which synchronises).
 *   allmap[7,H,W]: the rasterizer's second image output; viewmatrix = world_view_transform[4,4],
 *   projmatrix = full_proj_transform[4,4] (device, row-major as torch stores them);
 *   depth_ratio = pipe.depth_ratio.
 * Forward writes (device, caller-allocated):
 *   rend_alpha[1,H,W] rend_normal[3,H,W] (world) rend_normal_cam[3,H,W] rend_dist[1,H,W]
 *   surf_depth[1,H,W] surf_normal[3,H,W] (world, times alpha) surf_normal_cam[3,H,W] rend_depth[1,H,W]
 * Backward: any upstream gradient pointer may be NULL (= zeros); dL_dallmap[7,H,W] is fully written.
 * Where alpha == 0 the reference's autograd yields NaN for dL_dallmap[0:2] (0/0 in the division
 * backward); those pixels have no contributor and the rasterizer ignores their gradients; here the
 * masked terms contribute zero, so every written value is finite. */
int g4s_surface_forward(int W, int H, const float* allmap, const float* viewmatrix,
                        const float* projmatrix, double depth_ratio, float* rend_alpha,
                        float* rend_normal, float* rend_normal_cam, float* rend_dist,
                        float* surf_depth, float* surf_normal, float* surf_normal_cam,
                        float* rend_depth, void* stream);
int g4s_surface_backward(int W, int H, const float* allmap, const float* viewmatrix,
                         const float* projmatrix, double depth_ratio, const float* dL_drend_alpha,
                         const float* dL_drend_normal, const float* dL_drend_normal_cam,
                         const float* dL_drend_dist, const float* dL_dsurf_depth,
                         const float* dL_dsurf_normal, const float* dL_dsurf_normal_cam,
                         const float* dL_drend_depth, float* dL_dallmap, void* stream);

/* ---- introspection (tests, benchmarks) ----------------------------------------------------- */
/* Copies decoded views of the opaque buffers into caller-provided DEVICE arrays (any may be
 * NULL).  Lets stage-level parity tests compare against the oracle without knowing the layout.
 *   transMat[P,9] means2D[P,2] normal_opacity[P,4] rgb[P,3] depths[P] bbox[P,4] clamped[P,3](u8)
 *   tiles_touched[P](u32) */
int g4s_debug_decode_geom(int P, const void* geom_buffer, float* transMat, float* means2D,
                          float* normal_opacity, float* rgb, float* depths, float* bbox,
                          uint8_t* clamped, uint32_t* tiles_touched, void* stream);
/*   ranges[T,2](u32) final_T[3,H,W] n_contrib[2,H,W](u32); point_list[capacity](u32) */
int g4s_debug_decode_lists(int W, int H, const void* img_buffer, const void* binning_buffer,
                           int64_t capacity, uint32_t* ranges, float* final_T, uint32_t* n_contrib,
                           uint32_t* point_list, void* stream);
/* Work counters of one rendered view (secondary roofline) from its three scratch
 * buffers (valid after g4s_forward_render with num_rendered <= capacity), DEVICE uint64 stats[8]:
 *   [0] (pixel, surfel) pairs the forward blended  [1] sum over pixels of the last contributor's
 *   position in this library's (culled) tile lists  [2] pair slots = 256 * num_rendered
 *   [3] longest tile list  [4] pair evaluations this library's backward issues (32 lanes per
 *   (instance, 8x4 region) whose forward mask is non-zero)  [5..7] zero */
int g4s_debug_pair_stats(int W, int H, const void* geom_buffer, int P, const void* img_buffer,
                         const void* binning_buffer, int64_t capacity, uint64_t* stats, void* stream);
/* number of kernel launches issued by this library since process start (bench: gpu_launches) */
int64_t g4s_launch_count(void);
/* Per-stage CUDA-event timers on the launch stream (process-global, for bench.py / profiling;
 * the reference only times whole iterations, train_with_refine_depth.py:279-280,364,498).
 * While enabled every stage launch is bracketed by two events; completed brackets are folded
 * into a running mean without blocking.  g4s_profile_read waits for the last brackets and
 * fills mean_ms_out[i] (-1 when stage i never ran) and count_out[i] (may be NULL) since the
 * last g4s_profile_enable.  Stage order:
 * project_fwd, tile_scan, scatter, tile_sort, blend_fwd, acc_clear, blend_bwd, project_bwd. */
int g4s_profile_enable(int on);
/* Bit i set: stage i is bracketed while profiling is on (default: all).  Every bracket is two event records
 * on the stream, i.e. two points where consecutive kernels cannot overlap their launch: measured on a B200, bracketing
 * all eight stages of a 1.5 ms view costs 2.8 % of the throughput, so a benchmark brackets everything in an untimed
 * pass and only the kernel its roofline is about inside the timed region. */
int g4s_profile_select(unsigned stage_mask);
int g4s_profile_num_stages(void);
const char* g4s_profile_stage_name(int i);
int g4s_profile_read(float* mean_ms_out, int64_t* count_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* G4S_RASTERIZER_H_ */
