// kernels.cuh -- argument blocks and launchers shared between the translation units.
#pragma once
#include "common.cuh"

namespace g4s {

void count_launch();  // bumps the process-wide launch counter (api.cu)

struct ProjectArgs {
    int P, D, M, W, H, grid_x, grid_y, prefiltered;
    // raw != 0: the trainer's parameters before activation: shs = _features_dc [P,1,3],
    // sh_rest = _features_rest [P,M-1,3], opacities / scales / rotations un-activated, mip_filter [P] or null
    int raw; const float* sh_rest; const float* mip_filter;
    const float* means3D; const float* shs; const float* colors_precomp; const float* opacities;
    const float* scales; float scale_modifier; const float* rotations; const float* transMat_precomp;
    const float* view; const float* proj; const float* campos;
    int* radii;
    GeomView geom;
    uint32_t* tile_count;
    int32_t* counters;
};
void launch_project_fwd(const ProjectArgs& a, cudaStream_t s);

struct TileScanArgs {
    int num_tiles;
    uint32_t* tile_count;   // in: counts; out: zeroed (becomes the scatter cursor)
    uint32_t* tile_offset;  // [T+1]
    uint32_t* tile_order;   // [T] longest list first
    int32_t* counters;
};
void launch_tile_scan(const TileScanArgs& a, cudaStream_t s);

struct ScatterArgs {
    int P, grid_x;
    int64_t capacity;
    GeomView geom;
    const uint32_t* tile_offset;
    uint32_t* tile_cursor;
    unsigned long long* keys;
    const int32_t* counters;
};
void launch_scatter(const ScatterArgs& a, cudaStream_t s);

struct TileSortArgs {
    int num_tiles;
    int64_t capacity;
    const uint32_t* tile_offset;
    const uint32_t* tile_order;
    unsigned long long* keys;
    uint32_t* list;
    const int32_t* counters;
};
void launch_tile_sort(const TileSortArgs& a, cudaStream_t s);

struct BlendFwdArgs {
    int W, H, grid_x, grid_y;
    int64_t capacity;
    const uint32_t* tile_offset;
    const uint32_t* tile_order;
    const uint32_t* list;
    uint32_t* masks;       // [cap * 8] written, per tile [8 regions][n entries]: lanes of region w that blended entry j
    int fast_math;         // 0: IEEE division / expf, bit-identical to the reference; 1: rcp.approx / ex2.approx
    const float4* rec;
    const float* bg;
    float* final_T;        // [3][N]
    uint32_t* n_contrib;   // [2][N]
    float* out_color;      // [3][N]
    float* out_others;     // [7][N]
    const int32_t* counters;
};
void launch_blend_fwd(const BlendFwdArgs& a, cudaStream_t s);

struct BlendBwdArgs {
    int W, H, grid_x, grid_y;
    const uint32_t* tile_offset;
    const uint32_t* tile_order;
    const uint32_t* list;
    const uint32_t* masks;
    const float4* rec;
    const float* bg;
    const float* final_T;
    const uint32_t* n_contrib;
    const float* dL_dpix;     // [3][N]
    const float* dL_dothers;  // [7][N]
    float4* acc;              // [P][5], zeroed by the caller
    const int32_t* counters;
    int64_t capacity;
};
void launch_blend_bwd(const BlendBwdArgs& a, cudaStream_t s);

enum { ACC_MEANS3D = 1, ACC_SH = 2, ACC_OPACITY = 4, ACC_SCALES = 8, ACC_ROTATIONS = 16,
       ACC_MULTIMEM = 32 /* the accumulated outputs are NVSwitch multicast addresses: multimem.red */ };
struct ProjectBwdArgs {
    int P, D, M, W, H;
    int raw; const float* sh_rest; const float* opacities; const float* mip_filter; float* dL_dsh_rest;  // see ProjectArgs
    int accumulate;  // ACC_* bits: add into that output (visible rows only) instead of overwriting it
    const float* means3D; const float* shs; const float* scales; const float* rotations;
    const float* view; const float* proj; const float* campos;
    float focal_x, focal_y, tan_fovx, tan_fovy;
    const int* radii;
    GeomView geom;
    const float4* acc;
    const int32_t* counters;   // [CNT_N] of the forward (CNT_VISIBLE sizes geom.visible_list)
    float* dL_dmeans3D; float* dL_dmeans2D; float* dL_dsh; float* dL_dcolors; float* dL_dopacity;
    float* dL_dscales; float* dL_drots; float* dL_dtransMat;
};
void launch_project_bwd(const ProjectBwdArgs& a, cudaStream_t s);
void launch_acc_clear(int P, const int* radii, float4* acc, cudaStream_t s);

void launch_densify_stats(int P, const float* dL_dmeans2D, const int* radii, float* accum, float* denom,
                          int* max_radii, int multimem, cudaStream_t s);
void launch_multimem_allreduce(float* mc, size_t n_floats, int* mc_max, size_t n_ints, int rank, int world, cudaStream_t s);
void launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present, cudaStream_t s);

// ---- compute_mip_filter (gaussian_model.cu) ---------------------------------------------------
void launch_mip_filter(int P, const float* xyz, int C, const float* cams, float znear, float focal_length,
                       float sqrt_variance, float* filter, unsigned int* max_bits, cudaStream_t s);

// ---- photometric loss (loss.cu) ---------------------------------------------------------------
void launch_photometric_fwd(int W, int H, int C, const float* img, const float* gt, const float* window11, float lambda,
                            double* sums, float* dmaps, float* out3, cudaStream_t s);
void launch_photometric_bwd(int W, int H, int C, const float* img, const float* gt, const float* window11, float lambda,
                            const float* dmaps, const float* dL_dloss, float* dL_dimg, cudaStream_t s);

// ---- densification (densify.cu) ------------------------------------------------------------------
void launch_densify_classify(int P, const float* accum, const float* denom, const float* scaling, const float* opacity,
                             float grad_threshold, float dense_extent, float min_opacity, float big_ws, float child_scale,
                             uint8_t* flags, cudaStream_t s);
void launch_densify_gather(int P_new, int rest_w, const int* src_row, const uint8_t* kind, const int* sample_row, const float* samples,
                           float child_scale, const float* const* src, float* const* dst, cudaStream_t s);

// ---- image-space regularisers (regularizers.cu) -------------------------------------------------
void launch_normal2curv_fwd(int W, int H, const float* normal, const float* mask, float* curv, float* sg, cudaStream_t s);
void launch_normal2curv_bwd(int W, int H, const float* mask, const float* sg, const float* g_curv, float* g_normal, cudaStream_t s);
void launch_depth_order_fwd(int W, int H, const float* depth, const float* prior, const long long* shifts, float inv_extent,
                            int normalize, int log_space, float log_scale, float* per_pixel, double* sum, cudaStream_t s);
void launch_depth_order_bwd(int W, int H, const float* depth, const float* prior, const long long* shifts, float inv_extent,
                            int normalize, int log_space, float log_scale, const float* g, const float* g_scalar, float scale,
                            float* g_depth, cudaStream_t s);

// ---- render() post-processing (surface.cu) ----------------------------------------------------
struct SurfaceFwdArgs {
    int W, H;
    float r0, r1;            // (float)(1 - depth_ratio), (float)depth_ratio
    const float* allmap;     // [7][N]
    const float* view; const float* proj;
    float* rend_alpha; float* rend_normal; float* rend_normal_cam; float* rend_dist;
    float* surf_depth; float* surf_normal; float* surf_normal_cam; float* rend_depth;
};
void launch_surface_fwd(const SurfaceFwdArgs& a, cudaStream_t s);

struct SurfaceBwdArgs {
    int W, H;
    float r0, r1;
    const float* allmap;
    const float* view; const float* proj;
    // upstream gradients, any may be null (= zeros)
    const float* g_rend_alpha; const float* g_rend_normal; const float* g_rend_normal_cam; const float* g_rend_dist;
    const float* g_surf_depth; const float* g_surf_normal; const float* g_surf_normal_cam; const float* g_rend_depth;
    float* g_allmap;         // [7][N], every element written
};
void launch_surface_bwd(const SurfaceBwdArgs& a, cudaStream_t s);

}  // namespace g4s
