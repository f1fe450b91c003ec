// api.cu -- the C ABI declared in include/g4s_rasterizer.h.  Host orchestration only: carve the
// caller's opaque buffers, fill argument blocks, launch on the caller's stream.
// Replaces CR/rasterizer_impl.cu:198-448 (Rasterizer::forward / backward / markVisible) and the
// libtorch binding RAST/rasterize_points.cu:39-254.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/g4s_rasterizer.h"
#include "kernels.cuh"

namespace g4s {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static thread_local std::string g_error;
// 0: the forward decides and blends with IEEE division / expf (bit-identical to the reference);
// 1: rcp.approx / ex2.approx in the forward blend (opt-in, g4s_set_fast_math)
static std::atomic<int> g_fast_math{0};

static int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
static int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return G4S_OK;
    return fail(G4S_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
// debug == reference CHECK_CUDA (CR/auxiliary.h:295-302): synchronise after every stage and throw
static int stage_check(bool debug, cudaStream_t s, const char* stage) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, stage);
    if (debug) return check_cuda(cudaStreamSynchronize(s), stage);
    return G4S_OK;
}

// ---- optional per-stage CUDA-event timers (bench / profiling only; process-global) -------------
enum { ST_PROJECT_FWD = 0, ST_TILE_SCAN, ST_SCATTER, ST_TILE_SORT, ST_BLEND_FWD, ST_ACC_CLEAR, ST_BLEND_BWD,
       ST_PROJECT_BWD, ST_COUNT };
static const char* const kStageNames[ST_COUNT] = {"project_fwd", "tile_scan", "scatter", "tile_sort", "blend_fwd",
                                                  "acc_clear", "blend_bwd", "project_bwd"};
static bool g_profile = false;
static unsigned g_profile_mask = ~0u;     // stages that are bracketed while g_profile is on (g4s_profile_select)
// A small ring of event pairs per stage: a slot is folded into the running mean when it comes up
// for reuse, RING launches later, by which time it has completed even when the host runs ahead of
// the device by a whole view.
constexpr int EV_RING = 4;
static cudaEvent_t g_ev_begin[ST_COUNT][EV_RING], g_ev_end[ST_COUNT][EV_RING];
static bool g_ev_valid[ST_COUNT][EV_RING] = {{false}};
static unsigned g_ev_next[ST_COUNT] = {0};
static bool g_ev_created = false;
static double g_ev_sum_ms[ST_COUNT] = {0};
static long long g_ev_n[ST_COUNT] = {0};
static void harvest(int stage, int slot, bool wait) {
    if (!g_ev_valid[stage][slot]) return;
    if (wait) { if (cudaEventSynchronize(g_ev_end[stage][slot]) != cudaSuccess) return; }
    else if (cudaEventQuery(g_ev_end[stage][slot]) != cudaSuccess) { cudaGetLastError(); g_ev_valid[stage][slot] = false; return; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev_begin[stage][slot], g_ev_end[stage][slot]) == cudaSuccess) { g_ev_sum_ms[stage] += ms; g_ev_n[stage]++; }
    g_ev_valid[stage][slot] = false;
}
struct StageTimer {
    int stage, slot; cudaStream_t s; bool on;
    StageTimer(int st, cudaStream_t stream) : stage(st), slot(0), s(stream), on(false) {
        on = g_profile && ((g_profile_mask >> stage) & 1u);
        if (on) {
            slot = (int)(g_ev_next[stage]++ % EV_RING);
            harvest(stage, slot, false);
            cudaEventRecord(g_ev_begin[stage][slot], s);
        }
    }
    ~StageTimer() {
        if (on) { cudaEventRecord(g_ev_end[stage][slot], s); g_ev_valid[stage][slot] = true; }
    }
};
}  // namespace g4s

using namespace g4s;

extern "C" {

int g4s_profile_enable(int on) {
    if (on && !g_ev_created) {
        for (int i = 0; i < ST_COUNT; i++)
            for (int k = 0; k < EV_RING; k++)
                if (cudaEventCreate(&g_ev_begin[i][k]) != cudaSuccess || cudaEventCreate(&g_ev_end[i][k]) != cudaSuccess)
                    return fail(G4S_ECUDA, "g4s_profile_enable: cudaEventCreate failed");
        g_ev_created = true;
    }
    g_profile = on != 0;
    for (int i = 0; i < ST_COUNT; i++) {
        for (int k = 0; k < EV_RING; k++) g_ev_valid[i][k] = false;
        g_ev_sum_ms[i] = 0; g_ev_n[i] = 0; g_ev_next[i] = 0;
    }
    return G4S_OK;
}
int g4s_profile_select(unsigned stage_mask) { g_profile_mask = stage_mask; return G4S_OK; }
int g4s_profile_num_stages(void) { return ST_COUNT; }
const char* g4s_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int g4s_profile_read(float* mean_ms_out, int64_t* count_out, int n) {
    for (int i = 0; i < n && i < ST_COUNT; i++) {
        for (int k = 0; k < EV_RING; k++) harvest(i, k, true);
        mean_ms_out[i] = g_ev_n[i] ? (float)(g_ev_sum_ms[i] / (double)g_ev_n[i]) : -1.0f;
        if (count_out) count_out[i] = g_ev_n[i];
    }
    return G4S_OK;
}

int g4s_version(void) { return G4S_VERSION; }
int g4s_set_fast_math(int on) { return g_fast_math.exchange(on != 0 ? 1 : 0); }
int g4s_get_fast_math(void) { return g_fast_math.load(); }
const char* g4s_last_error(void) { return g_error.c_str(); }
int64_t g4s_launch_count(void) { return (int64_t)g_launches.load(); }

size_t g4s_geom_bytes(int P) { return geom_layout(P, nullptr, nullptr); }
size_t g4s_image_bytes(int W, int H) { return image_layout(W, H, nullptr, nullptr); }
size_t g4s_binning_bytes(int64_t capacity) { return bin_layout(capacity, nullptr, nullptr); }
size_t g4s_backward_scratch_bytes(int P) { return align_up((size_t)(P > 0 ? P : 1) * ACC_FLOATS * sizeof(float), 256); }
size_t g4s_backward_scratch_bytes_raw(int P) { return g4s_backward_scratch_bytes(P); }

}  // extern "C"

// raw != 0: shs = _features_dc, sh_rest = _features_rest, opacities / scales / rotations before activation
static int forward_plan_impl(int P, int D, int M, int W, int H, const float* means3D, const float* shs,
                             const float* colors_precomp, const float* opacities, const float* scales,
                             float scale_modifier, const float* rotations, const float* transMat_precomp,
                             const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                             int prefiltered, int* radii, void* geom_buffer,
                             void* img_buffer, int32_t* host_counts, void* stream, int debug,
                             int raw, const float* sh_rest, const float* mip_filter) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0) return fail(G4S_EINVAL, "g4s_forward_plan: bad P/W/H");
    if (!img_buffer || !geom_buffer) return fail(G4S_EINVAL, "g4s_forward_plan: null scratch buffer");
    if (P > 0) {
        if (!means3D || !opacities || !viewmatrix || !projmatrix || !radii)
            return fail(G4S_EINVAL, "g4s_forward_plan: null required input");
        if ((shs == nullptr) == (colors_precomp == nullptr))
            return fail(G4S_EINVAL, "Please provide excatly one of either SHs or precomputed colors!");
        const bool has_sr = scales != nullptr && rotations != nullptr;
        if (has_sr == (transMat_precomp != nullptr) || ((scales != nullptr) != (rotations != nullptr)))
            return fail(G4S_EINVAL, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        if (shs && !cam_pos) return fail(G4S_EINVAL, "g4s_forward_plan: campos required with SHs");
        if (shs && (M < (D + 1) * (D + 1))) return fail(G4S_EINVAL, "g4s_forward_plan: M < (D+1)^2");
        if (D < 0 || D > 3) return fail(G4S_EINVAL, "g4s_forward_plan: sh degree must be 0..3");
        if (shs && M > 16) return fail(G4S_EINVAL, "g4s_forward_plan: at most 16 SH coefficients (degree 3) are supported");
    }
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    if (gx > 65535 || gy > 65535) return fail(G4S_EINVAL, "g4s_forward_plan: image too large");
    GeomView geom;
    ImageView img;
    geom_layout(P, (char*)geom_buffer, &geom);
    image_layout(W, H, (char*)img_buffer, &img);
    const int T = gx * gy;
    int rc;
    if ((rc = check_cuda(cudaMemsetAsync(img.tile_count, 0, sizeof(uint32_t) * T, s), "memset tile_count"))) return rc;
    if ((rc = check_cuda(cudaMemsetAsync(img.counters, 0, sizeof(int32_t) * CNT_N, s), "memset counters"))) return rc;

    ProjectArgs pa;
    pa.P = P; pa.D = D; pa.M = M; pa.W = W; pa.H = H; pa.grid_x = gx; pa.grid_y = gy; pa.prefiltered = prefiltered;
    pa.means3D = means3D; pa.shs = shs; pa.colors_precomp = colors_precomp; pa.opacities = opacities;
    pa.scales = scales; pa.scale_modifier = scale_modifier; pa.rotations = rotations; pa.transMat_precomp = transMat_precomp;
    pa.view = viewmatrix; pa.proj = projmatrix; pa.campos = cam_pos;
    pa.radii = radii; pa.geom = geom; pa.tile_count = img.tile_count; pa.counters = img.counters;
    pa.raw = raw; pa.sh_rest = sh_rest; pa.mip_filter = mip_filter;
    { StageTimer tm(ST_PROJECT_FWD, s); launch_project_fwd(pa, s); }
    if ((rc = stage_check(debug, s, "project_fwd"))) return rc;

    TileScanArgs ta;
    ta.num_tiles = T; ta.tile_count = img.tile_count; ta.tile_offset = img.tile_offset;
    ta.tile_order = img.tile_order; ta.counters = img.counters;
    { StageTimer tm(ST_TILE_SCAN, s); launch_tile_scan(ta, s); }
    if ((rc = stage_check(debug, s, "tile_scan"))) return rc;

    if (host_counts) {
        if ((rc = check_cuda(cudaMemcpyAsync(host_counts, img.counters, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s),
                             "copy counters")))
            return rc;
    }
    if (debug && host_counts) {
        if ((rc = check_cuda(cudaStreamSynchronize(s), "copy counters"))) return rc;
        if (prefiltered && host_counts[3] != 0)
            return fail(G4S_ECUDA, "Point is filtered although prefiltered is set. This shouldn't happen!");
    }
    return G4S_OK;
}

extern "C" {

int g4s_forward_plan(int P, int D, int M, int W, int H, const float* means3D, const float* shs,
                     const float* colors_precomp, const float* opacities, const float* scales,
                     float scale_modifier, const float* rotations, const float* transMat_precomp,
                     const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                     float tan_fovx, float tan_fovy, int prefiltered, int* radii, void* geom_buffer,
                     void* img_buffer, int32_t* host_counts, void* stream, int debug) {
    (void)tan_fovx; (void)tan_fovy;
    return forward_plan_impl(P, D, M, W, H, means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
                             transMat_precomp, viewmatrix, projmatrix, cam_pos, prefiltered, radii, geom_buffer,
                             img_buffer, host_counts, stream, debug, 0, nullptr, nullptr);
}

int g4s_forward_plan_raw(int P, int D, int M, int W, int H, const float* xyz, const float* features_dc,
                         const float* features_rest, const float* opacity_raw, const float* scaling_raw,
                         float scale_modifier, const float* rotation_raw, const float* mip_filter,
                         const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                         float tan_fovx, float tan_fovy, int prefiltered, int* radii, void* geom_buffer,
                         void* img_buffer, int32_t* host_counts, void* stream, int debug) {
    (void)tan_fovx; (void)tan_fovy;
    if (P > 0) {
        if (!features_dc || !scaling_raw || !rotation_raw) return fail(G4S_EINVAL, "g4s_forward_plan_raw: null parameter tensor");
        if (M < 1 || (M > 1 && !features_rest)) return fail(G4S_EINVAL, "g4s_forward_plan_raw: features_rest required when M > 1");
    }
    return forward_plan_impl(P, D, M, W, H, xyz, features_dc, nullptr, opacity_raw, scaling_raw, scale_modifier, rotation_raw,
                             nullptr, viewmatrix, projmatrix, cam_pos, prefiltered, radii, geom_buffer, img_buffer,
                             host_counts, stream, debug, 1, features_rest, mip_filter);
}

// stage 1 of the render call: scatter + per-tile sort (everything between the plan and the blend)
int g4s_forward_bin(int P, int W, int H, const void* geom_buffer, void* img_buffer, void* binning_buffer, int64_t capacity,
                    void* stream, int debug) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(G4S_EINVAL, "g4s_forward_bin: bad sizes");
    if (!geom_buffer || !img_buffer || !binning_buffer) return fail(G4S_EINVAL, "g4s_forward_bin: null buffer");
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    GeomView geom;
    ImageView img;
    BinView bin;
    geom_layout(P, (char*)geom_buffer, &geom);
    image_layout(W, H, (char*)img_buffer, &img);
    bin_layout(capacity, (char*)binning_buffer, &bin);
    int rc;
    if (debug) {
        int32_t cnt[2];
        if ((rc = check_cuda(cudaMemcpyAsync(cnt, img.counters, sizeof(cnt), cudaMemcpyDeviceToHost, s), "read counters"))) return rc;
        if ((rc = check_cuda(cudaStreamSynchronize(s), "read counters"))) return rc;
        if ((int64_t)cnt[0] > capacity) return fail(G4S_ECAPACITY, "g4s_forward_render: num_rendered exceeds capacity");
    }
    ScatterArgs sa;
    sa.P = P; sa.grid_x = gx; sa.capacity = capacity; sa.geom = geom; sa.tile_offset = img.tile_offset;
    sa.tile_cursor = img.tile_count; sa.keys = bin.keys; sa.counters = img.counters;
    { StageTimer tm(ST_SCATTER, s); launch_scatter(sa, s); }
    if ((rc = stage_check(debug, s, "scatter"))) return rc;

    TileSortArgs ts;
    ts.num_tiles = gx * gy; ts.capacity = capacity; ts.tile_offset = img.tile_offset; ts.tile_order = img.tile_order;
    ts.keys = bin.keys; ts.list = bin.list; ts.counters = img.counters;
    { StageTimer tm(ST_TILE_SORT, s); launch_tile_sort(ts, s); }
    return stage_check(debug, s, "tile_sort");
}

// stage 2 of the render call: the blend
int g4s_forward_blend(int P, int W, int H, const float* background, const void* geom_buffer, void* img_buffer,
                      void* binning_buffer, int64_t capacity, float* out_color, float* out_others, void* stream, int debug) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(G4S_EINVAL, "g4s_forward_blend: bad sizes");
    if (!geom_buffer || !img_buffer || !binning_buffer || !out_color || !out_others || !background)
        return fail(G4S_EINVAL, "g4s_forward_blend: null buffer");
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    GeomView geom;
    ImageView img;
    BinView bin;
    geom_layout(P, (char*)geom_buffer, &geom);
    image_layout(W, H, (char*)img_buffer, &img);
    bin_layout(capacity, (char*)binning_buffer, &bin);
    BlendFwdArgs ba;
    ba.W = W; ba.H = H; ba.grid_x = gx; ba.grid_y = gy; ba.capacity = capacity;
    ba.tile_offset = img.tile_offset; ba.tile_order = img.tile_order; ba.list = bin.list; ba.masks = bin.masks; ba.rec = geom.rec;
    ba.bg = background; ba.final_T = img.final_T; ba.n_contrib = img.n_contrib;
    ba.out_color = out_color; ba.out_others = out_others; ba.counters = img.counters;
    ba.fast_math = g_fast_math.load(std::memory_order_relaxed);
    { StageTimer tm(ST_BLEND_FWD, s); launch_blend_fwd(ba, s); }
    return stage_check(debug, s, "blend_fwd");
}

int g4s_forward_render(int P, int W, int H, const float* background, const void* geom_buffer,
                       void* img_buffer, void* binning_buffer, int64_t capacity, float* out_color,
                       float* out_others, void* stream, int debug) {
    if (!out_color || !out_others || !background) return fail(G4S_EINVAL, "g4s_forward_render: null buffer");
    int rc = g4s_forward_bin(P, W, H, geom_buffer, img_buffer, binning_buffer, capacity, stream, debug);
    if (rc != G4S_OK) return rc;
    return g4s_forward_blend(P, W, H, background, geom_buffer, img_buffer, binning_buffer, capacity, out_color, out_others, stream, debug);
}

}  // extern "C"

static int backward_impl(int P, int D, int M, int W, int H, const float* background, const float* means3D,
                         const float* shs, const float* scales,
                         const float* rotations, const float* viewmatrix,
                         const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                         const int* radii, const void* geom_buffer, const void* binning_buffer, int64_t capacity,
                         const void* img_buffer, const float* dL_dout_color, const float* dL_dout_others,
                         float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors,
                         float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dtransMat,
                         int accumulate_mask, void* scratch, void* stream, int debug,
                         int raw, const float* sh_rest, const float* opacity_raw, const float* mip_filter, float* dL_dsh_rest) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || W <= 0 || H <= 0 || capacity < 0) return fail(G4S_EINVAL, "g4s_backward: bad sizes");
    if (P == 0) return G4S_OK;
    if (!geom_buffer || !binning_buffer || !img_buffer || !scratch || !dL_dout_color || !dL_dout_others ||
        !dL_dmeans3D || !dL_dmeans2D || !dL_dopacity || !dL_dscales || !dL_drotations ||
        !radii || !means3D || !background)
        return fail(G4S_EINVAL, "g4s_backward: null buffer");
    if (M > 0 && shs && !dL_dsh) return fail(G4S_EINVAL, "g4s_backward: dL_dsh required");
    if (M > 16) return fail(G4S_EINVAL, "g4s_backward: at most 16 SH coefficients (degree 3) are supported");
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    GeomView geom;
    ImageView img;
    BinView bin;
    geom_layout(P, (char*)geom_buffer, &geom);
    image_layout(W, H, (char*)img_buffer, &img);
    bin_layout(capacity, (char*)binning_buffer, &bin);
    int rc;
    float4* acc = (float4*)scratch;
    {
        StageTimer tm(ST_ACC_CLEAR, s);
        launch_acc_clear(P, radii, acc, s);
    }
    if ((rc = stage_check(debug, s, "acc_clear"))) return rc;

    BlendBwdArgs bb;
    bb.W = W; bb.H = H; bb.grid_x = gx; bb.grid_y = gy;
    bb.tile_offset = img.tile_offset; bb.tile_order = img.tile_order;
    bb.list = bin.list; bb.masks = bin.masks;
    bb.rec = geom.rec; bb.bg = background; bb.final_T = img.final_T; bb.n_contrib = img.n_contrib;
    bb.dL_dpix = dL_dout_color; bb.dL_dothers = dL_dout_others; bb.acc = acc;
    bb.counters = img.counters; bb.capacity = capacity;
    { StageTimer tm(ST_BLEND_BWD, s); launch_blend_bwd(bb, s); }
    if ((rc = stage_check(debug, s, "blend_bwd"))) return rc;

    ProjectBwdArgs pb;
    pb.P = P; pb.D = D; pb.M = M; pb.W = W; pb.H = H; pb.accumulate = accumulate_mask; pb.means3D = means3D; pb.shs = shs; pb.scales = scales; pb.rotations = rotations;
    pb.view = viewmatrix; pb.proj = projmatrix; pb.campos = cam_pos;
    pb.focal_y = H / (2.0f * tan_fovy);
    pb.focal_x = W / (2.0f * tan_fovx);
    pb.tan_fovx = tan_fovx; pb.tan_fovy = tan_fovy; pb.radii = radii; pb.geom = geom; pb.acc = acc; pb.counters = img.counters;
    pb.dL_dmeans3D = dL_dmeans3D; pb.dL_dmeans2D = dL_dmeans2D; pb.dL_dsh = (M > 0 && shs) ? dL_dsh : nullptr;
    pb.dL_dcolors = dL_dcolors; pb.dL_dopacity = dL_dopacity; pb.dL_dscales = dL_dscales; pb.dL_drots = dL_drotations;
    pb.dL_dtransMat = dL_dtransMat;
    pb.raw = raw; pb.sh_rest = sh_rest; pb.opacities = opacity_raw; pb.mip_filter = mip_filter; pb.dL_dsh_rest = dL_dsh_rest;
    { StageTimer tm(ST_PROJECT_BWD, s); launch_project_bwd(pb, s); }
    return stage_check(debug, s, "project_bwd");
}

extern "C" {

int g4s_backward(int P, int D, int M, int W, int H, const float* background, const float* means3D,
                 const float* shs, const float* colors_precomp, const float* scales, float scale_modifier,
                 const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                 const int* radii, const void* geom_buffer, const void* binning_buffer, int64_t capacity,
                 const void* img_buffer, const float* dL_dout_color, const float* dL_dout_others,
                 float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh, float* dL_dcolors,
                 float* dL_dopacity, float* dL_dscales, float* dL_drotations, float* dL_dtransMat,
                 int accumulate_mask, void* scratch, void* stream, int debug) {
    (void)scale_modifier; (void)colors_precomp; (void)transMat_precomp;
    if (P > 0 && (!dL_dmeans3D || !dL_dmeans2D || !dL_dopacity || !dL_dscales || !dL_drotations))
        return fail(G4S_EINVAL, "g4s_backward: null buffer");
    if (P > 0 && ((colors_precomp && !dL_dcolors) || (transMat_precomp && !dL_dtransMat)))
        return fail(G4S_EINVAL, "g4s_backward: a precomputed input needs its gradient buffer");
    return backward_impl(P, D, M, W, H, background, means3D, shs, scales, rotations, viewmatrix, projmatrix, cam_pos,
                         tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, capacity, img_buffer, dL_dout_color,
                         dL_dout_others, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, dL_dscales,
                         dL_drotations, dL_dtransMat, accumulate_mask, scratch, stream, debug, 0, nullptr, nullptr, nullptr, nullptr);
}

int g4s_backward_raw(int P, int D, int M, int W, int H, const float* background, const float* xyz,
                     const float* features_dc, const float* features_rest, const float* opacity_raw,
                     const float* scaling_raw, float scale_modifier, const float* rotation_raw,
                     const float* mip_filter, const float* viewmatrix, const float* projmatrix,
                     const float* cam_pos, float tan_fovx, float tan_fovy, const int* radii,
                     const void* geom_buffer, const void* binning_buffer, int64_t capacity, const void* img_buffer,
                     const float* dL_dout_color, const float* dL_dout_others, float* dL_dxyz, float* dL_dmeans2D,
                     float* dL_dfeatures_dc, float* dL_dfeatures_rest, float* dL_dopacity_raw, float* dL_dscaling_raw,
                     float* dL_drotation_raw, int accumulate_mask, void* scratch, void* stream, int debug) {
    (void)scale_modifier;
    if (P > 0 && (!features_dc || !opacity_raw || !scaling_raw || !rotation_raw || !dL_dxyz || !dL_dmeans2D || !dL_dfeatures_dc ||
                  !dL_dopacity_raw || !dL_dscaling_raw || !dL_drotation_raw || (M > 1 && (!features_rest || !dL_dfeatures_rest))))
        return fail(G4S_EINVAL, "g4s_backward_raw: null buffer");
    // the operator-only outputs (dL_dcolors_precomp, dL_dtransMat) have no raw counterpart: not written
    return backward_impl(P, D, M, W, H, background, xyz, features_dc, scaling_raw, rotation_raw, viewmatrix, projmatrix, cam_pos,
                         tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, capacity, img_buffer, dL_dout_color,
                         dL_dout_others, dL_dxyz, dL_dmeans2D, dL_dfeatures_dc, nullptr, dL_dopacity_raw, dL_dscaling_raw,
                         dL_drotation_raw, nullptr, accumulate_mask, scratch, stream, debug,
                         1, features_rest, opacity_raw, mip_filter, dL_dfeatures_rest);
}

int g4s_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
    (void)projmatrix;
    if (P < 0) return fail(G4S_EINVAL, "g4s_mark_visible: bad P");
    if (P == 0) return G4S_OK;
    if (!means3D || !viewmatrix || !present) return fail(G4S_EINVAL, "g4s_mark_visible: null buffer");
    launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "mark_visible");
}

static int densify_stats_impl(int P, const float* dL_dmeans2D, const int* radii, float* accum, float* denom,
                              int* max_radii, int multimem, void* stream) {
    if (P < 0) return fail(G4S_EINVAL, "g4s_densify_stats: bad P");
    if (P == 0) return G4S_OK;
    if (!dL_dmeans2D || !radii || !accum || !denom || !max_radii) return fail(G4S_EINVAL, "g4s_densify_stats: null buffer");
    launch_densify_stats(P, dL_dmeans2D, radii, accum, denom, max_radii, multimem, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "densify_stats");
}
int g4s_densify_stats(int P, const float* dL_dmeans2D, const int* radii, float* accum, float* denom,
                      int* max_radii, void* stream) {
    return densify_stats_impl(P, dL_dmeans2D, radii, accum, denom, max_radii, 0, stream);
}
int g4s_densify_stats_multimem(int P, const float* dL_dmeans2D, const int* radii, float* accum_mc, float* denom_mc,
                               int* max_radii_mc, void* stream) {
    return densify_stats_impl(P, dL_dmeans2D, radii, accum_mc, denom_mc, max_radii_mc, 1, stream);
}

int g4s_densify_classify(int P, const float* accum, const float* denom, const float* scaling_raw, const float* opacity_raw,
                         float grad_threshold, float dense_extent, float min_opacity, float big_world_size, int n_split,
                         uint8_t* flags, void* stream) {
    if (P < 0 || n_split < 1) return fail(G4S_EINVAL, "g4s_densify_classify: bad P / n_split");
    if (P == 0) return G4S_OK;
    if (!accum || !denom || !scaling_raw || !opacity_raw || !flags) return fail(G4S_EINVAL, "g4s_densify_classify: null buffer");
    launch_densify_classify(P, accum, denom, scaling_raw, opacity_raw, grad_threshold, dense_extent, min_opacity, big_world_size,
                            1.0f / (0.8f * (float)n_split), flags, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "densify_classify");
}
int g4s_densify_gather(int P_new, int rest_width, const int* src_row, const uint8_t* kind, const int* sample_row,
                       const float* samples, int n_split, const float* const* src_tensors, float* const* dst_tensors, void* stream) {
    if (P_new < 0 || rest_width < 0 || n_split < 1) return fail(G4S_EINVAL, "g4s_densify_gather: bad sizes");
    if (P_new == 0) return G4S_OK;
    if (!src_row || !kind || !sample_row || !src_tensors || !dst_tensors) return fail(G4S_EINVAL, "g4s_densify_gather: null buffer");
    for (int i = 0; i < 18; i += 3)
        if ((!src_tensors[i] || !dst_tensors[i]) && !(i == 6 && rest_width == 0))
            return fail(G4S_EINVAL, "g4s_densify_gather: null parameter tensor");
    launch_densify_gather(P_new, rest_width, src_row, kind, sample_row, samples, 1.0f / (0.8f * (float)n_split), src_tensors, dst_tensors,
                          (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "densify_gather");
}

int g4s_multimem_allreduce(float* sum_mc, int64_t n_floats, int* max_mc, int64_t n_ints, int rank, int world, void* stream) {
    if (n_floats < 0 || n_ints < 0 || world < 1 || rank < 0 || rank >= world || (n_floats & 3))
        return fail(G4S_EINVAL, "g4s_multimem_allreduce: bad sizes (n_floats must be a multiple of 4)");
    if (n_floats == 0 && n_ints == 0) return G4S_OK;
    if ((n_floats && !sum_mc) || (n_ints && !max_mc)) return fail(G4S_EINVAL, "g4s_multimem_allreduce: null buffer");
    if ((reinterpret_cast<uintptr_t>(sum_mc) & 15) != 0) return fail(G4S_EINVAL, "g4s_multimem_allreduce: sum_mc must be 16-byte aligned");
    launch_multimem_allreduce(sum_mc, (size_t)n_floats, max_mc, (size_t)n_ints, rank, world, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "multimem_allreduce");
}

int g4s_photometric_forward(int W, int H, int C, const float* image, const float* gt, const float* window11,
                            float lambda_dssim, double* sums, float* dmaps, float* out3, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0 || C <= 0 || C > 65535) return fail(G4S_EINVAL, "g4s_photometric_forward: bad W/H/C");
    if (!image || !gt || !window11 || !sums || !out3) return fail(G4S_EINVAL, "g4s_photometric_forward: null buffer");
    int rc;
    if ((rc = check_cuda(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s), "memset loss sums"))) return rc;
    launch_photometric_fwd(W, H, C, image, gt, window11, lambda_dssim, sums, dmaps, out3, s);
    return stage_check(false, s, "photometric_fwd");
}

int g4s_photometric_backward(int W, int H, int C, const float* image, const float* gt, const float* window11,
                             float lambda_dssim, const float* dmaps, const float* dL_dloss, float* dL_dimage, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0 || C <= 0 || C > 65535) return fail(G4S_EINVAL, "g4s_photometric_backward: bad W/H/C");
    if (!image || !gt || !window11 || !dmaps || !dL_dimage) return fail(G4S_EINVAL, "g4s_photometric_backward: null buffer");
    launch_photometric_bwd(W, H, C, image, gt, window11, lambda_dssim, dmaps, dL_dloss, dL_dimage, s);
    return stage_check(false, s, "photometric_bwd");
}

int g4s_mip_filter(int P, const float* xyz, int num_cameras, const float* cameras, float znear, float focal_length,
                   float sqrt_filter_variance, float* mip_filter, uint32_t* max_distance_bits, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P < 0 || num_cameras < 0) return fail(G4S_EINVAL, "g4s_mip_filter: bad P / num_cameras");
    if (P == 0) return G4S_OK;
    if (!xyz || !mip_filter || !max_distance_bits || (num_cameras > 0 && !cameras))
        return fail(G4S_EINVAL, "g4s_mip_filter: null buffer");
    if (!(focal_length > 0.0f)) return fail(G4S_EINVAL, "g4s_mip_filter: focal_length must be positive");
    int rc;
    if ((rc = check_cuda(cudaMemsetAsync(max_distance_bits, 0, sizeof(uint32_t), s), "memset max distance"))) return rc;
    launch_mip_filter(P, xyz, num_cameras, cameras, znear, focal_length, sqrt_filter_variance, mip_filter,
                      max_distance_bits, s);
    return stage_check(false, s, "mip_filter");
}

int g4s_surface_forward(int W, int H, const float* allmap, const float* viewmatrix, const float* projmatrix,
                        double depth_ratio, float* rend_alpha, float* rend_normal, float* rend_normal_cam, float* rend_dist,
                        float* surf_depth, float* surf_normal, float* surf_normal_cam, float* rend_depth, void* stream) {
    if (W < 0 || H < 0) return fail(G4S_EINVAL, "surface_forward: negative image size");
    if (W == 0 || H == 0) return G4S_OK;
    if (!allmap || !viewmatrix || !projmatrix || !rend_alpha || !rend_normal || !rend_normal_cam || !rend_dist ||
        !surf_depth || !surf_normal || !surf_normal_cam || !rend_depth)
        return fail(G4S_EINVAL, "surface_forward: null pointer");
    SurfaceFwdArgs a{W, H, (float)(1.0 - depth_ratio), (float)depth_ratio, allmap, viewmatrix, projmatrix,
                     rend_alpha, rend_normal, rend_normal_cam, rend_dist, surf_depth, surf_normal, surf_normal_cam, rend_depth};
    launch_surface_fwd(a, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "surface_fwd");
}

int g4s_surface_backward(int W, int H, const float* allmap, const float* viewmatrix, const float* projmatrix,
                         double depth_ratio, const float* dL_drend_alpha, const float* dL_drend_normal,
                         const float* dL_drend_normal_cam, const float* dL_drend_dist, const float* dL_dsurf_depth,
                         const float* dL_dsurf_normal, const float* dL_dsurf_normal_cam, const float* dL_drend_depth,
                         float* dL_dallmap, void* stream) {
    if (W < 0 || H < 0) return fail(G4S_EINVAL, "surface_backward: negative image size");
    if (W == 0 || H == 0) return G4S_OK;
    if (!allmap || !viewmatrix || !projmatrix || !dL_dallmap) return fail(G4S_EINVAL, "surface_backward: null pointer");
    SurfaceBwdArgs a{W, H, (float)(1.0 - depth_ratio), (float)depth_ratio, allmap, viewmatrix, projmatrix,
                     dL_drend_alpha, dL_drend_normal, dL_drend_normal_cam, dL_drend_dist, dL_dsurf_depth,
                     dL_dsurf_normal, dL_dsurf_normal_cam, dL_drend_depth, dL_dallmap};
    launch_surface_bwd(a, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "surface_bwd");
}

int g4s_normal2curv_forward(int W, int H, const float* normal, const float* mask, float* curv, float* sign_map, void* stream) {
    if (W < 0 || H < 0 || H > 65535) return fail(G4S_EINVAL, "g4s_normal2curv_forward: bad image size");
    if (W == 0 || H == 0) return G4S_OK;
    if (!normal || !curv) return fail(G4S_EINVAL, "g4s_normal2curv_forward: null buffer");
    launch_normal2curv_fwd(W, H, normal, mask, curv, sign_map, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "normal2curv_fwd");
}
int g4s_normal2curv_backward(int W, int H, const float* mask, const float* sign_map, const float* dL_dcurv, float* dL_dnormal,
                             void* stream) {
    if (W < 0 || H < 0 || H > 65535) return fail(G4S_EINVAL, "g4s_normal2curv_backward: bad image size");
    if (W == 0 || H == 0) return G4S_OK;
    if (!sign_map || !dL_dcurv || !dL_dnormal) return fail(G4S_EINVAL, "g4s_normal2curv_backward: null buffer");
    launch_normal2curv_bwd(W, H, mask, sign_map, dL_dcurv, dL_dnormal, (cudaStream_t)stream);
    return stage_check(false, (cudaStream_t)stream, "normal2curv_bwd");
}
int g4s_depth_order_forward(int W, int H, const float* depth, const float* prior_depth, const int64_t* pixel_shifts,
                            float scene_extent, int normalize_loss, int log_space, float log_scale, float* per_pixel,
                            double* sum, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0 || (int64_t)W * H > 0x7fffffff) return fail(G4S_EINVAL, "g4s_depth_order_forward: bad image size");
    if (!depth || !prior_depth || !pixel_shifts || (!per_pixel && !sum)) return fail(G4S_EINVAL, "g4s_depth_order_forward: null buffer");
    int rc;
    if (sum && (rc = check_cuda(cudaMemsetAsync(sum, 0, sizeof(double), s), "memset depth-order sum"))) return rc;
    launch_depth_order_fwd(W, H, depth, prior_depth, (const long long*)pixel_shifts, 1.0f / scene_extent, normalize_loss, log_space,
                           log_scale, per_pixel, sum, s);
    return stage_check(false, s, "depth_order_fwd");
}
int g4s_depth_order_backward(int W, int H, const float* depth, const float* prior_depth, const int64_t* pixel_shifts,
                             float scene_extent, int normalize_loss, int log_space, float log_scale, const float* dL_dper_pixel,
                             const float* dL_dloss, float scale, float* dL_ddepth, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (W <= 0 || H <= 0 || (int64_t)W * H > 0x7fffffff) return fail(G4S_EINVAL, "g4s_depth_order_backward: bad image size");
    if (!depth || !prior_depth || !pixel_shifts || !dL_ddepth) return fail(G4S_EINVAL, "g4s_depth_order_backward: null buffer");
    int rc;
    if ((rc = check_cuda(cudaMemsetAsync(dL_ddepth, 0, sizeof(float) * (size_t)W * H, s), "memset dL_ddepth"))) return rc;
    launch_depth_order_bwd(W, H, depth, prior_depth, (const long long*)pixel_shifts, 1.0f / scene_extent, normalize_loss, log_space,
                           log_scale, dL_dper_pixel, dL_dloss, scale, dL_ddepth, s);
    return stage_check(false, s, "depth_order_bwd");
}

// ---- introspection ---------------------------------------------------------------------------
namespace g4s {
__global__ void decode_geom_kernel(int P, GeomView geom, float* transMat, float* means2D, float* normal_opacity,
                                   float* rgb, float* depths, float* bbox, uint8_t* clamped, uint32_t* tiles_touched) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4* r = geom.rec + (size_t)i * REC_F4;
    const float4 q0 = r[0], q1 = r[1], q2 = r[2], q3 = r[3], q4 = r[4], q5 = r[5];
    if (transMat) {
        float* t = transMat + 9 * (size_t)i;
        t[0] = q1.x; t[1] = q1.y; t[2] = q1.z; t[3] = q1.w; t[4] = q2.x; t[5] = q2.y; t[6] = q2.z; t[7] = q2.w; t[8] = q3.x;
    }
    if (means2D) { means2D[2 * i] = q3.y; means2D[2 * i + 1] = q3.z; }
    if (normal_opacity) { normal_opacity[4 * i] = q4.x; normal_opacity[4 * i + 1] = q4.y; normal_opacity[4 * i + 2] = q4.z; normal_opacity[4 * i + 3] = q3.w; }
    if (rgb) { rgb[3 * i] = q4.w; rgb[3 * i + 1] = q5.x; rgb[3 * i + 2] = q5.y; }
    if (depths) depths[i] = geom.depth[i];
    (void)q5;
    if (bbox) { bbox[4 * i] = q0.x; bbox[4 * i + 1] = q0.y; bbox[4 * i + 2] = q0.z; bbox[4 * i + 3] = q0.w; }
    if (clamped) { const uint8_t m = geom.clamped[i]; clamped[3 * i] = m & 1; clamped[3 * i + 1] = (m >> 1) & 1; clamped[3 * i + 2] = (m >> 2) & 1; }
    if (tiles_touched) tiles_touched[i] = geom.ntiles[i];
}
__global__ void decode_ranges_kernel(int T, const uint32_t* tile_offset, uint32_t* ranges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    ranges[2 * i] = tile_offset[i];
    ranges[2 * i + 1] = tile_offset[i + 1];
}
// One CTA per tile, one warp per 8x8 block (two 8x4 regions), the same walk as blend_bwd_tall_kernel: an entry is
// replayed when either region's forward mask is non-zero, and then costs 64 pair slots (two pixels per lane).
//   out[0] pairs blended   out[1] sum over pixels of the last contributor's list position   out[2] pair slots (256 per entry)
//   out[3] longest list    out[4] pair slots the backward issues                             out[5] (block, entry) hits
__global__ void pair_stats_kernel(int W, int H, int grid_x, const uint32_t* tile_offset, const uint32_t* list,
                                  const uint32_t* masks, const float4* rec, const uint32_t* n_contrib,
                                  unsigned long long* out) {
    (void)list; (void)rec;
    const int tile = blockIdx.x, ty = tile / grid_x, tx = tile - ty * grid_x;
    const int sub = threadIdx.x >> 5, lane = threadIdx.x & 31;          // 4 warps: 8x8 blocks
    const int wU = 4 * (sub >> 1) + (sub & 1), wL = wU + 2;
    const int px = tx * TILE + (wU & 1) * REGION_W + (lane & 7);
    const int pyU = ty * TILE + (wU >> 1) * REGION_H + (lane >> 3), pyL = pyU + REGION_H;
    const uint32_t off = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - off);
    const int lastU = (px < W && pyU < H) ? (int)n_contrib[(size_t)W * pyU + px] : 0;
    const int lastL = (px < W && pyL < H) ? (int)n_contrib[(size_t)W * pyL + px] : 0;
    unsigned long long walked = (unsigned long long)lastU + lastL, blended = 0, issued = 0, hits = 0;
    int liveU = lastU, liveL = lastL;
    for (int o = 16; o; o >>= 1) {
        liveU = max(liveU, __shfl_xor_sync(~0u, liveU, o));
        liveL = max(liveL, __shfl_xor_sync(~0u, liveL, o));
    }
    liveU = min(liveU, n); liveL = min(liveL, n);
    for (int pos = lane; pos < max(liveU, liveL); pos += 32) {
        const uint32_t mU = pos < liveU ? masks[(size_t)off * 8 + (size_t)wU * n + pos] : 0u;
        const uint32_t mL = pos < liveL ? masks[(size_t)off * 8 + (size_t)wL * n + pos] : 0u;
        blended += __popc(mU) + __popc(mL);
        if (mU | mL) { issued += 64; hits += 1; }
    }
    for (int o = 16; o; o >>= 1) {
        blended += __shfl_xor_sync(~0u, blended, o);
        walked += __shfl_xor_sync(~0u, walked, o);
        issued += __shfl_xor_sync(~0u, issued, o);
        hits += __shfl_xor_sync(~0u, hits, o);
    }
    if (lane == 0) {
        atomicAdd(out + 0, blended);
        atomicAdd(out + 1, walked);
        atomicAdd(out + 4, issued);
        atomicAdd(out + 5, hits);
        if (sub == 0) {
            atomicAdd(out + 2, (unsigned long long)n * (TILE * TILE));
            atomicMax(out + 3, (unsigned long long)n);
        }
    }
}
}  // namespace g4s

int g4s_debug_decode_geom(int P, const void* geom_buffer, float* transMat, float* means2D, float* normal_opacity,
                          float* rgb, float* depths, float* bbox, uint8_t* clamped, uint32_t* tiles_touched, void* stream) {
    if (P <= 0) return G4S_OK;
    GeomView geom;
    geom_layout(P, (char*)geom_buffer, &geom);
    decode_geom_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, geom, transMat, means2D, normal_opacity, rgb,
                                                                          depths, bbox, clamped, tiles_touched);
    return stage_check(false, (cudaStream_t)stream, "decode_geom");
}

int g4s_debug_decode_lists(int W, int H, const void* img_buffer, const void* binning_buffer, int64_t capacity,
                           uint32_t* ranges, float* final_T, uint32_t* n_contrib, uint32_t* point_list, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    ImageView img;
    BinView bin;
    image_layout(W, H, (char*)img_buffer, &img);
    bin_layout(capacity, (char*)binning_buffer, &bin);
    const int T = ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    const size_t N = (size_t)W * H;
    int rc;
    if (ranges) decode_ranges_kernel<<<(T + 255) / 256, 256, 0, s>>>(T, img.tile_offset, ranges);
    if (final_T && (rc = check_cuda(cudaMemcpyAsync(final_T, img.final_T, 3 * N * sizeof(float), cudaMemcpyDeviceToDevice, s), "copy final_T"))) return rc;
    if (n_contrib && (rc = check_cuda(cudaMemcpyAsync(n_contrib, img.n_contrib, 2 * N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s), "copy n_contrib"))) return rc;
    if (point_list && binning_buffer && capacity > 0 &&
        (rc = check_cuda(cudaMemcpyAsync(point_list, bin.list, (size_t)capacity * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s), "copy list"))) return rc;
    return stage_check(false, s, "decode_lists");
}

int g4s_debug_pair_stats(int W, int H, const void* geom_buffer, int P, const void* img_buffer, const void* binning_buffer,
                         int64_t capacity, uint64_t* stats, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    GeomView geom;
    ImageView img;
    BinView bin;
    geom_layout(P, (char*)geom_buffer, &geom);
    image_layout(W, H, (char*)img_buffer, &img);
    bin_layout(capacity, (char*)binning_buffer, &bin);
    const int gx = (W + TILE - 1) / TILE, T = gx * ((H + TILE - 1) / TILE);
    int rc;
    if ((rc = check_cuda(cudaMemsetAsync(stats, 0, 8 * sizeof(uint64_t), s), "clear pair stats"))) return rc;
    pair_stats_kernel<<<T, 128, 0, s>>>(W, H, gx, img.tile_offset, bin.list, bin.masks, geom.rec, img.n_contrib,
                                             (unsigned long long*)stats);
    return stage_check(false, s, "pair_stats");
}

}  // extern "C"
