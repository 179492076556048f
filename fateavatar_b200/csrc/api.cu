// C ABI entry points of libfatesplat.so (see include/fatesplat.h for the boundary this replaces).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

// stage launchers (defined in the other translation units)
void fs_launch_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                          const float* rotations, const float* opacities, const float* shs,
                          const float* cov3D_precomp, const float* colors_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, int W, int H, float tan_fovx, float tan_fovy,
                          int prefiltered, int* radii, char* ws, const fs_workspace_layout& L, cudaStream_t stream);
void fs_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                            cudaStream_t stream);
void fs_launch_binning(int P, int W, int H, char* ws, const fs_workspace_layout& L, fs_frame_info* host_info_dev,
                       cudaStream_t stream);
void fs_launch_blend_forward(int W, int H, const float* bg, float* out_color, char* ws, const fs_workspace_layout& L,
                             cudaStream_t stream);
void fs_launch_backward(int P, int D, int M, const float* bg, int W, int H, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                        const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                        const int* radii, char* ws, const fs_workspace_layout& L, const float* dL_dpix,
                        float* dL_dmean2D, float* dL_dopacity, float* dL_dcolors, float* dL_dmean3D, float* dL_dcov3D,
                        float* dL_dsh, float* dL_dscale, float* dL_drot, cudaStream_t stream);
size_t fs_knn_workspace_bytes_impl(int P);
int fs_launch_knn(int P, const float* points, float* out, char* ws, size_t ws_bytes, cudaStream_t stream);

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local uint32_t g_tile_hint = 0;
static thread_local int g_early_notify = 1;
uint32_t fs_tile_hint() { return g_tile_hint; }

void fs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void fs_count_launch(int n) { g_launches += n; }

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
int fs_tuning(const char* env_name, int default_value) {
    static std::map<std::string, int> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(env_name);
    if (it != cache.end()) return it->second;
    const char* v = getenv(env_name);
    const int val = v ? atoi(v) : default_value;
    cache[env_name] = val;
    return val;
}

// ---- per-stage profiling ----------------------------------------------------------------------------------
// Developer / bench instrumentation (fs_profile_enable): process-wide, guarded by one mutex; events are kept per device.
#include <mutex>
#include <vector>
namespace {
struct ProfRec {
    int stage, device;
    cudaEvent_t start, stop;
};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool[64];
cudaEvent_t prof_event(int dev) {  // g_prof_mu held
    auto& pool = g_event_pool[dev & 63];
    if (!pool.empty()) {
        cudaEvent_t e = pool.back();
        pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

FsStageTimer::FsStageTimer(int stage, cudaStream_t st) : slot(-1), stream(st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    ProfRec r{stage, dev, prof_event(dev), prof_event(dev)};
    cudaEventRecord(r.start, stream);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
FsStageTimer::~FsStageTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].stop, stream);
}

int fs_num_sms() {  // per device: a process may drive more than one GPU
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

bool fs_first_use_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void fs_compute_layout(int P, int W, int H, size_t Rcap, fs_workspace_layout* L) {
    const size_t Tn = (size_t)((W + FS_TILE - 1) / FS_TILE) * ((H + FS_TILE - 1) / FS_TILE);
    const size_t Pn = (size_t)(P > 0 ? P : 0), A = 256;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off = align_up(off + bytes, A);
        return o;
    };
    memset(L, 0, sizeof(*L));
    // header + 256 per-SM slot counters of the forward blend + 16 work counters (FS_WORK_*, common.cuh)
    L->info = take(sizeof(fs_frame_info) + 1024 + 64);
    L->tile_count = take(Tn * 4 * FS_CNT_STRIDE);  // directly after the header: one memset clears both
    L->tile_cursor = take(Tn * 4 * FS_CNT_STRIDE);
    L->ranges = take(Tn * 8);
    L->big_tiles = take((Tn + 1) * 4);
    L->work_order = take(Tn * 4);
    const size_t Smax = Rcap / FS_SEG + Tn + 1;
    L->tile_meta = take(Tn * 16);
    L->seg_base = take((Tn + 1) * 4);
    L->seg_info = take(Smax * 8);
    L->ckpt = take(Smax * FS_TILE_PIX * 16);
    L->final_C = take((size_t)W * H * 16);
    L->depths = take(Pn * 4);
    L->cov3D = take(Pn * 24);
    L->splat = take(Pn * 48);
    L->clamped = take(Pn * 4);
    L->rect = take(Pn * 8);
    L->tiles_touched = take(Pn * 4);
    L->inst_keys = take(Rcap * 8);
    L->inst_keys_alt = L->inst_keys;  // the large-tile path sorts in place; kept for ABI stability
    L->point_list = take(Rcap * 4);
    L->inst_splat = take(Rcap * 48);
    L->final_T = take((size_t)W * H * 4);
    L->n_contrib = take((size_t)W * H * 4);
    L->bwd_counter = take(256 + 1024);  // work counter + 256 per-SM slot counters of the backward blend
    L->grad_acc = take(Pn * 48);
    L->pair_mask = take(Rcap * 32);
    L->instance_capacity = Rcap;
    L->total_bytes = off;
}

#define FS_CUDA_CHECK(call)                                                         \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            fs_set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return FS_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)

extern "C" {

size_t fs_workspace_bytes(int P, int width, int height, size_t instance_capacity) {
    fs_workspace_layout L;
    fs_compute_layout(P, width, height, instance_capacity, &L);
    return L.total_bytes;
}

int fs_get_workspace_layout(int P, int width, int height, size_t instance_capacity, fs_workspace_layout* out) {
    if (!out || P < 0 || width <= 0 || height <= 0) {
        fs_set_error("fs_get_workspace_layout: invalid argument");
        return FS_ERR_INVALID_ARGUMENT;
    }
    fs_compute_layout(P, width, height, instance_capacity, out);
    return FS_OK;
}

int fs_forward(int P, int D, int M, const float* d_background, int width, int height, const float* d_means3D,
               const float* d_shs, const float* d_colors_precomp, const float* d_opacities, const float* d_scales,
               float scale_modifier, const float* d_rotations, const float* d_cov3D_precomp,
               const float* d_viewmatrix, const float* d_projmatrix, const float* d_cam_pos, float tan_fovx,
               float tan_fovy, int prefiltered, float* d_out_color, int* d_radii, void* d_workspace,
               size_t workspace_bytes, size_t instance_capacity, fs_frame_info* h_info, void* stream) {
    g_launches = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0 || D < 0 || D > 3) {
        fs_set_error("fs_forward: invalid size (P=%d W=%d H=%d D=%d)", P, width, height, D);
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) {  // reference: nothing is launched, out_color keeps its initial zeros (rasterize_points.cu:81)
        if (h_info) memset(h_info, 0, sizeof(*h_info));
        return FS_OK;
    }
    if (!d_means3D || !d_opacities || !d_background || !d_viewmatrix || !d_projmatrix || !d_cam_pos || !d_out_color ||
        !d_radii || !d_workspace) {
        fs_set_error("fs_forward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_colors_precomp && (!d_shs || M <= 0 || (D + 1) * (D + 1) > M)) {
        // mirrors "For non-RGB, provide precomputed Gaussian colors!" class of errors (rasterizer_impl.cu:242-245)
        fs_set_error("fs_forward: need colors_precomp or SH coefficients with M >= (D+1)^2 (M=%d D=%d)", M, D);
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_cov3D_precomp && (!d_scales || !d_rotations)) {
        fs_set_error("fs_forward: need cov3D_precomp or scales+rotations");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if ((size_t)((width + 15) / 16) > 65535 || (size_t)((height + 15) / 16) > 65535 ||
        instance_capacity > 0xfffffff0ull) {
        fs_set_error("fs_forward: image or instance capacity too large");
        return FS_ERR_UNSUPPORTED;
    }
    fs_workspace_layout L;
    fs_compute_layout(P, width, height, instance_capacity, &L);
    if (workspace_bytes < L.total_bytes) {
        fs_set_error("fs_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total_bytes);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0 ||
        (d_rotations && (reinterpret_cast<uintptr_t>(d_rotations) & 15) != 0)) {
        fs_set_error("fs_forward: workspace must be 256-byte aligned and rotations 16-byte aligned");
        return FS_ERR_INVALID_ARGUMENT;
    }
    char* ws = static_cast<char*>(d_workspace);
    // header + per-tile histogram are contiguous: one memset node
    FS_CUDA_CHECK(cudaMemsetAsync(ws + L.info, 0, L.tile_cursor - L.info, st));
    fs_launch_preprocess(P, D, M, d_means3D, d_scales, scale_modifier, d_rotations, d_opacities, d_shs,
                         d_cov3D_precomp, d_colors_precomp, d_viewmatrix, d_projmatrix, d_cam_pos, width, height,
                         tan_fovx, tan_fovy, prefiltered, d_radii, ws, L, st);
    // If h_info is pinned memory the device can address, the scan kernel also stores R / overflow there directly
    // (early notification); the full header is still copied at the end of the frame.
    fs_frame_info* h_info_dev = nullptr;
    if (h_info && g_early_notify) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, h_info) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer != nullptr)
            h_info_dev = static_cast<fs_frame_info*>(attr.devicePointer);
        else
            cudaGetLastError();  // pageable memory: not an error, just no early notification
    }
    fs_launch_binning(P, width, height, ws, L, h_info_dev, st);
    fs_launch_blend_forward(width, height, d_background, d_out_color, ws, L, st);
    if (h_info)
        FS_CUDA_CHECK(cudaMemcpyAsync(h_info, ws + L.info, sizeof(fs_frame_info), cudaMemcpyDeviceToHost, st));
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int fs_backward(int P, int D, int M, const float* d_background, int width, int height, const float* d_means3D,
                const float* d_shs, const float* d_colors_precomp, const float* d_scales, float scale_modifier,
                const float* d_rotations, const float* d_cov3D_precomp, const float* d_viewmatrix,
                const float* d_projmatrix, const float* d_cam_pos, float tan_fovx, float tan_fovy,
                const int* d_radii, void* d_workspace, size_t workspace_bytes, size_t instance_capacity,
                const float* d_dL_dpix, float* d_dL_dmean2D, float* d_dL_dopacity, float* d_dL_dcolors,
                float* d_dL_dmean3D, float* d_dL_dcov3D, float* d_dL_dsh, float* d_dL_dscale, float* d_dL_drot,
                void* stream) {
    g_launches = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P < 0 || width <= 0 || height <= 0) {
        fs_set_error("fs_backward: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    if (!d_means3D || !d_background || !d_viewmatrix || !d_projmatrix || !d_cam_pos || !d_radii || !d_workspace ||
        !d_dL_dpix || !d_dL_dmean2D || !d_dL_dopacity || !d_dL_dcolors || !d_dL_dmean3D || !d_dL_dcov3D ||
        (M > 0 && !d_dL_dsh) || !d_dL_dscale || !d_dL_drot) {
        fs_set_error("fs_backward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    fs_workspace_layout L;
    fs_compute_layout(P, width, height, instance_capacity, &L);
    if (workspace_bytes < L.total_bytes) {
        fs_set_error("fs_backward: workspace too small (%zu < %zu bytes)", workspace_bytes, L.total_bytes);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    fs_launch_backward(P, D, M, d_background, width, height, d_means3D, d_shs, d_colors_precomp, d_scales,
                       scale_modifier, d_rotations, d_cov3D_precomp, d_viewmatrix, d_projmatrix, d_cam_pos, tan_fovx,
                       tan_fovy, d_radii, static_cast<char*>(d_workspace), L, d_dL_dpix, d_dL_dmean2D, d_dL_dopacity,
                       d_dL_dcolors, d_dL_dmean3D, d_dL_dcov3D, d_dL_dsh, d_dL_dscale, d_dL_drot, st);
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

int fs_mark_visible(int P, const float* d_means3D, const float* d_viewmatrix, const float* d_projmatrix,
                    uint8_t* d_present, void* stream) {
    (void)d_projmatrix;
    g_launches = 0;
    if (P < 0) {
        fs_set_error("fs_mark_visible: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    if (!d_means3D || !d_viewmatrix || !d_present) {
        fs_set_error("fs_mark_visible: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    fs_launch_mark_visible(P, d_means3D, d_viewmatrix, d_present, static_cast<cudaStream_t>(stream));
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

size_t fs_knn_workspace_bytes(int P) { return fs_knn_workspace_bytes_impl(P); }

int fs_knn_mean_dist2(int P, const float* d_points, float* d_mean_dist2, void* d_workspace, size_t workspace_bytes,
                      void* stream) {
    g_launches = 0;
    if (P < 0) {
        fs_set_error("fs_knn_mean_dist2: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    if (!d_points || !d_mean_dist2 || !d_workspace) {
        fs_set_error("fs_knn_mean_dist2: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    const int rc = fs_launch_knn(P, d_points, d_mean_dist2, static_cast<char*>(d_workspace), workspace_bytes,
                                 static_cast<cudaStream_t>(stream));
    if (rc != FS_OK) return rc;
    FS_CUDA_CHECK(cudaGetLastError());
    return FS_OK;
}

void fs_set_tile_hint(uint32_t max_tile_instances) { g_tile_hint = max_tile_instances; }
void fs_set_early_notify(int on) { g_early_notify = on != 0; }

void fs_profile_enable(int on) { g_prof_on.store(on != 0); }

int fs_profile_read(float* total_ms, int* counts, int n) {
    for (int i = 0; i < n; ++i) {
        total_ms[i] = 0.0f;
        counts[i] = 0;
    }
    std::lock_guard<std::mutex> lock(g_prof_mu);
    for (auto& r : g_prof) {
        if (cudaEventSynchronize(r.stop) != cudaSuccess) {
            fs_set_error("fs_profile_read: event sync failed");
            return FS_ERR_CUDA;
        }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        if (r.stage < n) {
            total_ms[r.stage] += ms;
            counts[r.stage] += 1;
        }
        g_event_pool[r.device & 63].push_back(r.start);
        g_event_pool[r.device & 63].push_back(r.stop);
    }
    g_prof.clear();
    return FS_OK;
}

int fs_last_launch_count(void) { return g_launches; }
const char* fs_last_error(void) { return g_err; }
const char* fs_version(void) { return "fatesplat 0.1 (sm_100a)"; }

}  // extern "C"
