// Densification statistics: model/fateavatar.py:734-737 (same in volume_rendering/gaussian_model.py:
// 418-420):   xyz_gradient_accum[filter] += ||viewspace.grad[filter, :2]||;   denom[filter] += 1
// Upstream this is boolean-mask indexing (nonzero + gather + norm + index_put, a host sync among them); here it is
// one masked elementwise pass over 16 bytes per splat, no synchronisation.
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256)
densify_stats_kernel(int P, const float* __restrict__ grad2d /*[P,3]*/, const uint8_t* __restrict__ filter,
                     float* __restrict__ accum, float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || !filter[i]) return;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1];
    accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.0f;
}
// Frame-sharded form: this frame's INCREMENTS written out of place (0 where the splat is not visible), so that they
// can sit in the gradient bucket and be summed over ranks; `radii` (int32, fs_forward's output) is the filter.
__global__ void __launch_bounds__(256)
densify_stats_inc_kernel(int P, const float* __restrict__ grad2d, const int* __restrict__ radii,
                         float* __restrict__ accum_inc, float* __restrict__ denom_inc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = radii[i] > 0;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1];
    accum_inc[i] = vis ? sqrtf(gx * gx + gy * gy) : 0.0f;
    denom_inc[i] = vis ? 1.0f : 0.0f;
}
}  // namespace

extern "C" int fs_densify_stats_inc(int P, const float* d_viewspace_grad, const int* d_radii, float* d_accum_inc,
                                    float* d_denom_inc, void* stream) {
    if (P < 0) {
        fs_set_error("fs_densify_stats_inc: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    if (!d_viewspace_grad || !d_radii || !d_accum_inc || !d_denom_inc) {
        fs_set_error("fs_densify_stats_inc: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    densify_stats_inc_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, d_viewspace_grad, d_radii, d_accum_inc, d_denom_inc);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_densify_stats_inc: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" int fs_densify_stats(int P, const float* d_viewspace_grad, const uint8_t* d_update_filter,
                                float* d_xyz_gradient_accum, float* d_denom, void* stream) {
    if (P < 0) {
        fs_set_error("fs_densify_stats: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    if (!d_viewspace_grad || !d_update_filter || !d_xyz_gradient_accum || !d_denom) {
        fs_set_error("fs_densify_stats: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        P, d_viewspace_grad, d_update_filter, d_xyz_gradient_accum, d_denom);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_densify_stats: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}
