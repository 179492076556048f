// Densification statistics: model/fateavatar.py:734-737 (same in volume_rendering/gaussian_model.py:
// 418-420):   xyz_gradient_accum[filter] += ||viewspace.grad[filter, :2]||;   denom[filter] += 1
// Upstream this is boolean-mask indexing (nonzero + gather + norm + index_put, a host sync among them); here it is
// one masked elementwise pass over 16 bytes per splat, no synchronisation.
#include <algorithm>

#include "common.cuh"

namespace {
// torch.norm(grad[:, :2], dim=-1) (model/fateavatar.py:734-737) evaluates sqrt(x0*x0 + x1*x1) with every operation rounded
// on its own (no FMA contraction; measured on the B200: tools/norm_check.py, 0 of 2^20 values differ).  The same sequence
// here makes the statistic -- and with it the multinomial densification draws of _uv_densify -- bit-identical to
// upstream's for the same gradient.
__device__ __forceinline__ float norm2_like_torch(float gx, float gy) {
    return __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
}

__global__ void __launch_bounds__(256)
densify_stats_kernel(int P, const float* __restrict__ grad2d /*[P,3]*/, const uint8_t* __restrict__ filter,
                     float* __restrict__ accum, float* __restrict__ denom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P || !filter[i]) return;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1];
    accum[i] += norm2_like_torch(gx, gy);
    denom[i] += 1.0f;
}
// Frame-sharded form: this frame's INCREMENTS written out of place (0 where the splat is not visible), so that they
// can sit in the gradient bucket and be summed over ranks; `radii` (int32, fs_forward's output) is the filter.
__global__ void __launch_bounds__(256)
densify_stats_inc_kernel(int P, const float* __restrict__ grad2d, const int* __restrict__ radii,
                         float* __restrict__ accum_inc, float* __restrict__ denom_inc) {
    fs::pdl_wait();  // launched while the per-Gaussian backward drains (fs_launch_pdl)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = radii[i] > 0;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1];
    accum_inc[i] = vis ? norm2_like_torch(gx, gy) : 0.0f;
    denom_inc[i] = vis ? 1.0f : 0.0f;
}
}  // namespace

// ---- per-frame camera ---------------------------------------------------------------------------------
// volume_rendering/camera_3dgs.py:53-72 builds the view matrix through two CPU 4x4 inverses, a host round trip and a GPU
// inverse per frame; in closed form  world_view = [[R, 0], [T, 1]],  full = world_view * P^T,  centre = -R T  it is 35
// floats of arithmetic: one warp, no host involvement, CUDA-graph capturable.
static __global__ void frame_camera_kernel(const float* __restrict__ cam_pose, const float* __restrict__ proj_t,
                                    float* __restrict__ view, float* __restrict__ full, float* __restrict__ campos) {
    __shared__ float v[16];
    const int t = threadIdx.x;
    if (t < 16) {
        const int i = t >> 2, j = t & 3;
        float x;
        if (i < 3) x = j < 3 ? cam_pose[4 * i + j] : 0.0f;        // R (camera-to-world rotation as cam_pose holds it)
        else x = j < 3 ? cam_pose[4 * j + 3] : 1.0f;              // T (world-to-camera translation)
        v[t] = x;
        view[t] = x;
    }
    __syncthreads();
    if (t < 16) {
        const int i = t >> 2, j = t & 3;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(v[4 * i + k], proj_t[4 * k + j], acc);
        full[t] = acc;
    } else if (t < 19) {
        const int i = t - 16;
        campos[i] = -(v[4 * i] * v[12] + v[4 * i + 1] * v[13] + v[4 * i + 2] * v[14]);
    }
}

extern "C" int fs_frame_camera(const float* d_cam_pose, const float* d_projection_t, float* d_world_view,
                               float* d_full_proj, float* d_camera_center, void* stream) {
    if (!d_cam_pose || !d_projection_t || !d_world_view || !d_full_proj || !d_camera_center) {
        fs_set_error("fs_frame_camera: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    frame_camera_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(d_cam_pose, d_projection_t, d_world_view,
                                                                        d_full_proj, d_camera_center);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_frame_camera: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

// ---- fused L1 image loss (training-loop plumbing) -------------------------------------------------------
// train/loss.py:103-105 (rgb_type 'l1'): loss = mean |x - t|.  As torch ops the forward and backward are ~9 launches over
// the 3 MB image (sub, abs, mean, fill, div, sign, mul, ...); here ONE pass writes the loss and d loss / d x =
// sign(x - t) / n.  Deterministic: per-CTA partial sums are added in CTA order by the last CTA to finish.
constexpr int kL1Threads = 256;
static __global__ void __launch_bounds__(kL1Threads)
l1_loss_kernel(size_t n, const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ grad,
               float* __restrict__ partial, unsigned int* __restrict__ counter, float* __restrict__ loss) {
    __shared__ float s_w[kL1Threads / 32];
    __shared__ bool s_last;
    const float inv_n = 1.0f / (float)n;
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * kL1Threads;
    const size_t n4 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(grad)) & 15u) ? 0 : n / 4;
    for (size_t i = (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n4; i += stride) {  // 128-bit body
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i), b = __ldg(reinterpret_cast<const float4*>(t) + i);
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        acc += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
        reinterpret_cast<float4*>(grad)[i] =
            make_float4(d0 > 0.0f ? inv_n : (d0 < 0.0f ? -inv_n : 0.0f), d1 > 0.0f ? inv_n : (d1 < 0.0f ? -inv_n : 0.0f),
                        d2 > 0.0f ? inv_n : (d2 < 0.0f ? -inv_n : 0.0f), d3 > 0.0f ? inv_n : (d3 < 0.0f ? -inv_n : 0.0f));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * kL1Threads + threadIdx.x; i < n; i += stride) {  // tail / unaligned
        const float d = x[i] - t[i];
        acc += fabsf(d);
        grad[i] = d > 0.0f ? inv_n : (d < 0.0f ? -inv_n : 0.0f);
    }
    // fixed-shape tree inside the CTA, per-CTA partials, and the last CTA to finish adds the partials with the same
    // fixed-shape tree: the result does not depend on which CTA finishes last
    auto block_sum = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
        __syncthreads();
        float b = 0.0f;
        if (threadIdx.x == 0)
            for (int w = 0; w < kL1Threads / 32; ++w) b += s_w[w];
        return b;  // valid in thread 0
    };
    const float mine = block_sum(acc);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = mine;
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float v = 0.0f;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += kL1Threads) v += *reinterpret_cast<volatile float*>(partial + b);
        const float total = block_sum(v);
        if (threadIdx.x == 0) {
            *loss = total * inv_n;
            *counter = 0u;
        }
    }
}

extern "C" size_t fs_l1_loss_workspace_bytes(void) { return 1024 * sizeof(float) + 256; }

extern "C" int fs_l1_loss(size_t n, const float* d_x, const float* d_target, float* d_grad, float* d_loss,
                          void* d_workspace, void* stream) {
    if (n == 0 || !d_x || !d_target || !d_grad || !d_loss || !d_workspace) {
        fs_set_error("fs_l1_loss: invalid argument");
        return FS_ERR_INVALID_ARGUMENT;
    }
    // two CTAs per SM: every CTA ends with one atomic on the same ticket word (~10 ns each at the L2), so the tail grows
    // with the CTA count, not with the image
    const int grid = (int)std::min<size_t>((n + kL1Threads * 4 - 1) / (kL1Threads * 4), (size_t)std::min(1024, 2 * fs_num_sms()));
    float* partial = static_cast<float*>(d_workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partial + 1024);  // zero-initialised by the caller once
    l1_loss_kernel<<<grid, kL1Threads, 0, static_cast<cudaStream_t>(stream)>>>(n, d_x, d_target, d_grad, partial, counter,
                                                                               d_loss);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_l1_loss: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

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
    fs_launch_pdl(densify_stats_inc_kernel, dim3((P + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream),
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
