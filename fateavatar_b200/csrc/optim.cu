// Optimiser step and splat-set maintenance of the FateAvatar optimise loop, on the device.
//
// Replaces, per step, the two torch.optim.Adam instances over 8 parameter groups of train/optim.py:15-35 (~10 small
// kernels per tensor) by ONE multi-tensor launch, and the optimiser-state surgery of model/fateavatar.py:610-732
// (_uv_densify: torch.cat of every parameter and Adam moment; _prune_low_opacity_points: boolean-mask indexing of 19
// tensors; _reset_opacity) by kernels that work in place on a capacity-allocated SoA, so that the live splat count P
// changes without any reallocation.
#include "common.cuh"

namespace {

constexpr int kAdamThreads = 256, kAdamPerThread = 4, kAdamChunk = kAdamThreads * kAdamPerThread;

struct AdamTable {
    float* p[FS_ADAM_MAX_TENSORS];
    const float* g[FS_ADAM_MAX_TENSORS];
    float* m[FS_ADAM_MAX_TENSORS];
    float* v[FS_ADAM_MAX_TENSORS];
    unsigned long long n[FS_ADAM_MAX_TENSORS];
    float lr[FS_ADAM_MAX_TENSORS];
    unsigned int first_chunk[FS_ADAM_MAX_TENSORS + 1];
    int count;
};

// torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False), the arithmetic of torch/optim/adam.py
// _single_tensor_adam:  exp_avg.lerp_(grad, 1 - beta1);  exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2);
// denom = exp_avg_sq.sqrt() / sqrt(1 - beta2^t) + eps;  param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1^t)).
// The step counters live in device memory (steps[k], advanced by the last CTA to leave), so the same recorded launch
// is valid for every replay of a CUDA graph.
__global__ void __launch_bounds__(kAdamThreads)
adam_kernel(const AdamTable t, int* __restrict__ steps, double beta1d, double beta2d, float beta2, float w1, float w2,
            float eps) {
    __shared__ float s_step_size, s_bc2_sqrt;
    int k = 0;
    while (k + 1 < t.count && blockIdx.x >= t.first_chunk[k + 1]) ++k;
    if (threadIdx.x == 0) {
        const int step = *reinterpret_cast<volatile int*>(steps + k) + 1;
        const double bc1 = 1.0 - pow(beta1d, (double)step), bc2 = 1.0 - pow(beta2d, (double)step);
        s_step_size = (float)((double)t.lr[k] / bc1);
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
    const size_t base = (size_t)(blockIdx.x - t.first_chunk[k]) * kAdamChunk;
    float* __restrict__ p = t.p[k];
    const float* __restrict__ g = t.g[k];
    float* __restrict__ m = t.m[k];
    float* __restrict__ v = t.v[k];
    const size_t n = t.n[k];
#pragma unroll
    for (int u = 0; u < kAdamPerThread; ++u) {
        const size_t i = base + (size_t)u * kAdamThreads + threadIdx.x;
        if (i < n) {
            const float gi = g[i];
            float mi = m[i], vi = v[i];
            mi = mi + w1 * (gi - mi);
            vi = vi * beta2 + (w2 * gi) * gi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[i] = p[i] + (-step_size) * (mi / denom);
            m[i] = mi;
            v[i] = vi;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int prev = atomicAdd(steps + FS_ADAM_MAX_TENSORS, 1);
        if (prev == (int)gridDim.x - 1) {
            steps[FS_ADAM_MAX_TENSORS] = 0;
            for (int j = 0; j < t.count; ++j) steps[j] += 1;
        }
    }
}

// ---- densification: append `n` children of sampled parents (model/fateavatar.py:610-672) ---------------------------
__global__ void __launch_bounds__(256)
splat_append_kernel(fs_splat_soa s, int P, int n, const long long* __restrict__ parents,
                    const float* __restrict__ new_bary) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t src = (size_t)parents[j], dst = (size_t)P + j;
    s.opacity[dst] = s.opacity[src];
    s.offset[dst] = s.offset[src];
#pragma unroll
    for (int c = 0; c < 3; ++c) s.color[3 * dst + c] = s.color[3 * src + c];
#pragma unroll
    for (int c = 0; c < 4; ++c) s.rotation[4 * dst + c] = s.rotation[4 * src + c];
#pragma unroll
    for (int c = 0; c < 3; ++c) s.scaling[3 * dst + c] = logf(expf(s.scaling[3 * src + c]) * 0.75f);
    const int width[5] = {1, 1, 3, 4, 3};  // opacity, offset, color, rotation, scaling
#pragma unroll
    for (int a = 0; a < 5; ++a)
        for (int c = 0; c < width[a]; ++c) {
            if (s.exp_avg[a]) s.exp_avg[a][width[a] * dst + c] = 0.0f;
            if (s.exp_avg_sq[a]) s.exp_avg_sq[a][width[a] * dst + c] = 0.0f;
        }
    s.face_index[dst] = s.face_index[src];
#pragma unroll
    for (int c = 0; c < 3; ++c) s.bary[3 * dst + c] = new_bary[3 * j + c];
    if (s.sample_flag) s.sample_flag[dst] = 1.0f;
}

__global__ void __launch_bounds__(256)
splat_zero_stats_kernel(fs_splat_soa s, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    s.accum[i] = 0.0f;
    s.denom[i] = 0.0f;
    if (s.max_radii2D) s.max_radii2D[i] = 0.0f;
}

// ---- pruning: stable stream compaction of every per-splat array (model/fateavatar.py:674-713) -----------------------
constexpr int kPruneThreads = 1024;
__device__ __forceinline__ bool keep_splat(float opacity_raw, float min_opacity) {
    const float sig = 1.0f / (1.0f + expf(-opacity_raw));  // torch.sigmoid
    return !(sig < min_opacity);
}

__global__ void __launch_bounds__(kPruneThreads)
prune_count_kernel(const float* __restrict__ opacity, int P, float min_opacity, int* __restrict__ block_counts) {
    const int i = blockIdx.x * kPruneThreads + threadIdx.x;
    const bool keep = i < P && keep_splat(opacity[i], min_opacity);
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
prune_scan_kernel(int* __restrict__ block_counts, int nblocks, int* __restrict__ d_new_P) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int c = i < nblocks ? block_counts[i] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += x;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const int excl = s_carry + s_warp[wid] + incl - c;
        if (i < nblocks) block_counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + c;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d_new_P = s_carry;
}

__global__ void __launch_bounds__(kPruneThreads)
prune_gather_kernel(fs_splat_soa in, fs_splat_soa out, int P, float min_opacity, const int* __restrict__ block_offsets) {
    __shared__ int s_warp[32];
    const int i = blockIdx.x * kPruneThreads + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool keep = i < P && keep_splat(in.opacity[i], min_opacity);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[wid] = __popc(mask);
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += x;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    if (!keep) return;
    const size_t src = (size_t)i, dst = (size_t)block_offsets[blockIdx.x] + s_warp[wid] + __popc(mask & ((1u << lane) - 1u));
    const int width[5] = {1, 1, 3, 4, 3};
    float* const pin[5] = {in.opacity, in.offset, in.color, in.rotation, in.scaling};
    float* const pout[5] = {out.opacity, out.offset, out.color, out.rotation, out.scaling};
#pragma unroll
    for (int a = 0; a < 5; ++a)
        for (int c = 0; c < width[a]; ++c) {
            pout[a][width[a] * dst + c] = pin[a][width[a] * src + c];
            if (in.exp_avg[a]) out.exp_avg[a][width[a] * dst + c] = in.exp_avg[a][width[a] * src + c];
            if (in.exp_avg_sq[a]) out.exp_avg_sq[a][width[a] * dst + c] = in.exp_avg_sq[a][width[a] * src + c];
        }
    out.face_index[dst] = in.face_index[src];
#pragma unroll
    for (int c = 0; c < 3; ++c) out.bary[3 * dst + c] = in.bary[3 * src + c];
    out.accum[dst] = in.accum[src];
    out.denom[dst] = in.denom[src];
    if (in.max_radii2D) out.max_radii2D[dst] = in.max_radii2D[src];
    if (in.sample_flag) out.sample_flag[dst] = in.sample_flag[src];
}

// ---- opacity reset (model/fateavatar.py:715-732) --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
opacity_reset_kernel(float* __restrict__ opacity, float* __restrict__ m, float* __restrict__ v, int P, float cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float sig = 1.0f / (1.0f + expf(-opacity[i]));
    const float y = fminf(sig, cap);
    opacity[i] = logf(y / (1.0f - y));  // tools/gs_utils/general_utils.py: inverse_sigmoid
    if (m) m[i] = 0.0f;
    if (v) v[i] = 0.0f;
}

bool soa_ok(const fs_splat_soa* s) {
    return s && s->opacity && s->offset && s->color && s->rotation && s->scaling && s->face_index && s->bary && s->accum &&
           s->denom;
}

}  // namespace

extern "C" int fs_adam_step(int n_tensors, const fs_adam_tensor* h_tensors, int* d_steps, double beta1, double beta2,
                            double eps, void* stream) {
    if (n_tensors < 1 || n_tensors > FS_ADAM_MAX_TENSORS || !h_tensors || !d_steps) {
        fs_set_error("fs_adam_step: 1..%d tensors, a descriptor array and the device step counters are required",
                     FS_ADAM_MAX_TENSORS);
        return FS_ERR_INVALID_ARGUMENT;
    }
    AdamTable t;
    memset(&t, 0, sizeof(t));
    unsigned int chunks = 0;
    for (int k = 0; k < n_tensors; ++k) {
        const fs_adam_tensor& d = h_tensors[k];
        if (!d.param || !d.grad || !d.exp_avg || !d.exp_avg_sq) {
            fs_set_error("fs_adam_step: tensor %d has a NULL pointer", k);
            return FS_ERR_INVALID_ARGUMENT;
        }
        t.p[k] = d.param;
        t.g[k] = d.grad;
        t.m[k] = d.exp_avg;
        t.v[k] = d.exp_avg_sq;
        t.n[k] = d.n;
        t.lr[k] = d.lr;
        t.first_chunk[k] = chunks;
        chunks += (unsigned int)((d.n + kAdamChunk - 1) / kAdamChunk);
    }
    t.first_chunk[n_tensors] = chunks;
    t.count = n_tensors;
    if (chunks == 0) return FS_OK;
    // scalars are rounded to fp32 the way torch rounds its Python-double scalars: 1 - beta in double first
    adam_kernel<<<chunks, kAdamThreads, 0, static_cast<cudaStream_t>(stream)>>>(t, d_steps, beta1, beta2, (float)beta2,
                                                                                (float)(1.0 - beta1), (float)(1.0 - beta2),
                                                                                (float)eps);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_adam_step: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" int fs_splat_append(const fs_splat_soa* soa, int P, int n, const long long* d_parents, const float* d_new_bary,
                               void* stream) {
    if (!soa_ok(soa) || P < 0 || n < 0 || (n > 0 && (!d_parents || !d_new_bary))) {
        fs_set_error("fs_splat_append: invalid argument");
        return FS_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n > 0) splat_append_kernel<<<(n + 255) / 256, 256, 0, st>>>(*soa, P, n, d_parents, d_new_bary);
    if (P + n > 0) splat_zero_stats_kernel<<<(P + n + 255) / 256, 256, 0, st>>>(*soa, P + n);  // statistics restart
    fs_count_launch(2);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_splat_append: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" size_t fs_splat_prune_workspace_bytes(int P) { return ((size_t)(P + kPruneThreads - 1) / kPruneThreads + 1) * sizeof(int); }

extern "C" int fs_splat_prune(const fs_splat_soa* in, const fs_splat_soa* out, int P, float min_opacity, int* d_new_P,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!soa_ok(in) || !soa_ok(out) || P < 0 || !d_new_P || !d_workspace ||
        workspace_bytes < fs_splat_prune_workspace_bytes(P)) {
        fs_set_error("fs_splat_prune: invalid argument or workspace too small");
        return FS_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nblocks = (P + kPruneThreads - 1) / kPruneThreads;
    int* counts = static_cast<int*>(d_workspace);
    if (nblocks == 0) {
        cudaMemsetAsync(d_new_P, 0, sizeof(int), st);
        return FS_OK;
    }
    prune_count_kernel<<<nblocks, kPruneThreads, 0, st>>>(in->opacity, P, min_opacity, counts);
    prune_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks, d_new_P);
    prune_gather_kernel<<<nblocks, kPruneThreads, 0, st>>>(*in, *out, P, min_opacity, counts);
    fs_count_launch(3);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_splat_prune: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" int fs_opacity_reset(float* d_opacity, float* d_exp_avg, float* d_exp_avg_sq, int P, float cap, void* stream) {
    if (P < 0 || (P > 0 && !d_opacity)) {
        fs_set_error("fs_opacity_reset: invalid argument");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (P == 0) return FS_OK;
    opacity_reset_kernel<<<(P + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_opacity, d_exp_avg, d_exp_avg_sq,
                                                                                         P, cap);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_opacity_reset: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}
