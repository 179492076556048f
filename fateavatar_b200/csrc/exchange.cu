// Peer-memory gradient exchange for the frame-sharded step: the per-step all-reduce of the flat gradient
// bucket done by ONE kernel over NVLink / NVSwitch instead of an NCCL call.
//
// Every rank keeps its bucket in a symmetric (peer-mapped) allocation.  After a cross-rank barrier each rank reads
// the SUM of all N copies:
//   * multicast path (NVSwitch / NVLS): multimem.ld_reduce -- the switch adds the N copies in flight, so a rank pulls
//     `n` floats once regardless of N;
//   * unicast path: plain 128-bit loads from the N peer pointers, summed in rank order (bitwise identical on all ranks).
// The sum lands in a rank-local buffer; the factor records of the FLAME delta gradients travel in the same bucket
// (every rank fills only its own slot, so the sum is the all-gather) and are expanded by fs_flame_expand_grads.
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}

__global__ void __launch_bounds__(512)
p2p_allreduce_multicast_kernel(const float* __restrict__ mc, size_t n4, float4* __restrict__ out) {
    // four switch reductions in flight per thread: the round trip through NVSwitch is long, the requests are cheap
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; ++u) out[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) out[i] = multimem_ld_reduce_add(mc + 4 * i);
}

__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// Two-shot all-reduce through the switch: this rank owns the float4 range [lo, hi) of the bucket, pulls its sum over
// all ranks with one in-switch reduction and pushes the result into EVERY rank's output with one multicast store.
// Per rank n/N floats cross the link in each direction, independent of N (the one-shot kernel moves n).
__global__ void __launch_bounds__(512)
p2p_reduce_scatter_bcast_kernel(const float* __restrict__ mc_in, float* __restrict__ mc_out, size_t lo, size_t hi) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < hi; i += 2 * stride) {
        const float4 a = multimem_ld_reduce_add(mc_in + 4 * i), b = multimem_ld_reduce_add(mc_in + 4 * (i + stride));
        multimem_st(mc_out + 4 * i, a);
        multimem_st(mc_out + 4 * (i + stride), b);
    }
    for (; i < hi; i += stride) multimem_st(mc_out + 4 * i, multimem_ld_reduce_add(mc_in + 4 * i));
}

__global__ void __launch_bounds__(512)
p2p_allreduce_unicast_kernel(int N, const float* const* __restrict__ peers, size_t offset, size_t n4,
                             float4* __restrict__ out) {
    const float4* src[FS_FLAME_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r) src[r] = reinterpret_cast<const float4*>(peers[r < N ? r : 0] + offset);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v[FS_FLAME_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)  // all peer loads in flight together
            if (r < N) v[r] = __ldcv(src[r] + i);     // peers' data changes every step: never a cached copy
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < FS_FLAME_MAX_RANKS; ++r)
            if (r < N) {
                s.x += v[r].x;
                s.y += v[r].y;
                s.z += v[r].z;
                s.w += v[r].w;
            }
        out[i] = s;
    }
}

}  // namespace

extern "C" int fs_p2p_allreduce(int N, const float* d_multicast, const float* const* d_peer_ptrs, size_t offset,
                                size_t n, float* d_out, void* stream) {
    if (N < 1 || N > FS_FLAME_MAX_RANKS || ((n | offset) & 3) != 0 || !d_out || (!d_multicast && !d_peer_ptrs)) {
        fs_set_error("fs_p2p_allreduce: invalid argument (1 <= N <= %d ranks, n and offset multiples of 4, a multicast or peer "
                     "pointer table)", FS_FLAME_MAX_RANKS);
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (n == 0) return FS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n4 = n / 4;
    const int grid = (int)std::min<size_t>((n4 + 511) / 512, (size_t)fs_num_sms() * 2);
    if (d_multicast)
        p2p_allreduce_multicast_kernel<<<grid, 512, 0, st>>>(d_multicast + offset, n4, reinterpret_cast<float4*>(d_out));
    else
        p2p_allreduce_unicast_kernel<<<grid, 512, 0, st>>>(N, d_peer_ptrs, offset, n4, reinterpret_cast<float4*>(d_out));
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_p2p_allreduce: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" int fs_p2p_reduce_scatter_bcast(int N, int rank, const float* d_multicast_in, float* d_multicast_out,
                                           size_t n, void* stream) {
    if (N < 1 || N > FS_FLAME_MAX_RANKS || rank < 0 || rank >= N || (n & 3) != 0 || !d_multicast_in || !d_multicast_out) {
        fs_set_error("fs_p2p_reduce_scatter_bcast: invalid argument (1 <= N <= %d, 0 <= rank < N, n a multiple of 4, "
                     "multicast addresses of the input and output regions)", FS_FLAME_MAX_RANKS);
        return FS_ERR_INVALID_ARGUMENT;
    }
    const size_t n4 = n / 4, per = (n4 + N - 1) / N;
    const size_t lo = std::min(n4, per * (size_t)rank), hi = std::min(n4, lo + per);
    if (hi == lo) return FS_OK;
    const int grid = (int)std::min<size_t>((hi - lo + 511) / 512, (size_t)fs_num_sms());
    p2p_reduce_scatter_bcast_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(d_multicast_in, d_multicast_out,
                                                                                         lo, hi);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_p2p_reduce_scatter_bcast: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}
