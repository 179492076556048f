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

// =====================================================================================================================
// Fused exchange of one frame-sharded step: cross-rank barrier + all-reduce of the splat-gradient part + expansion of
// the FLAME factor records into the dense delta gradients, in ONE kernel over peer memory (no NCCL call, no separate
// barrier / expand launches, CUDA-graph capturable).
//
//   * barrier: CTA 0 stores this step's epoch into every peer's arrive[rank] slot (st.release.sys); every CTA polls its
//     OWN rank's arrive[0..N) (ld.acquire.sys, local memory) until all N ranks have arrived -- CTAs never wait for each
//     other, only for remote ranks, so no co-residency is required.  The epoch lives in device memory and is advanced
//     by the last CTA to leave, so the same recorded launch works for every replay of a CUDA graph.
//   * splat part [0, n_splat): one-shot (every rank pulls the N copies: unicast 128-bit peer loads summed in rank order,
//     or multimem.ld_reduce through the switch) or two-shot (this rank reduces its 1/N slice in the switch and
//     multicast-stores it into everybody's output region; the last CTA then raises done[rank] on every peer and
//     fs_p2p_wait makes the consumer wait for all N of them).
//   * FLAME part: every rank reads all N factor records (its own slot of each bucket is the only non-zero one, so the
//     in-switch sum IS the gather) -- the small [betas | pose_feature] heads are staged in shared memory once per CTA,
//     the [dL/dv_shaped | dL/dv_posed] tails are read exactly once -- and streams the summed dense gradients
//     (26 MB at V = 5023, L = 400) to local memory while other CTAs are still on the wire.
// Even CTAs do reduce-then-expand, odd CTAs expand-then-reduce, so link traffic and local stores overlap.
namespace {

struct ExchangeArgs {
    int N, rank, algo;                 // algo: 0 one-shot unicast, 1 one-shot multicast, 2 two-shot (multicast)
    const float* const* peers;         // device array: unicast base of every rank's symmetric allocation
    const float* mc;                   // multicast base (algo 1, 2) or nullptr
    float* local;                      // this rank's unicast base
    size_t in_off, n_splat, rec_off, rec_stride, out_off, flags_off;
    float* out;                        // one-shot: rank-local destination of the splat part
    int V, L, l0, NP;
    float scale;
    float *d_dv, *d_ds, *d_dp;         // dense delta gradients (any may be null)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 scale4(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }

// flags (uint32) at local + flags_off:  [0, 8) arrive   [8, 16) done   [16] epoch   [17] CTAs finished
constexpr int kFlagArrive = 0, kFlagDone = 8, kFlagEpoch = 16, kFlagCtas = 17;
constexpr int kExThreads = 512;

__device__ __forceinline__ float4 load_sum4(const ExchangeArgs& a, size_t i4 /* float4 index inside the bucket */) {
    if (a.algo != 0) return multimem_ld_reduce_add(a.mc + a.in_off + 4 * i4);
    float4 v[FS_FLAME_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)
        if (r < a.N) v[r] = __ldcv(reinterpret_cast<const float4*>(a.peers[r] + a.in_off) + i4);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < FS_FLAME_MAX_RANKS; ++r)
        if (r < a.N) {
            s.x += v[r].x;
            s.y += v[r].y;
            s.z += v[r].z;
            s.w += v[r].w;
        }
    return s;
}

__device__ void exchange_reduce_part(const ExchangeArgs& a, int cta, int nctas) {
    const size_t n4 = a.n_splat / 4;
    size_t lo = 0, hi = n4;
    if (a.algo == 2) {  // this rank's slice
        const size_t per = (n4 + a.N - 1) / a.N;
        lo = min(n4, per * (size_t)a.rank);
        hi = min(n4, lo + per);
    }
    const size_t stride = (size_t)nctas * kExThreads;
    size_t i = lo + (size_t)cta * kExThreads + threadIdx.x;
    float4* out_local = reinterpret_cast<float4*>(a.out);
    float* mc_out = const_cast<float*>(a.mc) + a.out_off;
    for (; i + stride < hi; i += 2 * stride) {  // two requests in flight per thread
        const float4 x = scale4(load_sum4(a, i), a.scale), y = scale4(load_sum4(a, i + stride), a.scale);
        if (a.algo == 2) {
            multimem_st(mc_out + 4 * i, x);
            multimem_st(mc_out + 4 * (i + stride), y);
        } else {
            out_local[i] = x;
            out_local[i + stride] = y;
        }
    }
    for (; i < hi; i += stride) {
        const float4 x = scale4(load_sum4(a, i), a.scale);
        if (a.algo == 2) multimem_st(mc_out + 4 * i, x);
        else out_local[i] = x;
    }
}

// factor element `e` of rank r's record (record layout: [betas L | pose_feature NP | dL/dv_shaped 3V | dL/dv_posed 3V])
__device__ __forceinline__ float rec_ld(const ExchangeArgs& a, int r, size_t e) {
    return __ldcv(a.peers[r] + a.in_off + a.rec_off + (size_t)r * a.rec_stride + e);
}

__device__ void exchange_expand_part(const ExchangeArgs& a, int cta, int nctas, float* s_head /* [N][L + NP] */) {
    if (!a.d_dv && !a.d_ds && !a.d_dp) return;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n3 = 3 * a.V, LH = a.L + a.NP;
    // heads of all N records -> shared memory (read straight from each owner's slot: only that slot is non-zero)
    for (int e = t; e < a.N * LH; e += kExThreads) s_head[e] = rec_ld(a, e / LH, (size_t)(e % LH));
    __syncthreads();
    const size_t o_gs = (size_t)LH, o_gp = (size_t)LH + n3;
    if (a.d_dp) {
        for (int e = cta * kExThreads + t; e < n3; e += nctas * kExThreads) {
            float g[FS_FLAME_MAX_RANKS];
#pragma unroll
            for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r) g[r] = r < a.N ? rec_ld(a, r, o_gp + e) * a.scale : 0.0f;
            for (int i = 0; i < a.NP; ++i) {
                float o = 0.f;
#pragma unroll
                for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)
                    if (r < a.N) o += s_head[r * LH + a.L + i] * g[r];
                __stcs(a.d_dp + (size_t)i * n3 + e, o);
            }
        }
    }
    const int nw = nctas * (kExThreads / 32);
    const bool vec = (a.L & 3) == 0 && (a.l0 & 3) == 0 && (LH & 3) == 0;
    for (int row = cta * (kExThreads / 32) + wid; row < n3; row += nw) {
        float gs[FS_FLAME_MAX_RANKS], sum = 0.f;
#pragma unroll
        for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r) {
            gs[r] = r < a.N ? rec_ld(a, r, o_gs + row) * a.scale : 0.0f;
            sum += gs[r];
        }
        if (lane == 0 && a.d_dv) a.d_dv[row] = sum;
        if (!a.d_ds) continue;
        float* out = a.d_ds + (size_t)row * a.L;
        if (vec) {
            for (int c = lane; c < (a.L >> 2); c += 32) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (4 * c >= a.l0) {
#pragma unroll
                    for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)
                        if (r < a.N) {
                            const float4 b = reinterpret_cast<const float4*>(s_head + r * LH)[c];
                            o.x += b.x * gs[r];
                            o.y += b.y * gs[r];
                            o.z += b.z * gs[r];
                            o.w += b.w * gs[r];
                        }
                }
                __stcs(reinterpret_cast<float4*>(out) + c, o);
            }
        } else {
            for (int c = lane; c < a.L; c += 32) {
                float o = 0.f;
                if (c >= a.l0)
#pragma unroll
                    for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)
                        if (r < a.N) o += s_head[r * LH + c] * gs[r];
                __stcs(out + c, o);
            }
        }
    }
}

__global__ void __launch_bounds__(kExThreads)
p2p_exchange_kernel(const ExchangeArgs a) {
    extern __shared__ __align__(16) float s_head[];
    __shared__ uint32_t s_epoch;
    uint32_t* flags = reinterpret_cast<uint32_t*>(a.local + a.flags_off);
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(flags + kFlagEpoch) + 1u;
    __syncthreads();
    const uint32_t epoch = s_epoch;
    if (blockIdx.x == 0 && threadIdx.x < a.N) {
        // everything this rank wrote into its bucket (earlier kernels of this stream) becomes visible to the peers
        __threadfence_system();
        uint32_t* peer_flags = reinterpret_cast<uint32_t*>(const_cast<float*>(a.peers[threadIdx.x]) + a.flags_off);
        st_release_sys(peer_flags + kFlagArrive + a.rank, epoch);
    }
    if (threadIdx.x < a.N)
        while ((int32_t)(ld_acquire_sys(flags + kFlagArrive + threadIdx.x) - epoch) < 0) {
        }
    __syncthreads();
    const int half = gridDim.x / 2, odd = blockIdx.x & 1, idx = blockIdx.x >> 1;
    // even CTAs: wire first; odd CTAs: local stores first (gridDim.x is even)
    if (!odd) {
        exchange_reduce_part(a, idx, half);
        exchange_expand_part(a, blockIdx.x, gridDim.x, s_head);
    } else {
        exchange_expand_part(a, blockIdx.x, gridDim.x, s_head);
        exchange_reduce_part(a, half - 1 - idx, half);  // (mirrored so that the two halves meet in the middle)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this CTA's multicast stores are performed before it counts as finished
        const uint32_t prev = atomicAdd(flags + kFlagCtas, 1u);
        if (prev == gridDim.x - 1) {
            flags[kFlagCtas] = 0u;
            if (a.algo == 2)
                for (int r = 0; r < a.N; ++r) {
                    uint32_t* pf = reinterpret_cast<uint32_t*>(const_cast<float*>(a.peers[r]) + a.flags_off);
                    st_release_sys(pf + kFlagDone + a.rank, epoch);
                }
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(flags + kFlagEpoch) = epoch;
        }
    }
}

// two-shot: the consumer side of the completion handshake (every rank's slice has landed in this rank's output)
__global__ void p2p_wait_kernel(int N, uint32_t* flags) {
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(flags + kFlagEpoch);
    if (threadIdx.x < N)
        while ((int32_t)(ld_acquire_sys(flags + kFlagDone + threadIdx.x) - epoch) < 0) {
        }
}

}  // namespace

extern "C" size_t fs_p2p_exchange_flag_floats(void) { return 32; }

extern "C" int fs_p2p_exchange(int N, int rank, int algo, const float* const* d_peer_ptrs, const float* d_multicast,
                               float* d_local_base, size_t in_offset, size_t n_splat, size_t rec_offset,
                               size_t rec_stride, size_t out_offset, size_t flags_offset, float* d_out, int V, int L,
                               int l0, int NP, float scale, float* d_dL_ddelta_vertex, float* d_dL_ddelta_shapedirs,
                               float* d_dL_ddelta_posedirs, void* stream) {
    const bool flame = d_dL_ddelta_vertex || d_dL_ddelta_shapedirs || d_dL_ddelta_posedirs;
    if (N < 1 || N > FS_FLAME_MAX_RANKS || rank < 0 || rank >= N || algo < 0 || algo > 2 || !d_peer_ptrs || !d_local_base ||
        ((in_offset | n_splat | rec_offset | rec_stride | out_offset | flags_offset) & 3) != 0 ||
        (algo != 0 && !d_multicast) || (algo != 2 && !d_out) ||
        (flame && (V <= 0 || L <= 0 || NP < 0 || l0 < 0 || l0 > L || rec_stride < (size_t)L + NP + 6 * (size_t)V))) {
        fs_set_error("fs_p2p_exchange: invalid argument (1 <= N <= %d, offsets multiples of 4 floats, multicast base for "
                     "algo 1/2, local output for algo 0/1, record stride >= L + NP + 6V)", FS_FLAME_MAX_RANKS);
        return FS_ERR_INVALID_ARGUMENT;
    }
    ExchangeArgs a;
    a.N = N; a.rank = rank; a.algo = algo; a.peers = d_peer_ptrs; a.mc = d_multicast; a.local = d_local_base;
    a.in_off = in_offset; a.n_splat = n_splat; a.rec_off = rec_offset; a.rec_stride = rec_stride; a.out_off = out_offset;
    a.flags_off = flags_offset; a.out = d_out; a.V = V; a.L = L; a.l0 = l0; a.NP = NP; a.scale = scale;
    a.d_dv = d_dL_ddelta_vertex; a.d_ds = d_dL_ddelta_shapedirs; a.d_dp = d_dL_ddelta_posedirs;
    const size_t smem = flame ? (size_t)N * (L + NP) * sizeof(float) : 0;
    if (smem > 200 * 1024) {
        fs_set_error("fs_p2p_exchange: N * (L + NP) floats of record heads do not fit in shared memory");
        return FS_ERR_UNSUPPORTED;
    }
    static std::atomic<unsigned long long> attr_set{0};
    if (fs_first_use_on_device(attr_set))
        cudaFuncSetAttribute(p2p_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int grid = 2 * std::max(1, fs_tuning("FATESPLAT_EXCHANGE_CTAS_PER_SM", 1) * fs_num_sms());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p2p_exchange_kernel<<<grid, kExThreads, smem, st>>>(a);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_p2p_exchange: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

extern "C" int fs_p2p_wait(int N, float* d_local_base, size_t flags_offset, void* stream) {
    if (N < 1 || N > FS_FLAME_MAX_RANKS || !d_local_base || (flags_offset & 3) != 0) {
        fs_set_error("fs_p2p_wait: invalid argument");
        return FS_ERR_INVALID_ARGUMENT;
    }
    p2p_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        N, reinterpret_cast<uint32_t*>(d_local_base + flags_offset));
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_p2p_wait: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}
