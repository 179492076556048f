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
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    size_t in_off, n_splat, rec_off, rec_stride, out_off, flags_off, gather_off;
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
//                                        [20..21] t(barrier passed)  [24..25] sum ns waiting for peers  [26..27] sum ns
//                                        of work after the barrier  [28] calls   (device-side timing, fs_p2p_exchange_timing)
constexpr int kFlagArrive = 0, kFlagDone = 8, kFlagEpoch = 16, kFlagCtas = 17, kFlagReady = 18, kFlagTBar = 20, kFlagWait = 24,
              kFlagWork = 26, kFlagCalls = 28, kFlagStage = 32, kFlagCta0 = 34;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
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

// One remote round trip per CTA: every load a CTA needs from its peers -- its share of the splat part, the heads of the
// N factor records and the factor rows it will expand -- is issued before anything waits on one of them (NVLink loads
// cost microseconds each; a dependent chain of them was the whole kernel time in the first version).
constexpr int kStageUnroll = 8;

__device__ __forceinline__ void reduce_store(const ExchangeArgs& a, size_t i, float4 x) {
    x = scale4(x, a.scale);
    if (a.algo == 2) multimem_st(const_cast<float*>(a.mc) + a.out_off + 4 * i, x);
    else reinterpret_cast<float4*>(a.out)[i] = x;
}

// Gather once: the N factor records (122 KB each at FLAME size) cross the link ONCE per rank -- 4 CTAs per source rank
// copy them into this rank's local gather area and raise a local counter; every CTA then stages what it needs from
// local memory.  (Letting each of the ~300 CTAs pull the record heads from the peers itself put ~4 MB of redundant,
// same-address traffic on every link and made the remote round trip 50 us at 8 ranks.)
constexpr int kFetchPerRank = 4;

__device__ __forceinline__ void exchange_body(const ExchangeArgs& a, float* smem) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const bool flame = a.d_dv || a.d_ds || a.d_dp;
    const int n3 = 3 * a.V, LH = a.L + a.NP;
    const int rpc = (n3 + gridDim.x - 1) / gridDim.x;             // factor rows expanded by one CTA
    const int row0 = min(n3, (int)blockIdx.x * rpc), nrows = min(rpc, n3 - row0);
    float* s_head = smem;                                          // [N][LH]
    float* s_gs = smem + a.N * LH;                                 // [N][rpc]  dL/dv_shaped rows of this CTA
    float* s_gp = s_gs + a.N * rpc;                                // [N][rpc]  dL/dv_posed rows of this CTA
    uint32_t* flags = reinterpret_cast<uint32_t*>(a.local + a.flags_off);

    // ---- splat part: this CTA's float4s of [lo, hi) ----
    const size_t n4 = a.n_splat / 4;
    size_t lo = 0, hi = n4;
    if (a.algo == 2) {  // two-shot: only this rank's slice
        const size_t per = (n4 + a.N - 1) / a.N;
        lo = min(n4, per * (size_t)a.rank);
        hi = min(n4, lo + per);
    }
    const size_t stride = (size_t)gridDim.x * kExThreads;
    size_t i = lo + (size_t)blockIdx.x * kExThreads + t;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
    const bool h0 = i < hi, h1 = i + stride < hi;
    if (h0) r0 = load_sum4(a, i);            // in flight while the factor records are fetched
    if (h1) r1 = load_sum4(a, i + stride);

    const int fetchers = min((int)gridDim.x, kFetchPerRank * a.N);
    if (flame && (int)blockIdx.x < fetchers) {
        const int r = blockIdx.x / kFetchPerRank, q = blockIdx.x % kFetchPerRank;
        const int rec4 = (LH + 2 * n3 + 3) / 4, per = (rec4 + kFetchPerRank - 1) / kFetchPerRank;
        const int lo4 = min(rec4, q * per), hi4 = min(rec4, lo4 + per);
        const float4* src = reinterpret_cast<const float4*>(a.peers[r] + a.in_off + a.rec_off + (size_t)r * a.rec_stride);
        float4* dst = reinterpret_cast<float4*>(a.local + a.gather_off + (size_t)r * a.rec_stride);
        for (int base = lo4; base < hi4; base += kStageUnroll * kExThreads) {
            float4 v[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int e = base + u * kExThreads + t;
                if (e < hi4) v[u] = __ldcv(src + e);
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
                const int e = base + u * kExThreads + t;
                if (e < hi4) __stcg(dst + e, v[u]);
            }
        }
        __threadfence();
        __syncthreads();
        if (t == 0) atomicAdd(flags + kFlagReady, 1u);
    }
    if (h0) reduce_store(a, i, r0);
    if (h1) reduce_store(a, i + stride, r1);
    for (i += 2 * stride; i < hi; i += stride) reduce_store(a, i, load_sum4(a, i));
    if (!flame) return;
    if (t == 0) {
        while (*reinterpret_cast<volatile uint32_t*>(flags + kFlagReady) < (uint32_t)fetchers) {
        }
        __threadfence();
    }
    __syncthreads();
    {   // stage the heads of all records and this CTA's factor rows from the local gather area (L2)
        const float* g = a.local + a.gather_off;
        for (int e = t; e < a.N * LH; e += kExThreads) s_head[e] = __ldcg(g + (size_t)(e / LH) * a.rec_stride + (e % LH));
        for (int e = t; e < a.N * nrows; e += kExThreads) {
            const int r = e / nrows, k = e - r * nrows;
            s_gs[r * rpc + k] = __ldcg(g + (size_t)r * a.rec_stride + LH + row0 + k) * a.scale;
            s_gp[r * rpc + k] = __ldcg(g + (size_t)r * a.rec_stride + LH + n3 + row0 + k) * a.scale;
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && t == 0) {
        uint32_t* fl = reinterpret_cast<uint32_t*>(a.local + a.flags_off);
        *reinterpret_cast<unsigned long long*>(fl + kFlagStage) +=
            globaltimer_ns() - *reinterpret_cast<volatile unsigned long long*>(fl + kFlagTBar);
    }

    // ---- expansion of this CTA's rows, from shared memory only ----
    if (a.d_dp) {  // d_delta_posedirs[i][e] = sum_r pose_feature_r[i] * g_posed_r[e]
        for (int q = t; q < a.NP * nrows; q += kExThreads) {
            const int ii = q / nrows, k = q - ii * nrows;
            float o = 0.f;
#pragma unroll
            for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r)
                if (r < a.N) o += s_head[r * LH + a.L + ii] * s_gp[r * rpc + k];
            __stcs(a.d_dp + (size_t)ii * n3 + row0 + k, o);
        }
    }
    const bool vec = (a.L & 3) == 0 && (a.l0 & 3) == 0 && (LH & 3) == 0;
    for (int k = wid; k < nrows; k += kExThreads / 32) {
        float gs[FS_FLAME_MAX_RANKS], sum = 0.f;
#pragma unroll
        for (int r = 0; r < FS_FLAME_MAX_RANKS; ++r) {
            gs[r] = r < a.N ? s_gs[r * rpc + k] : 0.0f;
            sum += gs[r];
        }
        const int row = row0 + k;
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
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ uint32_t s_epoch;
    uint32_t* flags = reinterpret_cast<uint32_t*>(a.local + a.flags_off);
    unsigned long long t_entry = 0;
    if (threadIdx.x == 0) {
        s_epoch = *reinterpret_cast<volatile uint32_t*>(flags + kFlagEpoch) + 1u;
        if (blockIdx.x == 0) t_entry = globaltimer_ns();
    }
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
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned long long t = globaltimer_ns();
        *reinterpret_cast<volatile unsigned long long*>(flags + kFlagTBar) = t;
        *reinterpret_cast<unsigned long long*>(flags + kFlagWait) += t - t_entry;
    }
    exchange_body(a, s_dyn);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *reinterpret_cast<unsigned long long*>(flags + kFlagCta0) +=
            globaltimer_ns() - *reinterpret_cast<volatile unsigned long long*>(flags + kFlagTBar);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // this CTA's multicast stores are performed before it counts as finished
        const uint32_t prev = atomicAdd(flags + kFlagCtas, 1u);
        if (prev == gridDim.x - 1) {
            flags[kFlagCtas] = 0u;
            flags[kFlagReady] = 0u;
            if (a.algo == 2)
                for (int r = 0; r < a.N; ++r) {
                    uint32_t* pf = reinterpret_cast<uint32_t*>(const_cast<float*>(a.peers[r]) + a.flags_off);
                    st_release_sys(pf + kFlagDone + a.rank, epoch);
                }
            *reinterpret_cast<unsigned long long*>(flags + kFlagWork) +=
                globaltimer_ns() - *reinterpret_cast<volatile unsigned long long*>(flags + kFlagTBar);
            flags[kFlagCalls] += 1u;
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

extern "C" size_t fs_p2p_exchange_flag_floats(void) { return 64; }

// Device-side timing of the exchange kernel since the last reset (host call, synchronises the device): mean ns a rank
// spent waiting in the barrier for its peers, mean ns from the barrier to the last CTA's exit, number of calls.
extern "C" int fs_p2p_exchange_timing(float* d_local_base, size_t flags_offset, double* wait_ns, double* work_ns,
                                      int* calls, int reset) {
    uint32_t h[64];
    if (cudaMemcpy(h, d_local_base + flags_offset, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fs_set_error("fs_p2p_exchange_timing: copy failed");
        return FS_ERR_CUDA;
    }
    unsigned long long w, k;
    memcpy(&w, h + kFlagWait, 8);
    memcpy(&k, h + kFlagWork, 8);
    const int n = (int)h[kFlagCalls];
    if (calls) *calls = n;
    if (wait_ns) *wait_ns = n ? (double)w / n : 0.0;
    if (work_ns) *work_ns = n ? (double)k / n : 0.0;
    if (getenv("FATESPLAT_EXCHANGE_TRACE")) {
        unsigned long long s1, s2;
        memcpy(&s1, h + kFlagStage, 8);
        memcpy(&s2, h + kFlagCta0, 8);
        fprintf(stderr, "[fs_p2p_exchange] calls %d: wait %.1f us, CTA0 staged after %.1f us, CTA0 done after %.1f us, last CTA after %.1f us\n",
                n, n ? w / 1e3 / n : 0.0, n ? s1 / 1e3 / n : 0.0, n ? s2 / 1e3 / n : 0.0, n ? k / 1e3 / n : 0.0);
    }
    if (reset) cudaMemset(reinterpret_cast<uint32_t*>(d_local_base + flags_offset) + kFlagWait, 0, 12 * sizeof(uint32_t));
    return FS_OK;
}

extern "C" int fs_p2p_exchange(int N, int rank, int algo, const float* const* d_peer_ptrs, const float* d_multicast,
                               float* d_local_base, size_t in_offset, size_t n_splat, size_t rec_offset,
                               size_t rec_stride, size_t out_offset, size_t flags_offset, size_t gather_offset,
                               float* d_out, int V, int L, int l0, int NP, float scale, float* d_dL_ddelta_vertex, float* d_dL_ddelta_shapedirs,
                               float* d_dL_ddelta_posedirs, void* stream) {
    const bool flame = d_dL_ddelta_vertex || d_dL_ddelta_shapedirs || d_dL_ddelta_posedirs;
    if (N < 1 || N > FS_FLAME_MAX_RANKS || rank < 0 || rank >= N || algo < 0 || algo > 2 || !d_peer_ptrs || !d_local_base ||
        ((in_offset | n_splat | rec_offset | rec_stride | out_offset | flags_offset | gather_offset) & 3) != 0 ||
        (algo != 0 && !d_multicast) || (algo != 2 && !d_out) ||
        (flame && (V <= 0 || L <= 0 || NP < 0 || l0 < 0 || l0 > L || rec_stride < (size_t)L + NP + 6 * (size_t)V))) {
        fs_set_error("fs_p2p_exchange: invalid argument (1 <= N <= %d, offsets multiples of 4 floats, multicast base for "
                     "algo 1/2, local output for algo 0/1, record stride >= L + NP + 6V)", FS_FLAME_MAX_RANKS);
        return FS_ERR_INVALID_ARGUMENT;
    }
    ExchangeArgs a;
    a.N = N; a.rank = rank; a.algo = algo; a.peers = d_peer_ptrs; a.mc = d_multicast; a.local = d_local_base;
    a.in_off = in_offset; a.n_splat = n_splat; a.rec_off = rec_offset; a.rec_stride = rec_stride; a.out_off = out_offset;
    a.flags_off = flags_offset; a.gather_off = gather_offset; a.out = d_out; a.V = V; a.L = L; a.l0 = l0; a.NP = NP; a.scale = scale;
    a.d_dv = d_dL_ddelta_vertex; a.d_ds = d_dL_ddelta_shapedirs; a.d_dp = d_dL_ddelta_posedirs;
    const int grid = 2 * std::max(1, fs_tuning("FATESPLAT_EXCHANGE_CTAS_PER_SM", 1) * fs_num_sms());
    const int rpc = flame ? (3 * V + grid - 1) / grid : 0;
    const size_t smem = flame ? ((size_t)N * (L + NP) + 2 * (size_t)N * rpc) * sizeof(float) : 0;
    if (smem > 200 * 1024) {
        fs_set_error("fs_p2p_exchange: N * (L + NP) floats of record heads do not fit in shared memory");
        return FS_ERR_UNSUPPORTED;
    }
    static std::atomic<unsigned long long> attr_set{0};
    if (fs_first_use_on_device(attr_set))
        cudaFuncSetAttribute(p2p_exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FsStageTimer timer(FS_STAGE_EXCHANGE, st);
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
