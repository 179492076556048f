// Tile binning and per-tile depth sort.
//
// Replaces the reference's global pipeline  InclusiveSum -> duplicateWithKeys -> 64-bit DeviceRadixSort (6
// onesweep passes over R pairs) -> identifyTileRanges  (DGR rasterizer_impl.cu:70-138, 277-317) by a
// bucket sort:  per-tile histogram (done in preprocess) -> exclusive scan over Tn tiles (this file, one CTA;
// yields `ranges` for free) -> scatter of (depth,gaussian) keys into each tile's segment -> one CTA per tile
// sorts its segment in shared memory and gathers the 48-byte splat records into sorted order.
//
// Result equivalence: the reference sorts keys (tile<<32 | depth_bits) with a *stable* sort after emitting
// instances in ascending Gaussian index, so within a tile the order is (depth_bits, gaussian) ascending.
// Sorting the composite 64-bit key (depth_bits<<32 | gaussian) per tile gives the identical unique order,
// independent of the atomic scatter order => point_list and ranges are bit-exact.
//
// HBM traffic: 8*R (scatter) + 8*R + 4*R + 48*R (sort in/out + record gather) vs. the reference's 12*R emit +
// 144*R sort + 8*R ranges.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 1024;

// ---- K2: exclusive scan of per-tile counts; ranges; big-tile list; frame header ---------------------------
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int Tn, const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_cursor,
                 uint2* __restrict__ ranges, uint32_t* __restrict__ big_tiles, fs_frame_info* __restrict__ info,
                 uint32_t Rcap) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_wmax[32];
    __shared__ uint32_t s_nbig;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_nbig = 0;
    const int per = (Tn + kScanThreads - 1) / kScanThreads;
    const int beg = min(Tn, tid * per), end = min(Tn, beg + per);
    uint32_t sum = 0, mx = 0;
    for (int t = beg; t < end; ++t) {
        const uint32_t c = tile_count[t];
        sum += c;
        mx = max(mx, c);
    }
    // block exclusive scan of `sum`
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 31) s_warp[wid] = incl;
    if (lane == 0) s_wmax[wid] = mx;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;  // exclusive prefix of warp totals
        uint32_t m = s_wmax[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 31) {
            info->num_rendered = wi;
            info->overflow = (wi > Rcap) ? 1u : 0u;
            info->max_tile_instances = m;
        }
    }
    __syncthreads();
    uint32_t off = s_warp[wid] + (incl - sum);
    for (int t = beg; t < end; ++t) {
        const uint32_t c = tile_count[t];
        tile_cursor[t] = off;
        ranges[t] = c ? make_uint2(off, off + c) : make_uint2(0u, 0u);  // empty tiles stay (0,0) like the memset
        if (c > FS_SORT_SMEM_CAP) big_tiles[1 + atomicAdd(&s_nbig, 1u)] = (uint32_t)t;
        off += c;
    }
    __syncthreads();
    if (tid == 0) big_tiles[0] = s_nbig;
}

// ---- K3: scatter (depth, gaussian) keys into tile segments -------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int gx, const ushort4* __restrict__ rect, const float* __restrict__ depths,
               uint32_t* __restrict__ tile_cursor, unsigned long long* __restrict__ keys, uint32_t Rcap) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const ushort4 rc = rect[idx];
    if (rc.x >= rc.z || rc.y >= rc.w) return;
    const unsigned long long key = ((unsigned long long)__float_as_uint(depths[idx]) << 32) | (uint32_t)idx;
    for (int y = rc.y; y < rc.w; ++y)
        for (int x = rc.x; x < rc.z; ++x) {
            const uint32_t pos = atomicAdd(&tile_cursor[y * gx + x], 1u);
            if (pos < Rcap) keys[pos] = key;
        }
}

// ---- bitonic network with ascending-only compare-exchanges (virtual +inf padding needs no storage) --------
template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_ascending(KeyPtr s, uint32_t n, int tid, int nthreads) {
    if (n < 2) return;
    uint32_t m = 1;
    while (m < n) m <<= 1;
    const uint32_t half = m >> 1;
    for (uint32_t k = 2; k <= m; k <<= 1) {
        // flip step: partner mirrored inside each block of k
        for (uint32_t i = tid; i < half; i += nthreads) {
            const uint32_t blk = i / (k >> 1), o = i % (k >> 1);
            const uint32_t lo = blk * k + o, hi = blk * k + (k - 1 - o);
            if (hi < n) {
                const unsigned long long a = s[lo], b = s[hi];
                if (a > b) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < half; i += nthreads) {
                const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo + j;
                if (hi < n) {
                    const unsigned long long a = s[lo], b = s[hi];
                    if (a > b) {
                        s[lo] = b;
                        s[hi] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// write sorted ids and gather the splat records into sorted order (coalesced 16-byte stores)
__device__ __forceinline__ void emit_sorted(const unsigned long long* keys, uint32_t n, uint32_t start,
                                            const float4* __restrict__ splat, uint32_t* __restrict__ point_list,
                                            float4* __restrict__ inst_splat, int tid, int nthreads) {
    for (uint32_t i = tid; i < n; i += nthreads) point_list[start + i] = (uint32_t)keys[i];
    for (uint32_t e = tid; e < n * 3; e += nthreads) {
        const uint32_t i = e / 3, part = e - i * 3;
        const uint32_t g = (uint32_t)keys[i];
        inst_splat[(size_t)(start + i) * 3 + part] = __ldg(splat + (size_t)g * 3 + part);
    }
}

// ---- K4: one CTA per tile, segment <= FS_SORT_SMEM_CAP sorted in 32 KB of shared memory --------------------
constexpr int kSortThreads = 256;
__global__ void __launch_bounds__(kSortThreads)
tile_sort_kernel(const uint2* __restrict__ ranges, const unsigned long long* __restrict__ keys,
                 const float4* __restrict__ splat, uint32_t* __restrict__ point_list, float4* __restrict__ inst_splat,
                 uint32_t Rcap) {
    __shared__ unsigned long long s_keys[FS_SORT_SMEM_CAP];
    const uint2 r = ranges[blockIdx.x];
    const uint32_t n = r.y - r.x;
    if (n == 0 || n > FS_SORT_SMEM_CAP || r.y > Rcap) return;
    for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) s_keys[i] = keys[r.x + i];
    __syncthreads();
    bitonic_sort_ascending(s_keys, n, threadIdx.x, kSortThreads);
    emit_sorted(s_keys, n, r.x, splat, point_list, inst_splat, threadIdx.x, kSortThreads);
}

// ---- K4b: persistent CTAs over the (usually empty) list of oversized tiles ---------------------------------
// up to kBigSmemCap instances in dynamic shared memory; beyond that, the same network in global memory.
constexpr int kBigThreads = 1024;
constexpr uint32_t kBigSmemCap = 24576;  // 192 KB of 64-bit keys
__global__ void __launch_bounds__(kBigThreads)
big_tile_sort_kernel(const uint32_t* __restrict__ big_tiles, const uint2* __restrict__ ranges,
                     unsigned long long* __restrict__ keys, const float4* __restrict__ splat,
                     uint32_t* __restrict__ point_list, float4* __restrict__ inst_splat, uint32_t Rcap) {
    extern __shared__ __align__(16) unsigned long long d_keys[];
    const uint32_t nbig = big_tiles[0];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint2 r = ranges[big_tiles[1 + b]];
        const uint32_t n = r.y - r.x;
        if (r.y > Rcap) continue;
        if (n <= kBigSmemCap) {
            for (uint32_t i = threadIdx.x; i < n; i += kBigThreads) d_keys[i] = keys[r.x + i];
            __syncthreads();
            bitonic_sort_ascending(d_keys, n, threadIdx.x, kBigThreads);
            emit_sorted(d_keys, n, r.x, splat, point_list, inst_splat, threadIdx.x, kBigThreads);
        } else {
            unsigned long long* g = keys + r.x;
            bitonic_sort_ascending(g, n, threadIdx.x, kBigThreads);
            emit_sorted(g, n, r.x, splat, point_list, inst_splat, threadIdx.x, kBigThreads);
        }
        __syncthreads();
    }
}

}  // namespace

void fs_launch_binning(int P, int W, int H, char* ws, const fs_workspace_layout& L, cudaStream_t stream) {
    const int gx = (W + FS_TILE - 1) / FS_TILE, gy = (H + FS_TILE - 1) / FS_TILE, Tn = gx * gy;
    const uint32_t Rcap = (uint32_t)L.instance_capacity;
    auto* info = reinterpret_cast<fs_frame_info*>(ws + L.info);
    auto* tile_count = reinterpret_cast<uint32_t*>(ws + L.tile_count);
    auto* tile_cursor = reinterpret_cast<uint32_t*>(ws + L.tile_cursor);
    auto* ranges = reinterpret_cast<uint2*>(ws + L.ranges);
    auto* big = reinterpret_cast<uint32_t*>(ws + L.big_tiles);
    auto* keys = reinterpret_cast<unsigned long long*>(ws + L.inst_keys);
    auto* splat = reinterpret_cast<const float4*>(ws + L.splat);
    auto* point_list = reinterpret_cast<uint32_t*>(ws + L.point_list);
    auto* inst_splat = reinterpret_cast<float4*>(ws + L.inst_splat);

    tile_scan_kernel<<<1, kScanThreads, 0, stream>>>(Tn, tile_count, tile_cursor, ranges, big, info, Rcap);
    scatter_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, gx, reinterpret_cast<const ushort4*>(ws + L.rect),
                                                        reinterpret_cast<const float*>(ws + L.depths), tile_cursor,
                                                        keys, Rcap);
    tile_sort_kernel<<<Tn, kSortThreads, 0, stream>>>(ranges, keys, splat, point_list, inst_splat, Rcap);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(big_tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kBigSmemCap * sizeof(unsigned long long)));
        attr_set = true;
    }
    big_tile_sort_kernel<<<64, kBigThreads, kBigSmemCap * sizeof(unsigned long long), stream>>>(
        big, ranges, keys, splat, point_list, inst_splat, Rcap);
    fs_count_launch(4);
}
