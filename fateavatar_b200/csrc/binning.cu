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
// Per-tile sort: keys of one tile are spread over a narrow depth interval, so a linear bucket pass on the
// depth bits (shared-memory atomics give each key its rank inside its bucket), an exclusive scan and a tiny
// per-bucket insertion sort order the segment in O(n) work; pathological distributions (one bucket holding
// most keys) fall back to a bitonic network.  Every path ends in the same total order on the 64-bit key.
//
// HBM traffic: 8*R (scatter) + 8*R + 4*R + 48*R (sort in/out + record gather) vs. the reference's 12*R emit +
// 144*R sort + 8*R ranges.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
constexpr int kScanThreads = 1024;

// ---- K2: exclusive scan of per-tile counts; ranges; big-tile list; frame header ---------------------------
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(int Tn, const uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_cursor,
                 uint2* __restrict__ ranges, uint32_t* __restrict__ big_tiles, uint32_t* __restrict__ work_order,
                 uint32_t* __restrict__ seg_base, uint2* __restrict__ seg_info, uint4* __restrict__ tile_meta,
                 fs_frame_info* __restrict__ info, volatile fs_frame_info* host_info, uint32_t Rcap) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_wmax[32];
    __shared__ uint32_t s_nbig;
    __shared__ uint32_t s_bin[33];  // tiles per log2(count) class; class 0 = empty
    __shared__ uint32_t s_wseg[32];
    __shared__ uint32_t s_wfull[32];  // full (FS_SEG-position) segments: they are listed first, partial ones last
    __shared__ uint32_t s_nfull;
    fs::pdl_trigger();
    fs::pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_nbig = 0;
    if (tid < 33) s_bin[tid] = 0;
    __syncthreads();
    const int per = (Tn + kScanThreads - 1) / kScanThreads;
    const int beg = min(Tn, tid * per), end = min(Tn, beg + per);
    uint32_t sum = 0, mx = 0, segs = 0, fulls = 0;
    for (int t = beg; t < end; ++t) {
        const uint32_t c = tile_count[(size_t)t * FS_CNT_STRIDE];
        sum += c;
        segs += (c + FS_SEG - 1) / FS_SEG;
        fulls += c / FS_SEG;
        mx = max(mx, c);
        atomicAdd(&s_bin[c ? 32 - __clz(c) : 0], 1u);
    }
    uint32_t incl = sum, sincl = segs, fincl = fulls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        const uint32_t w = __shfl_up_sync(0xffffffffu, sincl, o);
        const uint32_t x = __shfl_up_sync(0xffffffffu, fincl, o);
        if (lane >= o) {
            incl += v;
            sincl += w;
            fincl += x;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 31) {
        s_warp[wid] = incl;
        s_wseg[wid] = sincl;
        s_wfull[wid] = fincl;
    }
    if (lane == 0) s_wmax[wid] = mx;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = s_warp[lane];
        uint32_t wi = w;
        uint32_t ws_ = s_wseg[lane];
        uint32_t wsi = ws_;
        uint32_t wf_ = s_wfull[lane];
        uint32_t wfi = wf_;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            const uint32_t v2 = __shfl_up_sync(0xffffffffu, wsi, o);
            const uint32_t v3 = __shfl_up_sync(0xffffffffu, wfi, o);
            if (lane >= o) {
                wi += v;
                wsi += v2;
                wfi += v3;
            }
        }
        s_warp[lane] = wi - w;  // exclusive prefix of warp totals
        s_wseg[lane] = wsi - ws_;
        s_wfull[lane] = wfi - wf_;
        if (lane == 31) s_nfull = wfi;
        if (lane == 31) {
            info->reserved[2] = wsi;  // depth segments in this frame (work units of the backward blend / 8)
            seg_base[Tn] = wsi;
        }
        uint32_t m = s_wmax[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 31) {
            info->num_rendered = wi;
            info->overflow = (wi > Rcap) ? 1u : 0u;
            info->max_tile_instances = m;
            if (host_info) {
                // early notification: the caller's pinned header learns R / overflow as soon as they exist, long
                // before the frame finishes, so a host that must know R (the reference's blocking read,
                // rasterizer_impl.cu:281) can keep queueing work behind this frame instead of idling the GPU.
                // num_rendered is written last: the host polls it.
                host_info->overflow = (wi > Rcap) ? 1u : 0u;
                host_info->max_tile_instances = m;
                host_info->num_visible = info->num_visible;
                __threadfence_system();
                host_info->num_rendered = wi;
            }
        }
    }
    __syncthreads();
    uint32_t off = s_warp[wid] + (incl - sum);
    uint32_t soff = s_wseg[wid] + (sincl - segs);
    // work list of the backward blend: every full segment first (tile order), then the partial ones -- the small
    // units come last, so the warps run out of work at about the same time
    uint32_t foff = s_wfull[wid] + (fincl - fulls);
    uint32_t poff = s_nfull + (soff - foff);
    for (int t = beg; t < end; ++t) {
        const uint32_t c = tile_count[(size_t)t * FS_CNT_STRIDE];
        seg_base[t] = soff;
        if (off + c <= Rcap) {  // an overflowed frame never reads these
            for (uint32_t k = 0; k < c / FS_SEG; ++k) seg_info[foff + k] = make_uint2((uint32_t)t, k);
            if (c % FS_SEG) seg_info[poff] = make_uint2((uint32_t)t, c / FS_SEG);
        }
        foff += c / FS_SEG;
        poff += (c % FS_SEG) ? 1u : 0u;
        soff += (c + FS_SEG - 1) / FS_SEG;
        tile_cursor[(size_t)t * FS_CNT_STRIDE] = off;
        ranges[t] = c ? make_uint2(off, off + c) : make_uint2(0u, 0u);  // empty tiles stay (0,0) like the memset
        tile_meta[t] = c ? make_uint4(off, off + c, seg_base[t], 0u) : make_uint4(0u, 0u, 0u, 0u);
        if (c > FS_SORT_SMEM_CAP) big_tiles[1 + atomicAdd(&s_nbig, 1u)] = (uint32_t)t;
        off += c;
    }
    __syncthreads();
    if (tid == 0) {
        big_tiles[0] = s_nbig;
        info->reserved[3] = (uint32_t)Tn - s_bin[0];  // non-empty tiles (work units of the backward blend)
        uint32_t o = 0;  // heaviest class first; s_bin becomes the running write cursor of each class
        for (int b = 32; b >= 0; --b) {
            const uint32_t n = s_bin[b];
            s_bin[b] = o;
            o += n;
        }
    }
    __syncthreads();
    for (int t = beg; t < end; ++t) {
        const uint32_t c = tile_count[(size_t)t * FS_CNT_STRIDE];
        work_order[atomicAdd(&s_bin[c ? 32 - __clz(c) : 0], 1u)] = (uint32_t)t;
    }
}

// ---- K3: scatter (depth, gaussian) keys into tile segments -------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_kernel(int P, int gx, const ushort4* __restrict__ rect, const float* __restrict__ depths,
               uint32_t* __restrict__ tile_cursor, u64* __restrict__ keys, uint32_t Rcap) {
    fs::pdl_trigger();
    fs::pdl_wait();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    u64 key = 0;
    if (idx < P) {
        rc = rect[idx];
        key = ((u64)__float_as_uint(depths[idx]) << 32) | (uint32_t)idx;
    }
    const int w = max(0, (int)rc.z - (int)rc.x), h = max(0, (int)rc.w - (int)rc.y);
    const int count = w * h;
    // A splat that covers many tiles would make its thread the tail of the whole kernel (one returning atomic per
    // tile, serially): such rectangles are handed to the warp, 32 tiles per round; small ones stay with their lane.
    constexpr int kOwnMax = 32;
    if (count <= kOwnMax) {
        for (int y = rc.y; y < rc.w; ++y)
            for (int x = rc.x; x < rc.z; ++x) {
                const uint32_t pos = atomicAdd(&tile_cursor[(size_t)(y * gx + x) * FS_CNT_STRIDE], 1u);
                if (pos < Rcap) keys[pos] = key;
            }
    }
    unsigned big = __ballot_sync(0xffffffffu, count > kOwnMax);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bx0 = __shfl_sync(0xffffffffu, (int)rc.x, src), by0 = __shfl_sync(0xffffffffu, (int)rc.y, src);
        const int bw = __shfl_sync(0xffffffffu, w, src), bn = __shfl_sync(0xffffffffu, count, src);
        const uint32_t klo = __shfl_sync(0xffffffffu, (uint32_t)key, src);
        const uint32_t khi = __shfl_sync(0xffffffffu, (uint32_t)(key >> 32), src);
        const u64 bkey = ((u64)khi << 32) | klo;
        for (int t = lane; t < bn; t += 32) {
            const int y = by0 + t / bw, x = bx0 + t % bw;
            const uint32_t pos = atomicAdd(&tile_cursor[(size_t)(y * gx + x) * FS_CNT_STRIDE], 1u);
            if (pos < Rcap) keys[pos] = bkey;
        }
    }
}

// ---- bitonic network with ascending-only compare-exchanges (virtual +inf padding needs no storage) --------
__device__ __forceinline__ void bitonic_sort_ascending(u64* s, uint32_t n, int tid, int nthreads) {
    if (n < 2) return;
    uint32_t m = 1, lg = 0;
    while (m < n) {
        m <<= 1;
        ++lg;
    }
    const uint32_t half = m >> 1;
    for (uint32_t lk = 1; lk <= lg; ++lk) {
        const uint32_t k = 1u << lk;
        for (uint32_t i = tid; i < half; i += nthreads) {  // flip step: partner mirrored inside each block of k
            const uint32_t blk = i >> (lk - 1), o = i & ((k >> 1) - 1);
            const uint32_t lo = (blk << lk) + o, hi = (blk << lk) + (k - 1 - o);
            if (hi < n) {
                const u64 a = s[lo], b = s[hi];
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
                    const u64 a = s[lo], b = s[hi];
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

// ---- linear-bucket sort of one tile segment held in shared memory ------------------------------------------
// Keys are unique (they embed the Gaussian id), so after the keys are grouped by depth bucket every key's final
// position is  bucket_start + #(keys of the same bucket that are smaller): one short, fully parallel scan of
// the key's own bucket -- no serial insertion tail.  s_grp[n] = bucket-grouped keys, s_fin[n] = sorted result.
// PER = ceil(CAP / THREADS) keys per thread.
template <int THREADS, int PER, int MAXB>
__device__ __forceinline__ void bucket_sort(const u64* __restrict__ gkeys, uint32_t n, u64* s_grp, u64* s_fin,
                                            uint32_t* s_cnt, uint32_t* s_red) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = THREADS / 32;
    // load + depth range
    uint32_t dmin = 0xffffffffu, dmax = 0u;
    u64 mine[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const uint32_t i = tid + k * THREADS;
        if (i < n) {
            mine[k] = gkeys[i];
            const uint32_t d = (uint32_t)(mine[k] >> 32);
            dmin = min(dmin, d);
            dmax = max(dmax, d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0) {
        s_red[wid] = dmin;
        s_red[NW + wid] = dmax;
    }
    uint32_t nb = 1;
    while (nb < n && nb < (uint32_t)MAXB) nb <<= 1;
    for (uint32_t i = tid; i < nb; i += THREADS) s_cnt[i] = 0;
    __syncthreads();
    dmin = s_red[0];
    dmax = s_red[NW];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
        dmin = min(dmin, s_red[w]);
        dmax = max(dmax, s_red[NW + w]);
    }
    // monotone map depth bits -> bucket (uint->float, * positive constant, truncation are all monotone)
    const float scale = (float)nb / ((float)(dmax - dmin) + 1.0f);
    uint32_t br[PER];  // bucket | rank<<16
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const uint32_t i = tid + k * THREADS;
        if (i < n) {
            const uint32_t d = (uint32_t)(mine[k] >> 32);
            const uint32_t b = min(nb - 1, (uint32_t)((float)(d - dmin) * scale));
            const uint32_t r = atomicAdd(&s_cnt[b], 1u);
            br[k] = b | (r << 16);
        }
    }
    __syncthreads();
    // exclusive scan of s_cnt[0..nb) in place, tracking the largest bucket
    const uint32_t per = (nb + THREADS - 1) / THREADS;
    const uint32_t b0 = min(nb, tid * per), b1 = min(nb, b0 + per);
    uint32_t sum = 0, big = 0;
    for (uint32_t b = b0; b < b1; ++b) {
        sum += s_cnt[b];
        big = max(big, s_cnt[b]);
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) big = max(big, __shfl_xor_sync(0xffffffffu, big, o));
    __syncthreads();  // all reads of s_red above are done
    if (lane == 31) s_red[wid] = incl;
    if (lane == 0) s_red[NW + wid] = big;
    __syncthreads();
    uint32_t woff = 0, maxb = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        if (w < wid) woff += s_red[w];
        maxb = max(maxb, s_red[NW + w]);
    }
    uint32_t off = woff + (incl - sum);
    for (uint32_t b = b0; b < b1; ++b) {
        const uint32_t c = s_cnt[b];
        s_cnt[b] = off;
        off += c;
    }
    if (tid == THREADS - 1) s_cnt[nb] = n;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const uint32_t i = tid + k * THREADS;
        if (i < n) s_grp[s_cnt[br[k] & 0xffffu] + (br[k] >> 16)] = mine[k];
    }
    __syncthreads();
    if (maxb <= 512) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const uint32_t i = tid + k * THREADS;
            if (i < n) {
                const uint32_t b = br[k] & 0xffffu;
                const uint32_t lo = s_cnt[b], hi = s_cnt[b + 1];
                uint32_t rank = lo;
                for (uint32_t j = lo; j < hi; ++j) rank += (s_grp[j] < mine[k]) ? 1u : 0u;
                s_fin[rank] = mine[k];
            }
        }
        __syncthreads();
    } else {  // one bucket holds most keys (extreme depth outliers): still exact, just slower
        bitonic_sort_ascending(s_grp, n, tid, THREADS);
        for (uint32_t i = tid; i < n; i += THREADS) s_fin[i] = s_grp[i];
        __syncthreads();
    }
}

// write sorted ids and gather the splat records into sorted order (coalesced 16-byte stores)
__device__ __forceinline__ void emit_sorted(const u64* keys, uint32_t n, uint32_t start,
                                            const float4* __restrict__ splat, uint32_t* __restrict__ point_list,
                                            float4* __restrict__ inst_splat, int tid, int nthreads) {
    for (uint32_t i = tid; i < n; i += nthreads) point_list[start + i] = (uint32_t)keys[i];
    for (uint32_t e = tid; e < n * 3; e += nthreads) {
        const uint32_t i = e / 3, part = e - i * 3;
        const uint32_t g = (uint32_t)keys[i];
        inst_splat[(size_t)(start + i) * 3 + part] = __ldg(splat + (size_t)g * 3 + part);
    }
}

// ---- K4: one CTA per tile, segments of up to FS_SORT_SMEM_CAP keys -----------------------------------------
#ifndef FS_SORT_THREADS
#define FS_SORT_THREADS 512
#endif
constexpr int kSortThreads = FS_SORT_THREADS;
constexpr int kSortPer = FS_SORT_SMEM_CAP / kSortThreads;
constexpr size_t kSortSmemBytes = (size_t)FS_SORT_SMEM_CAP * 16 + (FS_SORT_SMEM_CAP + 1) * 4 + 12;
__global__ void __launch_bounds__(kSortThreads)
tile_sort_kernel(const uint2* __restrict__ ranges, u64* __restrict__ keys, const float4* __restrict__ splat,
                 uint32_t* __restrict__ point_list, float4* __restrict__ inst_splat, uint32_t Rcap,
                 int big_kernel_follows) {
    extern __shared__ __align__(16) u64 s_sort[];
    u64* s_grp = s_sort;
    u64* s_out = s_sort + FS_SORT_SMEM_CAP;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_sort + 2 * FS_SORT_SMEM_CAP);
    __shared__ uint32_t s_red[2 * (kSortThreads / 32)];
    fs::pdl_trigger();
    fs::pdl_wait();
    const uint2 r = ranges[blockIdx.x];
    const uint32_t n = r.y - r.x;
    if (n == 0 || r.y > Rcap) return;
    if (n > FS_SORT_SMEM_CAP) {
        // oversized tile: normally left to big_tile_sort_kernel; when the caller's hint said no tile would be
        // this large and that launch was skipped, sort it here in global memory (slow, but always correct)
        if (!big_kernel_follows) {
            bitonic_sort_ascending(keys + r.x, n, threadIdx.x, kSortThreads);
            emit_sorted(keys + r.x, n, r.x, splat, point_list, inst_splat, threadIdx.x, kSortThreads);
        }
        return;
    }
    bucket_sort<kSortThreads, kSortPer, FS_SORT_SMEM_CAP>(keys + r.x, n, s_grp, s_out, s_cnt, s_red);
    emit_sorted(s_out, n, r.x, splat, point_list, inst_splat, threadIdx.x, kSortThreads);
}

// ---- K4b: persistent CTAs over the (usually empty) list of oversized tiles ---------------------------------
// n <= kBigBucketCap: bucket sort in dynamic shared memory; n <= kBigSmemCap: bitonic in shared memory;
// beyond: the same network directly on the global key segment.
constexpr int kBigThreads = 1024;
constexpr uint32_t kBigBucketCap = 8192;
constexpr uint32_t kBigSmemCap = 24576;  // 192 KB of 64-bit keys
constexpr size_t kBigSmemBytes = kBigSmemCap * sizeof(u64);
__global__ void __launch_bounds__(kBigThreads)
big_tile_sort_kernel(const uint32_t* __restrict__ big_tiles, const uint2* __restrict__ ranges,
                     u64* __restrict__ keys, const float4* __restrict__ splat, uint32_t* __restrict__ point_list,
                     float4* __restrict__ inst_splat, uint32_t Rcap) {
    extern __shared__ __align__(16) u64 d_smem[];
    __shared__ uint32_t s_red[2 * (kBigThreads / 32)];
    fs::pdl_trigger();
    fs::pdl_wait();
    const uint32_t nbig = big_tiles[0];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint2 r = ranges[big_tiles[1 + b]];
        const uint32_t n = r.y - r.x;
        if (r.y > Rcap) continue;
        if (n <= kBigBucketCap) {
            u64* s_grp = d_smem;                                                      // 64 KB
            u64* s_out = d_smem + kBigBucketCap;                                      // 64 KB
            uint32_t* s_cnt = reinterpret_cast<uint32_t*>(d_smem + 2 * kBigBucketCap);  // 32 KB + 4
            bucket_sort<kBigThreads, kBigBucketCap / kBigThreads, kBigBucketCap>(keys + r.x, n, s_grp, s_out, s_cnt,
                                                                               s_red);
            emit_sorted(s_out, n, r.x, splat, point_list, inst_splat, threadIdx.x, kBigThreads);
        } else if (n <= kBigSmemCap) {
            for (uint32_t i = threadIdx.x; i < n; i += kBigThreads) d_smem[i] = keys[r.x + i];
            __syncthreads();
            bitonic_sort_ascending(d_smem, n, threadIdx.x, kBigThreads);
            emit_sorted(d_smem, n, r.x, splat, point_list, inst_splat, threadIdx.x, kBigThreads);
        } else {
            u64* g = keys + r.x;
            bitonic_sort_ascending(g, n, threadIdx.x, kBigThreads);
            emit_sorted(g, n, r.x, splat, point_list, inst_splat, threadIdx.x, kBigThreads);
        }
        __syncthreads();
    }
}

}  // namespace

void fs_launch_binning(int P, int W, int H, char* ws, const fs_workspace_layout& L, fs_frame_info* host_info_dev,
                       cudaStream_t stream) {
    const int gx = (W + FS_TILE - 1) / FS_TILE, gy = (H + FS_TILE - 1) / FS_TILE, Tn = gx * gy;
    const uint32_t Rcap = (uint32_t)L.instance_capacity;
    auto* info = reinterpret_cast<fs_frame_info*>(ws + L.info);
    auto* tile_count = reinterpret_cast<uint32_t*>(ws + L.tile_count);
    auto* tile_cursor = reinterpret_cast<uint32_t*>(ws + L.tile_cursor);
    auto* ranges = reinterpret_cast<uint2*>(ws + L.ranges);
    auto* big = reinterpret_cast<uint32_t*>(ws + L.big_tiles);
    auto* work_order = reinterpret_cast<uint32_t*>(ws + L.work_order);
    auto* seg_base = reinterpret_cast<uint32_t*>(ws + L.seg_base);
    auto* seg_info = reinterpret_cast<uint2*>(ws + L.seg_info);
    auto* tile_meta = reinterpret_cast<uint4*>(ws + L.tile_meta);
    auto* keys = reinterpret_cast<u64*>(ws + L.inst_keys);
    auto* splat = reinterpret_cast<const float4*>(ws + L.splat);
    auto* point_list = reinterpret_cast<uint32_t*>(ws + L.point_list);
    auto* inst_splat = reinterpret_cast<float4*>(ws + L.inst_splat);
    // The oversized-tile kernel costs ~6 us of pure launch latency; skip it when the caller's hint (heaviest tile
    // of recent frames, fs_set_tile_hint) says no tile comes near the in-kernel capacity.  A wrong hint only
    // costs speed: tile_sort_kernel then sorts such a tile itself in global memory.
    const uint32_t hint = fs_tile_hint();
    const bool launch_big = hint == 0 || hint > (FS_SORT_SMEM_CAP * 3) / 4;
    {
        FsStageTimer t(FS_STAGE_TILE_SCAN, stream);
        fs_launch_pdl(tile_scan_kernel, dim3(1), dim3(kScanThreads), 0, stream, Tn, tile_count, tile_cursor, ranges, big,
                      work_order, seg_base, seg_info, tile_meta, info, (volatile fs_frame_info*)host_info_dev, Rcap);
    }
    {
        FsStageTimer t(FS_STAGE_SCATTER, stream);
        fs_launch_pdl(scatter_kernel, dim3((P + 255) / 256), dim3(256), 0, stream, P, gx,
                      reinterpret_cast<const ushort4*>(ws + L.rect), reinterpret_cast<const float*>(ws + L.depths),
                      tile_cursor, keys, Rcap);
    }
    {
        FsStageTimer t(FS_STAGE_TILE_SORT, stream);
        static std::atomic<unsigned long long> sort_attr_set{0};
        if (fs_first_use_on_device(sort_attr_set))
            cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSortSmemBytes);
        fs_launch_pdl(tile_sort_kernel, dim3(Tn), dim3(kSortThreads), kSortSmemBytes, stream, ranges, keys, splat, point_list,
                      inst_splat, Rcap, launch_big ? 1 : 0);
    }
    if (!launch_big) {
        fs_count_launch(3);
        return;
    }
    static std::atomic<unsigned long long> attr_set{0};
    if (fs_first_use_on_device(attr_set))
        cudaFuncSetAttribute(big_tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBigSmemBytes);
    {
        FsStageTimer t(FS_STAGE_BIG_TILE_SORT, stream);
        fs_launch_pdl(big_tile_sort_kernel, dim3(64), dim3(kBigThreads), kBigSmemBytes, stream, big, ranges, keys, splat,
                      point_list, inst_splat, Rcap);
    }
    fs_count_launch(4);
}
