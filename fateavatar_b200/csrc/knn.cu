// Mean squared distance to the 3 nearest neighbours of every point (simple-knn's distCUDA2).
//
// Replaces simple-knn/simple_knn.cu:186-222 (SimpleKNN::knn: 2 cub reductions with blocking copies, Morton
// codes, radix sort, 1024-point boxes, box scan per point) and spatial.cu:15-26.
//
// The reference's result is an *exact* 3-NN query (its box pruning is conservative), so any exact search
// returns the same three squared distances; per-pair arithmetic follows the reference's contraction
// mad(dz,dz, mad(dx,dx, dy*dy)) and the final (b0 + b1 + b2) / 3.
//
// B200 design: uniform grid instead of Morton boxes.  ~4 points per cell, counting-sort the points by cell
// (histogram + single-CTA scan + scatter, no host sync, no allocation), then each point searches rings of
// cells outward until the third-best distance is inside the searched shell.  O(P) work with coalesced
// cell-ordered queries instead of the reference's O(P * P/1024) box tests.
#include <cfloat>
#include <cmath>

#include "common.cuh"

namespace {

struct KnnParams {
    float mn[3];
    float inv_cs[3];  // cells per unit length (0 for a degenerate axis)
    float min_cs;     // smallest cell edge over non-degenerate axes (+inf if none)
    int G;
};

constexpr int kRedCtas = 148;

__global__ void __launch_bounds__(256) knn_bbox_kernel(int P, const float* __restrict__ pts, float* __restrict__ partial) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
    __shared__ float s[8][6];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s[wid][k] = mn[k];
            s[wid][3 + k] = mx[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, s[w][threadIdx.x]) : fmaxf(v, s[w][threadIdx.x]);
        partial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__device__ __forceinline__ void cell_of(const KnnParams& kp, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(kp.G - 1, max(0, (int)((x - kp.mn[0]) * kp.inv_cs[0])));
    cy = min(kp.G - 1, max(0, (int)((y - kp.mn[1]) * kp.inv_cs[1])));
    cz = min(kp.G - 1, max(0, (int)((z - kp.mn[2]) * kp.inv_cs[2])));
}

// finishes the bbox reduction (every CTA redundantly; 148*6 floats), assigns cells, builds the histogram
__global__ void __launch_bounds__(256)
knn_cell_kernel(int P, int G, const float* __restrict__ pts, const float* __restrict__ partial,
                KnnParams* __restrict__ params, uint32_t* __restrict__ cell_of_point,
                uint32_t* __restrict__ cell_count) {
    __shared__ KnnParams kp;
    __shared__ float s_b[6];
    if (threadIdx.x < 6) {
        float v = partial[threadIdx.x];
        for (int b = 1; b < kRedCtas; ++b)
            v = threadIdx.x < 3 ? fminf(v, partial[b * 6 + threadIdx.x]) : fmaxf(v, partial[b * 6 + threadIdx.x]);
        s_b[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float min_cs = __int_as_float(0x7f800000);
        for (int k = 0; k < 3; ++k) {
            const float ext = s_b[3 + k] - s_b[k];
            kp.mn[k] = s_b[k];
            if (ext > 0.0f && isfinite(ext)) {
                const float cs = ext / (float)G;
                kp.inv_cs[k] = (float)G / ext;
                min_cs = fminf(min_cs, cs);
            } else {
                kp.inv_cs[k] = 0.0f;
            }
        }
        kp.min_cs = min_cs;
        kp.G = G;
        if (blockIdx.x == 0) *params = kp;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int cx, cy, cz;
    cell_of(kp, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], cx, cy, cz);
    const uint32_t c = (uint32_t)((cz * G + cy) * G + cx);
    cell_of_point[i] = c;
    atomicAdd(&cell_count[c], 1u);
}

__global__ void __launch_bounds__(1024)
knn_scan_kernel(int n, const uint32_t* __restrict__ count, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int beg = min(n, tid * per), end = min(n, beg + per);
    uint32_t sum = 0;
    for (int t = beg; t < end; ++t) sum += count[t];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = s_warp[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t off = s_warp[wid] + (incl - sum);
    for (int t = beg; t < end; ++t) {
        start[t] = off;
        cursor[t] = off;
        off += count[t];
    }
    if (tid == 1023) start[n] = off;
}

__global__ void __launch_bounds__(256)
knn_scatter_kernel(int P, const float* __restrict__ pts, const uint32_t* __restrict__ cell_of_point,
                   uint32_t* __restrict__ cursor, float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint32_t pos = atomicAdd(&cursor[cell_of_point[i]], 1u);
    sorted[pos] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], __uint_as_float((uint32_t)i));
}

__device__ __forceinline__ void upd3(float d, float& b0, float& b1, float& b2) {  // updateKBest<3>
    if (b0 > d) { const float t = b0; b0 = d; d = t; }
    if (b1 > d) { const float t = b1; b1 = d; d = t; }
    if (b2 > d) { b2 = d; }
}

__global__ void __launch_bounds__(128)
knn_query_kernel(int P, const KnnParams* __restrict__ params, const uint32_t* __restrict__ start,
                 const float4* __restrict__ sorted, float* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P) return;
    const KnnParams kp = *params;
    const int G = kp.G;
    const float4 me = sorted[s];
    int cx, cy, cz;
    cell_of(kp, me.x, me.y, me.z, cx, cy, cz);
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int ring = 0; ring < G; ++ring) {
        const int z0 = max(0, cz - ring), z1 = min(G - 1, cz + ring);
        const int y0 = max(0, cy - ring), y1 = min(G - 1, cy + ring);
        const int x0 = max(0, cx - ring), x1 = min(G - 1, cx + ring);
        auto visit = [&](int x, int y, int z) {
            const uint32_t c = (uint32_t)((z * G + y) * G + x);
            const uint32_t pb = start[c], pe = start[c + 1];
            for (uint32_t p = pb; p < pe; ++p) {
                if ((int)p == s) continue;
                const float4 o = sorted[p];
                const float dx = __fsub_rn(o.x, me.x), dy = __fsub_rn(o.y, me.y), dz = __fsub_rn(o.z, me.z);
                upd3(fs::dot3(dx, dx, dy, dy, dz, dz), b0, b1, b2);
            }
        };
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                if (abs(z - cz) == ring || abs(y - cy) == ring) {  // z/y face of the shell: the whole x row
                    for (int x = x0; x <= x1; ++x) visit(x, y, z);
                } else {  // interior row: only the two x end caps belong to this ring
                    if (cx - ring >= 0) visit(cx - ring, y, z);
                    if (cx + ring <= G - 1) visit(cx + ring, y, z);
                }
            }
        // every unvisited point is at least ring * min_cs away along some axis
        const float reach = (float)ring * kp.min_cs * 0.999f;
        if (b2 < FLT_MAX && b2 <= reach * reach) break;
    }
    out[__float_as_uint(me.w)] = __fdiv_rn(__fadd_rn(__fadd_rn(b0, b1), b2), 3.0f);
}

inline int knn_grid_dim(int P) {
    int G = (int)cbrt((double)P / 4.0);
    if (G < 1) G = 1;
    if (G > 256) G = 256;
    return G;
}
inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

struct KnnLayout {
    size_t partial, params, cell_of_point, count, start, cursor, sorted, total;
};
inline KnnLayout knn_layout(int P) {
    const size_t G = knn_grid_dim(P), nc = G * G * G;
    KnnLayout L;
    size_t off = 0;
    L.partial = off; off = al(off + kRedCtas * 6 * 4);
    L.params = off; off = al(off + sizeof(KnnParams));
    L.cell_of_point = off; off = al(off + (size_t)P * 4);
    L.count = off; off = al(off + nc * 4);
    L.start = off; off = al(off + (nc + 1) * 4);
    L.cursor = off; off = al(off + nc * 4);
    L.sorted = off; off = al(off + (size_t)P * 16);
    L.total = off;
    return L;
}

}  // namespace

size_t fs_knn_workspace_bytes_impl(int P) { return knn_layout(P > 0 ? P : 1).total; }

int fs_launch_knn(int P, const float* points, float* out, char* ws, size_t ws_bytes, cudaStream_t stream) {
    const KnnLayout L = knn_layout(P);
    if (ws_bytes < L.total) {
        fs_set_error("fs_knn_mean_dist2: workspace too small (%zu < %zu bytes)", ws_bytes, L.total);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    const int G = knn_grid_dim(P), nc = G * G * G;
    auto* partial = reinterpret_cast<float*>(ws + L.partial);
    auto* params = reinterpret_cast<KnnParams*>(ws + L.params);
    auto* cop = reinterpret_cast<uint32_t*>(ws + L.cell_of_point);
    auto* count = reinterpret_cast<uint32_t*>(ws + L.count);
    auto* start = reinterpret_cast<uint32_t*>(ws + L.start);
    auto* cursor = reinterpret_cast<uint32_t*>(ws + L.cursor);
    auto* sorted = reinterpret_cast<float4*>(ws + L.sorted);
    cudaMemsetAsync(count, 0, (size_t)nc * 4, stream);
    knn_bbox_kernel<<<kRedCtas, 256, 0, stream>>>(P, points, partial);
    knn_cell_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, G, points, partial, params, cop, count);
    knn_scan_kernel<<<1, 1024, 0, stream>>>(nc, count, start, cursor);
    knn_scatter_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, points, cop, cursor, sorted);
    knn_query_kernel<<<(P + 127) / 128, 128, 0, stream>>>(P, params, start, sorted, out);
    fs_count_launch(5);
    return FS_OK;
}
