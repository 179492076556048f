// Front-to-back alpha blend of one 16x16 tile per CTA.
//
// Replaces DGR cuda_rasterizer/forward.cu:261-374 (renderCUDA).  Same per-pixel rule, evaluated in the same
// fp32 operation order, so final_T / n_contrib / colour agree with the reference:
//     power = -1/2 (A dx^2 + C dy^2) - B dx dy ; skip if power > 0
//     alpha = min(0.99, o * exp(power))        ; skip if alpha < 1/255
//     stop the pixel when T (1 - alpha) < 1e-4 ; else C += rgb * alpha * T, T *= 1 - alpha
//
// B200 design:
//   * the tile's splat records were gathered into depth order by the sort kernel, so a batch of 256 records
//     is one contiguous 12 KB block: it is staged with ONE cp.async.bulk (TMA 1-D, UBLKCP) per batch into a
//     2-stage shared-memory ring guarded by mbarriers -- no per-thread gather loads, no index indirection;
//   * each warp owns an 8x4 pixel block.  Per 32 records, every lane tests one record's conservative
//     alpha>=1/255 bounding box against the warp's block and the warp walks only the ballot survivors.
//     Skipped records could not have contributed to any of the warp's pixels, and positions in the list are
//     still counted, so n_contrib is unchanged.  This removes ~75-85 % of the exp/FMA work of the reference's
//     every-pixel-visits-every-record loop, which is what bounds this kernel (it is SFU/issue bound, not HBM
//     bound: algorithmic traffic is 48*R + 20*W*H + 8*Tn bytes).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBatch = 256;
constexpr int kStages = 2;

__global__ void __launch_bounds__(kThreads)
blend_forward_kernel(const uint2* __restrict__ ranges, const SplatRec* __restrict__ inst_splat, int W, int H,
                     const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, uint32_t Rcap) {
    __shared__ __align__(128) SplatRec s_rec[kStages][kBatch];
    __shared__ __align__(8) uint64_t s_full[kStages];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gx = (W + FS_TILE - 1) / FS_TILE;
    const int tile = blockIdx.x;
    const int tile_x = tile % gx, tile_y = tile / gx;
    // warp -> 8x4 pixel block, lane -> pixel
    const int bx = tile_x * FS_TILE + (wid & 1) * 8, by = tile_y * FS_TILE + (wid >> 1) * 4;
    const int px = bx + (lane & 7), py = by + (lane >> 3);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    // warp block bounds in pixel-centre coordinates
    const float wx0 = (float)bx, wx1 = (float)min(bx + 7, W - 1), wy0 = (float)by, wy1 = (float)min(by + 3, H - 1);

    uint2 range = ranges[tile];
    if (range.y > Rcap) range = make_uint2(0u, 0u);  // overflowed frame: flagged in the header, stay in bounds
    const uint32_t total = range.y - range.x;
    const int nbatches = (int)((total + kBatch - 1) / kBatch);

    if (tid == 0) {
        fs::mbar_init(&s_full[0], 1);
        fs::mbar_init(&s_full[1], 1);
        fs::mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int b) {
        const uint32_t cnt = min((uint32_t)kBatch, total - (uint32_t)b * kBatch);
        const uint32_t bytes = cnt * (uint32_t)sizeof(SplatRec);
        fs::mbar_expect_tx(&s_full[b & 1], bytes);
        fs::bulk_g2s(&s_rec[b & 1][0], inst_splat + range.x + (size_t)b * kBatch, bytes, &s_full[b & 1]);
    };
    if (tid == 0 && nbatches > 0) issue(0);

    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t last_contributor = 0;
    bool done = !inside;
    bool warp_done = __all_sync(0xffffffffu, done);

    int b = 0;
    for (; b < nbatches; ++b) {
        if (tid == 0 && b + 1 < nbatches) issue(b + 1);  // stage (b+1)&1 was released by the barrier below
        fs::mbar_wait(&s_full[b & 1], (uint32_t)(b >> 1) & 1u);
        const SplatRec* rec = s_rec[b & 1];
        const int cnt = (int)min((uint32_t)kBatch, total - (uint32_t)b * kBatch);
        if (!warp_done) {
            for (int c = 0; c < cnt; c += 32) {
                const int j = c + lane;
                bool hit = false;
                if (j < cnt) {
                    const float4 q0 = rec[j].q0;
                    hit = !(q0.z < 0.0f) &&
                          !(q0.x + q0.z < wx0 || q0.x - q0.z > wx1 || q0.y + q0.w < wy0 || q0.y - q0.w > wy1);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1;
                    if (!done) {
                        const float4 q0 = rec[c + bit].q0;
                        const float4 q1 = rec[c + bit].q1;
                        const float dx = fs::sub(q0.x, pxf), dy = fs::sub(q0.y, pyf);
                        const float power = fs::splat_power(dx, dy, q1.x, q1.y, q1.z);
                        if (power <= 0.0f) {
                            const float alpha = fminf(0.99f, fs::mul(q1.w, expf(power)));
                            if (alpha >= 1.0f / 255.0f) {
                                const float test_T = fs::mul(T, fs::sub(1.0f, alpha));
                                if (test_T < 0.0001f) {
                                    done = true;
                                } else {
                                    const float4 q2 = rec[c + bit].q2;
                                    C0 = fs::mad(T, fs::mul(alpha, q2.x), C0);
                                    C1 = fs::mad(T, fs::mul(alpha, q2.y), C1);
                                    C2 = fs::mad(T, fs::mul(alpha, q2.z), C2);
                                    T = test_T;
                                    last_contributor = (uint32_t)b * kBatch + (uint32_t)(c + bit) + 1u;
                                }
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) {
                    warp_done = true;
                    break;
                }
            }
        }
        // all warps finished with this stage (it may be refilled next iteration); stop when every pixel is done
        if (__syncthreads_and(warp_done)) {
            ++b;
            break;
        }
    }
    // a prefetch issued for batch `b` may still be in flight: drain it before the CTA retires
    if (b < nbatches && b > 0) fs::mbar_wait(&s_full[b & 1], (uint32_t)(b >> 1) & 1u);

    if (inside) {
        const size_t pid = (size_t)py * W + px;
        final_T[pid] = T;
        n_contrib[pid] = last_contributor;
        const size_t plane = (size_t)H * W;
        out_color[pid] = fs::mad(__ldg(bg_color + 0), T, C0);
        out_color[plane + pid] = fs::mad(__ldg(bg_color + 1), T, C1);
        out_color[2 * plane + pid] = fs::mad(__ldg(bg_color + 2), T, C2);
    }
}

}  // namespace

void fs_launch_blend_forward(int W, int H, const float* bg, float* out_color, char* ws, const fs_workspace_layout& L,
                             cudaStream_t stream) {
    const int gx = (W + FS_TILE - 1) / FS_TILE, gy = (H + FS_TILE - 1) / FS_TILE;
    blend_forward_kernel<<<gx * gy, kThreads, 0, stream>>>(
        reinterpret_cast<const uint2*>(ws + L.ranges), reinterpret_cast<const SplatRec*>(ws + L.inst_splat), W, H, bg,
        out_color, reinterpret_cast<float*>(ws + L.final_T), reinterpret_cast<uint32_t*>(ws + L.n_contrib),
        (uint32_t)L.instance_capacity);
    fs_count_launch(1);
}
