// Front-to-back alpha blend.
//
// Replaces DGR cuda_rasterizer/forward.cu:261-374 (renderCUDA).  Same per-pixel rule, evaluated in the same
// fp32 operation order, so final_T / n_contrib / colour agree with the reference:
//     power = -1/2 (A dx^2 + C dy^2) - B dx dy ; skip if power > 0
//     alpha = min(0.99, o * exp(power))        ; skip if alpha < 1/255
//     stop the pixel when T (1 - alpha) < 1e-4 ; else C += rgb * alpha * T, T *= 1 - alpha
//
// B200 design:
//   * work unit = one 8x4 pixel block of one tile, owned by ONE WARP.  Units are handed out heaviest-first
//     from a device work list (built by the tile scan) through an atomic counter to persistent warps, so the
//     few hundred dense tiles of a head spread evenly over all 592 SM sub-partitions instead of following
//     the launch order (a CTA-per-tile grid left the busiest sub-partition with 2.4x the average work);
//   * the tile's splat records were gathered into depth order by the sort kernel, so a batch of records is
//     one contiguous block: each warp streams it with cp.async.bulk (TMA 1-D, UBLKCP) into its private
//     2-stage shared-memory ring guarded by mbarriers -- no per-thread gather loads, no index indirection,
//     no CTA-wide barrier;
//   * per 32 records, every lane tests one record's conservative alpha>=1/255 bounding box against the
//     warp's block and the warp walks only the ballot survivors, four at a time so that the four
//     power/exp chains overlap.  Skipped records could not have contributed to any of the warp's pixels, and
//     positions in the list are still counted, so n_contrib is unchanged.
// The kernel is issue/SFU bound, not HBM bound: algorithmic traffic is 48*R + 20*W*H + 8*Tn bytes.
#include "common.cuh"

namespace {

constexpr int kWarps = 8;     // warps per CTA (independent; they only share the CTA's shared memory)
constexpr int kBatch = 64;    // records per stage
constexpr int kStages = 2;
#ifndef FS_FWD_GROUP
#define FS_FWD_GROUP 4
#endif
constexpr int kGroup = FS_FWD_GROUP;

struct WarpStage {
    SplatRec rec[kStages][kBatch];
};

struct Group {
    float alpha[kGroup];  // < 0: skip (not a survivor, power > 0, or alpha < 1/255)
    float4 col[kGroup];
    int jj[kGroup];
    bool any;
};

// Second, exact stage of the cull (the first is the bounding box of the alpha >= 1/255 ellipse): the largest alpha the
// splat reaches anywhere on the block's rectangle.  alpha >= 1/255 needs q(d) = 1/2 (A dx^2 + C dy^2) + B dx dy <=
// tau = ln(255 o), d = mean - pixel.  q is convex, so its minimum over the rectangle is 0 when the mean lies inside and
// otherwise sits on the edge(s) facing the mean: on the edge dx = X the minimiser is dy = clamp(-B X / C), and the
// interior of an edge that does not face the mean can never hold the minimum (there dq/dx = X det / C points inward).
// Continuous minimum <= minimum over the block's pixels, and 0.01 covers the fp32 rounding of both sides, so a record
// rejected here could not have been blended by any pixel of the block; positions in the list are still counted.
__device__ __forceinline__ bool ellipse_reaches_block(float mx, float my, float4 con_o, float wx0, float wx1, float wy0,
                                                      float wy1) {
    const float A = con_o.x, B = con_o.y, C = con_o.z;
    if (!(A > 0.0f && C > 0.0f && A * C - B * B > 0.0f)) return true;  // not an ellipse: leave it to the blend rule
    const float X0 = mx - wx1, X1 = mx - wx0, Y0 = my - wy1, Y1 = my - wy0;  // d ranges over [X0,X1] x [Y0,Y1]
    const bool xin = X0 <= 0.0f && X1 >= 0.0f, yin = Y0 <= 0.0f && Y1 >= 0.0f;
    if (xin && yin) return true;
    float qmin = 3.0e38f;
    if (!xin) {
        const float X = X0 > 0.0f ? X0 : X1;
        const float y = fminf(fmaxf(__fdividef(-B * X, C), Y0), Y1);
        qmin = 0.5f * (A * X * X + C * y * y) + B * X * y;
    }
    if (!yin) {
        const float Y = Y0 > 0.0f ? Y0 : Y1;
        const float x = fminf(fmaxf(__fdividef(-B * Y, A), X0), X1);
        qmin = fminf(qmin, 0.5f * (A * x * x + C * Y * Y) + B * x * Y);
    }
    const float tau = __logf(255.0f * con_o.w);
    return !(qmin > tau + 0.01f);
}

// extract the next (up to) four set bits of m and evaluate alpha for this lane's pixel
__device__ __forceinline__ void compute_group(Group& g, unsigned& m, int c, const SplatRec* __restrict__ rec, float pxf,
                                              float pyf) {
    bool live[kGroup];
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {  // branch-free extraction
        const int fbit = __ffs(m);
        live[k] = fbit != 0;
        g.jj[k] = live[k] ? c + fbit - 1 : c;
        m &= m - 1;
    }
    float4 q0[kGroup], q1[kGroup];
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
        q0[k] = rec[g.jj[k]].q0;
        q1[k] = rec[g.jj[k]].q1;
        g.col[k] = rec[g.jj[k]].q2;
    }
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
        const float dx = fs::sub(q0[k].x, pxf), dy = fs::sub(q0[k].y, pyf);
        const float power = fs::splat_power(dx, dy, q1[k].x, q1[k].y, q1[k].z);
        const float a = fminf(0.99f, fs::mul(q1[k].w, expf(power)));
        g.alpha[k] = (live[k] && power <= 0.0f && a >= 1.0f / 255.0f) ? a : -1.0f;
    }
    g.any = true;
}

// sequential transmittance update (front-to-back order), fully predicated
__device__ __forceinline__ void apply_group(const Group& g, uint32_t base, int c, int lane, float& T, float& C0,
                                            float& C1, float& C2, bool& done, uint32_t& last_contributor,
                                            uint32_t& my_mask) {
#pragma unroll
    for (int k = 0; k < kGroup; ++k) {
        const float test_T = fs::mul(T, fs::sub(1.0f, g.alpha[k]));
        const bool act = g.alpha[k] >= 0.0f && !done;
        const bool stop = act && test_T < 0.0001f;
        const bool apply = act && !stop;
        done = done || stop;
        const float n0 = fs::mad(T, fs::mul(g.alpha[k], g.col[k].x), C0);
        const float n1 = fs::mad(T, fs::mul(g.alpha[k], g.col[k].y), C1);
        const float n2 = fs::mad(T, fs::mul(g.alpha[k], g.col[k].z), C2);
        C0 = apply ? n0 : C0;
        C1 = apply ? n1 : C1;
        C2 = apply ? n2 : C2;
        T = apply ? test_T : T;
        last_contributor = apply ? base + (uint32_t)g.jj[k] + 1u : last_contributor;
        // which pixels of the block blended this record: kept by the lane that tested the record, stored once per
        // 32 records (pair_mask; the backward blend reads the decisions instead of re-deriving them)
        const unsigned bm = __ballot_sync(0xffffffffu, apply);
        my_mask |= (lane == g.jj[k] - c) ? bm : 0u;
    }
}

__global__ void __launch_bounds__(kWarps * 32)
blend_forward_kernel(const uint4* __restrict__ tile_meta, const uint32_t* __restrict__ work_order, int n_tiles,
                     const uint32_t* __restrict__ n_nonempty_tiles, uint32_t sm_count,
                     uint32_t* __restrict__ sm_slots, uint32_t* __restrict__ work_counter, const SplatRec* __restrict__ inst_splat, int W, int H,
                     float4* __restrict__ ckpt, float4* __restrict__ final_C,
                     const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ final_T,
                     uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ pair_mask,
                     uint32_t* __restrict__ work_stats, uint32_t Rcap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage* stages = reinterpret_cast<WarpStage*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(WarpStage) * kWarps);

    fs::pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // Dynamic balancing needs clearly more heavy units than warps (a warp that owns a single dense unit is
    // the critical path), so only as many CTAs stay active as there are ~2 dense units per warp; the launch
    // is sized for the largest case and surplus CTAs retire immediately.  CTA i runs on SM (i mod #SM).
    {
        const uint32_t dense_units = __ldg(n_nonempty_tiles) * 8u;
        const uint32_t want_per_sm = max(1u, dense_units / (2u * kWarps * sm_count));
        // placement-independent: the k-th CTA to arrive on an SM stays iff k < want_per_sm
        __shared__ uint32_t s_rank;
        if (threadIdx.x == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_rank = atomicAdd(&sm_slots[smid & 255u], 1u);
        }
        __syncthreads();
        if (s_rank >= want_per_sm) return;
    }
    SplatRec(*rec_ring)[kBatch] = stages[wid].rec;
    uint64_t* s_full = bars + wid * kStages;
    if (lane == 0) {
        fs::mbar_init(&s_full[0], 1);
        fs::mbar_init(&s_full[1], 1);
        fs::mbar_fence_init();
    }
    __syncwarp();
    uint32_t fills = 0;  // batches issued so far by this warp: stage = fills & 1, parity = (fills >> 1) & 1
    uint32_t n_box = 0;  // records that passed the box cull (warp-uniform)

    const int gx = (W + FS_TILE - 1) / FS_TILE;
    const float bg0 = __ldg(bg_color + 0), bg1 = __ldg(bg_color + 1), bg2 = __ldg(bg_color + 2);
    const uint32_t n_units = (uint32_t)n_tiles * 8u;

    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(work_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const int tile = (int)work_order[unit >> 3];
        const int blk = (int)(unit & 7u);  // 8x4 block inside the tile: 2 columns x 4 rows
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx = tile_x * FS_TILE + (blk & 1) * 8, by = tile_y * FS_TILE + (blk >> 1) * 4;
        const int px = bx + (lane & 7), py = by + (lane >> 3);
        const bool inside = px < W && py < H;
        const float pxf = (float)px, pyf = (float)py;
        const float wx0 = (float)bx, wx1 = (float)min(bx + 7, W - 1), wy0 = (float)by, wy1 = (float)min(by + 3, H - 1);

        const uint4 meta = tile_meta[tile];  // range and first checkpoint slot in one 16-byte load
        uint2 range = make_uint2(meta.x, meta.y);
        if (range.y > Rcap) range = make_uint2(0u, 0u);  // overflowed frame: flagged in the header, stay in bounds
        const uint32_t total = range.y - range.x;
        const int nbatches = (int)((total + kBatch - 1) / kBatch);

        auto issue = [&](int b) {  // lane 0 only
            const uint32_t cnt = min((uint32_t)kBatch, total - (uint32_t)b * kBatch);
            const uint32_t bytes = cnt * (uint32_t)sizeof(SplatRec);
            const uint32_t s = (fills + (uint32_t)b) & 1u;
            fs::mbar_expect_tx(&s_full[s], bytes);
            fs::bulk_g2s(&rec_ring[s][0], inst_splat + range.x + (size_t)b * kBatch, bytes, &s_full[s]);
        };
        if (lane == 0 && nbatches > 0) issue(0);

        float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
        uint32_t last_contributor = 0;
        bool done = !inside;
        bool warp_done = __all_sync(0xffffffffu, done);
        // checkpoints for the backward blend: per-pixel (T, C) before list positions k*FS_SEG, k = 1, 2, ...
        float4* ck = ckpt + (size_t)meta.z * FS_TILE_PIX + ((by - tile_y * FS_TILE) + (lane >> 3)) * FS_TILE +
                     (bx - tile_x * FS_TILE) + (lane & 7);
        uint32_t* mask_out = pair_mask + (size_t)range.x * 8u + (size_t)blk * total;

        int b = 0;
        for (; b < nbatches; ++b) {
            __syncwarp();  // every lane is done reading the stage that batch b+1 will overwrite
            if (lane == 0 && b + 1 < nbatches) issue(b + 1);
            if (b > 0 && (b * kBatch) % FS_SEG == 0) {
                ck[(size_t)(b * kBatch / FS_SEG) * FS_TILE_PIX] = make_float4(T, C0, C1, C2);
            }
            const uint32_t f = fills + (uint32_t)b;
            fs::mbar_wait(&s_full[f & 1u], (f >> 1) & 1u);
            const SplatRec* rec = rec_ring[f & 1u];
            const int cnt = (int)min((uint32_t)kBatch, total - (uint32_t)b * kBatch);
            for (int c = 0; c < cnt && !warp_done; c += 32) {
                const int j = c + lane;
                bool hit = false;
                if (j < cnt) {
                    const float4 q0 = rec[j].q0;
                    hit = !(q0.z < 0.0f) &&
                          !(q0.x + q0.z < wx0 || q0.x - q0.z > wx1 || q0.y + q0.w < wy0 || q0.y - q0.w > wy1);
                    if (hit) hit = ellipse_reaches_block(q0.x, q0.y, rec[j].q1, wx0, wx1, wy0, wy1);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                n_box += (uint32_t)__popc(m);
                uint32_t my_mask = 0u;
                // Survivors are walked four at a time.  The vote ends the basic block, so all four alphas are
                // needed at once and the scheduler overlaps the four power/exp chains instead of trailing them
                // behind the (short, sequential) transmittance chain; it also skips groups that touch no pixel.
                while (m) {
                    Group g;
                    compute_group(g, m, c, rec, pxf, pyf);
                    bool any_a = false;
#pragma unroll
                    for (int k = 0; k < kGroup; ++k) any_a |= g.alpha[k] >= 0.0f;
                    if (!__any_sync(0xffffffffu, any_a)) continue;
                    apply_group(g, (uint32_t)b * kBatch, c, lane, T, C0, C1, C2, done, last_contributor, my_mask);
                }
                if (j < cnt) mask_out[(size_t)b * kBatch + j] = my_mask;
                warp_done = __all_sync(0xffffffffu, done);
            }
            if (warp_done) {  // every pixel of the block has terminated: stop streaming
                ++b;
                break;
            }
        }
        // a prefetch issued for batch `b` may still be in flight: drain it so the ring can be reused
        if (b < nbatches && b > 0) {
            const uint32_t f = fills + (uint32_t)b;
            fs::mbar_wait(&s_full[f & 1u], (f >> 1) & 1u);
            ++b;
        }
        fills += (uint32_t)b;  // number of batches actually issued for this unit

        if (inside) {
            const size_t pid = (size_t)py * W + px;
            final_T[pid] = T;
            final_C[pid] = make_float4(C0, C1, C2, 0.0f);
            n_contrib[pid] = last_contributor;
            const size_t plane = (size_t)H * W;
            out_color[pid] = fs::mad(bg0, T, C0);
            out_color[plane + pid] = fs::mad(bg1, T, C1);
            out_color[2 * plane + pid] = fs::mad(bg2, T, C2);
        }
    }
    if (lane == 0 && n_box) atomicAdd(work_stats + FS_WORK_FWD_BOX, n_box);
}

}  // namespace

void fs_launch_blend_forward(int W, int H, const float* bg, float* out_color, char* ws, const fs_workspace_layout& L,
                             cudaStream_t stream) {
    const int gx = (W + FS_TILE - 1) / FS_TILE, gy = (H + FS_TILE - 1) / FS_TILE;
    FsStageTimer timer(FS_STAGE_BLEND_FWD, stream);
    const size_t smem = sizeof(WarpStage) * kWarps + sizeof(uint64_t) * kStages * kWarps;
    static std::atomic<unsigned long long> attr_set{0};
    if (fs_first_use_on_device(attr_set))
        cudaFuncSetAttribute(blend_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto* info = reinterpret_cast<fs_frame_info*>(ws + L.info);
    const int ctas_per_sm = fs_tuning("FATESPLAT_FWD_CTAS_PER_SM", 4);  // upper bound; see the kernel prologue
    const int grid = fs_num_sms() * ctas_per_sm;
    fs_launch_pdl(blend_forward_kernel, dim3(grid), dim3(kWarps * 32), smem, stream,
        reinterpret_cast<const uint4*>(ws + L.tile_meta), reinterpret_cast<const uint32_t*>(ws + L.work_order), gx * gy,
        &info->reserved[3], (uint32_t)fs_num_sms(), reinterpret_cast<uint32_t*>(info + 1), &info->reserved[1], reinterpret_cast<const SplatRec*>(ws + L.inst_splat), W, H,
        reinterpret_cast<float4*>(ws + L.ckpt),
        reinterpret_cast<float4*>(ws + L.final_C), bg, out_color,
        reinterpret_cast<float*>(ws + L.final_T), reinterpret_cast<uint32_t*>(ws + L.n_contrib),
        reinterpret_cast<uint32_t*>(ws + L.pair_mask), reinterpret_cast<uint32_t*>(ws + L.info + FS_WORK_FWD_OFFSET),
        (uint32_t)L.instance_capacity);
    fs_count_launch(1);
}
