// Backward of the splat render: blend backward (per tile) + fused per-Gaussian backward.
//
// Replaces DGR cuda_rasterizer/backward.cu:399-557 (renderCUDA bwd), :144-274 (computeCov2DCUDA),
// :346-396 (preprocessCUDA bwd) [+ :20-139 SH, :278-341 cov3D] and the nine torch::zeros of
// rasterize_points.cu:151-159.
//
// The hand-derived gradient keeps the reference's deviations from "autograd of the forward":
// no zeroing at the alpha = 0.99 clamp, T recovered by division, 1/(det^2 + 1e-7), frustum-clamp masks only
// on dL/dt.x, dL/dt.y, no quaternion-normalisation Jacobian, SH clamp via the saved flags, and pixels skip
// instances at positions >= n_contrib.
//
// B200 design of the blend backward (the dominant training kernel):
//   * same tiling as the forward: sorted 48-byte records streamed back-to-front with cp.async.bulk + mbarrier,
//     8x4 pixels per warp, ballot culling against the alpha >= 1/255 bounding box;
//   * the reference issues 9 global float atomics per contributing (pixel, Gaussian) pair.  Here the 9 partial
//     derivatives are butterfly-reduced over the warp, summed over the CTA's 8 warps in shared memory, and
//     flushed once per (tile, Gaussian) with three 128-bit vector reductions (red.global.add.v4.f32) into a
//     48-byte per-Gaussian accumulator -- ~100x fewer L2 atomic operations;
//   * the per-Gaussian kernel consumes that accumulator and writes every API gradient exactly once, so no
//     output tensor needs a zero-fill.
#include "common.cuh"

namespace {

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};


__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One step of a transposing warp reduction: N live values -> N/2, pairing lanes that differ in bit OFF.
// The lane with the bit set keeps the upper half.  After steps 16, 8, 4 on 32 values every lane holds the
// 4 values {i + (lane & 28)}, summed over its 8-lane class; two plain xor steps finish the sum.
template <int OFF, int N>
__device__ __forceinline__ void treduce_step(float* v, int lane) {
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const float send = up ? v[i] : v[i + N / 2];
        const float keep = up ? v[i + N / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
}

constexpr int kGroup = 4;

// grad_acc layout per Gaussian (12 floats): [0]=dmean2D.x [1]=dmean2D.y [2]=dconic.x [3]=dconic.y
//                                           [4]=dconic.w  [5]=dopacity  [6..8]=dcolor rgb  [9..11]=0
constexpr int kWarps = 8;   // independent warps per CTA
constexpr int kBatch = 64;  // records per stage
struct WarpStage {
    SplatRec rec[2][kBatch];
};

// Persistent warps pull (tile, depth segment, 8x4 block) units from an atomic work counter.  A depth segment is
// FS_SEG consecutive positions of the tile's sorted list; the forward kernel left, for every pixel, the
// transmittance and the colour accumulated behind each segment boundary (ckpt), so a unit can start its
// back-to-front walk at its own segment instead of at the end of the list.  This bounds the serial chain of a
// unit (the critical path when a dense tile's whole list belonged to one warp) and multiplies the number of
// units available to keep every SM sub-partition busy.
#ifndef FS_BWD_MIN_CTAS
#define FS_BWD_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(kWarps * 32, FS_BWD_MIN_CTAS)
blend_backward_kernel(const uint4* __restrict__ tile_meta, const uint2* __restrict__ seg_info,
                      const float4* __restrict__ ckpt,
                      const float4* __restrict__ final_C,
                      const uint32_t* __restrict__ n_segments, uint32_t sm_count,
                      uint32_t* __restrict__ sm_slots, uint32_t* __restrict__ work_counter,
                      const SplatRec* __restrict__ inst_splat, int W, int H, const float* __restrict__ bg_color,
                      const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                      const float* __restrict__ dL_dpix, float* __restrict__ grad_acc, uint32_t Rcap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpStage* stages = reinterpret_cast<WarpStage*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(WarpStage) * kWarps);

    fs::pdl_trigger();  // the per-Gaussian kernel may begin launching; it waits for this grid before reading
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    {   // keep ~2 dense units per active warp (see blend_forward.cu); surplus CTAs retire immediately
        const uint32_t dense_units = __ldg(n_segments) * 8u;
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
    uint64_t* s_full = bars + wid * 2;
    if (lane == 0) {
        fs::mbar_init(&s_full[0], 1);
        fs::mbar_init(&s_full[1], 1);
        fs::mbar_fence_init();
    }
    __syncwarp();
    uint32_t fills = 0;

    const int gx = (W + FS_TILE - 1) / FS_TILE;
    const float bg0 = __ldg(bg_color), bg1 = __ldg(bg_color + 1), bg2 = __ldg(bg_color + 2);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const uint32_t n_units = __ldg(n_segments) * 8u;
    const size_t plane = (size_t)H * W;

    for (;;) {
        uint32_t unit = 0;
        if (lane == 0) unit = atomicAdd(work_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const uint2 sg = seg_info[unit >> 3];
        const int tile = (int)sg.x;
        const uint32_t seg_lo = sg.y * FS_SEG;
        const int blk = (int)(unit & 7u);
        const int tile_x = tile % gx, tile_y = tile / gx;
        const int bx = tile_x * FS_TILE + (blk & 1) * 8, by = tile_y * FS_TILE + (blk >> 1) * 4;
        const int px = bx + (lane & 7), py = by + (lane >> 3);
        const bool inside = px < W && py < H;
        const float pxf = (float)px, pyf = (float)py;
        const float wx0 = (float)bx, wx1 = (float)min(bx + 7, W - 1), wy0 = (float)by, wy1 = (float)min(by + 3, H - 1);

        const uint4 meta = tile_meta[tile];
        uint2 range = make_uint2(meta.x, meta.y);
        if (range.y > Rcap) range = make_uint2(0u, 0u);

        const size_t pid = (size_t)py * W + px;
        const float T_final = inside ? final_T[pid] : 0.0f;
        const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
        float dpx = 0.f, dpy = 0.f, dpz = 0.f;
        if (inside) {
            dpx = dL_dpix[pid];
            dpy = dL_dpix[plane + pid];
            dpz = dL_dpix[2 * plane + pid];
        }
        const float bg_dot = bg0 * dpx + bg1 * dpy + bg2 * dpz;

        // this unit walks positions [seg_lo, seg_hi) back to front; positions >= the block's largest n_contrib
        // are never visited, so the stream starts at min(seg_hi, that)
        const uint32_t seg_hi = min(seg_lo + FS_SEG, range.y - range.x);
        uint32_t wl = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
        const uint32_t warp_last = min(wl, seg_hi);
        const int nbatches = warp_last > seg_lo ? (int)((warp_last - seg_lo + kBatch - 1) / kBatch) : 0;

        // batch k covers positions [lo_k, lo_k + cnt_k), walking down from warp_last to seg_lo
        auto batch_lo = [&](int k) { return (uint32_t)max((int)seg_lo, (int)warp_last - (k + 1) * kBatch); };
        auto batch_cnt = [&](int k) { return (warp_last - (uint32_t)k * kBatch) - batch_lo(k); };
        auto issue = [&](int k) {  // lane 0 only
            const uint32_t bytes = batch_cnt(k) * (uint32_t)sizeof(SplatRec);
            const uint32_t s = (fills + (uint32_t)k) & 1u;
            fs::mbar_expect_tx(&s_full[s], bytes);
            fs::bulk_g2s(&rec_ring[s][0], inst_splat + range.x + batch_lo(k), bytes, &s_full[s]);
        };
        if (lane == 0 && nbatches > 0) issue(0);

        float T = T_final;
        float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;  // accum_rec
        float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;  // last_color
        float last_alpha = 0.f;
        if (last_contributor > seg_hi) {
            // the pixel's list continues behind this segment: resume from the forward kernel's checkpoint
            const float4 c4 = ckpt[((size_t)meta.z + sg.y + 1) * FS_TILE_PIX +
                                   ((by - tile_y * FS_TILE) + (lane >> 3)) * FS_TILE + (bx - tile_x * FS_TILE) + (lane & 7)];
            const float4 fc = final_C[pid];
            const float inv = __fdividef(1.0f, c4.x);
            T = c4.x;
            ar0 = (fc.x - c4.y) * inv;  // colour accumulated behind the boundary
            ar1 = (fc.y - c4.z) * inv;
            ar2 = (fc.z - c4.w) * inv;
        }

        for (int kb = 0; kb < nbatches; ++kb) {
            __syncwarp();
            if (lane == 0 && kb + 1 < nbatches) issue(kb + 1);
            const uint32_t f = fills + (uint32_t)kb;
            fs::mbar_wait(&s_full[f & 1u], (f >> 1) & 1u);
            const SplatRec* rec = rec_ring[f & 1u];
            const uint32_t lo = batch_lo(kb);
            const int cnt = (int)batch_cnt(kb);
            for (int cb = ((cnt - 1) >> 5) << 5; cb >= 0; cb -= 32) {
                const int j = cb + lane;
                bool hit = false;
                if (j < cnt) {
                    const float4 q0 = rec[j].q0;
                    hit = !(q0.z < 0.0f) &&
                          !(q0.x + q0.z < wx0 || q0.x - q0.z > wx1 || q0.y + q0.w < wy0 || q0.y - q0.w > wy1);
                }
                unsigned m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    // four survivors per round, back to front
                    int jj[kGroup];
                    bool live[kGroup];
#pragma unroll
                    for (int k = 0; k < kGroup; ++k) {  // branch-free extraction, highest set bit first
                        const int bit = 31 - __clz(m);    // -1 when m == 0
                        live[k] = bit >= 0;
                        jj[k] = live[k] ? cb + bit : cb;
                        m &= ~(live[k] ? (1u << bit) : 0u);
                    }
                    // independent part: power, G, alpha for the four survivors
                    float G[kGroup], alpha[kGroup], dx[kGroup], dy[kGroup], inv[kGroup];
                    float4 q1[kGroup], q2[kGroup];
                    bool ok[kGroup];
                    bool any_ok = false;
#pragma unroll
                    for (int k = 0; k < kGroup; ++k) {
                        const int r = jj[k];
                        const float4 q0 = rec[r].q0;
                        q1[k] = rec[r].q1;
                        q2[k] = rec[r].q2;
                        dx[k] = fs::sub(q0.x, pxf);
                        dy[k] = fs::sub(q0.y, pyf);
                        const float power = fs::splat_power(dx[k], dy[k], q1[k].x, q1[k].y, q1[k].z);
                        G[k] = expf(power);
                        alpha[k] = fminf(0.99f, fs::mul(q1[k].w, G[k]));
                        ok[k] = live[k] && (lo + (uint32_t)r) < last_contributor && power <= 0.0f &&
                                alpha[k] >= 1.0f / 255.0f;
                        inv[k] = __fdividef(1.0f, 1.0f - alpha[k]);
                        any_ok |= ok[k];
                    }
                    if (!__any_sync(0xffffffffu, any_ok)) continue;
                    // sequential part: transmittance and the running "colour behind" recurrence
                    float Tk[kGroup], a0[kGroup], a1[kGroup], a2[kGroup];
#pragma unroll
                    for (int k = 0; k < kGroup; ++k) {
                        {   // predicated: no branches inside the dependent chain
                            const float oml = 1.0f - last_alpha;
                            const float n0 = last_alpha * lc0 + oml * ar0;
                            const float n1 = last_alpha * lc1 + oml * ar1;
                            const float n2 = last_alpha * lc2 + oml * ar2;
                            T = ok[k] ? T * inv[k] : T;
                            ar0 = ok[k] ? n0 : ar0;
                            ar1 = ok[k] ? n1 : ar1;
                            ar2 = ok[k] ? n2 : ar2;
                            lc0 = ok[k] ? q2[k].x : lc0;
                            lc1 = ok[k] ? q2[k].y : lc1;
                            lc2 = ok[k] ? q2[k].z : lc2;
                            last_alpha = ok[k] ? alpha[k] : last_alpha;
                        }
                        Tk[k] = T;
                        a0[k] = ar0;
                        a1[k] = ar1;
                        a2[k] = ar2;
                    }
                    // independent part: the nine partial derivatives per survivor
                    float v[32], e[kGroup];
#pragma unroll
                    for (int k = 0; k < kGroup; ++k) {
                        const float w = ok[k] ? 1.0f : 0.0f;
                        const float dchannel_dcolor = alpha[k] * Tk[k] * w;
                        float dL_dalpha = ((q2[k].x - a0[k]) * dpx + (q2[k].y - a1[k]) * dpy + (q2[k].z - a2[k]) * dpz) * Tk[k];
                        dL_dalpha += (-T_final * inv[k]) * bg_dot;
                        dL_dalpha *= w;
                        const float dL_dG = q1[k].w * dL_dalpha;
                        const float gdx = G[k] * dx[k], gdy = G[k] * dy[k];
                        const float dG_ddelx = -gdx * q1[k].x - gdy * q1[k].y;
                        const float dG_ddely = -gdy * q1[k].z - gdx * q1[k].y;
                        v[k * 8 + 0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[k * 8 + 1] = dL_dG * dG_ddely * ddely_dy;
                        v[k * 8 + 2] = -0.5f * gdx * dx[k] * dL_dG;
                        v[k * 8 + 3] = -0.5f * gdx * dy[k] * dL_dG;
                        v[k * 8 + 4] = -0.5f * gdy * dy[k] * dL_dG;
                        v[k * 8 + 5] = G[k] * dL_dalpha;
                        v[k * 8 + 6] = dchannel_dcolor * dpx;
                        v[k * 8 + 7] = dchannel_dcolor * dpy;
                        e[k] = dchannel_dcolor * dpz;
                    }
                    // transposing warp reduction: 36 values x 32 lanes in 42 shuffles (naive: 180)
                    treduce_step<16, 32>(v, lane);
                    treduce_step<8, 16>(v, lane);
                    treduce_step<4, 8>(v, lane);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 2);
                        v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
                    }
                    treduce_step<16, 4>(e, lane);
                    treduce_step<8, 2>(e, lane);
                    e[0] += __shfl_xor_sync(0xffffffffu, e[0], 4);
                    e[0] += __shfl_xor_sync(0xffffffffu, e[0], 2);
                    e[0] += __shfl_xor_sync(0xffffffffu, e[0], 1);
                    // lane (l & 3) == 0 owns values {(l & 28) .. +3} = survivor l>>3, half (l>>2)&1;
                    // lane (l & 7) == 0 also owns that survivor's ninth value
                    const int ks = lane >> 3;
                    const int js = ks == 0 ? jj[0] : ks == 1 ? jj[1] : ks == 2 ? jj[2] : jj[3];
                    const bool ls = ks == 0 ? live[0] : ks == 1 ? live[1] : ks == 2 ? live[2] : live[3];
                    if ((lane & 3) == 0 && ls) {
                        const uint32_t g = __float_as_uint(rec[js].q2.w);
                        float* dst = grad_acc + (size_t)g * 12;
                        if (v[0] != 0.f || v[1] != 0.f || v[2] != 0.f || v[3] != 0.f)
                            red_add_v4(dst + ((lane >> 2) & 1) * 4, v[0], v[1], v[2], v[3]);
                        if ((lane & 7) == 0 && e[0] != 0.f) atomicAdd(dst + 8, e[0]);
                    }
                }
            }
        }
        fills += (uint32_t)nbatches;
    }
}

// ---- fused per-Gaussian backward (cov2D -> cov3D -> scale/rot, projection, SH) -----------------------------
__global__ void __launch_bounds__(256)
preprocess_backward_kernel(int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
                           const float* __restrict__ shs, const uchar4* __restrict__ clamped,
                           const float* __restrict__ scales, const float* __restrict__ rotations,
                           float scale_modifier, const float* __restrict__ cov3Ds, const float* __restrict__ view,
                           const float* __restrict__ proj, int W, int H, float tan_fovx, float tan_fovy, float h_x,
                           float h_y, const float* __restrict__ campos, const float* __restrict__ grad_acc,
                           float* __restrict__ dL_dmean2D, float* __restrict__ dL_dopacity,
                           float* __restrict__ dL_dcolors, float* __restrict__ dL_dmeans3D,
                           float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh, float* __restrict__ dL_dscales,
                           float* __restrict__ dL_drots) {
    fs::pdl_trigger();  // a PDL-launched successor (fs_densify_stats_inc) may begin launching; it waits for this grid
    fs::pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    float gm[3] = {0, 0, 0};
    float dscale[3] = {0, 0, 0};
    float4 drot = make_float4(0, 0, 0, 0);
    float g2x = 0.f, g2y = 0.f, dop = 0.f, dcol[3] = {0, 0, 0};
    float* dsh = (M > 0) ? dL_dsh + (size_t)i * M * 3 : nullptr;
    const bool vis = radii[i] > 0;
    if (vis) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(grad_acc) + (size_t)i * 3);
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(grad_acc) + (size_t)i * 3 + 1);
        const float a8 = __ldg(grad_acc + (size_t)i * 12 + 8);
        g2x = a0.x;
        g2y = a0.y;
        const float dcx = a0.z, dcy = a0.w, dcz = a1.x;
        dop = a1.y;
        dcol[0] = a1.z;
        dcol[1] = a1.w;
        dcol[2] = a8;

        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        float cov[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) cov[k] = cov3Ds[(size_t)i * 6 + k];
        fs::Ewa e;
        fs::ewa_project(view, px, py, pz, h_x, h_y, tan_fovx, tan_fovy, cov, e);
        const float x_grad_mul = (e.txtz < -e.limx || e.txtz > e.limx) ? 0.f : 1.f;
        const float y_grad_mul = (e.tytz < -e.limy || e.tytz > e.limy) ? 0.f : 1.f;
        const float a = e.a, b = e.b, c = e.c;
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        const float T00 = e.T00, T01 = e.T01, T02 = e.T02, T10 = e.T10, T11 = e.T11, T12 = e.T12;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc);
            dcov[3] = (T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc);
            dcov[5] = (T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc);
            dcov[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
            dcov[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
            dcov[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        }
        const float V00 = cov[0], V01 = cov[1], V02 = cov[2], V11 = cov[3], V12 = cov[4], V22 = cov[5];
        const float r0x = T00 * V00 + T01 * V01 + T02 * V02, r0y = T00 * V01 + T01 * V11 + T02 * V12,
                    r0z = T00 * V02 + T01 * V12 + T02 * V22;
        const float r1x = T10 * V00 + T11 * V01 + T12 * V02, r1y = T10 * V01 + T11 * V11 + T12 * V12,
                    r1z = T10 * V02 + T11 * V12 + T12 * V22;
        const float dL_dT00 = 2 * r0x * dL_da + r1x * dL_db, dL_dT01 = 2 * r0y * dL_da + r1y * dL_db,
                    dL_dT02 = 2 * r0z * dL_da + r1z * dL_db;
        const float dL_dT10 = 2 * r1x * dL_dc + r0x * dL_db, dL_dT11 = 2 * r1y * dL_dc + r0y * dL_db,
                    dL_dT12 = 2 * r1z * dL_dc + r0z * dL_db;
        const float dL_dJ00 = view[0] * dL_dT00 + view[4] * dL_dT01 + view[8] * dL_dT02;
        const float dL_dJ02 = view[2] * dL_dT00 + view[6] * dL_dT01 + view[10] * dL_dT02;
        const float dL_dJ11 = view[1] * dL_dT10 + view[5] * dL_dT11 + view[9] * dL_dT12;
        const float dL_dJ12 = view[2] * dL_dT10 + view[6] * dL_dT11 + view[10] * dL_dT12;
        const float tz = 1.f / e.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * e.tx) * tz3 * dL_dJ02 +
                             (2 * h_y * e.ty) * tz3 * dL_dJ12;
        gm[0] = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        gm[1] = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        gm[2] = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        // 2D-mean gradient through the perspective projection (backward.cu:373-387)
        const float hx = proj[0] * px + proj[4] * py + proj[8] * pz + proj[12];
        const float hy = proj[1] * px + proj[5] * py + proj[9] * pz + proj[13];
        const float hw = proj[3] * px + proj[7] * py + proj[11] * pz + proj[15];
        const float m_w = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        gm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        gm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        gm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

        if (shs != nullptr) {  // backward.cu:20-139
            const float ox = px - campos[0], oy = py - campos[1], oz = pz - campos[2];
            const float len = sqrtf(ox * ox + oy * oy + oz * oz);
            const float x = ox / len, y = oy / len, z = oz / len;
            const uchar4 cl = clamped[i];
            const float dR[3] = {dcol[0] * (cl.x ? 0.f : 1.f), dcol[1] * (cl.y ? 0.f : 1.f),
                                 dcol[2] * (cl.z ? 0.f : 1.f)};
            const float* sh = shs + (size_t)i * M * 3;
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float dx3 = 0.f, dy3 = 0.f, dz3 = 0.f;
                const float dr = dR[ch];
#define SHC(k) sh[3 * (k) + ch]
#define DSH(k) dsh[3 * (k) + ch]
                DSH(0) = SH_C0 * dr;
                if (D > 0) {
                    DSH(1) = (-SH_C1 * y) * dr;
                    DSH(2) = (SH_C1 * z) * dr;
                    DSH(3) = (-SH_C1 * x) * dr;
                    dx3 = -SH_C1 * SHC(3);
                    dy3 = -SH_C1 * SHC(1);
                    dz3 = SH_C1 * SHC(2);
                    if (D > 1) {
                        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        DSH(4) = (SH_C2[0] * xy) * dr;
                        DSH(5) = (SH_C2[1] * yz) * dr;
                        DSH(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * dr;
                        DSH(7) = (SH_C2[3] * xz) * dr;
                        DSH(8) = (SH_C2[4] * (xx - yy)) * dr;
                        dx3 += SH_C2[0] * y * SHC(4) + SH_C2[2] * 2.f * -x * SHC(6) + SH_C2[3] * z * SHC(7) +
                               SH_C2[4] * 2.f * x * SHC(8);
                        dy3 += SH_C2[0] * x * SHC(4) + SH_C2[1] * z * SHC(5) + SH_C2[2] * 2.f * -y * SHC(6) +
                               SH_C2[4] * 2.f * -y * SHC(8);
                        dz3 += SH_C2[1] * y * SHC(5) + SH_C2[2] * 2.f * 2.f * z * SHC(6) + SH_C2[3] * x * SHC(7);
                        if (D > 2) {
                            DSH(9) = (SH_C3[0] * y * (3.f * xx - yy)) * dr;
                            DSH(10) = (SH_C3[1] * xy * z) * dr;
                            DSH(11) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dr;
                            DSH(12) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dr;
                            DSH(13) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dr;
                            DSH(14) = (SH_C3[5] * z * (xx - yy)) * dr;
                            DSH(15) = (SH_C3[6] * x * (xx - 3.f * yy)) * dr;
                            dx3 += (SH_C3[0] * SHC(9) * 3.f * 2.f * xy + SH_C3[1] * SHC(10) * yz +
                                    SH_C3[2] * SHC(11) * -2.f * xy + SH_C3[3] * SHC(12) * -3.f * 2.f * xz +
                                    SH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SHC(14) * 2.f * xz +
                                    SH_C3[6] * SHC(15) * 3.f * (xx - yy));
                            dy3 += (SH_C3[0] * SHC(9) * 3.f * (xx - yy) + SH_C3[1] * SHC(10) * xz +
                                    SH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) +
                                    SH_C3[3] * SHC(12) * -3.f * 2.f * yz + SH_C3[4] * SHC(13) * -2.f * xy +
                                    SH_C3[5] * SHC(14) * -2.f * yz + SH_C3[6] * SHC(15) * -3.f * 2.f * xy);
                            dz3 += (SH_C3[1] * SHC(10) * xy + SH_C3[2] * SHC(11) * 4.f * 2.f * yz +
                                    SH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) +
                                    SH_C3[4] * SHC(13) * 4.f * 2.f * xz + SH_C3[5] * SHC(14) * (xx - yy));
                        }
                    }
                }
#undef SHC
#undef DSH
                ddx += dx3 * dr;
                ddy += dy3 * dr;
                ddz += dz3 * dr;
            }
            for (int k = (D + 1) * (D + 1) * 3; k < M * 3; ++k) dsh[k] = 0.0f;  // inactive coefficients
            // dnormvdv (auxiliary.h:107-117)
            const float sum2 = ox * ox + oy * oy + oz * oz;
            const float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
            gm[0] += ((+sum2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv;
            gm[1] += (-ox * oy * ddx + (sum2 - oy * oy) * ddy - oz * oy * ddz) * inv;
            gm[2] += (-ox * oz * ddx - oy * oz * ddy + (sum2 - oz * oz) * ddz) * inv;
        }

        if (scales != nullptr) {  // backward.cu:278-341
            const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + i);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            const float Rg[3][3] = {{1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                                    {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                                    {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            const float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1],
                                scale_modifier * scales[3 * i + 2]};
            const float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dMt[3][3];  // dL_dMt[c][r] = dL_dM[r][c], dL_dM = 2 * M * dL_dSigma, M[c][r] = s[r] R[c][r]
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr)
                    dMt[rr][cc] = 2.0f * (s[rr] * Rg[0][rr] * dS[cc][0] + s[rr] * Rg[1][rr] * dS[cc][1] +
                                          s[rr] * Rg[2][rr] * dS[cc][2]);
#pragma unroll
            for (int k = 0; k < 3; ++k) dscale[k] = Rg[0][k] * dMt[k][0] + Rg[1][k] * dMt[k][1] + Rg[2][k] * dMt[k][2];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) dMt[k][rr] *= s[k];
            drot.x = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            drot.y = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                     4 * x * (dMt[2][2] + dMt[1][1]);
            drot.z = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                     4 * y * (dMt[2][2] + dMt[0][0]);
            drot.w = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                     4 * z * (dMt[1][1] + dMt[0][0]);
        }
    } else if (dsh != nullptr) {
        for (int k = 0; k < M * 3; ++k) dsh[k] = 0.0f;
    }
    if (vis && shs == nullptr && dsh != nullptr)
        for (int k = 0; k < M * 3; ++k) dsh[k] = 0.0f;

    dL_dmean2D[3 * i] = g2x;
    dL_dmean2D[3 * i + 1] = g2y;
    dL_dmean2D[3 * i + 2] = 0.0f;
    dL_dopacity[i] = dop;
    dL_dcolors[3 * i] = dcol[0];
    dL_dcolors[3 * i + 1] = dcol[1];
    dL_dcolors[3 * i + 2] = dcol[2];
    dL_dmeans3D[3 * i] = gm[0];
    dL_dmeans3D[3 * i + 1] = gm[1];
    dL_dmeans3D[3 * i + 2] = gm[2];
#pragma unroll
    for (int k = 0; k < 6; ++k) dL_dcov3D[(size_t)i * 6 + k] = dcov[k];
    dL_dscales[3 * i] = dscale[0];
    dL_dscales[3 * i + 1] = dscale[1];
    dL_dscales[3 * i + 2] = dscale[2];
    reinterpret_cast<float4*>(dL_drots)[i] = drot;
}

}  // namespace

void fs_launch_blend_backward_pipe(int W, int H, const float* bg, char* ws, const fs_workspace_layout& L,
                                   const float* dL_dpix, float* grad_acc, cudaStream_t stream);  // backward_pipe.cu

void fs_launch_backward(int P, int D, int M, const float* bg, int W, int H, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                        const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                        const int* radii, char* ws, const fs_workspace_layout& L, const float* dL_dpix,
                        float* dL_dmean2D, float* dL_dopacity, float* dL_dcolors, float* dL_dmean3D, float* dL_dcov3D,
                        float* dL_dsh, float* dL_dscale, float* dL_drot, cudaStream_t stream) {
    float* grad_acc = reinterpret_cast<float*>(ws + L.grad_acc);
    // work counter (256-byte slot) and the per-Gaussian accumulator are contiguous: one memset node
    cudaMemsetAsync(ws + L.bwd_counter, 0, (L.grad_acc - L.bwd_counter) + (size_t)P * 48, stream);
    if (fs_tuning("FATESPLAT_BWD_PIPE", 1)) {  // lanes own splats (backward_pipe.cu); 0 = lanes own pixels (below)
        FsStageTimer timer(FS_STAGE_BLEND_BWD, stream);
        fs_launch_blend_backward_pipe(W, H, bg, ws, L, dL_dpix, grad_acc, stream);
    } else {
        FsStageTimer timer(FS_STAGE_BLEND_BWD, stream);
        const size_t smem = sizeof(WarpStage) * kWarps + sizeof(uint64_t) * 2 * kWarps;
        static std::atomic<unsigned long long> attr_set{0};
        if (fs_first_use_on_device(attr_set))
            cudaFuncSetAttribute(blend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        auto* info = reinterpret_cast<fs_frame_info*>(ws + L.info);
        const int ctas_per_sm = fs_tuning("FATESPLAT_BWD_CTAS_PER_SM", 2);  // upper bound (100 regs/thread)
        const int grid = fs_num_sms() * ctas_per_sm;
        blend_backward_kernel<<<grid, kWarps * 32, smem, stream>>>(
            reinterpret_cast<const uint4*>(ws + L.tile_meta), reinterpret_cast<const uint2*>(ws + L.seg_info),
            reinterpret_cast<const float4*>(ws + L.ckpt),
            reinterpret_cast<const float4*>(ws + L.final_C), &info->reserved[2], (uint32_t)fs_num_sms(), reinterpret_cast<uint32_t*>(ws + L.bwd_counter + 256),
            reinterpret_cast<uint32_t*>(ws + L.bwd_counter),
            reinterpret_cast<const SplatRec*>(ws + L.inst_splat), W, H, bg,
            reinterpret_cast<const float*>(ws + L.final_T), reinterpret_cast<const uint32_t*>(ws + L.n_contrib),
            dL_dpix, grad_acc, (uint32_t)L.instance_capacity);
    }
    const float h_y = H / (2.0f * tan_fovy);
    const float h_x = W / (2.0f * tan_fovx);
    const float* cov = cov3D_precomp ? cov3D_precomp : reinterpret_cast<const float*>(ws + L.cov3D);
    const float* sh_in = colors_precomp ? nullptr : shs;
    FsStageTimer timer(FS_STAGE_PREPROCESS_BWD, stream);
    fs_launch_pdl(preprocess_backward_kernel, dim3((P + 255) / 256), dim3(256), 0, stream,
        P, D, M, means3D, radii, sh_in, reinterpret_cast<const uchar4*>(ws + L.clamped),
        cov3D_precomp ? nullptr : scales, rotations, scale_modifier, cov, viewmatrix, projmatrix, W, H, tan_fovx,
        tan_fovy, h_x, h_y, cam_pos, grad_acc, dL_dmean2D, dL_dopacity, dL_dcolors, dL_dmean3D, dL_dcov3D, dL_dsh,
        dL_dscale, dL_drot);
    fs_count_launch(2);
}
