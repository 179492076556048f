// Blend backward, splat-owner pipeline: the lanes of a warp own SPLATS, the pixels of an 8x4 block flow through them.
//
// Replaces DGR cuda_rasterizer/backward.cu:399-557 (renderCUDA backward).  Same pairs as the reference (a pixel
// skips positions >= n_contrib, power > 0, alpha < 1/255; no zeroing at the 0.99 clamp), same nine sums per
// Gaussian, written into the same 48-byte accumulator that preprocess_backward_kernel (backward.cu) consumes.
//
// Why: with lanes owning pixels (the reference's layout, and this repo's round-1 kernel) every (block, splat) pair ends
// in a cross-lane sum of nine values -- nine global atomics per pixel in the reference, a 36-value x 32-lane
// transposing reduction (42 shuffles, ~60 selects/adds, ~38 % of all instructions) in round 1.  Here a lane keeps ONE
// splat for 32 consecutive steps and sees the block's 32 pixels one after the other, so the nine sums accumulate in
// the lane's own registers and no cross-lane reduction exists at all.
//
//   * Schedule.  At step t lane l works on pixel (t - l) mod 32.  A pixel's running state (transmittance T and
//     the colour S accumulated in front of it) is handed from lane l to lane l+1 by one rotate (4 shuffles per
//     step) and wraps from lane 31 to lane 0 for the next batch of 32 splats, so the pixel meets the splats in
//     list order.  The pipeline never drains between batches or between work units: lane l swaps its finished
//     splat for the next batch's at step l of each 32-step window (its sums go to a per-lane slot in shared
//     memory and are flushed by all lanes together after the window: two red.global.add.v4.f32 and one scalar
//     reduction per (block, splat), skipped when zero).  The first splat of a work unit carries a flag: its lane
//     takes the pixels' state from the unit's table instead of from its neighbour.
//   * Front to back.  The walk runs in the forward's direction with the forward's own arithmetic, so T and S are
//     bit-identical to what the forward pass saw; the reference's "colour behind" recurrence is replaced by
//     (C_final - S_i) / (1 - alpha_i), with C_final from the forward (final_C).  A unit that starts at a depth
//     segment > 0 resumes from the forward's checkpoint (T, S) at that list position.
//   * Which pairs.  The forward blend left, per (block, list position), the 32-bit mask of the pixels that blended
//     the instance (pair_mask).  A splat enters the pipeline iff its mask is non-zero and a pair contributes iff
//     its bit is set: the backward never re-derives (and can never disagree with) the forward's decisions.
//   * Work units are (tile, depth segment of FS_SEG list positions, 8x4 block), pulled from an atomic work counter by
//     persistent warps (full segments first, binning.cu); the tile's sorted 48-byte records arrive by cp.async.bulk
//     (UBLKCP) into a per-warp 2-stage ring guarded by mbarriers; selected records are compacted into a per-warp queue in shared memory that
//     runs on across units (three pixel tables rotate; a table is reused once every splat that refers to it has
//     left the pipeline).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kChunk = 32;  // records per TMA stage = one selection round
constexpr int kQueue = 96;  // queue slots: previous batch (awaiting its flush), current batch, the one being filled
constexpr int kTabs = 3;    // pixel tables in rotation
constexpr uint32_t kNullId = 0xffffffffu;
constexpr uint32_t kTabBytes = 3 * 32 * 16;
constexpr uint32_t kFirstFlag = 0x80000000u;
#ifndef FS_PIPE_FAST_EXP
#define FS_PIPE_FAST_EXP 1
#endif

struct __align__(16) WarpSmem {
    SplatRec ring[2][kChunk];  // TMA destination
    SplatRec queue[kQueue];    // selected splats: q0 = {mean2D.x, mean2D.y, bits(pixel mask), bits(table offset | first-
                               // of-unit flag)}, q1 = conic + opacity, q2 = {r, g, b, bits(Gaussian id) or kNullId}
    float4 tab[kTabs][3][32];  // per unit and pixel: [0] {dL/dpix r,g,b, K}  [1] {px, py, -, -}
                               //                     [2] {T, S r,g,b} at the unit's first list position
    float4 finA[32], finB[32];
    float finC[32];
};

struct Pipe {
    float mx, my, cx, cy, cz, op, c0, c1, c2;
    uint32_t mask;  // bit 0 = the pixel of the current step
    uint32_t tab;   // byte offset of the pixel table of the splat's unit
    bool first;     // first splat of its unit: pixels enter with the unit's initial state
    float acc[9];   // sum g dx, sum g dy, sum g dx dx, sum g dx dy, sum g dy dy, sum g, sum w dL/dpix rgb
    float oT, oS0, oS1, oS2;  // state this lane hands to lane+1 at the next step
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ SplatRec null_record() {
    SplatRec r;
    r.q0 = make_float4(0.f, 0.f, 0.f, 0.f);  // mask 0, table 0 (always finite), not a first splat
    r.q1 = make_float4(0.f, 0.f, 0.f, 0.f);
    r.q2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(kNullId));
    return r;
}

// One 32-step window.  nxt = this lane's splat of the batch that enters the pipeline in this window.
template <bool ANYFIRST>
__device__ __forceinline__ void run_window(Pipe& P, WarpSmem* ws, const SplatRec* __restrict__ nxt, int lane) {
    const char* tabs = reinterpret_cast<const char*>(&ws->tab[0][0][0]);
    const int src = (lane + 31) & 31;
    uint32_t poff = ((32u - (uint32_t)lane) & 31u) << 4;  // 16 * pixel index
#pragma unroll 2
    for (int tm = 0; tm < 32; ++tm) {
        float T = __shfl_sync(0xffffffffu, P.oT, src);
        float S0 = __shfl_sync(0xffffffffu, P.oS0, src);
        float S1 = __shfl_sync(0xffffffffu, P.oS1, src);
        float S2 = __shfl_sync(0xffffffffu, P.oS2, src);
        if (tm == lane) {  // this lane's splat has seen all 32 pixels: park its sums, take the next one
            ws->finA[lane] = make_float4(P.acc[0], P.acc[1], P.acc[2], P.acc[3]);
            ws->finB[lane] = make_float4(P.acc[4], P.acc[5], P.acc[6], P.acc[7]);
            ws->finC[lane] = P.acc[8];
#pragma unroll
            for (int k = 0; k < 9; ++k) P.acc[k] = 0.0f;
            const float4 q0 = nxt->q0, q1 = nxt->q1, q2 = nxt->q2;
            P.mx = q0.x;
            P.my = q0.y;
            P.mask = __float_as_uint(q0.z);
            P.tab = __float_as_uint(q0.w) & ~kFirstFlag;
            P.first = (__float_as_uint(q0.w) & kFirstFlag) != 0u;
            P.cx = q1.x;
            P.cy = q1.y;
            P.cz = q1.z;
            P.op = q1.w;
            P.c0 = q2.x;
            P.c1 = q2.y;
            P.c2 = q2.z;
        }
        if (ANYFIRST) {
            if (P.first) {
                const float4 s = *reinterpret_cast<const float4*>(tabs + P.tab + 1024u + poff);
                T = s.x;
                S0 = s.y;
                S1 = s.z;
                S2 = s.w;
            }
        }
        const float4 pa = *reinterpret_cast<const float4*>(tabs + P.tab + poff);
        const float2 pb = *reinterpret_cast<const float2*>(tabs + P.tab + 512u + poff);
        const float dx = fs::sub(P.mx, pb.x), dy = fs::sub(P.my, pb.y);
        const float power = fs::splat_power(dx, dy, P.cx, P.cy, P.cz);
#if FS_PIPE_FAST_EXP
        // which pairs contribute is the forward's decision (mask), so G only enters values: ex2.approx (2 ulp) is enough
        float G;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(G) : "f"(power * 1.4426950408889634f));
#else
        const float G = expf(power);
#endif
        const float alpha = fminf(0.99f, fs::mul(P.op, G));
        const bool ok = (P.mask & 1u) != 0u;  // the forward blended this pair
        P.mask >>= 1;
        const float a = ok ? alpha : 0.0f;
        // the forward's own update (blend_forward.cu: apply_group): identical T and S
        const float oma = fs::sub(1.0f, a);
        const float nS0 = fs::mad(T, fs::mul(a, P.c0), S0);
        const float nS1 = fs::mad(T, fs::mul(a, P.c1), S1);
        const float nS2 = fs::mad(T, fs::mul(a, P.c2), S2);
        const float nT = fs::mul(T, oma);
        // dL/dalpha = T (c . dpix) - [ (C_final - S') . dpix + T_final bg . dpix ] / (1 - alpha)
        float inv;  // oma >= 0.01: no range handling needed
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(oma));
        const float cdp = P.c0 * pa.x + P.c1 * pa.y + P.c2 * pa.z;
        const float sdp = nS0 * pa.x + nS1 * pa.y + nS2 * pa.z;
        const float dLda = T * cdp - inv * (pa.w - sdp);
        const float g = ok ? G * dLda : 0.0f;
        const float w = a * T;
        const float gx = g * dx, gy = g * dy;
        P.acc[0] += gx;
        P.acc[1] += gy;
        P.acc[2] += gx * dx;
        P.acc[3] += gx * dy;
        P.acc[4] += gy * dy;
        P.acc[5] += g;
        P.acc[6] += w * pa.x;
        P.acc[7] += w * pa.y;
        P.acc[8] += w * pa.z;
        P.oT = nT;
        P.oS0 = nS0;
        P.oS1 = nS1;
        P.oS2 = nS2;
        poff = (poff + 16u) & 496u;
    }
}

// Sums of the batch that left the pipeline during the last window -> the per-Gaussian accumulator.
// Layout (consumed by backward.cu): [0]=dmean2D.x [1]=dmean2D.y [2]=dconic.x [3]=dconic.y [4]=dconic.w [5]=dopacity [6..8]=dcolor
__device__ __forceinline__ void flush_batch(const WarpSmem* ws, uint32_t slot0, int lane, float* __restrict__ grad_acc,
                                            uint32_t n_gaussians, float ddelx_dx, float ddely_dy) {
    const SplatRec* pr = &ws->queue[slot0 + lane];
    const uint32_t id = __float_as_uint(pr->q2.w);
    if (id >= n_gaussians) return;  // null splat (kNullId); also keeps a corrupt record from writing outside grad_acc
    const float4 q1 = pr->q1;
    const float4 fa = ws->finA[lane], fb = ws->finB[lane];
    const float fc = ws->finC[lane];
    const float o = q1.w, ho = -0.5f * q1.w;
    const float v0 = -(q1.x * fa.x + q1.y * fa.y) * o * ddelx_dx;
    const float v1 = -(q1.z * fa.y + q1.y * fa.x) * o * ddely_dy;
    const float v2 = ho * fa.z, v3 = ho * fa.w, v4 = ho * fb.x;
    float* dst = grad_acc + (size_t)id * 12;
    if (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f) red_add_v4(dst, v0, v1, v2, v3);
    if (v4 != 0.f || fb.y != 0.f || fb.z != 0.f || fb.w != 0.f) red_add_v4(dst + 4, v4, fb.y, fb.z, fb.w);
    if (fc != 0.f) atomicAdd(dst + 8, fc);
}

__global__ void __launch_bounds__(kWarps * 32, 2)
blend_backward_pipe_kernel(const uint4* __restrict__ tile_meta, const uint2* __restrict__ seg_info,
                           const float4* __restrict__ ckpt, const float4* __restrict__ final_C,
                           const uint32_t* __restrict__ n_segments, uint32_t sm_count, uint32_t* __restrict__ sm_slots,
                           uint32_t* __restrict__ work_counter, const SplatRec* __restrict__ inst_splat,
                           const uint32_t* __restrict__ pair_mask, int W, int H,
                           const float* __restrict__ bg_color, const float* __restrict__ final_T,
                           const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                           float* __restrict__ grad_acc, uint32_t n_gaussians, uint32_t Rcap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    fs::pdl_trigger();  // the per-Gaussian kernel may begin launching; it waits for this grid before reading
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    {   // same placement-independent CTA budget as blend_forward.cu
        const uint32_t dense_units = __ldg(n_segments) * 8u;
        const uint32_t want_per_sm = max(1u, dense_units / (2u * kWarps * sm_count));
        __shared__ uint32_t s_rank;
        if (threadIdx.x == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_rank = atomicAdd(&sm_slots[smid & 255u], 1u);
        }
        __syncthreads();
        if (s_rank >= want_per_sm) return;
    }
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(smem_raw) + wid;
    uint64_t* s_full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(WarpSmem) * kWarps) + wid * 2;
    if (lane == 0) {
        fs::mbar_init(&s_full[0], 1);
        fs::mbar_init(&s_full[1], 1);
        fs::mbar_fence_init();
    }
#pragma unroll
    for (int b = 0; b < kTabs; ++b)
#pragma unroll
        for (int k = 0; k < 3; ++k) ws->tab[b][k][lane] = make_float4(0.f, 0.f, 0.f, 0.f);  // finite for null splats
    __syncwarp();

    Pipe P;
    P.mx = P.my = P.cx = P.cy = P.cz = P.op = P.c0 = P.c1 = P.c2 = 0.0f;
    P.mask = 0u;
    P.tab = 0u;
    P.first = false;
#pragma unroll
    for (int k = 0; k < 9; ++k) P.acc[k] = 0.0f;
    P.oT = P.oS0 = P.oS1 = P.oS2 = 0.0f;

    uint32_t fills = 0;      // TMA stages waited for so far: stage = fills & 1, parity = (fills >> 1) & 1
    uint32_t head = 0;       // first queue slot of the batch being filled: 0, 32 or 64
    uint32_t pend = 0;       // splats queued behind `head`
    uint32_t prev = 0;       // first slot of the batch that is in the pipeline
    bool prev_valid = false;
    uint32_t windows = 0;    // windows run so far == index of the batch being filled
    uint32_t tab_next = 0;   // table buffer of the next unit that queues a splat
    uint32_t dead0 = 0, dead1 = 0, dead2 = 0;  // table b may be rewritten once `windows` >= dead_b

    const int gx = (W + FS_TILE - 1) / FS_TILE;
    const float bg0 = __ldg(bg_color), bg1 = __ldg(bg_color + 1), bg2 = __ldg(bg_color + 2);
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    const uint32_t n_units = __ldg(n_segments) * 8u;
    const size_t plane = (size_t)H * W;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // the unit being streamed (all warp-uniform)
    int kb = 0, nchunks = 0;           // next chunk / chunks of the unit
    uint32_t seg_lo = 0, warp_last = 0;
    const SplatRec* rec_base = inst_splat;   // the tile's first record
    const uint32_t* mask_in = pair_mask;     // the (tile, block)'s first mask word
    uint32_t mk_next = 0;              // this lane's mask word of chunk kb (prefetched)
    uint32_t my_tab = 0, tb = 0;
    bool first = true;                 // the unit has not queued a splat yet
    bool all_pulled = false, drained = false;
    uint32_t pending_unit = 0;         // a unit that was pulled but had to wait for a pixel table
    bool have_pending = false;
    uint32_t n_splats = 0, n_pairs = 0;  // work counters: splats taken in (uniform), blended pairs (per lane)

    auto issue = [&](int k) {  // lane 0 only: TMA of chunk k of the current unit into the ring
        const uint32_t lo = seg_lo + (uint32_t)k * kChunk;
        const uint32_t bytes = min((uint32_t)kChunk, warp_last - lo) * (uint32_t)sizeof(SplatRec);
        const uint32_t st = (fills + (uint32_t)(k - kb)) & 1u;
        fs::mbar_expect_tx(&s_full[st], bytes);
        fs::bulk_g2s(&ws->ring[st][0], rec_base + lo, bytes, &s_full[st]);
    };

    // One iteration = fill the queue up to a batch (pulling units, streaming chunks), then run ONE window: the window
    // code exists once (instruction-cache footprint), whatever made the batch complete.
    for (;;) {
        bool table_busy = false;
        while (pend < 32u && !all_pulled && !table_busy) {
            if (kb >= nchunks) {  // ---- next unit ----
                uint32_t unit = pending_unit;
                if (!have_pending) {
                    if (lane == 0) unit = atomicAdd(work_counter, 1u);
                    unit = __shfl_sync(0xffffffffu, unit, 0);
                }
                if (unit >= n_units) {
                    all_pulled = true;
                    break;
                }
                const uint2 sg = seg_info[unit >> 3];
                const int tile = (int)sg.x;
                const int blk = (int)(unit & 7u);
                const int tile_x = tile % gx, tile_y = tile / gx;
                const int bx = tile_x * FS_TILE + (blk & 1) * 8, by = tile_y * FS_TILE + (blk >> 1) * 4;
                const int px = bx + (lane & 7), py = by + (lane >> 3);
                const bool inside = px < W && py < H;
                const uint4 meta = tile_meta[tile];
                uint2 range = make_uint2(meta.x, meta.y);
                if (range.y > Rcap) range = make_uint2(0u, 0u);  // overflowed frame: flagged in the header
                const uint32_t total = range.y - range.x;
                const size_t pid = (size_t)py * W + px;
                const uint32_t last_contributor = inside ? n_contrib[pid] : 0u;
                // positions >= the block's largest n_contrib were never blended (and their masks never written)
                uint32_t wl = last_contributor;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
                const uint32_t lo0 = sg.y * FS_SEG;
                const uint32_t wlast = min(wl, min(lo0 + (uint32_t)FS_SEG, total));
                if (wlast <= lo0) {
                    have_pending = false;
                    continue;
                }
                // the unit's pixel table goes into the next buffer of the rotation, once nothing refers to it
                tb = tab_next;
                if (windows < (tb == 0u ? dead0 : (tb == 1u ? dead1 : dead2))) {
                    // rare (several tiny units in a row): run the batch as it is (completed with null splats) and
                    // come back to this unit
                    pending_unit = unit;
                    have_pending = true;
                    table_busy = true;
                    continue;
                }
                have_pending = false;
                seg_lo = lo0;
                warp_last = wlast;
                nchunks = (int)((warp_last - seg_lo + kChunk - 1) / kChunk);
                kb = 0;
                rec_base = inst_splat + range.x;
                mask_in = pair_mask + (size_t)range.x * 8u + (size_t)blk * total;
                if (lane == 0) issue(0);
                mk_next = (seg_lo + (uint32_t)lane < warp_last) ? __ldg(mask_in + seg_lo + lane) : 0u;
                float dpx = 0.f, dpy = 0.f, dpz = 0.f, K = 0.f;
                float4 st = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
                if (inside) {
                    dpx = dL_dpix[pid];
                    dpy = dL_dpix[plane + pid];
                    dpz = dL_dpix[2 * plane + pid];
                    const float4 fc = final_C[pid];
                    const float T_final = final_T[pid];
                    K = fc.x * dpx + fc.y * dpy + fc.z * dpz + T_final * (bg0 * dpx + bg1 * dpy + bg2 * dpz);
                    if (sg.y > 0)  // forward's (T, S) before list position seg_lo (blend_forward.cu)
                        st = ckpt[((size_t)meta.z + sg.y) * FS_TILE_PIX +
                                  ((by - tile_y * FS_TILE) + (lane >> 3)) * FS_TILE + (bx - tile_x * FS_TILE) + (lane & 7)];
                }
                ws->tab[tb][0][lane] = make_float4(dpx, dpy, dpz, K);
                ws->tab[tb][1][lane] = make_float4((float)px, (float)py, 0.0f, 0.0f);
                ws->tab[tb][2][lane] = st;
                my_tab = tb * kTabBytes;
                first = true;
                continue;
            }
            // ---- next chunk of the unit: select by the forward's masks, compact into the queue ----
            __syncwarp();  // every lane is done with the stage that chunk kb+1 overwrites
            if (lane == 0 && kb + 1 < nchunks) issue(kb + 1);
            const uint32_t lo = seg_lo + (uint32_t)kb * kChunk;
            const uint32_t mk = mk_next;
            mk_next = (lo + kChunk + (uint32_t)lane < warp_last) ? __ldg(mask_in + lo + kChunk + lane) : 0u;
            const unsigned m = __ballot_sync(0xffffffffu, mk != 0u);
            fs::mbar_wait(&s_full[fills & 1u], (fills >> 1) & 1u);  // always: the ring's parity counts every stage
            const SplatRec* rec = ws->ring[fills & 1u];
            ++fills;
            ++kb;
            n_pairs += (uint32_t)__popc(mk);
            if (m != 0u) {
                n_splats += (uint32_t)__popc(m);
                if (mk != 0u) {
                    const uint32_t rank = (uint32_t)__popc(m & lt_mask);
                    uint32_t slot = head + pend + rank;
                    slot = slot >= (uint32_t)kQueue ? slot - kQueue : slot;
                    SplatRec* dst = &ws->queue[slot];
                    const float4 q0 = rec[lane].q0;
                    const uint32_t tag = my_tab | ((first && rank == 0u) ? kFirstFlag : 0u);
                    dst->q0 = make_float4(q0.x, q0.y, __uint_as_float(mk), __uint_as_float(tag));
                    dst->q1 = rec[lane].q1;
                    dst->q2 = rec[lane].q2;
                }
                first = false;
                pend += (uint32_t)__popc(m);
            }
            if (kb >= nchunks && !first) {
                // the unit used its table: splats of the batch being filled leave the pipeline two windows on
                // (batch b runs in window b and its splats are swapped out during window b + 1)
                const uint32_t d = windows + (pend > 32u ? 3u : 2u);
                if (tb == 0u) dead0 = d; else if (tb == 1u) dead1 = d; else dead2 = d;
                tab_next = tb == (uint32_t)(kTabs - 1) ? 0u : tb + 1u;
            }
        }
        if (all_pulled && pend == 0u) {
            if (drained) break;
            drained = true;  // one last window of null splats: the batch in the pipeline leaves it
        }
        if (pend < 32u && (uint32_t)lane >= pend) ws->queue[head + lane] = null_record();
        __syncwarp();  // queue / table stores of other lanes are visible
        {
            const SplatRec* nxt = &ws->queue[head + lane];
            const bool nf = (__float_as_uint(nxt->q0.w) & kFirstFlag) != 0u;
            if (__any_sync(0xffffffffu, nf || P.first))
                run_window<true>(P, ws, nxt, lane);
            else
                run_window<false>(P, ws, nxt, lane);
            if (prev_valid) flush_batch(ws, prev, lane, grad_acc, n_gaussians, ddelx_dx, ddely_dy);
            prev = head;
            prev_valid = true;
            head = head == 64u ? 0u : head + 32u;
            ++windows;
            pend = pend >= 32u ? pend - 32u : 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
    if (lane == 0 && n_splats) {
        uint32_t* stats = work_counter + FS_WORK_BWD_OFFSET / 4;
        atomicAdd(stats + FS_WORK_BWD_SPLATS, n_splats);
        atomicAdd(stats + FS_WORK_BWD_PAIRS, n_pairs);
    }
}

}  // namespace

void fs_launch_blend_backward_pipe(int P, int W, int H, const float* bg, char* ws, const fs_workspace_layout& L,
                                   const float* dL_dpix, float* grad_acc, cudaStream_t stream) {
    const size_t smem = sizeof(WarpSmem) * kWarps + sizeof(uint64_t) * 2 * kWarps;
    static std::atomic<unsigned long long> attr_set{0};
    if (fs_first_use_on_device(attr_set))
        cudaFuncSetAttribute(blend_backward_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto* info = reinterpret_cast<fs_frame_info*>(ws + L.info);
    const int ctas_per_sm = fs_tuning("FATESPLAT_BWD_CTAS_PER_SM", 2);
    const int grid = fs_num_sms() * ctas_per_sm;
    blend_backward_pipe_kernel<<<grid, kWarps * 32, smem, stream>>>(
        reinterpret_cast<const uint4*>(ws + L.tile_meta), reinterpret_cast<const uint2*>(ws + L.seg_info),
        reinterpret_cast<const float4*>(ws + L.ckpt), reinterpret_cast<const float4*>(ws + L.final_C),
        &info->reserved[2], (uint32_t)fs_num_sms(), reinterpret_cast<uint32_t*>(ws + L.bwd_counter + 256),
        reinterpret_cast<uint32_t*>(ws + L.bwd_counter), reinterpret_cast<const SplatRec*>(ws + L.inst_splat),
        reinterpret_cast<const uint32_t*>(ws + L.pair_mask), W, H, bg,
        reinterpret_cast<const float*>(ws + L.final_T), reinterpret_cast<const uint32_t*>(ws + L.n_contrib), dL_dpix,
        grad_acc, (uint32_t)P, (uint32_t)L.instance_capacity);
}
