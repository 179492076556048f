// Forward per-Gaussian stage: project, EWA covariance, conic, radius, tile rectangle, SH colour, and the
// per-tile instance histogram that replaces the reference's tiles_touched prefix sum.
//
// Replaces DGR cuda_rasterizer/forward.cu:155-256 (preprocessCUDA) [+ :20-71, :74-113, :118-152],
// auxiliary.h:139-164 (in_frustum) and rasterizer_impl.cu:54-66 (checkFrustum).
//
// HBM-bound stage.  Algorithmic bytes per launch: 52*P + (40 + 12*M)*Pv  (SURVEY 8d).
// The 12-byte-stride inputs (xyz, scale, SH at M=1) are staged through shared memory with 128-bit
// loads; everything the later stages need is written as one 48-byte SplatRec per Gaussian.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

// Cooperative load of 3 floats per element for one CTA's 256 elements, 128-bit when aligned.
__device__ __forceinline__ void stage3(const float* __restrict__ g, float* s, int base, int P) {
    const int n = min(kThreads, P - base);
    const float* src = g + (size_t)base * 3;
    if (n == kThreads && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        if (threadIdx.x < 192) reinterpret_cast<float4*>(s)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(src) + threadIdx.x);
    } else {
        for (int i = threadIdx.x; i < n * 3; i += kThreads) s[i] = __ldg(src + i);
    }
}

// SH basis evaluation in the reference's operation order (forward.cu:20-71).
// s points at this Gaussian's coefficient k=0, channel ch; consecutive coefficients are 3 floats apart.
__device__ __forceinline__ float sh_channel(int deg, const float* __restrict__ s, float x, float y, float z) {
    using namespace fs;
    float res = mul(s[0], SH_C0);
    if (deg > 0) {
        res = mad(-mul(y, SH_C1), s[3], res);
        res = mad(mul(z, SH_C1), s[6], res);
        res = mad(-mul(x, SH_C1), s[9], res);
        if (deg > 1) {
            const float xx = mul(x, x), yy = mul(y, y), zz = mul(z, z);
            const float xy = mul(y, x), yz = mul(z, y), xz = mul(z, x);
            res = mad(mul(xy, SH_C2[0]), s[12], res);
            res = mad(mul(yz, SH_C2[1]), s[15], res);
            res = mad(mul(sub(sub(add(zz, zz), xx), yy), SH_C2[2]), s[18], res);
            res = mad(mul(xz, SH_C2[3]), s[21], res);
            res = mad(mul(sub(xx, yy), SH_C2[4]), s[24], res);
            if (deg > 2) {
                const float q = sub(mad(zz, 4.0f, -xx), yy);
                res = mad(mul(mul(y, SH_C3[0]), mad(xx, 3.0f, -yy)), s[27], res);
                res = mad(mul(mul(xy, SH_C3[1]), z), s[30], res);
                res = mad(mul(mul(y, SH_C3[2]), q), s[33], res);
                res = mad(mul(mul(z, SH_C3[3]), mad(yy, -3.0f, mad(xx, -3.0f, add(zz, zz)))), s[36], res);
                res = mad(mul(q, mul(x, SH_C3[4])), s[39], res);
                res = mad(mul(sub(xx, yy), mul(z, SH_C3[5])), s[42], res);
                res = mad(mul(mul(x, SH_C3[6]), mad(yy, -3.0f, xx)), s[45], res);
            }
        }
    }
    return res;
}

__global__ void __launch_bounds__(kThreads)
preprocess_kernel(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales,
                  float scale_modifier, const float* __restrict__ rotations, const float* __restrict__ opacities,
                  const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                  const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
                  const float* __restrict__ projmatrix, const float* __restrict__ cam_pos, int W, int H,
                  float tan_fovx, float tan_fovy, float focal_x, float focal_y, int prefiltered,
                  int* __restrict__ radii, float* __restrict__ depths, float* __restrict__ cov3Ds,
                  SplatRec* __restrict__ splat, uchar4* __restrict__ clamped, ushort4* __restrict__ rect,
                  uint32_t* __restrict__ tiles_touched, uint32_t* __restrict__ tile_count,
                  fs_frame_info* __restrict__ info) {
    using namespace fs;
    __shared__ __align__(16) float s_xyz[kThreads * 3];
    __shared__ __align__(16) float s_scale[kThreads * 3];
    __shared__ __align__(16) float s_sh[kThreads * 3];
    __shared__ float s_view[16], s_proj[16], s_cam[3];

    fs::pdl_trigger();  // the tile scan may begin launching; it waits for this grid before reading
    const int base = blockIdx.x * kThreads;
    const int idx = base + threadIdx.x;
    if (threadIdx.x < 16) {
        s_view[threadIdx.x] = viewmatrix[threadIdx.x];
        s_proj[threadIdx.x] = projmatrix[threadIdx.x];
    }
    if (threadIdx.x < 3) s_cam[threadIdx.x] = cam_pos[threadIdx.x];
    stage3(means3D, s_xyz, base, P);
    if (cov3D_precomp == nullptr) stage3(scales, s_scale, base, P);
    const bool sh_staged = (colors_precomp == nullptr) && (M == 1);
    if (sh_staged) stage3(shs, s_sh, base, P);
    __syncthreads();

    bool visible = false;
    int4 hist = make_int4(0, 0, 0, 0);  // tile rectangle this lane adds to the per-tile histogram
    const int gx = (W + FS_TILE - 1) / FS_TILE, gy = (H + FS_TILE - 1) / FS_TILE;
    if (idx < P) {
        const float px = s_xyz[3 * threadIdx.x], py = s_xyz[3 * threadIdx.x + 1], pz = s_xyz[3 * threadIdx.x + 2];
        int radius = 0;
        uint32_t ntiles = 0;
        ushort4 rc = make_ushort4(0, 0, 0, 0);
        const float zv = xform(s_view, 2, px, py, pz);
        if (!(zv <= 0.2f)) {  // auxiliary.h:154 (x/y frustum test is disabled in the reference)
            const float hx = xform(s_proj, 0, px, py, pz);
            const float hy = xform(s_proj, 1, px, py, pz);
            const float hw = xform(s_proj, 3, px, py, pz);
            const float p_w = __frcp_rn(add(hw, 0.0000001f));
            const float prx = mul(hx, p_w), pry = mul(hy, p_w);
            float cov[6];
            if (cov3D_precomp != nullptr) {
#pragma unroll
                for (int k = 0; k < 6; ++k) cov[k] = __ldg(cov3D_precomp + (size_t)idx * 6 + k);
            } else {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rotations) + idx);
                cov3d_from_scale_rot(s_scale[3 * threadIdx.x], s_scale[3 * threadIdx.x + 1],
                                     s_scale[3 * threadIdx.x + 2], scale_modifier, q, cov);
#pragma unroll
                for (int k = 0; k < 6; ++k) cov3Ds[(size_t)idx * 6 + k] = cov[k];
            }
            Ewa e;
            ewa_project(s_view, px, py, pz, focal_x, focal_y, tan_fovx, tan_fovy, cov, e);
            const float det = mad(e.a, e.c, -mul(e.b, e.b));
            if (det != 0.0f) {
                const float det_inv = __frcp_rn(det);
                const float conx = mul(e.c, det_inv), cony = mul(e.b, -det_inv), conz = mul(e.a, det_inv);
                const float mid = mul(add(e.a, e.c), 0.5f);
                const float sq = __fsqrt_rn(fmaxf(mad(mid, mid, -det), 0.1f));
                const float lam = fmaxf(add(mid, sq), sub(mid, sq));
                const float rad_f = ceilf(mul(__fsqrt_rn(lam), 3.0f));
                const int rad = __float2int_rz(rad_f);
                // ndc2Pix evaluated in double with one fused multiply-add (auxiliary.h:41-44)
                const float pix_x = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)prx, 1.0), (double)W, -1.0), 0.5));
                const float pix_y = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)pry, 1.0), (double)H, -1.0), 0.5));
                int x0, y0, x1, y1;
                tile_rect(pix_x, pix_y, rad, gx, gy, x0, y0, x1, y1);
                const uint32_t nt = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
                if (nt != 0) {
                    visible = true;
                    radius = rad;
                    ntiles = nt;
                    rc = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
                    float r = 0.f, g = 0.f, b = 0.f;
                    uchar4 cl = make_uchar4(0, 0, 0, 0);
                    if (colors_precomp == nullptr) {
                        const float dx = sub(px, s_cam[0]), dy = sub(py, s_cam[1]), dz = sub(pz, s_cam[2]);
                        const float len = __fsqrt_rn(dot3(dx, dx, dy, dy, dz, dz));
                        const float vx = __fdiv_rn(dx, len), vy = __fdiv_rn(dy, len), vz = __fdiv_rn(dz, len);
                        const float* sp = sh_staged ? (s_sh + 3 * threadIdx.x) : (shs + (size_t)idx * M * 3);
                        const float c0 = sh_channel(D, sp + 0, vx, vy, vz);
                        const float c1 = sh_channel(D, sp + 1, vx, vy, vz);
                        const float c2 = sh_channel(D, sp + 2, vx, vy, vz);
                        // clamp at 0 after +0.5; flag recorded for the backward (forward.cu:64-70)
                        cl.x = (c0 < -0.5f);
                        cl.y = (c1 < -0.5f);
                        cl.z = (c2 < -0.5f);
                        r = cl.x ? 0.0f : add(c0, 0.5f);
                        g = cl.y ? 0.0f : add(c1, 0.5f);
                        b = cl.z ? 0.0f : add(c2, 0.5f);
                    } else {
                        r = __ldg(colors_precomp + (size_t)idx * 3);
                        g = __ldg(colors_precomp + (size_t)idx * 3 + 1);
                        b = __ldg(colors_precomp + (size_t)idx * 3 + 2);
                    }
                    const float o = __ldg(opacities + idx);
                    // Conservative half-extent of {alpha >= 1/255}: the ellipse  d^T conic d <= 2*ln(255*o),
                    // whose axis-aligned half-sizes are sqrt(2 tau conic.z / det_c), sqrt(2 tau conic.x / det_c).
                    // 1% + 0.5 px margin absorbs fp32 rounding of the per-pixel power; NaN/indefinite -> +inf.
                    float ex, ey;
                    if (o < (1.0f / 255.0f)) {
                        ex = ey = -1.0f;  // alpha = min(0.99, o*G) < 1/255 for every G <= 1
                    } else {
                        const float tau2 = 2.0f * logf(255.0f * o);
                        const float dc = conx * conz - cony * cony;
                        if (dc > 0.0f && conx > 0.0f && conz > 0.0f && tau2 >= 0.0f) {
                            ex = sqrtf(tau2 * conz / dc) * 1.01f + 0.5f;
                            ey = sqrtf(tau2 * conx / dc) * 1.01f + 0.5f;
                        } else {
                            ex = ey = __int_as_float(0x7f800000);
                        }
                    }
                    depths[idx] = zv;
                    SplatRec rec;
                    rec.q0 = make_float4(pix_x, pix_y, ex, ey);
                    rec.q1 = make_float4(conx, cony, conz, o);
                    rec.q2 = make_float4(r, g, b, __uint_as_float((uint32_t)idx));
                    splat[idx] = rec;
                    clamped[idx] = cl;
                    hist = make_int4(x0, y0, x1, y1);
                }
            }
        } else if (prefiltered) {
            info->reserved[0] = 1;  // reference traps here (auxiliary.h:156-160); we flag instead
        }
        radii[idx] = radius;
        tiles_touched[idx] = ntiles;
        rect[idx] = rc;
    }
    {   // per-tile instance histogram; rectangles of many tiles are spread over the warp (see scatter_kernel)
        const int hw = hist.z - hist.x, cnt = hw * (hist.w - hist.y);
        constexpr int kOwnMax = 32;
        if (cnt <= kOwnMax)
            for (int y = hist.y; y < hist.w; ++y)
                for (int x = hist.x; x < hist.z; ++x) atomicAdd(&tile_count[(size_t)(y * gx + x) * FS_CNT_STRIDE], 1u);
        unsigned big = __ballot_sync(0xffffffffu, cnt > kOwnMax);
        const int lane = threadIdx.x & 31;
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int bx0 = __shfl_sync(0xffffffffu, hist.x, src), by0 = __shfl_sync(0xffffffffu, hist.y, src);
            const int bw = __shfl_sync(0xffffffffu, hw, src), bn = __shfl_sync(0xffffffffu, cnt, src);
            for (int t = lane; t < bn; t += 32)
                atomicAdd(&tile_count[(size_t)((by0 + t / bw) * gx + bx0 + t % bw) * FS_CNT_STRIDE], 1u);
        }
    }
    const unsigned vmask = __ballot_sync(0xffffffffu, visible);
    if ((threadIdx.x & 31) == 0 && vmask) atomicAdd(&info->num_visible, (uint32_t)__popc(vmask));
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                    uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float zv = fs::xform(viewmatrix, 2, means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    present[idx] = !(zv <= 0.2f);
}

}  // namespace

void fs_launch_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                          const float* rotations, const float* opacities, const float* shs,
                          const float* cov3D_precomp, const float* colors_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, int W, int H, float tan_fovx, float tan_fovy,
                          int prefiltered, int* radii, char* ws, const fs_workspace_layout& L, cudaStream_t stream) {
    const float focal_y = H / (2.0f * tan_fovy);
    const float focal_x = W / (2.0f * tan_fovx);
    FsStageTimer timer(FS_STAGE_PREPROCESS, stream);
    preprocess_kernel<<<(P + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
        P, D, M, means3D, scales, scale_modifier, rotations, opacities, shs, cov3D_precomp, colors_precomp, viewmatrix,
        projmatrix, cam_pos, W, H, tan_fovx, tan_fovy, focal_x, focal_y, prefiltered, radii,
        reinterpret_cast<float*>(ws + L.depths), reinterpret_cast<float*>(ws + L.cov3D),
        reinterpret_cast<SplatRec*>(ws + L.splat), reinterpret_cast<uchar4*>(ws + L.clamped),
        reinterpret_cast<ushort4*>(ws + L.rect), reinterpret_cast<uint32_t*>(ws + L.tiles_touched),
        reinterpret_cast<uint32_t*>(ws + L.tile_count), reinterpret_cast<fs_frame_info*>(ws + L.info));
    fs_count_launch(1);
}

void fs_launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                            cudaStream_t stream) {
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, means3D, viewmatrix, present);
    fs_count_launch(1);
}
