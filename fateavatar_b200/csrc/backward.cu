// Backward of the splat render: the fused per-Gaussian backward and the launcher of the two backward kernels.
//
// Replaces DGR cuda_rasterizer/backward.cu:144-274 (computeCov2DCUDA), :346-396 (preprocessCUDA bwd) [+ :20-139 SH,
// :278-341 cov3D] and the nine torch::zeros of rasterize_points.cu:151-159.  The blend backward (renderCUDA bwd,
// backward.cu:399-557) is backward_pipe.cu.
//
// The hand-derived gradient keeps the reference's deviations from "autograd of the forward":
// no zeroing at the alpha = 0.99 clamp, 1/(det^2 + 1e-7), frustum-clamp masks only on dL/dt.x, dL/dt.y, no
// quaternion-normalisation Jacobian, SH clamp via the saved flags, and pixels skip instances at positions >= n_contrib.
//
// The blend backward leaves nine sums per Gaussian in a 48-byte accumulator (grad_acc, three red.global.add.v4.f32
// worth of data per (block, Gaussian)); the per-Gaussian kernel below consumes it and writes every API gradient
// exactly once, so no output tensor needs a zero-fill.
#include "common.cuh"

namespace {

__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};


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

void fs_launch_blend_backward_pipe(int P, int W, int H, const float* bg, char* ws, const fs_workspace_layout& L,
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
    {
        FsStageTimer timer(FS_STAGE_BLEND_BWD, stream);
        fs_launch_blend_backward_pipe(P, W, H, bg, ws, L, dL_dpix, grad_acc, stream);
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
