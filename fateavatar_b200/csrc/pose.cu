// Per-splat pose stage: mesh vertices -> (means3D, scales, rotations, opacities) handed to the rasterizer.
//
// Fuses, per splat, what the reference runs as ~60 small torch kernels per frame:
//   model/fateavatar.py:225-240, 253-258      gathers by face_index, quaternion compose, shell offset
//   volume_rendering/mesh_compute.py:18-59    face frame (a0,a1,a2), face scale, un-normalised normal
//   volume_rendering/mesh_sampling.py:171-200 barycentric position
//   pytorch3d matrix_to_quaternion / quaternion_multiply (published 0.7.x algorithm)
//   volume_rendering/gaussian_model.py:105-128 activations exp / normalize / sigmoid
// and its hand-derived backward (d verts via 9 float atomics per splat; all per-splat parameter gradients
// written once).  HBM-bound: ~100 bytes in / 44 bytes out per splat forward.
#include "common.cuh"

namespace {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ V3 ld3(const float* p, long long i) { return mk(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }

struct Frame {  // forward intermediates of one face/splat
    V3 v0, v1, v2, e1, e2, a0, a1, a2, c, d, nrm;
    float l1, lc, ld, d11, dcc, ddd, t, fs;
};

__device__ __forceinline__ void face_frame(const float* __restrict__ verts, const long long* __restrict__ faces,
                                           long long f, Frame& F) {
    const long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    F.v0 = ld3(verts, i0);
    F.v1 = ld3(verts, i1);
    F.v2 = ld3(verts, i2);
    F.e1 = F.v1 - F.v0;
    F.e2 = F.v2 - F.v0;
    F.d11 = dot(F.e1, F.e1);
    F.l1 = sqrtf(fmaxf(F.d11, 1e-20f));
    F.a0 = F.e1 * (1.0f / F.l1);
    F.c = cross(F.a0, F.e2);
    F.dcc = dot(F.c, F.c);
    F.lc = sqrtf(fmaxf(F.dcc, 1e-20f));
    F.a1 = F.c * (1.0f / F.lc);
    F.d = cross(F.a1, F.a0);
    F.ddd = dot(F.d, F.d);
    F.ld = sqrtf(fmaxf(F.ddd, 1e-20f));
    F.a2 = F.d * (-1.0f / F.ld);
    F.t = dot(F.a2, F.e2);
    F.fs = 0.5f * (F.l1 + fabsf(F.t));
    F.nrm = cross(F.e1, F.e2);
}

// pytorch3d matrix_to_quaternion on M = [a0 a1 a2] (columns); returns the selected candidate row r and pieces
struct QFace {
    float q[4];
    float num[4];
    float qa, tr, den;
    int r;
};
__device__ __forceinline__ void mat_to_quat(const Frame& F, QFace& Q) {
    const float m00 = F.a0.x, m10 = F.a0.y, m20 = F.a0.z;
    const float m01 = F.a1.x, m11 = F.a1.y, m21 = F.a1.z;
    const float m02 = F.a2.x, m12 = F.a2.y, m22 = F.a2.z;
    const float t[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22,
                        1.0f - m00 - m11 + m22};
    float qa[4];
    int r = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        qa[k] = t[k] > 0.0f ? sqrtf(t[k]) : 0.0f;
        if (qa[k] > qa[r]) r = k;  // first maximum, as argmax
    }
    float num[4];
    if (r == 0) {
        num[0] = qa[0] * qa[0]; num[1] = m21 - m12; num[2] = m02 - m20; num[3] = m10 - m01;
    } else if (r == 1) {
        num[0] = m21 - m12; num[1] = qa[1] * qa[1]; num[2] = m10 + m01; num[3] = m02 + m20;
    } else if (r == 2) {
        num[0] = m02 - m20; num[1] = m10 + m01; num[2] = qa[2] * qa[2]; num[3] = m12 + m21;
    } else {
        num[0] = m10 - m01; num[1] = m20 + m02; num[2] = m21 + m12; num[3] = qa[3] * qa[3];
    }
    Q.r = r;
    Q.qa = qa[r];
    Q.tr = t[r];
    Q.den = 2.0f * fmaxf(qa[r], 0.1f);
    const float inv = 1.0f / Q.den;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        Q.num[k] = num[k];
        Q.q[k] = num[k] * inv;
    }
    // (pytorch3d >= 0.7.5 standardises here; the sign is irrelevant downstream because the product is standardised)
}

__global__ void __launch_bounds__(256)
pose_forward_kernel(int N, const float* __restrict__ verts, const long long* __restrict__ faces,
                    const long long* __restrict__ face_index, const float* __restrict__ bary,
                    const float* __restrict__ canon, const float* __restrict__ scaling_raw,
                    const float* __restrict__ rotation_raw, const float* __restrict__ offset_raw,
                    const float* __restrict__ opacity_raw, float shell_len, int resize_scale,
                    float* __restrict__ means3D, float* __restrict__ scales, float* __restrict__ rotations,
                    float* __restrict__ opacities) {
    fs::pdl_wait();  // launched while the FLAME skin kernel drains (fs_launch_pdl)
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long f = face_index[n];
    Frame F;
    face_frame(verts, faces, f, F);
    QFace Q;
    mat_to_quat(F, Q);
    const float b0 = bary[3 * n], b1 = bary[3 * n + 1], b2 = bary[3 * n + 2];
    const V3 pos = F.v0 * b0 + F.v1 * b1 + F.v2 * b2;
    const float lr = resize_scale ? logf(F.fs / canon[f]) : 0.0f;
    const float th = tanhf(offset_raw[n]);
    const V3 xyz = pos + F.nrm * (shell_len * th);
    means3D[3 * n] = xyz.x;
    means3D[3 * n + 1] = xyz.y;
    means3D[3 * n + 2] = xyz.z;
#pragma unroll
    for (int k = 0; k < 3; ++k) scales[3 * n + k] = expf(scaling_raw[3 * n + k] + lr);
    const float4 b = reinterpret_cast<const float4*>(rotation_raw)[n];
    const float aw = Q.q[0], ax = Q.q[1], ay = Q.q[2], az = Q.q[3];
    float ow = aw * b.x - ax * b.y - ay * b.z - az * b.w;
    float ox = aw * b.y + ax * b.x + ay * b.w - az * b.z;
    float oy = aw * b.z - ax * b.w + ay * b.x + az * b.y;
    float oz = aw * b.w + ax * b.z - ay * b.y + az * b.x;
    const float sg = ow < 0.0f ? -1.0f : 1.0f;
    ow *= sg; ox *= sg; oy *= sg; oz *= sg;
    const float inv = 1.0f / fmaxf(sqrtf(ow * ow + ox * ox + oy * oy + oz * oz), 1e-12f);
    reinterpret_cast<float4*>(rotations)[n] = make_float4(ow * inv, ox * inv, oy * inv, oz * inv);
    opacities[n] = 1.0f / (1.0f + expf(-opacity_raw[n]));
}

// d(x / max(|x|, eps)) : given dL/d(unit vector u) -> dL/dx
__device__ __forceinline__ V3 normalize_bwd(V3 u, float len, float dot_xx, V3 du) {
    if (dot_xx < 1e-20f) return du * (1.0f / len);  // clamped norm is a constant
    return (du - u * dot(u, du)) * (1.0f / len);
}

// dL/dverts[i] += v: rows are 12 bytes apart, so an even row starts 8-byte aligned (x,y as one red.v2 + z) and an odd
// row ends 8-byte aligned (x + y,z as one red.v2): two reductions per vertex instead of three
__device__ __forceinline__ void red_add3(float* __restrict__ base, long long i, V3 v) {
    float* p = base + 3 * i;
    if ((i & 1) == 0 && (reinterpret_cast<uintptr_t>(base) & 7u) == 0) {
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
        atomicAdd(p + 2, v.z);
    } else if ((reinterpret_cast<uintptr_t>(base) & 7u) == 0) {
        atomicAdd(p, v.x);
        asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p + 1), "f"(v.y), "f"(v.z) : "memory");
    } else {
        atomicAdd(p, v.x);
        atomicAdd(p + 1, v.y);
        atomicAdd(p + 2, v.z);
    }
}

__global__ void __launch_bounds__(256)
pose_backward_kernel(int N, const float* __restrict__ verts, const long long* __restrict__ faces,
                     const long long* __restrict__ face_index, const float* __restrict__ bary,
                     const float* __restrict__ canon, const float* __restrict__ scaling_raw,
                     const float* __restrict__ rotation_raw, const float* __restrict__ offset_raw,
                     const float* __restrict__ opacity_raw, float shell_len, int resize_scale,
                     const float* __restrict__ g_means3D, const float* __restrict__ g_scales,
                     const float* __restrict__ g_rotations, const float* __restrict__ g_opacities,
                     float* __restrict__ d_verts, float* __restrict__ d_scaling_raw, float* __restrict__ d_rotation_raw,
                     float* __restrict__ d_offset_raw, float* __restrict__ d_opacity_raw) {
    fs::pdl_trigger();  // the FLAME skin backward may begin launching; it waits for this grid before reading
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const long long f = face_index[n];
    Frame F;
    face_frame(verts, faces, f, F);
    QFace Q;
    mat_to_quat(F, Q);
    const float b0 = bary[3 * n], b1 = bary[3 * n + 1], b2 = bary[3 * n + 2];
    const float cf = canon[f];
    const float ratio = F.fs / cf;
    const float lr = resize_scale ? logf(ratio) : 0.0f;

    // opacity
    const float op = 1.0f / (1.0f + expf(-opacity_raw[n]));
    d_opacity_raw[n] = g_opacities[n] * op * (1.0f - op);
    // scales
    float d_logratio = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float s = expf(scaling_raw[3 * n + k] + lr);
        const float dl = g_scales[3 * n + k] * s;
        d_scaling_raw[3 * n + k] = dl;
        d_logratio += dl;
    }
    const float d_fs = resize_scale ? d_logratio / ratio / cf : 0.0f;

    // rotation: normalize -> standardize -> raw multiply
    const float4 b = reinterpret_cast<const float4*>(rotation_raw)[n];
    const float aw = Q.q[0], ax = Q.q[1], ay = Q.q[2], az = Q.q[3];
    float ow = aw * b.x - ax * b.y - ay * b.z - az * b.w;
    float ox = aw * b.y + ax * b.x + ay * b.w - az * b.z;
    float oy = aw * b.z - ax * b.w + ay * b.x + az * b.y;
    float oz = aw * b.w + ax * b.z - ay * b.y + az * b.x;
    const float sg = ow < 0.0f ? -1.0f : 1.0f;
    ow *= sg; ox *= sg; oy *= sg; oz *= sg;
    const float nq = sqrtf(ow * ow + ox * ox + oy * oy + oz * oz);
    const float4 gr = reinterpret_cast<const float4*>(g_rotations)[n];
    float dw, dx, dy, dz;
    if (nq > 1e-12f) {
        const float inv = 1.0f / nq;
        const float rw = ow * inv, rx = ox * inv, ry = oy * inv, rz = oz * inv;
        const float rg = rw * gr.x + rx * gr.y + ry * gr.z + rz * gr.w;
        dw = (gr.x - rw * rg) * inv; dx = (gr.y - rx * rg) * inv; dy = (gr.z - ry * rg) * inv; dz = (gr.w - rz * rg) * inv;
    } else {
        dw = gr.x * 1e12f; dx = gr.y * 1e12f; dy = gr.z * 1e12f; dz = gr.w * 1e12f;
    }
    dw *= sg; dx *= sg; dy *= sg; dz *= sg;
    // wrt b (= _rotation)
    reinterpret_cast<float4*>(d_rotation_raw)[n] =
        make_float4(dw * aw + dx * ax + dy * ay + dz * az, -dw * ax + dx * aw + dy * az - dz * ay,
                    -dw * ay - dx * az + dy * aw + dz * ax, -dw * az + dx * ay - dy * ax + dz * aw);
    // wrt a (= face quaternion)
    const float da[4] = {dw * b.x + dx * b.y + dy * b.z + dz * b.w, -dw * b.y + dx * b.x - dy * b.w + dz * b.z,
                         -dw * b.z + dx * b.w + dy * b.x - dz * b.y, -dw * b.w - dx * b.z + dy * b.y + dz * b.x};
    // candidate row r: q = num / den, den = 2 max(qa, 0.1), num[r] = qa^2, qa = sqrt(max(tr, 0))
    const float inv_den = 1.0f / Q.den;
    float dnum[4];
    float d_den = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dnum[k] = da[k] * inv_den;
        d_den -= da[k] * Q.num[k] * inv_den * inv_den;
    }
    float d_qa = (Q.qa > 0.1f) ? 2.0f * d_den : 0.0f;
    d_qa += 2.0f * Q.qa * dnum[Q.r];
    const float d_tr = (Q.tr > 0.0f && Q.qa > 0.0f) ? d_qa / (2.0f * Q.qa) : 0.0f;
    // dM (m_ij = a_j[i])
    float m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    const int r = Q.r;
    const float s00 = (r == 0 || r == 1) ? 1.0f : -1.0f, s11 = (r == 0 || r == 2) ? 1.0f : -1.0f,
                s22 = (r == 0 || r == 3) ? 1.0f : -1.0f;
    m[0][0] += s00 * d_tr; m[1][1] += s11 * d_tr; m[2][2] += s22 * d_tr;
    if (r == 0) {        // num = [qa^2, m21-m12, m02-m20, m10-m01]
        m[2][1] += dnum[1]; m[1][2] -= dnum[1]; m[0][2] += dnum[2]; m[2][0] -= dnum[2]; m[1][0] += dnum[3]; m[0][1] -= dnum[3];
    } else if (r == 1) { // [m21-m12, qa^2, m10+m01, m02+m20]
        m[2][1] += dnum[0]; m[1][2] -= dnum[0]; m[1][0] += dnum[2]; m[0][1] += dnum[2]; m[0][2] += dnum[3]; m[2][0] += dnum[3];
    } else if (r == 2) { // [m02-m20, m10+m01, qa^2, m12+m21]
        m[0][2] += dnum[0]; m[2][0] -= dnum[0]; m[1][0] += dnum[1]; m[0][1] += dnum[1]; m[1][2] += dnum[3]; m[2][1] += dnum[3];
    } else {             // [m10-m01, m20+m02, m21+m12, qa^2]
        m[1][0] += dnum[0]; m[0][1] -= dnum[0]; m[2][0] += dnum[1]; m[0][2] += dnum[1]; m[2][1] += dnum[2]; m[1][2] += dnum[2];
    }
    V3 d_a0 = mk(m[0][0], m[1][0], m[2][0]);
    V3 d_a1 = mk(m[0][1], m[1][1], m[2][1]);
    V3 d_a2 = mk(m[0][2], m[1][2], m[2][2]);

    // position / shell offset
    const V3 g = ld3(g_means3D, n);
    const float th = tanhf(offset_raw[n]);
    d_offset_raw[n] = dot(g, F.nrm) * shell_len * (1.0f - th * th);
    const V3 d_nrm = g * (shell_len * th);
    V3 d_e1 = cross(F.e2, d_nrm);   // nrm = e1 x e2
    V3 d_e2 = cross(d_nrm, F.e1);
    // face scale fs = (l1 + |t|)/2, t = a2 . e2
    const float d_t = (F.t > 0.0f ? 1.0f : (F.t < 0.0f ? -1.0f : 0.0f)) * 0.5f * d_fs;
    float d_l1 = 0.5f * d_fs;
    d_a2 = d_a2 + F.e2 * d_t;
    d_e2 = d_e2 + F.a2 * d_t;
    // a2 = -normalize(d), d = a1 x a0
    const V3 u = F.a2 * -1.0f;
    const V3 d_d = normalize_bwd(u, F.ld, F.ddd, d_a2 * -1.0f);
    d_a1 = d_a1 + cross(F.a0, d_d);
    d_a0 = d_a0 + cross(d_d, F.a1);
    // a1 = normalize(c), c = a0 x e2
    const V3 d_c = normalize_bwd(F.a1, F.lc, F.dcc, d_a1);
    d_a0 = d_a0 + cross(F.e2, d_c);
    d_e2 = d_e2 + cross(d_c, F.a0);
    // a0 = normalize(e1); l1 = |e1| also feeds the face scale
    d_e1 = d_e1 + normalize_bwd(F.a0, F.l1, F.d11, d_a0);
    if (F.d11 >= 1e-20f) d_e1 = d_e1 + F.a0 * d_l1;
    // vertices
    const V3 d_v0 = g * b0 - d_e1 - d_e2, d_v1 = g * b1 + d_e1, d_v2 = g * b2 + d_e2;
    const long long i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    red_add3(d_verts, i0, d_v0);
    red_add3(d_verts, i1, d_v1);
    red_add3(d_verts, i2, d_v2);
}

}  // namespace

extern "C" {

int fs_pose_forward(int N, int V, int F, const float* d_verts, const long long* d_faces,
                    const long long* d_face_index, const float* d_bary, const float* d_face_scale_canonical,
                    const float* d_scaling_raw, const float* d_rotation_raw, const float* d_offset_raw,
                    const float* d_opacity_raw, float shell_len, int resize_scale, float* d_means3D, float* d_scales,
                    float* d_rotations, float* d_opacities, void* stream) {
    if (N < 0 || V <= 0 || F <= 0) {
        fs_set_error("fs_pose_forward: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (N == 0) return FS_OK;
    if (!d_verts || !d_faces || !d_face_index || !d_bary || !d_face_scale_canonical || !d_scaling_raw ||
        !d_rotation_raw || !d_offset_raw || !d_opacity_raw || !d_means3D || !d_scales || !d_rotations || !d_opacities) {
        fs_set_error("fs_pose_forward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    FsStageTimer timer(FS_STAGE_POSE_FWD, static_cast<cudaStream_t>(stream));
    fs_launch_pdl(pose_forward_kernel, dim3((N + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream),
        N, d_verts, d_faces, d_face_index, d_bary, d_face_scale_canonical, d_scaling_raw, d_rotation_raw, d_offset_raw,
        d_opacity_raw, shell_len, resize_scale, d_means3D, d_scales, d_rotations, d_opacities);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_pose_forward: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

int fs_pose_backward(int N, int V, int F, const float* d_verts, const long long* d_faces,
                     const long long* d_face_index, const float* d_bary, const float* d_face_scale_canonical,
                     const float* d_scaling_raw, const float* d_rotation_raw, const float* d_offset_raw,
                     const float* d_opacity_raw, float shell_len, int resize_scale, const float* d_dL_dmeans3D,
                     const float* d_dL_dscales, const float* d_dL_drotations, const float* d_dL_dopacities,
                     float* d_dL_dverts, float* d_dL_dscaling_raw, float* d_dL_drotation_raw, float* d_dL_doffset_raw,
                     float* d_dL_dopacity_raw, void* stream) {
    if (N < 0 || V <= 0 || F <= 0) {
        fs_set_error("fs_pose_backward: invalid size");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_dL_dverts) {
        fs_set_error("fs_pose_backward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(d_dL_dverts, 0, (size_t)V * 3 * sizeof(float), st);
    if (N == 0) return FS_OK;
    if (!d_verts || !d_faces || !d_face_index || !d_bary || !d_face_scale_canonical || !d_scaling_raw ||
        !d_rotation_raw || !d_offset_raw || !d_opacity_raw || !d_dL_dmeans3D || !d_dL_dscales || !d_dL_drotations ||
        !d_dL_dopacities || !d_dL_dscaling_raw || !d_dL_drotation_raw || !d_dL_doffset_raw || !d_dL_dopacity_raw) {
        fs_set_error("fs_pose_backward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    FsStageTimer timer(FS_STAGE_POSE_BWD, st);
    pose_backward_kernel<<<(N + 255) / 256, 256, 0, st>>>(
        N, d_verts, d_faces, d_face_index, d_bary, d_face_scale_canonical, d_scaling_raw, d_rotation_raw, d_offset_raw,
        d_opacity_raw, shell_len, resize_scale, d_dL_dmeans3D, d_dL_dscales, d_dL_drotations, d_dL_dopacities,
        d_dL_dverts, d_dL_dscaling_raw, d_dL_drotation_raw, d_dL_doffset_raw, d_dL_dopacity_raw);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_pose_backward: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

}  // extern "C"
