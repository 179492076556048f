// Shared device helpers and workspace layout for libfatesplat (sm_100a).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fatesplat.h"

#define FS_TILE 16
#define FS_TILE_PIX 256
#ifndef FS_SEG
#define FS_SEG 256  // list positions per depth segment (unit of backward-blend parallelism along a tile's list)
#endif

// Splat record: 3 x float4 per Gaussian.
//   q0 = {mean2D.x, mean2D.y, extent.x, extent.y}   extent = conservative half-size of the region where
//                                                     alpha can reach 1/255 (<0: never contributes)
//   q1 = {conic.x, conic.y, conic.z, opacity}        == reference conic_opacity (forward.cu:253)
//   q2 = {r, g, b, bits(gaussian id)}                == reference rgb / colors_precomp (+ id for the backward)
struct __align__(16) SplatRec {
    float4 q0, q1, q2;
};
static_assert(sizeof(SplatRec) == 48, "splat record is 48 bytes");

namespace fs {

// ---- exact-sequence fp32 helpers (never contracted or re-associated by nvcc) --------------------------
// The op order mirrors the sm_100 SASS of the reference build ("Arithmetic contract").
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mad(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return __fmaf_rn(a2, b2, __fmaf_rn(a0, b0, __fmul_rn(a1, b1)));
}
// row vector [x y z 1] times column k of the flat 4x4 (DGR auxiliary.h:58-76)
__device__ __forceinline__ float xform(const float* __restrict__ m, int k, float x, float y, float z) {
    return __fadd_rn(dot3(x, m[k], y, m[4 + k], z, m[8 + k]), m[12 + k]);
}

__device__ __forceinline__ void tile_rect(float px, float py, int radius, int gx, int gy, int& x0, int& y0, int& x1,
                                          int& y1) {
    // DGR auxiliary.h:46-56 ; float->int conversions truncate and saturate (cvt.rzi.s32.f32)
    const float r = (float)radius;
    x0 = min(gx, max(0, __float2int_rz(mul(sub(px, r), 0.0625f))));
    y0 = min(gy, max(0, __float2int_rz(mul(sub(py, r), 0.0625f))));
    x1 = min(gx, max(0, __float2int_rz(mul(sub(add(add(px, r), 16.0f), 1.0f), 0.0625f))));
    y1 = min(gy, max(0, __float2int_rz(mul(sub(add(add(py, r), 16.0f), 1.0f), 0.0625f))));
}

struct Ewa {
    float T00, T01, T02, T10, T11, T12;
    float a, b, c;
    float tx, ty, tz, txtz, tytz, limx, limy;
};

// EWA projection of Sigma3 to screen space (DGR forward.cu:74-113; shared with backward.cu:144-199)
__device__ __forceinline__ void ewa_project(const float* __restrict__ view, float px, float py, float pz,
                                            float focal_x, float focal_y, float tan_fovx, float tan_fovy,
                                            const float* cov, Ewa& o) {
    const float tx = xform(view, 0, px, py, pz);
    const float ty = xform(view, 1, px, py, pz);
    const float tz = xform(view, 2, px, py, pz);
    const float limx = mul(tan_fovx, 1.3f), limy = mul(tan_fovy, 1.3f);
    const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
    const float cx = fminf(fmaxf(txtz, -limx), limx);
    const float cy = fminf(fmaxf(tytz, -limy), limy);
    const float tz2 = mul(tz, tz);
    const float J00 = __fdiv_rn(focal_x, tz);
    const float J02 = __fdiv_rn(mul(mul(tz, -cx), focal_x), tz2);
    const float J11 = __fdiv_rn(focal_y, tz);
    const float J12 = __fdiv_rn(mul(mul(tz, -cy), focal_y), tz2);
    const float m0 = view[0], m1 = view[1], m2 = view[2], m4 = view[4], m5 = view[5], m6 = view[6], m8 = view[8],
                m9 = view[9], m10 = view[10];
    o.T00 = mad(m2, J02, mul(m0, J00));
    o.T01 = mad(m6, J02, mul(m4, J00));
    o.T02 = mad(m10, J02, mul(m8, J00));
    o.T10 = mad(m2, J12, mul(m1, J11));
    o.T11 = mad(m6, J12, mul(m5, J11));
    o.T12 = mad(m10, J12, mul(m9, J11));
    const float c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3], c4 = cov[4], c5 = cov[5];
    const float A0_0 = dot3(o.T00, c0, o.T01, c1, o.T02, c2);
    const float A1_0 = dot3(o.T00, c1, o.T01, c3, o.T02, c4);
    const float A2_0 = dot3(o.T00, c2, o.T01, c4, o.T02, c5);
    const float A0_1 = dot3(o.T10, c0, o.T11, c1, o.T12, c2);
    const float A1_1 = dot3(o.T10, c1, o.T11, c3, o.T12, c4);
    const float A2_1 = dot3(o.T10, c2, o.T11, c4, o.T12, c5);
    o.a = add(dot3(o.T00, A0_0, o.T01, A1_0, o.T02, A2_0), 0.3f);
    o.b = dot3(o.T00, A0_1, o.T01, A1_1, o.T02, A2_1);
    o.c = add(dot3(o.T10, A0_1, o.T11, A1_1, o.T12, A2_1), 0.3f);
    o.tx = mul(cx, tz);
    o.ty = mul(cy, tz);
    o.tz = tz;
    o.txtz = txtz;
    o.tytz = tytz;
    o.limx = limx;
    o.limy = limy;
}

// Sigma3 = (S R)^T (S R) from scale and UN-normalised quaternion (r,x,y,z)  (DGR forward.cu:118-152)
__device__ __forceinline__ void cov3d_from_scale_rot(float s0, float s1, float s2, float mod, float4 q, float* cov) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float t_xz = mul(x, z), t_rx = mul(r, x), t_rz = mul(r, z), t_yy = mul(y, y), t_zz = mul(z, z);
    const float A = mad(r, y, t_xz), B = mad(-r, y, t_xz), C = mad(y, z, -t_rx), D = mad(y, z, t_rx);
    const float E = mad(x, y, -t_rz), F = mad(x, y, t_rz);
    const float G = mad(x, x, t_yy), Hh = add(t_yy, t_zz), I = mad(x, x, t_zz);
    const float R00 = sub(1.0f, add(Hh, Hh)), R11 = sub(1.0f, add(I, I)), R22 = sub(1.0f, add(G, G));
    const float sx = mul(s0, mod), sy = mul(s1, mod), sz = mul(s2, mod);
    const float m00 = mul(sx, R00), m01 = mul(sy, add(E, E)), m02 = mul(sz, add(A, A));
    const float m10 = mul(sx, add(F, F)), m11 = mul(sy, R11), m12 = mul(sz, add(C, C));
    const float m20 = mul(sx, add(B, B)), m21 = mul(sy, add(D, D)), m22 = mul(sz, R22);
    cov[0] = dot3(m00, m00, m01, m01, m02, m02);
    cov[1] = dot3(m00, m10, m01, m11, m02, m12);
    cov[2] = dot3(m00, m20, m01, m21, m02, m22);
    cov[3] = dot3(m10, m10, m11, m11, m12, m12);
    cov[4] = dot3(m10, m20, m11, m21, m12, m22);
    cov[5] = dot3(m20, m20, m21, m21, m22, m22);
}

// power of the 2D Gaussian at pixel offset d (DGR forward.cu:333-336), in the reference's contraction order
__device__ __forceinline__ float splat_power(float dx, float dy, float cx, float cy, float cz) {
    const float q = mad(dx, mul(dx, cx), mul(dy, mul(dy, cz)));
    return mad(q, -0.5f, -mul(dy, mul(dx, cy)));
}

// ---- mbarrier / bulk-async-copy (TMA 1-D) primitives ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (UBLKCP); bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with fs_launch_pdl may start while its predecessor in
// the stream is still draining; it must execute pdl_wait() before touching anything the predecessor wrote.
// pdl_trigger() lets the successor's launch begin early.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace fs

// ---- host side -------------------------------------------------------------------------------------------
struct FsLayout : fs_workspace_layout {};
void fs_compute_layout(int P, int W, int H, size_t Rcap, fs_workspace_layout* L);
void fs_set_error(const char* fmt, ...);
void fs_count_launch(int n);
int fs_num_sms();
// true exactly once per (call site, current device): cudaFuncSetAttribute is a per-device setting
bool fs_first_use_on_device(std::atomic<unsigned long long>& mask);
uint32_t fs_tile_hint();
int fs_tuning(const char* env_name, int default_value);  // integer tuning knob, read once from the environment

// Optional per-stage timing (fs_profile_enable): CUDA events recorded on the launching stream around a stage.
enum FsStage {
    FS_STAGE_PREPROCESS = 0,
    FS_STAGE_TILE_SCAN,
    FS_STAGE_SCATTER,
    FS_STAGE_TILE_SORT,
    FS_STAGE_BIG_TILE_SORT,
    FS_STAGE_BLEND_FWD,
    FS_STAGE_BLEND_BWD,
    FS_STAGE_PREPROCESS_BWD,
    FS_STAGE_KNN,
    FS_STAGE_POSE_FWD,
    FS_STAGE_POSE_BWD,
    FS_STAGE_FLAME_FWD,
    FS_STAGE_FLAME_BWD,
    FS_STAGE_EXCHANGE,
    FS_STAGE_COUNT
};
struct FsStageTimer {  // RAII: records start in the ctor and stop in the dtor when profiling is on
    FsStageTimer(int stage, cudaStream_t stream);
    ~FsStageTimer();
    int slot;
    cudaStream_t stream;
};

// Launch `kernel` so that it may overlap the tail of the previous kernel on `stream` (see fs::pdl_wait).
template <typename... KArgs, typename... Args>
inline cudaError_t fs_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = fs_tuning("FATESPLAT_PDL", 1) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Work counters of the blend kernels: uint32 words at workspace offset info + sizeof(fs_frame_info) + 1024 (forward,
// cleared with the header) and at bwd_counter + 16 bytes (backward, cleared with the backward's work counter).
#define FS_WORK_FWD_OFFSET (sizeof(fs_frame_info) + 1024)
#define FS_WORK_FWD_BOX 0    // (block, instance) pairs that passed the forward's box cull: each costs 32 exp evaluations
#define FS_WORK_BWD_OFFSET 16
#define FS_WORK_BWD_SPLATS 0 // (block, instance) pairs the backward pipeline took in (non-zero pair mask): 32 exp each
#define FS_WORK_BWD_PAIRS 1  // (pixel, instance) pairs that were blended == pairs with a gradient contribution

#ifndef FS_SORT_SMEM_CAP
#define FS_SORT_SMEM_CAP 4096  // instances a tile may hold to be sorted by the one-CTA-per-tile kernel
#endif
// Per-tile counters live 128 bytes apart: L2 atomics to one line serialise, and only a few hundred tiles are hot.
#define FS_CNT_STRIDE 32
