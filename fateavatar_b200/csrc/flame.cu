// FLAME linear blend skinning with personalised blendshape deltas.
//
// Replaces, for batch size 1 (the only size FateAvatar uses, model/fateavatar.py:207 "bs = 1, essentially"):
//   flame/FLAME.py:156-204  forward_with_delta_blendshape  (template + delta_vertex, shapedirs + delta, posedirs + delta)
//   flame/FLAME.py:131-154  forward                        (the same pose without the deltas: verts_orig)
//   flame/lbs.py:24-100     lbs: blendshape einsum (:63), joint regression (:67,207), Rodrigues with +1e-8
//                           inside the norm (:253-270), pose correctives (:75-80), kinematic chain (:285-342),
//                           skinning (:86-98)
// which upstream is ~40 small torch kernels per call, two calls per frame, each re-reading the 24 MB shapedirs
// tensor and materialising shapedirs + delta.  Here one frame is two kernels forward and two backward:
//
//   flame_blend_kernel          one warp per vertex streams the *active* coefficient range of shapedirs and
//                               delta_shapedirs once (128-bit loads) and produces v_shaped for BOTH paths (with and
//                               without deltas) plus per-CTA partial joint sums (fixed summation order);
//   flame_skin_kernel           every CTA finishes the joints, Rodrigues and the kinematic chain (a few hundred
//                               flops, cheaper than a third launch), then adds the pose correctives (posedirs +
//                               delta read once, coalesced) and skins its 32 vertices for both paths;
//   flame_skin_backward_kernel  dL/dv_posed = T^T g and per-CTA partials of dL/dA (the gradient that reaches the
//                               joints through the kinematic chain);
//   flame_blend_backward_kernel chain backward -> dL/dJ, dL/dv_shaped = dL/dv_posed + Jreg^T dL/dJ, then the three
//                               parameter gradients: delta_vertex, delta_posedirs (rank-1: pose_feature (x) g) and
//                               delta_shapedirs (rank-1: g (x) betas, 24 MB of streaming stores -- the only part of
//                               the stage that is HBM-bound).
// plus flame_coeff_partial/final_kernel (optional gradients of the expression / pose coefficients) and
// flame_expand_kernel (multi-GPU: dense delta gradients from the all-gathered rank-1 factors).
// Everything is deterministic (no float atomics).  fp32 throughout; sums run in a different order than cuBLAS's
// so parity with the reference is to tolerance (tests/test_flame.py), not bitwise.
#include "common.cuh"

namespace {

constexpr int kMaxJ = FS_FLAME_MAX_JOINTS;
constexpr int kBlendThreads = 1024;                 // forward: 32 warps per SM, one vertex per warp per iteration
constexpr int kBlendBwdThreads = 512;
constexpr int kSkinVerts = 32, kSkinCoords = 96;  // a skin CTA owns 32 vertices = 96 coordinates
constexpr int kSkinThreads = 256;                 // 2 x 96 split the pose-corrective sum, the rest help the prologue
constexpr int kSkinBwdThreads = 96;

struct Parents {
    int p[kMaxJ];
};

// Device-resident state of one forward call, kept in the workspace for the backward.  Index 0 = path with the
// deltas, 1 = path without (verts_orig).
struct FlameState {
    float J[2][kMaxJ][3];    // regressed joints
    float R[kMaxJ][9];       // local rotations (row-major)
    float Rg[2][kMaxJ][9];   // chained rotations
    float tg[2][kMaxJ][3];   // chained translations
    float A[2][kMaxJ][12];   // rows 0..2 of the relative transforms (lbs.py:336-340)
    float pf[(kMaxJ - 1) * 9];
};

struct FlameWs {  // byte offsets inside the workspace
    size_t state, v_shaped, v_posed, g_posed, jpart, dapart, bpart, pfpart, drloc, total;
    int nblk_blend, nblk_skin;
};

FlameWs flame_layout(int V) {
    FlameWs w;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    w.nblk_blend = fs_num_sms();
    w.nblk_skin = (V + kSkinVerts - 1) / kSkinVerts;
    size_t o = 0;
    w.state = o;
    o = al(o + sizeof(FlameState));
    w.v_shaped = o;
    o = al(o + (size_t)2 * V * 3 * sizeof(float));
    w.v_posed = o;
    o = al(o + (size_t)V * 3 * sizeof(float));
    w.g_posed = o;
    o = al(o + (size_t)V * 3 * sizeof(float));
    w.jpart = o;
    o = al(o + (size_t)w.nblk_blend * 2 * kMaxJ * 3 * sizeof(float));
    w.dapart = o;
    o = al(o + (size_t)w.nblk_skin * kMaxJ * 12 * sizeof(float));
    w.bpart = o;  // coefficient-gradient partials (fs_flame_backward_coeffs): [128 columns][#SM CTAs]
    o = al(o + (size_t)fs_num_sms() * 128 * sizeof(float));
    w.pfpart = o;
    o = al(o + (size_t)8 * 64 * sizeof(float));
    w.drloc = o;
    o = al(o + (size_t)kMaxJ * 9 * sizeof(float));
    w.total = o;
    return w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of `n` per-CTA partials of one slot, laid out [slot][n]: 8 consecutive threads share a slot, every load is
// independent (one memory round trip), fixed order => deterministic.  Valid in all 8 threads.
__device__ __forceinline__ float reduce_partials8(const float* __restrict__ part, int slot, int n, int stride, int sub) {
    float s = 0.f;
    for (int b0 = 0; b0 < n; b0 += 8 * 24) {  // one pass for up to 192 producer CTAs
        float v[24];
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            const int b = b0 + sub + 8 * i;
            v[i] = b < n ? part[(size_t)slot * stride + b] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 24; ++i) s += v[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// ---- forward 1: blendshapes ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlendThreads)
flame_blend_kernel(int V, int L, int l0, int J, const float* __restrict__ betas, const float* __restrict__ v_template,
                   const float* __restrict__ delta_vertex, const float* __restrict__ shapedirs,
                   const float* __restrict__ delta_shapedirs, const float* __restrict__ J_regressor,
                   float* __restrict__ v_shaped /*[2][3V]*/, float* __restrict__ jpart /*[2*3J][grid]*/) {
    __shared__ float s_j[kBlendThreads / 32][2 * kMaxJ * 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = gridDim.x * (kBlendThreads / 32);
    const int n = L - l0;
    const bool vec = ((L | l0) & 3) == 0;  // rows start 16-byte aligned and hold whole float4s
    const int nslot = 2 * J * 3;
    float jp0 = 0.f, jp1 = 0.f;  // joint partial sums of slots lane, lane + 32
    fs::pdl_wait();     // launched while the previous frame's last kernel (or the optimiser step) drains
    fs::pdl_trigger();  // the skin kernel may start its own prologue (rotations, pose correctives) right away; only
                        // after the wait, so that nothing it reads early can still be in flight

    for (int v = blockIdx.x * (kBlendThreads / 32) + wid; v < V; v += nw) {
        float jw[2];  // this lane's joint-regressor weights, requested together with the blendshape rows
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = lane + 32 * h;
            jw[h] = slot < nslot ? __ldg(J_regressor + (size_t)((slot % (3 * J)) / 3) * V + v) : 0.0f;
        }
        float ao[3] = {0.f, 0.f, 0.f}, ad[3] = {0.f, 0.f, 0.f};  // sum beta * S, sum beta * (S + dS)
        const size_t row0 = (size_t)v * 3 * L + l0;
        if (vec) {  // the three rows of a vertex are requested together: one memory round trip per 32 float4 columns
            const float4* b4 = reinterpret_cast<const float4*>(betas + l0);
            for (int c = lane; c < (n >> 2); c += 32) {
                float4 sv[3], dv[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    sv[k] = __ldg(reinterpret_cast<const float4*>(shapedirs + row0 + (size_t)k * L) + c);
                    dv[k] = delta_shapedirs ? __ldg(reinterpret_cast<const float4*>(delta_shapedirs + row0 + (size_t)k * L) + c)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float4 b = __ldg(b4 + c);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ao[k] += b.x * sv[k].x + b.y * sv[k].y + b.z * sv[k].z + b.w * sv[k].w;
                    ad[k] += b.x * (sv[k].x + dv[k].x) + b.y * (sv[k].y + dv[k].y) + b.z * (sv[k].z + dv[k].z) +
                             b.w * (sv[k].w + dv[k].w);
                }
            }
        } else {
            for (int c = lane; c < n; c += 32) {
                const float b = __ldg(betas + l0 + c);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float sv = __ldg(shapedirs + row0 + (size_t)k * L + c);
                    ao[k] += b * sv;
                    if (delta_shapedirs) ad[k] += b * (sv + __ldg(delta_shapedirs + row0 + (size_t)k * L + c));
                }
            }
        }
        const float tv = lane < 3 ? __ldg(v_template + 3 * v + lane) : 0.0f;
        const float dvv = (delta_vertex && lane < 3) ? __ldg(delta_vertex + 3 * v + lane) : 0.0f;
        float vs[2][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float so = warp_sum(ao[k]);
            const float sd = delta_shapedirs ? warp_sum(ad[k]) : so;
            const float t = __shfl_sync(0xffffffffu, tv, k), dd = __shfl_sync(0xffffffffu, dvv, k);
            vs[0][k] = (delta_vertex ? t + dd : t) + sd;
            vs[1][k] = t + so;
        }
        if (lane < 6) {
            const int path = lane / 3, k = lane % 3;
            const float val = k == 0 ? vs[path][0] : k == 1 ? vs[path][1] : vs[path][2];
            v_shaped[(size_t)path * 3 * V + 3 * v + k] = val;
        }
        // joint regression (lbs.py:207): slot = path * 3J + j * 3 + k
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = lane + 32 * h;
            if (slot < nslot) {
                const int path = slot / (3 * J), k = (slot % (3 * J)) % 3;
                const float a = path ? (k == 0 ? vs[1][0] : k == 1 ? vs[1][1] : vs[1][2])
                                     : (k == 0 ? vs[0][0] : k == 1 ? vs[0][1] : vs[0][2]);
                if (h == 0) jp0 += jw[0] * a; else jp1 += jw[1] * a;
            }
        }
    }
    if (lane < nslot) s_j[wid][lane] = jp0;
    if (lane + 32 < nslot) s_j[wid][lane + 32] = jp1;
    __syncthreads();
    if (threadIdx.x < nslot) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kBlendThreads / 32; ++w) s += s_j[w][threadIdx.x];
        jpart[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;  // [slot][CTA]: the consumer reads along CTAs
    }
}

// Rodrigues exactly as lbs.py:253-270: angle = ||r + 1e-8||, K from r / angle, R = I + sin K + (1 - cos) K K.
__device__ void rodrigues(const float* r, float* R) {
    const float ax = r[0] + 1e-8f, ay = r[1] + 1e-8f, az = r[2] + 1e-8f;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float rx = r[0] / angle, ry = r[1] / angle, rz = r[2] / angle;
    const float s = sinf(angle), c1 = 1.0f - cosf(angle);
    const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const float kk = K[a * 3] * K[b] + K[a * 3 + 1] * K[3 + b] + K[a * 3 + 2] * K[6 + b];
            R[a * 3 + b] = (a == b ? 1.0f : 0.0f) + s * K[a * 3 + b] + c1 * kk;
        }
}

// ---- forward 2: joints, chain, pose correctives, skinning ------------------------------------------------
__global__ void __launch_bounds__(kSkinThreads)
flame_skin_kernel(int V, int J, Parents parents, int nblk_blend, const float* __restrict__ pose,
                  const float* __restrict__ posedirs, const float* __restrict__ delta_posedirs,
                  const float* __restrict__ lbs_weights, const float* __restrict__ v_shaped,
                  const float* __restrict__ jpart, FlameState* __restrict__ state, float* __restrict__ v_posed_out,
                  float* __restrict__ verts, float* __restrict__ verts_orig, float* __restrict__ pose_feature_out,
                  float* __restrict__ transforms, float* __restrict__ transforms_orig) {
    __shared__ FlameState S;
    __shared__ float s_vp[2][kSkinCoords], s_po[2][kSkinCoords], s_pd[2][kSkinCoords];
    const int t = threadIdx.x;
    const int NP = (J - 1) * 9, n3 = 3 * V;
    fs::pdl_trigger();  // a PDL-launched successor (fs_pose_forward) may begin launching; it waits for this grid
    // ---- everything that does not depend on the blend kernel runs before the wait and overlaps it -------------
    if (t >= 192 && t < 192 + J) {
        const int j = t - 192;
        const float r[3] = {__ldg(pose + 3 * j), __ldg(pose + 3 * j + 1), __ldg(pose + 3 * j + 2)};
        rodrigues(r, S.R[j]);
        if (j > 0)
#pragma unroll
            for (int e = 0; e < 9; ++e) S.pf[(j - 1) * 9 + e] = S.R[j][e] - ((e % 4) == 0 ? 1.0f : 0.0f);
    }
    __syncthreads();
    const int tc = t % kSkinCoords, part = t / kSkinCoords;  // part 0/1: even/odd rows of posedirs; 2: helpers
    const int e = blockIdx.x * kSkinCoords + tc;             // coordinate index 3 v + k
    float wj[kMaxJ];
    if (part < 2 && e < n3) {
        float po = 0.f, pd = 0.f;  // pose_feature @ posedirs, pose_feature @ (posedirs + delta)
#pragma unroll 6
        for (int i = part; i < NP; i += 2) {
            const float p = __ldg(posedirs + (size_t)i * n3 + e);
            po += S.pf[i] * p;
            if (delta_posedirs) pd += S.pf[i] * (p + __ldg(delta_posedirs + (size_t)i * n3 + e));
        }
        s_po[part][tc] = po;
        s_pd[part][tc] = delta_posedirs ? pd : po;
        if (part == 0)
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) wj[j] = j < J ? __ldg(lbs_weights + (size_t)(e / 3) * J + j) : 0.0f;
    }
    fs::pdl_wait();
    float vsd = 0.f, vso = 0.f;
    if (part == 0 && e < n3) {
        vsd = v_shaped[e];
        vso = v_shaped[(size_t)n3 + e];
    }
    // finish the joint regression (lbs.py:207)
    for (int base = 0; base < 2 * J * 3; base += kSkinThreads / 8) {  // warp-uniform trip count (shuffles inside)
        const int slot = base + (t >> 3);
        const bool live = slot < 2 * J * 3;
        const float sum = reduce_partials8(jpart, live ? slot : 0, live ? nblk_blend : 0, nblk_blend, t & 7);
        if (live && (t & 7) == 0) (&S.J[0][0][0])[(slot / (3 * J)) * kMaxJ * 3 + (slot % (3 * J))] = sum;
    }
    __syncthreads();
    if (t < 64) {  // kinematic chain (lbs.py:285-342): warp p_ = path, lanes 0..8 = rotation entries, 9..11 = translation
        const int p_ = t >> 5, lane = t & 31;
        for (int j = 0; j < J; ++j) {  // parents[j] < j
            const int pa = parents.p[j];
            if (lane < 12) {
                float rel[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) rel[k] = S.J[p_][j][k] - (j > 0 ? S.J[p_][pa][k] : 0.0f);
                if (lane < 9) {
                    const int a = lane / 3, b = lane % 3;
                    const float* G = S.Rg[p_][pa];
                    S.Rg[p_][j][lane] = j == 0 ? S.R[0][lane]
                                               : G[a * 3] * S.R[j][b] + G[a * 3 + 1] * S.R[j][3 + b] + G[a * 3 + 2] * S.R[j][6 + b];
                } else {
                    const int a = lane - 9;
                    const float* G = S.Rg[p_][pa];
                    S.tg[p_][j][a] = j == 0 ? rel[a]
                                            : G[a * 3] * rel[0] + G[a * 3 + 1] * rel[1] + G[a * 3 + 2] * rel[2] + S.tg[p_][pa][a];
                }
            }
            __syncwarp();
        }
        for (int i = lane; i < J * 12; i += 32) {  // A = [Rg | tg - Rg J]
            const int j = i / 12, a = (i % 12) / 4, c = i % 4;
            const float* G = S.Rg[p_][j];
            S.A[p_][j][a * 4 + c] = c < 3 ? G[a * 3 + c]
                                          : S.tg[p_][j][a] - (G[a * 3] * S.J[p_][j][0] + G[a * 3 + 1] * S.J[p_][j][1] + G[a * 3 + 2] * S.J[p_][j][2]);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0) {  // publish the small results once
        float* dst = reinterpret_cast<float*>(state);
        const float* src = reinterpret_cast<const float*>(&S);
        for (int i = t; i < (int)(sizeof(FlameState) / sizeof(float)); i += kSkinThreads) dst[i] = src[i];
        if (pose_feature_out)
            for (int i = t; i < NP; i += kSkinThreads) pose_feature_out[i] = S.pf[i];
        for (int i = t; i < J * 16; i += kSkinThreads) {
            const int j = i >> 4, e = i & 15;
            if (transforms) transforms[i] = e < 12 ? S.A[0][j][e] : (e == 15 ? 1.0f : 0.0f);
            if (transforms_orig) transforms_orig[i] = e < 12 ? S.A[1][j][e] : (e == 15 ? 1.0f : 0.0f);
        }
    }

    if (part == 0) {
        float vpd = 0.f, vpo = 0.f;
        if (e < n3) {
            vpd = (s_pd[0][tc] + s_pd[1][tc]) + vsd;
            vpo = (s_po[0][tc] + s_po[1][tc]) + vso;
            v_posed_out[e] = vpd;
        }
        s_vp[0][tc] = vpd;
        s_vp[1][tc] = vpo;
    }
    __syncthreads();
    if (part == 0 && e < n3) {
        const int vl = tc / 3, k = tc % 3;
        float Td[4] = {0.f, 0.f, 0.f, 0.f}, To[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) {
            if (j < J) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    Td[c] += wj[j] * S.A[0][j][k * 4 + c];
                    To[c] += wj[j] * S.A[1][j][k * 4 + c];
                }
            }
        }
        verts[e] = Td[0] * s_vp[0][3 * vl] + Td[1] * s_vp[0][3 * vl + 1] + Td[2] * s_vp[0][3 * vl + 2] + Td[3];
        if (verts_orig)
            verts_orig[e] = To[0] * s_vp[1][3 * vl] + To[1] * s_vp[1][3 * vl + 1] + To[2] * s_vp[1][3 * vl + 2] + To[3];
    }
}

// ---- backward 1: through the skinning ----------------------------------------------------------------------
__global__ void __launch_bounds__(kSkinBwdThreads)
flame_skin_backward_kernel(int V, int J, const float* __restrict__ lbs_weights, const FlameState* __restrict__ state,
                           const float* __restrict__ v_posed, const float* __restrict__ dL_dverts,
                           float* __restrict__ g_posed, float* __restrict__ g_posed_user,
                           float* __restrict__ dapart /*[12J][grid]*/) {
    __shared__ float s_A[kMaxJ][12];
    __shared__ float s_g[kSkinBwdThreads], s_vp[kSkinBwdThreads], s_w[kSkinVerts][kMaxJ];
    const int t = threadIdx.x;
    const int v0 = blockIdx.x * kSkinVerts;
    const int e = blockIdx.x * kSkinBwdThreads + t, n3 = 3 * V;
    fs::pdl_wait();  // launched while the producer of dL/dverts (fs_pose_backward) drains
    for (int i = t; i < J * 12; i += kSkinBwdThreads) s_A[i / 12][i % 12] = state->A[0][i / 12][i % 12];
    s_g[t] = e < n3 ? dL_dverts[e] : 0.0f;
    s_vp[t] = e < n3 ? v_posed[e] : 0.0f;
    for (int i = t; i < kSkinVerts * J; i += kSkinBwdThreads) {
        const int vl = i / J, j = i % J;
        s_w[vl][j] = (v0 + vl) < V ? __ldg(lbs_weights + (size_t)(v0 + vl) * J + j) : 0.0f;
    }
    __syncthreads();
    if (e < n3) {  // dL/dv_posed[c] = sum_k T[k][c] g[k]
        const int vl = t / 3, c = t % 3;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float T = 0.f;
            for (int j = 0; j < J; ++j) T += s_w[vl][j] * s_A[j][k * 4 + c];
            acc += T * s_g[3 * vl + k];
        }
        g_posed[e] = acc;
        if (g_posed_user) g_posed_user[e] = acc;
    }
    if (t < J * 12) {  // dL/dA[j][k][c] = sum_v W[v][j] g[v][k] (c < 3 ? v_posed[v][c] : 1), this CTA's vertices
        const int j = t / 12, k = (t % 12) / 4, c = t % 4;
        float acc = 0.f;
        for (int vl = 0; vl < kSkinVerts; ++vl)
            acc += s_w[vl][j] * s_g[3 * vl + k] * (c < 3 ? s_vp[3 * vl + c] : 1.0f);
        dapart[(size_t)t * gridDim.x + blockIdx.x] = acc;  // [slot][CTA]
    }
    fs::pdl_trigger();
}

// Backward of the kinematic chain (lbs.py:285-342) on one warp, everything in shared memory.
//   A[j] = [Rg_j | tg_j - Rg_j J_j];  Rg_j = Rg_p R_j;  tg_j = Rg_p (J_j - J_p) + tg_p  (j > 0);  tg_0 = J_0
// in : dA (sum over vertices of W g (x) [v_posed; 1]);  out: dJ (joints) and, when dRloc != nullptr, the gradient of
// every joint's LOCAL rotation R_j (the path to the pose coefficients).  Lanes 0..8 own the entries of dRg, 9..11
// those of drel / dJ, 12..14 those of dtg, 16..24 those of dRloc.
__device__ __forceinline__ void chain_backward_warp(int lane, int J, const Parents& parents, const FlameState& S,
                                                    float (*s_dA)[12], float (*s_dRg)[9], float (*s_dtg)[3],
                                                    float (*s_dJ)[3], float (*s_dRloc)[9]) {
    for (int i = lane; i < J * 12; i += 32) {
        const int j = i / 12, r = i % 12;
        if (r < 9) s_dRg[j][r] = s_dA[j][(r / 3) * 4 + r % 3] - s_dA[j][(r / 3) * 4 + 3] * S.J[0][j][r % 3];
        else s_dtg[j][r - 9] = s_dA[j][(r - 9) * 4 + 3];
    }
    __syncwarp();
    for (int i = lane; i < J * 3; i += 32) {
        const int j = i / 3, b = i % 3;
        const float* G = S.Rg[0][j];
        s_dJ[j][b] = -(G[b] * s_dtg[j][0] + G[3 + b] * s_dtg[j][1] + G[6 + b] * s_dtg[j][2]);
    }
    __syncwarp();
    for (int j = J - 1; j >= 1; --j) {
        const int pa = parents.p[j];
        if (lane < 9) {  // dRg_p += dRg_j R_j^T + dtg_j (x) rel
            const int a = lane / 3, b = lane % 3;
            const float* Rj = S.R[j];
            const float rel = S.J[0][j][b] - S.J[0][pa][b];
            s_dRg[pa][lane] += s_dRg[j][a * 3] * Rj[b * 3] + s_dRg[j][a * 3 + 1] * Rj[b * 3 + 1] +
                               s_dRg[j][a * 3 + 2] * Rj[b * 3 + 2] + s_dtg[j][a] * rel;
        } else if (lane < 12) {  // drel = Rg_p^T dtg_j
            const int b = lane - 9;
            const float* Gp = S.Rg[0][pa];
            const float drel = Gp[b] * s_dtg[j][0] + Gp[3 + b] * s_dtg[j][1] + Gp[6 + b] * s_dtg[j][2];
            s_dJ[j][b] += drel;
            s_dJ[pa][b] -= drel;
        } else if (s_dRloc && lane >= 16 && lane < 25) {  // dR_j = Rg_p^T dRg_j (dRg_j is final: children came first)
            const int a = (lane - 16) / 3, b = (lane - 16) % 3;
            const float* Gp = S.Rg[0][pa];
            s_dRloc[j][lane - 16] = Gp[a] * s_dRg[j][b] + Gp[3 + a] * s_dRg[j][3 + b] + Gp[6 + a] * s_dRg[j][6 + b];
        }
        __syncwarp();
        if (lane >= 12 && lane < 15) s_dtg[pa][lane - 12] += s_dtg[j][lane - 12];
        __syncwarp();
    }
    if (lane < 3) s_dJ[0][lane] += s_dtg[0][lane];
    if (s_dRloc && lane < 9) s_dRloc[0][lane] = s_dRg[0][lane];
}

// ---- backward 2: chain backward + parameter gradients -------------------------------------------------------
__global__ void __launch_bounds__(kBlendBwdThreads)
flame_blend_backward_kernel(int V, int L, int l0, int J, Parents parents, int nblk_skin,
                            const float* __restrict__ betas, const float* __restrict__ J_regressor,
                            const FlameState* __restrict__ state, const float* __restrict__ g_posed,
                            const float* __restrict__ dapart, float* __restrict__ d_delta_vertex,
                            float* __restrict__ d_delta_shapedirs, float* __restrict__ d_delta_posedirs,
                            float* __restrict__ d_v_shaped, float* __restrict__ factor_header /*[L + NP] or null*/) {
    __shared__ float s_dA[kMaxJ][12];
    __shared__ float s_dJ[kMaxJ][3], s_dRg[kMaxJ][9], s_dtg[kMaxJ][3];
    __shared__ FlameState S;  // forward state (joints, rotations, chain): one round trip instead of one per use
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int NP = (J - 1) * 9, n3 = 3 * V;
    fs::pdl_trigger();  // the next frame's first kernel may begin launching; it waits for this grid
    for (int i = t; i < (int)(sizeof(FlameState) / sizeof(float)); i += kBlendBwdThreads)
        reinterpret_cast<float*>(&S)[i] = reinterpret_cast<const float*>(state)[i];  // written by the forward call
    fs::pdl_wait();
    for (int base = 0; base < J * 12; base += kBlendBwdThreads / 8) {  // warp-uniform trip count (shuffles inside)
        const int slot = base + (t >> 3);
        const bool live = slot < J * 12;
        const float sum = reduce_partials8(dapart, live ? slot : 0, live ? nblk_skin : 0, nblk_skin, t & 7);
        if (live && (t & 7) == 0) s_dA[slot / 12][slot % 12] = sum;
    }
    __syncthreads();
    if (wid == 0) chain_backward_warp(lane, J, parents, S, s_dA, s_dRg, s_dtg, s_dJ, nullptr);
    __syncthreads();

    if (factor_header && blockIdx.x == 0) {  // [betas | pose_feature]: the head of this rank's factor record
        for (int i = t; i < L; i += kBlendBwdThreads) factor_header[i] = i >= l0 ? __ldg(betas + i) : 0.0f;
        for (int i = t; i < NP; i += kBlendBwdThreads) factor_header[L + i] = S.pf[i];
    }
    // rank-1 gradient of delta_posedirs: pose_feature (x) dL/dv_posed, coalesced along the coordinate axis
    if (d_delta_posedirs) {
        const int nthreads = gridDim.x * kBlendBwdThreads;
        for (int e = blockIdx.x * kBlendBwdThreads + t; e < n3; e += nthreads) {
            const float g = g_posed[e];
            for (int i = 0; i < NP; ++i) __stcs(d_delta_posedirs + (size_t)i * n3 + e, S.pf[i] * g);
        }
    }
    // per coordinate row: dL/dv_shaped, delta_vertex, and the rank-1 gradient of delta_shapedirs.  A warp owns rows
    // w, w + nw, ...; lane i first computes row i's scalar (all loads of all rows in one round trip), then the warp
    // streams the rows out one after the other.
    const int nw = gridDim.x * (kBlendBwdThreads / 32);
    const int w0 = blockIdx.x * (kBlendBwdThreads / 32) + wid;
    const bool vec = (L & 3) == 0 && (l0 & 3) == 0;
    for (int rbase = w0; rbase < n3; rbase += 32 * nw) {
        const int my = rbase + lane * nw;
        float gs_l = 0.f;
        if (my < n3) {
            const int v = my / 3, k = my % 3;
            float jr[kMaxJ];
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) jr[j] = j < J ? __ldg(J_regressor + (size_t)j * V + v) : 0.0f;
            gs_l = g_posed[my];
#pragma unroll
            for (int j = 0; j < kMaxJ; ++j) gs_l += jr[j] * s_dJ[j < J ? j : 0][k];
            if (d_delta_vertex) d_delta_vertex[my] = gs_l;
            if (d_v_shaped) d_v_shaped[my] = gs_l;
        }
        if (d_delta_shapedirs) {
            for (int i = 0; i < 32; ++i) {
                const int r = rbase + i * nw;
                if (r >= n3) break;
                const float gs = __shfl_sync(0xffffffffu, gs_l, i);
                float* row = d_delta_shapedirs + (size_t)r * L;
                if (vec) {
                    const float4* b4 = reinterpret_cast<const float4*>(betas);
                    for (int c = lane; c < (L >> 2); c += 32) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (4 * c >= l0) {
                            const float4 b = __ldg(b4 + c);
                            o = make_float4(b.x * gs, b.y * gs, b.z * gs, b.w * gs);
                        }
                        __stcs(reinterpret_cast<float4*>(row) + c, o);
                    }
                } else {
                    for (int c = lane; c < L; c += 32) __stcs(row + c, c >= l0 ? __ldg(betas + c) * gs : 0.0f);
                }
            }
        }
    }
}

// ---- multi-GPU: dense delta gradients from the per-rank rank-1 factors --------------------------
// Data-parallel training needs sum_r dL_r/d(delta_shapedirs), 24 MB per rank, but each rank's term is rank-1:
// dL/dv_shaped_r (x) betas_r (and pose_feature_r (x) dL/dv_posed_r for delta_posedirs).  The ranks all-gather the
// factors (~120 KB each) and every rank expands the sum locally.  Factor record per rank:
//   [ betas L | pose_feature NP | dL/dv_shaped 3V | dL/dv_posed 3V ]
constexpr int kMaxRanks = FS_FLAME_MAX_RANKS;
__global__ void __launch_bounds__(kBlendBwdThreads)
flame_expand_kernel(int N, int V, int L, int l0, int NP, const float* __restrict__ factors, size_t stride, float scale,
                    float* __restrict__ d_delta_vertex, float* __restrict__ d_delta_shapedirs,
                    float* __restrict__ d_delta_posedirs) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n3 = 3 * V;
    const size_t o_pf = L, o_gs = (size_t)L + NP, o_gp = (size_t)L + NP + n3;
    if (d_delta_posedirs) {
        const int nthreads = gridDim.x * kBlendBwdThreads;
        for (int e = blockIdx.x * kBlendBwdThreads + t; e < n3; e += nthreads) {
            float g[kMaxRanks];
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r) g[r] = r < N ? factors[r * stride + o_gp + e] * scale : 0.0f;
            for (int i = 0; i < NP; ++i) {
                float o = 0.f;
#pragma unroll
                for (int r = 0; r < kMaxRanks; ++r)
                    if (r < N) o += __ldg(factors + r * stride + o_pf + i) * g[r];
                __stcs(d_delta_posedirs + (size_t)i * n3 + e, o);
            }
        }
    }
    const int nw = gridDim.x * (kBlendBwdThreads / 32);
    const bool vec = (L & 3) == 0 && (l0 & 3) == 0 && (stride & 3) == 0;
    for (int row = blockIdx.x * (kBlendBwdThreads / 32) + wid; row < n3; row += nw) {
        float gs[kMaxRanks], sum = 0.f;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) {
            gs[r] = r < N ? factors[r * stride + o_gs + row] * scale : 0.0f;
            sum += gs[r];
        }
        if (lane == 0 && d_delta_vertex) d_delta_vertex[row] = sum;
        if (!d_delta_shapedirs) continue;
        float* out = d_delta_shapedirs + (size_t)row * L;
        if (vec) {
            for (int c = lane; c < (L >> 2); c += 32) {
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (4 * c >= l0) {
#pragma unroll
                    for (int r = 0; r < kMaxRanks; ++r)
                        if (r < N) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(factors + r * stride) + c);
                            o.x += b.x * gs[r];
                            o.y += b.y * gs[r];
                            o.z += b.z * gs[r];
                            o.w += b.w * gs[r];
                        }
                }
                __stcs(reinterpret_cast<float4*>(out) + c, o);
            }
        } else {
            for (int c = lane; c < L; c += 32) {
                float o = 0.f;
                if (c >= l0)
#pragma unroll
                    for (int r = 0; r < kMaxRanks; ++r)
                        if (r < N) o += __ldg(factors + r * stride + c) * gs[r];
                __stcs(out + c, o);
            }
        }
    }
}

// ---- backward 3 (optional): gradients of the expression / pose COEFFICIENTS ---------------------------------------
// Needed only when the caller optimises per-frame tracking (train/base.py:113-151); FateAvatar's INSTA runs do not.
//   dL/dbetas[l] = sum_{v,k} dL/dv_shaped[v,k] (S + dS)[v,k,l]
//   dL/dpose     = Rodrigues^T ( chain path: Rg_p^T dRg_j   +   pose-feature path: (P + dP) dL/dv_posed )
// One chunk of <= 128 coefficients per launch: per-CTA partial sums in the workspace, fixed-order final reduction.
constexpr int kCoeffChunk = 128, kCoeffThreads = 512, kPfSegs = 8;

__global__ void __launch_bounds__(kCoeffThreads)
flame_coeff_partial_kernel(int V, int L, int J, Parents parents, int nblk_skin, int c0, int nc, int do_pose,
                           const float* __restrict__ shapedirs, const float* __restrict__ delta_shapedirs,
                           const float* __restrict__ posedirs, const float* __restrict__ delta_posedirs,
                           const float* __restrict__ J_regressor, const FlameState* __restrict__ state,
                           const float* __restrict__ g_posed, const float* __restrict__ dapart,
                           float* __restrict__ bpart /*[grid][128]*/, float* __restrict__ pfpart /*[kPfSegs][64]*/,
                           float* __restrict__ drloc_out /*[kMaxJ][9]*/) {
    __shared__ float s_dA[kMaxJ][12];
    __shared__ float s_dJ[kMaxJ][3], s_dRg[kMaxJ][9], s_dtg[kMaxJ][3], s_dRloc[kMaxJ][9];
    __shared__ FlameState S;
    __shared__ float s_part[kCoeffThreads / 32][kCoeffChunk];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int NP = (J - 1) * 9, n3 = 3 * V;
    for (int i = t; i < (int)(sizeof(FlameState) / sizeof(float)); i += kCoeffThreads)
        reinterpret_cast<float*>(&S)[i] = reinterpret_cast<const float*>(state)[i];
    for (int base = 0; base < J * 12; base += kCoeffThreads / 8) {
        const int slot = base + (t >> 3);
        const bool live = slot < J * 12;
        const float sum = reduce_partials8(dapart, live ? slot : 0, live ? nblk_skin : 0, nblk_skin, t & 7);
        if (live && (t & 7) == 0) s_dA[slot / 12][slot % 12] = sum;
    }
    __syncthreads();
    if (wid == 0) chain_backward_warp(lane, J, parents, S, s_dA, s_dRg, s_dtg, s_dJ, s_dRloc);
    __syncthreads();
    if (blockIdx.x == 0 && do_pose)
        for (int i = t; i < J * 9; i += kCoeffThreads) drloc_out[i] = s_dRloc[i / 9][i % 9];

    // this chunk of dL/dbetas: a warp per vertex, lane q-th column = c0 + lane + 32 q
    float acc[kCoeffChunk / 32] = {0.f, 0.f, 0.f, 0.f};
    const int nw = gridDim.x * (kCoeffThreads / 32);
    for (int v = blockIdx.x * (kCoeffThreads / 32) + wid; v < V; v += nw) {
        float gsk = 0.f;
        if (lane < 3) {
            gsk = g_posed[3 * v + lane];
            for (int j = 0; j < J; ++j) gsk += __ldg(J_regressor + (size_t)j * V + v) * s_dJ[j][lane];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float g = __shfl_sync(0xffffffffu, gsk, k);
            const size_t row = ((size_t)v * 3 + k) * L + c0;
#pragma unroll
            for (int q = 0; q < kCoeffChunk / 32; ++q) {
                const int col = lane + 32 * q;
                if (col < nc) {
                    float sv = __ldg(shapedirs + row + col);
                    if (delta_shapedirs) sv += __ldg(delta_shapedirs + row + col);
                    acc[q] += g * sv;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kCoeffChunk / 32; ++q) s_part[wid][lane + 32 * q] = acc[q];
    __syncthreads();
    if (t < kCoeffChunk) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kCoeffThreads / 32; ++w) sum += s_part[w][t];
        bpart[(size_t)t * gridDim.x + blockIdx.x] = sum;  // [column][CTA]
    }
    // pose-feature path: dpf[i] = sum_e (P + dP)[i,e] dL/dv_posed[e], split into kPfSegs segments per row
    if (do_pose) {
        const int gw = blockIdx.x * (kCoeffThreads / 32) + wid;
        for (int task = gw; task < NP * kPfSegs; task += nw) {
            const int i = task / kPfSegs, seg = task % kPfSegs;
            const int per = (n3 + kPfSegs - 1) / kPfSegs, e0 = seg * per, e1 = min(n3, e0 + per);
            float sum = 0.f;
            for (int e = e0 + lane; e < e1; e += 32) {
                float p = __ldg(posedirs + (size_t)i * n3 + e);
                if (delta_posedirs) p += __ldg(delta_posedirs + (size_t)i * n3 + e);
                sum += p * g_posed[e];
            }
            sum = warp_sum(sum);
            if (lane == 0) pfpart[seg * 64 + i] = sum;
        }
    }
}

// Rodrigues backward (lbs.py:253-270): R = I + sin(a) K(u) + (1 - cos a) K(u)^2, a = ||r + 1e-8||, u = r / a
__device__ void rodrigues_backward(const float* r, const float* dR, float* dr) {
    const float ex = r[0] + 1e-8f, ey = r[1] + 1e-8f, ez = r[2] + 1e-8f;
    const float a = sqrtf(ex * ex + ey * ey + ez * ez);
    const float u[3] = {r[0] / a, r[1] / a, r[2] / a};
    const float s = sinf(a), c = cosf(a);
    const float K[9] = {0.f, -u[2], u[1], u[2], 0.f, -u[0], -u[1], u[0], 0.f};
    float K2[9], dK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
    float dLds = 0.f, dLdm = 0.f;  // d/d sin, d/d (1 - cos)
    for (int i = 0; i < 9; ++i) {
        dLds += dR[i] * K[i];
        dLdm += dR[i] * K2[i];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {  // dK = s dR + (1 - c) (dR K^T + K^T dR)
            float x = 0.f;
            for (int k = 0; k < 3; ++k) x += dR[i * 3 + k] * K[j * 3 + k] + K[k * 3 + i] * dR[k * 3 + j];
            dK[i * 3 + j] = s * dR[i * 3 + j] + (1.0f - c) * x;
        }
    const float du[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    const float dLda = dLds * c + dLdm * s;
    const float dur = du[0] * r[0] + du[1] * r[1] + du[2] * r[2];
    const float k = (dLda - dur / (a * a)) / a;
    dr[0] = du[0] / a + k * ex;
    dr[1] = du[1] / a + k * ey;
    dr[2] = du[2] / a + k * ez;
}

__global__ void __launch_bounds__(256)
flame_coeff_final_kernel(int L, int l0, int J, int nblk, int c0, int nc, int do_pose, int first,
                         const float* __restrict__ pose, const float* __restrict__ bpart,
                         const float* __restrict__ pfpart, const float* __restrict__ drloc,
                         float* __restrict__ d_betas, float* __restrict__ d_pose) {
    const int t = threadIdx.x;
    if (d_betas) {
        if (first)
            for (int l = t; l < l0; l += 256) d_betas[l] = 0.0f;  // coefficients declared constant-zero by the caller
        for (int col = t; col < nc; col += 256) {
            float sum = 0.f;
            for (int b = 0; b < nblk; ++b) sum += bpart[(size_t)col * nblk + b];
            d_betas[c0 + col] = sum;
        }
    }
    if (do_pose && d_pose && t < J) {
        float dR[9], r[3] = {pose[3 * t], pose[3 * t + 1], pose[3 * t + 2]}, dr[3];
        for (int e = 0; e < 9; ++e) {
            float x = drloc[t * 9 + e];
            if (t > 0)
                for (int seg = 0; seg < kPfSegs; ++seg) x += pfpart[seg * 64 + (t - 1) * 9 + e];
            dR[e] = x;
        }
        rodrigues_backward(r, dR, dr);
        d_pose[3 * t] = dr[0];
        d_pose[3 * t + 1] = dr[1];
        d_pose[3 * t + 2] = dr[2];
    }
}

bool parents_ok(int J, const int* parents_host, Parents& P) {
    if (J < 1 || J > kMaxJ || !parents_host) return false;
    for (int j = 0; j < kMaxJ; ++j) P.p[j] = 0;
    for (int j = 0; j < J; ++j) {
        P.p[j] = parents_host[j];
        if (j > 0 && (parents_host[j] < 0 || parents_host[j] >= j)) return false;
    }
    return true;
}

}  // namespace

extern "C" {

size_t fs_flame_workspace_bytes(int V) { return V > 0 ? flame_layout(V).total : 0; }

int fs_flame_forward(int V, int L, int l0, int J, const int* parents_host, const float* d_betas, const float* d_pose,
                     const float* d_v_template, const float* d_delta_vertex, const float* d_shapedirs,
                     const float* d_delta_shapedirs, const float* d_posedirs, const float* d_delta_posedirs,
                     const float* d_J_regressor, const float* d_lbs_weights, float* d_verts, float* d_verts_orig,
                     float* d_pose_feature, float* d_transforms, float* d_transforms_orig, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
    Parents P;
    if (V <= 0 || L <= 0 || l0 < 0 || l0 > L || !parents_ok(J, parents_host, P)) {
        fs_set_error("fs_flame_forward: invalid size or kinematic tree (V=%d L=%d l0=%d J=%d; need parents[j] < j, J <= %d)",
                     V, L, l0, J, kMaxJ);
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_betas || !d_pose || !d_v_template || !d_shapedirs || !d_posedirs || !d_J_regressor || !d_lbs_weights ||
        !d_verts || !d_workspace) {
        fs_set_error("fs_flame_forward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    const FlameWs w = flame_layout(V);
    if (workspace_bytes < w.total) {
        fs_set_error("fs_flame_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) != 0) {
        fs_set_error("fs_flame_forward: workspace must be 256-byte aligned");
        return FS_ERR_INVALID_ARGUMENT;
    }
    // the 128-bit path needs 16-byte aligned tensor bases (torch allocations are); otherwise use scalars
    const bool aligned = ((reinterpret_cast<uintptr_t>(d_shapedirs) | reinterpret_cast<uintptr_t>(d_delta_shapedirs) |
                           reinterpret_cast<uintptr_t>(d_betas)) & 15) == 0;
    if (!aligned && ((L | l0) & 3) == 0) {
        fs_set_error("fs_flame_forward: shapedirs / delta_shapedirs / betas must be 16-byte aligned");
        return FS_ERR_INVALID_ARGUMENT;
    }
    char* ws = static_cast<char*>(d_workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FsStageTimer timer(FS_STAGE_FLAME_FWD, st);
    fs_launch_pdl(flame_blend_kernel, dim3(w.nblk_blend), dim3(kBlendThreads), 0, st,
        V, L, l0, J, d_betas, d_v_template, d_delta_vertex, d_shapedirs, d_delta_shapedirs, d_J_regressor,
        reinterpret_cast<float*>(ws + w.v_shaped), reinterpret_cast<float*>(ws + w.jpart));
    fs_launch_pdl(flame_skin_kernel, dim3(w.nblk_skin), dim3(kSkinThreads), 0, st, V, J, P, w.nblk_blend, d_pose,
                  d_posedirs, d_delta_posedirs, d_lbs_weights, reinterpret_cast<const float*>(ws + w.v_shaped),
                  reinterpret_cast<const float*>(ws + w.jpart), reinterpret_cast<FlameState*>(ws + w.state),
                  reinterpret_cast<float*>(ws + w.v_posed), d_verts, d_verts_orig, d_pose_feature, d_transforms,
                  d_transforms_orig);
    fs_count_launch(2);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_flame_forward: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

int fs_flame_backward(int V, int L, int l0, int J, const int* parents_host, const float* d_betas,
                      const float* d_J_regressor, const float* d_lbs_weights, const float* d_dL_dverts,
                      void* d_workspace, size_t workspace_bytes, float* d_dL_ddelta_vertex,
                      float* d_dL_ddelta_shapedirs, float* d_dL_ddelta_posedirs, float* d_dL_dv_shaped,
                      float* d_dL_dv_posed, float* d_factor_header, void* stream) {
    Parents P;
    if (V <= 0 || L <= 0 || l0 < 0 || l0 > L || !parents_ok(J, parents_host, P)) {
        fs_set_error("fs_flame_backward: invalid size or kinematic tree");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_betas || !d_J_regressor || !d_lbs_weights || !d_dL_dverts || !d_workspace) {
        fs_set_error("fs_flame_backward: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    const FlameWs w = flame_layout(V);
    if (workspace_bytes < w.total) {
        fs_set_error("fs_flame_backward: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    if (((L | l0) & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(d_betas) | reinterpret_cast<uintptr_t>(d_dL_ddelta_shapedirs)) & 15) != 0) {
        fs_set_error("fs_flame_backward: betas / dL_ddelta_shapedirs must be 16-byte aligned");
        return FS_ERR_INVALID_ARGUMENT;
    }
    char* ws = static_cast<char*>(d_workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FsStageTimer timer(FS_STAGE_FLAME_BWD, st);
    float* g_posed = reinterpret_cast<float*>(ws + w.g_posed);  // kept for fs_flame_backward_coeffs
    fs_launch_pdl(flame_skin_backward_kernel, dim3(w.nblk_skin), dim3(kSkinBwdThreads), 0, st,
        V, J, d_lbs_weights, reinterpret_cast<const FlameState*>(ws + w.state),
        reinterpret_cast<const float*>(ws + w.v_posed), d_dL_dverts, g_posed, d_dL_dv_posed,
        reinterpret_cast<float*>(ws + w.dapart));
    // enough CTAs to stream the 4 L V 3-byte delta_shapedirs gradient at full rate, few enough that the
    // redundant prologue (partials reduce + chain backward) stays negligible
    const int grid = d_dL_ddelta_shapedirs ? 2 * fs_num_sms() : fs_num_sms() / 2 + 1;
    fs_launch_pdl(flame_blend_backward_kernel, dim3(grid), dim3(kBlendBwdThreads), 0, st, V, L, l0, J, P, w.nblk_skin,
                  d_betas, d_J_regressor, reinterpret_cast<const FlameState*>(ws + w.state),
                  (const float*)g_posed, reinterpret_cast<const float*>(ws + w.dapart), d_dL_ddelta_vertex,
                  d_dL_ddelta_shapedirs, d_dL_ddelta_posedirs, d_dL_dv_shaped, d_factor_header);
    fs_count_launch(2);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_flame_backward: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

int fs_flame_backward_coeffs(int V, int L, int l0, int J, const int* parents_host, const float* d_pose,
                             const float* d_shapedirs, const float* d_delta_shapedirs, const float* d_posedirs,
                             const float* d_delta_posedirs, const float* d_J_regressor, void* d_workspace,
                             size_t workspace_bytes, float* d_dL_dbetas, float* d_dL_dpose, void* stream) {
    Parents P;
    if (V <= 0 || L <= 0 || l0 < 0 || l0 > L || !parents_ok(J, parents_host, P)) {
        fs_set_error("fs_flame_backward_coeffs: invalid size or kinematic tree");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_pose || !d_shapedirs || !d_posedirs || !d_J_regressor || !d_workspace) {
        fs_set_error("fs_flame_backward_coeffs: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    const FlameWs w = flame_layout(V);
    if (workspace_bytes < w.total) {
        fs_set_error("fs_flame_backward_coeffs: workspace too small (%zu < %zu bytes)", workspace_bytes, w.total);
        return FS_ERR_WORKSPACE_TOO_SMALL;
    }
    if (!d_dL_dbetas && !d_dL_dpose) return FS_OK;
    char* ws = static_cast<char*>(d_workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FsStageTimer timer(FS_STAGE_FLAME_BWD, st);
    const int grid = fs_num_sms();
    const int n = d_dL_dbetas ? L - l0 : 0;
    int launches = 0;
    for (int c0 = l0, first = 1; first || c0 < l0 + n; c0 += kCoeffChunk, first = 0) {
        const int nc = std::max(0, std::min(kCoeffChunk, l0 + n - c0));
        const int do_pose = first && d_dL_dpose;
        flame_coeff_partial_kernel<<<grid, kCoeffThreads, 0, st>>>(
            V, L, J, P, w.nblk_skin, c0, nc, do_pose, d_shapedirs, d_delta_shapedirs, d_posedirs, d_delta_posedirs,
            d_J_regressor, reinterpret_cast<const FlameState*>(ws + w.state),
            reinterpret_cast<const float*>(ws + w.g_posed), reinterpret_cast<const float*>(ws + w.dapart),
            reinterpret_cast<float*>(ws + w.bpart), reinterpret_cast<float*>(ws + w.pfpart),
            reinterpret_cast<float*>(ws + w.drloc));
        flame_coeff_final_kernel<<<1, 256, 0, st>>>(L, l0, J, grid, c0, nc, do_pose, first, d_pose,
                                                    reinterpret_cast<const float*>(ws + w.bpart),
                                                    reinterpret_cast<const float*>(ws + w.pfpart),
                                                    reinterpret_cast<const float*>(ws + w.drloc), d_dL_dbetas,
                                                    d_dL_dpose);
        launches += 2;
    }
    fs_count_launch(launches);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_flame_backward_coeffs: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

int fs_flame_expand_grads(int N, int V, int L, int l0, int NP, const float* d_factors, size_t rank_stride, float scale,
                          float* d_dL_ddelta_vertex, float* d_dL_ddelta_shapedirs, float* d_dL_ddelta_posedirs,
                          void* stream) {
    if (N < 1 || N > kMaxRanks || V <= 0 || L <= 0 || l0 < 0 || l0 > L || NP < 0 ||
        rank_stride < (size_t)L + NP + 6 * (size_t)V) {
        fs_set_error("fs_flame_expand_grads: invalid size (N=%d, at most %d ranks; stride >= L + NP + 6V)", N, kMaxRanks);
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (!d_factors) {
        fs_set_error("fs_flame_expand_grads: required pointer is NULL");
        return FS_ERR_INVALID_ARGUMENT;
    }
    if (((L | l0) & 3) == 0 && (rank_stride & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(d_factors) | reinterpret_cast<uintptr_t>(d_dL_ddelta_shapedirs)) & 15) != 0) {
        fs_set_error("fs_flame_expand_grads: factors / dL_ddelta_shapedirs must be 16-byte aligned");
        return FS_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FsStageTimer timer(FS_STAGE_FLAME_BWD, st);
    flame_expand_kernel<<<2 * fs_num_sms(), kBlendBwdThreads, 0, st>>>(N, V, L, l0, NP, d_factors, rank_stride, scale,
                                                                      d_dL_ddelta_vertex, d_dL_ddelta_shapedirs,
                                                                      d_dL_ddelta_posedirs);
    fs_count_launch(1);
    if (cudaGetLastError() != cudaSuccess) {
        fs_set_error("fs_flame_expand_grads: launch failed");
        return FS_ERR_CUDA;
    }
    return FS_OK;
}

}  // extern "C"
