/*
 * splat_oracle.c -- CPU restatement of the reference 3D-Gaussian-splatting rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA product in
 * fateavatar_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it.  The product never does.
 *
 * What it restates (reference = /root/reference/submodules/diff-gaussian-rasterization, "DGR"):
 *   orc_preprocess      DGR cuda_rasterizer/forward.cu:155-256 (preprocessCUDA), :118-152 (computeCov3D),
 *                       :74-113 (computeCov2D), :20-71 (computeColorFromSH), auxiliary.h:41-56,58-97,139-164
 *   orc_bin_sort        DGR cuda_rasterizer/rasterizer_impl.cu:70-111 (duplicateWithKeys), :277 (InclusiveSum),
 *                       :300-308 (stable radix sort on [tile|depth] keys), :116-138 (identifyTileRanges)
 *   orc_blend_forward   DGR cuda_rasterizer/forward.cu:261-374 (renderCUDA)
 *   orc_blend_backward  DGR cuda_rasterizer/backward.cu:399-557 (renderCUDA)
 *   orc_preprocess_backward  DGR backward.cu:144-274 (computeCov2DCUDA), :346-396 (preprocessCUDA),
 *                       :20-139 (computeColorFromSH), :278-341 (computeCov3D)
 *   orc_mark_visible    DGR rasterizer_impl.cu:54-66 (checkFrustum)
 *   orc_knn_mean_dist2  /root/reference/submodules/simple-knn/simple_knn.cu:148-222 (exact 3-NN mean squared distance)
 *
 * Arithmetic contract.  The reference is CUDA compiled by nvcc, which contracts a*b+c into FMA.  The
 * integer outputs of the path (radii, tile rectangles, point_list, ranges) depend on fp32 rounding, so
 * this file spells out every multiply / add / fma in the order found in the sm_100 SASS of the reference
 * build (oracle/_ref, see oracle/build_ref.py): dot products are fma(a2,b2, fma(a0,b0, a1*b1)), the
 * pixel mapping is evaluated in double with one fused multiply-add, and so on.  Compile with
 * -ffp-contract=off so the C compiler adds no contraction of its own.  The CUDA kernels use the same
 * sequence with __fmaf_rn/__fmul_rn/__fadd_rn, which is what makes "bit-exact tile/index outputs" testable
 * against both this oracle and the compiled reference.  expf is the one exception: CUDA's expf uses
 * MUFU.EX2, glibc's is correctly rounded to <1ulp; blend outputs are compared with a tolerance.
 *
 * Pinning.  The reference ships no tests or golden vectors for this path.  The oracle is
 * pinned against outputs of the compiled reference itself (oracle/_ref run on a B200), committed under
 * tests/golden/ together with the generating script tests/golden/make_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

/* ---- constants: DGR auxiliary.h:22-39 ---- */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

static inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}
/* row-vector * column k of a flat (transposed) 4x4: auxiliary.h:58-76 */
static inline float xform(const float* m, int k, float x, float y, float z) {
    return dot3(x, m[k], y, m[4 + k], z, m[8 + k]) + m[12 + k];
}
static inline int f2i_trunc(float f) { /* CUDA F2I.TRUNC saturates; NaN -> 0 */
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* tile rectangle: auxiliary.h:46-56 */
static inline void get_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1) {
    float r = (float)radius;
    *x0 = imin(gx, imax(0, f2i_trunc((px - r) * 0.0625f)));
    *y0 = imin(gy, imax(0, f2i_trunc((py - r) * 0.0625f)));
    *x1 = imin(gx, imax(0, f2i_trunc((((px + r) + 16.0f) - 1.0f) * 0.0625f)));
    *y1 = imin(gy, imax(0, f2i_trunc((((py + r) + 16.0f) - 1.0f) * 0.0625f)));
}

/* Sigma3 from scale/rotation: forward.cu:118-152.  q = (r,x,y,z), NOT normalised (forward.cu:127). */
static inline void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* cov) {
    float r = q[0], x = q[1], y = q[2], z = q[3];
    float t_xz = x * z, t_rx = r * x, t_rz = r * z, t_yy = y * y, t_zz = z * z;
    float A = fmaf(r, y, t_xz);   /* xz + ry */
    float B = fmaf(-r, y, t_xz);  /* xz - ry */
    float C = fmaf(y, z, -t_rx);  /* yz - rx */
    float D = fmaf(y, z, t_rx);   /* yz + rx */
    float E = fmaf(x, y, -t_rz);  /* xy - rz */
    float F = fmaf(x, y, t_rz);   /* xy + rz */
    float G = fmaf(x, x, t_yy);
    float Hh = t_yy + t_zz;
    float I = fmaf(x, x, t_zz);
    float R00 = 1.0f - (Hh + Hh), R11 = 1.0f - (I + I), R22 = 1.0f - (G + G);
    float sx = s[0] * mod, sy = s[1] * mod, sz = s[2] * mod;
    /* M = S * R (glm column-major); m<c><r> */
    float m00 = sx * R00, m01 = sy * (E + E), m02 = sz * (A + A);
    float m10 = sx * (F + F), m11 = sy * R11, m12 = sz * (C + C);
    float m20 = sx * (B + B), m21 = sy * (D + D), m22 = sz * R22;
    cov[0] = dot3(m00, m00, m01, m01, m02, m02);
    cov[1] = dot3(m00, m10, m01, m11, m02, m12);
    cov[2] = dot3(m00, m20, m01, m21, m02, m22);
    cov[3] = dot3(m10, m10, m11, m11, m12, m12);
    cov[4] = dot3(m10, m20, m11, m21, m12, m22);
    cov[5] = dot3(m20, m20, m21, m21, m22, m22);
}

/* intermediate of the EWA projection shared by forward and backward (forward.cu:74-113, backward.cu:144-199) */
typedef struct {
    float T00, T01, T02, T10, T11, T12; /* T<c><r>, glm column c row r */
    float a, b, c;                      /* Sigma2 (+0.3 on the diagonal) */
    float tx, ty, tz;                   /* clamped view-space mean */
    float txtz, tytz, limx, limy;
} ewa_t;

static inline void ewa_project(const float* view, float px, float py, float pz, float focal_x, float focal_y,
                               float tan_fovx, float tan_fovy, const float* cov, ewa_t* o) {
    float tx = xform(view, 0, px, py, pz);
    float ty = xform(view, 1, px, py, pz);
    float tz = xform(view, 2, px, py, pz);
    float limx = tan_fovx * 1.3f, limy = tan_fovy * 1.3f;
    float txtz = tx / tz, tytz = ty / tz;
    float cx = fminf(fmaxf(txtz, -limx), limx);
    float cy = fminf(fmaxf(tytz, -limy), limy);
    float tz2 = tz * tz;
    float J00 = focal_x / tz;
    float J02 = ((tz * -cx) * focal_x) / tz2;
    float J11 = focal_y / tz;
    float J12 = ((tz * -cy) * focal_y) / tz2;
    const float m0 = view[0], m1 = view[1], m2 = view[2], m4 = view[4], m5 = view[5], m6 = view[6], m8 = view[8],
                m9 = view[9], m10 = view[10];
    /* T = W * J with W = (m0,m4,m8 | m1,m5,m9 | m2,m6,m10) columns */
    o->T00 = fmaf(m2, J02, m0 * J00);
    o->T01 = fmaf(m6, J02, m4 * J00);
    o->T02 = fmaf(m10, J02, m8 * J00);
    o->T10 = fmaf(m2, J12, m1 * J11);
    o->T11 = fmaf(m6, J12, m5 * J11);
    o->T12 = fmaf(m10, J12, m9 * J11);
    const float c0 = cov[0], c1 = cov[1], c2 = cov[2], c3 = cov[3], c4 = cov[4], c5 = cov[5];
    float A0_0 = dot3(o->T00, c0, o->T01, c1, o->T02, c2);
    float A1_0 = dot3(o->T00, c1, o->T01, c3, o->T02, c4);
    float A2_0 = dot3(o->T00, c2, o->T01, c4, o->T02, c5);
    float A0_1 = dot3(o->T10, c0, o->T11, c1, o->T12, c2);
    float A1_1 = dot3(o->T10, c1, o->T11, c3, o->T12, c4);
    float A2_1 = dot3(o->T10, c2, o->T11, c4, o->T12, c5);
    o->a = dot3(o->T00, A0_0, o->T01, A1_0, o->T02, A2_0) + 0.3f;
    o->b = dot3(o->T00, A0_1, o->T01, A1_1, o->T02, A2_1);
    o->c = dot3(o->T10, A0_1, o->T11, A1_1, o->T12, A2_1) + 0.3f;
    o->tx = cx * tz;
    o->ty = cy * tz;
    o->tz = tz;
    o->txtz = txtz;
    o->tytz = tytz;
    o->limx = limx;
    o->limy = limy;
}

/* SH -> RGB: forward.cu:20-71.  sh points at this Gaussian's [M][3] block. Returns un-offset colour. */
static inline void sh_to_rgb(int deg, const float* sh, float x, float y, float z, float* out) {
    for (int ch = 0; ch < 3; ++ch) {
        const float* s = sh + ch; /* stride 3 between coefficients */
        float res = s[0] * SH_C0;
        if (deg > 0) {
            res = fmaf(-(y * SH_C1), s[3 * 1], res);
            res = fmaf((z * SH_C1), s[3 * 2], res);
            res = fmaf(-(x * SH_C1), s[3 * 3], res);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = y * x, yz = z * y, xz = z * x;
                res = fmaf(xy * SH_C2[0], s[3 * 4], res);
                res = fmaf(yz * SH_C2[1], s[3 * 5], res);
                res = fmaf((((zz + zz) - xx) - yy) * SH_C2[2], s[3 * 6], res);
                res = fmaf(xz * SH_C2[3], s[3 * 7], res);
                res = fmaf((xx - yy) * SH_C2[4], s[3 * 8], res);
                if (deg > 2) {
                    float q = fmaf(zz, 4.0f, -xx) - yy; /* 4zz - xx - yy */
                    res = fmaf((y * SH_C3[0]) * fmaf(xx, 3.0f, -yy), s[3 * 9], res);
                    res = fmaf((xy * SH_C3[1]) * z, s[3 * 10], res);
                    res = fmaf((y * SH_C3[2]) * q, s[3 * 11], res);
                    res = fmaf((z * SH_C3[3]) * fmaf(yy, -3.0f, fmaf(xx, -3.0f, zz + zz)), s[3 * 12], res);
                    res = fmaf(q * (x * SH_C3[4]), s[3 * 13], res);
                    res = fmaf((xx - yy) * (z * SH_C3[5]), s[3 * 14], res);
                    res = fmaf((x * SH_C3[6]) * fmaf(yy, -3.0f, xx), s[3 * 15], res);
                }
            }
        }
        out[ch] = res;
    }
}

/*
 * Per-Gaussian preprocessing.  forward.cu:155-256.
 * Outputs (all length-P arrays, zero/untouched semantics as the reference): radii and tiles_touched are
 * always written (0 when culled); the rest only for surviving Gaussians.
 */
void orc_preprocess(int P, int D, int M, const float* means3D, const float* scales, float scale_modifier,
                    const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                    const float* colors_precomp, const float* view, const float* proj, const float* campos, int W,
                    int H, float tan_fovx, float tan_fovy, int* radii, float* means2D, float* depths, float* cov3Ds,
                    float* rgb, float* conic_opacity, uint8_t* clamped, uint32_t* tiles_touched) {
    const float focal_y = H / (2.0f * tan_fovy);
    const float focal_x = W / (2.0f * tan_fovx);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        float zv = xform(view, 2, px, py, pz);
        if (zv <= 0.2f) continue; /* auxiliary.h:154 */
        float hx = xform(proj, 0, px, py, pz);
        float hy = xform(proj, 1, px, py, pz);
        float hw = xform(proj, 3, px, py, pz);
        float p_w = 1.0f / (hw + 0.0000001f);
        float prx = hx * p_w, pry = hy * p_w;
        const float* cov;
        if (cov3D_precomp) {
            cov = cov3D_precomp + 6 * i;
        } else {
            cov3d_from_scale_rot(scales + 3 * i, scale_modifier, rotations + 4 * i, cov3Ds + 6 * i);
            cov = cov3Ds + 6 * i;
        }
        ewa_t e;
        ewa_project(view, px, py, pz, focal_x, focal_y, tan_fovx, tan_fovy, cov, &e);
        float det = fmaf(e.a, e.c, -(e.b * e.b));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float conx = e.c * det_inv, cony = e.b * -det_inv, conz = e.a * det_inv;
        float mid = (e.a + e.c) * 0.5f;
        float sq = sqrtf(fmaxf(fmaf(mid, mid, -det), 0.1f));
        float lam = fmaxf(mid + sq, mid - sq);
        float rad_f = ceilf(sqrtf(lam) * 3.0f);
        int radius = (rad_f != rad_f) ? 0 : (rad_f >= 2147483648.0f ? INT_MAX : (int)rad_f);
        /* ndc2Pix in double, one fused multiply-add (auxiliary.h:41-44) */
        float pix_x = (float)(fma((double)prx + 1.0, (double)W, -1.0) * 0.5);
        float pix_y = (float)(fma((double)pry + 1.0, (double)H, -1.0) * 0.5);
        int x0, y0, x1, y1;
        get_rect(pix_x, pix_y, radius, gx, gy, &x0, &y0, &x1, &y1);
        uint32_t nt = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
        if (nt == 0) continue;
        if (!colors_precomp) {
            float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
            float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
            float c[3];
            sh_to_rgb(D, shs + (size_t)i * M * 3, dx / len, dy / len, dz / len, c);
            for (int ch = 0; ch < 3; ++ch) {
                int cl = !(c[ch] >= -0.5f); /* == (c+0.5 < 0); NaN counts as not clamped */
                if (c[ch] != c[ch]) cl = 0;
                clamped[3 * i + ch] = (uint8_t)cl;
                rgb[3 * i + ch] = cl ? 0.0f : c[ch] + 0.5f;
            }
        }
        depths[i] = zv;
        radii[i] = radius;
        means2D[2 * i] = pix_x;
        means2D[2 * i + 1] = pix_y;
        conic_opacity[4 * i] = conx;
        conic_opacity[4 * i + 1] = cony;
        conic_opacity[4 * i + 2] = conz;
        conic_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = nt;
    }
}

/* rasterizer_impl.cu:54-66 */
void orc_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present) {
    (void)proj;
    for (int i = 0; i < P; ++i) {
        float zv = xform(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        present[i] = (uint8_t)!(zv <= 0.2f);
    }
}

/*
 * Binning + sort.  Returns R = num_rendered.  Call with point_list == NULL to get R only.
 * keys: tile<<32 | depth bits; values emitted row-major over the rect in ascending Gaussian index
 * (rasterizer_impl.cu:98-109); stable sort (cub radix sort is stable) => ties keep ascending index.
 */
int64_t orc_bin_sort(int P, int W, int H, const int* radii, const float* means2D, const float* depths,
                     const uint32_t* tiles_touched, uint32_t* point_offsets, uint64_t* keys_sorted,
                     uint32_t* point_list, uint32_t* ranges /* [Tn][2] */) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    uint64_t acc = 0;
    for (int i = 0; i < P; ++i) {
        acc += tiles_touched[i];
        if (point_offsets) point_offsets[i] = (uint32_t)acc;
    }
    const int64_t R = (int64_t)acc;
    if (!point_list) return R;
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R == 0) return 0;
    uint64_t* k0 = (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * R);
    uint64_t* k1 = keys_sorted ? keys_sorted : (uint64_t*)malloc(sizeof(uint64_t) * R);
    uint32_t* v1 = point_list;
    size_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(means2D[2 * i], means2D[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, &depths[i], 4);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                k0[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                v0[off] = (uint32_t)i;
                ++off;
            }
    }
    /* LSD radix sort, 8 passes of 8 bits over the full 64-bit key (superset of the reference's 32+bit bits). */
    uint64_t *ka = k0, *kb = k1;
    uint32_t *va = v0, *vb = v1;
    for (int pass = 0; pass < 8; ++pass) {
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        int sh = pass * 8;
        for (int64_t j = 0; j < R; ++j) cnt[((ka[j] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t j = 0; j < R; ++j) {
            size_t dst = cnt[(ka[j] >> sh) & 255]++;
            kb[dst] = ka[j];
            vb[dst] = va[j];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    /* after 8 passes the result is back in (k0,v0) */
    memcpy(k1, k0, sizeof(uint64_t) * R);
    memcpy(v1, v0, sizeof(uint32_t) * R);
    for (int64_t j = 0; j < R; ++j) { /* rasterizer_impl.cu:116-138 */
        uint32_t cur = (uint32_t)(k1[j] >> 32);
        if (j == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(k1[j - 1] >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)j;
                ranges[2 * cur] = (uint32_t)j;
            }
        }
        if (j == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
    free(k0);
    free(v0);
    if (!keys_sorted) free(k1);
    return R;
}

static inline float splat_power(float mx, float my, float pxf, float pyf, float cx, float cy, float cz, float* dx,
                                float* dy) {
    float ddx = mx - pxf, ddy = my - pyf;
    *dx = ddx;
    *dy = ddy;
    /* -0.5*(cx*dx*dx + cz*dy*dy) - cy*dx*dy as contracted by nvcc (forward.cu:336) */
    float q = fmaf(ddx, ddx * cx, ddy * (ddy * cz));
    return fmaf(q, -0.5f, -(ddy * (ddx * cy)));
}

/* forward.cu:261-374.
 * `fragile` (optional, [H*W]) is set for pixels where one of the rule's three discrete decisions
 * that involve expf (alpha < 1/255, T(1-alpha) < 1e-4) was within rounding distance of its threshold: there a 1-ulp
 * difference between glibc's expf and CUDA's MUFU-based expf can flip the decision, so tests compare such
 * pixels with the looser bound of one dropped 1/255 contribution instead of 1e-5. */
/* Work counters of the reference blend forward (forward.cu:261-374) for one frame, for the pairs/s figure that sits
 * next to the HBM roofline: counts[0] = (pixel, list entry) pairs the reference kernel walks until
 * each pixel terminates, counts[1] = of those, pairs whose exponential is evaluated (power <= 0),
 * counts[2] = pairs that contribute colour (alpha >= 1/255, applied before termination). */
void orc_blend_pair_counts(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                           const float* conic_opacity, unsigned long long* counts) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    unsigned long long walked = 0, evals = 0, contrib = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : walked, evals, contrib)
    for (int t = 0; t < gx * gy; ++t) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                float T = 1.0f;
                for (uint32_t j = r0; j < r1; ++j) {
                    ++walked;
                    const uint32_t g = point_list[j];
                    const float* co = conic_opacity + 4 * (size_t)g;
                    float dx, dy;
                    float power = splat_power(means2D[2 * g], means2D[2 * g + 1], (float)x, (float)y, co[0], co[1], co[2], &dx, &dy);
                    if (power > 0.0f) continue;
                    ++evals;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break;
                    T = test_T;
                    ++contrib;
                }
            }
    }
    counts[0] = walked;
    counts[1] = evals;
    counts[2] = contrib;
}

void orc_blend_forward(int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                       const float* colors, const float* conic_opacity, const float* bg, float* out_color,
                       float* final_T, uint32_t* n_contrib, uint8_t* fragile) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; ++t) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const float pxf = (float)x, pyf = (float)y;
                float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
                uint32_t contributor = 0, last = 0;
                uint8_t frag = 0;
                for (uint32_t j = r0; j < r1; ++j) {
                    contributor++;
                    const uint32_t g = point_list[j];
                    const float* co = conic_opacity + 4 * (size_t)g;
                    float dx, dy;
                    float power = splat_power(means2D[2 * g], means2D[2 * g + 1], pxf, pyf, co[0], co[1], co[2], &dx, &dy);
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expf(power));
                    if (fabsf(alpha - 1.0f / 255.0f) < 1e-8f) frag = 1; /* expf differs by <= 3 ulp: |d alpha| < 2e-9 */
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1.0f - alpha);
                    if (fabsf(test_T - 0.0001f) < 3e-9f) frag = 1;
                    if (test_T < 0.0001f) break;
                    const float* c = colors + 3 * (size_t)g;
                    C0 = fmaf(T, alpha * c[0], C0);
                    C1 = fmaf(T, alpha * c[1], C1);
                    C2 = fmaf(T, alpha * c[2], C2);
                    T = test_T;
                    last = contributor;
                }
                const size_t pid = (size_t)y * W + x;
                final_T[pid] = T;
                n_contrib[pid] = last;
                if (fragile) fragile[pid] = frag;
                out_color[pid] = fmaf(bg[0], T, C0);
                out_color[(size_t)H * W + pid] = fmaf(bg[1], T, C1);
                out_color[2 * (size_t)H * W + pid] = fmaf(bg[2], T, C2);
            }
    }
}

/*
 * backward.cu:399-557.  Sums over pixels are accumulated in double (the reference uses float atomics in a
 * nondeterministic order; the double sum is the value they all approximate).
 * Outputs are [P,3] dL_dmean2D (z untouched), [P,4] dL_dconic (x,y,w used), [P] dL_dopacity, [P,3] dL_dcolors;
 * they are *accumulated into* (caller zero-fills), as the reference does.
 */
void orc_blend_backward(int P, int W, int H, const uint32_t* ranges, const uint32_t* point_list, const float* bg,
                        const float* means2D, const float* conic_opacity, const float* colors, const float* final_T,
                        const uint32_t* n_contrib, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                        float* dL_dopacity, float* dL_dcolors) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double* acc = (double*)calloc((size_t)P * 9, sizeof(double));
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; ++t) {
        const int tx = t % gx, ty = t / gx;
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        const uint32_t n = r1 - r0;
        if (n == 0) continue;
        double* loc = (double*)calloc((size_t)n * 9, sizeof(double));
        for (int ly = 0; ly < TILE; ++ly)
            for (int lx = 0; lx < TILE; ++lx) {
                const int x = tx * TILE + lx, y = ty * TILE + ly;
                if (x >= W || y >= H) continue;
                const size_t pid = (size_t)y * W + x;
                const float pxf = (float)x, pyf = (float)y;
                const float T_final = final_T[pid];
                float T = T_final;
                const uint32_t last = n_contrib[pid];
                float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
                const float dpix[3] = {dL_dpix[pid], dL_dpix[(size_t)H * W + pid], dL_dpix[2 * (size_t)H * W + pid]};
                for (uint32_t k = last; k-- > 0;) { /* positions last-1 .. 0 (entries >= last are skipped, :486-488) */
                    const uint32_t g = point_list[r0 + k];
                    const float* co = conic_opacity + 4 * (size_t)g;
                    float dx, dy;
                    float power = splat_power(means2D[2 * g], means2D[2 * g + 1], pxf, pyf, co[0], co[1], co[2], &dx, &dy);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.0f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    double* a = loc + (size_t)k * 9;
                    for (int ch = 0; ch < 3; ++ch) {
                        const float c = colors[3 * (size_t)g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dpix[ch];
                        a[6 + ch] += (double)(dchannel_dcolor * dpix[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0.f;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dpix[ch];
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    a[0] += (double)(dL_dG * dG_ddelx * ddelx_dx);
                    a[1] += (double)(dL_dG * dG_ddely * ddely_dy);
                    a[2] += (double)(-0.5f * gdx * dx * dL_dG);
                    a[3] += (double)(-0.5f * gdx * dy * dL_dG);
                    a[4] += (double)(-0.5f * gdy * dy * dL_dG);
                    a[5] += (double)(G * dL_dalpha);
                }
            }
        for (uint32_t k = 0; k < n; ++k) {
            const uint32_t g = point_list[r0 + k];
            for (int q = 0; q < 9; ++q) {
                double v = loc[(size_t)k * 9 + q];
                if (v != 0.0) {
#pragma omp atomic
                    acc[(size_t)g * 9 + q] += v;
                }
            }
        }
        free(loc);
    }
    for (int i = 0; i < P; ++i) {
        const double* a = acc + (size_t)i * 9;
        dL_dmean2D[3 * i] += (float)a[0];
        dL_dmean2D[3 * i + 1] += (float)a[1];
        dL_dconic[4 * i] += (float)a[2];
        dL_dconic[4 * i + 1] += (float)a[3];
        dL_dconic[4 * i + 3] += (float)a[4];
        dL_dopacity[i] += (float)a[5];
        dL_dcolors[3 * i] += (float)a[6];
        dL_dcolors[3 * i + 1] += (float)a[7];
        dL_dcolors[3 * i + 2] += (float)a[8];
    }
    free(acc);
}

/* backward.cu:20-139: SH backward.  dL_dmeans is accumulated into. */
static void sh_backward(int deg, int M, const float* sh, const uint8_t* clamped, float dirx, float diry, float dirz,
                        const float* dL_dcolor, float* dL_dmean, float* dL_dsh) {
    float len = sqrtf(dirx * dirx + diry * diry + dirz * dirz);
    float x = dirx / len, y = diry / len, z = dirz / len;
    float dRGB[3] = {dL_dcolor[0] * (clamped[0] ? 0.f : 1.f), dL_dcolor[1] * (clamped[1] ? 0.f : 1.f),
                     dL_dcolor[2] * (clamped[2] ? 0.f : 1.f)};
    float dx3[3] = {0, 0, 0}, dy3[3] = {0, 0, 0}, dz3[3] = {0, 0, 0};
    (void)M;
#define SH(k, ch) sh[3 * (k) + (ch)]
#define DSH(k, ch) dL_dsh[3 * (k) + (ch)]
    for (int ch = 0; ch < 3; ++ch) {
        DSH(0, ch) = SH_C0 * dRGB[ch];
        if (deg > 0) {
            DSH(1, ch) = (-SH_C1 * y) * dRGB[ch];
            DSH(2, ch) = (SH_C1 * z) * dRGB[ch];
            DSH(3, ch) = (-SH_C1 * x) * dRGB[ch];
            dx3[ch] = -SH_C1 * SH(3, ch);
            dy3[ch] = -SH_C1 * SH(1, ch);
            dz3[ch] = SH_C1 * SH(2, ch);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                DSH(4, ch) = (SH_C2[0] * xy) * dRGB[ch];
                DSH(5, ch) = (SH_C2[1] * yz) * dRGB[ch];
                DSH(6, ch) = (SH_C2[2] * (2.f * zz - xx - yy)) * dRGB[ch];
                DSH(7, ch) = (SH_C2[3] * xz) * dRGB[ch];
                DSH(8, ch) = (SH_C2[4] * (xx - yy)) * dRGB[ch];
                dx3[ch] += SH_C2[0] * y * SH(4, ch) + SH_C2[2] * 2.f * -x * SH(6, ch) + SH_C2[3] * z * SH(7, ch) +
                           SH_C2[4] * 2.f * x * SH(8, ch);
                dy3[ch] += SH_C2[0] * x * SH(4, ch) + SH_C2[1] * z * SH(5, ch) + SH_C2[2] * 2.f * -y * SH(6, ch) +
                           SH_C2[4] * 2.f * -y * SH(8, ch);
                dz3[ch] += SH_C2[1] * y * SH(5, ch) + SH_C2[2] * 2.f * 2.f * z * SH(6, ch) + SH_C2[3] * x * SH(7, ch);
                if (deg > 2) {
                    DSH(9, ch) = (SH_C3[0] * y * (3.f * xx - yy)) * dRGB[ch];
                    DSH(10, ch) = (SH_C3[1] * xy * z) * dRGB[ch];
                    DSH(11, ch) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dRGB[ch];
                    DSH(12, ch) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB[ch];
                    DSH(13, ch) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dRGB[ch];
                    DSH(14, ch) = (SH_C3[5] * z * (xx - yy)) * dRGB[ch];
                    DSH(15, ch) = (SH_C3[6] * x * (xx - 3.f * yy)) * dRGB[ch];
                    dx3[ch] += (SH_C3[0] * SH(9, ch) * 3.f * 2.f * xy + SH_C3[1] * SH(10, ch) * yz +
                                SH_C3[2] * SH(11, ch) * -2.f * xy + SH_C3[3] * SH(12, ch) * -3.f * 2.f * xz +
                                SH_C3[4] * SH(13, ch) * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * SH(14, ch) * 2.f * xz +
                                SH_C3[6] * SH(15, ch) * 3.f * (xx - yy));
                    dy3[ch] += (SH_C3[0] * SH(9, ch) * 3.f * (xx - yy) + SH_C3[1] * SH(10, ch) * xz +
                                SH_C3[2] * SH(11, ch) * (-3.f * yy + 4.f * zz - xx) +
                                SH_C3[3] * SH(12, ch) * -3.f * 2.f * yz + SH_C3[4] * SH(13, ch) * -2.f * xy +
                                SH_C3[5] * SH(14, ch) * -2.f * yz + SH_C3[6] * SH(15, ch) * -3.f * 2.f * xy);
                    dz3[ch] += (SH_C3[1] * SH(10, ch) * xy + SH_C3[2] * SH(11, ch) * 4.f * 2.f * yz +
                                SH_C3[3] * SH(12, ch) * 3.f * (2.f * zz - xx - yy) +
                                SH_C3[4] * SH(13, ch) * 4.f * 2.f * xz + SH_C3[5] * SH(14, ch) * (xx - yy));
                }
            }
        }
    }
#undef SH
#undef DSH
    float ddx = dx3[0] * dRGB[0] + dx3[1] * dRGB[1] + dx3[2] * dRGB[2];
    float ddy = dy3[0] * dRGB[0] + dy3[1] * dRGB[1] + dy3[2] * dRGB[2];
    float ddz = dz3[0] * dRGB[0] + dz3[1] * dRGB[1] + dz3[2] * dRGB[2];
    /* dnormvdv: auxiliary.h:107-117 */
    float sum2 = dirx * dirx + diry * diry + dirz * dirz;
    float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    dL_dmean[0] += ((+sum2 - dirx * dirx) * ddx - diry * dirx * ddy - dirz * dirx * ddz) * inv;
    dL_dmean[1] += (-dirx * diry * ddx + (sum2 - diry * diry) * ddy - dirz * diry * ddz) * inv;
    dL_dmean[2] += (-dirx * dirz * ddx - diry * dirz * ddy + (sum2 - dirz * dirz) * ddz) * inv;
}

/*
 * backward.cu:144-274 (cov2D backward) followed by :346-396 (projection/SH/cov3D backward).
 * dL_dmean2D [P,3], dL_dconic [P,4] and dL_dcolors [P,3] are inputs; dL_dmeans3D [P,3], dL_dcov3D [P,6],
 * dL_dsh [P,M,3], dL_dscales [P,3], dL_drots [P,4] are outputs (caller zero-fills; entries of culled
 * Gaussians stay zero).
 */
void orc_preprocess_backward(int P, int D, int M, const float* means3D, const int* radii, const float* shs,
                             const uint8_t* clamped, const float* scales, const float* rotations,
                             float scale_modifier, const float* cov3Ds, const float* view, const float* proj,
                             int W, int H, float tan_fovx, float tan_fovy, const float* campos,
                             const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolors,
                             float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots) {
    const float h_y = H / (2.0f * tan_fovy);
    const float h_x = W / (2.0f * tan_fovx);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        if (!(radii[i] > 0)) continue;
        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        const float* cov = cov3Ds + 6 * i;
        ewa_t e;
        ewa_project(view, px, py, pz, h_x, h_y, tan_fovx, tan_fovy, cov, &e);
        const float x_grad_mul = (e.txtz < -e.limx || e.txtz > e.limx) ? 0.f : 1.f;
        const float y_grad_mul = (e.tytz < -e.limy || e.tytz > e.limy) ? 0.f : 1.f;
        const float a = e.a, b = e.b, c = e.c;
        const float dcx = dL_dconic[4 * i], dcy = dL_dconic[4 * i + 1], dcz = dL_dconic[4 * i + 3];
        const float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        /* glm T[c][r] */
        const float T00 = e.T00, T01 = e.T01, T02 = e.T02, T10 = e.T10, T11 = e.T11, T12 = e.T12;
        float* dcov = dL_dcov3D + 6 * i;
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
        } else {
            for (int k = 0; k < 6; ++k) dcov[k] = 0;
        }
        /* Vrk[c][r], symmetric */
        const float V00 = cov[0], V01 = cov[1], V02 = cov[2], V11 = cov[3], V12 = cov[4], V22 = cov[5];
        float dL_dT00 = 2 * (T00 * V00 + T01 * V01 + T02 * V02) * dL_da + (T10 * V00 + T11 * V01 + T12 * V02) * dL_db;
        float dL_dT01 = 2 * (T00 * V01 + T01 * V11 + T02 * V12) * dL_da + (T10 * V01 + T11 * V11 + T12 * V12) * dL_db;
        float dL_dT02 = 2 * (T00 * V02 + T01 * V12 + T02 * V22) * dL_da + (T10 * V02 + T11 * V12 + T12 * V22) * dL_db;
        float dL_dT10 = 2 * (T10 * V00 + T11 * V01 + T12 * V02) * dL_dc + (T00 * V00 + T01 * V01 + T02 * V02) * dL_db;
        float dL_dT11 = 2 * (T10 * V01 + T11 * V11 + T12 * V12) * dL_dc + (T00 * V01 + T01 * V11 + T02 * V12) * dL_db;
        float dL_dT12 = 2 * (T10 * V02 + T11 * V12 + T12 * V22) * dL_dc + (T00 * V02 + T01 * V12 + T02 * V22) * dL_db;
        /* W[c][r]: W[0]=(m0,m4,m8) W[1]=(m1,m5,m9) W[2]=(m2,m6,m10) */
        float dL_dJ00 = view[0] * dL_dT00 + view[4] * dL_dT01 + view[8] * dL_dT02;
        float dL_dJ02 = view[2] * dL_dT00 + view[6] * dL_dT01 + view[10] * dL_dT02;
        float dL_dJ11 = view[1] * dL_dT10 + view[5] * dL_dT11 + view[9] * dL_dT12;
        float dL_dJ12 = view[2] * dL_dT10 + view[6] * dL_dT11 + view[10] * dL_dT12;
        float tz = 1.f / e.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * e.tx) * tz3 * dL_dJ02 +
                       (2 * h_y * e.ty) * tz3 * dL_dJ12;
        /* transformVec4x3Transpose (auxiliary.h:88-96); cov2D kernel *assigns* dL_dmeans (:273) */
        float gm[3];
        gm[0] = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        gm[1] = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        gm[2] = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        /* projection Jacobian of the 2D mean gradient: backward.cu:373-387 */
        float hx = proj[0] * px + proj[4] * py + proj[8] * pz + proj[12];
        float hy = proj[1] * px + proj[5] * py + proj[9] * pz + proj[13];
        float hw = proj[3] * px + proj[7] * py + proj[11] * pz + proj[15];
        float m_w = 1.0f / (hw + 0.0000001f);
        float mul1 = hx * m_w * m_w, mul2 = hy * m_w * m_w;
        const float g2x = dL_dmean2D[3 * i], g2y = dL_dmean2D[3 * i + 1];
        gm[0] += (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        gm[1] += (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        gm[2] += (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;

        if (shs)
            sh_backward(D, M, shs + (size_t)i * M * 3, clamped + 3 * i, px - campos[0], py - campos[1],
                        pz - campos[2], dL_dcolors + 3 * i, gm, dL_dsh + (size_t)i * M * 3);
        dL_dmeans3D[3 * i] = gm[0];
        dL_dmeans3D[3 * i + 1] = gm[1];
        dL_dmeans3D[3 * i + 2] = gm[2];

        if (scales) { /* backward.cu:278-341 */
            const float* q = rotations + 4 * i;
            float r = q[0], x = q[1], y = q[2], z = q[3];
            /* math-convention rotation Rm[row][col]; glm R[c][r] = Rm[r][c] transposed of the literal */
            float Rg[3][3] = {/* glm columns */
                              {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y)},
                              {2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x)},
                              {2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)}};
            float s[3] = {scale_modifier * scales[3 * i], scale_modifier * scales[3 * i + 1],
                          scale_modifier * scales[3 * i + 2]};
            float Mg[3][3]; /* M = S*R : M[c][r] = s[r]*R[c][r] */
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr) Mg[cc][rr] = s[rr] * Rg[cc][rr];
            float dS[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                              {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                              {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            /* dL_dM = 2 * M * dL_dSigma  (glm: (A*B)[c][r] = sum_k A[k][r]*B[c][k]) */
            float dM[3][3];
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr)
                    dM[cc][rr] = 2.0f * (Mg[0][rr] * dS[cc][0] + Mg[1][rr] * dS[cc][1] + Mg[2][rr] * dS[cc][2]);
            /* Rt = transpose(R): Rt[c][r] = R[r][c]; dL_dMt[c][r] = dM[r][c] */
            float dMt[3][3];
            for (int cc = 0; cc < 3; ++cc)
                for (int rr = 0; rr < 3; ++rr) dMt[cc][rr] = dM[rr][cc];
            for (int k = 0; k < 3; ++k)
                dL_dscales[3 * i + k] = Rg[0][k] * dMt[k][0] + Rg[1][k] * dMt[k][1] + Rg[2][k] * dMt[k][2];
            for (int k = 0; k < 3; ++k)
                for (int rr = 0; rr < 3; ++rr) dMt[k][rr] *= s[k];
            float* dq = dL_drots + 4 * i;
            dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
            dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) -
                    4 * x * (dMt[2][2] + dMt[1][1]);
            dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) -
                    4 * y * (dMt[2][2] + dMt[0][0]);
            dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) -
                    4 * z * (dMt[1][1] + dMt[0][0]);
        }
    }
}

/* ---- simple-knn: exact mean squared distance to the 3 nearest neighbours (simple_knn.cu:148-184) ---- */
static inline void upd3(float d, float* best) { /* updateKBest<3>, simple_knn.cu:132-146 */
    for (int j = 0; j < 3; ++j)
        if (best[j] > d) {
            float t = best[j];
            best[j] = d;
            d = t;
        }
}
static int cmp_cell(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}
/*
 * Exact 3-NN via a uniform grid (the reference's Morton/box pruning is also exact, so results agree up to the
 * per-pair distance arithmetic, which follows the nvcc contraction fma(dz,dz, fma(dx,dx, dy*dy))).
 * Missing neighbours (P < 4) leave FLT_MAX in the slots exactly like the reference.
 */
void orc_knn_mean_dist2(int P, const float* pts, float* out) {
    if (P <= 0) return;
    if (P <= 2048) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < P; ++i) {
            float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
            for (int j = 0; j < P; ++j) {
                if (j == i) continue;
                float dx = pts[3 * j] - pts[3 * i], dy = pts[3 * j + 1] - pts[3 * i + 1], dz = pts[3 * j + 2] - pts[3 * i + 2];
                upd3(dot3(dx, dx, dy, dy, dz, dz), best);
            }
            out[i] = (best[0] + best[1] + best[2]) / 3.0f;
        }
        return;
    }
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < P; ++i)
        for (int k = 0; k < 3; ++k) {
            mn[k] = fminf(mn[k], pts[3 * i + k]);
            mx[k] = fmaxf(mx[k], pts[3 * i + k]);
        }
    int G = (int)cbrt((double)P / 4.0);
    if (G < 1) G = 1;
    if (G > 512) G = 512;
    double ext[3], cs[3];
    for (int k = 0; k < 3; ++k) {
        ext[k] = (double)mx[k] - mn[k];
        if (ext[k] <= 0) ext[k] = 1e-30;
        cs[k] = ext[k] / G;
    }
    uint64_t* cell = (uint64_t*)malloc(sizeof(uint64_t) * P); /* (cell id << 32) | point */
    for (int i = 0; i < P; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (int)(((double)pts[3 * i + k] - mn[k]) / cs[k]);
            if (c[k] >= G) c[k] = G - 1;
            if (c[k] < 0) c[k] = 0;
        }
        cell[i] = ((uint64_t)((c[2] * G + c[1]) * G + c[0]) << 32) | (uint32_t)i;
    }
    qsort(cell, P, sizeof(uint64_t), cmp_cell);
    int* start = (int*)malloc(sizeof(int) * ((size_t)G * G * G + 1));
    {
        int p = 0;
        for (int c = 0; c <= G * G * G; ++c) {
            while (p < P && (int)(cell[p] >> 32) < c) ++p;
            start[c] = p;
        }
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < P; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (int)(((double)pts[3 * i + k] - mn[k]) / cs[k]);
            if (c[k] >= G) c[k] = G - 1;
            if (c[k] < 0) c[k] = 0;
        }
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int ring = 0; ring <= G; ++ring) {
            /* all cells at Chebyshev distance == ring */
            for (int dz = -ring; dz <= ring; ++dz)
                for (int dy = -ring; dy <= ring; ++dy)
                    for (int dx = -ring; dx <= ring; ++dx) {
                        if (imax(imax(abs(dx), abs(dy)), abs(dz)) != ring) continue;
                        int x = c[0] + dx, y = c[1] + dy, z = c[2] + dz;
                        if (x < 0 || y < 0 || z < 0 || x >= G || y >= G || z >= G) continue;
                        int cid = (z * G + y) * G + x;
                        for (int p = start[cid]; p < start[cid + 1]; ++p) {
                            int j = (int)(uint32_t)cell[p];
                            if (j == i) continue;
                            float ddx = pts[3 * j] - pts[3 * i], ddy = pts[3 * j + 1] - pts[3 * i + 1],
                                  ddz = pts[3 * j + 2] - pts[3 * i + 2];
                            upd3(dot3(ddx, ddx, ddy, ddy, ddz, ddz), best);
                        }
                    }
            /* anything in ring+1 or beyond is at least ring*min(cs) away (conservative, minus slack) */
            double reach = ring * fmin(cs[0], fmin(cs[1], cs[2]));
            reach *= 0.999;
            if (best[2] < FLT_MAX && (double)best[2] <= reach * reach) break;
        }
        out[i] = (best[0] + best[1] + best[2]) / 3.0f;
    }
    free(cell);
    free(start);
}
