"""ctypes/numpy front end of the C oracle (oracle/splat_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  Nothing under fateavatar_b200/ may import this module.

`forward()` mirrors CudaRasterizer::Rasterizer::forward (DGR cuda_rasterizer/rasterizer_impl.cu:198-336)
and returns every intermediate the reference keeps in geomBuffer / binningBuffer / imgBuffer
(rasterizer_impl.cu:155-194) as named arrays; `backward()` mirrors Rasterizer::backward (:340-434).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "splat_oracle.c")
        if (not os.path.exists(path)) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            import importlib.util

            spec = importlib.util.spec_from_file_location("_orc_build", os.path.join(_HERE, "build_oracle.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        _LIB = C.CDLL(path)
        _LIB.orc_bin_sort.restype = C.c_int64
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    if shape is not None:
        a = a.reshape(shape)
    return a


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def forward(means3D, opacities, bg, view, proj, campos, tanfovx, tanfovy, H, W, *, shs=None, sh_degree=0,
            colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0,
            stages="all"):
    """Run the oracle forward.  view/proj are the *transposed* matrices the reference API takes
    (flattened row-major => column-major for the kernels)."""
    L = lib()
    means3D = _f32(means3D, (-1, 3))
    P = means3D.shape[0]
    opacities = _f32(opacities, (-1,))
    shs_a = _f32(shs)
    M = 0 if shs_a is None or shs_a.size == 0 else shs_a.shape[1]
    if M == 0:
        shs_a = None
    cp = _f32(colors_precomp)
    if cp is not None and cp.size == 0:
        cp = None
    sc, ro, c3 = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    if sc is not None and sc.size == 0:
        sc = None
    if ro is not None and ro.size == 0:
        ro = None
    if c3 is not None and c3.size == 0:
        c3 = None
    view, proj, campos, bg = _f32(view, (16,)), _f32(proj, (16,)), _f32(campos, (3,)), _f32(bg, (3,))
    Tn = ((W + 15) // 16) * ((H + 15) // 16)
    st = dict(
        P=P, M=M, D=int(sh_degree), W=int(W), H=int(H),
        radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
        cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
        conic_opacity=np.zeros((P, 4), np.float32), clamped=np.zeros((P, 3), np.uint8),
        tiles_touched=np.zeros(P, np.uint32), point_offsets=np.zeros(P, np.uint32),
        ranges=np.zeros((Tn, 2), np.uint32),
    )
    st["inputs"] = dict(means3D=means3D, opacities=opacities, shs=shs_a, colors_precomp=cp, scales=sc, rotations=ro,
                        cov3D_precomp=c3, view=view, proj=proj, campos=campos, bg=bg, tanfovx=float(tanfovx),
                        tanfovy=float(tanfovy), scale_modifier=float(scale_modifier))
    if P == 0:
        st.update(R=0, point_list=np.zeros(0, np.uint32), keys=np.zeros(0, np.uint64),
                  color=np.broadcast_to(bg[:, None, None], (3, H, W)).copy() * 0.0,
                  final_T=np.zeros((H, W), np.float32), n_contrib=np.zeros((H, W), np.uint32),
                  fragile=np.zeros((H, W), np.uint8))
        return st
    L.orc_preprocess(C.c_int(P), C.c_int(int(sh_degree)), C.c_int(M), _p(means3D), _p(sc),
                     C.c_float(scale_modifier), _p(ro), _p(opacities), _p(shs_a), _p(c3), _p(cp), _p(view), _p(proj),
                     _p(campos), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy), _p(st["radii"]),
                     _p(st["means2D"]), _p(st["depths"]), _p(st["cov3D"]), _p(st["rgb"]), _p(st["conic_opacity"]),
                     _p(st["clamped"]), _p(st["tiles_touched"]))
    if stages == "preprocess":
        return st
    R = int(L.orc_bin_sort(C.c_int(P), C.c_int(W), C.c_int(H), _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]),
                           _p(st["tiles_touched"]), _p(st["point_offsets"]), None, None, None))
    st["R"] = R
    st["point_list"] = np.zeros(R, np.uint32)
    st["keys"] = np.zeros(R, np.uint64)
    L.orc_bin_sort(C.c_int(P), C.c_int(W), C.c_int(H), _p(st["radii"]), _p(st["means2D"]), _p(st["depths"]),
                   _p(st["tiles_touched"]), _p(st["point_offsets"]), _p(st["keys"]), _p(st["point_list"]),
                   _p(st["ranges"]))
    if stages == "sort":
        return st
    st["color"] = np.zeros((3, H, W), np.float32)
    st["final_T"] = np.zeros((H, W), np.float32)
    st["n_contrib"] = np.zeros((H, W), np.uint32)
    st["fragile"] = np.zeros((H, W), np.uint8)
    feat = cp if cp is not None else st["rgb"]
    L.orc_blend_forward(C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(st["means2D"]), _p(feat),
                        _p(st["conic_opacity"]), _p(bg), _p(st["color"]), _p(st["final_T"]), _p(st["n_contrib"]),
                        _p(st["fragile"]))
    return st


def backward(st, dL_dpix):
    """Oracle backward for a state returned by forward().  Returns the 8 gradients of
    RasterizeGaussiansBackwardCUDA (DGR rasterize_points.cu:117-196) plus dL_dconic."""
    L = lib()
    P, M, D, W, H = st["P"], st["M"], st["D"], st["W"], st["H"]
    inp = st["inputs"]
    g = dict(
        dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 2, 2), np.float32),
        dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
        dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
        dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
        dL_drotations=np.zeros((P, 4), np.float32),
    )
    if P == 0:
        return g
    dL_dpix = _f32(dL_dpix, (3, H, W))
    feat = inp["colors_precomp"] if inp["colors_precomp"] is not None else st["rgb"]
    L.orc_blend_backward(C.c_int(P), C.c_int(W), C.c_int(H), _p(st["ranges"]), _p(st["point_list"]), _p(inp["bg"]),
                         _p(st["means2D"]), _p(st["conic_opacity"]), _p(feat), _p(st["final_T"]), _p(st["n_contrib"]),
                         _p(dL_dpix), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
                         _p(g["dL_dcolors"]))
    cov = inp["cov3D_precomp"] if inp["cov3D_precomp"] is not None else st["cov3D"]
    L.orc_preprocess_backward(C.c_int(P), C.c_int(D), C.c_int(M), _p(inp["means3D"]), _p(st["radii"]), _p(inp["shs"]),
                              _p(st["clamped"]), _p(inp["scales"]), _p(inp["rotations"]),
                              C.c_float(inp["scale_modifier"]), _p(cov), _p(inp["view"]), _p(inp["proj"]), C.c_int(W),
                              C.c_int(H), C.c_float(inp["tanfovx"]), C.c_float(inp["tanfovy"]), _p(inp["campos"]),
                              _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]),
                              _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, view, proj):
    means3D = _f32(means3D, (-1, 3))
    out = np.zeros(means3D.shape[0], np.uint8)
    lib().orc_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(_f32(view, (16,))), _p(_f32(proj, (16,))), _p(out))
    return out.astype(bool)


def knn_mean_dist2(points):
    points = _f32(points, (-1, 3))
    out = np.zeros(points.shape[0], np.float32)
    lib().orc_knn_mean_dist2(C.c_int(points.shape[0]), _p(points), _p(out))
    return out


def pair_counts(st):
    """(walked, exp_evaluated, contributing) pixel-splat pairs of the reference blend forward for the frame `st`."""
    c = np.zeros(3, np.uint64)
    lib().orc_blend_pair_counts(C.c_int(st["W"]), C.c_int(st["H"]), _p(st["ranges"]), _p(st["point_list"]),
                                _p(st["means2D"]), _p(st["conic_opacity"]), _p(c))
    return dict(walked=int(c[0]), exp_evaluated=int(c[1]), contributing=int(c[2]))


def mask_fragile(o, dL_dpix):
    """Upstream gradient with the oracle's threshold-fragile pixels zeroed.  A fragile pixel's list may gain or lose
    one 1/255-contribution between glibc expf and CUDA expf (see compare_blend); feeding both sides a zero upstream
    gradient there keeps that single discrete flip out of a gradient comparison that is otherwise tight."""
    g = np.array(dL_dpix, np.float32, copy=True)
    g[:, o["fragile"].astype(bool)] = 0.0
    return g


def compare_blend(o, color, final_T=None, n_contrib=None, tol=1e-5, fragile_tol=6e-3, max_fragile_frac=5e-3):
    """Compare a CUDA blend result with the oracle state `o`.  Pixels the oracle flags as fragile (a discrete
    decision within rounding distance of its threshold, see orc_blend_forward) may differ by one dropped/added
    1/255-contribution; everything else must agree to `tol`.  Returns a dict of the measured maxima."""
    frag = o["fragile"].astype(bool)
    d = np.abs(np.asarray(color, np.float32) - o["color"]).max(0)
    res = dict(max_solid=float(d[~frag].max()) if (~frag).any() else 0.0,
               max_fragile=float(d[frag].max()) if frag.any() else 0.0, fragile_frac=float(frag.mean()))
    assert res["fragile_frac"] <= max_fragile_frac, res
    if res["max_solid"] > tol:  # say where: the pixel, both colours, its list length and transmittance in the oracle
        dd = np.where(frag, 0.0, d)
        y, x = np.unravel_index(int(dd.argmax()), dd.shape)
        res["worst_pixel"] = dict(y=int(y), x=int(x), got=np.asarray(color, np.float32)[:, y, x].tolist(),
                                  want=o["color"][:, y, x].tolist(), n_contrib=int(o["n_contrib"][y, x]),
                                  final_T=float(o["final_T"][y, x]), n_bad=int((dd > tol).sum()))
    assert res["max_solid"] <= tol, res
    assert res["max_fragile"] <= fragile_tol, res
    if final_T is not None:
        dT = np.abs(np.asarray(final_T, np.float32) - o["final_T"])
        assert (dT[~frag].max() if (~frag).any() else 0.0) <= tol, float(dT[~frag].max())
    if n_contrib is not None:
        bad = (np.asarray(n_contrib).astype(np.int64) != o["n_contrib"].astype(np.int64)) & ~frag
        assert bad.mean() <= 1e-5, float(bad.mean())
    return res
