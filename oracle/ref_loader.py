"""Loader for the compiled reference extensions in oracle/_ref (TEST INFRASTRUCTURE; see build_ref.py).

`ref_dgr()` / `ref_knn()` import the reference's own pybind modules under private names so they can coexist
with the drop-in; `ref_forward` / `ref_backward` call them the way DGR diff_gaussian_rasterization/__init__.py:
44-155 does and decode the opaque geom/binning/img byte buffers into named tensors using the chunk layout of
DGR cuda_rasterizer/rasterizer_impl.cu:155-194 (every array 128-byte aligned, in declaration order).
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mods = {}


def _load(name):
    if name in _mods:
        return _mods[name]
    path = os.path.join(_HERE, "_ref", name, "_C.so")
    if not os.path.exists(path):
        _mods[name] = None
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(f"{name}._C", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mods[name] = mod
    return mod


def ref_dgr():
    return _load("ref_dgr")


def ref_knn():
    return _load("ref_knn")


def available():
    return ref_dgr() is not None


def _al(x, a=128):
    return (x + a - 1) // a * a


def decode_buffers(geom, binning, img, P, R, W, H):
    import torch

    def take(buf, off, nbytes, dtype, shape):
        off = _al(off)
        return buf[off:off + nbytes].view(dtype).view(shape), off + nbytes

    out = {}
    off = 0
    out["depths"], off = take(geom, off, P * 4, torch.float32, (P,))
    out["clamped"], off = take(geom, off, P * 3, torch.uint8, (P, 3))
    out["internal_radii"], off = take(geom, off, P * 4, torch.int32, (P,))
    out["means2D"], off = take(geom, off, P * 8, torch.float32, (P, 2))
    out["cov3D"], off = take(geom, off, P * 24, torch.float32, (P, 6))
    out["conic_opacity"], off = take(geom, off, P * 16, torch.float32, (P, 4))
    out["rgb"], off = take(geom, off, P * 12, torch.float32, (P, 3))
    out["tiles_touched"], off = take(geom, off, P * 4, torch.int32, (P,))
    off = 0
    out["point_list"], off = take(binning, off, R * 4, torch.int32, (R,))
    _, off = take(binning, off, R * 4, torch.int32, (R,))
    out["point_list_keys"], off = take(binning, off, R * 8, torch.int64, (R,))
    N = W * H
    off = 0
    out["final_T"], off = take(img, off, N * 4, torch.float32, (H, W))
    out["n_contrib"], off = take(img, off, N * 4, torch.int32, (H, W))
    Tn = ((W + 15) // 16) * ((H + 15) // 16)
    rng, off = take(img, off, N * 8, torch.int32, (N, 2))
    out["ranges"] = rng[:Tn]
    return out


def ref_forward(t, cam, *, sh_degree=0, use_colors_precomp=False, scale_modifier=1.0):
    """t: dict of CUDA tensors (means3D, opacities, scales, rotations, shs [, colors_precomp], bg);
    cam: dict with CUDA viewmatrix/projmatrix/campos + tanfovx/tanfovy/W/H."""
    import torch

    C = ref_dgr()
    e = torch.Tensor([])
    sh = e if use_colors_precomp else t["shs"]
    cp = t["colors_precomp"] if use_colors_precomp else e
    args = (t["bg"], t["means3D"], cp, t["opacities"], t["scales"], t["rotations"], scale_modifier, e,
            cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], cam["H"], cam["W"], sh, sh_degree,
            cam["campos"], False, False)
    R, color, radii, geom, binning, img = C.rasterize_gaussians(*args)
    torch.cuda.synchronize()
    st = dict(R=R, color=color, radii=radii, geom=geom, binning=binning, img=img, args=args)
    st.update(decode_buffers(geom, binning, img, t["means3D"].shape[0], R, cam["W"], cam["H"]))
    return st


def ref_backward(st, dL_dpix):
    import torch

    C = ref_dgr()
    a = st["args"]
    (bg, means3D, cp, opac, scales, rots, smod, cov3, view, proj, tfx, tfy, H, W, sh, deg, campos, _pf, dbg) = a
    out = C.rasterize_gaussians_backward(bg, means3D, st["radii"], cp, scales, rots, smod, cov3, view, proj, tfx, tfy,
                                         dL_dpix, sh, deg, campos, st["geom"], st["R"], st["binning"], st["img"], dbg)
    torch.cuda.synchronize()
    names = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations")
    return dict(zip(names, out))
