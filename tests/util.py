"""Shared helpers for the parity tests."""
import numpy as np

GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations")


def oracle_forward(orc, sc, **kw):
    cam = sc["camera"]
    args = dict(shs=sc.get("shs"), sh_degree=sc.get("sh_degree", 0), scales=sc.get("scales"),
                rotations=sc.get("rotations"))
    args.update(kw)
    return orc.forward(sc["means3D"], sc["opacities"], sc["bg"], cam["viewmatrix"], cam["projmatrix"], cam["campos"],
                       cam["tanfovx"], cam["tanfovy"], cam["H"], cam["W"], **args)


def settings(R, cam, bg, deg, scale_modifier=1.0):
    return R.GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], bg, scale_modifier,
                                           cam["viewmatrix"], cam["projmatrix"], deg, cam["campos"], False, False)


def assert_grad_close(name, got, ref, rtol=2e-4):
    """Gradients are sums of many fp32 terms accumulated in a different order (the reference itself uses float
    atomics): compare with a tolerance relative to the largest reference magnitude of that tensor."""
    got = np.asarray(got, np.float64).reshape(np.asarray(ref).shape)
    ref = np.asarray(ref, np.float64)
    scale = max(float(np.abs(ref).max()), 1e-12) if ref.size else 1.0
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    assert err <= rtol * scale, f"{name}: max abs err {err:.3e} > {rtol:.0e} * {scale:.3e}"


def rasterizer_goldens(here):
    """The golden fixtures of the rasterizer (tests/golden/make_golden.py): the .npz files that carry a `scene` key.
    Other stages keep their own fixtures in the same directory (flame_*, pose_*, frame_*)."""
    import glob
    import os

    out = []
    for path in sorted(glob.glob(os.path.join(here, "golden", "*.npz"))):
        with np.load(path) as z:
            if "scene" in z.files:
                out.append(path)
    return out
