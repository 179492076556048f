"""Row R0 on the CPU: the reference's UNCHANGED volume_rendering/render_3dgs.py:render(..., device='cpu') executed on the
oracle-backed operator drop-in (oracle/cpu_dropin.py) versus this repo's restatement of that caller
(fateavatar_b200/render.py) on the same drop-in -- same image, radii, visibility and gradients, bit for bit; and the
config-1 CPU baseline that BASELINE.md 2b describes."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from fateavatar_b200 import render as rmod
from fateavatar_b200 import scenes
from oracle import cpu_dropin
from oracle import oracle as orc
from util import oracle_forward

REF_RENDER = "/root/reference/volume_rendering/render_3dgs.py"


def _camera_and_cloud(sc, requires_grad=True):
    cam = sc["camera"]
    t = lambda a: torch.from_numpy(a)
    camera = rmod.MiniCam(cam["W"], cam["H"], cam["fovy"], cam["fovx"], t(cam["viewmatrix"]), t(cam["projmatrix"]), t(cam["campos"]))
    leaf = lambda a: t(a.copy()).requires_grad_(requires_grad)
    # raw parameters whose activations (exp / normalize / sigmoid) reproduce the scene's tensors
    cloud = rmod.SplatCloud(leaf(sc["means3D"]), leaf(sc["shs"]), leaf(np.log(sc["scales"])), leaf(sc["rotations"]),
                            leaf(np.log(sc["opacities"] / (1 - sc["opacities"]))), max_sh_degree=sc["sh_degree"])
    return camera, cloud


def _grads(out, cloud, seed=0):
    g = torch.from_numpy(np.random.default_rng(seed).standard_normal(tuple(out["render"].shape)).astype(np.float32))
    (out["render"] * g).sum().backward()
    return [p.grad.clone() for p in (cloud._xyz, cloud._features, cloud._scaling, cloud._rotation, cloud._opacity)] + \
        [out["viewspace_points"].grad.clone()]


@pytest.mark.skipif(not os.path.exists(REF_RENDER), reason="reference tree not mounted")
@pytest.mark.parametrize("deg", [0, 2])
def test_unchanged_reference_render_equals_the_mirror_on_the_cpu_oracle(deg, monkeypatch):
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization", cpu_dropin)
    spec = importlib.util.spec_from_file_location("ref_render_3dgs_cpu", REF_RENDER)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    assert ref.GaussianRasterizer is cpu_dropin.GaussianRasterizer
    sc = scenes.head_scene(P=1500, W=96, H=80, sh_degree=deg, scale_mult=8.0, seed=20 + deg)
    bg = torch.from_numpy(sc["bg"])
    cam_a, cloud_a = _camera_and_cloud(sc)
    out_ref = ref.render(cam_a, cloud_a, bg, device="cpu")                      # the reference's own caller code
    cam_b, cloud_b = _camera_and_cloud(sc)
    out_new = rmod.render(cam_b, cloud_b, bg, device="cpu", rasterizer_module=cpu_dropin)
    assert set(out_ref) == set(out_new) == {"render", "viewspace_points", "visibility_filter", "radii"}
    assert torch.equal(out_ref["render"], out_new["render"]) and torch.equal(out_ref["radii"], out_new["radii"])
    assert torch.equal(out_ref["visibility_filter"], out_new["visibility_filter"])
    for a, b in zip(_grads(out_ref, cloud_a), _grads(out_new, cloud_b)):
        assert torch.equal(a, b)
    # and both equal a direct oracle call on the activated tensors
    o = oracle_forward(orc, sc)
    assert np.abs(out_ref["render"].detach().numpy() - o["color"]).max() <= 2e-6
    # override_color path (render_3dgs.py:56-64)
    col = torch.rand(1500, 3)
    o1 = ref.render(cam_a, cloud_a, bg, override_color=col, device="cpu")
    o2 = rmod.render(cam_b, cloud_b, bg, override_color=col, device="cpu", rasterizer_module=cpu_dropin)
    assert torch.equal(o1["render"], o2["render"])


def test_cpu_dropin_keeps_the_reference_argument_checks_and_config1_runs():
    sc = scenes.config1_scene(P=2000, W=128, H=128)                            # BASELINE config 1, reduced
    cam, cloud = _camera_and_cloud(sc, requires_grad=False)
    out = rmod.render(cam, cloud, torch.from_numpy(sc["bg"]), device="cpu", rasterizer_module=cpu_dropin)
    o = oracle_forward(orc, sc)
    assert out["render"].shape == (3, 128, 128) and np.array_equal(out["radii"].numpy(), o["radii"])
    rs = cpu_dropin.GaussianRasterizationSettings(16, 16, 0.2, 0.2, torch.ones(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                                  torch.zeros(3), False, False)
    r = cpu_dropin.GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(x, x, torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(x, x, torch.ones(4, 1), shs=torch.zeros(4, 1, 3))


REF_ROOT = "/root/reference"


def _load(name, path, monkeypatch):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, name, mod)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.exists(REF_RENDER), reason="reference tree not mounted")
def test_reference_gaussian_model_and_render_on_cpu_match_pose_stage_activations(monkeypatch):
    """The reference's own GaussianModel (volume_rendering/gaussian_model.py:105-128 getters) filled the way
    FateAvatar.forward fills it (model/fateavatar.py:244-258) and rendered by its own render(): equals this repo's
    SplatCloud + render on the same raw parameters.  plyfile and the CUDA knn extension are stubbed -- they are imported
    by that file but not used on this path."""
    import types

    stub_ply = types.ModuleType("plyfile")
    stub_ply.PlyData = stub_ply.PlyElement = object
    monkeypatch.setitem(sys.modules, "plyfile", stub_ply)
    knn_pkg, knn_c = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
    knn_c.distCUDA2 = lambda pts: torch.from_numpy(orc.knn_mean_dist2(pts.detach().cpu().numpy()))
    monkeypatch.setitem(sys.modules, "simple_knn", knn_pkg)
    monkeypatch.setitem(sys.modules, "simple_knn._C", knn_c)
    for pkg in ("tools", "tools.gs_utils"):  # the reference's package names, loaded from its files by path
        m = types.ModuleType(pkg)
        m.__path__ = []
        monkeypatch.setitem(sys.modules, pkg, m)
    for mod in ("general_utils", "system_utils", "sh_utils", "graphics_utils"):
        _load(f"tools.gs_utils.{mod}", f"{REF_ROOT}/tools/gs_utils/{mod}.py", monkeypatch)
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization", cpu_dropin)
    gm = _load("ref_gaussian_model_cpu", f"{REF_ROOT}/volume_rendering/gaussian_model.py", monkeypatch)
    ref = _load("ref_render_3dgs_cpu2", REF_RENDER, monkeypatch)

    sc = scenes.head_scene(P=1200, W=80, H=64, sh_degree=0, scale_mult=8.0, seed=31)
    cam, cloud = _camera_and_cloud(sc)
    gaussian = gm.GaussianModel(sh_degree=0)
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    gaussian._xyz, gaussian._features_dc = leaf(cloud._xyz), leaf(cloud._features)
    gaussian._features_rest = torch.zeros(1200, 0, 3)
    gaussian._scaling, gaussian._rotation, gaussian._opacity = leaf(cloud._scaling), leaf(cloud._rotation), leaf(cloud._opacity)
    bg = torch.from_numpy(sc["bg"])
    out_ref = ref.render(cam, gaussian, bg, device="cpu")
    out_new = rmod.render(cam, cloud, bg, device="cpu", rasterizer_module=cpu_dropin)
    assert torch.equal(out_ref["render"], out_new["render"]) and torch.equal(out_ref["radii"], out_new["radii"])
    g = torch.from_numpy(np.random.default_rng(1).standard_normal((3, 64, 80)).astype(np.float32))
    (out_ref["render"] * g).sum().backward()
    (out_new["render"] * g).sum().backward()
    for a, b in ((gaussian._xyz, cloud._xyz), (gaussian._features_dc, cloud._features), (gaussian._scaling, cloud._scaling),
                 (gaussian._rotation, cloud._rotation), (gaussian._opacity, cloud._opacity)):
        assert torch.equal(a.grad, b.grad)
