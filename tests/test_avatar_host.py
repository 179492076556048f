"""Caller-side mirror (fateavatar_b200/avatar.py): the closed-form FrameCamera against the reference's camera math
(tools/gs_utils/graphics_utils.py + volume_rendering/camera_3dgs.py:53-72) and, on the GPU, forward_frame / attach
against the separately tested operators."""
import importlib.util
import math
import os
import types

import numpy as np
import pytest
import torch

from fateavatar_b200 import avatar, scenes

REF_GU = "/root/reference/tools/gs_utils/graphics_utils.py"


def _rand_pose(seed):
    g = np.random.default_rng(seed)
    q = g.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R.astype(np.float32), g.standard_normal(3).astype(np.float32)


def test_frame_camera_matches_float64_camera_math():
    for seed in range(4):
        R, T = _rand_pose(seed)
        cam = avatar.FrameCamera(torch.from_numpy(R)[None], torch.from_numpy(T)[None], 0.35, 0.3, (96, 128))
        ref = scenes.make_camera(128, 96, 0.35, 0.3, R=R, T=T)
        assert (cam.image_width, cam.image_height) == (128, 96)
        assert np.abs(cam.world_view_transform.numpy() - ref["viewmatrix"]).max() <= 1e-6
        assert np.abs(cam.full_proj_transform.numpy() - ref["projmatrix"]).max() <= 2e-5 * np.abs(ref["projmatrix"]).max()
        assert np.abs(cam.camera_center.numpy() - ref["campos"]).max() <= 1e-5


@pytest.mark.skipif(not os.path.exists(REF_GU), reason="reference tree not mounted")
def test_frame_camera_matches_reference_graphics_utils():
    spec = importlib.util.spec_from_file_location("ref_graphics_utils", REF_GU)
    gu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gu)
    R, T = _rand_pose(7)
    Rt, Tt = torch.from_numpy(R), torch.from_numpy(T)
    view = gu.getWorld2View2_torch(Rt, Tt).transpose(0, 1)                     # camera_3dgs.py:53
    proj = gu.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=0.35, fovY=0.3).transpose(0, 1)
    full = view.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)                 # camera_3dgs.py:71
    center = view.inverse()[3, :3]                                             # camera_3dgs.py:72
    cam = avatar.FrameCamera(Rt[None], Tt[None], 0.35, 0.3, (64, 64))
    assert float((cam.world_view_transform - view).abs().max()) <= 1e-6        # the reference inverts twice in fp32
    assert float((cam.projection_matrix - proj).abs().max()) <= 1e-6 * float(proj.abs().max())
    assert float((cam.full_proj_transform - full).abs().max()) <= 2e-5 * float(full.abs().max())
    assert float((cam.camera_center - center).abs().max()) <= 1e-5


def test_quaternion_to_axis_angle_restatement_properties():
    g = torch.Generator().manual_seed(0)
    axis = torch.nn.functional.normalize(torch.randn(500, 3, generator=g, dtype=torch.float64), dim=-1)
    ang = torch.rand(500, 1, generator=g, dtype=torch.float64) * 3.0
    q = torch.cat([torch.cos(ang / 2), axis * torch.sin(ang / 2)], dim=-1)
    aa = avatar.quaternion_to_axis_angle(q)
    assert torch.allclose(aa, axis * ang, atol=1e-9)
    assert torch.allclose(avatar.quaternion_to_axis_angle(torch.tensor([[1.0, 0, 0, 0]], dtype=torch.float64)),
                          torch.zeros(1, 3, dtype=torch.float64))


@pytest.mark.gpu
def test_forward_frame_and_attach_equal_the_separate_operators(cuda_device):
    from fateavatar_b200 import flame, pose, rasterizer as R

    dev = cuda_device
    N = 20000
    p = scenes.pose_inputs(N=N, seed=5)
    f = scenes.flame_inputs(seed=5)
    d = lambda a: torch.from_numpy(a).to(dev)
    fl = types.SimpleNamespace(n_shape=f["n_shape"], n_exp=f["n_exp"], parents=torch.from_numpy(f["parents"]))
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
        setattr(fl, k, d(f[k]))
    from oracle import pose_oracle as po
    _, canon = po.compute_face_orientation(d(f["v_template"])[None], d(p["faces"]))
    par = lambda a: torch.nn.Parameter(d(a))
    model = types.SimpleNamespace(
        flame=fl, faces=d(p["faces"]), face_index=d(p["face_index"]), bary_coords=d(p["bary"]), face_scaling_canonical=canon,
        _scaling=par(p["scaling_raw"]), _rotation=par(p["rotation_raw"]), _offset=par(p["offset_raw"]), _opacity=par(p["opacity_raw"]),
        _features_dc=par(((np.random.default_rng(0).uniform(0, 1, (N, 1, 3)) - 0.5) / scenes.SH_C0).astype(np.float32)),
        delta_shapedirs=par(f["delta_shapedirs"]), delta_posedirs=par(f["delta_posedirs"]), delta_vertex=par(f["delta_vertex"]),
        cfg_model=types.SimpleNamespace(delta_blendshape=True, delta_vertex=True, resize_scale=True), shell_len=0.05,
        bg_color=torch.ones(3), img_res=(128, 160), device=dev, xyz_gradient_accum=torch.zeros(N, 1, device=dev),
        denom=torch.zeros(N, 1, device=dev), _add_densification_stats=None, forward=None)
    avatar.attach(model)
    Rm, T = np.diag([1.0, -1.0, -1.0]).astype(np.float32), np.array([0, 0, 1.25], np.float32)
    pose_c = np.eye(4, dtype=np.float32)
    pose_c[:3, :3], pose_c[:3, 3] = Rm, T
    inp = dict(cam_pose=d(pose_c)[None], fovx=torch.tensor([0.35]), fovy=torch.tensor([0.35]),
               flame_pose=d(f["pose"])[None], expression=d(f["betas"][300:])[None])
    out = model.forward(inp)
    assert out["rgb_image"].shape == (1, 3, 128, 160) and out["verts"].shape == (1, f["v_template"].shape[0], 3)
    assert out["raw_rot"].shape == (N, 3) and out["scale"].shape == (N, 3) and out["bs"] == 1
    loss = (out["rgb_image"] - 0.5).abs().mean() + 1e-2 * (out["verts"] - out["verts_orig"]).pow(2).sum()
    loss.backward()
    model._add_densification_stats(out["viewspace_points"][0], out["visibility_filter"][0])
    assert float(model.denom.sum()) == float(out["visibility_filter"][0].sum()) and float(model.xyz_gradient_accum.sum()) > 0
    for name in ("_scaling", "_rotation", "_offset", "_opacity", "_features_dc", "delta_shapedirs", "delta_posedirs", "delta_vertex"):
        assert getattr(model, name).grad is not None and torch.isfinite(getattr(model, name).grad).all(), name
    # the same frame from the separately tested operators
    fm = flame.model_tensors(fl)
    betas = torch.cat([torch.zeros(1, 300, device=dev), inp["expression"]], dim=1)
    verts, _, _, verts_orig, _ = flame.flame_lbs(fm, betas, inp["flame_pose"], model.delta_shapedirs, model.delta_posedirs,
                                                 model.delta_vertex, l0=300)
    xyz, sc, ro, op = pose.pose_splats(verts, model.faces, model.face_index, model.bary_coords, canon, model._scaling,
                                       model._rotation, model._offset, model._opacity, shell_len=0.05)
    cam = scenes.make_camera(160, 128, 0.35, 0.35, R=Rm, T=T)
    rs = R.GaussianRasterizationSettings(128, 160, cam["tanfovx"], cam["tanfovy"], torch.ones(3, device=dev), 1.0,
                                         d(cam["viewmatrix"]), d(cam["projmatrix"]), 0, d(cam["campos"]), False, False)
    img, radii = R.GaussianRasterizer(rs)(means3D=xyz, means2D=torch.zeros_like(xyz), shs=model._features_dc, opacities=op,
                                          scales=sc, rotations=ro)
    assert torch.equal(out["verts"], verts) and torch.equal(out["verts_orig"], verts_orig)
    assert float((out["rgb_image"][0].detach() - img.detach()).abs().max()) <= 1e-4  # camera matrices differ in the last fp32 bit
    assert float((out["radii"][0] != radii).float().mean()) <= 1e-3
