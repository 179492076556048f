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
    # exact mode repeats the reference's own library calls: every matrix carries the same fp32 bits
    ex = avatar.FrameCamera(Rt[None], Tt[None], 0.35, 0.3, (64, 64), exact=True)
    assert torch.equal(ex.world_view_transform, view) and torch.equal(ex.projection_matrix, proj)
    assert torch.equal(ex.full_proj_transform, full) and torch.equal(ex.camera_center, center)


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


REF = "/root/reference"
GOLDEN_FRAME = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frame_small.npz")
FRAME_CASE = dict(seed=51, n_lat=9, n_lon=16, N=900)   # tests/golden/make_frame_golden.py
FRAME_RES = (64, 80)


def oracle_frame(a, inp, leaves):
    """This repo's oracle composition of one frame on the CPU (what the fused GPU path is tested against):
    FrameCamera + flame_oracle + pose_oracle.pose_splats + the C oracle rasterizer behind the operator API."""
    from oracle import cpu_dropin, flame_oracle as fo, pose_oracle as po

    t = lambda x: torch.from_numpy(np.asarray(x))
    fm = {k: t(a[k]) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
    fm["parents"] = t(a["parents"])
    faces = t(a["faces"])
    _, canon = po.compute_face_orientation(fm["v_template"], faces)
    betas = torch.cat([torch.zeros(a["n_shape"]), inp["expression"][0]])
    verts, _, _ = fo.forward_with_delta_blendshape(fm, betas, inp["flame_pose"][0], leaves["delta_shapedirs"],
                                                   leaves["delta_posedirs"], leaves["delta_vertex"])
    verts_orig, _, _ = fo.forward_with_delta_blendshape(fm, betas, inp["flame_pose"][0])
    xyz, sc, ro, op = po.pose_splats(verts, faces, t(a["face_index"]), t(a["bary"]), canon, leaves["_scaling"], leaves["_rotation"],
                                     leaves["_offset"], leaves["_opacity"], shell_len=a["shell_len"])
    fx, fy = float(inp["fovx"][0]), float(inp["fovy"][0])
    cam = avatar.FrameCamera(inp["cam_pose"][:, :3, :3], inp["cam_pose"][:, :3, 3], fx, fy, FRAME_RES)
    rs = cpu_dropin.GaussianRasterizationSettings(FRAME_RES[0], FRAME_RES[1], math.tan(fx / 2), math.tan(fy / 2), torch.ones(3), 1.0,
                                                  cam.world_view_transform, cam.full_proj_transform, 0, cam.camera_center,
                                                  False, False)
    img, radii = cpu_dropin.GaussianRasterizer(rs)(means3D=xyz, means2D=torch.zeros_like(xyz), shs=leaves["_features_dc"],
                                                   opacities=op, scales=sc, rotations=ro)
    return img, radii, verts, verts_orig


@pytest.mark.skipif(not os.path.exists(f"{REF}/model/fateavatar.py"), reason="reference tree not mounted")
def test_reference_FateAvatar_forward_runs_on_cpu_and_matches_the_oracle_composition(monkeypatch):
    """The reference's UNCHANGED `FateAvatar.forward` (model/fateavatar.py:196-298) executed on the CPU through
    tests/ref_frame_harness.py must equal the composition this repo uses as oracle for the fused path (FrameCamera +
    flame_oracle + pose_oracle.pose_splats + rasterizer oracle): image, meshes, radii and every parameter gradient."""
    import ref_frame_harness as H

    FateAvatar, FLAME, mesh_compute = H.load_reference(monkeypatch)
    a = scenes.small_avatar(**FRAME_CASE)
    ref = H.build_reference_model(FateAvatar, FLAME, mesh_compute, a, FRAME_RES)
    inp = H.frame_input(a)
    out = FateAvatar.forward(ref, inp)                                     # <- the reference's own code
    assert out["rgb_image"].shape == (1, 3) + FRAME_RES and out["bs"] == 1 and int((out["radii"][0] > 0).sum()) > 500
    leaves = {n: getattr(ref, n) for n in H.PARAMS}
    img, radii, verts, verts_orig = oracle_frame(a, inp, leaves)
    # measured: image 2.4e-6, meshes 3e-8, radii identical, gradients <= 1e-5 (the reference inverts its camera twice)
    assert float((out["verts"][0] - verts).abs().max()) <= 1e-6 and float((out["verts_orig"][0] - verts_orig).abs().max()) <= 1e-6
    assert float((out["rgb_image"][0] - img).abs().max()) <= 2e-5
    assert float((out["radii"][0] != radii).float().mean()) <= 1e-3
    assert torch.equal(out["visibility_filter"][0], out["radii"][0] > 0)
    assert torch.equal(out["scale"], torch.exp(ref._scaling))
    assert torch.equal(out["raw_rot"], avatar.quaternion_to_axis_angle(ref._rotation))
    w = torch.from_numpy(np.random.default_rng(2).standard_normal((3,) + FRAME_RES).astype(np.float32))
    (out["rgb_image"][0] * w).sum().backward()
    g_ref = {n: getattr(ref, n).grad.clone() for n in H.PARAMS}
    for n in H.PARAMS:
        getattr(ref, n).grad = None
    (img * w).sum().backward()
    for n in H.PARAMS:
        a_, b_ = getattr(ref, n).grad, g_ref[n]
        assert float((a_ - b_).abs().max()) <= 1e-4 * max(float(b_.abs().max()), 1e-12), n


def test_oracle_frame_matches_golden_from_the_reference_forward():
    """tests/golden/frame_small.npz holds what the reference's own FateAvatar.forward produced (make_frame_golden.py)."""
    import ref_frame_harness as H

    gold = np.load(GOLDEN_FRAME)
    a = scenes.small_avatar(**FRAME_CASE)
    t = lambda x: torch.from_numpy(np.asarray(x))
    keys = dict(_scaling="scaling_raw", _rotation="rotation_raw", _offset="offset_raw", _opacity="opacity_raw", _features_dc="features_dc",
                delta_vertex="delta_vertex", delta_posedirs="delta_posedirs", delta_shapedirs="delta_shapedirs")
    leaves = {n: t(a[k]).clone().requires_grad_(True) for n, k in keys.items()}
    img, radii, verts, verts_orig = oracle_frame(a, H.frame_input(a), leaves)
    assert np.abs(img.detach().numpy() - gold["rgb_image"]).max() <= 2e-5 and np.abs(verts.detach().numpy() - gold["verts"]).max() <= 1e-6
    assert (radii.numpy() != gold["radii"]).mean() <= 1e-3
    w = torch.from_numpy(np.random.default_rng(2).standard_normal((3,) + FRAME_RES).astype(np.float32))
    (img * w).sum().backward()
    for n in H.PARAMS:
        g, want = leaves[n].grad.numpy(), gold["grad" + n]
        if n == "delta_shapedirs":
            assert not g[:, :, :300].any()
            g = g[:, :, 300:]
        assert np.abs(g - want).max() <= 1e-4 * max(np.abs(want).max(), 1e-12), n
