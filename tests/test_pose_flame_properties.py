"""Mathematical contracts of the FLAME and pose-stage oracles (CPU, float64): identities and rigid-motion equivariance
that hold for the reference's formulation whatever the implementation.  They pin the parts no reference fixture can
(the pytorch3d quaternion path) and guard the oracles the GPU kernels are compared with."""
import numpy as np
import torch

from fateavatar_b200 import scenes
from oracle import flame_oracle as fo
from oracle import pose_oracle as po


def _rot(seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(4, generator=g, dtype=torch.float64), dim=0)
    w, x, y, z = q
    return torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
                        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
                        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])


def _quat_to_mat(q):
    r, x, y, z = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


def _flame(seed=2, V=80):
    f = scenes.flame_inputs(seed=seed, V=V)
    t = lambda k: torch.from_numpy(f[k]).double()
    m = {k: t(k) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
    m["parents"] = torch.from_numpy(f["parents"])
    return f, m, t


def test_flame_rest_pose_and_blendshape_linearity():
    f, m, t = _flame()
    zero_pose = torch.zeros(15, dtype=torch.float64)
    # rest pose, zero coefficients: the template plus the vertex delta (Rodrigues' 1e-8 leaves ~1e-8 of rotation)
    v, pf, A = fo.forward_with_delta_blendshape(m, torch.zeros(400, dtype=torch.float64), zero_pose, delta_vertex=t("delta_vertex"))
    assert torch.allclose(v, m["v_template"] + t("delta_vertex"), atol=1e-7) and pf.abs().max() < 1e-7
    assert torch.allclose(A[:, :3, :3], torch.eye(3, dtype=torch.float64).expand(5, 3, 3), atol=1e-7)
    # at rest pose the mesh is linear in the coefficients
    b1, b2 = t("betas"), torch.roll(t("betas"), 7)
    va, _, _ = fo.forward_with_delta_blendshape(m, b1, zero_pose)
    vb, _, _ = fo.forward_with_delta_blendshape(m, b2, zero_pose)
    vab, _, _ = fo.forward_with_delta_blendshape(m, b1 + 2 * b2, zero_pose)
    v0, _, _ = fo.forward_with_delta_blendshape(m, torch.zeros(400, dtype=torch.float64), zero_pose)
    assert torch.allclose(vab - v0, (va - v0) + 2 * (vb - v0), atol=1e-7)
    # deltas add to the corresponding model tensors
    vd, _, _ = fo.forward_with_delta_blendshape(m, b1, t("pose"), t("delta_shapedirs"), t("delta_posedirs"), t("delta_vertex"))
    m2 = dict(m, v_template=m["v_template"] + t("delta_vertex"), shapedirs=m["shapedirs"] + t("delta_shapedirs"),
              posedirs=m["posedirs"] + t("delta_posedirs"))
    ve, _, _ = fo.forward_with_delta_blendshape(m2, b1, t("pose"))
    assert torch.allclose(vd, ve, atol=1e-14)


def test_flame_root_rotation_is_a_rigid_motion_about_the_root_joint():
    f, m, t = _flame(seed=5)
    pose = t("pose").clone()
    v1, pf1, A1 = fo.forward_with_delta_blendshape(m, t("betas"), pose)
    pose0 = pose.clone()
    pose0[:3] = 0.0
    v0, pf0, _ = fo.forward_with_delta_blendshape(m, t("betas"), pose0)
    R = fo.batch_rodrigues(pose[None, :3])[0]
    J0 = (m["J_regressor"] @ (m["v_template"] + torch.einsum("l,mkl->mk", t("betas"), m["shapedirs"])))[0]
    assert torch.allclose(v1, (v0 - J0) @ R.T + J0, atol=1e-7)     # the whole head turns about joint 0
    assert torch.allclose(pf1, pf0, atol=1e-12)                     # pose correctives ignore the root rotation
    d = lambda v: torch.cdist(v[:20], v[:20])
    assert torch.allclose(d(v1), d(v0), atol=1e-7)                  # distances preserved


def test_pose_stage_is_equivariant_under_rigid_motion_and_scaling():
    p = scenes.pose_inputs(N=400, seed=6)
    t = lambda k, dt=torch.float64: torch.from_numpy(p[k]).to(dt) if p[k].dtype.kind == "f" else torch.from_numpy(p[k])
    faces, fi = t("faces"), t("face_index")
    _, canon = po.compute_face_orientation(t("canon_verts"), faces)
    raw = [t(k) for k in ("scaling_raw", "rotation_raw", "offset_raw", "opacity_raw")]
    base = po.pose_splats(t("verts"), faces, fi, t("bary"), canon, *raw, shell_len=0.05)
    R, tr, s = _rot(3), torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64), 1.7
    moved = po.pose_splats((s * t("verts")) @ R.T + tr, faces, fi, t("bary"), canon, *raw, shell_len=0.05)
    xyz0, sc0, q0, op0 = base
    xyz1, sc1, q1, op1 = moved
    # face normals are un-normalised (area-weighted): the shell offset scales with s^2, the surface point with s
    _, vn = po.compute_face_orientation(t("verts"), faces)
    pos0 = (t("bary")[..., None] * t("verts")[faces[fi]]).sum(-2)
    off0 = xyz0 - pos0
    assert torch.allclose(xyz1, (s * pos0 + s * s * off0) @ R.T + tr, atol=1e-9)
    assert torch.allclose(sc1, s * sc0, rtol=1e-9)                  # splats grow with the mesh
    assert torch.equal(op1, op0)
    assert torch.allclose(_quat_to_mat(q1), R @ _quat_to_mat(q0), atol=1e-9)   # orientation follows the mesh
    assert np.allclose(q1.norm(dim=-1).numpy(), 1.0) and (q1[:, 0] >= 0).all()


def test_pose_stage_rotation_is_face_frame_times_local_rotation():
    """The contract of the (unpinned) pytorch3d pair matrix_to_quaternion + quaternion_multiply as FateAvatar uses it:
    R(splat) = [a0 a1 a2](face) . R(normalised _rotation)."""
    p = scenes.pose_inputs(N=300, seed=7)
    t = lambda k: torch.from_numpy(p[k]).double() if p[k].dtype.kind == "f" else torch.from_numpy(p[k])
    faces, fi = t("faces"), t("face_index")
    orient, scale = po.compute_face_orientation(t("verts"), faces)
    _, canon = po.compute_face_orientation(t("canon_verts"), faces)
    xyz, sc, q, op = po.pose_splats(t("verts"), faces, fi, t("bary"), canon, t("scaling_raw"), t("rotation_raw"),
                                    t("offset_raw"), t("opacity_raw"), shell_len=0.05)
    local = _quat_to_mat(torch.nn.functional.normalize(t("rotation_raw"), dim=-1))
    assert torch.allclose(_quat_to_mat(q), orient[fi] @ local, atol=1e-9)
    assert torch.allclose(orient.transpose(-1, -2) @ orient, torch.eye(3, dtype=torch.float64).expand_as(orient), atol=1e-9)
    assert torch.allclose(torch.linalg.det(orient), torch.ones(orient.shape[0], dtype=torch.float64), atol=1e-9)
    assert torch.allclose(sc, torch.exp(t("scaling_raw")) * (scale / canon)[fi], rtol=1e-12)
