"""Oracle of the per-splat pose stage.  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch restatement of what FateAvatar does between the FLAME vertices and the rasterizer call:

  model/fateavatar.py:225-233   face frames / scale / normals gathered per splat, matrix_to_quaternion
  model/fateavatar.py:235-240   barycentric splat positions (volume_rendering/mesh_sampling.py:171-200)
  model/fateavatar.py:253-258   _scaling + log(ratio), quaternion_multiply(q_face, _rotation),
                                pos + n_face * shell_len * tanh(_offset)
  volume_rendering/gaussian_model.py:105-128  activations: exp / F.normalize / sigmoid
  volume_rendering/mesh_compute.py:18-59      safe_normalize (eps 1e-20), compute_face_orientation, compute_face_normals

Gradients come from autograd of this restatement (run it in float64 for a clean reference).

Third-party arithmetic: matrix_to_quaternion / quaternion_multiply are pytorch3d 0.7.7 functions (README.md:40);
pytorch3d is not vendored under /root/reference and not installed here, so their published algorithm is
restated below -- PARITY UNPINNED at that boundary (no reference test or fixture covers it).  The final
rotation is invariant to the sign convention of the intermediate face quaternion because
quaternion_multiply standardises the product.  The mesh_compute functions themselves ARE checked against the
reference's own file when /root/reference is mounted (tests/test_pose.py).
"""
import torch


def safe_normalize(x, eps=1e-20):  # mesh_compute.py:18-22
    return x / torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=eps))


def length(x, eps=1e-20):
    return torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=eps))


def compute_face_orientation(verts, faces):  # mesh_compute.py:38-59 (return_scale=True)
    v0, v1, v2 = verts[..., faces[:, 0], :], verts[..., faces[:, 1], :], verts[..., faces[:, 2], :]
    a0 = safe_normalize(v1 - v0)
    a1 = safe_normalize(torch.cross(a0, v2 - v0, dim=-1))
    a2 = -safe_normalize(torch.cross(a1, a0, dim=-1))
    orientation = torch.cat([a0[..., None], a1[..., None], a2[..., None]], dim=-1)
    s0 = length(v1 - v0)
    s1 = (a2 * (v2 - v0)).sum(-1, keepdim=True).abs()
    return orientation, (s0 + s1) / 2


def compute_face_normals(verts, faces):  # mesh_compute.py:27-36 (NOT normalised)
    v0, v1, v2 = verts[..., faces[:, 0], :], verts[..., faces[:, 1], :], verts[..., faces[:, 2], :]
    return torch.cross(v1 - v0, v2 - v0, dim=-1)


def _sqrt_positive_part(x):  # pytorch3d.transforms.rotation_conversions
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def standardize_quaternion(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def matrix_to_quaternion(matrix):  # pytorch3d 0.7.x algorithm
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(matrix.shape[:-2] + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack(
        [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    flr = torch.tensor(0.1, dtype=q_abs.dtype)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    sel = torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return standardize_quaternion(cand[sel, :].reshape(matrix.shape[:-2] + (4,)))


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def quaternion_multiply(a, b):
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def pose_splats(verts, faces, face_index, bary, face_scaling_canonical, scaling_raw, rotation_raw, offset_raw,
                opacity_raw, shell_len=0.05, resize_scale=True):
    """verts [V,3], faces [F,3] long, face_index [N] long, bary [N,3], face_scaling_canonical [F,1] (or [F]).
    Returns the four activated tensors render() hands to the rasterizer: means3D [N,3], scales [N,3],
    rotations [N,4] (unit), opacities [N,1]."""
    orient, fscale = compute_face_orientation(verts, faces)
    normals = compute_face_normals(verts, faces)
    ratio = fscale / face_scaling_canonical.reshape(-1, 1)
    r_n = ratio[face_index]
    q_face = matrix_to_quaternion(orient[face_index])
    n_n = normals[face_index]
    fv = verts[faces[face_index]]  # [N,3(vertex),3]
    pos = (bary[..., None] * fv).sum(dim=-2)
    log_s = scaling_raw + torch.log(r_n) if resize_scale else scaling_raw
    rot = quaternion_multiply(q_face, rotation_raw)
    xyz = pos + n_n * shell_len * torch.tanh(offset_raw)
    return xyz, torch.exp(log_s), torch.nn.functional.normalize(rot), torch.sigmoid(opacity_raw)
