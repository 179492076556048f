"""Fused per-splat pose stage behind a torch autograd Function.

`pose_splats(...)` computes, in ONE kernel per direction, exactly what FateAvatar.forward does between the
FLAME vertices and render() (model/fateavatar.py:225-240, 253-258) plus the GaussianModel activations that
render() applies (volume_rendering/gaussian_model.py:105-128):

    face_orien_mat, face_scaling = compute_face_orientation(verts, faces, return_scale=True)
    face_normals                 = compute_face_normals(verts, faces)
    ratio    = (face_scaling / face_scaling_canonical)[face_index]
    q_face   = matrix_to_quaternion(face_orien_mat[face_index])
    pos      = reweight_verts_by_barycoords(verts, faces, face_index, bary_coords)
    scales   = exp(_scaling + log(ratio));  rots = normalize(quaternion_multiply(q_face, _rotation))
    xyz      = pos + face_normals[face_index] * shell_len * tanh(_offset);  opac = sigmoid(_opacity)

The four results are what GaussianRasterizer consumes.  Gradients flow to verts (and on to the FLAME
blendshape deltas through torch's autograd of the LBS) and to the four per-splat parameters.  See
INTEGRATION.md for the 12-line patch that calls it from model/fateavatar.py.  CUDA only; no CPU path.
"""
import torch

from . import _lib
from ._lib import FateSplatError


def pose_forward_raw(verts, faces, face_index, bary, canon, scaling_raw, rotation_raw, offset_raw, opacity_raw,
                     shell_len=0.05, resize_scale=True, out=None):
    """One fs_pose_forward call on contiguous CUDA tensors (no autograd).  Returns (xyz, scales, rots, opac)."""
    lib = _lib.load()
    dev = verts.device
    N, V, F = face_index.shape[0], verts.shape[-2], faces.shape[0]
    xyz, scales, rots, opac = out if out is not None else (
        torch.empty((N, 3), device=dev), torch.empty((N, 3), device=dev), torch.empty((N, 4), device=dev),
        torch.empty((N, 1), device=dev))
    rc = lib.fs_pose_forward(N, V, F, verts.data_ptr(), faces.data_ptr(), face_index.data_ptr(), bary.data_ptr(),
                             canon.data_ptr(), scaling_raw.data_ptr(), rotation_raw.data_ptr(), offset_raw.data_ptr(),
                             opacity_raw.data_ptr(), float(shell_len), int(bool(resize_scale)), xyz.data_ptr(),
                             scales.data_ptr(), rots.data_ptr(), opac.data_ptr(),
                             _lib.stream_ptr(dev))
    _lib.check(rc, "fs_pose_forward")
    return xyz, scales, rots, opac


def pose_backward_raw(verts, faces, face_index, bary, canon, scaling_raw, rotation_raw, offset_raw, opacity_raw,
                      g_xyz, g_scales, g_rots, g_opac, shell_len=0.05, resize_scale=True, out=None):
    """One fs_pose_backward call.  Returns (d_verts, d_scaling, d_rotation, d_offset, d_opacity)."""
    lib = _lib.load()
    dev = verts.device
    N, V, F = face_index.shape[0], verts.shape[-2], faces.shape[0]
    d_verts, d_sr, d_rr, d_of, d_op = out if out is not None else (
        torch.empty((V, 3), device=dev), torch.empty((N, 3), device=dev), torch.empty((N, 4), device=dev),
        torch.empty((N, 1), device=dev), torch.empty((N, 1), device=dev))
    rc = lib.fs_pose_backward(N, V, F, verts.data_ptr(), faces.data_ptr(), face_index.data_ptr(), bary.data_ptr(),
                              canon.data_ptr(), scaling_raw.data_ptr(), rotation_raw.data_ptr(), offset_raw.data_ptr(),
                              opacity_raw.data_ptr(), float(shell_len), int(bool(resize_scale)), g_xyz.data_ptr(),
                              g_scales.data_ptr(), g_rots.data_ptr(), g_opac.data_ptr(), d_verts.data_ptr(),
                              d_sr.data_ptr(), d_rr.data_ptr(), d_of.data_ptr(), d_op.data_ptr(),
                              _lib.stream_ptr(dev))
    _lib.check(rc, "fs_pose_backward")
    return d_verts, d_sr, d_rr, d_of, d_op


class _PoseSplats(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, scaling_raw, rotation_raw, offset_raw, opacity_raw, faces, face_index, bary, canon,
                shell_len, resize_scale):
        if not verts.is_cuda:
            raise FateSplatError("pose_splats needs CUDA tensors: fateavatar_b200 has no CPU path")
        lib = _lib.load()
        dev = verts.device
        v = verts.reshape(-1, 3).contiguous().float()
        sr, rr = scaling_raw.contiguous().float(), rotation_raw.contiguous().float()
        of, opr = offset_raw.contiguous().float(), opacity_raw.contiguous().float()
        fc, fi = faces.contiguous().long(), face_index.contiguous().long()
        bc, cn = bary.contiguous().float(), canon.reshape(-1).contiguous().float()
        N, V, F = fi.shape[0], v.shape[0], fc.shape[0]
        xyz = torch.empty((N, 3), device=dev)
        scales = torch.empty((N, 3), device=dev)
        rots = torch.empty((N, 4), device=dev)
        opac = torch.empty((N, 1), device=dev)
        with _lib.on_device(dev):
            rc = lib.fs_pose_forward(N, V, F, v.data_ptr(), fc.data_ptr(), fi.data_ptr(), bc.data_ptr(), cn.data_ptr(),
                                     sr.data_ptr(), rr.data_ptr(), of.data_ptr(), opr.data_ptr(), float(shell_len),
                                     int(bool(resize_scale)), xyz.data_ptr(), scales.data_ptr(), rots.data_ptr(),
                                     opac.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "fs_pose_forward")
        ctx.save_for_backward(v, sr, rr, of, opr, fc, fi, bc, cn)
        ctx.meta = (float(shell_len), int(bool(resize_scale)), verts.shape)
        return xyz, scales, rots, opac

    @staticmethod
    def backward(ctx, g_xyz, g_scales, g_rots, g_opac):
        lib = _lib.load()
        v, sr, rr, of, opr, fc, fi, bc, cn = ctx.saved_tensors
        shell_len, resize, vshape = ctx.meta
        dev = v.device
        N, V, F = fi.shape[0], v.shape[0], fc.shape[0]
        z = lambda g, shape: (torch.zeros(shape, device=dev) if g is None else g.contiguous().float())
        g_xyz, g_scales, g_rots, g_opac = z(g_xyz, (N, 3)), z(g_scales, (N, 3)), z(g_rots, (N, 4)), z(g_opac, (N, 1))
        d_verts = torch.empty((V, 3), device=dev)  # zero-filled by the library
        d_sr, d_rr = torch.empty_like(sr), torch.empty_like(rr)
        d_of, d_op = torch.empty_like(of), torch.empty_like(opr)
        with _lib.on_device(dev):
            rc = lib.fs_pose_backward(N, V, F, v.data_ptr(), fc.data_ptr(), fi.data_ptr(), bc.data_ptr(), cn.data_ptr(),
                                      sr.data_ptr(), rr.data_ptr(), of.data_ptr(), opr.data_ptr(), shell_len, resize,
                                      g_xyz.data_ptr(), g_scales.data_ptr(), g_rots.data_ptr(), g_opac.data_ptr(),
                                      d_verts.data_ptr(), d_sr.data_ptr(), d_rr.data_ptr(), d_of.data_ptr(),
                                      d_op.data_ptr(), _lib.stream_ptr(dev))
        _lib.check(rc, "fs_pose_backward")
        return d_verts.reshape(vshape), d_sr, d_rr, d_of, d_op, None, None, None, None, None, None


def pose_splats(verts, faces, face_index, bary_coords, face_scaling_canonical, scaling_raw, rotation_raw, offset_raw,
                opacity_raw, shell_len=0.05, resize_scale=True):
    """verts [V,3] or [1,V,3]; faces [F,3] long; face_index [N] long; bary_coords [N,3];
    face_scaling_canonical [F,1]; raw per-splat parameters as FateAvatar stores them.
    Returns (means3D [N,3], scales [N,3], rotations [N,4], opacities [N,1])."""
    return _PoseSplats.apply(verts, scaling_raw, rotation_raw, offset_raw, opacity_raw, faces, face_index,
                             bary_coords, face_scaling_canonical, shell_len, resize_scale)
