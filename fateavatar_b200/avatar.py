"""Caller-side mirror of `FateAvatar.forward` (model/fateavatar.py:196-298) on the fused operators.

The reference's per-frame forward is: build a Camera (two CPU 4x4 inverses and a host round trip per frame,
volume_rendering/camera_3dgs.py:53-72), FLAME skinning twice, ~60 torch ops that place the splats on the mesh, then
render().  `forward_frame(model, input)` produces the same output dict from the same model attributes with

    FrameCamera        closed-form view / projection / camera centre on the device (no inverse, no host sync)
    flame.flame_lbs    both FLAME meshes in one pass                                  (fs_flame_forward)
    pose.pose_splats   face frames, quaternions, barycentric positions, activations   (fs_pose_forward)
    GaussianRasterizer                                                               (fs_forward)

and `attach(model)` rebinds `model.forward` (plus the FLAME methods and the densification statistics) on an existing
reference FateAvatar instance, so train/iteration.py runs unchanged.  Everything stays differentiable through the
library's autograd Functions.  CUDA only.
"""
import math

import torch

from . import densify as _densify
from . import flame as _flame
from . import pose as _pose
from . import rasterizer as _rasterizer


class FrameCamera:
    """Attributes render() reads from volume_rendering/camera_3dgs.py:Camera, computed in closed form.

    The reference builds Rt = [[R^T, T], [0, 1]], inverts it twice on the CPU (getWorld2View2_torch with zero
    translate / unit scale is the identity on Rt) and inverts the view matrix once more for the camera centre.
    Here world_view_transform = Rt^T directly and camera_center = -R T, equal to the reference's up to the fp32
    rounding of its inverses."""

    _proj_cache = {}
    _const_cache = {}

    def __init__(self, R, T, FoVx, FoVy, img_res, znear=0.01, zfar=100.0, exact=False, cam_pose=None):
        """exact=True runs the reference's own sequence of library calls instead of the closed form -- Rt assembled on
        the host, torch.linalg.inv twice on the CPU (tools/gs_utils/graphics_utils.py:51-62), upload, bmm, and the GPU
        inverse for the camera centre (volume_rendering/camera_3dgs.py:53,71-72) -- so that view / projection matrices
        carry the same fp32 bits as upstream and `radii`, which ceil() a function of them, cannot flip.  It costs the
        host round trip the closed form exists to avoid and cannot be recorded into a CUDA graph."""
        R, T = R.reshape(3, 3), T.reshape(3)
        dev = R.device
        self.FoVx, self.FoVy = float(FoVx), float(FoVy)
        self.image_height, self.image_width = int(img_res[0]), int(img_res[1])
        self.znear, self.zfar = znear, zfar
        if exact:
            Rt = torch.zeros((4, 4))
            Rt[:3, :3] = R.transpose(0, 1)
            Rt[:3, 3] = T
            Rt[3, 3] = 1.0
            C2W = torch.linalg.inv(Rt)
            C2W[:3, 3] = (C2W[:3, 3] + torch.tensor([.0, .0, .0])) * 1.0
            view = torch.linalg.inv(C2W).float().transpose(0, 1).to(dev)
            proj = self._projection(dev)
            self.world_view_transform, self.projection_matrix = view, proj
            self.full_proj_transform = view.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
            self.camera_center = view.inverse()[3, :3]
            return
        if cam_pose is not None and cam_pose.is_cuda and cam_pose.dtype == torch.float32 and cam_pose.numel() == 16:
            # one kernel (fs_frame_camera) instead of ~8 small torch launches per frame
            from . import _lib

            cp = cam_pose.reshape(4, 4).contiguous()
            proj = self._projection(dev)
            buf = torch.empty(36, device=dev)
            with _lib.on_device(dev):
                rc = _lib.load().fs_frame_camera(cp.data_ptr(), proj.data_ptr(), buf.data_ptr(), buf.data_ptr() + 64,
                                                 buf.data_ptr() + 128, _lib.stream_ptr(dev))
            _lib.check(rc, "fs_frame_camera")
            self.world_view_transform, self.full_proj_transform = buf[:16].view(4, 4), buf[16:32].view(4, 4)
            self.projection_matrix, self.camera_center = proj, buf[32:35]
            return
        consts = FrameCamera._const_cache.get(str(dev))
        if consts is None:  # built once per device, outside any CUDA-graph capture (warm-up frames come first)
            consts = (torch.zeros(3, 1, device=dev), torch.ones(1, 1, device=dev))
            FrameCamera._const_cache[str(dev)] = consts
        R32, T32 = R.to(torch.float32), T.to(torch.float32)
        # world_view_transform = Rt^T = [[R, 0], [T, 1]]; assembled from device tensors only (capture-safe)
        view = torch.cat([torch.cat([R32, consts[0]], dim=1), torch.cat([T32[None], consts[1]], dim=1)], dim=0)
        self.world_view_transform = view
        proj = self._projection(dev)
        self.projection_matrix = proj
        self.full_proj_transform = view @ proj
        self.camera_center = -(R32 @ T32)

    def _projection(self, dev):
        """getProjectionMatrix (tools/gs_utils/graphics_utils.py:64-84) in the reference's own arithmetic, transposed;
        cached per (fov, clip planes, device)."""
        znear, zfar = self.znear, self.zfar
        key = (self.FoVx, self.FoVy, znear, zfar, str(dev))
        proj = FrameCamera._proj_cache.get(key)
        if proj is None:
            top, right = math.tan(self.FoVy / 2) * znear, math.tan(self.FoVx / 2) * znear
            bottom, left = -top, -right
            P = torch.zeros(4, 4)
            P[0, 0] = 2.0 * znear / (right - left)
            P[1, 1] = 2.0 * znear / (top - bottom)
            P[0, 2] = (right + left) / (right - left)
            P[1, 2] = (top + bottom) / (top - bottom)
            P[3, 2] = 1.0
            P[2, 2] = 1.0 * zfar / (zfar - znear)
            P[2, 3] = -(zfar * znear) / (zfar - znear)
            proj = P.transpose(0, 1).contiguous().to(dev)
            FrameCamera._proj_cache[key] = proj
        return proj


def quaternion_to_axis_angle(q):
    """pytorch3d.transforms.quaternion_to_axis_angle (published 0.7.x algorithm; the package is not vendored by the
    reference, so this restatement is parity-unpinned like the other quaternion helpers)."""
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    angles = 2 * half
    small = angles.abs() < 1e-6
    s = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half) / torch.where(small, torch.ones_like(angles), angles))
    return q[..., 1:] / s


_zeros = {}


def _cached_zeros(shape, device, dtype):
    """Read-only zeros, one allocation per (shape, device, dtype)."""
    key = (shape, str(device), dtype)
    z = _zeros.get(key)
    if z is None:
        z = _zeros[key] = torch.zeros(shape, device=device, dtype=dtype)
    return z


def forward_frame(model, input, exact_camera=None, extras=True):
    """`exact_camera` (default: `model.exact_camera` if set, else False) selects FrameCamera(exact=True).
    `extras=False` leaves out the two outputs that only the scale / rotation regularisers read ("scale" = exp(_scaling),
    "raw_rot" = quaternion_to_axis_angle(_rotation): ~15 elementwise launches per frame) for callers whose loss does not
    use them, `extras=("scale",)` computes only the named ones; everything else is computed regardless.
    `model`: an object with FateAvatar's attributes (flame, faces, face_index, bary_coords, face_scaling_canonical,
    _scaling, _rotation, _offset, _opacity, _features_dc, delta_shapedirs, delta_posedirs, delta_vertex, cfg_model,
    shell_len, bg_color, img_res, device); `input`: the dataset's dict (cam_pose, fovx, fovy, flame_pose, expression).
    Returns the dict of model/fateavatar.py:280-296."""
    cam_pose = input["cam_pose"]
    if exact_camera is None:
        exact_camera = bool(getattr(model, "exact_camera", False))
    camera = FrameCamera(cam_pose[:, :3, :3], cam_pose[:, :3, 3], input["fovx"][0], input["fovy"][0], model.img_res,
                         exact=exact_camera, cam_pose=None if exact_camera else cam_pose)
    flame_pose, expression = input["flame_pose"], input["expression"]
    bs = flame_pose.shape[0]
    if bs != 1:
        raise _rasterizer.FateSplatError("forward_frame renders one frame per call (the reference's batch size)")
    cfg = model.cfg_model
    fm = getattr(model, "_fs_flame_model", None)
    if fm is None:
        fm = model._fs_flame_model = _flame.model_tensors(model.flame)
    n_shape, n_exp = int(model.flame.n_shape), int(model.flame.n_exp)
    e = expression[:, :n_exp]
    betas = torch.cat([_cached_zeros((1, n_shape), e.device, e.dtype), e], dim=1)  # flame/FLAME.py:180
    verts, _, _, verts_orig, _ = _flame.flame_lbs(
        fm, betas, flame_pose,
        model.delta_shapedirs if cfg.delta_blendshape else None, model.delta_posedirs if cfg.delta_blendshape else None,
        model.delta_vertex if cfg.delta_vertex else None, l0=n_shape, want_orig=True)
    xyz, scales, rots, opac = _pose.pose_splats(verts, model.faces, model.face_index, model.bary_coords,
                                                model.face_scaling_canonical, model._scaling, model._rotation,
                                                model._offset, model._opacity, shell_len=model.shell_len,
                                                resize_scale=bool(cfg.resize_scale))
    # render_3dgs.py:22-27: a zero tensor whose .grad feeds the densifier.  No kernel reads or writes its values, so
    # every frame gets a fresh leaf over one shared block of zeros instead of a fill kernel
    screenspace = _cached_zeros(tuple(xyz.shape), xyz.device, xyz.dtype).detach().requires_grad_(True)
    settings = _rasterizer.GaussianRasterizationSettings(
        image_height=camera.image_height, image_width=camera.image_width, tanfovx=math.tan(camera.FoVx * 0.5),
        tanfovy=math.tan(camera.FoVy * 0.5), bg=model.bg_color.to(xyz.device), scale_modifier=1.0,
        viewmatrix=camera.world_view_transform, projmatrix=camera.full_proj_transform, sh_degree=0,
        campos=camera.camera_center, prefiltered=False, debug=False)
    image, radii = _rasterizer.GaussianRasterizer(settings)(means3D=xyz, means2D=screenspace, shs=model._features_dc,
                                                           opacities=opac, scales=scales, rotations=rots)
    want = ("scale", "raw_rot") if extras is True else (tuple(extras) if extras else ())
    out = {}
    if "scale" in want:
        out["scale"] = torch.exp(model._scaling)
    if "raw_rot" in want:
        out["raw_rot"] = quaternion_to_axis_angle(model._rotation)
    out.update({
        "rgb_image": image[None],
        "viewspace_points": [screenspace],
        "visibility_filter": [radii > 0],
        "radii": [radii],
        "bs": bs,
        "verts": verts,
        "verts_orig": verts_orig,
        "faces": model.faces,
    })
    return out


def attach(model, exact_camera=False):
    """Rebind the per-frame hot path of an existing reference FateAvatar instance: `forward`, the two FLAME methods and
    `_add_densification_stats`.  Nothing else of the model (densify / prune / checkpoints / inference) is touched."""
    _flame.attach(model.flame)
    _densify.attach(model)
    model.forward = lambda input: forward_frame(model, input, exact_camera=exact_camera)
    return model
