"""Host-side mirror of the reference render boundary, volume_rendering/render_3dgs.py:7-81.

`render(viewpoint_camera, pc, bg_color, scaling_modifier, override_color, device)` has the reference's
signature, argument meaning and return dict; it exists so tests/bench on a box without /root/reference can
drive the operators exactly the way FateAvatar does.  With fateavatar_b200.install() the reference's own
render_3dgs.py runs unchanged instead (tests/test_dropin_reference_gpu.py does that on the B200 from the staged
copy under oracle/_ref/pyref; tests/test_render_caller_cpu.py on the CPU oracle).

`rasterizer_module` lets the tests substitute another implementation of the same operator API (the CPU
oracle drop-in or the compiled reference) without touching this function.
"""
import math

import torch

from . import rasterizer as _default_rasterizer


class MiniCam:
    """Camera record with the attribute names render() reads (volume_rendering/camera_3dgs.py:76-87)."""

    def __init__(self, width, height, fovy, fovx, world_view_transform, full_proj_transform, camera_center=None,
                 znear=0.01, zfar=100.0):
        self.image_width = width
        self.image_height = height
        self.FoVy = fovy
        self.FoVx = fovx
        self.znear = znear
        self.zfar = zfar
        self.world_view_transform = world_view_transform
        self.full_proj_transform = full_proj_transform
        self.camera_center = camera_center if camera_center is not None else \
            torch.inverse(world_view_transform)[3][:3]


class SplatCloud:
    """Minimal stand-in for GaussianModel's getter surface (volume_rendering/gaussian_model.py:105-128):
    raw parameters in, activated properties out (exp / normalize / sigmoid)."""

    def __init__(self, xyz, features, scaling, rotation, opacity, max_sh_degree=0):
        self._xyz, self._features, self._scaling, self._rotation, self._opacity = xyz, features, scaling, rotation, opacity
        self.max_sh_degree = max_sh_degree

    get_xyz = property(lambda self: self._xyz)
    get_features = property(lambda self: self._features)
    get_scaling = property(lambda self: torch.exp(self._scaling))
    get_rotation = property(lambda self: torch.nn.functional.normalize(self._rotation))
    get_opacity = property(lambda self: torch.sigmoid(self._opacity))


def render(viewpoint_camera, pc, bg_color, scaling_modifier=1.0, override_color=None, device='cuda',
           rasterizer_module=None):
    mod = rasterizer_module or _default_rasterizer
    means3D = pc.get_xyz
    # zero tensor whose .grad receives the screen-space mean gradients (render_3dgs.py:22-27)
    screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device=device) + 0
    if screenspace_points.requires_grad:
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = mod.GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform, projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=pc.max_sh_degree, campos=viewpoint_camera.camera_center, prefiltered=False, debug=False)
    rasterizer = mod.GaussianRasterizer(raster_settings=raster_settings)
    shs = pc.get_features
    colors_precomp = None
    if override_color is not None:
        colors_precomp, shs = override_color, None
    rendered_image, radii = rasterizer(
        means3D=means3D, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
        opacities=pc.get_opacity, scales=pc.get_scaling, rotations=pc.get_rotation, cov3D_precomp=None)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}
