"""The reference's operator API (`diff_gaussian_rasterization`: GaussianRasterizationSettings, GaussianRasterizer)
implemented on the CPU ORACLE, with autograd.  TEST INFRASTRUCTURE ONLY.

It exists so the reference's *unchanged* volume_rendering/render_3dgs.py:render(..., device='cpu') can be executed in
this container (SURVEY 8d "CPU"): register this module as `diff_gaussian_rasterization` in sys.modules,
import the reference's render_3dgs.py, and its output is the reference-semantics result for that caller -- used to
check fateavatar_b200/render.py (row R0) against the reference's own code, and as the config-1 CPU baseline.
Mirrors DGR diff_gaussian_rasterization/__init__.py:21-220 (argument checks, gradient tuple order).
"""
from typing import NamedTuple

import numpy as np
import torch
import torch.nn as nn

from . import oracle as orc


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _np(t):
    return None if t is None or t.numel() == 0 else t.detach().cpu().float().numpy()


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        st = orc.forward(_np(means3D), _np(opacities), _np(rs.bg), _np(rs.viewmatrix), _np(rs.projmatrix), _np(rs.campos),
                         float(rs.tanfovx), float(rs.tanfovy), int(rs.image_height), int(rs.image_width), shs=_np(sh),
                         sh_degree=int(rs.sh_degree), scales=_np(scales), rotations=_np(rotations),
                         colors_precomp=_np(colors_precomp), cov3D_precomp=_np(cov3Ds_precomp),
                         scale_modifier=float(rs.scale_modifier))
        ctx.st = st
        ctx.have = tuple(t is not None and t.numel() > 0 for t in (sh, colors_precomp, scales, rotations, cov3Ds_precomp))
        radii = torch.from_numpy(st["radii"].astype(np.int32))
        ctx.mark_non_differentiable(radii)
        return torch.from_numpy(st["color"]), radii

    @staticmethod
    def backward(ctx, grad_color, _):
        g = orc.backward(ctx.st, grad_color.detach().cpu().float().numpy())
        t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k]))
        has_sh, has_cp, has_sc, has_ro, has_c3 = ctx.have
        return (t("dL_dmeans3D"), t("dL_dmeans2D"), t("dL_dsh") if has_sh else None, t("dL_dcolors") if has_cp else None,
                t("dL_dopacity"), t("dL_dscales") if has_sc else None, t("dL_drotations") if has_ro else None,
                t("dL_dcov3D") if has_c3 else None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        rs = self.raster_settings
        return torch.from_numpy(orc.mark_visible(_np(positions), _np(rs.viewmatrix), _np(rs.projmatrix)))

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        e = torch.Tensor([])
        return _Rasterize.apply(means3D, means2D, e if shs is None else shs, e if colors_precomp is None else colors_precomp,
                                opacities, e if scales is None else scales, e if rotations is None else rotations,
                                e if cov3D_precomp is None else cov3D_precomp, self.raster_settings)
