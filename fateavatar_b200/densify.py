"""Densification statistics on one fused kernel.

Mirrors `FateAvatar._add_densification_stats` (model/fateavatar.py:734-737) and
`GaussianModel.add_densification_stats` (volume_rendering/gaussian_model.py:418-420):

    self.xyz_gradient_accum[update_filter] += torch.norm(viewspace_point_tensor.grad[update_filter, :2], dim=-1, keepdim=True)
    self.denom[update_filter] += 1

`add_densification_stats(model, viewspace_point_tensor, update_filter)` has the reference method's arguments with
the model passed explicitly; `attach(model)` rebinds the method on an existing instance.  CUDA only.
"""
import torch

from . import _lib
from ._lib import FateSplatError


def densify_stats_raw(xyz_gradient_accum, denom, viewspace_grad, update_filter):
    """In place on contiguous CUDA tensors: accum [P,1] f32, denom [P,1] f32, viewspace_grad [P,3] f32,
    update_filter [P] bool/uint8."""
    if not xyz_gradient_accum.is_cuda:
        raise FateSplatError("densify_stats needs CUDA tensors: fateavatar_b200 has no CPU path")
    P = viewspace_grad.shape[0]
    if update_filter.dtype == torch.bool:
        update_filter = update_filter.view(torch.uint8)
    for t, n in ((xyz_gradient_accum, "xyz_gradient_accum"), (denom, "denom"), (viewspace_grad, "viewspace grad"),
                 (update_filter, "update_filter")):
        if not t.is_contiguous():
            raise FateSplatError(f"{n} must be contiguous")
    if xyz_gradient_accum.numel() != P or denom.numel() != P or update_filter.numel() != P or viewspace_grad.shape[1] != 3:
        raise FateSplatError("densify_stats: shape mismatch")
    dev = viewspace_grad.device
    with _lib.on_device(dev):
        rc = _lib.load().fs_densify_stats(P, viewspace_grad.data_ptr(), update_filter.data_ptr(),
                                          xyz_gradient_accum.data_ptr(), denom.data_ptr(),
                                          _lib.stream_ptr(dev))
    _lib.check(rc, "fs_densify_stats")


def add_densification_stats(model, viewspace_point_tensor, update_filter):
    densify_stats_raw(model.xyz_gradient_accum, model.denom, viewspace_point_tensor.grad, update_filter)


def attach(model):
    """Rebind `_add_densification_stats` (FateAvatar and the baselines) / `add_densification_stats` (GaussianModel)."""
    fn = lambda viewspace_point_tensor, update_filter: add_densification_stats(model, viewspace_point_tensor, update_filter)
    for name in ("_add_densification_stats", "add_densification_stats"):
        if hasattr(model, name):
            setattr(model, name, fn)
    return model
