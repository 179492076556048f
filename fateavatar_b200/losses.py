"""Fused image loss of the optimise loop (training-loop plumbing).

`l1_image_loss(image, target)` = train/loss.py:103-105 with rgb_type 'l1' (torch.nn.L1Loss, mean reduction): one kernel
writes the loss and its gradient image, where the torch formulation launches ~9 elementwise / reduction kernels over the
3 MB image per frame.  Deterministic (fixed-order partial sums).  CUDA only."""
import torch

from . import _lib
from ._lib import FateSplatError

_ws = {}


class _L1Image(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, target):
        if not image.is_cuda:
            raise FateSplatError("l1_image_loss needs CUDA tensors: fateavatar_b200 has no CPU path")
        lib = _lib.load()
        x, t = image.detach().contiguous().float(), target.detach().contiguous().float()
        if x.shape != t.shape:
            raise FateSplatError(f"l1_image_loss: shapes differ ({tuple(x.shape)} vs {tuple(t.shape)})")
        dev = x.device
        ws = _ws.get(dev)
        if ws is None:
            ws = _ws[dev] = torch.zeros(lib.fs_l1_loss_workspace_bytes(), dtype=torch.uint8, device=dev)
        grad, loss = torch.empty_like(x), torch.empty((), device=dev)
        with _lib.on_device(dev):
            rc = lib.fs_l1_loss(x.numel(), x.data_ptr(), t.data_ptr(), grad.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                _lib.stream_ptr(dev))
        _lib.check(rc, "fs_l1_loss")
        ctx.save_for_backward(grad)
        ctx.shape = image.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).view(ctx.shape), None


def l1_image_loss(image, target):
    """mean(|image - target|) with gradient to `image` (the target is data)."""
    return _L1Image.apply(image, target)
