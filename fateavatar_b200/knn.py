"""Drop-in for `simple_knn._C.distCUDA2` (simple-knn/ext.cpp:15-17, spatial.cu:15-26), backed by
libfatesplat.so's fs_knn_mean_dist2.  Same contract: float32 CUDA [P,3] in, float32 [P] out = mean of the
three smallest squared neighbour distances.  No CPU path."""
import torch

from . import _lib
from ._lib import FateSplatError


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise FateSplatError("distCUDA2 needs a CUDA tensor: fateavatar_b200 has no CPU path")
    if points.dtype != torch.float32:
        raise TypeError(f"points must be float32 (got {points.dtype})")
    lib = _lib.load()
    pts = points.contiguous()
    P = pts.shape[0]
    means = torch.zeros((P,), dtype=torch.float32, device=pts.device)  # reference: torch::full({P}, 0.0)
    if P == 0:
        return means
    nbytes = lib.fs_knn_workspace_bytes(P)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=pts.device)
    with _lib.on_device(pts.device):
        rc = lib.fs_knn_mean_dist2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(), nbytes,
                                   _lib.stream_ptr(pts.device))
        _lib.check(rc, "fs_knn_mean_dist2")
    return means
