"""Which rounding sequence does torch.norm(x[:, :2], dim=-1) use on the GPU?  (densification statistics, stats.cu)"""
import sys
import torch
sys.path.insert(0, '.')
from fateavatar_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
P = 1 << 20
grad = torch.randn(P, 3, device=dev, generator=g) * torch.exp(torch.randn(P, 1, device=dev, generator=g) * 3)
radii = torch.ones(P, dtype=torch.int32, device=dev)
a, d = torch.empty(P, 1, device=dev), torch.empty(P, 1, device=dev)
with _lib.on_device(dev):
    lib.fs_densify_stats_inc(P, grad.data_ptr(), radii.data_ptr(), a.data_ptr(), d.data_ptr(), _lib.stream_ptr(dev))
ref = torch.norm(grad[:, :2], dim=-1, keepdim=True)
x0, x1 = grad[:, 0:1], grad[:, 1:2]
cands = {
    "kernel": a,
    "sqrt(x0*x0 + x1*x1) separately rounded": torch.sqrt(x0 * x0 + x1 * x1),
    "float64 then round": (x0.double() ** 2 + x1.double() ** 2).sqrt().float(),
    "scaled: m*sqrt((x0/m)^2+(x1/m)^2)": (lambda m: m * torch.sqrt((x0 / m) ** 2 + (x1 / m) ** 2))(torch.maximum(x0.abs(), x1.abs())),
    "linalg.vector_norm": torch.linalg.vector_norm(grad[:, :2], dim=-1, keepdim=True),
    "hypot": torch.hypot(x0, x1),
}
torch.cuda.synchronize()
for k, v in cands.items():
    print(f"{k:45s} mismatches {int((v != ref).sum()):8d} of {P}")
