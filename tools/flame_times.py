#!/usr/bin/env python
"""FLAME stage on the GPU: fused kernels (raw C-ABI calls and through autograd) vs the reference's formulation as
plain torch ops (two lbs() calls per frame, as model/fateavatar.py:211-222 does)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from fateavatar_b200 import _lib, flame, scenes
from oracle import flame_oracle as fo
dev = torch.device("cuda:0")
f = scenes.flame_inputs(seed=0)
V, L, n_shape = f["v_template"].shape[0], f["shapedirs"].shape[-1], f["n_shape"]
d = lambda k: torch.from_numpy(f[k]).to(dev)
m = {k: d(k) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
mt = dict(m); mt["parents"] = torch.from_numpy(f["parents"]).to(dev)
m["parents"] = [int(x) for x in f["parents"]]
betas, pose = d("betas"), d("pose")
g = torch.randn(V, 3, device=dev)
deltas = {k: d(k).requires_grad_(True) for k in ("delta_vertex", "delta_shapedirs", "delta_posedirs")}

def fused_autograd():
    for t in deltas.values(): t.grad = None
    v, pf, A, vo, Ao = flame.flame_lbs(m, betas, pose, deltas["delta_shapedirs"], deltas["delta_posedirs"], deltas["delta_vertex"], l0=n_shape)
    (v[0] * g).sum().backward()

ws = torch.empty(_lib.load().fs_flame_workspace_bytes(V), dtype=torch.uint8, device=dev)
outg = [torch.empty(V, 3, device=dev), torch.empty(V, 3, L, device=dev), torch.empty(36, 3 * V, device=dev)]
def fused_raw():
    r = flame.flame_forward_raw(betas, pose, m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"],
                                m["lbs_weights"], deltas["delta_vertex"].detach(), deltas["delta_shapedirs"].detach(),
                                deltas["delta_posedirs"].detach(), l0=n_shape, workspace=ws)
    flame.flame_backward_raw(betas, m["J_regressor"], m["parents"], m["lbs_weights"], ws, g, (V, L), l0=n_shape, out=outg)

def torch_ops():
    for t in deltas.values(): t.grad = None
    v, _, _ = fo.forward_with_delta_blendshape(mt, betas, pose, deltas["delta_shapedirs"], deltas["delta_posedirs"], deltas["delta_vertex"])
    vo, _, _ = fo.forward_with_delta_blendshape(mt, betas, pose)
    (v * g).sum().backward()

# the oracle builds small tensors on the CPU default device; run it under a cuda default device
res = {}
for name, fn in (("fused_raw", fused_raw), ("fused_autograd", fused_autograd), ("torch_ops", torch_ops)):
    ctx = torch.device(dev) if name == "torch_ops" else torch.device("cpu")
    with ctx:
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): fn()
        e1.record(); torch.cuda.synchronize()
    res[name + "_fwd_bwd_us"] = round(1000 * e0.elapsed_time(e1) / 30, 1)
lib = _lib.load(); lib.fs_profile_enable(1); _lib.profile_read()
for _ in range(20): fused_raw()
torch.cuda.synchronize()
pr = _lib.profile_read(); lib.fs_profile_enable(0)
res["stage_us"] = {k: round(1000 * v[0] / v[1], 2) for k, v in pr.items() if v[1]}
print(json.dumps(res))
