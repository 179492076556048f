#!/usr/bin/env python
"""Fused pose kernel vs the same computation as plain torch ops (the reference's formulation) on the GPU."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from fateavatar_b200 import pose, scenes
from oracle import pose_oracle as po
dev = torch.device("cuda:0")
p = scenes.pose_inputs(N=100000, seed=0)
d = lambda k, req=False: torch.from_numpy(p[k]).to(dev).requires_grad_(req)
faces, fi, bary = d("faces"), d("face_index"), d("bary")
_, canon = po.compute_face_orientation(d("canon_verts"), faces)
def run(fn):
    verts = d("verts", True); leaves = [d(k, True) for k in ("scaling_raw", "rotation_raw", "offset_raw", "opacity_raw")]
    out = fn(verts, leaves)
    loss = sum(o.sum() for o in out); loss.backward()
def fused(verts, leaves):
    return pose.pose_splats(verts, faces, fi, bary, canon, *leaves, shell_len=0.05)
def torch_ops(verts, leaves):
    return po.pose_splats(verts, faces, fi, bary, canon, *leaves, shell_len=0.05)
res = {}
for name, fn in (("fused", fused), ("torch_ops", torch_ops)):
    for _ in range(5): run(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): run(fn)
    e1.record(); torch.cuda.synchronize()
    res[name + "_fwd_bwd_us"] = round(1000 * e0.elapsed_time(e1) / 30, 1)
print(json.dumps(res))
