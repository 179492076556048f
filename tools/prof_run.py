#!/usr/bin/env python
"""Tiny driver for ncu: N forward+backward frames of one scene through the C-ABI path (and the reference)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fateavatar_b200 import rasterizer as R, scenes
from oracle import ref_loader

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
impl = sys.argv[3] if len(sys.argv) > 3 else "new"
sc = {"c1": scenes.config1_scene, "c2": scenes.head_scene, "c5": scenes.stress_scene,
      "c2big": lambda: scenes.head_scene(scale_mult=4.0)}[which]()
dev = torch.device("cuda:0")
t = scenes.to_torch(sc, dev)
cam = t["camera"]
rs = R.GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], t["bg"], 1.0, cam["viewmatrix"],
                                     cam["projmatrix"], sc["sh_degree"], cam["campos"], False, False)
dpix = torch.from_numpy(np.random.default_rng(7).standard_normal((3, cam["H"], cam["W"])).astype(np.float32)).to(dev)
if os.environ.get("FATESPLAT_ASYNC") == "1":
    R.set_async(True)
for _ in range(iters):
    if impl == "new":
        c, r, s = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
        R.backward_raw(s, dpix)
    else:
        st = ref_loader.ref_forward(t, cam, sh_degree=sc["sh_degree"])
        ref_loader.ref_backward(st, dpix)
torch.cuda.synchronize()
print("done", which, impl)
