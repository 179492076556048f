#!/usr/bin/env python
"""Per-stage device times (CUDA events inside libfatesplat) for one scene; knobs come from the environment."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fateavatar_b200 import rasterizer as R, scenes, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sc = {"c1": scenes.config1_scene, "c2": scenes.head_scene, "c5": scenes.stress_scene,
      "c2big": lambda: scenes.head_scene(scale_mult=4.0)}[which]()
dev = torch.device("cuda:0")
t = scenes.to_torch(sc, dev)
cam = t["camera"]
rs = R.GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], t["bg"], 1.0, cam["viewmatrix"],
                                     cam["projmatrix"], sc["sh_degree"], cam["campos"], False, False)
dpix = torch.from_numpy(np.random.default_rng(7).standard_normal((3, cam["H"], cam["W"])).astype(np.float32)).to(dev)
R.set_async(True)
def step():
    c, r, s = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
    R.backward_raw(s, dpix)
    return s
for _ in range(5):
    s = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / iters
_lib.load().fs_profile_enable(1)
_lib.profile_read()
for _ in range(iters):
    step()
prof = _lib.profile_read()
_lib.load().fs_profile_enable(0)
st = {k: round(1000 * v[0] / max(v[1], 1), 1) for k, v in prof.items() if v[1]}
knobs = {k: v for k, v in os.environ.items() if k.startswith("FATESPLAT_")}
print(json.dumps(dict(scene=which, knobs=knobs, step_us=round(1000 * total, 1), stages_us=st, sum_us=round(sum(st.values()), 1))))
