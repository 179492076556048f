#!/usr/bin/env python
"""Where the host time of one end-to-end frame goes (operator API, host buffers): cProfile of bench.py's e2e step."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from fateavatar_b200 import flame, pose, rasterizer as R, scenes

class A: P = 100000; res = 512
args = A()
dev = torch.device("cuda:0")
frames = bench.make_frames(args, 8)
f0 = frames[0]
tdev = lambda a: torch.from_numpy(a).to(dev)
cam = {k: (tdev(v) if isinstance(v, np.ndarray) else v) for k, v in f0["camera"].items()}
faces, fidx, bary = tdev(f0["faces"]), tdev(f0["face_index"]), tdev(f0["bary"])
from oracle import pose_oracle as po
_, canon = po.compute_face_orientation(tdev(f0["canon_verts"]), faces)
canon = canon.reshape(-1).contiguous()
params = [tdev(f0[k]) for k in ("scaling_raw", "rotation_raw", "offset_raw", "opacity_raw")]
shs, bg = tdev(f0["shs"]), tdev(f0["bg"])
fmodel = {k: tdev(f0[k]) for k in bench.FLAME_KEYS}; fmodel["parents"] = f0["parents"]
n_shape = f0["n_shape"]
host = []
for f in frames:
    c = f["camera"]
    h = dict(expression=torch.from_numpy(f["betas"][n_shape:])[None], flame_pose=torch.from_numpy(f["pose"])[None],
             view=torch.from_numpy(c["viewmatrix"]), proj=torch.from_numpy(c["projmatrix"]),
             campos=torch.from_numpy(c["campos"]), target=torch.rand(3, 512, 512))
    host.append({k: v.contiguous().pin_memory() for k, v in h.items()})
out_img = torch.empty(3, 512, 512).pin_memory(); out_loss = torch.empty(1).pin_memory()
leaves = [p_.clone().requires_grad_(True) for p_ in params] + [shs.clone().requires_grad_(True)]
dleaves = {k: tdev(f0[k]).requires_grad_(True) for k in bench.DELTA_KEYS}
zeros_shape = torch.zeros(1, n_shape, device=dev)

def e2e_step(i, sync=True):
    h = host[i % 8]
    d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
    for p_ in leaves + list(dleaves.values()):
        p_.grad = None
    full_betas = torch.cat([zeros_shape, d["expression"]], dim=1)
    vts, _, _, vts_orig, _ = flame.flame_lbs(fmodel, full_betas, d["flame_pose"], dleaves["delta_shapedirs"],
                                             dleaves["delta_posedirs"], dleaves["delta_vertex"], l0=n_shape)
    xyz, sc, ro, op = pose.pose_splats(vts, faces, fidx, bary, canon, *leaves[:4], shell_len=f0["shell_len"])
    settings = R.GaussianRasterizationSettings(512, 512, cam["tanfovx"], cam["tanfovy"], bg, 1.0, d["view"], d["proj"], 0,
                                               d["campos"], False, False)
    screen = torch.zeros_like(xyz, requires_grad=True)
    img, radii = R.GaussianRasterizer(settings)(means3D=xyz, means2D=screen, shs=leaves[4], opacities=op, scales=sc, rotations=ro)
    loss = (img - d["target"]).abs().mean()
    loss.backward()
    out_img.copy_(img.detach(), non_blocking=True)
    out_loss.copy_(loss.detach().reshape(1), non_blocking=True)
    if sync:
        torch.cuda.current_stream().synchronize()

for i in range(10): e2e_step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(100): e2e_step(i)
torch.cuda.synchronize(); print("e2e step wall us:", 1e4 * (time.perf_counter() - t0))
t0 = time.perf_counter()
for i in range(100): e2e_step(i, sync=False)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host-only us per step (no final sync):", 1e4 * (t1 - t0), " drain:", 1e4 * (t2 - t1))
pr = cProfile.Profile(); pr.enable()
for i in range(100): e2e_step(i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
