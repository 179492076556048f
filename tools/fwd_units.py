#!/usr/bin/env python
"""Work distribution of the forward blend for one bench frame: per (tile, 8x4 block) unit, records walked until the
block's pixels have all terminated and bounding-box survivors among them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, types
import bench
from fateavatar_b200 import parallel, rasterizer as R

class A: P = 100000; res = 512
args = A(); dev = torch.device("cuda:0")
frames = bench.make_frames(args, 1); f0 = frames[0]
tdev = lambda a: torch.from_numpy(a).to(dev)
from oracle import pose_oracle as po, flame_oracle as fo
m = {k: tdev(f0[k]) for k in bench.FLAME_KEYS}; m["parents"] = torch.tensor(f0["parents"], device=dev)
with torch.device(dev):
    verts, _, _ = fo.forward_with_delta_blendshape(m, tdev(f0["betas"]), tdev(f0["pose"]), tdev(f0["delta_shapedirs"]), tdev(f0["delta_posedirs"]), tdev(f0["delta_vertex"]))
_, canon = po.compute_face_orientation(tdev(f0["canon_verts"]), tdev(f0["faces"]))
xyz, sc, ro, op = po.pose_splats(verts, tdev(f0["faces"]), tdev(f0["face_index"]), tdev(f0["bary"]), canon.reshape(-1, 1), tdev(f0["scaling_raw"]), tdev(f0["rotation_raw"]), tdev(f0["offset_raw"]), tdev(f0["opacity_raw"]), shell_len=f0["shell_len"])
cam = {k: (tdev(v) if isinstance(v, np.ndarray) else v) for k, v in f0["camera"].items()}
rs = R.GaussianRasterizationSettings(512, 512, cam["tanfovx"], cam["tanfovy"], tdev(f0["bg"]), 1.0, cam["viewmatrix"], cam["projmatrix"], 0, cam["campos"], False, False)
color, radii, st = R.forward_raw(rs, xyz.contiguous(), tdev(f0["shs"]), None, op.contiguous(), sc.contiguous(), ro.contiguous(), None)
torch.cuda.synchronize()
t = R.decode_workspace(st["workspace"], args.P, 512, 512, st["capacity"], st["num_rendered"])
ranges = t["ranges"].cpu().numpy(); recs = t["inst_splat"].cpu().numpy(); nc = t["n_contrib"].cpu().numpy(); fT = t["final_T"].cpu().numpy()
units = []
for tile in range(1024):
    lo, hi = ranges[tile]
    if hi <= lo: continue
    ty, tx = divmod(tile, 32)
    r = recs[lo:hi]
    for blk in range(8):
        bx, by = tx * 16 + (blk & 1) * 8, ty * 16 + (blk >> 1) * 4
        sub_nc, sub_T = nc[by:by + 4, bx:bx + 8], fT[by:by + 4, bx:bx + 8]
        # walked = whole list unless every pixel terminated (T < 1e-4 test) -> approximated by final_T small everywhere
        walked = hi - lo if (sub_T > 2e-4).any() else min(hi - lo, int(sub_nc.max()) + 32)
        rr = r[:walked]
        hit = ~(rr[:, 2] < 0) & ~((rr[:, 0] + rr[:, 2] < bx) | (rr[:, 0] - rr[:, 2] > bx + 7) | (rr[:, 1] + rr[:, 3] < by) | (rr[:, 1] - rr[:, 3] > by + 3))
        units.append((hi - lo, walked, int(hit.sum())))
u = np.array(units)
print("units", len(u), "records total", u[:, 0].sum(), "walked", u[:, 1].sum(), "survivors", u[:, 2].sum())
cost = u[:, 1] / 32 * 25 + u[:, 2] / 4 * 140   # warp instructions, rough
print("cost sum (M instr)", cost.sum() / 1e6, "max unit (k instr)", cost.max() / 1e3, "mean", cost.mean() / 1e3)
o = np.argsort(-cost)[:10]
print("top units (n, walked, survivors):", u[o].tolist())
print("survivor hit rate overall", u[:, 2].sum() / u[:, 1].sum(), " early-terminated units", int((u[:, 1] < u[:, 0]).sum()))
for q in (50, 90, 99, 100): print("cost percentile", q, np.percentile(cost, q) / 1e3)
