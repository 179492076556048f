import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from fateavatar_b200 import rasterizer as R, scenes
from oracle import oracle as orc
dev = torch.device("cuda:0")
sc = scenes.head_scene(P=5000, W=128, H=128, scale_mult=5.0, seed=1)
t = scenes.to_torch(sc, dev); cam = t["camera"]
rs = R.GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], t["bg"], 1.0, cam["viewmatrix"], cam["projmatrix"], 0, cam["campos"], False, False)
color, radii, st = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
torch.cuda.synchronize()
c = sc["camera"]
o = orc.forward(sc["means3D"], sc["opacities"], sc["bg"], c["viewmatrix"], c["projmatrix"], c["campos"], c["tanfovx"], c["tanfovy"], c["H"], c["W"], shs=sc["shs"], sh_degree=0, scales=sc["scales"], rotations=sc["rotations"])
taps = R.decode_workspace(st["workspace"], 5000, 128, 128, st["capacity"], st["num_rendered"])
col = color.cpu().numpy(); d = np.abs(col - o["color"]).max(0)
print("R", o["R"], "max tile", (o["ranges"][:,1]-o["ranges"][:,0]).max(), "n bad px", (d > 1e-5).sum(), "max", d.max())
nc = taps["n_contrib"].cpu().numpy(); print("n_contrib mismatches", (nc != o["n_contrib"].astype(np.int32)).sum())
fT = taps["final_T"].cpu().numpy(); print("final_T max diff", np.abs(fT - o["final_T"]).max())
ys, xs = np.nonzero(d > 1e-5)
print("bad pixels (x,y):", list(zip(xs[:20], ys[:20])))
ext = taps["extent"].cpu().numpy(); m2 = o["means2D"]; co = o["conic_opacity"]
for x, y in list(zip(xs, ys))[:3]:
    tile = (y // 16) * 8 + (x // 16)
    a, b = o["ranges"][tile]
    ids = o["point_list"][a:b]
    T = 1.0
    bx = (x // 8) * 8; by = (y // 4) * 4
    for pos, g in enumerate(ids):
        dx = m2[g, 0] - x; dy = m2[g, 1] - y
        power = -0.5 * (co[g, 0] * dx * dx + co[g, 2] * dy * dy) - co[g, 1] * dx * dy
        if power > 0: continue
        alpha = min(0.99, co[g, 3] * np.exp(power))
        if alpha < 1 / 255: continue
        ex, ey = ext[g]
        hit = not (ex < 0) and not (m2[g, 0] + ex < bx or m2[g, 0] - ex > bx + 7 or m2[g, 1] + ey < by or m2[g, 1] - ey > by + 3)
        if not hit:
            print(f"px({x},{y}) pos {pos} g {g}: alpha {alpha:.5f} power {power:.4f} T {T:.4f} CULLED! mean {m2[g]} ext {ext[g]} conic {co[g]} radii {o['radii'][g]}")
        if T * (1 - alpha) < 1e-4: break
        T *= 1 - alpha
