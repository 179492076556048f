#!/usr/bin/env python
"""Stress the forward for run-to-run determinism on the GPU: the same frame N times, every forward output compared
bit for bit with the first run (colour, final_T, n_contrib, point_list, ranges).  Gradients are float-atomic sums and
are compared with the parity tolerance instead.  Exercises the capacity re-run path by clearing the hint now and then."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fateavatar_b200 import rasterizer as R, scenes
from util import settings, GRAD_NAMES

def once(t, sc, dpix):
    cam = t["camera"]
    rs = settings(R, cam, t["bg"], sc["sh_degree"], 1.0)
    color, radii, st = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
    P = t["means3D"].shape[0]
    taps = R.decode_workspace(st["workspace"], P, cam["W"], cam["H"], st["capacity"], st["num_rendered"])
    out = {k: taps[k].clone() for k in ("final_T", "n_contrib", "point_list", "ranges")}
    out["color"] = color.clone(); out["radii"] = radii.clone()
    g = R.backward_raw(st, dpix)
    out.update({k: v.clone() for k, v in zip(GRAD_NAMES, g)})
    torch.cuda.synchronize()
    # the backward must leave every forward output alone (a stray reduction would land in one of them)
    assert torch.equal(color, out["color"]) and torch.equal(radii, out["radii"]), "backward modified a forward output"
    for k in ("final_T", "n_contrib", "point_list", "ranges"):
        assert torch.equal(taps[k], out[k]), f"backward modified {k}"
    return out

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda:0")
    cases = {"mirror_3k_96x80": scenes.head_scene(P=3000, W=96, H=80, scale_mult=6.0, seed=41),
             "smoke_5k_128": scenes.head_scene(P=5000, W=128, H=128, scale_mult=5.0, seed=1),
             "head_20k_300x200": scenes.head_scene(P=20000, W=300, H=200, scale_mult=3.0)}
    bad_total = 0
    for name, sc in cases.items():
        t = scenes.to_torch(sc, dev)
        cam = t["camera"]
        dpix = torch.randn(3, cam["H"], cam["W"], device=dev)
        ref = once(t, sc, dpix)
        bad = {}
        for i in range(n):
            if i % 7 == 3:
                R._capacity_hint.clear()  # next frame starts from the first-frame capacity guess again
            cur = once(t, sc, dpix)
            for k, v in cur.items():
                if k in GRAD_NAMES:
                    s = float(ref[k].abs().max().clamp_min(1e-12))
                    if float((v - ref[k]).abs().max()) > 2e-4 * s:
                        bad[k] = bad.get(k, 0) + 1
                elif not torch.equal(v.view(torch.int32) if v.dtype == torch.float32 else v,
                                     ref[k].view(torch.int32) if ref[k].dtype == torch.float32 else ref[k]):
                    bad[k] = bad.get(k, 0) + 1
        print(name, "runs", n, "R", int(ref["point_list"].numel()), "mismatches", bad or "none", flush=True)
        bad_total += sum(bad.values())
    sys.exit(1 if bad_total else 0)

if __name__ == "__main__":
    main()
