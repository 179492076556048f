#!/usr/bin/env python
"""Generate golden fixtures from the COMPILED REFERENCE (oracle/_ref, built by oracle/build_ref.py from the
sources under /root/reference) on a CUDA device:

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each fixture stores the reference's decoded per-stage outputs for a seeded scene from fateavatar_b200.scenes
(named below, regenerated at test time, so the fixture holds outputs only).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fateavatar_b200 import scenes  # noqa: E402

SCENES = {
    "c1_small": lambda: scenes.config1_scene(P=2000, W=128, H=96, seed=11),
    "head_sh0": lambda: scenes.head_scene(P=4000, W=160, H=144, sh_degree=0, scale_mult=4.0, seed=12),
    "head_sh3": lambda: scenes.head_scene(P=3000, W=128, H=128, sh_degree=3, scale_mult=5.0, seed=13),
}


def scene_from_name(name):
    return SCENES[name]()


def dpix_for(sc, seed):
    cam = sc["camera"]
    return np.random.default_rng(seed).standard_normal((3, cam["H"], cam["W"])).astype(np.float32)


def main():
    import torch

    from oracle import ref_loader

    assert ref_loader.available(), "oracle/_ref not built"
    dev = torch.device("cuda:0")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, mk in SCENES.items():
        sc = mk()
        t = scenes.to_torch(sc, dev)
        st = ref_loader.ref_forward(t, t["camera"], sh_degree=sc["sh_degree"])
        seed = 5
        g = ref_loader.ref_backward(st, torch.from_numpy(dpix_for(sc, seed)).to(dev))
        keep = dict(scene=name, R=st["R"], dL_dpix_seed=seed)
        for k in ("color", "radii", "depths", "means2D", "cov3D", "conic_opacity", "rgb", "tiles_touched", "ranges",
                  "point_list", "n_contrib", "final_T", "clamped"):
            keep[k] = st[k].cpu().numpy()
        for k, v in g.items():
            keep[k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **keep)
        print(name, "R", st["R"])


if __name__ == "__main__":
    main()
