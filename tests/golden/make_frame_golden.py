#!/usr/bin/env python
"""Whole-frame golden fixture made by executing the REFERENCE's own `FateAvatar.forward` (model/fateavatar.py:196-298)
on the CPU through tests/ref_frame_harness.py (this container only; /root/reference does not travel):

    python tests/golden/make_frame_golden.py        # writes tests/golden/frame_small.npz

Outputs only: image, both meshes, radii, and the gradients of every trained parameter for a seeded upstream gradient
(delta_shapedirs: the 100 expression columns; the 300 shape columns are identically zero).  Inputs are regenerated at
test time from fateavatar_b200.scenes.small_avatar(seed=51, n_lat=9, n_lon=16, N=900).  See the harness docstring for
what is substituted (CUDA rasterizer -> C oracle, pytorch3d quaternion helpers -> this repo's restatements).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_frame_harness as H  # noqa: E402
from fateavatar_b200 import scenes  # noqa: E402

CASE = dict(seed=51, n_lat=9, n_lon=16, N=900)
RES = (64, 80)


def main():
    patch = H.Patch()
    try:
        FateAvatar, FLAME, mesh_compute = H.load_reference(patch)
        a = scenes.small_avatar(**CASE)
        ref = H.build_reference_model(FateAvatar, FLAME, mesh_compute, a, RES)
        out = FateAvatar.forward(ref, H.frame_input(a))
        w = torch.from_numpy(np.random.default_rng(2).standard_normal((3,) + RES).astype(np.float32))
        (out["rgb_image"][0] * w).sum().backward()
        res = dict(rgb_image=out["rgb_image"][0].detach().numpy(), verts=out["verts"][0].detach().numpy(),
                   verts_orig=out["verts_orig"][0].detach().numpy(), radii=out["radii"][0].numpy())
        for n in H.PARAMS:
            g = getattr(ref, n).grad.numpy()
            assert n != "delta_shapedirs" or not g[:, :, :300].any()
            res["grad" + n] = g[:, :, 300:] if n == "delta_shapedirs" else g
    finally:
        patch.undo()
    path = os.path.join(HERE, "frame_small.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, {k: v.shape for k, v in res.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
