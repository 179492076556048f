#!/usr/bin/env python
"""Golden fixture for the FLAME skinning stage, made by importing the REFERENCE's own flame/lbs.py on the CPU
(this container only; /root/reference does not travel):

    python tests/golden/make_flame_golden.py        # writes tests/golden/flame_small.npz

The fixture holds outputs only (float32 forward of lbs() with and without the personalised deltas, and the
float64 autograd gradients of the three delta tensors for a seeded upstream gradient); the inputs are regenerated
at test time from fateavatar_b200.scenes.flame_inputs(seed=21, V=150).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fateavatar_b200 import scenes  # noqa: E402

REF_LBS = "/root/reference/flame/lbs.py"
CASE = dict(seed=21, V=150)


def load_ref_lbs():
    spec = importlib.util.spec_from_file_location("ref_flame_lbs", REF_LBS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_forward(ref, f, dtype, deltas):
    """Exactly the call sequence of flame/FLAME.py:180-202 (deltas) / :146-152 (no deltas), batch of one."""
    t = lambda k: torch.from_numpy(f[k]).to(dtype)
    betas, pose = t("betas")[None], t("pose")[None]
    vt, sd, pd = t("v_template")[None], t("shapedirs"), t("posedirs")
    leaves = None
    if deltas:
        leaves = [t(k).requires_grad_(True) for k in ("delta_vertex", "delta_shapedirs", "delta_posedirs")]
        vt = vt + leaves[0][None]
        sd = sd + leaves[1]
        pd = pd + leaves[2]
    verts, pf, A = ref.lbs(betas, pose, vt, sd, pd, t("J_regressor"), torch.from_numpy(f["parents"]), t("lbs_weights"),
                           dtype=dtype)
    return verts[0], pf[0], A[0], leaves


def main():
    ref = load_ref_lbs()
    f = scenes.flame_inputs(**CASE)
    out = {}
    v, pf, A, _ = ref_forward(ref, f, torch.float32, True)
    vo, _, Ao, _ = ref_forward(ref, f, torch.float32, False)
    out.update(verts=v.detach().numpy(), pose_feature=pf.detach().numpy(), A=A.detach().numpy(),
               verts_orig=vo.numpy(), A_orig=Ao.numpy())
    g = np.random.default_rng(5).standard_normal(v.shape).astype(np.float32)
    v64, _, _, leaves = ref_forward(ref, f, torch.float64, True)
    (v64 * torch.from_numpy(g).double()).sum().backward()
    out.update(d_delta_vertex=leaves[0].grad.numpy(), d_delta_shapedirs=leaves[1].grad.numpy().astype(np.float32),
               d_delta_posedirs=leaves[2].grad.numpy())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flame_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: a.shape for k, a in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
