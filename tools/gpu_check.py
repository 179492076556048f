#!/usr/bin/env python
"""Developer probe run on the GPU box: new kernels vs C oracle vs compiled reference, with timings.
Not part of the product or the test-suite; prints a report and writes gpurun_out/gpu_check.json."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from fateavatar_b200 import rasterizer as R
from fateavatar_b200 import scenes
from oracle import oracle as orc
from oracle import ref_loader

dev = torch.device("cuda:0")
report = {}


def settings(cam, bg, deg):
    return R.GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"], bg, 1.0,
                                           cam["viewmatrix"], cam["projmatrix"], deg, cam["campos"], False, False)


def timeit(fn, warm=5, iters=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def cmp(name, a, b, exact=False):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    if a.shape != b.shape:
        return f"{name}: SHAPE {a.shape} vs {b.shape}"
    if exact or a.dtype.kind in "iub":
        nm = int((a != b).sum())
        return f"{name}: mismatches {nm}/{a.size}"
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    bits = int((a.view(np.uint32) != b.view(np.uint32)).sum()) if a.dtype == np.float32 and b.dtype == np.float32 else -1
    return f"{name}: maxabs {d.max() if d.size else 0:.3e} (ref max {np.abs(b).max() if b.size else 0:.3e}) bit-diff {bits}/{a.size}"


def run_scene(tag, sc, do_oracle=True, do_bwd=True):
    print(f"\n===== {tag}: {sc['name']}", flush=True)
    rep = {}
    t = scenes.to_torch(sc, dev)
    cam = t["camera"]
    deg = sc["sh_degree"]
    rs = settings(cam, t["bg"], deg)
    color, radii, st = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
    torch.cuda.synchronize()
    P = t["means3D"].shape[0]
    taps = R.decode_workspace(st["workspace"], P, cam["W"], cam["H"], st["capacity"], st["num_rendered"])
    info = taps["info"].cpu().numpy()
    print("new: R", st["num_rendered"], "visible", int((radii > 0).sum()), "max tile", info[3], "launches", st["launches"])
    rep.update(R=int(st["num_rendered"]), visible=int((radii > 0).sum()), max_tile=int(info[3]), P=P)
    dpix = torch.from_numpy(np.random.default_rng(7).standard_normal((3, cam["H"], cam["W"])).astype(np.float32)).to(dev)
    grads = None
    if do_bwd:
        grads = R.backward_raw(st, dpix)
        torch.cuda.synchronize()
    gnames = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")
    vis = (radii > 0).cpu().numpy()

    if do_oracle:
        t0 = time.time()
        o = orc.forward(sc["means3D"], sc["opacities"], sc["bg"], sc["camera"]["viewmatrix"], sc["camera"]["projmatrix"],
                        sc["camera"]["campos"], cam["tanfovx"], cam["tanfovy"], cam["H"], cam["W"], shs=sc["shs"],
                        sh_degree=deg, scales=sc["scales"], rotations=sc["rotations"])
        print(f"oracle fwd {time.time() - t0:.3f}s  R {o['R']}")
        print(" vs oracle:", cmp("radii", radii, o["radii"]))
        for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
            print(" vs oracle:", cmp(k, taps[k][torch.from_numpy(vis)], o[k][vis]))
        print(" vs oracle:", cmp("tiles_touched", taps["tiles_touched"], o["tiles_touched"].astype(np.int32)))
        print(" vs oracle:", cmp("ranges", taps["ranges"], o["ranges"].astype(np.int32)))
        if o["R"] == st["num_rendered"]:
            print(" vs oracle:", cmp("point_list", taps["point_list"], o["point_list"].astype(np.int32)))
        print(" vs oracle:", cmp("n_contrib", taps["n_contrib"], o["n_contrib"].astype(np.int32)))
        print(" vs oracle:", cmp("final_T", taps["final_T"], o["final_T"]))
        print(" vs oracle:", cmp("color", color, o["color"]))
        if do_bwd:
            t0 = time.time()
            og = orc.backward(o, dpix.cpu().numpy())
            print(f"oracle bwd {time.time() - t0:.3f}s")
            for k, g in zip(gnames, grads):
                print(" vs oracle:", cmp(k, g.reshape(og[k].shape), og[k]))

    if ref_loader.available():
        tr = dict(t)
        rst = ref_loader.ref_forward(tr, cam, sh_degree=deg)
        print("ref: R", rst["R"])
        rep["R_ref"] = int(rst["R"])
        print(" vs ref:", cmp("radii", radii, rst["radii"]))
        vt = torch.from_numpy(vis).to(dev)
        for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
            print(" vs ref:", cmp(k, taps[k][vt], rst[k][vt]))
        print(" vs ref:", cmp("clamped", taps["clamped"][vt], rst["clamped"][vt]))
        print(" vs ref:", cmp("tiles_touched", taps["tiles_touched"], rst["tiles_touched"]))
        print(" vs ref:", cmp("ranges", taps["ranges"], rst["ranges"]))
        if rst["R"] == st["num_rendered"]:
            print(" vs ref:", cmp("point_list", taps["point_list"], rst["point_list"]))
        print(" vs ref:", cmp("n_contrib", taps["n_contrib"], rst["n_contrib"]))
        print(" vs ref:", cmp("final_T", taps["final_T"], rst["final_T"]))
        print(" vs ref:", cmp("color", color, rst["color"]))
        if do_bwd:
            rg = ref_loader.ref_backward(rst, dpix)
            for k, g in zip(gnames, grads):
                print(" vs ref:", cmp(k, g.reshape(rg[k].shape), rg[k]))
            if do_oracle:
                for k in gnames:
                    print(" ref vs oracle:", cmp(k, rg[k].reshape(og[k].shape), og[k]))
        # timings
        e = torch.Tensor([])

        def ref_f():
            return ref_loader.ref_dgr().rasterize_gaussians(*rst["args"])

        def ref_fb():
            Rr, c, rad, g, b, i = ref_loader.ref_dgr().rasterize_gaussians(*rst["args"])
            a = rst["args"]
            ref_loader.ref_dgr().rasterize_gaussians_backward(a[0], a[1], rad, a[2], a[4], a[5], a[6], a[7], a[8], a[9],
                                                              a[10], a[11], dpix, a[14], a[15], a[16], g, Rr, b, i, False)

        rep["ref_fwd_ms"] = timeit(ref_f)
        rep["ref_fwdbwd_ms"] = timeit(ref_fb)
        print(f"ref  fwd {rep['ref_fwd_ms']:.3f} ms   fwd+bwd {rep['ref_fwdbwd_ms']:.3f} ms")

    def new_f():
        return R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)

    def new_fb():
        c, r, s = R.forward_raw(rs, t["means3D"], t["shs"], None, t["opacities"], t["scales"], t["rotations"], None)
        R.backward_raw(s, dpix)

    rep["new_fwd_ms"] = timeit(new_f)
    rep["new_fwdbwd_ms"] = timeit(new_fb)
    print(f"new(sync)  fwd {rep['new_fwd_ms']:.3f} ms   fwd+bwd {rep['new_fwdbwd_ms']:.3f} ms")
    R.set_async(True)
    rep["new_async_fwd_ms"] = timeit(new_f)
    rep["new_async_fwdbwd_ms"] = timeit(new_fb)
    R.set_async(False)
    from fateavatar_b200 import _lib
    _lib.load().fs_profile_enable(1)
    _lib.profile_read()
    for _ in range(20):
        new_fb()
    prof = _lib.profile_read()
    _lib.load().fs_profile_enable(0)
    rep["stage_us"] = {k: round(1000 * v[0] / max(v[1], 1), 2) for k, v in prof.items() if v[1]}
    print("stage us:", rep["stage_us"])
    print(f"new(async) fwd {rep['new_async_fwd_ms']:.3f} ms   fwd+bwd {rep['new_async_fwdbwd_ms']:.3f} ms")
    report[tag] = rep


def main():
    print(torch.cuda.get_device_name(0), "ref available:", ref_loader.available(), "oracle threads", orc.num_threads())
    which = sys.argv[1:] or ["c1", "c2small", "c2", "c2sh3", "c5"]
    jobs = {
        "c1": lambda: run_scene("c1", scenes.config1_scene()),
        "smoke": lambda: run_scene("smoke", scenes.head_scene(P=5000, W=128, H=128, scale_mult=5.0, seed=1)),
        "c2small": lambda: run_scene("c2small", scenes.head_scene(P=20000, W=300, H=200, scale_mult=3.0)),
        "c2": lambda: run_scene("c2", scenes.head_scene()),
        "c2big": lambda: run_scene("c2big", scenes.head_scene(scale_mult=4.0)),
        "c2sh3": lambda: run_scene("c2sh3", scenes.head_scene(P=30000, sh_degree=3, scale_mult=2.0)),
        "c5": lambda: run_scene("c5", scenes.stress_scene(), do_oracle=True),
    }
    for w in which:
        try:
            jobs[w]()
        except Exception:
            traceback.print_exc()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report))


if __name__ == "__main__":
    main()
