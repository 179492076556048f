"""CPU tests of the oracle itself (no GPU): independent autograd cross-check, invariants, golden fixtures."""
import glob
import os

import numpy as np
import pytest

from fateavatar_b200 import scenes
from oracle import oracle as orc
from util import rasterizer_goldens, oracle_forward

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_oracle_matches_dense_autograd(deg):
    import dense_ref

    sc = scenes.head_scene(P=300, W=64, H=48, sh_degree=deg, scale_mult=12.0, seed=3 + deg)
    st = oracle_forward(orc, sc)
    dpix = np.random.default_rng(1).standard_normal((3, 48, 64)).astype(np.float32)
    g = orc.backward(st, dpix)
    d = dense_ref.dense_gradients(sc, sc["camera"], st["radii"], dpix)
    assert np.abs(d["color"] - st["color"]).max() < 2e-6
    for k in ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh", "dL_dmeans2D"):
        a, b = g[k].astype(np.float64), d[k].reshape(g[k].shape)
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(b).max(), 1e-9), k


def test_oracle_sort_invariants():
    sc = scenes.config1_scene(P=3000)
    st = oracle_forward(orc, sc)
    R = st["R"]
    assert R == int(st["tiles_touched"].sum()) == int(st["point_offsets"][-1])
    rng, pl = st["ranges"], st["point_list"]
    depth_bits = st["depths"].view(np.uint32)
    covered = 0
    for t in range(rng.shape[0]):
        a, b = int(rng[t, 0]), int(rng[t, 1])
        covered += b - a
        ids = pl[a:b].astype(np.int64)
        key = (depth_bits[ids].astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
        assert np.all(key[1:] > key[:-1])  # strictly increasing (depth bits, gaussian id)
    assert covered == R
    assert np.all(st["n_contrib"] <= (rng[:, 1] - rng[:, 0]).max())


def test_oracle_edge_cases():
    sc = scenes.config1_scene(P=500, W=50, H=37)  # not multiples of 16
    st = oracle_forward(orc, sc)
    assert st["color"].shape == (3, 37, 50) and np.isfinite(st["color"]).all()
    # everything behind the near plane -> nothing rendered, image == background
    sc2 = dict(sc)
    sc2["means3D"] = sc["means3D"] + np.array([0, 0, 10], np.float32)  # view z = -z + 2.5 < 0.2
    st2 = oracle_forward(orc, sc2)
    assert st2["R"] == 0 and (st2["radii"] == 0).all()
    assert np.allclose(st2["color"], 1.0)
    # P == 0
    st3 = orc.forward(np.zeros((0, 3), np.float32), np.zeros((0, 1), np.float32), sc["bg"], sc["camera"]["viewmatrix"],
                      sc["camera"]["projmatrix"], sc["camera"]["campos"], 0.2, 0.2, 16, 16,
                      shs=np.zeros((0, 1, 3), np.float32), scales=np.zeros((0, 3), np.float32),
                      rotations=np.zeros((0, 4), np.float32))
    assert st3["R"] == 0
    # colours may be supplied precomputed instead of SH
    col = np.random.default_rng(0).uniform(0, 1, (500, 3)).astype(np.float32)
    st4 = oracle_forward(orc, sc, shs=None, colors_precomp=col)
    assert np.array_equal(st4["radii"], st["radii"]) and not np.allclose(st4["color"], st["color"])


def test_oracle_mark_visible_and_knn():
    sc = scenes.config1_scene(P=2000)
    vis = orc.mark_visible(sc["means3D"], sc["camera"]["viewmatrix"], sc["camera"]["projmatrix"])
    assert vis.all()
    pts = np.random.default_rng(0).standard_normal((3000, 3)).astype(np.float32)
    got = orc.knn_mean_dist2(pts)  # grid path
    d2 = ((pts[:400, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    d2[np.arange(400), np.arange(400)] = np.inf
    ref = np.sort(d2, axis=1)[:, :3].mean(1)
    assert np.allclose(got[:400], ref, rtol=1e-5)
    small = orc.knn_mean_dist2(pts[:500])  # brute-force path
    d2 = ((pts[:500, None, :].astype(np.float64) - pts[None, :500, :]) ** 2).sum(-1)
    d2[np.arange(500), np.arange(500)] = np.inf
    assert np.allclose(small, np.sort(d2, axis=1)[:, :3].mean(1), rtol=1e-5)


GOLDEN = rasterizer_goldens(HERE)


@pytest.mark.parametrize("path", GOLDEN or [None])
def test_oracle_against_reference_golden(path):
    """Pins the oracle to outputs of the compiled reference run on a B200 (tests/golden/make_golden.py).
    Integer/index outputs and per-Gaussian floats must be bit-identical; blended outputs within 1e-6 (the only
    difference is CUDA's MUFU-based expf vs glibc's)."""
    if path is None:
        pytest.skip("no golden fixtures committed yet")
    import golden.make_golden as mg

    z = np.load(path)
    sc = mg.scene_from_name(str(z["scene"]))
    st = oracle_forward(orc, sc)
    assert st["R"] == int(z["R"])
    assert np.array_equal(st["radii"], z["radii"])
    assert np.array_equal(st["tiles_touched"].astype(np.int32), z["tiles_touched"])
    assert np.array_equal(st["ranges"].astype(np.int32), z["ranges"])
    assert np.array_equal(st["point_list"].astype(np.int32), z["point_list"])
    vis = st["radii"] > 0
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
        assert np.array_equal(st[k][vis].view(np.uint32), z[k][vis].view(np.uint32)), k
    orc.compare_blend(st, z["color"], z["final_T"], z["n_contrib"])
    if "dL_dpix_seed" in z:
        dpix = mg.dpix_for(sc, int(z["dL_dpix_seed"]))
        g = orc.backward(st, dpix)
        for k in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations"):
            ref = z[k].astype(np.float64)
            err = np.abs(g[k].astype(np.float64).reshape(ref.shape) - ref).max()
            assert err <= 2e-4 * max(np.abs(ref).max(), 1e-12), (k, err)


def test_pair_counts_match_a_python_recount():
    """orc_blend_pair_counts (the pairs/s figure of bench.py) against a plain-Python walk of the same lists."""
    sc = scenes.head_scene(P=80, W=24, H=20, scale_mult=30.0, seed=9)
    o = oracle_forward(orc, sc)
    got = orc.pair_counts(o)
    walked = evals = contrib = 0
    gx = 2
    for y in range(20):
        for x in range(24):
            a, b = o["ranges"][(y // 16) * gx + x // 16]
            T = np.float32(1.0)
            for j in range(int(a), int(b)):
                walked += 1
                g = o["point_list"][j]
                co, m = o["conic_opacity"][g], o["means2D"][g]
                dx, dy = np.float32(m[0] - x), np.float32(m[1] - y)
                power = np.float32(-0.5) * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy
                if power > 0:
                    continue
                evals += 1
                alpha = min(np.float32(0.99), np.float32(co[3] * np.exp(power)))
                if alpha < np.float32(1.0 / 255.0):
                    continue
                test_T = np.float32(T * (np.float32(1.0) - alpha))
                if test_T < np.float32(0.0001):
                    break
                T = test_T
                contrib += 1
    assert got["walked"] == walked and abs(got["exp_evaluated"] - evals) <= 2 and abs(got["contributing"] - contrib) <= 2
    assert got["contributing"] <= got["exp_evaluated"] <= got["walked"] <= int((o["ranges"][:, 1] - o["ranges"][:, 0]).max()) * 24 * 20
