"""FLAME skinning stage: the oracle is pinned against the reference's own flame/lbs.py (when the
tree is mounted) and against the committed fixture made from it; the fused kernels are compared with the oracle's
float64 forward / autograd backward and with the fixture, through the C ABI."""
import ctypes as C
import importlib.util
import os
import types

import numpy as np
import pytest
import torch

from fateavatar_b200 import _lib, scenes
from oracle import flame_oracle as fo

REF_LBS = "/root/reference/flame/lbs.py"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flame_small.npz")
GOLDEN_CASE = dict(seed=21, V=150)  # tests/golden/make_flame_golden.py
DELTAS = ("delta_vertex", "delta_shapedirs", "delta_posedirs")


def _model(f, dtype, device="cpu"):
    m = {k: torch.from_numpy(f[k]).to(dtype).to(device) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
    m["parents"] = torch.from_numpy(f["parents"])
    return m


def oracle_run(f, dtype=torch.float64, upstream=None, deltas=True):
    m = _model(f, dtype)
    t = lambda k: torch.from_numpy(f[k]).to(dtype)
    leaves = {k: t(k).requires_grad_(True) for k in DELTAS} if deltas else {}
    verts, pf, A = fo.forward_with_delta_blendshape(m, t("betas"), t("pose"), leaves.get("delta_shapedirs"),
                                                    leaves.get("delta_posedirs"), leaves.get("delta_vertex"))
    res = dict(verts=verts.detach().numpy(), pose_feature=pf.detach().numpy(), A=A.detach().numpy())
    if upstream is not None:
        (verts * torch.from_numpy(upstream).to(dtype)).sum().backward()
        res["grads"] = {k: leaves[k].grad.numpy() for k in DELTAS}
    return res


# ---------------------------------------------------------------------------------------------- CPU: the oracle
@pytest.mark.skipif(not os.path.exists(REF_LBS), reason="reference tree not mounted")
@pytest.mark.parametrize("deltas", [True, False])
def test_oracle_matches_reference_lbs_file(deltas):
    spec = importlib.util.spec_from_file_location("ref_flame_lbs", REF_LBS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    f = scenes.flame_inputs(seed=3, V=97)
    t = lambda k: torch.from_numpy(f[k]).double()
    leaves = [t(k).requires_grad_(True) for k in DELTAS] if deltas else [None] * 3
    vt = t("v_template")[None] + (leaves[0][None] if deltas else 0)
    sd = t("shapedirs") + (leaves[1] if deltas else 0)
    pd = t("posedirs") + (leaves[2] if deltas else 0)
    v, pf, A = ref.lbs(t("betas")[None], t("pose")[None], vt, sd, pd, t("J_regressor"), torch.from_numpy(f["parents"]),
                       t("lbs_weights"), dtype=torch.float64)
    g = np.random.default_rng(0).standard_normal((97, 3))
    o = oracle_run(f, torch.float64, upstream=g if deltas else None, deltas=deltas)
    np.testing.assert_allclose(o["verts"], v[0].detach().numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(o["pose_feature"], pf[0].detach().numpy(), rtol=0, atol=1e-14)
    np.testing.assert_allclose(o["A"], A[0].detach().numpy(), rtol=0, atol=1e-13)
    if deltas:
        (v[0] * torch.from_numpy(g)).sum().backward()
        for k, leaf in zip(DELTAS, leaves):
            np.testing.assert_allclose(o["grads"][k], leaf.grad.numpy(), rtol=0, atol=1e-12 * max(1.0, float(leaf.grad.abs().max())))


def test_oracle_matches_golden_fixture_from_reference():
    gold = np.load(GOLDEN)
    f = scenes.flame_inputs(**GOLDEN_CASE)
    g = np.random.default_rng(5).standard_normal((GOLDEN_CASE["V"], 3)).astype(np.float32)
    o = oracle_run(f, torch.float64, upstream=g)
    np.testing.assert_allclose(o["verts"], gold["verts"], rtol=0, atol=2e-7)       # fixture forward is fp32
    np.testing.assert_allclose(o["pose_feature"], gold["pose_feature"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(o["A"], gold["A"], rtol=0, atol=2e-7)
    o0 = oracle_run(f, torch.float64, deltas=False)
    np.testing.assert_allclose(o0["verts"], gold["verts_orig"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(o0["A"], gold["A_orig"], rtol=0, atol=2e-7)
    for k in DELTAS:
        ref = gold["d_" + k]
        np.testing.assert_allclose(o["grads"][k], ref, rtol=0, atol=2e-7 * max(1.0, float(np.abs(ref).max())))


def test_flame_abi_rejects_bad_arguments_without_touching_the_gpu():
    lib = _lib.load()
    bad_tree = (C.c_int * 5)(-1, 0, 3, 1, 1)  # parents[2] = 3 is not < 2
    ok_tree = (C.c_int * 5)(-1, 0, 1, 1, 1)
    nul = [None] * 10 + [None] * 5
    assert lib.fs_flame_forward(10, 8, 0, 5, bad_tree, *nul, None, 0, None) == -1
    assert b"kinematic" in lib.fs_last_error()
    assert lib.fs_flame_forward(0, 8, 0, 5, ok_tree, *nul, None, 0, None) == -1
    assert lib.fs_flame_forward(10, 8, 9, 5, ok_tree, *nul, None, 0, None) == -1          # l0 > L
    assert lib.fs_flame_forward(10, 8, 0, 9, ok_tree, *nul, None, 0, None) == -1          # J > FS_FLAME_MAX_JOINTS
    assert lib.fs_flame_backward(10, 8, 0, 5, bad_tree, None, None, None, None, None, 0, None, None, None, None, None, None, None) == -1
    assert lib.fs_flame_workspace_bytes(0) == 0


# ---------------------------------------------------------------------------------------------- GPU: the kernels
def _run_gpu(f, dev, l0, deltas=True, upstream=None, want_orig=True):
    from fateavatar_b200 import flame

    m = _model(f, torch.float32, dev)
    m["parents"] = [int(x) for x in f["parents"]]
    t = lambda k: torch.from_numpy(f[k]).to(dev)
    leaves = {k: t(k).requires_grad_(True) for k in DELTAS} if deltas else {}
    outs = flame.flame_lbs(m, t("betas")[None], t("pose")[None], leaves.get("delta_shapedirs"), leaves.get("delta_posedirs"),
                           leaves.get("delta_vertex"), l0=l0, want_orig=want_orig)
    res = dict(verts=outs[0][0].detach().cpu().numpy(), pose_feature=outs[1][0].cpu().numpy(), A=outs[2][0].cpu().numpy())
    if want_orig:
        res.update(verts_orig=outs[3][0].cpu().numpy(), A_orig=outs[4][0].cpu().numpy())
    if upstream is not None:
        (outs[0][0] * torch.from_numpy(upstream).to(dev)).sum().backward()
        res["grads"] = {k: leaves[k].grad.cpu().numpy() for k in DELTAS}
    torch.cuda.synchronize()
    return res


def _check(got, o, o0, atol_v=1e-6):
    # tolerance: fp32 sums of <= 400 + 36 + 5 terms in a different order than cuBLAS's; vertices are ~0.15 in size
    assert np.abs(got["verts"] - o["verts"]).max() <= atol_v
    assert np.abs(got["pose_feature"] - o["pose_feature"]).max() <= 2e-7
    assert np.abs(got["A"] - o["A"]).max() <= 1e-6
    if o0 is not None:
        assert np.abs(got["verts_orig"] - o0["verts"]).max() <= atol_v
        assert np.abs(got["A_orig"] - o0["A"]).max() <= 1e-6
    if "grads" in got:
        for k in DELTAS:
            ref = o["grads"][k]
            err = np.abs(got["grads"][k] - ref).max()
            assert err <= 1e-5 * max(float(np.abs(ref).max()), 1e-12), (k, err, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["golden_150", "flame_size_5002", "ragged_v_33"])
def test_flame_kernels_vs_oracle(case, cuda_device):
    f = {"golden_150": lambda: scenes.flame_inputs(**GOLDEN_CASE), "flame_size_5002": lambda: scenes.flame_inputs(seed=1),
         "ragged_v_33": lambda: scenes.flame_inputs(seed=2, V=33)}[case]()
    V = f["v_template"].shape[0]
    g = np.random.default_rng(5).standard_normal((V, 3)).astype(np.float32)
    got = _run_gpu(f, cuda_device, l0=f["n_shape"], upstream=g)
    _check(got, oracle_run(f, upstream=g), oracle_run(f, deltas=False))
    # structural zeros: coefficients below l0 never receive gradient
    assert not got["grads"]["delta_shapedirs"][:, :, : f["n_shape"]].any()
    if case == "golden_150":
        gold = np.load(GOLDEN)
        assert np.abs(got["verts"] - gold["verts"]).max() <= 1e-6
        assert np.abs(got["verts_orig"] - gold["verts_orig"]).max() <= 1e-6
        for k in DELTAS:
            ref = gold["d_" + k]
            assert np.abs(got["grads"][k] - ref).max() <= 1e-5 * float(np.abs(ref).max())


@pytest.mark.gpu
def test_flame_variants_no_deltas_dense_betas_scalar_path_three_joints(cuda_device):
    # no deltas at all (FLAME.forward), verts_orig not requested
    f = scenes.flame_inputs(seed=4, V=77, with_deltas=False)
    got = _run_gpu(f, cuda_device, l0=f["n_shape"], deltas=False, want_orig=False)
    _check(got, oracle_run(f, deltas=False), None)
    # shape coefficients non-zero, l0 = 0 (every column read)
    f = scenes.flame_inputs(seed=5, V=64)
    f["betas"][:300] = 0.3 * np.random.default_rng(1).standard_normal(300).astype(np.float32)
    g = np.random.default_rng(6).standard_normal((64, 3)).astype(np.float32)
    _check(_run_gpu(f, cuda_device, l0=0, upstream=g), oracle_run(f, upstream=g), oracle_run(f, deltas=False))
    # coefficient counts that are not multiples of 4 (scalar load path) and a 3-joint tree
    f = scenes.flame_inputs(seed=6, V=50, n_shape=7, n_exp=11, J=3)
    g = np.random.default_rng(7).standard_normal((50, 3)).astype(np.float32)
    _check(_run_gpu(f, cuda_device, l0=7, upstream=g), oracle_run(f, upstream=g), oracle_run(f, deltas=False))


@pytest.mark.gpu
def test_attach_rebinds_reference_flame_module_methods(cuda_device):
    """The two calls of model/fateavatar.py:211-222 on a FLAME-module look-alike: same results as the oracle,
    second call served from the first, gradients reach the delta parameters, expression grads refuse loudly."""
    from fateavatar_b200 import flame
    from fateavatar_b200._lib import FateSplatError

    f = scenes.flame_inputs(seed=8, V=120)
    mod = types.SimpleNamespace(n_shape=f["n_shape"], n_exp=f["n_exp"], parents=torch.from_numpy(f["parents"]))
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
        setattr(mod, k, torch.from_numpy(f[k]).to(cuda_device))
    flame.attach(mod)
    expr = torch.from_numpy(f["betas"][300:])[None].to(cuda_device)
    pose = torch.from_numpy(f["pose"])[None].to(cuda_device)
    leaves = {k: torch.from_numpy(f[k]).to(cuda_device).requires_grad_(True) for k in DELTAS}
    verts, pf, A = mod.forward_with_delta_blendshape(expression_params=expr, full_pose=pose, delta_shapedirs=leaves["delta_shapedirs"],
                                                     delta_posedirs=leaves["delta_posedirs"], delta_vertex=leaves["delta_vertex"])
    lc0 = _lib.load().fs_last_launch_count()
    verts_orig, _, _ = mod.forward(expression_params=expr, full_pose=pose)
    assert verts.shape == (1, 120, 3) and pf.shape == (1, 36) and A.shape == (1, 5, 4, 4) and verts_orig.shape == (1, 120, 3)
    g = np.random.default_rng(9).standard_normal((120, 3)).astype(np.float32)
    o, o0 = oracle_run(f, upstream=g), oracle_run(f, deltas=False)
    assert np.abs(verts[0].detach().cpu().numpy() - o["verts"]).max() <= 1e-6
    assert np.abs(verts_orig[0].cpu().numpy() - o0["verts"]).max() <= 1e-6
    ((verts[0] - verts_orig[0]) * torch.from_numpy(g).to(cuda_device)).sum().backward()  # mesh-loss style use of both
    for k in DELTAS:
        ref = o["grads"][k]
        assert np.abs(leaves[k].grad.cpu().numpy() - ref).max() <= 1e-5 * float(np.abs(ref).max())
    # a different pose is not served from the cache
    v2, _, _ = mod.forward(expression_params=expr, full_pose=pose * 0.5)
    assert np.abs(v2[0].cpu().numpy() - o0["verts"]).max() > 1e-5
    assert lc0 >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["flame_l0_300", "dense_betas_l0_0", "three_joints_odd_sizes"])
def test_expression_and_pose_coefficient_gradients(case, cuda_device):
    """fs_flame_backward_coeffs vs float64 autograd of the oracle: dL/dbetas (through blendshapes + joint regression)
    and dL/dpose (through pose correctives, kinematic chain and Rodrigues)."""
    from fateavatar_b200 import flame

    if case == "flame_l0_300":
        f, l0 = scenes.flame_inputs(seed=12, V=500), 300
    elif case == "dense_betas_l0_0":
        f, l0 = scenes.flame_inputs(seed=13, V=90), 0
        f["betas"][:300] = 0.2 * np.random.default_rng(2).standard_normal(300).astype(np.float32)
    else:
        f, l0 = scenes.flame_inputs(seed=14, V=61, n_shape=5, n_exp=9, J=3), 5
    V = f["v_template"].shape[0]
    g = np.random.default_rng(8).standard_normal((V, 3)).astype(np.float32)
    # oracle
    m64 = _model(f, torch.float64)
    t64 = lambda k: torch.from_numpy(f[k]).double()
    b64, p64 = t64("betas").requires_grad_(True), t64("pose").requires_grad_(True)
    v64, _, _ = fo.forward_with_delta_blendshape(m64, b64, p64, t64("delta_shapedirs"), t64("delta_posedirs"), t64("delta_vertex"))
    (v64 * torch.from_numpy(g).double()).sum().backward()
    want_b = b64.grad.numpy().copy()
    want_b[:l0] = 0.0  # coefficients below l0 are declared constant
    # kernels, through autograd
    m = _model(f, torch.float32, cuda_device)
    m["parents"] = [int(x) for x in f["parents"]]
    t = lambda k: torch.from_numpy(f[k]).to(cuda_device)
    betas, pose = t("betas")[None].requires_grad_(True), t("pose")[None].requires_grad_(True)
    dv = t("delta_vertex").requires_grad_(True)
    outs = flame.flame_lbs(m, betas, pose, t("delta_shapedirs"), t("delta_posedirs"), dv, l0=l0)
    (outs[0][0] * torch.from_numpy(g).to(cuda_device)).sum().backward()
    got_b, got_p = betas.grad[0].cpu().numpy(), pose.grad[0].cpu().numpy()
    assert betas.grad.shape == betas.shape and pose.grad.shape == pose.shape and dv.grad is not None
    assert np.abs(got_b - want_b).max() <= 2e-5 * np.abs(want_b).max(), (np.abs(got_b - want_b).max(), np.abs(want_b).max())
    assert np.abs(got_p - p64.grad.numpy()).max() <= 2e-5 * np.abs(p64.grad.numpy()).max()


@pytest.mark.gpu
def test_flame_loss_gradient_reaches_the_coefficients_through_verts_orig(cuda_device):
    """train/loss.py:197-201: flame_loss = mean((verts - verts_orig)^2).  With per-frame tracking optimisation the
    expression / pose coefficients require grad and upstream's flame.forward is differentiable w.r.t. them, so the
    -d verts_orig / d(coefficients) term must be present: against float64 autograd of the oracle."""
    from fateavatar_b200 import flame

    f, l0 = scenes.flame_inputs(seed=21, V=300), 300
    m64 = _model(f, torch.float64)
    t64 = lambda k: torch.from_numpy(f[k]).double()
    b64, p64 = t64("betas").requires_grad_(True), t64("pose").requires_grad_(True)
    v64, _, _ = fo.forward_with_delta_blendshape(m64, b64, p64, t64("delta_shapedirs"), t64("delta_posedirs"), t64("delta_vertex"))
    vo64, _, _ = fo.forward_with_delta_blendshape(m64, b64, p64)
    ((v64 - vo64) ** 2).mean().backward()
    want_b = b64.grad.numpy().copy()
    want_b[:l0] = 0.0
    m = _model(f, torch.float32, cuda_device)
    m["parents"] = [int(x) for x in f["parents"]]
    t = lambda k: torch.from_numpy(f[k]).to(cuda_device)
    betas, pose = t("betas")[None].requires_grad_(True), t("pose")[None].requires_grad_(True)
    outs = flame.flame_lbs(m, betas, pose, t("delta_shapedirs"), t("delta_posedirs"), t("delta_vertex"), l0=l0, want_orig=True)
    ((outs[0] - outs[3]) ** 2).mean().backward()
    got_b, got_p = betas.grad[0].cpu().numpy(), pose.grad[0].cpu().numpy()
    assert np.abs(got_b - want_b).max() <= 1e-4 * np.abs(want_b).max(), (np.abs(got_b - want_b).max(), np.abs(want_b).max())
    assert np.abs(got_p - p64.grad.numpy()).max() <= 1e-4 * np.abs(p64.grad.numpy()).max()
    # and through the attached reference-style methods (cache hit path included)
    mod = types.SimpleNamespace(n_shape=300, n_exp=100, parents=torch.from_numpy(f["parents"]),
                                **{k: t(k) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")})
    flame.attach(mod)
    expr = t("betas")[None, 300:].clone().requires_grad_(True)
    pose2 = t("pose")[None].clone().requires_grad_(True)
    v, _, _ = mod.forward_with_delta_blendshape(expr, pose2, t("delta_shapedirs"), t("delta_posedirs"), t("delta_vertex"))
    vo, _, _ = mod.forward(expr, pose2)
    ((v - vo) ** 2).mean().backward()
    assert np.abs(expr.grad[0].cpu().numpy() - want_b[300:]).max() <= 1e-4 * np.abs(want_b).max()
    assert np.abs(pose2.grad[0].cpu().numpy() - p64.grad.numpy()).max() <= 1e-4 * np.abs(p64.grad.numpy()).max()


@pytest.mark.gpu
def test_expand_factors_kernel_and_single_rank_identity(cuda_device):
    """fs_flame_expand_grads: (a) N = 3 random records vs the torch statement; (b) with N = 1 and this rank's own
    factors it reproduces the dense gradients fs_flame_backward writes (the rank-1 structure is exact)."""
    from fateavatar_b200 import flame

    V, L, NP = 333, 400, 36
    g = torch.Generator().manual_seed(1)
    rec = flame.factor_record_floats(V, L, NP)
    gathered = torch.randn(3, rec, generator=g).to(cuda_device)
    gathered[:, :300] = 0.0
    got = flame.expand_factors(gathered, V, L, NP, l0=300, scale=0.5)
    want = flame.expand_factors_reference(gathered, V, L, NP, scale=0.5)
    for a, b in zip(got, want):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    # scalar path: sizes that are not multiples of 4
    gathered = torch.randn(2, flame.factor_record_floats(50, 18, 18), generator=g).to(cuda_device)
    got = flame.expand_factors(gathered, 50, 18, 18, l0=0)
    want = flame.expand_factors_reference(gathered, 50, 18, 18)
    for a, b in zip(got, want):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())

    f = scenes.flame_inputs(seed=11, V=200)
    m = _model(f, torch.float32, cuda_device)
    m["parents"] = [int(x) for x in f["parents"]]
    t = lambda k: torch.from_numpy(f[k]).to(cuda_device)
    r = flame.flame_forward_raw(t("betas"), t("pose"), m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"],
                                m["parents"], m["lbs_weights"], t("delta_vertex"), t("delta_shapedirs"), t("delta_posedirs"),
                                l0=300)
    up = torch.randn(200, 3, generator=g).to(cuda_device)
    dv, ds, dp, gs, gp = flame.flame_backward_raw(t("betas"), m["J_regressor"], m["parents"], m["lbs_weights"], r["workspace"],
                                                  up, (200, 400), l0=300, factors=True)
    record = torch.zeros(1, flame.factor_record_floats(200, 400, 36), device=cuda_device)
    flame.pack_factors(record[0], t("betas"), r["pose_feature"], gs, gp)
    ev, es, ep = flame.expand_factors(record, 200, 400, 36, l0=300)
    assert torch.equal(ev, dv) and torch.equal(es, ds) and torch.equal(ep, dp)
    # the same record produced in place by one fs_flame_backward call
    rec2 = torch.zeros_like(record)
    flame.flame_backward_raw(t("betas"), m["J_regressor"], m["parents"], m["lbs_weights"], r["workspace"], up, (200, 400),
                             l0=300, want=(False, False, False), record=rec2[0])
    assert torch.equal(rec2, record)


@pytest.mark.skipif(not os.path.exists(REF_LBS), reason="reference tree not mounted")
def test_reference_FLAME_methods_run_on_cpu_and_match_the_oracle(monkeypatch):
    """flame/FLAME.py's own `forward_with_delta_blendshape` and `forward` (the two calls of model/fateavatar.py:211-222),
    executed unchanged on an instance whose buffers are the synthetic model (the constructor needs the licensed FLAME
    pickle, so it is bypassed), against the oracle's restatement of those methods."""
    import sys

    pkg = types.ModuleType("flame")
    pkg.__path__ = ["/root/reference/flame"]
    monkeypatch.setitem(sys.modules, "flame", pkg)
    for name in ("lbs", "FLAME"):
        spec = importlib.util.spec_from_file_location(f"flame.{name}", f"/root/reference/flame/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        monkeypatch.setitem(sys.modules, f"flame.{name}", mod)
        spec.loader.exec_module(mod)
    FLAME = sys.modules["flame.FLAME"].FLAME
    f = scenes.flame_inputs(seed=17, V=70)
    obj = FLAME.__new__(FLAME)
    torch.nn.Module.__init__(obj)
    obj.dtype, obj.n_shape, obj.n_exp = torch.float64, f["n_shape"], f["n_exp"]
    t = lambda k: torch.from_numpy(f[k]).double()
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
        obj.register_buffer(k, t(k))
    obj.register_buffer("parents", torch.from_numpy(f["parents"]))
    expr, pose = t("betas")[None, 300:], t("pose")[None]
    deltas = {k: t(k).requires_grad_(True) for k in DELTAS}
    v_ref, pf_ref, A_ref = obj.forward_with_delta_blendshape(expr, pose, deltas["delta_shapedirs"], deltas["delta_posedirs"],
                                                             deltas["delta_vertex"])
    vo_ref, _, Ao_ref = FLAME.forward(obj, expr, pose)
    g = np.random.default_rng(3).standard_normal((70, 3))
    (v_ref[0] * torch.from_numpy(g)).sum().backward()
    o, o0 = oracle_run(f, upstream=g), oracle_run(f, deltas=False)
    assert v_ref.shape == (1, 70, 3) and pf_ref.shape == (1, 36) and A_ref.shape == (1, 5, 4, 4)
    np.testing.assert_allclose(o["verts"], v_ref[0].detach().numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(o["A"], A_ref[0].detach().numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(o0["verts"], vo_ref[0].numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(o0["A"], Ao_ref[0].numpy(), rtol=0, atol=1e-13)
    for k in DELTAS:
        np.testing.assert_allclose(o["grads"][k], deltas[k].grad.numpy(), rtol=0, atol=1e-12 * max(1.0, float(deltas[k].grad.abs().max())))
