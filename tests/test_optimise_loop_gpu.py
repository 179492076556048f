"""Config 3 on the GPU path: the fused Adam, the in-place densify / prune / opacity reset and the
recorded optimise loop against the reference's own torch code -- torch.optim.Adam and the UNCHANGED FateAvatar._uv_densify
/ _prune_low_opacity_points / _reset_opacity of the staged model/fateavatar.py (oracle/_ref/pyref) -- under a shared
generator seed."""
import os
import types

import numpy as np
import pytest
import torch

import ref_frame_harness as H
from fateavatar_b200 import avatar, optimizer as fopt, scenes

pytestmark = pytest.mark.gpu
HAVE_REF = os.path.isdir(H.STAGED)
LRS = dict(fopt.OptimiseLoop.DEFAULTS)


def test_fused_l1_image_loss_matches_torch(cuda_device):
    from fateavatar_b200 import losses

    g = torch.Generator(device=cuda_device).manual_seed(3)
    for shape in ((3, 512, 512), (1, 3, 33, 47)):
        x = torch.rand(shape, device=cuda_device, generator=g).requires_grad_(True)
        t = torch.rand(shape, device=cuda_device, generator=g)
        t.view(-1)[:5] = x.detach().view(-1)[:5]  # exact ties: sign(0) = 0 on both sides
        (losses.l1_image_loss(x, t) * 1.7).backward()
        got_g, x.grad = x.grad.clone(), None
        ref = (x - t).abs().mean()
        (ref * 1.7).backward()
        assert abs(float(losses.l1_image_loss(x, t)) - float(ref)) <= 1e-6 * float(ref)
        assert float((got_g - x.grad).abs().max()) <= 1e-7 * float(x.grad.abs().max())
        a, b = losses.l1_image_loss(x, t), losses.l1_image_loss(x, t)
        assert float(a) == float(b)  # deterministic


def test_fused_adam_matches_torch_adam(cuda_device):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(0)
    shapes, lrs = [(1000, 1), (1000, 3), (777, 4), (5, 3, 40)], [0.05, 0.0025, 0.001, 1e-5]
    ref = [torch.nn.Parameter(torch.randn(s, device=dev, generator=g)) for s in shapes]
    mine = [p.detach().clone() for p in ref]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref, lrs)], lr=0.0)
    grads = [torch.zeros(s, device=dev) for s in shapes]
    adam = fopt.FusedAdam([dict(name=str(k), lr=lr, param=p.view(-1), grad=gr.view(-1), m=torch.zeros(p.numel(), device=dev),
                                v=torch.zeros(p.numel(), device=dev)) for k, (p, gr, lr) in enumerate(zip(mine, grads, lrs))], dev)
    for step in range(7):
        for p, gr in zip(ref, grads):
            gr.copy_(torch.randn(gr.shape, device=dev, generator=g) * 10.0 ** float(step - 3))
            p.grad = gr.clone()
        opt.step()
        adam.step()
    assert adam.steps[:4].tolist() == [7, 7, 7, 7]
    for k, (p, q) in enumerate(zip(ref, mine)):
        st = opt.state[p]
        assert float((p.detach() - q).abs().max()) <= 2e-6 * max(1.0, float(p.detach().abs().max())), k
        # (fp32, tolerance relative to each tensor's largest entry: nvcc contracts lerp / addcmul into FMAs its own way)
        for key, mine_t in (("exp_avg", adam.groups[k]["m"]), ("exp_avg_sq", adam.groups[k]["v"])):
            want = st[key].view(-1)
            assert float((want - mine_t).abs().max()) <= 2e-6 * float(want.abs().max()), (k, key)


def _ref_model_and_optim(a, dev, res):
    """A reference FateAvatar instance (staged, unchanged class) + the optimizer groups of train/optim.py:15-35."""
    import fateavatar_b200

    fateavatar_b200.install()
    import diff_gaussian_rasterization as dgr
    from simple_knn._C import distCUDA2

    patch = H.Patch()
    FateAvatar, FLAME, mesh_compute = H.load_reference(patch, root=H.STAGED, rasterizer=dgr, knn=distCUDA2)
    m = H.build_reference_model(FateAvatar, FLAME, mesh_compute, a, res, device=dev)
    P = m._scaling.shape[0]
    m.xyz_gradient_accum, m.denom = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)
    m.max_radii2D, m.sample_flag, m.num_points = torch.zeros(P, device=dev), torch.zeros(P, device=dev), P
    gs = torch.optim.Adam([{"params": [m._opacity], "lr": LRS["opacity_lr"], "name": "opacity"},
                           {"params": [m._offset], "lr": LRS["offset_lr"], "name": "offset"},
                           {"params": [m._features_dc], "lr": LRS["feature_dc_lr"], "name": "color"},
                           {"params": [m._rotation], "lr": LRS["rotation_lr"], "name": "rotation"},
                           {"params": [m._scaling], "lr": LRS["scaling_lr"], "name": "scaling"}], lr=0.0)
    bs = torch.optim.Adam([{"params": [m.delta_shapedirs], "lr": LRS["delta_shapedirs_lr"], "name": "delta_shapedirs"},
                           {"params": [m.delta_posedirs], "lr": LRS["delta_posedirs_lr"], "name": "delta_posedirs"},
                           {"params": [m.delta_vertex], "lr": 0.0001, "name": "delta_vertex"}], lr=0.0)
    return m, {"gs": gs, "bs": bs}, patch


def _my_model(a, dev, res):
    d = lambda x: torch.from_numpy(np.asarray(x)).to(dev)
    fl = types.SimpleNamespace(n_shape=a["n_shape"], n_exp=a["n_exp"], parents=torch.from_numpy(a["parents"]))
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
        setattr(fl, k, d(a[k]))
    from oracle import pose_oracle as po

    _, canon = po.compute_face_orientation(d(a["v_template"])[None], d(a["faces"]))
    par = lambda x: torch.nn.Parameter(d(x))
    P = a["face_index"].shape[0]
    return types.SimpleNamespace(
        flame=fl, faces=d(a["faces"]), face_index=d(a["face_index"]), bary_coords=d(a["bary"]), face_scaling_canonical=canon,
        _scaling=par(a["scaling_raw"]), _rotation=par(a["rotation_raw"]), _offset=par(a["offset_raw"]), _opacity=par(a["opacity_raw"]),
        _features_dc=par(a["features_dc"]), delta_shapedirs=par(a["delta_shapedirs"]), delta_posedirs=par(a["delta_posedirs"]),
        delta_vertex=par(a["delta_vertex"]), cfg_model=types.SimpleNamespace(delta_blendshape=True, delta_vertex=True, resize_scale=True),
        shell_len=a["shell_len"], bg_color=torch.ones(3, device=dev), img_res=res,
        xyz_gradient_accum=torch.zeros(P, 1, device=dev), denom=torch.zeros(P, 1, device=dev),
        max_radii2D=torch.zeros(P, device=dev), sample_flag=torch.zeros(P, device=dev), num_points=P)


def _assert_same_splat_set(ref, store, opt=None, exact=True, tol=0.0):
    P = store.P
    assert ref.num_points == P == ref._scaling.shape[0]
    assert torch.equal(ref.face_index, store.view("face_index")) and torch.equal(ref.bary_coords, store.view("bary"))
    # (the kernel's norm is torch.norm's rounding sequence, tools/norm_check.py; allclose kept as the assertion)
    assert torch.allclose(ref.xyz_gradient_accum, store.view("accum"), rtol=1e-6, atol=1e-12)
    assert torch.equal(ref.denom, store.view("denom"))
    assert torch.equal(ref.sample_flag, store.view("sample_flag"))
    for n, attr, w in fopt.FIELDS:
        a, b = getattr(ref, attr).detach().reshape(P, w), store.view(n)
        if exact and n != "scaling":
            assert torch.equal(a, b), n
        else:
            assert float((a - b).abs().max()) <= max(tol, 1e-6), n
        if opt is not None:
            st = opt.state.get(getattr(ref, attr), None)
            if st is not None:
                assert float((st["exp_avg"].reshape(P, w) - store.view("m_" + n)).abs().max()) <= tol + 1e-12, n
                assert float((st["exp_avg_sq"].reshape(P, w) - store.view("v_" + n)).abs().max()) <= tol + 1e-12, n


@pytest.mark.skipif(not HAVE_REF, reason="staged reference callers missing (oracle/stage_ref_py.py)")
def test_densify_prune_reset_match_the_unchanged_reference_methods(cuda_device):
    dev, res = cuda_device, (96, 96)
    a = scenes.small_avatar(seed=3, N=3000)
    ref, opts, patch = _ref_model_and_optim(a, dev, res)
    try:
        mine = _my_model(a, dev, res)
        store = fopt.SplatStore(mine, capacity=8000)
        g = torch.Generator(device=dev).manual_seed(5)
        # one Adam step on both sides so that moments exist (upstream only patches existing optimizer state)
        adam = fopt.fateavatar_adam(mine, store, LRS)
        for n, attr, w in fopt.FIELDS:
            gr = torch.randn(getattr(mine, attr).shape, device=dev, generator=g)
            getattr(mine, attr).grad, getattr(ref, attr).grad = gr.clone(), gr.clone()
        for n in ("delta_shapedirs", "delta_posedirs", "delta_vertex"):
            gr = torch.randn(getattr(mine, n).shape, device=dev, generator=g)
            getattr(mine, n).grad, getattr(ref, n).grad = gr.clone(), gr.clone()
        opts["gs"].step(), opts["bs"].step(), adam.step()
        stats = torch.rand(store.P, 1, device=dev, generator=g)
        ref.xyz_gradient_accum.copy_(stats), store.view("accum").copy_(stats)
        _assert_same_splat_set(ref, store, opts["gs"], exact=False, tol=1e-6)
        # _uv_densify (model/fateavatar.py:610-672): upstream draws from the global CUDA generator
        torch.cuda.manual_seed(1234)
        ref._uv_densify(opts["gs"], increase_num=500)
        store.uv_densify(500, generator=torch.Generator(device=dev).manual_seed(1234))
        assert store.P == 3500 and mine.num_points == 3500 and mine._scaling.shape == (3500, 3)
        _assert_same_splat_set(ref, store, opts["gs"], exact=False, tol=1e-6)
        # _prune_low_opacity_points (:674-713) with a threshold that removes a good part of the set
        ref._prune_low_opacity_points(opts["gs"], min_opacity=0.4)
        newP = store.prune_low_opacity(0.4)
        assert 500 < newP < 3300
        _assert_same_splat_set(ref, store, opts["gs"], exact=False, tol=1e-6)
        # _reset_opacity (:715-732)
        ref._reset_opacity(opts["gs"])
        store.reset_opacity()
        _assert_same_splat_set(ref, store, opts["gs"], exact=False, tol=1e-6)
        assert float(store.view("m_opacity").abs().max()) == 0.0 and float(torch.sigmoid(store.view("opacity")).max()) <= 0.0100001
    finally:
        patch.undo()


def _frames(a, res, n):
    frames = []
    for k in range(n):
        fa = scenes.flame_inputs(seed=100 + k, V=8, with_deltas=False)
        inp = H.frame_input(dict(a, betas=fa["betas"], pose=fa["pose"]), fovx=0.35, fovy=0.35, T=(0.0, 0.0, 1.25))
        tgt = torch.rand(3, *res, generator=torch.Generator().manual_seed(k))
        frames.append(dict(cam_pose=inp["cam_pose"], flame_pose=inp["flame_pose"], expression=inp["expression"], target=tgt))
    return [{k: v.contiguous().pin_memory() for k, v in f.items()} for f in frames]


FOV = [0.35]
LOOP_CFG = dict(densify_interval=4, prune_interval=6, opacity_reset_interval=13, increase_num=300, max_points_num=3200,
                min_opacity=0.2)


def _frame_loss(m, d):
    out = avatar.forward_frame(m, dict(cam_pose=d["cam_pose"], fovx=FOV, fovy=FOV, flame_pose=d["flame_pose"],
                                       expression=d["expression"]))
    sc = out["scale"]
    scale_regu = torch.relu(sc.max(dim=-1)[0] / sc.min(dim=-1)[0] - 9.0).mean()     # train/loss.py:143-149
    return (out["rgb_image"][0] - d["target"]).abs().mean() + 0.1 * scale_regu, out


@pytest.mark.skipif(not HAVE_REF, reason="staged reference callers missing (oracle/stage_ref_py.py)")
def test_optimise_loop_follows_the_reference_iteration_step(cuda_device):
    """iteration_step_fateavatar (train/iteration.py:33-85) restated with the reference's own model methods and
    torch.optim.Adam, against OptimiseLoop.  The reference side consumes the SAME per-frame gradients (its own backward
    sums float atomics in a run-dependent order, and Adam's g / sqrt(v) turns a last-bit difference of a vanishing
    gradient into a full learning-rate step, so two runs of upstream itself do not stay on one trajectory): everything
    after the gradient -- statistics, both Adam steps, densify / prune / reset and their order -- must then agree."""
    dev, res = cuda_device, (96, 96)
    a = scenes.small_avatar(seed=4, N=2500)
    cfg = LOOP_CFG
    ref, opts, patch = _ref_model_and_optim(a, dev, res)
    try:
        mine = _my_model(a, dev, res)
        host = _frames(a, res, 14)
        seen = {}

        def on_frame(loss, out, g):
            seen.update(loss=float(loss), g={k: v.detach().clone() for k, v in g.items()},
                        vs_grad=out["viewspace_points"][0].grad.detach().clone(), vis=out["visibility_filter"][0].clone())

        loop = fopt.OptimiseLoop(mine, _frame_loss, {k: v.to(dev) for k, v in host[0].items()}, training=cfg,
                                 generator=torch.Generator(device=dev).manual_seed(77), capture=False, on_frame=on_frame)
        torch.cuda.manual_seed(77)  # upstream's _uv_densify draws from the global generator
        sizes = []
        for t, h in enumerate(host):
            loop.step(h)
            loop.wait()
            # ---- reference side: train/iteration.py:47-85 on the same gradients ----
            g = seen["g"]
            for n, attr, w in fopt.FIELDS:
                getattr(ref, attr).grad = g[{"color": "features_dc"}.get(n, n)].reshape(getattr(ref, attr).shape).clone()
            for n in ("delta_vertex", "delta_shapedirs", "delta_posedirs"):
                getattr(ref, n).grad = g[n].reshape(getattr(ref, n).shape).clone()
            vs = types.SimpleNamespace(grad=seen["vs_grad"])
            ref._add_densification_stats(vs, seen["vis"])
            for o in opts.values():
                o.step()
            if t % cfg["densify_interval"] == 0 and ref.num_points < cfg["max_points_num"]:
                ref._uv_densify(opts["gs"], increase_num=min(cfg["max_points_num"] - ref.num_points, cfg["increase_num"]))
            if t % cfg["prune_interval"] == 0:
                ref._prune_low_opacity_points(opts["gs"], min_opacity=cfg["min_opacity"])
            if t % cfg["opacity_reset_interval"] == 0 and t != 0:
                ref._reset_opacity(opts["gs"])
            sizes.append(loop.store.P)
            assert ref.num_points == loop.store.P, (t, ref.num_points, sizes)
            assert torch.equal(ref.face_index, loop.store.view("face_index")), t
            assert torch.equal(ref.bary_coords, loop.store.view("bary")), t
            _assert_same_splat_set(ref, loop.store, opts["gs"], exact=False, tol=2e-5)
        assert len(set(sizes)) >= 4  # the set grew and shrank during the run
        for n in ("delta_vertex", "delta_shapedirs", "delta_posedirs"):
            x, y = getattr(ref, n).detach(), getattr(mine, n).detach()
            assert float((x - y).abs().max()) <= 2e-5 * max(float(x.abs().max()), 1e-6), n
    finally:
        patch.undo()


def test_recorded_optimise_loop_tracks_the_eager_one(cuda_device):
    """The same loop replayed from CUDA graphs (re-recorded whenever densify / prune change P) against the eager one:
    same splat counts within a handful (float atomics, see above), same loss curve."""
    dev, res = cuda_device, (96, 96)
    a = scenes.small_avatar(seed=4, N=2500)
    host = _frames(a, res, 14)
    runs = {}
    for graph in (False, True):
        mine = _my_model(a, dev, res)
        loop = fopt.OptimiseLoop(mine, _frame_loss, {k: v.to(dev) for k, v in host[0].items()}, training=LOOP_CFG,
                                 generator=torch.Generator(device=dev).manual_seed(77), capture=graph)
        losses, sizes = [], []
        for h in host:
            r = loop.step(h)
            loop.wait()
            losses.append(float(r["loss"][0]))
            sizes.append(loop.store.P)
        runs[graph] = (losses, sizes, loop)
    (l0, s0, _), (l1, s1, lg) = runs[False], runs[True]
    assert lg.recaptures >= 4 and len(set(s1)) >= 4
    assert max(abs(x - y) for x, y in zip(s0, s1)) <= 0.01 * max(s0)
    assert max(abs(x - y) for x, y in zip(l0, l1)) <= 2e-3 * max(l0)
    assert l1[-1] < l1[0]  # and it optimises


def test_checkpoint_round_trip_through_the_splat_store(cuda_device, tmp_path):
    """io.model_state / io.restore_splats(store=) (train/trainer.py:396-435, train/deserialize.py:7-40): a checkpoint taken
    after a densification is compact (P rows, not the store's capacity) and refills a fresh store of another size in
    place; Adam moments and densification statistics restart, as upstream."""
    from fateavatar_b200 import io as fio

    dev, res = cuda_device, (96, 96)

    class Holder(torch.nn.Module):  # a state_dict()-able view of the splat attributes
        def __init__(self, m):
            super().__init__()
            for k in ("_offset", "_features_dc", "_scaling", "_rotation", "_opacity"):
                setattr(self, k, getattr(m, k))
            self._features_rest = torch.nn.Parameter(torch.zeros(m._offset.shape[0], 0, 3, device=dev))
            self.register_buffer("face_index", m.face_index)
            self.register_buffer("bary_coords", m.bary_coords)

    m = _my_model(scenes.small_avatar(seed=5, N=2000), dev, res)
    store = fopt.SplatStore(m, capacity=4000)
    store.view("accum").uniform_(0.1, 1.0)
    store.uv_densify(300, generator=torch.Generator(device=dev).manual_seed(1))
    store.bind()
    assert store.P == 2300 and m._scaling.shape[0] == 2300
    state = fio.model_state(Holder(m))
    assert state["_scaling"].shape == (2300, 3) and state["_scaling"].untyped_storage().nbytes() == 2300 * 3 * 4
    path = str(tmp_path / "ck.pth")
    torch.save({"model": state}, path)

    m2 = _my_model(scenes.small_avatar(seed=6, N=1500), dev, res)
    store2 = fopt.SplatStore(m2, capacity=5000)
    store2.view("m_scaling").fill_(1.0)
    base_ptr = store2.arrays()["scaling"].data_ptr()
    h2 = Holder(m2)
    fio.restore_splats(h2, torch.load(path)["model"], store=store2)
    assert store2.P == 2300 and store2.model is m2 and m2.num_points == 2300
    assert store2.arrays()["scaling"].data_ptr() == base_ptr  # in place
    for n, attr, w in fopt.FIELDS:
        assert torch.equal(getattr(m2, attr).detach().reshape(2300, w), getattr(m, attr).detach().reshape(2300, w)), n
        assert not store2.arrays()["m_" + n].any() and not store2.arrays()["v_" + n].any()
    assert torch.equal(m2.face_index, m.face_index) and torch.equal(m2.bary_coords, m.bary_coords)
    assert not m2.xyz_gradient_accum.any() and m2.xyz_gradient_accum.shape == (2300, 1)
