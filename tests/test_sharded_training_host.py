"""Frame-sharded training, end to end on the CPU (world size 2, gloo): sampler -> per-rank frame (FLAME, splat
placement, rasterizer: the ORACLES under autograd) -> gradient exchange (flat bucket all-reduce + rank-1 factor records
expanded locally) -> Adam -> densification with rank-identical draws.  Checks what SURVEY section 4 layer (4) asks:
replicas stay identical, and equal a single process that sums the same frames' gradients."""
import os
import socket
import types

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fateavatar_b200 import flame, parallel, scenes
from oracle import flame_oracle as fo
from oracle import oracle as orc
from oracle import pose_oracle as po

N_FRAMES, STEPS, N0, RES = 6, 3, 150, 32
DELTAS = ("delta_vertex", "delta_shapedirs", "delta_posedirs")


class OracleRaster(torch.autograd.Function):
    """The C oracle rasterizer as an autograd op (tests only)."""

    @staticmethod
    def forward(ctx, xyz, scales, rots, opac, shs, cam, bg):
        st = orc.forward(xyz.detach().float().numpy(), opac.detach().float().numpy(), bg, cam["viewmatrix"], cam["projmatrix"],
                         cam["campos"], cam["tanfovx"], cam["tanfovy"], cam["H"], cam["W"], shs=shs.detach().float().numpy(),
                         sh_degree=0, scales=scales.detach().float().numpy(), rotations=rots.detach().float().numpy())
        ctx.st = st
        return torch.from_numpy(st["color"]).double()

    @staticmethod
    def backward(ctx, g):
        og = orc.backward(ctx.st, g.float().numpy())
        t = lambda k: torch.from_numpy(og[k]).double()
        return t("dL_dmeans3D"), t("dL_dscales"), t("dL_drotations"), t("dL_dopacity"), t("dL_dsh"), None, None


def build():
    verts, faces = scenes.ellipsoid_mesh(n_lat=5, n_lon=8)
    f = scenes.flame_inputs(seed=1, V=verts.shape[0], n_shape=4, n_exp=6, J=3)
    f["v_template"] = (verts - verts.mean(0)).astype(np.float32)
    g = torch.Generator().manual_seed(0)
    d = lambda a: torch.from_numpy(np.asarray(a)).double()
    fm = {k: d(f[k]) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights")}
    fm["parents"] = torch.from_numpy(f["parents"])
    P = lambda t: torch.nn.Parameter(t.double())
    m = types.SimpleNamespace(
        fm=fm, faces=torch.from_numpy(faces), face_index=torch.randint(0, faces.shape[0], (N0,), generator=g),
        bary_coords=(lambda u: u / u.sum(-1, keepdim=True))(torch.rand(N0, 3, generator=g, dtype=torch.float64) + 0.05),
        _scaling=P(torch.full((N0, 3), float(np.log(0.02)))), _rotation=P(torch.tensor([[1.0, 0, 0, 0]]).repeat(N0, 1)),
        _offset=P(torch.zeros(N0, 1)), _opacity=P(torch.randn(N0, 1, generator=g)), _features_dc=P(torch.randn(N0, 1, 3, generator=g)),
        delta_vertex=P(d(f["delta_vertex"])), delta_shapedirs=P(d(f["delta_shapedirs"])), delta_posedirs=P(d(f["delta_posedirs"])),
        xyz_gradient_accum=torch.zeros(N0, 1), denom=torch.zeros(N0, 1), max_radii2D=torch.zeros(N0), sample_flag=torch.zeros(N0),
        num_points=N0)
    _, m.canon = po.compute_face_orientation(fm["v_template"], m.faces)
    gs_opt = torch.optim.Adam([{"params": [getattr(m, a)], "name": n, "lr": 1e-2} for n, a in parallel._ATTR_OF_GROUP.items()])
    fl_opt = torch.optim.Adam([getattr(m, k) for k in DELTAS], lr=1e-4)
    frames = [scenes.flame_inputs(seed=100 + i, V=8, n_shape=4, n_exp=6, J=3, with_deltas=False) for i in range(N_FRAMES)]
    cam = scenes.make_camera(RES, RES, 0.35, 0.35, T=[0, 0, 1.0])
    return m, gs_opt, fl_opt, frames, cam


def frame_grads(m, fr, cam):
    """One frame forward + backward; leaves the gradients in .grad and returns the rank-1 factors of the delta grads."""
    for a in list(parallel._ATTR_OF_GROUP.values()) + list(DELTAS):
        getattr(m, a).grad = None
    betas, pose = torch.from_numpy(fr["betas"]).double(), torch.from_numpy(fr["pose"]).double()
    verts, pf, _ = fo.forward_with_delta_blendshape(m.fm, betas, pose, m.delta_shapedirs, m.delta_posedirs, m.delta_vertex)
    xyz, sc, ro, op = po.pose_splats(verts, m.faces, m.face_index, m.bary_coords, m.canon, m._scaling, m._rotation, m._offset,
                                     m._opacity, shell_len=0.02)
    img = OracleRaster.apply(xyz, sc, ro, op, m._features_dc, cam, np.ones(3, np.float32))
    target = torch.linspace(0, 1, 3 * RES * RES, dtype=torch.float64).view(3, RES, RES)
    ((img - target) ** 2).mean().backward()
    i = int(pf.abs().argmax())
    return dict(betas=betas, pose_feature=pf.detach(), dL_dv_shaped=m.delta_vertex.grad.clone(),
                dL_dv_posed=m.delta_posedirs.grad[i] / pf[i].detach())


def train(rank, world, queue=None):
    m, gs_opt, fl_opt, frames, cam = build()
    V, L, NP = m.fm["v_template"].shape[0], 10, 18
    sampler = parallel.FrameShardSampler(N_FRAMES, rank, world, seed=5)
    order = list(sampler)
    splat_attrs = list(parallel._ATTR_OF_GROUP.values())
    for step in range(STEPS):
        mine = [order[step]] if world > 1 else [parallel.FrameShardSampler(N_FRAMES, r, 2, seed=5).order()[r::2][step].item() for r in range(2)]
        splat_sum, records = None, []
        for fi in mine:  # one frame per rank; the single-process reference walks both ranks' frames
            fac = frame_grads(m, frames[fi], cam)
            flat = torch.cat([getattr(m, a).grad.reshape(-1) for a in splat_attrs])
            splat_sum = flat if splat_sum is None else splat_sum + flat
            rec = torch.zeros(flame.factor_record_floats(V, L, NP), dtype=torch.float64)
            records.append(flame.pack_factors(rec, fac["betas"], fac["pose_feature"], fac["dL_dv_shaped"], fac["dL_dv_posed"]))
        if world > 1:
            dist.all_reduce(splat_sum)
            gathered = torch.empty(world, records[0].numel(), dtype=torch.float64)
            dist.all_gather_into_tensor(gathered.view(-1), records[0])
        else:
            gathered = torch.stack(records)
        dv, ds, dp = flame.expand_factors_reference(gathered, V, L, NP)
        off = 0
        for a in splat_attrs:
            p = getattr(m, a)
            p.grad = splat_sum[off:off + p.numel()].view_as(p).clone()
            off += p.numel()
        m.delta_vertex.grad, m.delta_shapedirs.grad, m.delta_posedirs.grad = dv, ds, dp
        gs_opt.step()
        fl_opt.step()
        # stand-in for the per-frame densification statistics: every "rank" contributes its own increment
        for r in ([rank] if world > 1 else [0, 1]):
            m.xyz_gradient_accum += torch.rand(m.num_points, 1, generator=torch.Generator().manual_seed(1000 * r + step))
        if step == 1:  # densify from statistics summed over ranks, with rank-identical draws
            if world > 1:
                parallel.allreduce_densify_stats(m)
            parallel.uv_densify(m, gs_opt, 20, generator=parallel.synced_generator("cpu", 3, step))
    out = {a: getattr(m, a).detach().clone() for a in splat_attrs + list(DELTAS)}
    out["face_index"], out["bary"] = m.face_index.clone(), m.bary_coords.clone()
    if queue is not None:  # numpy: tensors would travel as shared-memory handles that die with the worker
        queue.put((rank, {k: v.numpy().copy() for k, v in out.items()}))
    return out


def _worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    train(rank, world, queue)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_keeps_replicas_identical_and_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), f"replicas diverged in {k}"   # bitwise: same sums in the same order
    assert res[0]["_opacity"].shape[0] == N0 + 20
    single = train(0, 1)
    for k in res[0]:
        a, b = res[0][k].astype(np.float64), single[k].double().numpy()
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-6 * max(np.abs(b).max(), 1.0), k
