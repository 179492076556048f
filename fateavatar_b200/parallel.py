"""Host-side plumbing for frame-sharded (data-parallel) training of one avatar.

The reference trains with batch size 1 on one GPU (train/trainer.py); sharding the monocular video over ranks needs
three things it never had to provide:

  * FrameShardSampler     rank r takes frames r, r+N, ... of a permutation every rank derives from the same seed, so
                          an epoch visits each frame once across the job and all ranks take the same number of steps
                          (a `sampler=` for the reference's DataLoader, train/dataset.py);
  * allreduce_densify_stats   densification statistics are per-frame sums (model/fateavatar.py:734-737): summed over
                          ranks before `_uv_densify` / prune look at them;
  * synced_generator      `_uv_densify` draws parents with torch.multinomial and barycentrics with torch.rand
                          (model/fateavatar.py:617-621); every rank must draw the SAME splats or the replicas diverge
                          in size, so the draws come from a generator seeded identically on all ranks;
  * uv_densify            `_uv_densify` itself with that generator as an argument.
"""
import torch


class FrameShardSampler(torch.utils.data.Sampler):
    def __init__(self, n_frames, rank, world, seed=0, shuffle=True):
        if not 0 <= rank < world:
            raise ValueError("rank must be in [0, world)")
        self.n, self.rank, self.world, self.seed, self.shuffle, self.epoch = int(n_frames), rank, world, seed, shuffle, 0
        self.per_rank = self.n // world  # frames beyond a multiple of `world` wait for the next epoch's permutation

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def order(self):
        if not self.shuffle:
            return torch.arange(self.n)
        g = torch.Generator()
        g.manual_seed(self.seed * 1000003 + self.epoch)
        return torch.randperm(self.n, generator=g)

    def __iter__(self):
        return iter(self.order()[self.rank::self.world][: self.per_rank].tolist())

    def __len__(self):
        return self.per_rank


def synced_generator(device, seed, step):
    """A torch.Generator in the same state on every rank (function of `seed` and the global step only)."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 2654435761 + int(step)) % (2 ** 63 - 1))
    return g


def allreduce_densify_stats(model, group=None):
    """Sum `xyz_gradient_accum` and `denom` over ranks in place (one collective over a flat copy)."""
    import torch.distributed as dist

    a, d = model.xyz_gradient_accum, model.denom
    flat = torch.cat([a.reshape(-1), d.reshape(-1)])
    dist.all_reduce(flat, group=group)
    a.copy_(flat[: a.numel()].view_as(a))
    d.copy_(flat[a.numel():].view_as(d))
    return model


# ---- densification with rank-identical draws ------------------------------------------------------------------------
_ATTR_OF_GROUP = {"opacity": "_opacity", "offset": "_offset", "color": "_features_dc", "rotation": "_rotation",
                  "scaling": "_scaling"}


def _grow_param_groups(optimizer, new_rows):
    """Append `new_rows[name]` to the single parameter of every named group, growing Adam's moments with zeros, and
    return {name: new Parameter}.  (What model/fateavatar.py:637-660 does inline.)"""
    out = {}
    for group in optimizer.param_groups:
        if len(group["params"]) != 1:
            raise ValueError("every splat parameter group holds exactly one tensor")
        old, rows = group["params"][0], new_rows[group["name"]]
        state = optimizer.state.pop(old, None)
        grown = torch.nn.Parameter(torch.cat((old.detach(), rows), dim=0).requires_grad_(True))
        if state is not None:
            for key in ("exp_avg", "exp_avg_sq"):
                state[key] = torch.cat((state[key], torch.zeros_like(rows)), dim=0)
            optimizer.state[grown] = state
        group["params"][0] = grown
        out[group["name"]] = grown
    return out


def uv_densify(model, gs_optimizer, increase_num=1000, generator=None):
    """`FateAvatar._uv_densify` (model/fateavatar.py:610-670) with the random draws taken from `generator`:
    parents ~ multinomial(xyz_gradient_accum) with replacement, new splats copy the parent's opacity / offset / colour
    / rotation, get 0.75x its scale, sit on the parent's face at fresh random barycentrics; Adam moments grow by
    zeros and the densification statistics restart.  With `synced_generator(...)` (and statistics summed over ranks
    by `allreduce_densify_stats`) every rank adds the same splats, so replicas stay identical."""
    dev = model.face_index.device
    parents = model.xyz_gradient_accum.squeeze(1).multinomial(increase_num, replacement=True, generator=generator)
    new_faces = model.face_index[parents]
    uvw = torch.rand((increase_num, 3), device=dev, generator=generator)
    new_bary = uvw / uvw.sum(dim=-1, keepdim=True)
    rows = {name: getattr(model, attr)[parents].detach() for name, attr in _ATTR_OF_GROUP.items() if name != "scaling"}
    rows["scaling"] = torch.log(torch.exp(model._scaling[parents].detach()) * 0.75)
    grown = _grow_param_groups(gs_optimizer, rows)
    for name, attr in _ATTR_OF_GROUP.items():
        setattr(model, attr, grown[name])
    model.face_index = torch.cat([model.face_index, new_faces], dim=0)
    model.bary_coords = torch.cat([model.bary_coords, new_bary], dim=0)
    if hasattr(model, "sample_flag"):
        model.sample_flag = torch.cat([model.sample_flag, torch.ones(increase_num, device=dev)], dim=0)
    n = model.num_points = model.bary_coords.shape[0]
    model.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
    model.denom = torch.zeros((n, 1), device=dev)
    model.max_radii2D = torch.zeros((n,), device=dev)
    return parents


# ---- the frame-sharded step -----------------------------------------------------------------------------------------
# (upstream has no counterpart: train/base.py:54-60 trains batch 1 on one GPU.)
SPLAT_PARTS = (("features_dc", 3), ("scaling", 3), ("rotation", 4), ("offset", 1), ("opacity", 1),
               ("accum_inc", 1), ("denom_inc", 1))
PARAM_ATTR = {"features_dc": "_features_dc", "scaling": "_scaling", "rotation": "_rotation", "offset": "_offset",
              "opacity": "_opacity"}
DELTA_NAMES = ("delta_vertex", "delta_shapedirs", "delta_posedirs")


class GradLayout:
    """Flat fp32 layout of what one data-parallel step exchanges:

        [ dL/d_features_dc 3P | dL/d_scaling 3P | dL/d_rotation 4P | dL/d_offset P | dL/d_opacity P |
          xyz_gradient_accum increment P | denom increment P ]            <- summed over ranks
        [ world x factor record of the FLAME delta gradients ]            <- gathered (each rank fills its own slot)

    Every part starts on a 16-byte boundary; the kernels write their gradients straight into views of it."""

    def __init__(self, P, V, L, NPF):
        from . import flame as _flame

        self.P, self.V, self.L, self.NPF = int(P), int(V), int(L), int(NPF)
        Pp = (self.P + 3) // 4 * 4
        self.offsets, off = {}, 0
        for name, k in SPLAT_PARTS:
            self.offsets[name] = (off, k)
            off += k * Pp
        self.n_splat = off
        self.rec_floats = _flame.factor_record_floats(self.V, self.L, self.NPF)

    def views(self, flat):
        P = self.P
        return {name: flat[o:o + k * P].view(P, k) for name, (o, k) in self.offsets.items()}


class ShardedStep:
    """One optimisation step of frame-sharded (data-parallel) FateAvatar training: every rank renders ITS frame, then
    ONE fused peer-memory kernel (exchange.StepExchange / fs_p2p_exchange) barriers the ranks, sums the splat gradients
    and densification-statistic increments over NVLink and rebuilds the summed dense FLAME delta gradients from the
    ranks' rank-1 factor records.  With one rank the exchange disappears and the gradients are used where they are.

        step = ShardedStep(model)                       # collective when torch.distributed is initialised
        for s, batch in enumerate(loader):              # FrameShardSampler: rank r sees frames r, r+N, ...
            loss, out = step.run_autograd(s, frame_loss_fn, batch)   # forward_frame + loss + backward + pack
            g = step.exchange(s)                        # {"scaling": [P,3], ..., "accum_inc": [P,1], "delta_vertex": ...}
            step.bind_grads(g); optimizer.step()        # p.grad := summed gradients (no copy)
            step.apply_densify_stats(g)                 # model.xyz_gradient_accum += ..., model.denom += ...

    `capture(...)` records the whole step (frame, pack, exchange) into two CUDA graphs (one per bucket parity) so a
    step is a single graph launch; `AbiFrame` is the same frame as a chain of raw C-ABI calls (no autograd).

    `model` needs FateAvatar's attributes (avatar.forward_frame's list)."""

    def __init__(self, model, group=None, device=None, algo=None, average=False):
        import torch.distributed as dist

        self.model = model
        self.device = torch.device(device) if device is not None else model._scaling.device
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        fl = model.flame
        V, L = int(fl.v_template.shape[0]), int(fl.shapedirs.shape[-1])
        NPF = (len(fl.parents) - 1) * 9
        self.l0 = int(fl.n_shape)
        self.layout = GradLayout(model._scaling.shape[0], V, L, NPF)
        self.scale = 1.0 / self.world if average else 1.0
        self.ex = None
        if self.world > 1:
            from .exchange import StepExchange

            self.ex = StepExchange(self.layout.n_splat, self.layout.rec_floats, self.device, group=group, algo=algo)
            # dense delta gradients, rebuilt by every exchange from the gathered factor records
            self.delta_grads = {n: torch.zeros_like(getattr(model, n)) for n in DELTA_NAMES
                                if getattr(model, n, None) is not None}
        else:
            self.local = torch.zeros(self.layout.n_splat, device=self.device)
            self.delta_grads = {}
        self._captured = None

    # -- where this rank's frame writes ------------------------------------------------------------------------------
    def fill_views(self, step):
        flat = self.ex.fill(step)[:self.layout.n_splat] if self.ex is not None else self.local
        return self.layout.views(flat)

    def record(self, step):
        """This rank's FLAME factor record for `step` (None with a single rank: dense gradients are written directly)."""
        return self.ex.record(step)[:self.layout.rec_floats] if self.ex is not None else None

    # -- the frame under autograd (public operator API) ---------------------------------------------------------------
    def run_autograd(self, step, frame_loss_fn, inputs):
        """frame_loss_fn(model, inputs) -> (loss, out) with out = avatar.forward_frame's dict.  Runs it, calls
        loss.backward(), and leaves this rank's contribution in the step's bucket."""
        import contextlib

        from . import _lib, flame as _flame

        m, rec = self.model, self.record(step)
        with (_flame.factor_record(rec) if rec is not None else contextlib.nullcontext()):
            loss, out = frame_loss_fn(m, inputs)
            loss.backward()
        v, P = self.fill_views(step), self.layout.P
        if self.ex is not None:  # pack: one multi-tensor copy of the five leaf gradients into the bucket
            torch._foreach_copy_([v[n] for n in PARAM_ATTR],
                                 [getattr(m, a).grad.reshape(P, -1) for a in PARAM_ATTR.values()])
        g2d, radii = out["viewspace_points"][0].grad, out["radii"][0]
        with _lib.on_device(self.device):
            rc = _lib.load().fs_densify_stats_inc(P, g2d.data_ptr(), radii.data_ptr(), v["accum_inc"].data_ptr(),
                                                  v["denom_inc"].data_ptr(), _lib.stream_ptr(self.device))
        _lib.check(rc, "fs_densify_stats_inc")
        return loss, out

    # -- the exchange -------------------------------------------------------------------------------------------------
    def exchange(self, step):
        """Sum over ranks.  Returns {part: tensor}: the five leaf gradients, the two statistic increments and the dense
        delta gradients, all rank-local, valid until the next exchange."""
        m, lay = self.model, self.layout
        if self.ex is None:
            g = dict(self.layout.views(self.local))
            for n, a in PARAM_ATTR.items():
                p = getattr(m, a)
                if p.grad is not None:
                    g[n] = p.grad
            for n in DELTA_NAMES:
                p = getattr(m, n, None)
                if p is not None and p.grad is not None:
                    g[n] = p.grad
            return g
        dg = self.delta_grads
        summed = self.ex.exchange(step, (lay.V, lay.L, self.l0, lay.NPF),
                                  tuple(dg.get(n) for n in DELTA_NAMES), scale=self.scale)
        g = dict(lay.views(summed))
        g.update(dg)
        return g

    def bind_grads(self, g):
        """Point .grad of every trained tensor at the summed gradients (attribute assignment, no copy)."""
        m = self.model
        for n, a in PARAM_ATTR.items():
            p = getattr(m, a)
            p.grad = g[n].view(p.shape)
        for n in DELTA_NAMES:
            p = getattr(m, n, None)
            if p is not None and n in g:
                p.grad = g[n].view(p.shape)

    def apply_densify_stats(self, g):
        m = self.model
        m.xyz_gradient_accum += g["accum_inc"]
        m.denom += g["denom_inc"]

    # -- whole step as CUDA graphs ------------------------------------------------------------------------------------
    def capture(self, frame_loss_fn, example_inputs, extra_outputs=None, warmup=3):
        """Record [frame + loss + backward + pack + exchange + read-back of the loss] once per bucket parity.  Afterwards
        `step(s, host_inputs)` is: async H2D of the frame's inputs, ONE graph launch; `wait()` returns the pinned
        outputs ({"loss": [1]} plus `extra_outputs(out)` entries)."""
        from .graph import CapturedStep

        leaves = [getattr(self.model, a) for a in PARAM_ATTR.values()]
        leaves += [getattr(self.model, n) for n in DELTA_NAMES if getattr(self.model, n, None) is not None]
        self._g = [None, None]

        def make(parity):
            def fn(inp):
                loss, out = self.run_autograd(parity, frame_loss_fn, inp)
                self._g[parity] = self.exchange(parity)
                res = {"loss": loss.detach().reshape(1)}
                if extra_outputs is not None:
                    res.update(extra_outputs(out))
                return res
            return fn

        self._captured = [CapturedStep(make(k), example_inputs, params=leaves, warmup=warmup, device=self.device)
                          for k in (0, 1)]
        return self

    def __call__(self, step, host_inputs=None):
        """Replay step `step` on `host_inputs` (copied in first) or, with None, on the inputs `prefetch` staged for it."""
        c = self._captured[step & 1]
        self._last = c
        return c(host_inputs)

    def prefetch(self, step, host_inputs):
        """Stage the pinned host inputs of a FUTURE step (normally step + 1, issued right after the launch of `step`) on a
        side stream, so that the H2D copy overlaps the running step instead of preceding the next one."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        self._captured[step & 1].prefetch(host_inputs, self._copy_stream)

    def wait(self, step=None):
        """Outputs of the last launched step (or of `step`): blocks only until THAT step has finished, so a caller may
        launch step s + 1 first and read step s afterwards -- the GPU then never idles between steps."""
        return (self._last if step is None else self._captured[step & 1]).wait()

    def grads(self, step):
        """After capture(): the gradient dict the replay of `step` refreshes (fixed tensors per parity)."""
        return self._g[step & 1]


class AbiFrame:
    """The same frame as a chain of raw C-ABI calls on device-resident tensors, no autograd (what bench.py's `value`
    times): fs_flame_forward -> fs_pose_forward -> fs_forward -> fs_backward -> fs_densify_stats_inc ->
    fs_pose_backward -> fs_flame_backward.  Outputs and scratch are allocated once; `run` writes this frame's leaf
    gradients into `views` (ShardedStep.fill_views) and either its FLAME factor record (`record`) or the dense delta
    gradients (`dense`)."""

    def __init__(self, model, raster_settings, betas, pose, dpix):
        from . import flame as _flame

        m = self.model = model
        dev = m._scaling.device
        self.rs, self.betas, self.pose, self.dpix = raster_settings, betas, pose, dpix
        self.fm = _flame.model_tensors(m.flame) if not isinstance(m.flame, dict) else m.flame
        P, V = m._scaling.shape[0], self.fm["v_template"].shape[0]
        self.P, self.V, self.L = P, V, self.fm["shapedirs"].shape[-1]
        self.l0 = int(m.flame.n_shape)
        self.pose_out = (torch.empty(P, 3, device=dev), torch.empty(P, 3, device=dev), torch.empty(P, 4, device=dev),
                         torch.empty(P, 1, device=dev))
        self.fl_out = None
        self.scratch = torch.empty(14 * P, device=dev)  # dL/d(xyz 3, scales 3, rotations 4, opacity 1, means2D 3)
        self.d_verts = torch.empty(V, 3, device=dev)
        self.launches = 0
        self.state = None

    def run(self, views, record=None, dense=None):
        from . import _lib, flame as _flame, pose as _pose, rasterizer as _R

        m, fm, P, V, L = self.model, self.fm, self.P, self.V, self.L
        f = lambda t: None if t is None else t.detach()
        self.fl_out = _flame.flame_forward_raw(
            self.betas, self.pose, fm["v_template"], fm["shapedirs"], fm["posedirs"], fm["J_regressor"], fm["parents"],
            fm["lbs_weights"], f(getattr(m, "delta_vertex", None)), f(getattr(m, "delta_shapedirs", None)),
            f(getattr(m, "delta_posedirs", None)), l0=self.l0, out=self.fl_out)
        verts = self.fl_out["verts"]
        raw = (m._scaling.detach(), m._rotation.detach(), m._offset.detach(), m._opacity.detach())
        canon = m.face_scaling_canonical.reshape(-1)
        xyz, sc, ro, op = _pose.pose_forward_raw(verts, m.faces, m.face_index, m.bary_coords, canon, *raw,
                                                 shell_len=m.shell_len, out=self.pose_out)
        color, radii, st = _R.forward_raw(self.rs, xyz, m._features_dc.detach(), None, op, sc, ro, None)
        s = self.scratch
        g = dict(means3D=s[0:3 * P], scales=s[3 * P:6 * P], rotations=s[6 * P:10 * P], opacity=s[10 * P:11 * P],
                 means2D=s[11 * P:14 * P], sh=views["features_dc"].view(-1))
        _R.backward_raw(st, self.dpix, out=g)
        lib = _lib.load()
        with _lib.on_device(xyz.device):
            rc = lib.fs_densify_stats_inc(P, g["means2D"].data_ptr(), radii.data_ptr(), views["accum_inc"].data_ptr(),
                                          views["denom_inc"].data_ptr(), _lib.stream_ptr(xyz.device))
        _lib.check(rc, "fs_densify_stats_inc")
        _pose.pose_backward_raw(verts, m.faces, m.face_index, m.bary_coords, canon, *raw, g["means3D"].view(P, 3),
                                g["scales"].view(P, 3), g["rotations"].view(P, 4), g["opacity"].view(P, 1),
                                shell_len=m.shell_len,
                                out=(self.d_verts, views["scaling"], views["rotation"], views["offset"], views["opacity"]))
        if record is not None:
            _flame.flame_backward_raw(self.betas, fm["J_regressor"], fm["parents"], fm["lbs_weights"],
                                      self.fl_out["workspace"], self.d_verts, (V, L), l0=self.l0,
                                      want=(False, False, False), record=record)
        else:
            _flame.flame_backward_raw(self.betas, fm["J_regressor"], fm["parents"], fm["lbs_weights"],
                                      self.fl_out["workspace"], self.d_verts, (V, L), l0=self.l0, out=dense)
        self.state = st
        # kernels + memset nodes: FLAME 2 + 2, pose 1 + (1 memset + 1), rasterizer (1 memset + n) + (1 memset + 2), stats 1
        self.launches = st["launches"] + st.get("launches_bwd", 0) + 2 + 4 + 3 + 1
        return color
