"""Optimiser step and splat-set maintenance of the FateAvatar optimise loop on the device.

Mirrors, for the GPU path,
    train/optim.py:11-37                    two torch.optim.Adam over 8 parameter groups   -> FusedAdam (fs_adam_step)
    model/fateavatar.py:610-672  _uv_densify                                               -> SplatStore.uv_densify
    model/fateavatar.py:674-713  _prune_low_opacity_points                                 -> SplatStore.prune_low_opacity
    model/fateavatar.py:715-732  _reset_opacity                                            -> SplatStore.reset_opacity
    train/iteration.py:21-89     iteration_step_fateavatar                                 -> OptimiseLoop.step

Upstream every densification torch.cat()s each parameter and both Adam moments and every prune boolean-mask-indexes 19
tensors, re-creating the nn.Parameters and patching optimizer.state.  Here everything that has one row per splat lives
in ONE capacity-allocated structure of arrays: densify appends rows in place, prune is a stable stream compaction into
the twin buffer set, and the model's attributes are re-bound to [:P] views -- no allocation, no optimizer-state surgery.
The random draws of _uv_densify stay torch.multinomial / torch.rand (same generator semantics as upstream, so a run is
reproducible against the reference's own code; tests/test_optimise_loop_gpu.py).  CUDA only.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FateSplatError

FIELDS = (("opacity", "_opacity", 1), ("offset", "_offset", 1), ("color", "_features_dc", 3),
          ("rotation", "_rotation", 4), ("scaling", "_scaling", 3))  # optimizer group name, model attribute, row width
ADAM_MAX = 8


class FsAdamTensor(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_size_t), ("lr", C.c_float)]


class FsSplatSoa(C.Structure):
    _fields_ = [("opacity", C.c_void_p), ("offset", C.c_void_p), ("color", C.c_void_p), ("rotation", C.c_void_p),
                ("scaling", C.c_void_p), ("exp_avg", C.c_void_p * 5), ("exp_avg_sq", C.c_void_p * 5),
                ("face_index", C.c_void_p), ("bary", C.c_void_p), ("accum", C.c_void_p), ("denom", C.c_void_p),
                ("max_radii2D", C.c_void_p), ("sample_flag", C.c_void_p)]


def _bind(lib):
    if getattr(lib, "_fs_optim_bound", False):
        return lib
    vp, i, f, sz, dbl = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_double
    lib.fs_adam_step.restype = i
    lib.fs_adam_step.argtypes = [i, C.POINTER(FsAdamTensor), vp, dbl, dbl, dbl, vp]
    lib.fs_splat_append.restype = i
    lib.fs_splat_append.argtypes = [C.POINTER(FsSplatSoa), i, i, vp, vp, vp]
    lib.fs_splat_prune_workspace_bytes.restype = sz
    lib.fs_splat_prune_workspace_bytes.argtypes = [i]
    lib.fs_splat_prune.restype = i
    lib.fs_splat_prune.argtypes = [C.POINTER(FsSplatSoa), C.POINTER(FsSplatSoa), i, f, vp, vp, sz, vp]
    lib.fs_opacity_reset.restype = i
    lib.fs_opacity_reset.argtypes = [vp, vp, vp, i, f, vp]
    lib._fs_optim_bound = True
    return lib


class SplatStore:
    """Everything with one row per splat, allocated once for `capacity` rows (two buffer sets: prune compacts from one
    into the other).  `bind()` points the model's attributes at the live [:P] views; the five parameters stay leaf
    nn.Parameters, so autograd, torch optimizers and the fused kernels all see the same memory."""

    def __init__(self, model, capacity):
        dev = model._scaling.device
        if not model._scaling.is_cuda:
            raise FateSplatError("SplatStore needs CUDA tensors: fateavatar_b200 has no CPU path")
        self.model, self.device = model, dev
        self.P = int(model._scaling.shape[0])
        self.capacity = int(max(capacity, self.P))
        cap = self.capacity

        def alloc():
            s = {n: torch.zeros(cap, w, device=dev) for n, _, w in FIELDS}
            s.update({"m_" + n: torch.zeros(cap, w, device=dev) for n, _, w in FIELDS})
            s.update({"v_" + n: torch.zeros(cap, w, device=dev) for n, _, w in FIELDS})
            s.update(face_index=torch.zeros(cap, dtype=torch.int64, device=dev), bary=torch.zeros(cap, 3, device=dev),
                     accum=torch.zeros(cap, 1, device=dev), denom=torch.zeros(cap, 1, device=dev),
                     max_radii2D=torch.zeros(cap, device=dev), sample_flag=torch.zeros(cap, device=dev))
            return s

        self.sets = [alloc(), alloc()]
        self.cur = 0
        a, P = self.sets[0], self.P
        for n, attr, w in FIELDS:
            a[n][:P].copy_(getattr(model, attr).detach().reshape(P, w))
        a["face_index"][:P].copy_(model.face_index)
        a["bary"][:P].copy_(model.bary_coords)
        for key, attr in (("accum", "xyz_gradient_accum"), ("denom", "denom"), ("max_radii2D", "max_radii2D"),
                          ("sample_flag", "sample_flag")):
            t = getattr(model, attr, None)
            if t is not None and t.numel() == P:
                a[key][:P].copy_(t.reshape(a[key][:P].shape))
        self.new_P = torch.zeros(1, dtype=torch.int32, device=dev)
        self.prune_ws = torch.empty(_bind(_lib.load()).fs_splat_prune_workspace_bytes(cap), dtype=torch.uint8, device=dev)
        self.bind()

    # -- views ------------------------------------------------------------------------------------------------------
    def arrays(self, k=None):
        return self.sets[self.cur if k is None else k]

    def view(self, name):
        return self.arrays()[name][:self.P]

    def bind(self):
        """(Re)point the model at the live rows.  Parameters are fresh leaf nn.Parameters over the SAME memory."""
        m, a, P = self.model, self.arrays(), self.P
        for n, attr, w in FIELDS:
            shape = (P, 1, 3) if attr == "_features_dc" else (P, w)
            setattr(m, attr, torch.nn.Parameter(a[n][:P].view(shape)))
        m.face_index, m.bary_coords = a["face_index"][:P], a["bary"][:P]
        m.xyz_gradient_accum, m.denom = a["accum"][:P], a["denom"][:P]
        m.max_radii2D, m.sample_flag = a["max_radii2D"][:P], a["sample_flag"][:P]
        m.num_points = P
        return m

    def load_rows(self, rows):
        """Refill the store from a checkpoint's splat attributes (io.restore_splats; train/deserialize.py:7-40): `rows`
        maps _offset / _features_dc / _scaling / _rotation / _opacity / face_index / bary_coords to tensors with P rows.
        As upstream, the optimizer state is not part of a checkpoint: both Adam moments and the densification statistics
        restart from zero.  In place -- the capacity arrays are kept."""
        P = int(rows["_offset"].shape[0])
        if P > self.capacity:
            raise FateSplatError(f"SplatStore capacity {self.capacity} < {P} checkpoint rows")
        a = self.arrays()
        for n, attr, w in FIELDS:
            a[n][:P].copy_(rows[attr].detach().reshape(P, w))
            a["m_" + n].zero_()
            a["v_" + n].zero_()
        a["face_index"][:P].copy_(rows["face_index"])
        a["bary"][:P].copy_(rows["bary_coords"])
        for key in ("accum", "denom", "max_radii2D", "sample_flag"):
            a[key].zero_()
        self.P = P
        return self.bind()

    def _soa(self, k=None):
        a = self.arrays(k)
        s = FsSplatSoa()
        for n, _, _ in FIELDS:
            setattr(s, n, a[n].data_ptr())
        s.exp_avg = (C.c_void_p * 5)(*[a["m_" + n].data_ptr() for n, _, _ in FIELDS])
        s.exp_avg_sq = (C.c_void_p * 5)(*[a["v_" + n].data_ptr() for n, _, _ in FIELDS])
        s.face_index, s.bary = a["face_index"].data_ptr(), a["bary"].data_ptr()
        s.accum, s.denom = a["accum"].data_ptr(), a["denom"].data_ptr()
        s.max_radii2D, s.sample_flag = a["max_radii2D"].data_ptr(), a["sample_flag"].data_ptr()
        return s

    # -- model/fateavatar.py:610-672 -----------------------------------------------------------------------------------
    def uv_densify(self, increase_num=1000, generator=None):
        """parents ~ multinomial(xyz_gradient_accum, increase_num, replacement) and fresh barycentrics uvw / sum(uvw) with
        uvw ~ U(0,1)^3 -- the reference's two draws, in its order, from `generator` (None: the global CUDA generator, as
        upstream) -- then ONE kernel appends the children in place."""
        n = int(increase_num)
        if self.P + n > self.capacity:
            raise FateSplatError(f"SplatStore capacity {self.capacity} < {self.P} + {n}: allocate for max_points_num")
        parents = self.view("accum").squeeze(1).multinomial(n, replacement=True, generator=generator)
        uvw = torch.rand((n, 3), device=self.device, generator=generator)
        new_bary = (uvw / uvw.sum(dim=-1, keepdim=True)).contiguous()
        lib = _bind(_lib.load())
        soa = self._soa()
        with _lib.on_device(self.device):
            rc = lib.fs_splat_append(C.byref(soa), self.P, n, parents.data_ptr(), new_bary.data_ptr(),
                                     _lib.stream_ptr(self.device))
        _lib.check(rc, "fs_splat_append")
        self.P += n
        self.bind()
        return parents

    # -- model/fateavatar.py:674-713 -----------------------------------------------------------------------------------
    def prune_low_opacity(self, min_opacity=0.05):
        lib = _bind(_lib.load())
        src, dst = self._soa(self.cur), self._soa(1 - self.cur)
        with _lib.on_device(self.device):
            rc = lib.fs_splat_prune(C.byref(src), C.byref(dst), self.P, float(min_opacity), self.new_P.data_ptr(),
                                    self.prune_ws.data_ptr(), self.prune_ws.numel(), _lib.stream_ptr(self.device))
        _lib.check(rc, "fs_splat_prune")
        self.P = int(self.new_P.item())  # the one host read of a prune (upstream: boolean-mask indexing syncs too)
        self.cur = 1 - self.cur
        self.bind()
        return self.P

    # -- model/fateavatar.py:715-732 -----------------------------------------------------------------------------------
    def reset_opacity(self, cap=0.01):
        a = self.arrays()
        with _lib.on_device(self.device):
            rc = _bind(_lib.load()).fs_opacity_reset(a["opacity"].data_ptr(), a["m_opacity"].data_ptr(),
                                                     a["v_opacity"].data_ptr(), self.P, float(cap),
                                                     _lib.stream_ptr(self.device))
        _lib.check(rc, "fs_opacity_reset")


class FusedAdam:
    """torch.optim.Adam(betas=(0.9, 0.999), eps=1e-8) over up to 8 tensors in ONE launch per step (fs_adam_step).
    `groups`: list of dicts {"name", "param": callable -> tensor, "grad": callable -> tensor, "m", "v": tensors or
    callables, "lr"} -- callables are resolved at every step(), so the splat tensors may be re-bound by SplatStore."""

    def __init__(self, groups, device, betas=(0.9, 0.999), eps=1e-8):
        if not 1 <= len(groups) <= ADAM_MAX:
            raise FateSplatError(f"FusedAdam takes 1..{ADAM_MAX} tensors")
        self.groups, self.betas, self.eps, self.device = groups, betas, eps, torch.device(device)
        self.steps = torch.zeros(ADAM_MAX + 1, dtype=torch.int32, device=self.device)

    def step(self):
        arr = (FsAdamTensor * len(self.groups))()
        r = lambda x: x() if callable(x) else x
        keep = []
        for k, g in enumerate(self.groups):
            p, gr, m, v = r(g["param"]), r(g["grad"]), r(g["m"]), r(g["v"])
            if gr is None:
                raise FateSplatError(f"FusedAdam: group {g['name']} has no gradient (every tensor of the table steps "
                                     "together; build the table from the tensors that are trained)")
            if not (p.is_contiguous() and gr.is_contiguous() and m.is_contiguous() and v.is_contiguous()):
                raise FateSplatError(f"FusedAdam: group {g['name']} must be contiguous")
            if not (p.numel() == gr.numel() <= m.numel() and m.numel() == v.numel()):
                raise FateSplatError(f"FusedAdam: group {g['name']} size mismatch")
            keep += [p, gr, m, v]
            arr[k] = FsAdamTensor(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), float(g["lr"]))
        with _lib.on_device(self.device):
            rc = _bind(_lib.load()).fs_adam_step(len(self.groups), arr, self.steps.data_ptr(), self.betas[0], self.betas[1],
                                                 self.eps, _lib.stream_ptr(self.device))
        _lib.check(rc, "fs_adam_step")


def _flat_grad(p):
    return None if p.grad is None else p.grad.reshape(-1)


def fateavatar_adam(model, store, lrs, delta_grads=None, grads=None):
    """The reference's optimizer groups (train/optim.py:15-35) on FusedAdam: the five splat tensors (moments in the
    store, so they follow densify / prune) + delta_shapedirs / delta_posedirs / delta_vertex.  `lrs`: dict with the
    config/fateavatar.yaml names (opacity_lr, offset_lr, feature_dc_lr, rotation_lr, scaling_lr, delta_shapedirs_lr,
    delta_posedirs_lr); delta_vertex uses 1e-4 like upstream.  `grads` / `delta_grads` optionally map a group name to
    a callable returning the gradient tensor (e.g. the summed gradients of parallel.ShardedStep); default: `.grad`."""
    lr_of = {"opacity": lrs["opacity_lr"], "offset": lrs["offset_lr"], "color": lrs["feature_dc_lr"],
             "rotation": lrs["rotation_lr"], "scaling": lrs["scaling_lr"]}
    grads, delta_grads = grads or {}, delta_grads or {}
    groups = []
    for n, attr, w in FIELDS:
        groups.append(dict(name=n, lr=lr_of[n], param=(lambda a=attr: getattr(model, a).detach().view(-1)),
                           grad=grads.get(n, (lambda a=attr: _flat_grad(getattr(model, a)))),
                           m=(lambda n=n: store.arrays()["m_" + n].view(-1)), v=(lambda n=n: store.arrays()["v_" + n].view(-1))))
    dev = model._scaling.device
    for n, lr in (("delta_shapedirs", lrs["delta_shapedirs_lr"]), ("delta_posedirs", lrs["delta_posedirs_lr"]),
                  ("delta_vertex", 0.0001)):
        p = getattr(model, n, None)
        if p is None:
            continue
        groups.append(dict(name=n, lr=lr, param=(lambda n=n: getattr(model, n).detach().view(-1)),
                           grad=delta_grads.get(n, (lambda n=n: _flat_grad(getattr(model, n)))),
                           m=torch.zeros(p.numel(), device=dev), v=torch.zeros(p.numel(), device=dev)))
    return FusedAdam(groups, dev)


class OptimiseLoop:
    """`iteration_step_fateavatar` (train/iteration.py:21-89) on the GPU path, one CUDA-graph launch per step:

        forward (avatar.forward_frame) -> loss -> backward -> [frame-sharded: fused exchange] -> densification statistics
        -> both Adam steps (FusedAdam)                                            -- recorded once, replayed per frame
        every densify_interval / prune_interval / opacity_reset_interval steps: SplatStore.uv_densify / prune_low_opacity
        / reset_opacity, in upstream's order and with its `global_step % interval == 0` rule, then re-record (P changed).

    `frame_loss_fn(model, inputs) -> (loss, out)` is the caller's loss on avatar.forward_frame's output; `training` holds
    config/fateavatar.yaml's `training:` keys.  With torch.distributed initialised every rank runs the same loop on its
    own frames (parallel.ShardedStep sums gradients and statistics; densify draws come from a generator that is in the
    same state on every rank)."""

    DEFAULTS = dict(feature_dc_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001, offset_lr=0.0016,
                    delta_shapedirs_lr=0.00001, delta_posedirs_lr=0.00001, opacity_reset_interval=60000,
                    densify_interval=3000, prune_interval=2000, min_opacity=0.005, increase_num=1000,
                    max_points_num=200000)

    def __init__(self, model, frame_loss_fn, example_inputs, training=None, generator=None, capture=True, log=None,
                 on_frame=None):
        """on_frame(loss, out, grads): optional observer called after the exchange and before the optimiser step of
        every live frame (eager mode only: Python callbacks are not part of a recorded graph)."""
        from . import parallel as _parallel

        self.on_frame = on_frame
        if on_frame is not None and capture:
            raise FateSplatError("OptimiseLoop: on_frame needs capture=False")

        self.cfg = dict(self.DEFAULTS, **(training or {}))
        self.model, self.frame_loss_fn, self.example = model, frame_loss_fn, example_inputs
        self.generator, self.use_graph, self.log = generator, capture, log or (lambda *_: None)
        self.device = model._scaling.device
        self.store = SplatStore(model, self.cfg["max_points_num"] + self.cfg["increase_num"])
        self._parallel = _parallel
        self.global_step = 0
        self._last = None
        self._staged_host = None
        self.recaptures = 0
        self.recording_s = 0.0  # wall clock spent (re)recording the step, for reports
        self._build()

    def _build(self):
        """(Re)create what depends on P: the sharded step (gradient layout) and the recorded graphs."""
        import time

        t0 = time.perf_counter()
        try:
            self._build_impl()
        finally:
            torch.cuda.synchronize(self.device)
            self.recording_s += time.perf_counter() - t0

    def _build_impl(self):
        m = self.model
        self.sharded = self._parallel.ShardedStep(m, device=self.device)
        sh = self.sharded
        if sh.world > 1:
            gsrc = {n: (lambda n=n, k=k: self._cur_grads()[k].reshape(-1)) for n, k in
                    (("opacity", "opacity"), ("offset", "offset"), ("color", "features_dc"), ("rotation", "rotation"),
                     ("scaling", "scaling"))}
            dsrc = {n: (lambda n=n: self._cur_grads()[n].reshape(-1)) for n in self._parallel.DELTA_NAMES}
        else:
            gsrc = dsrc = None
        if not hasattr(self, "adam"):
            self.adam = fateavatar_adam(m, self.store, self.cfg, delta_grads=dsrc, grads=gsrc)
        else:  # keep the delta moments and all step counters; only the gradient sources follow the new ShardedStep
            for g in self.adam.groups:
                src = (gsrc or {}).get(g["name"]) or (dsrc or {}).get(g["name"])
                if src is not None:
                    g["grad"] = src
        self._parity = 0
        self.captured = None
        if self.use_graph:
            from .graph import CapturedStep

            leaves = [getattr(m, a) for _, a, _ in FIELDS] + [getattr(m, n) for n in self._parallel.DELTA_NAMES
                                                              if getattr(m, n, None) is not None]
            # warm-up steps of the recording must not advance the optimiser: they run with the Adam launch disabled
            self._adam_on = False
            self.captured = []
            for parity in (0, 1):
                self._parity = parity
                self.captured.append(CapturedStep(lambda inp, parity=parity: self._body(parity, inp), self.example,
                                                  params=leaves, warmup=2, device=self.device, on_capture=self._enable_adam))
                self._adam_on = False
            self.recaptures += 1

    def _enable_adam(self):
        self._adam_on = True

    def _cur_grads(self):
        return self._g[self._parity]

    def _body(self, parity, inp):
        sh = self.sharded
        self._parity = parity
        loss, out = sh.run_autograd(parity, self.frame_loss_fn, inp)
        g = sh.exchange(parity)
        if not hasattr(self, "_g"):
            self._g = [None, None]
        self._g[parity] = g
        if getattr(self, "_adam_on", True):  # (not during the eager warm-up of a recording: it must leave no trace)
            if self.on_frame is not None:
                self.on_frame(loss, out, g)
            sh.apply_densify_stats(g)
            self.adam.step()
        return {"loss": loss.detach().reshape(1)}

    def prefetch(self, host_inputs):
        """Stage the pinned host inputs of the NEXT step() on a side stream (the pinned GT-frame prefetch of SURVEY 8f N3;
        upstream loads every frame synchronously, train/dataset.py:14-54): issued right after a step() it overlaps that
        step instead of preceding the next one.  The next call is then `step()` without arguments."""
        self._staged_host = host_inputs
        if self.captured is not None:
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            self.captured[self.global_step & 1].prefetch(host_inputs, self._copy_stream)

    def step(self, host_inputs=None):
        """One iteration on `host_inputs`, or on what `prefetch` staged.  Returns the pinned {"loss": [1]} of the step
        (valid after `wait()`, or after `last.wait()` for a caller that keeps `last = loop.last` and reads one step
        behind its launches)."""
        c, t = self.cfg, self.global_step
        if host_inputs is None:
            host_inputs = getattr(self, "_staged_host", None)
            if host_inputs is None:
                raise FateSplatError("OptimiseLoop.step(): no inputs given and nothing staged by prefetch()")
            if self.captured is not None and self.captured[t & 1]._staged is not None:
                host_inputs = None  # the side-stream copy into this recording is under way: replay waits for its event
        self._staged_host = None
        if self.captured is not None:
            cap = self.captured[t & 1]
            self._parity = t & 1
            res = cap(host_inputs)
            self._last = cap
        else:
            for _, a, _ in FIELDS:
                getattr(self.model, a).grad = None
            for n in self._parallel.DELTA_NAMES:
                if getattr(self.model, n, None) is not None:
                    getattr(self.model, n).grad = None
            self._adam_on = True
            res = self._body(t & 1, {k: v.to(self.device, non_blocking=True) for k, v in host_inputs.items()})
            self._last = None
        changed = False
        if t % c["densify_interval"] == 0:  # train/iteration.py:63-74
            old = self.store.P
            if old < c["max_points_num"]:
                self.store.uv_densify(min(c["max_points_num"] - old, c["increase_num"]), generator=self.generator)
                self.log(f"Do UV densification, Guassian splats: {old} --> {self.store.P}.")
                changed = True
        if t % c["prune_interval"] == 0:  # :77-82
            old = self.store.P
            changed |= self.store.prune_low_opacity(c["min_opacity"]) != old
            self.log(f"Prune low opacity points, Guassian splats: {old} --> {self.store.P}.")
        if t % c["opacity_reset_interval"] == 0 and t != 0:  # :84-85
            self.store.reset_opacity()
        if changed:
            if self._last is not None:
                self._last.wait()
            self._build()
        self.global_step += 1
        return res

    @property
    def last(self):
        """The recording the last step() replayed (None in eager mode): `last.wait()` blocks until THAT step is done."""
        return self._last

    def wait(self):
        if self._last is not None:
            return self._last.wait()
        torch.cuda.current_stream(self.device).synchronize()
