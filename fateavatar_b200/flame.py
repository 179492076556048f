"""FLAME skinning stage on the fused sm_100a kernels, behind the reference's own interface.

Mirrors flame/FLAME.py:131-204 (`FLAME.forward`, `FLAME.forward_with_delta_blendshape`) and flame/lbs.py:24-100:

    verts, pose_feature, transformations = flame.forward_with_delta_blendshape(expression, full_pose,
                                                delta_shapedirs, delta_posedirs, delta_vertex)
    verts_orig, _, _                     = flame.forward(expression, full_pose)

FateAvatar calls both, back to back, every frame (model/fateavatar.py:211-222).  `fs_flame_forward` produces both
results in one pass over the blendshape tensors; `attach(flame_module)` rebinds the two methods of an existing
reference FLAME module so the caller runs unchanged (the second call is served from the first one's by-product).
Gradients flow to delta_shapedirs / delta_posedirs / delta_vertex (the parameters FateAvatar trains) and, when they
require grad (per-frame tracking optimisation, train/base.py:113-151), to the expression and pose coefficients.
CUDA only; no CPU path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FateSplatError


def _parents_c(parents):
    p = [int(x) for x in (parents.tolist() if torch.is_tensor(parents) else list(parents))]
    return (C.c_int * len(p))(*p), len(p)


def _ptr(t):
    return None if t is None else t.data_ptr()


def flame_forward_raw(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights,
                      delta_vertex=None, delta_shapedirs=None, delta_posedirs=None, l0=0, want_orig=True,
                      workspace=None, out=None):
    """One fs_flame_forward call on contiguous fp32 CUDA tensors (no autograd).
    Returns dict(verts, verts_orig, pose_feature, transforms, transforms_orig, workspace); `out` may be the dict
    of a previous call, whose tensors are then reused instead of allocated."""
    lib = _lib.load()
    dev = v_template.device
    V, L = v_template.shape[0], shapedirs.shape[-1]
    pc, J = _parents_c(parents)
    nbytes = lib.fs_flame_workspace_bytes(V)
    if out is not None:
        verts, verts_orig, pf, A, A_orig, ws = (out[k] for k in ("verts", "verts_orig", "pose_feature", "transforms",
                                                                 "transforms_orig", "workspace"))
        with _lib.on_device(dev):
            rc = lib.fs_flame_forward(V, L, int(l0), J, pc, betas.data_ptr(), pose.data_ptr(), v_template.data_ptr(),
                                      _ptr(delta_vertex), shapedirs.data_ptr(), _ptr(delta_shapedirs), posedirs.data_ptr(),
                                      _ptr(delta_posedirs), J_regressor.data_ptr(), lbs_weights.data_ptr(),
                                      verts.data_ptr(), _ptr(verts_orig), pf.data_ptr(), A.data_ptr(), _ptr(A_orig),
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "fs_flame_forward")
        return out
    ws = workspace if workspace is not None and workspace.numel() >= nbytes else torch.empty(nbytes, dtype=torch.uint8, device=dev)
    verts = torch.empty((V, 3), device=dev)
    verts_orig = torch.empty((V, 3), device=dev) if want_orig else None
    pf = torch.empty(((J - 1) * 9,), device=dev)
    A = torch.empty((J, 4, 4), device=dev)
    A_orig = torch.empty((J, 4, 4), device=dev) if want_orig else None
    with _lib.on_device(dev):
        rc = lib.fs_flame_forward(V, L, int(l0), J, pc, betas.data_ptr(), pose.data_ptr(), v_template.data_ptr(),
                                  _ptr(delta_vertex), shapedirs.data_ptr(), _ptr(delta_shapedirs), posedirs.data_ptr(),
                                  _ptr(delta_posedirs), J_regressor.data_ptr(), lbs_weights.data_ptr(), verts.data_ptr(),
                                  _ptr(verts_orig), pf.data_ptr(), A.data_ptr(), _ptr(A_orig), ws.data_ptr(), ws.numel(),
                                  _lib.stream_ptr(dev))
    _lib.check(rc, "fs_flame_forward")
    return dict(verts=verts, verts_orig=verts_orig, pose_feature=pf, transforms=A, transforms_orig=A_orig, workspace=ws)


def flame_backward_raw(betas, J_regressor, parents, lbs_weights, workspace, dL_dverts, shapes, l0=0,
                       want=(True, True, True), out=None, factors=False, factor_out=None, record=None):
    """One fs_flame_backward call.  `shapes` = (V, L); `want` selects (delta_vertex, delta_shapedirs,
    delta_posedirs) gradients; `out` may supply preallocated tensors for them.  With factors=True also returns
    (dL_dv_shaped, dL_dv_posed), the [V,3] factors of the two rank-1 gradients (written into `factor_out` when
    given, e.g. views of an all-gather record).  `record` (a flat factor record, see factor_record_floats) makes the
    call fill this rank's whole record in place: [betas | pose_feature | dL_dv_shaped | dL_dv_posed]."""
    lib = _lib.load()
    dev = dL_dverts.device
    V, L = shapes
    pc, J = _parents_c(parents)
    o = list(out) if out is not None else [None, None, None]
    if want[0] and o[0] is None:
        o[0] = torch.empty((V, 3), device=dev)
    if want[1] and o[1] is None:
        o[1] = torch.empty((V, 3, L), device=dev)
    if want[2] and o[2] is None:
        o[2] = torch.empty(((J - 1) * 9, V * 3), device=dev)
    header = None
    if record is not None:
        NPr = (J - 1) * 9
        header = record
        factor_out = (record[L + NPr:L + NPr + 3 * V], record[L + NPr + 3 * V:L + NPr + 6 * V])
    if factor_out is not None:
        gs, gp = factor_out
        factors = True
    else:
        gs = torch.empty((V, 3), device=dev) if factors else None
        gp = torch.empty((V, 3), device=dev) if factors else None
    with _lib.on_device(dev):
        rc = lib.fs_flame_backward(V, L, int(l0), J, pc, betas.data_ptr(), J_regressor.data_ptr(), lbs_weights.data_ptr(),
                                   dL_dverts.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                   _ptr(o[0]) if want[0] else None, _ptr(o[1]) if want[1] else None,
                                   _ptr(o[2]) if want[2] else None, _ptr(gs), _ptr(gp), _ptr(header),
                                   _lib.stream_ptr(dev))
    _lib.check(rc, "fs_flame_backward")
    return (o[0], o[1], o[2]) + ((gs, gp) if factors else ())


_factor_sink = [None]


class factor_record:
    """`with flame.factor_record(record):` -- while active, the backward of flame_lbs writes this frame's rank-1 factor
    record [betas | pose_feature | dL/dv_shaped | dL/dv_posed] (factor_record_floats) into `record` INSTEAD of the
    dense delta gradients (26 MB at FLAME size), and autograd delivers no gradient to delta_vertex / delta_shapedirs /
    delta_posedirs.  Frame-sharded training exchanges the records and expands their sum (parallel.ShardedStep)."""

    def __init__(self, record):
        self.record, self.prev = record, None

    def __enter__(self):
        self.prev, _factor_sink[0] = _factor_sink[0], self.record
        return self

    def __exit__(self, *exc):
        _factor_sink[0] = self.prev
        return False


class _FlameLBS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, betas, pose, delta_vertex, delta_shapedirs, delta_posedirs, model, l0, want_orig):
        if not model["v_template"].is_cuda:
            raise FateSplatError("flame_lbs needs CUDA tensors: fateavatar_b200 has no CPU path")
        f = lambda t: None if t is None else t.detach().contiguous().float()
        b, p = f(betas).reshape(-1), f(pose).reshape(-1)
        dv, ds, dp = f(delta_vertex), f(delta_shapedirs), f(delta_posedirs)
        r = flame_forward_raw(b, p, model["v_template"], model["shapedirs"], model["posedirs"], model["J_regressor"],
                              model["parents"], model["lbs_weights"], dv, ds, dp, l0=l0, want_orig=want_orig)
        ctx.model, ctx.l0, ctx.ws, ctx.betas, ctx.pose = model, l0, r["workspace"], b, p
        ctx.deltas = (ds, dp)  # only read again when the coefficients themselves are being optimised
        ctx.in_shapes = (betas.shape, pose.shape)
        ctx.have = (delta_vertex is not None, delta_shapedirs is not None, delta_posedirs is not None)
        outs = (r["verts"], r["pose_feature"], r["transforms"])
        if want_orig:
            outs += (r["verts_orig"], r["transforms_orig"])
        ctx.mark_non_differentiable(*outs[1:])
        ctx.set_materialize_grads(False)  # no zero-filled gradients for the four by-products on every backward
        return outs

    @staticmethod
    def backward(ctx, g_verts, *_unused):
        m = ctx.model
        V, L = m["v_template"].shape[0], m["shapedirs"].shape[-1]
        want = tuple(h and ctx.needs_input_grad[2 + i] for i, h in enumerate(ctx.have))
        want_b, want_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if g_verts is None or not (any(want) or want_b or want_p):
            return (None,) * 8
        if _factor_sink[0] is not None and any(want) and not (want_b or want_p):
            flame_backward_raw(ctx.betas, m["J_regressor"], m["parents"], m["lbs_weights"], ctx.ws,
                               g_verts.contiguous().float().reshape(V, 3), (V, L), l0=ctx.l0,
                               want=(False, False, False), record=_factor_sink[0])
            return (None,) * 8
        gdv, gds, gdp = flame_backward_raw(ctx.betas, m["J_regressor"], m["parents"], m["lbs_weights"], ctx.ws,
                                           g_verts.contiguous().float().reshape(V, 3), (V, L), l0=ctx.l0, want=want)
        gb = gp = None
        if want_b or want_p:  # per-frame tracking optimisation (train/base.py:113-151)
            gb, gp = flame_backward_coeffs_raw(m, ctx.pose, ctx.deltas[0], ctx.deltas[1], ctx.ws, (V, L), l0=ctx.l0,
                                               want=(want_b, want_p))
            gb = gb.reshape(ctx.in_shapes[0]) if gb is not None else None
            gp = gp.reshape(ctx.in_shapes[1]) if gp is not None else None
        return gb, gp, gdv, gds, gdp, None, None, None


def flame_backward_coeffs_raw(model, pose, delta_shapedirs, delta_posedirs, workspace, shapes, l0=0, want=(True, True)):
    """One fs_flame_backward_coeffs call (after flame_backward_raw on the same workspace): gradients of the
    blendshape coefficients [L] (zeros below l0) and of the axis-angle pose [J*3]."""
    lib = _lib.load()
    dev = pose.device
    V, L = shapes
    pc, J = _parents_c(model["parents"])
    gb = torch.empty((L,), device=dev) if want[0] else None
    gp = torch.empty((J * 3,), device=dev) if want[1] else None
    with _lib.on_device(dev):
        rc = lib.fs_flame_backward_coeffs(V, L, int(l0), J, pc, pose.data_ptr(), model["shapedirs"].data_ptr(),
                                          _ptr(delta_shapedirs), model["posedirs"].data_ptr(), _ptr(delta_posedirs),
                                          model["J_regressor"].data_ptr(), workspace.data_ptr(), workspace.numel(),
                                          _ptr(gb), _ptr(gp), _lib.stream_ptr(dev))
    _lib.check(rc, "fs_flame_backward_coeffs")
    return gb, gp


def model_tensors(flame_module):
    """The buffers flame/FLAME.py:72-107 registers, as contiguous fp32 CUDA tensors (parents stays a host list)."""
    g = lambda n: getattr(flame_module, n).detach().contiguous().float()
    return dict(v_template=g("v_template"), shapedirs=g("shapedirs"), posedirs=g("posedirs"),
                J_regressor=g("J_regressor"), lbs_weights=g("lbs_weights"),
                parents=[int(x) for x in flame_module.parents.tolist()])


def flame_lbs(model, betas, pose, delta_shapedirs=None, delta_posedirs=None, delta_vertex=None, l0=0, want_orig=True):
    """model: dict from `model_tensors` (or the same keys built by hand); betas [L] or [1,L]; pose [J*3] or [1,J*3].
    Returns (verts [1,V,3], pose_feature [1,(J-1)*9], transformations [1,J,4,4]) and, with want_orig, additionally
    (verts_orig [1,V,3], transformations_orig [1,J,4,4]) -- the same pose without the deltas."""
    if betas.dim() == 2 and betas.shape[0] != 1:
        raise FateSplatError("flame_lbs handles one frame per call (the reference's batch size); loop over frames")
    outs = _FlameLBS.apply(betas, pose, delta_vertex, delta_shapedirs, delta_posedirs, model, int(l0), bool(want_orig))
    if want_orig and (betas.requires_grad or pose.requires_grad):
        # Per-frame tracking optimisation (train/base.py:113-151): upstream's flame.forward(expression, pose) is
        # differentiable w.r.t. the coefficients (flame_loss = (verts - verts_orig)^2, train/loss.py:197-201), the
        # by-product of the fused pass is not -- so the undeformed mesh gets its own differentiable pass here.
        o2 = _FlameLBS.apply(betas, pose, None, None, None, model, int(l0), False)
        outs = (outs[0], outs[1], outs[2], o2[0], o2[2])
    return tuple(o[None] for o in outs)


def attach(flame_module):
    """Rebind `forward_with_delta_blendshape` and `forward` of a reference FLAME module (flame/FLAME.py) to the
    fused kernels.  model/fateavatar.py:211-222 then runs unchanged: the first call computes both meshes, the
    second returns the by-product when it is asked for the same (expression, pose) tensors."""
    model = model_tensors(flame_module)
    n_shape, n_exp = int(flame_module.n_shape), int(flame_module.n_exp)
    cache = {}

    def betas_of(expression_params):
        e = expression_params[:, :n_exp]
        return torch.cat([torch.zeros(e.shape[0], n_shape, device=e.device, dtype=e.dtype), e], dim=1)

    def key_of(expression_params, full_pose):
        # the entry keeps strong references to the two tensors, so neither address can be recycled while it is cached
        return (expression_params, expression_params._version, full_pose, full_pose._version)

    def same(a, b):
        return a is not None and a[0] is b[0] and a[1] == b[1] and a[2] is b[2] and a[3] == b[3]

    def forward_with_delta_blendshape(expression_params, full_pose, delta_shapedirs=None, delta_posedirs=None,
                                      delta_vertex=None):
        v, pf, A, vo, Ao = flame_lbs(model, betas_of(expression_params), full_pose, delta_shapedirs, delta_posedirs,
                                     delta_vertex, l0=n_shape, want_orig=True)
        cache.clear()
        cache["key"], cache["val"] = key_of(expression_params, full_pose), (vo, pf, Ao)
        return v, pf, A

    def forward(expression_params, full_pose):
        hit = cache.get("val") if same(cache.get("key"), key_of(expression_params, full_pose)) else None
        cache.clear()
        if hit is not None:
            return hit
        v, pf, A = flame_lbs(model, betas_of(expression_params), full_pose, l0=n_shape, want_orig=False)
        return v, pf, A

    flame_module.forward_with_delta_blendshape = forward_with_delta_blendshape
    flame_module.forward = forward
    return flame_module


# ---- data-parallel exchange of the delta gradients in factored form -----------------------------

def factor_record_floats(V, L, NP):
    """Floats per rank record [betas L | pose_feature NP | dL_dv_shaped 3V | dL_dv_posed 3V], padded to 4."""
    return (L + NP + 6 * V + 3) // 4 * 4


def pack_factors(record, betas, pose_feature, dL_dv_shaped, dL_dv_posed):
    """Fill one rank's record (a flat float tensor of factor_record_floats) from the outputs of
    flame_forward_raw (pose_feature) and flame_backward_raw(..., factors=True)."""
    L, NP, n3 = betas.numel(), pose_feature.numel(), dL_dv_shaped.numel()
    record[:L].copy_(betas.reshape(-1))
    record[L:L + NP].copy_(pose_feature.reshape(-1))
    record[L + NP:L + NP + n3].copy_(dL_dv_shaped.reshape(-1))
    record[L + NP + n3:L + NP + 2 * n3].copy_(dL_dv_posed.reshape(-1))
    return record


def expand_factors_reference(gathered, V, L, NP, scale=1.0):
    """Plain-torch statement of what fs_flame_expand_grads computes (any device): used by the host-side tests."""
    n3 = 3 * V
    b, pf = gathered[:, :L], gathered[:, L:L + NP]
    gs, gp = gathered[:, L + NP:L + NP + n3], gathered[:, L + NP + n3:L + NP + 2 * n3]
    return ((scale * gs.sum(0)).view(V, 3), scale * torch.einsum("re,rl->el", gs, b).view(V, 3, L),
            scale * torch.einsum("ri,re->ie", pf, gp))


def expand_factors(gathered, V, L, NP, l0=0, scale=1.0, out=None):
    """gathered [N, record] (the all-gather result, CUDA) -> dense (d_delta_vertex [V,3], d_delta_shapedirs [V,3,L],
    d_delta_posedirs [NP,3V]) summed over the N ranks, via fs_flame_expand_grads."""
    if not gathered.is_cuda:
        raise FateSplatError("expand_factors needs CUDA tensors: fateavatar_b200 has no CPU path")
    dev = gathered.device
    N, stride = gathered.shape
    o = list(out) if out is not None else [torch.empty((V, 3), device=dev), torch.empty((V, 3, L), device=dev),
                                            torch.empty((NP, 3 * V), device=dev)]
    with _lib.on_device(dev):
        rc = _lib.load().fs_flame_expand_grads(N, V, L, int(l0), NP, gathered.data_ptr(), stride, float(scale),
                                               _ptr(o[0]), _ptr(o[1]), _ptr(o[2]),
                                               _lib.stream_ptr(dev))
    _lib.check(rc, "fs_flame_expand_grads")
    return tuple(o)


def allgather_delta_grads(record, V, L, NP, l0=0, scale=1.0, out=None, group=None, gathered=None):
    """All-gather every rank's factor record (~120 KB) and expand the summed dense delta gradients locally:
    the same result as all-reducing the 26 MB dense gradients, at 1/200 of the wire traffic."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if gathered is None:
        gathered = torch.empty((world, record.numel()), device=record.device, dtype=record.dtype)
    dist.all_gather_into_tensor(gathered.view(-1), record, group=group)
    return expand_factors(gathered, V, L, NP, l0=l0, scale=scale, out=out)
