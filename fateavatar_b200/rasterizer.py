"""Drop-in for the `diff_gaussian_rasterization` Python operator API, backed by libfatesplat.so.

Mirrors DGR diff_gaussian_rasterization/__init__.py:21-220 -- same names, arguments, return values, error
behaviour and autograd contract -- so volume_rendering/render_3dgs.py:3,33-76 and
model/baseline/monogaussianavatar.py:394-421 run unchanged:

  * GaussianRasterizationSettings: NamedTuple with the reference's 12 fields (__init__.py:157-169)
  * GaussianRasterizer(raster_settings).forward(means3D, means2D, opacities, shs=None, colors_precomp=None,
        scales=None, rotations=None, cov3D_precomp=None) -> (color[3,H,W], radii[P] int32)   (:187-220)
    raising Exception unless exactly one of shs/colors_precomp and one of (scales,rotations)/cov3D_precomp
  * GaussianRasterizer.markVisible(positions) -> bool [P]                                     (:176-185)
  * autograd: gradients are returned in the reference's order (means3D, means2D, sh, colors_precomp,
    opacities, scales, rotations, cov3Ds_precomp, None)  (:143-153); `means2D` is the dummy tensor whose
    .grad receives the NDC-scaled screen-space gradient used for densification statistics.

What differs underneath (see include/fatesplat.h): one caller-owned workspace instead of three resizable
byte tensors, launches on torch's *current* stream (the reference uses the legacy default stream), and the
instance count R comes back through pinned memory.  Two modes:

  FATESPLAT_ASYNC=0 (default)  wait for R like the reference's blocking cudaMemcpy (rasterizer_impl.cu:281) -- but
                               only until the scan kernel has stored it into the pinned header, i.e. ~40 us into
                               the frame, so the host keeps queueing work while the frame renders; if the
                               workspace was too small the frame is re-run with a larger one, so results never
                               depend on the capacity guess.
  FATESPLAT_ASYNC=1            no host synchronisation at all; R is checked lazily (next call / backward)
                               and an overflow raises FateSplatError instead of returning a truncated frame.
"""
import ctypes as C
import os
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import FateSplatError, FsFrameInfo


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ---- workspace / capacity bookkeeping ---------------------------------------------------------------------

_ASYNC = os.environ.get("FATESPLAT_ASYNC", "0") == "1"
_capacity_hint = {}   # (device index, W, H, P) -> instances seen recently
_tile_hint = {}       # (device index, W, H, P) -> heaviest tile seen recently (fs_set_tile_hint)
_pinned = {}          # device index -> (pinned uint8 tensor viewed as FsFrameInfo slots, next slot, events)
_N_SLOTS = 64
_INFO_BYTES = C.sizeof(FsFrameInfo)
capture_headers = []  # (pinned header, hint key, capacity) of every forward recorded under CUDA-graph capture
_capture_header_pool = []


def reserve_capture_headers(n=4):
    """Pre-allocate pinned frame headers for forwards that will be recorded under CUDA-graph capture."""
    while len(_capture_header_pool) < n:
        _capture_header_pool.append(torch.zeros(_INFO_BYTES, dtype=torch.uint8).pin_memory())


def set_async(flag: bool):
    """Switch the no-host-sync mode on/off at run time (same as FATESPLAT_ASYNC)."""
    global _ASYNC
    _ASYNC = bool(flag)


def _pinned_slots(dev):
    ent = _pinned.get(dev)
    if ent is None:
        buf = torch.zeros(_N_SLOTS * _INFO_BYTES, dtype=torch.uint8).pin_memory()
        ent = {"buf": buf, "next": 0, "pending": []}
        _pinned[dev] = ent
    return ent


def _slot_info(ent, slot):
    return FsFrameInfo.from_address(ent["buf"].data_ptr() + slot * _INFO_BYTES)


def _initial_capacity(P, key):
    """Instance capacity of the next frame: 1.5x the largest recent frame of this (device, W, H, P) once one has been
    seen (the reference sizes its binning buffers from the exact count, rasterizer_impl.cu:281-287; an overflow here
    is flagged and the frame re-run, never silent), 8 instances per splat for the very first frame."""
    hint = _capacity_hint.get(key, 0)
    cap = int(hint * 1.5) + 4096 if hint else 8 * P + 65536
    return (cap + 1023) // 1024 * 1024


# ---- persistent workspace arena ---------------------------------------------------------------------------
# The reference allocates its three scratch buffers inside every call (rasterize_points.cu:68-78, resized through
# std::function callbacks).  Here a frame makes no allocator call at all: workspaces come from a grow-only pool per
# (device, stream).  A slot is handed out as a view of its backing tensor and counts as busy for as long as ANY
# tensor shares its storage -- the state returned by forward_raw, autograd's saved tensors, decode_workspace views --
# so it returns to the pool by itself when the last of them is dropped (after the backward, normally).
_arena = {}          # (device index, raw stream) -> list of uint8 backing tensors
_ARENA_MAX_SLOTS = 64
arena_stats = {"hits": 0, "allocs": 0}


def _storage_users(t):
    return torch._C._storage_Use_Count(t.untyped_storage()._cdata)


def _arena_get(nbytes, dev, di, stream_ptr):
    slots = _arena.setdefault((di, stream_ptr), [])
    small = -1
    for k, (t, idle) in enumerate(slots):
        if _storage_users(t) <= idle:
            if t.numel() >= nbytes:
                arena_stats["hits"] += 1
                return t[:nbytes] if t.numel() != nbytes else t.view(-1)
            small = k
    arena_stats["allocs"] += 1
    t = torch.empty(max(nbytes, 0) if small < 0 else max(nbytes, int(slots[small][0].numel() * 1.25)), dtype=torch.uint8,
                    device=dev)
    ent = (t, _storage_users(t))
    if small >= 0:
        slots[small] = ent  # grow-only: the undersized free slot is replaced
    elif len(slots) < _ARENA_MAX_SLOTS:
        slots.append(ent)
    return t[:nbytes] if t.numel() != nbytes else t.view(-1)


def trim_workspace_arena():
    """Drop every idle workspace slot (e.g. after densification changed P for good)."""
    for key, slots in _arena.items():
        slots[:] = [(t, idle) for t, idle in slots if _storage_users(t) > idle]


def _capacity_for_bytes(lib, P, W, H, ws):
    """Largest instance capacity whose layout fits the caller-provided workspace tensor."""
    if not (ws.is_cuda and ws.dtype == torch.uint8 and ws.is_contiguous()):
        raise FateSplatError("workspace= must be a contiguous uint8 CUDA tensor")
    base, n = lib.fs_workspace_bytes(P, W, H, 0), ws.numel()
    per = (lib.fs_workspace_bytes(P, W, H, 1 << 20) - base) / float(1 << 20)  # bytes per instance, incl. checkpoints
    cap = int(max(0, n - base - 65536) / per) // 1024 * 1024
    while cap > 0 and lib.fs_workspace_bytes(P, W, H, cap) > n:
        cap -= 1024
    if cap <= 0:
        raise FateSplatError(f"workspace= of {n} bytes is too small for P={P}, {W}x{H}")
    return cap


def _drain_pending(ent, dev, block=False):
    """Async mode: look at completed frames, update the capacity hint, raise on overflow."""
    keep = []
    for slot, ev, key in ent["pending"]:
        if block:
            ev.synchronize()
        if ev.query():
            info = _slot_info(ent, slot)
            _capacity_hint[key] = max(int(info.num_rendered), int(_capacity_hint.get(key, 0) * 0.9))
            _tile_hint[key] = max(int(info.max_tile_instances), int(_tile_hint.get(key, 0) * 0.9))
            if info.overflow:
                _capacity_hint[key] = int(info.num_rendered)
                ent["pending"] = [p for p in ent["pending"] if p[0] != slot]
                raise FateSplatError(
                    f"a frame rendered in FATESPLAT_ASYNC mode overflowed its workspace (R={info.num_rendered}); "
                    "its image/gradients are incomplete. Re-run the step (capacity has been raised) or use the "
                    "default synchronous mode.")
        else:
            keep.append((slot, ev, key))
    ent["pending"] = keep


_POISON = 0xFFFFFFFF


def _wait_num_rendered(info, dev):
    """Synchronous mode: block until this frame's instance count is known, like the reference's blocking read of
    num_rendered (rasterizer_impl.cu:281) -- but only until the scan kernel has stored it into the pinned header
    (fs_forward's early notification), not until the frame has finished, so the host keeps the stream fed."""
    spins = 0
    while info.num_rendered == _POISON:
        spins += 1
        if (spins & 1023) == 0 and torch.cuda.current_stream(dev).query():  # stream drained: it must have arrived
            if info.num_rendered == _POISON:
                raise FateSplatError("fs_forward finished without reporting num_rendered (pinned header not written)")
    return info


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _prep(t, name, dev):
    """float32 contiguous CUDA tensor on `dev`, or None for the reference's empty placeholder tensors."""
    if t is None or t.numel() == 0:
        return None
    if not t.is_cuda:
        raise FateSplatError(f"{name} must be a CUDA tensor: fateavatar_b200 has no CPU path")
    if t.device != dev:
        raise FateSplatError(f"{name} is on {t.device}, expected {dev}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")  # reference: data<float>() throws
    return t.contiguous()


def forward_raw(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                workspace=None):
    """One fs_forward call.  Returns (color, radii, state) where `state` carries the workspace the backward and
    the parity taps need.  This is the C-ABI path with device-resident tensors; the autograd Function and
    bench.py both go through it.  The workspace comes from the persistent arena (no allocator call per frame) unless
    the caller passes its own uint8 CUDA tensor as `workspace=` (then its size fixes the instance capacity)."""
    lib = _lib.load()
    if means3D.ndim != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:57-59
    if not means3D.is_cuda:
        raise FateSplatError("means3D must be a CUDA tensor: fateavatar_b200 has no CPU path")
    dev = means3D.device
    rs = raster_settings
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    m3 = _prep(means3D, "means3D", dev) if P else means3D.contiguous()
    sh_c = _prep(sh, "sh", dev)
    cp_c = _prep(colors_precomp, "colors_precomp", dev)
    op_c = _prep(opacities, "opacities", dev)
    sc_c = _prep(scales, "scales", dev)
    ro_c = _prep(rotations, "rotations", dev)
    c3_c = _prep(cov3Ds_precomp, "cov3D_precomp", dev)
    bg = _prep(rs.bg, "bg", dev)
    view = _prep(rs.viewmatrix, "viewmatrix", dev)
    proj = _prep(rs.projmatrix, "projmatrix", dev)
    campos = _prep(rs.campos, "campos", dev)
    M = 0 if sh_c is None else int(sh_c.shape[1])
    D = int(rs.sh_degree)

    color = torch.zeros((3, H, W), dtype=torch.float32, device=dev) if P == 0 else \
        torch.empty((3, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev) if P == 0 else \
        torch.empty((P,), dtype=torch.int32, device=dev)
    num_rendered = 0
    own_ws = workspace
    workspace = torch.empty(0, dtype=torch.uint8, device=dev)
    capacity = 0
    launches = 0
    if P != 0:
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (di, W, H, P)
        ent = _pinned_slots(di)
        with _lib.on_device(dev):
            stream_ptr = _lib.stream_ptr(dev)
            capturing = torch.cuda.is_current_stream_capturing()
            if _ASYNC and not capturing:
                _drain_pending(ent, di)
            capacity = _initial_capacity(P, key)
            if own_ws is not None:
                capacity = _capacity_for_bytes(lib, P, W, H, own_ws)
            lib.fs_set_tile_hint(int(_tile_hint.get(key, 0) * 1.25))
            lib.fs_set_early_notify(0 if (_ASYNC or capturing) else 1)  # nobody polls R in the no-host-sync modes
            while capturing:
                # CUDA-graph capture (fateavatar_b200.graph): the launches are recorded, not run, so nothing can be
                # waited for.  The frame gets a generous fixed capacity and its own pinned header, which every
                # replay refreshes; CapturedStep.check() reads it after a replay and raises on overflow.
                if own_ws is None:
                    capacity = max(capacity, (2 * int(_capacity_hint.get(key, 0)) + 1023) // 1024 * 1024)
                nbytes = lib.fs_workspace_bytes(P, W, H, capacity)
                # graph-private memory: allocated once at capture time, replays make no allocator call
                workspace = own_ws if own_ws is not None else torch.empty(nbytes, dtype=torch.uint8, device=dev)
                if not _capture_header_pool:
                    raise FateSplatError("CUDA-graph capture needs pinned headers reserved beforehand "
                                         "(use fateavatar_b200.graph.CapturedStep, or reserve_capture_headers())")
                header = _capture_header_pool.pop()  # pinned memory cannot be allocated while capturing
                capture_headers.append((header, key, capacity))
                rc = lib.fs_forward(P, D, M, _ptr(bg), W, H, _ptr(m3), _ptr(sh_c), _ptr(cp_c), _ptr(op_c),
                                    _ptr(sc_c), float(rs.scale_modifier), _ptr(ro_c), _ptr(c3_c), _ptr(view),
                                    _ptr(proj), _ptr(campos), float(rs.tanfovx), float(rs.tanfovy),
                                    int(bool(rs.prefiltered)), color.data_ptr(), radii.data_ptr(),
                                    workspace.data_ptr(), nbytes, capacity, header.data_ptr(), stream_ptr)
                _lib.check(rc, "fs_forward")
                launches += lib.fs_last_launch_count()
                num_rendered = -1
                break
            while not capturing:
                nbytes = lib.fs_workspace_bytes(P, W, H, capacity)
                workspace = own_ws if own_ws is not None else _arena_get(nbytes, dev, di, stream_ptr)
                slot = ent["next"]
                ent["next"] = (slot + 1) % _N_SLOTS
                if _ASYNC and any(p[0] == slot for p in ent["pending"]):
                    _drain_pending(ent, di, block=True)
                h_info = ent["buf"].data_ptr() + slot * _INFO_BYTES
                info = _slot_info(ent, slot)
                info.num_rendered = _POISON
                rc = lib.fs_forward(P, D, M, _ptr(bg), W, H, _ptr(m3), _ptr(sh_c), _ptr(cp_c), _ptr(op_c),
                                    _ptr(sc_c), float(rs.scale_modifier), _ptr(ro_c), _ptr(c3_c), _ptr(view),
                                    _ptr(proj), _ptr(campos), float(rs.tanfovx), float(rs.tanfovy),
                                    int(bool(rs.prefiltered)), color.data_ptr(), radii.data_ptr(),
                                    workspace.data_ptr(), nbytes, capacity, h_info, stream_ptr)
                _lib.check(rc, "fs_forward")
                launches += lib.fs_last_launch_count()
                if _ASYNC:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(dev))
                    ent["pending"].append((slot, ev, key))
                    num_rendered = -1
                    break
                _wait_num_rendered(info, dev)
                num_rendered = int(info.num_rendered)
                _capacity_hint[key] = max(num_rendered, int(_capacity_hint.get(key, 0) * 0.9))
                _tile_hint[key] = max(int(info.max_tile_instances), int(_tile_hint.get(key, 0) * 0.9))
                if not info.overflow:
                    break
                if own_ws is not None:
                    raise FateSplatError(f"the caller's workspace holds {capacity} instances but the frame has "
                                         f"{num_rendered}: pass a larger one (fs_workspace_bytes)")
                capacity = (int(num_rendered * 1.25) + 1023) // 1024 * 1024  # re-run, exact results
                del workspace  # the undersized slot goes back to the arena before the larger one is taken
        if rs.debug:
            torch.cuda.synchronize(dev)  # surfaces asynchronous CUDA errors like CHECK_CUDA(debug)
    empty = torch.empty(0, device=dev)
    state = dict(
        raster_settings=rs, num_rendered=num_rendered, capacity=capacity, dims=(P, D, M, H, W), launches=launches,
        workspace=workspace, radii=radii,
        tensors=(cp_c if cp_c is not None else empty, m3, sc_c if sc_c is not None else empty,
                 ro_c if ro_c is not None else empty, c3_c if c3_c is not None else empty, radii,
                 sh_c if sh_c is not None else empty, workspace, bg, view, proj, campos),
    )
    return color, radii, state


def backward_raw(state, grad_out_color, out=None):
    """One fs_backward call for a state returned by forward_raw.  Returns the reference's 8 gradient tensors
    (means2D, colors, opacity, means3D, cov3D, sh, scales, rotations) as in rasterize_points.cu:195.
    `out` may map any of means3D/means2D/colors/opacity/cov3D/sh/scales/rotations to preallocated contiguous
    float32 tensors (e.g. views of one flat all-reduce bucket) that receive the gradients in place."""
    lib = _lib.load()
    rs = state["raster_settings"]
    P, D, M, H, W = state["dims"]
    cp_c, m3, sc_c, ro_c, c3_c, radii, sh_c, workspace, bg, view, proj, campos = state["tensors"]
    dev = m3.device
    opts = dict(dtype=torch.float32, device=dev)
    alloc = torch.zeros if P == 0 else torch.empty  # every element is written by fs_backward when P > 0
    out = out or {}

    def buf(name, shape):
        t = out.get(name)
        if t is None:
            return alloc(shape, **opts)
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == int(torch.Size(shape).numel()), name
        return t.view(shape)

    g_means3D = buf("means3D", (P, 3))
    g_means2D = buf("means2D", (P, 3))
    g_colors = buf("colors", (P, 3))
    g_opacity = buf("opacity", (P, 1))
    g_cov3D = buf("cov3D", (P, 6))
    g_sh = buf("sh", (P, M, 3))
    g_scales = buf("scales", (P, 3))
    g_rot = buf("rotations", (P, 4))
    if P != 0:
        if grad_out_color.dtype != torch.float32:
            grad_out_color = grad_out_color.float()
        dpix = grad_out_color.contiguous()
        with _lib.on_device(dev):
            if _ASYNC and not torch.cuda.is_current_stream_capturing():
                di = dev.index if dev.index is not None else torch.cuda.current_device()
                _drain_pending(_pinned_slots(di), di, block=True)  # the frame must not have overflowed
            rc = lib.fs_backward(P, D, M, _ptr(bg), W, H, _ptr(m3), _ptr(sh_c), _ptr(cp_c), _ptr(sc_c),
                                 float(rs.scale_modifier), _ptr(ro_c), _ptr(c3_c), _ptr(view), _ptr(proj),
                                 _ptr(campos), float(rs.tanfovx), float(rs.tanfovy), radii.data_ptr(),
                                 workspace.data_ptr(), workspace.numel(), state["capacity"], dpix.data_ptr(),
                                 g_means2D.data_ptr(), g_opacity.data_ptr(), g_colors.data_ptr(),
                                 g_means3D.data_ptr(), g_cov3D.data_ptr(), _ptr(g_sh), g_scales.data_ptr(),
                                 g_rot.data_ptr(), _lib.stream_ptr(dev))
            _lib.check(rc, "fs_backward")
            state["launches_bwd"] = lib.fs_last_launch_count()
        if rs.debug:
            torch.cuda.synchronize(dev)
    return g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        color, radii, state = forward_raw(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations,
                                          cov3Ds_precomp)
        ctx.state = {k: v for k, v in state.items() if k not in ("tensors", "workspace", "radii")}
        ctx.num_rendered = state["num_rendered"]
        ctx.save_for_backward(*state["tensors"])
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)  # no zero-filled "gradient" of radii on every backward
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        if grad_out_color is None:
            return (None,) * 9
        state = dict(ctx.state)
        state["tensors"] = ctx.saved_tensors
        P, D, M, H, W = state["dims"]
        cp_c, _m3, sc_c, ro_c, c3_c = state["tensors"][:5]
        g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rot = backward_raw(state,
                                                                                                  grad_out_color)
        # reference order (DGR __init__.py:143-153)
        return (g_means3D, g_means2D, g_sh if M > 0 else None, g_colors if cp_c.numel() else None, g_opacity,
                g_scales if sc_c.numel() else None, g_rot if ro_c.numel() else None,
                g_cov3D if c3_c.numel() else None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # boolean mask of points passing the reference's near-plane test (DGR __init__.py:176-185)
        with torch.no_grad():
            rs = self.raster_settings
            if not positions.is_cuda:
                raise FateSplatError("positions must be a CUDA tensor: fateavatar_b200 has no CPU path")
            pos = positions.contiguous().float()
            P = pos.shape[0]
            present = torch.zeros((P,), dtype=torch.uint8, device=pos.device)
            if P:
                view = _prep(rs.viewmatrix, "viewmatrix", pos.device)
                proj = _prep(rs.projmatrix, "projmatrix", pos.device)
                if view is None or proj is None or view.numel() != 16 or proj.numel() != 16:
                    raise FateSplatError("markVisible needs 4x4 viewmatrix / projmatrix in the raster settings")
                with _lib.on_device(pos.device):
                    rc = _lib.load().fs_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                     present.data_ptr(), _lib.stream_ptr(pos.device))
                    _lib.check(rc, "fs_mark_visible")
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   raster_settings)


# ---- parity taps ------------------------------------------------------------------------------------------

def decode_workspace(workspace, P, W, H, capacity, num_rendered=None):
    """Named views into a forward workspace: the analogue of decoding the reference's geom/binning/img buffers
    (DGR rasterizer_impl.cu:155-194).  Used by tests and smoke() for bit-exact tile/index comparisons."""
    L = _lib.workspace_layout(P, W, H, capacity)
    Tn = ((W + 15) // 16) * ((H + 15) // 16)

    def view(off, nbytes, dtype, shape):
        return workspace[off:off + nbytes].view(dtype).view(shape)

    info = view(L.info, 32, torch.int32, (8,))
    R = int(info[0].item()) if num_rendered is None or num_rendered < 0 else int(num_rendered)
    R = min(R, int(capacity))  # an overflowed frame holds at most `capacity` instances
    splat = view(L.splat, P * 48, torch.float32, (P, 12))
    out = dict(
        num_rendered=R, info=info,
        depths=view(L.depths, P * 4, torch.float32, (P,)),
        cov3D=view(L.cov3D, P * 24, torch.float32, (P, 6)),
        means2D=splat[:, 0:2], extent=splat[:, 2:4], conic_opacity=splat[:, 4:8], rgb=splat[:, 8:11],
        clamped=view(L.clamped, P * 4, torch.uint8, (P, 4))[:, :3],
        rect=view(L.rect, P * 8, torch.int16, (P, 4)),
        tiles_touched=view(L.tiles_touched, P * 4, torch.int32, (P,)),
        tile_count=view(L.tile_count, Tn * 128, torch.int32, (Tn, 32))[:, 0],  # counters are 128 B apart
        ranges=view(L.ranges, Tn * 8, torch.int32, (Tn, 2)),
        point_list=view(L.point_list, R * 4, torch.int32, (R,)),
        inst_splat=view(L.inst_splat, R * 48, torch.float32, (R, 12)),
        final_T=view(L.final_T, W * H * 4, torch.float32, (H, W)),
        n_contrib=view(L.n_contrib, W * H * 4, torch.int32, (H, W)),
        # work counters of the blend kernels (common.cuh FS_WORK_*): [0] (block, instance) pairs behind the forward's
        # box cull; after a backward: [0] (block, instance) pairs it took in, [1] blended (pixel, instance) pairs
        work_forward=view(L.info + 32 + 1024, 64, torch.int32, (16,)),
        work_backward=view(L.bwd_counter + 16, 8, torch.int32, (2,)),
    )
    return out
