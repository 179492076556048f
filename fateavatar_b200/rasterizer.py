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
_capacity_hint = {}   # (device index, W, H) -> instances seen recently
_tile_hint = {}       # (device index, W, H) -> heaviest tile seen recently (fs_set_tile_hint)
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


def _initial_capacity(P, W, H, dev):
    hint = _capacity_hint.get((dev, W, H), 0)
    cap = max(8 * P + 65536, int(hint * 1.5) + 4096)
    return (cap + 1023) // 1024 * 1024


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


def forward_raw(raster_settings, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp):
    """One fs_forward call.  Returns (color, radii, state) where `state` carries the workspace the backward and
    the parity taps need.  This is the C-ABI path with device-resident tensors; the autograd Function and
    bench.py both go through it."""
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
    workspace = torch.empty(0, dtype=torch.uint8, device=dev)
    capacity = 0
    launches = 0
    if P != 0:
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (di, W, H)
        ent = _pinned_slots(di)
        with _lib.on_device(dev):
            stream_ptr = _lib.stream_ptr(dev)
            capturing = torch.cuda.is_current_stream_capturing()
            if _ASYNC and not capturing:
                _drain_pending(ent, di)
            capacity = _initial_capacity(P, W, H, di)
            lib.fs_set_tile_hint(int(_tile_hint.get(key, 0) * 1.25))
            while capturing:
                # CUDA-graph capture (fateavatar_b200.graph): the launches are recorded, not run, so nothing can be
                # waited for.  The frame gets a generous fixed capacity and its own pinned header, which every
                # replay refreshes; CapturedStep.check() reads it after a replay and raises on overflow.
                capacity = max(capacity, (2 * int(_capacity_hint.get(key, 0)) + 1023) // 1024 * 1024)
                nbytes = lib.fs_workspace_bytes(P, W, H, capacity)
                workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
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
                workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
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
                capacity = (int(num_rendered * 1.25) + 1023) // 1024 * 1024  # re-run, exact results
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
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
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
                with _lib.on_device(pos.device):
                    rc = _lib.load().fs_mark_visible(P, pos.data_ptr(), rs.viewmatrix.contiguous().data_ptr(),
                                                     rs.projmatrix.contiguous().data_ptr(), present.data_ptr(),
                                                     _lib.stream_ptr(pos.device))
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
    )
    return out
