"""Whole-frame CUDA-graph capture for training loops built on the fateavatar_b200 operators.

One FateAvatar frame is ~17 kernel launches of this library plus a few dozen small torch kernels (loss, autograd
glue); driven from Python the frame is host-bound (~0.8 ms of interpreter/autograd time for ~0.4 ms of GPU work).
`CapturedStep` records the frame once -- forward, loss, backward, result read-back -- into a CUDA graph and replays
it with a single launch, the B200-native alternative to a tracing compiler:

    def frame(inp):                                   # ordinary code: flame_lbs, pose_splats, GaussianRasterizer, loss
        ...; loss.backward(); return {"loss": loss.detach(), "image": img.detach()}
    step = CapturedStep(frame, example_inputs, params=[...all leaf parameters...])
    for batch in loader:                              # batch: dict of (pinned) host tensors with the example's shapes
        out = step(batch)                             # H2D copies + one graph launch + D2H into pinned outputs
        step.wait(); optimizer.step()                 # .grad of every parameter is refreshed in place by the replay
                                                      # (keep them attached: no zero_grad(set_to_none=True))

Requirements while captured: static shapes (re-capture after densification changes P), no host synchronisation
inside `frame` (the rasterizer runs in its no-host-sync mode with a fixed instance capacity of 2x the largest
frame seen during warm-up; `check()` raises FateSplatError if a replayed frame overflowed it).
"""
import torch

from . import rasterizer as _R
from ._lib import FateSplatError, FsFrameInfo


class CapturedStep:
    def __init__(self, frame_fn, example_inputs, params=(), warmup=3, device=None, on_capture=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self._staged = None
        self._done = torch.cuda.Event()
        self.params = list(params)
        self.static_in = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in example_inputs.items()}
        for k, v in example_inputs.items():
            self.static_in[k].copy_(v)
        # eager warm-up on a side stream (the documented torch recipe): sizes the rasterizer's capacity / tile hints
        was_async = _R._ASYNC
        _R.set_async(False)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                for p in self.params:
                    p.grad = None
                out = frame_fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        for p in self.params:
            p.grad = None
        # pinned memory cannot be allocated while capturing: outputs and frame headers are reserved now
        self.host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
        del out
        _R.reserve_capture_headers(8)
        _R.set_async(True)
        n0 = len(_R.capture_headers)
        self.graph = torch.cuda.CUDAGraph()
        if on_capture is not None:
            on_capture()  # e.g. enable launches that must not run during the eager warm-up (an optimiser step)
        try:
            with torch.cuda.graph(self.graph):
                out = frame_fn(self.static_in)
                self.static_out = dict(out)
                # results leave through pinned host memory inside the graph: one launch covers the read-back too
                for k, v in out.items():
                    self.host_out[k].copy_(v, non_blocking=True)
        finally:
            _R.set_async(was_async)
        # the gradient tensors the replay refreshes in place: keep them attached (never set .grad to None afterwards)
        self.grads = [p.grad for p in self.params]
        self.headers = _R.capture_headers[n0:]
        del _R.capture_headers[n0:]

    def __call__(self, host_inputs=None):
        """Copy this frame's inputs in (asynchronously, from pinned host tensors), replay, return the pinned outputs
        (valid after `wait()`).  With host_inputs=None the inputs staged by `prefetch()` are used."""
        if host_inputs is not None:
            for k, v in host_inputs.items():
                self.static_in[k].copy_(v, non_blocking=True)
        elif self._staged is not None:
            torch.cuda.current_stream(self.device).wait_event(self._staged)
            self._staged = None
        self.graph.replay()
        self._done.record(torch.cuda.current_stream(self.device))
        return self.host_out

    def prefetch(self, host_inputs, stream):
        """Input prefetch (the pinned GT-frame prefetch of SURVEY 8f N3; upstream loads every frame synchronously,
        train/dataset.py:14-54): copy the NEXT frame's pinned host inputs into this step's static input tensors on
        `stream` while another recorded step is still running.  The copies are ordered on the device behind this
        recording's previous replay (its inputs may still be in use: with two alternating recordings the prefetch for
        step i+1 is issued while step i-1 can still be running), so the host never has to wait before calling this."""
        with torch.cuda.stream(stream):
            stream.wait_event(self._done)  # no-op before the first replay
            for k, v in host_inputs.items():
                self.static_in[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._staged = ev

    def wait(self):
        """Block until THIS recording's last replay has finished (later launches on the stream keep running), check it
        for workspace overflow and return its pinned outputs."""
        self._done.synchronize()
        self.check()
        return self.host_out

    def check(self):
        """After a completed replay: raise if any captured rasterizer frame overflowed its fixed capacity."""
        for header, key, capacity in self.headers:
            info = FsFrameInfo.from_address(header.data_ptr())
            if info.overflow:
                _R._capacity_hint[key] = int(info.num_rendered)
                raise FateSplatError(f"a replayed frame needed {info.num_rendered} instances but was captured with "
                                     f"capacity {capacity}: build a new CapturedStep (the hint has been raised)")

    def num_rendered(self):
        return [int(FsFrameInfo.from_address(h.data_ptr()).num_rendered) for h, _, _ in self.headers]
