"""graph.CapturedStep plumbing on the GPU: the input prefetch must never overwrite inputs a running replay still reads."""
import pytest
import torch

from fateavatar_b200 import graph as fgraph

pytestmark = pytest.mark.gpu


def test_prefetch_is_ordered_behind_the_previous_replay_of_the_same_recording(cuda_device):
    dev = cuda_device
    big = torch.randn(2048, 2048, device=dev)

    def frame(inp):  # a few ms of unrelated work first, the input is read late
        z = big
        for _ in range(20):
            z = (z @ big) * 1e-3
        return {"y": inp["x"] * 2.0 + z.sum() * 0.0}

    host = [torch.full((1 << 20,), float(i)).pin_memory() for i in range(12)]
    rec = [fgraph.CapturedStep(frame, {"x": host[0].to(dev)}, params=(), warmup=1, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    rec[0].prefetch({"x": host[0]}, copy_stream)
    seen = []
    for i in range(10):
        r = rec[i & 1]
        r(None)                                                   # replay step i on its staged input
        rec[(i + 1) & 1].prefetch({"x": host[i + 1]}, copy_stream)  # step i-1 (same recording) may still be running
        if i >= 1:
            out = rec[(i - 1) & 1].wait()
            seen.append((i - 1, float(out["y"][0]), float(out["y"][-1])))
    for i, first, last in seen:
        assert first == 2.0 * i and last == 2.0 * i, seen
