"""World-size-2 gloo test (CPU) of the host-side logic of the frame-sharded step: every rank owns its own
frames, gradients land in views of one flat bucket, and one all-reduce per step yields the sum over ranks'
frames.  The per-rank gradients come from the oracle here (no GPU in this container); on the GPU the same
bucket layout is filled by fs_backward (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fateavatar_b200 import scenes
from oracle import oracle as orc
from util import oracle_forward

WIDTHS = dict(means3D=3, means2D=3, sh=3, opacity=1, scales=3, rotations=4)  # bench.py bucket layout
KEYS = dict(means3D="dL_dmeans3D", means2D="dL_dmeans2D", sh="dL_dsh", opacity="dL_dopacity", scales="dL_dscales",
            rotations="dL_drotations")


def frame_grads(rank, P):
    sc = scenes.head_scene(seed=100 * rank, P=P, W=64, H=64, scale_mult=14.0)
    st = oracle_forward(orc, sc)
    dpix = np.random.default_rng(rank).standard_normal((3, 64, 64)).astype(np.float32)
    return orc.backward(st, dpix)


def _worker(rank, world, port, P, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bucket = torch.zeros(P * sum(WIDTHS.values()))
    views, off = {}, 0
    for k, w in WIDTHS.items():
        views[k] = bucket[off:off + P * w]
        off += P * w
    g = frame_grads(rank, P)
    for k, v in views.items():
        v.copy_(torch.from_numpy(g[KEYS[k]].reshape(-1)))
    dist.all_reduce(bucket)
    if rank == 0:
        out.put(bucket.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharded_gradient_bucket_world2():
    P, world = 300, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros_like(got)
    for r in range(world):
        g = frame_grads(r, P)
        off = 0
        for k, w in WIDTHS.items():
            want[off:off + P * w] += g[KEYS[k]].reshape(-1)
            off += P * w
    assert np.allclose(got, want, rtol=1e-6, atol=1e-7)
