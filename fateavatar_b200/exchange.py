"""Peer-memory gradient exchange for frame-sharded (data-parallel) training: one kernel over NVLink / NVSwitch.

    bucket = SymmetricBucket(n_floats, device)       # collective: allocates + rendezvous on the default group
    for step in ...:
        b = bucket.local(step)                       # this step's fill target (two buffers used alternately)
        ... kernels write this rank's gradients into views of b ...
        summed = bucket.all_reduce(step)             # one barrier + fs_p2p_allreduce -> rank-local tensor

The buckets live in torch's symmetric memory (CUDA VMM, peer-mapped; multicast-bound when the box has
NVSwitch/NVLS); `fs_p2p_allreduce` reads the sum of the N copies with multimem.ld_reduce (or N unicast peer loads).
Synchronisation is one symmetric-memory signal-pad barrier per step: it proves every rank has finished writing the
bucket of step s, and -- because consecutive steps use different buffers -- by the time a rank refills buffer
s % 2 at step s + 2 every rank has passed the barrier of step s + 1, i.e. has finished reading it.  No NCCL call
sits in the step.
"""
import os

import torch

from . import _lib
from ._lib import FateSplatError


class SymmetricBucket:
    def __init__(self, n_floats, device, group=None, use_multicast=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        if not dist.is_initialized():
            raise FateSplatError("SymmetricBucket needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.n = (int(n_floats) + 63) // 64 * 64
        self.device = torch.device(device)
        self.both = symm.empty(2 * self.n, dtype=torch.float32, device=self.device)
        self.both.zero_()
        self.hdl = symm.rendezvous(self.both, self.group)
        self.out = torch.empty(self.n, dtype=torch.float32, device=self.device)
        if use_multicast is None:
            # in-switch reduction reads n floats once per rank but each request takes the long way round; plain peer
            # loads read N copies.  Measured on B200/NVLink5 at 6 MB: N = 2 -> 23 us unicast vs 31 us multicast.
            env = os.environ.get("FATESPLAT_P2P_MULTICAST", "auto")
            use_multicast = self.world >= 4 if env == "auto" else env == "1"
        mc = int(self.hdl.multicast_ptr) if use_multicast else 0
        self.multicast_ptr = mc if mc else None
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)

    def local(self, step):
        """This rank's bucket for `step` (a view of the symmetric allocation; fill it in place)."""
        k = step & 1
        return self.both[k * self.n:(k + 1) * self.n]

    def all_reduce(self, step, n=None):
        """Sum over ranks of local(step)[:n] -> out[:n] (rank-local).  Stream-ordered on torch's current stream."""
        n = self.n if n is None else (int(n) + 3) // 4 * 4
        with torch.cuda.device(self.device):
            self.hdl.barrier(channel=0)  # every rank's bucket of this step is complete and visible
            rc = _lib.load().fs_p2p_allreduce(self.world, self.multicast_ptr, self.peer_ptrs_dev, (step & 1) * self.n, n,
                                              self.out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(rc, "fs_p2p_allreduce")
        return self.out[:n]
