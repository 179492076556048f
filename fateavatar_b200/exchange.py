"""Peer-memory gradient exchange for frame-sharded (data-parallel) training: one kernel over NVLink / NVSwitch.

    bucket = SymmetricBucket(n_floats, device)       # collective: allocates + rendezvous on the default group
    for step in ...:
        b = bucket.local(step)                       # this step's fill target (two buffers used alternately)
        ... kernels write this rank's gradients into views of b ...
        summed = bucket.all_reduce(step)             # one barrier + fs_p2p_allreduce -> rank-local tensor

The buckets live in torch's symmetric memory (CUDA VMM, peer-mapped; multicast-bound when the box has
NVSwitch/NVLS); `fs_p2p_allreduce` reads the sum of the N copies with multimem.ld_reduce (or N unicast peer loads).
From 4 ranks on the two-shot form takes over (fs_p2p_reduce_scatter_bcast: each rank reduces its 1/N slice in the
switch and multicast-stores it to all ranks, n/N floats per rank on the wire, one more barrier).
Synchronisation of the one-shot form is one symmetric-memory signal-pad barrier per step: it proves every rank has finished writing the
bucket of step s, and -- because consecutive steps use different buffers -- by the time a rank refills buffer
s % 2 at step s + 2 every rank has passed the barrier of step s + 1, i.e. has finished reading it.  No NCCL call
sits in the step.
"""
import os

import torch

from . import _lib
from ._lib import FateSplatError


class SymmetricBucket:
    """algo: "one_shot" (every rank reads the sum of all copies: unicast peer loads or, `use_multicast`, in-switch
    reduction), "two_shot" (each rank reduces its 1/N slice in the switch and multicast-stores it to everybody; two
    barriers, n/N floats per rank on the wire) or None = pick by world size."""

    def __init__(self, n_floats, device, group=None, use_multicast=None, algo=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        if not dist.is_initialized():
            raise FateSplatError("SymmetricBucket needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n = (int(n_floats) + 63) // 64 * 64
        self.device = torch.device(device)
        # [ input buffer of even steps | input buffer of odd steps | output of the two-shot algorithm ]
        self.both = symm.empty(3 * self.n, dtype=torch.float32, device=self.device)
        self.both.zero_()
        self.hdl = symm.rendezvous(self.both, self.group)
        self.out_local = torch.empty(self.n, dtype=torch.float32, device=self.device)
        mc_base = int(self.hdl.multicast_ptr)
        if algo is None:
            # measured on B200 / NVLink 5 at 6.2 MB: N = 2: 22 us one-shot unicast, 31 us one-shot multicast, 37 us
            # two-shot; N = 8: 84 / 83 us one-shot, 33 us two-shot (NCCL: 39 / 58 us) -> two-shot from 4 ranks
            algo = os.environ.get("FATESPLAT_P2P_ALGO", "auto")
            if algo == "auto":
                algo = "two_shot" if (self.world >= 4 and mc_base) else "one_shot"
        if algo == "two_shot" and not mc_base:
            raise FateSplatError("the two-shot exchange needs NVLink multicast (NVSwitch/NVLS), not available here")
        self.algo = algo
        if use_multicast is None:
            env = os.environ.get("FATESPLAT_P2P_MULTICAST", "auto")
            use_multicast = self.world >= 4 if env == "auto" else env == "1"
        self.multicast_base = mc_base if mc_base else None
        self.multicast_ptr = mc_base if (mc_base and use_multicast) else None
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)

    def local(self, step):
        """This rank's bucket for `step` (a view of the symmetric allocation; fill it in place)."""
        k = step & 1
        return self.both[k * self.n:(k + 1) * self.n]

    def all_reduce(self, step, n=None):
        """Sum over ranks of local(step)[:n] -> a rank-local tensor of n floats.  Stream-ordered on torch's current
        stream."""
        n = self.n if n is None else (int(n) + 3) // 4 * 4
        lib = _lib.load()
        with _lib.on_device(self.device):
            stream = _lib.stream_ptr(self.device)
            self.hdl.barrier(channel=0)  # every rank's bucket of this step is complete and visible
            if self.algo == "two_shot":
                rc = lib.fs_p2p_reduce_scatter_bcast(self.world, self.rank, self.multicast_base + 4 * (step & 1) * self.n,
                                                     self.multicast_base + 4 * 2 * self.n, n, stream)
                _lib.check(rc, "fs_p2p_reduce_scatter_bcast")
                self.hdl.barrier(channel=1)  # every rank's slice has landed in everybody's output region
                return self.both[2 * self.n:2 * self.n + n]
            rc = lib.fs_p2p_allreduce(self.world, self.multicast_ptr, self.peer_ptrs_dev, (step & 1) * self.n, n,
                                      self.out_local.data_ptr(), stream)
            _lib.check(rc, "fs_p2p_allreduce")
        return self.out_local[:n]


class StepExchange:
    """The per-step exchange of frame-sharded training as ONE kernel (fs_p2p_exchange): cross-rank barrier, all-reduce
    of the splat part of the bucket and expansion of the N FLAME factor records into the dense delta gradients.

        ex = StepExchange(n_splat, rec_floats, device)       # collective (symmetric allocation + rendezvous)
        b  = ex.fill(step)                                   # this step's bucket: [n_splat | world x rec_stride]
        ... kernels write this rank's splat gradients into b[:n_splat] and its factor record into ex.record(step) ...
        summed = ex.exchange(step, flame_dims, delta_out)    # one launch (+ fs_p2p_wait for the two-shot form)

    Consecutive steps alternate between two input buckets, which makes the single in-kernel barrier sufficient (see
    the module docstring).  No host synchronisation, no NCCL call, CUDA-graph capturable."""

    ALGOS = {"one_shot": 0, "one_shot_mc": 1, "two_shot": 2}

    def __init__(self, n_splat, rec_floats, device, group=None, algo=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        if not dist.is_initialized():
            raise FateSplatError("StepExchange needs an initialised torch.distributed process group")
        lib = _lib.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = torch.device(device)
        pad = lambda n, a: (int(n) + a - 1) // a * a
        self.n_splat, self.rec_stride = pad(n_splat, 64), pad(rec_floats, 64)
        self.bucket = self.n_splat + self.world * self.rec_stride
        self.out_off = 2 * self.bucket
        self.flags_off = self.out_off + self.n_splat
        self.gather_off = self.flags_off + pad(lib.fs_p2p_exchange_flag_floats(), 64)
        total = self.gather_off + self.world * self.rec_stride
        self.mem = symm.empty(total, dtype=torch.float32, device=self.device)
        self.mem.zero_()
        self.hdl = symm.rendezvous(self.mem, self.group)
        mc = int(self.hdl.multicast_ptr)
        if algo is None:
            algo = os.environ.get("FATESPLAT_P2P_ALGO", "auto")
        if algo == "auto":  # measured on B200 / NVLink 5: in-switch two-shot wins from 4 ranks
            algo = "two_shot" if (self.world >= 4 and mc) else "one_shot"
        if algo not in self.ALGOS:
            raise FateSplatError(f"unknown exchange algorithm {algo!r} (one of {sorted(self.ALGOS)})")
        if algo != "one_shot" and not mc:
            raise FateSplatError(f"{algo} needs NVLink multicast (NVSwitch/NVLS), not available here")
        self.algo = algo
        self.multicast = mc or None
        self.peer_ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.out_local = torch.empty(self.n_splat, dtype=torch.float32, device=self.device)
        torch.cuda.synchronize(self.device)
        self.hdl.barrier(channel=0)  # every rank's flags / buckets are zero before anybody signals

    def fill(self, step):
        k = step & 1
        return self.mem[k * self.bucket:(k + 1) * self.bucket]

    def record(self, step):
        """This rank's slot among the N factor records of `step`'s bucket."""
        o = (step & 1) * self.bucket + self.n_splat + self.rank * self.rec_stride
        return self.mem[o:o + self.rec_stride]

    def exchange(self, step, flame_dims=None, delta_out=(None, None, None), scale=1.0):
        """flame_dims = (V, L, l0, NP); delta_out = (d_delta_vertex [V,3], d_delta_shapedirs [V,3,L], d_delta_posedirs
        [NP,3V]) receive the summed dense gradients.  Returns the summed splat part (n_splat floats, rank-local)."""
        lib = _lib.load()
        V, L, l0, NP = flame_dims if flame_dims is not None else (0, 0, 0, 0)
        p = lambda t: None if t is None else t.data_ptr()
        a = self.ALGOS[self.algo]
        with _lib.on_device(self.device):
            st = _lib.stream_ptr(self.device)
            rc = lib.fs_p2p_exchange(self.world, self.rank, a, self.peer_ptrs_dev, self.multicast, self.mem.data_ptr(),
                                     (step & 1) * self.bucket, self.n_splat, self.n_splat, self.rec_stride, self.out_off,
                                     self.flags_off, self.gather_off, None if a == 2 else self.out_local.data_ptr(),
                                     V, L, l0, NP,
                                     float(scale), p(delta_out[0]), p(delta_out[1]), p(delta_out[2]), st)
            _lib.check(rc, "fs_p2p_exchange")
            if a == 2:
                _lib.check(lib.fs_p2p_wait(self.world, self.mem.data_ptr(), self.flags_off, st), "fs_p2p_wait")
                return self.mem[self.out_off:self.out_off + self.n_splat]
        return self.out_local

    def timing(self, reset=True):
        """Device-side clock of the exchange kernel since the last reset: {"wait_us": mean time this rank spent in the
        barrier waiting for its peers, "work_us": mean time from the barrier to the kernel's end, "calls"}."""
        import ctypes as C

        w, k, n = C.c_double(), C.c_double(), C.c_int()
        with _lib.on_device(self.device):
            rc = _lib.load().fs_p2p_exchange_timing(self.mem.data_ptr(), self.flags_off, C.byref(w), C.byref(k),
                                                    C.byref(n), int(reset))
        _lib.check(rc, "fs_p2p_exchange_timing")
        return {"wait_us": w.value / 1e3, "work_us": k.value / 1e3, "calls": n.value}

    def summed(self):
        """Where exchange() leaves the summed splat part (fixed address: valid to alias before the first call)."""
        return self.mem[self.out_off:self.out_off + self.n_splat] if self.algo == "two_shot" else self.out_local
