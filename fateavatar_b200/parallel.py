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
                          in size, so the draws come from a generator seeded identically on all ranks.
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
