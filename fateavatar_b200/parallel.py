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
                          in size, so the draws come from a generator seeded identically on all ranks;
  * uv_densify            `_uv_densify` itself with that generator as an argument.
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


# ---- densification with rank-identical draws ------------------------------------------------------------------------
_ATTR_OF_GROUP = {"opacity": "_opacity", "offset": "_offset", "color": "_features_dc", "rotation": "_rotation",
                  "scaling": "_scaling"}


def _grow_param_groups(optimizer, new_rows):
    """Append `new_rows[name]` to the single parameter of every named group, growing Adam's moments with zeros, and
    return {name: new Parameter}.  (What model/fateavatar.py:637-660 does inline.)"""
    out = {}
    for group in optimizer.param_groups:
        if len(group["params"]) != 1:
            raise ValueError("every splat parameter group holds exactly one tensor")
        old, rows = group["params"][0], new_rows[group["name"]]
        state = optimizer.state.pop(old, None)
        grown = torch.nn.Parameter(torch.cat((old.detach(), rows), dim=0).requires_grad_(True))
        if state is not None:
            for key in ("exp_avg", "exp_avg_sq"):
                state[key] = torch.cat((state[key], torch.zeros_like(rows)), dim=0)
            optimizer.state[grown] = state
        group["params"][0] = grown
        out[group["name"]] = grown
    return out


def uv_densify(model, gs_optimizer, increase_num=1000, generator=None):
    """`FateAvatar._uv_densify` (model/fateavatar.py:610-670) with the random draws taken from `generator`:
    parents ~ multinomial(xyz_gradient_accum) with replacement, new splats copy the parent's opacity / offset / colour
    / rotation, get 0.75x its scale, sit on the parent's face at fresh random barycentrics; Adam moments grow by
    zeros and the densification statistics restart.  With `synced_generator(...)` (and statistics summed over ranks
    by `allreduce_densify_stats`) every rank adds the same splats, so replicas stay identical."""
    dev = model.face_index.device
    parents = model.xyz_gradient_accum.squeeze(1).multinomial(increase_num, replacement=True, generator=generator)
    new_faces = model.face_index[parents]
    uvw = torch.rand((increase_num, 3), device=dev, generator=generator)
    new_bary = uvw / uvw.sum(dim=-1, keepdim=True)
    rows = {name: getattr(model, attr)[parents].detach() for name, attr in _ATTR_OF_GROUP.items() if name != "scaling"}
    rows["scaling"] = torch.log(torch.exp(model._scaling[parents].detach()) * 0.75)
    grown = _grow_param_groups(gs_optimizer, rows)
    for name, attr in _ATTR_OF_GROUP.items():
        setattr(model, attr, grown[name])
    model.face_index = torch.cat([model.face_index, new_faces], dim=0)
    model.bary_coords = torch.cat([model.bary_coords, new_bary], dim=0)
    if hasattr(model, "sample_flag"):
        model.sample_flag = torch.cat([model.sample_flag, torch.ones(increase_num, device=dev)], dim=0)
    n = model.num_points = model.bary_coords.shape[0]
    model.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
    model.denom = torch.zeros((n, 1), device=dev)
    model.max_radii2D = torch.zeros((n,), device=dev)
    return parents
