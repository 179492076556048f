import sys, os, traceback
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
import test_optimise_loop_gpu as T
from fateavatar_b200 import avatar, optimizer as fopt, scenes
import ref_frame_harness as H
dev = torch.device('cuda:0'); res=(96,96)
a = scenes.small_avatar(seed=4, N=2500)
cfg = dict(densify_interval=4, prune_interval=6, opacity_reset_interval=9, increase_num=300, max_points_num=3200, min_opacity=0.2)
ref, opts, patch = T._ref_model_and_optim(a, dev, res)
avatar.attach(ref)
mine = T._my_model(a, dev, res)
fov=[0.35]
def frame_loss(m, d):
    out = avatar.forward_frame(m, dict(cam_pose=d["cam_pose"], fovx=fov, fovy=fov, flame_pose=d["flame_pose"], expression=d["expression"]))
    return (out["rgb_image"][0] - d["target"]).abs().mean(), out
inp = H.frame_input(a, fovx=0.35, fovy=0.35, T=(0.0, 0.0, 1.25))
h = dict(cam_pose=inp["cam_pose"], flame_pose=inp["flame_pose"], expression=inp["expression"], target=torch.rand(3,*res))
h = {k: v.contiguous().pin_memory() for k,v in h.items()}
loop = fopt.OptimiseLoop(mine, frame_loss, {k: v.to(dev) for k,v in h.items()}, training=cfg, generator=torch.Generator(device=dev).manual_seed(77), capture=False)
for t in range(3):
    d = {k: v.to(dev) for k, v in h.items()}
    out = ref(dict(cam_pose=d["cam_pose"], fovx=fov, fovy=fov, flame_pose=d["flame_pose"], expression=d["expression"]))
    loss = (out["rgb_image"][0] - d["target"]).abs().mean()
    for o in opts.values(): o.zero_grad(set_to_none=True)
    loss.backward()
    ref._add_densification_stats(out["viewspace_points"][0], out["visibility_filter"][0])
    for o in opts.values(): o.step()
    if t % 4 == 0: ref._uv_densify(opts["gs"], increase_num=300)
    if t % 6 == 0: ref._prune_low_opacity_points(opts["gs"], min_opacity=0.2)
    try:
        loop.step(h); loop.wait()
    except Exception as e:
        traceback.print_exc()
    for n, attr, w in fopt.FIELDS:
        p = getattr(mine, attr)
        print(t, attr, tuple(p.shape), p.requires_grad, p.is_leaf, None if p.grad is None else tuple(p.grad.shape))
    print('P', loop.store.P, ref.num_points)
