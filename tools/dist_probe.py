#!/usr/bin/env python
"""torchrun probe: cost of the collectives the frame-sharded step uses, alone and back to back with compute."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
res = {}
def timeit(fn, n=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return round(1000 * e0.elapsed_time(e1) / n, 1), round(1e6 * (t1 - t0) / n, 1)
for mb in (0.25, 1.0, 6.2, 32.0):
    buf = torch.zeros(int(mb * 1e6 / 4), device=dev)
    res[f"all_reduce_{mb}MB_us(gpu,host)"] = timeit(lambda: dist.all_reduce(buf))
a = torch.zeros(1 << 20, device=dev)
def compute_then_ar():
    for _ in range(10): a.add_(1.0)
    dist.all_reduce(buf6)
buf6 = torch.zeros(int(6.2e6 / 4), device=dev)
res["10_small_kernels_us"] = timeit(lambda: [a.add_(1.0) for _ in range(10)])
res["10_small_kernels+all_reduce_6.2MB_us"] = timeit(compute_then_ar)
if rank == 0: print(json.dumps(res, indent=1))
dist.destroy_process_group()
