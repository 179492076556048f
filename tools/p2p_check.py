#!/usr/bin/env python
"""torchrun check of the peer-memory exchange: fs_p2p_allreduce (multicast and unicast) vs NCCL all_reduce, and timing.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from fateavatar_b200.exchange import SymmetricBucket
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
res = {"world": world}
n = 1_561_000
for name, mc, algo in (("multicast", True, "one_shot"), ("unicast", False, "one_shot"), ("two_shot", True, "two_shot")):
    try:
        sb = SymmetricBucket(n, dev, use_multicast=mc, algo=algo)
        res[name + "_ptr"] = bool(sb.multicast_ptr)
        errs = []
        for it in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * it + rank)
            x = torch.randn(sb.n, device=dev, generator=g)
            sb.local(it).copy_(x)
            want = x.clone(); dist.all_reduce(want)
            got = sb.all_reduce(it)
            errs.append(float((got[:n] - want[:n]).abs().max()))
        res[name + "_max_err"] = max(errs)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(20): sb.all_reduce(i)
        e0.record()
        for i in range(200): sb.all_reduce(i)
        e1.record(); torch.cuda.synchronize()
        res[name + "_us"] = round(1000 * e0.elapsed_time(e1) / 200, 1)
    except Exception as ex:
        res[name + "_error"] = repr(ex)[:300]
buf = torch.zeros(n, device=dev)
for _ in range(20): dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize()
res["nccl_us"] = round(1000 * e0.elapsed_time(e1) / 200, 1)
if rank == 0: print(json.dumps(res))
dist.destroy_process_group()
