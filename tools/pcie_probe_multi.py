#!/usr/bin/env python3
"""Platform ceiling for the multi-GPU end-to-end number: every rank copies a 1.8 GB PCM-sized buffer from its
GPU into its own pinned host buffer at the same time (and a 164 MB upload the other way); prints per-rank
and aggregate GB/s.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe_multi.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mp3_b200 import multi  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if "--numa" in sys.argv:
    multi.bind_to_gpu_numa(local)
dist.init_process_group("gloo", rank=rank, world_size=world)
n, nu = 1807220736, 163920896
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
hu = torch.empty(nu, dtype=torch.uint8).pin_memory()
du = torch.empty(nu, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(with_upload, reps=6):
    def go():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        if with_upload:
            with torch.cuda.stream(s2):
                du.copy_(hu, non_blocking=True)
    go()
    torch.cuda.synchronize()
    dist.barrier()
    t = time.perf_counter()
    for _ in range(reps):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    out = [None] * world
    dist.all_gather_object(out, dt)
    return out


for up in (False, True):
    ts = timed(up)
    if rank == 0:
        per = [n / t / 1e9 for t in ts]
        print("N=%d D2H 1.8 GB per rank%s: per-rank GB/s %s | aggregate %.1f GB/s (slowest rank %.1f ms)" % (
            world, " + 164 MB H2D" if up else "", " ".join("%.1f" % p for p in per), world * n / max(ts) / 1e9, max(ts) * 1e3), flush=True)
dist.destroy_process_group()
