"""The N > 1 path on CPU: partition / merge logic, and a world_size-2 gloo run of the rank-side
protocol bench.py uses (shard by rank, local work, barrier, max-over-ranks of the timing, sum of
the work) -- with the oracle standing in for the GPU decode."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_and_balances():
    from mp3_b200.multi import merge, partition
    rng = np.random.default_rng(1)
    for n, k in ((1, 1), (5, 2), (100, 8), (3, 8), (1024, 4), (0, 3)):
        sizes = rng.integers(1, 1000, n)
        r = partition(sizes, k)
        assert len(r) == k and r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
        if n >= 8 * k:
            tot = [sizes[lo:hi].sum() for lo, hi in r]
            assert max(tot) - min(tot) <= 2 * sizes.max()
        parts = [list(range(lo, hi)) for lo, hi in r]
        assert merge(parts, r, n) == list(range(n))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mp3_b200 import synth
    from mp3_b200.multi import partition
    from oracle import oracle
    cfgs = [dict(nframes=4, seed=500 + i, blocks=1) for i in range(6)]
    streams = [synth.make_stream(**c) for c in cfgs]
    lo, hi = partition([len(s) for s in streams], world)[rank]
    dist.barrier()
    secs = 0.0
    checks = []
    for s in streams[lo:hi]:
        d = oracle.decode(s)
        secs += d.samples / d.sample_rate
        checks.append(float(np.abs(d.pcm).sum()))
    t = torch.tensor([0.001 * (rank + 1)], dtype=torch.float64)   # stand-in for the device time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([secs, float(hi - lo)], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    q.put((rank, float(t[0]), float(tot[0]), float(tot[1]), checks, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_gloo_protocol():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, a0, n0, c0, rng0), (r1, t1, a1, n1, c1, rng1) = res
    assert t0 == t1 == 0.002                 # max over ranks
    assert a0 == a1 and n0 == n1 == 6.0      # whole-job sums agree on every rank
    assert rng0[1] == rng1[0] and rng0[0] == 0 and rng1[1] == 6
    assert abs(a0 - 6 * 4 * 1152 / 44100.0) < 1e-9
