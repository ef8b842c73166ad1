#!/usr/bin/env python3
"""Measure pinned-memory D2H / H2D bandwidth alone and concurrently (context for the e2e number)."""
import time
import torch

n = 1807220736
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(163920896, dtype=torch.uint8).pin_memory()
d2 = torch.empty(163920896, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def d2h():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


def both():
    d2h()
    h2d()


def d2h_chunks():
    with torch.cuda.stream(s1):
        k = 8
        c = n // k
        for i in range(k):
            h[i * c:(i + 1) * c].copy_(d[i * c:(i + 1) * c], non_blocking=True)


t = run(d2h)
print("D2H 1.8 GB: %.2f ms  %.1f GB/s" % (t * 1e3, n / t / 1e9))
t = run(d2h_chunks)
print("D2H 8 chunks: %.2f ms  %.1f GB/s" % (t * 1e3, n / t / 1e9))
t = run(h2d)
print("H2D 164 MB: %.2f ms  %.1f GB/s" % (t * 1e3, 163920896 / t / 1e9))
t = run(both)
print("both concurrently: %.2f ms" % (t * 1e3))
