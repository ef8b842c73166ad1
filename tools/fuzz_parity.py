#!/usr/bin/env python3
"""Exploratory: damaged streams through the GPU path and the oracle; report where they differ.
usage: fuzz_parity.py [n_variants_per_stream]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

nvar = int(sys.argv[1]) if len(sys.argv) > 1 else 6
rng = np.random.default_rng(1234)
base = synth.make_workload("cfg3", 6, 24) + synth.make_workload("cfg4", 9, 24) + [
    synth.make_stream(layer=2, bitrate_kbps=192, nframes=10, seed=5)]
bad, what = [], []
for k, s in enumerate(base):
    a = np.frombuffer(s, np.uint8).copy()
    for v in range(nvar):
        b = a.copy()
        n = int(rng.integers(1, 12))
        idx = rng.integers(0, b.size, n)
        b[idx] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
        bad.append(b.tobytes())
        what.append((k, "flip%d" % n))
    bad.append(a[: rng.integers(1, a.size)].tobytes()); what.append((k, "trunc"))
    bad.append(a[rng.integers(1, 500):].tobytes()); what.append((k, "head"))
with mp3_b200.Decoder(device=0, pcm_format=mp3_b200.PCM_F32) as dec:
    dec.decode_batch(bad)
    arena = dec.fetch_pcm()
    nbad = 0
    for i, s in enumerate(bad):
        r = oracle.decode(s)
        inf = dec.stream_info(i)
        if r.rc != 0:
            ok = inf.frames == 0
            msg = "oracle finds no stream; gpu frames %d" % inf.frames
        elif (inf.frames, inf.samples, inf.channels) != (r.frames, r.samples, r.channels):
            ok = False
            msg = "shape gpu %s oracle %s" % ((inf.frames, inf.samples, inf.channels), (r.frames, r.samples, r.channels))
        else:
            got = dec.stream_pcm(i, arena).astype(np.float64)
            d = got - r.pcm.T
            sc = max(1.0, np.abs(r.pcm).max())
            mx = np.abs(d).max() / sc if d.size else 0.0
            ok = mx <= 2.0 ** -14
            msg = "max err %.3g (scale %.3g) first bad sample %s" % (mx, sc, int(np.argmax(np.abs(d).max(axis=1) > 2.0 ** -14 * sc)) if not ok else -1)
        if not ok:
            nbad += 1
            print("MISMATCH stream %d %s: %s" % (i, what[i], msg))
    print("%d / %d damaged streams differ" % (nbad, len(bad)))
