#!/usr/bin/env python3
"""Exploratory: the incremental interface (open / enqueue in random pieces / decode / fetch) against the
one-shot batch decode of the same bytes, random stream shapes: the concatenated PCM must be bit-identical.
usage: fuzz_incremental.py [n_streams] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
with mp3_b200.Decoder(device=0, pcm_format=mp3_b200.PCM_F32) as dec:
    for k in range(n):
        layer = int(rng.choice([0, 0, 0, 2, 1]))
        rate = int(rng.choice([44100, 48000, 32000, 22050, 24000, 16000] + ([] if layer else [11025, 12000, 8000])))
        lsf = rate < 32000
        cfg = dict(nframes=int(rng.integers(3, 40)), seed=int(rng.integers(1, 10 ** 6)), sample_rate=rate,
                   mode=int(rng.choice([0, 1, 2, 3])), crc=int(rng.integers(0, 2)))
        if layer:
            cfg.update(layer=layer, bitrate_kbps=(64 if lsf else 192) if layer == 2 else (128 if lsf else 256))
            if layer == 2 and cfg["mode"] == 3:
                cfg["bitrate_kbps"] = 64 if lsf else 96
        else:
            cfg.update(blocks=int(rng.integers(0, 2)), mixed_pct=int(rng.choice([0, 30])), bitrate_kbps=64 if lsf else 128,
                       fill_lo_pct=int(rng.choice([30, 85])), tag=int(rng.choice([0, 0, 1])))
            if rng.integers(0, 3) == 0:
                cfg.update(vbr_min_kbps=32 if lsf else 64, vbr_max_kbps=128 if lsf else 256)
        try:
            s = synth.make_stream(**cfg)
        except ValueError:
            continue
        junk = bytes(rng.integers(0, 256, int(rng.integers(0, 300)), dtype=np.uint8)) if rng.integers(0, 4) == 0 else b""
        s = junk + s
        try:
            dec.decode_batch([s])
        except Exception as e:
            print('DECODE FAILED', cfg, 'junk', len(junk), 'bytes', len(s), e)
            raise
        whole = dec.stream_pcm(0, dec.fetch_pcm()).copy()
        h = dec.open_stream()
        pos, got = 0, []
        hi = int(rng.choice([64, 700, 5000]))
        while pos < len(s):
            m = int(rng.integers(1, hi))
            h.enqueue(s[pos: pos + m])
            pos += m
            dec.decode_streams()
            inf = h.info()
            if inf.samples:
                got.append(h.fetch(inf.samples))
        h.close()
        cat = np.concatenate(got) if got else np.zeros((0, whole.shape[1] if whole.ndim == 2 else 1), np.float32)
        ok = cat.shape == whole.shape and np.array_equal(cat, whole)
        if not ok:
            bad += 1
            print("MISMATCH", k, cfg, "junk", len(junk), cat.shape, whole.shape,
                  (float(np.abs(cat - whole).max()) if cat.shape == whole.shape else None))
print("%d / %d streams differ" % (bad, n))
