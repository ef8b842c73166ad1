#!/usr/bin/env python3
"""Small decode (all stream families + corrupted input, both pipelines, both indexers) to run under
`compute-sanitizer --tool memcheck|racecheck|initcheck` on the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases  # noqa: E402
import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

allc = dict(cases.FF)
allc.update(cases.EXTRA)
names = sorted(allc)[::3]
streams = [synth.make_stream(**dict(allc[n], nframes=6)) for n in names]
rng = np.random.default_rng(5)
for s in list(streams[:6]):
    a = np.frombuffer(s, np.uint8).copy()
    idx = rng.integers(0, a.size, 30)
    a[idx] ^= 0x55
    streams.append(a.tobytes())
    streams.append(s[: len(s) // 2 + 3])
streams += [b"", b"\xff\xfb\x90\x00" * 50]
for pipe in (mp3_b200.PIPE_FUSED, mp3_b200.PIPE_STAGED):
    for idx in (mp3_b200.INDEX_DEVICE, mp3_b200.INDEX_HOST):
        for fmt in (mp3_b200.PCM_S16, mp3_b200.PCM_F32):
            with mp3_b200.Decoder(device=0, pcm_format=fmt, pipeline=pipe, indexer=idx) as dec:
                dec.decode_batch(streams)
                a = dec.fetch_pcm()
                print(pipe, idx, fmt, a.size, float(np.abs(a.astype(np.float64)).sum()))
print("done")
