#!/usr/bin/env python3
"""Exploratory: the exact output operators (planar copy, sentence boundaries, window energies) on a batch of
random stream shapes against their numpy definitions, both PCM formats, gapless on and off.
usage: fuzz_ops.py [n_streams] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402
from oracle import segments as seg_oracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
streams = []
while len(streams) < n:
    layer = int(rng.choice([0, 0, 0, 2, 1]))
    rate = int(rng.choice([44100, 48000, 32000, 22050, 24000, 16000] + ([] if layer else [11025, 12000, 8000])))
    lsf = rate < 32000
    cfg = dict(nframes=int(rng.integers(1, 60)), seed=int(rng.integers(1, 10 ** 6)), sample_rate=rate,
               mode=int(rng.choice([0, 1, 3])), level_lo_db=int(rng.choice([8, 30, 60])), level_hi_db=int(rng.choice([61, 70, 90])))
    if layer:
        cfg.update(layer=layer, bitrate_kbps=(64 if lsf else 192) if layer == 2 else (128 if lsf else 256))
        if layer == 2 and cfg["mode"] == 3:
            cfg["bitrate_kbps"] = 64 if lsf else 96
    else:
        cfg.update(blocks=int(rng.integers(0, 2)), bitrate_kbps=64 if lsf else 128, tag=int(rng.choice([0, 1])),
                   tag_lame=1, enc_delay=int(rng.integers(0, 1200)), enc_padding=int(rng.integers(0, 1500)))
    try:
        streams.append(synth.make_stream(**cfg))
    except ValueError:
        pass
bad = 0
for fmt in (mp3_b200.PCM_S16, mp3_b200.PCM_F32):
    for gapless in (False, True):
        with mp3_b200.Decoder(device=0, pcm_format=fmt, gapless=gapless) as dec:
            dec.decode_batch(streams)
            arena = dec.fetch_pcm()
            pl = dec.planar()
            params = dict(threshold=int(rng.integers(20, 3000)), min_silence_ms=int(rng.choice([10, 50, 300])),
                          min_sentence_ms=int(rng.choice([0, 40, 200])))
            segs = dec.segments(**params)
            for i in range(len(streams)):
                inf = dec.stream_info(i)
                if not inf.frames or inf.samples <= 0:
                    continue
                pcm = dec.stream_pcm(i, arena)
                got = pl[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].reshape(inf.channels, inf.samples).T
                ok = np.array_equal(got, pcm)
                E, W = seg_oracle.window_energy(pcm, inf.sample_rate)
                gE, gW = dec.window_energy(i)
                ok = ok and gW == W and np.array_equal(gE, E)
                ok = ok and np.array_equal(segs[i], seg_oracle.segments(pcm, inf.sample_rate, **params))
                if not ok:
                    bad += 1
                    print("MISMATCH stream", i, "fmt", fmt, "gapless", gapless, params)
print("%d mismatches over %d streams x 4 configurations" % (bad, len(streams)))
