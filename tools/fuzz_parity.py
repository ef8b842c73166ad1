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
rng = np.random.default_rng(int(os.environ.get('FUZZ_SEED', '1234')))
base = synth.make_workload("cfg3", 6, 24) + synth.make_workload("cfg4", 9, 24) + [
    synth.make_stream(layer=2, bitrate_kbps=192, nframes=10, seed=5),
    synth.make_stream(layer=2, bitrate_kbps=64, sample_rate=24000, mode=3, nframes=10, seed=6, crc=1),
    synth.make_stream(layer=1, bitrate_kbps=256, nframes=20, seed=7),
    synth.make_stream(layer=1, bitrate_kbps=128, sample_rate=32000, mode=1, nframes=20, seed=8, crc=1),
    synth.make_stream(sample_rate=8000, bitrate_kbps=16, blocks=1, nframes=16, seed=9),
    synth.make_stream(sample_rate=11025, bitrate_kbps=32, mode=1, blocks=1, mixed_pct=30, nframes=16, seed=10, crc=1),
    synth.make_stream(nframes=16, seed=11, crc=1, mode=1, blocks=1),
    synth.make_stream(nframes=16, seed=12, tag=1, tag_lame=1, enc_delay=576, enc_padding=1000)]
# real-signal streams from the in-tree encoder (long count1 regions, bit reservoir up to 511 bytes, window switching)
from mp3_b200 import signals  # noqa: E402
base += [synth.encode_pcm(signals.stereo(44100, 0.5), 44100, 128), synth.encode_pcm(signals.to_s16(signals.speech(48000, 0.5) * 0.8), 48000, 96),
         synth.encode_pcm(signals.to_s16(signals.castanets(44100, 0.6)), 44100, 160, short_blocks=True)]
crc = int(os.environ.get("FUZZ_CRC", "0"))

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
with mp3_b200.Decoder(device=0, pcm_format=mp3_b200.PCM_F32, verify_crc=bool(crc)) as dec:
    dec.decode_batch(bad)
    arena = dec.fetch_pcm()
    nbad = 0
    failed = []
    for i, s in enumerate(bad):
        r = oracle.decode(s, verify_crc=bool(crc))
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
            failed.append(i)
    print("%d / %d damaged streams differ" % (nbad, len(bad)))

# where in the pipeline do the failing streams start to differ (staged pipeline, stage dumps)
for i in failed[:8]:
    r = oracle.decode(bad[i], dumps=True, verify_crc=bool(crc))
    if r.rc != 0:
        continue
    with mp3_b200.Decoder(device=0, pcm_format=mp3_b200.PCM_F32, verify_crc=bool(crc), pipeline=mp3_b200.PIPE_STAGED,
                          keep_stages=True) as dec:
        dec.decode_batch([bad[i]])
        pcm = dec.stream_pcm(0, dec.fetch_pcm()).astype(np.float64)
        line = "  stream %d staged: pcm max err %.3g;" % (i, np.abs(pcm - r.pcm.T).max() if pcm.shape == r.pcm.T.shape else -1)
        for name, st, ref in (("sf", mp3_b200.STAGE_SF, r.sf), ("is", mp3_b200.STAGE_IS, r.is_), ("xr", mp3_b200.STAGE_XR, r.xr),
                              ("sb", mp3_b200.STAGE_SB, r.sb)):
            got = dec.stage(st)[: r.units].reshape(r.units, -1).astype(np.float64)
            ref = np.asarray(ref).reshape(r.units, -1).astype(np.float64)
            e = np.abs(got - ref).max(axis=1)
            tol = 0 if name in ("sf", "is") else 2e-5 * max(1.0, np.abs(ref).max())
            w = np.nonzero(e > tol)[0]
            line += " %s: %s" % (name, "ok" if w.size == 0 else "unit %d of %d (err %.3g, |ref| %.3g)" % (w[0], r.units, e[w[0]], np.abs(ref[w[0]]).max()))
        print(line)
