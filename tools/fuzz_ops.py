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

# ---- time stretch at random ratios: chosen offsets bit-exact against oracle/wsola.py, PCM within rounding
from oracle import wsola  # noqa: E402

bad_ts = n_ts = 0
ratios = [(1, 2), (3, 4), (5, 4), (7, 10), (2, 1), (1, 4), (9, 8), (13, 16), (3, 1)]
sub = [s for s in streams[:24]]
for fmt in (mp3_b200.PCM_S16, mp3_b200.PCM_F32):
    with mp3_b200.Decoder(device=0, pcm_format=fmt, gapless=True) as dec:
        dec.decode_batch(sub)
        arena = dec.fetch_pcm().copy()
        for num, den in [ratios[int(k)] for k in rng.choice(len(ratios), 4, replace=False)]:
            dec.time_stretch(num, den)
            out, where = dec.fetch_stretched()
            for i in range(len(sub)):
                inf = dec.stream_info(i)
                if not inf.frames or inf.samples <= 0:
                    continue
                pcm = dec.stream_pcm(i, arena)
                if fmt == mp3_b200.PCM_S16:
                    s16, x = pcm.astype(np.int64), pcm.astype(np.float64) / 32768.0
                else:
                    x, s16 = pcm.astype(np.float64), wsola.to_s16(pcm)
                ref, offs = wsola.wsola(x, s16, inf.sample_rate, num, den)
                got_offs, hop = dec.stretch_offsets(i)
                off, cnt = where[i]
                got = out[off: off + cnt * inf.channels].reshape(cnt, inf.channels).astype(np.float64)
                n_ts += 1
                ok = np.array_equal(got_offs, offs) and cnt == ref.shape[0]
                if ok and fmt == mp3_b200.PCM_F32:
                    ok = np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
                elif ok:
                    ok = np.abs(got - np.clip(np.rint(ref * 32768.0), -32768, 32767)).max() <= 1
                if not ok:
                    bad_ts += 1
                    print("STRETCH MISMATCH stream", i, "fmt", fmt, "speed", num, den)
print("%d stretch mismatches over %d (stream, ratio, format) cases" % (bad_ts, n_ts))

# ---- sample-rate conversion: tensor-core path forced on against the FP32 kernels (1 LSB)
bad_rs = n_rs = 0
for rate in (48000, 96000, 44100):
    res = {}
    for tc in ("1", "0"):
        os.environ["MP3B_RS_TC"] = tc
        with mp3_b200.Decoder(device=0, pcm_format=mp3_b200.PCM_S16) as dec:
            dec.decode_batch(streams)
            dec.resample(rate)
            res[tc] = dec.fetch_resampled()
    a, b = res["1"][0].astype(np.int32), res["0"][0].astype(np.int32)
    n_rs += 1
    if res["1"][1] != res["0"][1] or np.abs(a - b).max() > 1:
        bad_rs += 1
        print("RESAMPLE MISMATCH rate", rate, int(np.abs(a - b).max()))
os.environ.pop("MP3B_RS_TC", None)
print("%d resample mismatches over %d rates (tensor-core path against the FP32 kernels, %d streams)" % (bad_rs, n_rs, len(streams)))
