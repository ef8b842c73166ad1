#!/usr/bin/env python3
"""Time the output-side operators on the cfg2 batch (1,024 x 10 s stereo s16): mp3b_batch_resample
(44.1 -> 48 / 22.05 kHz) and mp3b_batch_time_stretch (half speed); CUDA events on the context's stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

streams = synth.make_workload("cfg2", 1024, 383)
dec = mp3_b200.Decoder(device=0)
st = torch.cuda.Stream()
dec.set_stream(st.cuda_stream)
dec.decode_batch(streams)
for rate in (48000, 22050):
    for _ in range(2):
        dec.resample(rate)
    dec.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        dec.resample(rate)
    e1.record(st)
    dec.sync()
    ms = e0.elapsed_time(e1) / 5
    audio = 1024 * 383 * 1152 / 44100.0
    print("44100 -> %d: %.3f ms per batch, %.2e x realtime" % (rate, ms, audio / (ms * 1e-3)))

for num, den in ((1, 2), (3, 4)):
    dec.time_stretch(num, den)
    dec.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        dec.time_stretch(num, den)
    e1.record(st)
    dec.sync()
    ms = e0.elapsed_time(e1) / 3
    audio = 1024 * 383 * 1152 / 44100.0
    print("stretch speed %d/%d: %.3f ms per batch, %.2e x realtime (of input audio)" % (num, den, ms, audio / (ms * 1e-3)))

import ctypes  # noqa: E402
L = dec.L
L.mp3b_batch_planar.argtypes = [ctypes.c_void_p]
for _ in range(2):
    L.mp3b_batch_planar(dec.ctx)
dec.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10):
    L.mp3b_batch_planar(dec.ctx)
e1.record(st)
dec.sync()
ms = e0.elapsed_time(e1) / 10
nbytes = 2 * 1024 * 383 * 1152 * 2 * 2  # read + write of the s16 stereo arena
print("planar: %.3f ms per batch, %.0f GB/s of HBM traffic (read + write)" % (ms, nbytes / (ms * 1e-3) / 1e9))

L.mp3b_batch_segments.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
for _ in range(2):
    L.mp3b_batch_segments(dec.ctx, 328, 300, 200)
dec.sync()
e0.record(st)
for _ in range(10):
    L.mp3b_batch_segments(dec.ctx, 328, 300, 200)
e1.record(st)
dec.sync()
ms = e0.elapsed_time(e1) / 10
nbytes = 1024 * 383 * 1152 * 2 * 2  # one read of the s16 stereo arena
print("segments: %.3f ms per batch, %.0f GB/s of PCM read" % (ms, nbytes / (ms * 1e-3) / 1e9))
