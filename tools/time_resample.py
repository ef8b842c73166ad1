#!/usr/bin/env python3
"""Time mp3b_batch_resample on the cfg2 batch (1,024 x 10 s, 44.1 -> 48 kHz): CUDA events on the
context's stream."""
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
