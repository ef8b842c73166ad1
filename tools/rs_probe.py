#!/usr/bin/env python3
"""Exploratory: one cfg2 batch, then mp3b_batch_resample(48000) twice (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

streams = synth.make_workload("cfg2", 256, 383)
with mp3_b200.Decoder(device=0) as dec:
    dec.decode_batch(streams)
    dec.sync()
    for _ in range(2):
        dec.resample(int(sys.argv[1]) if len(sys.argv) > 1 else 48000)
    dec.sync()
