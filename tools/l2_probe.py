#!/usr/bin/env python3
"""Exploratory: one Layer II batch (256 x 10 s) decoded twice (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

base = [synth.make_stream(layer=2, bitrate_kbps=192, nframes=383, seed=1000 + i) for i in range(32)]
streams = [base[i % 32] for i in range(256)]
with mp3_b200.Decoder(device=0) as dec:
    for _ in range(2):
        dec.decode_batch(streams)
    dec.sync()
