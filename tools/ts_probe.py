#!/usr/bin/env python3
"""Exploratory: one cfg2-type batch (256 streams), then half-speed time stretch twice (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

streams = synth.make_workload("cfg2", 1024, 383)
with mp3_b200.Decoder(device=0) as dec:
    dec.decode_batch(streams)
    dec.sync()
    for _ in range(2):
        dec.time_stretch(1, 2)
    dec.sync()
