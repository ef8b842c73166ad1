#!/usr/bin/env python3
"""Time a Layer II batch (1,024 x 10 s, 44.1 kHz stereo 192 kbit/s) through the C-ABI, device-resident
input, CUDA events on the context's stream; per-kernel times come from an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

base = [synth.make_stream(layer=2, bitrate_kbps=192, nframes=383, seed=1000 + i) for i in range(64)]
streams = [base[i % 64] for i in range(1024)]
packed, offs = mp3_b200.pack_streams(streams)
dec = mp3_b200.Decoder(device=0)
st = torch.cuda.Stream()
dec.set_stream(st.cuda_stream)
d_raw = torch.empty(packed.size + 64, dtype=torch.uint8, device="cuda")
d_raw[: packed.size].copy_(torch.from_numpy(packed))
torch.cuda.synchronize()
for _ in range(3):
    dec.decode_packed(d_raw.data_ptr(), offs, where=mp3_b200.DEVICE, sync=False)
dec.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(5):
    dec.decode_packed(d_raw.data_ptr(), offs, where=mp3_b200.DEVICE, sync=False)
e1.record(st)
dec.sync()
ms = e0.elapsed_time(e1) / 5
audio = 1024 * 383 * 1152 / 44100.0
print("layer II, 1024 x 10 s: %.3f ms per batch, %.2e x realtime" % (ms, audio / (ms * 1e-3)))
