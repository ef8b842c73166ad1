#!/usr/bin/env python3
"""The sample-rate converter on the cfg2 batch (1,024 x 10 s stereo s16), tensor-core path against the FP32 kernels
(MP3B_RS_TC = 1 / 0), for several output rates; CUDA events on the context's stream; also checks that the two paths
agree within 1 LSB.  One JSON line per rate."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mp3_b200  # noqa: E402
from mp3_b200 import synth  # noqa: E402

streams = synth.make_workload("cfg2", 1024, 383)
audio = 1024 * 383 * 1152 / 44100.0
res = {}
outs = {}
for tc in (("1",) if os.environ.get("RS_TC_ONLY") else ("1", "0")):
    os.environ["MP3B_RS_TC"] = tc
    dec = mp3_b200.Decoder(device=0)
    st = torch.cuda.Stream()
    dec.set_stream(st.cuda_stream)
    dec.decode_batch(streams)
    for rate in ([int(a) for a in sys.argv[1:]] or [48000, 96000, 88200]):
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
        out, where = dec.fetch_resampled()
        nbytes = out.nbytes + 1024 * 383 * 1152 * 2 * 2  # algorithmic: one read of the arena, one write of the result
        res.setdefault(rate, {})["tc" if tc == "1" else "fp32"] = {"ms": round(ms, 3), "x_realtime": round(audio / (ms * 1e-3)),
                                                                  "effective_gbs": round(nbytes / (ms * 1e-3) / 1e9)}
        if rate == 48000:
            outs[tc] = out[: 1 << 26].astype(np.int32)
    dec.close()
if "0" in outs and "1" in outs:
    d = np.abs(outs["1"] - outs["0"])
    res["agreement_48000"] = {"max_abs_diff_lsb": int(d.max()), "fraction_differing": float(np.mean(d != 0))}
print(json.dumps(res))
