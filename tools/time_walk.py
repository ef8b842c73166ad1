#!/usr/bin/env python3
"""Index-stage time (frame walk + side info + main-data compaction, CUDA events) of the serial and the time-parallel
frame walk: one long stream (an hour at 128 kbit/s: 138 k frames), and the cfg2 / cfg3 batches.  Also the cfg1
latency: one 10-s stream, host bytes in -> host PCM out."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def index_ms(m, streams, mode, seg=None, reps=4):
    os.environ["MP3B_WALK"] = mode
    if seg:
        os.environ["MP3B_WALK_SEG"] = str(seg)
    else:
        os.environ.pop("MP3B_WALK_SEG", None)
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.set_stage_timing(True)
        v, tot = [], []
        for _ in range(reps):
            dec.decode_batch(streams)
            st = dec.stats()
            v.append(st.ms_index)
            tot.append(st.ms_total)
        fr = dec.stats().frames
    return {"mode": mode, "seg": seg, "frames": int(fr), "ms_index": round(min(v[1:]), 3), "ms_total": round(min(tot[1:]), 3)}


def main():
    import mp3_b200 as m
    from mp3_b200 import synth
    out = {}
    hour = synth.make_stream(nframes=138000, seed=5, mode=1, bitrate_kbps=128, blocks=1)
    out["one_hour_stream"] = [index_ms(m, [hour], "serial"), index_ms(m, [hour], "par"), index_ms(m, [hour], "par", 16368),
                              index_ms(m, [hour], "par", 1008)]
    for wl in ("cfg2", "cfg3"):
        s = synth.make_workload(wl, 1024)
        out[wl] = [index_ms(m, s, "serial"), index_ms(m, s, "par"), index_ms(m, s, "par", 16368)]
    # cfg1 latency: one 10-s stream, pageable host bytes in, host PCM out, wall clock around the C-ABI calls
    os.environ["MP3B_WALK"] = "auto"
    os.environ.pop("MP3B_WALK_SEG", None)
    one = synth.make_workload("cfg1", 1)
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.set_stage_timing(True)
        lat = []
        for _ in range(12):
            t = time.perf_counter()
            dec.decode_batch(one)
            pcm = dec.fetch_pcm()
            lat.append((time.perf_counter() - t) * 1e3)
        st = dec.stats()
        out["cfg1_latency"] = {"host_to_host_ms_median": round(float(np.median(lat[2:])), 3), "min": round(min(lat[2:]), 3),
                               "gpu_ms_total": round(st.ms_total, 3), "ms_index": round(st.ms_index, 3),
                               "ms_huffman": round(st.ms_huffman, 3), "ms_fused": round(st.ms_fused, 3),
                               "audio_s": 10.005, "pcm_samples": int(pcm.size)}
    t = time.perf_counter()
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.set_stage_timing(True)
        for _ in range(3):
            t = time.perf_counter()
            dec.decode_batch([hour])
            dec.sync()
            dt = time.perf_counter() - t
        st = dec.stats()
        out["one_hour_decode"] = {"wall_ms": round(dt * 1e3, 2), "gpu_ms_total": round(st.ms_total, 3), "ms_index": round(st.ms_index, 3),
                                  "x_realtime": round(st.frames * 1152 / 44100.0 / dt, 0)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
