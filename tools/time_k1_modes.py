#!/usr/bin/env python3
"""The three Huffman kernels (MP3B_K1_MODE = chunk | sorted | warp) by batch size: stage time from the library's CUDA
events (stage timing on), for 1 / 4 / 16 / 64 / 1024 streams of the cfg2 shape (383 frames each).  One JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import mp3_b200 as m
    from mp3_b200 import synth
    allst = synth.make_workload("cfg2", 1024)
    out = {}
    for n in (1, 4, 16, 64, 1024):
        row = {}
        for mode in ("chunk", "sorted", "warp"):
            os.environ["MP3B_K1_MODE"] = mode
            with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
                dec.set_stage_timing(True)
                v = []
                for _ in range(6):
                    dec.decode_batch(allst[:n])
                    v.append(dec.stats().ms_huffman)
                row[mode] = round(min(v[1:]), 4)
                row["units"] = int(dec.stats().units)
        out["%d_streams" % n] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
