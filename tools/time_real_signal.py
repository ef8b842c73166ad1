#!/usr/bin/env python3
"""The cfg2 shape on REAL-SIGNAL streams: 1,024 x 10 s, 44.1 kHz stereo, 128 kbit/s CBR, encoded by the in-tree encoder
(gen/l3gen.c::l3enc_stream) from synthetic music + speech (32 distinct streams, each used 32 times), decoded with the
stage events on; next to the same batch of generator streams (random spectra, the workload bench.py times).
One JSON object."""
import concurrent.futures as cf
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def one(seed):
    from mp3_b200 import signals, synth
    return synth.encode_pcm(signals.stereo(44100, 10.0, 1 + 2 * seed), 44100, 128)


def run(m, streams):
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.set_stage_timing(True)
        best = None
        for _ in range(6):
            dec.decode_batch(streams)
            st = dec.stats()
            if best is None or st.ms_total < best["ms_total"]:
                best = {"ms_total": round(st.ms_total, 3), "ms_index": round(st.ms_index, 3), "ms_huffman": round(st.ms_huffman, 3),
                        "ms_fused": round(st.ms_fused, 3), "units": int(st.units)}
        return best


def main():
    import mp3_b200 as m
    from mp3_b200 import synth
    synth.build()
    with cf.ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        distinct = list(ex.map(one, range(32)))
    real = [distinct[i % 32] for i in range(1024)]
    gen = synth.make_workload("cfg2", 1024, 383)
    out = {"real_signal_128k": run(m, real), "generator_cfg2": run(m, gen),
           "note": "same shape (1,024 x 10 s, 44.1 kHz stereo, 128 kbit/s CBR, long blocks); real-signal streams carry fewer and "
                   "shorter code words per granule than the generator's random spectra"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
