#!/usr/bin/env python3
"""GPU timeline of a few pipelined decode calls (cfg2, device-resident input) from CUPTI through torch.profiler: every
kernel / copy / memset of one step with its start offset, duration and the idle gap in front of it on the device.
usage: timeline.py [workload] [out.json]"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    import mp3_b200 as m
    from mp3_b200 import synth
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    streams = synth.make_workload(wl, 1024)
    packed, offs = m.pack_streams(streams)
    d_raw = torch.empty(packed.size + 64, dtype=torch.uint8, device="cuda")
    d_raw[: packed.size].copy_(torch.from_numpy(packed))
    ts = torch.cuda.Stream()
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.set_stream(ts.cuda_stream)
        for _ in range(4):
            dec.decode_packed(d_raw.data_ptr(), offs, where=m.DEVICE, sync=False)
        dec.sync()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
            for _ in range(6):
                dec.decode_packed(d_raw.data_ptr(), offs, where=m.DEVICE, sync=False)
            dec.sync()
        path = os.path.join(tempfile.mkdtemp(), "trace.json")
        prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    # one step = from one k_side_parse to the next
    starts = [i for i, e in enumerate(ev) if "k_side_parse" in e["name"]]
    out = []
    if len(starts) >= 4:
        a, b = starts[2], starts[3]
        t0 = ev[a]["ts"]
        busy_end = max(e["ts"] + e["dur"] for e in ev[:a]) if a else t0
        for e in ev[a:b]:
            gap = e["ts"] - busy_end
            nm = e["name"]
            k = nm.find("k_")
            nm = nm[k:].split("(")[0].split("<")[0] if k >= 0 else nm.split("(")[0]
            out.append({"name": nm[:40], "stream": e.get("args", {}).get("stream"), "start_us": round(e["ts"] - t0, 1),
                        "dur_us": round(e["dur"], 1), "idle_before_us": round(max(gap, 0.0), 1)})
            busy_end = max(busy_end, e["ts"] + e["dur"])
        period = ev[b]["ts"] - t0
        res = {"workload": wl, "step_us": round(period, 1), "idle_us": round(sum(o["idle_before_us"] for o in out), 1), "ops": out}
    else:
        res = {"error": "no steps found", "n_events": len(ev)}
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
