#!/usr/bin/env python3
"""Turn gpurun_out/ ncu captures into the committed summaries under profiles/.

usage: summarise_profiles.py <tag> <launches.csv> <full.ncu-rep> <bench.json>
writes profiles/<tag>_launches.csv (copy), <tag>_launches_summary.txt, <tag>_ncu_full_summary.txt,
<tag>_bench.json and refreshes profiles/traffic.json (per-launch DRAM bytes of the timed kernels).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def short(name):
    return name.split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "nsecond": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3,
            "s": v * 1e6, "second": v * 1e6}[unit]


def main():
    tag, launches, rep, bench = sys.argv[1:5]
    os.makedirs(PROF, exist_ok=True)
    shutil.copy(launches, os.path.join(PROF, tag + "_launches.csv"))
    shutil.copy(bench, os.path.join(PROF, tag + "_bench.json"))
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(short(r[ik]), []).append(to_us(r[iv], r[iu]))
    tot = sum(sum(v) for v in agg.values())
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot_ours = sum(sum(v) for v in ours.values())
    with open(os.path.join(PROF, tag + "_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ; command: python bench.py --steps 5 --warmup 3\n")
        f.write("per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n\n")
        f.write("%-44s %8s %12s %10s %12s\n" % ("kernel", "launches", "mean us", "share", "share(ours)"))
        for k, v in agg.items():
            so = "%.1f%%" % (100 * sum(v) / tot_ours) if k in ours else "-"
            f.write("%-44s %8d %12.1f %9.1f%% %12s\n" % (k[:44], len(v), sum(v) / len(v), 100 * sum(v) / tot, so))
        b = json.load(open(bench))
        f.write("\nlive CUDA-event stage times of the same command (bench.py, ms): %s\n" % json.dumps(b.get("stage_ms")))
        sm = b.get("stage_ms", {})
        live = {k: v for k, v in sm.items() if v}
        lt = sum(live.values())
        if lt:
            f.write("live shares: " + ", ".join("%s %.1f%%" % (k, 100 * v / lt) for k, v in live.items()) + "\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    traffic = {}
    with open(os.path.join(PROF, tag + "_ncu_full_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on ; command: python bench.py --steps 5 --warmup 3\n\n")
        for row in r[2:]:
            for k in keys:
                if k in hdr:
                    f.write("%-64s %s %s\n" % (k, row[hdr.index(k)][:80], units[hdr.index(k)]))
            f.write("\n")
            name = short(row[hdr.index("Kernel Name")])
            def gb(key):
                v, u = float(row[hdr.index(key)].replace(",", "")), units[hdr.index(key)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            key = "huffman" if "huffman" in name else ("fused" if "backend" in name else name)
            traffic[key] = {"kernel": name, "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                            "dram_read": gb("dram__bytes_read.sum"), "dram_write": gb("dram__bytes_write.sum"),
                            "source": tag + "_ncu_full_summary.txt"}
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, tag + "_launches_summary.txt")).read())
    print(open(os.path.join(PROF, tag + "_ncu_full_summary.txt")).read())


if __name__ == "__main__":
    main()
