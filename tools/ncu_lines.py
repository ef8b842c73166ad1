#!/usr/bin/env python3
"""Summarise an .ncu-rep by CUDA source line: share of executed warp instructions and of stall samples.

usage: ncu_lines.py report.ncu-rep [top_n] [kernel-regex]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units = r[0], r[1]
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "launch__grid_size"]
    for row in r[2:]:
        for k in keys:
            if k in hdr:
                print("%-66s %s %s" % (k, row[hdr.index(k)][:90], units[hdr.index(k)]))
        print()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    lines, tot, ts = [], 0, 0
    for row in rows:
        if row and row[0] == "Line No":
            h = row
            iI, iS = h.index("Instructions Executed"), h.index("# Samples")
            continue
        if h is None or not row or row[0] == "":
            continue
        try:
            ln, n, s = int(row[0]), int(row[iI]), int(row[iS])
        except ValueError:
            continue
        lines.append((n, s, ln, row[1].strip()[:100]))
        tot += n
        ts += s
    print("total warp instructions", tot, "samples", ts)
    for n, s, ln, txt in sorted(lines, reverse=True)[:top]:
        print("%6.2f%% inst %6.2f%% smp  L%-4d %s" % (100.0 * n / max(tot, 1), 100.0 * s / max(ts, 1), ln, txt))
    print("--- by stall samples")
    for n, s, ln, txt in sorted(lines, key=lambda x: -x[1])[:top // 2]:
        print("%6.2f%% inst %6.2f%% smp  L%-4d %s" % (100.0 * n / max(tot, 1), 100.0 * s / max(ts, 1), ln, txt))


if __name__ == "__main__":
    main()
