#!/usr/bin/env python3
"""Per source line of one kernel in an .ncu-rep: share of warp instructions, stall samples and the
average number of active threads per instruction (divergence).

usage: ncu_lines2.py report.ncu-rep kernel-substring [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, filt = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur = fn = h = None
    data = {}
    for row in csv.reader(io.StringIO(src)):
        if not row:
            continue
        if row[0] == "File Path":
            cur, h = row[1], None
            continue
        if row[0] == "Function Name":
            fn = row[1]
            continue
        if row[0] == "Line No":
            h = row
            iI, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
            continue
        if h is None or cur is None or filt not in (fn or ""):
            continue
        try:
            ln, n, t, s = int(row[0]), int(row[iI]), int(row[iT]), int(row[iS])
        except (ValueError, IndexError):
            continue
        d = data.setdefault((cur.split("/")[-1], ln), [0, 0, 0, row[1].strip()[:88]])
        d[0] += n
        d[1] += t
        d[2] += s
    tot = sum(v[0] for v in data.values())
    tots = sum(v[2] for v in data.values())
    tthr = sum(v[1] for v in data.values())
    print("kernel filter %r: warp inst %d, thread inst %d (avg %.1f / inst), samples %d" % (
        filt, tot, tthr, tthr / max(tot, 1), tots))
    for k, v in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.2f%% inst %5.2f%% smp thr %4.1f  %s:%d  %s" % (
            100.0 * v[0] / max(tot, 1), 100.0 * v[2] / max(tots, 1), v[1] / max(v[0], 1), k[0], k[1], v[3]))


if __name__ == "__main__":
    main()
