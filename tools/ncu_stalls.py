#!/usr/bin/env python3
"""Stall-reason totals and the most-stalled SASS instructions (with context) of a kernel in an .ncu-rep.

usage: ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h = r[0]
    for row in r[2:]:
        print(row[h.index("Kernel Name")][:60])
        for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                  "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                  "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                  "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
                  "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
                  "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "launch__registers_per_thread"):
            if k in h:
                print("   %-66s %s" % (k, row[h.index(k)]))
        st = []
        for k, v in zip(h, row):
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((float(v.replace(",", "")), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(a for a, _ in st) or 1
        print("   stalls: " + ", ".join("%s %.1f%%" % (k, 100 * a / tot) for a, k in sorted(st, reverse=True) if a / tot > 0.01))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, data = None, []
    for row in rows:
        if row and row[0] == "Address":
            h = row
            continue
        if h is None or len(row) < len(h):
            continue
        data.append(row)
    if not data:
        return
    iS, iSrc, iI = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "(Not" not in c]
    top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:top_n]
    for i in top:
        for row in data[max(0, i - 4): i + 2]:
            s = sorted(((h[k][6:], int(row[k])) for k in sc if row[k] not in ("", "0")), key=lambda x: -x[1])
            print(row[0][-5:], "%6s %9s" % (row[iS], row[iI]), row[iSrc][:64].ljust(64), s[:2])
        print()


if __name__ == "__main__":
    main()
