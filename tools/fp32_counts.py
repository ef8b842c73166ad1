#!/usr/bin/env python3
"""Executed FP32 flops of the kernels in an .ncu-rep, from the per-instruction thread counts of its SASS view
(FFMA = 2 flop, FFMA2 = 4, FADD / FMUL = 1, FADD2 / FMUL2 = 2 per predicated-on thread), per decoded unit.

usage: fp32_counts.py report.ncu-rep <workload> <units_per_launch>  -> updates profiles/fp32_counts.json
bench.py divides `flop_per_unit` x units by the kernel's CUDA-event time for roofline.executed."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOP = {"FFMA": 2, "FFMA2": 4, "FADD": 1, "FMUL": 1, "FADD2": 2, "FMUL2": 2}


def main():
    rep, workload, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    out = {}
    kernel, h = None, None
    acc = None
    for row in rows:
        if row and row[0] in ("Function Name", "Kernel Name"):
            kernel = row[1]
            continue
        if row and row[0] == "Address":
            h = row
            acc = out.setdefault(kernel or "k_backend (only kernel of the report)", collections.Counter())
            continue
        if h is None or len(row) < len(h):
            continue
        s = row[h.index("Source")].split()
        if not s:
            continue
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        try:
            n = int(row[h.index("Predicated-On Thread Instructions Executed")])
            w = int(row[h.index("Instructions Executed")])
        except ValueError:
            continue
        acc["warp_inst"] += w
        if op in FLOP:
            acc["flop"] += FLOP[op] * n
            acc[op] += n
    path = os.path.join(ROOT, "profiles", "fp32_counts.json")
    try:
        db = json.load(open(path))
    except Exception:  # noqa: BLE001
        db = {}
    for k, a in out.items():
        name = "fused" if "k_backend" in k else ("huffman" if "k_huffman" in k else k[:40])
        db.setdefault(workload, {})[name] = {
            "flop_per_unit": a["flop"] / units, "warp_instructions_per_unit": a["warp_inst"] / units,
            "thread_ops_per_unit": {o: a[o] / units for o in FLOP if a[o]},
            "source": os.path.basename(rep), "units_per_launch": units,
            "how": "ncu --set full --import-source on: predicated-on thread counts per SASS opcode; FFMA 2, FFMA2 4, "
                   "FADD / FMUL 1, FADD2 / FMUL2 2 flop"}
        print(name, json.dumps(db[workload][name]))
    json.dump(db, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
