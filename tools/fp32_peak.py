#!/usr/bin/env python3
"""Run tools/fp32_peak (the FFMA microbenchmark) while sampling the SM clock through NVML, and write
profiles/fp32_peak.json -- the FP32 denominator bench.py uses for the back-end kernel's roofline."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    exe = os.path.join(ROOT, "tools", "fp32_peak")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(exe + ".cu"):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", exe + ".cu", "-o", exe])
    clocks, reasons, stop = [], set(), []

    def sample():
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(0)
            while not stop:
                clocks.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                                (0x4, "sw_power_cap")):
                    if r & bit:
                        reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            reasons.add("nvml_unavailable: %s" % e)

    t = threading.Thread(target=sample, daemon=True)
    t.start()
    out = subprocess.check_output([exe] + sys.argv[1:2], text=True)
    stop.append(1)
    t.join(timeout=2)
    res = json.loads(out)
    if clocks:
        s = sorted(clocks)
        res["sm_mhz_median_under_load"] = s[len(s) // 2]
        res["sm_mhz_max_seen"] = s[-1]
    res["throttle_reasons"] = sorted(reasons)
    res["how"] = ("tools/fp32_peak.cu: 8 independent FFMA chains per thread, grid = SMs x resident CTAs, 20000 x 64 "
                  "FFMAs per thread per launch; burst = best of 10 launches, sustained = back to back for the stated time")
    path = os.path.join(ROOT, "profiles", "fp32_peak.json")
    json.dump(res, open(path, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
