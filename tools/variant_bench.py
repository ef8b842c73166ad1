#!/usr/bin/env python3
"""Kernel-tuning aid: run the parity tests and the stage timings for several builds of the library in one GPU call.

  python tools/variant_bench.py name[:ENV=val,...] ...     (name = mp3_b200/variants/libmp3b_<name>.so, or "main")

For each variant: the parity / golden / poison GPU tests (results must not change), then
bench.py --no-e2e --no-sweep --no-cpu for cfg2 / cfg3 / cfg4 and the stage times.  One JSON line per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        lib = os.path.join(ROOT, "mp3_b200", "libmp3b.so" if name == "main" else "variants/libmp3b_%s.so" % name)
        env["MP3B_LIB"] = lib
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        res = {"variant": spec}
        t = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "tests/test_gpu_parity.py",
                            "tests/test_gpu_golden.py", "tests/test_gpu_poison.py"], cwd=ROOT, env=env,
                           capture_output=True, text=True)
        res["tests"] = t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-300:]
        if t.returncode != 0:
            res["tests_tail"] = t.stdout[-1500:]
        for wl in ("cfg2", "cfg3", "cfg4"):
            b = subprocess.run([sys.executable, "bench.py", "--no-e2e", "--no-sweep", "--no-cpu", "--steps", "20",
                                "--workload", wl], cwd=ROOT, env=env, capture_output=True, text=True)
            try:
                d = json.loads(b.stdout.strip().splitlines()[-1])
                res[wl] = {"ms": round(d["ms_per_step"], 4), "huffman": round(d["stage_ms"]["huffman"], 4),
                           "fused": round(d["stage_ms"]["fused"], 4), "index": round(d["stage_ms"]["index"], 4),
                           "parity": d["parity_checked"]}
            except Exception:  # noqa: BLE001
                res[wl] = {"error": (b.stderr or b.stdout)[-400:]}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
