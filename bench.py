#!/usr/bin/env python3
"""bench.py -- decoded audio-seconds per second for the batched MP3 decode hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one pass of the hot path over one batch of synthetic streams (BASELINE.json configs[1]:
1,024 concurrent 44.1 kHz stereo 128 kbps streams, 383 frames = 10 s each).  One process per GPU
(torchrun for N > 1); streams shard across ranks with no collective (weak scaling: every rank
decodes its own 1,024 streams); the only torch.distributed traffic is the barrier and the
max-over-ranks of the timing.

  value    = audio-seconds decoded per second, inputs (raw MP3 bytes) resident in HBM, PCM left in HBM
  e2e      = the same metric through the C-ABI with pinned HOST buffers: H2D of the MP3 bytes and D2H
             of the PCM inside the timed region
  roofline = the dominant kernel's algorithmic bytes / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline = the scalar from-spec oracle (oracle/l3_oracle.c) on the host cores, bounded sample

`--impl reference` times the reference's CPU implementation of the path.  The reference repository
ships no code (/root/reference/README.md:1-84), so that arm is the oracle port on all host threads.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decoded_audio_seconds_per_second"
UNIT = "audio-s/s"


# ------------------------------------------------------------------------------------ helpers
def workload(name, nstreams, nframes, seed, distinct):
    from mp3_b200 import synth
    t = time.time()
    streams = synth.make_workload(name, nstreams, nframes, seed=seed, distinct=distinct)
    return streams, time.time() - t


WORKLOADS = {
    "cfg1": "44.1 kHz stereo 128 kbps CBR, long blocks, bit reservoir in use (BASELINE.json configs[0])",
    "cfg2": "44.1 kHz stereo 128 kbps CBR, long blocks, bit reservoir in use (BASELINE.json configs[1])",
    "cfg3": "44.1 kHz 320 kbps joint stereo (MS+IS), long/short/mixed/switching windows, reservoir swept "
            "(BASELINE.json configs[2])",
    "cfg4": "MPEG-2 LSF 22.05/24 kHz + MPEG-1 44.1 kHz, VBR 8-320 kbps, mono and stereo mixed (BASELINE.json configs[3])",
    "cfg5": "100k-stream sweep of the cfg2 type, 128 frames each, streams sharded over the GPUs "
            "(BASELINE.json configs[4])",
}
DEFAULT_FRAMES = {"cfg5": 128}
DEFAULT_STREAMS = {"cfg1": 1, "cfg5": 100000}


def cpu_sample_size(nstreams, cores):
    """Streams for the bounded CPU sample: about 10-30 s of CPU work (a 10-s stream costs the oracle
    ~40 ms on one core), never more than the workload has."""
    return max(1, min(nstreams, cores * 32))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm, self.ok = index, False, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons)}


def cpu_oracle_throughput(streams, threads, min_seconds=0.0):
    """Decode `streams` with the oracle on `threads` host threads; returns audio-s/s."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    oracle.lib()
    info = oracle.decode(streams[0], want_pcm=False)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:  # l3o_decode releases the GIL (ctypes)
        res = list(ex.map(lambda s: oracle.decode(s, want_pcm=True).samples, streams))
    dt = time.perf_counter() - t0
    return sum(res) / float(info.sample_rate) / dt, dt


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsample = cpu_sample_size(1 << 30, cores)
    streams, _ = workload(args.workload, nsample, args.frames, args.seed, None)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_oracle_throughput(streams, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    v = float(np.mean([a for a, _ in vals]))
    ms = float(np.mean([b for _, b in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d-stream sample x %d frames, %s" % (
            args.workload, nsample, args.frames or DEFAULT_FRAMES.get(args.workload, 383), WORKLOADS[args.workload])},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d streams per step, all host threads; the reference repository has no code, "
                                   "so this is the from-spec oracle port" % nsample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world):
    import torch
    import mp3_b200
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    numa = None
    if world > 1 and not args.no_numa:  # before any pinned allocation: first touch decides the node
        from mp3_b200 import multi
        numa = multi.bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # cfg2 (the headline config): weak scaling, every rank decodes its own 1,024 streams.
    # cfg5 (the 100k-stream sweep): strong scaling, the 100k streams are sharded over the ranks.
    strong = args.workload == "cfg5"
    nstreams = args.streams
    if strong:
        total = args.streams or DEFAULT_STREAMS["cfg5"]
        nstreams = total // world + (1 if rank < total % world else 0)
        if args.distinct is None:
            args.distinct = 1024  # bounds host-side generation; the decoder still decodes every stream
    streams, gen_s = workload(args.workload, nstreams, args.frames, args.seed + 100000 * rank, args.distinct)
    packed, offs = mp3_b200.pack_streams(streams)
    nbytes_in = int(packed.size)

    dec = mp3_b200.Decoder(device=local, pcm_format=mp3_b200.PCM_S16, pipeline=args.pipeline_id)
    # a dedicated (non-default) stream: the library launches on it and the timing events are
    # recorded on it.  (Handle 0 would mean "the context's own stream" to mp3b_ctx_set_stream.)
    tstream = torch.cuda.Stream()
    assert tstream.cuda_stream != 0
    dec.set_stream(tstream.cuda_stream)

    # inputs resident in HBM
    d_raw = torch.empty(nbytes_in + 64, dtype=torch.uint8, device="cuda")
    d_raw[:nbytes_in].copy_(torch.from_numpy(packed))
    torch.cuda.synchronize()

    def step_device():
        dec.decode_packed(d_raw.data_ptr(), offs, where=mp3_b200.DEVICE, sync=False)

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device()
    dec.sync()
    st0 = dec.stats()
    infos = [dec.stream_info(i) for i in range(len(streams))]
    audio_s = sum(i.samples / float(i.sample_rate) for i in infos if i.frames)
    units = st0.units
    pcm_bytes = st0.pcm_bytes
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {k: 0.0 for k in ("index", "huffman", "requant", "imdct", "overlap", "synth", "fused")}
    e0.record(tstream)
    for _ in range(args.steps):
        step_device()
        if args.stage_times:
            dec.sync()
            st = dec.stats()
            for k in stage_ms:
                stage_ms[k] += getattr(st, "ms_" + k)
    e1.record(tstream)
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    launches = dec.stats().kernel_launches

    # per-stage CUDA-event times (library events on the launching stream), separate untimed passes
    if not args.stage_times:
        for _ in range(max(1, min(args.steps, 3))):
            step_device()
            dec.sync()
            st = dec.stats()
            for k in stage_ms:
                stage_ms[k] += getattr(st, "ms_" + k)
        nst = max(1, min(args.steps, 3))
    else:
        nst = args.steps
    stage_ms = {k: v / nst for k, v in stage_ms.items()}

    # ---- end to end: pinned host in, pinned host out
    do_e2e = not args.no_e2e and pcm_bytes < (8 << 30)  # cfg5 would need 59 GB of pinned host memory
    ms_e2e, checksum = float("nan"), None
    if do_e2e:
        ms_e2e, checksum = run_e2e(args, mp3_b200, dec, packed, offs, nbytes_in, pcm_bytes, tstream, barrier)
    sampler.stop_flag = True
    finish(args, rank, world, dist, dec, streams, units, nbytes_in, pcm_bytes, audio_s, ms_dev, ms_e2e, stage_ms,
           launches, sampler, gen_s, checksum, strong, numa)


def run_e2e(args, mp3_b200, dec, packed, offs, nbytes_in, pcm_bytes, tstream, barrier):
    import torch
    h_in = mp3_b200.PinnedBuffer(nbytes_in + 64)
    h_in.view(np.uint8)[:nbytes_in] = packed
    h_out = mp3_b200.PinnedBuffer(pcm_bytes + 64)
    pcm_elems = pcm_bytes // 2

    dec.set_pcm_sink(h_out.ptr, pcm_elems)  # D2H of each wave overlaps the next wave's kernels

    def step_e2e():
        dec.decode_packed(h_in.ptr, offs, where=mp3_b200.HOST, sync=False)

    for _ in range(min(args.warmup, 3)):
        step_e2e()
    dec.sync()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(tstream)
    for _ in range(args.steps):
        step_e2e()
    dec.flush()  # the stream (and so the end event) waits for every D2H copy of every step
    e3.record(tstream)
    barrier()
    dec.set_pcm_sink(0, 0)
    ms_e2e = e2.elapsed_time(e3) / args.steps
    checksum = int(h_out.view(np.int16, pcm_elems)[:: max(1, pcm_elems // 4096)].astype(np.int64).sum())
    return ms_e2e, checksum


def finish(args, rank, world, dist, dec, streams, units, nbytes_in, pcm_bytes, audio_s, ms_dev, ms_e2e, stage_ms,
           launches, sampler, gen_s, checksum, strong, numa=None):
    import torch
    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        tot = torch.tensor([audio_s, float(nbytes_in), float(pcm_bytes), float(units)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        audio_all, in_all, pcm_all, units_all = [float(x) for x in tot]
    else:
        audio_all, in_all, pcm_all, units_all = audio_s, float(nbytes_in), float(pcm_bytes), float(units)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # dominant kernel and its algorithmic bytes per unit:
        # what the stage must read and write by definition, s16 PCM out
        in_per_unit = nbytes_in / max(units, 1)
        alg = {"huffman": in_per_unit + 1152.0 + 40.0,           # compressed bits in; int16 spectrum + scalefactors out
               "requant": 1152.0 + 2304.0, "imdct": 2304.0 + 4608.0, "overlap": 4608.0 + 2304.0,
               "synth": 2304.0 + 1152.0,
               "fused": 1152.0 + 40.0 + 1152.0}                  # int16 spectrum + scalefactors in; s16 PCM out
        kern = {k: v for k, v in stage_ms.items() if k != "index" and v > 0}
        if kern:
            dom = max(kern, key=kern.get)
            achieved = alg[dom] * units / (kern[dom] * 1e-3) / 1e9
        else:  # a decode cut into several waves has no per-stage events: whole pipeline only
            dom = "pipeline"
            alg[dom] = in_per_unit + 1152.0
            achieved = alg[dom] * units / (ms_dev * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = prof.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        cores = os.cpu_count() or 1
        nsample = cpu_sample_size(len(streams), cores)
        cpu_v, cpu_dt = cpu_oracle_throughput(streams[:nsample], cores)
        line = {
            "metric": METRIC, "value": audio_all / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "%s: %d streams/GPU x %d frames, %s" % (
                    args.workload, len(streams), args.frames or DEFAULT_FRAMES.get(args.workload, 383),
                    WORKLOADS[args.workload]),
                "streams_per_gpu": len(streams), "distinct_streams": args.distinct or len(streams),
                "pcm": "s16 interleaved", "pipeline": args.pipeline, "indexer": "device",
                "l2_policy": "inputs_larger_than_l2 (%.0f MB in, %.0f MB PCM out per step)" % (nbytes_in / 1e6, pcm_bytes / 1e6),
                "timing": "torch.cuda.Event on the stream the library launches on",
                "host_binding": numa or "none",
            },
            "pcm_gbs": pcm_all / (ms_dev * 1e-3) / 1e9,
            "x_realtime_per_gpu": audio_all / world / (ms_dev * 1e-3),
            "stage_ms": stage_ms,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_unit": alg[dom], "units_per_launch": units,
                         "kernel_ms": kern.get(dom),
                         "whole_pipeline": {"algorithmic_bytes_per_unit": in_per_unit + 1152.0,
                                            "achieved": (in_per_unit + 1152.0) * units / (ms_dev * 1e-3) / 1e9,
                                            "frac": (in_per_unit + 1152.0) * units / (ms_dev * 1e-3) / 1e9 / hbm_peak},
                         "note": "the back end is FP32 / shared-memory issue bound, not HBM bound (ncu: issue 74 %, LSU "
                                 "67 %, FMA 37 %, DRAM 10 %); its DRAM traffic is at the algorithmic bytes (profiles/)"},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "first %d streams of the workload (%.1f s of wall time on %d threads = %.0f s of "
                                       "CPU work), oracle/l3_oracle.c, one stream per host thread" % (
                                           nsample, cpu_dt, cores, cpu_dt * min(cores, nsample))},
            "e2e": ({"value": audio_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(in_all),
                     "d2h_bytes_per_step": int(pcm_all), "ms_per_step": ms_e2e,
                     "pcm_gbs": pcm_all / (ms_e2e * 1e-3) / 1e9} if ms_e2e == ms_e2e else None),
            "gpu_launches": int(launches * args.steps),
            "clocks": sampler.result(),
            "gen_seconds": gen_s, "pcm_checksum": checksum,
        }
        print(json.dumps(line))
    dec.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--streams", type=int, default=None, help="streams per GPU (default: the workload's)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--distinct", type=int, default=None, help="generate only this many distinct streams and repeat")
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--pipeline", default="default", choices=["default", "fused", "staged"])
    ap.add_argument("--stage-times", action="store_true", help="sync after every step to read per-stage events")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to their GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.pipeline_id = {"default": None, "fused": 0, "staged": 1}[args.pipeline]
    if args.frames is None:
        args.frames = DEFAULT_FRAMES.get(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
