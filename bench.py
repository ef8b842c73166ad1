#!/usr/bin/env python3
"""bench.py -- decoded audio-seconds per second for the batched MP3 decode hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one pass of the hot path over one batch of synthetic streams (BASELINE.json configs[1]:
1,024 concurrent 44.1 kHz stereo 128 kbps streams, 383 frames = 10 s each).  One process per GPU
(torchrun for N > 1); streams shard across ranks with no collective (weak scaling: every rank
decodes its own 1,024 streams); the only torch.distributed traffic is the barrier and the
max-over-ranks of the timing.

  value        = audio-seconds decoded per second, inputs (raw MP3 bytes) resident in HBM, PCM left in HBM
  e2e          = the same metric through the C-ABI with pinned HOST buffers: H2D of the MP3 bytes and D2H
                 of the PCM inside the timed region; e2e.pcie_floor_ms = the same two copies alone
  splits       = SURVEY.md 8(d) (i)-(iv): decode kernels only / + H2D / + indexing / + D2H
  roofline     = the dominant kernel (the fused back end): executed and direct-form-effective FP32 flops
                 against the MEASURED FP32 peak (profiles/fp32_peak.json), and the HBM view of the whole
                 pipeline at SURVEY.md 8(d)'s 1256.5 algorithmic bytes per unit
  parity_checked = streams compared with the oracle inside this run, before anything is timed
  sweep_cfg5   = BASELINE.json configs[4]: 100,000 streams x 128 frames sharded over the N ranks (strong scaling)
  cpu_baseline = the scalar from-spec oracle (oracle/l3_oracle.c) on the host cores, bounded sample;
                 cpu_lines = oracle 1 thread / all threads and FFmpeg mp3float 1 thread / all threads

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
FUSED_IDEAL_PCM_BYTES = 1152.0          # 576 s16 samples out per unit
DIRECT_FORM_FLOP_PER_UNIT = 137088.0    # 67,968 MAC, direct (definition) form


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


def workload_string(args, streams_per_gpu):
    """The same string in both arms (the driver compares them)."""
    return "%s: %d streams/GPU x %d frames, %s" % (
        args.workload, streams_per_gpu, args.frames or DEFAULT_FRAMES.get(args.workload, 383), WORKLOADS[args.workload])


def default_streams(args, world=1, rank=0):
    if args.workload == "cfg5":
        total = args.streams or DEFAULT_STREAMS["cfg5"]
        return total // world + (1 if rank < total % world else 0)
    return args.streams or DEFAULT_STREAMS.get(args.workload, 1024)


def cpu_sample_size(nstreams, cores):
    """Streams for the bounded CPU sample: about 10-30 s of CPU work per pass (a 10-s stream costs the oracle
    ~40 ms on one core), never more than the workload has."""
    return max(1, min(nstreams, cores * 64))


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


# ------------------------------------------------------------------------------------ CPU lines
def cpu_oracle_throughput(streams, threads):
    """Decode `streams` with the oracle on `threads` host threads; returns (audio-s/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    oracle.lib()
    oracle.decode(streams[0], want_pcm=False)

    def one(s):
        d = oracle.decode(s, want_pcm=True)
        return d.samples / float(d.sample_rate or 1)

    t0 = time.perf_counter()
    if threads <= 1:
        res = [one(s) for s in streams]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:  # l3o_decode releases the GIL (ctypes)
            res = list(ex.map(one, streams))
    dt = time.perf_counter() - t0
    return sum(res) / dt, dt


def _ffmpeg_line_lib():
    """tools/ff_line.c: a C loop around libavcodec's mp3float (an independent production decoder; not the oracle)."""
    import ctypes
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ffmpeg_ref
    if not ffmpeg_ref.available():
        return None
    src, lib = os.path.join(ROOT, "tools", "ff_line.c"), os.path.join(ROOT, "tools", "libffline.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", src, "-o", lib, "-ldl"])
    L = ctypes.CDLL(lib)
    L.ffl_decode_stream.restype = ctypes.c_long
    L.ffl_decode_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p]
    return L if L.ffl_init() == 0 else None


def cpu_ffmpeg_throughput(streams, threads):
    """FFmpeg mp3float over `streams`; returns (audio-s/s, seconds) or None when libavcodec is not on the box."""
    from concurrent.futures import ThreadPoolExecutor
    try:
        L = _ffmpeg_line_lib()
    except Exception:  # noqa: BLE001
        L = None
    if L is None:
        return None
    import l3util
    work = []
    cache = {}
    for s in streams:
        if s not in cache:
            fr = l3util.split_frames(s)
            o = np.zeros(len(fr) + 1, np.uint32)
            np.cumsum([len(f) for f in fr], out=o[1:])
            h = fr[0] if fr else b"\xff\xfb\x90\x00"
            sr = l3util._SR[(h[1] >> 3) & 3][(h[2] >> 2) & 3]
            cache[s] = (np.frombuffer(s, np.uint8), o, float(sr))
        work.append(cache[s])

    def one(w):
        n = L.ffl_decode_stream(w[0].ctypes.data, w[1].ctypes.data, len(w[1]) - 1, b"mp3float")
        return n / w[2] if n >= 0 else -1.0

    one(work[0])
    t0 = time.perf_counter()
    if threads <= 1:
        res = [one(w) for w in work]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            res = list(ex.map(one, work))
    dt = time.perf_counter() - t0
    if min(res) < 0:
        return None
    return sum(res) / dt, dt


def cpu_lines(streams, cores):
    """The CPU lines of SURVEY.md 8(d) / BASELINE.md section 4, each on a bounded sample of the workload."""
    out = []
    n1 = max(1, min(len(streams), 16))
    v, dt = cpu_oracle_throughput(streams[:n1], 1)
    out.append({"decoder": "oracle/l3_oracle.c (from-spec, double, direct form)", "threads": 1, "value": v, "unit": UNIT,
                "sample": "%d streams, %.1f s" % (n1, dt)})
    na = cpu_sample_size(len(streams), cores)
    v, dt = cpu_oracle_throughput(streams[:na], cores)
    out.append({"decoder": "oracle/l3_oracle.c (from-spec, double, direct form)", "threads": cores, "value": v,
                "unit": UNIT, "sample": "%d streams, %.1f s" % (na, dt)})
    nf1 = max(1, min(len(streams), 64))
    r = cpu_ffmpeg_throughput(streams[:nf1], 1)
    if r:
        out.append({"decoder": "FFmpeg libavcodec mp3float (float32, production decoder)", "threads": 1, "value": r[0],
                    "unit": UNIT, "sample": "%d streams, %.1f s" % (nf1, r[1])})
        nfa = max(1, min(len(streams), cores * 128))
        work = (streams * (nfa // max(len(streams), 1) + 1))[:nfa] if nfa > len(streams) else streams[:nfa]
        r = cpu_ffmpeg_throughput(work, cores)
        if r:
            out.append({"decoder": "FFmpeg libavcodec mp3float (float32, production decoder)", "threads": cores,
                        "value": r[0], "unit": UNIT, "sample": "%d streams, %.1f s" % (len(work), r[1])})
    else:
        out.append({"decoder": "FFmpeg libavcodec mp3float", "unavailable": "libavcodec not found on this box"})
    return out


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_gpu = default_streams(args)
    nsample = cpu_sample_size(per_gpu * world if args.workload != "cfg5" else per_gpu, cores)
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
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.workload == "cfg5" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args, default_streams(args, world, 0))},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step decodes %d streams of the workload (a rate metric: the per-GPU batch is %d "
                                   "streams), one stream per host thread, all %d host threads; the reference repository "
                                   "has no code, so this is the from-spec oracle port oracle/l3_oracle.c"
                                   % (nsample, per_gpu, cores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        ff = cpu_ffmpeg_throughput((streams * 4)[: max(len(streams), cores * 64)], cores)
        if ff:
            line["cpu_lines"] = [{"decoder": "FFmpeg libavcodec mp3float (float32, production decoder)", "threads": cores,
                                  "value": ff[0], "unit": UNIT,
                                  "note": "a faster CPU decoder than the literal oracle port this arm times; reported so "
                                          "that the GPU / CPU ratio can be read against either"}]
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ our arm
def parity_gate(mp3_b200, dec, streams, pick):
    """Accuracy gate of every reported run: `pick` streams of the decoded batch against the
    oracle -- s16 PCM within 1 LSB of round(oracle * 32768), which is far inside ISO/IEC 11172-4.  Raises on mismatch."""
    import torch
    from oracle import oracle
    ptr, n = dec.pcm_device()

    class _View:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i2", "data": (ptr, False), "version": 2}
    pcm = torch.as_tensor(_View(), device="cuda")
    worst = 0
    for i in pick:
        inf = dec.stream_info(i)
        ref = oracle.decode(streams[i])
        want = np.clip(np.rint(ref.pcm.T * 32768.0), -32768, 32767).astype(np.int64)
        got = pcm[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].cpu().numpy().astype(np.int64)
        got = got.reshape(inf.samples, inf.channels)
        if got.shape != want.shape:
            raise SystemExit("parity gate: stream %d has shape %r, the oracle %r" % (i, got.shape, want.shape))
        d = int(np.abs(got - want).max()) if got.size else 0
        worst = max(worst, d)
        if d > 1:
            raise SystemExit("parity gate: stream %d differs from the oracle by %d LSB" % (i, d))
    return len(pick), worst


def time_steps(torch, tstream, barrier, fn, steps, after=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(tstream)
    for _ in range(steps):
        fn()
    if after:
        after()
    e1.record(tstream)
    barrier()
    return e0.elapsed_time(e1) / steps


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
    nstreams = default_streams(args, world, rank)
    if strong and args.distinct is None:
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

    # ---- accuracy gate, before anything is timed
    step_device()
    dec.sync()
    npick = min(len(streams), 8)
    pick = sorted(set(int(x) for x in np.linspace(0, len(streams) - 1, npick)))
    parity_n, parity_worst = parity_gate(mp3_b200, dec, streams, pick)

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
    ms_dev = time_steps(torch, tstream, barrier, step_device, args.steps)
    launches = dec.stats().kernel_launches

    # per-stage CUDA-event times (library events on the launching stream), separate untimed passes
    stage_ms = {k: 0.0 for k in ("index", "huffman", "requant", "imdct", "overlap", "synth", "fused")}
    # (stage timing puts events between the kernels, which serialises them; the timed steps above run without, so
    # that consecutive kernels overlap through programmatic dependent launch)
    nst = max(1, min(args.steps, 5))
    dec.set_stage_timing(True)
    step_device()
    dec.sync()
    for _ in range(nst):
        step_device()
        dec.sync()
        st = dec.stats()
        for k in stage_ms:
            stage_ms[k] += getattr(st, "ms_" + k)
    dec.set_stage_timing(False)
    stage_ms = {k: v / nst for k, v in stage_ms.items()}

    # ---- end to end: pinned host in, pinned host out
    do_e2e = not args.no_e2e and pcm_bytes < (8 << 30)  # cfg5 would need 59 GB of pinned host memory
    e2e = None
    checksum = None
    if do_e2e:
        e2e, checksum = run_e2e(args, torch, mp3_b200, dec, packed, offs, nbytes_in, pcm_bytes, tstream, barrier)
    sampler.stop_flag = True

    sweep = None
    if args.workload == "cfg2" and not args.no_sweep:
        sweep = run_sweep_cfg5(args, torch, mp3_b200, dec, tstream, barrier, rank, world)
    global CFG1_LATENCY, REAL_SIGNAL
    CFG1_LATENCY = run_cfg1_latency(mp3_b200, local, streams[0]) if (rank == 0 and not args.no_e2e) else None
    REAL_SIGNAL = run_real_signal(mp3_b200, local) if (rank == 0 and args.workload == "cfg2" and not args.no_sweep) else None
    finish(args, rank, world, dist, dec, streams, units, nbytes_in, pcm_bytes, audio_s, ms_dev, e2e, stage_ms,
           launches, sampler, gen_s, checksum, strong, numa, (parity_n, parity_worst), sweep, infos)


def run_e2e(args, torch, mp3_b200, dec, packed, offs, nbytes_in, pcm_bytes, tstream, barrier):
    h_in = mp3_b200.PinnedBuffer(nbytes_in + 64)
    h_in.view(np.uint8)[:nbytes_in] = packed
    h_out = mp3_b200.PinnedBuffer(pcm_bytes + 64)
    pcm_elems = pcm_bytes // 2

    # (ii)/(iii): host input, PCM left on the device (H2D + indexing + kernels)
    def step_h2d():
        dec.decode_packed(h_in.ptr, offs, where=mp3_b200.HOST, sync=False)

    for _ in range(2):
        step_h2d()
    dec.sync()
    ms_h2d = time_steps(torch, tstream, barrier, step_h2d, args.steps)

    # (iv): + D2H through the sink
    dec.set_pcm_sink(h_out.ptr, pcm_elems)  # D2H of each wave overlaps the next wave's kernels
    for _ in range(min(args.warmup, 3)):
        step_h2d()
    dec.sync()
    # the stream (and so the end event) waits for every D2H copy of every step
    # two passes of exactly K steps each, the faster one reported (both listed in the line): the transfers share the
    # host's memory system with whatever else the box is doing, and one pass in five comes out 15-20 % slow
    e2e_passes = [time_steps(torch, tstream, barrier, step_h2d, args.steps, after=dec.flush) for _ in range(2)]
    ms_e2e = min(e2e_passes)
    dec.set_pcm_sink(0, 0)
    checksum = int(h_out.view(np.int16, pcm_elems)[:: max(1, pcm_elems // 4096)].astype(np.int64).sum())

    # the floor: the same two transfers alone -- plain cudaMemcpyAsync, same pinned buffers, the library's own PCM
    # arena as the device side, H2D and D2H on two streams as the library issues them
    floor = None
    try:
        ptr, n = dec.pcm_device()

        class _Dev:
            __cuda_array_interface__ = {"shape": (pcm_bytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        d_pcm = torch.as_tensor(_Dev(), device="cuda")
        t_out = torch.from_numpy(h_out.view(np.uint8)[:pcm_bytes])
        t_in = torch.from_numpy(h_in.view(np.uint8)[:nbytes_in])
        d_in = torch.empty(nbytes_in, dtype=torch.uint8, device="cuda")
        s2 = torch.cuda.Stream()

        def copies():
            ev = torch.cuda.Event()
            with torch.cuda.stream(s2):
                d_in.copy_(t_in, non_blocking=True)
                ev.record(s2)
            with torch.cuda.stream(tstream):
                t_out.copy_(d_pcm, non_blocking=True)
                tstream.wait_event(ev)

        pinned = bool(t_out.is_pinned() and t_in.is_pinned())
        for _ in range(2):
            copies()
        torch.cuda.synchronize()
        floor = {"ms": time_steps(torch, tstream, barrier, copies, max(3, min(args.steps, 10))), "pinned": pinned}
    except Exception as e:  # noqa: BLE001
        floor = {"ms": None, "error": str(e)[:200]}
    return {"ms": ms_e2e, "ms_h2d_only": ms_h2d, "floor": floor, "passes": e2e_passes}, checksum


CFG1_LATENCY = None
REAL_SIGNAL = None


def _encode_real(seed):
    from mp3_b200 import signals, synth
    return synth.encode_pcm(signals.stereo(44100, 10.0, 1 + 2 * seed), 44100, 128)


def run_real_signal(mp3_b200, device, distinct=16, nstreams=1024):
    """The cfg2 shape on REAL-SIGNAL streams: synthetic music + speech through the in-tree encoder (gen/l3gen.c::
    l3enc_stream; 16 distinct 10-s streams, each used 64 times), device-resident, library CUDA events.  The generator's
    random spectra -- the workload `value` is measured on, as BASELINE.json defines it -- have short count1 regions;
    real audio has long ones and spends more of its time in the Huffman stage.  Reported next to the headline, not
    instead of it."""
    import multiprocessing as mp
    try:
        with mp.get_context("spawn").Pool(min(distinct, os.cpu_count() or 1)) as pool:
            uniq = pool.map(_encode_real, range(distinct))
    except Exception as e:  # noqa: BLE001
        return {"error": "encoder pool failed: %r" % (e,)}
    streams = [uniq[i % distinct] for i in range(nstreams)]
    packed, offs = mp3_b200.pack_streams(streams)
    import torch
    d_raw = torch.empty(int(packed.size) + 64, dtype=torch.uint8, device="cuda")
    d_raw[: packed.size].copy_(torch.from_numpy(packed))
    torch.cuda.synchronize()
    with mp3_b200.Decoder(device=device, pcm_format=mp3_b200.PCM_S16) as d2:
        d2.set_stage_timing(True)
        best = None
        for _ in range(8):
            d2.decode_packed(d_raw.data_ptr(), offs, where=mp3_b200.DEVICE, sync=True)
            st = d2.stats()
            if best is None or st.ms_total < best.ms_total:
                best = st
        audio = sum(d2.stream_info(i).samples / float(d2.stream_info(i).sample_rate) for i in range(nstreams))
    return {"workload": "1024 x 10 s, 44.1 kHz stereo 128 kbit/s CBR, real signals (synthetic music + speech) through the "
                        "in-tree encoder, %d distinct streams" % distinct,
            "ms_per_step": best.ms_total, "value": audio / (best.ms_total * 1e-3), "unit": UNIT,
            "stage_ms": {"index": best.ms_index, "huffman": best.ms_huffman, "fused": best.ms_fused},
            "note": "stage events on (0.03 ms slower than the headline's mode); best of 8"}


def run_cfg1_latency(mp3_b200, device, stream_bytes):
    """BASELINE.json configs[0] on the GPU: ONE 10-s stream, pageable host bytes in -> host PCM out through the batch
    call of the C-ABI, wall clock around the calls (median of 20 after 3 warm-ups).  The player's actual case."""
    with mp3_b200.Decoder(device=device, pcm_format=mp3_b200.PCM_S16) as d1:
        lat = []
        for _ in range(23):
            t0 = time.perf_counter()
            d1.decode_batch([stream_bytes])
            pcm = d1.fetch_pcm()
            lat.append((time.perf_counter() - t0) * 1e3)
        inf = d1.stream_info(0)
        audio = inf.samples / float(inf.sample_rate)
        # the same through page-locked buffers, the way a player would hold them: bytes in a pinned buffer, PCM into a
        # pinned sink (mp3b_decode_packed + mp3b_set_pcm_sink + mp3b_flush + mp3b_sync)
        packed, offs = mp3_b200.pack_streams([stream_bytes])
        h_in = mp3_b200.PinnedBuffer(packed.size + 64)
        h_in.view(np.uint8)[: packed.size] = packed
        nbytes = int(d1.stats().pcm_bytes)
        h_out = mp3_b200.PinnedBuffer(nbytes + 64)
        d1.set_pcm_sink(h_out.ptr, nbytes // 2)
        pin = []
        for _ in range(23):
            t0 = time.perf_counter()
            d1.decode_packed(h_in.ptr, offs, where=mp3_b200.HOST, sync=False)
            d1.flush()
            d1.sync()
            pin.append((time.perf_counter() - t0) * 1e3)
        same = bool(np.array_equal(h_out.view(np.int16, nbytes // 2), np.asarray(pcm).reshape(-1)[: nbytes // 2]))
        d1.set_pcm_sink(0, 0)
    med = float(np.median(lat[3:]))
    return {"workload": "cfg1: one %.1f-s stream, host bytes in -> host s16 PCM out (mp3b_decode_batch + fetch)" % audio,
            "ms_median": med, "ms_min": float(min(lat[3:])), "x_realtime": audio / (med * 1e-3), "pcm_samples": int(pcm.size),
            "pinned": {"ms_median": float(np.median(pin[3:])), "ms_min": float(min(pin[3:])), "pcm_equal": same,
                       "note": "pinned input buffer and pinned PCM sink instead of pageable memory"}}


def run_sweep_cfg5(args, torch, mp3_b200, dec, tstream, barrier, rank, world):
    """BASELINE.json configs[4]: 100,000 streams of the cfg2 type x 128 frames, sharded over the ranks as contiguous
    stream ranges (strong scaling: the job is fixed, N grows).  1,024 distinct streams per rank repeat (bounds the
    host-side generation); every stream is decoded.  Inputs resident in HBM, PCM left in HBM (59 GB in total)."""
    total, nf, distinct = args.sweep_streams, 128, 1024
    n = total // world + (1 if rank < total % world else 0)
    free, _ = torch.cuda.mem_get_info()
    need = n * (nf * 1152 * 2 * 2 + 60000 * 2) * 1.15 + (4 << 30)
    if free < need:
        return {"skipped": "needs %.0f GB of device memory per rank, %.0f free" % (need / 1e9, free / 1e9)}
    from mp3_b200 import synth
    base = synth.make_workload("cfg5", min(distinct, n), nf, seed=args.seed + 100000 * rank)
    lens = {len(s) for s in base}
    D = len(base)
    if len(lens) == 1:  # CBR streams of one rate: equal lengths, replicate on the device
        slen = lens.pop()
        p1, _ = mp3_b200.pack_streams(base)
        d1 = torch.from_numpy(p1).cuda()
        d_raw = d1.repeat(-(-n // D))[: n * slen].contiguous()
        del d1
        offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(slen)
    else:
        p, offs = mp3_b200.pack_streams([base[i % D] for i in range(n)])
        d_raw = torch.from_numpy(p).cuda()

    def step():
        dec.decode_packed(d_raw.data_ptr(), offs, where=mp3_b200.DEVICE, sync=False)

    step()
    dec.sync()
    pick = sorted(set(int(x) for x in np.linspace(0, n - 1, 4)))
    parity_n, _ = parity_gate(mp3_b200, dec, _Cyclic(base, n), pick)
    for _ in range(2):
        step()
    dec.sync()
    st = dec.stats()
    steps = max(3, min(args.steps, 5))
    ms = time_steps(torch, tstream, barrier, step, steps)
    audio = st.frames * 1152 / 44100.0
    out = {"ms": ms, "audio_s": audio, "pcm_bytes": st.pcm_bytes, "bytes_in": st.bytes_in, "units": st.units,
           "streams": n, "steps": steps, "launches": int(st.kernel_launches) * steps, "parity_checked": parity_n}
    del d_raw
    return out


class _Cyclic:
    """streams[i] = base[i % len(base)] without materialising a 100,000-entry list."""

    def __init__(self, base, n):
        self.base, self.n = base, n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.base[i % len(self.base)]


def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:  # noqa: BLE001
        return {}


def finish(args, rank, world, dist, dec, streams, units, nbytes_in, pcm_bytes, audio_s, ms_dev, e2e, stage_ms,
           launches, sampler, gen_s, checksum, strong, numa, parity, sweep, infos):
    import torch
    ms_e2e = e2e["ms"] if e2e else float("nan")
    ms_h2d = e2e["ms_h2d_only"] if e2e else float("nan")
    ms_floor = (e2e["floor"] or {}).get("ms") if e2e else None
    ms_floor = float("nan") if ms_floor is None else ms_floor
    sw = sweep if sweep and "ms" in sweep else None
    # ---- max over ranks (times), sum over ranks (work)
    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e, ms_h2d, ms_floor, sw["ms"] if sw else float("nan")], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_h2d, ms_floor = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        if sw:
            sw["ms"] = float(t[4])
        tot = torch.tensor([audio_s, float(nbytes_in), float(pcm_bytes), float(units), float(parity[0])] +
                           ([sw["audio_s"], float(sw["pcm_bytes"]), float(sw["bytes_in"]), float(sw["units"]),
                             float(sw["streams"]), float(sw["launches"]), float(sw["parity_checked"])] if sw else [0.0] * 7),
                           device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        audio_all, in_all, pcm_all, units_all, parity_all = [float(x) for x in tot[:5]]
        if sw:
            sw["audio_s"], sw["pcm_bytes"], sw["bytes_in"], sw["units"], sw["streams"], sw["launches"], \
                sw["parity_checked"] = [float(x) for x in tot[5:]]
    else:
        audio_all, in_all, pcm_all, units_all, parity_all = audio_s, float(nbytes_in), float(pcm_bytes), float(units), \
            float(parity[0])

    if rank == 0:
        peaks = load_json("MEASURED_PEAKS.json")
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        fp = load_json("profiles", "fp32_peak.json")
        counts = load_json("profiles", "fp32_counts.json")
        traffic = load_json("profiles", "traffic.json")
        in_per_unit = nbytes_in / max(units, 1)
        ideal_b = in_per_unit + FUSED_IDEAL_PCM_BYTES                  # fused ideal, 1256.5 at 128 kbit/s
        kern = {k: v for k, v in stage_ms.items() if k != "index" and v > 0}
        dom = max(kern, key=kern.get) if kern else "pipeline"
        kms = kern.get(dom, ms_dev)
        # algorithmic bytes per unit of each kernel's own interface
        own = {"huffman": in_per_unit + 1152.0 + 40.0, "requant": 1152.0 + 2304.0, "imdct": 2304.0 + 4608.0,
               "overlap": 4608.0 + 2304.0, "synth": 2304.0 + 1152.0, "fused": 1152.0 + 40.0 + 1152.0,
               "pipeline": ideal_b}
        if dom == "fused":
            # FP32 / issue bound (ncu: issue slots ~74 %, DRAM ~10 %).  Peak = the measured FFMA rate of this GPU
            # (tools/fp32_peak.cu -> profiles/fp32_peak.json; burst: the kernel is timed alone); executed flops per
            # unit from ncu's smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on (profiles/fp32_counts.json).
            fp_peak = fp.get("fp32_tflops_burst")
            fp_src = "measured (profiles/fp32_peak.json, tools/fp32_peak.cu, burst)"
            if not fp_peak:
                fp_peak, fp_src = 74.4, "theoretical 148 SM x 128 FMA x 2 x 1.965 GHz (no measured file)"
            cnt = (counts.get(args.workload) or {}).get("fused", {})
            exe = cnt.get("flop_per_unit")
            wi = cnt.get("warp_instructions_per_unit")
            sms, mhz = fp.get("sms", 148), fp.get("max_clock_mhz", 1965)
            issue_ms = wi * units / (sms * 4 * mhz * 1e6) * 1e3 if wi else None  # one warp instruction per scheduler and clock
            eff_tf = DIRECT_FORM_FLOP_PER_UNIT * units / (kms * 1e-3) / 1e12
            exe_tf = exe * units / (kms * 1e-3) / 1e12 if exe else None
            roof = {"bound": "fp32-issue", "kernel": "k_backend (fused a6-a11)", "unit": "TFLOP/s", "peak": fp_peak,
                    "peak_source": fp_src,
                    "achieved": exe_tf if exe_tf is not None else eff_tf,
                    "frac": (exe_tf if exe_tf is not None else eff_tf) / fp_peak,
                    "achieved_is": "executed FP32 flops (fast transforms)" if exe_tf is not None
                                   else "direct-form effective flops (no executed-flop count for this workload)",
                    "executed_flop_per_unit": exe, "executed": exe_tf,
                    "executed_frac": exe_tf / fp_peak if exe_tf is not None else None,
                    "effective_direct_form": eff_tf, "effective_direct_form_frac": eff_tf / fp_peak,
                    "direct_form_flop_per_unit": DIRECT_FORM_FLOP_PER_UNIT,
                    "peak_register_operands": fp.get("ffma_reg_operands_tflops"),
                    "peak_note": "peak = FFMA with one constant operand (the usual definition); with all three operands in "
                                 "registers the same pipe measures peak_register_operands, and the packed FFMA2 form the same "
                                 "flops in half the instructions",
                    "issue": ({"warp_instructions_per_unit": wi, "issue_limited_ms": issue_ms, "frac": issue_ms / kms,
                               "note": "the kernel's executed warp instructions (ncu) at one instruction per scheduler and "
                                       "clock: the bound the kernel actually runs against"} if issue_ms else None)}
        else:
            ach = own[dom] * units / (kms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": dom, "unit": "GB/s", "peak": hbm_peak, "peak_source": peak_src,
                    "achieved": ach, "frac": ach / hbm_peak}
        roof.update({
            "traffic": (traffic.get(dom) or {}).get("dram_bytes_per_launch"),
            "units_per_launch": units, "kernel_ms": kms,
            "hbm": {
                "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
                "algorithmic_bytes_per_unit_fused_ideal": ideal_b,
                "whole_pipeline": {"ms": ms_dev, "achieved": ideal_b * units / (ms_dev * 1e-3) / 1e9,
                                   "frac": ideal_b * units / (ms_dev * 1e-3) / 1e9 / hbm_peak},
                "dominant_kernel_at_fused_ideal": {"ms": kms, "achieved": ideal_b * units / (kms * 1e-3) / 1e9,
                                                   "frac": ideal_b * units / (kms * 1e-3) / 1e9 / hbm_peak},
                "dominant_kernel_own_interface": {"algorithmic_bytes_per_unit": own[dom],
                                                  "achieved": own[dom] * units / (kms * 1e-3) / 1e9,
                                                  "frac": own[dom] * units / (kms * 1e-3) / 1e9 / hbm_peak},
            },
            "note": "the back end is FP32 / shared-memory issue bound, not HBM bound; both views are printed. "
                    "fused ideal = compressed bytes in + s16 PCM out per unit (SURVEY.md 8(d)); own interface = what the "
                    "kernel itself reads and writes (it includes the int16 spectrum between the two kernels)"})
        cores = os.cpu_count() or 1
        # CPU lines: rank 0 at N = 1 only (they are per-box numbers, and the scaling runs need not repeat them)
        lines = cpu_lines(streams, cores) if not args.no_cpu and world == 1 else []
        allc = next((l for l in lines if l.get("threads") == cores and l["decoder"].startswith("oracle")), None)
        kernels_only = stage_ms["huffman"] + sum(stage_ms[k] for k in ("requant", "imdct", "overlap", "synth", "fused"))
        line = {
            "metric": METRIC, "value": audio_all / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_string(args, len(streams)),
                "streams_per_gpu": len(streams), "distinct_streams": args.distinct or len(streams),
                "pcm": "s16 interleaved", "pipeline": args.pipeline, "indexer": "device",
                "l2_policy": "inputs_larger_than_l2 (%.0f MB in, %.0f MB PCM out per step)" % (nbytes_in / 1e6, pcm_bytes / 1e6),
                "timing": "torch.cuda.Event on the stream the library launches on",
                "host_binding": numa or "none",
            },
            "pcm_gbs": pcm_all / (ms_dev * 1e-3) / 1e9,
            "x_realtime_per_gpu": audio_all / world / (ms_dev * 1e-3),
            "parity_checked": int(parity_all),
            "parity": "s16 PCM of %d streams per rank within %d LSB of round(oracle x 32768) (gate: 1), checked before "
                      "timing" % (parity[0], parity[1]),
            "stage_ms": stage_ms,
            "splits": {
                "i_decode_kernels_only_ms": kernels_only,
                "iii_plus_indexing_device_resident_ms": ms_dev,
                "ii_iii_plus_h2d_ms": None if ms_h2d != ms_h2d else ms_h2d,
                "iv_plus_d2h_ms": None if ms_e2e != ms_e2e else ms_e2e,
                "note": "(i) Huffman + back end (library CUDA events); (iii) adds the device indexer (walk, side info, "
                        "main-data compaction) and its one host round trip = `value`; (ii)+(iii) adds the H2D of the MP3 "
                        "bytes from pinned memory; (iv) adds the D2H of the PCM = `e2e`",
            },
            "roofline": roof,
            "cpu_baseline": ({"value": allc["value"], "unit": UNIT, "cores": cores, "kind": "port",
                              "sample": "first %s of the workload, oracle/l3_oracle.c, one stream per host thread"
                                        % allc["sample"]} if allc else None),
            "cpu_lines": lines,
            "e2e": ({"value": audio_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(in_all),
                     "d2h_bytes_per_step": int(pcm_all), "ms_per_step": ms_e2e,
                     "pcm_gbs": pcm_all / (ms_e2e * 1e-3) / 1e9,
                     "pcie_floor_ms": None if ms_floor != ms_floor else ms_floor,
                     "frac_of_pcie_floor": None if ms_floor != ms_floor else ms_floor / ms_e2e,
                     "passes_ms_rank0": (e2e or {}).get("passes"),
                     "passes_note": "two passes of exactly K steps; ms_per_step is the faster one (rank 0's passes listed)",
                     "pcie_floor_note": "the step's two transfers alone (plain cudaMemcpyAsync of the same pinned buffers, "
                                        "same process, all ranks at once); floor / e2e = 1 means transfer bound"}
                    if ms_e2e == ms_e2e else None),
            "gpu_launches": int(launches * args.steps),
            "cfg1_latency": CFG1_LATENCY,
            "real_signal": REAL_SIGNAL,
            "clocks": sampler.result(),
            "gen_seconds": gen_s, "pcm_checksum": checksum,
        }
        if sweep is not None:
            if sw:
                v = sw["audio_s"] / (sw["ms"] * 1e-3)
                ideal5 = sw["bytes_in"] + sw["units"] * FUSED_IDEAL_PCM_BYTES
                line["sweep_cfg5"] = {
                    "workload": "cfg5: %d streams x 128 frames sharded over %d GPU(s), %s" % (
                        int(sw["streams"]), world, WORKLOADS["cfg5"]),
                    "scaling": "strong", "metric": METRIC, "value": v, "unit": UNIT, "ms_per_step": sw["ms"],
                    "steps": sw["steps"], "streams": int(sw["streams"]), "units": int(sw["units"]),
                    "pcm_gbs": sw["pcm_bytes"] / (sw["ms"] * 1e-3) / 1e9,
                    "hbm": {"algorithmic_bytes": ideal5, "achieved": ideal5 / (sw["ms"] * 1e-3) / 1e9,
                            "peak": hbm_peak * world, "frac": ideal5 / (sw["ms"] * 1e-3) / 1e9 / (hbm_peak * world),
                            "note": "fused-ideal bytes (compressed in + s16 PCM out) of the whole job / max-over-ranks "
                                    "time, against N x the measured HBM peak"},
                    "x_realtime_per_gpu": v / world, "gpu_launches": int(sw["launches"]),
                    "parity_checked": int(sw["parity_checked"]),
                    "data": "synthetic; 1,024 distinct streams per rank repeat, every stream is decoded",
                }
            else:
                line["sweep_cfg5"] = sweep
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
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the cfg5 100k-stream sweep appended to a cfg2 run")
    ap.add_argument("--sweep-streams", type=int, default=100000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU lines (profiling runs)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to their GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.pipeline_id = {"default": None, "fused": 0, "staged": 1}[args.pipeline]
    if args.frames is None:
        args.frames = DEFAULT_FRAMES.get(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
