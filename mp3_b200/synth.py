"""Synthetic Layer III stream generator (ctypes binding of gen/l3gen.c) and the named workloads
of BASELINE.json `configs`.

The reference repository ships no audio and no generator (/root/reference/README.md:1-84), and
there is no MP3 encoder on the box, so all inputs are synthesised: random, spec-valid frames.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gen", "l3gen.c")
LIB = os.path.join(HERE, "gen", "libl3gen.so")


class Cfg(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("sample_rate", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("bitrate_kbps", ctypes.c_int32), ("vbr_min_kbps", ctypes.c_int32),
                ("vbr_max_kbps", ctypes.c_int32), ("nframes", ctypes.c_int32), ("blocks", ctypes.c_int32),
                ("mixed_pct", ctypes.c_int32), ("reservoir", ctypes.c_int32), ("fill_lo_pct", ctypes.c_int32),
                ("fill_hi_pct", ctypes.c_int32), ("mode_ext_mask", ctypes.c_int32), ("crc", ctypes.c_int32),
                ("lsf_avoid_illegal_ispos", ctypes.c_int32), ("level_lo_db", ctypes.c_int32),
                ("level_hi_db", ctypes.c_int32), ("only_table", ctypes.c_int32), ("scfsi_pct", ctypes.c_int32),
                ("max_linbits_value", ctypes.c_int32), ("mixed_free", ctypes.c_int32), ("tag", ctypes.c_int32),
                ("tag_lame", ctypes.c_int32), ("enc_delay", ctypes.c_int32), ("enc_padding", ctypes.c_int32),
                ("layer", ctypes.c_int32)]


DEFAULTS = dict(seed=20261018, sample_rate=44100, mode=0, bitrate_kbps=128, vbr_min_kbps=0, vbr_max_kbps=0,
                nframes=383, blocks=0, mixed_pct=0, reservoir=1, fill_lo_pct=85, fill_hi_pct=100,
                mode_ext_mask=15, crc=0, lsf_avoid_illegal_ispos=0, level_lo_db=14, level_hi_db=40,
                only_table=0, scfsi_pct=25, max_linbits_value=0, mixed_free=0, tag=0, tag_lame=0, enc_delay=0,
                enc_padding=0, layer=0)


def build(force=False):
    deps = [SRC, os.path.join(HERE, "csrc", "iso_tables.h"), os.path.join(HERE, "csrc", "iso_tables_gen.h"),
            os.path.join(HERE, "csrc", "iso_tables_l2.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.l3gen_stream.restype = ctypes.c_size_t
        _lib.l3gen_stream.argtypes = [ctypes.POINTER(Cfg), ctypes.c_void_p, ctypes.c_size_t]
        _lib.l3gen_max_bytes.restype = ctypes.c_size_t
        _lib.l3gen_max_bytes.argtypes = [ctypes.POINTER(Cfg)]
        _lib.l3enc_stream2.restype = ctypes.c_size_t
        _lib.l3enc_stream2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_size_t]
    return _lib


def encode_pcm(pcm, sample_rate=44100, bitrate_kbps=128, short_blocks=False):
    """Encode real audio: pcm = int16 array [samples] (mono) or [samples, 2]; returns the MPEG-1 Layer III stream
    (gen/l3gen.c::l3enc_stream2: CBR, bit reservoir in use, no psychoacoustic model; short_blocks: window switching on
    attacks -- start / short / stop windows --, else long blocks only)."""
    L = _load()
    a = np.ascontiguousarray(pcm, dtype=np.int16)
    nch = 1 if a.ndim == 1 else a.shape[1]
    n = a.shape[0]
    cap = (144 * bitrate_kbps * 1000 // sample_rate + 1) * ((n + 1151) // 1152 + 2) + 64
    buf = np.zeros(cap, np.uint8)
    m = L.l3enc_stream2(a.ctypes.data_as(ctypes.c_void_p), n, nch, sample_rate, bitrate_kbps, 1 if short_blocks else 0,
                        buf.ctypes.data_as(ctypes.c_void_p), cap)
    if m == 0:
        raise ValueError("encoder rejected the arguments")
    return buf[:m].tobytes()


def make_stream(**kw):
    """Generate one stream; returns bytes.  Keyword arguments override DEFAULTS."""
    L = _load()
    d = dict(DEFAULTS)
    d.update(kw)
    cfg = Cfg(**d)
    cap = L.l3gen_max_bytes(ctypes.byref(cfg)) + 64
    buf = np.zeros(cap, np.uint8)
    n = L.l3gen_stream(ctypes.byref(cfg), buf.ctypes.data_as(ctypes.c_void_p), cap)
    if n == 0:
        raise ValueError("generator rejected config %r" % (d,))
    return buf[:n].tobytes()


# ---------------------------------------------------------------- BASELINE.json workloads
def workload_cfgs(name, nstreams=None, nframes=None, seed=20261018):
    """Per-stream generator configs for the named workloads (BASELINE.json `configs`).

    cfg1: one 44.1 kHz stereo 128 kbps CBR stream, 383 frames (10 s), long blocks
    cfg2: 1024 such streams (seed + stream id)
    cfg3: 320 kbps joint stereo (MS+IS), mixed long/short/switching windows, reservoir swept
    cfg4: MPEG-2 LSF 22.05/24 kHz + MPEG-1 44.1 kHz, VBR, mono and stereo mixed
    cfg5: 100k streams of the cfg2 type x 128 frames
    """
    out = []
    if name == "cfg1":
        n, nf = nstreams or 1, nframes or 383
        for i in range(n):
            out.append(dict(seed=seed + i, nframes=nf))
    elif name == "cfg2":
        n, nf = nstreams or 1024, nframes or 383
        for i in range(n):
            out.append(dict(seed=seed + i, nframes=nf))
    elif name == "cfg3":
        n, nf = nstreams or 1024, nframes or 383
        for i in range(n):
            out.append(dict(seed=seed + i, nframes=nf, bitrate_kbps=320, mode=1, blocks=1, mixed_pct=25,
                            fill_lo_pct=35, fill_hi_pct=100))
    elif name == "cfg4":
        n, nf = nstreams or 1024, nframes or 383
        for i in range(n):
            k = i % 3
            mode = 3 if (i // 3) % 2 == 0 else (1 if (i // 6) % 2 else 0)
            if k == 0:
                out.append(dict(seed=seed + i, nframes=nf, sample_rate=22050, vbr_min_kbps=8, vbr_max_kbps=160,
                                mode=mode, blocks=1, fill_lo_pct=50))
            elif k == 1:
                out.append(dict(seed=seed + i, nframes=nf, sample_rate=24000, vbr_min_kbps=8, vbr_max_kbps=160,
                                mode=mode, blocks=1, fill_lo_pct=50))
            else:
                out.append(dict(seed=seed + i, nframes=nf, sample_rate=44100, vbr_min_kbps=32, vbr_max_kbps=320,
                                mode=mode, blocks=1, mixed_pct=25, fill_lo_pct=50))
    elif name == "cfg5":
        n, nf = nstreams or 100000, nframes or 128
        for i in range(n):
            out.append(dict(seed=seed + i, nframes=nf))
    else:
        raise KeyError(name)
    return out


def make_workload(name, nstreams=None, nframes=None, seed=20261018, distinct=None):
    """List of stream byte strings for a named workload.

    `distinct` bounds the number of distinct streams generated (the rest repeat cyclically); the
    bench says so in its `config` when it uses this to bound host-side generation time."""
    from concurrent.futures import ThreadPoolExecutor
    cfgs = workload_cfgs(name, nstreams, nframes, seed)
    _load()
    nd = len(cfgs) if distinct is None else min(distinct, len(cfgs))
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:  # the C call releases the GIL
        base = list(ex.map(lambda c: make_stream(**c), cfgs[:nd]))
    return [base[i % nd] for i in range(len(cfgs))]
