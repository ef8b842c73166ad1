"""ctypes binding of the CPU oracle (oracle/l3_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under mp3_b200/ imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class Info(ctypes.Structure):
    _fields_ = [("sample_rate", ctypes.c_int), ("channels", ctypes.c_int), ("lsf", ctypes.c_int),
                ("frames", ctypes.c_long), ("units", ctypes.c_long), ("samples", ctypes.c_long),
                ("concealed_frames", ctypes.c_long)]


class Tag(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("has_lame", ctypes.c_int), ("frames", ctypes.c_uint), ("bytes", ctypes.c_uint),
                ("enc_delay", ctypes.c_int), ("enc_padding", ctypes.c_int), ("first_sample", ctypes.c_long),
                ("num_samples", ctypes.c_long)]


def build(force=False):
    srcs = [os.path.join(HERE, "l3_oracle.c"),
            os.path.join(HERE, "..", "mp3_b200", "csrc", "iso_tables.h"),
            os.path.join(HERE, "..", "mp3_b200", "csrc", "iso_tables_gen.h")]
    if not force and os.path.exists(LIB) and all(
            not os.path.exists(s) or os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.check_call(["make", "-C", HERE, "-s", "-B", "liboracle.so"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.l3o_decode.restype = ctypes.c_int
        _lib.l3o_decode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Info), ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
        _lib.l3o_dwin.restype = ctypes.c_double
        _lib.l3o_init()
    return _lib


def parse_tag(data):
    """Xing / Info / LAME / VBRI tag of a stream and the gapless window it implies (l3o_parse_tag)."""
    L = lib()
    data = bytes(data)
    buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
    t = Tag()
    rc = L.l3o_parse_tag(buf, ctypes.c_size_t(len(data)), ctypes.byref(t))
    return t if rc == 0 else None


class Decoded:
    """Result of decode(): pcm [channels, samples] float64 (full scale +-1), plus optional dumps."""


def decode(data, dumps=False, want_pcm=True, verify_crc=False):
    L = lib()
    L.l3o_set_verify_crc(1 if verify_crc else 0)  # process-wide switch: tests that use it are single-threaded
    data = bytes(data)
    n = len(data)
    buf = (ctypes.c_uint8 * max(n, 1)).from_buffer_copy(data if n else b"\0")
    cap_frames = n // 24 + 2          # smallest legal Layer III frame is 24 bytes
    cap = cap_frames * 1152
    info = Info()
    pcm = np.zeros((cap, 2), np.float64) if want_pcm else None
    max_units = cap_frames * 4
    d_is = np.zeros((max_units, 576), np.int16) if dumps else None
    d_sf = np.zeros((max_units, 40), np.uint8) if dumps else None
    # dumps="int": only the integer stages (Huffman output, scalefactors) -- what the full-size parity tests need
    d_xr = np.zeros((max_units, 576), np.float64) if dumps and dumps != "int" else None
    d_sb = np.zeros((max_units, 576), np.float64) if dumps and dumps != "int" else None

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    # pcm is laid out [sample][channel]; for mono the oracle writes stride 1
    r = L.l3o_decode(buf, n, ctypes.byref(info), ptr(pcm), cap, ptr(d_is), ptr(d_sf), ptr(d_xr), ptr(d_sb))
    out = Decoded()
    out.rc = r
    out.info = info
    out.sample_rate, out.channels = info.sample_rate, info.channels
    out.frames, out.units, out.samples = info.frames, info.units, info.samples
    out.concealed_frames = info.concealed_frames
    if r != 0:
        out.pcm = np.zeros((0, 0))
        return out
    if want_pcm:
        flat = pcm.reshape(-1)[: info.samples * info.channels]
        out.pcm = flat.reshape(info.samples, info.channels).T.copy()
    if dumps:
        u = info.units
        out.is_ = d_is[:u]
        out.sf = d_sf[:u]
        if d_xr is not None:
            out.xr = d_xr[:u]
            out.sb = d_sb[:u].reshape(u, 18, 32)
    return out
