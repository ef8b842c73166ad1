"""The time-parallel frame walk's logic on the CPU: tests/c/walk_emu.cpp runs the phase functions of
mp3_b200/csrc/walk_par.h (the ones the CUDA kernels k_walk_* call) one thread after the other; the frame
table must equal the serial host walk (mp3b_index_stream_host) for every segment length, on healthy and damaged streams,
and healthy streams must need no repair at the default segment length."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("walk") / "libwalkemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "mp3_b200", "csrc"),
                           os.path.join(ROOT, "tests", "c", "walk_emu.cpp"), "-o", so])
    L = ctypes.CDLL(so)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.walk_emu.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32,
                           ctypes.c_void_p, ctypes.c_uint32, u32p, u32p, u32p, u32p, u32p, u32p, u32p]
    return L


def run(L, data, seg):
    data = bytes(data)
    buf = np.frombuffer(data if data else b"\0", np.uint8)
    cap = len(data) // 24 + 2
    out = np.zeros((cap, 4), np.uint32)
    v = [ctypes.c_uint32() for _ in range(5)]
    tag = (ctypes.c_uint32 * 4)()
    bad = ctypes.c_uint32()
    rc = L.walk_emu(buf.ctypes.data, len(data), seg, 0, 0, 0, out.ctypes.data, cap, ctypes.byref(v[0]), ctypes.byref(v[1]),
                    ctypes.byref(v[2]), ctypes.byref(v[3]), ctypes.byref(v[4]), tag, ctypes.byref(bad))
    assert rc == 0
    n = v[0].value
    return out[:n], bad.value, tag[0]


def serial(data):
    import mp3_b200
    fr, info, tag = mp3_b200.index_stream_host(data)
    return fr, tag


def plant_false_entries(s, every=2500, start=2000):
    """Copies of the stream's own second header written into the main data of later frames, each followed by another
    copy one frame length on: positions that look exactly like the chain to a walk that starts in the middle (a wrong
    guess of the time-parallel walk), while the chain itself never reads them."""
    import mp3_b200
    fr, _, _ = mp3_b200.index_stream_host(s)
    a = bytearray(s)
    if len(fr) < 8:
        return bytes(a)
    hdr = bytes(a[int(fr["offset"][1]): int(fr["offset"][1]) + 4])
    flen = int(fr["offset"][2]) - int(fr["offset"][1])
    starts = fr["offset"].astype(np.int64)
    for x in range(start, len(a) - 2 * flen - 8, every):
        # keep clear of the real headers and side info: 40 bytes behind a frame start at least, both copies
        ok = all(np.min(np.abs(starts - y - d)) > 44 for y in (x, x + flen) for d in (0,))
        if ok:
            a[x: x + 4] = hdr
            a[x + flen: x + flen + 4] = hdr
    return bytes(a)


@pytest.fixture(scope="module")
def streams(synth_mod):
    allc = dict(cases.FF)
    allc.update(cases.EXTRA)
    allc.update(cases.L2)
    allc.update(cases.L1)
    good = [synth_mod.make_stream(**allc[n]) for n in sorted(allc)]
    good.append(synth_mod.make_stream(nframes=300, seed=9, mode=1, bitrate_kbps=320, blocks=1, mixed_pct=25, fill_lo_pct=35))
    good.append(synth_mod.make_stream(nframes=400, seed=10, vbr_min_kbps=32, vbr_max_kbps=320, blocks=1))
    good.append(synth_mod.make_stream(nframes=8, seed=5, tag=2, tag_lame=1, enc_delay=576, enc_padding=1000, mode=1))
    rng = np.random.default_rng(77)
    bad = []
    id3 = b"ID3\x03\x00\x00" + bytes([0, 0, 3, 10]) + bytes(394)
    for s in good[::2]:
        a = np.frombuffer(s, np.uint8).copy()
        for n in (1, 5, 30, 200):
            b = a.copy()
            idx = rng.integers(0, b.size, n)
            b[idx] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
            bad.append(b.tobytes())
        bad.append(a[: rng.integers(1, a.size)].tobytes())
        bad.append(a[rng.integers(1, 700):].tobytes())
        bad.append(id3 + s)
        bad.append(s[:900] + bytes(rng.integers(0, 256, 1500, dtype=np.uint8)) + s[900:])
    # wrong guesses by construction: false chain entries inside the main data, every few segments
    bad += [plant_false_entries(good[-3]), plant_false_entries(good[-2], every=1700, start=900)]
    bad += [b"", b"\xff" * 40, bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), bytes([0xFF, 0xFB, 0x90, 0x00]) * 300]
    return good, bad


@pytest.mark.parametrize("seg", [48, 240, 1032, 4096, 65520])
def test_emulated_parallel_walk_equals_serial(seg, emu, streams):
    good, bad = streams
    for k, s in enumerate(good + bad):
        want, wtag = serial(s)
        got, nbad, tkind = run(emu, s, seg)
        assert got.shape[0] == want.shape[0], (k, seg)
        assert np.array_equal(got[:, 0], want["offset"]) and np.array_equal(got[:, 1], want["payload_offset"]) and \
            np.array_equal(got[:, 2], want["header"]), (k, seg)
        assert (tkind & 0xff) == wtag.kind


def test_healthy_streams_need_no_repair(emu, streams):
    good, _ = streams
    for k, s in enumerate(good):
        _, nbad, _ = run(emu, s, 4096)
        assert nbad == 0, k
