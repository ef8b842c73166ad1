#!/usr/bin/env python3
"""Regenerate tests/golden/ff_golden.npz.

For a handful of generator cases: the stream bytes and the PCM that FFmpeg's `mp3float` decoder
(libavcodec 62.11, the only independent MP3 decoder on the build box; see tests/ffmpeg_ref.py)
produced for them, as float32.  The fixture lets the oracle be pinned on machines where that
library is absent.  The reference repository holds no vectors of its own
(/root/reference has no tests or audio), so these are the golden vectors for the path.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import ffmpeg_ref  # noqa: E402
import l3util  # noqa: E402
from mp3_b200 import synth  # noqa: E402

PICK = ["cfg1_long_cbr128", "mixed_blocks", "ms_plus_intensity", "cfg3_320k_joint", "mono", "lsf22_stereo",
        "lsf24_joint", "lsf16_mono", "vbr_32_320", "48k_crc", "m25_12k_joint", "m25_8k_stereo_vbr"]


def main():
    assert ffmpeg_ref.available(), "libavcodec not found"
    out = {}
    for name in PICK:
        kw = dict(cases.FF[name])
        kw["nframes"] = 6
        s = synth.make_stream(**kw)
        frames = l3util.split_frames(s)
        nch = 1 if kw.get("mode", 0) == 3 else 2
        pcm, per = ffmpeg_ref.decode_frames(frames, nch)
        assert all(p is not None for p in per)
        out[name + ".mp3"] = np.frombuffer(s, np.uint8)
        out[name + ".pcm"] = pcm.astype(np.float32)
    # Layer II / Layer I: FFmpeg's mp2float / mp1float
    for group, codec in ((cases.L2, b"mp2float"), (cases.L1, b"mp1float")):
        for name in sorted(group)[:3]:
            kw = dict(group[name])
            kw["nframes"] = 6
            s = synth.make_stream(**kw)
            frames = l3util.split_frames(s)
            nch = 1 if kw.get("mode", 0) == 3 else 2
            pcm, per = ffmpeg_ref.decode_frames(frames, nch, codec)
            assert all(p is not None for p in per)
            out[name + ".mp3"] = np.frombuffer(s, np.uint8)
            out[name + ".pcm"] = pcm.astype(np.float32)
    # real signals through the in-tree encoder (gen/l3gen.c::l3enc_stream): music + speech, bit reservoir in use
    from mp3_b200 import signals  # noqa: E402
    for name, sr, nch, kbps in (("enc_stereo_44k_128", 44100, 2, 128), ("enc_mono_44k_64", 44100, 1, 64)):
        src = signals.stereo(sr, 0.6) if nch == 2 else signals.to_s16(signals.speech(sr, 0.6) * 0.8)
        s = synth.encode_pcm(src, sr, kbps)
        frames = l3util.split_frames(s)
        pcm, per = ffmpeg_ref.decode_frames(frames, nch)
        assert all(p is not None for p in per)
        out[name + ".mp3"] = np.frombuffer(s, np.uint8)
        out[name + ".pcm"] = pcm.astype(np.float32)
    # the survey's known-answer frames
    kat0 = bytes.fromhex("fffb9000") + bytes(413)
    side = bytes.fromhex("0000000401690021080000000D20042100000001A40084200000003480108400")
    kat1 = bytes.fromhex("fffb9000") + side + bytes([0x50]) + bytes(417 - 4 - 32 - 1)
    pcm, _ = ffmpeg_ref.decode_frames([kat1, kat0, kat0], 2)
    out["kat1.mp3"] = np.frombuffer(kat1 + kat0 + kat0, np.uint8)
    out["kat1.pcm"] = pcm.astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ff_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
