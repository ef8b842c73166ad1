"""ctypes driver for the FFmpeg `mp3float` decoder bundled in the opencv wheel.

Independent second opinion for the oracle: the only MP3 decoder on the
box that was not written in this repository.  No headers are installed, so struct fields are
read at the hand-verified offsets from the survey (AVPacket.data @24, .size @32;
AVFrame.extended_data @96, .nb_samples @112, .format @116).  Tests that use this skip when the
library cannot be loaded.
"""
import ctypes
import glob
import os
import sys

import numpy as np

_lib = None
_tried = False


def _find_dir():
    for base in sys.path:
        for sub in ("opencv_python_headless.libs", "opencv_python.libs"):
            d = os.path.join(base, sub)
            if glob.glob(os.path.join(d, "libavcodec*.so*")):
                return d
    return None


def libavcodec_path():
    d = _find_dir()
    if not d:
        return None
    return glob.glob(os.path.join(d, "libavcodec*.so*"))[0]


def load():
    """Return (avcodec, avutil) CDLLs or None."""
    global _lib, _tried
    if _tried:
        return _lib
    _tried = True
    d = _find_dir()
    if not d:
        return None
    pending = sorted(glob.glob(os.path.join(d, "*.so*")))
    loaded = {}
    for _ in range(6):
        rest = []
        for f in pending:
            try:
                loaded[os.path.basename(f)] = ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                rest.append(f)
        pending = rest
        if not pending:
            break
    avc = next((v for k, v in loaded.items() if k.startswith("libavcodec")), None)
    avu = next((v for k, v in loaded.items() if k.startswith("libavutil")), None)
    if avc is None or avu is None:
        return None
    avu.av_log_set_level(-8)
    avc.avcodec_find_decoder_by_name.restype = ctypes.c_void_p
    avc.avcodec_find_decoder_by_name.argtypes = [ctypes.c_char_p]
    avc.avcodec_alloc_context3.restype = ctypes.c_void_p
    avc.avcodec_alloc_context3.argtypes = [ctypes.c_void_p]
    avc.avcodec_open2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    avc.av_packet_alloc.restype = ctypes.c_void_p
    avc.avcodec_send_packet.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    avc.avcodec_receive_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    avc.avcodec_free_context.argtypes = [ctypes.c_void_p]
    avu.av_frame_alloc.restype = ctypes.c_void_p
    avu.av_frame_unref.argtypes = [ctypes.c_void_p]
    _lib = (avc, avu)
    return _lib


def available():
    return load() is not None


def decode_frames(frames, nch, codec_name=b"mp3float"):
    """Decode a list of `bytes` (one frame each, in stream order) with one decoder context
    (mp3float for Layer III, mp2float for Layer II).

    Returns float32 array [nch, nsamples] (raw codec output, no delay trimming)."""
    avc, avu = load()
    codec = avc.avcodec_find_decoder_by_name(codec_name)
    assert codec
    ctx = avc.avcodec_alloc_context3(codec)
    assert avc.avcodec_open2(ctx, codec, None) == 0
    pkt = avc.av_packet_alloc()
    frm = avu.av_frame_alloc()
    out = []
    for fb in frames:
        buf = ctypes.create_string_buffer(bytes(fb) + b"\0" * 64, len(fb) + 64)
        ctypes.c_void_p.from_address(pkt + 24).value = ctypes.addressof(buf)
        ctypes.c_int.from_address(pkt + 32).value = len(fb)
        r = avc.avcodec_send_packet(ctx, pkt)
        if r != 0:
            out.append(None)
            continue
        r = avc.avcodec_receive_frame(ctx, frm)
        if r != 0:
            out.append(None)
            continue
        n = ctypes.c_int.from_address(frm + 112).value
        ext = ctypes.c_void_p.from_address(frm + 96).value
        chans = []
        for c in range(nch):
            ptr = ctypes.c_void_p.from_address(ext + 8 * c).value
            chans.append(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), (n,)).copy())
        out.append(np.stack(chans))
        avu.av_frame_unref(frm)
    cp = ctypes.c_void_p(ctx)
    avc.avcodec_free_context(ctypes.byref(cp))
    good = [o for o in out if o is not None]
    if not good:
        return np.zeros((nch, 0), np.float32), out
    n = good[0].shape[1]
    full = [o if o is not None else np.zeros((nch, n), np.float32) for o in out]
    return np.concatenate(full, axis=1), out
