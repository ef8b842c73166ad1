"""mp3_b200 -- B200-native batched MPEG-1/2 Layer III decoder.

Python host-side mirror of the C-ABI in include/mp3b.h (ctypes).  The reference project
(lxm0851/mp3) is a scripting-language audio player with no published source
(/root/reference/README.md:2,44), so this is the binding such a player would use: open a context
on a GPU, hand it MP3 byte strings, get PCM back.

There is deliberately no CPU fallback: if libmp3b.so is missing, or no CUDA device is usable,
the calls raise.
"""
import ctypes
import os

import numpy as np

from ._build import LIB, build_lib  # noqa: F401

PCM_S16, PCM_F32 = 0, 1
HOST, DEVICE = 0, 1
INDEX_DEVICE, INDEX_HOST = 0, 1
PIPE_FUSED, PIPE_STAGED = 0, 1
STAGE_FRAMES, STAGE_UNITDESC, STAGE_IS, STAGE_SF, STAGE_XR, STAGE_SB, STAGE_MAINDATA = range(7)


class Mp3bError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("mp3b: %s (status %d)" % (msg, status))
        self.status = status


class Opts(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("pcm_format", ctypes.c_int32), ("indexer", ctypes.c_int32),
                ("pipeline", ctypes.c_int32), ("host_threads", ctypes.c_int32), ("keep_stages", ctypes.c_int32),
                ("async_index", ctypes.c_int32), ("gapless", ctypes.c_int32),
                ("verify_crc", ctypes.c_int32)]


class TagInfo(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("has_lame", ctypes.c_int32), ("frames", ctypes.c_uint32),
                ("bytes", ctypes.c_uint32), ("enc_delay", ctypes.c_int32), ("enc_padding", ctypes.c_int32),
                ("first_sample", ctypes.c_int64), ("num_samples", ctypes.c_int64)]


class StreamInfo(ctypes.Structure):
    _fields_ = [("sample_rate", ctypes.c_int32), ("channels", ctypes.c_int32), ("lsf", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("frames", ctypes.c_int64), ("samples", ctypes.c_int64),
                ("concealed_frames", ctypes.c_int64), ("pcm_offset", ctypes.c_int64),
                ("total_samples", ctypes.c_int64)]


class Stats(ctypes.Structure):
    _fields_ = [("streams", ctypes.c_int64), ("frames", ctypes.c_int64), ("granules", ctypes.c_int64),
                ("units", ctypes.c_int64), ("bytes_in", ctypes.c_int64), ("pcm_bytes", ctypes.c_int64),
                ("concealed_frames", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("ms_index", ctypes.c_float), ("ms_huffman", ctypes.c_float), ("ms_requant", ctypes.c_float),
                ("ms_imdct", ctypes.c_float), ("ms_overlap", ctypes.c_float), ("ms_synth", ctypes.c_float),
                ("ms_fused", ctypes.c_float), ("ms_total", ctypes.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "mp3b_abi_version", "mp3b_device_count", "mp3b_opts_default", "mp3b_ctx_create", "mp3b_ctx_destroy",
    "mp3b_ctx_set_stream", "mp3b_ctx_set_stage_timing", "mp3b_resample_tc_plan", "mp3b_strerror", "mp3b_last_error", "mp3b_host_alloc", "mp3b_host_free", "mp3b_decode_batch",
    "mp3b_decode_packed", "mp3b_sync", "mp3b_flush", "mp3b_batch_stream_info", "mp3b_batch_tag_info", "mp3b_batch_pcm_device_ptr",
    "mp3b_batch_fetch_pcm", "mp3b_get_stats", "mp3b_set_pcm_sink", "mp3b_stream_open", "mp3b_stream_close", "mp3b_stream_enqueue",
    "mp3b_decode", "mp3b_stream_get_info", "mp3b_stream_fetch_pcm", "mp3b_stream_pcm_device_ptr",
    "mp3b_debug_stage", "mp3b_index_stream_host", "mp3b_batch_resample", "mp3b_batch_resampled_info",
    "mp3b_batch_resampled_device_ptr", "mp3b_batch_fetch_resampled", "mp3b_resample_filter",
    "mp3b_batch_time_stretch", "mp3b_batch_stretched_info", "mp3b_batch_stretched_device_ptr",
    "mp3b_batch_fetch_stretched", "mp3b_batch_stretch_offsets", "mp3b_batch_planar",
    "mp3b_batch_planar_device_ptr", "mp3b_batch_fetch_planar", "mp3b_seek_plan",
    "mp3b_batch_segments", "mp3b_batch_fetch_segments", "mp3b_batch_window_energy",
]

_lib = None


def load_library():
    """Load libmp3b.so (never builds implicitly; call build_lib() / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    # MP3B_LIB: load another build of the library (kernel-tuning variants, tools/variant_bench.py)
    LIB = os.environ.get("MP3B_LIB") or globals()["LIB"]
    if not os.path.exists(LIB):
        raise Mp3bError(-5, "libmp3b.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback")
    L = ctypes.CDLL(LIB)
    vp, i32, u64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t
    L.mp3b_strerror.restype = ctypes.c_char_p
    L.mp3b_last_error.restype = ctypes.c_char_p
    L.mp3b_last_error.argtypes = [vp]
    L.mp3b_host_alloc.restype = vp
    L.mp3b_host_alloc.argtypes = [sz]
    L.mp3b_host_free.argtypes = [vp]
    L.mp3b_ctx_create.argtypes = [i32, ctypes.POINTER(Opts), ctypes.POINTER(vp)]
    L.mp3b_ctx_destroy.argtypes = [vp]
    L.mp3b_ctx_set_stream.argtypes = [vp, vp]
    L.mp3b_ctx_set_stage_timing.argtypes = [vp, ctypes.c_int]
    L.mp3b_opts_default.argtypes = [ctypes.POINTER(Opts)]
    L.mp3b_decode_batch.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(sz), i32]
    L.mp3b_decode_packed.argtypes = [vp, vp, ctypes.POINTER(u64), i32, i32]
    L.mp3b_sync.argtypes = [vp]
    L.mp3b_flush.argtypes = [vp]
    L.mp3b_batch_stream_info.argtypes = [vp, i32, ctypes.POINTER(StreamInfo)]
    L.mp3b_batch_tag_info.argtypes = [vp, i32, ctypes.POINTER(TagInfo)]
    L.mp3b_batch_pcm_device_ptr.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u64)]
    L.mp3b_batch_fetch_pcm.argtypes = [vp, vp, u64, i32, ctypes.POINTER(u64)]
    L.mp3b_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.mp3b_set_pcm_sink.argtypes = [vp, vp, u64]
    L.mp3b_stream_open.argtypes = [vp, ctypes.POINTER(vp)]
    L.mp3b_stream_close.argtypes = [vp]
    L.mp3b_stream_enqueue.argtypes = [vp, vp, sz]
    L.mp3b_decode.argtypes = [vp]
    L.mp3b_stream_get_info.argtypes = [vp, ctypes.POINTER(StreamInfo)]
    L.mp3b_stream_fetch_pcm.argtypes = [vp, vp, sz, i32, ctypes.POINTER(sz)]
    L.mp3b_stream_pcm_device_ptr.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(sz)]
    L.mp3b_debug_stage.argtypes = [vp, i32, vp, u64, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(u64)]
    _lib = L
    return L


class PinnedBuffer:
    """Page-locked host memory from mp3b_host_alloc, viewed as a numpy array."""

    def __init__(self, nbytes):
        self._L = load_library()
        self.nbytes = int(nbytes)
        self.ptr = self._L.mp3b_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise Mp3bError(-6, "pinned allocation of %d bytes failed" % nbytes)
        self.array = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ctypes.c_uint8)), (max(self.nbytes, 1),))

    def view(self, dtype, count=None):
        a = self.array[: self.nbytes].view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self.ptr:
            self._L.mp3b_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Stream:
    """open / enqueue / decode / fetch interface of one MP3 stream (mp3b_stream_*)."""

    def __init__(self, dec):
        self.dec = dec
        h = ctypes.c_void_p()
        dec._ck(dec.L.mp3b_stream_open(dec.ctx, ctypes.byref(h)))
        self.h = h

    def enqueue(self, data):
        data = bytes(data)
        buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
        self.dec._ck(self.dec.L.mp3b_stream_enqueue(self.h, buf, len(data)))

    def info(self):
        inf = StreamInfo()
        rc = self.dec.L.mp3b_stream_get_info(self.h, ctypes.byref(inf))
        if rc not in (0, -2):  # -2: no frame seen yet
            self.dec._ck(rc)
        return inf

    def fetch(self, max_samples):
        inf = self.info()
        dt = np.int16 if self.dec.pcm_format == PCM_S16 else np.float32
        out = np.zeros((max_samples, max(inf.channels, 1)), dt)
        got = ctypes.c_size_t()
        self.dec._ck(self.dec.L.mp3b_stream_fetch_pcm(self.h, out.ctypes.data_as(ctypes.c_void_p), max_samples, HOST,
                                                     ctypes.byref(got)))
        return out[: got.value]

    def close(self):
        if self.h:
            self.dec.L.mp3b_stream_close(self.h)
            self.h = None


class FrameRec(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint32), ("payload_offset", ctypes.c_uint32), ("header", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


def index_stream_host(data):
    """Frame index of one stream computed on the host (no GPU): (frames[n] structured array, StreamInfo,
    TagInfo), or (empty, info, tag) when no MPEG audio frame is found."""
    L = load_library()
    data = bytes(data)
    n = len(data)
    buf = (ctypes.c_uint8 * max(n, 1)).from_buffer_copy(data if n else b"\0")
    cap = n // 24 + 2
    frames = (FrameRec * cap)()
    got = ctypes.c_size_t()
    info, tag = StreamInfo(), TagInfo()
    L.mp3b_index_stream_host.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(StreamInfo),
                                         ctypes.POINTER(TagInfo)]
    rc = L.mp3b_index_stream_host(buf, n, frames, cap, ctypes.byref(got), ctypes.byref(info), ctypes.byref(tag))
    if rc not in (0, -2):
        raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
    arr = np.frombuffer(frames, dtype=np.dtype([("offset", "<u4"), ("payload_offset", "<u4"), ("header", "<u4"),
                                                ("reserved", "<u4")]), count=got.value).copy()
    return arr, info, tag


class Seek(ctypes.Structure):
    _fields_ = [("byte_offset", ctypes.c_uint64), ("first_frame", ctypes.c_uint32), ("target_frame", ctypes.c_uint32),
                ("discard_samples", ctypes.c_int64)]


def seek_plan(data, target_sample, frames=None):
    """Where to start decoding `data` so that the PCM from target_sample on equals a decode from the start
    (include/mp3b.h: mp3b_seek_plan).  Returns a Seek; decode data[seek.byte_offset:] and drop the first
    seek.discard_samples samples per channel."""
    L = load_library()
    data = bytes(data)
    if frames is None:
        frames, _, _ = index_stream_host(data)
    frames = np.ascontiguousarray(frames)
    buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
    out = Seek()
    L.mp3b_seek_plan.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int64,
                                 ctypes.POINTER(Seek)]
    rc = L.mp3b_seek_plan(buf, len(data), frames.ctypes.data_as(ctypes.c_void_p), len(frames), int(target_sample),
                          ctypes.byref(out))
    if rc != 0:
        raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
    return out


def resample_tc_plan(in_rate, out_rate):
    """The tensor-core resampler's plan: list of per-kind coefficient matrices [128, K] (float32), or None if the rate
    pair is not served by that path."""
    L = load_library()
    L.mp3b_resample_tc_plan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    nk, kp = ctypes.c_int(), ctypes.c_int()
    rc = L.mp3b_resample_tc_plan(in_rate, out_rate, 0, None, 0, ctypes.byref(nk), ctypes.byref(kp))
    if rc == -4:  # MP3B_E_UNSUPPORTED
        return None
    if rc != 0:
        raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
    out = []
    for k in range(nk.value):
        a = np.empty((128, kp.value), np.float32)
        rc = L.mp3b_resample_tc_plan(in_rate, out_rate, k, a.ctypes.data_as(ctypes.c_void_p), a.size, ctypes.byref(nk),
                                     ctypes.byref(kp))
        if rc != 0:
            raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
        out.append(a)
    return out


def resample_filter(in_rate, out_rate):
    """The FIR the library uses for in_rate -> out_rate: (taps[L, taps_per_phase] float32, L, M)."""
    L = load_library()
    L.mp3b_resample_filter.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                       ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    n, l, m, t = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = L.mp3b_resample_filter(in_rate, out_rate, None, 0, ctypes.byref(n), ctypes.byref(l), ctypes.byref(m),
                                ctypes.byref(t))
    if rc not in (0, -3):
        raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
    taps = np.empty(n.value, np.float32)
    rc = L.mp3b_resample_filter(in_rate, out_rate, taps.ctypes.data_as(ctypes.c_void_p), taps.size, ctypes.byref(n),
                                ctypes.byref(l), ctypes.byref(m), ctypes.byref(t))
    if rc != 0:
        raise Mp3bError(rc, L.mp3b_strerror(rc).decode())
    return taps.reshape(l.value, t.value), l.value, m.value


class Decoder:
    """One context on one GPU (mp3b_ctx).  Not thread-safe; use one per GPU."""

    def __init__(self, device=0, pcm_format=PCM_S16, indexer=INDEX_DEVICE, pipeline=None, host_threads=0,
                 keep_stages=False, async_index=None, gapless=False, verify_crc=False):
        self.L = load_library()
        o = Opts()
        self.L.mp3b_opts_default(ctypes.byref(o))
        o.pcm_format, o.indexer, o.host_threads, o.keep_stages = pcm_format, indexer, host_threads, int(keep_stages)
        if pipeline is not None:
            o.pipeline = pipeline
        if async_index is not None:
            o.async_index = int(bool(async_index))
        o.gapless = int(bool(gapless))
        o.verify_crc = int(bool(verify_crc))
        self.pcm_format = pcm_format
        ctx = ctypes.c_void_p()
        rc = self.L.mp3b_ctx_create(device, ctypes.byref(o), ctypes.byref(ctx))
        if rc != 0:
            raise Mp3bError(rc, self.L.mp3b_strerror(rc).decode() + " -- no CPU fallback exists")
        self.ctx = ctx
        self.nstreams = 0

    def _ck(self, rc):
        if rc != 0:
            extra = self.L.mp3b_last_error(self.ctx).decode() if self.ctx else ""
            raise Mp3bError(rc, self.L.mp3b_strerror(rc).decode() + (": " + extra if extra else ""))

    def close(self):
        if self.ctx:
            self.L.mp3b_ctx_destroy(self.ctx)
            self.ctx = None

    def set_stream(self, cuda_stream_handle):
        """Enqueue on the caller's CUDA stream (an integer cudaStream_t, e.g. torch's .cuda_stream)."""
        self._ck(self.L.mp3b_ctx_set_stream(self.ctx, ctypes.c_void_p(cuda_stream_handle)))

    def set_stage_timing(self, on):
        """Per-stage CUDA events between the kernels (stats().ms_huffman ...); off by default: the kernels then
        overlap through programmatic dependent launch."""
        self._ck(self.L.mp3b_ctx_set_stage_timing(self.ctx, 1 if on else 0))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- bulk interface
    def decode_batch(self, streams, sync=True):
        """streams: list of bytes-like.  Gathers on the host (mp3b_decode_batch)."""
        n = len(streams)
        keep = [np.frombuffer(bytes(s) if not isinstance(s, (bytes, bytearray, np.ndarray)) else s, np.uint8)
                for s in streams]
        ptrs = (ctypes.c_void_p * max(n, 1))(*[k.ctypes.data if k.size else None for k in keep])
        lens = (ctypes.c_size_t * max(n, 1))(*[k.size for k in keep])
        self._ck(self.L.mp3b_decode_batch(self.ctx, ptrs, lens, n))
        self.nstreams = n
        if sync:
            self.sync()

    def decode_packed(self, base_ptr, offsets, where=HOST, sync=True):
        """base_ptr: integer address (host pinned / pageable, or device); offsets: uint64[n+1]."""
        offsets = np.ascontiguousarray(offsets, np.uint64)
        n = offsets.size - 1
        self._ck(self.L.mp3b_decode_packed(self.ctx, ctypes.c_void_p(base_ptr),
                                           offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), n, where))
        self.nstreams = n
        if sync:
            self.sync()

    def sync(self):
        self._ck(self.L.mp3b_sync(self.ctx))

    def flush(self):
        """Order the outstanding sink copies into the context's stream (non-blocking)."""
        self._ck(self.L.mp3b_flush(self.ctx))

    def stream_info(self, i):
        inf = StreamInfo()
        rc = self.L.mp3b_batch_stream_info(self.ctx, i, ctypes.byref(inf))
        if rc not in (0, -2):
            self._ck(rc)
        return inf

    def tag_info(self, i):
        """Xing / Info / LAME / VBRI tag of stream i and the gapless window it implies."""
        t = TagInfo()
        self._ck(self.L.mp3b_batch_tag_info(self.ctx, i, ctypes.byref(t)))
        return t

    def pcm_device(self):
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(self.L.mp3b_batch_pcm_device_ptr(self.ctx, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def fetch_pcm(self, out=None):
        """Whole-batch PCM arena as a flat numpy array (int16 or float32)."""
        _, n = self.pcm_device()
        dt = np.int16 if self.pcm_format == PCM_S16 else np.float32
        if out is None:
            out = np.empty(n, dt)
        got = ctypes.c_uint64()
        self._ck(self.L.mp3b_batch_fetch_pcm(self.ctx, out.ctypes.data_as(ctypes.c_void_p), out.size, HOST,
                                             ctypes.byref(got)))
        self.sync()
        return out[: got.value]

    def fetch_pcm_into(self, host_ptr, cap_elems):
        got = ctypes.c_uint64()
        self._ck(self.L.mp3b_batch_fetch_pcm(self.ctx, ctypes.c_void_p(host_ptr), cap_elems, HOST, ctypes.byref(got)))
        return got.value

    def set_pcm_sink(self, host_ptr, cap_elems):
        """Stream every decode's PCM to this (pinned) host buffer, overlapped with the kernels."""
        self._ck(self.L.mp3b_set_pcm_sink(self.ctx, ctypes.c_void_p(host_ptr) if host_ptr else None, cap_elems))

    def stream_pcm(self, i, arena=None):
        """PCM of stream i as [samples, channels]."""
        inf = self.stream_info(i)
        if arena is None:
            arena = self.fetch_pcm()
        n = inf.samples * inf.channels
        return arena[inf.pcm_offset: inf.pcm_offset + n].reshape(inf.samples, max(inf.channels, 1))

    def resample(self, out_rate):
        """Convert every stream of the last batch to out_rate (async); see fetch_resampled()."""
        self._ck(self.L.mp3b_batch_resample(self.ctx, int(out_rate)))

    def fetch_resampled(self):
        """(flat arena, [(offset_elems, samples)] per stream) of the resampled batch."""
        L = self.L
        L.mp3b_batch_resampled_device_ptr.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                                      ctypes.POINTER(ctypes.c_uint64)]
        L.mp3b_batch_fetch_resampled.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_uint64)]
        L.mp3b_batch_resampled_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                                ctypes.POINTER(ctypes.c_int64)]
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(L.mp3b_batch_resampled_device_ptr(self.ctx, ctypes.byref(p), ctypes.byref(n)))
        out = np.empty(n.value, np.int16 if self.pcm_format == PCM_S16 else np.float32)
        got = ctypes.c_uint64()
        self._ck(L.mp3b_batch_fetch_resampled(self.ctx, out.ctypes.data_as(ctypes.c_void_p), out.size, HOST,
                                              ctypes.byref(got)))
        self.sync()
        where = []
        for i in range(self.nstreams):
            o, c = ctypes.c_int64(), ctypes.c_int64()
            self._ck(L.mp3b_batch_resampled_info(self.ctx, i, ctypes.byref(o), ctypes.byref(c)))
            where.append((o.value, c.value))
        return out, where

    def planar(self):
        """Planar copy of the last batch: returns the flat arena in which stream i, channel c starts at
        pcm_offset + c * samples."""
        L = self.L
        L.mp3b_batch_planar.argtypes = [ctypes.c_void_p]
        L.mp3b_batch_fetch_planar.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_uint64)]
        self._ck(L.mp3b_batch_planar(self.ctx))
        _, n = self.pcm_device()
        out = np.zeros(n, np.int16 if self.pcm_format == PCM_S16 else np.float32)
        got = ctypes.c_uint64()
        self._ck(L.mp3b_batch_fetch_planar(self.ctx, out.ctypes.data_as(ctypes.c_void_p), out.size, HOST,
                                           ctypes.byref(got)))
        self.sync()
        return out

    def segments(self, threshold=328, min_silence_ms=300, min_sentence_ms=200):
        """Sentence boundaries of the last batch (include/mp3b.h: mp3b_batch_segments): a list with one
        int64 array [n, 2] of {first sample, end sample} per stream."""
        L = self.L
        L.mp3b_batch_segments.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.mp3b_batch_fetch_segments.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                ctypes.POINTER(ctypes.c_size_t)]
        self._ck(L.mp3b_batch_segments(self.ctx, int(threshold), int(min_silence_ms), int(min_sentence_ms)))
        out = []
        n = ctypes.c_size_t()
        for i in range(self.nstreams):
            rc = L.mp3b_batch_fetch_segments(self.ctx, i, None, 0, ctypes.byref(n))
            if rc not in (0, -3):
                self._ck(rc)
            a = np.zeros((n.value, 2), np.int64)
            if n.value:
                self._ck(L.mp3b_batch_fetch_segments(self.ctx, i, a.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
            out.append(a)
        return out

    def window_energy(self, i):
        """(energies uint64[nwin], window length in samples) of stream i, after segments()."""
        L = self.L
        L.mp3b_batch_window_energy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                               ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)]
        n, w = ctypes.c_size_t(), ctypes.c_int()
        rc = L.mp3b_batch_window_energy(self.ctx, int(i), None, 0, ctypes.byref(n), ctypes.byref(w))
        if rc not in (0, -3):
            self._ck(rc)
        a = np.zeros(n.value, np.uint64)
        if n.value:
            self._ck(L.mp3b_batch_window_energy(self.ctx, int(i), a.ctypes.data_as(ctypes.c_void_p), n.value,
                                                ctypes.byref(n), ctypes.byref(w)))
        return a, w.value

    def time_stretch(self, num, den):
        """WSOLA time-scale modification of the last batch: speed = num / den (1, 2 = half speed); async."""
        self._ck(self.L.mp3b_batch_time_stretch(self.ctx, int(num), int(den)))

    def fetch_stretched(self):
        """(flat arena, [(offset_elems, samples)] per stream) of the stretched batch."""
        L = self.L
        L.mp3b_batch_stretched_device_ptr.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                                      ctypes.POINTER(ctypes.c_uint64)]
        L.mp3b_batch_fetch_stretched.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_uint64)]
        L.mp3b_batch_stretched_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                                ctypes.POINTER(ctypes.c_int64)]
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(L.mp3b_batch_stretched_device_ptr(self.ctx, ctypes.byref(p), ctypes.byref(n)))
        out = np.empty(n.value, np.int16 if self.pcm_format == PCM_S16 else np.float32)
        got = ctypes.c_uint64()
        self._ck(L.mp3b_batch_fetch_stretched(self.ctx, out.ctypes.data_as(ctypes.c_void_p), out.size, HOST,
                                              ctypes.byref(got)))
        self.sync()
        where = []
        for i in range(self.nstreams):
            o, c = ctypes.c_int64(), ctypes.c_int64()
            self._ck(L.mp3b_batch_stretched_info(self.ctx, i, ctypes.byref(o), ctypes.byref(c)))
            where.append((o.value, c.value))
        return out, where

    def stretch_offsets(self, i):
        """(alignment offsets chosen per output segment of stream i, hop)."""
        L = self.L
        L.mp3b_batch_stretch_offsets.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int)]
        n, hop = ctypes.c_size_t(), ctypes.c_int()
        rc = L.mp3b_batch_stretch_offsets(self.ctx, i, None, 0, ctypes.byref(n), ctypes.byref(hop))
        if rc not in (0, -3):
            self._ck(rc)
        out = np.zeros(n.value, np.int32)
        if n.value:
            self._ck(L.mp3b_batch_stretch_offsets(self.ctx, i, out.ctypes.data_as(ctypes.c_void_p), out.size,
                                                  ctypes.byref(n), ctypes.byref(hop)))
        return out, hop.value

    def stats(self):
        st = Stats()
        self._ck(self.L.mp3b_get_stats(self.ctx, ctypes.byref(st)))
        return st

    def stage(self, which):
        es, cnt = ctypes.c_uint32(), ctypes.c_uint64()
        self._ck(self.L.mp3b_debug_stage(self.ctx, which, None, 0, ctypes.byref(es), ctypes.byref(cnt)))
        dt = {STAGE_IS: np.int16, STAGE_SF: np.uint8, STAGE_XR: np.float32, STAGE_SB: np.float32,
              STAGE_MAINDATA: np.uint8, STAGE_FRAMES: np.uint32, STAGE_UNITDESC: np.uint8}[which]
        nbytes = es.value * cnt.value
        out = np.empty(nbytes // np.dtype(dt).itemsize, dt)
        self._ck(self.L.mp3b_debug_stage(self.ctx, which, out.ctypes.data_as(ctypes.c_void_p), nbytes, None, None))
        if which in (STAGE_IS, STAGE_XR):
            return out.reshape(-1, 576)
        if which == STAGE_SF:
            return out.reshape(-1, 40)
        if which == STAGE_SB:
            return out.reshape(-1, 18, 32)
        if which == STAGE_FRAMES:
            return out.reshape(-1, 4)
        if which == STAGE_UNITDESC:
            return out.reshape(-1, 32)
        return out

    # ---- stream interface
    def open_stream(self):
        return Stream(self)

    def decode_streams(self):
        self._ck(self.L.mp3b_decode(self.ctx))
        self.sync()


def pack_streams(streams):
    """Concatenate byte strings; returns (uint8 array, uint64 offsets[n+1])."""
    lens = np.array([len(s) for s in streams], np.uint64)
    offs = np.zeros(len(streams) + 1, np.uint64)
    np.cumsum(lens, out=offs[1:])
    buf = np.empty(int(offs[-1]), np.uint8)
    for s, o in zip(streams, offs[:-1]):
        buf[int(o): int(o) + len(s)] = np.frombuffer(s, np.uint8)
    return buf, offs
