"""Small pure-Python helpers for tests: frame splitting and the ISO 11172-4 accuracy criterion."""
import numpy as np

_BR = [[0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0],
       [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160, 0]]
_BR2 = [[0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 0], _BR[1]]  # Layer II
_BR1 = [[0, 32, 64, 96, 128, 160, 192, 224, 256, 288, 320, 352, 384, 416, 448, 0],
        [0, 32, 48, 56, 64, 80, 96, 112, 128, 144, 160, 176, 192, 224, 256, 0]]                        # Layer I
_SR = {3: [44100, 48000, 32000], 2: [22050, 24000, 16000], 0: [11025, 12000, 8000]}  # by version bits


def frame_len(h):
    """Length in bytes of the Layer III frame whose 4 header bytes are h, or 0 if invalid."""
    if h[0] != 0xFF or (h[1] & 0xE0) != 0xE0:
        return 0
    ver = (h[1] >> 3) & 3
    layer = (h[1] >> 1) & 3  # 1 = Layer III, 2 = Layer II, 3 = Layer I
    if ver == 1 or layer == 0:
        return 0
    lsf = 0 if ver == 3 else 1  # MPEG-2 and MPEG-2.5 share the LSF syntax
    bri, sri = h[2] >> 4, (h[2] >> 2) & 3
    if bri in (0, 15) or sri == 3:
        return 0
    pad = (h[2] >> 1) & 1
    if layer == 3:
        return (12 * _BR1[lsf][bri] * 1000 // _SR[ver][sri] + pad) * 4
    if layer == 2:
        return 144 * _BR2[lsf][bri] * 1000 // _SR[ver][sri] + pad
    return (72 if lsf else 144) * _BR[lsf][bri] * 1000 // _SR[ver][sri] + pad


def split_frames(data):
    out, p = [], 0
    while p + 4 <= len(data):
        n = frame_len(data[p:p + 4])
        if n == 0 or p + n > len(data):
            p += 1
            continue
        out.append(data[p:p + n])
        p += n
    return out


RMS_LIMIT = 2.0 ** -15 / np.sqrt(12.0)   # ISO/IEC 11172-4 full accuracy: rms of the difference
MAX_LIMIT = 2.0 ** -14                    # and maximum absolute difference, full scale = +-1.0


def iso_compliance(test, ref):
    """Return (rms_error, max_abs_error) between two PCM arrays in full-scale +-1 units."""
    d = np.asarray(test, np.float64) - np.asarray(ref, np.float64)
    if d.size == 0:
        return 0.0, 0.0
    return float(np.sqrt(np.mean(d * d))), float(np.max(np.abs(d)))


def assert_iso_full_accuracy(test, ref, what=""):
    rms, mx = iso_compliance(test, ref)
    assert rms < RMS_LIMIT and mx <= MAX_LIMIT, "%s: rms %.3g (limit %.3g) max %.3g (limit %.3g)" % (
        what, rms, RMS_LIMIT, mx, MAX_LIMIT)
    return rms, mx


def patch_side_info(frame, unit, part2_3_length=None, big_values=None):
    """Return a copy of a Layer III frame (no CRC fix-up) with fields of granule-channel `unit` (in side-info
    order: [granule][channel]) overwritten: 11172-3 2.4.1.7 / 13818-3 2.4.1.7 bit layout."""
    b = bytearray(frame)
    lsf = ((b[1] >> 3) & 3) != 3
    mono = (b[3] >> 6) == 3
    crc = not (b[1] & 1)
    base = (4 + (2 if crc else 0)) * 8
    if lsf:
        base += 9 if mono else 10
        stride = 63
    else:
        base += 18 if mono else 20
        stride = 59

    def put(pos, n, v):
        for i in range(n):
            bit = (v >> (n - 1 - i)) & 1
            byte, sh = (pos + i) >> 3, 7 - ((pos + i) & 7)
            b[byte] = (b[byte] & ~(1 << sh)) | (bit << sh)
    if part2_3_length is not None:
        put(base + unit * stride, 12, part2_3_length)
    if big_values is not None:
        put(base + unit * stride + 12, 9, big_values)
    return bytes(b)
