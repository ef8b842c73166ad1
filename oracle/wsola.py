"""CPU oracle of the WSOLA time-scale modification (mp3_b200/csrc/k_stretch.cu).

TEST INFRASTRUCTURE ONLY (tests/ imports it; nothing under mp3_b200/ does).  The reference repository
describes slow playback (/root/reference/README.md:46) but has no code for it, so this is a numpy
restatement of the textbook waveform-similarity overlap-add (Verhelst & Roelands 1993) with the
parameters the product fixes: hop Hs = 512 / 256 / 128 by sample rate, frame 2 Hs, search radius Hs / 2,
an int8 alignment signal, a coarse-to-fine search (every 4th offset on every 2nd sample, then +-3 in
full), first-maximum tie break, periodic-Hann cross-fade.  float64 throughout.
"""
import numpy as np


def hop_of(sample_rate):
    return 512 if sample_rate >= 32000 else (256 if sample_rate >= 16000 else 128)


def to_s16(x):
    """Full-scale float -> s16 the way the device does it (round to nearest even, saturate)."""
    return np.clip(np.rint(np.asarray(x, np.float64) * 32768.0), -32768, 32767).astype(np.int64)


def align_signal(s16):
    """s16 [n, channels] -> the int8-range alignment signal."""
    s16 = np.asarray(s16, np.int64)
    v = (s16[:, 0] + s16[:, 1]) >> 9 if s16.shape[1] == 2 else s16[:, 0] >> 8
    return np.clip(v, -127, 127)


def wsola(x, s16, sample_rate, num, den):
    """x: float64 [n, channels] full scale (what is cross-faded); s16: the same PCM as s16 integers (what
    is aligned).  Returns (y [floor(n den / num), channels], offsets per output segment)."""
    x = np.asarray(x, np.float64)
    n, nch = x.shape
    hs = hop_of(sample_rate)
    N, R = 2 * hs, hs // 2
    out_n = n * den // num
    nseg = -(-out_n // hs)
    c = align_signal(s16)

    def seg(arr, start, length):
        out = np.zeros((length,) + arr.shape[1:], arr.dtype)
        lo, hi = max(start, 0), min(start + length, n)
        if hi > lo:
            out[lo - start: hi - start] = arr[lo:hi]
        return out

    y = np.zeros((nseg * hs, nch))
    w = 0.5 - 0.5 * np.cos(np.pi * np.arange(hs) / hs)
    offs = np.zeros(nseg, np.int64)
    p_prev = 0
    for m in range(nseg):
        if m == 0:
            p = 0
            y[:hs] = seg(x, 0, hs)
        else:
            a = (m * hs * num) // den
            t = seg(c, p_prev + hs, N)
            region = seg(c, a - R, N + 2 * R)
            win = np.lib.stride_tricks.sliding_window_view(region, N)   # win[c] = region[c : c + N], d = c - R
            # coarse: every fourth candidate, every second sample; fine: +-3 around the winner, in full
            coarse = win[::4, ::2] @ t[::2]
            c0 = 4 * int(np.argmax(coarse))                              # argmax returns the first maximum
            lo, hi = max(c0 - 3, 0), min(c0 + 3, 2 * R)
            d = lo + int(np.argmax(win[lo: hi + 1] @ t)) - R
            offs[m] = d
            p = a + d
            y[m * hs: (m + 1) * hs] = (1.0 - w)[:, None] * seg(x, p_prev + hs, hs) + w[:, None] * seg(x, p, hs)
        p_prev = p
    return y[:out_n], offs
