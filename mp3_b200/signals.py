"""Deterministic "real" audio for the encoder tests: a music-like and a speech-like signal (numpy only)."""
import numpy as np


def music(sr, seconds, seed=1):
    """Chords of decaying harmonic notes over a bass line, a little noise: tonal, wide-band, transient onsets."""
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    t = np.arange(n) / sr
    x = np.zeros(n)
    onsets = np.arange(0.0, seconds, 0.25)
    for k, t0 in enumerate(onsets):
        f0 = 110.0 * 2 ** (rng.integers(0, 24) / 12.0)
        i0 = int(t0 * sr)
        tt = t[i0:] - t0
        env = np.exp(-tt * (3.0 + 4.0 * rng.random())) * (1.0 - np.exp(-tt * 400.0))
        note = sum(np.sin(2 * np.pi * f0 * h * tt + rng.random() * 6.28) / h ** 1.3 for h in range(1, 9) if f0 * h < sr * 0.45)
        x[i0:] += 0.25 * env * note
    x += 0.15 * np.sin(2 * np.pi * 55.0 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 0.5 * t))
    x += 0.004 * rng.standard_normal(n)
    return x / max(1.0, np.abs(x).max() / 0.9)


def speech(sr, seconds, seed=2):
    """Voiced segments (a glottal pulse train through three moving formants) between pauses and fricative noise:
    the kind of signal the reference's listening-practice player plays."""
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    x = np.zeros(n)
    pos = 0
    while pos < n:
        dur = int(sr * (0.08 + 0.25 * rng.random()))
        kind = rng.integers(0, 4)
        seg = np.zeros(dur)
        tt = np.arange(dur) / sr
        if kind <= 1:  # voiced
            f0 = 100.0 + 80.0 * rng.random() + 15.0 * np.sin(2 * np.pi * 3.0 * tt)
            ph = 2 * np.pi * np.cumsum(f0) / sr
            src = sum(np.sin(h * ph) / h for h in range(1, 30))
            forms = [(300 + 500 * rng.random(), 80), (900 + 1400 * rng.random(), 120), (2300 + 900 * rng.random(), 200)]
            f = np.fft.rfftfreq(dur, 1.0 / sr)
            H = sum(1.0 / (1.0 + ((f - fc) / bw) ** 2) for fc, bw in forms)
            seg = np.fft.irfft(np.fft.rfft(src) * H, dur)
            seg *= np.hanning(dur) ** 0.5
        elif kind == 2:  # fricative
            f = np.fft.rfftfreq(dur, 1.0 / sr)
            seg = np.fft.irfft(np.fft.rfft(rng.standard_normal(dur)) * (f > 2500) * (f < 9000), dur) * np.hanning(dur) * 0.4
        m = np.abs(seg).max()
        if m > 0:
            seg = seg / m * (0.3 + 0.5 * rng.random())
        x[pos: pos + dur] += seg[: max(0, min(dur, n - pos))]
        pos += dur + int(sr * 0.02 * rng.integers(0, 6))
    return x


def castanets(sr, seconds, seed=5):
    """Sharp noise bursts over a quiet tone: attacks that make an encoder switch to short windows."""
    rng = np.random.default_rng(seed)
    n = int(sr * seconds)
    x = 0.05 * np.sin(2 * np.pi * 440.0 * np.arange(n) / sr)
    for t0 in np.arange(0.15, seconds, 0.21):
        i0 = int(t0 * sr)
        m = min(2000, n - i0)
        x[i0: i0 + m] += 0.7 * rng.standard_normal(m) * np.exp(-np.arange(m) / 300.0)
    return x


def to_s16(x):
    return np.clip(np.round(np.asarray(x) * 32767.0), -32768, 32767).astype(np.int16)


def stereo(sr, seconds, seed=1):
    """music left-ish, speech right-ish, some of each in both"""
    m, s = music(sr, seconds, seed), speech(sr, seconds, seed + 1)
    return to_s16(np.stack([0.8 * m + 0.15 * s, 0.25 * m + 0.7 * s], axis=1) * 0.9)


CODEC_DELAY = 1057  # 528 + 1 (decoder) + 528 (encoder filterbank): where the input reappears in the decoded stream


def snr_db(decoded, pcm16, delay=CODEC_DELAY, skip=2304):
    """decoded: float [samples, ch] (full scale 1), pcm16: the encoder's input.  SNR over the common part."""
    ref = pcm16.astype(np.float64) / 32768.0
    if ref.ndim == 1:
        ref = ref[:, None]
    m = min(decoded.shape[0] - delay, ref.shape[0])
    err = decoded[delay: delay + m] - ref[:m]
    sl = slice(skip, m - skip)
    return 10.0 * np.log10(np.sum(ref[:m][sl] ** 2) / np.sum(err[sl] ** 2)), float(np.abs(err[sl]).max())
