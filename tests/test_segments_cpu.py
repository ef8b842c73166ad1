"""The sentence-boundary oracle (oracle/segments.py) on constructed signals: its definition is what the CUDA
path is held to (tests/test_gpu_segments.py), so its own behaviour is pinned here on the CPU."""
import numpy as np

from oracle import segments as sg


def tone(n, amp, sr=44100):
    t = np.arange(n) / sr
    return np.rint(amp * np.sin(2 * np.pi * 440 * t)).astype(np.int16)


def test_pauses_split_and_short_gaps_do_not():
    sr, W = 44100, 441
    parts = [(50, 0), (100, 8000), (40, 0), (100, 8000), (10, 0), (60, 8000), (35, 10), (5, 8000), (80, 0)]  # windows, amplitude
    x = np.concatenate([tone(w * W, a) for w, a in parts])
    pcm = np.stack([x, x], axis=1)
    E, w = sg.window_energy(pcm, sr)
    assert w == W and len(E) == sum(p[0] for p in parts)
    assert E[0] == 0 and E[60] == np.sum(x[60 * W: 61 * W].astype(np.int64) ** 2)
    segs = sg.segments(pcm, sr, threshold=328, min_silence_ms=300, min_sentence_ms=200)
    # 400-ms pause splits, 100-ms gap does not, 50-ms burst after a 350-ms pause is dropped
    assert segs.tolist() == [[50 * W, 150 * W], [190 * W, 360 * W]]
    # a shorter minimum keeps the burst; a longer pause requirement merges everything up to the burst
    assert len(sg.segments(pcm, sr, 328, 300, 0)) == 3
    assert sg.segments(pcm, sr, 328, 410, 200).tolist() == [[50 * W, 400 * W]]


def test_mono_float_ragged_tail_and_rounding():
    sr = 8000
    x = np.concatenate([np.zeros(405), tone(1000, 3000, sr).astype(np.float64), np.zeros(80), [0.75, -0.75, 0.5]])
    pcm = (x / 32768.0).astype(np.float32)[:, None]
    E, W = sg.window_energy(pcm, sr)
    assert W == 80 and len(E) == (len(x) + 79) // 80
    assert E[-1] == 1 + 1 + 0  # 0.75 -> 1 and -0.75 -> -1 (round half to even: 0.5 -> 0), 3-sample last window
    segs = sg.segments(pcm, sr, threshold=100, min_silence_ms=10, min_sentence_ms=0)
    assert segs[0].tolist() == [400, 1440]
    assert sg.segments(np.zeros((0, 2), np.int16), 44100).shape == (0, 2)
