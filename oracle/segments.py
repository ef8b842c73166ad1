"""Oracle for mp3b_batch_segments (test infrastructure only): the definition in include/mp3b.h /
mp3_b200/csrc/k_segments.cu restated with numpy and a plain loop.  The reference has no code for sentence
detection (/root/reference/README.md:46 only names the feature), so parity here is GPU == this restatement
of OUR definition, bit for bit (integer arithmetic throughout)."""
import numpy as np


def to_s16(pcm):
    """[samples, channels] s16 or float (full scale +-1) -> int64 s16 values (float: round half to even, saturate)."""
    pcm = np.asarray(pcm)
    if pcm.dtype == np.int16:
        return pcm.astype(np.int64)
    return np.clip(np.rint(pcm.astype(np.float32) * np.float32(32768.0)), -32768, 32767).astype(np.int64)


def window_energy(pcm, sample_rate):
    x = to_s16(pcm)
    m = (x[:, 0] + x[:, 1]) >> 1 if x.shape[1] == 2 else x[:, 0]
    W = sample_rate // 100
    nwin = (len(m) + W - 1) // W
    sq = np.zeros(nwin * W, np.int64)
    sq[: len(m)] = m * m
    return sq.reshape(nwin, W).sum(axis=1).astype(np.uint64), W


def segments(pcm, sample_rate, threshold=328, min_silence_ms=300, min_sentence_ms=200):
    E, W = window_energy(pcm, sample_rate)
    N = len(pcm)
    G, S = min_silence_ms // 10, max(1, min_sentence_ms // 10)
    out, in_seg, start, last, run = [], False, 0, 0, 0
    for k in range(len(E)):
        cnt = min(W, N - k * W)
        if int(E[k]) > threshold * threshold * cnt:
            if not in_seg:
                in_seg, start = True, k
            last, run = k + 1, 0
        elif in_seg:
            run += 1
            if run >= G:
                if last - start >= S:
                    out.append((start * W, min(last * W, N)))
                in_seg = False
    if in_seg and last - start >= S:
        out.append((start * W, min(last * W, N)))
    return np.array(out, np.int64).reshape(-1, 2)
