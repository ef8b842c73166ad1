"""GPU: the flow INTEGRATION.md shows for the reference's one described feature ("slow-speed listening,
repeat each sentence", /root/reference/README.md:46): decode -> sentence boundaries -> time stretch ->
slices per sentence, and a seek to a sentence start that reproduces the same PCM."""
import numpy as np
import pytest

from test_gpu_segments import PATTERN, speech_like

pytestmark = pytest.mark.gpu


def test_slow_listening_flow(synth_mod):
    import mp3_b200 as m
    mp3_bytes = speech_like(synth_mod, PATTERN, 100)
    with m.Decoder(device=0) as dec:
        dec.decode_batch([mp3_bytes])
        pcm = dec.stream_pcm(0, dec.fetch_pcm()).copy()
        sentences = dec.segments(threshold=328, min_silence_ms=300, min_sentence_ms=200)[0]
        assert len(sentences) == 3
        dec.time_stretch(3, 4)
        slow, where = dec.fetch_stretched()
        off, n = where[0]
        assert n == pcm.shape[0] * 4 // 3
        slow = slow[off: off + n * pcm.shape[1]].reshape(n, pcm.shape[1])
        for first, end in sentences:
            a, b = slow[first * 4 // 3: end * 4 // 3], pcm[first:end]
            assert a.shape[0] == end * 4 // 3 - first * 4 // 3
            # the slowed sentence carries the sentence's energy (same loudness, 4/3 the duration)
            ra = np.sqrt(np.mean(a.astype(np.float64) ** 2))
            rb = np.sqrt(np.mean(b.astype(np.float64) ** 2))
            assert 0.7 * rb < ra < 1.3 * rb
        # the pauses stay pauses
        q0, q1 = int(sentences[0][1]) + 4410, int(sentences[1][0]) - 4410
        assert np.abs(slow[q0 * 4 // 3: q1 * 4 // 3].astype(np.int32)).max() < 400
        # seek to the third sentence: identical PCM from there on
        t = int(sentences[2][0])
        sk = m.seek_plan(mp3_bytes, t)
        dec.decode_batch([mp3_bytes[sk.byte_offset:]])
        part = dec.stream_pcm(0, dec.fetch_pcm())
        assert np.array_equal(part[sk.discard_samples:], pcm[t:])
