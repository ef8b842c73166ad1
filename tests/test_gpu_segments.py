"""GPU: sentence boundaries (mp3b_batch_segments) against the numpy restatement of the same definition
(oracle/segments.py): window energies and segment lists bit-exact, s16 and f32, gapless on and off, all
layers, plus the properties of the definition on speech-like streams spliced from loud and quiet frames."""
import numpy as np
import pytest

import l3util

pytestmark = pytest.mark.gpu


def speech_like(synth_mod, pattern, seed, **kw):
    """Splice self-contained frames (no bit reservoir) of a loud and a near-silent stream: pattern is a list
    of (n_frames, loud) runs."""
    n = sum(p[0] for p in pattern)
    loud = l3util.split_frames(synth_mod.make_stream(nframes=n, seed=seed, reservoir=0, level_lo_db=8, level_hi_db=24, **kw))
    quiet = l3util.split_frames(synth_mod.make_stream(nframes=n, seed=seed + 1, reservoir=0, level_lo_db=84, level_hi_db=90, **kw))
    out, k = [], 0
    for cnt, is_loud in pattern:
        out += (loud if is_loud else quiet)[k: k + cnt]
        k += cnt
    return b"".join(out)


PATTERN = [(6, 0), (30, 1), (20, 0), (25, 1), (5, 0), (25, 1), (40, 0), (3, 1), (30, 0), (12, 1)]


@pytest.fixture(scope="module")
def streams(synth_mod):
    return [speech_like(synth_mod, PATTERN, 100),
            speech_like(synth_mod, PATTERN, 102, mode=3),
            speech_like(synth_mod, PATTERN, 104, sample_rate=22050, bitrate_kbps=64, mode=1),
            speech_like(synth_mod, PATTERN, 106, sample_rate=8000, bitrate_kbps=16, mode=3),
            speech_like(synth_mod, PATTERN, 108, layer=2, bitrate_kbps=192),
            speech_like(synth_mod, [(3 * a, b) for a, b in PATTERN], 110, layer=1, bitrate_kbps=256),
            synth_mod.make_stream(nframes=30, seed=112),                                       # no pauses at all
            synth_mod.make_stream(nframes=30, seed=113, level_lo_db=86, level_hi_db=90),       # nothing but silence
            b"\x00" * 500,                                                                      # no stream
            synth_mod.make_stream(nframes=20, seed=114, tag=1, tag_lame=1, enc_delay=576, enc_padding=900)]


@pytest.mark.parametrize("gapless", [False, True])
@pytest.mark.parametrize("fmt", ["s16", "f32"])
def test_segments_and_energies_bit_exact(fmt, gapless, streams):
    import mp3_b200 as m
    from oracle import segments as seg_oracle
    with m.Decoder(device=0, pcm_format=m.PCM_S16 if fmt == "s16" else m.PCM_F32, gapless=gapless) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        for params in (dict(), dict(threshold=100, min_silence_ms=100, min_sentence_ms=0),
                       dict(threshold=2000, min_silence_ms=500, min_sentence_ms=400)):
            got = dec.segments(**params)
            assert len(got) == len(streams)
            for i in range(len(streams)):
                inf = dec.stream_info(i)
                if not inf.frames or inf.samples <= 0:
                    assert got[i].shape == (0, 2)
                    continue
                pcm = dec.stream_pcm(i, arena)
                E, W = seg_oracle.window_energy(pcm, inf.sample_rate)
                gE, gW = dec.window_energy(i)
                assert gW == W and np.array_equal(gE, E), i
                want = seg_oracle.segments(pcm, inf.sample_rate, **params)
                assert np.array_equal(got[i], want), (i, params, got[i], want)


def test_segments_follow_the_splice_pattern(streams):
    """The default parameters find the sentences the pattern holds: pauses of 20 / 40 / 30 frames split, the
    5-frame gap (130 ms) does not, the 3-frame burst (78 ms) is dropped as shorter than 200 ms."""
    import mp3_b200 as m
    with m.Decoder(device=0) as dec:
        dec.decode_batch(streams[:2])
        for i, segs in enumerate(dec.segments()):
            spf = 1152
            want = [(6, 36), (56, 111), (184, 196)]  # frames; the decoder's 528 + 1 samples of delay shift everything a little
            assert len(segs) == len(want), segs
            for (a, b), (fa, fb) in zip(segs, want):
                assert abs(a - fa * spf) <= 1152 and abs(b - fb * spf) <= 1152, (i, segs)
            assert np.all(segs[1:, 0] > segs[:-1, 1])
        # silence and continuous sound
        dec.decode_batch(streams[6:8])
        segs = dec.segments()
        assert len(segs[0]) == 1 and segs[0][0, 0] <= 1152 and segs[0][0, 1] >= 29 * 1152
        assert len(segs[1]) == 0
