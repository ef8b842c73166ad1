"""mp3b_seek_plan (host logic, no GPU): decoding from the planned offset and dropping the planned number of
samples gives, from the target on, exactly what a decode from the stream's start gives.  Checked here with
the oracle as the decoder (the same double-precision operations in the same order, so equality is exact);
tests/test_gpu_seek.py repeats it through the CUDA path."""
import numpy as np
import pytest

import l3util

CASES = {
    "m1_stereo_reservoir": dict(nframes=40, seed=31),
    "m1_joint_320_short": dict(nframes=40, seed=32, mode=1, bitrate_kbps=320, blocks=1, mixed_pct=25, fill_lo_pct=35),
    "m1_mono_vbr": dict(nframes=40, seed=33, mode=3, vbr_min_kbps=32, vbr_max_kbps=160, blocks=1),
    "lsf_22k": dict(nframes=40, seed=34, sample_rate=22050, bitrate_kbps=64, mode=1, blocks=1),
    "m25_8k": dict(nframes=40, seed=35, sample_rate=8000, bitrate_kbps=16, blocks=1, mode=3),
    "tagged": dict(nframes=40, seed=36, tag=1, tag_lame=1, enc_delay=576, enc_padding=700),
    "crc": dict(nframes=40, seed=37, crc=1),
    "layer2": dict(nframes=24, seed=38, layer=2, bitrate_kbps=192),
    "layer1": dict(nframes=60, seed=39, layer=1, bitrate_kbps=256),
}


@pytest.fixture(scope="module")
def m():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


@pytest.mark.parametrize("name", sorted(CASES))
def test_seek_plan_reproduces_the_tail_exactly(name, m, synth_mod, oracle_mod):
    s = synth_mod.make_stream(**CASES[name])
    full = oracle_mod.decode(s)
    frames, info, _ = m.index_stream_host(s)
    assert len(frames) == full.frames
    spf = full.samples // full.frames
    rng = np.random.default_rng(7)
    targets = [0, 1, spf - 1, spf, 2 * spf + 17, full.samples - 1] + list(rng.integers(0, full.samples, 6))
    for t in targets:
        sk = m.seek_plan(s, int(t), frames)
        assert sk.target_frame == t // spf and sk.first_frame <= sk.target_frame
        assert sk.byte_offset == frames["offset"][sk.first_frame]
        assert sk.target_frame - sk.first_frame <= 14  # a bounded pre-roll: warm-up + at most 511 bytes of reservoir
        part = oracle_mod.decode(s[sk.byte_offset:])
        assert part.samples == full.samples - sk.first_frame * spf
        assert np.array_equal(part.pcm[:, sk.discard_samples:], full.pcm[:, t:]), (name, t)


def test_seek_plan_rejects_bad_arguments(m, synth_mod):
    s = synth_mod.make_stream(nframes=5, seed=1)
    frames, info, _ = m.index_stream_host(s)
    with pytest.raises(m.Mp3bError):
        m.seek_plan(s, 5 * 1152, frames)       # past the end
    with pytest.raises(m.Mp3bError):
        m.seek_plan(s, -1, frames)
    with pytest.raises(m.Mp3bError):
        m.seek_plan(s, 0, frames[:0])


def test_seek_plan_property(m, synth_mod, oracle_mod):
    """Random stream shapes and targets (hypothesis): the planned slice reproduces the tail exactly, the
    pre-roll stays bounded, and a slice never starts after its target."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(seed=st.integers(1, 10 ** 6), rate=st.sampled_from([44100, 48000, 32000, 22050, 16000, 11025, 8000]),
           mode=st.sampled_from([0, 1, 3]), blocks=st.sampled_from([0, 1]), vbr=st.booleans(), frac=st.floats(0.0, 0.999))
    def check(seed, rate, mode, blocks, vbr, frac):
        lsf = rate < 32000
        cfg = dict(nframes=24, seed=seed, sample_rate=rate, mode=mode, blocks=blocks, bitrate_kbps=64 if lsf else 128,
                   fill_lo_pct=30)
        if vbr:
            cfg.update(vbr_min_kbps=32 if lsf else 64, vbr_max_kbps=128 if lsf else 256)
        s = synth_mod.make_stream(**cfg)
        full = oracle_mod.decode(s)
        frames, info, _ = m.index_stream_host(s)
        t = int(frac * full.samples)
        sk = m.seek_plan(s, t, frames)
        spf = full.samples // full.frames
        assert sk.first_frame <= sk.target_frame == t // spf
        part = oracle_mod.decode(s[sk.byte_offset:])
        assert np.array_equal(part.pcm[:, sk.discard_samples:], full.pcm[:, t:])

    check()
