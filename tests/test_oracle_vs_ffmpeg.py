"""Differential test: the from-spec oracle against FFmpeg's mp3float on generated streams.

This is the independent pin for every non-derivable table (all code books, count1 A/B, linbits,
sfb partitions, LSF scalefactor partitions, window D[]): generator, oracle and FFmpeg would all have
to agree on a wrong value for an error to go unnoticed.
"""
import numpy as np
import pytest

import cases
import ffmpeg_ref
import l3util

pytestmark = pytest.mark.skipif(not ffmpeg_ref.available(), reason="libavcodec not loadable")


@pytest.mark.parametrize("name", sorted(cases.FF))
def test_oracle_vs_mp3float(name, oracle_mod, synth_mod):
    s = synth_mod.make_stream(**cases.FF[name])
    d = oracle_mod.decode(s, dumps=True)
    frames = l3util.split_frames(s)
    assert len(frames) == d.frames == cases.FF[name]["nframes"]
    pcm, per = ffmpeg_ref.decode_frames(frames, d.channels)
    assert all(p is not None for p in per)
    assert pcm.shape == d.pcm.shape
    assert np.abs(d.pcm).max() > 1e-3, "degenerate (silent) test stream"
    rms, mx = l3util.iso_compliance(pcm, d.pcm)
    assert rms < 5e-7 and mx < 1e-5, (rms, mx)
    if name.startswith("table_"):
        t = cases.FF[name]["only_table"]
        lin = [0] * 16 + [1, 2, 3, 4, 6, 8, 10, 13, 4, 5, 6, 7, 8, 9, 11, 13]
        dim = {1: 2, 2: 3, 3: 3, 5: 4, 6: 4, 7: 6, 8: 6, 9: 6, 10: 8, 11: 8, 12: 8, 13: 16, 15: 16}.get(t, 16)
        maxv = dim - 1 + ((1 << lin[t]) - 1 if lin[t] else 0)
        assert np.abs(d.is_).max() == maxv, "the largest value of the book was not exercised"


@pytest.mark.parametrize("name", sorted(cases.L2))
def test_oracle_layer2_vs_mp2float(name, oracle_mod, synth_mod):
    """Layer II: the allocation / quantisation tables, grouping, scalefactors and joint-stereo bound of the
    oracle against FFmpeg's mp2float on generated streams."""
    s = synth_mod.make_stream(**cases.L2[name])
    d = oracle_mod.decode(s, dumps=True)
    frames = l3util.split_frames(s)
    assert len(frames) == d.frames == cases.L2[name]["nframes"] and d.samples == 1152 * d.frames
    pcm, per = ffmpeg_ref.decode_frames(frames, d.channels, b"mp2float")
    assert all(p is not None for p in per)
    assert pcm.shape == d.pcm.shape
    assert np.abs(d.pcm).max() > 1e-3, "degenerate (silent) test stream"
    rms, mx = l3util.iso_compliance(pcm, d.pcm)
    scale = max(1.0, float(np.abs(d.pcm).max()))
    assert rms < 5e-7 * scale and mx < 1e-5 * scale, (rms, mx)


@pytest.mark.parametrize("name", sorted(cases.L1))
def test_oracle_layer1_vs_mp1float(name, oracle_mod, synth_mod):
    """Layer I: allocation, scalefactors, requantisation and the joint-stereo bound of the oracle against
    FFmpeg's mp1float on generated streams (384 samples per frame)."""
    s = synth_mod.make_stream(**cases.L1[name])
    d = oracle_mod.decode(s, dumps=True)
    frames = l3util.split_frames(s)
    assert len(frames) == d.frames == cases.L1[name]["nframes"] and d.samples == 384 * d.frames
    pcm, per = ffmpeg_ref.decode_frames(frames, d.channels, b"mp1float")
    assert all(p is not None for p in per)
    assert pcm.shape == d.pcm.shape
    assert np.abs(d.pcm).max() > 1e-3, "degenerate (silent) test stream"
    rms, mx = l3util.iso_compliance(pcm, d.pcm)
    scale = max(1.0, float(np.abs(d.pcm).max()))
    assert rms < 5e-7 * scale and mx < 1e-5 * scale, (rms, mx)
