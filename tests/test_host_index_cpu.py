"""Host logic without a GPU: the frame walk shared by the host and device indexers (l3_defs.h:
l3_frame_at / l3_parse_hdr / l3_parse_tag, driven by api.cu: host_index_stream) through
mp3b_index_stream_host, against the oracle's own scan and under property tests."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import cases
import l3util


@pytest.fixture(scope="module")
def m():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


def check_consistent(frames, data):
    """Every indexed frame is a valid Layer III frame inside the buffer; frames do not overlap; the
    main-data offsets are the running sum of the frames' payload sizes."""
    end, payload = 0, 0
    for f in frames:
        off, hdr = int(f["offset"]), int(f["header"])
        hb = hdr.to_bytes(4, "big")
        n = l3util.frame_len(hb)
        assert n > 0 and off >= end and off + n <= len(data)
        assert data[off:off + 4] == hb
        assert int(f["payload_offset"]) == payload
        lsf = ((hb[1] >> 3) & 3) != 3
        mono = (hb[3] >> 6) == 3
        side = (9 if mono else 17) if lsf else (17 if mono else 32)
        payload += n - 4 - (0 if hb[1] & 1 else 2) - side
        end = off + n


@pytest.mark.parametrize("name", sorted(cases.FF)[:24] + sorted(cases.EXTRA))
def test_index_matches_oracle_scan(name, m, oracle_mod, synth_mod):
    kw = dict(cases.FF.get(name) or cases.EXTRA[name])
    s = synth_mod.make_stream(**kw)
    for data in (s, b"\x00\xff\xfb" * 7 + s, b"ID3\x04\x00\x00\x00\x00\x01\x00" + bytes(128) + s, s[:-5], s[3:]):
        frames, info, tag = m.index_stream_host(data)
        d = oracle_mod.decode(data, want_pcm=False)
        assert len(frames) == d.frames
        assert (info.sample_rate, info.channels, info.samples) == (d.sample_rate, d.channels, d.samples)
        check_consistent(frames, data)
        t = oracle_mod.parse_tag(data)
        assert (tag.kind, tag.first_sample, tag.num_samples) == (t.kind, t.first_sample, t.num_samples)


@settings(max_examples=150, deadline=None)
@given(st.binary(min_size=0, max_size=3000))
def test_random_bytes_never_break_the_walk(m, data):
    frames, info, tag = m.index_stream_host(data)
    check_consistent(frames, data)
    assert info.frames == len(frames)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(0, 40), st.integers(0, 60), st.lists(st.integers(0, 4000), max_size=6))
def test_damaged_streams_index_like_the_oracle(m, oracle_mod, synth_mod, seed, junk, cut, flips):
    """Junk in front, a truncated tail and flipped bytes: the walk still agrees with the oracle's scan."""
    s = bytearray(synth_mod.make_stream(nframes=6, seed=seed % 1000, blocks=1, mode=seed % 4,
                                        bitrate_kbps=[64, 128, 320][seed % 3]))
    for f in flips:
        if f < len(s):
            s[f] ^= 1 << (seed % 8)
    data = bytes([0x55]) * junk + bytes(s[: len(s) - cut])
    frames, info, tag = m.index_stream_host(data)
    check_consistent(frames, data)
    d = oracle_mod.decode(data, want_pcm=False)
    assert len(frames) == (d.frames if d.rc == 0 else 0)
