"""GPU: tag frame parsing in the device / host frame walk and the gapless window, through the C-ABI,
against the oracle's own parser (oracle/l3_oracle.c::l3o_parse_tag) on the same bytes."""
import numpy as np
import pytest

import l3util
from test_tags_cpu import CASES, FIELDS, silent_frames, tag_frame

pytestmark = pytest.mark.gpu

GEN = [
    dict(tag=2, tag_lame=1, enc_delay=576, enc_padding=1200, nframes=10, seed=1),
    dict(tag=1, tag_lame=1, enc_delay=1105, enc_padding=700, vbr_min_kbps=32, vbr_max_kbps=320, nframes=9, seed=2, blocks=1),
    dict(tag=1, tag_lame=0, nframes=7, seed=3, mode=1, blocks=1),
    dict(tag=3, enc_delay=64, sample_rate=22050, bitrate_kbps=16, mode=3, nframes=12, seed=4),
    dict(tag=2, tag_lame=1, enc_delay=0, enc_padding=4095, sample_rate=8000, bitrate_kbps=8, mode=3, nframes=12, seed=5),
    dict(tag=0, nframes=5, seed=6),
    dict(tag=2, tag_lame=1, enc_delay=2000, enc_padding=3000, nframes=3, seed=7, crc=1),
]


def _streams(synth_mod):
    out = [synth_mod.make_stream(**kw) for kw in GEN]
    for name, hdr, side, spf in CASES:
        out.append(tag_frame(hdr, side, b"Info", 0x0F, 9, 1234, lame=(576, 1105)) + silent_frames(hdr, 9))
        out.append(b"ID3\x03\x00\x00\x00\x00\x00\x10" + bytes(16) +
                   tag_frame(hdr, side, b"Xing", 0x01, 77, 0) + silent_frames(hdr, 4))
    hdr = bytes.fromhex("fffa9000")
    out.append(tag_frame(hdr, 32, b"Info", 0x0F, 5, 10, lame=(1000, 2000), crc=True, shift_for_crc=True) + silent_frames(hdr, 5))
    return out


@pytest.mark.parametrize("indexer", ["device", "host"])
def test_tag_info_and_gapless_window(indexer, synth_mod, oracle_mod):
    import mp3_b200 as m
    streams = _streams(synth_mod)
    idx = m.INDEX_DEVICE if indexer == "device" else m.INDEX_HOST
    with m.Decoder(device=0, pcm_format=m.PCM_F32, indexer=idx) as full, \
            m.Decoder(device=0, pcm_format=m.PCM_F32, indexer=idx, gapless=True) as gap:
        full.decode_batch(streams)
        gap.decode_batch(streams)
        a_full, a_gap = full.fetch_pcm(), gap.fetch_pcm()
        assert np.array_equal(a_full, a_gap)  # the arena is the same, only the window differs
        for i, s in enumerate(streams):
            want = oracle_mod.parse_tag(s)
            got = full.tag_info(i)
            assert tuple(getattr(got, k) for k in FIELDS) == tuple(getattr(want, k) for k in FIELDS), i
            ref = oracle_mod.decode(s).pcm.T
            whole = full.stream_pcm(i, a_full)
            assert whole.shape == ref.shape
            win = gap.stream_pcm(i, a_gap)
            assert gap.stream_info(i).samples == want.num_samples == win.shape[0]
            assert np.array_equal(win, whole[want.first_sample: want.first_sample + want.num_samples])
            l3util.assert_iso_full_accuracy(win, ref[want.first_sample: want.first_sample + want.num_samples], "stream %d" % i)


def test_crc_verification_matches_oracle(synth_mod, oracle_mod):
    """opts.verify_crc: a protected frame whose CRC-16 fails is concealed (and counted) exactly like the
    oracle's 11172-3 2.4.3.1 check does; intact protected streams are unaffected."""
    import mp3_b200 as m
    cfgs = [dict(nframes=8, seed=90, crc=1, bitrate_kbps=192, sample_rate=48000),
            dict(nframes=10, seed=91, crc=1, mode=1, blocks=1),
            dict(nframes=12, seed=92, crc=1, sample_rate=22050, bitrate_kbps=64, mode=3, blocks=1),
            dict(nframes=6, seed=93, crc=0)]
    streams = [synth_mod.make_stream(**c) for c in cfgs]
    damaged = []
    for k, s in enumerate(streams):
        frames = l3util.split_frames(s)
        b = bytearray(s)
        for j in (2, 4):
            off = sum(len(f) for f in frames[:j])
            b[off + 6 + 3 + k] ^= 0x04 << (k % 3)  # a bit inside the side info of frames 2 and 4
        damaged.append(bytes(b))
    for verify in (False, True):
        with m.Decoder(device=0, pcm_format=m.PCM_F32, verify_crc=verify) as dec:
            for batch in (streams, damaged):
                dec.decode_batch(batch)
                arena = dec.fetch_pcm()
                refs = [oracle_mod.decode(s, verify_crc=verify) for s in batch]
                assert dec.stats().concealed_frames == sum(r.concealed_frames for r in refs)
                if verify and batch is damaged:
                    assert [r.concealed_frames for r in refs][:3] == [2, 2, 2]
                for i, r in enumerate(refs):
                    l3util.assert_iso_full_accuracy(dec.stream_pcm(i, arena), r.pcm.T, "stream %d verify=%s" % (i, verify))
