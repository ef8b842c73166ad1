"""Tag frame (Xing / Info / LAME / VBRI) parsing and the gapless window: the oracle against frames built
byte by byte here from the published tag layouts, and against the generator's tag frames."""
import struct

import pytest

FIELDS = ("kind", "has_lame", "frames", "bytes", "enc_delay", "enc_padding", "first_sample", "num_samples")


def tag_frame(hdr, side_len, ident=b"Info", flags=0x0F, frames=0, nbytes=0, lame=None, crc=False, shift_for_crc=False):
    """One Layer III frame with zero side info whose payload is a Xing/Info tag.  lame = (delay, padding)."""
    from l3util import frame_len
    n = frame_len(hdr)
    f = bytearray(n)
    f[:4] = hdr
    at = 4 + side_len + (2 if (crc and shift_for_crc) else 0)
    body = bytearray(ident) + struct.pack(">I", flags)
    if flags & 1:
        body += struct.pack(">I", frames)
    if flags & 2:
        body += struct.pack(">I", nbytes)
    if flags & 4:
        body += bytes(range(100))
    if flags & 8:
        body += struct.pack(">I", 50)
    if lame is not None:
        d, p = lame
        ext = bytearray(b"LAME3.99r") + bytes([0x13, 180]) + bytes(8) + bytes([0x20, 128])
        ext += bytes([d >> 4, ((d & 15) << 4) | (p >> 8), p & 255])
        body += ext + bytes(12)
    f[at:at + len(body)] = body
    return bytes(f)


def silent_frames(hdr, n):
    from l3util import frame_len
    return (bytes(hdr) + bytes(frame_len(hdr) - 4)) * n


CASES = [
    # name, header bytes, side-info length, samples per frame
    ("m1_stereo", bytes.fromhex("fffb9000"), 32, 1152),
    ("m1_mono", bytes.fromhex("fffb90c0"), 17, 1152),
    ("lsf_stereo", bytes.fromhex("fff38000"), 17, 576),
    ("lsf_mono", bytes.fromhex("fff380c0"), 9, 576),
    ("m25_stereo", bytes.fromhex("ffe38000"), 17, 576),
]


@pytest.mark.parametrize("name,hdr,side,spf", CASES)
def test_info_lame_known_answers(name, hdr, side, spf, oracle_mod):
    audio = silent_frames(hdr, 9)
    s = tag_frame(hdr, side, b"Info", 0x0F, 9, 1234, lame=(576, 1105)) + audio
    t = oracle_mod.parse_tag(s)
    got = tuple(getattr(t, k) for k in FIELDS)
    assert got == (2, 1, 9, 1234, 576, 1105, spf + 576 + 529, 9 * spf - 576 - 1105)
    # Xing with the frame count only, no LAME extension: only the tag frame is dropped
    s = tag_frame(hdr, side, b"Xing", 0x01, 77, 0) + audio
    t = oracle_mod.parse_tag(s)
    assert tuple(getattr(t, k) for k in FIELDS) == (1, 0, 77, 0, 0, 0, spf, 9 * spf)
    # no tag at all
    t = oracle_mod.parse_tag(audio)
    assert tuple(getattr(t, k) for k in FIELDS) == (0, 0, 0, 0, 0, 0, 0, 9 * spf)
    # padding larger than what was decoded: the window clamps at zero length, never negative
    s = tag_frame(hdr, side, b"Info", 0x03, 1, 99, lame=(4095, 4095)) + silent_frames(hdr, 1)
    t = oracle_mod.parse_tag(s)
    assert t.num_samples == 0 and 0 <= t.first_sample <= 2 * spf


def test_tag_behind_crc_word_and_id3(oracle_mod):
    hdr = bytes.fromhex("fffa9000")  # protection bit 0: a CRC word follows the header
    for shift in (False, True):
        s = tag_frame(hdr, 32, b"Info", 0x0F, 5, 10, lame=(1000, 2000), crc=True, shift_for_crc=shift) + silent_frames(hdr, 5)
        id3 = b"ID3\x03\x00\x00\x00\x00\x00\x10" + bytes(16)
        t = oracle_mod.parse_tag(id3 + s)
        assert (t.kind, t.has_lame, t.enc_delay, t.enc_padding) == (2, 1, 1000, 2000), shift


def test_vbri_known_answer(oracle_mod):
    hdr = bytes.fromhex("fffb9000")
    f = bytearray(bytes(hdr) + bytes(413))
    f[36:36 + 18] = b"VBRI" + struct.pack(">HHHII", 1, 2000, 75, 123456, 321)
    t = oracle_mod.parse_tag(bytes(f) + silent_frames(hdr, 3))
    assert (t.kind, t.has_lame, t.frames, t.bytes, t.enc_delay, t.first_sample, t.num_samples) == (
        3, 0, 321, 123456, 2000, 1152, 3 * 1152)


@pytest.mark.parametrize("kw", [
    dict(tag=2, tag_lame=1, enc_delay=576, enc_padding=1200),
    dict(tag=1, tag_lame=1, enc_delay=1105, enc_padding=700, vbr_min_kbps=32, vbr_max_kbps=320),
    dict(tag=1, tag_lame=0),
    dict(tag=3, enc_delay=64, sample_rate=22050, bitrate_kbps=16, mode=3),
    dict(tag=2, tag_lame=1, enc_delay=0, enc_padding=4095, sample_rate=8000, bitrate_kbps=8, mode=3),
])
def test_generator_tags_round_trip(kw, oracle_mod, synth_mod):
    s = synth_mod.make_stream(nframes=8, seed=77, **kw)
    t = oracle_mod.parse_tag(s)
    d = oracle_mod.decode(s, want_pcm=False)
    assert d.frames == 9  # the tag frame decodes as a (silent) frame of its own
    assert t.kind == kw["tag"] and t.frames == 8 and t.bytes == len(s)
    assert t.has_lame == (1 if kw.get("tag_lame") and kw["tag"] != 3 else 0)
    assert t.enc_delay == kw.get("enc_delay", 0)
    spf = d.samples // d.frames
    if t.has_lame:
        assert t.enc_padding == kw["enc_padding"]
        assert t.first_sample == min(spf + kw["enc_delay"] + 529, d.samples)
        assert t.num_samples == max(0, min(8 * spf - kw["enc_delay"] - kw["enc_padding"], d.samples - t.first_sample))
    else:
        assert (t.first_sample, t.num_samples) == (spf, 8 * spf)
