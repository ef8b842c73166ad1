"""Oracle vs the known-answer vectors of SURVEY.md section 8(c) and the committed FFmpeg golden PCM."""
import os

import numpy as np
import pytest

import l3util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ff_golden.npz")

KAT0 = bytes.fromhex("fffb9000") + bytes(413)
KAT1 = (bytes.fromhex("fffb9000")
        + bytes.fromhex("0000000401690021080000000D20042100000001A40084200000003480108400")
        + bytes([0x50]) + bytes(417 - 4 - 32 - 1))


def test_kat0_silence(oracle_mod):
    d = oracle_mod.decode(KAT0 * 3)
    assert d.frames == 3 and d.channels == 2 and d.sample_rate == 44100 and d.samples == 3456
    assert np.all(d.pcm == 0.0)


def test_kat1_single_line(oracle_mod):
    # numbers measured from FFmpeg mp3float in the survey session
    d = oracle_mod.decode(KAT1 + KAT0 + KAT0, dumps=True)
    left, right = d.pcm
    assert d.is_[0, 0] == 1 and np.count_nonzero(d.is_) == 1
    assert np.all(right == 0.0)
    nz = np.nonzero(np.abs(left) > 1e-12)[0]
    assert nz[0] == 1 and nz[-1] == 1631
    k = int(np.argmax(np.abs(left)))
    assert k == 922 and abs(left[k] - (-0.8535740971565247)) < 2e-6
    assert abs(np.sum(left * left) - 288.0059985) < 1e-3
    for idx, val in ((240, 0.04973330349), (528, -0.07968214154), (800, -0.74641698599),
                     (1152, -0.50985199213), (1300, -0.12107632309)):
        assert abs(left[idx] - val) < 3e-6


def test_kat1_gain_step(oracle_mod):
    # global_gain 170 instead of 210 scales by exactly 2^-10
    side = bytearray(KAT1[4:36])
    # global_gain field of gr0/ch0 starts at bit 9+3+8+12+9 = 41
    bits = int.from_bytes(side, "big")
    total = 256
    shift = total - 41 - 8
    bits = (bits & ~(0xFF << shift)) | (170 << shift)
    f = KAT1[:4] + bits.to_bytes(32, "big") + KAT1[36:]
    a = oracle_mod.decode(KAT1 + KAT0 + KAT0).pcm[0]
    b = oracle_mod.decode(f + KAT0 + KAT0).pcm[0]
    assert np.allclose(b * 1024.0, a, rtol=0, atol=1e-15)


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture missing")
def test_oracle_matches_committed_ffmpeg_golden(oracle_mod):
    z = np.load(GOLD)
    names = sorted(k[:-4] for k in z.files if k.endswith(".mp3"))
    assert len(names) >= 10
    for name in names:
        d = oracle_mod.decode(z[name + ".mp3"].tobytes())
        ref = z[name + ".pcm"]
        assert d.pcm.shape == ref.shape, name
        # FFmpeg computes in float32: allow its rounding noise, far below the ISO limits
        rms, mx = l3util.iso_compliance(d.pcm, ref)
        assert rms < 5e-7 and mx < 1e-5, (name, rms, mx)


def test_crc16_check_value(oracle_mod):
    """CRC-16 with polynomial 0x8005, preset 0xFFFF, MSB first, no final xor: check value of "123456789"."""
    import ctypes
    L = oracle_mod.lib()
    L.l3o_crc16.restype = ctypes.c_uint
    L.l3o_crc16.argtypes = [ctypes.c_uint, ctypes.c_char_p, ctypes.c_size_t]
    assert L.l3o_crc16(0xFFFF, b"123456789", 72) == 0xAEE7


def test_crc_verification_conceals_exactly_the_damaged_frame(oracle_mod, synth_mod):
    import l3util
    s = synth_mod.make_stream(nframes=8, seed=90, crc=1, bitrate_kbps=192, sample_rate=48000)
    good = oracle_mod.decode(s, verify_crc=True)
    assert good.concealed_frames == 0  # the generator's own CRC routine agrees with the oracle's
    frames = l3util.split_frames(s)
    off = sum(len(f) for f in frames[:3])
    bad = bytearray(s)
    bad[off + 6 + 5] ^= 0x10  # one bit of frame 3's side info (inside a global_gain field)
    d_off = oracle_mod.decode(bytes(bad))
    d_on = oracle_mod.decode(bytes(bad), verify_crc=True)
    assert d_off.concealed_frames == 0 and d_on.concealed_frames == 1
    assert d_on.pcm.shape == d_off.pcm.shape == good.pcm.shape
