"""WSOLA time stretch (mp3b_batch_time_stretch): the numpy oracle's own properties (CPU) and the CUDA
kernel against it (GPU): alignment offsets bit-exact, PCM within float rounding / 1 LSB."""
import numpy as np
import pytest

from oracle import wsola


def test_oracle_preserves_pitch_and_changes_duration():
    sr, f0 = 44100, 440.0
    t = np.arange(sr) / sr
    x = (0.5 * np.sin(2 * np.pi * f0 * t))[:, None]
    for num, den in ((1, 2), (3, 4), (5, 4), (2, 1)):
        y, offs = wsola.wsola(x, wsola.to_s16(x), sr, num, den)
        assert y.shape[0] == sr * den // num
        seg = y[4096: 4096 + 16384, 0] * np.hanning(16384)
        peak = np.argmax(np.abs(np.fft.rfft(seg))) * sr / 16384.0
        assert abs(peak - f0) < 4.0, (num, den, peak)            # same pitch
        assert 0.45 < np.abs(y[4096:-4096]).max() < 0.56         # no gaps, no doubling
        assert np.abs(offs).max() <= wsola.hop_of(sr) // 2
    # speed 1 on a signal without periodicity: every frame continues the previous one, the identity
    x = np.random.default_rng(7).uniform(-0.5, 0.5, (sr // 2, 2))
    y, offs = wsola.wsola(x, wsola.to_s16(x), sr, 1, 1)
    assert np.array_equal(offs, np.zeros_like(offs)) and np.allclose(y, x, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["f32", "s16"])
@pytest.mark.parametrize("speed", [(1, 2), (3, 4), (3, 2)])
def test_stretch_matches_oracle(fmt, speed, synth_mod):
    import mp3_b200 as m
    num, den = speed
    cfgs = [dict(nframes=12, seed=1), dict(nframes=14, seed=2, mode=3, sample_rate=22050, bitrate_kbps=32),
            dict(nframes=16, seed=3, sample_rate=8000, bitrate_kbps=16, mode=1, blocks=1),
            dict(nframes=8, seed=5, tag=2, tag_lame=1, enc_delay=576, enc_padding=1000, mode=1)]
    streams = [synth_mod.make_stream(**c) for c in cfgs] + [b"\x00" * 100]
    with m.Decoder(device=0, pcm_format=m.PCM_F32 if fmt == "f32" else m.PCM_S16, gapless=True) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm().copy()
        dec.time_stretch(num, den)
        out, where = dec.fetch_stretched()
        assert where[-1][1] == 0
        for i in range(len(cfgs)):
            inf = dec.stream_info(i)
            pcm = dec.stream_pcm(i, arena)
            if fmt == "s16":
                s16 = pcm.astype(np.int64)
                x = pcm.astype(np.float64) / 32768.0
            else:
                x = pcm.astype(np.float64)
                s16 = wsola.to_s16(pcm)
            ref, offs = wsola.wsola(x, s16, inf.sample_rate, num, den)
            got_offs, hop = dec.stretch_offsets(i)
            assert hop == wsola.hop_of(inf.sample_rate)
            assert np.array_equal(got_offs, offs), i
            off, cnt = where[i]
            assert cnt == ref.shape[0] == inf.samples * den // num
            got = out[off: off + cnt * inf.channels].reshape(cnt, inf.channels).astype(np.float64)
            if fmt == "f32":
                assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
            else:
                assert np.abs(got - np.clip(np.rint(ref * 32768.0), -32768, 32767)).max() <= 1
