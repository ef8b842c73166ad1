"""Sample-rate conversion (mp3b_batch_resample): the filter's properties (CPU) and the CUDA kernel
against scipy's upfirdn -- an independent implementation of "zero-stuff by L, filter, keep every M-th
sample" -- evaluated in float64 on the filter the library reports (GPU)."""
import numpy as np
import pytest


def full_filter(taps, L):
    T = taps.shape[1]
    h = np.zeros(L * T)
    for p in range(L):
        h[p::L][:T] = taps[p]
    return h  # h[p + j L] = taps[p][j]; centred at (T // 2) * L


@pytest.mark.parametrize("rin,rout", [(44100, 48000), (48000, 44100), (44100, 22050), (8000, 48000), (22050, 44100),
                                      (32000, 44100), (44100, 44100)])
def test_filter_properties(rin, rout):
    import mp3_b200
    taps, L, M = mp3_b200.resample_filter(rin, rout)
    g = np.gcd(rin, rout)
    assert (L, M) == (rout // g, rin // g)
    if rin == rout:
        assert taps.shape == (1, 1) and taps[0, 0] == 1.0
        return
    T = taps.shape[1]
    h = full_filter(taps.astype(np.float64), L)
    D = (T // 2) * L
    assert np.allclose(h[: D + 1], h[2 * D:: -1][: D + 1], atol=1e-7)      # linear phase, centre at D
    assert abs(h.sum() / L - 1.0) < 1e-4                                    # unity gain at DC
    n = 1 << 20
    H = np.abs(np.fft.rfft(h, n)) / L
    f = np.fft.rfftfreq(n)
    ny = 0.5 / max(L, M)
    assert np.abs(20 * np.log10(H[f < 0.8 * ny])).max() < 0.01             # flat pass band
    assert 20 * np.log10(H[f >= ny].max()) < -85.0                          # nothing left to alias / image


@pytest.mark.parametrize("rin,rout,nk", [(44100, 48000, 5), (22050, 44100, 1), (32000, 48000, 3), (8000, 48000, 3),
                                         (44100, 96000, 5), (24000, 44100, 147)])
def test_tensor_core_plan_is_the_polyphase_filter(rin, rout, nk):
    """The tensor-core resampler's host-side plan (no GPU needed): every tile kind's banded matrix, applied to the input
    window of a tile of that kind, gives the 128 outputs of the polyphase definition -- the fp16 split of the
    coefficients (c1 + c2) costs less than 1e-6 of the filter's gain."""
    import mp3_b200
    taps, L, M = mp3_b200.resample_filter(rin, rout)
    T = taps.shape[1]
    half = T // 2
    mats = mp3_b200.resample_tc_plan(rin, rout)
    assert mats is not None and len(mats) == nk
    K = mats[0].shape[1]
    assert K % 16 == 0 and K <= 192
    rng = np.random.default_rng(rin + rout)
    x = rng.integers(-32768, 32768, 40000).astype(np.float64)
    for t in list(range(min(nk, 7))) + [nk + 1, 3 * nk + 2]:
        k = t % nk
        n0 = 128 * t
        base = (n0 * M + half * L) // L - (T - 1)
        win = np.array([x[base + c] if 0 <= base + c < x.size else 0.0 for c in range(K)])
        got = mats[k].astype(np.float64) @ win
        ref = np.empty(128)
        for r in range(128):
            u = (n0 + r) * M + half * L
            q, p = divmod(u, L)
            ref[r] = sum(float(taps[p, j]) * (x[q - j] if 0 <= q - j < x.size else 0.0) for j in range(T))
        assert np.abs(got - ref).max() < 1e-6 * 32768 * np.abs(taps).sum(axis=1).max(), (t, np.abs(got - ref).max())
    assert mp3_b200.resample_tc_plan(48000, 44100) is None      # downwards: the window does not fit, FP32 kernels
    assert mp3_b200.resample_tc_plan(32000, 44100) is None      # 441 kinds: too many


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["f32", "s16"])
@pytest.mark.parametrize("out_rate", [48000, 44100, 32000, 22100, 22050, 16000])
def test_resample_matches_upfirdn(fmt, out_rate, synth_mod):
    from scipy.signal import upfirdn
    import mp3_b200 as m
    cfgs = [dict(nframes=10, seed=1), dict(nframes=10, seed=2, mode=3, sample_rate=22050, bitrate_kbps=32),
            dict(nframes=12, seed=3, sample_rate=8000, bitrate_kbps=16, mode=1), dict(nframes=6, seed=4, sample_rate=48000),
            dict(nframes=8, seed=5, tag=2, tag_lame=1, enc_delay=576, enc_padding=1000)]
    streams = [synth_mod.make_stream(**c) for c in cfgs] + [b"not an mp3 stream"]
    with m.Decoder(device=0, pcm_format=m.PCM_F32 if fmt == "f32" else m.PCM_S16, gapless=True) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm().copy()
        dec.resample(out_rate)
        out, where = dec.fetch_resampled()
        assert where[-1][1] == 0
        for i in range(len(cfgs)):
            inf = dec.stream_info(i)
            x = dec.stream_pcm(i, arena).astype(np.float64)
            if fmt == "s16":
                x /= 32768.0
            taps, L, M = m.resample_filter(inf.sample_rate, out_rate)
            T = taps.shape[1]
            D = (T // 2) * L
            nout = -(-inf.samples * L // M)
            off, cnt = where[i]
            assert cnt == nout
            got = out[off: off + cnt * inf.channels].reshape(cnt, inf.channels).astype(np.float64)
            if L == 1 and M == 1:
                ref = x
            else:
                # y[n] = sum_i x[i] h[n M + D - i L]: upfirdn computes sum_i x[i] g[m M - i L]; shifting the
                # filter by r = (-D) mod M puts n M + D on its output grid at m = n + (D + r) / M
                r = (-D) % M
                g = np.concatenate([np.zeros(r), full_filter(taps.astype(np.float64), L)])
                k = (D + r) // M
                ref = np.stack([upfirdn(g, x[:, c], up=L, down=M)[k: k + nout] for c in range(inf.channels)], axis=1)
                if ref.shape[0] < nout:
                    ref = np.pad(ref, ((0, nout - ref.shape[0]), (0, 0)))
            if fmt == "f32":
                assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
            else:
                want = np.clip(np.round(ref * 32768.0), -32768, 32767)
                assert np.abs(got - want).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("out_rate", [48000, 44100, 96000])
def test_tensor_core_path_matches_upfirdn(out_rate, synth_mod, monkeypatch):
    """The tensor-core resampler (stereo s16, forced on with MP3B_RS_TC=1 -- by default it serves large batches only)
    against the same float64 reference, within the same 1 LSB, and against the FP32 kernels' output: streams of several
    input rates, lengths that are not a multiple of the 128-output tile, a stream shorter than one tile, a mono stream
    in between (which stays on the FP32 kernel)."""
    from scipy.signal import upfirdn
    import mp3_b200 as m
    cfgs = [dict(nframes=40, seed=11, mode=1), dict(nframes=3, seed=12), dict(nframes=9, seed=13, mode=3),
            dict(nframes=25, seed=14, sample_rate=22050, bitrate_kbps=64), dict(nframes=1, seed=15, sample_rate=32000),
            dict(nframes=30, seed=16, sample_rate=32000, mode=1, bitrate_kbps=192), dict(nframes=17, seed=17, sample_rate=8000,
                                                                                        bitrate_kbps=16, mode=1),
            dict(nframes=21, seed=18, sample_rate=48000)]
    streams = [synth_mod.make_stream(**c) for c in cfgs]
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("MP3B_RS_TC", tc)
        with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
            dec.decode_batch(streams)
            arena = dec.fetch_pcm().copy()
            dec.resample(out_rate)
            out, where = dec.fetch_resampled()
            outs[tc] = (out.copy(), where)
            if tc == "0":
                continue
            for i in range(len(cfgs)):
                inf = dec.stream_info(i)
                x = dec.stream_pcm(i, arena).astype(np.float64) / 32768.0
                taps, L, M = m.resample_filter(inf.sample_rate, out_rate)
                D = (taps.shape[1] // 2) * L
                nout = -(-inf.samples * L // M)
                off, cnt = where[i]
                assert cnt == nout
                got = out[off: off + cnt * inf.channels].reshape(cnt, inf.channels).astype(np.float64)
                if L == 1 and M == 1:
                    ref = x
                else:
                    r = (-D) % M
                    g = np.concatenate([np.zeros(r), full_filter(taps.astype(np.float64), L)])
                    k = (D + r) // M
                    ref = np.stack([upfirdn(g, x[:, c], up=L, down=M)[k: k + nout] for c in range(inf.channels)], axis=1)
                    if ref.shape[0] < nout:
                        ref = np.pad(ref, ((0, nout - ref.shape[0]), (0, 0)))
                want = np.clip(np.round(ref * 32768.0), -32768, 32767)
                assert np.abs(got - want).max() <= 1, (i, cfgs[i])
    a, b = outs["1"][0].astype(np.int32), outs["0"][0].astype(np.int32)
    assert outs["1"][1] == outs["0"][1]
    assert np.abs(a - b).max() <= 1          # both within 1 LSB of the exact value's rounding ...
    assert np.mean(a != b) < 0.02            # ... and they differ only where that value sits at a rounding boundary


@pytest.mark.gpu
def test_unusable_rate_pair_is_refused(synth_mod):
    """A rate pair whose ratio needs more than 4,096 phases is refused with an error, not attempted."""
    import mp3_b200 as m
    with m.Decoder(device=0) as dec:
        dec.decode_batch([synth_mod.make_stream(nframes=4, seed=1)])
        with pytest.raises(m.Mp3bError):
            dec.resample(47999)
        dec.resample(48000)  # the context is still usable
        out, where = dec.fetch_resampled()
        assert where[0][1] > 0
