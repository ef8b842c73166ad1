"""GPU: more streams than a CUDA grid's y / z dimension holds (65,535): every operator that maps streams onto
grid dimensions or per-stream job tables -- decode, planar copy, sentence boundaries, sample-rate conversion,
time stretch -- on 70,000 two-frame streams (copies of 7 distinct ones, so every copy must equal its first)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, DISTINCT = 70000, 7


@pytest.fixture(scope="module")
def many(synth_mod):
    base = [synth_mod.make_stream(nframes=2, seed=900 + k, mode=(3 if k % 3 == 2 else 0),
                                  **(dict(sample_rate=22050, bitrate_kbps=64) if k % 2 else {})) for k in range(DISTINCT)]
    return base, [base[i % DISTINCT] for i in range(N)]


def test_output_operators_on_70k_streams(many):
    import mp3_b200 as m
    base, streams = many
    probe = [0, 1, 2, 3, 4, 5, 6, 65534, 65535, 65536, 65537, N - 2, N - 1]
    with m.Decoder(device=0) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        infos = [dec.stream_info(i) for i in probe]
        for i, inf in zip(probe, infos):
            assert inf.frames == 2
            assert np.array_equal(dec.stream_pcm(i, arena), dec.stream_pcm(i % DISTINCT, arena)), i
        assert arena.any()
        # planar
        pl = dec.planar()
        for i, inf in zip(probe, infos):
            got = pl[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].reshape(inf.channels, inf.samples)
            assert np.array_equal(got.T, dec.stream_pcm(i, arena)), i
        # sentence boundaries: energies of a copy equal those of the first
        segs = dec.segments(threshold=50, min_silence_ms=20, min_sentence_ms=0)
        assert len(segs) == N
        for i in probe:
            assert np.array_equal(segs[i], segs[i % DISTINCT]), i
            assert np.array_equal(dec.window_energy(i)[0], dec.window_energy(i % DISTINCT)[0]), i
        assert any(len(segs[i]) for i in probe)
        # sample-rate conversion and time stretch
        dec.resample(48000)
        rs, where = dec.fetch_resampled()
        for i in probe:
            (o, c), (o0, c0) = where[i], where[i % DISTINCT]
            ch = dec.stream_info(i).channels
            assert c == c0 and c > 0 and np.array_equal(rs[o: o + c * ch], rs[o0: o0 + c0 * ch]), i
        dec.time_stretch(2, 3)
        ts, where = dec.fetch_stretched()
        for i in probe:
            (o, c), (o0, c0) = where[i], where[i % DISTINCT]
            ch = dec.stream_info(i).channels
            assert c == c0 and c > 0 and np.array_equal(ts[o: o + c * ch], ts[o0: o0 + c0 * ch]), i
