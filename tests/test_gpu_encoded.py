"""Real-signal streams (music-like and speech-like audio through the in-tree encoder, bit reservoir in use) through the
CUDA path: Huffman output and scalefactors bit-exact against the oracle on every Huffman kernel, PCM under the ISO/IEC
11172-4 full-accuracy criterion, the decoded audio equal to the encoder's INPUT within the noise the bit rate allows and
without a click, one-shot and through the incremental interface -- alone and in a batch with generator streams."""
import numpy as np
import pytest

import l3util
import signals
from test_encoder_cpu import CASES, _input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def encoded(synth_mod, oracle_mod):
    srcs, streams, refs = [], [], []
    for name, sr, nch, kbps, seconds, min_snr in CASES:
        p = _input(sr, nch, seconds)
        s = synth_mod.encode_pcm(p, sr, kbps)
        srcs.append(p)
        streams.append(s)
        refs.append(oracle_mod.decode(s, dumps=True))
    return srcs, streams, refs


def test_window_switching_stream(synth_mod, oracle_mod):
    """Attacks coded with start / short / stop windows by the in-tree encoder: bit-exact through Huffman, ISO accuracy."""
    import mp3_b200 as m
    x = signals.castanets(44100, 2.0)
    pcm16 = signals.to_s16(np.stack([x, 0.6 * x], axis=1))
    s = synth_mod.encode_pcm(pcm16, 44100, 192, short_blocks=True)
    r = oracle_mod.decode(s, dumps=True)
    with m.Decoder(device=0, pcm_format=m.PCM_F32, keep_stages=True) as dec:
        dec.decode_batch([s])
        assert np.array_equal(dec.stage(m.STAGE_IS)[: r.units], r.is_)
        got = dec.stream_pcm(0, dec.fetch_pcm()).astype(np.float64)
        l3util.assert_iso_full_accuracy(got, r.pcm.T, "window switching")
        snr, worst = signals.snr_db(got, pcm16)
        assert snr > 8.0


@pytest.mark.parametrize("k1", ["auto", "chunk", "warp", "sorted"])
def test_encoded_streams_bit_exact_and_iso(k1, encoded, synth_mod, monkeypatch):
    import mp3_b200 as m
    srcs, streams, refs = encoded
    if k1 != "auto":
        monkeypatch.setenv("MP3B_K1_MODE", k1)
    filler = [synth_mod.make_stream(nframes=30, seed=300 + i, mode=1, blocks=1, bitrate_kbps=192) for i in range(3)]
    batch = [filler[0]] + streams[:2] + [filler[1]] + streams[2:] + [filler[2]]
    where = [1, 2, 4, 5, 6]
    with m.Decoder(device=0, pcm_format=m.PCM_F32, keep_stages=True) as dec:
        dec.decode_batch(batch)
        arena = dec.fetch_pcm()
        is_, sf = dec.stage(m.STAGE_IS), dec.stage(m.STAGE_SF)
        assert dec.stats().concealed_frames == 0
        for (name, sr, nch, kbps, seconds, min_snr), src, r, w in zip(CASES, srcs, refs, where):
            ub = dec.stream_info(w).pcm_offset // 576
            assert np.array_equal(is_[ub: ub + r.units], r.is_), name
            assert np.array_equal(sf[ub: ub + r.units], r.sf), name
            got = dec.stream_pcm(w, arena).astype(np.float64)
            l3util.assert_iso_full_accuracy(got, r.pcm.T, name)
            snr, worst = signals.snr_db(got, src)
            assert snr > min_snr - 3.0 and worst < 0.25, (name, snr, worst)


def test_encoded_streams_s16_and_incremental(encoded):
    import mp3_b200 as m
    srcs, streams, refs = encoded
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm().copy()
        one_shot = [dec.stream_pcm(i, arena).copy() for i in range(len(streams))]
        for got, r in zip(one_shot, refs):
            want = np.clip(np.rint(r.pcm.T * 32768.0), -32768, 32767).astype(np.int64)
            assert np.abs(got.astype(np.int64) - want).max() <= 1
        hs = [dec.open_stream() for _ in streams]
        parts = [[] for _ in streams]
        for p in range(0, max(len(s) for s in streams), 1500):
            for h, s in zip(hs, streams):
                if p < len(s):
                    h.enqueue(s[p: p + 1500])
            dec.decode_streams()
            for j, h in enumerate(hs):
                inf = h.info()
                if inf.samples:
                    parts[j].append(h.fetch(inf.samples))
        for j, h in enumerate(hs):
            assert np.array_equal(np.concatenate(parts[j]), one_shot[j])
            h.close()
