"""The parity suite again with MP3B_DEBUG_POISON=1: the library then fills every scratch and output buffer (frame
table, unit descriptors, main-data arena, Huffman output, scalefactors, vector counts, staged intermediates, subband
samples, the PCM arena) with 0xFF before each decode.  Buffers are grow-only and reused from call to call, and the
Huffman kernel does not write the all-zero tail of a spectrum (the back end re-creates it from the vector count):
a read of anything the current call did not write would be silent on a fresh (zeroed) allocation -- here it decodes
0xFFFF lines or NaN samples and fails the comparison.  Batches of different shapes alternate on ONE context, so
every buffer holds stale data of another batch as well."""
import numpy as np
import pytest

import cases
import l3util

pytestmark = pytest.mark.gpu

ALL = dict(cases.FF)
ALL.update(cases.EXTRA)
ALL.update(cases.L2)
ALL.update(cases.L1)
NAMES = sorted(ALL)


@pytest.fixture(scope="module", autouse=True)
def poison_env():
    import os
    old = os.environ.get("MP3B_DEBUG_POISON")
    os.environ["MP3B_DEBUG_POISON"] = "1"   # read by mp3b_ctx_create
    yield
    if old is None:
        del os.environ["MP3B_DEBUG_POISON"]
    else:
        os.environ["MP3B_DEBUG_POISON"] = old


@pytest.fixture(scope="module")
def batches(synth_mod, oracle_mod):
    streams = [synth_mod.make_stream(**ALL[n]) for n in NAMES]
    refs = [oracle_mod.decode(s, dumps="int") for s in streams]
    a = list(range(len(streams)))
    b = a[::-1][: len(a) // 2]           # a different shape: half the streams, reversed
    c = [k for k in a if k % 3 == 0]
    return streams, refs, [a, b, c, a]


def _check_pcm(dec, arena, order, refs, fmt):
    for i, k in enumerate(order):
        r = refs[k]
        got = dec.stream_pcm(i, arena)
        assert got.shape == r.pcm.T.shape, NAMES[k]
        if fmt == "s16":
            want = np.clip(np.rint(r.pcm.T * 32768.0), -32768, 32767).astype(np.int64)
            assert np.abs(got.astype(np.int64) - want).max() <= 1, NAMES[k]
        elif NAMES[k] == "loud_clipping":
            assert np.abs(got - r.pcm.T).max() / np.abs(r.pcm).max() < 1e-5
        else:
            assert np.all(np.isfinite(got)), NAMES[k]
            l3util.assert_iso_full_accuracy(got.astype(np.float64), r.pcm.T, NAMES[k])


@pytest.mark.parametrize("fmt", ["f32", "s16"])
@pytest.mark.parametrize("pipe", ["fused", "staged"])
def test_poisoned_buffers_alternating_batches(pipe, fmt, batches):
    import mp3_b200 as m
    streams, refs, orders = batches
    with m.Decoder(device=0, pcm_format=m.PCM_F32 if fmt == "f32" else m.PCM_S16,
                   pipeline=m.PIPE_FUSED if pipe == "fused" else m.PIPE_STAGED) as dec:
        for order in orders:
            dec.decode_batch([streams[k] for k in order])
            arena = dec.fetch_pcm()
            if fmt == "f32":
                assert np.all(np.isfinite(arena)), "an element of the PCM arena was never written"
            _check_pcm(dec, arena, order, refs, fmt)


def test_poisoned_huffman_output_is_bit_exact(batches):
    import mp3_b200 as m
    streams, refs, orders = batches
    with m.Decoder(device=0, pcm_format=m.PCM_F32, keep_stages=True) as dec:
        for order in orders[:3]:
            dec.decode_batch([streams[k] for k in order])
            is_, sf = dec.stage(m.STAGE_IS), dec.stage(m.STAGE_SF)
            for i, k in enumerate(order):
                if ALL[NAMES[k]].get("layer", 3) != 3:
                    continue
                ub = dec.stream_info(i).pcm_offset // 576
                r = refs[k]
                assert np.array_equal(is_[ub: ub + r.units], r.is_), NAMES[k]
                assert np.array_equal(sf[ub: ub + r.units], r.sf), NAMES[k]


def test_poisoned_waves_sink_and_tiles(batches, monkeypatch):
    """Small waves (the wave-sized scratch is reused inside one call), short tiles (state re-derived from warm-up
    granules) and the PCM sink (two alternating arenas)."""
    import mp3_b200 as m
    streams, refs, orders = batches
    monkeypatch.setenv("MP3B_WAVE_UNITS", "150")
    monkeypatch.setenv("MP3B_FUSED_TILE", "3")
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        cap = sum(r.pcm.size for r in refs) + 576 * 2 * len(refs)
        sink = m.PinnedBuffer(cap * 2 + 64)
        dec.set_pcm_sink(sink.ptr, cap)
        for order in orders:
            dec.decode_batch([streams[k] for k in order])
            n = dec.pcm_device()[1]
            _check_pcm(dec, sink.view(np.int16, n).copy(), order, refs, "s16")
        dec.set_pcm_sink(0, 0)


def test_poisoned_incremental_streams(batches):
    import mp3_b200 as m
    streams, refs, _ = batches
    pick = list(range(0, len(streams), 4))
    rng = np.random.default_rng(5)
    with m.Decoder(device=0, pcm_format=m.PCM_F32) as dec:
        hs = [dec.open_stream() for _ in pick]
        pos = [0] * len(pick)
        got = [[] for _ in pick]
        while any(pos[j] < len(streams[k]) for j, k in enumerate(pick)):
            for j, k in enumerate(pick):
                n = int(rng.integers(1, 3000))
                hs[j].enqueue(streams[k][pos[j]: pos[j] + n])
                pos[j] += n
            dec.decode_streams()
            for j in range(len(pick)):
                inf = hs[j].info()
                if inf.samples:
                    got[j].append(hs[j].fetch(inf.samples))
        for j, k in enumerate(pick):
            cat = np.concatenate(got[j]).astype(np.float64)
            assert cat.shape == refs[k].pcm.T.shape, NAMES[k]
            if NAMES[k] == "loud_clipping":
                assert np.abs(cat - refs[k].pcm.T).max() / np.abs(refs[k].pcm).max() < 1e-5
            else:
                l3util.assert_iso_full_accuracy(cat, refs[k].pcm.T, NAMES[k])
        for h in hs:
            h.close()
