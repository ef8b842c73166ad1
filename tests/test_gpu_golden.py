"""The committed golden vectors (tests/golden/ff_golden.npz: stream bytes + the PCM FFmpeg's mp3float / mp2float /
mp1float decoders produced for them, see tests/golden/make_golden.py) through the CUDA path directly: every
stream, both pipelines, both indexers, float and s16 output, the bulk and the incremental interface.  FFmpeg
computes in float32, as the kernels do, so the bound is the ISO/IEC 11172-4 full-accuracy criterion with a wide
margin; the s16 output must be within 1 LSB of round(golden * 32768)."""
import os

import numpy as np
import pytest

import l3util

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ff_golden.npz")


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLD)
    names = sorted(k[:-4] for k in z.files if k.endswith(".mp3"))
    assert len(names) >= 19
    return names, [z[n + ".mp3"].tobytes() for n in names], [z[n + ".pcm"].astype(np.float64).T for n in names]


@pytest.mark.parametrize("indexer", ["device", "host"])
@pytest.mark.parametrize("pipe", ["fused", "staged"])
def test_golden_float(pipe, indexer, golden):
    import mp3_b200 as m
    names, streams, refs = golden
    with m.Decoder(device=0, pcm_format=m.PCM_F32, pipeline=m.PIPE_FUSED if pipe == "fused" else m.PIPE_STAGED,
                   indexer=m.INDEX_DEVICE if indexer == "device" else m.INDEX_HOST) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        assert dec.stats().concealed_frames == 0
        for i, (n, ref) in enumerate(zip(names, refs)):
            got = dec.stream_pcm(i, arena).astype(np.float64)
            assert got.shape == ref.shape, n
            rms, mx = l3util.iso_compliance(got, ref)
            # two float32 implementations of the same arithmetic: far inside ISO (rms 8.8e-6, max 6.1e-5)
            assert rms < 2e-6 and mx < 3e-5, (n, rms, mx)


@pytest.mark.parametrize("pipe", ["fused", "staged"])
def test_golden_s16(pipe, golden):
    import mp3_b200 as m
    names, streams, refs = golden
    with m.Decoder(device=0, pcm_format=m.PCM_S16, pipeline=m.PIPE_FUSED if pipe == "fused" else m.PIPE_STAGED) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        for i, (n, ref) in enumerate(zip(names, refs)):
            got = dec.stream_pcm(i, arena).astype(np.int64)
            want = np.clip(np.rint(ref * 32768.0), -32768, 32767).astype(np.int64)
            assert got.shape == want.shape, n
            assert np.abs(got - want).max() <= 1, n


def test_golden_incremental(golden):
    import mp3_b200 as m
    names, streams, refs = golden
    with m.Decoder(device=0, pcm_format=m.PCM_F32) as dec:
        hs = [dec.open_stream() for _ in streams]
        got = [[] for _ in streams]
        for p in range(0, max(len(s) for s in streams), 777):
            for h, s in zip(hs, streams):
                if p < len(s):
                    h.enqueue(s[p: p + 777])
            dec.decode_streams()
            for j, h in enumerate(hs):
                inf = h.info()
                if inf.samples:
                    got[j].append(h.fetch(inf.samples))
        for j, (n, ref) in enumerate(zip(names, refs)):
            cat = np.concatenate(got[j]).astype(np.float64)
            assert cat.shape == ref.shape, n
            rms, mx = l3util.iso_compliance(cat, ref)
            assert rms < 2e-6 and mx < 3e-5, (n, rms, mx)
        for h in hs:
            h.close()
