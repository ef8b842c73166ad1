"""GPU: Layer II (MP2) and Layer I (MP1) streams through the C-ABI against the oracle (which FFmpeg's
mp2float / mp1float pin on the CPU side, tests/test_oracle_vs_ffmpeg.py), mixed with Layer III streams in
one batch."""
import numpy as np
import pytest

import cases
import l3util

pytestmark = pytest.mark.gpu

ALL = dict(cases.L2)
ALL.update(cases.L1)
NAMES = sorted(ALL)


@pytest.fixture(scope="module")
def batch(synth_mod, oracle_mod):
    streams = [synth_mod.make_stream(**ALL[n]) for n in NAMES]
    # Layer III neighbours in the same batch: the two paths share the unit / PCM layout
    streams.insert(2, synth_mod.make_stream(**cases.FF["cfg3_320k_joint"]))
    streams.append(synth_mod.make_stream(**cases.FF["lsf16_mono"]))
    refs = [oracle_mod.decode(s) for s in streams]
    return streams, refs


@pytest.mark.parametrize("pipe", ["fused", "staged"])
@pytest.mark.parametrize("indexer", ["device", "host"])
def test_layer2_pcm_iso_full_accuracy(pipe, indexer, batch):
    import mp3_b200 as m
    streams, refs = batch
    with m.Decoder(device=0, pcm_format=m.PCM_F32, pipeline=m.PIPE_FUSED if pipe == "fused" else m.PIPE_STAGED,
                   indexer=m.INDEX_DEVICE if indexer == "device" else m.INDEX_HOST) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        for i, r in enumerate(refs):
            inf = dec.stream_info(i)
            assert (inf.sample_rate, inf.channels, inf.frames, inf.samples) == (r.sample_rate, r.channels, r.frames, r.samples)
            got = dec.stream_pcm(i, arena).astype(np.float64)
            scale = max(1.0, float(np.abs(r.pcm).max()))
            l3util.assert_iso_full_accuracy(got / scale, r.pcm.T / scale, "stream %d" % i)


def test_layer2_s16_and_sink(batch):
    import mp3_b200 as m
    streams, refs = batch
    with m.Decoder(device=0, pcm_format=m.PCM_S16) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm().copy()
        for i, r in enumerate(refs):
            want = np.clip(np.rint(r.pcm.T * 32768.0), -32768, 32767)
            assert np.abs(dec.stream_pcm(i, arena).astype(np.float64) - want).max() <= 1
        sink = m.PinnedBuffer(arena.nbytes + 64)
        dec.set_pcm_sink(sink.ptr, arena.size)
        dec.decode_batch(streams)
        dec.sync()
        assert np.array_equal(sink.view(np.int16, arena.size), arena)
        dec.set_pcm_sink(0, 0)


def test_layer2_incremental_stream(batch):
    """open / enqueue in pieces / decode / fetch: Layer II frames are self-contained, only the synthesis
    history has to be carried (two granules replayed)."""
    import mp3_b200 as m
    streams, refs = batch
    rng = np.random.default_rng(3)
    for k in (0, 1, 4, 5, 12):  # Layer I (384-sample frames), Layer II, and a Layer III neighbour
        s, r = streams[k], refs[k]
        with m.Decoder(device=0, pcm_format=m.PCM_F32) as dec:
            dec.decode_batch([s])
            whole = dec.stream_pcm(0, dec.fetch_pcm()).copy()
            h = dec.open_stream()
            pos, got = 0, []
            while pos < len(s):
                n = int(rng.integers(50, 3000))
                h.enqueue(s[pos: pos + n])
                pos += n
                dec.decode_streams()
                inf = h.info()
                if inf.samples:
                    got.append(h.fetch(inf.samples))
            cat = np.concatenate(got)
            assert cat.shape == whole.shape and np.array_equal(cat, whole), k
            h.close()


def test_allocation_asking_for_more_bits_than_the_frame_holds(synth_mod, oracle_mod):
    """A damaged Layer I / II frame whose allocation field demands more sample bits than the frame has:
    bits past the frame's end read as zero in both implementations (found by tools/fuzz_parity.py)."""
    import mp3_b200 as m
    bad = []
    for cfg, fill in ((dict(layer=1, bitrate_kbps=256, nframes=12, seed=7), 0xEE),
                      (dict(layer=1, bitrate_kbps=64, sample_rate=48000, mode=3, nframes=12, seed=8), 0xDD),
                      (dict(layer=2, bitrate_kbps=192, nframes=8, seed=5), 0xFF),
                      (dict(layer=2, bitrate_kbps=32, sample_rate=32000, mode=3, nframes=8, seed=6), 0xFF),
                      (dict(layer=2, bitrate_kbps=160, sample_rate=24000, nframes=8, seed=9), 0xFF)):
        s = synth_mod.make_stream(**cfg)
        frames = l3util.split_frames(s)
        b = bytearray(s)
        for k in (1, len(frames) - 1):  # a middle frame and the stream's last one
            off = sum(len(f) for f in frames[:k])
            for i in range(off + 4, off + 4 + 32):
                b[i] = fill
        bad.append(bytes(b))
    refs = [oracle_mod.decode(s) for s in bad]
    with m.Decoder(device=0, pcm_format=m.PCM_F32) as dec:
        dec.decode_batch(bad)
        arena = dec.fetch_pcm()
        for i, r in enumerate(refs):
            inf = dec.stream_info(i)
            assert (inf.frames, inf.samples) == (r.frames, r.samples)
            got = dec.stream_pcm(i, arena).astype(np.float64)
            scale = max(1.0, float(np.abs(r.pcm).max()))
            l3util.assert_iso_full_accuracy(got / scale, r.pcm.T / scale, "stream %d" % i)
