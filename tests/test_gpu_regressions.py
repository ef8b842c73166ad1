"""Regression tests for defects found by review (round-1 ADVICE.md), each through the C-ABI on the GPU:

  * an odd-length mono stream in front of a stereo stream must not misalign the stereo stream's output in the
    stretched / resampled arenas (k_stretch / k_resample store a stereo frame as one word);
  * closing one stream between mp3b_decode() and the fetch of another must not drop the other's PCM;
  * an ID3v2 tag that arrives in pieces (its body full of bytes that look like syncs) and a first frame whose
    confirming next header has not arrived yet must not fix a false stream identity;
  * Layer II streams at both ends of a Layer III batch (the dense subband-sample buffer).
"""
import numpy as np
import pytest

import l3util
from oracle import wsola

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mp3b():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


@pytest.mark.parametrize("fmt", ["s16", "f32"])
@pytest.mark.parametrize("speed", [(5, 4), (7, 10), (3, 4), (1, 1)])
def test_odd_mono_stream_in_front_of_a_stereo_stream(fmt, speed, mp3b, synth_mod):
    num, den = speed
    # 2 mono frames at 5/4 = 1843 samples (odd); the gapless window of the tagged mono stream is odd at every speed
    cfgs = [dict(nframes=2, seed=1, mode=3),
            dict(nframes=6, seed=2, mode=1),
            dict(nframes=5, seed=3, mode=3, tag=2, tag_lame=1, enc_delay=577, enc_padding=1000),
            dict(nframes=6, seed=4),
            dict(nframes=3, seed=5, mode=3, sample_rate=22050, bitrate_kbps=32),
            dict(nframes=8, seed=6, sample_rate=22050, bitrate_kbps=64)]
    streams = [synth_mod.make_stream(**c) for c in cfgs]
    f = mp3b.PCM_F32 if fmt == "f32" else mp3b.PCM_S16
    with mp3b.Decoder(device=0, pcm_format=f, gapless=True) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm().copy()
        dec.time_stretch(num, den)
        out, where = dec.fetch_stretched()
        odd = 0
        for i in range(len(cfgs)):
            inf = dec.stream_info(i)
            pcm = dec.stream_pcm(i, arena)
            x = pcm.astype(np.float64) / (32768.0 if fmt == "s16" else 1.0)
            s16 = pcm.astype(np.int64) if fmt == "s16" else wsola.to_s16(pcm)
            ref, _ = wsola.wsola(x, s16, inf.sample_rate, num, den)
            off, cnt = where[i]
            assert off % 8 == 0 and cnt == ref.shape[0]
            odd += (cnt * inf.channels) & 1
            got = out[off: off + cnt * inf.channels].reshape(cnt, inf.channels).astype(np.float64)
            if fmt == "f32":
                assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
            else:
                assert np.abs(got - np.clip(np.rint(ref * 32768.0), -32768, 32767)).max() <= 1
        assert odd > 0, "the batch must hold an odd-length stream in front of a stereo one"
        # the same layout rule in the resampled arena
        dec.resample(48000)
        out, where = dec.fetch_resampled()
        for i, (off, cnt) in enumerate(where):
            assert off % 8 == 0
            inf = dec.stream_info(i)
            assert cnt == -(-inf.samples * 48000 // inf.sample_rate)
        dec.decode_batch(streams[:1])  # the context survived (a misaligned store is a sticky CUDA error)
        assert dec.stream_info(0).frames == 2


def test_closing_one_stream_keeps_the_others_pcm(mp3b, synth_mod):
    a = synth_mod.make_stream(nframes=10, seed=11, blocks=1)
    b = synth_mod.make_stream(nframes=12, seed=12, mode=1, bitrate_kbps=192, blocks=1)
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch([b])
        whole = dec.stream_pcm(0).copy()
        ha, hb = dec.open_stream(), dec.open_stream()
        got = []
        cut = len(b) // 2
        ha.enqueue(a)
        hb.enqueue(b[:cut])
        dec.decode_streams()
        ha.close()                       # between the decode and the fetch of stream B
        inf = hb.info()
        assert inf.samples > 0
        got.append(hb.fetch(inf.samples))
        hb.enqueue(b[cut:])
        dec.decode_streams()
        inf = hb.info()
        got.append(hb.fetch(inf.samples))
        cat = np.concatenate(got)
        assert cat.shape == whole.shape and np.array_equal(cat, whole)
        hb.close()


def test_id3_tag_arriving_in_pieces_with_false_syncs(mp3b, synth_mod):
    s = synth_mod.make_stream(nframes=12, seed=21, sample_rate=22050, bitrate_kbps=64, blocks=1)
    # a tag body made of headers of a DIFFERENT stream family (MPEG-1 44.1 kHz stereo 128k), chained at the right
    # distance so that each is "confirmed" by the next: a walk that looked inside would lock onto them
    fake = (bytes([0xFF, 0xFB, 0x90, 0x00]) + bytes(413)) * 8
    n = len(fake)
    id3 = b"ID3\x04\x00\x00" + bytes([(n >> 21) & 127, (n >> 14) & 127, (n >> 7) & 127, n & 127]) + fake
    data = id3 + s
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch([s])
        whole = dec.stream_pcm(0).copy()
        for piece in (7, 300, 1000):
            h = dec.open_stream()
            got = []
            for p in range(0, len(data), piece):
                h.enqueue(data[p: p + piece])
                dec.decode_streams()
                inf = h.info()
                if inf.samples:
                    assert inf.sample_rate == 22050
                    got.append(h.fetch(inf.samples))
            cat = np.concatenate(got)
            assert cat.shape == whole.shape and np.array_equal(cat, whole), piece
            h.close()


def test_first_frame_waits_for_its_confirming_header(mp3b, synth_mod):
    s = synth_mod.make_stream(nframes=6, seed=22, blocks=1)
    frames = l3util.split_frames(s)
    # junk that holds one valid-looking 48 kHz mono header whose "frame" ends exactly where the enqueued bytes end
    junk = bytes([0xFF, 0xFB, 0x94, 0xC0]) + bytes(384 - 4)
    assert l3util.frame_len(junk[:4]) == 384
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch([junk + s])
        inf1 = dec.stream_info(0)
        whole = dec.stream_pcm(0).copy()
        assert inf1.sample_rate == 44100 and inf1.channels == 2 and inf1.frames == len(frames)
        h = dec.open_stream()
        h.enqueue(junk)                  # the false candidate is complete, its look-ahead is not there yet
        dec.decode_streams()
        assert h.info().samples == 0
        got = []
        h.enqueue(s)
        dec.decode_streams()
        inf = h.info()
        assert inf.sample_rate == 44100 and inf.channels == 2
        got.append(h.fetch(inf.samples))
        cat = np.concatenate(got)
        assert cat.shape == whole.shape and np.array_equal(cat, whole)
        h.close()


@pytest.mark.parametrize("pipe", ["fused", "staged"])
def test_layer2_streams_at_both_ends_of_a_layer3_batch(pipe, mp3b, synth_mod, oracle_mod):
    l2a = synth_mod.make_stream(layer=2, bitrate_kbps=192, nframes=6, seed=31)
    l2b = synth_mod.make_stream(layer=2, bitrate_kbps=64, mode=3, nframes=7, seed=32)
    l1 = synth_mod.make_stream(layer=1, bitrate_kbps=384, nframes=11, seed=33)
    l3 = [synth_mod.make_stream(nframes=9, seed=40 + k, blocks=1, mode=k % 4) for k in range(6)]
    streams = [l2a] + l3[:3] + [l1] + l3[3:] + [l2b]
    pl = mp3b.PIPE_STAGED if pipe == "staged" else mp3b.PIPE_FUSED
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, pipeline=pl) as dec:
        for _ in range(2):  # twice: the second call reuses the tile table behind its event
            dec.decode_batch(streams)
            arena = dec.fetch_pcm()
            for i, s in enumerate(streams):
                ref = oracle_mod.decode(s)
                got = dec.stream_pcm(i, arena).astype(np.float64)
                assert got.shape == ref.pcm.T.shape
                l3util.assert_iso_full_accuracy(got, ref.pcm.T, "stream %d" % i)


def test_bad_options_are_rejected(mp3b):
    import ctypes
    L = mp3b.load_library()
    for field, val in (("pcm_format", 2), ("pcm_format", -1), ("indexer", 5), ("pipeline", 9), ("host_threads", -3)):
        o = mp3b.Opts()
        L.mp3b_opts_default(ctypes.byref(o))
        setattr(o, field, val)
        ctx = ctypes.c_void_p()
        assert L.mp3b_ctx_create(0, ctypes.byref(o), ctypes.byref(ctx)) == -1, field
        assert not ctx.value
