"""The time-parallel frame walk (k_walk_first / _segments / _stitch / _compact: the bytes cut into segments that are walked
speculatively, a thread each, and then stitched per stream) must leave exactly the table of the serial walk (k_index_walk, one thread per
stream): frame records, stream records, tags, concealment, PCM -- on healthy streams of every kind, on damaged ones
(where guesses are wrong and segments are repaired), with segments shorter than a frame (the chain skips segments),
through the incremental interface, and on one long stream."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

ALL = dict(cases.FF)
ALL.update(cases.EXTRA)
ALL.update(cases.L2)
ALL.update(cases.L1)


@pytest.fixture(scope="module")
def streams(synth_mod):
    rng = np.random.default_rng(4242)
    good = [synth_mod.make_stream(**ALL[n]) for n in sorted(ALL)]
    out = list(good)
    id3 = b"ID3\x03\x00\x00" + bytes([0, 0, 3, 10]) + bytes(394)
    for s in good[::3]:
        a = np.frombuffer(s, np.uint8).copy()
        for n in (1, 4, 20, 120):                       # bit flips: headers, side info, main data
            b = a.copy()
            idx = rng.integers(0, b.size, n)
            b[idx] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
            out.append(b.tobytes())
        out.append(a[: rng.integers(1, a.size)].tobytes())   # truncated
        out.append(a[rng.integers(1, 700):].tobytes())        # lost head
        out.append(id3 + s)                                   # tag in front
        out.append(s[:900] + bytes(rng.integers(0, 256, 1500, dtype=np.uint8)) + s[900:])   # junk inside
    long_ = [synth_mod.make_stream(nframes=300, seed=9, mode=1, bitrate_kbps=320, blocks=1, mixed_pct=25, fill_lo_pct=35),
             synth_mod.make_stream(nframes=400, seed=10, vbr_min_kbps=32, vbr_max_kbps=320, blocks=1)]
    # wrong guesses by construction: false chain entries inside the main data, every few segments
    out += [plant_false_entries(long_[0]), plant_false_entries(long_[1], every=1700, start=900)]
    out += [b"", b"\xff" * 40, bytes(rng.integers(0, 256, 5000, dtype=np.uint8)), bytes([0xFF, 0xFB, 0x90, 0x00]) * 300]
    return out


from test_walk_par_cpu import plant_false_entries  # noqa: E402


def _decode(mp3b, streams, monkeypatch, mode, seg=None):
    monkeypatch.setenv("MP3B_WALK", mode)
    if seg:
        monkeypatch.setenv("MP3B_WALK_SEG", str(seg))
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, keep_stages=True) as dec:
        dec.decode_batch(streams)
        pcm = dec.fetch_pcm().copy()
        frames = dec.stage(mp3b.STAGE_FRAMES).copy()
        infos = [(i.sample_rate, i.channels, i.frames, i.samples, i.pcm_offset)
                 for i in (dec.stream_info(k) for k in range(len(streams)))]
        tags = [(t.kind, t.has_lame, t.frames, t.bytes, t.first_sample, t.num_samples)
                for t in (dec.tag_info(k) for k in range(len(streams)))]
        conc = dec.stats().concealed_frames
    return frames, infos, tags, conc, pcm


@pytest.fixture(scope="module")
def mp3b():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


@pytest.fixture(scope="module")
def serial(mp3b, streams):
    mp = pytest.MonkeyPatch()
    try:
        return _decode(mp3b, streams, mp, "serial")
    finally:
        mp.undo()


@pytest.mark.parametrize("seg", [48, 240, 1032, 4104, 65520])
def test_parallel_walk_equals_serial_walk(seg, mp3b, streams, serial, monkeypatch):
    frames, infos, tags, conc, pcm = _decode(mp3b, streams, monkeypatch, "par", seg)
    assert infos == serial[1]
    assert np.array_equal(frames, serial[0])
    assert tags == serial[2] and conc == serial[3]
    assert np.array_equal(pcm, serial[4])


def test_parallel_walk_incremental(mp3b, streams, monkeypatch):
    pick = [s for s in streams if len(s) > 3000][:10]
    monkeypatch.setenv("MP3B_WALK", "serial")
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        dec.decode_batch(pick)
        arena = dec.fetch_pcm()
        whole = [dec.stream_pcm(i, arena).copy() for i in range(len(pick))]
    monkeypatch.setenv("MP3B_WALK", "par")
    monkeypatch.setenv("MP3B_WALK_SEG", "504")
    rng = np.random.default_rng(1)
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        hs = [dec.open_stream() for _ in pick]
        pos = [0] * len(pick)
        got = [[] for _ in pick]
        while any(pos[j] < len(s) for j, s in enumerate(pick)):
            for j, s in enumerate(pick):
                n = int(rng.integers(1, 2500))
                hs[j].enqueue(s[pos[j]: pos[j] + n])
                pos[j] += n
            dec.decode_streams()
            for j in range(len(pick)):
                inf = hs[j].info()
                if inf.samples:
                    got[j].append(hs[j].fetch(inf.samples))
        for j in range(len(pick)):
            # (a stream whose only frame could not be confirmed while streaming emits nothing; the batch decode does)
            cat = np.concatenate(got[j]) if got[j] else np.zeros((0, whole[j].shape[1]), np.int16)
            if cat.shape[0] or whole[j].shape[0] > 1152:
                assert cat.shape == whole[j].shape and np.array_equal(cat, whole[j]), j
        for h in hs:
            h.close()


def test_long_stream_is_walked_in_parallel(mp3b, synth_mod, monkeypatch):
    """A long stream takes the parallel walk by default (batch shape); same table and PCM as the serial walk, and the
    index stage is far shorter (the serial chain is one dependent load per frame)."""
    s = synth_mod.make_stream(nframes=20000, seed=5, mode=1, bitrate_kbps=128, blocks=1)   # 8.7 minutes
    res = {}
    for mode in ("serial", "auto"):
        monkeypatch.setenv("MP3B_WALK", mode)
        with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, keep_stages=True) as dec:
            for _ in range(4):
                dec.decode_batch([s])
            res[mode] = (dec.stage(mp3b.STAGE_FRAMES).copy(), dec.fetch_pcm().copy(), dec.stats().ms_index)
    assert res["serial"][0].shape[0] == 20000
    assert np.array_equal(res["serial"][0], res["auto"][0]) and np.array_equal(res["serial"][1], res["auto"][1])
    print("index stage: serial walk %.3f ms, parallel walk %.3f ms" % (res["serial"][2], res["auto"][2]))
    assert res["auto"][2] < 0.5 * res["serial"][2]
