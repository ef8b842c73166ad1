"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the CPU
oracle on the same generated streams.

  * frame index, scalefactors and Huffman output: bit-exact
  * requantised spectrum, subband samples: float32 vs the oracle's double, relative to the block scale
  * PCM: ISO/IEC 11172-4 full accuracy (rms < 2^-15/sqrt(12), max |err| <= 2^-14, full scale 1.0);
    s16 output: |diff| <= 1 LSB against round(oracle * 32768)
"""
import os

import numpy as np
import pytest

import cases
import l3util

# the time-parallel frame walk is four kernels instead of one (these batches take the serial walk unless it is forced)
WALK_EXTRA = 3 if os.environ.get("MP3B_WALK") == "par" else 0

pytestmark = pytest.mark.gpu

ALL = dict(cases.FF)
ALL.update(cases.EXTRA)
NAMES = sorted(ALL)


@pytest.fixture(scope="module")
def mp3b():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


@pytest.fixture(scope="module")
def batch(mp3b, synth_mod, oracle_mod):
    streams = [synth_mod.make_stream(**ALL[n]) for n in NAMES]
    refs = [oracle_mod.decode(s, dumps=True) for s in streams]
    return streams, refs


@pytest.fixture(scope="module", params=["device_index", "host_index"])
def decoded(request, mp3b, batch):
    streams, refs = batch
    idx = mp3b.INDEX_DEVICE if request.param == "device_index" else mp3b.INDEX_HOST
    dec = mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, indexer=idx, pipeline=mp3b.PIPE_STAGED, keep_stages=True)
    dec.decode_batch(streams)
    out = dict(dec=dec, pcm=dec.fetch_pcm(), is_=dec.stage(mp3b.STAGE_IS), sf=dec.stage(mp3b.STAGE_SF),
               xr=dec.stage(mp3b.STAGE_XR), sb=dec.stage(mp3b.STAGE_SB), frames=dec.stage(mp3b.STAGE_FRAMES))
    yield out
    dec.close()


def _ubase(dec, i):
    inf = dec.stream_info(i)
    return inf, inf.pcm_offset // 576


def test_stream_info_and_frame_table(decoded, batch):
    streams, refs = batch
    dec = decoded["dec"]
    fbase = 0
    for i, (s, r) in enumerate(zip(streams, refs)):
        inf = dec.stream_info(i)
        assert (inf.sample_rate, inf.channels, inf.frames, inf.samples) == (r.sample_rate, r.channels, r.frames, r.samples)
        fr = decoded["frames"][fbase: fbase + r.frames]
        py = l3util.split_frames(s)
        offs = np.cumsum([0] + [len(f) for f in py[:-1]])
        assert np.array_equal(fr[:, 0], offs.astype(np.uint32)), NAMES[i]
        assert np.all(fr[:, 3] == i)
        fbase += r.frames
    assert dec.stats().concealed_frames == sum(r.concealed_frames for r in refs)


@pytest.mark.parametrize("k", range(len(NAMES)))
def test_huffman_and_scalefactors_bit_exact(k, decoded, batch):
    _, refs = batch
    inf, ub = _ubase(decoded["dec"], k)
    r = refs[k]
    assert np.array_equal(decoded["sf"][ub: ub + r.units], r.sf), NAMES[k]
    assert np.array_equal(decoded["is_"][ub: ub + r.units], r.is_), NAMES[k]


@pytest.mark.parametrize("mode,env", [("chunk", {}), ("warp", {}), ("sorted", {"MP3B_K1_WARPS": "3"}), ("sorted", {"MP3B_K1_WARPS": "32"}),
                                      ("sorted", {"MP3B_K1_REGION": "64"})])
def test_huffman_variants_bit_exact(mode, env, mp3b, batch, monkeypatch):
    """Both Huffman kernels -- the default chunked one (a CTA sorts its chunk) and the sorted one (blocks of 2,048 units
    ordered by length, persistent warps pull groups of 32; two passes), the latter also with few warps and with a
    shared-memory part too small for the groups (register-window reader over global memory) -- give the oracle's
    integers."""
    streams, refs = batch
    monkeypatch.setenv("MP3B_K1_MODE", mode)
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, pipeline=mp3b.PIPE_FUSED, keep_stages=True) as dec:
        dec.decode_batch(streams)
        is_, sf = dec.stage(mp3b.STAGE_IS), dec.stage(mp3b.STAGE_SF)
        assert dec.stats().kernel_launches == (10 if mode == "sorted" else 7) + WALK_EXTRA
        for k, r in enumerate(refs):
            ub = dec.stream_info(k).pcm_offset // 576
            assert np.array_equal(sf[ub: ub + r.units], r.sf), NAMES[k]
            assert np.array_equal(is_[ub: ub + r.units], r.is_), NAMES[k]


@pytest.mark.parametrize("k", range(len(NAMES)))
def test_spectrum_and_subbands(k, decoded, batch):
    _, refs = batch
    inf, ub = _ubase(decoded["dec"], k)
    r = refs[k]
    for name, got, ref in (("xr", decoded["xr"][ub: ub + r.units], r.xr),
                           ("sb", decoded["sb"][ub: ub + r.units].reshape(r.units, 576), r.sb.reshape(r.units, 576))):
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-30)
        err = np.abs(got.astype(np.float64) - ref) / scale
        assert err.max() < 2e-5, (NAMES[k], name, err.max())


@pytest.mark.parametrize("k", range(len(NAMES)))
def test_pcm_iso_full_accuracy(k, decoded, batch):
    _, refs = batch
    dec = decoded["dec"]
    got = dec.stream_pcm(k, decoded["pcm"]).astype(np.float64)
    ref = refs[k].pcm.T
    assert got.shape == ref.shape
    if NAMES[k] == "loud_clipping":
        # far above full scale: the absolute ISO bound is meaningless, check relative error instead
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
    else:
        l3util.assert_iso_full_accuracy(got, ref, NAMES[k])


@pytest.mark.parametrize("tile", [0, 1, 3, 5, 16])
def test_fused_pipeline_pcm(tile, mp3b, batch, monkeypatch):
    """The fused back end (default pipeline) against the oracle, for several tile lengths: tile
    boundaries re-derive the overlap / synthesis state from two warm-up granules."""
    streams, refs = batch
    if tile:
        monkeypatch.setenv("MP3B_FUSED_TILE", str(tile))
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, pipeline=mp3b.PIPE_FUSED) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        assert dec.stats().kernel_launches == 7 + WALK_EXTRA  # walk, publish, side_parse, payload_copy, huffman, backend, publish
        for k, r in enumerate(refs):
            got = dec.stream_pcm(k, arena).astype(np.float64)
            ref = r.pcm.T
            assert got.shape == ref.shape
            if NAMES[k] == "loud_clipping":
                assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
            else:
                l3util.assert_iso_full_accuracy(got, ref, "%s tile=%d" % (NAMES[k], tile))


def test_fused_equals_staged_s16(mp3b, batch):
    streams, _ = batch
    out = []
    for pipe in (mp3b.PIPE_STAGED, mp3b.PIPE_FUSED):
        with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, pipeline=pipe) as dec:
            dec.decode_batch(streams)
            out.append(dec.fetch_pcm().astype(np.int32))
    assert out[0].shape == out[1].shape
    assert np.abs(out[0] - out[1]).max() <= 1


@pytest.mark.parametrize("pipe", ["staged", "fused"])
def test_s16_output_rounding_and_saturation(pipe, mp3b, batch):
    streams, refs = batch
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16,
                      pipeline=mp3b.PIPE_STAGED if pipe == "staged" else mp3b.PIPE_FUSED) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        clipped = 0
        for k, r in enumerate(refs):
            got = dec.stream_pcm(k, arena).astype(np.int64)
            ref = np.clip(np.rint(r.pcm.T * 32768.0), -32768, 32767).astype(np.int64)
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 1, NAMES[k]
            clipped += int(np.count_nonzero(np.abs(got) >= 32767))
        assert clipped > 0, "the loud case should exercise saturation"


def test_stream_interface_matches_batch(mp3b, batch):
    streams, refs = batch
    pick = [0, 3, 9]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        hs = [dec.open_stream() for _ in pick]
        for h, k in zip(hs, pick):
            s = streams[k]
            h.enqueue(s[: len(s) // 3])
            h.enqueue(s[len(s) // 3:])
        dec.decode_streams()
        for h, k in zip(hs, pick):
            inf = h.info()
            assert inf.samples == refs[k].samples
            a = h.fetch(1000)
            b = h.fetch(inf.samples)
            got = np.concatenate([a, b]).astype(np.float64)
            assert got.shape[0] == inf.samples
            assert h.fetch(10).shape[0] == 0
            l3util.assert_iso_full_accuracy(got, refs[k].pcm.T, NAMES[k])
        for h in hs:
            h.close()


def test_garbage_id3_and_truncation(mp3b, synth_mod, oracle_mod):
    s = synth_mod.make_stream(nframes=10, seed=77, blocks=1)
    id3 = b"ID3\x03\x00\x00" + bytes([0, 0, 1, 10]) + bytes(138)
    variants = [id3 + s, b"\x00\xff\x12junk" * 7 + s, s[:-100], s[5:], b"", b"\xff" * 50, s[: 4], s + s[:300]]
    refs = [oracle_mod.decode(v) for v in variants]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        for idx in (mp3b.INDEX_DEVICE,):
            dec.decode_batch(variants)
            arena = dec.fetch_pcm()
            for k, r in enumerate(refs):
                inf = dec.stream_info(k)
                assert inf.frames == r.frames, k
                if r.frames:
                    got = dec.stream_pcm(k, arena).astype(np.float64)
                    l3util.assert_iso_full_accuracy(got, r.pcm.T, "variant %d" % k)
            assert dec.stats().concealed_frames == sum(r.concealed_frames for r in refs)


@pytest.mark.parametrize("pipe", ["staged", "fused"])
def test_waves_give_identical_pcm(pipe, mp3b, batch, monkeypatch):
    streams, _ = batch
    pl = mp3b.PIPE_STAGED if pipe == "staged" else mp3b.PIPE_FUSED
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, pipeline=pl) as dec:
        dec.decode_batch(streams)
        one = dec.fetch_pcm().copy()
    monkeypatch.setenv("MP3B_WAVE_UNITS", "150")
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, pipeline=pl) as dec:
        dec.decode_batch(streams)
        many = dec.fetch_pcm().copy()
    assert np.array_equal(one, many)


def test_pcm_sink_matches_fetch(mp3b, batch):
    """The overlapped D2H path (mp3b_set_pcm_sink) delivers the same bytes as an explicit fetch."""
    streams, _ = batch
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        dec.decode_batch(streams)
        ref = dec.fetch_pcm().copy()
        sink = mp3b.PinnedBuffer(ref.nbytes + 64)
        sink.view(np.uint8)[:] = 0xEE
        dec.set_pcm_sink(sink.ptr, ref.size)
        for _ in range(2):                    # twice: the second call must wait for the first copies
            dec.decode_batch(streams)
        got = sink.view(np.int16, ref.size).copy()
        dec.set_pcm_sink(0, 0)
        assert np.array_equal(got, ref)
        small = mp3b.PinnedBuffer(1024)
        dec.set_pcm_sink(small.ptr, 10)
        with pytest.raises(mp3b.Mp3bError) as e:
            dec.decode_batch(streams)
        assert e.value.status == -3
        dec.set_pcm_sink(0, 0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_incremental_stream_equals_one_shot(seed, mp3b, batch):
    """open / enqueue (arbitrary piece sizes) / decode / fetch, many times: the concatenated PCM is
    bit-identical to a one-shot decode, for every stream family (reservoir, LSF, short blocks ...)."""
    streams, refs = batch
    rng = np.random.default_rng(seed)
    pick = list(range(0, len(streams), 5))[:12]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch([streams[k] for k in pick])
        arena = dec.fetch_pcm()
        whole = [dec.stream_pcm(i, arena).copy() for i in range(len(pick))]
        hs = [dec.open_stream() for _ in pick]
        pos = [0] * len(pick)
        got = [[] for _ in pick]
        for step in range(400):
            alive = False
            for j, k in enumerate(pick):
                s = streams[k]
                if pos[j] < len(s):
                    n = int(rng.integers(1, 4000)) if step % 3 else int(rng.integers(1, 200))
                    hs[j].enqueue(s[pos[j]: pos[j] + n])
                    pos[j] += n
                    alive = True
            dec.decode_streams()
            for j in range(len(pick)):
                inf = hs[j].info()
                if inf.samples:
                    a = hs[j].fetch(inf.samples)
                    assert a.shape[0] == inf.samples
                    got[j].append(a)
            if not alive:
                break
        for j, k in enumerate(pick):
            cat = np.concatenate(got[j]) if got[j] else np.zeros((0, whole[j].shape[1]), np.float32)
            assert cat.shape == whole[j].shape, (NAMES[k], cat.shape, whole[j].shape)
            assert np.array_equal(cat, whole[j]), NAMES[k]
            assert hs[j].info().total_samples == whole[j].shape[0]
        for h in hs:
            h.close()


def test_run_ahead_walk_back_to_back_calls(mp3b, batch):
    """opts.async_index: the frame walk of call N+1 runs on a private stream while call N's kernels are
    still queued, on double-buffered tables.  A large call X followed immediately (no synchronisation)
    by a small call Y into the same PCM sink: the head of the sink must hold Y's PCM and the tail, which
    Y does not reach, X's -- decoded while Y's walk ran ahead -- both equal to the stream-ordered mode."""
    streams, _ = batch
    big = [streams, streams[::-1], streams[1:] + streams[:1]]
    small = [streams[:3], streams[5:7], streams[-2:]]
    want = {}
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, async_index=False) as dec:
        for k, st in enumerate(big + small):
            dec.decode_batch(st)
            want[k] = dec.fetch_pcm().copy()

    def pinned(st):
        packed, offs = mp3b.pack_streams(st)
        b = mp3b.PinnedBuffer(packed.size + 64)
        b.view(np.uint8)[:packed.size] = packed
        return b, offs

    ins = [pinned(st) for st in big + small]
    cap = max(w.size for w in want.values())
    sink = mp3b.PinnedBuffer(cap * 2 + 64)
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16, async_index=True) as dec:
        dec.set_pcm_sink(sink.ptr, cap)
        for rep in range(2):
            for x in range(len(big)):
                for y in range(len(big), len(big) + len(small)):
                    dec.decode_packed(ins[x][0].ptr, ins[x][1], where=mp3b.HOST, sync=False)
                    dec.decode_packed(ins[y][0].ptr, ins[y][1], where=mp3b.HOST, sync=False)
                    dec.sync()
                    got = sink.view(np.int16, cap)
                    ny, nx = want[y].size, want[x].size
                    assert ny < nx
                    assert np.array_equal(got[:ny], want[y]), (x, y, "head")
                    assert np.array_equal(got[ny:nx], want[x][ny:]), (x, y, "tail")
        dec.set_pcm_sink(0, 0)


def test_two_contexts_on_two_threads(mp3b, batch):
    """mp3_b200.multi.MultiDecoder: one context per entry, each driven by its own thread (here twice the
    same GPU; with several GPUs the same code shards a batch) -- same PCM as one context."""
    from mp3_b200 import multi
    streams, _ = batch
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        want = [dec.stream_pcm(i, arena).copy() for i in range(len(streams))]
    md = multi.MultiDecoder([0, 0], pcm_format=mp3b.PCM_S16)
    try:
        for _ in range(2):
            got = md.decode(streams)
            assert len(got) == len(want)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
    finally:
        md.close()
