"""Parity at BASELINE.json's full sizes, through size-independent properties and oracle samples.

  cfg2 (1024 x 383 frames), cfg3 (320 kbit/s joint stereo, switching windows, reservoir), cfg4 (LSF +
  VBR + mono/stereo mix): EVERY stream against the oracle over its full length -- Huffman output and
  scalefactors bit-exact, PCM under ISO 11172-4 --, every stream's length / rate / channel count;
  cfg5 (100,000 streams x 128 frames) at full size: shapes, replica equality over every stream, 64 oracle checks; and
  * replication: the same stream at different batch positions decodes to identical PCM;
  * linearity: lowering every global_gain by 4 halves the float PCM exactly (requantiser gain is a
    power of two, all later stages are linear);
  * corruption: bit flips and truncation never fault the device, sizes stay consistent.
"""
import numpy as np
import pytest

import l3util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mp3b():
    import mp3_b200
    mp3_b200.load_library()
    return mp3_b200


def _check_all(mp3b, oracle_mod, streams):
    """Every stream of the batch against the oracle: Huffman output and scalefactors bit-exact, PCM under
    ISO/IEC 11172-4.  One fused decode with keep_stages (the product kernels; the Huffman kernel then also writes
    the all-zero tails it normally skips, so that the dump is complete); the oracle runs on all host threads."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, keep_stages=True) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        is_ = dec.stage(mp3b.STAGE_IS)
        sf = dec.stage(mp3b.STAGE_SF)
        infos = [dec.stream_info(i) for i in range(len(streams))]
        st = dec.stats()
    distinct = {}
    for i, s in enumerate(streams):
        distinct.setdefault(s, []).append(i)

    def check(item):
        s, idx = item
        ref = oracle_mod.decode(s, dumps="int")
        want = ref.pcm.T
        worst = (0.0, 0.0)
        for i in idx:
            inf = infos[i]
            assert (inf.frames, inf.samples, inf.channels, inf.sample_rate) == (
                ref.frames, ref.samples, ref.channels, ref.sample_rate), i
            ub = inf.pcm_offset // 576
            assert np.array_equal(is_[ub: ub + ref.units], ref.is_), "Huffman output of stream %d" % i
            assert np.array_equal(sf[ub: ub + ref.units], ref.sf), "scalefactors of stream %d" % i
            got = arena[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].reshape(inf.samples, inf.channels)
            worst = max(worst, l3util.assert_iso_full_accuracy(got, want, "stream %d" % i))
        return worst

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:  # l3o_decode and numpy release the GIL
        worst = list(ex.map(check, distinct.items()))
    return infos, st, max(w[0] for w in worst), max(w[1] for w in worst)


@pytest.mark.parametrize("name,nstreams", [("cfg2", 1024), ("cfg3", 1024), ("cfg4", 1024)])
def test_full_workloads_every_stream_against_the_oracle(name, nstreams, mp3b, synth_mod, oracle_mod):
    streams = synth_mod.make_workload(name, nstreams)
    infos, st, rms, mx = _check_all(mp3b, oracle_mod, streams)
    print("%s: worst rms %.3g (limit %.3g), worst max %.3g (limit %.3g)" % (name, rms, l3util.RMS_LIMIT, mx, l3util.MAX_LIMIT))
    assert st.streams == nstreams and st.concealed_frames == 0
    for inf, s in zip(infos, streams):
        assert inf.frames == 383
    if name == "cfg2":
        assert st.units == 1568768 and all(i.channels == 2 and i.sample_rate == 44100 for i in infos)
    if name == "cfg4":
        assert {i.sample_rate for i in infos} == {22050, 24000, 44100}
        assert {i.channels for i in infos} == {1, 2}


def test_cfg5_full_size(mp3b, synth_mod, oracle_mod):
    """BASELINE.json configs[4] at its full size on one GPU: 100,000 streams x 128 frames (51.2 M units, 59 GB of s16
    PCM, 26 waves).  1,024 distinct streams repeat; every one of the 100,000 must report the right shape, every
    copy must equal the first copy of its stream bit for bit (compared on the device), and 64 streams spread over the
    batch are checked against the oracle."""
    torch = pytest.importorskip("torch")
    free, _ = torch.cuda.mem_get_info()
    if free < 100 * (1 << 30):
        pytest.skip("needs ~75 GB of device memory")
    N, D, NF = 100000, 1024, 128
    base = synth_mod.make_workload("cfg5", D, NF)
    assert len({len(s) for s in base}) == 1, "CBR streams of one rate: equal lengths (the device-side view relies on it)"
    packed1, _ = mp3b.pack_streams(base)
    slen = len(base[0])
    d1 = torch.from_numpy(packed1).cuda()
    reps = -(-N // D)
    d_raw = d1.repeat(reps)[: N * slen].contiguous()
    del d1
    offs = np.arange(N + 1, dtype=np.uint64) * np.uint64(slen)
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        dec.decode_packed(d_raw.data_ptr(), offs, where=mp3b.DEVICE)
        st = dec.stats()
        assert st.streams == N and st.frames == N * NF and st.units == N * NF * 4 and st.concealed_frames == 0
        per = NF * 1152 * 2
        for i in range(N):
            inf = dec.stream_info(i)
            assert (inf.frames, inf.samples, inf.channels, inf.sample_rate, inf.pcm_offset) == (
                NF, NF * 1152, 2, 44100, i * per), i
        ptr, n = dec.pcm_device()
        assert n == N * per

        class _View:  # the library's PCM arena as a torch tensor (zero copy)
            __cuda_array_interface__ = {"shape": (N, per), "typestr": "<i2", "data": (ptr, False), "version": 2}
        pcm = torch.as_tensor(_View(), device="cuda")
        first = pcm[:D]
        assert bool(first.any())
        for k in range(1, reps):
            blk = pcm[k * D: min((k + 1) * D, N)]
            assert torch.equal(blk, first[: blk.shape[0]]), "replica block %d differs" % k
        pick = sorted(set(int(x) for x in np.linspace(0, N - 1, 64)))
        for i in pick:
            ref = oracle_mod.decode(base[i % D])
            got = pcm[i].cpu().numpy().reshape(-1, 2).astype(np.int64)
            want = np.clip(np.rint(ref.pcm.T * 32768.0), -32768, 32767).astype(np.int64)
            assert got.shape == want.shape and np.abs(got - want).max() <= 1, i
        del pcm, first
    del d_raw
    torch.cuda.empty_cache()


def test_replicated_streams_decode_identically(mp3b, synth_mod):
    base = synth_mod.make_workload("cfg3", 8, 200)
    streams = [base[i % 8] for i in range(512)]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_S16) as dec:
        dec.decode_batch(streams)
        arena = dec.fetch_pcm()
        first = [dec.stream_pcm(i, arena).copy() for i in range(8)]
        for i in range(8, 512, 37):
            assert np.array_equal(dec.stream_pcm(i, arena), first[i % 8]), i


def _lower_gain(stream, delta):
    """Return a copy of an MPEG-1 stereo/mono Layer III stream with every global_gain lowered by delta."""
    out = bytearray(stream)
    p = 0
    for f in l3util.split_frames(stream):
        mono = (f[3] >> 6) == 3
        crc = 0 if (f[1] & 1) else 2
        base = (p + 4 + crc) * 8
        hdr_bits = 9 + (5 if mono else 3) + (4 if mono else 8)
        for u in range(2 if mono else 4):
            pos = base + hdr_bits + 59 * u + 21
            v = 0
            for k in range(8):
                v = (v << 1) | ((out[(pos + k) >> 3] >> (7 - ((pos + k) & 7))) & 1)
            v = max(v - delta, 0)
            for k in range(8):
                bit = (v >> (7 - k)) & 1
                idx, sh = (pos + k) >> 3, 7 - ((pos + k) & 7)
                out[idx] = (out[idx] & ~(1 << sh)) | (bit << sh)
        p += len(f)
    return bytes(out)


def test_gain_linearity_is_exact(mp3b, synth_mod):
    streams = synth_mod.make_workload("cfg3", 16, 100)
    half = [_lower_gain(s, 4) for s in streams]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch(streams)
        a = dec.fetch_pcm().copy()
        dec.decode_batch(half)
        b = dec.fetch_pcm().copy()
    assert a.shape == b.shape and np.abs(a).max() > 1e-3
    assert np.array_equal(b * 2.0, a)


def test_corrupted_streams_never_fault(mp3b, synth_mod):
    rng = np.random.default_rng(20261018)
    base = synth_mod.make_workload("cfg3", 8, 60) + synth_mod.make_workload("cfg4", 9, 60)
    bad = []
    for s in base:
        a = np.frombuffer(s, np.uint8).copy()
        for _ in range(6):                       # bit flips anywhere (headers, side info, main data)
            b = a.copy()
            idx = rng.integers(0, b.size, 40)
            b[idx] ^= (1 << rng.integers(0, 8, 40)).astype(np.uint8)
            bad.append(b.tobytes())
        bad.append(a[: rng.integers(1, a.size)].tobytes())          # truncation
        bad.append(a[rng.integers(1, 500):].tobytes())               # lost head
        bad.append(rng.integers(0, 256, 3000, dtype=np.uint8).tobytes())  # noise
        bad.append(bytes([0xFF, 0xFB, 0x90, 0x00]) * 200)            # headers only
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        for _ in range(2):
            dec.decode_batch(bad)
            arena = dec.fetch_pcm()
            assert np.all(np.isfinite(arena))
            tot = 0
            for i in range(len(bad)):  # streams back to back (a Layer I stream is padded to a whole granule)
                inf = dec.stream_info(i)
                assert tot <= inf.pcm_offset < tot + 576 * 2 or inf.frames == 0
                tot = max(tot, inf.pcm_offset + inf.samples * inf.channels)
            assert tot <= arena.size < tot + 576 * 2
        good = synth_mod.make_stream(nframes=8, seed=3)
        dec.decode_batch([good])               # the context is still healthy afterwards
        assert dec.stream_info(0).frames == 8


def test_damaged_streams_decode_like_the_oracle(mp3b, synth_mod, oracle_mod):
    """Parity does not stop at valid input: bit flips (headers, side info, main data, light and heavy),
    truncation and a lost head must give the oracle's frame count, sync decisions, concealment and PCM --
    the two implementations share the policy (sync confirmation, stream consistency, frames whose
    main_data_begin reaches before the stream are concealed, out-of-bits handling in Huffman)."""
    rng = np.random.default_rng(99)
    base = synth_mod.make_workload("cfg3", 5, 24) + synth_mod.make_workload("cfg4", 9, 24) + [
        synth_mod.make_stream(layer=2, bitrate_kbps=192, nframes=10, seed=5),
        synth_mod.make_stream(layer=2, sample_rate=24000, bitrate_kbps=64, mode=1, nframes=10, seed=6),
        synth_mod.make_stream(layer=1, bitrate_kbps=256, mode=1, nframes=20, seed=7)]
    bad = []
    for s in base:
        a = np.frombuffer(s, np.uint8).copy()
        for n in (1, 2, 3, 5, 8, 13, 40):
            b = a.copy()
            idx = rng.integers(0, b.size, n)
            b[idx] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
            bad.append(b.tobytes())
        bad.append(a[: rng.integers(1, a.size)].tobytes())
        bad.append(a[rng.integers(1, 500):].tobytes())
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch(bad)
        arena = dec.fetch_pcm()
        concealed = 0
        for i, s in enumerate(bad):
            r = oracle_mod.decode(s)
            inf = dec.stream_info(i)
            if r.rc != 0:
                assert inf.frames == 0, i
                continue
            concealed += r.concealed_frames
            assert (inf.frames, inf.samples, inf.channels, inf.sample_rate) == (r.frames, r.samples, r.channels, r.sample_rate), i
            got = dec.stream_pcm(i, arena).astype(np.float64)
            scale = max(1.0, float(np.abs(r.pcm).max()))
            assert np.abs(got - r.pcm.T).max() <= 2.0 ** -14 * scale, i
        assert dec.stats().concealed_frames == concealed


def test_side_info_that_overruns_the_stream(mp3b, synth_mod, oracle_mod):
    """Two damage patterns at the END of a stream, where the next stream's bytes follow in the arena
    (found by tools/fuzz_parity.py): (1) big_values inflated so that the last Huffman codes run into /
    across the unit's end -- the bits past the stream's end read as zero; (2) part2_3_length inflated so
    that the frame claims more main data than it can hold -- the frame is concealed.  Both in the middle
    of a stream too.  Each damaged stream is followed by a healthy one full of set bits."""
    cfgs = [dict(nframes=10, seed=21), dict(nframes=10, seed=22, mode=3, blocks=1),
            dict(nframes=10, seed=23, mode=1, bitrate_kbps=320, blocks=1, mixed_pct=25),
            dict(nframes=10, seed=24, sample_rate=22050, bitrate_kbps=64, mode=1, blocks=1),
            dict(nframes=10, seed=25, sample_rate=8000, bitrate_kbps=16, blocks=1, mode=3)]
    batch, expect_concealed = [], 0
    for c in cfgs:
        s = synth_mod.make_stream(**c)
        frames = l3util.split_frames(s)
        lsf = c.get("sample_rate", 44100) < 32000
        nunits = (1 if lsf else 2) * (1 if c.get("mode", 0) == 3 else 2)
        for what in ("big_values", "part2_3_length"):
            for where in (len(frames) - 1, 4):
                fr = list(frames)
                for u in range(nunits):
                    fr[where] = (l3util.patch_side_info(fr[where], u, big_values=288) if what == "big_values"
                                 else l3util.patch_side_info(fr[where], u, part2_3_length=4095))
                batch.append(b"".join(fr))
                batch.append(s)
    refs = [oracle_mod.decode(s) for s in batch]
    assert sum(r.concealed_frames for r in refs) >= len(cfgs) * 2  # the inflated lengths are concealed
    for fmt_pipe in (mp3b.PIPE_FUSED, mp3b.PIPE_STAGED):
        with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32, pipeline=fmt_pipe) as dec:
            dec.decode_batch(batch)
            arena = dec.fetch_pcm()
            for i, r in enumerate(refs):
                inf = dec.stream_info(i)
                assert (inf.frames, inf.samples) == (r.frames, r.samples), i
                got = dec.stream_pcm(i, arena).astype(np.float64)
                scale = max(1.0, float(np.abs(r.pcm).max()))
                assert np.abs(got - r.pcm.T).max() <= 2.0 ** -14 * scale, i
            assert dec.stats().concealed_frames == sum(r.concealed_frames for r in refs)


def test_one_long_stream(mp3b, synth_mod, oracle_mod):
    """A single 6,000-frame stream (2.6 minutes, joint stereo, mixed block types, bit reservoir in use): the
    frame chain is walked by one thread and the stream is cut into many tiles -- every sample against the
    oracle, both output formats, plus the sentence / planar / seek operators on it."""
    s = synth_mod.make_stream(nframes=6000, seed=77, mode=1, bitrate_kbps=192, blocks=1, mixed_pct=20, fill_lo_pct=40)
    ref = oracle_mod.decode(s)
    assert ref.frames == 6000 and ref.concealed_frames == 0
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        dec.decode_batch([s])
        got = dec.stream_pcm(0, dec.fetch_pcm()).astype(np.float64)
        assert got.shape == ref.pcm.T.shape
        l3util.assert_iso_full_accuracy(got, ref.pcm.T, "long stream")
        pl = dec.planar()
        assert np.array_equal(pl.reshape(2, -1).T, dec.stream_pcm(0, dec.fetch_pcm()))
        t = 4321 * 1152 + 77
        sk = mp3b.seek_plan(s, t)
        full = dec.stream_pcm(0, dec.fetch_pcm()).copy()
        dec.decode_batch([s[sk.byte_offset:]])
        part = dec.stream_pcm(0, dec.fetch_pcm())
        assert np.array_equal(part[sk.discard_samples:], full[t:])


def test_batches_of_large_units(mp3b, synth_mod, oracle_mod):
    """Low sample rate, high bitrate, mono: up to 1,440 bytes of frame per granule-channel.  Such a batch on
    its own makes the Huffman kernel's bit stage as large as it gets (found by tools/fuzz_incremental.py: the
    launch once asked for more shared memory than the kernel may have)."""
    cfgs = [dict(nframes=10, seed=758025, sample_rate=12000, mode=3, bitrate_kbps=64, fill_lo_pct=30,
                 vbr_min_kbps=32, vbr_max_kbps=128),
            dict(nframes=12, seed=5, sample_rate=8000, mode=3, bitrate_kbps=160, blocks=1),
            dict(nframes=12, seed=6, sample_rate=16000, mode=3, bitrate_kbps=160, fill_lo_pct=20),
            dict(nframes=12, seed=7, sample_rate=32000, mode=3, bitrate_kbps=320, blocks=1, mixed_pct=30)]
    with mp3b.Decoder(device=0, pcm_format=mp3b.PCM_F32) as dec:
        for c in cfgs:  # each on its own: the stage is sized by the batch's average unit
            s = synth_mod.make_stream(**c)
            ref = oracle_mod.decode(s)
            dec.decode_batch([s])
            got = dec.stream_pcm(0, dec.fetch_pcm()).astype(np.float64)
            assert got.shape == ref.pcm.T.shape, c
            scale = max(1.0, float(np.abs(ref.pcm).max()))
            l3util.assert_iso_full_accuracy(got / scale, ref.pcm.T / scale, str(c))
