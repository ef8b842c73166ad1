"""GPU: mp3b_seek_plan through the CUDA path -- the PCM decoded from the planned offset equals, bit for bit
from the target on, the PCM of a decode from the stream's start (one batch holds the whole streams and all
their seek slices; batch and incremental interface)."""
import numpy as np
import pytest

from test_seek_cpu import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["f32", "s16"])
def test_seek_slices_equal_the_full_decode(fmt, synth_mod):
    import mp3_b200 as m
    rng = np.random.default_rng(11)
    names = sorted(CASES)
    streams = [synth_mod.make_stream(**CASES[n]) for n in names]
    batch, plan = list(streams), []
    for k, s in enumerate(streams):
        frames, info, _ = m.index_stream_host(s)
        for t in [0, info.samples - 1] + list(rng.integers(0, info.samples, 5)):
            sk = m.seek_plan(s, int(t), frames)
            plan.append((k, int(t), sk.discard_samples, len(batch)))
            batch.append(s[sk.byte_offset:])
    with m.Decoder(device=0, pcm_format=m.PCM_F32 if fmt == "f32" else m.PCM_S16) as dec:
        dec.decode_batch(batch)
        arena = dec.fetch_pcm()
        for k, t, drop, j in plan:
            full, part = dec.stream_pcm(k, arena), dec.stream_pcm(j, arena)
            assert part.shape[0] - drop == full.shape[0] - t
            assert np.array_equal(part[drop:], full[t:]), (names[k], t)
        # the same through the incremental interface: open at the planned offset, feed in pieces
        for k, t, drop, j in plan[::5]:
            s = batch[j]
            h = dec.open_stream()
            got, pos = [], 0
            while pos < len(s):
                n = int(rng.integers(200, 4000))
                h.enqueue(s[pos: pos + n])
                pos += n
                dec.decode_streams()
                inf = h.info()
                if inf.samples:
                    got.append(h.fetch(inf.samples))
            h.close()
            dec.decode_batch(batch[k: k + 1])
            full = dec.stream_pcm(0, dec.fetch_pcm())
            assert np.array_equal(np.concatenate(got)[drop:], full[t:]), (names[k], t)
