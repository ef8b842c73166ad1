"""GPU: the planar copy of the decoded batch is the interleaved arena transposed per stream."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["s16", "f32"])
@pytest.mark.parametrize("gapless", [False, True])
def test_planar_is_the_transpose(fmt, gapless, synth_mod):
    import mp3_b200 as m
    cfgs = [cases.FF["cfg1_long_cbr128"], cases.FF["mono"], cases.FF["lsf22_stereo"], cases.L2["l2_44k_192_stereo"],
            cases.L1["l1_32k_128_mono_crc"], cases.L1["l1_44k_384_stereo"],
            dict(nframes=9, seed=5, tag=2, tag_lame=1, enc_delay=577, enc_padding=1001)]
    streams = [synth_mod.make_stream(**c) for c in cfgs] + [b"no audio here"]
    with m.Decoder(device=0, pcm_format=m.PCM_S16 if fmt == "s16" else m.PCM_F32, gapless=gapless) as dec:
        dec.decode_batch(streams)
        inter = dec.fetch_pcm().copy()
        plan = dec.planar()
        assert plan.shape == inter.shape and plan.dtype == inter.dtype
        for i in range(len(cfgs)):
            inf = dec.stream_info(i)
            a = inter[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].reshape(inf.samples, inf.channels)
            b = plan[inf.pcm_offset: inf.pcm_offset + inf.samples * inf.channels].reshape(inf.channels, inf.samples)
            assert np.abs(a).max() > 0 and np.array_equal(a.T, b), i
