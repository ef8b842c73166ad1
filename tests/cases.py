"""Named generator cases shared by the CPU and GPU parity tests.

`FF` cases stay inside what FFmpeg's mp3float models (see DESIGN.md "Known decoder divergences"):
no mixed -> pure-short transitions, LSF intensity positions legal and <= 15, no LSF mixed blocks.
`EXTRA` cases are legal streams outside that set; only the oracle and the CUDA path decode them.
"""
FF = {
    "cfg1_long_cbr128": dict(nframes=24),
    "short_switching": dict(blocks=1, nframes=24, seed=5),
    "mixed_blocks": dict(blocks=1, mixed_pct=50, nframes=24, seed=6),
    "ms_only": dict(mode=1, mode_ext_mask=4, nframes=16, seed=7, blocks=1),
    "intensity_only": dict(mode=1, mode_ext_mask=2, nframes=16, seed=8, blocks=1, mixed_pct=30),
    "ms_plus_intensity": dict(mode=1, mode_ext_mask=8, nframes=16, seed=9, blocks=1, mixed_pct=30),
    "cfg3_320k_joint": dict(mode=1, bitrate_kbps=320, blocks=1, mixed_pct=25, fill_lo_pct=35, nframes=24, seed=10),
    "mono": dict(mode=3, nframes=16, seed=11, blocks=1),
    "dual_channel": dict(mode=2, nframes=16, seed=21, blocks=1),
    "lsf22_stereo": dict(sample_rate=22050, bitrate_kbps=64, nframes=24, seed=12, blocks=1),
    "lsf24_joint": dict(sample_rate=24000, bitrate_kbps=96, nframes=24, seed=13, blocks=1, mode=1,
                        lsf_avoid_illegal_ispos=1),
    "lsf16_mono": dict(sample_rate=16000, bitrate_kbps=32, nframes=24, seed=14, blocks=1, mode=3),
    "vbr_32_320": dict(vbr_min_kbps=32, vbr_max_kbps=320, nframes=24, seed=15, blocks=1),
    "lsf_vbr_8_160": dict(sample_rate=22050, vbr_min_kbps=8, vbr_max_kbps=160, nframes=24, seed=22, blocks=1,
                          mode=1, lsf_avoid_illegal_ispos=1),
    "48k_crc": dict(sample_rate=48000, bitrate_kbps=192, nframes=16, seed=16, blocks=1, crc=1),
    "32k_joint": dict(sample_rate=32000, bitrate_kbps=96, nframes=16, seed=17, blocks=1, mode=1),
    "low_bitrate_32k": dict(bitrate_kbps=32, nframes=24, seed=18, blocks=1, mode=1),
    "no_reservoir": dict(reservoir=0, nframes=16, seed=19, blocks=1),
    "sparse_frames": dict(fill_lo_pct=0, fill_hi_pct=30, nframes=24, seed=20, blocks=1, mode=1),
    # MPEG-2.5 (version bits 00): LSF syntax at 11.025 / 12 / 8 kHz; 8 kHz has its own band partition
    "m25_11k_stereo": dict(sample_rate=11025, bitrate_kbps=32, nframes=24, seed=41, blocks=1),
    "m25_12k_joint": dict(sample_rate=12000, bitrate_kbps=48, nframes=24, seed=42, blocks=1, mode=1,
                          lsf_avoid_illegal_ispos=1),
    "m25_8k_mono": dict(sample_rate=8000, bitrate_kbps=16, nframes=24, seed=43, blocks=1, mode=3),
    "m25_8k_stereo_vbr": dict(sample_rate=8000, vbr_min_kbps=8, vbr_max_kbps=64, nframes=24, seed=44, blocks=1),
}
# one case per big_values code book (table_select), so every codeword family is exercised
for _t in [1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31]:
    FF["table_%02d" % _t] = dict(only_table=_t, nframes=6, seed=100 + _t, bitrate_kbps=320)

# Layer II (MP2): one case per allocation table (3-B.2a..d, LSF), mono / stereo / joint with every bound
L2 = {
    "l2_44k_192_stereo": dict(layer=2, bitrate_kbps=192, nframes=10, seed=51),             # table 1 (30 subbands)
    "l2_48k_128_joint": dict(layer=2, sample_rate=48000, bitrate_kbps=128, mode=1, nframes=10, seed=52),  # table 0
    "l2_44k_64_mono": dict(layer=2, bitrate_kbps=64, mode=3, nframes=10, seed=53),           # table 0
    "l2_32k_48_mono": dict(layer=2, sample_rate=32000, bitrate_kbps=48, mode=3, nframes=10, seed=54),    # table 3
    "l2_44k_64_joint": dict(layer=2, bitrate_kbps=64, mode=1, nframes=10, seed=55),          # table 2 (8 subbands)
    "l2_44k_384_dual_crc": dict(layer=2, bitrate_kbps=384, mode=2, crc=1, nframes=8, seed=56),
    "l2_lsf24_64_joint": dict(layer=2, sample_rate=24000, bitrate_kbps=64, mode=1, nframes=10, seed=57),  # table 4
    "l2_lsf16_32_mono": dict(layer=2, sample_rate=16000, bitrate_kbps=32, mode=3, nframes=10, seed=58),
    "l2_sparse": dict(layer=2, bitrate_kbps=256, nframes=8, seed=59, fill_lo_pct=5, fill_hi_pct=40),
}

# Layer I (MP1): 384-sample frames
L1 = {
    "l1_44k_384_stereo": dict(layer=1, bitrate_kbps=384, nframes=20, seed=61),
    "l1_48k_256_joint": dict(layer=1, sample_rate=48000, bitrate_kbps=256, mode=1, nframes=20, seed=62),
    "l1_32k_128_mono_crc": dict(layer=1, sample_rate=32000, bitrate_kbps=128, mode=3, crc=1, nframes=19, seed=63),
    "l1_lsf22_96_stereo": dict(layer=1, sample_rate=22050, bitrate_kbps=96, nframes=20, seed=64),
    "l1_lsf16_48_mono": dict(layer=1, sample_rate=16000, bitrate_kbps=48, mode=3, nframes=17, seed=65),
}

EXTRA = {
    "mixed_free_transitions": dict(blocks=1, mixed_pct=50, mixed_free=1, nframes=24, seed=31, mode=1),
    "lsf_intensity_full_range": dict(sample_rate=24000, bitrate_kbps=96, nframes=24, seed=32, blocks=1, mode=1),
    "lsf16_joint_vbr": dict(sample_rate=16000, vbr_min_kbps=8, vbr_max_kbps=160, nframes=24, seed=33, blocks=1, mode=1),
    "loud_clipping": dict(level_lo_db=-12, level_hi_db=0, nframes=12, seed=34, blocks=1),
}
