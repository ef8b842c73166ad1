"""The test signals live in the package (bench.py and tools use them too)."""
from mp3_b200.signals import *  # noqa: F401,F403
from mp3_b200.signals import CODEC_DELAY, castanets, music, snr_db, speech, stereo, to_s16  # noqa: F401
