"""mp3_b200 -- B200-native batched MPEG-1/2 Layer III decoder (host-side Python mirror of the C-ABI)."""
