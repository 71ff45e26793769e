"""beatheritage_b200 -- B200-native (sm_100a) drop-in for BeatHeritage's audio frontend.

Only what the hot path needs: the host-side mirror of the reference module
(`MelSpectrogram`, reference osuT5/osuT5/model/spectrogram.py), the fused-segmentation helpers,
the ctypes binding of the C ABI (`include/bhmel.h`) and the CUDA sources under `csrc/`.
"""
from .spectrogram import MelSpectrogram  # noqa: F401

__all__ = ["MelSpectrogram"]
__version__ = "0.1.0"
