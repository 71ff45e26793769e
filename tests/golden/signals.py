"""Seeded synthetic audio used by the golden fixtures, the parity tests and bench.py.

numpy Generator(PCG64) streams are stable across platforms and numpy versions, so the GPU box
regenerates exactly the inputs the fixtures were made from (only outputs are stored).
Value range follows the reference loader: peak-normalised to [-1, 1]
(reference osuT5/osuT5/dataset/data_utils.py:95-97).
"""
from __future__ import annotations

import numpy as np

SR = 16000


def noise(B: int, N: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.random((B, N), dtype=np.float32) * 2.0 - 1.0).astype(np.float32)


def sine(B: int, N: int, freq: float = 440.0, amp: float = 0.5) -> np.ndarray:
    t = np.arange(N, dtype=np.float64) / SR
    x = amp * np.sin(2.0 * np.pi * freq * t)
    return np.broadcast_to(x.astype(np.float32), (B, N)).copy()


def music(N: int, seed: int) -> np.ndarray:
    """'Music-like' mix: a few decaying harmonic notes on a beat grid + pink-ish noise,
    peak-normalised to 1.0 like load_audio_file does.  1-D [N]."""
    rng = np.random.default_rng(seed)
    t = np.arange(N, dtype=np.float64) / SR
    x = np.zeros(N, dtype=np.float64)
    beat = 0.5  # 120 bpm
    n_beats = int(t[-1] / beat) + 1
    for i in range(n_beats):
        f0 = 110.0 * 2.0 ** (rng.integers(0, 36) / 12.0)
        start = int(i * beat * SR)
        seg = t[start:start + int(0.45 * SR)] - t[start]
        env = np.exp(-6.0 * seg)
        note = sum(np.sin(2 * np.pi * f0 * h * seg) / h for h in (1, 2, 3, 4) if f0 * h < SR / 2)
        x[start:start + len(seg)] += 0.6 * env * note
    white = rng.standard_normal(N)
    # one-pole low-pass of white noise as a cheap pink-ish floor
    pink = np.empty(N)
    acc = 0.0
    a = 0.98
    # vectorised IIR via cumulative products would lose precision; a python loop is too slow
    # for 1-hour inputs, so use a block FFT shaping instead (deterministic given the seed)
    spec = np.fft.rfft(white)
    fr = np.fft.rfftfreq(N, 1.0 / SR)
    spec /= np.sqrt(np.maximum(fr, 20.0))
    pink = np.fft.irfft(spec, n=N)
    pink *= 0.05 / (np.abs(pink).max() + 1e-12)
    x += pink
    x *= 1.0 / np.max(np.abs(x))
    return x.astype(np.float32)


def impulse(N: int, pos: int, amp: float = 1.0) -> np.ndarray:
    x = np.zeros((1, N), dtype=np.float32)
    x[0, pos] = amp
    return x
