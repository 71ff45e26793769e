"""CPU oracle for the osuT5 MelSpectrogram hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the arithmetic the reference reaches through
`osuT5/osuT5/model/spectrogram.py:38-49,79-82` (reference) when `implementation="torchaudio"`:

    torchaudio.transforms.MelSpectrogram(sample_rate, n_fft, n_mels, hop_length, center=True,
                                         f_min, f_max, pad_mode)        (spectrogram.py:40-49)
      -> torch.stft(center=True, pad_mode, window=hann_window(n_fft), onesided=True)
      -> abs().pow(2)                                   (torchaudio/functional/functional.py `spectrogram`)
      -> matmul(spec^T, melscale_fbanks(htk, norm=None))^T   (torchaudio/transforms/_transforms.py `MelScale`)
    torch.log1p if log_scale                            (spectrogram.py:80-81)
    permute(0, 2, 1) -> [B, T, n_mels]                  (spectrogram.py:82)

The arithmetic itself lives in third-party packages that are NOT vendored under /root/reference:
torchaudio (requirements.txt:21, unpinned; 2.11.0+cu128 in this image) and torch 2.11.0+cu128.
The reference ships no tests or golden vectors for this path, so parity
is pinned by running the reference module itself in the build container
(`tests/golden/make_golden.py` imports `/root/reference/osuT5/osuT5/model/spectrogram.py` by
path and stores its outputs as fixtures; `tests/test_oracle.py` checks this file against them).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package.  The product (`beatheritage_b200/`) never does.
"""
from __future__ import annotations

import math

import numpy as np

N_FFT = 1024
HOP = 128


# --------------------------------------------------------------------------------------
# a6: filterbank -- torchaudio/functional/functional.py melscale_fbanks (@518),
#     _hz_to_mel (@425), _mel_to_hz (@459), _create_triangular_filterbank (@492)
# --------------------------------------------------------------------------------------
def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace(dtype=float32) on CPU: step computed in fp32; the first half is
    fma(step, i, start) and the second half fma(-step, steps-1-i, end) (ATen RangeFactories
    kernel, compiled with FMA contraction -- probed bit-exact against torch 2.11 for every
    parameter set in SURVEY.md appendix B).  fma(a, b, c) with fp32 inputs is emulated
    as float32(float64(a)*float64(b) + float64(c)): the product is exact in fp64."""
    start32, end32 = np.float32(start), np.float32(end)
    if steps == 1:
        return np.array([start32], dtype=np.float32)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    i = np.arange(steps)
    lo = (np.float64(start32) + np.float64(step) * i).astype(np.float32)
    hi = (np.float64(end32) - np.float64(step) * (steps - 1 - i)).astype(np.float32)
    return np.where(i < steps // 2, lo, hi).astype(np.float32)


def hz_to_mel_htk(freq: float) -> float:
    """torchaudio `_hz_to_mel(mel_scale="htk")`: python-float (fp64) arithmetic."""
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                    dtype=np.float32) -> np.ndarray:
    """htk mel scale, norm=None triangular filterbank, shape [n_freqs, n_mels].

    dtype=float32 follows torchaudio's fp32 tensor arithmetic step by step; dtype=float64 is
    the exact-arithmetic arbiter."""
    if dtype == np.float32:
        all_freqs = _linspace_f32(0, sample_rate // 2, n_freqs)
        m_pts = _linspace_f32(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
        # 700.0 * (10.0 ** (mels / 2595.0) - 1.0) on an fp32 tensor
        # (correctly-rounded fp32 pow: evaluate in fp64, round once)
        expo = (m_pts / np.float32(2595.0)).astype(np.float32)
        p10 = np.power(10.0, expo.astype(np.float64)).astype(np.float32)
        f_pts = (np.float32(700.0) * (p10 - np.float32(1.0))).astype(np.float32)
    else:
        all_freqs = np.linspace(0, sample_rate // 2, n_freqs, dtype=np.float64)
        m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2, dtype=np.float64)
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    return fb.astype(dtype)


# --------------------------------------------------------------------------------------
# a4 (window): torch.hann_window(n_fft, periodic=True)
# --------------------------------------------------------------------------------------
def hann_window(n_fft: int = N_FFT, dtype=np.float32) -> np.ndarray:
    n = np.arange(n_fft, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(dtype)


# --------------------------------------------------------------------------------------
# a3: centre padding -- torch/functional.py stft `if center:` F.pad(x, [n_fft//2]*2, pad_mode)
# --------------------------------------------------------------------------------------
def center_pad(x: np.ndarray, n_fft: int, pad_mode: str) -> np.ndarray:
    p = n_fft // 2
    if pad_mode == "reflect":
        if x.shape[-1] <= p:
            # F.pad raises: "Padding size should be less than the corresponding input dimension"
            raise RuntimeError(
                f"reflect padding ({p}) must be smaller than the input length ({x.shape[-1]})")
        return np.pad(x, [(0, 0), (p, p)], mode="reflect")
    if pad_mode == "constant":
        return np.pad(x, [(0, 0), (p, p)], mode="constant")
    raise ValueError(f"unsupported pad_mode {pad_mode!r}")


# --------------------------------------------------------------------------------------
# a2..a9: the whole forward
# --------------------------------------------------------------------------------------
def mel_forward(x: np.ndarray, *, n_mels: int = 80, f_min: float = 20.0, f_max: float = 8000.0,
                sample_rate: int = 16000, n_fft: int = N_FFT, hop: int = HOP,
                pad_mode: str = "reflect", log_scale: bool = True, dtype=np.float64,
                fb: np.ndarray | None = None, window: np.ndarray | None = None,
                frame_chunk: int = 8192) -> np.ndarray:
    """[B, N] -> [B, N//hop + 1, n_mels].  dtype selects the accumulation type: float64 is the
    arbiter; float32 mirrors the reference's fp32 path (FFT in fp32 via complex64 rfft)."""
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("expected a [batch, samples] array")
    B, N = x.shape
    xp = center_pad(x.astype(dtype), n_fft, pad_mode)
    w = hann_window(n_fft, np.float32).astype(dtype) if window is None else np.asarray(window, dtype)
    if fb is None:
        fb = melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate, np.float32)
    fbm = np.asarray(fb).astype(dtype)
    T = N // hop + 1
    out = np.empty((B, T, fbm.shape[1]), dtype=dtype)
    frames_view = np.lib.stride_tricks.sliding_window_view(xp, n_fft, axis=1)[:, ::hop]  # [B,T,n_fft]
    for b in range(B):
        for t0 in range(0, T, frame_chunk):
            fr = frames_view[b, t0:t0 + frame_chunk] * w                # a4: frame * window
            spec = np.fft.rfft(fr, axis=-1)                              # a4: one-sided DFT
            if dtype == np.float32:
                spec = spec.astype(np.complex64)
            power = (spec.real * spec.real + spec.imag * spec.imag).astype(dtype)   # a5
            mel = power @ fbm                                            # a7
            if log_scale:
                mel = np.log1p(mel)                                      # a8 (spectrogram.py:80-81)
            out[b, t0:t0 + frame_chunk] = mel                            # a9: [B, T, M] layout
    return out


# --------------------------------------------------------------------------------------
# a10: window producers (inference) -- osuT5/osuT5/inference/preprocessor.py:12-21, 41-71, 94-102
# --------------------------------------------------------------------------------------
def segment_params(src_seq_len: int = 4096, hop: int = HOP, lookback: float = 0.5,
                   lookahead: float = 0.4, parallel: bool = False) -> tuple[int, int]:
    """(samples_per_sequence, sequence_stride) as Preprocessor.__init__ computes them
    (preprocessor.py:14-21); note the float truncation in the stride."""
    samples_per_sequence = (src_seq_len - 1) * hop
    stride = int(samples_per_sequence * (1 - lookback - lookahead))
    if parallel:
        stride = samples_per_sequence
    return samples_per_sequence, stride


def segment(samples: np.ndarray, samples_per_sequence: int, stride: int,
            begin_pad: int = 0, end_pad: int = 0) -> np.ndarray:
    """Preprocessor.segment (preprocessor.py:58-71): right-pad so the strided windows tile the song
    exactly, then take windows every `stride`.  `sequence_times` and the start/end-time trimming
    (preprocessor.py:72-90) are restated by segment_times_and_trim below."""
    s = np.pad(np.asarray(samples), [begin_pad, end_pad])
    if len(s) < samples_per_sequence:
        padding = samples_per_sequence - len(s)
    else:
        rem = (len(s) - samples_per_sequence) % stride
        padding = 0 if rem == 0 else stride - rem
    s = np.pad(s, [0, padding])
    view = np.lib.stride_tricks.sliding_window_view(s, samples_per_sequence)[::stride]
    return np.ascontiguousarray(view, dtype=np.float32)


def segment_times_and_trim(n_windows: int, samples_per_sequence: int, stride: int, sample_rate: int = 16000,
                           lookback: float = 0.5, lookahead: float = 0.4, start_time=None, end_time=None):
    """The rest of Preprocessor.segment (preprocessor.py:72-90) and the constants it uses
    (preprocessor.py:22-25): `sequence_times` and the start_time / end_time trimming.
    Returns (first_window, n_kept, sequence_times int32 of the kept windows).

    torch.arange(0, W * ms, ms) with Python floats is a float32 tensor whose element i is
    float32(i * ms) (the product formed in double), and `.to(int32)` truncates; the float32
    rounding is visible for long songs (values >= 2^20 ms lose their 1/16 ms fractions), so it is
    restated, not idealised.  torch.searchsorted compares the int32 boundaries with the Python
    float as real numbers (probed: 3275 < 3275.5)."""
    ms_per_stride = stride * 1000 / sample_rate                            # :22
    ms_per_sequence = samples_per_sequence * 1000 / sample_rate            # :23
    lookback_max_time = lookback * ms_per_sequence                         # :24
    lookahead_max_time = (1 - lookahead) * ms_per_sequence                 # :25
    n = int(np.ceil((n_windows * ms_per_stride) / ms_per_stride)) if n_windows > 0 else 0   # arange's length
    times = (np.arange(n, dtype=np.float64) * ms_per_stride).astype(np.float32).astype(np.int32)   # :72-73
    first = 0
    if start_time is not None:                                             # :75-82
        first = int(np.searchsorted(times.astype(np.float64), start_time - lookahead_max_time, side="right"))
        if first == len(times):
            first -= 1
        times = times[first:]
    kept = len(times)
    if end_time is not None:                                               # :83-90
        kept = int(np.searchsorted(times.astype(np.float64), end_time - lookback_max_time, side="left"))
        if kept == 0:
            kept += 1
        times = times[:kept]
    return first, len(times), times


# --------------------------------------------------------------------------------------
# a10: window producers (training) -- osuT5/osuT5/dataset/ors_dataset.py:243-262, 303-343, 563-590
# --------------------------------------------------------------------------------------
def dataset_windows(samples: np.ndarray, src_seq_len: int = 4096, hop: int = HOP,
                    offset: int = 0, gen_start_frame: int = 0) -> np.ndarray:
    """_get_frames + _create_sequences (frame slicing only) + _pad_frame_sequence:
    pad to a hop multiple (a full extra hop when already aligned, ors_dataset.py:256),
    cut (src_seq_len-1)-hop-frame windows from `offset`, zero-pad the last one, flatten."""
    s = np.asarray(samples, dtype=np.float32)
    s = np.pad(s, [0, hop - len(s) % hop])
    frames = s.reshape(-1, hop)
    fsl = src_seq_len - 1
    n_frames = len(frames)
    out = []
    # gen_start_frame = round(lookback * frame_seq_len) (ors_dataset.py:230); 0 in configs/train
    for start in range(offset, n_frames - gen_start_frame, fsl):
        chunk = frames[start:min(start + fsl, n_frames)]
        padded = np.zeros((fsl, hop), dtype=np.float32)
        padded[:len(chunk)] = chunk
        out.append(padded.reshape(-1))
    return np.stack(out) if out else np.zeros((0, fsl * hop), np.float32)
