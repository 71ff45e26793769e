"""CPU restatement of the reference's `implementation="nnAudio"` arithmetic -- TEST INFRASTRUCTURE ONLY.

**PARITY UNPINNED.**  The reference reaches this arithmetic through
`osuT5/osuT5/model/spectrogram.py:50-61` (reference):

    from nnAudio import features
    features.MelSpectrogram(sr, n_fft, n_mels, hop_length, center=True, fmin, fmax, pad_mode)

nnAudio is a third-party dependency (reference requirements.txt:3, **unpinned**); it is neither
vendored under /root/reference nor installed in this image, and there is no network to fetch it.
The reference has no tests or golden vectors for this path either.  So
this file restates nnAudio's *published* algorithm (the 0.3.x line: `features/mel.py`
`MelSpectrogram`, `features/stft.py` `STFT`, `utils.py` `create_fourier_kernels`,
`librosa_functions.py` `mel` / `mel_frequencies`) as the author of this file knows it, and nothing
here has been checked against nnAudio itself.  What IS checked (tests/test_oracle.py):

* the Slaney-scale, area-normalised filterbank against torchaudio's independent implementation of
  the same published definition (`melscale_fbanks(norm="slaney", mel_scale="slaney")`, which
  torchaudio's own tests hold against librosa);
* the conv-STFT (explicit `frames @ (cos|sin * window)` products) against the FFT form in
  `oracle/mel_oracle.py`, which IS pinned to the reference.

The restated pipeline, all defaults as nnAudio's MelSpectrogram passes them down:

    STFT(n_fft, win_length=n_fft, freq_bins=None, hop_length, window="hann", freq_scale="no",
         center=True, pad_mode, trainable=False, output_format="Magnitude")
      x = ConstantPad1d(n_fft//2, 0)(x)  |  ReflectionPad1d(n_fft//2)(x)
      spec_imag = conv1d(x, wsin, stride=hop);  spec_real = conv1d(x, wcos, stride=hop)
          wsin[k,0,n] = sin(2 pi k n / n_fft) * window[n],  wcos likewise, k = 0 .. n_fft/2
          window = scipy.signal.get_window("hann", n_fft, fftbins=True)   (periodic Hann)
      magnitude = sqrt(spec_real^2 + spec_imag^2)
    spec = magnitude ** 2.0                                             (power = 2.0)
    melspec = matmul(mel_basis, spec)                                   -> [B, n_mels, T]
          mel_basis = mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1)   [n_mels, n_fft/2+1] f32
    (spectrogram.py:79-82 then applies log1p if log_scale and permute(0, 2, 1).)

Only `tests/` may import this file.
"""
from __future__ import annotations

import numpy as np

from .mel_oracle import HOP, N_FFT, center_pad, hann_window


# librosa_functions.py `hz_to_mel` / `mel_to_hz`, htk=False (Slaney's Auditory Toolbox scale):
# linear below 1 kHz (200/3 Hz per mel), logarithmic above (27 mels per factor 6.4).
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel_slaney(f):
    f = np.asarray(f, np.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(f, _MIN_LOG_HZ) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz_slaney(m):
    m = np.asarray(m, np.float64)
    lin = _F_SP * m
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa_functions.py `mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=1)`:
    triangles on the Slaney scale, each scaled by 2 / (its band width in Hz); arithmetic in fp64,
    stored as float32 [n_mels, n_fft//2 + 1]."""
    weights = np.zeros((n_mels, n_fft // 2 + 1), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, n_fft // 2 + 1, endpoint=True)
    mel_f = mel_to_hz_slaney(np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]          # float32 *= float64: product in fp64, rounded once
    return weights


def fourier_kernels(n_fft: int = N_FFT) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """utils.py `create_fourier_kernels(freq_scale="no")` + stft.py's window product:
    (wsin, wcos) float32 [n_fft//2+1, 1, n_fft] already multiplied by the window, and the float32
    window itself."""
    s = np.arange(0, n_fft, 1.0)
    k = np.arange(n_fft // 2 + 1)[:, None]
    ksin = np.sin(2 * np.pi * k * s / n_fft).astype(np.float32)
    kcos = np.cos(2 * np.pi * k * s / n_fft).astype(np.float32)
    window = hann_window(n_fft, np.float32)
    return (ksin * window)[:, None, :], (kcos * window)[:, None, :], window


def mel_forward(x: np.ndarray, *, n_mels: int = 388, f_min: float = 0.0, f_max: float = 8000.0,
                sample_rate: int = 16000, n_fft: int = N_FFT, hop: int = HOP, pad_mode: str = "constant",
                log_scale: bool = False, dtype=np.float64, basis: np.ndarray | None = None,
                kernels: tuple | None = None) -> np.ndarray:
    """[B, N] -> [B, N//hop + 1, n_mels], the conv-STFT written out as matrix products."""
    x = np.asarray(x)
    B, N = x.shape
    xp = center_pad(x.astype(dtype), n_fft, pad_mode)
    wsin, wcos, _ = fourier_kernels(n_fft) if kernels is None else kernels
    wsin, wcos = wsin[:, 0, :].astype(dtype), wcos[:, 0, :].astype(dtype)
    mb = (mel_basis(sample_rate, n_fft, n_mels, f_min, f_max) if basis is None else np.asarray(basis)).astype(dtype)
    frames = np.lib.stride_tricks.sliding_window_view(xp, n_fft, axis=1)[:, ::hop]     # [B, T, n_fft]
    out = np.empty((B, N // hop + 1, mb.shape[0]), dtype=dtype)
    for b in range(B):
        re = frames[b] @ wcos.T
        im = frames[b] @ wsin.T
        mag = np.sqrt(re * re + im * im)
        mel = (mag ** 2.0) @ mb.T
        out[b] = np.log1p(mel) if log_scale else mel
    return out
