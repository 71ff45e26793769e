"""Torch-CPU port of the reference MelSpectrogram path -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Same operator sequence the reference runs on CPU (osuT5/osuT5/model/spectrogram.py:38-49,79-82 ->
torchaudio.transforms.MelSpectrogram -> torch.stft -> abs().pow(2) -> matmul(fb) -> log1p ->
permute), written against torch only so it runs on the GPU box where /root/reference does not
exist.  It is what `bench.py`'s cpu_baseline and `--impl reference` legs time ("port"), with all
host threads torch can use.  The product never imports it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import mel_oracle


class TorchPortMel(torch.nn.Module):
    def __init__(self, log_scale=True, sample_rate=16000, n_fft=1024, n_mels=80, hop_length=128,
                 f_min=20, f_max=8000, pad_mode="reflect"):
        super().__init__()
        self.log_scale, self.n_fft, self.hop, self.pad_mode = log_scale, n_fft, hop_length, pad_mode
        self.register_buffer("window", torch.hann_window(n_fft))
        fb = mel_oracle.melscale_fbanks(n_fft // 2 + 1, float(f_min), float(f_max), n_mels,
                                        sample_rate, np.float32)
        self.register_buffer("fb", torch.from_numpy(fb))

    @torch.no_grad()
    def forward(self, samples: torch.Tensor) -> torch.Tensor:
        spec = torch.stft(samples, n_fft=self.n_fft, hop_length=self.hop, win_length=self.n_fft,
                          window=self.window, center=True, pad_mode=self.pad_mode,
                          normalized=False, onesided=True, return_complex=True)
        power = spec.abs().pow(2.0)                                   # [B, F, T]
        mel = torch.matmul(power.transpose(-1, -2), self.fb).transpose(-1, -2)   # [B, M, T]
        if self.log_scale:
            mel = torch.log1p(mel)
        return mel.permute(0, 2, 1)                                   # [B, T, M]
