"""CPU restatement of the encoder's convolutional stem -- TEST INFRASTRUCTURE ONLY.

Follows the reference encoder, osuT5/osuT5/model/custom_transformers/modeling_ropewhisper.py:

    self.conv1 = nn.Conv1d(self.num_mel_bins, embed_dim, kernel_size=3, padding=1)          (:1135)
    self.conv2 = nn.Conv1d(embed_dim, embed_dim, kernel_size=3, stride=2, padding=1)        (:1136)
    inputs_embeds = nn.functional.gelu(self.conv1(input_features))                          (:1206)
    inputs_embeds = nn.functional.gelu(self.conv2(inputs_embeds))                           (:1207)
    inputs_embeds = inputs_embeds.permute(0, 2, 1)                                          (:1209)

as the reference runs it for inference: the model in bfloat16 (inference.py:486-489 casts
everything but the spectrogram), i.e. bf16 weights and biases, every op's result rounded to bf16,
fp32 accumulation inside the convolution (cuDNN) and fp32 evaluation inside GELU (ATen's opmath).
The arithmetic lives in torch (Conv1d / gelu), which IS installed here, so this file simply calls the
same torch ops on the CPU in fp32 and applies the bf16 roundings at the same points; what differs
from the GPU reference is only the fp32 summation order inside the convolution.

Pinning: the reference has no tests or golden vectors for the encoder and
its RoPE encoder does not construct under the installed transformers; the stem is
two stock torch modules, so `tests/test_gpu_stem.py` pins this file against those very modules
(`torch.nn.Conv1d` + `torch.nn.functional.gelu` in bf16 on the GPU) on the same inputs.

Only `tests/` may import this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def conv_gelu(x_bct: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, stride: int) -> torch.Tensor:
    """One stem stage on [B, C, T] float32 values that are bf16-representable: conv (fp32 accumulate,
    bf16 weights / bias) -> bf16 -> exact GELU in fp32 -> bf16 (returned as float32)."""
    conv = F.conv1d(x_bct, _bf16_round(weight.float()), _bf16_round(bias.float()), stride=stride, padding=1)
    return _bf16_round(F.gelu(_bf16_round(conv)))


def conv_stem(x_btc: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
              return_hidden: bool = False):
    """x_btc [B, T, C] (bf16 or bf16-representable float) channels last -> [B, T/2, D] float32 holding
    bf16 values; `return_hidden` also returns gelu(conv1) as [B, T, D]."""
    x = _bf16_round(x_btc.detach().cpu().float()).swapaxes(1, 2)        # the reference's swapaxes(1, 2)
    h = conv_gelu(x, w1.detach().cpu(), b1.detach().cpu(), 1)
    y = conv_gelu(h, w2.detach().cpu(), b2.detach().cpu(), 2)
    y = y.permute(0, 2, 1).contiguous()
    return (y, h.permute(0, 2, 1).contiguous()) if return_hidden else y
