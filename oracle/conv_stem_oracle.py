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


def folded_bias(cond_bc: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, n_var: int) -> torch.Tensor:
    """The per-window bias of the split conv1 (include/bhstem.h, bhstem_forward_split), restated on the CPU in
    float64: the reference repeats each window's conditioning vector over all T frames before conv1
    (osuT5/osuT5/model/modeling_mapperatorinator.py:368-370), so those channels add, per tap,
    S_tap[b][n] = sum_c w1[n][n_var + c][tap] * cond[b][c] -- all three taps inside the window, taps 1 and 2 at its
    first frame (tap 0 reads conv1's zero padding, modeling_ropewhisper.py:1135 `padding=1`), taps 0 and 1 at its last.
    cond_bc [B, C - n_var], w1 [D, C, 3], b1 [D] (values as the bf16 model holds them) -> [B, 3, D] float64:
    rows = (interior frames, frame 0, frame T - 1)."""
    w = _bf16_round(w1.detach().cpu().float()).double()[:, n_var:, :]
    s = torch.einsum("nct,bc->btn", w, _bf16_round(cond_bc.detach().cpu().float()).double())      # [B, 3 taps, D]
    b = _bf16_round(b1.detach().cpu().float()).double()
    return torch.stack([b + s[:, 0] + s[:, 1] + s[:, 2], b + s[:, 1] + s[:, 2], b + s[:, 0] + s[:, 1]], dim=1)


def split_conv1_preactivation(frames_btn: torch.Tensor, cond_bc: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor) -> torch.Tensor:
    """conv1's pre-activation computed the split way in float64: the convolution over the n_var time-varying
    channels plus folded_bias by (window, edge) -> [B, T, D].  tests/test_oracle.py checks that this equals conv1 over
    the concatenated input [frames | cond repeated over T] exactly (up to float64 rounding)."""
    n_var = frames_btn.shape[2]
    x = _bf16_round(frames_btn.detach().cpu().float()).double().swapaxes(1, 2)
    w = _bf16_round(w1.detach().cpu().float()).double()[:, :n_var, :]
    y = F.conv1d(x, w, None, stride=1, padding=1).swapaxes(1, 2)                                   # [B, T, D]
    fb = folded_bias(cond_bc, w1, b1, n_var)
    T = y.shape[1]
    y = y + fb[:, 0:1]
    y[:, 0] += fb[:, 1] - fb[:, 0]
    y[:, T - 1] += fb[:, 2] - fb[:, 0]
    return y
