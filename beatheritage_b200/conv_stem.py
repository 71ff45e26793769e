"""Host-side mirror of the encoder's convolutional stem.

The reference encoder (osuT5/osuT5/model/custom_transformers/modeling_ropewhisper.py:1135-1136,
1206-1209; the stock HF WhisperEncoder is identical here) starts with

    inputs_embeds = gelu(conv1(input_features))        # Conv1d(C_in, D, 3, padding=1) on [B, C_in, T]
    inputs_embeds = gelu(conv2(inputs_embeds))         # Conv1d(D, D, 3, stride=2, padding=1)
    inputs_embeds = inputs_embeds.permute(0, 2, 1)     # [B, T/2, D]

`ConvStem` keeps the parameter names (`conv1.weight`, `conv1.bias`, `conv2.weight`, `conv2.bias`)
so it loads from the encoder's state dict, and computes the three lines with two tcgen05 implicit
GEMM launches behind the C ABI in include/bhstem.h.  It consumes the CHANNELS-LAST encoder input
[B, T, C_in] bf16 -- what `MelSpectrogram.forward_encoder_input(..., channels_first=False)` returns,
i.e. the tensor the reference has *before* its `swapaxes(1, 2)` (modeling_mapperatorinator.py:
368-376) -- so that swapaxes and the permute above both disappear.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
import threading

import torch
import torch.nn as nn

from . import _stem_lib

_handle_lock = threading.Lock()


class ConvStem(nn.Module):
    def __init__(self, num_mel_bins: int, d_model: int):
        """num_mel_bins = encoder input channels (config.num_mel_bins: mel + conditioning channels),
        d_model = config.d_model (reference modeling_ropewhisper.py:1129-1136)."""
        super().__init__()
        if num_mel_bins % 8 or d_model % 128:
            raise ValueError("the sm_100a stem needs num_mel_bins % 8 == 0 and d_model % 128 == 0")
        self.conv1 = nn.Conv1d(num_mel_bins, d_model, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(d_model, d_model, kernel_size=3, stride=2, padding=1)
        self._handles: dict[int, int] = {}
        self._stamp: dict[int, tuple] = {}
        self._retired: list[int] = []     # handles of superseded parameters: freed with the module, never earlier
        self._split: dict[int, tuple] = {}  # device -> (handle, n_var) prepared for forward_split

    @classmethod
    def from_encoder(cls, encoder: nn.Module) -> "ConvStem":
        """Build from any module that carries `conv1` / `conv2` like the reference encoder."""
        stem = cls(encoder.conv1.in_channels, encoder.conv1.out_channels)
        stem.load_state_dict({k: v for k, v in encoder.state_dict().items() if k.startswith(("conv1.", "conv2."))})
        return stem.to(encoder.conv1.weight.device)

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"], state["_stamp"], state["_retired"], state["_split"] = {}, {}, [], {}
        return state

    def _param_stamp(self) -> tuple:
        ps = (self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias)
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _handle_for(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        stamp = self._param_stamp()
        with _handle_lock:
            h = self._handles.get(idx)
            if h is not None and self._stamp.get(idx) == stamp:
                return h
            lib = _stem_lib.lib()
            if h is not None:
                # Parameters were reloaded / edited: repack into a NEW handle.  Another thread may still
                # be inside bhstem_forward with the old one (the lock is released before the launch), so
                # it is only retired here and destroyed together with the module.
                self._retired.append(h)
                del self._handles[idx]
            host = [p.detach().to("cpu", torch.float32).contiguous()
                    for p in (self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias)]
            fp = ctypes.POINTER(ctypes.c_float)
            out = ctypes.c_void_p()
            with torch.cuda.device(idx):
                _stem_lib.check(lib.bhstem_create(self.conv1.in_channels, self.conv1.out_channels,
                                                  *[ctypes.cast(t.data_ptr(), fp) for t in host], ctypes.byref(out)))
            if getattr(self, "_variant", 2) != 2:
                _stem_lib.check(lib.bhstem_set_option(out.value, 1, self._variant))
            if getattr(self, "_epi", None) is not None:
                _stem_lib.check(lib.bhstem_set_option(out.value, 3, self._epi))
            if getattr(self, "_deep", None) is not None:
                _stem_lib.check(lib.bhstem_set_option(out.value, 5, self._deep))
            if not getattr(self, "_small", True):
                _stem_lib.check(lib.bhstem_set_option(out.value, 4, 0))
            if not getattr(self, "_pdl", True):
                _stem_lib.check(lib.bhstem_set_option(out.value, 2, 0))
            self._handles[idx] = out.value
            self._stamp[idx] = stamp
            self._split.pop(idx, None)
            return out.value

    def _split_handle_for(self, device: torch.device, n_var: int) -> int:
        """The handle with its time-varying conv1 channels repacked (bhstem_prepare_split, once per handle)."""
        h = self._handle_for(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        with _handle_lock:
            done = self._split.get(idx)
            if done is not None and done != (h, n_var):
                raise RuntimeError(f"this ConvStem already runs split with {done[1]} time-varying channels")
            if done is None:
                with torch.cuda.device(idx):
                    _stem_lib.check(_stem_lib.lib().bhstem_prepare_split(h, n_var))
                self._split[idx] = (h, n_var)
        return h

    VARIANTS = {"tap_boxes": 0, "shared_taps": 1, "cta_pairs": 2}

    def set_variant(self, name: str) -> None:
        """Kernel schedule (A/B runs; same results within the bf16 tolerance): "cta_pairs" (default: tcgen05
        cta_group::2 where d_model % 256 == 0), "shared_taps", "tap_boxes".  bhstem_set_option, include/bhstem.h."""
        self._variant = self.VARIANTS[name]
        lib = _stem_lib.lib()
        for h in self._handles.values():
            _stem_lib.check(lib.bhstem_set_option(h, 1, self._variant))

    def set_epilogue_warps(self, conv1: int = 8, conv2: int = 8, split_conv1: int = 16) -> None:
        """Epilogue warps of the CTA-pair kernel per stage, 8 or 16 each (A/B runs; same bits).
        BHSTEM_OPT_EPILOGUE_WARPS, include/bhstem.h."""
        self._epi = conv1 | (conv2 << 8) | (split_conv1 << 16)
        lib = _stem_lib.lib()
        for h in self._handles.values():
            _stem_lib.check(lib.bhstem_set_option(h, 3, self._epi))

    def set_deep_a_ring(self, mask: int) -> None:
        """3 activation + 6 weight stages (bit set, the default for conv1 / conv2) or 2 + 8 in the CTA-pair kernel, per
        stage (bit 0 conv1, bit 1 conv2, bit 2 split conv1; A/B runs, same bits).  BHSTEM_OPT_DEEP_A_RING."""
        self._deep = int(mask)
        lib = _stem_lib.lib()
        for h in self._handles.values():
            _stem_lib.check(lib.bhstem_set_option(h, 5, self._deep))

    def set_small_batch_tiles(self, on: bool) -> None:
        """128-column tiles for launches that would leave half the SMs idle (default on; same bits).
        BHSTEM_OPT_SMALL_BATCH_TILES, include/bhstem.h."""
        self._small = bool(on)
        lib = _stem_lib.lib()
        for h in self._handles.values():
            _stem_lib.check(lib.bhstem_set_option(h, 4, int(self._small)))

    def set_pdl(self, on: bool) -> None:
        """Programmatic dependent launch (default on): a kernel's prologue overlaps the previous kernel's tail;
        all global accesses still wait for that kernel.  BHSTEM_OPT_PDL, include/bhstem.h."""
        self._pdl = bool(on)
        lib = _stem_lib.lib()
        for h in self._handles.values():
            _stem_lib.check(lib.bhstem_set_option(h, 2, int(self._pdl)))

    def __del__(self):
        try:
            lib = _stem_lib.lib()
            for h in [*self._handles.values(), *self._retired]:
                lib.bhstem_destroy(h)
        except Exception:
            pass

    def launch_count(self) -> int:
        lib = _stem_lib.lib()
        return sum(int(lib.bhstem_launch_count(h)) for h in self._handles.values())

    def _check(self, x: torch.Tensor) -> None:
        if x.dim() != 3 or x.shape[2] != self.conv1.in_channels:
            raise RuntimeError(f"expected channels-last input [B, T, {self.conv1.in_channels}], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("beatheritage_b200.ConvStem has no CPU path: move the batch to a CUDA (sm_100) device")
        if x.dtype != torch.bfloat16:
            raise RuntimeError("ConvStem computes in bfloat16 (the reference model's inference dtype): pass a bfloat16 tensor")
        if x.shape[1] < 2 or x.shape[1] % 2:
            raise RuntimeError("the number of frames T must be even")

    @torch.no_grad()
    def forward(self, x: torch.Tensor, hidden: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """x [B, T, C_in] bf16 channels last -> [B, T/2, D] bf16
        == gelu(conv2(gelu(conv1(x.swapaxes(1, 2))))).permute(0, 2, 1).
        `hidden` [B, T, D] and `out` [B, T/2, D] (bf16, contiguous, same device) may be passed in to keep
        the two allocations out of a serving loop."""
        self._check(x)
        x = x.contiguous()
        B, T, _ = x.shape
        D = self.conv1.out_channels
        for name, buf, shape in (("hidden", hidden, (B, T, D)), ("out", out, (B, T // 2, D))):
            if buf is not None and not (tuple(buf.shape) == shape and buf.dtype == torch.bfloat16 and buf.is_contiguous()
                                        and buf.device == x.device):
                raise RuntimeError(f"{name} must be a contiguous bfloat16 tensor {shape} on {x.device}")
        if hidden is None:
            hidden = torch.empty((B, T, D), dtype=torch.bfloat16, device=x.device)
        y = out if out is not None else torch.empty((B, T // 2, D), dtype=torch.bfloat16, device=x.device)
        h = self._handle_for(x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _stem_lib.check(_stem_lib.lib().bhstem_forward(h, x.data_ptr(), B, T, hidden.data_ptr(), y.data_ptr(), stream))
        return y

    @torch.no_grad()
    def forward_split(self, frames: torch.Tensor, cond: torch.Tensor, hidden: torch.Tensor | None = None,
                      out: torch.Tensor | None = None, bias_scratch: torch.Tensor | None = None) -> torch.Tensor:
        """The stem on the reference's encoder input WITHOUT building it: `frames` [B, T, n_mels] bf16 (what
        `MelSpectrogram.forward_into` writes into a dense bf16 buffer) and `cond` [B, C_in - n_mels] bf16, the
        concatenated conditioning embeddings of each window, which the reference repeats over the T frames
        (modeling_mapperatorinator.py:368-370) before conv1 (modeling_ropewhisper.py:1206).  Equals

            x = torch.cat([frames, cond.unsqueeze(1).expand(-1, T, -1)], dim=-1)
            self(x)

        up to the fp32 summation order inside conv1: the time-constant channels are folded into a per-window
        bias (three small sums per output channel), conv1 multiplies n_mels instead of C_in channels.
        `bias_scratch` [B, 3, D] float32 may be passed in like `hidden` / `out` (it receives the folded bias:
        interior frames, frame 0, frame T - 1).  bhstem_forward_split, include/bhstem.h."""
        C, D = self.conv1.in_channels, self.conv1.out_channels
        if frames.dim() != 3 or cond.dim() != 2 or frames.shape[0] != cond.shape[0] or frames.shape[2] + cond.shape[1] != C:
            raise RuntimeError(f"expected frames [B, T, n] and cond [B, {C} - n], got {tuple(frames.shape)} and {tuple(cond.shape)}")
        if not (frames.is_cuda and cond.is_cuda and frames.device == cond.device):
            raise RuntimeError("beatheritage_b200.ConvStem has no CPU path: move the batch to a CUDA (sm_100) device")
        if frames.dtype != torch.bfloat16 or cond.dtype != torch.bfloat16:
            raise RuntimeError("ConvStem computes in bfloat16 (the reference model's inference dtype): pass bfloat16 tensors")
        B, T, n_var = frames.shape
        if T < 2 or T % 2:
            raise RuntimeError("the number of frames T must be even")
        if n_var % 8 or n_var < 8 or cond.shape[1] < 8:
            raise RuntimeError("the number of time-varying channels must be a multiple of 8 below C_in")
        frames, cond = frames.contiguous(), cond.contiguous()
        for name, buf, shape in (("hidden", hidden, (B, T, D)), ("out", out, (B, T // 2, D))):
            if buf is not None and not (tuple(buf.shape) == shape and buf.dtype == torch.bfloat16 and buf.is_contiguous()
                                        and buf.device == frames.device):
                raise RuntimeError(f"{name} must be a contiguous bfloat16 tensor {shape} on {frames.device}")
        if hidden is None:
            hidden = torch.empty((B, T, D), dtype=torch.bfloat16, device=frames.device)
        y = out if out is not None else torch.empty((B, T // 2, D), dtype=torch.bfloat16, device=frames.device)
        if bias_scratch is not None and not (tuple(bias_scratch.shape) == (B, 3, D) and bias_scratch.dtype == torch.float32
                                             and bias_scratch.is_contiguous() and bias_scratch.device == frames.device):
            raise RuntimeError(f"bias_scratch must be a contiguous float32 tensor {(B, 3, D)} on {frames.device}")
        bias3 = bias_scratch if bias_scratch is not None else torch.empty((B, 3, D), dtype=torch.float32, device=frames.device)
        h = self._split_handle_for(frames.device, n_var)
        with torch.cuda.device(frames.device):
            stream = torch.cuda.current_stream(frames.device).cuda_stream
            _stem_lib.check(_stem_lib.lib().bhstem_forward_split(h, frames.data_ptr(), cond.data_ptr(), B, T,
                                                                 bias3.data_ptr(), hidden.data_ptr(), y.data_ptr(), stream))
        return y

    @torch.no_grad()
    def forward_stage(self, stage: int, x: torch.Tensor) -> torch.Tensor:
        """conv1 + GELU (stage 1: [B, T, C_in] -> [B, T, D]) or conv2 + GELU (stage 2: [B, T, D] ->
        [B, T/2, D]) alone, channels last; for tests and profiling."""
        D = self.conv1.out_channels
        if stage == 1:
            self._check(x)
        elif not (x.dim() == 3 and x.shape[2] == D and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 2 == 0):
            raise RuntimeError(f"stage 2 expects a bfloat16 CUDA tensor [B, even T, {D}]")
        x = x.contiguous()
        B, T, _ = x.shape
        out = torch.empty((B, T if stage == 1 else T // 2, D), dtype=torch.bfloat16, device=x.device)
        h = self._handle_for(x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _stem_lib.check(_stem_lib.lib().bhstem_forward_stage(h, stage, x.data_ptr(), B, T, out.data_ptr(), stream))
        return out
