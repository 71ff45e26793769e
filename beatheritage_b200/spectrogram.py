"""Host-side mirror of the reference audio frontend.

`MelSpectrogram` keeps the interface of reference osuT5/osuT5/model/spectrogram.py:7-83 --
same constructor signature (including the misspelt `n_ftt` and the positional order the call
sites rely on, reference osuT5/osuT5/model/modeling_mapperatorinator.py:56-66 and
osuT5/dataloading.py:81-90), same `forward(samples[B, N]) -> [B, N // hop + 1, n_mels]`
float32 contract, same state-dict keys (`transform.spectrogram.window`,
`transform.mel_scale.fb`) -- but the work is one fused CUDA kernel behind the C ABI in
include/bhmel.h.  PyTorch is only plumbing here: device memory, the current stream, and the
`torch.library` custom op that keeps `torch.compile(model)` (reference osuT5/train.py:101-102)
from graph-breaking.  There is no CPU fallback: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
import math
import threading
import weakref

import torch
import torch.nn as nn

from . import _lib

N_FFT = 1024
HOP = 128


# ----------------------------------------------------------------------------------------
# Buffers, computed with the same torch ops torchaudio uses so the state dict is identical
# ----------------------------------------------------------------------------------------
def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """htk mel scale, norm=None triangular filterbank [n_freqs, n_mels] float32.

    Restates torchaudio.functional.melscale_fbanks (functional.py @518 with `_hz_to_mel` @425,
    `_mel_to_hz` @459, `_create_triangular_filterbank` @492) operation by operation in fp32 torch
    arithmetic, which makes the buffer bit-identical to the one the reference module registers."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    rising = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    falling = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(rising, falling))


def melscale_fbanks_slaney(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """Slaney mel scale, area-normalised triangular filterbank [n_freqs, n_mels] float32 -- the
    transpose of the `mel_basis` nnAudio's MelSpectrogram builds with its defaults `htk=False,
    norm=1` (reference spectrogram.py:52-61 passes neither).  UNPINNED against nnAudio itself (not
    vendored, not installable here: SURVEY.md 8c / N4); restated from the published definition
    (linear 200/3 Hz per mel below 1 kHz, 27 mels per factor 6.4 above; each triangle scaled by
    2 / band width in Hz; fp64 arithmetic rounded once to fp32) and held against torchaudio's
    independent implementation of it in tests/test_host.py."""
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp

    def to_mel(f: float) -> float:
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2, dtype=torch.float64)
    f_pts = torch.where(mels >= min_log_mel, min_log_hz * torch.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    all_freqs = torch.linspace(0, sample_rate / 2, n_freqs, dtype=torch.float64)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    tri = torch.clamp(torch.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)
    tri = tri.to(torch.float32).to(torch.float64)                 # the triangles are stored in fp32 first
    return (tri * (2.0 / (f_pts[2:] - f_pts[:-2])).unsqueeze(0)).to(torch.float32)


class _WindowHolder(nn.Module):
    """Carries the `window` buffer under the name torchaudio's Spectrogram uses."""

    def __init__(self, n_fft: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(n_fft), persistent=True)


class _FbHolder(nn.Module):
    """Carries the `fb` buffer under the name torchaudio's MelScale uses."""

    def __init__(self, fb: torch.Tensor):
        super().__init__()
        self.register_buffer("fb", fb, persistent=True)


class _Transform(nn.Module):
    """Namespace module so the state-dict keys read `transform.spectrogram.window` and
    `transform.mel_scale.fb` exactly like torchaudio.transforms.MelSpectrogram's."""

    def __init__(self, n_fft: int, fb: torch.Tensor):
        super().__init__()
        self.spectrogram = _WindowHolder(n_fft)
        self.mel_scale = _FbHolder(fb)


# ----------------------------------------------------------------------------------------
# torch.library custom op (opaque to dynamo/inductor, shape-inferable)
# ----------------------------------------------------------------------------------------
_registry: "weakref.WeakValueDictionary[int, MelSpectrogram]" = weakref.WeakValueDictionary()
_registry_lock = threading.Lock()
_handle_lock = threading.Lock()
_next_key = [1]


@torch.library.custom_op("beatheritage_b200::mel_forward", mutates_args=(), device_types="cuda")
def _mel_forward_op(samples: torch.Tensor, module_key: int, n_mels: int) -> torch.Tensor:
    mod = _registry.get(module_key)
    if mod is None:
        raise RuntimeError("beatheritage_b200::mel_forward: the owning MelSpectrogram module is gone")
    return mod._launch(samples)


@_mel_forward_op.register_fake
def _(samples, module_key, n_mels):
    return samples.new_empty((samples.shape[0], samples.shape[1] // HOP + 1, n_mels), dtype=torch.float32)


def _no_backward(ctx, grad):
    raise RuntimeError(
        "beatheritage_b200.MelSpectrogram is a forward-only replacement: the reference never differentiates "
        "through its frontend (no parameters; the input never requires grad), so no backward is provided. "
        "Detach the samples, or keep torchaudio's transform for an input that needs gradients.")


_mel_forward_op.register_autograd(_no_backward)


# ----------------------------------------------------------------------------------------
# State dict in nnAudio's layout (implementation="nnAudio", nnaudio_arithmetic="published").
# nnAudio 0.3.x registers (names restated from its published source, UNPINNED like the rest of N4):
#   transform.mel_basis         [n_mels, n_fft/2+1]     the filterbank, mel-major
#   transform.stft.wsin / wcos  [n_fft/2+1, 1, n_fft]   sin / cos DFT kernels times the window
#   transform.stft.window_mask  [1, n_fft, 1]           the window
# The kernel keeps working from `fb` [n_fft/2+1, n_mels] and `window` [n_fft]; the hooks translate.
# ----------------------------------------------------------------------------------------
def _dft_kernels(window: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    n_fft = window.numel()
    ang = 2 * math.pi * torch.arange(n_fft // 2 + 1, dtype=torch.float64).unsqueeze(1) \
        * torch.arange(n_fft, dtype=torch.float64) / n_fft
    w = window.detach().to("cpu", torch.float32)
    return ((torch.sin(ang).to(torch.float32) * w).unsqueeze(1), (torch.cos(ang).to(torch.float32) * w).unsqueeze(1))


def _nnaudio_save_hook(module, state_dict, prefix, local_metadata):
    fb = state_dict.pop(prefix + "transform.mel_scale.fb")
    window = state_dict.pop(prefix + "transform.spectrogram.window")
    wsin, wcos = _dft_kernels(window)
    state_dict[prefix + "transform.mel_basis"] = fb.t().contiguous()
    state_dict[prefix + "transform.stft.wsin"] = wsin.to(window.device)
    state_dict[prefix + "transform.stft.wcos"] = wcos.to(window.device)
    state_dict[prefix + "transform.stft.window_mask"] = window.reshape(1, -1, 1)


def _nnaudio_load_hook(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
    """Accept nnAudio's buffers: `mel_basis` becomes `fb`, the window is read back from the k = 0 row
    of `wcos` (cos 0 = 1) or from `window_mask`, and every other `transform.*` key nnAudio may carry
    (wsin, inverse kernels) is consumed after checking that the kernels are the plain windowed DFT
    -- a *trained* STFT (nnAudio trainable_STFT) is not what the FFT kernel computes."""
    tp = prefix + "transform."
    ours = {tp + "spectrogram.window", tp + "mel_scale.fb"}
    theirs = [k for k in state_dict if k.startswith(tp) and k not in ours]
    if not theirs:
        return
    got = {k[len(tp):]: state_dict.pop(k) for k in theirs}
    n_fft, n_bins = module.n_fft, module.n_fft // 2 + 1
    if "mel_basis" in got:
        mb = got["mel_basis"]
        if tuple(mb.shape) != (module.n_mels, n_bins):
            error_msgs.append(f"{tp}mel_basis: expected shape {(module.n_mels, n_bins)}, got {tuple(mb.shape)}")
            return
        state_dict[tp + "mel_scale.fb"] = mb.detach().to(torch.float32).t().contiguous()
    window = None
    if "stft.wcos" in got and tuple(got["stft.wcos"].shape) == (n_bins, 1, n_fft):
        window = got["stft.wcos"][0, 0, :].detach().to(torch.float32).clone()
    elif "stft.window_mask" in got and got["stft.window_mask"].numel() == n_fft:
        window = got["stft.window_mask"].detach().to(torch.float32).reshape(n_fft).clone()
    if window is not None:
        if "stft.wcos" in got and "stft.wsin" in got:
            wsin, wcos = _dft_kernels(window)
            dev = max(float((got["stft.wsin"].detach().cpu().float() - wsin).abs().max()),
                      float((got["stft.wcos"].detach().cpu().float() - wcos).abs().max()))
            if dev > 1e-4:
                error_msgs.append(f"{tp}stft.wsin/wcos are not window * sin/cos(2 pi k n / n_fft) (max deviation "
                                  f"{dev:.3g}): a trained STFT is not supported by the FFT kernel")
                return
        state_dict[tp + "spectrogram.window"] = window


# ----------------------------------------------------------------------------------------
class MelSpectrogram(nn.Module):
    #: nnAudio arithmetic cannot be pinned (the package is neither vendored in the reference nor
    #: installed here), so implementation="nnAudio" raises unless the integrator opts in:
    #:   nnaudio_arithmetic = "published"   nnAudio's published defaults (Slaney scale, area-normalised
    #:       filterbank `melscale_fbanks_slaney`, periodic Hann conv-STFT == windowed DFT), state dict
    #:       under nnAudio's buffer names; buffers loaded from a checkpoint replace the restated tables,
    #:       so a checkpoint that carries `mel_basis` / `wcos` pins the arithmetic by itself;
    #:   nnaudio_arithmetic = "torchaudio"  (or the older allow_nnaudio_as_torchaudio = True) the
    #:       torchaudio arithmetic (htk, no normalisation) with these parameters.
    nnaudio_arithmetic: str | None = None
    allow_nnaudio_as_torchaudio = False

    def __init__(
        self,
        implementation: str = "nnAudio",
        log_scale: bool = False,
        sample_rate: int = 16000,
        n_ftt: int = 2048,
        n_mels: int = 512,
        hop_length: int = 128,
        f_min: int = 0,
        f_max: int = 8000,
        pad_mode: str = "constant",
    ):
        """Melspectrogram transformation layer on B200 (see module docstring).

        Arguments mirror the reference (spectrogram.py:8-33): implementation, log_scale,
        sample_rate, n_ftt (STFT size), n_mels, hop_length, f_min, f_max, pad_mode."""
        super().__init__()
        assert implementation in ["torchaudio", "nnAudio"], f"Unsupported implementation: {implementation}"
        mode = None
        if implementation == "nnAudio":
            mode = self.nnaudio_arithmetic or ("torchaudio" if self.allow_nnaudio_as_torchaudio else None)
            if mode not in ("published", "torchaudio"):
                raise NotImplementedError(
                    "implementation='nnAudio': nnAudio's arithmetic is unpinned (not vendored by the reference, "
                    "not installable offline); set MelSpectrogram.nnaudio_arithmetic = 'published' for nnAudio's "
                    "published defaults (Slaney filterbank; buffers from a checkpoint take precedence) or "
                    "'torchaudio' for the torchaudio arithmetic with these parameters")
        if n_ftt != N_FFT or hop_length != HOP:
            raise ValueError(
                f"the sm_100a kernel is compiled for n_ftt={N_FFT}, hop_length={HOP} (constant in every reference "
                f"config); got n_ftt={n_ftt}, hop_length={hop_length}")
        if pad_mode not in ("reflect", "constant"):
            raise ValueError(f"unsupported pad_mode {pad_mode!r} (reference configs use 'reflect' or 'constant')")
        self.implementation = implementation
        self.log_scale = log_scale
        self.sample_rate = sample_rate
        self.n_fft = n_ftt
        self.n_mels = n_mels
        self.hop_length = hop_length
        self.f_min = f_min
        self.f_max = f_max
        self.pad_mode = pad_mode
        self._nnaudio_layout = mode == "published"
        make_fb = melscale_fbanks_slaney if self._nnaudio_layout else melscale_fbanks_htk
        fb = make_fb(n_ftt // 2 + 1, float(f_min), float(f_max), n_mels, sample_rate)
        self.transform = _Transform(n_ftt, fb)
        if self._nnaudio_layout:
            # scipy.signal.get_window("hann", n_fft, fftbins=True).astype(float32): fp64 arithmetic rounded
            # once, a few 1e-9 away from torch.hann_window's fp32 evaluation
            n = torch.arange(n_ftt, dtype=torch.float64)
            self.transform.spectrogram.window.copy_((0.5 - 0.5 * torch.cos(2 * math.pi * n / n_ftt)).to(torch.float32))
            self.register_load_state_dict_pre_hook(_nnaudio_load_hook)
            self.register_state_dict_post_hook(_nnaudio_save_hook)
        self._handles: dict[int, int] = {}        # device index -> bhmel_handle*
        self._stamp: dict[int, tuple] = {}        # device index -> buffer versions the handle was built from
        self._bulk = True
        self._variant = _lib.KERNEL_WARP_SPECIALIZED
        self._static_mel = 1
        self._register()

    def _register(self) -> None:
        with _registry_lock:
            self._key = _next_key[0]
            _next_key[0] += 1
            _registry[self._key] = self

    # handles are process-local device resources: never copied or pickled with the module
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_handles"], state["_stamp"] = {}, {}
        state.pop("_key", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._register()

    # -- handle management -------------------------------------------------------------
    def _buffer_stamp(self) -> tuple:
        w, fb = self.transform.spectrogram.window, self.transform.mel_scale.fb
        return (w.data_ptr(), w._version, fb.data_ptr(), fb._version)

    def _handle_for(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        stamp = self._buffer_stamp()
        with _handle_lock:
            h = self._handles.get(idx)
            if h is not None and self._stamp.get(idx) == stamp:
                return h
            lib = _lib.lib()
            fb = self.transform.mel_scale.fb.detach().to("cpu", torch.float32).contiguous()
            win = self.transform.spectrogram.window.detach().to("cpu", torch.float32).contiguous()
            if tuple(fb.shape) != (self.n_fft // 2 + 1, self.n_mels) or tuple(win.shape) != (self.n_fft,):
                raise RuntimeError(f"unexpected buffer shapes fb{tuple(fb.shape)} window{tuple(win.shape)}")
            fp = ctypes.POINTER(ctypes.c_float)
            with torch.cuda.device(idx):
                if h is None:
                    prm = _lib.BhmelParams(
                        self.sample_rate, self.n_fft, self.hop_length, self.n_mels, float(self.f_min),
                        float(self.f_max), _lib.PAD_REFLECT if self.pad_mode == "reflect" else _lib.PAD_CONSTANT,
                        int(bool(self.log_scale)), ctypes.cast(fb.data_ptr(), fp), ctypes.cast(win.data_ptr(), fp))
                    out = ctypes.c_void_p()
                    _lib.check(lib.bhmel_create(ctypes.byref(prm), ctypes.byref(out)))
                    h = out.value
                    self._handles[idx] = h
                    if not self._bulk:
                        _lib.check(lib.bhmel_set_option(h, _lib.OPT_BULK_COPY, 0))
                    _lib.check(lib.bhmel_set_option(h, _lib.OPT_KERNEL, self._variant))
                    if int(getattr(self, "_static_mel", 1)) != 1:
                        _lib.check(lib.bhmel_set_option(h, _lib.OPT_STATIC_MEL, int(self._static_mel)))
                    if not getattr(self, "_pdl", True):
                        _lib.check(lib.bhmel_set_option(h, _lib.OPT_PDL, 0))
                else:   # buffers were reloaded / edited: refresh the device tables
                    _lib.check(lib.bhmel_set_fb(h, ctypes.cast(fb.data_ptr(), fp)))
                    _lib.check(lib.bhmel_set_window(h, ctypes.cast(win.data_ptr(), fp)))
            self._stamp[idx] = stamp
            return h

    def __del__(self):
        try:
            lib = _lib.lib()
            for h in self._handles.values():
                lib.bhmel_destroy(h)
        except Exception:
            pass

    def launch_count(self) -> int:
        """Total kernel launches issued by this module (all devices)."""
        lib = _lib.lib()
        return sum(int(lib.bhmel_launch_count(h)) for h in self._handles.values())

    def set_bulk_copy(self, enabled: bool) -> None:
        """Debug/tuning switch: stage interior tiles with the TMA bulk copy (default) or not."""
        self._bulk = bool(enabled)
        for h in self._handles.values():
            _lib.check(_lib.lib().bhmel_set_option(h, _lib.OPT_BULK_COPY, int(enabled)))

    def set_kernel_variant(self, variant: str) -> None:
        """'ws' (default: warp-specialised FFT / mel roles), 'barrier' (stage-by-stage CTA schedule)
        or 'warp' (independent per-warp pipelines); all three produce bit-identical results."""
        self._variant = {"warp": _lib.KERNEL_INDEPENDENT_WARPS, "barrier": _lib.KERNEL_BARRIER,
                         "ws": _lib.KERNEL_WARP_SPECIALIZED}[variant]
        for h in self._handles.values():
            _lib.check(_lib.lib().bhmel_set_option(h, _lib.OPT_KERNEL, self._variant))

    def set_static_mel(self, enabled) -> None:
        """Debug / A-B switch: False / 0 forces the generic mel stage even for the baked reference
        filterbanks, True / 1 is the default (direct generated forms), 2 gives P0 its hybrid form, which
        is bit-identical to the generic stage (see BHMEL_OPT_STATIC_MEL in include/bhmel.h)."""
        self._static_mel = int(enabled)
        for h in self._handles.values():
            _lib.check(_lib.lib().bhmel_set_option(h, _lib.OPT_STATIC_MEL, int(self._static_mel)))

    @staticmethod
    def host_chunk_plan(batch: int, n_samples: int, pcm16: bool = False) -> list:
        """Rows per chunk of one forward_host call (bhmel_host_chunk_plan, include/bhmel.h): what a benchmark
        needs to time plain copies in the entry's own pattern."""
        lib = _lib.lib()
        n = int(lib.bhmel_host_chunk_plan(batch, n_samples, _lib.IN_PCM16 if pcm16 else _lib.IN_F32, None, 0))
        buf = (ctypes.c_int64 * max(n, 1))()
        lib.bhmel_host_chunk_plan(batch, n_samples, _lib.IN_PCM16 if pcm16 else _lib.IN_F32, buf, n)
        return [int(buf[i]) for i in range(n)]

    def set_pdl(self, enabled: bool) -> None:
        """A-B switch: programmatic dependent launch of the fused kernel (default on; BHMEL_OPT_PDL in
        include/bhmel.h).  Results and stream-order semantics do not depend on it."""
        self._pdl = bool(enabled)
        for h in self._handles.values():
            _lib.check(_lib.lib().bhmel_set_option(h, _lib.OPT_PDL, int(self._pdl)))

    # -- forward -----------------------------------------------------------------------
    def _check_input(self, samples: torch.Tensor) -> torch.Tensor:
        if samples.dim() != 2:
            # the reference fails in permute(0, 2, 1) for anything but [batch, samples]
            raise RuntimeError(f"expected samples of shape [batch, n_samples], got {tuple(samples.shape)}")
        if samples.shape[0] == 0 or samples.shape[1] == 0:
            raise RuntimeError("empty input")
        if self.pad_mode == "reflect" and samples.shape[1] <= self.n_fft // 2:
            # torch.stft's F.pad(..., "reflect") raises for the reference
            raise RuntimeError(
                f"Padding size should be less than the corresponding input dimension, but got: padding "
                f"({self.n_fft // 2}, {self.n_fft // 2}) at dimension 1 of input {list(samples.shape)}")
        return samples

    def _launch(self, samples: torch.Tensor) -> torch.Tensor:
        B, N = samples.shape
        x = samples
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if x.stride(1) != 1 or (B > 1 and x.stride(0) < N):
            x = x.contiguous()
        stride = x.stride(0) if B > 1 else N
        T = N // self.hop_length + 1
        y = torch.empty((B, T, self.n_mels), dtype=torch.float32, device=x.device)
        h = self._handle_for(x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().bhmel_forward(h, x.data_ptr(), B, N, stride, y.data_ptr(), stream))
        return y

    def forward(self, samples: torch.Tensor) -> torch.Tensor:
        """Convert a batch of audio windows [batch, n_samples] into log-mel frames
        [batch, n_samples // hop_length + 1, n_mels] (reference spectrogram.py:63-83)."""
        self._check_input(samples)
        if not samples.is_cuda:
            raise RuntimeError(
                "beatheritage_b200.MelSpectrogram has no CPU path: move the batch to a CUDA (sm_100) device, "
                "or use forward_host() for host-resident batches")
        if (type(samples) is torch.Tensor and not samples.requires_grad and not torch.compiler.is_compiling()
                and not torch._C._is_tracing() and torch._C._len_torch_dispatch_stack() == 0):
            # eager call on a plain tensor: same launch without the dispatcher round trip of the custom op
            # (about half of the per-call host time for one model-context window)
            return self._launch(samples)
        return _mel_forward_op(samples, self._key, self.n_mels)

    @torch.no_grad()
    def forward_into(self, samples: torch.Tensor, out: torch.Tensor, channel_offset: int = 0) -> torch.Tensor:
        """Write the mel frames of `samples` [B, N] into `out[:, :, channel_offset:channel_offset+n_mels]`
        where `out` is a float32 or bfloat16 CUDA tensor [B, N // hop + 1, C >= n_mels] whose last dim
        is contiguous.  This is the reference's `frames.to(dtype)` + `torch.cat([frames, cond...])`
        (modeling_mapperatorinator.py:352, 369-370) without the extra passes: allocate the encoder
        input once, let the frontend fill the mel channels, fill the conditioning channels yourself."""
        self._check_input(samples)
        if not (samples.is_cuda and out.is_cuda and out.device == samples.device):
            raise RuntimeError("forward_into expects CUDA tensors on the same device")
        B, N = samples.shape
        T = N // self.hop_length + 1
        if out.dim() != 3 or out.shape[0] != B or out.shape[1] != T or out.stride(2) != 1:
            raise RuntimeError(f"out must be [B={B}, T={T}, C] with a contiguous last dim, got {tuple(out.shape)}")
        if not (0 <= channel_offset and channel_offset + self.n_mels <= out.shape[2]):
            raise RuntimeError("channel_offset + n_mels exceeds out's channel count")
        if out.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("out must be float32 or bfloat16")
        x = samples if samples.dtype == torch.float32 else samples.to(torch.float32)
        if x.stride(1) != 1 or (B > 1 and x.stride(0) < N):
            x = x.contiguous()
        h = self._handle_for(x.device)
        desc = _lib.BhmelOutDesc(out.data_ptr() + channel_offset * out.element_size(),
                                 _lib.OUT_BF16 if out.dtype == torch.bfloat16 else _lib.OUT_F32,
                                 out.stride(1), out.stride(0) if B > 1 else T * out.stride(1))
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().bhmel_forward_ex(h, x.data_ptr(), B, N, x.stride(0) if B > 1 else N,
                                                   ctypes.byref(desc), stream))
        return out

    @torch.no_grad()
    def forward_encoder_input(self, samples: torch.Tensor, conds=(), dtype: torch.dtype = torch.bfloat16,
                              channels_first: bool = False) -> torch.Tensor:
        """The encoder input the reference assembles after the frontend
        (modeling_mapperatorinator.py:351-352, 368-376), in one go:

            frames = self(samples).to(dtype)
            x = torch.concatenate([frames] + [c.unsqueeze(1).expand(-1, T, -1) for c in conds], dim=-1)
            x = torch.swapaxes(x, 1, 2).contiguous()          # only if channels_first

        `conds`: the conditioning embeddings, each [B, D_i] (any float dtype; converted to `dtype`
        like torch.concatenate would require them to be).  Returns [B, T, C] or [B, C, T] with
        C = n_mels + sum(D_i); bit-identical to the torch ops above."""
        self._check_input(samples)
        if not samples.is_cuda:
            raise RuntimeError("forward_encoder_input expects a CUDA tensor")
        if dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("dtype must be float32 or bfloat16")
        B, N = samples.shape
        T = N // self.hop_length + 1
        conds = [conds] if isinstance(conds, torch.Tensor) else list(conds)
        for c in conds:
            if c.dim() != 2 or c.shape[0] != B or c.device != samples.device:
                raise RuntimeError("every conditioning tensor must be [B, D] on the samples' device")
        cond = torch.cat([c.to(dtype) for c in conds], dim=1).contiguous() if conds else None
        n_cond = cond.shape[1] if cond is not None else 0
        C = self.n_mels + n_cond
        out = torch.empty((B, C, T) if channels_first else (B, T, C), dtype=dtype, device=samples.device)
        x = samples if samples.dtype == torch.float32 else samples.to(torch.float32)
        if x.stride(1) != 1 or (B > 1 and x.stride(0) < N):
            x = x.contiguous()
        h = self._handle_for(x.device)
        # channels first stages [B, T, n_mels] first; the torch allocator keeps that stream-ordered / thread-safe
        scratch = torch.empty((B, T, self.n_mels), dtype=dtype, device=x.device) if channels_first else None
        desc = _lib.BhmelEncoderInputDesc(out.data_ptr(), _lib.OUT_BF16 if dtype == torch.bfloat16 else _lib.OUT_F32,
                                          _lib.LAYOUT_BCT if channels_first else _lib.LAYOUT_BTC,
                                          cond.data_ptr() if cond is not None else None, n_cond,
                                          scratch.data_ptr() if scratch is not None else None)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().bhmel_forward_encoder_input(h, x.data_ptr(), B, N, x.stride(0) if B > 1 else N,
                                                              ctypes.byref(desc), stream))
        return out

    @torch.no_grad()
    def forward_host(self, samples: torch.Tensor, out: torch.Tensor | None = None, device: int | None = None,
                     scales: torch.Tensor | None = None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """End-to-end call for HOST batches (reference osuT5/dataloading.py:128-130 calls the module
        with CPU tensors): chunked H2D copy, kernel and D2H copy overlapped inside the library.

        samples  CPU [B, N], float32 -- or int16 PCM, converted on the device as
                 float32(pcm) * scales[b] (the reference's cast + peak normalisation,
                 osuT5/osuT5/dataset/data_utils.py:95-97), so only 2 bytes/sample cross PCIe;
        scales   CPU float32 [B] for int16 input (None: 1.0);
        out      optional CPU result buffer [B, T, n_mels], float32 or bfloat16 (pinned recommended).
        Returns the CPU tensor."""
        self._check_input(samples)
        if samples.is_cuda:
            raise RuntimeError("forward_host expects a CPU tensor")
        pcm = samples.dtype == torch.int16
        x = samples if (pcm or samples.dtype == torch.float32) else samples.to(torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        B, N = x.shape
        T = N // self.hop_length + 1
        if out is None:
            out = torch.empty((B, T, self.n_mels), dtype=out_dtype, pin_memory=True)
        if not (out.is_contiguous() and tuple(out.shape) == (B, T, self.n_mels)
                and out.dtype in (torch.float32, torch.bfloat16) and not out.is_cuda):
            raise RuntimeError("out must be a contiguous CPU float32/bfloat16 tensor [B, T, n_mels]")
        sc_ptr = None
        if pcm and scales is not None:
            scales = scales.to(torch.float32).contiguous()
            if scales.numel() != B or scales.is_cuda:
                raise RuntimeError("scales must be a CPU tensor with one entry per row")
            sc_ptr = scales.data_ptr()
        idx = torch.cuda.current_device() if device is None else device
        h = self._handle_for(torch.device("cuda", idx))
        io = _lib.BhmelHostIO(x.data_ptr(), _lib.IN_PCM16 if pcm else _lib.IN_F32, sc_ptr, out.data_ptr(),
                              _lib.OUT_BF16 if out.dtype == torch.bfloat16 else _lib.OUT_F32)
        with torch.cuda.device(idx):
            _lib.check(_lib.lib().bhmel_forward_host_ex(h, ctypes.byref(io), B, N, x.stride(0) if B > 1 else N))
        return out

    @torch.no_grad()
    def peak_scale(self, pcm: torch.Tensor) -> torch.Tensor:
        """`1.0 / max|pcm|` of a device-resident int16 song as a float32 CUDA scalar tensor [1] -- the
        factor of the reference's `samples *= 1.0 / np.max(np.abs(samples))`
        (osuT5/osuT5/dataset/data_utils.py:94-96), reduced on the device."""
        if pcm.dim() != 1 or not pcm.is_cuda or pcm.dtype != torch.int16:
            raise RuntimeError("peak_scale expects a 1-D int16 CUDA tensor")
        pcm = pcm.contiguous()
        scale = torch.empty(1, dtype=torch.float32, device=pcm.device)
        h = self._handle_for(pcm.device)
        with torch.cuda.device(pcm.device):
            stream = torch.cuda.current_stream(pcm.device).cuda_stream
            _lib.check(_lib.lib().bhmel_peak_scale_pcm16(h, pcm.data_ptr(), pcm.numel(), scale.data_ptr(), stream))
        return scale

    def forward_gather(self, song: torch.Tensor, first_offset: int, stride: int, n_windows: int,
                       window_len: int, normalize=False) -> torch.Tensor:
        """Fused segmentation + forward: window w covers song[first_offset + w*stride : ... + window_len]
        with zeros past the end of `song` -- what Preprocessor.segment materialises on the host
        (reference osuT5/osuT5/inference/preprocessor.py:58-71, 94-102) -- without ever building
        the [W, window_len] batch.

        `song`: 1-D CUDA tensor, float32, or int16 PCM kept resident at 2 bytes per sample.  For int16,
        `normalize` selects the reference loader's peak normalisation (data_utils.py:94-96): True
        reduces the peak on the device, a float32 CUDA tensor [1] supplies the scale, False uses 1.0;
        sample i enters the transform as float32(pcm[i]) * scale."""
        if song.dim() != 1 or not song.is_cuda:
            raise RuntimeError("forward_gather expects a 1-D CUDA tensor")
        if self.pad_mode == "reflect" and window_len <= self.n_fft // 2:
            raise RuntimeError("window_len too short for reflect padding")
        T = window_len // self.hop_length + 1
        if song.dtype == torch.int16:
            x = song.contiguous()
            if isinstance(normalize, torch.Tensor):
                if normalize.dtype != torch.float32 or normalize.device != x.device or normalize.numel() != 1:
                    raise RuntimeError("normalize must be a float32 tensor with one element on the song's device")
                scale = normalize.contiguous()
            else:
                scale = self.peak_scale(x) if normalize else None
            y = torch.empty((n_windows, T, self.n_mels), dtype=torch.float32, device=x.device)
            scratch = torch.empty(x.numel(), dtype=torch.float32, device=x.device)   # stream-ordered, thread-safe
            h = self._handle_for(x.device)
            with torch.cuda.device(x.device):
                stream = torch.cuda.current_stream(x.device).cuda_stream
                _lib.check(_lib.lib().bhmel_forward_gather_pcm16(
                    h, x.data_ptr(), x.numel(), scale.data_ptr() if scale is not None else None, first_offset,
                    stride, n_windows, window_len, y.data_ptr(), scratch.data_ptr(), stream))
            return y
        if normalize is not False:
            raise RuntimeError("normalize applies to int16 songs only")
        x = song if song.dtype == torch.float32 else song.to(torch.float32)
        x = x.contiguous()
        y = torch.empty((n_windows, T, self.n_mels), dtype=torch.float32, device=x.device)
        h = self._handle_for(x.device)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream(x.device).cuda_stream
            _lib.check(_lib.lib().bhmel_forward_gather(h, x.data_ptr(), x.numel(), first_offset, stride,
                                                       n_windows, window_len, y.data_ptr(), stream))
        return y

    def extra_repr(self) -> str:
        return (f"implementation={self.implementation!r}, log_scale={self.log_scale}, sample_rate={self.sample_rate}, "
                f"n_ftt={self.n_fft}, n_mels={self.n_mels}, hop_length={self.hop_length}, f_min={self.f_min}, "
                f"f_max={self.f_max}, pad_mode={self.pad_mode!r}")
