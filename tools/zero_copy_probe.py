"""Would writing the mel frames straight to pinned host memory from the kernel (zero copy) beat the copy
engine's D2H while the H2D copies of the next chunks run?  One bench step's bytes: 256 windows.
  (a) H2D only (b) H2D + copy-engine D2H (the host entry's pattern) (c) H2D + kernel writing to mapped host memory
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram, _lib  # noqa: E402

dev = torch.device("cuda", 0)
B, N, T, M = 256, 524160, 4096, 80
mel = MelSpectrogram("torchaudio", True, 16000, 1024, M, 128, 20, 8000, "reflect").to(dev)
h_in = torch.rand(B, N).mul_(2).sub_(1).pin_memory()
h_out = torch.empty(B, T, M, pin_memory=True)
d_in = torch.empty(B, N, device=dev)
d_out = torch.empty(B, T, M, device=dev)
rows = 16
streams = [torch.cuda.Stream() for _ in range(3)]
handle = mel._handle_for(dev)
lib = _lib.lib()
fp = ctypes.POINTER(ctypes.c_float)


def step(mode):
    for k, b0 in enumerate(range(0, B, rows)):
        s = streams[k % 3]
        with torch.cuda.stream(s):
            d_in[b0:b0 + rows].copy_(h_in[b0:b0 + rows], non_blocking=True)
            if mode == "h2d":
                continue
            if mode == "zero_copy":      # kernel output pointer = mapped pinned host memory
                yptr = h_out.data_ptr() + b0 * T * M * 4
                _lib.check(lib.bhmel_forward(handle, ctypes.cast(d_in.data_ptr() + b0 * N * 4, fp), rows, N, N,
                                             ctypes.cast(yptr, fp), ctypes.c_void_p(s.cuda_stream)))
            else:
                _lib.check(lib.bhmel_forward(handle, ctypes.cast(d_in.data_ptr() + b0 * N * 4, fp), rows, N, N,
                                             ctypes.cast(d_out.data_ptr() + b0 * T * M * 4, fp), ctypes.c_void_p(s.cuda_stream)))
                h_out[b0:b0 + rows].copy_(d_out[b0:b0 + rows], non_blocking=True)


ref = None
for mode in ("h2d", "copy_engine", "zero_copy", "copy_engine", "zero_copy"):
    step(mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step(mode)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    extra = ""
    if mode != "h2d":
        if ref is None:
            ref = h_out.clone()
        extra = f"  same result: {bool(torch.equal(h_out, ref))}"
    print(f"{mode:12s} {dt * 1e3:7.2f} ms per step  -> {B * N / 16000 / dt / 1e6:.3f} M audio-s/s{extra}", flush=True)
