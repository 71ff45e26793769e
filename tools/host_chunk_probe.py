#!/usr/bin/env python3
"""End-to-end host-buffer throughput (f32/f32 and int16/bf16) of whichever libbhmel BHMEL_LIB points at:
A/B of the host pipeline's slot count / chunk size (-DBHMEL_HOST_SLOTS, -DBHMEL_HOST_CHUNK_MB builds made
with tools/build_variant.sh).   BHMEL_LIB=build/libbhmel_x.so python tools/host_chunk_probe.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

P0 = ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect")
B, N, T, M = 256, 524160, 4096, 80


def main():
    dev = torch.device("cuda", 0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0))
    except Exception:
        pass
    mel = MelSpectrogram(*P0).to(dev)
    x = torch.rand(B, N, device=dev) * 2 - 1
    h_in = torch.empty(B, N, dtype=torch.float32, pin_memory=True)
    h_in.copy_(x)
    h_out = torch.empty(B, T, M, dtype=torch.float32, pin_memory=True)
    pcm = (x * 32767.0).to(torch.int16).cpu().pin_memory()
    scales = torch.full((B,), 1.0 / 32767.0)
    out16 = torch.empty(B, T, M, dtype=torch.bfloat16, pin_memory=True)
    res = {"lib": os.environ.get("BHMEL_LIB", "default")}
    for name, fn in (("f32", lambda: mel.forward_host(h_in, out=h_out)),
                     ("pcm16_bf16", lambda: mel.forward_host(pcm, out=out16, scales=scales))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(8):
                fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) / 8)
        ts.sort()
        res[name + "_ms"] = [round(t * 1e3, 3) for t in ts]
        res[name + "_audio_s_per_s"] = round(B * N / 16000 / ts[1])
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
