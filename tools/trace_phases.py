#!/usr/bin/env python3
"""Phase timeline of one SM of the warp-specialised kernel (debug build with -DBHMEL_TRACE).

    BHMEL_EXTRA_NVCC=-DBHMEL_TRACE python -m beatheritage_b200.build --force   # then, on the GPU box:
    python tools/trace_phases.py [--out gpurun_out/trace.json]

Block 0 records clock64() at the phase boundaries of 4 consecutive tiles (iterations 64..67) for each
of its 16 warps; this script runs one bench-sized launch, fetches the table and prints per-phase
durations.  FFT warps: ev0 tile start, then per pair q=0,1: 1+5q after loads+pass A, 2+5q after the
transposes, 3+5q after pass B, 4+5q after separation + P stores.  Mel warps: ev0 before / ev1 after the
p_full wait, ev2 after the mel dot products, ev3 staging barrier, ev4 after the global stores, ev5 end.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.json"))
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--static-mel", type=int, default=1, help="BHMEL_OPT_STATIC_MEL (0: generic mel stage)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
mel = MelSpectrogram("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect").to(dev)
mel.set_static_mel(args.static_mel)
x = torch.rand(args.batch, 524160, device=dev) * 2 - 1
for _ in range(3):
    y = mel(x)
torch.cuda.synchronize()
lib = _lib.lib()
fn = lib.bhmel_debug_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn.restype = ctypes.c_int
n = 16 * 4 * 16
buf = np.zeros(n, dtype=np.int64)
rc = fn(buf.ctypes.data, n)
assert rc == 0, rc
t = buf.reshape(16, 4, 16)
base = t[:, 0, 0][t[:, 0, 0] > 0].min()
rel = np.where(t > 0, t - base, -1)
os.makedirs(os.path.dirname(args.out), exist_ok=True)
json.dump(rel.tolist(), open(args.out, "w"))
print("FFT warps (cycles rel. to first event): it, ev0, [A, T, B, S] x 2 pairs")
for w in range(8):
    for i in range(4):
        print(f"w{w} it{i}", " ".join(f"{v:7d}" for v in rel[w, i, :11]))
print("mel warps: it, before wait, after wait, dots, staged, stored, end")
for w in range(8, 16):
    for i in range(4):
        print(f"w{w} it{i}", " ".join(f"{v:7d}" for v in rel[w, i, :6]))
