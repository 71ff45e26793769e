"""Randomised cross-check of the three kernel schedules (bit-identical by construction with the generic /
mel stage; the warp-specialised kernel's direct mel stages -- 80 / 128 / 388 mels here -- are held to a few
ulp of them) over many
shapes: ragged lengths, batch sizes around multiples of the SM count, both pad modes, several
filterbanks, module and gather mode, aligned and unaligned rows.  Catches pipeline (mbarrier
parity / tile hand-off) bugs that only show for particular tile counts.

    python tools/fuzz_variants.py [--seconds 60] [--seed 0]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
dev = torch.device("cuda", 0)
mods = {}
for n_mels, pad, log, fmin in ((80, "reflect", True, 20), (80, "constant", True, 20), (128, "reflect", True, 20),
                               (388, "constant", False, 0), (33, "reflect", True, 0)):
    mods[(n_mels, pad)] = MelSpectrogram("torchaudio", log, 16000, 1024, n_mels, 128, fmin, 8000, pad).to(dev)
keys = list(mods)
DIRECT = {80: True, 128: True, 388: False}      # n_mels whose default ws mel stage is a direct form -> log scale?


def close(y, ref, key, what):
    """Bit-identical unless the set's default stage is a direct form; then a few ulp."""
    if key[0] not in DIRECT:
        assert torch.equal(y, ref), what
    elif DIRECT[key[0]]:
        assert float((y - ref).abs().max()) <= 2e-6, what
    else:
        assert bool(((y - ref).abs() <= 4e-6 * ref.abs()).all()), what


t_end = time.time() + a.seconds
n_cases = n_frames = 0
while time.time() < t_end:
    key = keys[rng.integers(len(keys))]
    m = mods[key]
    lo = 513 if key[1] == "reflect" else 1
    kind = rng.integers(4)
    if kind == 0:      # many short rows: tile counts around multiples of 148
        B, N = int(rng.integers(1, 700)), int(rng.integers(lo, 9000))
    elif kind == 1:    # few long rows
        B, N = int(rng.integers(1, 6)), int(rng.integers(100000, 700000))
    else:              # medium
        B, N = int(rng.integers(1, 64)), int(rng.integers(lo, 70000))
    off = int(rng.integers(0, 4))
    stride = N + int(rng.integers(0, 5))
    base = torch.rand(B * stride + 8, device=dev) * 2 - 1
    x = base[off:off + B * stride].view(B, stride)[:, :N]
    outs = []
    for variant in ("barrier", "ws", "warp"):
        m.set_kernel_variant(variant)
        if variant == "ws":
            m.set_static_mel(False)
        outs.append(m(x))
        m.set_static_mel(True)
    m.set_kernel_variant("ws")
    y_default = m(x)          # the shipped configuration: static / direct mel stage where one is baked
    if kind == 3:      # gather mode over the same buffer
        wlen = int(rng.integers(lo, 40000))
        gstride = int(rng.integers(1, 30000))
        W = int(rng.integers(1, 40))
        first = int(rng.integers(0, 1000))
        gouts = []
        for variant in ("barrier", "ws", "warp"):
            m.set_kernel_variant(variant)
            if variant == "ws":
                m.set_static_mel(False)
            gouts.append(m.forward_gather(base, first, gstride, W, wlen))
            m.set_static_mel(True)
        m.set_kernel_variant("ws")
        close(m.forward_gather(base, first, gstride, W, wlen), gouts[0], key, ("gather default", key, first, gstride, W, wlen))
        torch.cuda.synchronize()
        assert torch.equal(gouts[0], gouts[1]) and torch.equal(gouts[0], gouts[2]), ("gather", key, first, gstride, W, wlen)
        assert torch.isfinite(gouts[1]).all()
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]), ("ws != barrier", key, B, N, off, stride)
    assert torch.equal(outs[0], outs[2]), ("warp != barrier", key, B, N, off, stride)
    assert torch.isfinite(outs[1]).all()
    close(y_default, outs[0], key, ("ws default != barrier", key, B, N, off, stride))
    n_cases += 1
    n_frames += B * (N // 128 + 1)
    m.set_kernel_variant("ws")
print(f"fuzz ok: {n_cases} cases, {n_frames} frames, three schedules bit-identical, direct stages within a few ulp")
