"""A/B of BHSTEM_OPT_SMALL_BATCH_TILES on 1-3 windows (C5 sequential serving): full and split stem, CUDA-event medians."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402
from tools.bench_stem import timed  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
stem = ConvStem(464, 768).to(dev)
for B in (1, 2, 3):
    x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
    frames, cond = x[:, :, :80].contiguous(), x[:, 0, 80:].contiguous()
    hid = torch.empty(B, 4096, 768, dtype=torch.bfloat16, device=dev)
    out = torch.empty(B, 2048, 768, dtype=torch.bfloat16, device=dev)
    for on in (False, True, False, True):
        stem.set_small_batch_tiles(on)
        c1 = timed(lambda: stem.forward_stage(1, x), 10, 100)
        c2 = timed(lambda: stem.forward_stage(2, hid), 10, 100)
        full = timed(lambda: stem(x, hidden=hid, out=out), 10, 100)
        split = timed(lambda: stem.forward_split(frames, cond, hidden=hid, out=out), 10, 100)
        print(f"B={B} small-batch tiles {'on ' if on else 'off'}: conv1 {c1:.4f}  conv2 {c2:.4f}  full {full:.4f}  split {split:.4f} ms")
