"""A/B of BHSTEM_OPT_DEEP_A_RING (3 activation + 6 weight stages against 2 + 8) for conv1 and conv2 of the full stem,
alternating in one process; CUDA-event medians, same bits checked.  `python tools/stem_ring_ab.py [B ...]`"""
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
for B in [int(a) for a in sys.argv[1:]] or [16, 46]:
    x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
    hid = stem.forward_stage(1, x)
    outs = {}
    for rep in range(3):
        for mask in (0, 3):
            stem.set_deep_a_ring(mask)
            c1 = timed(lambda: stem.forward_stage(1, x), 5, 40)
            c2 = timed(lambda: stem.forward_stage(2, hid), 5, 40)
            outs[mask] = stem(x).clone()
            print(f"B={B:3d} rings { {0: '2A+8W (conv1: 3 A)', 3: '3A+6W (conv1: 5 A)'}[mask] }: conv1 {c1:.4f}  conv2 {c2:.4f} ms")
    print(f"B={B:3d} same bits: {torch.equal(outs[0], outs[3])}")
