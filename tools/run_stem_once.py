"""A few launches of the conv stem on the C5 shapes, for ncu (`-k regex:bhstem`): B windows x 4096
frames x 464 channels -> 768.  `python tools/run_stem_once.py [B] [reps]`; prints the median ms."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
torch.manual_seed(0)
stem = ConvStem(464, 768).to(dev)
if "nosmall" in sys.argv:
    stem.set_small_batch_tiles(False)
x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = stem(x)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
flop = 2.0 * B * 4096 * 768 * 3 * 464 + 2.0 * B * 2048 * 768 * 3 * 768
ms = sorted(ts)[len(ts) // 2]
print(f"B={B} median {ms:.4f} ms  {flop / ms / 1e9:.1f} TFLOP/s  finite={bool(torch.isfinite(y.float()).all())}")
