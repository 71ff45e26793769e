"""A few launches of the split conv stem (bhstem_forward_split) on the C5 shapes, for ncu (`-k regex:bhstem`):
B windows x 4096 frames x (80 time-varying + 384 folded) channels -> 768.
`python tools/run_stem_split_once.py [B] [reps] [epilogue warps of the split conv1: 8 | 16]`; prints the median ms."""
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
if len(sys.argv) > 3 and sys.argv[3].isdigit():
    stem.set_epilogue_warps(8, 8, int(sys.argv[3]))
frames = (torch.randn(B, 4096, 80, device=dev) * 1.5).to(torch.bfloat16)
cond = (torch.randn(B, 384, device=dev) * 1.5).to(torch.bfloat16)
hid = torch.empty(B, 4096, 768, dtype=torch.bfloat16, device=dev)
out = torch.empty(B, 2048, 768, dtype=torch.bfloat16, device=dev)
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = stem.forward_split(frames, cond, hidden=hid, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"B={B} median {ms:.4f} ms  finite={bool(torch.isfinite(y.float()).all())}")
