"""A/B of the CTA-pair kernel's epilogue width (BHSTEM_OPT_EPILOGUE_WARPS) on the C5 shapes: the full stem
(464 channels) and the split stem (80 + 384 folded), 8 against 16 epilogue warps per stage.
`python tools/stem_split_ab.py [B ...]`; CUDA-event medians, same bits checked."""
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
for B in [int(a) for a in sys.argv[1:]] or [6, 16, 46]:
    x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
    x[:, :, 80:] = x[:, :1, 80:]
    frames, cond = x[:, :, :80].contiguous(), x[:, 0, 80:].contiguous()
    hid = torch.empty(B, 4096, 768, dtype=torch.bfloat16, device=dev)
    out = torch.empty(B, 2048, 768, dtype=torch.bfloat16, device=dev)
    res, outs = {}, {}
    for name, (e1, e2, es) in {"8/8/8": (8, 8, 8), "16/16/16": (16, 16, 16)}.items():
        stem.set_epilogue_warps(e1, e2, es)
        res[name] = (timed(lambda: stem.forward_stage(1, x)), timed(lambda: stem.forward_stage(2, hid)),
                     timed(lambda: stem.forward_split(frames, cond, hidden=hid, out=out)),
                     timed(lambda: stem(x, hidden=hid, out=out)))
        outs[name] = (stem.forward_split(frames, cond).clone(), stem(x).clone())
    same = all(torch.equal(a, b) for a, b in zip(outs["8/8/8"], outs["16/16/16"]))
    for name, (c1, c2, sp, full) in res.items():
        print(f"B={B:3d} epilogue warps {name:9s} conv1 {c1:.4f}  conv2 {c2:.4f}  full stem {full:.4f}  split stem {sp:.4f} "
              f"(split conv1 ~ {sp - c2:.4f}) ms")
    print(f"B={B:3d} same bits: {same}")
