#!/usr/bin/env python3
"""A/B of programmatic dependent launch (BHMEL_OPT_PDL / BHSTEM_OPT_PDL): back-to-back launches of the
fused frontend kernel and of the conv stem, timed with CUDA events around the whole loop.
    python tools/pdl_ab.py [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402

P0 = ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect")


def loop_ms(fn, reps, rounds=5):
    for _ in range(max(3, reps // 10)):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        best.append(a.elapsed_time(b) / reps)
    best.sort()
    return best[len(best) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    mel = MelSpectrogram(*P0).to(dev)
    for W in (1, 6, 46, 256):
        x = torch.rand(W, 524160, device=dev) * 2 - 1
        out = torch.empty(W, 4096, 80, device=dev)
        reps = 400 if W <= 6 else (100 if W <= 46 else 40)
        r = {"what": "frontend P0", "windows": W}
        for on in (True, False, True, False):
            mel.set_pdl(on)
            r.setdefault("pdl_ms" if on else "plain_ms", []).append(loop_ms(lambda: mel.forward_into(x, out), reps))
        mel.set_pdl(True)
        rows.append(r)
        print(json.dumps(r), flush=True)
    torch.manual_seed(0)
    stem = ConvStem(464, 768).to(dev)
    for B in (1, 6, 16, 46):
        x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
        hidden = torch.empty(B, 4096, 768, dtype=torch.bfloat16, device=dev)
        out = torch.empty(B, 2048, 768, dtype=torch.bfloat16, device=dev)
        flop = 2.0 * B * 4096 * 768 * 3 * 464 + 2.0 * B * 2048 * 768 * 3 * 768
        r = {"what": "conv stem", "windows": B}
        for on in (True, False, True, False):
            stem.set_pdl(on)
            ms = loop_ms(lambda: stem(x, hidden=hidden, out=out), 200 if B <= 6 else 50)
            r.setdefault("pdl_ms" if on else "plain_ms", []).append(ms)
            r.setdefault("pdl_tflops" if on else "plain_tflops", []).append(flop / ms / 1e9)
        stem.set_pdl(True)
        rows.append(r)
        print(json.dumps(r), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"gpu": torch.cuda.get_device_name(0), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
