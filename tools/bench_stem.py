"""Times the tcgen05 conv stem (include/bhstem.h) against torch's own bf16 Conv1d + gelu (cuDNN) on
the C5 shapes: whisper-small dims, 464 input channels, 4096 frames.

    python tools/bench_stem.py [--json out.json]

CUDA-event times, 5 warm-up + 30 timed calls each; prints TFLOP/s against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402


def timed(fn, warm=5, reps=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    ap.add_argument("--variant", default="cta_pairs", choices=["shared_taps", "tap_boxes", "cta_pairs"])
    ap.add_argument("--no-pdl", action="store_true", help="plain stream order instead of programmatic dependent launch")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("bf16_tflops", 1640.9))
    torch.manual_seed(0)
    C, D, T = 464, 768, 4096
    stem = ConvStem(C, D).to(dev)
    stem.set_variant(args.variant)
    if args.no_pdl:
        stem.set_pdl(False)
    ref = ConvStem(C, D).to(dev).to(torch.bfloat16)
    ref.load_state_dict(stem.state_dict())
    out = {"gpu": torch.cuda.get_device_name(0), "dims": {"c_in": C, "d_model": D, "frames": T}, "rows": []}
    for B in (1, 6, 16, 46):
        x = (torch.randn(B, T, C, device=dev) * 1.5).to(torch.bfloat16)
        x_bct = x.swapaxes(1, 2).contiguous()

        def torch_stem():
            h = torch.nn.functional.gelu(ref.conv1(x_bct))
            return torch.nn.functional.gelu(ref.conv2(h)).permute(0, 2, 1)

        flop1 = 2.0 * B * T * D * 3 * C
        flop2 = 2.0 * B * (T // 2) * D * 3 * D
        ms = timed(lambda: stem(x))
        ms1 = timed(lambda: stem.forward_stage(1, x))
        h = stem.forward_stage(1, x)
        ms2 = timed(lambda: stem.forward_stage(2, h))
        ms_t = timed(torch_stem)
        # split conv1: 80 time-varying channels + 384 folded into a per-window bias (bhstem_forward_split)
        frames, cond = x[:, :, :80].contiguous(), x[:, 0, 80:].contiguous()
        hid = torch.empty(B, T, D, dtype=torch.bfloat16, device=dev)
        yo = torch.empty(B, T // 2, D, dtype=torch.bfloat16, device=dev)
        ms_split = timed(lambda: stem.forward_split(frames, cond, hidden=hid, out=yo))
        row = {"batch": B, "ours_ms": ms, "conv1_ms": ms1, "conv2_ms": ms2, "torch_cudnn_bf16_ms": ms_t,
               "ours_tflops": (flop1 + flop2) / ms / 1e9, "conv1_tflops": flop1 / ms1 / 1e9,
               "conv2_tflops": flop2 / ms2 / 1e9, "frac_of_measured_bf16_peak": (flop1 + flop2) / ms / 1e9 / peak,
               "speedup_vs_torch": ms_t / ms, "split_ms": ms_split, "split_conv1_ms": ms_split - ms2,
               "split_speedup_vs_full": ms / ms_split, "split_speedup_vs_torch": ms_t / ms_split}
        out["rows"].append(row)
        print(json.dumps(row))
    out["peak_bf16_tflops"] = peak
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
