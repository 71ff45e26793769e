#!/usr/bin/env python3
""""one process with 8 devices / streams, or 8 processes -- measure both if cheap".
bench.py is the 8-process form (torchrun); this is the one-process form: ONE Python thread drives all
visible GPUs, one module handle, input and output shard per device, launches issued round-robin.
Device-resident C3 batch (256 windows per device per step), CUDA events per device, max over devices.
    python tools/one_process_multi_gpu.py [--steps 50] [--json out.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402

P0 = ("torchaudio", True, 16000, 1024, 80, 128, 20, 8000, "reflect")
B, N, SR = 256, 524160, 16000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--json")
    args = ap.parse_args()
    n_dev = torch.cuda.device_count()
    rows = []
    for n in [k for k in (1, 2, 4, 8) if k <= n_dev]:
        devs = [torch.device("cuda", i) for i in range(n)]
        mel = MelSpectrogram(*P0)
        xs, ys, mods = [], [], []
        for d in devs:
            g = torch.Generator(device=d).manual_seed(1234 + d.index)
            xs.append(torch.rand(B, N, device=d, generator=g).mul_(2).sub_(1))
            ys.append(torch.empty(B, N // 128 + 1, 80, device=d))
            mods.append(MelSpectrogram(*P0).to(d))
        for _ in range(5):
            for m, x, y in zip(mods, xs, ys):
                m.forward_into(x, y)
        for d in devs:
            torch.cuda.synchronize(d)
        ev = []
        for d in devs:
            with torch.cuda.device(d):
                ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
        for d, (a, _) in zip(devs, ev):
            with torch.cuda.device(d):
                a.record()
        for _ in range(args.steps):
            for m, x, y in zip(mods, xs, ys):
                m.forward_into(x, y)
        for d, (_, b) in zip(devs, ev):
            with torch.cuda.device(d):
                b.record()
        for d in devs:
            torch.cuda.synchronize(d)
        ms = max(a.elapsed_time(b) for a, b in ev) / args.steps
        ref = mods[0](xs[0][:1])
        same = all(torch.equal(m(x[:1].to(m_dev)).cpu(), MelSpectrogram(*P0).to(m_dev)(x[:1]).cpu())
                   for m, x, m_dev in zip(mods, xs, devs))
        row = {"devices": n, "ms_per_step_max_over_devices": ms, "audio_s_per_s": n * B * N / SR / (ms / 1e3),
               "per_device_audio_s_per_s": B * N / SR / (ms / 1e3), "deterministic_per_device": bool(same),
               "finite": bool(torch.isfinite(ref).all())}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del xs, ys, mods
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"what": "one process, one thread, N devices (SURVEY.md 8e)", "gpu": torch.cuda.get_device_name(0),
                       "windows_per_device_per_step": B, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
