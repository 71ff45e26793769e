"""Diagnose run-to-run variance of bench.py's timed loop: per-step CUDA-event times with and without
the NVML sampler thread."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from beatheritage_b200 import MelSpectrogram

dev = torch.device("cuda", 0)
mel = MelSpectrogram(*bench.P0).to(dev)
xs = []
for b in range(4):
    g = torch.Generator(device=dev).manual_seed(1234 + b)
    xs.append(torch.rand(bench.BATCH, bench.WINDOW, device=dev, generator=g).mul_(2).sub_(1))
torch.cuda.synchronize()
def loop(n, sampler):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ctx = bench.ClockSampler(0) if sampler else None
    if ctx: ctx.__enter__()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(n):
        y = mel(xs[i % 4])
        evs[i + 1].record()
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    if ctx: ctx.__exit__()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
    s = sorted(ts)
    print(f"sampler={sampler} n={n} enqueue {1e3*t_enq:.1f} ms  total {sum(ts):.1f} ms  mean {sum(ts)/n:.3f}  median {s[n//2]:.3f}  "
          f"p90 {s[int(.9*n)]:.3f}  max {s[-1]:.3f}  first10 {[round(t,2) for t in ts[:10]]}", flush=True)
    if ctx: print("   clocks", ctx.summary(), flush=True)
for i in range(10): mel(xs[i % 4])
loop(200, True)
loop(200, False)
loop(200, True)
