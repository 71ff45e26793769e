#!/usr/bin/env python3
"""bench.py -- log-mel audio-seconds/sec of the B200 frontend (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (torch port)

Workload (SURVEY.md 8d, C3 "dataset preprocessing"): every step is one batch of 256 model-context
windows [256, 524160] float32 (P0: 80 mels, reflect, log1p) -> [256, 4096, 80]; inputs are
synthetic uniform(-1, 1) audio generated on the device from seed 1234 + 1000*rank + batch and
rotate over 4 distinct 537 MB buffers, so each step reads data far larger than the 126 MB L2.
Ranks process independent shards (weak scaling, no collective on the data path); time is the
max over ranks of the CUDA-event time of the K timed steps.

One JSON line is printed by rank 0.  `value` is device-resident throughput; `e2e` is the same
metric through the host-buffer entry (`MelSpectrogram.forward_host` -> bhmel_forward_host) with
the H2D and D2H copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WINDOW = 524160                  # (4096 - 1) * 128 samples = 32.76 s   (reference preprocessor.py:14-17)
SR = 16000
FRAMES = WINDOW // 128 + 1       # 4096
N_MELS = 80
BATCH = 256                      # windows per step
N_INPUT_BUFFERS = 4
METRIC = "log-mel audio-sec/sec"
UNIT = "audio-s/s"
P0 = ("torchaudio", True, SR, 1024, N_MELS, 128, 20, 8000, "reflect")
ALGO_BYTES_PER_WINDOW = 4 * WINDOW + 4 * FRAMES * N_MELS      # 832 B/frame, SURVEY.md 8d


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def fp32_roof(frames_per_s: float, clocks: dict) -> dict:
    """Achieved algorithmic fp32 throughput of the frontend against the CUDA-core FMA peak of the device at the SM
    clock sampled under load (148 SMs x 128 lanes x 2 flop; SURVEY.md 8d's second roofline)."""
    flop_per_frame = 30.5e3
    try:
        import torch
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        sms = 148
    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
    peak = sms * 128 * 2 * mhz * 1e6 / 1e12
    achieved = frames_per_s * flop_per_frame / 1e12
    return {"achieved_tflops": achieved, "peak_tflops": peak, "frac": achieved / peak, "flop_per_frame": flop_per_frame,
            "sm_mhz": mhz, "sms": sms}


def reduce_over_ranks(elapsed: float, units: float):
    """(max over ranks of elapsed, sum over ranks of units).  Works for gloo (CPU) and nccl."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return elapsed, units
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    u = torch.tensor([units], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and clock-event (throttle) reasons through NVML every few ms while
    the timed region runs (same fields as the profiling recipe's nvidia-smi clocks line)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.samples, self.stop_flag, self.thread = index, period_s, [], False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            # prime every query once: the first NVML calls of a process take tens of ms and hold a
            # driver lock that stalls kernel launches -- never let that happen inside a timed region
            for _ in range(3):
                pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
                pynvml.nvmlDeviceGetPowerUsage(self.handle)
                try:
                    pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        pmax = max(p for _, p, _ in self.samples)
        busy = [s for s in self.samples if s[1] >= 0.5 * pmax] or self.samples
        reasons = set()
        for _, _, r in busy:
            for bit, name in self.REASONS.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(s for s, _, _ in busy), "sm_max_mhz": self.sm_max,
                "reasons": sorted(reasons), "samples": len(self.samples), "samples_under_load": len(busy),
                "power_w_max": pmax}


# ------------------------------------------------------------------------------------------
def load_cpu_reference():
    """(module, kind): the reference's OWN MelSpectrogram module (oracle/_ref/spectrogram.py, copied
    unmodified by __graft_entry__.build(); kind "reference") on CPU with P0's arguments, else the torch
    port of the same operator sequence (oracle/torch_port.py; kind "port")."""
    try:
        from oracle import ref_loader
        ref = ref_loader.load_shipped_reference_class()(*P0)      # lazily imports torchaudio.transforms
        ref.eval()
        return ref, "reference"
    except Exception:      # copy absent (never built in the container) or torchaudio missing
        from oracle.torch_port import TorchPortMel
        return TorchPortMel(), "port"


def cpu_port_throughput(budget_s: float, windows: int, warmup: int = 1, min_passes: int = 3):
    """Times the reference's CPU path (the reference module itself when oracle/_ref holds it, else the
    torch port of torchaudio MelSpectrogram + log1p + permute) on this host with every core torch will
    use; returns (audio-s/s mean, best, passes, threads, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    port, kind = load_cpu_reference()
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(windows, WINDOW, generator=g) * 2 - 1
    for _ in range(warmup):
        port(x)
    times = []
    t_end = time.perf_counter() + budget_s
    with torch.no_grad():
        while len(times) < min_passes or time.perf_counter() < t_end:
            t0 = time.perf_counter()
            port(x)
            times.append(time.perf_counter() - t0)
            if len(times) >= 200:
                break
    audio = windows * WINDOW / SR
    return audio / statistics.mean(times), audio / min(times), len(times), torch.get_num_threads(), kind


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    windows = 16                                   # bounded sample of the 256-window step
    port, kind = load_cpu_reference()
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(windows, WINDOW, generator=g) * 2 - 1
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            port(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            port(x)
        dt = time.perf_counter() - t0
    value = args.steps * windows * WINDOW / SR / dt
    sample = f"{windows} of the {BATCH} windows of one step per step ([{windows}, {WINDOW}] f32), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("reference CPU path = the reference's own osuT5/osuT5/model/spectrogram.py (unmodified copy in "
                 "oracle/_ref, implementation=\"torchaudio\")" if kind == "reference" else
                 "reference CPU path = torchaudio MelSpectrogram arithmetic restated with torch ops "
                 "(oracle/torch_port.py; oracle/_ref is absent)") + "; runs on rank 0 only, all host threads",
    }
    emit(line)
    return 0


def workload_config(n_gpus):
    return {
        "workload": "C3 dataset-preprocessing batch: 256 model-context windows [256, 524160] f32 -> [256, 4096, 80] "
                    "per step per GPU, P0 (torchaudio arithmetic, 80 mels, f_min 20, reflect, log1p)",
        "windows_per_step_per_gpu": BATCH, "samples_per_window": WINDOW, "n_mels": N_MELS,
        "audio_seconds_per_step_per_gpu": BATCH * WINDOW / SR,
        "sharding": f"{n_gpus} independent rank(s), no data-path collective",
        "l2": "no flush needed: each step streams 537 MB in + 336 MB out, inputs rotate over 4 distinct buffers",
    }


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # bind this rank to the CPUs next to its GPU so pinned host buffers are NUMA-local (matters
    # for the host-buffer e2e leg when several ranks share the host); undone for the CPU baseline
    numa_bound = False
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_rank
        if visible:
            ids = [v for v in visible.split(",") if v.strip() != ""]
            if local_rank < len(ids) and ids[local_rank].strip().isdigit():
                phys = int(ids[local_rank])
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        numa_bound = True
    except Exception:
        pass

    from beatheritage_b200 import MelSpectrogram
    mel = MelSpectrogram(*P0).to(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # ---- synthetic inputs, generated on the device
    xs = []
    for b in range(N_INPUT_BUFFERS):
        g = torch.Generator(device=dev).manual_seed(1234 + 1000 * rank + b)
        xs.append(torch.rand(BATCH, WINDOW, device=dev, generator=g).mul_(2).sub_(1))
    torch.cuda.synchronize(dev)

    # ---- device-resident throughput ------------------------------------------------------
    # W warm-up steps as asked, then keep stepping (still untimed) until the GPU has been busy for
    # ~0.4 s so clocks, allocator and launch path are in steady state before the K timed steps
    for i in range(args.warmup):
        mel(xs[i % N_INPUT_BUFFERS])
    torch.cuda.synchronize(dev)
    t_warm = time.perf_counter()
    extra_warm = 0
    while time.perf_counter() - t_warm < 0.4:
        for i in range(20):
            mel(xs[i % N_INPUT_BUFFERS])
        torch.cuda.synchronize(dev)
        extra_warm += 20
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    for ev in [start, stop] + marks:     # events are created lazily at first record(): do that now,
        ev.record()                      # not inside the timed region
    torch.cuda.synchronize(dev)
    # The K timed steps are measured `--repeats` times (default 3), each time bracketed by barrier +
    # synchronize on both sides and timed with CUDA events on the launching stream; the MEDIAN
    # repeat is reported (all repeats are listed in the JSON line).  A fresh box occasionally
    # stalls a single launch for 25-150 ms (observed only in the first process after the box
    # comes up), which would otherwise decide the whole figure.
    clocks = ClockSampler(local_rank)
    repeats = []
    launches = 0
    for rep in range(max(1, args.repeats)):
        barrier()
        # untimed runway: a few steps queued ahead of the start event (no synchronisation in between)
        # so that the CPU has enqueued every timed step before the GPU reaches them
        runway = max(1, min(20, args.steps // 4))
        for i in range(runway):
            mel(xs[i % N_INPUT_BUFFERS])
        launches0 = mel.launch_count()
        start.record()
        marks[0].record()
        for i in range(args.steps):          # asynchronous launches: the CPU runs far ahead of the GPU
            y = mel(xs[i % N_INPUT_BUFFERS])
            marks[i + 1].record()
        stop.record()
        if rep == 0:
            with clocks:                     # sample clocks while the GPU works through the timed steps
                barrier()
        else:
            barrier()
        launches = mel.launch_count() - launches0
        per = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
        t_rep, _ = reduce_over_ranks(start.elapsed_time(stop) / 1e3, 0.0)      # max over ranks
        repeats.append({"ms_per_step": 1e3 * t_rep / args.steps, "median": sorted(per)[len(per) // 2],
                        "max": max(per), "argmax": per.index(max(per))})
    # the MEDIAN repeat is reported (best-of-N biases the figure upward); best and all repeats are side fields
    order = sorted(range(len(repeats)), key=lambda i: repeats[i]["ms_per_step"])
    best = order[len(order) // 2]
    fastest = order[0]
    t_max = repeats[best]["ms_per_step"] * args.steps / 1e3
    audio_local = args.steps * BATCH * WINDOW / SR
    _, audio_total = reduce_over_ranks(0.0, audio_local)
    value = audio_total / t_max
    clock_summary = clocks.summary()

    # ---- parity spot check (untimed, after the timed region): two windows against the CPU port
    #      on EVERY rank (each rank's own shard), all-reduced with MAX
    from oracle.torch_port import TorchPortMel
    port = TorchPortMel()
    port.fb.copy_(mel.transform.mel_scale.fb.cpu())
    port.window.copy_(mel.transform.spectrogram.window.cpu())
    yp = mel(xs[0][:2].contiguous())
    ref = port(xs[0][:2].cpu())
    parity = float((yp.cpu() - ref).abs().max())
    if world > 1:
        pt = torch.tensor([parity], dtype=torch.float64, device=dev)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        parity = float(pt.item())

    # ---- end to end: host buffers in, host buffers out -----------------------------------
    n_e2e = max(1, min(args.steps, 20))
    host_in = [torch.empty(BATCH, WINDOW, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    for hb, xb in zip(host_in, xs):
        hb.copy_(xb)
    host_out = torch.empty(BATCH, FRAMES, N_MELS, dtype=torch.float32, pin_memory=True)
    for i in range(min(args.warmup, 3)):
        mel.forward_host(host_in[i % 2], out=host_out)
    barrier()
    e2e_l0 = mel.launch_count()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        mel.forward_host(host_in[i % 2], out=host_out)
    torch.cuda.synchronize(dev)
    e2e_elapsed = time.perf_counter() - t0
    e2e_launches = mel.launch_count() - e2e_l0
    e2e_t, e2e_audio = reduce_over_ranks(e2e_elapsed, n_e2e * BATCH * WINDOW / SR)
    e2e_ok = bool(torch.isfinite(host_out[-1, -1]).all())
    del host_in

    # ---- plain-copy ceiling of this host for the SAME bytes: every rank at once, the host entry's own pattern
    #      (bhmel_host_chunk_plan's row chunks round-robin over 3 streams, one cudaMemcpyAsync H2D and one D2H
    #      per chunk, no kernel in between), pinned buffers.  e2e can at best equal this figure.
    def copy_ceiling(h_in, h_out, d_in, d_out, n_steps, pcm16=False):
        streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        plan = MelSpectrogram.host_chunk_plan(BATCH, WINDOW, pcm16=pcm16)
        assert sum(plan) == BATCH
        def one_step():
            b0 = 0
            for k, rows in enumerate(plan):
                with torch.cuda.stream(streams[k % 3]):
                    d_in[b0:b0 + rows].copy_(h_in[b0:b0 + rows], non_blocking=True)
                    h_out[b0:b0 + rows].copy_(d_out[b0:b0 + rows], non_blocking=True)
                b0 += rows
        one_step()
        best = None
        for _ in range(3):          # a ceiling: the best of three rounds (host-memory traffic of the box varies)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_steps):
                one_step()
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            t, audio = reduce_over_ranks(dt, n_steps * BATCH * WINDOW / SR)
            if best is None or audio / t > best[1] / best[0]:
                best = (t, audio)
        return best

    d_out32 = torch.empty(BATCH, FRAMES, N_MELS, dtype=torch.float32, device=dev)
    host_in1 = torch.empty(BATCH, WINDOW, dtype=torch.float32, pin_memory=True)
    cc_t, cc_audio = copy_ceiling(host_in1, host_out, xs[0], d_out32, max(3, n_e2e // 2))
    copy_ceiling_f32 = cc_audio / cc_t
    del host_in1, d_out32

    # ---- same end-to-end call with the next-row ingest/egress types: int16 PCM in (scaled on the
    #      device like load_audio_file does) and bfloat16 out (what the encoder consumes); every rank --
    pcm = (xs[0] * 32767.0).to(torch.int16).cpu().pin_memory()
    scales = torch.full((BATCH,), 1.0 / 32767.0)
    out16 = torch.empty(BATCH, FRAMES, N_MELS, dtype=torch.bfloat16, pin_memory=True)
    for _ in range(2):
        mel.forward_host(pcm, out=out16, scales=scales)
    barrier()
    t0 = time.perf_counter()
    n_typed = max(1, min(args.steps, 10))
    for _ in range(n_typed):
        mel.forward_host(pcm, out=out16, scales=scales)
    torch.cuda.synchronize(dev)
    typed_t, typed_audio = reduce_over_ranks(time.perf_counter() - t0, n_typed * BATCH * WINDOW / SR)
    d_pcm = torch.empty(BATCH, WINDOW, dtype=torch.int16, device=dev)
    d_out16 = torch.empty(BATCH, FRAMES, N_MELS, dtype=torch.bfloat16, device=dev)
    cc16_t, cc16_audio = copy_ceiling(pcm, out16, d_pcm, d_out16, max(3, n_typed // 2), pcm16=True)
    e2e_typed = {"value": typed_audio / typed_t, "unit": UNIT, "input": "int16 PCM + per-row scale",
                 "output": "bfloat16", "h2d_bytes_per_step": BATCH * WINDOW * 2,
                 "d2h_bytes_per_step": BATCH * FRAMES * N_MELS * 2, "n_gpus_measured": world,
                 "plain_copy_ceiling": cc16_audio / cc16_t,
                 "fraction_of_plain_copy_ceiling": (typed_audio / typed_t) / (cc16_audio / cc16_t)}
    del pcm, out16, d_pcm, d_out16

    # ---- smaller configs, for context (rank 0): C2 46-window song and the 10 s clip ----------
    extra = {}
    if rank == 0 and e2e_typed is not None:
        extra["e2e_pcm16_in_bf16_out"] = e2e_typed
    if rank == 0:
        def timed(fn, reps):
            fn(); torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps
        x46 = xs[0][:46]
        ms46 = timed(lambda: mel(x46), 20)
        extra["c2_song_46_windows_ms"] = ms46
        extra["c2_song_46_windows_audio_s_per_s"] = 46 * WINDOW / SR / (ms46 / 1e3)
        x6 = xs[1][:6]
        ms6 = timed(lambda: mel(x6), 20)
        extra["c2_song_6_windows_ms"] = ms6
        xc = xs[2][:1, :160000].contiguous()
        extra["c1_10s_clip_us"] = 1e3 * timed(lambda: mel(xc), 50)
        # the reference's other parameter sets, same 256-window batch: one row per set
        psets = {}
        for name, a in (("P128", ("torchaudio", True, SR, 1024, 128, 128, 20, 8000, "reflect")),
                        ("P1", ("torchaudio", False, SR, 1024, 388, 128, 0, 8000, "constant")),
                        ("T5", ("torchaudio", False, SR, 1024, 512, 128, 0, 8000, "constant"))):
            try:
                m2 = MelSpectrogram(*a).to(dev)
                y2 = torch.empty(BATCH, FRAMES, a[4], device=dev)
                ms = timed(lambda: m2.forward_into(xs[0], y2), 20)
                nbytes = BATCH * (4 * WINDOW + 4 * FRAMES * a[4])
                psets[name] = {"n_mels": a[4], "f_min": a[6], "pad_mode": a[8], "log_scale": a[1], "ms_per_step": ms,
                               "audio_s_per_s": BATCH * WINDOW / SR / (ms / 1e3),
                               "algorithmic_bytes_per_frame": 512 + 4 * a[4], "algorithmic_GBps": nbytes / ms / 1e6}
                del m2, y2
            except Exception as e:  # noqa: BLE001
                psets[name] = {"unavailable": f"{type(e).__name__}: {e}"}
        extra["psets"] = psets
        torch.cuda.empty_cache()
        # next row N3: frontend -> assembled channels-last encoder input -> tcgen05 conv stem
        # (libbhstem.so), 6 and 46 windows (C2), whisper-small dims (80 mel + 384 conditioning channels -> 768)
        try:
            from beatheritage_b200.conv_stem import ConvStem
            torch.manual_seed(0)
            stem = ConvStem(N_MELS + 384, 768).to(dev)
            try:
                bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
            except Exception:
                bf16_peak = 1590.0          # B200_PROFILING.md fallback
            stem_rows = {}
            for nw, xw in ((6, x6), (46, x46)):
                cond = torch.randn(nw, 384, device=dev)
                enc_in = mel.forward_encoder_input(xw, [cond], dtype=torch.bfloat16, channels_first=False)
                hid = torch.empty(nw, FRAMES, 768, dtype=torch.bfloat16, device=dev)
                yst = torch.empty(nw, FRAMES // 2, 768, dtype=torch.bfloat16, device=dev)
                ms_stem = timed(lambda: stem(enc_in, hidden=hid, out=yst), 20)
                ms_all = timed(lambda: stem(mel.forward_encoder_input(xw, [cond], dtype=torch.bfloat16,
                                                                      channels_first=False)), 20)
                flop = 2.0 * nw * FRAMES * 768 * 3 * (N_MELS + 384) + 2.0 * nw * (FRAMES // 2) * 768 * 3 * 768
                # split conv1 (bhstem_forward_split): the 384 time-constant channels folded into a per-window
                # bias, the [nw, 4096, 464] encoder input never built; same outputs within the bf16 tolerance
                frames16 = torch.empty(nw, FRAMES, N_MELS, dtype=torch.bfloat16, device=dev)
                mel.forward_into(xw, frames16)
                cond16 = cond.to(torch.bfloat16)
                ms_split = timed(lambda: stem.forward_split(frames16, cond16, hidden=hid, out=yst), 20)
                ms_split_all = timed(lambda: stem.forward_split(mel.forward_into(xw, frames16), cond16, hidden=hid,
                                                                out=yst), 20)
                flop_split = 2.0 * nw * FRAMES * 768 * 3 * N_MELS + 2.0 * nw * (FRAMES // 2) * 768 * 3 * 768
                stem_rows[f"{nw}_windows"] = {
                    "stem_ms": ms_stem, "stem_tflops": flop / ms_stem / 1e9,
                    "stem_frac_of_measured_bf16_peak": flop / ms_stem / 1e9 / bf16_peak,
                    "frontend_plus_assembly_plus_stem_ms": ms_all,
                    "split_stem_ms": ms_split, "split_stem_executed_tflops": flop_split / ms_split / 1e9,
                    "split_stem_speedup": ms_stem / ms_split,
                    "frontend_plus_split_stem_ms": ms_split_all}
            extra["conv_stem_n3"] = {"dims": "464 -> 768 channels, 4096 -> 2048 frames, bf16, fp32 accumulate",
                                     "kernel": "bhstem_conv_gelu_pair_kernel (tcgen05 cta_group::2 / TMEM / TMA)",
                                     "launches": stem.launch_count(), **stem_rows}
        except Exception as e:  # noqa: BLE001 -- the headline line must not depend on the next-row library
            extra["conv_stem_n3"] = {"unavailable": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if numa_bound:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        windows = 16
        mean_v, best_v, passes, threads, kind = cpu_port_throughput(args.cpu_budget, windows)
        what = ("the reference's own module (oracle/_ref/spectrogram.py)" if kind == "reference"
                else "torch CPU port of the reference path (oracle/torch_port.py)")
        cpu = {"value": mean_v, "best": best_v, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"{passes} passes over [{windows}, {WINDOW}] f32 (a 16-window slice of the 256-window step), "
                         f"{what}, ~{args.cpu_budget:.0f} s budget"}

    if rank == 0:
        ms_per_step = 1e3 * t_max / args.steps
        achieved = BATCH * ALGO_BYTES_PER_WINDOW / (t_max / args.steps) / 1e9
        peak, peak_src = 6538.0, "fallback"
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic, ncu_facts = None, {}
        try:
            ncu_facts = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = ncu_facts["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step,
            "ms_per_step_median": repeats[best]["median"], "ms_per_step_max": repeats[best]["max"],
            "timed_repeats": repeats, "reported_repeat": best, "reported_repeat_is": "median of the repeats",
            "ms_per_step_best_repeat": repeats[fastest]["ms_per_step"],
            "extra_untimed_warmup_steps": extra_warm, "untimed_runway_steps": runway, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_audio / e2e_t, "unit": UNIT, "h2d_bytes_per_step": BATCH * WINDOW * 4,
                    "d2h_bytes_per_step": BATCH * FRAMES * N_MELS * 4, "steps": n_e2e, "launches": e2e_launches,
                    "api": "MelSpectrogram.forward_host -> bhmel_forward_host (pinned host buffers)", "numa_bound": numa_bound,
                    "finite": e2e_ok,
                    "plain_copy_ceiling": copy_ceiling_f32,
                    "plain_copy_ceiling_is": f"the same {BATCH * WINDOW * 4 + BATCH * FRAMES * N_MELS * 4} bytes per step per rank as plain pinned "
                                             f"cudaMemcpyAsync H2D + D2H in the host entry's chunk pattern, no kernel, all {world} rank(s) at once, "
                                             "best of 3 rounds, expressed in audio-s/s",
                    "fraction_of_plain_copy_ceiling": (e2e_audio / e2e_t) / copy_ceiling_f32},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "bhmel_logmel_ws_kernel",
                         "algorithmic_bytes_per_launch": BATCH * ALGO_BYTES_PER_WINDOW,
                         "note": "fp32 CUDA-core FFT: the FP32 issue rate, not HBM, bounds this kernel (DESIGN.md)",
                         "static_from": "profiles/traffic.json (one ncu --set full capture of this kernel, not live): "
                                        "traffic, ncu_issue_slots_busy_pct, ncu_warp_instructions_per_launch",
                         "ncu_capture": (f"{ncu_facts.get('kernel')} -- profiles/{ncu_facts.get('source')}" if ncu_facts else None),
                         # the roof that actually binds: ~30.5 kflop of fp32 CUDA-core work per frame
                         # (real 1024-point FFT 25.6 k, window 1.0 k, power 1.5 k, banded mel 2.0 k, log1p 0.3 k)
                         # against the FMA peak of the SMs at the clock sampled during the timed region
                         "fp32": fp32_roof(BATCH * FRAMES / (t_max / args.steps), clock_summary),
                         "ncu_issue_slots_busy_pct": ncu_facts.get("issue_slots_busy_pct"),
                         "ncu_warp_instructions_per_launch": ncu_facts.get("warp_instructions_per_launch")},
            "cpu_baseline": cpu,
            "clocks": clock_summary,
            "parity_max_abs_err_vs_cpu_port": parity, "parity_ranks_checked": world,
            "extra": extra,
        }
        emit(line)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------
def run_config(args):
    """--config c1|c2|c4|c5: the other measurement configs of SURVEY.md 8(d) on ONE GPU, one JSON line in
    the same contract (the measurement code lives in tools/bench_configs.py).  `value` is window
    audio-seconds per second of the config's main row; `details` carries every row."""
    import torch
    rank, _, _ = dist_env()
    if rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    ctx = bc.Ctx()
    l0 = ctx.mel.launch_count()
    name = args.config
    if name == "c1":
        d = bc.run_c1(ctx)
        workload = "C1: one 10 s 16 kHz mono clip [1, 160000] f32 -> [1, 1251, 80], latency of one call"
        value, ms = 10.0 / (d["ours_us"] / 1e6), d["ours_us"] / 1e3
    elif name == "c2":
        d = bc.run_c2(ctx)
        workload = "C2: one 3-min song segmented per Preprocessor.segment: 46 overlapped windows [46, 524160] (and 6 non-overlapped)"
        row = d["sequential_46"]
        value, ms = row["window_audio_s_per_s_module"], row["module_ms"]
    elif name == "c4":
        d = bc.run_c4(ctx)
        workload = "C4: 1-hour continuous audio (57.6 M samples) streamed in overlapping windows, stride / batch sweeps; main row: stride 52415, fused gather"
        row = d["stride_sweep"][0]
        value, ms = row["gather_window_audio_s_per_s"], row["gather_ms"]
    else:
        d = bc.run_c5(ctx)
        workload = ("C5: inference slice, 3-min song -> segment -> H2D -> frontend -> encoder input -> random-init "
                    "WhisperEncoder; main row: fused encoder input, 6 windows per step")
        row = d.get("ours_fused_encoder_input_parallel_b6", {})
        ms = row.get("total_ms")
        value = (46 * WINDOW / SR) / (ms / 1e3) if ms else None
    emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": None, "warmup": None, "ms_per_step": ms,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": workload, "name": name}, "gpu_launches": ctx.mel.launch_count() - l0, "details": d})
    return 0


_JSON_OUT = None


def _claim_stdout() -> None:
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL prints its version
    banner there under torchrun) is sent to stderr; emit() writes the line to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--repeats", type=int, default=3, help="how many times the K timed steps are measured (the median repeat is reported)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline timing (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", choices=["c1", "c2", "c3", "c4", "c5"], default="c3",
                    help="c3 (default) is the headline dataset-preprocessing batch; the others are SURVEY.md 8(d)'s "
                         "single-GPU configs (one JSON line each)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _, _, world = dist_env()
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_config(args) if args.config != "c3" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
