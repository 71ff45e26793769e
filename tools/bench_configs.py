#!/usr/bin/env python3
"""Measurement configs C1..C5 of SURVEY.md section 8(d) on one B200 (bench.py covers C3, the headline).

    python tools/bench_configs.py [--out profiles/r1_configs.json] [--skip-c5]

C1  10 s clip: latency (us) of one call, ours vs torchaudio on the same GPU (the CPU figure is
    bench.py's cpu_baseline; this script never touches oracle/).
C2  3-min song: 6 (parallel) and 46 (sequential, stride 52 415) windows; module mode and fused gather.
C4  1-hour audio streamed in overlapping windows: stride sweep x {materialised batch, fused gather},
    and a batch-size sweep of the module interface (achieved algorithmic GB/s vs batch).
C5  inference slice: song -> segment -> H2D -> frontend -> bf16 -> cat 384 conditioning channels ->
    swapaxes -> random-init HF WhisperEncoder (whisper-small dims, 464 input channels); frontend
    share of wall time for the reference's GPU frontend (torchaudio: cuFFT + cuBLAS) and ours.
The "reference GPU path" is torchaudio.transforms.MelSpectrogram + log1p + permute on CUDA, i.e.
what the reference module runs when the model sits on a GPU; informative, not the published baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from beatheritage_b200 import MelSpectrogram  # noqa: E402
from beatheritage_b200.segment import segment_plan  # noqa: E402

WINDOW, SR, HOP, M = 524160, 16000, 128, 80
P0 = ("torchaudio", True, SR, 1024, M, HOP, 20, 8000, "reflect")
dev = torch.device("cuda", 0)


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


class TorchaudioGpu(torch.nn.Module):
    """The reference module's own arithmetic on CUDA (spectrogram.py:38-49, 79-82)."""

    def __init__(self):
        super().__init__()
        import torchaudio
        self.t = torchaudio.transforms.MelSpectrogram(sample_rate=SR, n_fft=1024, n_mels=M, hop_length=HOP,
                                                      center=True, f_min=20, f_max=8000, pad_mode="reflect")

    def forward(self, x):
        return torch.log1p(self.t(x)).permute(0, 2, 1)


def algo_bytes(n_in_samples, n_windows):
    return 4 * n_in_samples + 4 * n_windows * (WINDOW // HOP + 1) * M


class Ctx:
    """Modules and inputs shared by the per-config functions (built lazily)."""

    def __init__(self):
        self.mel = MelSpectrogram(*P0).to(dev)
        try:
            self.ta = TorchaudioGpu().to(dev)
            self.ta_error = None
        except Exception as e:   # torchaudio missing
            self.ta, self.ta_error = None, str(e)
        self._song = None

    @property
    def song(self):
        if self._song is None:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from tests.golden import signals
            self._song = signals.music(2_880_000, seed=1)
        return self._song


def materialise(song_t, plan):
    """What Preprocessor.segment builds on the host: zero-pad to plan.padded_len, strided windows."""
    sp = torch.nn.functional.pad(song_t, (0, plan.padded_len - song_t.numel()))
    return sp.as_strided((plan.n_windows, plan.window_len), (plan.stride, 1)).contiguous()


def run_c1(ctx):
    mel, ta = ctx.mel, ctx.ta
    g = torch.Generator().manual_seed(0)
    x1 = torch.rand(1, 160000, generator=g) * 2 - 1
    x1d = x1.to(dev)
    c1 = {"ours_us": 1e3 * timed(lambda: mel(x1d), 200)}
    if ta is not None:
        c1["torchaudio_gpu_us"] = 1e3 * timed(lambda: ta(x1d), 200)
        c1["max_abs_diff_vs_torchaudio_gpu"] = float((mel(x1d) - ta(x1d)).abs().max())


    return c1


def run_c2(ctx):
    mel, ta = ctx.mel, ctx.ta
    song = ctx.song
    song_d = torch.from_numpy(song).to(dev)
    c2 = {}
    for name, parallel in (("sequential_46", False), ("parallel_6", True)):
        plan = segment_plan(len(song), parallel=parallel)
        seq = materialise(song_d, plan)
        ms_mod = timed(lambda: mel(seq), 50)
        ms_gat = timed(lambda: mel.forward_gather(song_d, 0, plan.stride, plan.n_windows, plan.window_len), 50)
        same = bool(torch.equal(mel(seq), mel.forward_gather(song_d, 0, plan.stride, plan.n_windows, plan.window_len)))
        entry = {"windows": plan.n_windows, "module_ms": ms_mod, "gather_ms": ms_gat, "gather_equals_module": same,
                 "window_audio_s_per_s_module": plan.n_windows * WINDOW / SR / (ms_mod / 1e3),
                 "song_s_per_s_gather": len(song) / SR / (ms_gat / 1e3)}
        if ta is not None:
            entry["torchaudio_gpu_ms"] = timed(lambda: ta(seq), 20)
        c2[name] = entry


    return c2


def run_c4(ctx):
    mel, ta = ctx.mel, ctx.ta
    n_hour = 57_600_000
    gd = torch.Generator(device=dev).manual_seed(2)
    hour = torch.rand(n_hour, device=dev, generator=gd).mul_(2).sub_(1)
    c4 = {"stride_sweep": [], "batch_sweep": []}
    for stride in (52415, 131040, 262080, 524160):
        rem = (n_hour - WINDOW) % stride
        padded = n_hour + (0 if rem == 0 else stride - rem)
        W = (padded - WINDOW) // stride + 1
        ms_g = timed(lambda: mel.forward_gather(hour, 0, stride, W, WINDOW), 5, warm=2)
        entry = {"stride": stride, "windows": W, "gather_ms": ms_g,
                 "gather_window_audio_s_per_s": W * WINDOW / SR / (ms_g / 1e3),
                 "gather_song_s_per_s": n_hour / SR / (ms_g / 1e3),
                 "gather_algorithmic_GBps": algo_bytes(n_hour, W) / (ms_g / 1e3) / 1e9}
        # materialised batch (what Preprocessor.segment builds): [W, 524160] f32
        hp = torch.nn.functional.pad(hour, (0, padded - n_hour))
        batch = hp.as_strided((W, WINDOW), (stride, 1)).contiguous()
        ms_m = timed(lambda: mel(batch), 5, warm=2)
        entry.update({"module_ms": ms_m, "module_window_audio_s_per_s": W * WINDOW / SR / (ms_m / 1e3),
                      "module_algorithmic_GBps": algo_bytes(W * WINDOW, W) / (ms_m / 1e3) / 1e9,
                      "materialised_input_MB": W * WINDOW * 4 / 1e6})
        if stride == 52415:
            entry["gather_equals_module"] = bool(torch.equal(mel(batch[:64]), mel.forward_gather(hour, 0, stride, 64, WINDOW)))
            for b in (1, 4, 16, 64, 256, W):
                xb = batch[:b]
                ms = timed(lambda: mel(xb), 20 if b <= 64 else 5, warm=2)
                c4["batch_sweep"].append({"batch": b, "ms": ms, "window_audio_s_per_s": b * WINDOW / SR / (ms / 1e3),
                                          "algorithmic_GBps": algo_bytes(b * WINDOW, b) / (ms / 1e3) / 1e9})
        del batch, hp
        c4["stride_sweep"].append(entry)
    del hour
    torch.cuda.empty_cache()


    return c4


def run_c5(ctx):
    mel, ta = ctx.mel, ctx.ta
    song = ctx.song
    song_d = torch.from_numpy(song).to(dev)
    try:
        from transformers import WhisperConfig
        from transformers.models.whisper.modeling_whisper import WhisperEncoder
        torch.manual_seed(0)
        cfg = WhisperConfig(d_model=768, encoder_layers=12, encoder_attention_heads=12, encoder_ffn_dim=3072,
                            num_mel_bins=M + 384, max_source_positions=2048)
        enc = WhisperEncoder(cfg).to(dev).to(torch.bfloat16).eval()
        plan = segment_plan(len(song))
        seq_host = materialise(song_d, plan).cpu().pin_memory()
        cond = torch.randn(1, 1, 384, device=dev, dtype=torch.bfloat16)

        def run(frontend, batch, fused=False):
            """fused=True: mel + conditioning channels + channels-first layout in one library call
            (forward_encoder_input); the 'frontend' time then includes the assembly."""
            t_front = t_total = 0.0
            for i in range(0, seq_host.shape[0], batch):
                xb = seq_host[i:i + batch]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                xd = xb.to(dev, non_blocking=True)                    # server.py:42
                if fused:
                    fr = frontend.forward_encoder_input(xd, [cond[:, 0].expand(xd.shape[0], -1)], channels_first=True)
                else:
                    fr = frontend(xd)                                 # modeling_mapperatorinator.py:351
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                if not fused:
                    fr = fr.to(torch.bfloat16)                        # :352
                    fr = torch.cat([fr, cond.expand(fr.shape[0], fr.shape[1], -1)], dim=-1)   # :369-370
                    fr = fr.swapaxes(1, 2)                            # :375-376
                with torch.no_grad():
                    enc(fr)                                           # encoder
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                t_front += t1 - t0
                t_total += t2 - t0
            return t_front, t_total

        c5 = {"encoder": "HF WhisperEncoder random init, d_model 768, 12 layers, 464 input channels, bf16",
              "windows": int(seq_host.shape[0])}
        fronts = {"ours": mel}
        if ta is not None:
            fronts["torchaudio_gpu"] = ta
        with torch.no_grad():
            for fname, f in fronts.items():
                for mode, batch in (("sequential_b1", 1), ("parallel_b6", 6)):
                    run(f, batch)
                    tf, tt = run(f, batch)
                    c5[f"{fname}_{mode}"] = {"frontend_ms_incl_h2d": 1e3 * tf, "total_ms": 1e3 * tt,
                                             "frontend_share": tf / tt}
            for mode, batch in (("sequential_b1", 1), ("parallel_b6", 6)):
                run(mel, batch, fused=True)
                tf, tt = run(mel, batch, fused=True)
                c5[f"ours_fused_encoder_input_{mode}"] = {"frontend_plus_assembly_ms_incl_h2d": 1e3 * tf,
                                                          "total_ms": 1e3 * tt, "frontend_share": tf / tt}
        # N3: the encoder's conv stem (modeling_ropewhisper.py:1206-1209 / the stock encoder's first lines)
        # behind the frontend: reference ops end to end vs forward_encoder_input (channels last) + ConvStem
        from beatheritage_b200.conv_stem import ConvStem
        stem = ConvStem.from_encoder(enc)
        gelu = torch.nn.functional.gelu
        stem_res = {}
        with torch.no_grad():
            for mode, batch in (("sequential_b1", 1), ("parallel_b6", 6)):
                xd = seq_host[:batch].to(dev)
                cvec = cond[:, 0].expand(batch, -1)

                def ref_ops():
                    fr = mel(xd).to(torch.bfloat16)
                    fr = torch.cat([fr, cond.expand(batch, fr.shape[1], -1)], dim=-1).swapaxes(1, 2)
                    return gelu(enc.conv2(gelu(enc.conv1(fr)))).permute(0, 2, 1)

                def ref_stem_only(fr_bct):
                    return gelu(enc.conv2(gelu(enc.conv1(fr_bct)))).permute(0, 2, 1)

                def ours():
                    return stem(mel.forward_encoder_input(xd, [cvec], channels_first=False))

                # split form: bf16 frames straight from the frontend, conditioning channels folded into a bias
                frames16 = torch.empty(batch, 4096, 80, dtype=torch.bfloat16, device=dev)
                hid16 = torch.empty(batch, 4096, enc.conv1.out_channels, dtype=torch.bfloat16, device=dev)
                out16 = torch.empty(batch, 2048, enc.conv1.out_channels, dtype=torch.bfloat16, device=dev)
                cvec16 = cvec.to(torch.bfloat16).contiguous()

                def ours_split():
                    return stem.forward_split(mel.forward_into(xd, frames16), cvec16, hidden=hid16, out=out16)

                fr_btc = mel.forward_encoder_input(xd, [cvec], channels_first=False)
                fr_bct = fr_btc.swapaxes(1, 2).contiguous()
                a, b = ref_ops().float(), ours().float()
                b_split = ours_split().float().clone()
                split_graph_ms = None
                try:
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            ours_split()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    torch.cuda.synchronize()
                    sgraph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(sgraph):
                        sg_out = ours_split()
                    sgraph.replay()
                    torch.cuda.synchronize()
                    if torch.equal(sg_out.float(), b_split):
                        split_graph_ms = timed(sgraph.replay, 50)
                except Exception as e:  # noqa: BLE001
                    split_graph_ms = f"unavailable: {type(e).__name__}: {e}"
                # the same chain captured once in a CUDA graph (the serving loop: no host work per call)
                graph_ms = None
                try:
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            ours()
                    torch.cuda.current_stream(dev).wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        g_out = ours()
                    graph.replay()
                    torch.cuda.synchronize()
                    if torch.equal(g_out.float(), b):
                        graph_ms = timed(graph.replay, 50)
                except Exception as e:  # noqa: BLE001
                    graph_ms = f"unavailable: {type(e).__name__}: {e}"
                stem_res[mode] = {
                    "ours_frontend_to_stem_cuda_graph_ms": graph_ms,
                    "reference_ops_frontend_to_stem_ms": timed(ref_ops, 20),
                    "ours_frontend_to_stem_ms": timed(ours, 20),
                    "stem_only_torch_cudnn_ms": timed(lambda: ref_stem_only(fr_bct), 20),
                    "stem_only_ours_ms": timed(lambda: stem(fr_btc), 20),
                    "max_abs_diff": float((a - b).abs().max()),
                    "max_abs_value": float(a.abs().max()),
                    "ours_split_frontend_to_stem_ms": timed(ours_split, 20),
                    "ours_split_frontend_to_stem_cuda_graph_ms": split_graph_ms,
                    "split_max_abs_diff_vs_reference_ops": float((a - b_split).abs().max()),
                }
        c5["conv_stem_N3"] = stem_res
        return c5
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_configs.json"))
    ap.add_argument("--skip-c5", action="store_true")
    args = ap.parse_args()
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    ctx = Ctx()
    if ctx.ta is None:
        res["torchaudio_gpu"] = f"unavailable: {ctx.ta_error}"
    res["C1_clip_10s"] = run_c1(ctx)
    res["C2_song_3min"] = run_c2(ctx)
    res["C4_one_hour_stream"] = run_c4(ctx)
    if not args.skip_c5:
        res["C5_inference_slice"] = run_c5(ctx)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
