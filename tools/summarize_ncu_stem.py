"""ncu report of the conv stem (two launches: conv1, conv2) -> profiles/<tag>_raw.csv + <tag>_summary.md.
Usage: python tools/summarize_ncu_stem.py gpurun_out/prof_stem.ncu-rep profiles/r1_stem [B] [split]
With `split` the report holds the three launches of bhstem_forward_split (folded bias, split conv1, conv2) taken with
`python tools/run_stem_split_once.py B`."""
import csv
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
SPLIT = len(sys.argv) > 4 and sys.argv[4] == "split"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
T, C, D = 4096, 464, 768
flops = [2.0 * B * T * D * 3 * C, 2.0 * B * (T // 2) * D * 3 * D]
algo = [2.0 * (B * T * C + 3 * D * C + B * T * D), 2.0 * (B * T * D + 3 * D * D + B * (T // 2) * D)]
lines = [f"# ncu summary: {os.path.basename(rep)}", "",
         f"`python tools/run_stem_once.py {B}`: {B} windows x {T} frames x {C} channels -> {D}; launch 1 = conv1 + GELU, "
         "launch 2 = conv2 (stride 2) + GELU.  `--set full --clock-control none`; durations under ncu are cold-cache and "
         "serialised -- the bench numbers are CUDA-event times outside the profiler.", ""]
NV = 80
if SPLIT:
    flops = [2.0 * B * D * 3 * (C - NV), 2.0 * B * T * D * 3 * NV, flops[1]]
    algo = [2.0 * (3 * D * (C - NV) + B * (C - NV)) + 4.0 * B * 3 * D,
            2.0 * (B * T * NV + 3 * D * NV + B * T * D) + 4.0 * B * 3 * D, algo[1]]
    lines[2] = (f"`python tools/run_stem_split_once.py {B}`: {B} windows x {T} frames x ({NV} time-varying + {C - NV} folded) "
                f"channels -> {D}; launch 1 = folded bias (the time-constant channels' three per-tap sums per window and "
                "output channel), launch 2 = split conv1 + GELU (16 epilogue warps), launch 3 = conv2 (stride 2) + GELU.  "
                "`--set full --clock-control none`; durations under ncu are cold-cache and serialised -- the bench numbers "
                "are CUDA-event times outside the profiler.")
for n, r in enumerate(rows[2:2 + len(flops)]):
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    dur_us = float(d["gpu__time_duration.sum"]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(u["gpu__time_duration.sum"], 1)
    lines += [f"## launch {n + 1}: `{d.get('Kernel Name', '?')[:80]}`", "",
              f"algorithmic work {flops[n] / 1e9:.1f} GFLOP, algorithmic bytes {algo[n] / 1e6:.1f} MB "
              f"-> {flops[n] / dur_us / 1e6:.0f} TFLOP/s under ncu", "", "| metric | value | unit |", "|---|---|---|"]
    for k in keys:
        if k in d:
            lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
    lines += ["", "warp stall reasons (warps stalled per issued instruction, > 0.05):", ""]
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.05:
                lines.append(f"* {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.2f}")
    lines.append("")
open(out + "_summary.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
