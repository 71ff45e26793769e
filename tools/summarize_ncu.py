"""Turn an ncu report into the committed artefacts under profiles/:
  <tag>_raw.csv          the raw page (all metrics) of the profiled launch(es)
  <tag>_summary.md       headline metrics, stall breakdown, per-role sample shares
  traffic.json           DRAM bytes per launch for bench.py's roofline.traffic (if --traffic)
Usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_ws [--traffic]
"""
import csv
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
d = dict(zip(hdr, rows[2]))
u = dict(zip(hdr, units))


def f(k):
    try:
        return float(d[k])
    except Exception:
        return float("nan")


def to_bytes(k):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u.get(k, "byte"), 1)
    return f(k) * scale


keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "smsp__average_warp_latency_per_inst_issued.ratio"]
lines = [f"# ncu summary: {os.path.basename(rep)}", "",
         f"kernel: `{d.get('Kernel Name', '?')}`  (one launch of the bench workload: 256 windows x 524160 samples, P0)", "",
         "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in d:
        lines.append(f"| {k} | {d[k]} | {u.get(k, '')} |")
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
lines += ["", f"DRAM traffic per launch: {traffic / 1e6:.1f} MB (algorithmic 872.3 MB: 536.7 in + 335.5 out)", "",
          "## warp stall reasons (warps stalled per issued instruction)", "", "| reason | ratio |", "|---|---|"]
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and f(k) > 0.02:
        lines.append(f"| {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {f(k):.3f} |")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
if len(srows) > 2:
    sh = srows[1]
    ix = {h: i for i, h in enumerate(sh)}
    sass = [r for r in srows[2:] if len(r) > 5 and r[ix["Address"]].strip().startswith("0x")]
    stall = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in sass) or 1
    toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in sass) or 1
    marks = [n for n, r in enumerate(sass) if "USETMAXREG" in r[ix["Source"]] or "BAR.SYNC" in r[ix["Source"]]]
    bounds = [0] + marks + [len(sass)]
    lines += ["", "## sample / instruction shares between control markers (USETMAXREG / BAR.SYNC) in SASS order", "",
              "| SASS lines | contains | samples % | instructions % | top stall reasons |", "|---|---|---|---|---|"]
    for a, b in zip(bounds[:-1], bounds[1:]):
        seg = sass[a:b]
        if not seg:
            continue
        s = sum(int(r[ix["# Samples"]] or 0) for r in seg)
        i = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
        if s < 0.005 * tot:
            continue
        st = {}
        for r in seg:
            for c in stall:
                st[c] = st.get(c, 0) + int(r[ix[c]] or 0)
        top = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
        txt = " ".join(r[ix["Source"]] for r in seg)
        tags = [t for t in ("SHFL", "UBLKCP", "LDGSTS", "MUFU.LG2", "STG", "SYNCS") if t in txt]
        lines.append(f"| {a}-{b} | {' '.join(tags)} | {100 * s / tot:.1f} | {100 * i / toti:.1f} | {top} |")
open(out + "_summary.md", "w").write("\n".join(lines) + "\n")
if "--traffic" in sys.argv:
    json.dump({"dram_bytes_per_launch": traffic, "source": os.path.basename(out) + "_raw.csv",
               "kernel": d.get("Kernel Name", "?"), "duration_ms_under_ncu": f("gpu__time_duration.sum"),
               "issue_slots_busy_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
               "warp_instructions_per_launch": f("smsp__inst_executed.sum")},
              open(os.path.join(os.path.dirname(out), "traffic.json"), "w"), indent=1)
print("\n".join(lines[:40]))
