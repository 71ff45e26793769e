"""Summarise an ncu report: headline metrics + stall samples per kernel region (by SASS markers)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
d = dict(zip(hdr, rows[2]))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(d[k] or 0) > 0.02:
        print(f"  {k[34:-28]:30s} {float(d[k]):.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[2:] if len(r) > 5 and r[ix["Address"]].strip().startswith("0x")]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
# region boundaries: BAR.SYNC instructions and the SHFL-containing loop
marks = []
for n, r in enumerate(sass):
    txt = r[ix["Source"]]
    if "BAR.SYNC" in txt:
        marks.append((n, "BAR"))
tot = sum(int(r[ix["# Samples"]] or 0) for r in sass)
toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in sass)
print("total samples", tot, "instr", toti, "SASS lines", len(sass), "BAR at", [m[0] for m in marks])
bounds = [0] + [m[0] + 1 for m in marks] + [len(sass)]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = sass[a:b]
    s = sum(int(r[ix["# Samples"]] or 0) for r in seg)
    i = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    st = {}
    for r in seg:
        for c in stall_cols:
            st[c] = st.get(c, 0) + int(r[ix[c]] or 0)
    top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
    has_shfl = any("SHFL" in r[ix["Source"]] for r in seg)
    print(f"lines {a:5d}-{b:5d} {'FFT' if has_shfl else '   '} samples {100 * s / max(tot, 1):5.1f}% instr {100 * i / max(toti, 1):5.1f}%  "
          + ", ".join(f"{k[6:]}={100 * v / max(s, 1):.0f}%" for k, v in top))
if len(sys.argv) > 2:   # dump hottest instructions
    hot = sorted(sass, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2])]
    for r in hot:
        print(r[ix["# Samples"]], r[ix["Source"]].strip()[:90])
