#!/usr/bin/env python3
"""Generates mel_static_gen.h: the mel projection of the baked filterbanks (bhmel_fb_baked.h) as
straight-line code -- weights are FFMA immediates, every power-spectrum block is loaded once per
warp, zero weights cost nothing, no descriptors.

Bit-identical to the generic band_dot2 path (bhmel_kernel.cuh) by construction: per filter, chain c
accumulates bins k = c (mod 4) in ascending order with fmaf, and the chains are combined as
(a0 + a1) + (a2 + a3); chains that only ever see zero weights are exactly +0 in the generic path
and are dropped here.

The 8 warps of the mel role take contiguous runs of filters (balanced by a small DP on an
instruction-count model); a warp reads the union of its filters' bin blocks once.

    python gen_mel_static.py [-o mel_static_gen.h]
"""
import argparse
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
N_WARPS = 8         # warps of the mel role
STATIC_WARPS = 4    # of which this many run generated code; the rest run the generic stage on the
                    # remaining (highest, widest) filters.  All-static (8) is 28 KB of code and falls
                    # out of the instruction cache next to the FFT role.


def read_baked(path):
    txt = open(path).read()
    tables = {}
    for m in re.finditer(r"static const unsigned kBaked(\w+)\[\]\[3\] = \{(.*?)\};", txt, re.S):
        name = m.group(1)
        n_mels = int(re.search(rf"kBaked{name}Mels = (\d+)", txt).group(1))
        ent = [(int(a), int(b), int(c, 16)) for a, b, c in re.findall(r"\{(\d+), (\d+), 0x([0-9a-f]+)u\}", m.group(2))]
        tables[name] = (n_mels, ent)
    return tables


def run_cost(filters, lo, hi):
    """Instruction-count model of one warp's run of filters [lo, hi)."""
    blocks = set()
    nnz = 0
    for f in range(lo, hi):
        for k in filters[f]:
            blocks.add(k // 4)
        nnz += len(filters[f])
    return nnz + 8 * (hi - lo) + 3 * len(blocks)   # a block load also costs 4 shared-memory wavefronts


def partition(filters, n_parts):
    n = len(filters)
    INF = 1 << 60
    best = [[INF] * (n + 1) for _ in range(n_parts + 1)]
    cut = [[0] * (n + 1) for _ in range(n_parts + 1)]
    best[0][0] = 0
    for p in range(1, n_parts + 1):
        for j in range(p, n + 1):
            for i in range(p - 1, j):
                c = max(best[p - 1][i], run_cost(filters, i, j))
                if c < best[p][j]:
                    best[p][j], cut[p][j] = c, i
    bounds = [n]
    for p in range(n_parts, 0, -1):
        bounds.append(cut[p][bounds[-1]])
    return bounds[::-1]


def emit_table(name, n_mels, ent):
    filters = [dict() for _ in range(n_mels)]   # filter -> {bin: bits}
    for k, m, bits in ent:
        filters[m][k] = bits
    bounds = partition(filters, N_WARPS)[:STATIC_WARPS + 1]
    L = []
    L.append(f"// {name}: {n_mels} filters, {len(ent)} non-zero weights; static warp runs {bounds}; filters")
    L.append(f"// {bounds[-1]}..{n_mels - 1} stay on the generic stage (warps {STATIC_WARPS}..{N_WARPS - 1})")
    L.append(f"constexpr int kStatic{name}Warps = {STATIC_WARPS};")
    L.append(f"constexpr int kStatic{name}Filters = {bounds[-1]};")
    L.append("template <bool kLog>")
    L.append(f"__device__ __forceinline__ void mel_static_{name}(const float4* __restrict__ prow, float* __restrict__ orow, int mw) {{")
    L.append("  switch (mw) {")
    for w in range(STATIC_WARPS):
        lo, hi = bounds[w], bounds[w + 1]
        L.append(f"    case {w}: {{   // filters {lo}..{hi - 1}")
        L.append("      float4 x;")
        blocks = sorted({k // 4 for f in range(lo, hi) for k in filters[f]})
        live = {}   # (filter, chain) -> assigned
        last_block = {f: max(filters[f]) // 4 for f in range(lo, hi) if filters[f]}
        for f in range(lo, hi):
            if not filters[f]:   # all-zero filter: the generic path stores log1p(0) = 0
                L.append(f"      orow[{f}] = 0.f;")
        for g in blocks:
            L.append(f"      x = prow[{g}];")
            for f in range(lo, hi):
                for c in range(4):
                    k = 4 * g + c
                    if k in filters[f]:
                        wlit = f"__uint_as_float(0x{filters[f][k]:08x}u)"
                        comp = "xyzw"[c]
                        if (f, c) in live:
                            L.append(f"      a{f}_{c} = fmaf(x.{comp}, {wlit}, a{f}_{c});")
                        else:
                            L.append(f"      float a{f}_{c} = __fmul_rn(x.{comp}, {wlit});")   # never contracted into a later add
                            live[(f, c)] = True
                if last_block.get(f) == g:
                    lo_pair = [f"a{f}_{c}" for c in (0, 1) if (f, c) in live]
                    hi_pair = [f"a{f}_{c}" for c in (2, 3) if (f, c) in live]
                    parts = []
                    for pr in (lo_pair, hi_pair):
                        if len(pr) == 2:
                            parts.append(f"({pr[0]} + {pr[1]})")
                        elif len(pr) == 1:
                            parts.append(pr[0])
                    expr = " + ".join(parts)
                    L.append(f"      {{ float v = {expr}; if constexpr (kLog) v = fast_log1p(v); orow[{f}] = v; }}")
        L.append("    } break;")
    L.append("    default: break;")
    L.append("  }")
    L.append("}")
    L.append("")
    return L


def main():
    global N_WARPS, STATIC_WARPS
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--out", default=os.path.join(HERE, "mel_static_gen.h"))
    ap.add_argument("--warps", type=int, default=N_WARPS, help="warps of the mel role")
    ap.add_argument("--static-warps", type=int, default=STATIC_WARPS)
    args = ap.parse_args()
    N_WARPS, STATIC_WARPS = args.warps, args.static_warps
    tables = read_baked(os.path.join(HERE, "bhmel_fb_baked.h"))
    L = ["// mel_static_gen.h -- generated by gen_mel_static.py; do not edit.", "#pragma once", "", "namespace bhmel {", ""]
    for name, (n_mels, ent) in tables.items():
        L += emit_table(name, n_mels, ent)
    L.append("}  // namespace bhmel")
    open(args.out, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    main()
