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
M_CHUNK = 96        # filters per output chunk (kMelChunk of bhmel_tables.h)
STAGE_COLS = 32     # direct form: filters per private staging block (kStageCols of bhmel_kernel_ws.cuh)
CHAIN1_MAX = 6      # direct form: bands up to this many bins accumulate in one chain, longer ones in two
PACKED = True       # direct form: a filter's fully covered 4-bin block runs as LDCU.128 (four weights from constant
                    # memory into uniform registers) + two FFMA2 (packed fp32 pairs: even bin -> chain 0, odd bin ->
                    # chain 1) instead of four immediate FFMAs: 48 instead of 64 bytes of code, same arithmetic
STATIC_WARPS = 4    # P0: of which this many run generated code; the rest run the generic stage on the
                    # remaining (highest, widest) filters.  All-static (8) is 28 KB of code and falls
                    # out of the instruction cache next to the FFT role.


def read_baked(path):
    """name -> (id, n_mels, [(bin, mel, bits), ...]) from the compact tables of bhmel_fb_baked.h."""
    txt = open(path).read()

    def arr(kind, name):
        body = re.search(rf"kBaked{name}{kind}\[\] = \{{(.*?)\}};", txt, re.S).group(1)
        return [int(v.rstrip("u"), 0) for v in body.replace("\n", " ").split(",") if v.strip()]

    tables = {}
    order = re.findall(r'\{"(\w+)", (\d+), kBaked', txt)
    for name, ident in order:
        n_mels = int(re.search(rf"kBaked{name}Mels = (\d+)", txt).group(1))
        start, count, bits = arr("Start", name), arr("Count", name), arr("Bits", name)
        assert len(start) == len(count) == n_mels and len(bits) == sum(count)
        ent, pos = [], 0
        for m in range(n_mels):
            for j in range(count[m]):
                ent.append((start[m] + j, m, bits[pos]))
                pos += 1
        tables[name] = (int(ident), n_mels, ent)
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


def emit_run(L, filters, lo, hi, col0, ind):
    """Straight-line code for filters [lo, hi): every 4-bin block their bands touch is loaded once;
    results go to orow[filter - col0]."""
    L.append(f"{ind}float4 x;")
    blocks = sorted({k // 4 for f in range(lo, hi) for k in filters[f]})
    live = {}   # (filter, chain) -> assigned
    last_block = {f: max(filters[f]) // 4 for f in range(lo, hi) if filters[f]}
    for f in range(lo, hi):
        if not filters[f]:   # all-zero filter: the generic path stores log1p(0) = 0
            L.append(f"{ind}orow[{f - col0}] = 0.f;")
    for g in blocks:
        L.append(f"{ind}x = prow[{g}];")
        for f in range(lo, hi):
            for c in range(4):
                k = 4 * g + c
                if k in filters[f]:
                    wlit = f"__uint_as_float(0x{filters[f][k]:08x}u)"
                    comp = "xyzw"[c]
                    if (f, c) in live:
                        L.append(f"{ind}a{f}_{c} = fmaf(x.{comp}, {wlit}, a{f}_{c});")
                    else:
                        L.append(f"{ind}float a{f}_{c} = __fmul_rn(x.{comp}, {wlit});")   # never contracted into a later add
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
                L.append(f"{ind}{{ float v = {expr}; if constexpr (kLog) v = fast_log1p(v); orow[{f - col0}] = v; }}")


def emit_table(name, n_mels, ent):
    """Hybrid form (P0): STATIC_WARPS of the 8 mel warps run generated code for the lowest filters, the
    others the generic stage on the pair tables of the rest (instruction-cache sweet spot)."""
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
        emit_run(L, filters, lo, hi, 0, "      ")
        L.append("    } break;")
    L.append("    default: break;")
    L.append("  }")
    L.append("}")
    L.append("")
    return L


def partition4(filters, n_parts):
    """partition() with every cut at a multiple of 4 filters (the direct form stores aligned groups of 4)."""
    groups = [None] * (len(filters) // 4)
    n = len(groups)
    INF = 1 << 60
    best = [[INF] * (n + 1) for _ in range(n_parts + 1)]
    cut = [[0] * (n + 1) for _ in range(n_parts + 1)]
    best[0][0] = 0
    cost = {}
    for p in range(1, n_parts + 1):
        for j in range(p, n + 1):
            for i in range(p - 1, j):
                if (i, j) not in cost:
                    cost[(i, j)] = run_cost(filters, 4 * i, 4 * j)
                c = max(best[p - 1][i], cost[(i, j)])
                if c < best[p][j]:
                    best[p][j], cut[p][j] = c, i
    bounds = [n]
    for p in range(n_parts, 0, -1):
        bounds.append(cut[p][bounds[-1]])
    return [4 * b for b in bounds[::-1]]


def emit_table_direct(name, n_mels, ent):
    """Direct form: all mel warps run generated code, a contiguous 4-aligned run of filters per warp, cut
    into PARTS of at most STAGE_COLS filters.  A filter accumulates in ONE chain (bands of at most
    CHAIN1_MAX bins) or two (even / odd bins), in ascending bin order, so there are almost no adds left:
    the code of a part is its non-zeros as FMAs + one load per power block + one 128-bit store per four
    finished filters into the warp's PRIVATE staging block [32 frames][STAGE_COLS].  The kernel loops
    over the parts and writes each staged block out with one shared, compact loop (mel_flush: the
    log1p epilogue and the output type conversion happen THERE, once, not per filter in the generated
    code; row segments of up to 128 contiguous bytes) -- no chunk loop over the whole role, no cross-warp
    barriers.  kRun<name>[warp * parts + part] = {first filter, filters} of every part.  (Not
    bit-identical to the generic stage's four-chain order; tests hold it to the oracle and to the
    generic stage within a few ulp.)"""
    assert n_mels % 4 == 0
    filters = [dict() for _ in range(n_mels)]
    for k, m, bits in ent:
        filters[m][k] = bits
    bounds = partition4(filters, N_WARPS)
    n_parts = max((bounds[w + 1] - bounds[w] + STAGE_COLS - 1) // STAGE_COLS for w in range(N_WARPS))
    runs = []
    wtab = []           # packed form: four weight bit patterns per fully covered (filter, block)
    L = [f"// {name}: {n_mels} filters, {len(ent)} non-zero weights; direct form, warp runs {bounds}, {n_parts} part(s)",
         f"constexpr int kParts{name} = {n_parts};"]
    body = [f"__device__ __forceinline__ void mel_direct_{name}(const float4* __restrict__ prow, float* __restrict__ srow, int mw, int part) {{",
            f"  switch (mw * {n_parts} + part) {{"]
    for w in range(N_WARPS):
        for part in range(n_parts):
            lo = bounds[w] + part * STAGE_COLS
            hi = min(lo + STAGE_COLS, bounds[w + 1])
            if lo >= hi:
                runs.append((0, 0))
                continue
            runs.append((lo, hi - lo))
            ind = "      "
            body.append(f"    case {w * n_parts + part}: {{   // filters {lo}..{hi - 1}")
            body.append(f"{ind}float4 x;")
            blocks = sorted({k // 4 for f in range(lo, hi) for k in filters[f]})
            last_block = {f: max(filters[f]) // 4 for f in range(lo, hi) if filters[f]}
            nch = {f: (1 if len(filters[f]) <= CHAIN1_MAX else 2) for f in range(lo, hi)}
            live, done = set(), set()
            state = {"next": lo}

            def drain():   # stage finished groups of four in ascending order
                while state["next"] < hi and all(f in done for f in range(state["next"], state["next"] + 4)):
                    m0 = state["next"]
                    body.append(f"{ind}mel_stage4(srow, {m0 - lo}, v{m0}, v{m0 + 1}, v{m0 + 2}, v{m0 + 3});")
                    state["next"] = m0 + 4

            for f in range(lo, hi):
                if not filters[f]:   # all-zero filter: log1p(0) = 0
                    body.append(f"{ind}const float v{f} = 0.f;")
                    done.add(f)
            drain()
            for g in blocks:
                body.append(f"{ind}x = prow[{g}];")
                for f in range(lo, hi):
                    if PACKED and nch[f] == 2 and all(4 * g + c in filters[f] for c in range(4)):
                        idx = len(wtab) // 4
                        wtab += [filters[f][4 * g + c] for c in range(4)]
                        if (f, 0) not in live and (f, 1) not in live:
                            body.append(f"{ind}float a{f}_0, a{f}_1; mel_mul2x2(a{f}_0, a{f}_1, x, kWpk{name}[{idx}]);")
                        else:
                            for ch in (0, 1):
                                if (f, ch) not in live:
                                    body.append(f"{ind}float a{f}_{ch} = 0.f;")
                            body.append(f"{ind}mel_fma2x2(a{f}_0, a{f}_1, x, kWpk{name}[{idx}]);")
                        live.add((f, 0))
                        live.add((f, 1))
                        if last_block.get(f) == g:
                            body.append(f"{ind}const float v{f} = a{f}_0 + a{f}_1;")
                            done.add(f)
                        continue
                    for c in range(4):
                        k = 4 * g + c
                        if k in filters[f]:
                            ch = (k & 1) if nch[f] == 2 else 0
                            wlit = f"__uint_as_float(0x{filters[f][k]:08x}u)"
                            comp = "xyzw"[c]
                            if (f, ch) in live:
                                body.append(f"{ind}a{f}_{ch} = fmaf(x.{comp}, {wlit}, a{f}_{ch});")
                            else:
                                body.append(f"{ind}float a{f}_{ch} = __fmul_rn(x.{comp}, {wlit});")
                                live.add((f, ch))
                    if last_block.get(f) == g:
                        expr = " + ".join(f"a{f}_{ch}" for ch in (0, 1) if (f, ch) in live)
                        body.append(f"{ind}const float v{f} = {expr};")     # the log epilogue is applied once, in mel_flush
                        done.add(f)
                drain()
            assert state["next"] == hi
            body.append("    } break;")
    body += ["    default: break;", "  }", "}", ""]
    L.append(f"static __constant__ unsigned short kRun{name}[{N_WARPS * n_parts}][2] = {{" +
             ", ".join(f"{{{a}, {b}}}" for a, b in runs) + "};")
    if wtab:
        import struct
        vals = [struct.unpack("<f", struct.pack("<I", b))[0] for b in wtab]
        for b, v in zip(wtab, vals):      # nine significant digits round-trip a float32
            assert struct.unpack("<I", struct.pack("<f", float(f"{v:.9e}")))[0] == b
        L.append(f"// {len(wtab) // 4} fully covered blocks run packed ({len(wtab)} of the {len(ent)} weights)")
        L.append(f"static __constant__ float4 kWpk{name}[{len(wtab) // 4}] = {{")
        for i in range(0, len(vals), 4):
            L.append("  {" + ", ".join(f"{v:.9e}f" for v in vals[i:i + 4]) + "},")
        L.append("};")
    return L + body


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
    L += [
        "// One fully covered 4-bin block of a two-chain filter: chain 0 += x.x w.x, chain 1 += x.y w.y, then chain 0 += x.z w.z,",
        "// chain 1 += x.w w.w -- four IEEE fp32 FMAs.  On the device they are two packed fma.rn.f32x2 (SASS FFMA2) whose weight",
        "// pairs sit in uniform registers loaded by ONE LDCU.128 from constant memory; the host (lane emulator) runs the same",
        "// four fmaf.  mel_mul2x2 starts both chains (first pair multiplied, not accumulated).",
        "#ifdef __CUDA_ARCH__",
        "__device__ __forceinline__ unsigned long long mel_pk(float lo, float hi) {",
        "  unsigned long long r;",
        "  asm(\"mov.b64 %0, {%1, %2};\" : \"=l\"(r) : \"f\"(lo), \"f\"(hi));",
        "  return r;",
        "}",
        "__device__ __forceinline__ void mel_fma2x2(float& a0, float& a1, const float4& x, const float4& w) {",
        "  unsigned long long a = mel_pk(a0, a1);",
        "  asm(\"fma.rn.f32x2 %0, %1, %2, %0;\" : \"+l\"(a) : \"l\"(mel_pk(x.x, x.y)), \"l\"(mel_pk(w.x, w.y)));",
        "  asm(\"fma.rn.f32x2 %0, %1, %2, %0;\" : \"+l\"(a) : \"l\"(mel_pk(x.z, x.w)), \"l\"(mel_pk(w.z, w.w)));",
        "  asm(\"mov.b64 {%0, %1}, %2;\" : \"=f\"(a0), \"=f\"(a1) : \"l\"(a));",
        "}",
        "__device__ __forceinline__ void mel_mul2x2(float& a0, float& a1, const float4& x, const float4& w) {",
        "  unsigned long long a;",
        "  asm(\"mul.rn.f32x2 %0, %1, %2;\" : \"=l\"(a) : \"l\"(mel_pk(x.x, x.y)), \"l\"(mel_pk(w.x, w.y)));",
        "  asm(\"fma.rn.f32x2 %0, %1, %2, %0;\" : \"+l\"(a) : \"l\"(mel_pk(x.z, x.w)), \"l\"(mel_pk(w.z, w.w)));",
        "  asm(\"mov.b64 {%0, %1}, %2;\" : \"=f\"(a0), \"=f\"(a1) : \"l\"(a));",
        "}",
        "#else",
        "inline void mel_fma2x2(float& a0, float& a1, const float4& x, const float4& w) {",
        "  a0 = fmaf(x.x, w.x, a0); a1 = fmaf(x.y, w.y, a1); a0 = fmaf(x.z, w.z, a0); a1 = fmaf(x.w, w.w, a1);",
        "}",
        "inline void mel_mul2x2(float& a0, float& a1, const float4& x, const float4& w) {",
        "  a0 = x.x * w.x; a1 = x.y * w.y; a0 = fmaf(x.z, w.z, a0); a1 = fmaf(x.w, w.w, a1);",
        "}",
        "#endif",
        "",
    ] if PACKED else []
    for name, (ident, n_mels, ent) in tables.items():
        if name == "P0":
            L += emit_table(name, n_mels, ent)
        L += emit_table_direct(name, n_mels, ent)
    # kStatic -> direct stage: the baked table's id for every set but P0, whose id (1) is the hybrid
    # form; its direct form is kStatic = kNumBaked + 1 (A/B experiments, BHMEL_OPT_STATIC_MEL = 2)
    L += [f"constexpr int kStaticP0Direct = {len(tables) + 1};"]
    sel = {name: (ident if name != "P0" else len(tables) + 1) for name, (ident, _, _) in tables.items()}
    L += ["template <int kStatic>", "__device__ __forceinline__ constexpr int mel_direct_parts() {"]
    for name in tables:
        L.append(f"  if (kStatic == {sel[name]}) return kParts{name};")
    L += ["  return 0;", "}",
          "// {first filter, filter count} of part `idx` = warp * parts + part",
          "template <int kStatic>", "__device__ __forceinline__ void mel_direct_run(int idx, int& m0, int& ncols) {",
          "  m0 = 0; ncols = 0;"]
    for name in tables:
        L.append(f"  if constexpr (kStatic == {sel[name]}) {{ m0 = kRun{name}[idx][0]; ncols = kRun{name}[idx][1]; }}")
    L += ["}",
          "template <int kStatic>",
          "__device__ __forceinline__ void mel_direct(const float4* __restrict__ prow, float* __restrict__ srow, int mw, int part) {"]
    for name in tables:
        L.append(f"  if constexpr (kStatic == {sel[name]}) mel_direct_{name}(prow, srow, mw, part);")
    L += ["}", "", "}  // namespace bhmel"]
    open(args.out, "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    main()
