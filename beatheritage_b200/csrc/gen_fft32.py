#!/usr/bin/env python3
"""Generates fft32_gen.h: straight-line, fully register-resident 32-point complex FFT passes.

The 1024-point complex transform of one frame PAIR (frame 2j in the real part, frame 2j+1 in
the imaginary part) is split 32 x 32: pass A is a 32-point DFT inside each thread over the
slow sample index, pass B a 32-point DFT inside each thread over the fast sample index after a
warp-wide transpose.  Everything here is compile-time: twiddles are literal constants, all
array indices are static (so the arrays live in registers), and butterflies with a non-trivial
twiddle use the 6-FMA (Linzer-Feig) form:

    X  = E + W*O      Xr = fma(c, Or, fma( s, Oi, Er)),  Xi = fma(c, Oi, fma(-s, Or, Ei))
    X' = E - W*O      X'r = fma(2, Er, -Xr),             X'i = fma(2, Ei, -Xi)        (W = c - i s)

pass A fuses the (half-)Hann window into its first butterfly stage (3 instructions per real
pair), pass B fuses the inter-pass twiddle W_1024^(k1*n2) into its first stage (10 per pair).

Usage: python gen_fft32.py > fft32_gen.h   (build.py runs it; the output is not committed)
       python gen_fft32.py --packed > ...   additionally emits fft32x2_pass_a/b, the same schedule on
       packed (re, im) register pairs with Blackwell's FFMA2/FADD2/FMUL2 (experiment)
"""
import math
import sys

N = 32


class Emit:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.count = 0          # arithmetic instructions (roughly: one per emitted op)

    def tmp(self, expr):
        name = f"t{self.n}"
        self.n += 1
        self.count += 1
        self.lines.append(f"    const float {name} = {expr};")
        return name


def flit(x):
    return f"{x:.9e}f"


def butterfly(em, E, O, k, n):
    """E, O: (re, im) names. Returns (X, X') for twiddle W_n^k = exp(-2 pi i k / n)."""
    er, ei = E
    orr, oi = O
    k %= n
    if k == 0:
        return ((em.tmp(f"{er} + {orr}"), em.tmp(f"{ei} + {oi}")),
                (em.tmp(f"{er} - {orr}"), em.tmp(f"{ei} - {oi}")))
    if 4 * k == n:       # W = -i : W*O = (Oi, -Or)
        return ((em.tmp(f"{er} + {oi}"), em.tmp(f"{ei} - {orr}")),
                (em.tmp(f"{er} - {oi}"), em.tmp(f"{ei} + {orr}")))
    th = 2.0 * math.pi * k / n
    c, s = math.cos(th), math.sin(th)
    xr = em.tmp(f"fmaf({flit(c)}, {orr}, fmaf({flit(s)}, {oi}, {er}))")
    xi = em.tmp(f"fmaf({flit(c)}, {oi}, fmaf({flit(-s)}, {orr}, {ei}))")
    em.count += 2       # the two nested fmaf
    yr = em.tmp(f"fmaf(2.0f, {er}, -{xr})")
    yi = em.tmp(f"fmaf(2.0f, {ei}, -{xi})")
    return (xr, xi), (yr, yi)


def dit(em, vals, base2):
    """Recursive radix-2 decimation in time. vals: list of complex names in natural input order.
    base2(i0, i1) emits the size-2 transform of inputs i0, i1 (which may be fused with a
    pre-multiplication) and returns (X0, X1)."""
    n = len(vals)
    if n == 2:
        return list(base2(vals[0], vals[1]))
    E = dit(em, vals[0::2], base2)
    O = dit(em, vals[1::2], base2)
    out = [None] * n
    for k in range(n // 2):
        out[k], out[k + n // 2] = butterfly(em, E[k], O[k], k, n)
    return out


def gen_pass_a(em):
    """in: v[36] real samples (v[m] -> frame A sample m, v[m+4] -> frame B sample m),
    w[32] window (already scaled by 0.5).  out: re/im[k1], k1 natural order."""
    def base2(i0, i1):          # i0, i1 are integer sample slots m, m+16
        outs = []
        for off in (0, 4):      # real part from frame A, imaginary part from frame B
            p = em.tmp(f"w[{i0}] * v[{i0 + off}]")
            s = em.tmp(f"fmaf(w[{i1}], v[{i1 + off}], {p})")
            d = em.tmp(f"fmaf(-w[{i1}], v[{i1 + off}], {p})")
            outs.append((s, d))
        (sr, dr), (si, di) = outs
        return (sr, si), (dr, di)
    return dit(em, list(range(N)), base2)


def gen_pass_b(em):
    """in: ur/ui[n2] (transposed pass-A output), tr/ti[n2] = W_1024^(lane*n2) (cos, -sin).
    out: re/im[k2] natural order."""
    def base2(i0, i1):
        # A = tw[i0]*u[i0] ; out0 = A + tw[i1]*u[i1] ; out1 = 2A - out0
        if i0 == 0:
            ar, ai = "ur[0]", "ui[0]"          # W^0 = 1
        else:
            ar = em.tmp(f"fmaf(tr[{i0}], ur[{i0}], -(ti[{i0}] * ui[{i0}]))")
            ai = em.tmp(f"fmaf(tr[{i0}], ui[{i0}], ti[{i0}] * ur[{i0}])")
            em.count += 2
        xr = em.tmp(f"fmaf(tr[{i1}], ur[{i1}], fmaf(-ti[{i1}], ui[{i1}], {ar}))")
        xi = em.tmp(f"fmaf(tr[{i1}], ui[{i1}], fmaf(ti[{i1}], ur[{i1}], {ai}))")
        em.count += 2
        yr = em.tmp(f"fmaf(2.0f, {ar}, -{xr})")
        yi = em.tmp(f"fmaf(2.0f, {ai}, -{xi})")
        return (xr, xi), (yr, yi)
    return dit(em, list(range(N)), base2)


# --------------------------------------------------------------------------------------------
# Packed flavour: the same butterfly schedule emitted with Blackwell's packed fp32 instructions
# (fma/add/mul.rn.f32x2 -> SASS FFMA2/FADD2/FMUL2).  A complex value lives in one 64-bit register
# pair (re, im); ptxas folds the half swaps / per-half negations / scalar broadcasts that the
# mov.b64 pack-unpack sequences below express into operand modifiers (.LO_HI, .NP, R.F32), so a
# butterfly costs half the issue slots of the scalar form.  The arithmetic (operation order and
# rounding points) is IDENTICAL to the scalar flavour, so both give the same bits.
class EmitX2(Emit):
    def tmp64(self, expr):
        name = f"z{self.n}"
        self.n += 1
        self.count += 1
        self.lines.append(f"    const unsigned long long {name} = {expr};")
        return name


def x2_swap_np(o):      # (o.im, -o.re)
    return f"pk_im_nre({o})"


def x2_swap_pn(o):      # (-o.im, o.re)
    return f"pk_nim_re({o})"


def butterfly_x2(em, E, O, k, n):
    k %= n
    if k == 0:
        return em.tmp64(f"add2({E}, {O})"), em.tmp64(f"sub2({E}, {O})")
    if 4 * k == n:       # W = -i : W*O = (Oi, -Or)
        return em.tmp64(f"add2({E}, {x2_swap_np(O)})"), em.tmp64(f"add2({E}, {x2_swap_pn(O)})")
    th = 2.0 * math.pi * k / n
    c, s = math.cos(th), math.sin(th)
    inner = em.tmp64(f"fma2({x2_swap_np(O)}, bc({flit(s)}), {E})")     # (s*Oi + Er, -s*Or + Ei)
    x = em.tmp64(f"fma2({O}, bc({flit(c)}), {inner})")
    y = em.tmp64(f"fma2({E}, bc(2.0f), neg2({x}))")
    return x, y


def dit_x2(em, vals, base2):
    n = len(vals)
    if n == 2:
        return list(base2(vals[0], vals[1]))
    E = dit_x2(em, vals[0::2], base2)
    O = dit_x2(em, vals[1::2], base2)
    out = [None] * n
    for k in range(n // 2):
        out[k], out[k + n // 2] = butterfly_x2(em, E[k], O[k], k, n)
    return out


def gen_pass_a_x2(em):
    """in: v[36], w[32] (as in the scalar flavour).  out: z[k1] packed (re, im)."""
    def base2(i0, i1):
        p = em.tmp64(f"mul2(pk(v[{i0}], v[{i0 + 4}]), bc(w[{i0}]))")
        z1 = f"pk(v[{i1}], v[{i1 + 4}])"
        s = em.tmp64(f"fma2({z1}, bc(w[{i1}]), {p})")
        d = em.tmp64(f"fma2({z1}, bc(-w[{i1}]), {p})")
        return s, d
    return dit_x2(em, list(range(N)), base2)


def gen_pass_b_x2(em):
    """in: u[n2] packed, tr/ti[n2].  out: z[k2] packed."""
    def base2(i0, i1):
        if i0 == 0:
            a = "u[0]"
        else:
            m = em.tmp64(f"mul2(pk_swap(u[{i0}]), pk(-ti[{i0}], ti[{i0}]))")           # (-ti*ui, ti*ur)
            a = em.tmp64(f"fma2(u[{i0}], bc(tr[{i0}]), {m})")
        inner = em.tmp64(f"fma2(pk_swap(u[{i1}]), pk(-ti[{i1}], ti[{i1}]), {a})")
        x = em.tmp64(f"fma2(u[{i1}], bc(tr[{i1}]), {inner})")
        y = em.tmp64(f"fma2({a}, bc(2.0f), neg2({x}))")
        return x, y
    return dit_x2(em, list(range(N)), base2)


X2_HELPERS = r"""
// ---- packed fp32 pair helpers (device only; sm_100+) ----
__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long bc(float s) { return pk(s, s); }
__device__ __forceinline__ unsigned long long pk_swap(unsigned long long v) { float a, b; upk(v, a, b); return pk(b, a); }
__device__ __forceinline__ unsigned long long pk_im_nre(unsigned long long v) { float a, b; upk(v, a, b); return pk(b, -a); }
__device__ __forceinline__ unsigned long long pk_nim_re(unsigned long long v) { float a, b; upk(v, a, b); return pk(-b, a); }
__device__ __forceinline__ unsigned long long neg2(unsigned long long v) { float a, b; upk(v, a, b); return pk(-a, -b); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
"""


def main():
    out = []
    out.append("// GENERATED by gen_fft32.py -- do not edit.  Straight-line 32-point FFT passes.")
    out.append("#pragma once")
    out.append("#ifndef BHMEL_HD")
    out.append("#define BHMEL_HD inline")
    out.append("#endif")
    out.append("namespace bhmel {")
    for name, gen, sig in (
        ("fft32_pass_a", gen_pass_a,
         "const float (&v)[36], const float (&w)[32], float (&re)[32], float (&im)[32]"),
        ("fft32_pass_b", gen_pass_b,
         "const float (&ur)[32], const float (&ui)[32], const float (&tr)[32], const float (&ti)[32], "
         "float (&re)[32], float (&im)[32]"),
    ):
        em = Emit()
        res = gen(em)
        out.append(f"// ~{em.count} arithmetic instructions")
        out.append(f"BHMEL_HD void {name}({sig}) {{")
        out.extend(em.lines)
        for k, (r, i) in enumerate(res):
            out.append(f"    re[{k}] = {r}; im[{k}] = {i};")
        out.append("}")
    if "--packed" not in sys.argv:      # the packed flavour is an experiment, off by default
        out.append("}  // namespace bhmel")
        sys.stdout.write("\n".join(out) + "\n")
        return
    out.append("#ifdef __CUDACC__")
    out.append(X2_HELPERS)
    for name, gen, sig in (
        ("fft32x2_pass_a", gen_pass_a_x2,
         "const float (&v)[36], const float (&w)[32], unsigned long long (&z)[32]"),
        ("fft32x2_pass_b", gen_pass_b_x2,
         "const unsigned long long (&u)[32], const float (&tr)[32], const float (&ti)[32], unsigned long long (&z)[32]"),
    ):
        em = EmitX2()
        res = gen(em)
        out.append(f"// ~{em.count} packed instructions")
        out.append(f"__device__ __forceinline__ void {name}({sig}) {{")
        out.extend(em.lines)
        for k, r in enumerate(res):
            out.append(f"    z[{k}] = {r};")
        out.append("}")
    out.append("#endif  // __CUDACC__")
    out.append("}  // namespace bhmel")
    sys.stdout.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
