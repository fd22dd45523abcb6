#!/usr/bin/env python
"""Print the traced pocketfft DCT programs (pocketfft_dag.py) as CUDA codelets.

Output: vcf_b200/csrc/dct_codelets.cuh (committed; tests/test_codegen.py checks
that it is in sync with this generator).  One straight-line, fully unrolled
device function per length N in {4, 8, 16, 32} and direction, templated on the
working type T (float | double | float2 = two independent transforms per
instruction through Blackwell's packed add/mul/fma.f32x2) and on EXACT:

* EXACT=true  -- every node is one IEEE round-to-nearest operation issued
  through __fadd_rn/__fmul_rn/... (never contracted by nvcc), so the result is
  bit-identical to scipy.fftpack.dct/idct(norm='ortho') in the same precision;
* EXACT=false -- same DAG with plain operators, nvcc may fuse a*b+c.  Outputs
  0 and N/2 of the forward transform contain no multiply-then-add chain, so
  they stay bit-exact (those are the tie-prone rational coefficients of
  SURVEY.md 7.3); the others move by at most a few ulp.

Each codelet works in place on a register array and leaves output k scaled by
sgn[k] * 2^exp[k] (exact lazy scaling, see pocketfft_dag.py); the tables are
emitted as constexpr so the caller folds them into its next multiply.
"""
from __future__ import annotations

import os
import sys

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vcf_b200.codegen import pocketfft_dag as D  # noqa: E402

SIZES = (4, 8, 16, 32)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "dct_codelets.cuh")


def _hex32(v) -> str:
    return float(np.float32(v)).hex() + "f"


def _hex64(v) -> str:
    return float(np.float64(v)).hex()


def _const(sym) -> str:
    return f"vcfb::konst<T>({_hex32(D.const_value(sym, np.float32))}, {_hex64(D.const_value(sym, np.float64))})"


def emit_codelet(n: int, inverse: bool) -> str:
    g, outs = D.trace_dct(n, inverse)
    live = D.live_nodes(g, outs)
    name = f"dct{n}_{'inv' if inverse else 'fwd'}"
    cnt = D.op_counts(g, outs)
    L = []
    L.append(f"// {name}: {cnt['add'] + cnt['sub']} add/sub, {cnt['mul']} mul, {cnt['fma2']} exact-product fma")
    L.append("template <typename T, bool EXACT>")
    L.append(f"__device__ __forceinline__ void {name}(T (&v)[{n}]) {{")
    L.append("  using O = vcfb::Ops<T, EXACT>;")
    ref = {}
    for idx, nd in enumerate(g.nodes):
        if idx not in live:
            continue
        op = nd[0]
        if op == "in":
            ref[idx] = f"i{nd[1]}"
            L.append(f"  const T i{nd[1]} = v[{nd[1]}];")
        elif op == "add":
            ref[idx] = f"t{idx}"
            L.append(f"  const T t{idx} = O::add({ref[nd[1]]}, {ref[nd[2]]});")
        elif op == "sub":
            ref[idx] = f"t{idx}"
            L.append(f"  const T t{idx} = O::sub({ref[nd[1]]}, {ref[nd[2]]});")
        elif op == "mul":
            ref[idx] = f"t{idx}"
            L.append(f"  const T t{idx} = O::mul({ref[nd[1]]}, {_const(nd[2])});")
        elif op == "fma2":
            _, a, k, b, sa, sb = nd
            ref[idx] = f"t{idx}"
            cv = (-1.0 if sa < 0 else 1.0) * float(2.0 ** k)
            c = f"vcfb::konst<T>({cv!r}f, {cv!r})"
            bb = ref[b] if sb > 0 else f"O::neg({ref[b]})"
            L.append(f"  const T t{idx} = O::fma({ref[a]}, {c}, {bb});")
    for k, o in enumerate(outs):
        L.append(f"  v[{k}] = {ref[o.node]};")
    L.append("}")
    exps = ", ".join(str(o.exp) for o in outs)
    sgns = ", ".join(str(o.sign) for o in outs)
    L.append(f"struct {name}_meta {{")
    L.append(f"  static constexpr int N = {n};")
    L.append("  __host__ __device__ static constexpr int exp(int k) {")
    L.append(f"    constexpr int e[{n}] = {{{exps}}};")
    L.append("    return e[k];")
    L.append("  }")
    L.append("  __host__ __device__ static constexpr int sgn(int k) {")
    L.append(f"    constexpr int s[{n}] = {{{sgns}}};")
    L.append("    return s[k];")
    L.append("  }")
    L.append("};")
    return "\n".join(L)


def emit_pruned_codelet(n: int, inverse: bool, nin: int) -> str:
    """The same program with inputs nin..n-1 known to be exact zeros.  Every operation that has a
    zero operand is exact (x + 0 = x, 0 * c = 0, a * 2^k + 0 = a * 2^k), so dropping it changes
    no bit of any non-zero value (only the sign of an exact zero can differ, which no later
    operation turns into a different non-zero value).  Same output scaling as the full codelet."""
    g, outs = D.trace_dct(n, inverse)
    live = D.live_nodes(g, outs)
    name = f"dct{n}_{'inv' if inverse else 'fwd'}_low{nin}"
    L = []
    ZERO = None
    form = {}            # node -> None (exact zero) | (expression name, sign)
    nops = 0

    def emit(idx, expr):
        nonlocal nops
        nops += 1
        L.append(f"  const T t{idx} = {expr};")
        return f"t{idx}"

    def addsub(idx, x, y):          # value of x + y for resolved operands (name, sign)
        (xn, xs), (yn, ys) = x, y
        if xs > 0 and ys > 0:
            return (emit(idx, f"O::add({xn}, {yn})"), 1)
        if xs > 0 and ys < 0:
            return (emit(idx, f"O::sub({xn}, {yn})"), 1)
        if xs < 0 and ys > 0:
            return (emit(idx, f"O::sub({yn}, {xn})"), 1)
        return (emit(idx, f"O::add({xn}, {yn})"), -1)      # (-x) + (-y) = -(x + y), exactly

    for idx, nd in enumerate(g.nodes):
        if idx not in live:
            continue
        op = nd[0]
        if op == "in":
            if nd[1] >= nin:
                form[idx] = ZERO
            else:
                L.append(f"  const T i{nd[1]} = v[{nd[1]}];")
                form[idx] = (f"i{nd[1]}", 1)
        elif op in ("add", "sub"):
            a, b = form[nd[1]], form[nd[2]]
            if op == "sub" and b is not ZERO:
                b = (b[0], -b[1])
            if a is ZERO and b is ZERO:
                form[idx] = ZERO
            elif a is ZERO:
                form[idx] = b
            elif b is ZERO:
                form[idx] = a
            else:
                form[idx] = addsub(idx, a, b)
        elif op == "mul":
            a = form[nd[1]]
            form[idx] = ZERO if a is ZERO else (emit(idx, f"O::mul({a[0]}, {_const(nd[2])})"), a[1])
        elif op == "fma2":
            _, a, k, b, sa, sb = nd
            fa, fb = form[a], form[b]
            if fa is ZERO and fb is ZERO:
                form[idx] = ZERO
            elif fa is ZERO:
                form[idx] = (fb[0], fb[1] * (1 if sb > 0 else -1))
            else:
                cv = (-1.0 if sa * fa[1] < 0 else 1.0) * float(2.0 ** k)
                c = f"vcfb::konst<T>({cv!r}f, {cv!r})"
                if fb is ZERO:
                    form[idx] = (emit(idx, f"O::mul({fa[0]}, {c})"), 1)         # exact: power of two
                else:
                    bb = fb[0] if sb * fb[1] > 0 else f"O::neg({fb[0]})"
                    form[idx] = (emit(idx, f"O::fma({fa[0]}, {c}, {bb})"), 1)
    body = L
    H = [f"// {name}: inputs {nin}..{n - 1} are exact zeros -- {nops} operations",
         "template <typename T, bool EXACT>",
         f"__device__ __forceinline__ void {name}(T (&v)[{n}]) {{",
         "  using O = vcfb::Ops<T, EXACT>;"]
    T = []
    for k, o in enumerate(outs):
        f = form[o.node]
        if f is ZERO:
            T.append(f"  v[{k}] = T(0);")
        elif f[1] > 0:
            T.append(f"  v[{k}] = {f[0]};")
        else:
            T.append(f"  v[{k}] = O::neg({f[0]});")
    return "\n".join(H + body + T + ["}"])


# pruned variants emitted: (n, inverse, number of leading inputs that may be non-zero)
PRUNED = ((8, True, 2), (8, True, 4), (16, True, 2), (16, True, 4))


def generate() -> str:
    parts = [
        "// GENERATED by vcf_b200/codegen/gen_cuda.py -- do not edit.",
        "// Operation-exact restatement of pocketfft's DCT-II/III (scipy.fftpack.dct/idct,",
        "// norm='ortho'); see vcf_b200/codegen/pocketfft_dag.py for provenance.",
        "#pragma once",
        '#include "exact_ops.cuh"',
        "",
        "namespace vcfb {",
        "",
    ]
    for n in SIZES:
        for inv in (False, True):
            parts.append(emit_codelet(n, inv))
            parts.append("")
    for n, inv, nin in PRUNED:
        parts.append(emit_pruned_codelet(n, inv, nin))
        parts.append("")
    # dispatcher
    parts.append("template <int N, bool INV> struct Dct;")
    for n in SIZES:
        for inv in (False, True):
            nm = f"dct{n}_{'inv' if inv else 'fwd'}"
            parts.append(f"template <> struct Dct<{n}, {'true' if inv else 'false'}> {{")
            parts.append(f"  using meta = {nm}_meta;")
            parts.append("  template <typename T, bool EXACT>")
            parts.append(f"  __device__ __forceinline__ static void run(T (&v)[{n}]) {{ {nm}<T, EXACT>(v); }}")
            parts.append("};")
    parts.append("")
    parts.append("}  // namespace vcfb")
    parts.append("")
    return "\n".join(parts)


def main():
    src = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        sys.exit(0 if cur == src else 1)
    with open(OUT, "w") as f:
        f.write(src)
    print(f"wrote {OUT} ({len(src.splitlines())} lines)")


if __name__ == "__main__":
    main()
