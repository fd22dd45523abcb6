"""Operation-exact restatement of pocketfft's real-data DCT-II / DCT-III as a DAG.

Why this exists.  The reference's block DCT is ``scipy.fftpack.dct/idct(...,
norm='ortho')`` applied per axis (external DCT2D package, called at
/root/reference/src/2D-DCT.py:303 and :440; in-repo corroboration
src/IPP_DCT.py:257-263).  scipy implements it with pocketfft (C++, bundled in
scipy 1.18.1 as scipy.fft._pocketfft.pypocketfft; not part of /root/reference).
With 8-bit input the coefficients at (u,v) in {0,B/2}^2 are exact rationals that
often land exactly on a multiple of the quantisation step, so the truncated
index depends on the last-ulp rounding of that library.  The
only way to be *bit-exact* with the reference on a GPU -- without any CPU
fallback -- is to execute the same sequence of individually rounded IEEE
operations.  This module restates pocketfft's published algorithm for that
purpose:

* ``T_dcst23<T0>::exec`` (type 2 and type 3, cosine),
* ``rfftp<T0>`` with the radix-4 / radix-2 passes ``radf4, radf2, radb4,
  radb2`` and its factorisation order,
* the twiddle generator ``sincos_2pibyn`` (double-precision two-table product,
  rounded to T0) and the ``1/sqrt(2N)`` normalisation computed in long double.

It is *traced*, not executed: the passes run on symbolic values and record a
straight-line program (add / sub / mul-by-constant / fma-with-power-of-two).
Two exactness-preserving simplifications are applied while tracing:

* multiplication by +-2^k is exact, so it is carried as a lazy (sign, exponent)
  pair on each value instead of being emitted.  ``a*2^k + b`` with different
  lazy exponents becomes one fused multiply-add whose product is exact -- the
  result is bit-identical to the reference's separately rounded sequence;
* IEEE negation is exact and round-to-nearest is sign-symmetric, so signs are
  carried lazily too and additions turn into subtractions where needed.

The same DAG is (a) evaluated with numpy in float32/float64 and compared
bit-for-bit with scipy in tests/test_pocketfft_dag.py, and (b) printed as CUDA
device code by ``gen_cuda.py``.  No scaling is ever applied to an *inexact*
product, so nothing here changes a rounding.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

# ----------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------

_LD = np.longdouble
_PI_LD = _LD("3.141592653589793238462643383279502884197")
_SQRT2_LD = _LD("1.414213562373095048801688724209698")
_HSQT2_LD = _LD("0.707106781186547524400844362104849")


class SinCos2PiByN:
    """pocketfft ``sincos_2pibyn<T>`` for T in {float, double} (Thigh = double):
    exp(2*pi*i*idx/N) as the product of two table entries evaluated in double,
    then rounded to T."""

    def __init__(self, n: int):
        self.N = n
        ang = float(_LD(0.25) * _PI_LD / _LD(n))
        nval = (n + 2) // 2
        shift = 1
        while (1 << shift) * (1 << shift) < nval:
            shift += 1
        self.shift = shift
        self.mask = (1 << shift) - 1
        self.v1 = [(1.0, 0.0)] + [self._calc(i, n, ang) for i in range(1, self.mask + 1)]
        n2 = (nval + self.mask) // (self.mask + 1)
        self.v2 = [(1.0, 0.0)] + [self._calc(i * (self.mask + 1), n, ang) for i in range(1, n2)]

    @staticmethod
    def _calc(x: int, n: int, ang: float):
        x <<= 3
        if x < 4 * n:
            if x < 2 * n:
                if x < n:
                    return (math.cos(float(x) * ang), math.sin(float(x) * ang))
                return (math.sin(float(2 * n - x) * ang), math.cos(float(2 * n - x) * ang))
            x -= 2 * n
            if x < n:
                return (-math.sin(float(x) * ang), math.cos(float(x) * ang))
            return (-math.cos(float(2 * n - x) * ang), math.sin(float(2 * n - x) * ang))
        x = 8 * n - x
        if x < 2 * n:
            if x < n:
                return (math.cos(float(x) * ang), -math.sin(float(x) * ang))
            return (math.sin(float(2 * n - x) * ang), -math.cos(float(2 * n - x) * ang))
        x -= 6 * n
        if x < n:
            return (-math.sin(float(x) * ang), -math.cos(float(x) * ang))
        return (-math.cos(float(2 * n - x) * ang), -math.sin(float(2 * n - x) * ang))

    def get(self, idx: int):
        """(re, im) in double, i.e. before the final rounding to T."""
        if 2 * idx <= self.N:
            x1 = self.v1[idx & self.mask]
            x2 = self.v2[idx >> self.shift]
            return (x1[0] * x2[0] - x1[1] * x2[1], x1[0] * x2[1] + x1[1] * x2[0])
        idx = self.N - idx
        x1 = self.v1[idx & self.mask]
        x2 = self.v2[idx >> self.shift]
        return (x1[0] * x2[0] - x1[1] * x2[1], -(x1[0] * x2[1] + x1[1] * x2[0]))


def const_value(sym, dtype):
    """Numeric value of a symbolic constant in working dtype (np.float32/64)."""
    dt = np.dtype(dtype).type
    kind = sym[0]
    if kind == "rtw":            # rfftp twiddle: ("rtw", N, idx, 0|1)
        _, n, idx, part = sym
        return dt(SinCos2PiByN(n).get(idx)[part])
    if kind == "dtw":            # T_dcst23 twiddle[i] = tw4N[i+1].r : ("dtw", N, i)
        _, n, i = sym
        return dt(SinCos2PiByN(4 * n).get(i + 1)[0])
    if kind == "fct":            # 1/sqrt(2N) in long double, rounded to T
        _, n = sym
        return dt(_LD(1) / np.sqrt(_LD(2 * n)))
    if kind == "scaled":         # exact variant of another constant: sign * 2^k * value
        _, inner, sign, k = sym
        return dt(sign * 2.0 ** k) * const_value(inner, dtype)
    if kind == "sqrt2":
        return dt(_SQRT2_LD)
    if kind == "hsqt2":
        return dt(_HSQT2_LD)
    raise KeyError(sym)


def _pow2_exponent(x: float):
    """k if x == 2^k exactly, else None."""
    if x <= 0:
        return None
    m, e = math.frexp(x)
    return e - 1 if m == 0.5 else None


# ----------------------------------------------------------------------------
# DAG
# ----------------------------------------------------------------------------

@dataclass(frozen=True)
class Val:
    """sign * 2^exp * node  (sign and exp are lazy, exact)."""
    node: int
    sign: int = 1
    exp: int = 0


class Graph:
    """Straight-line program.  Node kinds:
    ("in", i) | ("add", a, b) | ("sub", a, b) | ("mul", a, csym)
    | ("fma2", a, k, b, sa, sb)  meaning  sa*a*2^k + sb*b  (one rounding)."""

    def __init__(self):
        self.nodes = []
        self._cse = {}

    def _mk(self, key):
        n = self._cse.get(key)
        if n is None:
            n = len(self.nodes)
            self.nodes.append(key)
            self._cse[key] = n
        return n

    # -- value constructors ---------------------------------------------------
    def input(self, i: int, exp: int = 0) -> Val:
        return Val(self._mk(("in", i)), 1, exp)

    @staticmethod
    def neg(a: Val) -> Val:
        return Val(a.node, -a.sign, a.exp)

    @staticmethod
    def scale2(a: Val, k: int) -> Val:
        return Val(a.node, a.sign, a.exp + k)

    def mulc(self, a: Val, sym) -> Val:
        return Val(self._mk(("mul", a.node, sym)), a.sign, a.exp)

    def add(self, a: Val, b: Val) -> Val:
        if a.exp == b.exp:
            if a.sign == b.sign:
                x, y = sorted((a.node, b.node))     # IEEE add commutes
                return Val(self._mk(("add", x, y)), a.sign, a.exp)
            if a.sign > 0:
                return Val(self._mk(("sub", a.node, b.node)), 1, a.exp)
            return Val(self._mk(("sub", b.node, a.node)), 1, a.exp)
        # different lazy exponents: scale the one with the larger exponent
        if a.exp < b.exp:
            a, b = b, a
        k = a.exp - b.exp
        sa, sb = a.sign, b.sign
        s = 1
        if sa < 0 and sb < 0:
            s, sa, sb = -1, 1, 1
        return Val(self._mk(("fma2", a.node, k, b.node, sa, sb)), s, b.exp)

    def sub(self, a: Val, b: Val) -> Val:
        return self.add(a, self.neg(b))

    def pm(self, c: Val, d: Val):
        """pocketfft PM(a,b,c,d): a=c+d; b=c-d."""
        return self.add(c, d), self.sub(c, d)

    def mulpm(self, c, d, e: Val, f: Val):
        """pocketfft MULPM(a,b,c,d,e,f): a=c*e+d*f; b=c*f-d*e  (c,d constants)."""
        a = self.add(self.mulc(e, c), self.mulc(f, d))
        b = self.sub(self.mulc(f, c), self.mulc(e, d))
        return a, b


# ----------------------------------------------------------------------------
# rfftp passes (pocketfft_hdronly.h, class rfftp)
# ----------------------------------------------------------------------------

def _factorize(n: int):
    fact = []
    length = n
    while length % 4 == 0:
        fact.append(4)
        length >>= 2
    if length % 2 == 0:
        length >>= 1
        fact.append(2)
        fact[0], fact[-1] = fact[-1], fact[0]
    if length != 1:
        raise ValueError("only power-of-two lengths are restated")
    return fact


class _RFFTP:
    def __init__(self, g: Graph, n: int):
        self.g = g
        self.n = n
        self.fact = _factorize(n)
        # twiddle symbols per factor: tw[k][(j-1)*(ido-1)+2i-2] = twid[j*l1*i].r, +1 -> .i
        self.tw = []
        l1 = 1
        for k, ip in enumerate(self.fact):
            ido = n // (l1 * ip)
            tw = {}
            if k < len(self.fact) - 1:
                for j in range(1, ip):
                    for i in range(1, (ido - 1) // 2 + 1):
                        tw[(j - 1) * (ido - 1) + 2 * i - 2] = ("rtw", n, j * l1 * i, 0)
                        tw[(j - 1) * (ido - 1) + 2 * i - 1] = ("rtw", n, j * l1 * i, 1)
            self.tw.append(tw)
            l1 *= ip

    # --- forward (r2hc) -----------------------------------------------------
    def radf2(self, ido, l1, cc, wa):
        g = self.g
        ch = [None] * self.n
        WA = lambda x, i: wa[i + x * (ido - 1)]
        CC = lambda a, b, c: cc[a + ido * (b + l1 * c)]
        def CH(a, b, c, v): ch[a + ido * (b + 2 * c)] = v
        for k in range(l1):
            s, d = g.pm(CC(0, k, 0), CC(0, k, 1))
            CH(0, 0, k, s); CH(ido - 1, 1, k, d)
        if ido & 1 == 0:
            for k in range(l1):
                CH(0, 1, k, g.neg(CC(ido - 1, k, 1)))
                CH(ido - 1, 0, k, CC(ido - 1, k, 0))
        if ido <= 2:
            return ch
        for k in range(l1):
            for i in range(2, ido, 2):
                ic = ido - i
                tr2, ti2 = g.mulpm(WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
                s, d = g.pm(CC(i - 1, k, 0), tr2)
                CH(i - 1, 0, k, s); CH(ic - 1, 1, k, d)
                s, d = g.pm(ti2, CC(i, k, 0))
                CH(i, 0, k, s); CH(ic, 1, k, d)
        return ch

    def radf4(self, ido, l1, cc, wa):
        g = self.g
        ch = [None] * self.n
        hsqt2 = ("hsqt2",)
        WA = lambda x, i: wa[i + x * (ido - 1)]
        CC = lambda a, b, c: cc[a + ido * (b + l1 * c)]
        def CH(a, b, c, v): ch[a + ido * (b + 4 * c)] = v
        for k in range(l1):
            tr1, d = g.pm(CC(0, k, 3), CC(0, k, 1)); CH(0, 2, k, d)
            tr2, d = g.pm(CC(0, k, 0), CC(0, k, 2)); CH(ido - 1, 1, k, d)
            s, d = g.pm(tr2, tr1); CH(0, 0, k, s); CH(ido - 1, 3, k, d)
        if ido & 1 == 0:
            for k in range(l1):
                ti1 = g.neg(g.mulc(g.add(CC(ido - 1, k, 1), CC(ido - 1, k, 3)), hsqt2))
                tr1 = g.mulc(g.sub(CC(ido - 1, k, 1), CC(ido - 1, k, 3)), hsqt2)
                s, d = g.pm(CC(ido - 1, k, 0), tr1); CH(ido - 1, 0, k, s); CH(ido - 1, 2, k, d)
                s, d = g.pm(ti1, CC(ido - 1, k, 2)); CH(0, 3, k, s); CH(0, 1, k, d)
        if ido <= 2:
            return ch
        for k in range(l1):
            for i in range(2, ido, 2):
                ic = ido - i
                cr2, ci2 = g.mulpm(WA(0, i - 2), WA(0, i - 1), CC(i - 1, k, 1), CC(i, k, 1))
                cr3, ci3 = g.mulpm(WA(1, i - 2), WA(1, i - 1), CC(i - 1, k, 2), CC(i, k, 2))
                cr4, ci4 = g.mulpm(WA(2, i - 2), WA(2, i - 1), CC(i - 1, k, 3), CC(i, k, 3))
                tr1, tr4 = g.pm(cr4, cr2)
                ti1, ti4 = g.pm(ci2, ci4)
                tr2, tr3 = g.pm(CC(i - 1, k, 0), cr3)
                ti2, ti3 = g.pm(CC(i, k, 0), ci3)
                s, d = g.pm(tr2, tr1); CH(i - 1, 0, k, s); CH(ic - 1, 3, k, d)
                s, d = g.pm(ti1, ti2); CH(i, 0, k, s); CH(ic, 3, k, d)
                s, d = g.pm(tr3, ti4); CH(i - 1, 2, k, s); CH(ic - 1, 1, k, d)
                s, d = g.pm(tr4, ti3); CH(i, 2, k, s); CH(ic, 1, k, d)
        return ch

    # --- backward (hc2r) ----------------------------------------------------
    def radb2(self, ido, l1, cc, wa):
        g = self.g
        ch = [None] * self.n
        WA = lambda x, i: wa[i + x * (ido - 1)]
        CC = lambda a, b, c: cc[a + ido * (b + 2 * c)]
        def CH(a, b, c, v): ch[a + ido * (b + l1 * c)] = v
        for k in range(l1):
            s, d = g.pm(CC(0, 0, k), CC(ido - 1, 1, k)); CH(0, k, 0, s); CH(0, k, 1, d)
        if ido & 1 == 0:
            for k in range(l1):
                CH(ido - 1, k, 0, g.scale2(CC(ido - 1, 0, k), 1))
                CH(ido - 1, k, 1, g.neg(g.scale2(CC(0, 1, k), 1)))
        if ido <= 2:
            return ch
        for k in range(l1):
            for i in range(2, ido, 2):
                ic = ido - i
                s, tr2 = g.pm(CC(i - 1, 0, k), CC(ic - 1, 1, k)); CH(i - 1, k, 0, s)
                ti2, d = g.pm(CC(i, 0, k), CC(ic, 1, k)); CH(i, k, 0, d)
                a, b = g.mulpm(WA(0, i - 2), WA(0, i - 1), ti2, tr2)
                CH(i, k, 1, a); CH(i - 1, k, 1, b)
        return ch

    def radb4(self, ido, l1, cc, wa):
        g = self.g
        ch = [None] * self.n
        sqrt2 = ("sqrt2",)
        WA = lambda x, i: wa[i + x * (ido - 1)]
        CC = lambda a, b, c: cc[a + ido * (b + 4 * c)]
        def CH(a, b, c, v): ch[a + ido * (b + l1 * c)] = v
        for k in range(l1):
            tr2, tr1 = g.pm(CC(0, 0, k), CC(ido - 1, 3, k))
            tr3 = g.scale2(CC(ido - 1, 1, k), 1)
            tr4 = g.scale2(CC(0, 2, k), 1)
            s, d = g.pm(tr2, tr3); CH(0, k, 0, s); CH(0, k, 2, d)
            s, d = g.pm(tr1, tr4); CH(0, k, 3, s); CH(0, k, 1, d)
        if ido & 1 == 0:
            for k in range(l1):
                ti1, ti2 = g.pm(CC(0, 3, k), CC(0, 1, k))
                tr2, tr1 = g.pm(CC(ido - 1, 0, k), CC(ido - 1, 2, k))
                CH(ido - 1, k, 0, g.scale2(tr2, 1))                       # tr2+tr2
                CH(ido - 1, k, 1, g.mulc(g.sub(tr1, ti1), sqrt2))
                CH(ido - 1, k, 2, g.scale2(ti2, 1))                       # ti2+ti2
                CH(ido - 1, k, 3, g.neg(g.mulc(g.add(tr1, ti1), sqrt2)))
        if ido <= 2:
            return ch
        for k in range(l1):
            for i in range(2, ido, 2):
                ic = ido - i
                tr2, tr1 = g.pm(CC(i - 1, 0, k), CC(ic - 1, 3, k))
                ti1, ti2 = g.pm(CC(i, 0, k), CC(ic, 3, k))
                tr4, ti3 = g.pm(CC(i, 2, k), CC(ic, 1, k))
                tr3, ti4 = g.pm(CC(i - 1, 2, k), CC(ic - 1, 1, k))
                s, cr3 = g.pm(tr2, tr3); CH(i - 1, k, 0, s)
                s, ci3 = g.pm(ti2, ti3); CH(i, k, 0, s)
                cr4, cr2 = g.pm(tr1, tr4)
                ci2, ci4 = g.pm(ti1, ti4)
                a, b = g.mulpm(WA(0, i - 2), WA(0, i - 1), ci2, cr2); CH(i, k, 1, a); CH(i - 1, k, 1, b)
                a, b = g.mulpm(WA(1, i - 2), WA(1, i - 1), ci3, cr3); CH(i, k, 2, a); CH(i - 1, k, 2, b)
                a, b = g.mulpm(WA(2, i - 2), WA(2, i - 1), ci4, cr4); CH(i, k, 3, a); CH(i - 1, k, 3, b)
        return ch

    def exec(self, c, fct_sym, r2hc: bool):
        n, nf = self.n, len(self.fact)
        p = list(c)
        if r2hc:
            l1 = n
            for k1 in range(nf):
                k = nf - k1 - 1
                ip = self.fact[k]
                ido = n // l1
                l1 //= ip
                p = (self.radf4 if ip == 4 else self.radf2)(ido, l1, p, self.tw[k])
        else:
            l1 = 1
            for k in range(nf):
                ip = self.fact[k]
                ido = n // (ip * l1)
                p = (self.radb4 if ip == 4 else self.radb2)(ido, l1, p, self.tw[k])
                l1 *= ip
        # copy_and_norm
        k2 = _pow2_exponent(float(const_value(fct_sym, np.float64)))
        if k2 is not None and _pow2_exponent(float(const_value(fct_sym, np.float32))) == k2:
            return [self.g.scale2(v, k2) for v in p]
        return [self.g.mulc(v, fct_sym) for v in p]


# ----------------------------------------------------------------------------
# T_dcst23::exec, cosine, ortho=True  (scipy.fftpack.dct/idct norm='ortho')
# ----------------------------------------------------------------------------

def trace_dct(n: int, inverse: bool, in_exp: int = 0):
    """Return (graph, outputs) for the length-n orthonormal DCT-II
    (inverse=False: scipy.fftpack.dct(x, norm='ortho')) or DCT-III
    (inverse=True: scipy.fftpack.idct(x, norm='ortho')).
    Inputs carry the lazy exponent in_exp (value = node * 2^in_exp)."""
    g = Graph()
    c = [g.input(i, in_exp) for i in range(n)]
    plan = _RFFTP(g, n)
    fct = ("fct", n)
    ns2 = (n + 1) // 2
    tw = lambda i: ("dtw", n, i)
    if not inverse:                                   # type 2
        c[0] = g.scale2(c[0], 1)
        if n & 1 == 0:
            c[n - 1] = g.scale2(c[n - 1], 1)
        for k in range(1, n - 1, 2):                  # MPINPLACE(c[k+1], c[k])
            a, b = c[k + 1], c[k]
            c[k + 1] = g.sub(a, b)
            c[k] = g.add(a, b)
        c = plan.exec(c, fct, False)
        k, kc = 1, n - 1
        while k < ns2:
            t1 = g.add(g.mulc(c[kc], tw(k - 1)), g.mulc(c[k], tw(kc - 1)))
            t2 = g.sub(g.mulc(c[k], tw(k - 1)), g.mulc(c[kc], tw(kc - 1)))
            c[k] = g.scale2(g.add(t1, t2), -1)
            c[kc] = g.scale2(g.sub(t1, t2), -1)
            k += 1; kc -= 1
        if n & 1 == 0:
            c[ns2] = g.mulc(c[ns2], tw(ns2 - 1))
        c[0] = g.scale2(g.mulc(c[0], ("sqrt2",)), -1)   # c[0] *= sqrt2*0.5
    else:                                             # type 3
        c[0] = g.mulc(c[0], ("sqrt2",))
        k, kc = 1, n - 1
        while k < ns2:
            t1 = g.add(c[k], c[kc])
            t2 = g.sub(c[k], c[kc])
            c[k] = g.add(g.mulc(t2, tw(k - 1)), g.mulc(t1, tw(kc - 1)))
            c[kc] = g.sub(g.mulc(t1, tw(k - 1)), g.mulc(t2, tw(kc - 1)))
            k += 1; kc -= 1
        if n & 1 == 0:
            c[ns2] = g.scale2(g.mulc(c[ns2], tw(ns2 - 1)), 1)   # *= 2*twiddle
        c = plan.exec(c, fct, True)
        for k in range(1, n - 1, 2):                  # MPINPLACE(c[k], c[k+1])
            a, b = c[k], c[k + 1]
            c[k] = g.sub(a, b)
            c[k + 1] = g.add(a, b)
    return make_unfusable(g, c)


# ----------------------------------------------------------------------------
# make multiply -> add chains unfusable
# ----------------------------------------------------------------------------

def _scaled(sym, sign, k):
    if sym[0] == "scaled":
        return _scaled(sym[1], sign * sym[2], k + sym[3])
    return sym if (sign == 1 and k == 0) else ("scaled", sym, sign, k)


def make_unfusable(g: Graph, outs):
    """Rewrite the program so that no addition consumes the result of a constant multiply.

    ptxas 12.9 fuses a packed ``mul.rn.f32x2`` into a following ``add.rn.f32x2`` (and into
    ``fma.rn.f32x2`` with a factor of +1) despite the explicit rounding modifier, and
    ``-fmad=false`` does not stop it; it does not touch ``fma(p, -1, a)``.  So every
    ``a + c*x`` becomes ``a - (-c)*x`` (IEEE negation is exact and round-to-nearest is
    sign-symmetric: same bits), and ``2^k*(c*x) + b`` becomes ``(2^k*c)*x + b`` first (the
    power of two commutes with the rounding of the product).  Multiplies feeding more than
    one consumer get a private negated copy.  The result is a new (graph, outputs)."""
    live = live_nodes(g, outs)
    ng = Graph()
    m = {}

    def mulneg(idx):
        nd = g.nodes[idx]
        return ng._mk(("mul", m[nd[1]], _scaled(nd[2], -1, 0)))

    for idx, nd in enumerate(g.nodes):
        if idx not in live:
            continue
        op = nd[0]
        if op == "in":
            m[idx] = ng._mk(nd)
        elif op == "mul":
            m[idx] = ng._mk(("mul", m[nd[1]], nd[2]))
        elif op == "add":
            a, b = nd[1], nd[2]
            if g.nodes[b][0] == "mul":
                m[idx] = ng._mk(("sub", m[a], mulneg(b)))
            elif g.nodes[a][0] == "mul":
                m[idx] = ng._mk(("sub", m[b], mulneg(a)))
            else:
                m[idx] = ng._mk(("add", m[a], m[b]))
        elif op == "sub":
            m[idx] = ng._mk(("sub", m[nd[1]], m[nd[2]]))
        elif op == "fma2":
            _, a, k, b, sa, sb = nd
            if g.nodes[a][0] == "mul":          # (2^k * c) * x, exact
                an = g.nodes[a]
                if sb > 0:                       # sa*2^k*c*x + b  ->  b - (-sa*2^k*c)*x
                    pa = ng._mk(("mul", m[an[1]], _scaled(an[2], -sa, k)))
                    m[idx] = ng._mk(("sub", m[b], pa))
                else:                            # 2^k*c*x - b      (sa is +1 here by construction)
                    pa = ng._mk(("mul", m[an[1]], _scaled(an[2], sa, k)))
                    m[idx] = ng._mk(("sub", pa, m[b]))
            else:
                m[idx] = ng._mk(("fma2", m[a], k, m[b], sa, sb))
    nouts = [Val(m[o.node], o.sign, o.exp) for o in outs]
    # post-condition: no add / fma2 has a multiply operand
    for nd in ng.nodes:
        if nd[0] == "add":
            assert ng.nodes[nd[1]][0] != "mul" and ng.nodes[nd[2]][0] != "mul"
        if nd[0] == "fma2":
            assert ng.nodes[nd[1]][0] != "mul"
    return ng, nouts


# ----------------------------------------------------------------------------
# numpy evaluation (the bit-exactness check against scipy)
# ----------------------------------------------------------------------------

def live_nodes(g: Graph, outs):
    live = set()
    stack = [o.node for o in outs]
    while stack:
        n = stack.pop()
        if n in live:
            continue
        live.add(n)
        nd = g.nodes[n]
        if nd[0] in ("add", "sub"):
            stack += [nd[1], nd[2]]
        elif nd[0] == "mul":
            stack.append(nd[1])
        elif nd[0] == "fma2":
            stack += [nd[1], nd[3]]
    return live


def evaluate(g: Graph, outs, x: np.ndarray, dtype, in_exp: int = 0, contract: bool = False):
    """Evaluate the traced program on x[..., n] (true input values; the lazy
    input exponent in_exp is removed first, exactly) with one rounding per node
    in the working dtype.  contract=True additionally emulates the fused
    multiply-adds gen_cuda emits in its non-exact mode (float32 only; the
    product is formed in float64, which is exact for float32 operands)."""
    dt = np.dtype(dtype).type
    live = live_nodes(g, outs)
    uses = {}
    if contract:
        for n in live:
            nd = g.nodes[n]
            if nd[0] in ("add", "sub"):
                for a in (nd[1], nd[2]):
                    uses[a] = uses.get(a, 0) + 1
            elif nd[0] == "mul":
                uses[nd[1]] = uses.get(nd[1], 0) + 1
            elif nd[0] == "fma2":
                uses[nd[1]] = uses.get(nd[1], 0) + 1
                uses[nd[3]] = uses.get(nd[3], 0) + 1
        for o in outs:
            uses[o.node] = uses.get(o.node, 0) + 1
    vals = {}
    for n, nd in enumerate(g.nodes):
        if n not in live:
            continue
        op = nd[0]
        if op == "in":
            vals[n] = (x[..., nd[1]].astype(dtype) * dt(2.0 ** (-in_exp))).astype(dtype)
        elif op in ("add", "sub"):
            a, b = nd[1], nd[2]
            if contract and dtype == np.float32:
                fused = None
                for first, second in ((a, b), (b, a)):
                    fn = g.nodes[first]
                    if fn[0] == "mul" and uses.get(first, 0) == 1:
                        fused = (first, second)
                        break
                if fused is not None:
                    first, second = fused
                    fn = g.nodes[first]
                    prod = vals[fn[1]].astype(np.float64) * np.float64(const_value(fn[2], dtype))
                    oth = vals[second].astype(np.float64)
                    if op == "add":
                        r = prod + oth
                    else:
                        r = (prod - oth) if first == a else (oth - prod)
                    vals[n] = r.astype(np.float32)
                    continue
            vals[n] = (vals[a] + vals[b]) if op == "add" else (vals[a] - vals[b])
        elif op == "mul":
            vals[n] = vals[nd[1]] * const_value(nd[2], dtype)
        elif op == "fma2":
            _, a, k, b, sa, sb = nd
            vals[n] = dt(sa * 2.0 ** k) * vals[a] + dt(sb) * vals[b]
        assert vals[n].dtype == np.dtype(dtype), (nd, vals[n].dtype)
    res = [dt(o.sign * 2.0 ** o.exp) * vals[o.node] for o in outs]
    return np.stack(res, axis=-1)


def op_counts(g: Graph, outs):
    live = live_nodes(g, outs)
    cnt = {"add": 0, "sub": 0, "mul": 0, "fma2": 0}
    for n in live:
        op = g.nodes[n][0]
        if op in cnt:
            cnt[op] += 1
    return cnt
