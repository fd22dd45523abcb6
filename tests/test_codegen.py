"""The committed CUDA codelets must be exactly what the generator prints."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_codelets_in_sync_with_generator():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "vcf_b200", "codegen", "gen_cuda.py"), "--check"], cwd=ROOT)
    assert r.returncode == 0, "run `make codelets` and commit vcf_b200/csrc/dct_codelets.cuh"


def _run_emitted(src: str, x, dtype):
    """Interpret one emitted codelet (straight-line `const T t = O::op(...)` lines) with numpy
    scalars of `dtype`: every operation individually rounded, like the EXACT=true CUDA path."""
    import re
    import numpy as np
    env = {}
    out = [None] * len(x)

    def val(tok):
        tok = tok.strip()
        m = re.match(r"O::neg\((\w+)\)", tok)
        if m:
            return -env[m.group(1)]
        m = re.match(r"vcfb::konst<T>\(([^,]+)f, ([^)]+)\)", tok)
        if m:
            return dtype(float.fromhex(m.group(1))) if dtype is np.float32 else dtype(float.fromhex(m.group(2))) \
                if "p" in m.group(2) else dtype(float(m.group(2)))
        return env[tok]

    for line in src.splitlines():
        line = line.strip()
        m = re.match(r"const T (\w+) = v\[(\d+)\];", line)
        if m:
            env[m.group(1)] = dtype(x[int(m.group(2))])
            continue
        m = re.match(r"const T (\w+) = O::(\w+)\((.*)\);", line)
        if m:
            name, op, args = m.groups()
            parts, depth, cur = [], 0, ""
            for ch in args:
                if ch == "," and depth == 0:
                    parts.append(cur); cur = ""
                else:
                    depth += ch == "("; depth -= ch == ")"; cur += ch
            parts.append(cur)
            a = [val(p) for p in parts]
            if op == "add":
                env[name] = dtype(a[0] + a[1])
            elif op == "sub":
                env[name] = dtype(a[0] - a[1])
            elif op == "mul":
                env[name] = dtype(a[0] * a[1])
            elif op == "fma":      # only exact products (powers of two) appear
                env[name] = dtype(dtype(a[0] * a[1]) + a[2])
            continue
        m = re.match(r"v\[(\d+)\] = (.*);", line)
        if m:
            out[int(m.group(1))] = dtype(0) if m.group(2).startswith("T(0)") else val(m.group(2))
    return out


def test_pruned_inverse_codelets_are_bit_exact():
    """dct8_inv_low2 / _low4, dct16_inv_low2 / _low4 (inputs known to be zero dropped) against the real scipy:
    bitwise, float32 and float64 (the lazy output scale of dct8_inv is 2^-2, of dct16_inv 1)."""
    import numpy as np
    import scipy.fftpack as sf
    sys.path.insert(0, ROOT)
    from vcf_b200.codegen import gen_cuda as G
    rng = np.random.default_rng(11)
    for n, nin, scale in ((8, 2, 0.25), (8, 4, 0.25), (16, 2, 1.0), (16, 4, 1.0)):
        src = G.emit_pruned_codelet(n, True, nin)
        for dtype in (np.float64, np.float32):
            for _ in range(400):
                x = np.zeros(n, dtype=dtype)
                x[:nin] = rng.integers(-32768, 32768, size=nin)
                if rng.random() < 0.3:
                    x[rng.integers(0, nin)] = 0
                got = np.array(_run_emitted(src, x, dtype), dtype=dtype) * dtype(scale)
                want = sf.idct(x, norm="ortho")
                assert want.dtype == dtype
                assert np.array_equal(got, want), (nin, dtype, x, got, want)


def test_dag_programs_in_sync_with_generator():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "vcf_b200", "codegen", "gen_dag_programs.py"), "--check"], cwd=ROOT)
    assert r.returncode == 0, "run `python vcf_b200/codegen/gen_dag_programs.py` and commit vcf_b200/csrc/dag_programs.inc"


def test_dag_programs_are_bit_exact():
    """The interpreted programs of csrc/kernels_anyb.cu (block sizes 2, 64, 128 of the reference's -L search,
    src/2D-DCT.py:538) through the numpy twin of the device interpreter, against the real scipy: bitwise,
    float32 and float64, both directions."""
    import numpy as np
    import scipy.fftpack as sf
    sys.path.insert(0, ROOT)
    from vcf_b200.codegen import gen_dag_programs as G
    rng = np.random.default_rng(5)
    for n in G.SIZES:
        for inverse in (False, True):
            ops, consts, outs, nslots = G.program(n, inverse)
            assert nslots <= 138
            for dtype in (np.float32, np.float64):
                x = rng.integers(-2048, 2048, size=(3000, n)).astype(dtype)
                x[::7] *= dtype(0.25)
                x[1::11, n // 2:] = 0
                got = G.interpret(ops, consts, outs, nslots, x, dtype)
                want = (sf.idct if inverse else sf.dct)(x, norm="ortho", axis=-1)
                assert want.dtype == dtype
                assert np.array_equal(got, want), (n, inverse, dtype)
