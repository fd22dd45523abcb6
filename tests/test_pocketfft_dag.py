"""The traced pocketfft program (vcf_b200/codegen/pocketfft_dag.py) must equal
scipy.fftpack.dct/idct(norm='ortho') BIT FOR BIT in float32 and float64.  This is
what the CUDA codelets are generated from; a scipy upgrade that changes
pocketfft's operation order or twiddles makes this test fail."""
import numpy as np
import pytest
import scipy.fftpack as fp

from vcf_b200.codegen import pocketfft_dag as D


def _bits(a):
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


@pytest.mark.parametrize("n", [4, 8, 16, 32])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_dag_matches_scipy_bitwise(n, inverse, dt):
    rng = np.random.default_rng(n * 7 + inverse)
    g, outs = D.trace_dct(n, inverse)
    fn = fp.idct if inverse else fp.dct
    m = 50000
    wide = (rng.standard_normal((m, n)) * np.exp(rng.uniform(-4, 4, (m, n)))).astype(dt)
    quarter = (rng.integers(-512, 512, (m, n)) / 4).astype(dt)     # YCoCg samples
    ints = (rng.integers(-128, 128, (m, n)) * rng.integers(1, 65, (m, 1))).astype(dt)  # q*k
    for x in (wide, quarter, ints):
        ref = fn(x, norm="ortho", axis=-1)
        got = D.evaluate(g, outs, x, dt)
        assert ref.dtype == got.dtype == dt
        assert np.array_equal(_bits(ref), _bits(got))


@pytest.mark.parametrize("n", [8, 16])
def test_two_dimensional_rational_positions_are_exact_even_when_contracted(n):
    """Outputs 0 and N/2 contain no multiply-add chain, so fusing multiply-adds
    (VCFB_F_CONTRACT) cannot move the tie-prone coefficients of SURVEY.md 7.3."""
    rng = np.random.default_rng(3)
    g, outs = D.trace_dct(n, False)
    x = (rng.integers(-512, 512, (20000, n)) / 4).astype(np.float32)
    ref = fp.dct(x, norm="ortho", axis=-1)
    got = D.evaluate(g, outs, x, np.float32, contract=True)
    assert np.array_equal(_bits(ref[:, 0]), _bits(got[:, 0]))
    assert np.array_equal(_bits(ref[:, n // 2]), _bits(got[:, n // 2]))
    assert np.abs(ref - got).max() < 2e-4


def test_lazy_input_exponent_is_exact():
    rng = np.random.default_rng(4)
    x = (rng.integers(-512, 512, (1000, 8)) / 4).astype(np.float32)
    g0, o0 = D.trace_dct(8, False, in_exp=0)
    g2, o2 = D.trace_dct(8, False, in_exp=-2)
    assert np.array_equal(D.evaluate(g0, o0, x, np.float32), D.evaluate(g2, o2, x, np.float32, in_exp=-2))
    assert [o.exp - 2 for o in o0] == [o.exp for o in o2]


def test_op_counts_documented_in_design():
    g, o = D.trace_dct(8, False)
    c = D.op_counts(g, o)
    assert (c["add"] + c["sub"], c["mul"], c["fma2"]) == (38, 18, 0)
