"""The kernel's branch-free f64 division / square root (suite_kernel.cuh div_fast, div_finish, sqrt_fast)
against the device's own `/` and sqrt() operators, bit for bit, wherever their acceptance test passes:
random magnitudes over the whole exponent range, prices / volumes / sums of the size the suite sees,
and the special values (zeros, subnormals, infinities, NaN) that must be rejected or still agree."""
import ctypes as C

import numpy as np
import pytest

import polars_quant_b200 as pq
from polars_quant_b200 import _native as N

pytestmark = pytest.mark.gpu


def _run(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    res = (C.c_uint64 * 4)()
    N.check(N.lib().pqb_selftest_divsqrt(pq.get_engine(0)._h, a.ctypes.data_as(C.c_void_p),
                                         b.ctypes.data_as(C.c_void_p), a.size, res))
    return [int(x) for x in res]


def test_fast_division_and_sqrt_equal_the_operators_on_typical_magnitudes():
    rng = np.random.default_rng(11)
    n = 1 << 21
    a = np.concatenate([rng.lognormal(4.0, 2.0, n // 2), rng.normal(0.0, 1e6, n // 2)])
    b = np.concatenate([rng.lognormal(0.0, 3.0, n // 2), rng.integers(1, 300, n // 2).astype(np.float64)])
    bad_div, bad_sqrt, n_div, n_sqrt = _run(a, b)
    assert bad_div == 0 and bad_sqrt == 0
    assert n_div > 0.99 * n                      # the fast path is THE path on such data
    assert n_sqrt > 0.7 * n                      # (the negative half of the normal draws is rejected)


def test_fast_division_and_sqrt_equal_the_operators_on_random_bit_patterns():
    rng = np.random.default_rng(12)
    n = 1 << 21
    a = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64)
    b = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64)
    bad_div, bad_sqrt, n_div, n_sqrt = _run(a, b)
    assert bad_div == 0 and bad_sqrt == 0
    assert n_div > 0.2 * n and n_sqrt > 0.4 * n


def test_special_values_are_rejected_or_agree():
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072014e-308,
                   1.7976931348623157e308, 1e-300, 1e300, 3.0, 1.0 / 3.0, 100.0, 2.0 ** -969, 2.0 ** -970,
                   2.0 ** 1017, 2.0 ** 1016, 2.0 ** -1022, 2.0 ** 1023])
    a, b = np.meshgrid(sp, sp)
    bad_div, bad_sqrt, _, _ = _run(a.ravel(), b.ravel())
    assert bad_div == 0 and bad_sqrt == 0
