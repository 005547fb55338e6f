"""The table-based logarithm of the save-moments kernel (csrc/rowops.cu: fast_log_normal / fast_log_normal_rcp), restated in
numpy with the same constants and operation order: absolute error of the algorithm itself, without a GPU.  The GPU
tests (tests/test_gpu_ops.py::test_save_moments_*) check the kernel's sums against numpy's log."""

import numpy as np
import pytest


def _split(x):
    bits = x.view(np.int64)
    hi = (bits >> 32).astype(np.int64)
    m = ((bits & 0x000FFFFFFFFFFFFF) | 0x3FF0000000000000).view(np.float64)
    return hi, m


def _poly(r):
    q = r * (-1.0 / 8.0) + 1.0 / 7.0
    for c in (-1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5):
        q = r * q + c
    return (r * r) * q + r


def _fma_m1(m, inv):  # fma(m, inv, -1) is exact to one rounding: emulate with extended precision
    return (m.astype(np.longdouble) * inv - 1).astype(np.float64)


def fast_log_normal(x):
    hi, m = _split(x)
    inv = 1.0 / (1.0 + np.arange(128) / 128.0)
    tab_l = -np.log(inv)
    idx = (hi >> 13) & 127
    r = _fma_m1(m, inv[idx])
    return ((hi >> 20) - 1023).astype(np.float64) * 0.693147180559945309417232 + (tab_l[idx] + _poly(r)), r


def fast_log_normal_rcp(x, seed_error):
    hi, m = _split(x)
    r0 = (1.0 / m) * (1.0 + seed_error)  # stand-in for MUFU.RCP64H: any value near 1/m must give a correct result
    rh = (r0.view(np.int64) >> 32) & 0xFFFFE000
    rh = np.minimum(np.maximum(rh, 0x3FE00000), 0x3FF00000)
    qm = (rh << 32).astype(np.int64).view(np.float64)
    j = (rh - 0x3FE00000) >> 13
    ltab = np.array([-np.log(0.5 * (1.0 + k / 128.0)) for k in range(128)] + [0.0])
    r = _fma_m1(m, qm)
    return ((hi >> 20) - 1023).astype(np.float64) * 0.693147180559945309417232 + (ltab[j] + _poly(r)), r


def _samples():
    rng = np.random.default_rng(1)
    return np.concatenate([10.0 ** rng.uniform(-300, 300, 200000), rng.uniform(0.4, 2.5, 200000),
                           [1.0, 2.0, 0.5, 1 - 2.0**-53, 1 + 2.0**-52, np.nextafter(2.0, 0.0), 2.2250738585072014e-308]])


def _check(got, x):
    ref = np.log(x.astype(np.longdouble)).astype(np.float64)
    err = np.abs(got - ref)
    big = np.abs(ref) > 0.5
    assert np.max(err[big] / np.abs(ref[big])) < 1.5e-15
    assert np.max(err[~big]) < 3e-16  # near x = 1 the error is absolute: what a sum of -f log f needs


def test_table_log_matches_log():
    x = _samples()
    got, r = fast_log_normal(x)
    assert r.min() >= -1e-16 and r.max() <= 2.0**-7
    _check(got, x)


@pytest.mark.parametrize("seed_error", [0.0, 2.0**-20, -(2.0**-20), 2.0**-12, -(2.0**-12)])
def test_reciprocal_seed_log_is_independent_of_the_seed(seed_error):
    x = _samples()
    got, r = fast_log_normal_rcp(x, seed_error)
    assert np.abs(r).max() <= 2.0**-7 + 2.0**-11
    _check(got, x)
