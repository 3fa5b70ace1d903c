"""The deterministic libm behind every bit-reproducible result (oracle/detmath.h and
libprosic_b200/csrc/detmath.cuh: msun exp / expm1 / log / log1p as fixed sequences of correctly
rounded binary64 operations).

Why it exists: the reference keeps an observation iff prob_ref != prob_alt EXACTLY
(src/variants/evidence/realignment/mod.rs:303), so the last bit of exp / ln_1p decides a discrete
output.  The oracle and the library's literal log-space kernel (VLR_HMM_LITERAL) must therefore
share one libm; the platform's (glibc) differs between versions and builds."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from libprosic_b200 import _ffi


def _inputs(fn, n, seed):
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    if fn == 0:      # exp: arguments of ln_add_exp / ln_sum_exp (<= 0), plus the general range
        x = np.concatenate([-745.0 * u[:n // 3], -50.0 * u[n // 3:2 * n // 3] ** 2, 700.0 * (u[2 * n // 3:] - 0.5)])
    elif fn == 1:    # expm1: ln_one_minus_exp uses it on [-0.693, 0]
        x = np.concatenate([-0.6932 * u[:n // 2], 30.0 * (u[n // 2:] - 0.5)])
    elif fn == 2:    # log
        x = np.concatenate([u[:n // 3], np.exp(700.0 * (u[n // 3:2 * n // 3] - 0.5)), 1.0 + (u[2 * n // 3:] - 0.5) * 1e-3])
    else:            # log1p: sums of exp(p - pmax) and -exp(p)
        x = np.concatenate([2.0 * u[:n // 3], -0.9999 * u[n // 3:2 * n // 3], np.exp(-40.0 * u[2 * n // 3:])])
    special = {0: [-np.inf, 0.0, -0.0, 1e-300, -745.2, 709.7, 710.0], 1: [-np.inf, 0.0, -1e-20, -0.693, -0.35],
               2: [0.0, 1.0, np.inf, 4.9e-324, 2.2e-308], 3: [-1.0, 0.0, 1e-20, -1e-20, 1.0, np.inf]}[fn]
    return np.concatenate([x, np.array(special)])


def _ulps(a, b):
    fin = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~fin & ~np.isnan(a)], b[~fin & ~np.isnan(b)])
    ia, ib = a[fin].view(np.int64), b[fin].view(np.int64)
    return np.abs(ia - ib).max() if fin.any() else 0


@pytest.mark.parametrize("fn,npf", [(0, np.exp), (1, np.expm1), (2, np.log), (3, np.log1p)])
def test_oracle_libm_is_within_one_ulp_of_the_platform_libm(fn, npf):
    x = _inputs(fn, 300000, 10 + fn)
    with np.errstate(all="ignore"):
        want = npf(x)
    assert _ulps(O.ref_math(fn, x), want) <= 1


@pytest.mark.parametrize("fn", [0, 1, 2, 3])
def test_library_host_libm_is_bit_identical_to_the_oracle_libm(fn):
    """Two separately written copies of the same published algorithm (the product never includes
    oracle/): they must agree bit for bit."""
    lib = _ffi.load()
    x = _inputs(fn, 300000, 20 + fn)
    got = np.empty_like(x)
    assert lib.vlr_ref_math(None, fn, x.size, x.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p)) == 0
    want = O.ref_math(fn, x)
    assert np.array_equal(got.view(np.int64), want.view(np.int64))


def test_constants_of_the_literal_kernel():
    # kLnConfusion (pairhmm_kernel.cuh) and Prob(0.5) (myers.cu) as the oracle's libm gives them
    assert O.ref_math(2, [0.3333])[0] == -1.098712293668443
    assert O.ref_math(2, [0.5])[0] == -0.6931471805599453


@pytest.mark.gpu
@pytest.mark.parametrize("fn", [0, 1, 2, 3])
def test_device_libm_is_bit_identical_to_the_host_libm(ctx, fn):
    x = _inputs(fn, 1000000, 30 + fn)
    dev = ctx.ref_math(fn, x, on_device=True)
    host = ctx.ref_math(fn, x, on_device=False)
    assert np.array_equal(dev.view(np.int64), host.view(np.int64))
    assert np.array_equal(host.view(np.int64), O.ref_math(fn, x).view(np.int64))
