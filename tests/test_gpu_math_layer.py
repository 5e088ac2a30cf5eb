"""The device math layer (csrc/bb_math.cuh: bb_sincospi, bb_atan, bb_rcp_pos - what every per-bin loop uses in place of
numpy's sin / cos / arctan and of IEEE division) against 50-digit arithmetic, through the C ABI (bb_math_device)."""
import ctypes

import mpmath as mp
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
mp.mp.dps = 50


def _run(function, x, width):
    import torch
    from bilby_b200 import _lib
    h = _lib.Handle()
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    out = torch.empty(len(x) * width, dtype=torch.float64, device="cuda")
    _lib.check(h.lib.bb_math_device(h.ptr, function, xd.data_ptr(), len(x), out.data_ptr(),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(len(x), width)


def test_sincospi_vs_50_digits():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-2, 2, 2000), rng.uniform(-1e5, 1e5, 2000), rng.uniform(-1e9, 1e9, 1000),
                        rng.uniform(-2.0 ** 49, 2.0 ** 49, 500),
                        [0.0, -0.0, 0.25, -0.25, 0.5, 1.0, 1.5, 2.0, 1e-300, 1073741824.0, 2.0 ** 40 + 0.25,
                         2.0 ** 49 + 0.5]])
    got = _run(0, x, 2)
    ref = np.array([[float(mp.sinpi(mp.mpf(float(v)))), float(mp.cospi(mp.mpf(float(v))))] for v in x])
    err = np.abs(got - ref)
    assert err.max() < 2.5e-16, (err.max(), x[np.unravel_index(err.argmax(), err.shape)[0]])
    # beyond 2^50 half turns a double has no quarter-turn information: NaN, like inf and nan themselves
    bad = _run(0, np.array([2.0 ** 50, -2.0 ** 52, 1e300, np.inf, -np.inf, np.nan]), 2)
    assert np.isnan(bad).all()


def test_atan_and_reciprocal_vs_50_digits():
    rng = np.random.default_rng(2)
    y = np.concatenate([rng.uniform(-3, 3, 3000), rng.standard_cauchy(2000) * 10, 10.0 ** rng.uniform(-300, 300, 500),
                        [0.0, 0.41421356237309503, 0.4142135623730951, 2.414213562373095, 2.4142135623730954, 1e308]])
    got = _run(1, y, 1)[:, 0]
    ref = np.array([float(mp.atan(mp.mpf(float(v)))) for v in y])
    assert np.max(np.abs(got - ref)) < 4.5e-16
    a = np.concatenate([rng.uniform(0.5, 2.0, 2000), 10.0 ** rng.uniform(-280, 280, 2000)])
    got = _run(2, a, 1)[:, 0]
    assert np.max(np.abs(got * a - 1.0)) < 4.5e-16
