"""The FP64 tensor-core contraction (csrc/bb_gemm.cuh, mma.sync.m8n8k4.f64) through the C ABI vs numpy: complex and
real, ragged tile edges, K-segments (detectors along the contraction axis), batches, accumulation, alpha.
It stands where the reference uses `@` / einsum: roq.py:604-651, base.py:305-346, roq.py:849-918."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(cplx, m, n, k, n_seg=1, n_batch=1, accumulate=False, alpha=1.0, pad=0, seed=0):
    import torch
    from bilby_b200 import _lib
    h = _lib.Handle()
    rng = np.random.default_rng(seed)
    lda, ldb, ldc = n_seg * k + pad, n_seg * k + 2 * pad, n + pad

    def rand(*shape):
        x = rng.standard_normal(shape)
        return x + 1j * rng.standard_normal(shape) if cplx else x
    a, b = rand(n_batch, m, lda), rand(n_batch, n, ldb)
    c0 = rand(n_batch, m, ldc)
    ref = np.array(c0) if accumulate else np.zeros_like(c0)
    for s in range(n_seg):
        ref[:, :, :n] += alpha * np.einsum("bmk,bnk->bmn", a[:, :, s * k:(s + 1) * k], b[:, :, s * k:(s + 1) * k])
    if not accumulate:
        ref[:, :, n:] = c0[:, :, n:]                      # padding columns are never written
    dt = torch.complex128 if cplx else torch.float64
    ad, bd, cd = (torch.from_numpy(np.ascontiguousarray(x)).to(dt).cuda() for x in (a, b, c0))
    _lib.check(h.lib.bb_contract_device(h.ptr, int(cplx), m, n, k, n_seg, k, k, n_batch, m * lda, n * ldb, m * ldc,
                                        float(alpha), ad.data_ptr(), lda, bd.data_ptr(), ldb, int(accumulate),
                                        cd.data_ptr(), ldc, None))
    torch.cuda.synchronize()
    got = cd.cpu().numpy()
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) < 1e-13 * scale * max(1, (n_seg * k) ** 0.5), (cplx, m, n, k)


@pytest.mark.parametrize("cplx", [True, False])
def test_contract_shapes(cplx):
    _run(cplx, 64, 128, 16)                    # exactly one complex tile, one slab
    _run(cplx, 1, 1, 4)                        # smallest: one k-step
    _run(cplx, 67, 131, 37, pad=1)             # ragged in every dimension, odd K, odd leading dimensions
    _run(cplx, 300, 523, 256, n_batch=3)       # many tiles per CTA stream, batches (detectors)
    _run(cplx, 129, 1000, 72, n_seg=3, alpha=0.5, accumulate=True, pad=4)   # K-segments + accumulate (calibration curves)
    _run(cplx, 2000, 96, 1000)                 # long K: many refills of the 3-stage ring


def test_contract_many_tiles_per_cta():
    _run(True, 64 * 40, 128 * 12, 64)          # 480 tiles over <= 148 CTAs: cursor wraps tiles mid-ring
    _run(False, 128 * 30, 128 * 10, 48)
