"""Launches bb_contract_device on a few shapes (run under `ncu --metrics gpu__time_duration.sum -k regex:bb_gemm_nt`):
the per-launch durations give the GEMM's TFLOP/s as a function of K and of the tile count."""
import sys
import torch
sys.path.insert(0, ".")
from bilby_b200 import _lib

h = _lib.Handle()
SHAPES = [(True, 9472, 3072, 256), (True, 9472, 3072, 4096), (True, 9472, 1024, 4096), (True, 9472 * 2, 3072, 64),
          (False, 9472, 1024, 4096), (False, 18944, 2048, 1024), (True, 64 * 148, 128, 16384)]
for cplx, m, n, k in SHAPES:
    dt = torch.complex128 if cplx else torch.float64
    a = torch.randn(m, k, dtype=dt, device="cuda")
    b = torch.randn(n, k, dtype=dt, device="cuda")
    c = torch.empty(m, n, dtype=dt, device="cuda")
    for _ in range(2):
        _lib.check(h.lib.bb_contract_device(h.ptr, int(cplx), m, n, k, 1, 0, 0, 1, 0, 0, 0, 1.0, a.data_ptr(), k, b.data_ptr(),
                                            k, 0, c.data_ptr(), n, None))
    torch.cuda.synchronize()
    print("shape", cplx, m, n, k, "flop", (8 if cplx else 2) * m * n * k)
