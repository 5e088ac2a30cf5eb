"""Test infrastructure for the BASELINE-size parity tests (tests/test_gpu_baseline_size.py): the oracle built on the
SAME data as a product likelihood, and a fork pool that evaluates it over the host cores (the reference's own
fan-out, bilby/core/sampler/base_sampler.py:772-800)."""
import multiprocessing
import os

import numpy as np

from oracle import cbc_likelihood as ocl
from oracle import cbc_reduced as ocr

_POOL = {}


def _eval(p):
    return _POOL["like"].log_likelihood_ratio(p)


def _eval_snrs(p):
    return _POOL["like"].log_likelihood_ratio(p, return_snrs=True)


def oracle_map(olike, draws, n, snrs=False, processes=None):
    """lnL (or the per-detector (<d|h>, <h|h>) list) of the oracle for the first n rows of a dict of arrays."""
    plist = [{k: float(np.asarray(v)[i]) for k, v in draws.items()} for i in range(n)]
    processes = processes or min(os.cpu_count() or 1, 32)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _POOL["like"] = olike
    fn = _eval_snrs if snrs else _eval
    if processes == 1:
        out = [fn(p) for p in plist]
    else:
        with multiprocessing.get_context("fork").Pool(processes) as pool:
            out = pool.map(fn, plist, chunksize=max(1, n // (processes * 4)))
    _POOL.pop("like")
    return out if snrs else np.array(out)


def oracle_ifos_like(ifos, calibration_points=0):
    """Oracle interferometers carrying exactly the product interferometers' data, band and PSD source."""
    out = []
    for ifo in ifos:
        o = ocl.OracleInterferometer(ifo.name, ifo.sampling_frequency, ifo.duration, ifo.start_time,
                                     minimum_frequency=ifo.minimum_frequency, maximum_frequency=ifo.maximum_frequency)
        assert np.array_equal(o.power_spectral_density_array, ifo.power_spectral_density_array)
        assert np.array_equal(o.frequency_mask, ifo.frequency_mask)
        o.frequency_domain_strain = np.array(ifo.frequency_domain_strain)
        if calibration_points:
            o.calibration = ocl.OracleCubicSpline(f"recalib_{ifo.name}_", ifo.minimum_frequency, ifo.maximum_frequency,
                                                  calibration_points)
        out.append(o)
    return out


def scale_of(lnl, hh_total):
    """lnL passes through zero, so the 1e-8 gate is relative to max(|lnL|, 1/2 sum rho_opt^2)."""
    return np.maximum(np.abs(lnl), 0.5 * np.asarray(hh_total))


def total_optimal_snr_squared(like, draws, cal=None):
    import torch
    rows = torch.from_numpy(np.ascontiguousarray(like.pack(draws))).cuda()
    cal_dev = None if cal is None else torch.from_numpy(np.ascontiguousarray(cal)).cuda()
    return like.inner_products_batch(rows, cal_dev)[..., 2].sum(dim=1).cpu().numpy()
