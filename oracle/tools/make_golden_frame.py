"""BUILD TOOL - golden vectors for the detector-based sky frame / detector time reference
(GravitationalWaveTransient.get_sky_frame_parameters, base.py:1091-1137) from the UNMODIFIED reference.
    PYTHONPATH=oracle/standins:/root/reference python oracle/tools/make_golden_frame.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

import bilby  # noqa: E402
from oracle import cbc_likelihood as ocl  # noqa: E402
from make_golden import build  # noqa: E402

bilby.core.utils.logger.setLevel("ERROR")


def main():
    inj, start_time, wfg, ifos = build(4.0, 2048.0, ["H1", "L1", "V1"], noise_seed=88170235)
    like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, reference_frame="H1L1", time_reference="H1")
    n = 32
    rng = np.random.default_rng(11)
    draws = ocl.draw_bbh_prior(n, np.random.default_rng(20261017))
    draws.pop("ra"), draws.pop("dec")
    t = draws.pop("geocent_time")
    draws["zenith"] = np.arccos(rng.uniform(-1, 1, n))
    draws["azimuth"] = rng.uniform(0, 2 * np.pi, n)
    draws["H1_time"] = t
    res = dict(start_time=start_time, detectors=np.array(["H1", "L1", "V1"]))
    sky = np.zeros((n, 3))
    lnl = np.zeros(n)
    for i in range(n):
        p = {k: float(v[i]) for k, v in draws.items()}
        s = like.get_sky_frame_parameters(p)
        sky[i] = s["ra"], s["dec"], s["geocent_time"]
        lnl[i] = like.log_likelihood_ratio(p)
    for k in draws:
        res["param_" + k] = draws[k]
    res["sky"] = sky
    res["lnl_none"] = lnl
    # time reference only (sky frame)
    like = bilby.gw.likelihood.GravitationalWaveTransient(ifos, wfg, time_reference="L1")
    d2 = ocl.draw_bbh_prior(n, np.random.default_rng(20261017))
    d2["L1_time"] = d2.pop("geocent_time")
    res["lnl_L1_time"] = np.array([like.log_likelihood_ratio({k: float(v[i]) for k, v in d2.items()}) for i in range(n)])
    for ifo in ifos:
        res[f"strain_{ifo.name}"] = ifo.frequency_domain_strain
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sky_frame_4s_H1L1V1.npz"), **res)
    print(sky[:3], lnl[:3], res["lnl_L1_time"][:3])


if __name__ == "__main__":
    main()
