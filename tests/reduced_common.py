"""Shared set-up for the reduced-order likelihood tests (relative binning, ROQ): rebuilds the data of
tests/golden/{relbin,roq}_*.npz (noise from the committed seed, injection through the oracle) for the oracle
and for the CUDA path.  Test infrastructure only."""
import functools
import os

import numpy as np

from oracle import cbc_likelihood as ocl
from oracle import cbc_reduced as ocr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["H1", "L1", "V1"]


def load(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    return g, draws


def bbh_conv(p):
    return ocl.convert_to_lal_binary_black_hole_parameters(p)


def oracle_ifos(g, inj, source, wa, lambdas=False, maximum_frequency=None):
    """Oracle interferometers carrying seeded Gaussian noise + the injection (same recipe as
    oracle/tools/make_golden_reduced.py::make_ifos + inject_signal_from_waveform_polarizations)."""
    fs, duration, start = float(g["sampling_frequency"]), float(g["duration"]), float(g["start_time"])
    rng = np.random.default_rng(int(g["noise_seed"]))
    ifos = [ocl.OracleInterferometer(n, fs, duration, start, maximum_frequency=maximum_frequency) for n in NAMES]
    for o in ifos:
        o.set_gaussian_noise(rng)
    conv = bbh_conv(inj)
    args = [conv[k] for k in ocl.SOURCE_ARGS]
    if lambdas:
        args += [inj["lambda_1"], inj["lambda_2"]]
    pols = source(ifos[0].frequency_array, *args, **wa)
    for o in ifos:
        o.frequency_domain_strain = o.frequency_domain_strain + o.get_detector_response(pols, conv)
    return ifos


def injection_of(g):
    return {k[4:]: float(g[k]) for k in g.files if k.startswith("inj_")}


@functools.lru_cache(maxsize=None)
def distance_phase_table():
    """The 400 x 800 distance+phase lookup table for a PowerLaw(alpha=2) prior on [d, 50 d]: identical for
    (100, 5000) and (10, 500) Mpc because only d_ref/d and p(d) dd enter (base.py:994-1018)."""
    cache = os.path.join(os.environ.get("TMPDIR", "/tmp"), "bb200_oracle_dp_table.npy")
    if os.path.exists(cache):
        return np.load(cache)
    stub = ocl.OracleLikelihood.__new__(ocl.OracleLikelihood)
    stub.phase_marginalization = True
    prior = ocl.OraclePowerLaw(2, 100.0, 5000.0)
    stub.distance_prior = prior
    stub._distance_array = np.linspace(prior.minimum, prior.maximum, int(1e4))
    stub.distance_prior_array = np.array([prior.prob(d) for d in stub._distance_array])
    stub._ref_dist = prior.rescale(0.5)
    table = stub.create_lookup_table(processes=min(8, os.cpu_count() or 1))
    try:
        np.save(cache, table)
    except OSError:
        pass
    return table


def relbin_oracle(g, bns, **kw):
    inj = injection_of(g)
    approx = str(g["approximant"])
    wa = dict(waveform_approximant=approx, reference_frequency=50.0, minimum_frequency=20.0)
    grid_source = ocl.lal_binary_neutron_star if bns else ocl.lal_binary_black_hole
    ifos = oracle_ifos(g, inj, grid_source, wa, lambdas=bns)
    draws = {k[6:]: g[k] for k in g.files if k.startswith("param_")}
    fid = {k: float(v[0]) for k, v in draws.items() if k != "time_jitter"}
    model = ocr.lal_binary_neutron_star_relative_binning if bns else ocr.lal_binary_black_hole_relative_binning
    return ocr.OracleRelativeBinning(ifos, fid, source_model=model, waveform_arguments=wa, **kw), ifos


def roq_oracle(g, **kw):
    inj = dict(ocl.INJECTION)
    fmax = float(g["maximum_frequency"])
    wa_full = dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0, minimum_frequency=20.0)
    ifos = oracle_ifos(g, inj, ocl.lal_binary_black_hole, wa_full, maximum_frequency=fmax)
    t_inj = inj["geocent_time"]
    return ocr.OracleROQ(
        ifos, g["linear_matrix"].astype(complex), g["quadratic_matrix"].astype(complex),
        g["frequency_nodes_linear"], g["frequency_nodes_quadratic"],
        time_prior=ocl.OracleUniform(t_inj - 0.1, t_inj + 0.1),
        waveform_arguments=dict(waveform_approximant="IMRPhenomD", reference_frequency=20.0),
        optimal_snrs=list(g["optimal_snrs"]), **kw), ifos


def multiband_oracle(g, bns, **kw):
    """Oracle multi-banded likelihood on the data of tests/golden/multiband_*.npz (oracle/tools/make_golden_multiband.py)."""
    from oracle import cbc_multiband as ocm
    inj = injection_of(g)
    approx = str(g["approximant"])
    wa_full = dict(waveform_approximant=approx, reference_frequency=50.0, minimum_frequency=20.0)
    grid_source = ocl.lal_binary_neutron_star if bns else ocl.lal_binary_black_hole
    ifos = oracle_ifos(g, inj, grid_source, wa_full, lambdas=bns)
    model = ocm.binary_neutron_star_frequency_sequence if bns else ocm.binary_black_hole_frequency_sequence
    return ocm.OracleMultiband(ifos, float(g["reference_chirp_mass"]), source_model=model,
                               waveform_arguments=dict(waveform_approximant=approx, reference_frequency=50.0),
                               geocent_time_prior=tuple(g["geocent_time_prior"]), **kw), ifos
