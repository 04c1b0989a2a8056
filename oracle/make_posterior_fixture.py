"""TEST INFRASTRUCTURE ONLY -- posterior-moment fixture for the bench path of BASELINE config #2.

    python -m oracle.make_posterior_fixture            # writes tests/golden/radon_posterior_oracle.npz

Runs 32 independent chains x (500 tune + 500 draws) of the Radon model through oracle/nuts_numpy.py (the NumPy
restatement that tests/test_oracle_vs_reference.py proves bit-identical to the reference's own NUTS files) with the
reference's default initialisation (jitter+adapt_diag: U(-1,1) jitter, QuadPotentialDiagAdapt(n, mean, ones, 10),
pymc/sampling/mcmc.py:1890-1894) and stores, per unconstrained parameter, the posterior mean and sd with their Monte
Carlo standard errors (rank-normalised bulk ESS).  The GPU test
(tests/test_gpu_fullsize.py::test_radon_bench_path_posterior_matches_oracle_run) checks that a device-Philox,
cold-adaptation GPU run agrees within 4 MCSE -- the statistical check the reference's own sampler tests use
(tests/sampler_fixtures.py:158-171, tests/step_methods/hmc/test_nuts.py).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHAINS, TUNE, DRAWS, SEED = 32, 500, 500, 20260923


def _chain(job):
    c, seed = job
    from oracle import logp_numpy, nuts_numpy
    from pymc_b200 import models

    spec = models.radon()
    f = logp_numpy.make_logp(spec)
    rng = np.random.default_rng([seed, c])
    q0 = spec.initial_point() + rng.uniform(-1, 1, spec.n)
    mass = nuts_numpy.DiagMass(np.ones(spec.n), adapt=True, initial_mean=q0.copy(), initial_weight=10)
    o = nuts_numpy.Oracle(f, mass)
    o.setup_chain(np.random.default_rng([seed, c, 1]))
    qs, st = o.run(q0, TUNE, DRAWS)
    return qs[TUNE:], int(st["diverging"][TUNE:].sum())


def main():
    from pymc_b200 import diagnostics

    with mp.get_context("spawn").Pool(min(CHAINS, os.cpu_count() or 1)) as pool:
        out = pool.map(_chain, [(c, SEED) for c in range(CHAINS)])
    x = np.stack([o[0] for o in out])  # [chains, draws, n]
    ess = diagnostics.ess_bulk(x)
    rhat = diagnostics.rhat(x) if hasattr(diagnostics, "rhat") else np.zeros(x.shape[-1])
    mean, sd = x.mean((0, 1)), x.std((0, 1))
    # Monte Carlo error of the VARIANCE estimate: (x - mean)^2 is itself a chain, with its own (smaller) effective sample
    # size -- the funnel directions (log sigma_b) have heavy-tailed squared deviations
    dev2 = (x - mean) ** 2
    ess_var = diagnostics.ess_bulk(dev2)
    var_mcse = dev2.std((0, 1)) / np.sqrt(ess_var)
    path = os.path.join(ROOT, "tests", "golden", "radon_posterior_oracle.npz")
    np.savez_compressed(path, mean=mean, sd=sd, ess=ess, mcse_mean=sd / np.sqrt(ess), mcse_sd=sd / np.sqrt(2 * ess),
                        var=sd**2, ess_var=ess_var, var_mcse=var_mcse, chains=CHAINS, tune=TUNE, draws=DRAWS, seed=SEED, divergences=sum(o[1] for o in out), rhat=rhat)
    print(f"radon posterior fixture: {CHAINS} x ({TUNE}+{DRAWS}), min ESS {ess.min():.0f}, max rhat {np.max(rhat):.4f}, "
          f"divergences {sum(o[1] for o in out)} -> {path}")


if __name__ == "__main__":
    main()
