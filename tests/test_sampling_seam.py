"""CPU: seam B4 (``sample_b200_nuts``) above the C ABI, with the device engine replaced by the oracle-backed stand-in:
stream derivation per chain, init (jitter + mean start point), result packaging and names, reproducibility, and chain
sharding over torch.distributed (gloo, world_size 2) giving the same posterior as a single process."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from b200_helpers import OracleEngine
from pymc_b200 import models, sampling

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(chains=3, seed=11, **kw):
    spec = models.eight_schools()
    return sampling.sample_b200_nuts(20, tune=40, chains=chains, random_seed=seed, model=OracleEngine(spec), momentum="numpy",
                                     keep_untransformed=True, **kw)


def test_whole_run_sampler_conventions_and_reproducibility():
    a, b = run(), run()
    assert np.array_equal(a.unconstrained, b.unconstrained)  # same seed => identical posterior (test_mcmc_external.py:83)
    assert set(a.posterior) == {"mu", "tau", "theta_t"} and a.posterior["theta_t"].shape == (3, 20, 8)
    assert {"diverging", "energy", "tree_depth", "n_steps", "acceptance_rate", "lp", "step_size"} <= set(a.sample_stats)
    assert a.sample_stats["diverging"].dtype == bool and a.sample_stats["n_steps"].shape == (3, 20)
    assert a.attrs["tuning_steps"] == 40 and a.attrs["inference_library"] == "pymc_b200" and "sampling_time" in a.attrs
    assert np.all(a.posterior["tau"] > 0) and np.allclose(np.log(a.posterior["tau"]), a.unconstrained[..., 1])
    c = run(seed=12)
    assert not np.array_equal(a.unconstrained, c.unconstrained)
    w = run(discard_tuned_samples=False)
    assert w.warmup_posterior["mu"].shape == (3, 40) and np.array_equal(w.unconstrained, a.unconstrained)
    only = run(var_names=["mu"])
    assert set(only.posterior) == {"mu"}


def test_chains_are_the_reference_chains_for_their_streams():
    """Chain c of the whole-run sampler == the oracle (== the reference) run on chain c's own streams and start."""
    from oracle import logp_numpy, nuts_numpy
    from pymc_b200 import rng as brng

    spec = models.eight_schools()
    chains, seed, tune, draws = 3, 5, 30, 10
    res = sampling.sample_b200_nuts(draws, tune=tune, chains=chains, random_seed=seed, model=OracleEngine(spec), momentum="numpy",
                                    keep_untransformed=True)
    step_rngs, pot_rngs, jitter_seeds = brng.chain_generators(seed, chains)
    q0 = sampling.initial_points(spec, chains, jitter_seeds)
    f = logp_numpy.make_logp(spec)
    for c in range(chains):
        m = nuts_numpy.DiagMass(np.ones(spec.n), adapt=True, initial_mean=q0.mean(axis=0), initial_weight=10)  # mcmc.py:1890-1894
        o = nuts_numpy.Oracle(f, m)
        o.rng, o.mass.rng = step_rngs[c], pot_rngs[c]
        qs, _ = o.run(q0[c], tune, draws)
        assert np.array_equal(qs[tune:], res.unconstrained[c])


def test_init_adapt_full_hands_the_reference_potential_to_the_engine():
    """init="jitter+adapt_full" (mcmc.py:1997-2005): QuadPotentialFullAdapt(n, mean of the start points, eye, 10) per chain."""
    from oracle import logp_numpy, nuts_numpy
    from pymc_b200 import rng as brng

    spec = models.eight_schools()
    chains, seed, tune, draws = 2, 8, 25, 6
    res = sampling.sample_b200_nuts(draws, tune=tune, chains=chains, random_seed=seed, model=OracleEngine(spec), momentum="numpy",
                                    keep_untransformed=True, init="jitter+adapt_full")
    step_rngs, pot_rngs, jitter_seeds = brng.chain_generators(seed, chains)
    q0 = sampling.initial_points(spec, chains, jitter_seeds)
    f = logp_numpy.make_logp(spec)
    for c in range(chains):
        o = nuts_numpy.Oracle(f, nuts_numpy.DenseAdaptMass(spec.n, q0.mean(axis=0), np.eye(spec.n), 10))
        o.rng, o.mass.rng = step_rngs[c], pot_rngs[c]
        qs, _ = o.run(q0[c], tune, draws)
        assert np.array_equal(qs[tune:], res.unconstrained[c])
    import pytest

    with pytest.raises(ValueError, match="adapt_full"):
        sampling.sample_b200_nuts(2, tune=2, chains=1, random_seed=1, model=OracleEngine(spec), momentum="numpy", init="jitter+nope")


def test_log_likelihood_group_from_the_pointwise_pass():
    """idata_kwargs={"log_likelihood": True} (sampling/jax.py:660-673): the sampler records unconstrained draws and asks the
    engine's pointwise pass for log p(y_i | draw) of every likelihood factor; here the engine is the oracle-backed stand-in."""
    from scipy import stats as st

    from oracle import ir_numpy
    from pymc_b200 import engine, ir

    m = ir.radon_ir(60, 5, 3)

    class IrStandIn(OracleEngine):
        def __init__(self, model):
            self.ir, self.spec = model, engine._ir_spec(model)
            self.n, self.f = model.n, ir_numpy.make_logp(model)
            self.calls = []

        def pointwise_loglik(self, draws, lik=0):
            self.calls.append((np.shape(draws), lik))
            c = self.ir.constrain(np.asarray(draws))
            L = self.ir.likelihoods[lik]
            county, floor = L.terms[1].factors[1][1], L.terms[2].coef
            mu = (c["mu_a"][..., None] + c["sigma_a"][..., None] * c["a"][..., county]
                  + (c["mu_b"][..., None] + c["sigma_b"][..., None] * c["b"][..., county]) * floor)
            return st.norm(mu, c["eps"][..., None]).logpdf(L.y)

    eng = IrStandIn(m)
    res = sampling.sample_b200_nuts(5, tune=8, chains=2, random_seed=4, model=eng, momentum="numpy",
                                    idata_kwargs={"log_likelihood": True}, compute_convergence_checks=False)
    name = m.likelihoods[0].name
    assert "log_likelihood" in res.groups() and set(res.log_likelihood) == {name}
    assert res.log_likelihood[name].shape == (2, 5, 60) and eng.calls == [((2, 5, m.n), 0)]
    assert np.all(np.isfinite(res.log_likelihood[name])) and np.all(res.posterior["eps"] > 0)
    plain = sampling.sample_b200_nuts(5, tune=8, chains=2, random_seed=4, model=IrStandIn(m), momentum="numpy",
                                      compute_convergence_checks=False)
    assert "log_likelihood" not in plain.groups()
    for k in res.posterior:  # asking for the group does not change the draws
        assert np.array_equal(res.posterior[k], plain.posterior[k])
    # a Deterministic of the IR appears in the posterior group, computed from the constrained draws
    m.deterministics.append(ir.Deterministic("a_scaled", m.var("a").size, [ir.Term([("sigma_a_log__", None), ("a", None)])]))
    det = sampling.sample_b200_nuts(5, tune=8, chains=2, random_seed=4, model=IrStandIn(m), momentum="numpy",
                                    compute_convergence_checks=False)
    np.testing.assert_allclose(det.posterior["a_scaled"], det.posterior["sigma_a"][..., None] * det.posterior["a"], rtol=1e-14)
    assert np.array_equal(det.posterior["a"], plain.posterior["a"])


def test_init_map_starts_every_chain_at_the_optimum_with_the_negated_hessian():
    """init="map" (mcmc.py:1981-1985): find_MAP by L-BFGS-B on the engine's logp/dlogp, cov = -Hessian (finite differences of
    the gradient, one batched call), QuadPotentialFull(cov), initial_points = [map] * chains."""
    spec = models.logistic(n_rows=200, n_features=4, seed=9)
    X, y = spec.data["X"], spec.data["y"].astype(np.float64)
    eng = OracleEngine(spec)
    q, cov = sampling.map_and_neg_hessian(eng, spec.initial_point())
    lp, g = eng.logp_dlogp(q[None])
    assert np.max(np.abs(g[0])) <= 1e-4  # a stationary point of logp
    p = 1.0 / (1.0 + np.exp(-(X @ q)))
    want = X.T @ (X * (p * (1 - p))[:, None]) + np.eye(4)  # -Hessian of logp for beta ~ Normal(0, 1), Bernoulli-logit likelihood
    np.testing.assert_allclose(cov, want, rtol=1e-6, atol=1e-8)
    res = sampling.sample_b200_nuts(6, tune=10, chains=3, random_seed=2, model=eng, momentum="numpy", keep_untransformed=True,
                                    init="map", compute_convergence_checks=False)
    assert res.unconstrained.shape == (3, 6, 4)
    assert np.allclose(eng.last_q0, np.broadcast_to(q, (3, 4)), rtol=0, atol=1e-12)  # every chain starts at the MAP
    np.testing.assert_allclose(eng.dense_cov, want, rtol=1e-6, atol=1e-8)


def test_advi_fit_and_the_advi_init_branches():
    """pymc_b200.advi.fit_meanfield finds the mean / sd of a factorised Gaussian target with the reference's optimiser and
    stopping rule; init="advi" -> QuadPotentialDiag(std**2) and start points drawn from the approximation,
    init="advi+adapt_diag" -> QuadPotentialDiagAdapt(n, mean, std**2, 50) (mcmc.py:1913-1958)."""
    from pymc_b200 import advi

    m, sd = np.array([1.0, -2.0, 0.5, 3.0]), np.array([0.5, 2.0, 1.0, 0.25])

    def f(q):
        q = np.atleast_2d(q)
        r = (q - m) / sd
        return -0.5 * np.sum(r * r, axis=1), -(q - m) / sd**2

    mu, s_, it = advi.fit_meanfield(f, np.zeros(4), seed=1)
    assert 1000 < it < 200_000  # stopped by the parameter-convergence rule, not by the iteration cap
    # the reference's stopping rule (parameters moved < 1e-2 in 100 iterations) ends the fit at optimiser-step accuracy
    assert np.all(np.abs(mu - m) < 0.5) and np.all((s_ > 0.6 * sd) & (s_ < 2.0 * sd))
    mu2, s2, it2 = advi.fit_meanfield(f, np.zeros(4), seed=1)
    assert it2 == it and np.array_equal(mu2, mu) and np.array_equal(s2, s_)  # reproducible for a seed

    spec = models.std_normal(5)
    for init, kind, weight in (("advi", "diag", None), ("advi+adapt_diag", "diag_adapt", 50.0)):
        eng = OracleEngine(spec)
        res = sampling.sample_b200_nuts(5, tune=12, chains=3, random_seed=3, model=eng, momentum="numpy", keep_untransformed=True,
                                        init=init, nuts_kwargs={"n_init": 4000}, compute_convergence_checks=False)
        assert res.unconstrained.shape == (3, 5, 5) and eng.last_call["mass"] == kind
        v0 = eng.last_call["var0"]
        assert v0.shape == (3, 5) and np.all(v0 == v0[0]) and np.all((v0[0] > 0.3) & (v0[0] < 2.5))  # std**2 of N(0, 1), roughly
        assert not np.array_equal(eng.last_q0[0], eng.last_q0[1])  # start points are draws of the approximation
        if weight is not None:
            assert eng.last_call["mass_initial_weight"] == weight and np.all(np.abs(eng.last_call["mean0"]) < 0.5)


WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch.distributed as dist
    from b200_helpers import OracleEngine
    from pymc_b200 import models, sampling
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    spec = models.eight_schools()
    res = sampling.sample_b200_nuts(12, tune=25, chains=5, random_seed=3, model=OracleEngine(spec), momentum="numpy",
                                    keep_untransformed=True, gather="all")
    np.save(sys.argv[2], res.unconstrained)
    assert res.sample_stats["n_steps"].shape == (5, 12) and res.attrs["chains_held"] == (0, 5)
    r0 = sampling.sample_b200_nuts(12, tune=25, chains=5, random_seed=3, model=OracleEngine(spec), momentum="numpy",
                                   keep_untransformed=True)  # default: all chains on rank 0, the others keep their shard
    if dist.get_rank() == 0:
        assert np.array_equal(r0.unconstrained, res.unconstrained)
    else:
        lo, hi = r0.attrs["chains_held"]
        assert (lo, hi) != (0, 5) and np.array_equal(r0.unconstrained, res.unconstrained[lo:hi])
    dist.destroy_process_group()
    print("ok")
    """
)


def test_sharded_run_world_size_2_gloo_equals_single_process(tmp_path):
    port = 29700 + (os.getpid() % 200)
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    outs = [str(tmp_path / f"r{r}.npy") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, str(script), str(r), outs[r]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
    spec = models.eight_schools()
    single = sampling.sample_b200_nuts(12, tune=25, chains=5, random_seed=3, model=OracleEngine(spec), momentum="numpy",
                                       keep_untransformed=True)
    for o in outs:  # every rank holds all chains after the gather, identical to the unsharded run
        assert np.array_equal(np.load(o), single.unconstrained)
