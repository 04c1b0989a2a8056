"""-m gpu: chain-state export / import (b200_chain_state: BaseHMC.sampling_state of the reference, hmc/base_hmc.py:61-71,
quadpotential.py:189-208, step_sizes.py:26-38), partial schedules, pooled warm-up, the fused constrain step and the
sampling API's init / step variants."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(spec, C, seed):
    from pymc_b200 import rng as brng

    sr, pr, js = brng.chain_generators(seed, C)
    q0 = np.stack([spec.initial_point() + np.random.default_rng(s).uniform(-1, 1, spec.n) for s in js])
    return q0, sr, pr


@pytest.mark.parametrize("kind", ["diag_adapt", "diag_adapt_grad", "hmc"])
def test_a_run_split_into_slices_is_bit_identical_to_the_uninterrupted_run(kind):
    from pymc_b200 import engine, models
    from pymc_b200 import rng as brng

    spec = models.radon()
    cm = engine.CompiledModel(spec)
    C, tune, draws = 9, 45, 15
    q0, sr, pr = _setup(spec, C, 5)
    kw = dict(mean0=q0, philox_seed=99, adaptation_window=20, discard_window=5)
    if kind == "hmc":
        kw.update(sampler="hmc", target_accept=0.65, mass="diag_adapt")
    else:
        kw.update(mass=kind)
    full = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=tune, draws=draws, **kw)
    st = brng.pack_pcg64(sr)
    state, begin, parts = None, 0, []
    for cut in (13, 40, 52, tune + draws):  # crosses estimator windows and the tuning / sampling switch
        nxt = engine.ChainState(C, spec.n)
        parts.append(cm.nuts_run(q0, st, tune=tune, draws=draws, iter_begin=begin, iter_count=cut - begin, resume=state, save=nxt, **kw))
        state, begin = nxt, cut
    assert state.iter_count == tune + draws
    assert np.array_equal(np.concatenate([p.draws for p in parts], axis=1), full.draws)
    for k in full.stats:
        assert np.array_equal(np.concatenate([p.stats[k] for p in parts], axis=1), full.stats[k], equal_nan=True), k
    assert np.array_equal(parts[-1].summary["final_var"], full.summary["final_var"])
    assert np.array_equal(parts[-1].summary["grad_evals"], full.summary["grad_evals"])
    assert np.array_equal(state.log_bar, np.log(full.summary["final_step_size"])) or np.allclose(np.exp(state.log_bar), full.summary["final_step_size"], rtol=1e-15)


def test_store_warmup_false_slices_record_only_sampling_iterations():
    from pymc_b200 import engine, models
    from pymc_b200 import rng as brng

    spec = models.eight_schools()
    cm = engine.CompiledModel(spec)
    q0, sr, _ = _setup(spec, 4, 1)
    full = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=30, draws=10, mean0=q0, philox_seed=1, store_warmup=False)
    st = brng.pack_pcg64(sr)
    s1 = engine.ChainState(4, spec.n)
    a = cm.nuts_run(q0, st, tune=30, draws=10, mean0=q0, philox_seed=1, store_warmup=False, iter_begin=0, iter_count=33, save=s1)
    b = cm.nuts_run(q0, st, tune=30, draws=10, mean0=q0, philox_seed=1, store_warmup=False, iter_begin=33, resume=s1)
    assert a.draws.shape[1] == 3 and b.draws.shape[1] == 7
    assert np.array_equal(np.concatenate([a.draws, b.draws], axis=1), full.draws)


def test_pooled_warmup_gives_every_chain_the_pooled_mass_matrix_and_a_correct_posterior():
    from pymc_b200 import engine, models, parallel
    from pymc_b200 import rng as brng

    n = 12
    spec = models.std_normal(n)
    cm = engine.CompiledModel(spec)
    C = 256
    q0, sr, _ = _setup(spec, C, 11)
    res = parallel.pooled_warmup_run(cm, q0, brng.pack_pcg64(sr), tune=320, draws=200, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(),
                                     philox_seed=5)
    fv = res.summary["final_var"]
    assert np.all(fv == fv[0])                       # one pooled estimate for all chains
    assert np.all(np.abs(fv[0] - 1.0) < 0.05)        # ~25k pooled draws per window: far tighter than a per-chain estimate
    x = res.draws.reshape(-1, n)
    assert np.all(np.abs(x.mean(0)) < 0.03) and np.all(np.abs(x.var(0) - 1.0) < 0.04)
    eps = res.summary["final_step_size"]
    per_chain = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=320, draws=200, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(),
                            philox_seed=5, store_warmup=False)
    # the point of pooling: step sizes (hence tree sizes and per-chain run times) spread less than with per-chain adaptation
    assert np.std(np.log(eps)) < np.std(np.log(per_chain.summary["final_step_size"]))


def test_constrained_draws_recorded_by_the_kernel_equal_host_transform():
    from pymc_b200 import engine, ir
    from pymc_b200 import rng as brng

    m = ir.stochvol_ir(T=60)  # log, interval and identity transforms
    cm = engine.CompiledModel(m)
    q0, sr, _ = _setup(cm.spec, 3, 2)
    q0 = cm.spec.initial_point() + 0.1 * (q0 - cm.spec.initial_point())
    a = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=15, draws=10, mean0=q0, philox_seed=4)
    b = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=15, draws=10, mean0=q0, philox_seed=4, constrain=True)
    want = cm.spec.constrain(a.draws)
    got = cm.spec.split_rv(b.draws)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-14, atol=0)
    assert np.array_equal(a.stats["tree_size"], b.stats["tree_size"])


def test_sampling_api_init_and_step_variants_and_ir_models():
    import pymc_b200
    from pymc_b200 import ir

    m = ir.varying_intercept_logistic_ir()
    a = pymc_b200.sample_b200_nuts(150, tune=300, chains=16, random_seed=2, model=m)
    assert set(a.posterior) == {"mu_alpha", "sigma_alpha", "alpha", "beta"} and np.all(a.posterior["sigma_alpha"] > 0)
    assert set(a.groups()) >= {"posterior", "sample_stats", "observed_data", "constant_data"} and "y" in a.observed_data
    g = pymc_b200.sample_b200_nuts(150, tune=300, chains=16, random_seed=2, model=m, init="jitter+adapt_diag_grad")
    h = pymc_b200.sample_b200_nuts(150, tune=300, chains=16, random_seed=2, model=m, step="hmc")
    for other in (g, h):
        for k in ("mu_alpha", "beta"):
            se = np.sqrt(a.posterior[k].var((0, 1)) / 300 + other.posterior[k].var((0, 1)) / 300)
            assert np.all(np.abs(a.posterior[k].mean((0, 1)) - other.posterior[k].mean((0, 1))) < 5 * se), k
    assert h.sample_stats["tree_depth"].max() == 0 and h.sample_stats["n_steps"].min() >= 1
    es = pymc_b200.sample_b200_nuts(100, tune=200, chains=8, random_seed=1, model=ir.eight_schools_ir())  # routed to the hand kernel
    assert 3.0 < es.posterior["mu"].mean() < 6.0 and es.observed_data["y"].shape == (8,)
