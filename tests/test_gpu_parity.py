"""-m gpu: the CUDA engine (through the C ABI) against the oracle and the committed golden vectors.

Tolerances (fp64): logp/grad <= 1e-10 relative (north_star asks <= 1e-6); accepted positions of the
fixed-step-size replays <= 1e-9 absolute with every discrete tree statistic identical; adaptive warm-up
is chaotic in the reference itself (a 1e-15 perturbation of the start decorrelates the reference's own
chain after ~35 draws), so it is checked draw by draw from the golden pre-draw state.
"""
import numpy as np
import pytest

from b200_helpers import (CONTINUOUS, SPEC_OF, discrete_equal, gpu_free_run, gpu_single_draws, relerr,
                           stream_states)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def compiled():
    from pymc_b200 import engine

    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = engine.CompiledModel(SPEC_OF[name]())
        return cache[name]

    return get


MODELS = ["std_normal_fixed", "eight_schools_fixed", "radon_fixed", "radon_small_adapt",
          "std_normal_team_fixed", "stochvol_small_fixed", "stochvol_fixed",  # these three: chain-per-CTA kernels
          "logistic_small_fixed", "mvgauss_dense_fixed"]  # lock-step (GEMM-shaped) engine


@pytest.mark.parametrize("name", MODELS)
def test_logp_grad_matches_oracle(compiled, name):
    from oracle import logp_numpy

    cm = compiled(name)
    spec = cm.spec
    f = logp_numpy.make_logp(spec)
    rng = np.random.default_rng(0)
    nq, amp = (257, 2.0) if spec.n < 1000 else (19, 0.5)
    Q = spec.initial_point() + rng.uniform(-amp, amp, (nq, spec.n))  # ragged vs the 8-warp CTA
    lp, g = cm.logp_dlogp(Q)
    lo = np.array([f(q)[0] for q in Q])
    go = np.array([f(q)[1] for q in Q])
    assert relerr(lp, lo) <= 1e-12
    scale = np.max(np.abs(go), axis=1, keepdims=True)
    assert np.max(np.abs(g - go) / scale) <= 1e-12
    one_lp, one_g = cm.logp_dlogp(Q[3])
    if spec.name in ("logistic", "mvgauss"):  # batched contraction: the reduction order depends on the batch width
        assert relerr(one_lp, lp[3]) <= 1e-13 and np.max(np.abs(one_g - g[3])) <= 1e-12 * scale[3]
    else:
        assert one_lp == lp[3] and np.array_equal(one_g, g[3])


def test_eight_schools_known_answer(compiled):
    """SURVEY 8c loader self-check values (produced by the verbatim reference + NumPy logp)."""
    lp, g = compiled("eight_schools_fixed").logp_dlogp(np.zeros(10))
    assert abs(lp - (-43.43563727714813)) <= 1e-12
    np.testing.assert_allclose(g[:3], [0.46353275, 0.92307692, 0.12444444], atol=1e-8)


@pytest.mark.parametrize("eps", [0.01, 0.1])
@pytest.mark.parametrize("n_steps", [1, 2, 3, 4, 20])
def test_leapfrog_reversible(compiled, eps, n_steps):
    """tests/step_methods/hmc/test_hmc.py:49-74 (test_leapfrog_reversible): n steps forward then back, rtol 1e-5."""
    cm = compiled("radon_fixed")
    d_rng = np.random.default_rng(5)
    C = 33
    q0 = cm.spec.initial_point() + d_rng.uniform(-0.3, 0.3, (C, cm.n))
    p0 = d_rng.standard_normal((C, cm.n))
    var = np.exp(d_rng.uniform(-1, 1, (C, cm.n)))
    s = cm.leapfrog(q0, p0, var, eps * 0.1, 0)
    f = cm.leapfrog(s["q"], s["p"], var, eps * 0.1, n_steps, grad=s["grad"])
    assert np.all(f["idx"] == n_steps)
    b = cm.leapfrog(f["q"], f["p"], var, -eps * 0.1, n_steps, grad=f["grad"], idx=f["idx"])
    assert np.all(b["idx"] == 0)
    np.testing.assert_allclose(b["q"], q0, rtol=1e-5)
    np.testing.assert_allclose(b["p"], p0, rtol=1e-5)
    np.testing.assert_allclose(b["energy"], s["energy"], rtol=1e-9)


@pytest.mark.parametrize("name", ["eight_schools_fixed", "radon_fixed", "stochvol_small_fixed"])
def test_leapfrog_matches_oracle(compiled, name):
    from oracle import logp_numpy, nuts_numpy

    cm = compiled(name)
    spec = cm.spec
    rng = np.random.default_rng(9)
    C = 5
    q0 = spec.initial_point() + rng.uniform(-0.2, 0.2, (C, spec.n))
    p0 = rng.standard_normal((C, spec.n))
    var = np.exp(rng.uniform(-0.5, 0.5, (C, spec.n)))
    eps = np.array([0.01, -0.01, 0.02, 0.005, -0.015])
    out = cm.leapfrog(q0, p0, var, eps, 7)
    f = logp_numpy.make_logp(spec)
    for c in range(C):
        o = nuts_numpy.Oracle(f, nuts_numpy.DiagMass(var[c]))
        s = o._start_state(q0[c], p0[c])
        for _ in range(7):
            s = o._leapfrog(eps[c], s)
        assert relerr(out["q"][c], s.q) <= 1e-11 and relerr(out["p"][c], s.p) <= 1e-10
        assert abs(out["energy"][c] - s.energy) <= 1e-9 * abs(s.energy)
        assert out["idx"][c] == s.idx


@pytest.mark.parametrize("name", ["std_normal_fixed", "eight_schools_fixed", "radon_fixed", "std_normal_team_fixed",
                                  "stochvol_small_fixed", "stochvol_fixed", "logistic_small_fixed",
                                  "mvgauss_dense_fixed"])
def test_nuts_fixed_step_identical_draws(compiled, golden, name):
    """Same stream seeds, same fixed step size and mass matrix => same accepted draws as the reference."""
    d = golden(name)
    res, states = gpu_free_run(compiled(name), d, name)
    for c in range(len(d["seeds"])):
        st = {k: v[c] for k, v in res.stats.items()}
        assert discrete_equal(st, d, c).all(), f"{name} chain {c}: tree statistics differ from the reference"
        assert np.max(np.abs(res.draws[c] - d["draws_q"][c])) <= 1e-9
        for k in CONTINUOUS:
            if k in ("energy_error", "max_energy_error"):  # differences of energies: absolute, on the energy scale
                scale = max(1.0, float(np.max(np.abs(d["stat_energy"][c]))))
                assert np.max(np.abs(st[k] - d["stat_" + k][c])) <= 1e-9 * scale, k
            else:
                assert relerr(st[k], d["stat_" + k][c]) <= 1e-8, k
    assert np.all(res.summary["bad_energy_at"] == -1)
    # grad evaluations: tree_size + 1 per draw (compute_state), SURVEY 3.2.  The lock-step engine carries the accepted
    # proposal's (logp, grad) into the next draw, so only the very first start state is evaluated.
    extra = 1 if compiled(name).spec.name in ("logistic", "mvgauss") else res.stats["tree_size"].shape[1]
    assert np.array_equal(res.summary["grad_evals"], res.stats["tree_size"].sum(axis=1) + extra)


@pytest.mark.parametrize("name", ["eight_schools_fixed", "radon_fixed"])
def test_stream_consumption_order(compiled, golden, name):
    """After T-1 draws the device stream must sit exactly where the reference's Generator sat before draw T-1
    (one uniform per doubling, per completed merge and per top-level pick: SURVEY 8a a15)."""
    d = golden(name)
    T = int(d["draws"])
    res, states = gpu_free_run(compiled(name), {**d, "draws": np.int64(T - 1), "z": d["z"][:, : T - 1]}, name)
    want = stream_states(d["pre_rng"][:, T - 1])
    assert np.array_equal(states.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("name", ["eight_schools_adapt", "radon_adapt", "radon_small_adapt", "stochvol_small_adapt",
                                  "logistic_small_adapt"])
def test_nuts_single_draw_replay_of_adaptive_run(compiled, golden, name):
    """Every draw of a full adaptive reference run (huge early step sizes, divergences, depth caps) replayed
    from its golden pre-draw state."""
    d = golden(name)
    dq, st = gpu_single_draws(compiled(name), d, name, chain=0)
    ok = discrete_equal(st, d, 0)
    T = len(ok)
    # a trajectory integrated with a wildly unstable step size amplifies ulp differences within ONE draw;
    # allow a handful of such flips but nothing systematic
    assert ok.mean() >= 0.97, f"{name}: only {ok.sum()} of {T} draws reproduce the reference tree"
    err = np.max(np.abs(dq - d["draws_q"][0]), axis=1)[ok]
    assert np.quantile(err, 0.95) <= 1e-9 and err.max() <= 1e-4, (np.quantile(err, 0.95), err.max())
    assert np.array_equal(st["diverging"][ok], d["stat_diverging"][0][ok])
    assert relerr(st["mean_tree_accept"][ok], d["stat_mean_tree_accept"][0][ok]) <= 1e-4


@pytest.mark.parametrize("name", ["eight_schools_warm_adapt", "radon_warm_adapt"])
def test_nuts_adaptation_from_warm_state(compiled, golden, name):
    """Dual averaging + Welford windows (two window switches) with adaptation ON, from a tuned state."""
    d = golden(name)
    res, _ = gpu_free_run(compiled(name), d, name)
    st = {k: v[0] for k, v in res.stats.items()}
    ok = discrete_equal(st, d, 0)
    first_bad = int(np.argmin(ok)) if not ok.all() else len(ok)
    # adaptation feeds the step size back into the dynamics, so ulp differences grow exponentially even from a
    # warm state (the error curve is smooth: no jump at the Welford window switches at t=102 and t=203)
    assert first_bad >= 120, f"{name}: diverged from the reference at draw {first_bad}"
    m = slice(0, 100)
    assert relerr(st["step_size"][m], d["stat_step_size"][0][m]) <= 1e-5
    assert np.max(np.abs(res.draws[0][m] - d["draws_q"][0][m])) <= 1e-4
    m = slice(0, 30)
    assert relerr(st["step_size"][m], d["stat_step_size"][0][m]) <= 1e-8
    assert np.max(np.abs(res.draws[0][m] - d["draws_q"][0][m])) <= 1e-7
    if first_bad == len(ok):  # still on the reference's path at the end: the adapted mass matrix agrees
        assert relerr(res.summary["final_var"][0], d["final_var"][0]) <= 1e-2


@pytest.mark.parametrize("name", ["eight_schools_adapt", "radon_adapt"])
def test_nuts_cold_adaptation_prefix(compiled, golden, name):
    d = golden(name)
    res, _ = gpu_free_run(compiled(name), d, name, draws=12)
    for c in range(len(d["seeds"])):
        st = {k: v[c] for k, v in res.stats.items()}
        dd = {k: (v[:, :12] if k.startswith("stat_") else v) for k, v in d.items()}
        assert discrete_equal(st, dd, c).all()
        assert relerr(st["step_size"], d["stat_step_size"][c][:12]) <= 1e-9
        assert np.max(np.abs(res.draws[c] - d["draws_q"][c][:12])) <= 1e-7


def test_dense_mass_step_size_adaptation_prefix(compiled, golden):
    """QuadPotentialFull (fixed dense covariance) + dual averaging, lock-step engine: the first iterations of the
    reference run (one mass GEMM per leapfrog instead of the reference's two: rounding-level differences only)."""
    name = "mvgauss_dense_stepadapt"
    d = golden(name)
    T = 25
    res, _ = gpu_free_run(compiled(name), d, name, draws=T)
    st = {k: v[0] for k, v in res.stats.items()}
    dd = {k: (v[:, :T] if k.startswith("stat_") else v) for k, v in d.items()}
    assert discrete_equal(st, dd, 0).all()
    assert relerr(st["step_size"], d["stat_step_size"][0][:T]) <= 1e-8
    assert np.max(np.abs(res.draws[0] - d["draws_q"][0][:T])) <= 1e-7


def test_device_philox_momentum_matches_host_replica(compiled):
    """momentum_source=DEVICE_PHILOX draws the normals of pymc_b200.rng.philox_normal (NumPy replica)."""
    from pymc_b200 import rng as brng

    cm = compiled("eight_schools_fixed")
    n, C, T, key = cm.n, 3, 6, 0x1234_5678_9ABC_DEF0
    q0 = np.zeros((C, n))
    sr, _, _ = brng.chain_generators(5, C)
    s1, s2 = brng.pack_pcg64(sr), brng.pack_pcg64(sr)
    z = brng.philox_normal(key, np.arange(C)[:, None, None], np.arange(T)[None, :, None], np.arange(n)[None, None, :])
    a = cm.nuts_run(q0, s1, tune=0, draws=T, mass="diag", adapt_step_size=False, philox_seed=key)
    b = cm.nuts_run(q0, s2, tune=0, draws=T, mass="diag", adapt_step_size=False, z=z)
    assert np.array_equal(a.stats["tree_size"], b.stats["tree_size"])
    assert np.max(np.abs(a.draws - b.draws)) <= 1e-9


def test_bad_initial_energy_freezes_chain(compiled):
    from pymc_b200 import rng as brng

    cm = compiled("eight_schools_fixed")
    q0 = np.zeros((2, cm.n))
    q0[1, 1] = 800.0  # tau = exp(800) = inf -> logp = -inf/NaN -> "Bad initial energy" (base_hmc.py:205-224)
    sr, _, _ = brng.chain_generators(1, 2)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=5, mass="diag", adapt_step_size=False, philox_seed=1)
    assert res.summary["bad_energy_at"][0] == -1 and res.summary["bad_energy_at"][1] == 0
    assert np.isnan(res.draws[1]).all() and np.isfinite(res.draws[0]).all()


def test_chain_results_do_not_depend_on_batch(compiled):
    """A chain's draws depend only on its own inputs, not on how many chains share the launch (sharding invariance)."""
    from pymc_b200 import rng as brng

    cm = compiled("radon_fixed")
    C = 37
    rng = np.random.default_rng(3)
    q0 = cm.spec.initial_point() + rng.uniform(-1, 1, (C, cm.n))
    sr, _, _ = brng.chain_generators(17, C)
    st = brng.pack_pcg64(sr)
    # the Philox stream is keyed by chain index, so give the sub-batch the same noise explicitly
    z = brng.philox_normal(99, np.arange(C)[:, None, None], np.arange(30)[None, :, None], np.arange(cm.n)[None, None, :])
    full = cm.nuts_run(q0, st.copy(), tune=20, draws=10, z=z, mean0=q0)
    sub = cm.nuts_run(q0[30:35], st[30:35].copy(), tune=20, draws=10, z=z[30:35], mean0=q0[30:35])
    assert np.array_equal(full.draws[30:35], sub.draws)
    assert np.array_equal(full.stats["tree_size"][30:35], sub.stats["tree_size"])


def test_sampler_statistics_std_normal(compiled):
    """tests/sampler_fixtures.py style check: N(0, I) target, pooled mean/variance and adapted step size."""
    from pymc_b200 import engine, models
    from pymc_b200 import rng as brng

    cm = engine.CompiledModel(models.std_normal(10))
    C = 512
    rng = np.random.default_rng(0)
    q0 = rng.uniform(-1, 1, (C, 10))
    sr, _, _ = brng.chain_generators(42, C)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=300, draws=200, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(),
                      store_warmup=False, philox_seed=7)
    x = res.draws.reshape(-1, 10)
    assert np.all(np.abs(x.mean(0)) < 0.02) and np.all(np.abs(x.var(0) - 1.0) < 0.03)
    acc = res.stats["mean_tree_accept"].mean()
    assert 0.7 < acc < 0.92, acc
    assert res.stats["diverging"].sum() == 0
    fv = res.summary["final_var"]  # Welford over ~100-200 draws: sd ~ 0.12 per entry
    assert abs(fv.mean() - 1.0) < 0.03 and np.all(np.abs(fv - 1.0) < 0.9)


def test_radon_full_size_run_properties(compiled):
    """BASELINE config 2 shape (2048 chains), shortened: size-independent properties of a healthy run."""
    from pymc_b200 import rng as brng

    cm = compiled("radon_fixed")
    C = 2048
    rng = np.random.default_rng(1)
    q0 = cm.spec.initial_point() + rng.uniform(-1, 1, (C, cm.n))
    sr, _, _ = brng.chain_generators(2024, C)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=300, draws=100, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(),
                      store_warmup=False, philox_seed=11)
    assert np.all(res.summary["bad_energy_at"] == -1)
    assert np.isfinite(res.draws).all()
    ts = res.stats["tree_size"]
    assert ts.min() >= 1 and ts.max() <= 1023
    assert np.all(res.stats["depth"] <= 10)
    # tree_size is consistent with depth: 2^(depth-1) <= tree_size <= 2^depth - 1
    dep = res.stats["depth"].astype(np.int64)
    assert np.all(ts <= (1 << dep) - 1) and np.all(ts >= (1 << (dep - 1)) - 0)
    assert res.stats["diverging"].mean() < 0.02
    eps = res.summary["final_step_size"]
    assert 0.02 < np.median(eps) < 1.0
    # chains agree with each other (R-hat-like): between-chain variance of chain means is small
    mu_a = res.draws[:, :, 0]
    B = mu_a.mean(1).var() * mu_a.shape[1]
    W = mu_a.var(1).mean()
    assert B / W < 3.0


def test_reference_sampler_drives_cuda_logp_through_seam_b1(compiled):
    """Seam B1: the (restated) reference NUTS calling the CUDA logp/grad one point at a time gives the same
    chain as with the NumPy logp -- correctness of a1 behind the reference's own tree builder."""
    from oracle import logp_numpy, nuts_numpy
    from pymc_b200.step import B200LogpDlogp

    cm = compiled("eight_schools_fixed")
    f_np = logp_numpy.make_logp(cm.spec)
    f_cu = B200LogpDlogp(cm)
    outs = []
    for f in (f_np, f_cu._pytensor_function):
        o = nuts_numpy.Oracle(f, nuts_numpy.DiagMass(np.ones(cm.n)), adapt_step_size=False)
        o.setup_chain(np.random.default_rng(3))
        o.tune = False
        outs.append(o.run(np.zeros(cm.n), 0, 12))
    assert np.array_equal(outs[0][1]["tree_size"], outs[1][1]["tree_size"])
    assert np.max(np.abs(outs[0][0] - outs[1][0])) <= 1e-9


def test_sample_b200_nuts_public_api(compiled):
    """Seam B4: signature/return conventions of sample_jax_nuts; seed reproducibility (test_mcmc_external.py:83)."""
    import pymc_b200
    from pymc_b200 import models

    spec = models.eight_schools()
    a = pymc_b200.sample_b200_nuts(200, tune=300, chains=8, random_seed=11, model=spec, keep_untransformed=True)
    b = pymc_b200.sample_b200_nuts(200, tune=300, chains=8, random_seed=11, model=spec, keep_untransformed=True)
    assert set(a.posterior) == {"mu", "tau", "theta_t"}
    assert a.posterior["mu"].shape == (8, 200) and a.posterior["theta_t"].shape == (8, 200, 8)
    assert {"diverging", "energy", "tree_depth", "n_steps", "acceptance_rate", "lp", "step_size"} <= set(a.sample_stats)
    assert {"sampling_time", "tuning_steps"} <= set(a.attrs) and a.attrs["tuning_steps"] == 300
    assert np.array_equal(a.unconstrained, b.unconstrained)
    assert np.all(a.posterior["tau"] > 0)
    # posterior sanity vs the well-known Eight Schools answer (mu ~ 4.4, tau median ~ 2.7)
    assert 3.0 < a.posterior["mu"].mean() < 6.0
    c = pymc_b200.sample_b200_nuts(50, tune=100, chains=2, random_seed=5, model=spec, momentum="numpy",
                                   discard_tuned_samples=False)
    assert c.warmup_posterior["mu"].shape == (2, 100) and c.posterior["mu"].shape == (2, 50)
