"""HamiltonianMC (hmc/hmc.py:143-200, incl. the uniform(0.85, 1.15) step-size jitter drawn from the
chain's step stream) and init="jitter+adapt_diag_grad" (QuadPotentialDiagAdaptExp, quadpotential.py:493-579).

Goldens come from the verbatim reference (oracle/make_golden.py f3).  CPU: the oracle restatements reproduce them bit for
bit.  GPU: the persistent kernel follows the reference chain over a cold-start prefix (adaptive dynamics are chaotic, see
DESIGN.md), with every discrete statistic identical."""
import numpy as np
import pytest

from b200_helpers import relerr, start_states
from pymc_b200 import models

SPEC = {"eight_schools": models.eight_schools, "radon": models.radon}


def _gen(state_u64x4):
    g = np.random.default_rng(0)
    st = g.bit_generator.state
    st["state"]["state"] = (int(state_u64x4[0]) << 64) | int(state_u64x4[1])
    st["state"]["inc"] = (int(state_u64x4[2]) << 64) | int(state_u64x4[3])
    st["has_uint32"], st["uinteger"] = 0, 0
    g.bit_generator.state = st
    return g


@pytest.mark.parametrize("name", ["eight_schools", "radon"])
def test_oracle_hmc_reproduces_reference_golden(golden, name):
    from oracle import logp_numpy, nuts_numpy

    d = golden(name + "_hmc_adapt")
    spec = SPEC[name]()
    f = logp_numpy.make_logp(spec)
    for c in range(len(d["seeds"])):
        mass = nuts_numpy.DiagMass(np.ones(spec.n), adapt=True, initial_mean=d["q0"][c].copy(), initial_weight=10)
        o = nuts_numpy.Oracle(f, mass, sampler="hmc", target_accept=0.65)
        o.rng = _gen(d["pre_rng"][c][0])
        qs, st = o.run(d["q0"][c], int(d["tune"]), int(d["draws"]), z=d["z"][c])
        assert np.array_equal(st["tree_size"], d["stat_tree_size"][c])
        assert np.array_equal(qs, d["draws_q"][c])
        assert np.array_equal(st["mean_tree_accept"], d["stat_mean_tree_accept"][c])
        assert np.array_equal(st["index_in_trajectory"] > 0, d["stat_accepted"][c].astype(bool))


@pytest.mark.parametrize("name", ["eight_schools", "radon"])
def test_oracle_adapt_diag_grad_reproduces_reference_golden(golden, name):
    from oracle import logp_numpy, nuts_numpy

    d = golden(name + "_adapt_grad")
    spec = SPEC[name]()
    f = logp_numpy.make_logp(spec)
    for c in range(len(d["seeds"])):
        mass = nuts_numpy.DiagMassExp(spec.n, alpha=float(d["alpha"]), stop_adaptation=int(d["stop_adaptation"]),
                                      discard_window=int(d["discard_window"]), initial_mean=d["q0"][c].copy())
        o = nuts_numpy.Oracle(f, mass)
        o.rng = _gen(d["pre_rng"][c][0])
        qs, st = o.run(d["q0"][c], int(d["tune"]), int(d["draws"]), z=d["z"][c])
        assert np.array_equal(st["tree_size"], d["stat_tree_size"][c])
        assert np.array_equal(qs, d["draws_q"][c])
        assert np.array_equal(mass.var, d["final_var"][c])


# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eight_schools", "radon"])
def test_gpu_hmc_follows_the_reference_chain(golden, name):
    from pymc_b200 import engine

    d = golden(name + "_hmc_adapt")
    cm = engine.CompiledModel(SPEC[name]())
    T = 12  # cold-start prefix (all tuning iterations)
    res = cm.nuts_run(d["q0"], start_states(d), tune=T, draws=0, z=np.ascontiguousarray(d["z"][:, :T]), sampler="hmc",
                      target_accept=0.65, mass="diag_adapt", mean0=d["q0"], var0=np.ones_like(d["q0"]))
    for c in range(len(d["seeds"])):
        assert np.array_equal(res.stats["tree_size"][c], d["stat_tree_size"][c][:T])  # n_steps incl. the jitter draw
        assert np.array_equal(res.stats["index_in_trajectory"][c] > 0, d["stat_accepted"][c][:T].astype(bool))
        assert np.array_equal(res.stats["diverging"][c].astype(bool), d["stat_diverging"][c][:T].astype(bool))
        assert np.max(np.abs(res.draws[c] - d["draws_q"][c][:T])) <= 1e-7
        assert relerr(res.stats["mean_tree_accept"][c], d["stat_mean_tree_accept"][c][:T]) <= 1e-6
        assert relerr(res.stats["step_size"][c], d["stat_step_size"][c][:T]) <= 1e-8
        assert np.array_equal(res.summary["grad_evals"][c], d["stat_tree_size"][c][:T].sum() + T)


@pytest.mark.gpu
def test_gpu_hmc_stream_position_and_sampling_phase(golden):
    """The whole golden run (tuning + sampling): the device stream ends where the reference's Generator ends as long as
    the chain is on the reference path; here checked on the Eight Schools chain that stays on it."""
    from pymc_b200 import engine

    d = golden("eight_schools_hmc_adapt")
    cm = engine.CompiledModel(models.eight_schools())
    tune, draws = int(d["tune"]), int(d["draws"])
    res = cm.nuts_run(d["q0"], start_states(d), tune=tune, draws=draws, z=d["z"], sampler="hmc", target_accept=0.65,
                      mass="diag_adapt", mean0=d["q0"], var0=np.ones_like(d["q0"]))
    ok = [np.array_equal(res.stats["tree_size"][c], d["stat_tree_size"][c]) for c in range(len(d["seeds"]))]
    assert any(ok), "no chain stayed on the reference path for 50 HMC transitions"
    c = ok.index(True)
    assert np.max(np.abs(res.draws[c] - d["draws_q"][c])) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["eight_schools", "radon"])
def test_gpu_adapt_diag_grad_follows_the_reference_chain(golden, name):
    from pymc_b200 import engine

    d = golden(name + "_adapt_grad")
    cm = engine.CompiledModel(SPEC[name]())
    T = 32 if name == "radon" else 40  # covers the start of the estimators (draw 10) and the first mass updates (draw 21 on)
    kw = dict(mass="diag_adapt_grad", mass_alpha=float(d["alpha"]), stop_adaptation=int(d["stop_adaptation"]),
              discard_window=int(d["discard_window"]))
    res = cm.nuts_run(d["q0"], start_states(d), tune=T, draws=0, z=np.ascontiguousarray(d["z"][:, :T]), **kw)
    first_bad = []
    for c in range(len(d["seeds"])):
        same = (res.stats["tree_size"][c] == d["stat_tree_size"][c][:T]) & (res.stats["depth"][c] == d["stat_depth"][c][:T])
        first_bad.append(int(np.argmin(same)) if not same.all() else T)
    from test_gpu_fullsize import _report

    _report("first_divergent_draw/" + name + "_adapt_grad", {"first_tree_mismatch_per_chain": first_bad, "draws": T,
                                                             "mass_updates_start_at_draw": 21})
    # The gradient-based mass matrix feeds back into the dynamics from draw 21 on.  Eight Schools stays on the reference
    # path well past it (the update rule itself is model-independent); a cold-start Radon chain leaves the reference path
    # by chaos around draw 27 with ANY adaptation (parity_report: first_divergent_draw/radon_adapt), so it only has to get there.
    assert max(first_bad) >= (28 if name == "eight_schools" else 15), first_bad
    c = int(np.argmax(first_bad))
    m = slice(0, min(first_bad[c], 26))
    tol = 1e-6 if name == "eight_schools" else 1e-3   # the cold-start Radon chain amplifies ulp differences ~10x per 4 draws
    assert np.max(np.abs(res.draws[c][m] - d["draws_q"][c][m])) <= tol
    assert np.max(np.abs(res.draws[c][:12] - d["draws_q"][c][:12])) <= 1e-7
    assert relerr(res.stats["step_size"][c][m], d["stat_step_size"][c][m]) <= (1e-7 if name == "eight_schools" else 1e-3)


@pytest.mark.gpu
def test_gpu_adapt_diag_grad_final_mass_matrix_from_a_tune_only_run(golden):
    """final_var after a tuning-only run includes the update that follows the last iteration (needs one more gradient)."""
    from oracle import logp_numpy, nuts_numpy
    from pymc_b200 import engine
    from pymc_b200 import rng as brng

    spec = models.eight_schools()
    cm = engine.CompiledModel(spec)
    f = logp_numpy.make_logp(spec)
    T = 25
    sr, pr, _ = brng.chain_generators(3, 1)
    q0 = spec.initial_point()[None] + 0.1
    z = brng.momentum_noise(pr, T, spec.n)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=T, draws=0, z=z, mass="diag_adapt_grad", discard_window=5, mass_alpha=0.05)
    mass = nuts_numpy.DiagMassExp(spec.n, alpha=0.05, discard_window=5, initial_mean=q0[0].copy())
    o = nuts_numpy.Oracle(f, mass)
    o.rng = sr[0]
    qs, st = o.run(q0[0], T, 0, z=z[0])
    if np.array_equal(st["tree_size"], res.stats["tree_size"][0]):  # on the oracle's path: the mass matrices agree
        assert relerr(res.summary["final_var"][0], mass.var) <= 1e-6
    assert np.all(np.isfinite(res.summary["final_var"])) and np.all(res.summary["final_var"] > 0)


# ------------------------------------------------------------------------------------------------------------------------
# dense mass matrix on a non-Gaussian model (QuadPotentialFull / QuadPotentialFullInv are model-independent in the reference)
# ------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["full", "fullinv"])
def test_oracle_dense_mass_on_radon_reproduces_reference_golden(golden, tag):
    from oracle import logp_numpy, nuts_numpy

    d = golden(f"radon_dense_{tag}_fixed")
    spec = models.radon()
    f = logp_numpy.make_logp(spec)
    for c in range(len(d["seeds"])):
        mass = nuts_numpy.DenseMass(d["cov"]) if tag == "full" else nuts_numpy.DenseInvMass(d["A"])
        o = nuts_numpy.Oracle(f, mass, adapt_step_size=False)
        o.da = nuts_numpy.DualAveraging(float(d["eps"][c]))
        o.rng, o.tune = _gen(d["pre_rng"][c][0]), False
        qs, st = o.run(d["q0"][c], 0, int(d["draws"]), z=d["z"][c])
        assert np.array_equal(st["tree_size"], d["stat_tree_size"][c])
        assert np.max(np.abs(qs - d["draws_q"][c])) <= 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["full", "fullinv"])
def test_gpu_dense_mass_on_radon_matches_reference(golden, tag):
    """Any model under a dense mass matrix advances in lock step: its own fused logp+grad kernel per batch + one fp64
    tensor-core GEMM (Sigma . grad) per leapfrog.  Same draws as the reference's QuadPotentialFull / FullInv."""
    from b200_helpers import discrete_equal
    from pymc_b200 import engine

    d = golden(f"radon_dense_{tag}_fixed")
    cm = engine.CompiledModel(models.radon())
    if tag == "full":
        cm.set_dense_mass(cov=d["cov"])
    else:
        cm.set_dense_mass(inverse=d["A"])
    res = cm.nuts_run(d["q0"], start_states(d), tune=0, draws=int(d["draws"]), z=d["z"], mass="dense", adapt_step_size=False,
                      eps0=d["eps"])
    for c in range(len(d["seeds"])):
        st = {k: v[c] for k, v in res.stats.items()}
        assert discrete_equal(st, d, c).all(), (tag, c, st["tree_size"], d["stat_tree_size"][c])
        assert np.max(np.abs(res.draws[c] - d["draws_q"][c])) <= 1e-8
        assert relerr(st["energy"], d["stat_energy"][c]) <= 1e-8
