"""from_pymc: lowering a pm.Model to ModelIR.  PyMC / PyTensor are not importable in the build image, so the
whole module is skipped there; wherever PyMC is installed it checks the lowering against the model's own compiled
logp/dlogp (the seam the IR replaces, pymc/model/core.py:464-529)."""
import numpy as np
import pytest

pm = pytest.importorskip("pymc")


def _check(model, rtol=1e-9):
    from oracle import ir_numpy
    from pymc_b200 import from_pymc

    m = from_pymc(model)
    f = model.logp_dlogp_function(ravel_inputs=True)
    f.set_extra_values({})
    g = ir_numpy.make_logp(m)
    rng = np.random.default_rng(0)
    for _ in range(3):
        q = m.initial_point() + rng.uniform(-0.5, 0.5, m.n)
        lp_ref, dlp_ref = f(q)
        lp, dlp = g(q)
        np.testing.assert_allclose(lp, lp_ref, rtol=rtol)
        np.testing.assert_allclose(dlp, dlp_ref, rtol=1e-7, atol=1e-9)
    return m


def test_eight_schools_lowers_and_specialises():
    from pymc_b200 import ir

    y = np.array([28.0, 8.0, -3.0, 7.0, -1.0, 1.0, 18.0, 12.0])
    s = np.array([15.0, 10.0, 16.0, 11.0, 9.0, 11.0, 10.0, 18.0])
    with pm.Model() as model:
        mu = pm.Normal("mu", 0, 5)
        tau = pm.HalfCauchy("tau", 5)
        theta_t = pm.Normal("theta_t", 0, 1, shape=8)
        pm.Normal("y", mu + tau * theta_t, s, observed=y)
    m = _check(model)
    assert [v.name for v in m.vars] == ["mu", "tau_log__", "theta_t"]
    assert ir.specialise(m) is not None


def test_radon_shape_lowers():
    rng = np.random.default_rng(1)
    J, N = 6, 50
    county = rng.integers(0, J, N)
    floor = (rng.random(N) < 0.3).astype(float)
    y = rng.normal(1.0, 0.8, N)
    with pm.Model() as model:
        mu_a = pm.Normal("mu_a", 0.0, 100**2)
        sigma_a = pm.HalfCauchy("sigma_a", 5)
        mu_b = pm.Normal("mu_b", 0.0, 100**2)
        sigma_b = pm.HalfCauchy("sigma_b", 5)
        a = pm.Normal("a", 0, 1, shape=J)
        b = pm.Normal("b", 0, 1, shape=J)
        eps = pm.HalfCauchy("eps", 5)
        pm.Normal("radon", (mu_a + a[county] * sigma_a) + (mu_b + b[county] * sigma_b) * floor, eps, observed=y)
    _check(model)


def test_hierarchical_logistic_lowers():
    rng = np.random.default_rng(2)
    N, G, K = 80, 5, 3
    g = rng.integers(0, G, N)
    X = rng.standard_normal((N, K))
    y = (rng.random(N) < 0.5).astype(int)
    with pm.Model() as model:
        mu = pm.Normal("mu", 0, 2)
        s = pm.HalfNormal("s", 1)
        alpha = pm.Normal("alpha", mu, s, shape=G)
        beta = pm.Normal("beta", 0, 2.5, shape=K)
        pm.Bernoulli("y", logit_p=alpha[g] + pm.math.dot(X, beta), observed=y)
    _check(model)


def test_unsupported_distribution_is_refused():
    from pymc_b200 import from_pymc

    with pm.Model() as model:
        pm.Weibull("w", 1.0, 2.0)
    with pytest.raises(NotImplementedError):
        from_pymc(model)
