"""Mean-field ADVI for the ``init="advi"``, ``"advi+adapt_diag"`` and ``"advi_map"`` branches of ``init_nuts``
(pymc/sampling/mcmc.py:1913-1980).

The reference fits ``pm.fit(method="advi", n=200000, obj_optimizer=pm.adagrad_window, callbacks=[CheckParametersConvergence(
tolerance=1e-2, diff="absolute"), CheckParametersConvergence(tolerance=1e-2, diff="relative")])`` and then uses the
approximation's mean / standard deviation as the mass matrix and draws of it as start points.  This module restates exactly
that much of the variational subsystem, on top of the engine's batched ``logp_dlogp``:

* the family: ``q(z) = Normal(mu, softplus(rho)^2)`` over the raveled unconstrained vector, ``mu = start``, ``rho = 0``
  (``MeanFieldGroup.create_shared_params``, variational/approximations.py:79-99);
* the objective: one-sample reparameterised estimate of ``KL(q || p)``: ``z = mu + sigma eps``,
  ``loss = log q(z) - log p(z)``, whose gradients are ``d/dmu = -grad logp(z)`` and
  ``d/dsigma = -grad logp(z) . eps - 1 / sigma`` (the ``mu`` dependence of ``log q`` cancels pathwise);
* the optimiser: ``adagrad_window(learning_rate=1e-3, epsilon=0.1, n_win=10)`` (variational/updates.py:542-585): each
  parameter is moved by ``lr * g / sqrt(sum of its last n_win squared gradients + epsilon)``;
* the stopping rule: every 100 iterations the flattened parameters ``[mu, rho]`` are compared with their values 100 iterations
  earlier; the fit stops when the infinity norm of the absolute difference, or of the relative difference
  ``(|d| + 1e-6) / (|prev| + 1e-6)``, is below ``tolerance`` (variational/callbacks.py:29-87).

The noise comes from a NumPy ``Generator`` (the reference uses PyTensor's RNG ops, whose stream is PyTensor-internal: like the
start-point jitter, SURVEY a17, parity with the reference's ADVI draws is unpinned -- the fixed point is the same).
"""
from __future__ import annotations

import numpy as np


def _softplus(x):
    return np.maximum(x, 0.0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def fit_meanfield(logp_dlogp, start, *, n: int = 200_000, seed=None, learning_rate: float = 1e-3, epsilon: float = 0.1,
                  n_win: int = 10, tolerance: float = 1e-2, every: int = 100):
    """-> (mean[n], std[n], iterations run).  ``logp_dlogp``: batched ``q[C, n] -> (logp[C], grad[C, n])`` (the engine's)."""
    rng = np.random.default_rng(seed)
    mu = np.array(start, dtype=np.float64).reshape(-1)
    d = mu.size
    rho = np.zeros(d)
    accu = np.zeros((2, d, n_win))  # squared-gradient windows of mu and rho
    slot = 0
    prev = None
    it = 0
    for it in range(1, int(n) + 1):
        sigma = _softplus(rho)
        eps = rng.standard_normal(d)
        z = mu + sigma * eps
        _, g = logp_dlogp(z[None, :])
        g = np.asarray(g[0], dtype=np.float64)
        if not np.all(np.isfinite(g)):
            raise FloatingPointError(f"ADVI: non-finite gradient of logp at iteration {it}")
        grads = (-g, (-(g * eps) - 1.0 / sigma) * _sigmoid(rho))  # d loss / d mu, d loss / d rho
        for k, (param, gr) in enumerate(((mu, grads[0]), (rho, grads[1]))):
            accu[k, :, slot] = gr * gr
            param -= learning_rate * gr / np.sqrt(accu[k].sum(axis=-1) + epsilon)
        slot = slot + 1 if slot + 1 < n_win else 0
        cur = np.concatenate([mu, rho])
        if prev is None:
            prev = cur
            continue
        if it % every or it < every:
            continue
        diff = np.abs(cur - prev)
        rel = (diff + 1e-6) / (np.abs(prev) + 1e-6)
        prev = cur
        if diff.max() < tolerance or rel.max() < tolerance:
            break
    return mu, _softplus(rho), it
