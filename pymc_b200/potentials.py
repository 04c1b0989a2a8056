"""Seam B2, inbound: a user-supplied ``QuadPotential`` object (``pm.NUTS(potential=...)``, hmc/base_hmc.py:82-169) configures
the engine's mass matrix.

The reference passes the potential object itself to the integrator, which calls ``velocity / energy / random / update`` on it
once or twice per leapfrog (hmc/integration.py:68-145, quadpotential.py:121-181).  On the device the potentials live INSIDE the
kernels, so the object is not called: it is read -- class and constructor state -- and mapped onto the engine's mass kinds:

    QuadPotentialDiag(v)                         :582-630   ->  mass="diag",            var0 = v
    QuadPotentialDiagAdapt(n, mean, diag, w, ..) :211-355   ->  mass="diag_adapt",      mean0, var0, weight, windows
    QuadPotentialDiagAdaptExp(.., use_grads=True):493-579   ->  mass="diag_adapt_grad", alpha, stop_adaptation, discard window
    QuadPotentialFull(cov)                       :680-725   ->  mass="dense",           set_dense_mass(cov=cov)
    QuadPotentialFullInv(A)                      :633-677   ->  mass="dense",           set_dense_mass(inverse=A)
    QuadPotentialFullAdapt(n, mean, cov, w, ..)  :748-845   ->  mass="dense_adapt",     mean0, var0 = diag(cov), weight, windows

Matching is by class name and attributes (duck typing: PyMC is not importable in the build image); options the kernels do not
implement (``early_update``, a growing diagonal window, ``use_grads=False``, a non-diagonal initial covariance) raise
``NotImplementedError`` naming the option instead of being ignored.  A user SUBCLASS with its own ``velocity`` cannot be run on
the device and is refused the same way.
"""
from __future__ import annotations

import numpy as np

_KNOWN = ("QuadPotentialDiag", "QuadPotentialDiagAdapt", "QuadPotentialDiagAdaptExp", "QuadPotentialFull", "QuadPotentialFullInv",
          "QuadPotentialFullAdapt")


def engine_kwargs(potential, n: int) -> dict:
    """-> keyword arguments for ``CompiledModel.nuts_run`` (``mass``, ``var0`` [n], ``mean0`` [n], windows ...) plus, for a fixed
    dense matrix, ``dense_cov`` or ``dense_inverse`` ([n, n]) for ``CompiledModel.set_dense_mass``."""
    kind = type(potential).__name__
    if kind not in _KNOWN:
        raise NotImplementedError(f"potential {kind}: only the reference's own QuadPotential classes {_KNOWN} can be mapped onto "
                                  "the device kernels (a subclass with its own velocity()/update() runs in Python only)")

    def vec(a, what):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        if a.shape != (n,):
            raise ValueError(f"{kind}.{what}: expected {n} values, got {a.shape}")
        return a

    if kind == "QuadPotentialDiag":
        return {"mass": "diag", "var0": vec(potential.v, "v")}
    if kind in ("QuadPotentialDiagAdapt", "QuadPotentialDiagAdaptExp"):
        if getattr(potential, "_early_update", False):
            raise NotImplementedError(f"{kind}(early_update=True) is not implemented by the kernels")
        out = {"var0": vec(potential._initial_diag, "initial_diag"), "mean0": vec(potential._initial_mean, "initial_mean"),
               "discard_window": int(potential._discard_window)}
        if kind == "QuadPotentialDiagAdaptExp":
            if not getattr(potential, "_use_grads", False):
                raise NotImplementedError("QuadPotentialDiagAdaptExp(use_grads=False) is not implemented (init_nuts always passes "
                                          "use_grads=True, sampling/mcmc.py:1905-1912)")
            stop = potential._stop_adaptation
            out.update(mass="diag_adapt_grad", mass_alpha=float(potential._alpha),
                       stop_adaptation=None if not np.isfinite(stop) else int(stop))
            return out
        if float(potential.adaptation_window_multiplier) != 1.0:
            raise NotImplementedError("QuadPotentialDiagAdapt with adaptation_window_multiplier != 1: the diagonal kernels keep the "
                                      "window fixed (the reference's default)")
        out.update(mass="diag_adapt", mass_initial_weight=float(potential._initial_weight),
                   adaptation_window=int(potential.adaptation_window))
        return out
    if kind == "QuadPotentialFull":
        cov = np.asarray(potential._cov, dtype=np.float64)
        if cov.shape != (n, n):
            raise ValueError(f"QuadPotentialFull: covariance is {cov.shape}, the model has {n} unconstrained values")
        return {"mass": "dense", "dense_cov": cov}
    if kind == "QuadPotentialFullInv":
        L = np.asarray(potential.L, dtype=np.float64)  # chol(A, lower)
        if L.shape != (n, n):
            raise ValueError(f"QuadPotentialFullInv: factor is {L.shape}, the model has {n} unconstrained values")
        A = L @ L.T
        return {"mass": "dense", "dense_inverse": 0.5 * (A + A.T)}
    # QuadPotentialFullAdapt
    cov0 = np.asarray(potential._initial_cov, dtype=np.float64)
    if cov0.shape != (n, n):
        raise ValueError(f"QuadPotentialFullAdapt: initial covariance is {cov0.shape}, the model has {n} unconstrained values")
    if np.any(cov0 - np.diag(np.diag(cov0)) != 0.0):
        raise NotImplementedError("QuadPotentialFullAdapt with a non-diagonal initial covariance (the engine starts every chain "
                                  "from diag(var0); init_nuts passes the identity)")
    return {"mass": "dense_adapt", "var0": np.diag(cov0).copy(), "mean0": vec(potential._initial_mean, "initial_mean"),
            "mass_initial_weight": float(potential._initial_weight), "adaptation_window": int(potential.adaptation_window),
            "window_multiplier": float(potential.adaptation_window_multiplier), "update_window": int(potential._update_window)}
