"""Seams B1/B3: objects the reference's own step machinery accepts.

``B200LogpDlogp`` satisfies what ``GradientSharedStep`` / ``CpuLeapfrogIntegrator`` read from a
``ValueGradFunction`` (pymc/step_methods/arraystep.py:175-205, hmc/integration.py:42-66), so the UNMODIFIED
reference NUTS can run on the CUDA logp/grad one point at a time.  That proves correctness behind the
reference's own sampler; speed comes from the whole-run seam (``sample_b200_nuts``).
"""
from __future__ import annotations

import numpy as np

from .engine import CompiledModel


class B200LogpDlogp:
    def __init__(self, compiled: CompiledModel):
        self._cm = compiled
        self._raveled_inputs = True
        self.dtype = "float64"
        self._extra_vars_shared = {}
        self.trust_input = True

    def _pytensor_function(self, q):
        lp, g = self._cm.logp_dlogp(np.asarray(q, dtype=np.float64))
        return lp, g

    def set_extra_values(self, point):  # no non-gradient (shared) variables in the supported models
        pass

    def __call__(self, q):
        return self._pytensor_function(q)


class B200NUTS:
    """Seam B3: a step-method object with the ``BlockedStep`` / ``NUTS`` protocol that
    ``pm.sample(step=...)``'s chain loop drives (``_iter_sample``, pymc/sampling/mcmc.py:1503-1578):

        step.setup_chain(rng, tune, draws); step.tune = bool(tune); step.reset_tuning()
        for i in range(tune + draws):
            if i == tune: step.stop_tuning()
            point, stats = step.step(point)

    One chain per object, like the reference.  ``setup_chain`` announces ``tune`` and ``draws``, so the WHOLE chain runs
    on the device inside the first ``step`` call (one ``b200_nuts_run``) and the following calls hand its iterations out
    one by one.  Streams are derived as the reference derives them (``self.rng = rng``; the potential draws its momentum
    normals from ``rng.spawn(1)[0]``, hmc/base_hmc.py:300-302), so with ``momentum="numpy"`` a chain equals the chain the
    reference's own ``NUTS`` produces from the same generator, and ``rng`` is left where the reference would leave it.
    Defaults follow ``BaseHMC.__init__`` (hmc/base_hmc.py:82-169): ``QuadPotentialDiagAdapt(n, zeros, ones, 10)``.
    """

    name = "nuts"
    default_blocked = True
    # NUTS.stats_dtypes_shapes (hmc/nuts.py:110-130); "warning" carries None here
    stats_dtypes_shapes = {
        "depth": (np.int64, []), "step_size": (np.float64, []), "mean_tree_accept": (np.float64, []),
        "step_size_bar": (np.float64, []), "tree_size": (np.float64, []), "diverging": (bool, []),
        "divergences": (int, []), "energy_error": (np.float64, []), "energy": (np.float64, []),
        "max_energy_error": (np.float64, []), "model_logp": (np.float64, []), "process_time_diff": (np.float64, []),
        "perf_counter_diff": (np.float64, []), "perf_counter_start": (np.float64, []), "largest_eigval": (np.float64, []),
        "smallest_eigval": (np.float64, []), "index_in_trajectory": (np.int64, []), "reached_max_treedepth": (bool, []),
        "warning": (object, None),
    }

    def __init__(self, model, *, max_treedepth=10, early_max_treedepth=8, target_accept=0.8, step_scale=0.25, gamma=0.05,
                 k=0.75, t0=10, Emax=1000, adapt_step_size=True, potential_mean=None, potential_var=None,
                 potential_weight=10.0, momentum="numpy", rng=None):
        self._cm = model if hasattr(model, "nuts_run") else CompiledModel(model)
        self.spec = self._cm.spec
        self.vars = list(self.spec.vars)
        n = self.spec.n
        self._kw = dict(max_treedepth=max_treedepth, early_max_treedepth=early_max_treedepth, target_accept=target_accept,
                        step_scale=step_scale, gamma=gamma, k=k, t0=t0, Emax=Emax, adapt_step_size=adapt_step_size,
                        mass_initial_weight=potential_weight)
        self._mean0 = np.zeros(n) if potential_mean is None else np.asarray(potential_mean, dtype=np.float64).reshape(n)
        self._var0 = np.ones(n) if potential_var is None else np.asarray(potential_var, dtype=np.float64).reshape(n)
        if momentum not in ("numpy", "device"):
            raise ValueError("momentum must be 'numpy' or 'device'")
        self._momentum = momentum
        self.rng = np.random.default_rng(rng)
        self.tune = True
        self.iter_count = 0
        self.divergences = 0
        self._tune_n = self._draws_n = None
        self._res = None
        self._base, self._served, self._bad_at, self._resume_eps, self._pot_rng = 0, None, -1, None, None
        self._state = None

    # ---- BlockedStep protocol (step_methods/compound.py:132-250) -------------------------------------------------
    @staticmethod
    def competence(var, has_grad):
        return 3 if has_grad and np.issubdtype(np.dtype(getattr(var, "dtype", "float64")), np.floating) else 0

    def setup_chain(self, rng, tune: int, draws: int) -> None:
        self.rng = rng
        self._tune_n, self._draws_n = int(tune), int(draws)
        self._res = None
        self._base, self._served, self._bad_at, self._resume_eps, self._pot_rng = 0, None, -1, None, None

    def reset_tuning(self, start=None):
        self.iter_count = 0
        self.divergences = 0
        self.tune = True
        self._res = None
        self._base, self._served, self._bad_at, self._resume_eps = 0, None, -1, None

    def stop_tuning(self):
        if self._tune_n is not None and self.iter_count != self._tune_n and self._res is not None:
            raise RuntimeError(f"stop_tuning() at iteration {self.iter_count}, but setup_chain announced tune={self._tune_n}: "
                               "the chain already ran on the device with that schedule")
        self.tune = False

    def _ravel(self, point):
        return np.concatenate([np.ravel(np.asarray(point[v.name], dtype=np.float64)) for v in self.vars])

    def _run_chain(self, q0):
        """Launch the REMAINING schedule of the chain from q0 (iteration ``self.iter_count`` onwards)."""
        from . import rng as brng

        if self._tune_n is None:
            raise RuntimeError("setup_chain(rng, tune, draws) must be called before step(): the chain runs as one device launch")
        T = self._tune_n + self._draws_n - self.iter_count
        if T <= 0:
            raise RuntimeError("step() called more often than setup_chain announced (tune + draws)")
        # a chain whose tuning was stopped before this launch samples only
        tune = max(0, self._tune_n - self.iter_count) if self.tune else 0
        states = brng.pack_pcg64([self.rng])
        z, seed = None, 0
        if self._momentum == "numpy":
            if self._pot_rng is None:
                self._pot_rng = self.rng.spawn(1)[0]  # hmc/base_hmc.py:300-302
            z = brng.momentum_noise([self._pot_rng], T, self.spec.n)
        else:
            seed = int(self.rng.spawn(1)[0].integers(2**63))
        kw = dict(self._kw)
        eps0 = None
        if self._resume_eps is not None:  # re-launch mid-chain: continue from the step size adapted so far
            eps0 = np.array([self._resume_eps])
        from .engine import ChainState

        self._state = ChainState(1, self.spec.n) if hasattr(self._cm, "_h") else None  # engines that export chain state
        extra = {"save": self._state} if self._state is not None else {}
        self._res = self._cm.nuts_run(q0[None, :], states, tune=tune, draws=T - tune, z=z, philox_seed=seed, store_warmup=True,
                                      mass="diag_adapt", mean0=self._mean0[None, :], var0=self._var0[None, :], eps0=eps0, **kw,
                                      **extra)
        self._base = self.iter_count
        self._served = q0.copy()
        brng.unpack_pcg64(states, [self.rng])  # the host generator continues where the device stopped
        self._bad_at = int(self._res.summary["bad_energy_at"][0])

    def step(self, point):
        import time

        t0 = time.perf_counter()
        q_in = self._ravel(point)
        if self._res is not None and not np.array_equal(q_in, self._served):
            # BaseHMC.astep re-reads q0 on every call (hmc/base_hmc.py:196-202).  The chain was run ahead from the point
            # of the previous call; a driver that changed the point in between (CompoundStep with other step methods,
            # hand-written loops) gets a fresh launch of the remaining schedule from ITS point, never stale draws.
            import warnings

            warnings.warn("B200NUTS.step: the incoming point differs from the previously returned draw; re-launching the "
                          "remaining schedule from it (mass-matrix adaptation restarts, the step size carries over)",
                          RuntimeWarning, stacklevel=2)
            i_prev = self.iter_count - self._base - 1
            if i_prev >= 0:
                key = "step_size" if self.tune else "step_size_bar"
                self._resume_eps = float(self._res.stats[key][0, i_prev])
            self._res = None
        if self._res is None:
            self._run_chain(q_in)
        i = self.iter_count - self._base
        if i >= np.asarray(self._res.draws).shape[1]:
            raise RuntimeError("step() called more often than setup_chain announced (tune + draws)")
        if self._bad_at >= 0 and i >= self._bad_at:
            # raised at the iteration where the reference raises it (base_hmc.py:205-224), after the valid draws before it
            from .sampling import SamplingError

            raise SamplingError(f"Bad initial energy at iteration {self.iter_count}: check any log probabilities that are "
                                "inf or nan (model.debug())")
        q = np.asarray(self._res.draws[0, i])
        self._served = q.copy()
        st = {k: v[0, i] for k, v in self._res.stats.items()}
        diverging = bool(st["diverging"])
        if not self.tune:
            self.divergences += diverging
        self.iter_count += 1
        new_point = dict(point)
        for v in self.vars:
            new_point[v.name] = q[v.offset : v.offset + v.size].reshape(np.shape(point[v.name])).copy()
        t1 = time.perf_counter()
        stats = {
            "depth": int(st["depth"]), "step_size": float(st["step_size"]),
            "mean_tree_accept": float(st["mean_tree_accept"]), "step_size_bar": float(st["step_size_bar"]),
            "tree_size": float(st["tree_size"]), "diverging": diverging, "divergences": self.divergences,
            "energy_error": float(st["energy_error"]), "energy": float(st["energy"]),
            "max_energy_error": float(st["max_energy_error"]), "model_logp": float(st["model_logp"]),
            "process_time_diff": t1 - t0, "perf_counter_diff": t1 - t0, "perf_counter_start": t0,
            "largest_eigval": np.nan, "smallest_eigval": np.nan,  # QuadPotential.stats() (quadpotential.py:177-178)
            "index_in_trajectory": int(st["index_in_trajectory"]), "reached_max_treedepth": bool(st["reached_max_treedepth"]),
            "warning": None,
        }
        return new_point, [stats]

    @property
    def sampling_state(self):
        """What ``BaseHMC.sampling_state`` carries (hmc/base_hmc.py:61-71): counters, the step-size adaptation state
        (``StepSizeState``, step_sizes.py:26-38) and the potential's state (``QuadPotentialDiagAdaptState`` with its two
        ``WeightedVarianceState``, quadpotential.py:189-208, :396-403).  The chain runs ahead on the device, so the
        adaptation state is the one at the END of the launched schedule (exported by the kernel, ``b200_chain_state``)."""
        out = {"iter_count": self.iter_count, "tune": self.tune, "divergences": self.divergences}
        st = getattr(self, "_state", None)
        if st is not None and self._res is not None:
            out["step_adapt"] = {"log_step": float(st.log_step[0]), "log_bar": float(st.log_bar[0]), "hbar": float(st.hbar[0]),
                                 "count": int(st.da_count[0])}
            out["potential"] = {"_var": st.var[0].copy(), "_n_samples": int(st.n_samples[0]),
                                "adaptation_window": int(st.window[0]),
                                "_foreground_var": {"n_samples": float(st.fg_n[0]), "mean": st.fg_mean[0].copy(), "raw_var": st.fg_m2[0].copy()},
                                "_background_var": {"n_samples": float(st.bg_n[0]), "mean": st.bg_mean[0].copy(), "raw_var": st.bg_m2[0].copy()}}
            out["launched_through_iteration"] = int(st.iter_count) + self._base
        return out


def from_pymc(model):
    """Lower a ``pm.Model`` to a ``pymc_b200.ir.ModelIR``: see ``pymc_b200.frontend``."""
    from .frontend import from_pymc as _from_pymc

    return _from_pymc(model)
