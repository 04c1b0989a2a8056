"""Python face of the C-ABI engine: compiled model handle + the three hot-path entry points.

``CompiledModel`` is the object the reference gets from ``Model.logp_dlogp_function(ravel_inputs=True)``
(pymc/model/core.py:464-529): a handle whose call evaluates ``q -> (logp, dlogp)``, here batched over
many points and backed by the model's hand-written CUDA function instead of a PyTensor C thunk.
PyTorch is used only for device allocations when the caller wants results to stay in HBM.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import B200Error  # noqa: F401  (re-export)
from .models import ModelSpec


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ir_spec(ir) -> ModelSpec:
    """The ModelSpec face (layout, transforms, initial point) of a model that runs on the generic IR device function."""
    from . import models

    vars_, n = models._layout([(v.name, v.rv_name, v.size, v.transform, v.bounds) for v in ir.vars])
    return ModelSpec(models.KIND_IR, n, vars_, data={}, meta={"initial_point": ir.initial_point(), "ir": ir})


class ChainState:
    """Per-chain sampler state between calls (b200_chain_state; BaseHMC.sampling_state of the reference, hmc/base_hmc.py:61-71
    with step_sizes.py:26-38 and quadpotential.py:189-208, :396-403): host NumPy arrays, one row per chain."""

    def __init__(self, chains: int, n: int):
        self.chains, self.n = chains, n
        for name, dt, vec in _lib.STATE_FIELDS:
            setattr(self, name, np.zeros((chains, n) if vec else (chains,), dtype=dt))
        self.iter_count = 0  # iterations completed (the `iter_begin` of the next call)

    def c_struct(self):
        st = _lib.ChainStateC()
        for name, _, _ in _lib.STATE_FIELDS:
            setattr(st, name, getattr(self, name).ctypes.data)
        return st

    def as_dict(self):
        return {name: getattr(self, name).copy() for name, _, _ in _lib.STATE_FIELDS} | {"iter_count": self.iter_count}


class CompiledModel:
    """Observed data resident in HBM + dispatch to the model's fused logp/grad device function.

    ``spec`` is either a ``pymc_b200.ir.ModelIR`` (the general interface: any model of the closed factor set; routed to
    a hand-specialised kernel when ``ir.specialise`` recognises its shape, otherwise evaluated by the generic IR device
    function) or a ``models.ModelSpec`` naming one of the hand-written kernels directly (the GEMM-shaped configs)."""

    def __init__(self, spec, device: int | None = None, specialise: bool | str = True):
        from . import ir as _ir

        self.ir = None
        if isinstance(spec, _ir.ModelIR):
            self.ir = spec
            fast = _ir.specialise(spec, extended=(specialise == "all")) if specialise else None
            spec = fast if fast is not None else _ir_spec(spec)
        self.spec = spec
        self.n = spec.n
        self._lib = _lib.load()
        _lib.require_gpu()
        if device is not None:
            _lib.check(self._lib.b200_set_device(device))
        self.device = 0 if device is None else int(device)
        d = _lib.ModelDesc()
        d.kind, d.n = spec.kind, spec.n
        keep = []

        def hold(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data

        data = spec.data
        if spec.name == "eight_schools":
            d.n_obs = len(data["y"])
            d.y, d.aux = hold(data["y"], np.float64), hold(data["sigma"], np.float64)
        elif spec.name == "radon":
            d.n_obs, d.n_groups = len(data["y"]), spec.meta["n_counties"]
            d.x, d.y = hold(data["floor"], np.float64), hold(data["y"], np.float64)
            d.idx = hold(data["county_idx"], np.int32)
        elif spec.name == "logistic":
            d.n_obs = data["X"].shape[0]
            d.x, d.y_u8 = hold(data["X"], np.float64), hold(data["y"], np.uint8)
        elif spec.name == "stochvol":
            d.n_obs = len(data["y"])
            d.y = hold(data["y"], np.float64)
        elif spec.name == "mvgauss":
            import scipy.linalg as sl

            L = np.ascontiguousarray(data["L"], dtype=np.float64)
            LinvT = data["LinvT"] if "LinvT" in data else sl.solve_triangular(L, np.eye(L.shape[0]), lower=True).T
            d.x, d.aux = hold(data["prec"], np.float64), hold(data["cov"], np.float64)
            d.m1, d.m2 = hold(LinvT, np.float64), hold(L, np.float64)
            d.scalar0 = spec.meta["logdet_L"]
        elif spec.name == "ir":
            c_ir, keep_ir = _lib.build_ir(_ir.lower(self.ir))
            keep += keep_ir + [c_ir]
            d.ir = C.addressof(c_ir)
        handle = C.c_void_p()
        _lib.check(self._lib.b200_model_create(C.byref(d), C.byref(handle)))
        self._h = handle
        # backward transforms per element, for draws recorded in constrained space (nuts_run(constrain=True))
        kind, lo, hi = np.zeros(spec.n, dtype=np.int8), np.zeros(spec.n), np.ones(spec.n)
        for v in spec.vars:
            sl = slice(v.offset, v.offset + v.size)
            if v.transform == "log":
                kind[sl] = 1
            elif v.transform == "interval":
                kind[sl], lo[sl], hi[sl] = 2, v.bounds[0], v.bounds[1]
        self.supports_constrain = hasattr(self._lib, "b200_model_set_transforms")
        if self.supports_constrain:
            _lib.check(self._lib.b200_model_set_transforms(self._h, kind.ctypes.data, lo.ctypes.data, hi.ctypes.data))

    def pointwise_loglik(self, draws, lik: int = 0):
        """log p(y_i | draw) for every draw: ``draws[..., n]`` (unconstrained) -> ``[..., N]`` (IR models on the generic
        device function).  The `log_likelihood` group pm.compute_log_likelihood builds (pymc/stats/log_density.py:31-77)."""
        if self.spec.name != "ir":
            raise NotImplementedError("pointwise_loglik needs a model compiled from ModelIR with specialise=False")
        q = _f64(draws)
        lead = q.shape[:-1]
        q2 = q.reshape(-1, self.n)
        N = len(self.ir.likelihoods[lik].y)
        out = np.empty((q2.shape[0], N))
        _lib.check(self._lib.b200_pointwise_loglik(self._h, int(lik), q2.ctypes.data, q2.shape[0], out.ctypes.data, _lib.MEM_HOST, None))
        return out.reshape(lead + (N,))

    def set_dense_mass(self, cov=None, *, inverse=None) -> None:
        """Fixed dense mass matrix for ``nuts_run(mass="dense")``: ``cov`` -> QuadPotentialFull(cov) (quadpotential.py:680-725),
        ``inverse=A`` -> QuadPotentialFullInv(A) (:633-677).  The Cholesky factors are computed here (SciPy)."""
        import scipy.linalg as sl

        n = self.n
        if (cov is None) == (inverse is None):
            raise ValueError("pass exactly one of cov= or inverse=")
        eye = np.eye(n)
        if cov is not None:
            S = np.ascontiguousarray(cov, dtype=np.float64).reshape(n, n)
            L = sl.cholesky(S, lower=True)
            mp0 = sl.solve_triangular(L, eye, lower=True).T.copy()  # L^-T: p0 = solve_triangular(L.T, z)
            mv0 = L
        else:
            A = np.ascontiguousarray(inverse, dtype=np.float64).reshape(n, n)
            LA = sl.cholesky(A, lower=True)
            S = sl.cho_solve((LA, True), eye)                          # velocity = cho_solve(L, x) = A^-1 x
            S = 0.5 * (S + S.T)
            mp0 = LA                                                   # random() = L . normal
            mv0 = sl.solve_triangular(LA, eye, lower=True).T.copy()    # v0 = A^-1 L z = L^-T z
        S, mp0, mv0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (S, mp0, mv0))
        _lib.check(self._lib.b200_model_set_dense_mass(self._h, S.ctypes.data, mp0.ctypes.data, mv0.ctypes.data))

    def set_precision(self, mode: str) -> None:
        """"fp64" (parity mode, default) or "tc_fp16x2": the dense contractions of the logistic GLM on the tcgen05 tensor
        cores with split-fp16 operands (performance mode; gradient ~1e-7 relative, see include/b200nuts.h)."""
        code = {"fp64": _lib.PRECISION_FP64, "tc_fp16x2": _lib.PRECISION_TC_FP16X2}[mode]
        _lib.check(self._lib.b200_model_set_precision(self._h, code))
        self.precision = mode

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.b200_model_destroy(h)

    # ------------------------------------------------------------------------------------------
    # a1: batched ValueGradFunction._pytensor_function
    # ------------------------------------------------------------------------------------------
    def logp_dlogp(self, q):
        """q[C, n] (or [n]) -> (logp[C], grad[C, n]); NumPy in, NumPy out."""
        q = _f64(q)
        single = q.ndim == 1
        q2 = q.reshape(-1, self.n)
        Cn = q2.shape[0]
        logp = np.empty(Cn)
        grad = np.empty((Cn, self.n))
        _lib.check(self._lib.b200_logp_dlogp(self._h, q2.ctypes.data, Cn, logp.ctypes.data, grad.ctypes.data, _lib.MEM_HOST, None))
        return (float(logp[0]), grad[0]) if single else (logp, grad)

    # ------------------------------------------------------------------------------------------
    # a2/a3: CpuLeapfrogIntegrator.compute_state / .step, batched, diagonal potential
    # ------------------------------------------------------------------------------------------
    def leapfrog(self, q, p, var, eps, n_steps, *, grad=None, idx=None):
        """Advance C states by ``n_steps`` leapfrogs of signed size ``eps[C]`` (n_steps=0: start state only).

        Returns dict(q, p, v, grad, energy, logp, idx) -- the fields of integration.State."""
        q, p, var = (np.array(_f64(x).reshape(-1, self.n)) for x in (q, p, var))
        Cn = q.shape[0]
        eps = np.broadcast_to(_f64(eps), (Cn,)).copy()
        if n_steps > 0 and grad is None:
            start = self.leapfrog(q, p, var, eps, 0)
            grad = start["grad"]
        g = np.zeros((Cn, self.n)) if grad is None else np.array(_f64(grad).reshape(Cn, self.n))
        v = np.empty((Cn, self.n))
        energy, logp = np.empty(Cn), np.empty(Cn)
        ix = np.zeros(Cn, dtype=np.int64) if idx is None else np.array(idx, dtype=np.int64).reshape(Cn)
        _lib.check(
            self._lib.b200_leapfrog(
                self._h, var.ctypes.data, eps.ctypes.data, int(n_steps), Cn, q.ctypes.data, p.ctypes.data,
                v.ctypes.data, g.ctypes.data, energy.ctypes.data, logp.ctypes.data, ix.ctypes.data, _lib.MEM_HOST, None,
            )
        )
        return dict(q=q, p=p, v=v, grad=g, energy=energy, logp=logp, idx=ix)

    # ------------------------------------------------------------------------------------------
    # a12+a13+a16: the whole multi-chain NUTS run
    # ------------------------------------------------------------------------------------------
    def nuts_run(self, q0, rng_states, *, tune, draws, var0=None, mean0=None, eps0=None, z=None, store_warmup=True,
                 mass="diag_adapt", adapt_step_size=True, step_scale=0.25, target_accept=0.8, gamma=0.05,
                 k=0.75, t0=10.0, Emax=1000.0, max_treedepth=10, early_max_treedepth=8,
                 mass_initial_weight=10.0, adaptation_window=101, discard_window=50, philox_seed=0,
                 device_outputs=False, stats=True, chain_offset=0, pinned_outputs=False, reuse_outputs=False,
                 sampler="nuts", path_length=2.0, max_steps=1024, mass_alpha=0.02, stop_adaptation=None, constrain=False,
                 iter_begin=0, iter_count=None, resume=None, save=None, window_multiplier=None, update_window=1):
        """Run C chains for tune+draws NUTS iterations inside one persistent kernel.

        ``sampler="hmc"`` runs HamiltonianMC (hmc/hmc.py: ``path_length``, ``max_steps``; pass ``target_accept=0.65`` for
        its default); ``mass="diag_adapt_grad"`` is init="jitter+adapt_diag_grad" (``mass_alpha``, ``stop_adaptation``);
        ``mass="dense_adapt"`` is QuadPotentialFullAdapt (init="adapt_full"): a dense covariance per chain estimated from
        the tuning draws (``adaptation_window``, ``window_multiplier`` default 2, ``update_window``); initial covariance
        diag(``var0``), initial mean ``mean0``, weight ``mass_initial_weight``; ``summary["final_cov"]`` holds the result.
        ``iter_begin`` / ``iter_count`` run a slice of the ``tune + draws`` schedule; ``resume`` / ``save`` are ``ChainState``
        objects (host arrays) carrying the chains between calls: a run split into slices is bit-identical to the
        uninterrupted run (persistent engine; outputs and ``z`` then cover the slice only).
        ``rng_states``: structured array (``_lib.PCG64_DTYPE``) of the chains' NumPy PCG64 step streams
        (see ``pymc_b200.rng``); updated in place.  ``z``: optional momentum noise [C, tune+draws, n]
        (NumPy ``Generator.normal`` for draw-parity with the reference); otherwise generated on device.
        ``device_outputs=True`` keeps draws/stats as torch CUDA tensors (no D2H copy).
        ``pinned_outputs=True`` returns host arrays backed by a per-model pool of pinned (page-locked) buffers
        that is REUSED by the next call of the same shape (fast D2H; copy what you want to keep);
        ``reuse_outputs=True`` does the same for device outputs (no allocator call inside a timed loop).
        Both are opt-in: the default returns fresh arrays the caller owns."""
        is_torch = lambda a: a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")  # noqa: E731
        if not is_torch(q0):
            q0 = _f64(q0).reshape(-1, self.n)
        Cn = q0.shape[0]
        n_iter = (tune + draws - iter_begin) if iter_count is None else int(iter_count)
        Ttot = n_iter  # iterations of THIS call (z and the outputs cover exactly these)
        rec_lo = iter_begin if store_warmup else max(tune, iter_begin)
        T = max(0, iter_begin + n_iter - rec_lo)
        cfg = _lib.NutsCfg()
        cfg.chains, cfg.tune, cfg.draws = Cn, int(tune), int(draws)
        cfg.max_treedepth, cfg.early_max_treedepth = int(max_treedepth), int(early_max_treedepth)
        cfg.adapt_step_size = int(bool(adapt_step_size))
        cfg.mass_kind = {"diag": _lib.MASS_DIAG, "diag_adapt": _lib.MASS_DIAG_ADAPT, "dense": _lib.MASS_DENSE,
                         "diag_adapt_grad": _lib.MASS_DIAG_ADAPT_GRAD, "dense_adapt": _lib.MASS_DENSE_ADAPT}[mass]
        cfg.mass_update_window = int(update_window)
        cfg.adaptation_window_multiplier = 0.0 if window_multiplier is None else float(window_multiplier)
        cfg.sampler = {"nuts": _lib.SAMPLER_NUTS, "hmc": _lib.SAMPLER_HMC}[sampler]
        cfg.path_length, cfg.max_steps = float(path_length), int(max_steps)
        cfg.mass_alpha = float(mass_alpha)
        cfg.stop_adaptation = -1 if stop_adaptation is None else int(stop_adaptation)
        cfg.iter_begin, cfg.iter_count = int(iter_begin), int(n_iter)
        keep_state = []
        if resume is not None or save is not None:
            if device_outputs:
                raise ValueError("chain-state import/export uses host arrays: call with device_outputs=False")
            for obj, name in ((resume, "resume"), (save, "save")):
                if obj is not None:
                    if obj.chains != Cn or obj.n != self.n:
                        raise ValueError(f"{name}: ChainState is for {obj.chains} chains x {obj.n}, this call has {Cn} x {self.n}")
                    cs = obj.c_struct()
                    keep_state.append(cs)
                    setattr(cfg, name, C.addressof(cs))
        cfg.constrain_draws = int(bool(constrain))  # draws come back in CONSTRAINED space (transform fused into the record step)
        cfg.momentum_source = _lib.MOMENTUM_DEVICE_PHILOX if z is None else _lib.MOMENTUM_HOST_BUFFER
        cfg.store_warmup = int(bool(store_warmup))
        cfg.chain_offset = int(chain_offset)
        cfg.step_scale, cfg.target_accept, cfg.gamma, cfg.k, cfg.t0, cfg.Emax = (
            float(step_scale), float(target_accept), float(gamma), float(k), float(t0), float(Emax))
        cfg.mass_initial_weight = float(mass_initial_weight)
        cfg.adaptation_window, cfg.discard_window = int(adaptation_window), int(discard_window)
        cfg.philox_seed = int(philox_seed) & 0xFFFFFFFFFFFFFFFF
        if rng_states.dtype != _lib.PCG64_DTYPE or rng_states.shape != (Cn,):
            raise ValueError("rng_states must be a (chains,) array of _lib.PCG64_DTYPE")

        tdt = None
        if device_outputs or pinned_outputs or reuse_outputs:
            import torch

            tdt = {np.int32: torch.int32, np.uint8: torch.uint8, np.float64: torch.float64, np.int64: torch.int64}
        # Output buffers.  The kernels write EVERY element of draws / stats / summary (iterations a frozen chain never ran
        # get NaN / zero sentinels in-kernel), so buffers are allocated uninitialised and may be reused between calls:
        # `reuse_outputs` (device) and `pinned_outputs` (host) keep one set per (shape, dtype) in the model handle, which
        # takes torch's allocator and cudaHostAlloc out of a timed loop.
        pool = self.__dict__.setdefault("_out_pool", {}) if (reuse_outputs or pinned_outputs) else None
        counter = [0]

        def pooled(make, shape, dt, space):
            if pool is None:
                return make()
            counter[0] += 1
            key = (space, counter[0], tuple(shape), np.dtype(dt).str)
            if key not in pool:
                pool[key] = make()
            return pool[key]

        if device_outputs:
            dev = torch.device("cuda", torch.cuda.current_device())
            mem = _lib.MEM_DEVICE

            def to_dev(a):  # NumPy -> HBM; torch CUDA tensors are used in place (inputs already resident)
                if a is None:
                    return None
                if is_torch(a):
                    return a.to(device=dev, dtype=torch.float64).contiguous()
                return torch.as_tensor(_f64(a), device=dev)

            q0_b, var0_b, mean0_b, z_b = to_dev(q0), to_dev(var0), to_dev(mean0), to_dev(z)
            eps0_b = to_dev(eps0)
            rng_b = torch.as_tensor(rng_states.view(np.uint64).reshape(Cn, 4).view(np.int64), device=dev)
            mk = lambda shape, dt: pooled(lambda: torch.empty(shape, dtype=tdt[dt], device=dev), shape, dt, "dev")  # noqa: E731
        else:
            for a in (q0, var0, mean0, eps0, z):
                if is_torch(a) and a.is_cuda:
                    raise ValueError("device_outputs=False takes host (NumPy) inputs; pass device_outputs=True for CUDA tensors")
            if is_torch(q0):
                q0 = _f64(q0.numpy()).reshape(-1, self.n)
            mem = _lib.MEM_HOST
            q0_b = q0
            var0_b = None if var0 is None else _f64(var0).reshape(Cn, self.n)
            mean0_b = None if mean0 is None else _f64(mean0).reshape(Cn, self.n)
            z_b = None if z is None else _f64(z).reshape(Cn, Ttot, self.n)
            eps0_b = None if eps0 is None else np.broadcast_to(_f64(eps0), (Cn,)).copy()
            rng_b = rng_states
            if pinned_outputs:
                def pinned(shape, dt):  # page-locked memory on the GPU's own NUMA node (parallel.near_gpu)
                    from . import parallel

                    with parallel.near_gpu(self.device):
                        return torch.empty(tuple(shape), dtype=tdt[dt], pin_memory=True).numpy()

                mk = lambda shape, dt: pooled(lambda: pinned(shape, dt), shape, dt, "pin")  # noqa: E731
            else:
                mk = lambda shape, dt: np.empty(shape, dtype=dt)  # noqa: E731
        draws_b = mk((Cn, T, self.n), np.float64)

        st, st_arr = _lib.Stats(), {}
        if stats:
            for name, dt in _lib.STAT_FIELDS:
                st_arr[name] = mk((Cn, T), dt)
                setattr(st, name, _lib.ptr(st_arr[name]))
        sm, sm_arr = _lib.ChainSummary(), {}
        for name, dt in _lib.SUMMARY_FIELDS:
            if name == "final_cov":
                if mass != "dense_adapt":
                    continue
                sm_arr[name] = mk((Cn, self.n, self.n), dt)
            else:
                sm_arr[name] = mk((Cn, self.n) if name == "final_var" else (Cn,), dt)
            setattr(sm, name, _lib.ptr(sm_arr[name]))

        _lib.check(
            self._lib.b200_nuts_run(
                self._h, C.byref(cfg), _lib.ptr(q0_b), _lib.ptr(var0_b), _lib.ptr(mean0_b), _lib.ptr(eps0_b), _lib.ptr(rng_b),
                _lib.ptr(z_b), _lib.ptr(draws_b), C.byref(st), C.byref(sm), mem, None,
            )
        )
        if device_outputs:
            rng_states[:] = rng_b.cpu().numpy().view(np.uint64).reshape(Cn, 4).view(_lib.PCG64_DTYPE).reshape(Cn)
        if save is not None:
            save.iter_count = iter_begin + n_iter
        ms, launches = _lib.last_kernel_ms()
        return NutsResult(draws=draws_b, stats=st_arr, summary=sm_arr, kernel_ms=ms, launches=launches,
                          tune=tune, n_draws=draws, store_warmup=store_warmup)


@dataclass
class NutsResult:
    draws: object  # [C, T, n] unconstrained positions (NumPy or torch CUDA)
    stats: dict
    summary: dict
    kernel_ms: float
    launches: int
    tune: int
    n_draws: int
    store_warmup: bool

    @property
    def grad_evals(self) -> int:
        """Leapfrog gradient evaluations (sum of tree_size, the reference's own stat, hmc/nuts.py:485)."""
        ts = self.stats["tree_size"]
        return int(ts.sum())
