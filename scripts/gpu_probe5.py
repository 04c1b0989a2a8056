"""Stochastic volatility (config 4) throughput probe: 256 chains, short run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib
spec = models.stochvol(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
for C in (148, 256):
    q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
    p0 = r.standard_normal((C, spec.n)); var = np.ones((C, spec.n))
    s = cm.leapfrog(q0, p0, var, 1e-4, 0)
    cm.leapfrog(s["q"], s["p"], var, 1e-4, 200, grad=s["grad"])
    ms, _ = _lib.last_kernel_ms()
    print(f"[stochvol leapfrog] C={C}: {C*200/ms/1e3:.2f} M evals/s")
    sr, pr, _ = brng.chain_generators(123, C)
    for hot in (0, 1, 2):
        os.environ["B200_NUTS_HOT"] = str(hot)
        res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=10, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                          max_treedepth=6, early_max_treedepth=6, philox_seed=3, stats=True)
        ge = res.grad_evals
        print(f"[stochvol nuts depth 6] C={C} hot={hot}: {res.kernel_ms:.1f} ms {ge/res.kernel_ms/1e3:.2f} M evals/s")
C = 256
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(5, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=200, draws=50, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(), philox_seed=5)
ge = res.summary["grad_evals"] - 250
print(f"[stochvol adaptive 200+50] {res.kernel_ms:.0f} ms, {ge.sum()/res.kernel_ms/1e3:.2f} M evals/s, mean depth {res.stats['depth'].mean():.2f}, div {res.stats['diverging'].mean():.4f}, bad {np.sum(res.summary['bad_energy_at']>=0)}")
