"""Perf decomposition probe: pure leapfrog kernel vs fixed-depth NUTS vs adaptive NUTS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib

spec = models.radon(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
for C in (1184, 2048, 4096):
    q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
    p0 = r.standard_normal((C, spec.n)); var = np.ones((C, spec.n))
    s = cm.leapfrog(q0, p0, var, 1e-4, 0)
    for steps in (200, 1000):
        cm.leapfrog(s["q"], s["p"], var, 1e-4, steps, grad=s["grad"])
        ms, _ = _lib.last_kernel_ms()
        print(f"[leapfrog] C={C} steps={steps}: {ms:.2f} ms -> {C*steps/ms/1e3:.1f} M evals/s")
for C in (1184, 2048):
    q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
    sr, pr, _ = brng.chain_generators(123, C)
    for depth in (3, 5, 7):
        st = brng.pack_pcg64(sr)
        res = cm.nuts_run(q0, st, tune=0, draws=40, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                          max_treedepth=depth, early_max_treedepth=depth, philox_seed=3)
        ge = res.grad_evals
        print(f"[nuts fixed depth {depth}] C={C}: {res.kernel_ms:.1f} ms, evals {ge} ({ge/C/40:.1f}/draw), {ge/res.kernel_ms/1e3:.1f} M evals/s")
C = 2048
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=300, draws=100, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(), philox_seed=5)
per_chain = res.stats["tree_size"].sum(1)
print(f"[nuts adaptive] {res.kernel_ms:.1f} ms; evals per chain: mean {per_chain.mean():.0f} min {per_chain.min()} max {per_chain.max()} p99 {np.quantile(per_chain,0.99):.0f}; {res.grad_evals/res.kernel_ms/1e3:.1f} M evals/s")
w = res.stats["tree_size"][:, 300:]
print("   post-warmup evals/draw mean", w.mean(), "depth mean", res.stats["depth"][:,300:].mean())
