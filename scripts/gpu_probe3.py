"""A/B probe of build variants (B200_LIB) on the Radon config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib
spec = models.radon(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
C = 2048
q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
p0 = r.standard_normal((C, spec.n)); var = np.ones((C, spec.n))
s = cm.leapfrog(q0, p0, var, 1e-4, 0)
cm.leapfrog(s["q"], s["p"], var, 1e-4, 1000, grad=s["grad"])
ms, _ = _lib.last_kernel_ms()
tag = os.environ.get("B200_LIB", "default")[-14:]
print(f"[{tag}] leapfrog: {C*1000/ms/1e3:.1f} M evals/s")
sr, pr, _ = brng.chain_generators(123, C)
for wpb, hot in ((4, 0), (4, 1), (4, 2), (8, 1)):
    os.environ["B200_NUTS_WPB"] = str(wpb); os.environ["B200_NUTS_HOT"] = str(hot)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=40, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                      max_treedepth=5, early_max_treedepth=5, philox_seed=3)
    ge = res.grad_evals
    print(f"[{tag}] nuts fixed depth 5 wpb={wpb} hot={hot}: {res.kernel_ms:.1f} ms {ge/res.kernel_ms/1e3:.1f} M evals/s")
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=300, draws=100, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(), philox_seed=5)
pc = res.stats["tree_size"].sum(1)
print(f"[{tag}] nuts adaptive 300+100: {res.kernel_ms:.1f} ms, {res.grad_evals/res.kernel_ms/1e3:.1f} M evals/s; per-chain evals mean {pc.mean():.0f} max {pc.max()}")
