"""Short single-GPU target for ncu: one leapfrog launch and one NUTS launch on the Radon config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
spec = models.radon(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
C = int(os.environ.get("NCU_C", 2048))
q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
p0 = r.standard_normal((C, spec.n)); var = np.ones((C, spec.n))
s = cm.leapfrog(q0, p0, var, 1e-4, 0)
cm.leapfrog(s["q"], s["p"], var, 1e-4, 100, grad=s["grad"])
sr, pr, _ = brng.chain_generators(123, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=10, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                  max_treedepth=5, early_max_treedepth=5, philox_seed=3)
print("nuts ms", res.kernel_ms, "evals", res.grad_evals)
