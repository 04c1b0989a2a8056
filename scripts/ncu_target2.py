"""ncu target: the TAIL regime -- one chain (warp) per SM, depth-6 trees, Radon."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
spec = models.radon(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
C = 148
os.environ["B200_NUTS_WPB"] = "1"
q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=20, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                  max_treedepth=6, early_max_treedepth=6, philox_seed=3)
print("nuts ms", res.kernel_ms, "evals", res.grad_evals, "per-warp us/eval", res.kernel_ms*1e3/(res.grad_evals/C))
