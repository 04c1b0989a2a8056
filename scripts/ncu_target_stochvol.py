"""ncu target for BASELINE config 4: nuts_warp_kernel<StochVolModel,12,8> (chain = CTA), 256 chains, shortened run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
spec = models.stochvol(); cm = engine.CompiledModel(spec)
C = int(os.environ.get("NCU_C", 256))
r = np.random.default_rng(1)
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(5, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=60, draws=20, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(), philox_seed=5)
ge = res.summary["grad_evals"]
print("nuts ms", res.kernel_ms, "grad_evals_incl_start", int(ge.sum()), "leapfrog_evals", int(res.stats["tree_size"].sum()),
      "mean depth", float(res.stats["depth"].mean()))
