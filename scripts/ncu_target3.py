"""ncu targets for the lock-step (GEMM-shaped) engine: a few batched leapfrogs at full size.
usage: ncu_target3.py logistic|logistic_tc|mvgauss|mvgauss_tc [tune draws]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
which = sys.argv[1]
tune, draws = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3, 2)
r = np.random.default_rng(1)
if which in ("logistic", "logistic_tc"):
    spec = models.logistic(); C, kw = 512, {}
else:
    spec = models.mvgauss(cache_dir="/dev/shm/b200_cache"); C, kw = 256, dict(mass="dense")
cm = engine.CompiledModel(spec)
if which in ("logistic_tc", "mvgauss_tc"):
    cm.set_precision("tc_fp16x2")
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(5, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=tune, draws=draws, philox_seed=5, **kw)
print(which, "kernel ms", res.kernel_ms, "launches", res.launches, "evals", int(res.stats["tree_size"].sum()))
