"""Radon: batched leapfrog throughput, solo-warp latency (1 chain per SM) and fixed-depth NUTS, per build variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib
spec = models.radon(); cm = engine.CompiledModel(spec)
r = np.random.default_rng(1)
tag = os.path.basename(os.environ.get("B200_LIB", "default"))
for C in (2048, 148):
    q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
    p0 = r.standard_normal((C, spec.n)); var = np.ones((C, spec.n))
    s = cm.leapfrog(q0, p0, var, 1e-4, 0)
    cm.leapfrog(s["q"], s["p"], var, 1e-4, 1000, grad=s["grad"])
    ms, _ = _lib.last_kernel_ms()
    print(f"[{tag}] leapfrog C={C}: {C*1000/ms/1e3:.1f} M evals/s; {ms*1e3/1000:.2f} us per step", flush=True)
C = 148
os.environ["B200_NUTS_WPB"] = "1"
q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=20, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                  max_treedepth=6, early_max_treedepth=6, philox_seed=3)
print(f"[{tag}] solo nuts depth 6: per-warp us/eval {res.kernel_ms*1e3/(res.grad_evals/C):.3f}", flush=True)
C = 2048
os.environ["B200_NUTS_WPB"] = "4"
q0 = spec.initial_point() + r.uniform(-0.1, 0.1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=0, draws=40, mass="diag", adapt_step_size=False, eps0=np.full(C, 1e-5),
                  max_treedepth=5, early_max_treedepth=5, philox_seed=3)
print(f"[{tag}] nuts fixed depth 5, 2048 chains: {res.grad_evals/res.kernel_ms/1e3:.1f} M evals/s", flush=True)
