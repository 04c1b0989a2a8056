"""Tail / CTA-granularity probe on the bench workload (Radon, 2048 chains, 1000+1000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
spec = models.radon(); cm = engine.CompiledModel(spec)
C = 2048
r = np.random.default_rng(1)
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
m0 = np.broadcast_to(q0.mean(0), q0.shape).copy()
for wpb, hot in ((7, 1), (7, 2), (7, 3), (4, 2)):
    os.environ["B200_NUTS_WPB"] = str(wpb); os.environ["B200_NUTS_HOT"] = str(hot)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=1000, draws=1000, mean0=m0, philox_seed=5, store_warmup=False, device_outputs=False, stats=False)
    ge = res.summary["grad_evals"] - 2000
    print(f"wpb={wpb} hot={hot}: {res.kernel_ms:.1f} ms, {ge.sum()/res.kernel_ms/1e3:.1f} M evals/s; per-chain evals mean {ge.mean():.0f} p50 {np.median(ge):.0f} p99 {np.quantile(ge,.99):.0f} max {ge.max()}")
