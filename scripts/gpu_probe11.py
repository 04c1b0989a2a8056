"""e2e anatomy on the bench workload: wall time vs kernel time per call, outputs written directly to pinned host memory
(default) or staged through a device buffer + D2H copy (B200_NO_DIRECT_HOST_WRITES=1), and pure device outputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pymc_b200 import models, engine, rng as brng
spec = models.radon(); cm = engine.CompiledModel(spec)
C = 2048
r = np.random.default_rng(1)
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
m0 = np.broadcast_to(q0.mean(0), q0.shape).copy()
def run(tag, **kw):
    for rep in range(3):
        st = brng.pack_pcg64(sr)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = cm.nuts_run(q0, st, tune=1000, draws=1000, mean0=m0, philox_seed=5, store_warmup=False, **kw)
        torch.cuda.synchronize(); w = (time.perf_counter() - t0) * 1e3
        print(f"[{tag}] rep {rep}: wall {w:.1f} ms, kernel {res.kernel_ms:.1f} ms, host-side {w - res.kernel_ms:.1f} ms", flush=True)
run("device outputs", device_outputs=True)
run("pinned, direct writes", pinned_outputs=True)
os.environ["B200_NO_DIRECT_HOST_WRITES"] = "1"
run("pinned, staged + D2H", pinned_outputs=True)
run("pinned, staged, no stats", pinned_outputs=True, stats=False)
