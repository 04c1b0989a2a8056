"""Full bench workload (Radon, 2048 chains, 1000+1000): per-chain work distribution and kernel time per build variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng
spec = models.radon(); cm = engine.CompiledModel(spec)
C = int(os.environ.get("CHAINS", 2048))
r = np.random.default_rng(1)
q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
sr, pr, _ = brng.chain_generators(123, C)
m0 = np.broadcast_to(q0.mean(0), q0.shape).copy()
tag = os.path.basename(os.environ.get("B200_LIB", "default"))
cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(4, 2)]
for wpb, hot in cfgs:
    os.environ["B200_NUTS_WPB"] = str(wpb); os.environ["B200_NUTS_HOT"] = str(hot)
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=1000, draws=1000, mean0=m0, philox_seed=5, store_warmup=True, stats=True)
    ge = res.summary["grad_evals"] - 2000
    ts = res.stats["tree_size"]
    print(f"[{tag}] wpb={wpb} hot={hot}: {res.kernel_ms:.1f} ms, {ge.sum()/res.kernel_ms/1e3:.1f} M evals/s; per-chain evals mean {ge.mean():.0f} p50 {np.median(ge):.0f} p99 {np.quantile(ge,.99):.0f} max {ge.max()}"
          f" | warmup evals mean {ts[:, :1000].sum(1).mean():.0f} max {ts[:, :1000].sum(1).max()} | sampling mean {ts[:, 1000:].sum(1).mean():.0f} max {ts[:, 1000:].sum(1).max()}"
          f" | solo-us/eval if straggler-bound {res.kernel_ms*1e3/ge.max():.2f}", flush=True)
