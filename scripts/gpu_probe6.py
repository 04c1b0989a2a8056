"""Lock-step engine probe: logistic (config 3) and dense Gaussian (config 5) at reduced + full size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib
r = np.random.default_rng(1)
which = sys.argv[1:] or ["logi_small", "mv_small"]
def run(cm, spec, C, tune, draws, **kw):
    q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
    sr, pr, _ = brng.chain_generators(5, C)
    t0 = time.time()
    res = cm.nuts_run(q0, brng.pack_pcg64(sr), tune=tune, draws=draws, philox_seed=5, **kw)
    wall = time.time() - t0
    ge = int(res.stats["tree_size"].sum())
    print(f"  C={C} {tune}+{draws}: kernel {res.kernel_ms:.0f} ms wall {wall:.2f}s launches {res.launches} grad-evals {ge} -> {ge/res.kernel_ms/1e3:.3f} M evals/s; "
          f"mean depth {res.stats['depth'].mean():.2f} div {res.stats['diverging'].mean():.4f} acc {res.stats['mean_tree_accept'][:, -draws or None:].mean():.3f} bad {np.sum(res.summary['bad_energy_at']>=0)}", flush=True)
    return res
for w in which:
    if w == "logi_small":
        spec = models.logistic(n_rows=100_000, n_features=128); cm = engine.CompiledModel(spec)
        print("[logistic 1e5 x 128]"); run(cm, spec, 512, 60, 20, mean0=None)
    if w == "logi_full":
        t0 = time.time(); spec = models.logistic(); print("spec", time.time() - t0); cm = engine.CompiledModel(spec)
        print("[logistic 1e6 x 128]"); run(cm, spec, 512, 30, 10)
    if w == "mv_small":
        spec = models.mvgauss(n=2000); cm = engine.CompiledModel(spec)
        print("[mvgauss n=2000 dense]"); run(cm, spec, 256, 60, 20, mass="dense")
    if w == "mv_full":
        t0 = time.time(); spec = models.mvgauss(); print("spec", time.time() - t0); cm = engine.CompiledModel(spec)
        print("[mvgauss n=10000 dense]"); run(cm, spec, 256, 30, 10, mass="dense")
