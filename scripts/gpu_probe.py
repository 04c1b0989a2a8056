"""Exploratory GPU perf probe (prints diagnostics)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib

# throughput probe: radon 2048 chains, 100+100
spec = models.radon(); cm = engine.CompiledModel(spec)
for C in (2048,):
    r = np.random.default_rng(1)
    q0 = spec.initial_point() + r.uniform(-1, 1, (C, spec.n))
    sr, pr, _ = brng.chain_generators(123, C)
    states = brng.pack_pcg64(sr)
    for wpb, hot in ((4, 2), (8, 1), (8, 2), (4, 1), (2, 3)):
        os.environ["B200_NUTS_WPB"] = str(wpb); os.environ["B200_NUTS_HOT"] = str(hot)
        t = time.time()
        res = cm.nuts_run(q0, states.copy(), tune=200, draws=100, mean0=np.broadcast_to(q0.mean(0), q0.shape).copy(), philox_seed=5)
        dt = time.time() - t
        ge = res.grad_evals
        print(f"[perf] radon C={C} wpb={wpb} hot={hot}: kernel {res.kernel_ms:.1f} ms, wall {dt:.2f}s, tree evals {ge}, {ge/res.kernel_ms/1e3:.2f} M evals/s, "
              f"mean depth {res.stats['depth'].mean():.2f} div {res.stats['diverging'][:,200:].sum()} step {res.summary['final_step_size'].mean():.4f}")
lib = _lib.load()
import ctypes
tf = ctypes.c_double(); lib.b200_measure_fp64_tflops(ctypes.byref(tf)); print("[peak] fp64 DFMA TFLOP/s", tf.value)
