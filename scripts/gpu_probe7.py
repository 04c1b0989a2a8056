"""Where does the first lock-step call spend its time on a fresh box?"""
import os, sys, time
t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, _lib
print("imports", time.time() - t0, flush=True)
t0 = time.time(); cm0 = engine.CompiledModel(models.eight_schools()); cm0.logp_dlogp(np.zeros(10)); print("eight schools create+logp", time.time() - t0, flush=True)
spec = models.logistic(n_rows=400, n_features=8)
t0 = time.time(); cm = engine.CompiledModel(spec); print("logistic create (cublasCreate)", time.time() - t0, flush=True)
t0 = time.time(); cm.logp_dlogp(np.zeros((4, 8))); print("first dgemm", time.time() - t0, flush=True)
t0 = time.time(); cm.logp_dlogp(np.zeros((4, 8))); print("second", time.time() - t0, flush=True)
