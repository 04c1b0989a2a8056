import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from b200_helpers import SPEC_OF, gpu_free_run, discrete_equal
from pymc_b200 import engine
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ("eight_schools_warm_adapt", "radon_warm_adapt"):
    d = dict(np.load(os.path.join(G, name + ".npz")))
    cm = engine.CompiledModel(SPEC_OF[name]())
    res, _ = gpu_free_run(cm, d, name)
    st = {k: v[0] for k, v in res.stats.items()}
    ok = discrete_equal(st, d, 0)
    err = np.max(np.abs(res.draws[0] - d["draws_q"][0]), axis=1)
    es = np.abs(st["step_size"] - d["stat_step_size"][0]) / d["stat_step_size"][0]
    ea = np.abs(st["mean_tree_accept"] - d["stat_mean_tree_accept"][0])
    print(name, "first bad", int(np.argmin(ok)) if not ok.all() else -1)
    for t in list(range(0, 60, 6)) + list(range(60, 280, 10)):
        print(f"  t={t:3d} ok={bool(ok[t])} |dq|={err[t]:.2e} step_rel={es[t]:.2e} accept_abs={ea[t]:.2e} depth={st['depth'][t]} gold_depth={d['stat_depth'][0][t]}")
    fv = res.summary["final_var"][0]
    print("  final var rel err", np.max(np.abs(fv - d["final_var"][0]) / d["final_var"][0]))
