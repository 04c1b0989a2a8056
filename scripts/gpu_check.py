"""Exploratory GPU check (prints diagnostics; the asserting versions live in tests/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymc_b200 import models, engine, rng as brng, _lib
from oracle import logp_numpy, nuts_numpy

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def relerr(a, b):
    return float(np.max(np.abs(a - b) / (1e-300 + np.maximum(np.abs(a), np.abs(b)))))

for name, spec in [("std_normal", models.std_normal(100)), ("eight_schools", models.eight_schools()), ("radon", models.radon()),
                   ("radon_small", models.radon(40, 7, 5))]:
    cm = engine.CompiledModel(spec)
    f = logp_numpy.make_logp(spec)
    r = np.random.default_rng(0)
    Q = spec.initial_point() + r.uniform(-1.5, 1.5, (64, spec.n))
    lp, g = cm.logp_dlogp(Q)
    lo = np.array([f(q)[0] for q in Q]); go = np.array([f(q)[1] for q in Q])
    print(f"[logp] {name}: logp rel {relerr(lp, lo):.2e} grad maxabs {np.max(np.abs(g-go)):.2e} rel-to-norm {np.max(np.abs(g-go))/np.max(np.abs(go)):.2e}")

def run_case(fname, spec):
    d = np.load(os.path.join(G, fname + ".npz"))
    cm = engine.CompiledModel(spec)
    C = len(d["seeds"]); tune, draws, adapt = int(d["tune"]), int(d["draws"]), bool(d["adapt"])
    step_rngs = [np.random.default_rng(int(s)) for s in d["seeds"]]
    for g_ in step_rngs: g_.spawn(1)  # potential stream spawned by setup_chain (does not advance the bit stream)
    states = brng.pack_pcg64(step_rngs)
    kw = {}
    if adapt:
        kw.update(mass="diag_adapt", mean0=d["q0"], var0=np.ones_like(d["q0"]), adapt_step_size=True)
    else:
        eps = d["eps"]
        scale = 0.25 if np.isnan(eps[0]) else None
        if scale is None:
            # per-chain eps: run chains separately
            pass
        kw.update(mass="diag", var0=d["var"], adapt_step_size=False)
    if fname == "radon_small_adapt":
        kw.update(max_treedepth=6, early_max_treedepth=4)
    outs = []
    if (not adapt) and not np.isnan(d["eps"][0]):
        for c in range(C):
            res = cm.nuts_run(d["q0"][c:c+1], states[c:c+1], tune=tune, draws=draws, z=d["z"][c:c+1],
                              step_scale=float(d["eps"][c]) * spec.n**0.25, **{k: (v[c:c+1] if isinstance(v, np.ndarray) else v) for k, v in kw.items()})
            outs.append(res)
        dq = np.concatenate([o.draws for o in outs]); st = {k: np.concatenate([o.stats[k] for o in outs]) for k in outs[0].stats}
    else:
        res = cm.nuts_run(d["q0"], states, tune=tune, draws=draws, z=d["z"], **kw)
        dq, st = res.draws, res.stats
        print("   kernel ms", res.kernel_ms, "bad", res.summary["bad_energy_at"])
    for c in range(C):
        same = (st["tree_size"][c] == d["stat_tree_size"][c]) & (st["depth"][c] == d["stat_depth"][c]) & (st["index_in_trajectory"][c] == d["stat_index_in_trajectory"][c])
        first_bad = int(np.argmin(same)) if not same.all() else -1
        err = np.max(np.abs(dq[c] - d["draws_q"][c]), axis=1)
        upto = first_bad if first_bad >= 0 else len(err)
        print(f"[nuts] {fname} chain {c}: discrete stats identical for {'ALL' if first_bad<0 else first_bad} of {len(same)} draws; "
              f"max |dq| before that {err[:upto].max() if upto else 0:.2e}; energy rel {relerr(st['energy'][c][:upto], d['stat_energy'][c][:upto]) if upto else 0:.2e}; "
              f"step_size rel {relerr(st['step_size'][c][:upto], d['stat_step_size'][c][:upto]) if upto else 0:.2e}")
    return dq, st, d

run_case("std_normal_fixed", models.std_normal(100))
run_case("eight_schools_fixed", models.eight_schools())
run_case("eight_schools_adapt", models.eight_schools())
run_case("radon_fixed", models.radon())
run_case("radon_small_adapt", models.radon(40, 7, 5))
run_case("radon_adapt", models.radon())

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
