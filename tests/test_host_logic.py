"""CPU: host-side logic -- model specs, transforms, diagnostics, chain sharding (gloo, world_size 2)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from pymc_b200 import diagnostics, models, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_model_layouts_follow_registration_order():
    rd = models.radon()
    assert [v.name for v in rd.vars] == ["mu_a", "sigma_a_log__", "mu_b", "sigma_b_log__", "a", "b", "eps_log__"]
    assert rd.n == 175 and rd.vars[4].offset == 4 and rd.vars[6].offset == 174
    es = models.eight_schools()
    assert [v.name for v in es.vars] == ["mu", "tau_log__", "theta_t"] and es.n == 10
    sv = models.stochvol(T=20)
    assert sv.n == 23 and sv.vars[1].transform == "interval"


def test_constrain_applies_backward_transforms():
    sv = models.stochvol(T=5)
    q = np.zeros((2, 3, sv.n))
    q[..., 1] = 0.3
    q[..., 2] = -1.0
    out = sv.constrain(q)
    assert out["phi"].shape == (2, 3) and np.allclose(out["phi"], 2 / (1 + np.exp(-0.3)) - 1)
    assert np.allclose(out["sigma"], np.exp(-1.0)) and out["h"].shape == (2, 3, 5)


def test_radon_synthetic_data_shape():
    c, x, y = models.radon_data()
    assert len(y) == 919 and c.max() == 84 and np.bincount(c).min() >= 1 and set(np.unique(x)) <= {0.0, 1.0}
    c2, x2, y2 = models.radon_data()
    assert np.array_equal(c, c2) and np.array_equal(y, y2)  # fixed seed


def test_chain_range_partitions_all_chains():
    for chains, world in [(2048, 8), (10, 4), (3, 8), (256, 1)]:
        got = [parallel.chain_range(chains, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == chains
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [hi - lo for lo, hi in got]
        assert max(sizes) - min(sizes) <= 1


def test_ess_of_iid_and_ar1():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(8, 1000, 2))
    e = diagnostics.ess_bulk(x)
    assert np.all(e > 6000) and np.all(e < 10000)
    rho = 0.8
    y = np.empty((8, 2000))
    y[:, 0] = rng.normal(size=8)
    for t in range(1, 2000):
        y[:, t] = rho * y[:, t - 1] + np.sqrt(1 - rho**2) * rng.normal(size=8)
    theory = 8 * 2000 * (1 - rho) / (1 + rho)
    e = diagnostics.ess_bulk(y)[0]
    assert 0.7 * theory < e < 1.3 * theory
    assert abs(diagnostics.rhat(y)[0] - 1.0) < 0.02
    assert diagnostics.rhat(y + np.arange(8)[:, None])[0] > 1.5


def test_ess_torch_matches_numpy():
    import torch

    rng = np.random.default_rng(1)
    x = np.cumsum(rng.normal(size=(6, 400, 5)), axis=1) * 0.05 + rng.normal(size=(6, 400, 5))
    a = diagnostics.ess_bulk(x)
    b = diagnostics.ess_bulk_torch(torch.as_tensor(x), param_chunk=2).numpy()
    np.testing.assert_allclose(a, b, rtol=1e-6)


def test_rhat_torch_and_convergence_summary_match_numpy():
    import torch

    rng = np.random.default_rng(2)
    x = np.cumsum(rng.normal(size=(6, 300, 5)), axis=1) * 0.05 + rng.normal(size=(6, 300, 5))
    x[:, :, 3] += np.arange(6)[:, None] * 0.7  # one parameter whose chains disagree
    a = diagnostics.rhat(x)
    b = diagnostics.rhat_torch(torch.as_tensor(x), param_chunk=2).numpy()
    np.testing.assert_allclose(a, b, rtol=1e-6)
    assert a[3] > 1.2 and a[3] > a[[0, 1, 2, 4]].max()
    e, r = diagnostics.convergence_summary(x)
    assert abs(e - np.nanmin(diagnostics.ess_bulk(x))) < 1e-9 and abs(r - a.max()) < 1e-12
    e2, r2 = diagnostics.convergence_summary(x, large=10)  # "large" path: GPU if visible, else the same host path
    assert abs(e2 - e) <= 1e-6 * e and abs(r2 - r) <= 1e-6


WORKER = textwrap.dedent(
    """
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from pymc_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    chains, T, n = 5, 4, 3
    lo, hi = parallel.my_chain_range(chains)
    full = np.arange(chains * T * n, dtype=np.float64).reshape(chains, T, n)
    stats = {{"tree_size": (np.arange(chains * T, dtype=np.int32).reshape(chains, T))[lo:hi]}}
    d, s = parallel.gather_chains(full[lo:hi].copy(), stats, chains, dst=None)   # every rank gets everything
    assert np.array_equal(d, full), d
    assert np.array_equal(s["tree_size"], np.arange(chains * T, dtype=np.int32).reshape(chains, T))
    d0, s0 = parallel.gather_chains(full[lo:hi].copy(), stats, chains)            # default: rank 0 only
    if dist.get_rank() == 0:
        assert np.array_equal(d0, full) and np.array_equal(s0["tree_size"], np.arange(chains * T, dtype=np.int32).reshape(chains, T))
    else:
        assert d0 is None and s0 is None
    # pooled Welford over the chains of both ranks equals the single-process pooling
    rs = np.random.default_rng(5)
    cnt, mu, m2 = rs.integers(5, 50, chains).astype(float), rs.standard_normal((chains, n)), rs.random((chains, n)) * 10
    N, pm_, pM2 = parallel.pool_welford(cnt[lo:hi], mu[lo:hi], m2[lo:hi])
    N1 = cnt.sum(); m1 = (cnt[:, None] * mu).sum(0) / N1; M1 = (m2 + cnt[:, None] * (mu - m1) ** 2).sum(0)
    assert abs(N - N1) < 1e-12 and np.allclose(pm_, m1, rtol=1e-13) and np.allclose(pM2, M1, rtol=1e-11)
    # bench.py's N > 1 e2e leg: the per-chain scalar summaries of an engine result (which also carries [chains, n] vectors)
    import bench
    summ = {{"grad_evals": np.arange(lo, hi, dtype=np.int64), "bad_energy_at": np.full(hi - lo, -1, dtype=np.int32),
            "final_step_size": np.linspace(0.1, 0.2, chains)[lo:hi], "final_var": np.ones((hi - lo, n))}}
    sm, _ = parallel.gather_chains(bench.summary_matrix(summ), {{}}, chains, dst=0)
    if dist.get_rank() == 0:
        assert sm.shape == (chains, 3) and np.array_equal(sm[:, 0], np.arange(chains)) and np.all(sm[:, 1] == -1)
    else:
        assert sm is None
    assert parallel.max_over_ranks(float(dist.get_rank())) == 1.0
    assert parallel.sum_over_ranks(1.5) == 3.0
    dist.destroy_process_group()
    print("ok", dist.is_initialized())
    """
)


def test_chain_sharding_and_gather_world_size_2_gloo(tmp_path):
    port = 29500 + (os.getpid() % 500)
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert out.strip().startswith("ok")
