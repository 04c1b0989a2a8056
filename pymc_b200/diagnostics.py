"""Rank-normalised bulk ESS and split R-hat (Vehtari, Gelman, Simpson, Carpenter, Buerkner 2021).

The reference delegates these to arviz_stats (pymc/stats/convergence.py:108-131, pymc/stats/__init__.py:30-37),
which is not part of the reference tree nor installed here; this restates the published algorithm
(BASELINE.json's metric needs ESS/sec).  Two back-ends with identical maths:
  * NumPy (host)           -- `ess_bulk`, `rhat`
  * torch (device, fp64)   -- `ess_bulk_torch`: draws never leave HBM; used by bench.py
Input layout everywhere: x[chains, draws, ...params].
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------
# NumPy
# ---------------------------------------------------------------------------------------------
def _split(x):
    C, T = x.shape[:2]
    h = T // 2
    return np.concatenate([x[:, :h], x[:, T - h :]], axis=0)


def _rank_normalise(x):
    """Fractional ranks over all chains -> normal scores, Blom offset 3/8 (paper eq. 14)."""
    from scipy import stats as st

    shape = x.shape
    flat = x.reshape(shape[0] * shape[1], -1)
    r = st.rankdata(flat, axis=0, method="average")
    z = st.norm.ppf((r - 0.375) / (flat.shape[0] + 0.25))
    return z.reshape(shape)


def _autocov(x):
    """Biased autocovariance along axis 1 via FFT."""
    T = x.shape[1]
    m = 1 << int(np.ceil(np.log2(2 * T)))
    xc = x - x.mean(axis=1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=1)[:, :T]
    return ac / T


def _ess_core(x):
    """ESS of x[chains, draws, ...] with Geyer's initial monotone positive sequence."""
    C, T = x.shape[:2]
    P = int(np.prod(x.shape[2:], dtype=np.int64)) if x.ndim > 2 else 1
    x = x.reshape(C, T, P)
    acov = _autocov(x)  # [C, T, P]
    chain_mean = x.mean(axis=1)
    mean_var = acov[:, 0].mean(axis=0) * T / (T - 1.0)
    var_plus = mean_var * (T - 1.0) / T
    if C > 1:
        var_plus = var_plus + chain_mean.var(axis=0, ddof=1)
    out = np.empty(P)
    for j in range(P):
        if not np.isfinite(var_plus[j]) or var_plus[j] <= 0:
            out[j] = np.nan
            continue
        rho = 1.0 - (mean_var[j] - acov[:, :, j].mean(axis=0)) / var_plus[j]
        rho[0] = 1.0
        # sums of adjacent pairs, truncated at the first negative pair, then made monotone
        npair = T // 2
        pair = rho[0 : 2 * npair : 2] + rho[1 : 2 * npair : 2]
        neg = np.nonzero(pair < 0)[0]
        k = neg[0] if len(neg) else npair
        pair = np.minimum.accumulate(pair[:k])
        tau = -1.0 + 2.0 * pair.sum()
        tau = max(tau, 1.0 / np.log10(C * T))
        out[j] = C * T / tau
    return out.reshape(x.shape[2:]) if P > 1 or x.ndim > 2 else out


def ess_bulk(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    return _ess_core(_rank_normalise(_split(x)))


def rhat(x):
    """Rank-normalised split R-hat (max of bulk and folded)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]

    def _r(z):
        C, T = z.shape[:2]
        W = z.var(axis=1, ddof=1).mean(axis=0)
        B = T * z.mean(axis=1).var(axis=0, ddof=1)
        return np.sqrt(((T - 1.0) / T * W + B / T) / W)

    s = _split(x)
    bulk = _r(_rank_normalise(s))
    folded = _r(_rank_normalise(np.abs(s - np.median(s.reshape(-1, *s.shape[2:]), axis=0))))
    return np.maximum(bulk, folded)


# ---------------------------------------------------------------------------------------------
# torch (device)
# ---------------------------------------------------------------------------------------------
def ess_bulk_torch(x, param_chunk: int = 16):
    """Same estimator on a torch tensor x[chains, draws, params] (any device), chunked over params."""
    import torch

    C0, T0, P = x.shape
    h = T0 // 2
    out = torch.empty(P, dtype=torch.float64, device=x.device)
    for p0 in range(0, P, param_chunk):
        xs = x[:, :, p0 : p0 + param_chunk].to(torch.float64)
        xs = torch.cat([xs[:, :h], xs[:, T0 - h :]], dim=0)  # split chains
        C, T, Pc = xs.shape
        flat = xs.reshape(C * T, Pc)
        # average ranks are not needed for continuous draws: ordinal ranks (ties have measure zero)
        order = flat.argsort(dim=0)
        ranks = torch.empty_like(flat)
        ar = torch.arange(1, C * T + 1, dtype=torch.float64, device=x.device).unsqueeze(1).expand(-1, Pc)
        ranks.scatter_(0, order, ar)
        pfrac = (ranks - 0.375) / (C * T + 0.25)
        z = (torch.special.ndtri(pfrac)).reshape(C, T, Pc)
        m = 1 << int(np.ceil(np.log2(2 * T)))
        zc = z - z.mean(dim=1, keepdim=True)
        f = torch.fft.rfft(zc, n=m, dim=1)
        acov = torch.fft.irfft(f * f.conj(), n=m, dim=1)[:, :T] / T
        mean_var = acov[:, 0].mean(dim=0) * T / (T - 1.0)
        var_plus = mean_var * (T - 1.0) / T + z.mean(dim=1).var(dim=0, unbiased=True)
        rho = 1.0 - (mean_var.unsqueeze(0) - acov.mean(dim=0)) / var_plus.unsqueeze(0)  # [T, Pc]
        rho[0] = 1.0
        npair = T // 2
        pair = rho[0 : 2 * npair : 2] + rho[1 : 2 * npair : 2]  # [npair, Pc]
        negative = (pair < 0).to(torch.int64)
        alive = (negative.cumsum(dim=0) == 0).to(torch.float64)  # 1 before the first negative pair
        pair = torch.cummin(pair, dim=0).values
        tau = -1.0 + 2.0 * (pair * alive).sum(dim=0)
        tau = torch.clamp(tau, min=1.0 / float(np.log10(C * T)))
        out[p0 : p0 + param_chunk] = C * T / tau
    return out


def rhat_torch(x, param_chunk: int = 16):
    """Rank-normalised split R-hat (max of bulk and folded) on a torch tensor x[chains, draws, params], any device."""
    import torch

    C0, T0, P = x.shape
    h = T0 // 2
    out = torch.empty(P, dtype=torch.float64, device=x.device)

    def normal_scores(flat):  # [N, Pc] -> normal scores of the ordinal ranks (continuous draws: ties have measure zero)
        N, Pc = flat.shape
        order = flat.argsort(dim=0)
        ranks = torch.empty_like(flat)
        ar = torch.arange(1, N + 1, dtype=torch.float64, device=x.device).unsqueeze(1).expand(-1, Pc)
        ranks.scatter_(0, order, ar)
        return torch.special.ndtri((ranks - 0.375) / (N + 0.25))

    def r(z):  # z[C, T, Pc]
        C, T = z.shape[:2]
        W = z.var(dim=1, unbiased=True).mean(dim=0)
        B = T * z.mean(dim=1).var(dim=0, unbiased=True)
        return torch.sqrt(((T - 1.0) / T * W + B / T) / W)

    for p0 in range(0, P, param_chunk):
        xs = x[:, :, p0 : p0 + param_chunk].to(torch.float64)
        s = torch.cat([xs[:, :h], xs[:, T0 - h :]], dim=0)
        C, T, Pc = s.shape
        bulk = r(normal_scores(s.reshape(C * T, Pc)).reshape(C, T, Pc))
        med = s.reshape(C * T, Pc).median(dim=0).values
        folded = r(normal_scores((s - med).abs().reshape(C * T, Pc)).reshape(C, T, Pc))
        out[p0 : p0 + param_chunk] = torch.maximum(bulk, folded)
    return out


def convergence_summary(x, large: int = 5_000_000):
    """(min bulk ESS, max R-hat) of x[chains, draws, params].  Arrays with more than `large` elements go through the
    torch back-end on the GPU when one is visible (a 2048 x 1000 x 175 run costs minutes in NumPy, about a second there)."""
    x = np.asarray(x)
    if x.ndim == 2:
        x = x[:, :, None]
    if x.size > large:
        try:
            import torch

            if torch.cuda.is_available():
                e, rh = [], []
                step = max(1, int(2.0e8 // (x.shape[0] * x.shape[1])))  # params per host->device slab (~1.6 GB)
                for p0 in range(0, x.shape[2], step):
                    t = torch.as_tensor(np.ascontiguousarray(x[:, :, p0 : p0 + step]), device="cuda")
                    e.append(ess_bulk_torch(t).cpu())
                    rh.append(rhat_torch(t).cpu())
                return float(np.nanmin(torch.cat(e).numpy())), float(np.nanmax(torch.cat(rh).numpy()))
        except Exception as e:  # any device-side problem: the host path below is always available, but say so
            import warnings

            warnings.warn(f"convergence_summary: device path failed ({type(e).__name__}: {e}); falling back to the NumPy "
                          "estimator, which takes minutes for large posteriors", RuntimeWarning, stacklevel=2)
    return float(np.nanmin(ess_bulk(x))), float(np.nanmax(rhat(x)))
