"""Multi-GPU: chains are the unit of sharding (one process per GPU, torch.distributed for plumbing).

The reference never exchanges anything between NUTS chains (per-chain adaptation,
hmc/quadpotential.py:335-355; one OS process per chain, sampling/parallel.py:477-589), so the data path
has NO collective: rank r runs a contiguous block of chains with streams derived on every rank from the
same ``SeedSequence`` spawn (sampling/mcmc.py:907), which makes results independent of the number of
GPUs.  The only communication is the final gather of draws and sampler stats (all_gather over
NCCL/NVLink, or gloo on CPU in tests).
"""
from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def world_size() -> int:
    d = _dist()
    return d.get_world_size() if d else 1


def rank() -> int:
    d = _dist()
    return d.get_rank() if d else 0


def chain_range(chains: int, r: int, world: int) -> tuple[int, int]:
    """Contiguous block of chains owned by rank r: sizes differ by at most one."""
    base, extra = divmod(chains, world)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def my_chain_range(chains: int) -> tuple[int, int]:
    return chain_range(chains, rank(), world_size())


def gather_chains(draws, stats: dict, chains: int):
    """All-gather per-rank shards [c_r, T, ...] into [chains, T, ...] on every rank.

    NCCL needs device tensors and equal shard shapes: shards are padded to the largest block."""
    import torch

    d = _dist()
    world = d.get_world_size()
    backend = d.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sizes = [chain_range(chains, r, world) for r in range(world)]
    cmax = max(hi - lo for lo, hi in sizes)

    def gather(a):
        was_np = isinstance(a, np.ndarray)
        t = torch.as_tensor(a).to(dev)
        if t.shape[0] < cmax:
            pad = torch.zeros((cmax - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
            t = torch.cat([t, pad], dim=0)
        outs = [torch.empty_like(t) for _ in range(world)]
        d.all_gather(outs, t.contiguous())
        full = torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)
        return full.cpu().numpy() if was_np else full

    return gather(draws), {k: gather(v) for k, v in stats.items()}


def max_over_ranks(x: float) -> float:
    """Device-side timing rule: a multi-GPU time is the max over ranks."""
    d = _dist()
    if not d:
        return x
    import torch

    dev = torch.device("cuda", torch.cuda.current_device()) if d.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float) -> float:
    d = _dist()
    if not d:
        return x
    import torch

    dev = torch.device("cuda", torch.cuda.current_device()) if d.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())
