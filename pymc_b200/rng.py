"""Per-chain random streams, derived exactly as the reference derives them (host side, NumPy).

Reference order:
  * ``rngs = get_random_generator(random_seed).spawn(chains)``          sampling/mcmc.py:907
  * ``random_seed_list = [rng.integers(2**30) for rng in rngs]``         sampling/mcmc.py:908  (consumes one draw)
  * ``step.setup_chain(rng, ...)`` -> ``step.rng = rng``                 step_methods/compound.py:233-250
  * ``potential.set_rng(step.rng.spawn(1)[0])``                          hmc/base_hmc.py:300-302
The `step` stream feeds the tree (direction, multinomial picks); the `potential` stream feeds the
momentum normals.  The device replays the `step` stream bit-exactly (csrc/rng.cuh: Pcg64).
"""
from __future__ import annotations

import numpy as np

from ._lib import PCG64_DTYPE

_M64 = (1 << 64) - 1


def chain_generators(random_seed, chains: int):
    """-> (step_rngs, potential_rngs, jitter_seeds) as NumPy Generators / ints, one per chain."""
    if isinstance(random_seed, np.random.Generator):
        root = random_seed
    else:
        root = np.random.default_rng(random_seed)
    rngs = root.spawn(chains)
    jitter_seeds = [int(r.integers(2**30)) for r in rngs]
    pots = [r.spawn(1)[0] for r in rngs]
    return rngs, pots, jitter_seeds


def philox_key(random_seed) -> int:
    """63-bit key of the device momentum noise (Philox4x32): an independent child of the ROOT seed sequence, on a spawn
    key no chain stream uses -- full entropy, and not a function of any chain's jitter draws."""
    if isinstance(random_seed, np.random.Generator):
        ss = random_seed.bit_generator.seed_seq
        base = np.random.SeedSequence(ss.entropy, spawn_key=tuple(ss.spawn_key) + (0xB200,))
    else:
        base = np.random.SeedSequence(random_seed, spawn_key=(0xB200,))
    w = base.generate_state(2, dtype=np.uint64)
    return int(w[0] >> np.uint64(1))


def pack_pcg64(generators) -> np.ndarray:
    """NumPy PCG64 Generators -> structured array of (state, inc) split into 64-bit halves."""
    out = np.empty(len(generators), dtype=PCG64_DTYPE)
    for i, g in enumerate(generators):
        bg = g.bit_generator
        if type(bg).__name__ != "PCG64":
            raise TypeError("the device replays NumPy's PCG64; got " + type(bg).__name__)
        st = bg.state
        # a buffered 32-bit half (left by rng.integers(2**30), mcmc.py:908) is only ever consumed by 32-bit
        # requests; the sampler draws doubles (next_uint64), so it is irrelevant to the replay and kept as is.
        s, inc = st["state"]["state"], st["state"]["inc"]
        out[i] = (s >> 64, s & _M64, inc >> 64, inc & _M64)
    return out


def unpack_pcg64(states: np.ndarray, generators) -> None:
    """Write device-advanced states back into the NumPy Generators (resume on the host)."""
    for rec, g in zip(states, generators):
        st = g.bit_generator.state
        st["state"]["state"] = (int(rec["state_hi"]) << 64) | int(rec["state_lo"])
        st["state"]["inc"] = (int(rec["inc_hi"]) << 64) | int(rec["inc_lo"])
        g.bit_generator.state = st


def momentum_noise(potential_rngs, iterations: int, n: int) -> np.ndarray:
    """z[C, iterations, n]: what ``potential.random()`` would draw, one ``normal(size=n)`` per iteration
    (hmc/quadpotential.py:325, :619)."""
    z = np.empty((len(potential_rngs), iterations, n))
    for c, g in enumerate(potential_rngs):
        for t in range(iterations):
            z[c, t] = g.normal(size=n)
    return z


# ---- host replica of the device Philox stream (csrc/rng.cuh: philox_normal) ----------------------
def philox_normal(key: int, chain, draw, elem):
    """NumPy replica of the device momentum generator; broadcastable integer arrays in, float64 out."""
    c0 = np.asarray(elem, dtype=np.uint64) & 0xFFFFFFFF
    c1 = np.asarray(draw, dtype=np.uint64) & 0xFFFFFFFF
    c2 = np.asarray(chain, dtype=np.uint64) & 0xFFFFFFFF
    c0, c1, c2 = np.broadcast_arrays(c0, c1, c2)
    c0, c1, c2 = c0.copy(), c1.copy(), c2.copy()
    c3 = np.full_like(c0, 0x4E555453)
    k0 = np.uint64(key & 0xFFFFFFFF)
    k1 = np.uint64((key >> 32) & 0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
    m32 = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> s32) ^ c1 ^ k0) & m32
        n1 = p1 & m32
        n2 = ((p0 >> s32) ^ c3 ^ k1) & m32
        n3 = p0 & m32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & m32, (k1 + W1) & m32
    a = (c0 << s32) | c1
    b = (c2 << s32) | c3
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    u2 = (b >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
