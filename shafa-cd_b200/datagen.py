"""Synthetic inputs of the shapes BASELINE.json names.

Data specification only: numpy `default_rng` streams with fixed seeds so that the
parity tests, the reference arm and the CUDA arm all see the same bytes.  Every
generator keeps >= 2 distinct symbols per block (a single-symbol N-mode block
crashes the reference decoder, /root/reference/src/modules/d.c:533-536).
"""
from __future__ import annotations

import numpy as np

KIB = 1024
MIB = 1024 * 1024
BLOCK_K = 655360      # -b K  (utils/file.h:9)
BLOCK_m = 8388608     # -b m  (utils/file.h:10)
BLOCK_M = 67108864    # -b M  (utils/file.h:11)
BLOCK_DEFAULT = 65536  # no -b (shafa.c:304-305)

_TEXT_ALPHABET = np.frombuffer(
    b" etaoinshrdlcumwfgypbvkjxqzETAOINSHRDLCUMWFGYPBVKJXQZ0123456789.,\n", dtype=np.uint8
)


def zipf(n: int, seed: int = 12345) -> np.ndarray:
    """C1: skewed bytes, p(k) ~ (k+1)^-1.2 over a permuted alphabet (N mode)."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 257) ** 1.2
    p /= p.sum()
    perm = rng.permutation(256).astype(np.uint8)
    out = np.empty(n, dtype=np.uint8)
    step = 64 * MIB
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = perm[rng.choice(256, size=hi - lo, p=p)]
    return out


def textlike(n: int, seed: int = 2024) -> np.ndarray:
    """C2: 66-symbol 1/rank text with a run of 4-16 spaces every 80 bytes (R mode)."""
    rng = np.random.default_rng(seed)
    syms = _TEXT_ALPHABET
    p = 1.0 / np.arange(1, len(syms) + 1)
    p /= p.sum()
    out = np.empty(n, dtype=np.uint8)
    step = 64 * MIB
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = syms[rng.choice(len(syms), size=hi - lo, p=p)]
    if n > 112:
        pos = np.arange(0, n - 32, 80)
        length = rng.integers(4, 17, size=pos.size)
        for l in range(4, 17):
            pp = pos[length == l]
            for k in range(l):
                out[pp + k] = 32
    return out


def runheavy(n: int, seed: int = 99) -> np.ndarray:
    """C3: long byte runs (log-normal lengths), 32 run symbols incl. 0x00 (R mode)."""
    rng = np.random.default_rng(seed)
    q = 1.0 / np.arange(1, 33)
    q /= q.sum()
    out = np.empty(n, dtype=np.uint8)
    filled = 0
    while filled < n:
        want = min(n - filled, 256 * MIB)
        m = max(16, want // 40)
        L = np.minimum(np.maximum(1, rng.lognormal(mean=4.0, sigma=1.5, size=m).astype(np.int64)), 100000)
        s = rng.choice(32, size=m, p=q).astype(np.uint8)
        chunk = np.repeat(s, L)[:want]
        out[filled:filled + chunk.size] = chunk
        filled += chunk.size
    return out


def uniform(n: int, seed: int = 7) -> np.ndarray:
    """C4: incompressible bytes (N mode, run F with -c f)."""
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def mixed(n: int, seed: int = 5, segment: int = 64 * MIB) -> np.ndarray:
    """C5: segments cycling run-heavy / zipf / text / uniform; the first segment is
    run-heavy so that block 0 selects R mode."""
    gens = (runheavy, zipf, textlike, uniform)
    out = np.empty(n, dtype=np.uint8)
    i = 0
    for lo in range(0, n, segment):
        hi = min(n, lo + segment)
        out[lo:hi] = gens[i % 4](hi - lo, seed=seed + i)
        i += 1
    return out


def mixed4m(n: int, seed: int = 5) -> np.ndarray:
    """`mixed` with 4 MiB segments (test-sized)."""
    return mixed(n, seed=seed, segment=4 * MIB)


def mixed_tiled(n: int, seed: int = 5, segment: int = 64 * MIB) -> np.ndarray:
    """C5 at its full 16 GiB: the same cycle of 64 MiB run-heavy / zipf / text / uniform segments as
    `mixed`, but built from eight base segments (two per kind, seeds seed .. seed+7) so that the input
    can be produced in seconds on the GPU box: segment i is base (i % 8) rotated by (i * 1000003) mod
    `segment` bytes -- the same statistics per kind, different bytes in every block.  Segment 0 is the
    unrotated run-heavy base, so block 0 selects R mode."""
    gens = (runheavy, zipf, textlike, uniform)
    bases = [gens[k % 4](segment, seed=seed + k) for k in range(8)]
    out = np.empty(n, dtype=np.uint8)
    i = 0
    for lo in range(0, n, segment):
        hi = min(n, lo + segment)
        base = bases[i % 8]
        r = (i * 1000003) % segment
        m = hi - lo
        first = min(m, segment - r)
        out[lo:lo + first] = base[r:r + first]
        if first < m:
            out[lo + first:hi] = base[:m - first]
        i += 1
    return out


GENERATORS = {"zipf": zipf, "textlike": textlike, "runheavy": runheavy, "uniform": uniform,
              "mixed": mixed, "mixed4m": mixed4m, "mixed_tiled": mixed_tiled}
