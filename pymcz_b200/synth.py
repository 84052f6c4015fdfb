"""Synthetic galaxy tables for the benchmark / parity configurations (host side, numpy only).

The generator is the one SURVEY.md section 8(d) fixes (so that builder and judge use the same
table): three template line-flux rows in the reference's ``alllines`` order
(/root/reference/pyMCZ/mcz.py:45), jittered per galaxy, float32 like the reference's reader
(/root/reference/pyMCZ/mcz.py:150).
"""
from __future__ import annotations

import numpy as np

#: template rows: exampledata row 1, exampledata row 2 (missing lines filled), a low-metallicity row
TEMPLATES = np.array([
    [2.875, 1.251, 0.064 / 3, 0.064, 0.1, 4.026, 0.781, 0.821, 0.573, 0.30, 0.75],
    [1.842, 0.958, 0.302 / 3, 0.302, 0.127, 4.746, 1.642, 0.941, 0.543, 0.30, 0.75],
    [1.5, 1.0, 4.0 / 3, 4.0, 0.03, 2.95, 0.06, 0.12, 0.09, 0.10, 0.25],
])

#: BASELINE.json configs 3-5: (galaxies, nsample, md)
CONFIGS = {
    3: dict(G=10_000, nsample=10_000, md="all"),
    4: dict(G=1_000_000, nsample=2_000, md="all"),
    5: dict(G=100_000, nsample=10_000, md="KD02"),
}


def synthetic_table(G: int, config: int = 4, start: int = 0, stop: int | None = None):
    """(meas, err) float32 [G, 11] for BASELINE config ``config``.

    ``start/stop`` slice the table by global galaxy index *after* generating the full-G random
    stream, so every rank of a sharded run sees the same global table.
    """
    rng = np.random.Generator(np.random.PCG64(1000 + config))
    jit = rng.uniform(-0.3, 0.3, (G, 11))
    jit[:, [1, 5]] = rng.uniform(-0.05, 0.05, (G, 2))
    norm = rng.uniform(-0.5, 0.5, (G, 1))
    snr = rng.uniform(10, 50, (G, 11))
    T = TEMPLATES[np.arange(G) % 3]
    meas = (T * 10.0 ** (jit + norm)).astype(np.float32)
    err = (meas / snr).astype(np.float32)
    stop = G if stop is None else stop
    return np.ascontiguousarray(meas[start:stop]), np.ascontiguousarray(err[start:stop])


def shard_bounds(G: int, world: int, rank: int):
    """Contiguous galaxy range [lo, hi) of ``rank``."""
    lo = (G * rank) // world
    hi = (G * (rank + 1)) // world
    return lo, hi
