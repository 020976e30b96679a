"""Deterministic synthetic inputs shared by the golden generator, the tests and bench.py.

Integer-hash based (no dependence on any RNG library's stream), NaN-free, no signed-zero mix, so that the
bit-exact min/max claims are well defined."""
import numpy as np


def lcg_uniform(n, seed, lo=1e-3, hi=1.0):
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    h = (h ^ (h >> np.uint64(15))) * np.uint64(2246822519) & np.uint64(0xFFFFFFFF)
    h = (h ^ (h >> np.uint64(13))) & np.uint64(0xFFFFFFFF)
    return lo + (hi - lo) * (h.astype(np.float64) / 4294967296.0)


def hx(a):
    return [float(v).hex() for v in np.asarray(a, dtype=np.float64).ravel()]


def unhx(lst):
    return np.array([float.fromhex(s) for s in lst], dtype=np.float64)


def poisson_counts(ncells, mean=16.0, seed=42, lo=0, hi=64):
    """particles per cell ~ Poisson(mean) clamped to [lo,hi] (config C4); numpy Generator(PCG64) stream."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.clip(rng.poisson(mean, ncells), lo, hi).astype(np.uint32)
