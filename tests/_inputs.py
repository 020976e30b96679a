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


# ---- id-based redistribution inputs: the two generators of the reference's own tests ----
def xs_simple_case(world: int, id_min: int, id_max: int, rotation: int):
    """tests/xs_data_move_simple.cpp:30-79 : contiguous id blocks before; after = blocks grown/shrunk by one at the
    edges (odd/even ranks) and rotated.  Returns (id_min, id_max, ids_before[rank], ids_after[rank])."""
    if (id_max - id_min) // world < 4:
        id_max = id_min + world * 4
    rng = id_max - id_min
    before, after = [], []
    for rank in range(world):
        b0, b1 = (rng * rank) // world, (rng * (rank + 1)) // world
        a0, a1 = b0, b1
        if rank % 2 == 1:
            if a0 > 0:
                a0 -= 1
            if a1 < rng:
                a1 += 1
        else:
            if 0 < a0 < rng:
                a0 += 1
            if 0 < a1 < rng:
                a1 -= 1
        before.append(np.arange(b0, b1, dtype=np.uint64) + np.uint64(id_min))
        after.append(((np.arange(a0, a1, dtype=np.uint64) + np.uint64(rotation)) % np.uint64(rng)) + np.uint64(id_min))
    return id_min, id_max, before, after


def xs_random_case(world: int, seed: int, n: int, perms: int):
    """tests/xs_data_move_random.cpp:34-97 in spirit (numpy generator instead of std::mt19937): ids seed..seed+n shuffled by
    `perms` random swaps, cut into `world` contiguous pieces at random positions; a second round of swaps and cuts gives the
    distribution after."""
    rs = np.random.RandomState(seed)
    ids = np.arange(n, dtype=np.uint64) + np.uint64(seed)

    def swaps(a):
        a = a.copy()
        for _ in range(perms):
            i, j = rs.randint(0, n, 2)
            a[i], a[j] = a[j], a[i]
        return a

    def cuts():
        c = np.sort(np.concatenate([[0, n], rs.randint(0, n + 1, world - 1)])).astype(np.int64)
        return c

    before_all = swaps(ids)
    cb = cuts()
    after_all = swaps(before_all)
    ca = cuts()
    before = [before_all[cb[r]:cb[r + 1]].copy() for r in range(world)]
    after = [after_all[ca[r]:ca[r + 1]].copy() for r in range(world)]
    return seed, seed + n, before, after
