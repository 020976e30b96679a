"""CPU, build container only: the C restatement against the LIVE reference (oracle/_ref, compiled from
/root/reference's own sources) on seeded inputs at sizes the golden file does not hold."""
import numpy as np
import pytest

from tests._inputs import lcg_uniform, poisson_counts


def test_axpy_1m(orc, ref):
    n = 1_000_003
    x, y = lcg_uniform(n, 1, -1, 1), lcg_uniform(n, 2, -1, 1)
    yo, yr = y.copy(), y.copy()
    orc.axpy(yo, x, 0.5)
    ref.axpy(yr, x, 0.5)
    assert np.array_equal(yo, yr)
    assert np.array_equal(yo, y + 0.5 * x)  # unfused mul+add == numpy


def test_value_add_and_lanes(orc, ref):
    a = lcg_uniform(257 * 129, 3).reshape(257, 129)
    ao, ar = a.copy(), a.copy()
    orc.value_add(ao, 1.1)
    ref.value_add(ar, 1.1)
    assert np.array_equal(ao, ar)
    b = lcg_uniform(257 * 129, 4).reshape(257, 129)
    for mode in (0, 1, 2):
        o1, o2, r1, r2 = a.copy(), b.copy(), a.copy(), b.copy()
        orc.lane_sequence(o1, o2, 1.1, 1.2, 1.3, 1.4)
        ref.lane_sequence(r1, r2, 1.1, 1.2, 1.3, 1.4, mode=mode, use_tasks=(mode == 0))
        assert np.array_equal(o1, r1) and np.array_equal(o2, r2)


def test_reduce_min_16m(orc, ref):
    d = lcg_uniform(16 << 20, 7)
    m_ref, _ = ref.reduce_min_chunks(d)
    assert orc.reduce_min(d) == m_ref == d.min()


def test_soatl_layout_random_sequences(orc, ref):
    rng = np.random.default_rng(0)
    for cfg, desc in ref.SOATL_CONFIGS.items():
        sizes = [int(s) for s in rng.integers(0, 5000, 40)]
        caps, nbytes, offs = ref.soatl_layout(cfg, sizes)
        cap = 0
        for s, c, b, o in zip(sizes, caps, nbytes, offs):
            cap = orc.update_capacity(s, cap, desc["chunk"])
            total, oo = orc.layout(desc["align"], desc["fields"], cap)
            assert (cap, total) == (c, b)
            assert orc.storage_size_calc(desc["align"], desc["chunk"], desc["fields"], cap) == b
            if cap:
                assert oo == o


def test_soatl_dist_100k(orc, ref):
    n = 100_003
    rx, ry, rz = lcg_uniform(n, 10, 0, 1), lcg_uniform(n, 11, 0, 1), lcg_uniform(n, 12, 0, 1)
    e_ref, _ = ref.soatl_dist(rx, ry, rz)
    e = orc.soatl_dist(rx, ry, rz)
    assert np.isnan(e[0]) and np.array_equal(e, e_ref, equal_nan=True)


def test_soatl_rmw8_100k(orc, ref):
    n = 100_003
    cols = [lcg_uniform(n, 20 + k, 0, 1) for k in range(8)]
    c = [0.125 * (k + 1) for k in range(8)]
    r = [col.copy() for col in cols]
    ref.soatl_rmw8(r, 1.0000001, c, align=256)
    cap = orc.update_capacity(n, 0, 16)
    total, offs = orc.layout(256, [8] * 8, cap)
    st = np.zeros(total, np.uint8)
    for k in range(8):
        st[offs[k]:offs[k] + 8 * n] = cols[k].view(np.uint8)
    orc.soatl_rmw(st, 256, 16, 8, n, cap, 1.0000001, c)
    for k in range(8):
        assert np.array_equal(st[offs[k]:offs[k] + 8 * n].view(np.float64), r[k])


def test_cell_sum_32k_cells(orc, ref):
    counts = poisson_counts(32 ** 3)
    e = lcg_uniform(int(counts.sum()), 40, -1, 1)
    s_ref, t_ref, nbytes, _ = ref.cell_sum(counts, e)
    fields = [8, 8, 8, 8]
    total_bytes, cap, off = orc.cell_grid_layout(counts, 64, 16, fields, 64)
    assert total_bytes == nbytes
    arena = np.zeros(total_bytes, np.uint8)
    starts = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)[:-1]]).astype(np.int64)
    for c in range(counts.size):
        n = int(counts[c])
        if n:
            a = int(off[c]) + 3 * ((int(cap[c]) * 8 + 63) // 64 * 64)
            arena[a:a + 8 * n] = e[starts[c]:starts[c] + n].view(np.uint8)
    sums, total = orc.cell_sum(arena, off, counts, cap, 64, fields, 3)
    assert np.array_equal(sums, s_ref)
    assert abs(total - t_ref) <= 1e-12 * np.abs(e).sum()
    # field-major restatement (values cell after cell + offsets): same per-cell bits as the reference's FieldArrays loop
    begin = np.zeros(counts.size + 1, np.uint64)
    begin[1:] = np.cumsum(counts, dtype=np.uint64)
    sp, tp = orc.cell_sum_packed(e, begin)
    assert np.array_equal(sp, s_ref)
    assert abs(tp - t_ref) <= 1e-12 * np.abs(e).sum()


def test_parallel_for_3d(orc, ref):
    n = 16
    mo, mr = np.zeros(n ** 3), np.zeros(n ** 3)
    orc.parallel_for_3d(mo, n, (2, 3, 3), (n - 1, n - 3, n - 3))
    ref.parallel_for_3d(mr, n, (2, 3, 3), (n - 1, n - 3, n - 3))
    assert np.array_equal(mo, mr)
