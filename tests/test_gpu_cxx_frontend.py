"""GPU (-m gpu): programs written against the reference's C++ API (tests/cxx/frontend_test.cu: the scenarios of the
onikaParallelTutorial plugins, tests/parallel_queue_lane.cu, tests/benchmark.cpp, tests/soatlserializetest.cpp) compiled
against THIS repo's include/onika front end + libonika_b200.so, run on the GPU, outputs checked against the oracle and
the golden vectors.  The generic functor path (user functors through block_parallel_for / parallel_for) is what is
exercised here; the typed fast paths are covered by test_gpu_parity.py."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests._inputs import lcg_uniform

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")
SUM_RTOL = 1e-12


@pytest.fixture(scope="module")
def prog():
    from onika_b200 import cxx
    return cxx.build_all()[0]


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def run(prog, tmp_path, scenario, *args, expect_rc=0):
    out = os.path.join(tmp_path, scenario + ".bin")
    r = subprocess.run([prog, scenario, out, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert r.returncode == expect_rc, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return np.fromfile(out, dtype=np.uint8)


def test_parallel_for_axpy_and_memset(prog, tmp_path, orc):
    for n in (1, 100003, 1_000_001):
        d = run(prog, tmp_path, "axpy", n).view(np.float64)
        x, y = lcg_uniform(n, 1, -1, 1), lcg_uniform(n, 2, -1, 1)
        orc.axpy(y, x, 0.5)
        assert np.array_equal(d[:n], y)
        assert np.array_equal(d[n:], np.full(n, 3.25))


def test_block_parallel_for_value_add_1d_2d_workstealing(prog, tmp_path, orc):
    rows, cols = 257, 129
    d = run(prog, tmp_path, "value_add", rows, cols).view(np.float64).reshape(3, rows, cols)
    a = lcg_uniform(rows * cols, 7).reshape(rows, cols)
    for k in range(3):
        orc.value_add(a, 1.1)
        assert np.array_equal(d[k], a), k


@pytest.mark.parametrize("mode", ["hybrid-lane", "hybrid-delayed-lane", "hybrid-sequence", "singletask-dependency"])
@pytest.mark.parametrize("kernel_1d", [1, 0])
def test_parallel_queue_lane_scenarios(prog, tmp_path, orc, mode, kernel_1d):
    rows, cols = (1024, 1024) if kernel_1d else (256, 192)
    if mode == "singletask-dependency" and not kernel_1d:
        pytest.skip("single-task space {64..980}^2 is written for the 1024^2 arrays")
    d = run(prog, tmp_path, "lanes", mode, rows, cols, kernel_1d).view(np.float64).reshape(2, rows, cols)
    a1, a2 = lcg_uniform(rows * cols, 8).reshape(rows, cols), lcg_uniform(rows * cols, 9).reshape(rows, cols)
    orc.lane_sequence(a1, a2, 1.1, 1.2, 1.3, 1.4)      # GPU kernel then host-only kernel in each lane, in order
    assert np.array_equal(d[0], a1) and np.array_equal(d[1], a2)


def test_automatic_lanes_from_declared_accesses(prog, tmp_path):
    """any_lane() + declared accesses -> lane DAG (parallel_execution_queue.h:pre_process_queue): independent arrays
    overlap on different lanes, an operation touching two arrays waits for both lanes, and the outcome equals serial FIFO
    execution (the reference serialises these operations: src/parallel/parallel_execution_queue.cpp:185-207)."""
    rows, cols, iters = 64, 4096, 2000
    d = run(prog, tmp_path, "autolanes", rows, cols, iters).view(np.float64)
    n = rows * cols
    A, B, C = lcg_uniform(n, 8), lcg_uniform(n, 9), lcg_uniform(n, 10)
    A += 1.1
    for _ in range(iters):
        B += 1.3
    A += 1.2
    B += A
    A += 1.4
    for _ in range(3):
        C += 0.7
    assert np.array_equal(d[:n], A) and np.array_equal(d[n:2 * n], B) and np.array_equal(d[2 * n:3 * n], C)
    lanes, waits = d[3 * n:3 * n + 6].astype(int), d[3 * n + 6:3 * n + 12].astype(int)
    assert lanes[0] == lanes[2] == lanes[3] == lanes[4]           # chain on A
    assert lanes[1] != lanes[0] and lanes[5] not in (lanes[0], lanes[1])
    assert list(waits) == [0, 0, 0, 1, 0, 0]                      # only B += A carries a cross-lane edge (to B's lane)


def test_automatic_lanes_order_write_only_and_read_only_declarations(prog, tmp_path):
    """one-sided declarations under any_lane(): a WO producer followed by an RO consumer (read-after-write), then RO
    consumers followed by a WO producer (write-after-read, readers on two lanes) -- results must equal serial FIFO execution.
    The reference's printed test concurrent_data_access() keeps its own answers (no 'conflict' for writer/reader)."""
    rows, cols, iters = 64, 4096, 2000
    d = run(prog, tmp_path, "autolanes_raw_war", rows, cols, iters).view(np.float64)
    n = rows * cols
    A, B, C, D = lcg_uniform(n, 8), lcg_uniform(n, 9), lcg_uniform(n, 10), lcg_uniform(n, 11)
    for _ in range(iters):
        A += 1.1
    B += A
    for _ in range(iters):
        C += 1.3
    C += A
    A += 1.4
    D += A
    assert np.array_equal(d[:n], A) and np.array_equal(d[n:2 * n], B) and np.array_equal(d[2 * n:3 * n], C) and np.array_equal(d[3 * n:4 * n], D)
    lanes, waits = d[4 * n:4 * n + 6].astype(int), d[4 * n + 6:4 * n + 12].astype(int)
    assert lanes[1] == lanes[0]                                   # the consumer follows its producer's lane
    assert lanes[2] != lanes[0]                                   # the independent slow kernel overlaps
    assert waits[3] >= 1 and waits[4] >= 1                        # C += A waits for A's lane; the overwrite waits for the reader on C's lane
    assert list(d[4 * n + 12:4 * n + 16]) == [0.0, 1.0, 1.0, 0.0]  # reference rule (printed) vs ordering rule


def test_automatic_lanes_overlap_like_explicit_lanes(prog, tmp_path):
    """few-block kernels on two arrays: any_lane()+accesses must overlap them like explicit lanes 0/1 do (and both beat one
    lane); results stay exact: 9 sequential runs of (k1,k2) per array"""
    rows, cols, iters = 4, 8192, 20000
    d = run(prog, tmp_path, "autolanes_perf", rows, cols, iters).view(np.float64)
    one_lane, explicit, auto = d[:3]
    assert auto < 0.75 * one_lane and auto < 1.25 * explicit, (one_lane, explicit, auto)
    n = rows * cols
    A, B = lcg_uniform(n, 8), lcg_uniform(n, 9)
    for _ in range(9):
        for v in (1.1, 1.2):
            for _ in range(iters):
                A += v
        for v in (1.3, 1.4):
            for _ in range(iters):
                B += v
    assert np.array_equal(d[3:3 + n], A) and np.array_equal(d[3 + n:], B)


def test_queue_capture_and_replay(prog, tmp_path):
    """queue << capture(g) ... << flush records the lane DAG (GPU kernels, a host-only functor as a host node, the
    return-data bracket of a reduction) into one CUDA graph; every replay must equal one more serial pass"""
    rows, cols, replays, nsmall = 128, 512, 3, 200
    d = run(prog, tmp_path, "graph", rows, cols, replays, nsmall).view(np.float64)
    n = rows * cols
    A, B = lcg_uniform(n, 8), lcg_uniform(n, 9)
    vmin, vmax, vsum, cnt = np.finfo(np.float64).max, -np.finfo(np.float64).max, 0.0, 0
    for _ in range(replays):
        A += 1.1
        B += 1.3
        A += 1.2
        B += 1.4
        vmin, vmax, vsum, cnt = min(vmin, A.min()), max(vmax, A.max()), vsum + float(np.sum(A)), cnt + n
    assert np.array_equal(d[:n], A) and np.array_equal(d[n:2 * n], B)
    r = d[2 * n:2 * n + 6]
    assert r[0] == vmin and r[1] == vmax and r[3] == cnt          # accumulated through the return-data channel, replay after replay
    assert abs(r[2] - vsum) <= SUM_RTOL * abs(vsum)
    assert r[4] == 5 and r[5] == 2                                # 5 operations on 2 lanes (arrays A and B)
    direct_ms, graph_ms = d[2 * n + 6:2 * n + 8]
    small = d[2 * n + 8:]
    assert np.array_equal(small, np.full(8 * 256, 9.0 * nsmall))
    assert graph_ms < direct_ms, (direct_ms, graph_ms)


def test_cell_sums_through_the_generic_functor_path(prog, tmp_path, orc):
    """config C4 written the application way (one block per cell, ONIKA_CU_BLOCK_SIMD_FOR + block_reduce_add) with 128- and
    32-thread blocks, next to the typed field-major kernel on the same column: typed sums are the in-order sums bit for
    bit, the functor's tree sums agree within rounding"""
    ncells = 20000
    d = run(prog, tmp_path, "cells_generic_perf", ncells).view(np.float64)
    sizes = np.floor(lcg_uniform(ncells, 42, 8.0, 25.0)).astype(np.uint64)
    begin = np.zeros(ncells + 1, np.uint64)
    begin[1:] = np.cumsum(sizes, dtype=np.uint64)
    values = lcg_uniform(int(begin[-1]), 40, -1.0, 1.0)
    want, _ = orc.cell_sum_packed(values, begin)
    typed, g128, g32 = (d[6 + k * ncells:6 + (k + 1) * ncells] for k in range(3))
    assert np.array_equal(typed, want)
    mag = np.add.reduceat(np.abs(values), begin[:-1].astype(np.int64))
    assert np.all(np.abs(g128 - want) <= SUM_RTOL * mag) and np.all(np.abs(g32 - want) <= SUM_RTOL * mag)
    # sub-block items (several cells per CTA, the same functor source): every cell written, same sums within rounding
    for k in range(3, 6):
        sub = d[6 + k * ncells:6 + (k + 1) * ncells]
        assert sub.size == ncells and not np.isnan(sub).any()
        assert np.all(np.abs(sub - want) <= SUM_RTOL * mag)
    tail = d[6 + 6 * ncells:]                      # sub-blocks of 4 and 2 threads: two timings, then their sums
    assert tail.size == 2 + 2 * ncells
    for k in range(2):
        sub = tail[2 + k * ncells:2 + (k + 1) * ncells]
        assert not np.isnan(sub).any() and np.all(np.abs(sub - want) <= SUM_RTOL * mag)


def test_3d_spaces(prog, tmp_path, orc, gold):
    d = run(prog, tmp_path, "grid3d").view(np.float64)
    assert np.array_equal(d[:64], np.ones(64))          # block_parallel_for over {0..4}^3: every cell once
    m = np.zeros(16 ** 3)
    orc.parallel_for_3d(m, 16, (2, 3, 3), (15, 13, 13))
    assert np.array_equal(d[64:], m)
    assert [int(i) for i in np.nonzero(d[64:])[0]] == gold["parallel_for_3d"]["nonzero"]


def test_return_data_reduction_and_callback(prog, tmp_path, orc):
    rows, cols = 777, 1001
    r = run(prog, tmp_path, "reduce", rows, cols).view(np.float64)
    d = lcg_uniform(rows * cols, 17, -1.0, 1.0)
    assert r[0] == orc.reduce_min(d) and r[1] == orc.reduce_max(d)
    assert abs(r[2] - orc.reduce_sum_ld(d)) <= SUM_RTOL * np.abs(d).sum()
    assert r[3] == rows * cols and r[4] == 1.0


def test_cell_grid_of_field_arrays_generic_path(prog, tmp_path, orc):
    ncells = 4096
    raw = run(prog, tmp_path, "cells", ncells)
    hdr = raw[:24].view(np.uint64)
    body = raw[24:].view(np.float64)
    counts = body[:ncells].astype(np.uint32)
    sums, total = body[ncells:2 * ncells], body[2 * ncells]
    assert hdr[0] == ncells and hdr[1] == counts.sum()
    nbytes, cap, off = orc.cell_grid_layout(counts, 64, 16, [8, 8, 8, 8], 64)
    assert hdr[2] == nbytes                               # same bytes as one reference FieldArrays per cell
    e = lcg_uniform(int(counts.sum()), 40, -1, 1)
    starts = np.concatenate([[0], np.cumsum(counts.astype(np.int64))[:-1]])
    want = np.array([np.add.reduce(e[s:s + c]) if c else 0.0 for s, c in zip(starts, counts)])
    seq = np.zeros(ncells)
    for c in range(ncells):                               # strict particle order, like the functor
        s = 0.0
        for v in e[starts[c]:starts[c] + counts[c]]:
            s += v
        seq[c] = s
    assert np.array_equal(sums, seq)
    assert abs(total - seq.sum()) <= SUM_RTOL * np.abs(e).sum()
    assert np.allclose(want, seq)


def test_element_list_space_prolog_epilog(prog, tmp_path):
    r = run(prog, tmp_path, "listed").view(np.float64)
    d = lcg_uniform(1000, 3)
    d[::3] *= 2.0
    assert np.array_equal(r[:1000], d)
    assert r[1000] == 1.0 and r[1001] == 1.0


@pytest.mark.parametrize("which", ["typed", "generic"])
def test_parallel_reduce_helpers_on_lanes(tmp_path, orc, which):
    """onika/parallel/reduction.h: reductions queued behind the kernels producing their input, two lanes; typed kernel
    (onk_reduce_f64) and, with the fast paths compiled out, the generic one-block functor body."""
    import sys
    from onika_b200 import cxx
    exe = cxx.program_path("frontend_test" if which == "typed" else "frontend_test_generic")
    if not os.path.exists(exe):
        cxx.build_all()
    rows, cols = 513, 1027
    r = run(exe, tmp_path, "reduce_helpers", rows, cols).view(np.float64)
    a1 = lcg_uniform(rows * cols, 8, -1.0, 1.0) + 1.1
    a2 = lcg_uniform(rows * cols, 9, -1.0, 1.0) + 1.3
    for k, a in ((0, a1), (3, a2)):
        assert r[k] == orc.reduce_min(a) and r[k + 1] == orc.reduce_max(a)
        assert abs(r[k + 2] - orc.reduce_sum_ld(a)) <= SUM_RTOL * np.abs(a).sum()
    assert r[6] == sys.float_info.max and r[7] == -sys.float_info.max and r[8] == 0.0


def test_cuda_helper_headers_memcpy_memmove_spans_math(prog, tmp_path):
    rows, rb = 333, 1000
    raw = run(prog, tmp_path, "helpers", rows)
    n = rows * rb
    dst, mv, res = raw[:n], raw[n:2 * n], raw[2 * n:].view(np.float64)
    i = np.arange(n, dtype=np.uint64)
    src = ((i * 131 + 7) >> 3).astype(np.uint8)
    mv0 = ((i * 29 + 1) >> 2).astype(np.uint8)
    want_dst = np.full(n, 0xEE, np.uint8)
    want_mv = mv0.copy()
    for r in range(rows):
        want_dst[r * rb + r % 7: r * rb + r % 7 + rb - 8] = src[r * rb + r % 5: r * rb + r % 5 + rb - 8]
        want_mv[r * rb + 3: r * rb + 3 + rb - 8] = mv0[r * rb: r * rb + rb - 8]
    assert np.array_equal(dst, want_dst) and np.array_equal(mv, want_mv)
    c = [0.5, -0.25, 0.125, 1.5, -2.0]
    table = lcg_uniform(64, 4, -1.0, 1.0)
    want = np.zeros(rows)
    for r in range(rows):
        acc = 0.0
        for k, ck in enumerate(c):
            acc += ck * table[(r + k) % 64]
        want[r] = min(max(acc, -1.0), 1.0) + 2.0 ** (r % 5)
    assert np.array_equal(res, want * 2.0)        # doubled by a user kernel launched with ONIKA_CU_LAUNCH_KERNEL on lane 2


def test_soatl_container_layout_serialize_and_device_apply(prog, tmp_path, orc, gold):
    n = 13
    raw = run(prog, tmp_path, "soatl", n)
    u = raw[: 6 * 8 * 8].view(np.uint64).reshape(6, 8)
    case = [c for c in gold["soatl_layout"] if c["cfg"] == 3 and c["sizes"] == [17, 33, 10, 0, 100, 40]][0]
    for k in range(6):
        assert u[k, 0] == case["capacity"][k] and u[k, 1] == case["storage_bytes"][k]
        assert list(u[k, 2:]) == case["offsets"][k]
    pos = 6 * 8 * 8
    wire_size = int(raw[pos:pos + 8].view(np.uint64)[0])
    pos += 8
    g = gold["soatl_serialize"]
    assert raw[pos:pos + wire_size].tobytes().hex() == g["packed_hex"]      # byte-identical to the reference's wire format
    pos += wire_size
    e, e_named = raw[pos:pos + 8 * n].view(np.float64), raw[pos + 8 * n:].view(np.float64)
    want = orc.soatl_dist(lcg_uniform(n, 10, 0, 1), lcg_uniform(n, 11, 0, 1), lcg_uniform(n, 12, 0, 1))
    assert np.isnan(e[0]) and np.array_equal(e, want, equal_nan=True) and np.array_equal(e_named, want, equal_nan=True)
    # a larger sweep of the device lambda
    n = 100_003
    raw = run(prog, tmp_path, "soatl", n)
    e, e_named = raw[-16 * n:-8 * n].view(np.float64), raw[-8 * n:].view(np.float64)
    want = orc.soatl_dist(lcg_uniform(n, 10, 0, 1), lcg_uniform(n, 11, 0, 1), lcg_uniform(n, 12, 0, 1))
    assert np.array_equal(e, want, equal_nan=True)              # extended lambda on the device
    assert np.array_equal(e_named, want, equal_nan=True)        # named functor type with ApplyFunctorTraits::CudaCompatible


@pytest.mark.parametrize("n", [1, 1023, 1024, 100_003, 1 << 20])
def test_soatl_parallel_apply_simd_vectorised_device_kernel(prog, tmp_path, orc, n):
    """both C3 operators through the generic vectorised kernel of soatl/compute.h (256-bit accesses, 4 elements per thread,
    scalar tail): the 8-field read-modify-write sweep and the tests/benchmark.cpp operator, bit for bit against the oracle"""
    d = run(prog, tmp_path, "soatl_apply", n).view(np.float64)
    assert d.size == 9 * n
    s, c = 1.0000001, np.array([0.125 * (k + 1) for k in range(8)])
    for k in range(8):
        col = lcg_uniform(n, 20 + k, 0.0, 1.0)
        assert np.array_equal(d[k * n:(k + 1) * n], col * s + c[k]), k       # one rounded multiply, one rounded add (oracle rule)
    # the same through the oracle's packed-storage sweep (orc_soatl_rmw_f64) for one layout
    cap = ((n + 15) // 16) * 16
    stride = ((cap * 8 + 255) // 256) * 256
    st = np.zeros(8 * stride, dtype=np.uint8)
    for k in range(8):
        st[k * stride:k * stride + 8 * n].view(np.float64)[:] = lcg_uniform(n, 20 + k, 0.0, 1.0)
    orc.soatl_rmw(st, 256, 16, 8, n, cap, s, c)
    for k in range(8):
        assert np.array_equal(d[k * n:(k + 1) * n], st[k * stride:k * stride + 8 * n].view(np.float64)), k
    want = orc.soatl_dist(lcg_uniform(n, 10, 0, 1), lcg_uniform(n, 11, 0, 1), lcg_uniform(n, 12, 0, 1))
    assert np.array_equal(d[8 * n:], want, equal_nan=True)


def test_block_vocabulary_in_ordinary_ctas_and_in_ctas_of_sub_blocks(prog, tmp_path):
    """ONIKA_CU_THREAD_IDX / BLOCK_SIZE / BLOCK_DIMS / BLOCK_IDX / GRID_SIZE / BLOCK_SIMD_FOR / BLOCK_SYNC[_OR|_AND] and
    block_reduce_add/min/max mean the functor's BLOCK whether a CTA holds one block (reference launch shape) or several
    sub-blocks (SubBlockItems): same functor source, same per-item answers"""
    nitems = 1000
    d = run(prog, tmp_path, "block_vocabulary", nitems).view(np.float64).reshape(7, nitems, 8)
    values = lcg_uniform(nitems * 128, 21, -1.0, 1.0)
    for k, bs in enumerate((8, 8, 16, 16, 32, 32, 128)):
        r = d[k]
        v = values[: nitems * bs].reshape(nitems, bs)                                           # item i reads values[i*bs + tid]
        assert np.all(r[:, 0] == bs) and np.all(r[:, 7] == bs), (k, bs)                         # block size and dims of the (sub-)block
        assert np.all(r[:, 1] == bs * (bs + 1) / 2)                                             # every thread index 0..bs-1 exactly once
        assert np.array_equal(r[:, 2], v.min(axis=1)) and np.array_equal(r[:, 3], v.max(axis=1))
        assert np.all(r[:, 4] == 3.0)                                                            # OR: yes, AND: yes, OR of nothing: no
        assert np.all(r[:, 5] == 3 * bs + 1) and np.all(r[:, 6] == 1.0)
    assert np.array_equal(d[0], d[1]) and np.array_equal(d[2], d[3]) and np.array_equal(d[4], d[5])   # packed == unpacked, bit for bit


def _multi_gpu_expected(orc, n, world):
    """what every rank of tests/cxx/multi_gpu_test must write: fused min/max/sum of the sharded array (twice), then the
    all_reduce_multi results -- rank-order folds from the identity"""
    d = lcg_uniform(n, 7, -1.0, 1.0)
    a = 0.0
    f = np.float64(0.0)
    cnt = 0.0
    for r in range(world):
        a = a + 0.5 * (r + 1)
        f = f + np.float64(np.float32(1.25) * np.float32(r + 1))
        cnt = cnt + (10 + r)
    return d, [a, float(np.float32(f)), float(int(cnt)), -3.0, -3.0 + (world - 1)]


def run_multi_gpu_program(tmp_path, world, n, transport=None):
    from onika_b200 import cxx
    exe = cxx.program_path("multi_gpu_test")
    if not os.path.exists(exe):
        cxx.build_all()
    rdv = os.path.join(tmp_path, f"rdv{world}_{n}_{transport}")
    os.makedirs(rdv, exist_ok=True)
    procs = [subprocess.Popen([exe, str(r), str(world), rdv, str(n), os.path.join(rdv, f"out{r}.bin")] + ([str(transport)] if transport is not None else []),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE)
             for r in range(world)]
    outs = []
    for r, p in enumerate(procs):
        so, se = p.communicate(timeout=300)
        assert p.returncode == 0, (r, p.returncode, so[-1000:], se[-1000:])
        outs.append(np.fromfile(os.path.join(rdv, f"out{r}.bin"), dtype=np.float64))
    return outs


def check_multi_gpu_outputs(outs, orc, n, world):
    d, multi = _multi_gpu_expected(orc, n, world)
    mag = np.abs(d).sum()
    for o in outs:
        assert o.size == 12
        for rep in range(2):
            mn, mx, sm = o[3 * rep:3 * rep + 3]
            assert mn == orc.reduce_min(d) and mx == orc.reduce_max(d)
            assert abs(sm - orc.reduce_sum_ld(d)) <= SUM_RTOL * mag
        assert list(o[0:3]) == list(o[3:6])                         # reproducible
        assert list(o[6:11]) == multi and o[11] == 1.0
        assert np.array_equal(o, outs[0])                           # same bits on every rank


def test_cxx_multi_gpu_helper_world_of_one(tmp_path, orc):
    """onika/parallel/multi_gpu.h (all_reduce_multi in the reference's call shape, fused reduce + exchange) with one rank:
    the whole protocol runs, the folds start from the identity"""
    n = 1_000_003
    check_multi_gpu_outputs(run_multi_gpu_program(tmp_path, 1, n), orc, n, 1)
    check_multi_gpu_outputs(run_multi_gpu_program(tmp_path, 1, n, transport=1), orc, n, 1)     # buffers in host shared memory


@pytest.mark.parametrize("n", [1, 7, 1000, 1024, 4099, 200_003])
def test_soatl_apply_mixed_types_device_equals_host_loop(prog, tmp_path, n):
    """vectorised apply kernel with fp64 / u8 / i32 / fp32 columns side by side (32-, 4-, 16-, 16-byte accesses per thread),
    read-only columns by value and by const reference, a sub-range off the vector alignment, sizes below one tile: the device
    sweep leaves exactly the bytes the host loop (the reference's sequential semantics) leaves"""
    raw = run(prog, tmp_path, "soatl_apply_mixed", n)
    pos = 0
    for width in (8, 1, 4, 4):
        dev, host = raw[pos:pos + n * width], raw[pos + n * width:pos + 2 * n * width]
        assert np.array_equal(dev, host), width
        pos += 2 * n * width
    assert pos == raw.size
    x = raw[:n * 8].view(np.float64)
    t = np.floor(lcg_uniform(n, 51, 0.0, 200.0)).astype(np.uint8)
    assert np.array_equal(x, lcg_uniform(n, 50, -1.0, 1.0) * 0.25 + t.astype(np.float64))      # and what the operator says


def test_sub_block_items_over_element_lists_and_tiny_spaces(prog, tmp_path, orc):
    """SubBlockItems with an indexed 1-D space (every other cell), with fewer items than a CTA holds sub-blocks, with empty
    cells and with an empty space: listed cells get their sums, the others stay untouched"""
    ncells = 1001
    d = run(prog, tmp_path, "sub_block_listed", ncells).view(np.float64)
    sizes = np.floor(lcg_uniform(ncells, 42, 0.0, 25.0)).astype(np.uint64)
    begin = np.zeros(ncells + 1, np.uint64)
    begin[1:] = np.cumsum(sizes, dtype=np.uint64)
    values = lcg_uniform(int(begin[-1]), 40, -1.0, 1.0)
    want, _ = orc.cell_sum_packed(values, begin)
    mag = np.add.reduceat(np.abs(np.append(values, 0.0)), np.minimum(begin[:-1].astype(np.int64), values.size)) * (sizes > 0) + 1e-300
    listed, few, ws = d[:ncells], d[ncells:2 * ncells], d[2 * ncells:]
    assert ws.size == ncells and np.all(np.abs(ws - want) <= SUM_RTOL * mag)       # fixed grid + work stealing, sub-blocks
    assert np.all(listed[1::2] == -7.0) and np.all(np.abs(listed[0::2] - want[0::2]) <= SUM_RTOL * mag[0::2])
    assert np.all(few[5:] == -7.0) and np.all(np.abs(few[:5] - want[:5]) <= SUM_RTOL * mag[:5])
    assert np.all(listed[0::2][sizes[0::2] == 0] == 0.0)           # empty cells: exactly zero


@pytest.mark.parametrize("dims", [(13, 11, 7), (40, 33, 20), (2, 2, 3)])
def test_sub_block_items_over_2d_and_3d_boxes(prog, tmp_path, orc, dims):
    """SubBlockItems on 2D/3D execution spaces (reference trampoline: block_parallel_for_adapter.h:95-104, one configured
    8x8x4 tile per coordinate): with opts.max_block_size <= 32 a CTA holds 256/bs coordinates, each run by a 1-D sub-block.
    Sub-boxes with non-zero starts; cells outside the box stay untouched; in-cell sums (<= 25 values through a butterfly or a
    block tree) against the oracle's particle-order sums within SUM_RTOL; the block size the functor saw is added to every
    result (16 packed, 256 / 64 unpacked); the trait without a small max_block_size keeps the configured tile"""
    nx, ny, nz = dims
    ncells = nx * ny * nz
    d = run(prog, tmp_path, "sub_block_box", nx, ny, nz).view(np.float64)
    assert d.size == 5 * ncells
    sizes = np.floor(lcg_uniform(ncells, 42, 0.0, 25.0)).astype(np.uint64)
    begin = np.zeros(ncells + 1, np.uint64)
    begin[1:] = np.cumsum(sizes, dtype=np.uint64)
    values = lcg_uniform(int(begin[-1]), 40, -1.0, 1.0)
    want, _ = orc.cell_sum_packed(values, begin)
    mag = np.add.reduceat(np.abs(np.append(values, 0.0)), np.minimum(begin[:-1].astype(np.int64), values.size)) * (sizes > 0) + 1e-300
    want = want.reshape(nz, ny, nx)
    mag = mag.reshape(nz, ny, nx)
    in3 = np.zeros((nz, ny, nx), bool)
    in3[2:nz, 0:ny - 1, 1:nx] = True
    in2 = np.zeros((nz, ny, nx), bool)
    in2[0, 1:ny, 0:nx - 1] = True
    for k, (inside, bs) in enumerate([(in3, 256.0), (in3, 16.0), (in2, 64.0), (in2, 16.0), (in3, 256.0)]):
        got = d[k * ncells:(k + 1) * ncells].reshape(nz, ny, nx)
        assert np.all(got[~inside] == -7.0), k
        assert np.all(np.abs(got[inside] - bs - want[inside]) <= SUM_RTOL * (mag[inside] + bs)), k
    # packed and unpacked runs agree to the same tolerance; empty cells give exactly the block size
    packed3 = d[ncells:2 * ncells].reshape(nz, ny, nx)
    assert np.all(packed3[in3 & (sizes.reshape(nz, ny, nx) == 0)] == 16.0)


def test_host_only_functor_sees_the_openmp_team_as_its_grid(prog, tmp_path):
    """a CudaCompatible = false functor with per-block accumulators indexed by ONIKA_CU_BLOCK_IDX (reference meaning on the host:
    cuda.h:221-224, grid = OpenMP team, block index = thread rank): the unsynchronised `+=` per block adds up exactly"""
    env = dict(os.environ, OMP_NUM_THREADS="6")
    out = os.path.join(tmp_path, "hbs.bin")
    r = subprocess.run([prog, "host_block_scratch", out, "200000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-800:], r.stderr[-800:])
    d = np.fromfile(out, dtype=np.float64)
    total, expect, team, used, bad = d[:5]
    assert bad == 0 and total == expect
    assert team == 6 and 1 <= used <= 6          # the lane's worker thread runs the region with the caller's OpenMP settings
    # opts.omp_scheduling = OMP_SCHED_STATIC (reference: block_parallel_for_adapter.h:296-303): thread t owns items [t*ceil(n/T), ...)
    owner = d[5:]
    assert owner.size == 4096 and np.all(np.diff(owner) >= 0) and owner[0] == 0 and owner[-1] == 5      # contiguous runs, in thread order
    assert set(np.bincount(owner.astype(np.int64)).tolist()) <= {682, 683}                                # of near-equal length


def test_host_compiler_and_nvcc_translation_units_share_one_context():
    """a plugin = .cpp files compiled by g++ plus .cu files compiled by nvcc: the default CUDA context is created by the g++
    half and read by the nvcc half (tests/cxx/mixed_tu_*): every shared type has one layout, the SM count that sizes fixed
    grids is the device's"""
    from onika_b200 import cxx
    if not os.path.exists(cxx.MIXED_TU_BIN):
        cxx.build_mixed_tu_test()
    r = subprocess.run([cxx.MIXED_TU_BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-1000:], r.stderr[-1000:])
    assert "mixed TU ok" in r.stdout


# ---------------------------------------------------------------- the reference's own tutorial plugins, compiled unchanged
def _refplugins():
    from onika_b200 import cxx
    if not os.path.exists(cxx.REFPLUGIN_BIN):
        pytest.skip("tests/cxx/_refplugins/run_reference_tutorial not built (needs /root/reference at build time)")
    return cxx.REFPLUGIN_BIN


def _run_op(tmp_path, op, *args):
    exe = _refplugins()
    out = os.path.join(tmp_path, op + ".bin")
    dump = [a.replace("@OUT", out) for a in args]
    r = subprocess.run([exe, op, *dump], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    return out, r.stdout


def _load_array(path):
    raw = np.fromfile(path, dtype=np.uint8)
    rows, cols = (int(v) for v in raw[:16].view(np.uint64))
    return raw[16:].view(np.float64).reshape(rows, cols)


def _chain(x, v, iters):
    """BlockParallelIterativeValueAddFunctor (extras/block_parallel_value_add_functor.h:60-93) in numpy"""
    y = np.power(x, np.sin(x))
    for _ in range(iters):
        x = x + y + v
        y = np.power(x, np.sin(x))
    return x * y


def test_reference_plugin_synchronous_and_async_block_parallel(tmp_path):
    """plugins/onikaParallelTutorial/{synchronous,async}_block_parallel.cu, unchanged sources"""
    a = lcg_uniform(300 * 200, 8).reshape(300, 200)
    for op in ("synchronous_block_parallel", "async_block_parallel"):
        out, _ = _run_op(tmp_path, op, "fill:my_array=300x200", "my_value=1.1", "dump:my_array=@OUT")
        assert np.array_equal(_load_array(out), a + 1.1)


@pytest.mark.parametrize("op", ["concurrent_block_parallel", "hybrid_async_block_parallel", "automatic_concurrent_parallel"])
def test_reference_plugin_lane_operators(tmp_path, op):
    """two arrays x (iterative kernel, then value-add kernel) on lanes 0/1; the hybrid variant runs kernel 2 on the host"""
    a = lcg_uniform(128 * 96, 8).reshape(128, 96)
    v, iters = 0.25, 3
    out, stdout = _run_op(tmp_path, op, "fill:array1=128x96", "fill:array2=128x96", f"value={v}", f"iter_count={iters}",
                          "dump:array1=@OUT")
    want = _chain(a, v, iters) + v
    got = _load_array(out)
    # pow/sin come from libm on the host and from libdevice on the GPU: last-ulp differences, not bit parity
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    assert "All operations have terminated" in stdout


def test_reference_plugin_parallel_for_benchmarks(tmp_path):
    _, out = _run_op(tmp_path, "parallel_for_benchmark", "samples=4096", "iterations=8")
    assert "Benchmark passed" in out
    _, out = _run_op(tmp_path, "parallel_for_3d_benchmark", "pfor3d_side=16", "pfor3d_block_side=4")
    # the tutorial prints the 16^3 coverage mask of parallel_for over {2,3,3}..{15,13,13}: 13*10*10 ones
    body = out.split("parallel_for 3D test")[-1]
    ones = sum(line.split(":")[1].split().count("1") for line in body.splitlines() if " J=" in line)
    assert ones == 13 * 10 * 10


MSP_DIR = os.path.join(os.path.dirname(__file__), "cxx", "msp")


def _run_msp(tmp_path, msp, *args):
    exe = _refplugins()
    out = os.path.join(tmp_path, os.path.basename(msp) + ".bin")
    r = subprocess.run([exe, "--msp", msp, *[a.replace("@OUT", out) for a in args]], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    return out, r.stdout


def test_reference_plugins_from_msp_batch_connected_slots(tmp_path):
    """.msp flat batch: three unchanged tutorial operators in a task group, `my_array` connected by name"""
    out, stdout = _run_msp(tmp_path, os.path.join(MSP_DIR, "tutorial_chain.msp"), "dump:my_array=@OUT")
    assert "3 operator(s), task_group" in stdout
    want = np.zeros((1024, 1024))            # first operator allocates 1024x1024 zeros (synchronous_block_parallel.cu:28-31)
    for v in (1.1, 1.2, 1.3):
        want += v
    assert np.array_equal(_load_array(out), want)


def test_reference_plugins_from_msp_example_shape(tmp_path):
    """the layout of data/exemples/concurrent_block_parallel.msp (comments, flow maps, a second operator re-using the
    arrays of the first through the connected array1/array2 slots)"""
    out, stdout = _run_msp(tmp_path, os.path.join(MSP_DIR, "concurrent_lanes.msp"), "lanes_then_auto", "dump:array2=@OUT")
    assert stdout.count("All operations have terminated") == 2
    x = _chain(np.zeros((4, 256)), 2.3, 2) + 2.3           # concurrent_block_parallel
    x = _chain(x, 0.5, 1) + 0.5                            # automatic_concurrent_parallel on the same arrays
    assert np.allclose(_load_array(out), x, rtol=1e-10, atol=0)     # pow/sin: libm vs libdevice, few iterations (the map is chaotic)


def _reference_msp(*parts):
    from onika_b200 import cxx
    path = os.path.join(cxx.REFERENCE_MSP_DIR, *parts)
    if not os.path.exists(path):
        pytest.skip(f"{path} not staged (needs /root/reference at build time)")
    return path


def _run_app(tmp_path, *args):
    exe = _refplugins()
    out = os.path.join(tmp_path, "app.bin")
    r = subprocess.run([exe, "--app", *[a.replace("@OUT2", out + "2").replace("@OUT", out) for a in args]], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.returncode, r.stdout[-2500:], r.stderr[-1500:])
    return out, r.stdout


def test_reference_application_chain_with_nested_batches(tmp_path):
    """the reference's own data/config/main-config.msp (unchanged) as the base document -- its simulation chain
    message / mpi_comm_world / init_cuda / global / unit_system / init / run / finalize_cuda -- with a user file layered on
    top that turns `run` into nested batches of unchanged tutorial plugins; `my_array` is connected through all of them"""
    out, stdout = _run_app(tmp_path, "--config", _reference_msp("config", "main-config.msp"), os.path.join(MSP_DIR, "nested_app.msp"), "dump:my_array=@OUT")
    graph = [ln.strip() for ln in stdout.split("simulation graph:")[1].splitlines() if ln.strip()]
    names = [g.split()[0] for g in graph[:14]]
    assert names == ["simulation", "message", "mpi_comm_world", "init_cuda", "global", "unit_system", "nop", "my_run", "first_stage",
                     "synchronous_block_parallel", "second_stage", "async_block_parallel", "synchronous_block_parallel", "finalize_cuda"]
    assert "[batch nested_run]" in stdout and "first_stage [batch first_stage, task_group]" in stdout
    assert "init_cuda:" in stdout and "SMs" in stdout and stdout.rstrip().endswith("finalize_cuda")
    want = np.zeros((1024, 1024))
    for v in (1.1, 1.2, 1.3):
        want += v
    assert np.array_equal(_load_array(out), want)


def test_reference_example_msp_unchanged(tmp_path):
    """the reference's own example input, data/exemples/concurrent_block_parallel.msp layered on data/config/main-config.msp,
    run the way the reference's application runs it (`onika-exec <file> --run run_test`, apps/onika-exec.cpp:21-30): the
    task-group batch holds the unchanged concurrent_block_parallel plugin with the file's parameters (4 x 65536 elements,
    65536 iterations of the pow/sin map on two lanes).  The map is chaotic over that many iterations, so the arrays are
    checked for what must hold exactly: shape, both arrays identical (same operations from the same zeros), finite, and the
    last value-add applied; bit parity of the same sequence with short iteration counts is test_parallel_queue_lane_scenarios"""
    cfg, ex = _reference_msp("config", "main-config.msp"), _reference_msp("exemples", "concurrent_block_parallel.msp")
    out1, stdout = _run_app(tmp_path, "--config", cfg, ex, "--run", "run_test", "dump:array1=@OUT", "dump:array2=@OUT2")
    assert "run_test [batch concurrent_block_parallel_test, task_group]" in stdout and "All operations have terminated" in stdout
    a1, a2 = _load_array(out1), _load_array(out1 + "2")
    assert a1.shape == (4, 65536) and np.array_equal(a1, a2) and np.isfinite(a1).all()
    assert np.all(a1 == a1[0, 0])                                 # every element went through the same map from 0.0
    # without the override `run` stays the reference's `nop`: the chain runs, the example batch does not
    _, stdout = _run_app(tmp_path, "--config", cfg, ex)
    assert "All operations have terminated" not in stdout and stdout.rstrip().endswith("finalize_cuda")


def _run_reference_test(name, stdin):
    from onika_b200 import cxx
    exe = cxx.reference_test_path(name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe], input=stdin, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    return r.stdout


@pytest.mark.parametrize("uvm", [1, 0])
def test_reference_gpu_uvm_benchmark_unchanged(uvm):
    """tests/gpu_uvm_benchmark.cu compiled unchanged: managed or explicit device memory, a kernel launched with
    ONIKA_CU_LAUNCH_KERNEL iterating x <- x*x - 0.25 ten million times (converges to (1 - sqrt 2)/2 from [0,1))"""
    out = _run_reference_test("gpu_uvm_benchmark", f"0 {uvm} 5\n")
    line = [ln for ln in out.splitlines() if ln.startswith("result[5]")][0]
    vhost, vcuda = (float(v) for v in line.split("=")[1].split("/"))
    assert abs(vhost - 5.0 / (1024 * 1024)) < 1e-9            # printed with 6 significant digits
    assert abs(vcuda - (1.0 - 2.0 ** 0.5) / 2.0) < 1e-6


def test_reference_gpu_reduce_min_fp64_unchanged():
    """tests/gpu_reduce_min_fp64.cu compiled unchanged: its CPU reduction (the oracle's restated fixture, minimum 0 at i = 0)
    runs, its GPU part is the reference's TODO; device allocation / copy / synchronise macros are exercised"""
    out = _run_reference_test("gpu_reduce_min_fp64", "1 0 0\n")
    assert "GPU reduction not implemented yet" in out
    line = [ln for ln in out.splitlines() if ln.startswith("result =")][0]
    assert [float(v) for v in line.split("=")[1].split("/")] == [0.0, 0.0]


@pytest.mark.parametrize("mode", ["hybrid-lane", "hybrid-delayed-lane", "hybrid-sequence", "singletask-dependency"])
@pytest.mark.parametrize("omp_mode", ["task", "pfor"])
@pytest.mark.parametrize("dims", ["1d", "2d"])
def test_reference_parallel_queue_lane_unchanged(mode, omp_mode, dims):
    """tests/parallel_queue_lane.cu compiled unchanged (4 queue modes x OpenMP task / parallel-for context x 1D / 2D kernels):
    the program only prints (the reference checks nothing); results of the same scenarios are verified bit for bit by
    test_parallel_queue_lane_scenarios above"""
    from onika_b200 import cxx
    exe = cxx.reference_test_path("parallel_queue_lane")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe, mode, omp_mode, dims], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
    assert "Synchronize parallel operations" in r.stdout and r.stdout.rstrip().endswith("done")
    assert ("1D" if dims == "1d" else "2D") + " kernel parallelization space" in r.stdout


@pytest.mark.parametrize("name,args", [("soatupletest", ["10"]), ("soatupletest", ["1000", "3"]), ("soatlserializetest", ["10000"]), ("soatlserializetest", ["7"])])
def test_reference_soatl_tests_unchanged(tmp_path, name, args):
    """tests/soatupletest.cpp and tests/soatlserializetest.cpp compiled unchanged (their own declare_fields.h, asserts on):
    find_index_of_id static_asserts, tuple copies, clear() -> null pointers, copy variants and the storage_ptr() file round
    trip with exact equality (CMakeLists.txt registers them with 10 and 10000)"""
    from onika_b200 import cxx
    exe = cxx.reference_test_path(name)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300, cwd=tmp_path)     # serialize.dat lands in cwd
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-1500:])
