"""Multi-GPU parity checks of the hot path against the oracle, callable from any process group (one process per GPU):
tests/test_gpu_multi.py spawns two workers around it, bench.py runs it at N > 1 outside every timed region and reports the
outcome under `checks` -- the driver's test box has one GPU, so this is where the multi-GPU paths are pinned on hardware.

Every check computes the expected value with the oracle on the UNSHARDED input (each rank regenerates it from the same
seed) and compares this rank's device results with it; results that must be identical on all ranks are all-gathered and
compared bit for bit.  run_checks() returns {name: "ok" | "FAILED ..."}; it never raises for a mismatch, so that all
ranks keep walking through the same collectives."""
from __future__ import annotations

import numpy as np

SUM_RTOL = 1e-12      # order-dependent fp64 sums: |err| <= 1e-12 * sum|x| (north_star tolerance)


def _same_on_all_ranks(values, what):
    """bit-identical on every rank: all-gather the raw 8-byte patterns.  Collective; returns a list of error strings (empty
    when all ranks agree) instead of raising, so that callers can finish their own collectives first"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.float64).view(np.int64).copy()).cuda()
    got = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(got, t)
    return [f"{what}: rank {r} holds different bits than rank 0" for r, g in enumerate(got) if not torch.equal(g, got[0])]


class _Errors(list):
    """assertions that must not unbalance the collectives of the other ranks are collected and raised at the end of a check"""

    def check(self, cond, msg):
        if not cond:
            self.append(msg)

    def done(self):
        assert not self, "; ".join(self[:4])
        return "ok"


def check_fused_reduce(rank, world, x, orc, n=4_000_003):
    """onk_reduce_xchg_f64: MIN / MAX bit-exact against the oracle on the whole input; SUM against the rank-order fold of the
    per-rank partials (oracle order inside a rank) and a long-double reference, <= 1e-12; same bits on all ranks; repeated
    calls exercise the packet epochs"""
    import torch
    from onika_b200 import distributed as D, _capi
    from tests._inputs import lcg_uniform
    d = lcg_uniform(n, 7, -1.0, 1.0)
    b, e = D.shard_range(n, rank, world)
    shard = torch.from_numpy(d[b:e]).cuda()
    mag = float(np.abs(d).sum())
    fold = 0.0
    for r in range(world):
        rb, re = D.shard_range(n, r, world)
        fold = fold + orc.reduce_sum(d[rb:re])                     # rank-order fold of oracle-order partials
    want = {"min": orc.reduce_min(d), "max": orc.reduce_max(d)}
    got_all = []
    err = _Errors()
    for rep in range(3):
        got = {}
        for name, op in (("min", _capi.OP_MIN), ("max", _capi.OP_MAX), ("sum", _capi.OP_SUM)):
            out = x.reduce(op, shard, lane=0)
            torch.cuda.synchronize()
            got[name] = float(out.item())
        err.check(got["min"] == want["min"] and got["max"] == want["max"], f"fused min/max differ from the oracle (rep {rep})")
        err.check(abs(got["sum"] - fold) <= SUM_RTOL * mag, f"fused sum vs rank-order fold of oracle partials: {got['sum']!r} vs {fold!r}")
        err.check(abs(got["sum"] - orc.reduce_sum_ld(d)) <= SUM_RTOL * mag, "fused sum vs long-double reference")
        got_all += [got["min"], got["max"], got["sum"]]
    err.check(got_all[0:3] == got_all[3:6] == got_all[6:9], "fused reduction is not reproducible run to run")
    err += _same_on_all_ranks(got_all, "fused reduction")
    return err.done()


def check_allreduce_multi(rank, world, x):
    """onk_xchg_allreduce_f64 (the mpi::all_reduce_multi counterpart): 8 values, 3 operators, rank-order fold from the identity"""
    import torch
    from onika_b200 import distributed as D, _capi
    base = np.array([0.1, -2.5, 3.0, 1e-3, 7.75, -1e300, 2.0 ** -1040, 123456.789])
    per_rank = [base * (r + 1) + r * 0.37 for r in range(world)]
    got_all = []
    err = _Errors()
    for name, op in (("sum", _capi.OP_SUM), ("min", _capi.OP_MIN), ("max", _capi.OP_MAX)):
        for count in (8, 5, 1):
            v = torch.from_numpy(per_rank[rank][:count].copy()).cuda()
            x.allreduce(op, v, lane=0)
            torch.cuda.synchronize()
            got = v.cpu().numpy()
            want = np.array([D.fold_partials(op, [float(per_rank[r][k]) for r in range(world)]) for k in range(count)])
            err.check(np.array_equal(got, want), f"allreduce {name} x{count}: {got} vs {want}")
            got_all += got.tolist()
    err += _same_on_all_ranks(got_all, "allreduce_multi")
    return err.done()


def check_c4_zslabs(rank, world, x, orc, side=40):
    """config C4 sharded by z-slabs: per-cell sums bit-exact in both layouts, one all-reduce of the totals"""
    import torch
    from onika_b200 import distributed as D, cellgrid, _capi
    from tests._inputs import lcg_uniform, poisson_counts
    counts = poisson_counts(side ** 3)
    e = lcg_uniform(int(counts.sum()), 40, -1.0, 1.0)
    z0, z1 = D.shard_cells_z(side, rank, world)
    c0, c1 = z0 * side * side, z1 * side * side
    first = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    mine_counts, mine_e = counts[c0:c1], e[first[c0]:first[c1]]
    begin = first.astype(np.uint64)
    want, total = orc.cell_sum_packed(e, begin)
    err = _Errors()
    if c1 > c0:
        cg = cellgrid.CellGrid(mine_counts, 64, 16, (8, 8, 8, 8), 64)
        cg.upload(cg.host_arena_with_field(3, mine_e))
        s_arena, t_arena = cg.cell_sum(3)
        fm = cellgrid.FieldMajorCellGrid(mine_counts, nfields=1)
        fm.set_field(0, mine_e)
        s_fm, t_fm = fm.cell_sum(0)
        torch.cuda.synchronize()
        err.check(np.array_equal(s_arena.cpu().numpy(), want[c0:c1]), "per-cell sums (SOATL blocks) differ from the oracle")
        err.check(np.array_equal(s_fm.cpu().numpy(), want[c0:c1]), "per-cell sums (field-major) differ from the oracle")
        totals = torch.cat([t_arena.reshape(1), t_fm.reshape(1)])
    else:
        totals = torch.zeros(2, dtype=torch.float64, device="cuda")
    x.allreduce(_capi.OP_SUM, totals, lane=0)
    torch.cuda.synchronize()
    t = totals.cpu().numpy()
    mag = float(np.abs(e).sum())
    err.check(abs(t[0] - total) <= SUM_RTOL * mag and abs(t[1] - total) <= SUM_RTOL * mag, f"global total {t} vs {total}")
    err += _same_on_all_ranks(t, "C4 totals")
    return err.done()


def check_c5_sharded(rank, world, x1, x2, orc, rows=1025, cols=513):
    """config C5 scaled down: kernel1 -> kernel2 per array on two lanes, rows sharded, min and sum of each array combined
    across the GPUs inside the reduction kernels; arrays bit-exact against the oracle's lane sequence"""
    import torch
    import onika_b200 as ob
    from onika_b200 import distributed as D, parallel as P, _capi
    from tests._inputs import lcg_uniform
    a1 = lcg_uniform(rows * cols, 8, -1, 1).reshape(rows, cols)
    a2 = lcg_uniform(rows * cols, 9, -1, 1).reshape(rows, cols)
    b, e = D.shard_range(rows, rank, world)
    t1, t2 = torch.from_numpy(a1[b:e].copy()).cuda(), torch.from_numpy(a2[b:e].copy()).cuda()
    torch.cuda.synchronize()
    q_ = P.ParallelExecutionQueue()
    q_ << P.set_lane(1) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t1, 1.1)) \
       << P.set_lane(2) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t2, 1.3)) \
       << P.set_lane(1) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t1, 1.2)) \
       << P.set_lane(2) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t2, 1.4)) << P.flush
    r = [x1.reduce(_capi.OP_MIN, t1, lane=1), x1.reduce(_capi.OP_SUM, t1, lane=1),
         x2.reduce(_capi.OP_MIN, t2, lane=2), x2.reduce(_capi.OP_SUM, t2, lane=2)]
    q_ << P.synchronize
    ob.lane_sync()
    orc.lane_sequence(a1, a2, 1.1, 1.2, 1.3, 1.4)
    err = _Errors()
    err.check(np.array_equal(t1.cpu().numpy(), a1[b:e]) and np.array_equal(t2.cpu().numpy(), a2[b:e]), "C5 arrays differ from the oracle")
    m1, s1, m2, s2 = (float(v.item()) for v in r)
    err.check(m1 == a1.min() and m2 == a2.min(), "C5 minima differ")
    err.check(abs(s1 - orc.reduce_sum_ld(a1.ravel())) <= SUM_RTOL * np.abs(a1).sum(), "C5 sum of array 1")
    err.check(abs(s2 - orc.reduce_sum_ld(a2.ravel())) <= SUM_RTOL * np.abs(a2).sum(), "C5 sum of array 2")
    err += _same_on_all_ranks([m1, s1, m2, s2], "C5 reductions")
    return err.done()


def check_xs_move(rank, world, x, orc, n=1 << 20):
    """id-based redistribution of 2^20 elements between the GPUs (onk_xs_plan_build + onk_xs_move): byte-identical to the
    oracle's communication scheme + data_move"""
    import torch
    from onika_b200 import distributed as D
    from tests._inputs import xs_random_case
    id_min, id_max, before, after = xs_random_case(world, 7, n, 1 << 14)
    nb_all = sum(b.size for b in before)
    b0 = sum(b.size for b in before[:rank])
    rs = np.random.RandomState(3)
    all8 = rs.rand(nb_all)
    all32 = rs.randint(0, 256, (nb_all, 32)).astype(np.uint8)
    vals8, vals32 = all8[b0:b0 + before[rank].size], all32[b0:b0 + before[rank].size]
    mv = D.IdRedistribution(x, id_min, id_max)
    ib = torch.from_numpy(before[rank].astype(np.int64)).cuda()
    ia = torch.from_numpy(after[rank].astype(np.int64)).cuda()
    mv.plan(ib, ia, lane=0)
    na = after[rank].size
    w8, w32 = D.PeerWindow(max(na, 1) * 8, x), D.PeerWindow(max(na, 1) * 32, x)
    mv.move(torch.from_numpy(vals8.copy()).cuda(), w8, lane=0)
    mv.move(torch.from_numpy(vals32.copy()).cuda(), w32, lane=0)
    torch.cuda.synchronize()
    out8 = w8.tensor(torch.float64).cpu().numpy()[:na].copy()
    out32 = w32.tensor(torch.uint8).cpu().numpy()[:na * 32].reshape(na, 32).copy()
    sch = orc.xs_comm_scheme(id_min, id_max, before, after)
    want8, want32 = orc.xs_data_move(sch, all8), orc.xs_data_move(sch, all32)
    a0, a1 = int(sch.after_off[rank]), int(sch.after_off[rank + 1])
    err = _Errors()
    err.check(np.array_equal(out8, want8[a0:a1]) and np.array_equal(out32, want32[a0:a1]), "moved arrays differ from the oracle")
    err.check(mv.send_counts() == sch.send_count[rank].tolist(), "send counts differ from the oracle's scheme")
    import torch.distributed as dist
    dist.barrier()
    mv.close(); w8.close(); w32.close()
    return err.done()


def run_checks(rank: int, world: int, which=None, transport=None) -> dict:
    """all checks (or those named in `which`) on the calling process group; lane 0 must be usable.  Collective.
    transport: exchange buffers in "device" memory (CUDA IPC, default) or in "host" shared memory (no peer access needed)"""
    import os
    import torch
    import oracle
    from onika_b200 import distributed as D
    orc = oracle.orc
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    orc.set_num_threads(max(1, ncpu // max(1, world)))
    x, x2 = D.PeerExchange(transport=transport), D.PeerExchange(transport=transport)
    out = {}
    table = {
        "fused_reduce_min_max_sum": lambda: check_fused_reduce(rank, world, x, orc),
        "allreduce_multi_8_values_3_ops": lambda: check_allreduce_multi(rank, world, x),
        "c4_cell_sums_z_slabs": lambda: check_c4_zslabs(rank, world, x, orc),
        "c5_lane_sequence_sharded_rows": lambda: check_c5_sharded(rank, world, x, x2, orc),
        "xs_move_2^20_ids": lambda: check_xs_move(rank, world, x, orc),
    }
    import torch.distributed as dist
    for name, fn in table.items():
        if which is None or name in which:
            try:
                out[name] = fn()
            except AssertionError as ex:          # raised at the end of a check, after its collectives: the ranks stay in step
                out[name] = f"FAILED on rank {rank}: {ex}"[:400]
    torch.cuda.synchronize()
    dist.barrier()
    x.close(); x2.close()
    return out
