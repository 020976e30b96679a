"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the fused reduction + cross-GPU combine over NVLink peer memory
(onk_reduce_xchg_f64) against the oracle on the unsharded data, one process per GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onika_b200 as ob
    from onika_b200 import distributed as D, _capi
    from tests._inputs import lcg_uniform
    ob.init(rank)
    ob.bind_lane_to_torch_stream(0)
    d = lcg_uniform(n, 7, -1.0, 1.0)
    b, e = D.shard_range(n, rank, world)
    shard = torch.from_numpy(d[b:e]).cuda()
    x = D.PeerExchange()
    res = {}
    for rep in range(3):                      # repeated calls exercise the epoch / parity protocol
        for name, op in (("min", _capi.OP_MIN), ("max", _capi.OP_MAX), ("sum", _capi.OP_SUM)):
            out = x.reduce(op, shard)
            torch.cuda.synchronize()
            res[(name, rep)] = float(out.item())
    # same answer as the NCCL path
    part = ob.parallel.reduce_min(shard)
    nccl = float(D.combine_across_ranks(_capi.OP_MIN, part).item())
    # all_reduce_multi counterpart: 5 scalars, one operator, in place
    base = np.array([0.1, -2.5, 3.0, 1e-3, 7.75])
    for name, op in (("sum", _capi.OP_SUM), ("min", _capi.OP_MIN), ("max", _capi.OP_MAX)):
        v = torch.from_numpy(base * (rank + 1) + rank).cuda()
        x.allreduce(op, v)
        torch.cuda.synchronize()
        res[("multi", name)] = v.cpu().numpy().tolist()
    q.put((rank, res, nccl))
    dist.barrier()
    x.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4_000_003, 17])
def test_fused_reduce_exchange_two_gpus(n, orc):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from tests._inputs import lcg_uniform
    d = lcg_uniform(n, 7, -1.0, 1.0)
    for rank, res, nccl in got:
        for rep in range(3):
            assert res[("min", rep)] == orc.reduce_min(d) and res[("max", rep)] == orc.reduce_max(d)
            assert abs(res[("sum", rep)] - orc.reduce_sum_ld(d)) <= 1e-12 * np.abs(d).sum()
        assert nccl == orc.reduce_min(d)
        base = np.array([0.1, -2.5, 3.0, 1e-3, 7.75])
        per_rank = [base * (r + 1) + r for r in range(world)]
        assert res[("multi", "sum")] == ((0.0 + per_rank[0]) + per_rank[1]).tolist()      # rank-order fold from the identity
        assert res[("multi", "min")] == np.minimum(per_rank[0], per_rank[1]).tolist()
        assert res[("multi", "max")] == np.maximum(per_rank[0], per_rank[1]).tolist()
    assert got[0][1] == got[1][1]             # identical bits on both ranks, every repetition


def _worker_c5(rank, world, port, rows, cols, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onika_b200 as ob
    from onika_b200 import distributed as D, parallel as P, _capi
    from tests._inputs import lcg_uniform
    ob.init(rank)
    a1 = lcg_uniform(rows * cols, 8, -1, 1).reshape(rows, cols)
    a2 = lcg_uniform(rows * cols, 9, -1, 1).reshape(rows, cols)
    b, e = D.shard_range(rows, rank, world)                  # contiguous row shard of both arrays
    t1, t2 = torch.from_numpy(a1[b:e].copy()).cuda(), torch.from_numpy(a2[b:e].copy()).cuda()
    torch.cuda.synchronize()
    x1, x2 = D.PeerExchange(), D.PeerExchange()              # one exchange per lane: calls on one exchange are ordered by its lane
    q_ = P.ParallelExecutionQueue()
    q_ << P.set_lane(1) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t1, 1.1)) \
       << P.set_lane(2) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t2, 1.3)) \
       << P.set_lane(1) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t1, 1.2)) \
       << P.set_lane(2) << P.block_parallel_for(e - b, P.BlockParallelValueAddFunctor(t2, 1.4)) << P.flush
    r = [x1.reduce(_capi.OP_MIN, t1, lane=1), x1.reduce(_capi.OP_SUM, t1, lane=1),
         x2.reduce(_capi.OP_MIN, t2, lane=2), x2.reduce(_capi.OP_SUM, t2, lane=2)]
    q_ << P.synchronize
    ob.lane_sync()
    q.put((rank, b, e, t1.cpu().numpy(), t2.cpu().numpy(), [float(v.item()) for v in r]))
    dist.barrier()
    x1.close(); x2.close()
    dist.destroy_process_group()


def test_config_c5_lane_sequence_sharded_over_two_gpus(orc):
    """BASELINE config C5 (scaled down): kernel1 -> kernel2 per array on lanes 1/2, rows sharded over 2 GPUs, min and sum
    of each array combined across the GPUs inside the reduction kernels."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests._inputs import lcg_uniform
    world, rows, cols = 2, 1025, 513
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_c5, args=(r, world, port, rows, cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a1 = lcg_uniform(rows * cols, 8, -1, 1).reshape(rows, cols)
    a2 = lcg_uniform(rows * cols, 9, -1, 1).reshape(rows, cols)
    orc.lane_sequence(a1, a2, 1.1, 1.2, 1.3, 1.4)
    assert np.array_equal(np.concatenate([g[3] for g in got]), a1) and np.array_equal(np.concatenate([g[4] for g in got]), a2)
    for g in got:
        m1, s1, m2, s2 = g[5]
        assert m1 == a1.min() and m2 == a2.min()
        assert abs(s1 - orc.reduce_sum_ld(a1.ravel())) <= 1e-12 * np.abs(a1).sum()
        assert abs(s2 - orc.reduce_sum_ld(a2.ravel())) <= 1e-12 * np.abs(a2).sum()
    assert got[0][5] == got[1][5]


# ---------------------------------------------------------------- id-based redistribution over peer memory
def _worker_xs(rank, world, port, cases, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onika_b200 as ob
    from onika_b200 import distributed as D
    from tests._inputs import xs_random_case, xs_simple_case
    ob.init(rank)
    ob.bind_lane_to_torch_stream(0)
    x = D.PeerExchange()
    res = {}
    for ci, (kind, args) in enumerate(cases):
        id_min, id_max, before, after = (xs_simple_case if kind == "simple" else xs_random_case)(*args)
        nb_all = sum(b.size for b in before)
        b0 = sum(b.size for b in before[:rank])
        rs = np.random.RandomState(3)
        vals8 = rs.rand(nb_all)[b0:b0 + before[rank].size]
        vals32 = rs.randint(0, 256, (nb_all, 32)).astype(np.uint8)[b0:b0 + before[rank].size]
        mv = D.IdRedistribution(x, id_min, id_max)
        ib = torch.from_numpy(before[rank].astype(np.int64)).cuda()
        ia = torch.from_numpy(after[rank].astype(np.int64)).cuda()
        mv.plan(ib, ia, lane=0)
        na = after[rank].size
        w8, w32 = D.PeerWindow(max(na, 1) * 8, x), D.PeerWindow(max(na, 1) * 32, x)
        for rep in range(2):                                   # windows are re-used: leading barrier of move()
            mv.move(torch.from_numpy(vals8).cuda(), w8, lane=0)
            mv.move(torch.from_numpy(vals32).cuda(), w32, lane=0)
        torch.cuda.synchronize()
        res[ci] = (w8.tensor(torch.float64).cpu().numpy()[:na].copy(), w32.tensor(torch.uint8).cpu().numpy()[:na * 32].reshape(na, 32).copy(),
                   mv.send_counts())
        # a rejected plan is rejected on every rank (rank 1 drops one id after)
        if ci == 0 and world > 1:
            bad = ia[:-1].contiguous() if rank == 1 else ia
            try:
                mv.plan(ib, bad, lane=0)
                res["rejected"] = False
            except ValueError:
                res["rejected"] = True
        dist.barrier()
        mv.close(); w8.close(); w32.close()
    q.put((rank, res))
    dist.barrier()
    x.close()
    dist.destroy_process_group()


def test_id_redistribution_two_gpus(orc):
    import torch
    import torch.multiprocessing as mp
    from tests._inputs import xs_random_case, xs_simple_case
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cases = [("simple", (2, 0, 107, 3)), ("simple", (2, 1023, 1023 + 100000, 77)), ("random", (2, 1976, 110, 1000)),
             ("random", (2, 0, 100000, 1000)), ("random", (2, 7, 1 << 20, 1 << 18))]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_xs, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for ci, (kind, args) in enumerate(cases):
        id_min, id_max, before, after = (xs_simple_case if kind == "simple" else xs_random_case)(*args)
        sch = orc.xs_comm_scheme(id_min, id_max, before, after)
        nb_all = sum(b.size for b in before)
        rs = np.random.RandomState(3)
        want8 = orc.xs_data_move(sch, rs.rand(nb_all))
        want32 = orc.xs_data_move(sch, rs.randint(0, 256, (nb_all, 32)).astype(np.uint8))
        for rank in range(world):
            a0, a1 = int(sch.after_off[rank]), int(sch.after_off[rank + 1])
            out8, out32, counts = got[rank][ci]
            assert np.array_equal(out8, want8[a0:a1]) and np.array_equal(out32, want32[a0:a1]), (ci, rank)
            assert counts == sch.send_count[rank].tolist()
    assert got[0]["rejected"] and got[1]["rejected"]


# ---------------------------------------------------------------- config C4 sharded by z-slabs: no halo, one total
def _worker_c4(rank, world, port, side, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onika_b200 as ob
    from onika_b200 import distributed as D, cellgrid, _capi
    from tests._inputs import lcg_uniform, poisson_counts
    ob.init(rank)
    ob.bind_lane_to_torch_stream(0)
    counts = poisson_counts(side ** 3)
    e = lcg_uniform(int(counts.sum()), 40, -1.0, 1.0)
    z0, z1 = D.shard_cells_z(side, rank, world)                    # cells are ordered z-major: a slab is a contiguous cell range
    c0, c1 = z0 * side * side, z1 * side * side
    first = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
    mine_counts, mine_e = counts[c0:c1], e[first[c0]:first[c1]]
    x = D.PeerExchange()
    out = {}
    # reference layout (per-cell FieldArrays blocks) and field-major layout: same per-cell bits, one fused all-reduce of both totals
    cg = cellgrid.CellGrid(mine_counts, 64, 16, (8, 8, 8, 8), 64)
    cg.upload(cg.host_arena_with_field(3, mine_e))
    s_arena, t_arena = cg.cell_sum(3)
    fm = cellgrid.FieldMajorCellGrid(mine_counts, nfields=1)
    fm.set_field(0, mine_e)
    s_fm, t_fm = fm.cell_sum(0)
    totals = torch.cat([t_arena, t_fm])
    x.allreduce(_capi.OP_SUM, totals)
    torch.cuda.synchronize()
    out = (c0, c1, s_arena.cpu().numpy(), s_fm.cpu().numpy(), totals.cpu().numpy())
    q.put((rank, out))
    dist.barrier()
    x.close()
    dist.destroy_process_group()


def test_cell_grid_z_slabs_two_gpus(orc):
    import torch
    import torch.multiprocessing as mp
    from tests._inputs import lcg_uniform, poisson_counts
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, side = 2, 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_c4, args=(r, world, port, side, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    counts = poisson_counts(side ** 3)
    e = lcg_uniform(int(counts.sum()), 40, -1.0, 1.0)
    begin = np.zeros(counts.size + 1, np.uint64)
    begin[1:] = np.cumsum(counts, dtype=np.uint64)
    want, total = orc.cell_sum_packed(e, begin)
    for rank in range(world):
        c0, c1, s_arena, s_fm, totals = got[rank]
        assert np.array_equal(s_arena, want[c0:c1]) and np.array_equal(s_fm, want[c0:c1])      # per-cell sums: bit-exact, both layouts
        assert abs(totals[0] - total) <= 1e-12 * np.abs(e).sum() and abs(totals[1] - total) <= 1e-12 * np.abs(e).sum()
    assert np.array_equal(got[0][4], got[1][4])                                                # same bits on both ranks
    assert got[0][1] == got[1][0] and got[0][0] == 0 and got[1][1] == side ** 3


# ---------------------------------------------------------------- the checks bench.py reports under `checks` at N > 1
def _worker_checks(rank, world, port, q, transport=None, which=None):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import onika_b200 as ob
    from tests import _multigpu_checks
    ob.init(rank)
    ob.bind_lane_to_torch_stream(0)
    q.put((rank, _multigpu_checks.run_checks(rank, world, which=which, transport=transport)))
    dist.barrier()
    dist.destroy_process_group()


def test_bench_multi_gpu_checks_two_gpus():
    """tests/_multigpu_checks.py (what `bench.py --gpus N` runs outside its timed regions): all five checks on 2 GPUs"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_checks, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        assert len(got[rank]) == 5 and all(v == "ok" for v in got[rank].values()), got[rank]


def test_exchange_through_host_shared_memory_two_gpus():
    """ONK_XCHG_TRANSPORT_HOST: the exchange buffers live in pinned host memory shared through POSIX shared memory (the fallback
    where CUDA IPC / peer access is refused): fused reductions, the k-value all-reduce, the C4 totals and the C5 sequence give the
    same bits as over NVLink"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    which = ("fused_reduce_min_max_sum", "allreduce_multi_8_values_3_ops", "c4_cell_sums_z_slabs", "c5_lane_sequence_sharded_rows")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_checks, args=(r, world, port, q, "host", which)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        assert len(got[rank]) == 4 and all(v == "ok" for v in got[rank].values()), got[rank]


def test_cxx_multi_gpu_helper_two_processes(tmp_path, orc):
    """the C++ front end's multi-GPU helper (include/onika/parallel/multi_gpu.h) without MPI or Python in the data path: two
    processes of tests/cxx/multi_gpu_test, handles exchanged through the file rendezvous"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.test_gpu_cxx_frontend import run_multi_gpu_program, check_multi_gpu_outputs
    for n in (4_000_003, 5):
        check_multi_gpu_outputs(run_multi_gpu_program(str(tmp_path), 2, n), orc, n, 2)
    check_multi_gpu_outputs(run_multi_gpu_program(str(tmp_path), 2, 100_001, transport=1), orc, 100_001, 2)      # host shared memory
