"""CPU, world_size 2 over gloo: the N>1 host logic of the path -- contiguous sharding of elements / rows / cells and the
rank-order combination of reduction partials -- against the oracle on the unsharded data.  (Kernels are not involved:
each rank's partial is produced by the oracle here; on GPUs it comes from onk_reduce_f64.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._inputs import lcg_uniform


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from onika_b200 import distributed as D, _capi
    d = lcg_uniform(n, 7, -1.0, 1.0)
    b, e = D.shard_range(n, rank, world)
    shard = d[b:e]
    out = {}
    for name, op, fn in (("min", _capi.OP_MIN, oracle.orc.reduce_min), ("max", _capi.OP_MAX, oracle.orc.reduce_max),
                         ("sum", _capi.OP_SUM, lambda a: oracle.orc.reduce_sum(a, 4))):
        part = torch.tensor([fn(shard)], dtype=torch.float64)
        out[name] = float(D.combine_across_ranks(op, part)[0])
    # rows of an Array2D sharded contiguously: value-add on the shard only
    rows, cols = 37, 11
    a = lcg_uniform(rows * cols, 3).reshape(rows, cols)
    rb, re_ = D.shard_range(rows, rank, world)
    mine = a[rb:re_].copy()
    oracle.orc.value_add(mine, 1.1)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rb, re_, mine))
    out["rows"] = gathered
    q.put((rank, b, e, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_000_003, 5])
def test_world2_sharding_and_partial_combination(n, orc):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = lcg_uniform(n, 7, -1.0, 1.0)
    # shards tile the range exactly
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n
    for rank, b, e, out in res:
        assert out["min"] == orc.reduce_min(d) and out["max"] == orc.reduce_max(d)       # bit-exact
        assert abs(out["sum"] - orc.reduce_sum_ld(d)) <= 1e-12 * np.abs(d).sum()
    assert res[0][3]["sum"] == res[1][3]["sum"]                                          # same bits on every rank
    rows, cols = 37, 11
    a = lcg_uniform(rows * cols, 3).reshape(rows, cols)
    orc.value_add(a, 1.1)
    rebuilt = np.concatenate([blk for _, _, blk in sorted(res[0][3]["rows"], key=lambda t: t[0])])
    assert np.array_equal(rebuilt, a)


def test_shard_range_properties():
    from onika_b200 import distributed as D
    for n in (0, 1, 7, 128, 2 ** 28, 10_000_000):
        for world in (1, 2, 4, 8):
            edges = [D.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    assert D.fold_partials(0, [3.0, float("nan"), 2.0]) == 2.0
    assert D.fold_partials(2, [0.1, 0.2, 0.3]) == (0.1 + 0.2) + 0.3


# ---- handle exchange of peer-mapped buffers (exchange buffers, id directories, output windows): host plumbing only ----
def _worker_handles(rank, world, port, fail_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from onika_b200 import distributed as D
    from onika_b200._capi import ONK_IPC_HANDLE_BYTES
    seen = {}

    def export(buf):                      # stands in for onk_window_export / onk_xs_plan_export
        if rank == fail_rank:
            return -2
        buf.raw = bytes([rank + 1]) * ONK_IPC_HANDLE_BYTES
        return 0

    def connect(all_handles):             # stands in for onk_window_connect: receives world handles in rank order
        seen["handles"] = all_handles
        return 0

    try:
        D._share_handles(export, connect, None, "a test buffer")
        q.put((rank, "ok", seen.get("handles")))
    except RuntimeError as e:
        q.put((rank, "raised", str(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_world2_peer_handle_exchange_is_collective_safe(fail_rank):
    from onika_b200._capi import ONK_IPC_HANDLE_BYTES
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_handles, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if fail_rank < 0:
        want = bytes([1]) * ONK_IPC_HANDLE_BYTES + bytes([2]) * ONK_IPC_HANDLE_BYTES
        assert [r[1] for r in res] == ["ok", "ok"] and all(r[2] == want for r in res)
    else:                                 # one rank could not export: every rank raises, nobody hangs in a collective
        assert [r[1] for r in res] == ["raised", "raised"]
