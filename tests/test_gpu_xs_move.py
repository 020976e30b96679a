"""GPU (-m gpu): id-based redistribution (onk_xs_*: directory over peer memory + direct scatter into peer windows) against
the oracle's restatement of the reference's communication_scheme_from_ids + data_move, through the C ABI.
One-GPU cases here (world = 1: the directory, routes, verification and the scatter kernels run, peers are absent);
tests/test_gpu_multi.py holds the 2-GPU cases."""
import numpy as np
import pytest

from tests._inputs import xs_random_case, xs_simple_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xch():
    import torch
    from onika_b200 import runtime
    from onika_b200.distributed import PeerExchange
    runtime.init(0)
    runtime.bind_lane_to_torch_stream(0)
    x = PeerExchange()
    yield x
    torch.cuda.synchronize()
    x.close()


def _move(xch, id_min, id_max, before, after, values, dtype):
    import torch
    from onika_b200.distributed import IdRedistribution, PeerWindow
    mv = IdRedistribution(xch, id_min, id_max)
    ib = torch.from_numpy(before.astype(np.int64)).cuda()
    ia = torch.from_numpy(after.astype(np.int64)).cuda()
    mv.plan(ib, ia, lane=0)
    v = torch.from_numpy(values).cuda()
    elem = values.dtype.itemsize * int(np.prod(values.shape[1:], dtype=np.int64))
    guard = 512                                              # windows are zero-filled: bytes past the payload must stay zero
    out = PeerWindow(after.size * elem + guard, xch)
    mv.move(v, out, lane=0)
    torch.cuda.synchronize()
    raw = out.tensor(torch.uint8).cpu().numpy()
    assert not raw[after.size * elem:].any(), "scatter wrote past the last element of the window"
    got = raw[:after.size * elem].view(values.dtype).reshape((after.size,) + values.shape[1:])
    counts = mv.send_counts()
    mv.close()
    out.close()
    return got, counts


@pytest.mark.parametrize("case", [("simple", (1, 0, 10, 0)), ("simple", (1, 1023, 1023 + 100000, 77)), ("random", (1, 0, 110, 1000)),
                                  ("random", (1, 1976, 100000, 1000)), ("random", (1, 5, 1 << 20, 1 << 16))])
def test_single_rank_move_equals_oracle(xch, orc, case):
    import torch
    kind, args = case
    id_min, id_max, before, after = (xs_simple_case if kind == "simple" else xs_random_case)(*args)
    sch = orc.xs_comm_scheme(id_min, id_max, before, after)
    n = before[0].size
    rs = np.random.RandomState(3)
    for values, dtype in ((rs.rand(n), torch.float64),                                   # 8-byte elements
                          (rs.randint(0, 256, (n, 32)).astype(np.uint8), torch.uint8),     # 32-byte records (4 doubles worth)
                          (rs.randint(0, 256, (n, 3)).astype(np.uint8), torch.uint8),      # odd size: byte pieces
                          (rs.rand(n, 3).astype(np.float32), torch.float32),               # 12 bytes: 4-byte pieces
                          (before[0].astype(np.int64), torch.int64)):                      # the reference's own check: ids onto ids
        got, counts = _move(xch, id_min, id_max, before[0], after[0], values, dtype)
        want = orc.xs_data_move(sch, values)
        assert np.array_equal(got, want)
        assert counts == sch.send_count[0].tolist()
    assert np.array_equal(_move(xch, id_min, id_max, before[0], after[0], before[0].astype(np.int64), torch.int64)[0], after[0].astype(np.int64))


def test_ids_that_do_not_match_up_are_rejected(xch):
    import torch
    from onika_b200.distributed import IdRedistribution
    mv = IdRedistribution(xch, 10, 18)
    ids = torch.arange(10, 18, dtype=torch.int64, device="cuda")
    cases = {5: (ids, torch.cat([ids[:7], ids.new_tensor([99])])),                 # out of range (and 17 without destination: 5 wins)
             4: (ids, torch.cat([ids, ids[:1]])),                                  # duplicate after
             3: (torch.cat([ids, ids[:1]]), ids),                                  # duplicate before
             2: (ids, ids[:7]),                                                    # id without destination
             1: (ids[:7], ids)}                                                    # id without source
    for code, (b, a) in cases.items():
        with pytest.raises(ValueError, match=IdRedistribution.ERRORS[code]):
            mv.plan(b.contiguous(), a.contiguous(), lane=0)
    mv.plan(ids, ids.flip(0).contiguous(), lane=0)      # the plan object stays usable after a rejected build
    assert mv.send_counts() == [8]
    mv.close()


def test_single_rank_allreduce_and_barrier_are_identities(xch):
    import torch
    from onika_b200 import _capi
    v = torch.tensor([1.5, -2.0, 3.25], dtype=torch.float64, device="cuda")
    for op in (_capi.OP_SUM, _capi.OP_MIN, _capi.OP_MAX):
        w = v.clone()
        xch.allreduce(op, w, lane=0)
        xch.barrier(lane=0)
        torch.cuda.synchronize()
        assert torch.equal(w, v)
    with pytest.raises(Exception):
        xch.allreduce(_capi.OP_SUM, torch.zeros(9, dtype=torch.float64, device="cuda"), lane=0)      # more than 8 values
    import ctypes as C
    L = _capi.lib()
    rank, world = C.c_int(-1), C.c_int(-1)
    _capi.check(L.onk_xchg_info(xch.handle, C.byref(rank), C.byref(world)))
    assert (rank.value, world.value) == (0, 1)
    _capi.check(L.onk_device_sync())                      # everything queued on every lane has completed
    idle = C.c_int(0)
    _capi.check(L.onk_lane_query(_capi.LANE_ALL, C.byref(idle)))
    assert idle.value == 1


def test_several_fields_move_with_one_pair_of_barriers(xch, orc):
    """onk_xs_move_fields: the fields of a particle (positions as 24-byte records, an 8-byte id, a 1-byte type) travel with the
    same plan in one call; each must equal the oracle's data_move of that array"""
    import torch
    from onika_b200.distributed import IdRedistribution, PeerWindow
    id_min, id_max, before, after = xs_random_case(1, 9, 50000, 20000)
    sch = orc.xs_comm_scheme(id_min, id_max, before, after)
    n = before[0].size
    rs = np.random.RandomState(5)
    fields = [rs.rand(n, 3), before[0].astype(np.int64), rs.randint(0, 256, n).astype(np.uint8)]
    mv = IdRedistribution(xch, id_min, id_max)
    mv.plan(torch.from_numpy(before[0].astype(np.int64)).cuda(), torch.from_numpy(after[0].astype(np.int64)).cuda(), lane=0)
    wins = [PeerWindow(after[0].size * f[0:1].nbytes, xch) for f in fields]
    mv.move_fields([torch.from_numpy(f).cuda() for f in fields], wins, lane=0)
    torch.cuda.synchronize()
    for f, w in zip(fields, wins):
        got = w.tensor(torch.uint8).cpu().numpy()[:after[0].size * f[0:1].nbytes].view(f.dtype).reshape((after[0].size,) + f.shape[1:])
        assert np.array_equal(got, orc.xs_data_move(sch, f))
    assert np.array_equal(wins[1].tensor(torch.int64).cpu().numpy()[:after[0].size], after[0].astype(np.int64))   # ids onto ids
    mv.close()
    for w in wins:
        w.close()
