"""CPU: the oracle's id-based redistribution (oracle.c:orc_xs_comm_scheme / orc_xs_data_move) against
 (1) the reference's own acceptance property (tests/xs_data_move_simple.cpp:98-117, xs_data_move_random.cpp:121-139: moving the
     `before` ids with the computed scheme must reproduce the `after` ids) on the parameter grid of CMakeLists.txt:681-711,
 (2) the UNMODIFIED reference run on in-process ranks (oracle/_ref, built where /root/reference exists): index tables and
     moved payloads bit for bit,
 (3) golden tables committed under tests/golden/ (generated from (2) by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from tests._inputs import xs_random_case, xs_simple_case

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")


def _check_property(orc, id_min, id_max, before, after):
    sch = orc.xs_comm_scheme(id_min, id_max, before, after)
    moved = orc.xs_data_move(sch, np.concatenate(before))
    assert np.array_equal(moved, np.concatenate(after))
    # tables are permutations of each rank's buffer slots, and counts agree pairwise
    for r in range(len(before)):
        s = sch.send_indices[int(sch.before_off[r]):int(sch.before_off[r + 1])]
        v = sch.recv_indices[int(sch.after_off[r]):int(sch.after_off[r + 1])]
        assert np.array_equal(np.sort(s), np.arange(s.size)) and np.array_equal(np.sort(v), np.arange(v.size))
    assert np.array_equal(sch.send_count, sch.recv_count.T)
    return sch


@pytest.mark.parametrize("world", [1, 7, 8])
@pytest.mark.parametrize("rotation", [0, 1, 77])
@pytest.mark.parametrize("start", [0, 1, 1023])
@pytest.mark.parametrize("count", [10, 110, 100000])
def test_simple_grid_moves_before_ids_onto_after_ids(orc, world, rotation, start, count):
    _check_property(orc, *xs_simple_case(world, start, start + count, rotation))


@pytest.mark.parametrize("args", [(3, 0, 20, 3), (3, 0, 107, 3)])
def test_simple_adhoc(orc, args):
    _check_property(orc, *xs_simple_case(*args))


@pytest.mark.parametrize("world", [1, 7, 8])
@pytest.mark.parametrize("seed", [0, 1976])
@pytest.mark.parametrize("count", [10, 110, 100000])
@pytest.mark.parametrize("perms", [10, 1000])
def test_random_grid_moves_before_ids_onto_after_ids(orc, world, seed, count, perms):
    _check_property(orc, *xs_random_case(world, seed, count, perms))


def test_range_utils(orc):
    """include/onika/mpi/xs_data_move_range_utils.h:96-108 (the reference's unit test of the id -> rendezvous rank map)"""
    lo, hi = 5, 1000
    for nprocs in range(1, 16):
        starts = [orc.xs_range_start(r, nprocs, lo, hi) for r in range(nprocs)] + [hi]
        for i in range(lo, hi):
            p = orc.xs_process_for_id(i, nprocs, lo, hi)
            assert starts[p] <= i < starts[p + 1]


def test_reference_own_unit_tests_pass(ref):
    run, failed = ref.run_unit_tests()
    assert run >= 2 and failed == 0


def test_missing_id_is_an_error(orc):
    with pytest.raises(ValueError):
        orc.xs_comm_scheme(0, 8, [np.arange(0, 4, dtype=np.uint64), np.arange(4, 8, dtype=np.uint64)],
                           [np.arange(0, 4, dtype=np.uint64), np.arange(4, 7, dtype=np.uint64)])


CASES = [("simple", (1, 0, 10, 0)), ("simple", (3, 0, 107, 3)), ("simple", (7, 1, 111, 77)), ("simple", (8, 1023, 1023 + 100000, 1)),
         ("random", (7, 1976, 110, 1000)), ("random", (8, 0, 100000, 1000)), ("random", (2, 3, 4097, 5000))]


@pytest.mark.parametrize("kind,args", CASES)
def test_oracle_equals_reference(orc, ref, kind, args):
    id_min, id_max, before, after = (xs_simple_case if kind == "simple" else xs_random_case)(*args)
    nb = sum(len(b) for b in before)
    payload8 = np.random.RandomState(1).rand(nb)
    payload32 = np.random.RandomState(2).randint(0, 256, (nb, 32)).astype(np.uint8)
    rsch, rout8, rout32 = ref.xs_data_move(id_min, id_max, before, after, payload8, payload32)
    osch = orc.xs_comm_scheme(id_min, id_max, before, after)
    for a, b, name in zip(osch.tables(), rsch.tables(), ("send_indices", "recv_indices", "send_count", "send_displ", "recv_count", "recv_displ")):
        assert np.array_equal(a, b), name
    assert np.array_equal(orc.xs_data_move(osch, payload8), rout8)
    assert np.array_equal(orc.xs_data_move(osch, payload32), rout32)


def test_oracle_equals_golden_tables(orc):
    with open(GOLD) as f:
        g = json.load(f)["xs_data_move"]
    for case in g:
        maker = xs_simple_case if case["kind"] == "simple" else xs_random_case
        id_min, id_max, before, after = maker(*case["args"])
        sch = orc.xs_comm_scheme(id_min, id_max, before, after)
        assert sch.send_indices.tolist() == case["send_indices"] and sch.recv_indices.tolist() == case["recv_indices"]
        assert sch.send_count.tolist() == case["send_count"] and sch.recv_displ.tolist() == case["recv_displ"]
