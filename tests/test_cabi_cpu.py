"""CPU: the C-ABI library loads, exports every symbol the header declares, its host arithmetic agrees with the oracle
and the golden vectors, and the compute path fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import onika_b200 as ob
from onika_b200 import _capi, parallel, soatl, cellgrid

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")


def _no_gpu():
    n = C.c_int(0)
    rc = _capi.lib().onk_device_count(C.byref(n))
    return rc != 0 or n.value == 0


def test_library_exports_every_declared_symbol():
    hs = _capi.header_symbols()
    assert len(hs) >= 40
    assert set(hs) == set(_capi.PROTOTYPES), set(hs) ^ set(_capi.PROTOTYPES)
    L = _capi.lib()
    for name in hs:
        assert hasattr(L, name), name
    assert L.onk_abi_version() == 2


def test_only_cuda_runtime_is_linked_no_torch():
    import subprocess
    out = subprocess.run(["ldd", _capi._build.LIB], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out


def test_no_cpu_fallback_without_gpu():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(ob.OnikaError) as ei:
        ob.init(0)
    assert ei.value.code == _capi.ERR_NO_DEVICE
    d = np.ones(16)
    out = np.zeros(1)
    rc = _capi.lib().onk_reduce_f64(_capi.OP_MIN, d.ctypes.data, 16, out.ctypes.data, 0)
    assert rc == _capi.ERR_NO_DEVICE and out[0] == 0.0   # nothing was computed on the host
    with pytest.raises(ob.OnikaError):
        parallel.host_reduce(_capi.OP_MIN, d)


def test_argument_errors_are_reported_not_aborted():
    L = _capi.lib()
    assert L.onk_device_count(None) == _capi.ERR_ARG
    assert b"null" in L.onk_last_error()


def test_soatl_host_arithmetic_matches_golden_and_oracle(orc):
    with open(GOLD) as f:
        gold = json.load(f)
    for case in gold["soatl_layout"]:
        cap = 0
        for s, gcap, gbytes, goffs in zip(case["sizes"], case["capacity"], case["storage_bytes"], case["offsets"]):
            prev = cap
            cap = soatl.update_capacity(s, prev, case["chunk"])
            assert cap == gcap == orc.update_capacity(s, prev, case["chunk"])
            total, offs = soatl.layout(case["align"], case["fields"], cap)
            assert total == gbytes
            if cap:
                assert offs == goffs
    rng = np.random.default_rng(1)
    for _ in range(2000):
        s, c, ch = int(rng.integers(0, 100000)), int(rng.integers(0, 100000)), int(rng.choice([1, 4, 8, 16]))
        c = c // ch * ch
        assert soatl.update_capacity(s, c, ch) == orc.update_capacity(s, c, ch)


def test_cell_grid_layout_matches_oracle(orc):
    from tests._inputs import poisson_counts
    counts = poisson_counts(5000)
    for align, chunk, fields, cell_align in [(64, 16, [8, 8, 8, 8], 64), (256, 16, [8, 8, 8, 8], 256), (64, 8, [8, 1, 8, 4, 8, 8], 64)]:
        t, cap, off = cellgrid.cell_grid_layout(counts, align, chunk, fields, cell_align)
        to, capo, offo = orc.cell_grid_layout(counts, align, chunk, fields, cell_align)
        assert t == to and np.array_equal(cap, capo) and np.array_equal(off, offo)


class _Rec:
    log = []

    def __init__(self, name):
        self.name = name

    def launch(self, n, lane):
        _Rec.log.append((self.name, n, lane))


def test_queue_semantics_lanes_fifo_and_delay_queues(monkeypatch):
    """ordering rules of the reference queue (parallel_execution_queue.cpp:94-111,242-271; parallel_execution_operators.h:62-112),
    checked on the host logic alone (launches are recorded instead of executed)."""
    monkeypatch.setattr(parallel.ParallelExecutionQueue, "wait", lambda self, lane=-2: self.m_exec_list.clear())
    P = parallel
    _Rec.log = []
    q = P.ParallelExecutionQueue()
    ops = [P.block_parallel_for(10 + i, _Rec(f"k{i}")) for i in range(4)]
    q << P.set_lane(0) << ops[0] << P.set_lane(1) << ops[1] << P.set_lane(0) << ops[2] << P.set_lane(1) << ops[3]
    assert _Rec.log == []                       # nothing runs before flush
    q << P.flush
    assert _Rec.log == [("k0", 10, 0), ("k1", 11, 1), ("k2", 12, 0), ("k3", 13, 1)]
    assert q.m_lane == P.DEFAULT_EXECUTION_LANE  # flush resets the lane
    q << P.synchronize
    assert q.empty()
    # delay queues keep their lanes when spliced (tests/parallel_queue_lane.cu:87-96)
    _Rec.log = []
    qa, qb = P.ParallelExecutionQueueBase(), P.ParallelExecutionQueueBase()
    o = [P.block_parallel_for(1, _Rec(f"d{i}")) for i in range(4)]
    qa << P.set_lane(0) << o[0] << P.set_lane(1) << o[1]
    qb << P.set_lane(0) << o[2] << P.set_lane(1) << o[3]
    q << qa << qb << P.flush << P.synchronize
    assert _Rec.log == [("d0", 1, 0), ("d1", 1, 1), ("d2", 1, 0), ("d3", 1, 1)]
    # default lane -> 0 ; selective scheduling by lane
    _Rec.log = []
    a, b = P.block_parallel_for(1, _Rec("a")), P.block_parallel_for(1, _Rec("b"))
    q << a << P.set_lane(3) << b
    q.schedule_all(3)
    assert _Rec.log == [("b", 1, 3)] and len(q.m_queue_list) == 1
    q.schedule_all()
    assert _Rec.log[-1] == ("a", 1, 0)
    q.wait()
    # wrappers are move-only
    w = P.block_parallel_for(1, _Rec("w"))
    q << w
    with pytest.raises(RuntimeError):
        q << w
    q << P.flush << P.synchronize
    # invalid lane is an error, as in the reference (parallel_execution_queue.cpp:35-39)
    bad = P.block_parallel_for(1, _Rec("bad"))
    q << P.set_lane(300) << bad
    with pytest.raises(ValueError):
        q << P.flush
    q.m_queue_list.clear()


def test_ctypes_prototypes_have_the_arity_of_the_header_declarations():
    """a parameter added to a C declaration (e.g. the column length of onk_cell_sum_packed_f64) must show up in the ctypes
    table too: compare argument counts, and pointer-ness argument by argument"""
    import ctypes as C
    import re
    with open(_capi.HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    decls = dict(re.findall(r"\b(onk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text))
    assert set(decls) == set(_capi.PROTOTYPES)
    for name, params in decls.items():
        plist = [] if params.strip() in ("", "void") else [p.strip() for p in params.split(",")]
        res, args = _capi.PROTOTYPES[name]
        assert len(plist) == len(args), (name, plist, args)
        for p, a in zip(plist, args):
            is_ptr_c = ("*" in p) or ("[" in p)
            is_ptr_py = a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents") or (isinstance(a, type) and issubclass(a, C._Pointer))
            assert is_ptr_c == is_ptr_py, (name, p, a)
