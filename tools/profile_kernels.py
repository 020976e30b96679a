#!/usr/bin/env python
"""Launch every typed kernel a few times at the BASELINE config sizes (for ncu captures and quick device timings).

  python tools/profile_kernels.py [--reps 5] [--only cells,axpy,...]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    import torch
    import onika_b200 as ob
    from onika_b200 import parallel as P, soatl, cellgrid
    from tests._inputs import poisson_counts
    ob.init(0)
    ob.bind_lane_to_torch_stream(0)
    # L2 flush by READING 1 GiB (clean lines): a write-flush would leave 126 MB of dirty lines whose eviction gets billed
    # to the next kernel
    flush = torch.ones(1 << 27, dtype=torch.float64, device="cuda")
    flush_out = torch.empty(1, dtype=torch.float64, device="cuda")

    def timed(name, fn, nbytes, flush_l2=False):
        if only and name not in only:
            return
        ts = []
        for _ in range(args.reps + 2):
            if flush_l2:
                P.ReduceFunctor(ob.OP_MIN, flush, flush_out).launch(flush.numel(), 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts = sorted(ts[2:])
        t = ts[len(ts) // 2]
        print(f"{name:28s} {t * 1e6:10.1f} us   {nbytes / t / 1e9:8.1f} GB/s  (min {nbytes / ts[-1] / 1e9:.0f} max {nbytes / ts[0] / 1e9:.0f})", flush=True)

    n = 10_000_000
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
    timed("axpy", lambda: P.AxpyFunctor(y, x, 0.5).launch(n, 0), 24 * n, True)
    n = 1 << 28
    big = torch.rand(n, dtype=torch.float64, device="cuda")
    o = torch.empty(1, dtype=torch.float64, device="cuda")
    timed("reduce_min", lambda: P.ReduceFunctor(ob.OP_MIN, big, o).launch(n, 0), 8 * n)
    timed("reduce_sum", lambda: P.ReduceFunctor(ob.OP_SUM, big, o).launch(n, 0), 8 * n)
    timed("axpy_big", lambda: P.AxpyFunctor(big[: n // 2], big[n // 2:], 0.5).launch(n // 2, 0), 24 * (n // 2))
    timed("value_add", lambda: P.BlockParallelValueAddFunctor(big.view(1 << 14, 1 << 14), 1.1).launch(1 << 14, 0), 16 * n)
    timed("memset", lambda: P.MemSetFunctor(big, 1.5).launch(n, 0), 8 * n)
    del big
    fa = soatl.make_hybrid_field_arrays(256, 16, [(f"f{k}", "f8") for k in range(8)])
    n = 64 << 20
    fa.resize(n)
    fa.storage.view(torch.float64).uniform_(0, 1)
    c = [0.125 * (k + 1) for k in range(8)]
    timed("soatl_rmw", lambda: fa.parallel_apply_rmw(1.0000001, c), 128 * n)
    fa.clear()
    fb = soatl.make_hybrid_field_arrays(64, 16, [("rx", "f8"), ("ry", "f8"), ("rz", "f8"), ("e", "f8")])
    n = 64 << 20
    fb.resize(n)
    fb.storage.view(torch.float64).uniform_(0.1, 1)
    timed("soatl_dist", lambda: fb.parallel_apply_dist("e", "rx", "ry", "rz", 0.05, 0.06, 0.07), 32 * n)
    fb.clear()
    # SOATL serialisation: <64,8,e,atype,rx,mid,ry,rz> -> <1,1,...> wire layout and back (tests/soatlserializetest.cpp shape)
    fields = [("e", "f8"), ("atype", "u1"), ("rx", "f8"), ("mid", "i4"), ("ry", "f8"), ("rz", "f8")]
    n = (32 << 20) + int(os.environ.get("PACK_ODD", "0"))
    src = soatl.make_hybrid_field_arrays(64, 8, fields); src.resize(n)
    src.storage.random_(0, 255)
    wire = soatl.make_hybrid_field_arrays(1, 1, fields); wire.resize(n)
    back = soatl.make_hybrid_field_arrays(64, 8, fields); back.resize(n)
    timed("soatl_pack", lambda: wire.copy_from(src), 2 * 37 * n)
    timed("soatl_unpack", lambda: back.copy_from(wire), 2 * 37 * n)
    if not only or "soatl_pack" in only:
        torch.cuda.synchronize()
        assert all(torch.equal(back[nm][:n].view(torch.uint8), src[nm][:n].view(torch.uint8)) for nm, _ in fields)
    src.clear(); wire.clear(); back.clear()
    counts = poisson_counts(128 ** 3)
    cg = cellgrid.CellGrid(counts, 64, 16, (8, 8, 8, 8), 64)
    cg.arena.view(torch.float64).uniform_(-1, 1)
    alg = 8 * int(counts.sum()) + cg.ncells * 24
    timed("cells", lambda: cg.cell_sum(3), alg, True)
    timed("cells_f0", lambda: cg.cell_sum(0), alg, True)
    timed("cells_all", lambda: cg.cell_sum_all_fields(), 32 * int(counts.sum()) + cg.ncells * 48, True)
    # raw back-to-back launches through the C ABI with preallocated outputs (the arena is 1.5 GB >> L2)
    import ctypes as C
    from onika_b200 import _capi
    from onika_b200.runtime import ptr
    sums = torch.empty(cg.ncells, dtype=torch.float64, device="cuda")
    tot = torch.zeros(1, dtype=torch.float64, device="cuda")
    fb_ = (C.c_uint32 * 4)(8, 8, 8, 8)
    def raw(field):
        _capi.check(_capi.lib().onk_cell_sum_f64(ptr(cg.arena), ptr(cg.cell_off), ptr(cg.cell_size), ptr(cg.cell_cap), cg.ncells, 64, fb_, 4, field,
                                                 ptr(sums), ptr(tot), 0))
    if not only or "cells_raw" in only:
        for field in (3, 0):
            for _ in range(3):
                raw(field)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                raw(field)
            e1.record(); e1.synchronize()
            t = e0.elapsed_time(e1) / 20 * 1e-3
            print(f"cells_raw field {field}            {t * 1e6:10.1f} us   {alg / t / 1e9:8.1f} GB/s", flush=True)
    del cg
    # the same cells in the field-major layout (onk_cell_sum_packed_f64)
    fm = cellgrid.FieldMajorCellGrid(counts, nfields=1)
    fm.columns[0].uniform_(-1, 1)
    alg_fm = 8 * fm.nparticles + fm.ncells * 16
    timed("cells_fm", lambda: fm.cell_sum(0), alg_fm, True)
    sums_fm = torch.empty(fm.ncells, dtype=torch.float64, device="cuda")
    timed("cells_fm_raw", lambda: _capi.check(_capi.lib().onk_cell_sum_packed_f64(ptr(fm.columns[0]), fm.nparticles, ptr(fm.cell_begin), fm.ncells, ptr(sums_fm), ptr(tot), 0)),
          alg_fm, True)
    # a plain stream over the same number of bytes, for reference (short kernels do not reach the 2 GiB rate)
    same = torch.rand(alg_fm // 8, dtype=torch.float64, device="cuda")
    timed("reduce_min_same_bytes", lambda: P.ReduceFunctor(ob.OP_MIN, same, o).launch(same.numel(), 0), 8 * same.numel(), True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
