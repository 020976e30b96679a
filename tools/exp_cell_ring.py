"""A/B of the ring shapes of onk_cell_sum_f64 (ONK_CELL_SHAPE, ONK_CELL_KERNEL=ldg for the register-staged kernel) at 128^3 and
256^3 cells, with an exact check (integer-valued particles: any order gives the same sum)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import onika_b200 as ob
from onika_b200 import parallel as P, cellgrid
from tests._inputs import poisson_counts
ob.init(0)
flush = torch.ones(1 << 27, dtype=torch.float64, device="cuda"); fo = torch.empty(1, dtype=torch.float64, device="cuda")
for side in (128, 256):
    counts = poisson_counts(side ** 3)
    if side == 128:
        counts[7] = 0; counts[100] = 64; counts[5000] = 33
    cg = cellgrid.CellGrid(counts, 64, 16, (8, 8, 8, 8), 64)
    a = cg.arena.view(torch.float64)
    a.copy_(torch.randint(-8, 9, (a.numel(),), device="cuda", dtype=torch.int32).to(torch.float64))
    sums, total = cg.cell_sum(3)
    torch.cuda.synchronize()
    # expected: per-cell sum of the first `size` doubles of the e column
    off = torch.from_numpy((cg.field_byte_offsets(3) // np.uint64(8)).astype(np.int64)).cuda()
    size = torch.from_numpy(counts.astype(np.int64)).cuda()
    idx = torch.arange(64, device="cuda")[None, :]
    want = torch.zeros(cg.ncells, dtype=torch.float64, device="cuda")
    for lo in range(0, cg.ncells, 1 << 21):
        hi = min(cg.ncells, lo + (1 << 21))
        g = (off[lo:hi, None] + idx).clamp_(max=a.numel() - 1)
        v = a[g] * (idx < size[lo:hi, None])
        want[lo:hi] = v.sum(dim=1)
    ok = bool(torch.equal(sums, want)) and float(total.item()) == float(want.sum().item())
    alg = 8 * int(counts.sum()) + cg.ncells * 24
    ts = []
    for _ in range(9):
        P.ReduceFunctor(ob.OP_MIN, flush, fo).launch(flush.numel(), 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cg.cell_sum(3); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:]); t = ts[len(ts)//2] * 1e-3
    print(f"ONK_CELL_KERNEL={os.environ.get('ONK_CELL_KERNEL','async')} ONK_CELL_SHAPE={os.environ.get('ONK_CELL_SHAPE','0')} side {side}: {t*1e6:.1f} us {alg/t/1e9:.0f} GB/s algorithmic, exact={'ok' if ok else 'MISMATCH'}", flush=True)
    del cg, a, want
    torch.cuda.empty_cache()
