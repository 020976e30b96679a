"""A/B of the CTA shapes of onk_cell_sum_fields_f64 (ONK_CELLF_SHAPE) at 128^3 and 256^3 cells."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import onika_b200 as ob
from onika_b200 import parallel as P, cellgrid
from tests._inputs import poisson_counts
ob.init(0)
flush = torch.ones(1 << 27, dtype=torch.float64, device="cuda"); fo = torch.empty(1, dtype=torch.float64, device="cuda")
for side in (128, 256):
    counts = poisson_counts(side ** 3)
    cg = cellgrid.CellGrid(counts, 64, 16, (8, 8, 8, 8), 64); cg.arena.view(torch.float64).uniform_(-1, 1)
    ts = []
    for _ in range(9):
        P.ReduceFunctor(ob.OP_MIN, flush, fo).launch(flush.numel(), 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cg.cell_sum_all_fields(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:]); t = ts[len(ts)//2] * 1e-3
    print(f"ONK_CELLF_SHAPE={os.environ.get('ONK_CELLF_SHAPE','0')} side {side}: {t*1e6:.1f} us arena stream {(cg.arena_bytes + cg.ncells*48)/t/1e9:.0f} GB/s", flush=True)
    del cg
