import os, sys, torch
sys.path.insert(0, os.getcwd())
import onika_b200 as ob
from onika_b200 import parallel as P, cellgrid
from tests._inputs import poisson_counts
ob.init(0)
flush = torch.ones(1 << 27, dtype=torch.float64, device="cuda"); fo = torch.empty(1, dtype=torch.float64, device="cuda")
for side in (128, 256):
    counts = poisson_counts(side ** 3)
    fm = cellgrid.FieldMajorCellGrid(counts, nfields=1); fm.columns[0].uniform_(-1, 1)
    alg = 8 * fm.nparticles + fm.ncells * 16
    ts = []
    for _ in range(9):
        P.ReduceFunctor(ob.OP_MIN, flush, fo).launch(flush.numel(), 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fm.cell_sum(0); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:]); t = ts[len(ts)//2] * 1e-3
    print(f"ONK_PK_FOLD={os.environ.get('ONK_PK_FOLD','0')} side {side}: {t*1e6:.1f} us {alg/t/1e9:.0f} GB/s (min {ts[0]*1e3:.1f} max {ts[-1]*1e3:.1f} us)", flush=True)
    del fm
