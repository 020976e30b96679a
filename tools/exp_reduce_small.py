"""Which reduction kernel for medium inputs (the per-GPU shard of a strong-scaling run)?  ONK_RED_KERNEL=ldg forces the
register-staged kernel; default = TMA-staged above 2 tiles per SM.  Back-to-back launches (as bench.py's strong section)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import onika_b200 as ob
from onika_b200 import parallel as P
ob.init(0)
out = torch.empty(1, dtype=torch.float64, device="cuda")
for lg in (23, 24, 25, 26, 27):
    n = 1 << lg
    d = torch.rand(n, dtype=torch.float64, device="cuda")
    f = P.ReduceFunctor(ob.OP_MIN, d, out)
    for _ in range(5): f.launch(n, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): f.launch(n, 0)
    e1.record(); e1.synchronize()
    t = e0.elapsed_time(e1) / 50 * 1e-3
    print(f"ONK_RED_KERNEL={os.environ.get('ONK_RED_KERNEL','tma')} ONK_RED_CTAS={os.environ.get('ONK_RED_CTAS','-')} 2^{lg} doubles ({n*8>>20} MiB): {t*1e6:.1f} us  {n*8/t/1e9:.0f} GB/s", flush=True)
