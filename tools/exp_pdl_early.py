"""Back-to-back reductions of one shard (what a rank does in bench.py's strong-scaling section) under the ONK_PDL_EARLY modes:
0 = dependents start when the predecessor completed, 1 = released early (resident while the predecessor drains), 2 = released
early + L2 prefetch of their first tiles.  Also the two-lane C5-like pattern (value-add, value-add, min, sum per lane).
  for m in 0 1 2; do ONK_PDL_EARLY=$m python tools/exp_pdl_early.py; done"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import onika_b200 as ob
from onika_b200 import parallel as P
ob.init(0)
mode = os.environ.get("ONK_PDL_EARLY", "0")
out = torch.empty(4, dtype=torch.float64, device="cuda")
res = []
for lg in (24, 25, 26, 28):
    n = 1 << lg
    d = torch.rand(n, dtype=torch.float64, device="cuda")
    want = float(d.min())
    f = P.ReduceFunctor(ob.OP_MIN, d, out[:1])
    for _ in range(10): f.launch(n, 0)
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f.launch(n, 0)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) / 50 * 1e3)
    assert float(out[0]) == want
    ts.sort()
    res.append(f"2^{lg} ({n*8>>20} MiB): {ts[2]:.1f} us (best {ts[0]:.1f})")
    del d
print(f"ONK_PDL_EARLY={mode} single lane:", " | ".join(res), flush=True)
# two lanes, C5-like: per lane value-add x2, min, sum over an 8192 x 8192 array
rows = 8192
a = [torch.rand(rows, rows, dtype=torch.float64, device="cuda") for _ in range(2)]
def sequence():
    for lane in (0, 1):
        P.BlockParallelValueAddFunctor(a[lane], 1.1).launch(rows, lane)
        P.BlockParallelValueAddFunctor(a[lane], 1.2).launch(rows, lane)
        P.ReduceFunctor(ob.OP_MIN, a[lane].view(-1), out[2 * lane:2 * lane + 1]).launch(rows * rows, lane)
        P.ReduceFunctor(ob.OP_SUM, a[lane].view(-1), out[2 * lane + 1:2 * lane + 2]).launch(rows * rows, lane)
import time
for _ in range(3): sequence()
ob.lane_sync()
torch.cuda.synchronize()
ts = []
for rep in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): sequence()
    ob.lane_sync()
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / 10 * 1e3)
ts.sort()
print(f"ONK_PDL_EARLY={mode} two lanes (2 x 8192^2: add, add, min, sum): {ts[3]:.4f} ms per sequence (best {ts[0]:.4f})", flush=True)
