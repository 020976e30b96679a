import sys, time
sys.path.insert(0, "/root/repo")
import torch, onika_b200 as ob
from onika_b200 import parallel as P
ob.init(0); ob.bind_lane_to_torch_stream(0)
n = 1 << 28
h = torch.rand(n, dtype=torch.float64).pin_memory()
out = torch.empty(1, dtype=torch.float64, device="cuda")
want = float(h.min())
for name, fn in (("staged host_reduce", lambda: P.host_reduce(ob.OP_MIN, h)),
                 ("zero-copy kernel  ", lambda: (P.ReduceFunctor(ob.OP_MIN, h, out).launch(n, 0), torch.cuda.synchronize(), out.item())[-1])):
    for _ in range(2): r = fn()
    assert r == want, (name, r, want)
    t0 = time.perf_counter()
    for _ in range(5): r = fn()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt*1e3:8.2f} ms  {n*8/dt/1e9:7.2f} GB/s")
