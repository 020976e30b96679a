import torch, sys
sys.path.insert(0, '/root/repo')
import onika_b200 as ob
from onika_b200 import parallel as P
ob.init(0); ob.bind_lane_to_torch_stream(0)
flush = torch.empty(64 << 20, dtype=torch.float64, device='cuda')
for n in (37_742_428, 1 << 25, 1 << 26, 1 << 28):
    x = torch.rand(n, dtype=torch.float64, device='cuda')
    out = torch.empty(1, dtype=torch.float64, device='cuda')
    ts = []
    for _ in range(12):
        flush.sum()           # read 512 MiB: evicts x from L2 without leaving dirty lines
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); P.reduce_min(x, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"reduce_min n={n} bytes={8*n/1e6:.0f} MB  median {ts[len(ts)//2]:.1f} us  best {ts[0]:.1f} us  -> {8*n/ts[len(ts)//2]/1e6:.0f} GB/s")
