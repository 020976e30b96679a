"""bench.py's strong-scaling section (2^28 and 2^31 doubles in total) and its C5 lane sequence, alone, under torchrun: the A/B
harness for launch-gap work (ONK_PDL_EARLY modes) that does not need the rest of the bench line.
  for m in 0 2; do ONK_PDL_EARLY=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29511 tools/exp_strong.py; done"""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist
import onika_b200 as ob
from onika_b200 import parallel as P, distributed as D
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ob.init(local)
ob.bind_lane_to_torch_stream(0)
xchg, lanes = None, None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    made = [D.PeerExchange(transport="device") for _ in range(3)]
    xchg, lanes = made[0], (made[1], made[2])
K, W = int(os.environ.get("K", "20")), 5
out = {"ONK_PDL_EARLY": os.environ.get("ONK_PDL_EARLY", "default"), "world": world}
s = bench.strong_scaling(K, W, world, rank, torch, dist, ob, P, D, xchg)
out["strong"] = {k: {q: v[q] for q in ("T1_ms", "TN_ms", "speedup")} for k, v in s.items() if isinstance(v, dict)}
c = bench.c5_lane_sequence(K, W, world, rank, torch, dist, ob, P, D, lanes)
out["C5"] = {q: c[q] for q in ("T1_ms", "TN_ms", "speedup")}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
