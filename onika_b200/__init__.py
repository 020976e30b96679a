"""onika_b200 -- B200-native data-parallel execution layer behind onika's block_parallel_for / parallel_for /
SOATL interface.  The product is libonika_b200.so (hand-written sm_100a CUDA + a C ABI, include/onika_b200.h) and the
C++ front end in include/onika/.  This Python package is the thin host-side mirror used by tests and bench.py:
PyTorch supplies device memory, streams and torch.distributed -- plumbing, not the product."""
from . import _capi
from ._capi import OnikaError, OP_MIN, OP_MAX, OP_SUM, LANE_DEFAULT, LANE_ALL
from .runtime import (init, finalize, device_info, launch_count, lane_sync, bind_lane_to_torch_stream)
from . import parallel, soatl, cellgrid

__all__ = ["_capi", "OnikaError", "OP_MIN", "OP_MAX", "OP_SUM", "LANE_DEFAULT", "LANE_ALL", "init", "finalize",
           "device_info", "launch_count", "lane_sync", "bind_lane_to_torch_stream", "parallel", "soatl", "cellgrid"]
