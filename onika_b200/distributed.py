"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed over NCCL; gloo on CPU for the tests).

The path shards without any data-path collective: rank g owns the contiguous range [g*N/G, (g+1)*N/G) of elements /
cells / rows.  The only exchange is the combination of one fp64 partial per rank and per reduction -- the
counterpart of the reference's mpi::all_reduce_multi (include/onika/mpi/all_reduce_multi.h:30-37):
  * min / max : order independent -> one all_reduce(MIN|MAX) on the device tensor, on the kernel's stream;
  * sum       : all_gather of the G partials + a fold in RANK ORDER (onk_combine_f64 on the device, or the same fold on
                the host for G scalars), so the result is bit-reproducible whatever tree NCCL would use.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from ._capi import OP_MIN, OP_MAX, OP_SUM

IDENTITY = {OP_MIN: 1.7976931348623157e308, OP_MAX: -1.7976931348623157e308, OP_SUM: 0.0}


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block partition [begin, end) of n units over `world` ranks"""
    assert 0 <= rank < world
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_cells_z(nz: int, rank: int, world: int) -> Tuple[int, int]:
    """z-slab partition of a cell grid (no halo: per-cell work is local)"""
    return shard_range(nz, rank, world)


def fold_partials(op: int, partials: Sequence[float]) -> float:
    """rank-order fold with the reference's comparison semantics (tests/gpu_reduce_min_fp64.cu:48-53)"""
    acc = IDENTITY[op]
    for v in partials:
        if op == OP_MIN:
            acc = v if v < acc else acc
        elif op == OP_MAX:
            acc = v if v > acc else acc
        else:
            acc = acc + v
    return acc


def combine_across_ranks(op: int, partial, group=None):
    """partial: 1-element float64 tensor (device or CPU).  Returns a 1-element tensor holding the global result on
    every rank.  Asynchronous on the current stream for device tensors."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial
    if op in (OP_MIN, OP_MAX):
        dist.all_reduce(partial, op=dist.ReduceOp.MIN if op == OP_MIN else dist.ReduceOp.MAX, group=group)
        return partial
    world = dist.get_world_size(group)
    gathered = torch.empty(world, dtype=torch.float64, device=partial.device)
    dist.all_gather_into_tensor(gathered, partial, group=group) if partial.is_cuda else \
        dist.all_gather(list(gathered.split(1)), partial, group=group)
    if partial.is_cuda:
        from . import parallel
        return parallel.combine(OP_SUM, gathered, out=partial)
    partial[0] = fold_partials(OP_SUM, [float(v) for v in gathered])
    return partial


class PeerExchange:
    """Fused reduction + cross-GPU combine over NVLink peer memory (onk_reduce_xchg_f64).

    One exchange buffer per rank in device memory, shared between the processes of the node through CUDA IPC: the
    handles travel once through torch.distributed (host-side plumbing); afterwards every reduction is ONE kernel per
    rank -- its last CTA stores the rank's partial into every peer's buffer, publishes a release flag, waits for the
    peers' flags and folds the partials in rank order.  No NCCL call on the data path.  All ranks must issue the same
    sequence of reduce() calls."""

    def __init__(self, group=None):
        import ctypes as C
        import torch.distributed as dist
        from ._capi import check, lib, ONK_IPC_HANDLE_BYTES
        self._C, self._check, self._lib = C, check, lib()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.handle = C.c_void_p()
        check(self._lib.onk_xchg_create(self.rank, self.world, C.byref(self.handle)))
        if self.world > 1:
            mine = C.create_string_buffer(ONK_IPC_HANDLE_BYTES)
            rc = self._lib.onk_xchg_export(self.handle, mine)
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine.raw) if rc == 0 else None, group=group)   # collective: always reached
            if rc != 0 or any(g is None for g in gathered):
                raise RuntimeError("CUDA IPC export of the exchange buffer failed on some rank")
            rc = self._lib.onk_xchg_connect(self.handle, b"".join(gathered))
            oks = [None] * self.world
            dist.all_gather_object(oks, rc == 0, group=group)          # every rank has mapped every buffer before first use
            if not all(oks):
                raise RuntimeError("CUDA IPC open of a peer exchange buffer failed on some rank (no peer access?)")

    def reduce(self, op: int, data, n=None, out=None, lane: int = -1):
        import torch
        from .runtime import ptr
        if n is None:
            n = data.numel()
        if out is None:
            out = torch.empty(1, dtype=torch.float64, device=data.device)
        self._check(self._lib.onk_reduce_xchg_f64(op, ptr(data), n, ptr(out), self.handle, lane))
        return out

    def close(self):
        if self.handle:
            self._check(self._lib.onk_xchg_destroy(self.handle))
            self.handle = None
