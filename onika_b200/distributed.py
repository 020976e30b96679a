"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed over NCCL; gloo on CPU for the tests).

The path shards without any data-path collective: rank g owns the contiguous range [g*N/G, (g+1)*N/G) of elements /
cells / rows.  The only exchange is the combination of one fp64 partial per rank and per reduction -- the
counterpart of the reference's mpi::all_reduce_multi (include/onika/mpi/all_reduce_multi.h:30-37):
  * min / max : order independent -> one all_reduce(MIN|MAX) on the device tensor, on the kernel's stream;
  * sum       : all_gather of the G partials + a fold in RANK ORDER (onk_combine_f64 on the device, or the same fold on
                the host for G scalars), so the result is bit-reproducible whatever tree NCCL would use.
"""
from __future__ import annotations

from typing import Sequence, Tuple

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
    every rank.  For device tensors the collective is asynchronous on TORCH's current stream: the kernel that produced
    `partial` must have been queued on a lane bound to that stream (lane 0 after runtime.init(), or any lane passed to
    bind_lane_to_torch_stream) -- otherwise synchronise the producing lane first (lane_sync)."""
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

    def __init__(self, group=None, transport=None):
        """transport: None = the library's default (device memory over NVLink, or ONK_XCHG_TRANSPORT), "device", or "host"
        (pinned host memory in POSIX shared memory: no peer access needed, a few microseconds slower per exchange)"""
        import ctypes as C
        import torch.distributed as dist
        from ._capi import check, lib, XCHG_TRANSPORT_DEVICE, XCHG_TRANSPORT_HOST
        self._C, self._check, self._lib = C, check, lib()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.handle = C.c_void_p()
        if transport is None:
            check(self._lib.onk_xchg_create(self.rank, self.world, C.byref(self.handle)))
        else:
            t = {"device": XCHG_TRANSPORT_DEVICE, "host": XCHG_TRANSPORT_HOST}[transport]
            check(self._lib.onk_xchg_create_ex(self.rank, self.world, t, C.byref(self.handle)))
        _share_handles(lambda buf: self._lib.onk_xchg_export(self.handle, buf),
                       lambda allh: self._lib.onk_xchg_connect(self.handle, allh), group, "the exchange buffer")

    def allreduce(self, op: int, values, lane: int = -1):
        """in-place all-reduce of 1..8 doubles held in a device tensor (onk_xchg_allreduce_f64): the counterpart of
        mpi::all_reduce_multi; rank-order fold, same bits on every rank; collective"""
        from .runtime import ptr
        self._check(self._lib.onk_xchg_allreduce_f64(self.handle, op, ptr(values), values.numel(), lane))
        return values

    def barrier(self, lane: int = -1) -> None:
        """device-side barrier between the ranks' lanes (onk_xchg_barrier); collective"""
        self._check(self._lib.onk_xchg_barrier(self.handle, lane))

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


def _share_handles(export, connect, group, what: str) -> None:
    """export the local CUDA IPC handle, all-gather the handles (host plumbing), open the peers'; collective-safe: every
    rank reaches every collective even when a local step failed, then all raise together"""
    import ctypes as C
    import torch.distributed as dist
    from ._capi import ONK_IPC_HANDLE_BYTES
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return
    mine = C.create_string_buffer(ONK_IPC_HANDLE_BYTES)
    rc = export(mine)
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(mine.raw) if rc == 0 else None, group=group)
    if rc != 0 or any(g is None for g in gathered):
        raise RuntimeError(f"CUDA IPC export of {what} failed on some rank")
    rc = connect(b"".join(gathered))
    oks = [None] * world
    dist.all_gather_object(oks, rc == 0, group=group)
    if not all(oks):
        raise RuntimeError(f"CUDA IPC open of a peer's {what} failed on some rank (no peer access?)")


class PeerWindow:
    """`nbytes` of device memory of this rank mapped into every peer process (onk_window_*): the landing zone of
    IdRedistribution.move().  Collective constructor."""

    def __init__(self, nbytes: int, exchange: "PeerExchange", group=None):
        import ctypes as C
        from ._capi import check, lib
        self._C, self._check, self._lib = C, check, lib()
        self.nbytes = int(nbytes)
        self.handle = C.c_void_p()
        check(self._lib.onk_window_create(self.nbytes, exchange.rank, exchange.world, C.byref(self.handle)))
        _share_handles(lambda buf: self._lib.onk_window_export(self.handle, buf),
                       lambda allh: self._lib.onk_window_connect(self.handle, allh), group, "an output window")

    def tensor(self, dtype, shape=None):
        """the local window viewed as a torch tensor (no copy)"""
        import torch
        C = self._C
        p = C.c_void_p()
        self._check(self._lib.onk_window_ptr(self.handle, C.byref(p), None))
        itemsize = torch.empty(0, dtype=dtype).element_size()
        n = self.nbytes // itemsize
        if n == 0:
            return torch.empty(shape if shape is not None else (0,), dtype=dtype, device="cuda")
        from .runtime import tensor_from_ptr
        t = tensor_from_ptr(p.value, n, dtype, owner=self)
        return t.reshape(shape) if shape is not None else t

    def close(self):
        if self.handle:
            self._check(self._lib.onk_window_destroy(self.handle))
            self.handle = None


class IdRedistribution:
    """Id-based redistribution of elements between the GPUs of a node: the counterpart of the reference's
    communication_scheme_from_ids + data_move (include/onika/mpi/xs_data_move.h:31-50, xs_data_move_impl.h:44-113).

        mv = IdRedistribution(exchange, id_min, id_max)            # collective
        mv.plan(ids_before, ids_after)                             # collective; raises on every rank if ids do not match up
        out = PeerWindow(n_after * elem_bytes, exchange)           # collective
        mv.move(values_before, out)                                # collective: out.tensor(...)[j] = value of the same id

    Every id of [id_min, id_max) must appear exactly once before and once after (the reference aborts otherwise)."""

    ERRORS = {1: "an id of the range has no source", 2: "an id has no destination", 3: "an id appears twice before",
              4: "an id appears twice after", 5: "an id lies outside the id range"}

    def __init__(self, exchange: "PeerExchange", id_min: int, id_max: int, group=None):
        import ctypes as C
        from ._capi import check, lib
        self._C, self._check, self._lib = C, check, lib()
        self.exchange, self.group = exchange, group
        self.handle = C.c_void_p()
        check(self._lib.onk_xs_plan_create(exchange.handle, int(id_min), int(id_max), C.byref(self.handle)))
        _share_handles(lambda buf: self._lib.onk_xs_plan_export(self.handle, buf),
                       lambda allh: self._lib.onk_xs_plan_connect(self.handle, allh), group, "the id directory")
        self.n_before = self.n_after = 0

    def plan(self, ids_before, ids_after, lane: int = -1) -> None:
        import torch
        from .runtime import ptr
        C = self._C
        for t in (ids_before, ids_after):
            if t.dtype not in (torch.int64, torch.uint64) or not t.is_cuda or not t.is_contiguous():
                raise ValueError("ids must be contiguous 64-bit integer CUDA tensors")
        status = C.c_int(0)
        rc = self._lib.onk_xs_plan_build(self.handle, ptr(ids_before), ids_before.numel(), ptr(ids_after), ids_after.numel(), lane, C.byref(status))
        if rc != 0 and status.value != 0:
            raise ValueError("id redistribution: " + self.ERRORS.get(status.value, f"error {status.value}"))
        self._check(rc)
        self.n_before, self.n_after = ids_before.numel(), ids_after.numel()

    def send_counts(self):
        """elements this rank sends to each rank (the reference's send_count)"""
        C = self._C
        buf = (C.c_uint64 * self.exchange.world)()
        self._check(self._lib.onk_xs_plan_send_counts(self.handle, buf))
        return [int(v) for v in buf]

    def move(self, values, out: PeerWindow, lane: int = -1) -> None:
        from .runtime import ptr
        if values.shape[0] != self.n_before or not values.is_contiguous():
            raise ValueError("values must be contiguous with one row per local element before")
        elem = values.element_size() * (values.numel() // max(values.shape[0], 1)) if values.shape[0] else values.element_size()
        self._check(self._lib.onk_xs_move(self.handle, out.handle, elem, ptr(values), lane))

    def move_fields(self, values_list, outs, lane: int = -1) -> None:
        """several arrays that travel with the same ids (onk_xs_move_fields): one pair of barriers for all of them"""
        from .runtime import ptr
        C = self._C
        k = len(values_list)
        if k == 0 or k != len(outs):
            raise ValueError("one output window per array")
        elems = []
        for v in values_list:
            if v.shape[0] != self.n_before or not v.is_contiguous():
                raise ValueError("values must be contiguous with one row per local element before")
            elems.append(v.element_size() * (v.numel() // max(v.shape[0], 1)) if v.shape[0] else v.element_size())
        c_outs = (C.c_void_p * k)(*[o.handle for o in outs])
        c_elems = (C.c_uint64 * k)(*elems)
        c_ins = (C.c_void_p * k)(*[ptr(v) for v in values_list])
        self._check(self._lib.onk_xs_move_fields(self.handle, k, c_outs, c_elems, c_ins, lane))

    def close(self):
        if self.handle:
            self._check(self._lib.onk_xs_plan_destroy(self.handle))
            self.handle = None
