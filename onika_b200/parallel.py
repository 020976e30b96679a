"""Host-side mirror of onika::parallel (include/onika/parallel in the reference): functors, parallel_for /
block_parallel_for, ParallelExecutionQueue with lanes and the `<<` operators.

The reference type-erases an arbitrary C++ functor into the ParallelExecutionContext; in Python the functors are the
recognised patterns that have a typed sm_100a fast path in libonika_b200.so (arbitrary user functors are the business
of the C++ front end, include/onika/parallel/block_parallel_for.h in this repo).  Execution semantics follow the
reference: an operation that is not pushed onto a queue executes synchronously when its wrapper is released
(parallel_execution_operators.h:125-146); operations pushed with `queue << set_lane(i) << op` are scheduled at
`<< flush` in FIFO order per lane and may overlap across lanes (parallel_execution_queue.cpp:242-271).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, List, Optional

from . import _capi
from ._capi import check, lib, OP_MIN, OP_MAX, OP_SUM, LANE_DEFAULT
from .runtime import ptr

DEFAULT_EXECUTION_LANE = -1      # include/onika/parallel/constants.h:36
UNDEFINED_EXECUTION_LANE = -2    # :37
MAX_EXECUTION_LANES = 256        # :38
ONIKA_AUTO_LANE_CYCLE_HINT = 16  # include/onika/parallel/parallel_execution_queue.h:30-32


# ------------------------------------------------------------------ functors (typed fast paths)
@dataclass
class AxpyFunctor:
    """y[i] += a*x[i]   (config C1; element functor for parallel_for)"""
    y: Any
    x: Any
    a: float

    def launch(self, n: int, lane: int) -> None:
        check(lib().onk_axpy_f64(ptr(self.y), ptr(self.x), self.a, n, lane))


@dataclass
class MemSetFunctor:
    """data[i] = value   (reference: include/onika/parallel/memset.h:30-40); fp64 here"""
    data: Any
    value: float

    def launch(self, n: int, lane: int) -> None:
        import struct
        pattern = struct.unpack("<Q", struct.pack("<d", self.value))[0]
        check(lib().onk_parallel_memset(ptr(self.data), n, pattern, 8, lane))


@dataclass
class BlockParallelValueAddFunctor:
    """a[i][j] += v, row i per block (reference: extras/block_parallel_value_add_functor.h:15-58)"""
    array: Any          # 2-D, row-major, contiguous
    value: float

    def launch(self, n: int, lane: int) -> None:
        rows, cols = self.array.shape
        assert n <= rows
        check(lib().onk_value_add_f64(ptr(self.array), n, cols, self.value, lane))


@dataclass
class ReduceFunctor:
    """min / max / sum of data[0:n] into out[0] (device).  MIN follows tests/gpu_reduce_min_fp64.cu:46-55."""
    op: int
    data: Any
    out: Any

    def launch(self, n: int, lane: int) -> None:
        check(lib().onk_reduce_f64(self.op, ptr(self.data), n, ptr(self.out), lane))


# ------------------------------------------------------------------ execution context / wrapper
@dataclass
class ParallelExecutionContext:
    """What the reference keeps per operation (parallel_execution_context.h:87-210), reduced to what the typed paths need."""
    tag: str = ""
    sub_tag: str = ""
    default_queue: Optional["ParallelExecutionQueue"] = None
    lane: int = UNDEFINED_EXECUTION_LANE
    functor: Any = None
    n: int = 0
    user_cb: Optional[Callable[[], None]] = None


class ParallelExecutionWrapper:
    """Move-only holder (parallel_execution_operators.h:40-58): executes synchronously on the default queue when
    released without having been pushed onto a queue."""

    def __init__(self, pec: ParallelExecutionContext):
        self.pec: Optional[ParallelExecutionContext] = pec

    def take(self) -> ParallelExecutionContext:
        pec, self.pec = self.pec, None
        if pec is None:
            raise RuntimeError("parallel operation was already consumed (wrappers are move-only)")
        return pec

    def run(self) -> None:
        """explicit form of the wrapper destructor: wait, enqueue on the default lane, schedule, wait"""
        if self.pec is None:
            return
        q = self.pec.default_queue or ParallelExecutionQueue.default_queue()
        q.wait()
        q.set_lane(DEFAULT_EXECUTION_LANE)
        q.enqueue(self.take())
        q.schedule_all()
        q.wait()

    def __del__(self):
        try:
            self.run()
        except Exception:
            pass


def block_parallel_for(n: int, functor, pec: Optional[ParallelExecutionContext] = None) -> ParallelExecutionWrapper:
    """block_parallel_for(N, functor, exec_ctx)   (reference: include/onika/parallel/block_parallel_for.h:163-173)"""
    pec = pec or ParallelExecutionContext()
    pec.functor, pec.n = functor, int(n)
    return ParallelExecutionWrapper(pec)


def parallel_for(n: int, functor, pec: Optional[ParallelExecutionContext] = None) -> ParallelExecutionWrapper:
    """parallel_for(N, functor, exec_ctx)   (reference: include/onika/parallel/parallel_for.h:91-112)"""
    return block_parallel_for(n, functor, pec)


def parallel_memset(data, n: int, value: float, pec: Optional[ParallelExecutionContext] = None) -> ParallelExecutionWrapper:
    """parallel_memset(data, N, value, exec_ctx)   (reference: include/onika/parallel/memset.h:47-51)"""
    return parallel_for(n, MemSetFunctor(data, value), pec)


# ------------------------------------------------------------------ queue + stream operators
@dataclass
class set_lane:
    lane: int = UNDEFINED_EXECUTION_LANE
    cycle: int = ONIKA_AUTO_LANE_CYCLE_HINT


def any_lane(cycle: int = ONIKA_AUTO_LANE_CYCLE_HINT) -> set_lane:
    return set_lane(UNDEFINED_EXECUTION_LANE, cycle)


class _Flush:
    pass


class _Synchronize:
    pass


flush = _Flush()
synchronize = _Synchronize()


class ParallelExecutionQueueBase:
    """Delay queue (parallel_execution_queue.h:66-87): holds operations with their lanes until spliced into a real queue."""

    def __init__(self):
        self.m_lane = DEFAULT_EXECUTION_LANE
        self.m_auto_lane_cycle = ONIKA_AUTO_LANE_CYCLE_HINT
        self.m_queue_list: List[ParallelExecutionContext] = []

    def set_lane(self, lane: int, cycle: int = ONIKA_AUTO_LANE_CYCLE_HINT) -> None:
        self.m_lane = lane
        self.m_auto_lane_cycle = min(32, cycle)   # ONIKA_PARALLEL_QUEUE_MAX_LANES, parallel_execution_queue.cpp:74-78

    def enqueue(self, pec: ParallelExecutionContext, from_other_queue: bool = False) -> None:
        if not from_other_queue:
            pec.lane = self.m_lane               # parallel_execution_queue.cpp:94-111
        self.m_queue_list.append(pec)

    def __lshift__(self, item):
        if isinstance(item, ParallelExecutionWrapper):
            self.enqueue(item.take())
        elif isinstance(item, set_lane):
            self.set_lane(item.lane, item.cycle)
        elif isinstance(item, ParallelExecutionQueueBase):
            while item.m_queue_list:              # splice, keeping each operation's lane (parallel_execution_operators.h:78-89)
                self.enqueue(item.m_queue_list.pop(0), from_other_queue=True)
        else:
            raise TypeError(f"cannot push {type(item)} onto a parallel execution queue")
        return self


class ParallelExecutionQueue(ParallelExecutionQueueBase):
    _default: Optional["ParallelExecutionQueue"] = None

    def __init__(self):
        super().__init__()
        self.m_exec_list: List[ParallelExecutionContext] = []

    @classmethod
    def default_queue(cls) -> "ParallelExecutionQueue":
        if cls._default is None:
            cls._default = cls()
        return cls._default

    def _pre_process_queue(self) -> None:
        # automatic lane assignment for any_lane() operations: round robin (parallel_execution_queue.cpp:151-169)
        cycle = 0
        for pec in self.m_queue_list:
            if pec.lane == DEFAULT_EXECUTION_LANE:
                pec.lane = 0
            if pec.lane == UNDEFINED_EXECUTION_LANE and self.m_lane == UNDEFINED_EXECUTION_LANE:
                pec.lane = cycle
                cycle = (cycle + 1) % self.m_auto_lane_cycle

    def schedule(self, pec: ParallelExecutionContext) -> None:
        lane = pec.lane
        if lane in (UNDEFINED_EXECUTION_LANE, DEFAULT_EXECUTION_LANE):
            lane = 0                              # parallel_execution_queue.cpp:283-290
        if not (0 <= lane < MAX_EXECUTION_LANES):
            raise ValueError(f"Invalid execution stream id ({lane})")
        pec.lane = lane
        pec.functor.launch(pec.n, lane)

    def schedule_all(self, lane: int = UNDEFINED_EXECUTION_LANE) -> None:
        self._pre_process_queue()
        keep = []
        for pec in self.m_queue_list:
            if lane < 0 or pec.lane == lane:
                self.schedule(pec)
                self.m_exec_list.append(pec)
            else:
                keep.append(pec)
        self.m_queue_list = keep

    def wait(self, lane: int = UNDEFINED_EXECUTION_LANE) -> None:
        lanes = sorted({p.lane for p in self.m_exec_list if lane < 0 or p.lane == lane})
        for l in lanes:
            check(lib().onk_lane_sync(l))
        done = [p for p in self.m_exec_list if lane < 0 or p.lane == lane]
        self.m_exec_list = [p for p in self.m_exec_list if not (lane < 0 or p.lane == lane)]
        for p in done:
            if p.user_cb:
                p.user_cb()

    def query_status(self, lane: int = UNDEFINED_EXECUTION_LANE) -> bool:
        if self.m_queue_list:
            return False
        for l in sorted({p.lane for p in self.m_exec_list if lane < 0 or p.lane == lane}):
            idle = C.c_int()
            check(lib().onk_lane_query(l, C.byref(idle)))
            if not idle.value:
                return False
        self.wait(lane)
        return True

    def empty(self) -> bool:
        return not self.m_queue_list and not self.m_exec_list

    def __lshift__(self, item):
        if item is flush:
            self.schedule_all()
            self.set_lane(DEFAULT_EXECUTION_LANE)   # parallel_execution_operators.h:107-112
            return self
        if item is synchronize:
            self.wait()
            return self
        return super().__lshift__(item)


# ------------------------------------------------------------------ direct typed calls (device-resident data)
def reduce(op: int, data, n: Optional[int] = None, out=None, lane: int = LANE_DEFAULT):
    """Asynchronous device reduction; returns the 1-element device tensor that will hold the result."""
    import torch
    if n is None:
        n = data.numel()
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=data.device)
    check(lib().onk_reduce_f64(op, ptr(data), n, ptr(out), lane))
    return out


def reduce_min(data, **kw):
    return reduce(OP_MIN, data, **kw)


def reduce_max(data, **kw):
    return reduce(OP_MAX, data, **kw)


def reduce_sum(data, **kw):
    return reduce(OP_SUM, data, **kw)


def combine(op: int, partials, out=None, lane: int = LANE_DEFAULT):
    """Fixed (index) order combine of per-rank partials; counterpart of mpi::all_reduce_multi."""
    import torch
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=partials.device)
    check(lib().onk_combine_f64(op, ptr(partials), partials.numel(), ptr(out), lane))
    return out


# ------------------------------------------------------------------ host-buffer calls (numpy / pinned tensors)
def host_reduce(op: int, host_array) -> float:
    out = C.c_double()
    n = host_array.numel() if hasattr(host_array, "numel") else host_array.size
    check(lib().onk_host_reduce_f64(op, ptr(host_array), n, C.cast(C.byref(out), C.c_void_p)))
    return out.value


def host_axpy(y_host, x_host, a: float) -> None:
    n = y_host.numel() if hasattr(y_host, "numel") else y_host.size
    check(lib().onk_host_axpy_f64(ptr(y_host), ptr(x_host), a, n))


def host_value_add(array_host, value: float) -> None:
    rows, cols = array_host.shape
    check(lib().onk_host_value_add_f64(ptr(array_host), rows, cols, value))
