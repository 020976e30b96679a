"""Context, lanes and small helpers over the C ABI."""
from __future__ import annotations

import ctypes as C

from . import _capi
from ._capi import check, lib


def init(device: int = 0, bind_torch_stream: bool = True) -> None:
    """onk_init: one process drives one device (reference: init_cuda device selection, init_cuda.cpp:84-111).

    Ordering with torch: the library's lanes are non-blocking streams of their own, so nothing orders a lane's kernels with
    torch operations on the same tensors unless they share a stream.  By default lane 0 (the default lane of every Python
    call here) is therefore bound to torch's current stream when torch is in use: torch ops, lane-0 kernels and
    torch.distributed collectives (which run on torch's current stream) then execute in program order.  Work put on OTHER
    lanes must be ordered explicitly (lane_sync, onk_lane_wait_lane, or bind that lane too).  bind_torch_stream=False keeps
    lane 0 on its own stream."""
    check(lib().onk_init(device))
    if bind_torch_stream:
        import sys
        torch = sys.modules.get("torch")
        # only when torch drives the same device (one process per GPU: torch.cuda.set_device(local) before init(local))
        if torch is not None and torch.cuda.is_available() and torch.cuda.current_device() == device:
            bind_lane_to_torch_stream(0)


def finalize() -> None:
    check(lib().onk_finalize())


def device_info() -> dict:
    dev, sms, mem = C.c_int(), C.c_int(), C.c_size_t()
    name = C.create_string_buffer(256)
    check(lib().onk_device_info(C.byref(dev), C.byref(sms), C.byref(mem), name, 256))
    return {"device": dev.value, "sm_count": sms.value, "total_mem": mem.value, "name": name.value.decode()}


def launch_count() -> int:
    n = C.c_uint64()
    check(lib().onk_launch_count(C.byref(n)))
    return n.value


def lane_sync(lane: int = _capi.LANE_ALL) -> None:
    check(lib().onk_lane_sync(lane))


def lane_idle(lane: int = _capi.LANE_ALL) -> bool:
    idle = C.c_int()
    check(lib().onk_lane_query(lane, C.byref(idle)))
    return bool(idle.value)


def lane_stream(lane: int) -> int:
    s = C.c_void_p()
    check(lib().onk_lane_stream(lane, C.byref(s)))
    return s.value or 0


def bind_lane_to_torch_stream(lane: int, stream=None) -> None:
    """Run `lane` on a torch stream (default: torch's current stream).  This is what ORDERS the lane's kernels with torch
    operations and torch.distributed collectives on the same tensors (they all run on that stream, in program order); it also
    makes torch.cuda.Event timing see the kernels.  init() does it for lane 0; a lane that is not bound runs concurrently
    with torch's stream and needs explicit synchronisation."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    check(lib().onk_lane_bind_stream(lane, C.c_void_p(stream.cuda_stream)))


def ptr(t) -> C.c_void_p:
    """Device (or pinned host) pointer of a torch tensor / numpy array / int."""
    if isinstance(t, int):
        return C.c_void_p(t)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    if hasattr(t, "ctypes"):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(f"cannot take the address of {type(t)}")


class _DeviceBytes:
    """raw device memory owned by the library, exposed through __cuda_array_interface__ so torch can view it"""

    def __init__(self, address: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 2}
        self._owner = owner          # keeps the allocation's handle alive as long as the view lives


def tensor_from_ptr(address: int, count: int, dtype, owner=None):
    """`count` elements of `dtype` at device address `address` as a torch tensor (no copy)"""
    import torch
    itemsize = torch.empty(0, dtype=dtype).element_size()
    raw = torch.as_tensor(_DeviceBytes(int(address), count * itemsize, owner), device="cuda")
    return raw.view(dtype)
