"""ctypes binding of the C ABI declared in include/onika_b200.h (libonika_b200.so, built in-tree by _build.py).

There is no fallback: if the shared library is missing it is built with nvcc; if that fails, importing raises.
Compute entry points fail with OnikaError(ONK_ERR_NO_DEVICE) on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import re

from . import _build

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "onika_b200.h")

OK, ERR_NO_DEVICE, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5
LANE_DEFAULT, LANE_ALL, MAX_LANES = -1, -2, 256
MEM_DEVICE, MEM_MANAGED, MEM_PINNED = 0, 1, 2
OP_MIN, OP_MAX, OP_SUM = 0, 1, 2
ONK_IPC_HANDLE_BYTES, ONK_XCHG_MAX_WORLD = 64, 16
XCHG_TRANSPORT_DEVICE, XCHG_TRANSPORT_HOST = 0, 1

vp, u64, u32, i32, dbl, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_size_t
P = C.POINTER

# name -> (restype, argtypes).  Kept in one table so tests can check it against the header.
PROTOTYPES = {
    "onk_abi_version": (i32, []),
    "onk_last_error": (C.c_char_p, []),
    "onk_device_count": (i32, [P(i32)]),
    "onk_init": (i32, [i32]),
    "onk_finalize": (i32, []),
    "onk_device_info": (i32, [P(i32), P(i32), P(sz), C.c_char_p, sz]),
    "onk_device_sync": (i32, []),
    "onk_lane_stream": (i32, [i32, P(vp)]),
    "onk_lane_bind_stream": (i32, [i32, vp]),
    "onk_lane_sync": (i32, [i32]),
    "onk_lane_query": (i32, [i32, P(i32)]),
    "onk_lane_wait_lane": (i32, [i32, i32]),
    "onk_lane_host_func": (i32, [i32, vp, vp]),
    "onk_alloc": (i32, [sz, sz, i32, P(vp)]),
    "onk_free": (i32, [vp, sz]),
    "onk_is_gpu_addressable": (i32, [vp, P(i32)]),
    "onk_memcpy": (i32, [vp, vp, sz, i32]),
    "onk_memset_bytes": (i32, [vp, i32, sz, i32]),
    "onk_prefetch": (i32, [vp, sz, i32, i32]),
    "onk_op_create": (i32, [P(vp)]),
    "onk_op_destroy": (i32, [vp]),
    "onk_op_device_scratch": (i32, [vp, P(vp)]),
    "onk_op_begin": (i32, [vp, i32, vp, sz, i32]),
    "onk_op_end": (i32, [vp, vp, sz, vp, vp]),
    "onk_op_elapsed_ms": (i32, [vp, P(C.c_float)]),
    "onk_launch": (i32, [vp, P(C.c_uint), P(C.c_uint), sz, i32, P(vp)]),
    "onk_launch_ex": (i32, [vp, P(C.c_uint), P(C.c_uint), sz, i32, P(vp), i32]),
    "onk_launch_count": (i32, [P(u64)]),
    "onk_graph_begin": (i32, [i32, P(i32), i32]),
    "onk_graph_end": (i32, [P(vp)]),
    "onk_graph_launch": (i32, [vp, i32]),
    "onk_graph_destroy": (i32, [vp]),
    "onk_reduce_f64": (i32, [i32, vp, u64, vp, i32]),
    "onk_combine_f64": (i32, [i32, vp, u32, vp, i32]),
    "onk_xchg_create": (i32, [i32, i32, P(vp)]),
    "onk_xchg_create_ex": (i32, [i32, i32, i32, P(vp)]),
    "onk_xchg_export": (i32, [vp, C.c_char_p]),
    "onk_xchg_connect": (i32, [vp, C.c_char_p]),
    "onk_xchg_destroy": (i32, [vp]),
    "onk_reduce_xchg_f64": (i32, [i32, vp, u64, vp, vp, i32]),
    "onk_xchg_allreduce_f64": (i32, [vp, i32, vp, i32, i32]),
    "onk_xchg_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32)]),
    "onk_xchg_barrier": (i32, [vp, i32]),
    "onk_xchg_status": (i32, [vp, P(u64)]),
    "onk_window_create": (i32, [C.c_size_t, i32, i32, C.POINTER(vp)]),
    "onk_window_export": (i32, [vp, C.c_char_p]),
    "onk_window_connect": (i32, [vp, C.c_char_p]),
    "onk_window_ptr": (i32, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
    "onk_window_destroy": (i32, [vp]),
    "onk_xs_plan_create": (i32, [vp, u64, u64, C.POINTER(vp)]),
    "onk_xs_plan_export": (i32, [vp, C.c_char_p]),
    "onk_xs_plan_connect": (i32, [vp, C.c_char_p]),
    "onk_xs_plan_build": (i32, [vp, vp, u64, vp, u64, i32, C.POINTER(i32)]),
    "onk_xs_plan_send_counts": (i32, [vp, C.POINTER(u64)]),
    "onk_xs_move": (i32, [vp, vp, u64, vp, i32]),
    "onk_xs_move_fields": (i32, [vp, i32, C.POINTER(vp), C.POINTER(u64), C.POINTER(vp), i32]),
    "onk_xs_plan_destroy": (i32, [vp]),
    "onk_axpy_f64": (i32, [vp, vp, dbl, u64, i32]),
    "onk_parallel_memset": (i32, [vp, u64, u64, i32, i32]),
    "onk_value_add_f64": (i32, [vp, u64, u64, dbl, i32]),
    "onk_soatl_update_capacity": (u64, [u64, u64, u64]),
    "onk_soatl_layout": (u64, [u64, P(u32), i32, u64, P(u64)]),
    "onk_soatl_rmw_f64": (i32, [vp, u64, u64, i32, u64, u64, dbl, P(dbl), i32]),
    "onk_soatl_dist_f64": (i32, [vp, vp, vp, vp, u64, dbl, dbl, dbl, i32]),
    "onk_soatl_repack": (i32, [vp, u64, u64, vp, u64, u64, P(u32), i32, u64, i32]),
    "onk_cell_grid_layout": (u64, [P(u32), u64, u64, u64, P(u32), i32, u64, P(u32), P(u64)]),
    "onk_cell_sum_f64": (i32, [vp, vp, vp, vp, u64, u64, P(u32), i32, i32, vp, vp, i32]),
    "onk_cell_sum_packed_f64": (i32, [vp, u64, vp, u64, vp, vp, i32]),
    "onk_cell_sum_fields_f64": (i32, [vp, u64, vp, vp, vp, u64, u64, P(u32), i32, vp, vp, i32]),
    "onk_host_reduce_f64": (i32, [i32, vp, u64, vp]),
    "onk_host_axpy_f64": (i32, [vp, vp, dbl, u64]),
    "onk_host_value_add_f64": (i32, [vp, u64, u64, dbl]),
}


class OnikaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"onika_b200 error {code}: {message}")
        self.code = code


def header_symbols() -> list[str]:
    """Every function the C header declares."""
    with open(HEADER) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(onk_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != OK:
        raise OnikaError(code, lib().onk_last_error().decode(errors="replace"))
