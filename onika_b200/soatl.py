"""Host-side mirror of onika::soatl (include/onika/soatl in the reference): packed struct-of-arrays storage in HBM.

`FieldArrays(align, chunk, fields)` reproduces FieldArraysWithAllocator<A,C,...> (field_arrays.h:78-585): ONE packed
allocation, column k at sum_{j<k} roundup(capacity*sizeof(T_j), A), capacity a multiple of C that follows the
reference's log growth policy, so a raw dump of `storage` is byte-compatible with the reference's
storage_ptr()/storage_size().  The storage lives in device memory (a torch uint8 tensor = plumbing); all arithmetic
on it runs in libonika_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import check, lib, LANE_DEFAULT
from .runtime import ptr

DEFAULT_ALIGNMENT = 64      # SimdRequirements<double>::alignment with AVX-512 (memory/simd.h:42-54)
DEFAULT_CHUNK_SIZE = 16     # SimdRequirements<float>::chunksize
MINIMUM_CUDA_ALIGNMENT = 256  # memory/allocator.h:61-65

_NP = {"f8": np.float64, "f4": np.float32, "u1": np.uint8, "i1": np.int8, "i2": np.int16, "i4": np.int32, "i8": np.int64, "u4": np.uint32}


def update_capacity(size: int, capacity: int, chunk: int) -> int:
    return int(lib().onk_soatl_update_capacity(size, capacity, chunk))


def layout(align: int, field_bytes: Sequence[int], capacity: int) -> Tuple[int, List[int]]:
    fb = (C.c_uint32 * len(field_bytes))(*field_bytes)
    offs = (C.c_uint64 * max(1, len(field_bytes)))()
    total = lib().onk_soatl_layout(align, fb, len(field_bytes), capacity, offs)
    return int(total), [int(offs[k]) for k in range(len(field_bytes))]


class FieldArrays:
    """fields: sequence of (name, dtype-code) e.g. [("rx","f8"),("ry","f8"),("atype","u1")]"""

    def __init__(self, align: int = DEFAULT_ALIGNMENT, chunk: int = DEFAULT_CHUNK_SIZE, fields: Sequence[Tuple[str, str]] = (), device="cuda"):
        assert align > 0 and align & (align - 1) == 0, "alignment must be a power of 2"
        assert chunk > 0
        self.align, self.chunk = align, chunk
        self.names = [f[0] for f in fields]
        self.codes = [f[1] for f in fields]
        self.field_bytes = [np.dtype(_NP[c]).itemsize for c in self.codes]
        self.device = device
        self._size = 0
        self._capacity = 0
        self.storage = None      # torch.uint8 tensor of storage_size() bytes (None when empty, like the reference's nullptr)

    # --- queries (field_arrays.h:253-266) ---
    def size(self) -> int:
        return self._size

    def capacity(self) -> int:
        return self._capacity

    def empty(self) -> bool:
        return self._size == 0

    def chunk_ceil(self) -> int:
        return (self._size + self.chunk - 1) // self.chunk * self.chunk

    def storage_size(self) -> int:
        return layout(self.align, self.field_bytes, self._capacity)[0]

    def offsets(self) -> List[int]:
        return layout(self.align, self.field_bytes, self._capacity)[1]

    def payload_bytes(self) -> int:
        return sum(self.field_bytes) * self._size

    # --- resize / clear (field_arrays.h:200-211, 513-580) ---
    def resize(self, n: int) -> None:
        import torch
        if n == self._size:
            return
        new_cap = update_capacity(n, self._capacity, self.chunk)
        if new_cap != self._capacity:
            old_storage, old_cap, keep = self.storage, self._capacity, min(n, self._size)
            if new_cap == 0:
                self.storage = None
            else:
                total, _ = layout(self.align, self.field_bytes, new_cap)
                self.storage = torch.empty(total, dtype=torch.uint8, device=self.device)
                assert self.storage.data_ptr() % max(self.align, 1) == 0
                if old_storage is not None and keep > 0:
                    fb = (C.c_uint32 * len(self.field_bytes))(*self.field_bytes)
                    check(lib().onk_soatl_repack(ptr(self.storage), self.align, new_cap, ptr(old_storage), self.align, old_cap,
                                                 fb, len(self.field_bytes), keep, LANE_DEFAULT))
                    check(lib().onk_lane_sync(LANE_DEFAULT))
            self._capacity = new_cap
        self._size = n

    def clear(self) -> None:
        self.resize(0)

    # --- column access (operator[](FieldId), field_arrays.h:296-306) ---
    def column(self, name: str):
        """typed torch view over the whole column (capacity elements)"""
        import torch
        k = self.names.index(name)
        off = self.offsets()[k]
        nbytes = self._capacity * self.field_bytes[k]
        tdt = {"f8": torch.float64, "f4": torch.float32, "u1": torch.uint8, "i1": torch.int8, "i2": torch.int16,
               "i4": torch.int32, "i8": torch.int64, "u4": torch.int32}[self.codes[k]]
        return self.storage[off:off + nbytes].view(tdt)

    def __getitem__(self, name: str):
        return self.column(name)

    def set_column(self, name: str, host_values: np.ndarray) -> None:
        import torch
        col = self.column(name)
        col[: len(host_values)].copy_(torch.from_numpy(np.ascontiguousarray(host_values)).view(col.dtype))

    # --- compute (soatl/compute.h) ---
    def parallel_apply_rmw(self, s: float, c: Sequence[float], lane: int = LANE_DEFAULT) -> None:
        """soatl::parallel_apply_simd(f, arrays, all fp64 fields) with f: x = x*s + c_k (config C3)."""
        assert all(code == "f8" for code in self.codes), "rmw sweep is defined on fp64 field sets"
        if self._capacity == 0:
            return
        cc = (C.c_double * len(c))(*c)
        check(lib().onk_soatl_rmw_f64(ptr(self.storage), self.align, self.chunk, len(self.codes), self._size, self._capacity,
                                      s, cc, lane))

    def parallel_apply_dist(self, e: str, rx: str, ry: str, rz: str, ax: float, ay: float, az: float, lane: int = LANE_DEFAULT) -> None:
        """the stock tests/benchmark.cpp kernel through parallel_apply_simd (3R + 1W, sweeps whole chunks)."""
        if self._capacity == 0:
            return
        check(lib().onk_soatl_dist_f64(ptr(self.column(e)), ptr(self.column(rx)), ptr(self.column(ry)), ptr(self.column(rz)),
                                       self.chunk_ceil(), ax, ay, az, lane))

    def copy_from(self, other: "FieldArrays", lane: int = LANE_DEFAULT) -> None:
        """soatl::copy(dst, src) for identical field lists in (possibly) different layouts (copy.h:29-97)."""
        assert self.codes == other.codes
        self.resize(other.size())
        if other.size() == 0:
            return
        fb = (C.c_uint32 * len(self.field_bytes))(*self.field_bytes)
        check(lib().onk_soatl_repack(ptr(self.storage), self.align, self._capacity, ptr(other.storage), other.align, other._capacity,
                                     fb, len(self.field_bytes), other.size(), lane))


def make_hybrid_field_arrays(align: int = DEFAULT_ALIGNMENT, chunk: int = DEFAULT_CHUNK_SIZE, fields=(), device="cuda") -> FieldArrays:
    """make_hybrid_field_arrays(cst::align<A>, cst::chunk<C>, ids...)  (field_arrays.h:587-601)"""
    return FieldArrays(align, chunk, fields, device)
