"""Cell grid (config C4): the per-cell variable-length SOATL blocks exaNBody operators run block_parallel_for over,
packed back to back in one HBM arena with offset/size/capacity tables."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from ._capi import check, lib, LANE_DEFAULT
from .runtime import ptr


def cell_grid_layout(cell_size: np.ndarray, align: int, chunk: int, field_bytes: Sequence[int], cell_align: int):
    cs = np.ascontiguousarray(cell_size, dtype=np.uint32)
    fb = (C.c_uint32 * len(field_bytes))(*field_bytes)
    cap = np.zeros(cs.size, np.uint32)
    off = np.zeros(cs.size, np.uint64)
    total = lib().onk_cell_grid_layout(cs.ctypes.data_as(C.POINTER(C.c_uint32)), cs.size, align, chunk, fb, len(field_bytes), cell_align,
                                       cap.ctypes.data_as(C.POINTER(C.c_uint32)), off.ctypes.data_as(C.POINTER(C.c_uint64)))
    return int(total), cap, off


class CellGrid:
    """cells of FieldArrays<align,chunk,fields...> (all fields given by byte size; fp64 = 8)."""

    def __init__(self, cell_size: np.ndarray, align: int = 64, chunk: int = 16, field_bytes: Sequence[int] = (8, 8, 8, 8),
                 cell_align: int = 64, device="cuda"):
        import torch
        self.align, self.chunk, self.field_bytes, self.cell_align = align, chunk, list(field_bytes), cell_align
        self.cell_size_host = np.ascontiguousarray(cell_size, dtype=np.uint32)
        self.ncells = self.cell_size_host.size
        self.arena_bytes, self.cell_cap_host, self.cell_off_host = cell_grid_layout(self.cell_size_host, align, chunk, field_bytes, cell_align)
        self.device = device
        self.arena = torch.zeros(max(self.arena_bytes, 1), dtype=torch.uint8, device=device)
        self.cell_size = torch.from_numpy(self.cell_size_host.view(np.int32)).to(device)
        self.cell_cap = torch.from_numpy(self.cell_cap_host.view(np.int32)).to(device)
        self.cell_off = torch.from_numpy(self.cell_off_host.view(np.int64)).to(device)

    def field_byte_offsets(self, field: int) -> np.ndarray:
        """host: byte address (within the arena) of element 0 of `field` for every cell"""
        a = self.align
        cap = self.cell_cap_host.astype(np.uint64)
        off = self.cell_off_host.copy()
        for k in range(field):
            off += (cap * np.uint64(self.field_bytes[k]) + np.uint64(a - 1)) // np.uint64(a) * np.uint64(a)
        return off

    def host_arena_with_field(self, field: int, values_concat: np.ndarray) -> np.ndarray:
        """host-side construction of the arena bytes with one fp64 field filled cell after cell (test/bench input)"""
        arena = np.zeros(max(self.arena_bytes, 1), np.uint8)
        a64 = arena[: self.arena_bytes - self.arena_bytes % 8].view(np.float64) if self.arena_bytes >= 8 else np.zeros(0, np.float64)
        starts = self.field_byte_offsets(field) // np.uint64(8)
        sizes = self.cell_size_host.astype(np.int64)
        # element index of every particle: start[cell] + local index
        cell_of = np.repeat(np.arange(self.ncells, dtype=np.int64), sizes)
        first = np.concatenate([[0], np.cumsum(sizes)[:-1]]) if self.ncells else np.zeros(0, np.int64)
        local = np.arange(int(sizes.sum()), dtype=np.int64) - np.repeat(first, sizes)
        a64[starts.astype(np.int64)[cell_of] + local] = values_concat
        return arena

    def upload(self, arena_host: np.ndarray) -> None:
        import torch
        self.arena.copy_(torch.from_numpy(arena_host))

    def cell_sum(self, field: int, with_total: bool = True, lane: int = LANE_DEFAULT):
        """block_parallel_for(ncells, per-cell sum of `field`); returns (cell_sum tensor, total tensor or None)"""
        import torch
        sums = torch.empty(max(self.ncells, 1), dtype=torch.float64, device=self.device)[: self.ncells]
        total = torch.zeros(1, dtype=torch.float64, device=self.device) if with_total else None
        fb = (C.c_uint32 * len(self.field_bytes))(*self.field_bytes)
        check(lib().onk_cell_sum_f64(ptr(self.arena), ptr(self.cell_off), ptr(self.cell_size), ptr(self.cell_cap), self.ncells, self.align,
                                     fb, len(self.field_bytes), field, ptr(sums), ptr(total) if with_total else None, lane))
        return sums, total


    def cell_sum_all_fields(self, with_totals: bool = True, lane: int = LANE_DEFAULT):
        """per-cell sums of every (fp64) field in one pass over the arena (onk_cell_sum_fields_f64); returns
        (ncells x nfields tensor, nfields totals or None)"""
        import torch
        nf = len(self.field_bytes)
        sums = torch.empty((max(self.ncells, 1), nf), dtype=torch.float64, device=self.device)[: self.ncells]
        totals = torch.zeros(nf, dtype=torch.float64, device=self.device) if with_totals else None
        fb = (C.c_uint32 * nf)(*self.field_bytes)
        check(lib().onk_cell_sum_fields_f64(ptr(self.arena), self.arena_bytes, ptr(self.cell_off), ptr(self.cell_size), ptr(self.cell_cap), self.ncells,
                                            self.align, fb, nf, ptr(sums), ptr(totals) if with_totals else None, lane))
        return sums, totals


class FieldMajorCellGrid:
    """Cell grid stored FIELD-MAJOR: every field is one contiguous column over all cells, cell c owning
    [cell_begin[c], cell_begin[c+1]) in each column.  The B200-first layout of config C4: per-cell work streams the
    columns it needs at full HBM rate instead of picking 128-byte lines out of 512-byte per-cell blocks.
    Per-cell results are bit-identical to the per-cell FieldArrays layout (same particle order)."""

    def __init__(self, cell_size: np.ndarray, nfields: int = 4, device="cuda"):
        import torch
        self.cell_size_host = np.ascontiguousarray(cell_size, dtype=np.uint32)
        self.ncells = self.cell_size_host.size
        begin = np.zeros(self.ncells + 1, dtype=np.int64)
        np.cumsum(self.cell_size_host, dtype=np.int64, out=begin[1:])
        self.cell_begin_host = begin
        self.nparticles = int(begin[-1])
        self.device = device
        self.cell_begin = torch.from_numpy(begin).to(device)
        self.columns = [torch.zeros(max(self.nparticles, 1), dtype=torch.float64, device=device)[: self.nparticles] for _ in range(nfields)]

    def set_field(self, field: int, values_concat) -> None:
        import torch
        self.columns[field].copy_(torch.as_tensor(values_concat, dtype=torch.float64))

    def cell_sum(self, field: int, with_total: bool = True, lane: int = LANE_DEFAULT):
        """per-cell sum of `field` (onk_cell_sum_packed_f64); returns (cell_sum tensor, total tensor or None)"""
        import torch
        sums = torch.empty(max(self.ncells, 1), dtype=torch.float64, device=self.device)[: self.ncells]
        total = torch.zeros(1, dtype=torch.float64, device=self.device) if with_total else None
        check(lib().onk_cell_sum_packed_f64(ptr(self.columns[field]), self.nparticles, ptr(self.cell_begin), self.ncells, ptr(sums),
                                            ptr(total) if with_total else None, lane))
        return sums, total
