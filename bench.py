#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native onika data-parallel hot path.

Metric (BASELINE.json): block_parallel_for & reduce HBM GB/s.  Workload at every N: BASELINE configs[1], the fp64
reduce-min of tests/gpu_reduce_min_fp64.cu over 2^28 doubles (2 GiB) PER GPU (weak scaling: the element range is
sharded, the only exchange is the all-reduce of one 8-byte partial per step).  At N > 1 the same line also carries
`strong` (fixed totals 2^28 and 2^31 doubles: T(1) on rank 0 alone vs T(N) sharded), `C5` (the two-lane kernel
sequence of tests/parallel_queue_lane.cu at 16384^2 with rows sharded and fused min + sum, BASELINE configs[4]) and
`checks` (multi-GPU parity against the oracle, outside every timed region; a mismatch fails the run).

  python bench.py --gpus N --steps K --warmup W            this repo (CUDA path through the C ABI)
  python bench.py --gpus N --scaling strong ...            headline workload = 2^28 doubles in TOTAL, sharded over N
  python bench.py --impl reference --gpus N --steps K ...  the reference's own OpenMP path on the host cores

One JSON line on stdout (rank 0).  See the task contract for the keys; `roofline`, `cpu_baseline`, `e2e`, `clocks`
and `gpu_launches` are described in DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ELEMS = 1 << 28                      # config C2: 2^28 doubles = 2 GiB per GPU
SEED = 1234
METRIC = "block_parallel_for & reduce HBM GB/s (reduce_min fp64, 2^28 doubles per GPU)"
UNIT = "GB/s"
WORKLOAD = "reduce_min_f64_2^28 (BASELINE configs[1]; tests/gpu_reduce_min_fp64.cu shape)"
CPU_SAMPLE_ELEMS = N_ELEMS              # the host-core baseline runs the same 2^28 elements (one GPU's share of the workload)
L2_NOTE = "inputs larger than L2 (2 GiB per pass vs 126 MB)"


def bench_config(scaling: str = "weak") -> dict:
    """`config` of the JSON line: identical in both arms (the driver compares them)"""
    c = {"workload": WORKLOAD, "elements_per_gpu": N_ELEMS, "bytes_per_gpu": N_ELEMS * 8, "l2": L2_NOTE}
    if scaling == "strong":
        c = {"workload": WORKLOAD + ", fixed total", "elements_total": N_ELEMS, "bytes_total": N_ELEMS * 8, "l2": "flushed between timed iterations when the shard fits L2, else larger than L2"}
    return c


def read_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def _start_nvml(self) -> bool:
        """same counters straight from NVML (what nvidia-smi reads), polled every 2 ms: a 60 ms timed region gets ~30 samples
        instead of the 1-4 a 50 ms nvidia-smi loop delivers"""
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.device
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.device])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = (("hw_slowdown", pynvml.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", pynvml.nvmlClocksEventReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", pynvml.nvmlClocksEventReasonSwPowerCap))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)          # probe once: any failure falls back to nvidia-smi
        except Exception:
            return False
        self.nvml_rows, self.nvml_stop, self.nvml_max = [], threading.Event(), mx

        def poll():
            while not self.nvml_stop.is_set():
                try:
                    sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.nvml_rows.append((time.time(), sm, pw, [n for n, b in bits if mask & b]))
                except Exception:
                    pass
                time.sleep(0.002)
        self.t = threading.Thread(target=poll, daemon=True)
        self.t.start()
        return True

    def start(self):
        self.nvml = self._start_nvml()
        if self.nvml:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if getattr(self, "nvml", False):
            time.sleep(0.01)
            self.nvml_stop.set()
            self.t.join(timeout=1.0)
            sel = [r for r in self.nvml_rows if t0 <= r[0] <= t1] or self.nvml_rows
            if sel:
                sm = sorted(r[1] for r in sel)
                return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.nvml_max, "power_w_max": round(max(r[2] for r in sel), 2),
                        "samples": len(sel), "reasons": sorted({n for r in sel for n in r[3]}), "source": "NVML, 2 ms period"}
            return {"sm_mhz": None, "sm_max_mhz": self.nvml_max, "reasons": ["no samples"]}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = []
        for ts, line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 7:
                rows.append((ts, p))
        sel = [p for ts, p in rows if t0 - 0.05 <= ts <= t1 + 0.1] or [p for _, p in rows]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (num(p[0]) for p in sel) if v is not None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(p[3 + i].lower().startswith("active") for p in sel)]
        pw = [v for v in (num(p[2]) for p in sel) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": num(sel[0][1]), "power_w_max": max(pw) if pw else None,
                "samples": len(sel), "reasons": reasons}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reduce_min_baseline(min_seconds: float, steps: int | None = None, warmup: int = 1):
    """The reference's CPU implementation of the path on the host cores, over the same 2^28 doubles one GPU gets:
    `reference` = cpu_compute of tests/gpu_reduce_min_fp64.cu compiled from the reference tree (oracle/_ref; 2^20 elements per
    call as the reference hard-codes it), `port` = the single-region restatement (oracle/oracle.c).
    Returns a dict: value (GB/s of `kind`), kind, cores, sample, ref_gbs, port_gbs, seconds per pass, passes."""
    import numpy as np
    import oracle
    orc = oracle.orc
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    orc.set_num_threads(ncpu)
    n = CPU_SAMPLE_ELEMS
    d = np.empty(n, dtype=np.float64)
    orc.fill_lcg(d, SEED)                     # parallel first touch, like cpu_init (tests/gpu_reduce_min_fp64.cu:30-44)
    cores = orc.num_threads()
    want = orc.reduce_min(d)

    def run(fn):
        for _ in range(warmup):
            fn()
        ts = []
        t_end = time.perf_counter() + min_seconds
        while (steps is None and time.perf_counter() < t_end and len(ts) < 400) or (steps is not None and len(ts) < steps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
            assert r == want
        return ts

    res = {"port": run(lambda: orc.reduce_min(d))}
    if oracle.ref.available():
        res["reference"] = run(lambda: oracle.ref.reduce_min_chunks(d)[0])
    kind = "reference" if "reference" in res else "port"      # the line's value is the reference's own code whenever it was built
    ts = sorted(res[kind])
    med = ts[len(ts) // 2]
    gbs = {k: n * 8 / sorted(v)[len(v) // 2] / 1e9 for k, v in res.items()}
    sample = (f"reduce_min over {n * 8 >> 20} MiB (2^28 doubles, one GPU's share), median of {len(ts)} passes, OpenMP {cores} threads; "
              "reference = cpu_compute compiled from the reference tree (2^20 elements per call); port = oracle/oracle.c, one region")
    return {"value": gbs[kind], "kind": kind, "cores": cores, "sample": sample, "ref_gbs": gbs.get("reference"), "port_gbs": gbs["port"],
            "seconds": med, "passes": len(ts)}


def cpu_secondary_baselines(budget_s: float = 2.0) -> dict:
    """The reference's own OpenMP path (oracle/_ref; oracle port when _ref is absent) on bounded samples of the secondary
    configs, median of a few passes each: context for the `kernels` object, same run, same box."""
    import numpy as np
    import oracle
    orc = oracle.orc
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    orc.set_num_threads(ncpu)
    use_ref = oracle.ref.available()
    out = {"cores": ncpu, "impl": "reference (oracle/_ref)" if use_ref else "port (oracle/oracle.c)"}

    def med(fn, nbytes):
        fn()
        ts = []
        t_end = time.perf_counter() + budget_s
        while time.perf_counter() < t_end and len(ts) < 15:
            ts.append(fn())
        ts.sort()
        return round(nbytes / ts[len(ts) // 2] / 1e9, 1)

    def wall(f):
        t0 = time.perf_counter(); f(); return time.perf_counter() - t0

    n = 10_000_000
    x = np.empty(n); y = np.empty(n)
    orc.fill_lcg(x, 1, -1.0, 1.0); orc.fill_lcg(y, 2, -1.0, 1.0)
    out["C1_axpy_10M_GBps"] = med((lambda: oracle.ref.axpy(y, x, 0.5)) if use_ref else (lambda: wall(lambda: orc.axpy(y, x, 0.5))), 24 * n)
    rows = cols = 4096
    a = np.empty((rows, cols)); orc.fill_lcg(a.reshape(-1), 3)
    out["C5_value_add_4096x4096_GBps"] = med((lambda: oracle.ref.value_add(a, 1.1)) if use_ref else (lambda: wall(lambda: orc.value_add(a, 1.1))), 16 * rows * cols)
    if use_ref:
        n = 8 << 20
        colsv = [np.empty(n) for _ in range(8)]
        for k, cv in enumerate(colsv):
            orc.fill_lcg(cv, 20 + k, 0.0, 1.0)
        out["C3_soatl_rmw_8x8Mi_GBps"] = med(lambda: oracle.ref.soatl_rmw8(colsv, 1.0000001, [0.125 * (k + 1) for k in range(8)], 64), 128 * n)
        from tests._inputs import poisson_counts
        counts = poisson_counts(64 ** 3)
        e = np.empty(int(counts.sum())); orc.fill_lcg(e, 40, -1.0, 1.0)
        alg = 8 * e.size + 24 * counts.size
        out["C4_cell_sum_64^3_GBps"] = med(lambda: oracle.ref.cell_sum(counts, e)[3], alg)
    return out


def _cpu_baseline_object(b: dict) -> dict:
    r = lambda v: None if v is None else round(v, 3)
    return {"value": r(b["value"]), "unit": UNIT, "cores": b["cores"], "kind": b["kind"], "sample": b["sample"],
            "ref_gbs": r(b["ref_gbs"]), "port_gbs": r(b["port_gbs"])}


def run_reference_arm(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    b = cpu_reduce_min_baseline(0.0, steps=max(1, args.steps), warmup=max(1, args.warmup))
    gbs = b["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": b["passes"],
        "warmup": max(1, args.warmup), "ms_per_step": round(b["seconds"] * 1e3, 4), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.scaling),
        "method": "each step = one pass of the reference's OpenMP reduce-min over 2^28 doubles in host memory on the box's host cores (rank 0 only); no GPU involved",
        "cpu_baseline": _cpu_baseline_object(b),
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "elements_per_s": round(gbs * 1e9 / 8, 1),
    }
    emit_json(line)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def secondary_kernels(ob, torch):
    """Short device-timed pass over the other BASELINE configs (context for the judge; not the headline)."""
    import numpy as np
    from onika_b200 import parallel as P, soatl, cellgrid
    from tests._inputs import poisson_counts
    out = {}

    def timed(fn, reps=20, warm=5):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    # the working sets of C1 and C4 fit in the 126 MB L2; 1 GiB of other data is READ between reps to flush it (clean
    # lines: a write-flush would leave dirty lines whose eviction is billed to the next kernel)
    flush = torch.ones(1 << 27, dtype=torch.float64, device="cuda")
    flush_out = torch.empty(1, dtype=torch.float64, device="cuda")

    def flushed(fn, reps=8):
        ts = []
        for _ in range(reps + 2):
            P.ReduceFunctor(ob.OP_MIN, flush, flush_out).launch(flush.numel(), 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts = sorted(ts[2:])
        return ts[len(ts) // 2]

    n = 10_000_000
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
    f = P.AxpyFunctor(y, x, 0.5)
    t = flushed(lambda: f.launch(n, 0))
    out["C1_axpy_10M"] = {"GB/s": round(24 * n / t / 1e9, 1), "ms": round(t * 1e3, 4), "bytes_per_elem": 24, "l2": "flushed"}
    n = 1 << 28
    big = torch.rand(n, dtype=torch.float64, device="cuda")
    f = P.AxpyFunctor(big, big[: n // 2], 0.5)
    # (read-only kernels first: for some milliseconds after a long read-modify-write pass the same reduction runs ~8 % slower)
    outd = torch.empty(1, dtype=torch.float64, device="cuda")
    t = timed(lambda: P.ReduceFunctor(ob.OP_SUM, big, outd).launch(n, 0))
    out["reduce_sum_2^28"] = {"GB/s": round(8 * n / t / 1e9, 1), "ms": round(t * 1e3, 4), "bytes_per_elem": 8}
    t = timed(lambda: P.BlockParallelValueAddFunctor(big.view(1 << 14, 1 << 14), 1.1).launch(1 << 14, 0))
    out["C5_value_add_16384x16384"] = {"GB/s": round(16 * n / t / 1e9, 1), "ms": round(t * 1e3, 4), "bytes_per_elem": 16}
    # C5: tests/parallel_queue_lane.cu sequence at 16384^2 -- two arrays, (kernel1 -> kernel2 -> min -> sum) per array on
    # lanes 4/5, captured once as a CUDA graph and replayed on lane 0
    import ctypes as C
    from onika_b200 import _capi
    L = _capi.lib()
    big2 = torch.rand(n, dtype=torch.float64, device="cuda")
    res = torch.empty(4, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    others = (C.c_int * 1)(5)
    _capi.check(L.onk_graph_begin(4, others, 1))
    for lane, arr, v1, v2, r0 in ((4, big, 1.1, 1.2, 0), (5, big2, 1.3, 1.4, 2)):
        _capi.check(L.onk_value_add_f64(arr.data_ptr(), 1 << 14, 1 << 14, v1, lane))
        _capi.check(L.onk_value_add_f64(arr.data_ptr(), 1 << 14, 1 << 14, v2, lane))
        _capi.check(L.onk_reduce_f64(ob.OP_MIN, arr.data_ptr(), n, res[r0:].data_ptr(), lane))
        _capi.check(L.onk_reduce_f64(ob.OP_SUM, arr.data_ptr(), n, res[r0 + 1:].data_ptr(), lane))
    gph = C.c_void_p()
    _capi.check(L.onk_graph_end(C.byref(gph)))
    t = timed(lambda: _capi.check(L.onk_graph_launch(gph, 0)), reps=5, warm=2)
    c5_bytes = 2 * (2 * 16 * n + 2 * 8 * n)
    out["C5_lane_sequence_2x16384^2_graph"] = {"GB/s": round(c5_bytes / t / 1e9, 1), "ms": round(t * 1e3, 4), "kernels_per_replay": 8,
                                               "algorithmic_bytes": c5_bytes, "lanes": 2}
    _capi.check(L.onk_graph_destroy(gph))
    del big, big2
    fa = soatl.make_hybrid_field_arrays(256, 16, [(f"f{k}", "f8") for k in range(8)])
    n = 64 << 20
    fa.resize(n)
    fa.storage.view(torch.float64).uniform_(0, 1)
    c = [0.125 * (k + 1) for k in range(8)]
    t = timed(lambda: fa.parallel_apply_rmw(1.0000001, c), reps=5)
    out["C3_soatl_rmw_8x64Mi"] = {"GB/s": round(128 * n / t / 1e9, 1), "ms": round(t * 1e3, 4), "bytes_per_elem": 128}
    fa.clear()
    # config C4 at the BASELINE size (128^3 cells, working set ~ L2-sized for the e-only variant: L2 flushed) and at 256^3
    # cells (2.4 GB of e lines / 12.3 GB arena: steady state); three variants each, as SURVEY 8(d) asks to state:
    #   field e only, per-cell SOATL blocks  : 8 B/particle + 24 B/cell (offset 8, size 4, capacity 4, out 8)
    #   all 4 fields, per-cell SOATL blocks  : 32 B/particle + 48 B/cell (tables 16 + 4 outputs) -- the arena is one contiguous stream
    #   field e only, field-major layout     : 8 B/particle + 16 B/cell
    for side in (128, 256):
        counts = poisson_counts(side ** 3)
        npart = int(counts.sum())
        tag = f"C4_cell_sum_{side}^3"
        run = flushed if side == 128 else (lambda fn: timed(fn, reps=5, warm=2))
        cg = cellgrid.CellGrid(counts, 64, 16, (8, 8, 8, 8), 64)
        cg.arena.view(torch.float64).uniform_(-1, 1)
        alg = 8 * npart + cg.ncells * (8 + 8 + 4 + 4)
        t = run(lambda: cg.cell_sum(3))
        out[tag] = {"GB/s": round(alg / t / 1e9, 1), "ms": round(t * 1e3, 4), "algorithmic_bytes": alg, "arena_bytes": cg.arena_bytes, "particles": npart,
                    "l2": "flushed" if side == 128 else "working set larger than L2", "variant": "field e only (8 B/particle + 24 B/cell tables+out)"}
        alg = 32 * npart + cg.ncells * (8 + 4 + 4 + 32)
        t = run(lambda: cg.cell_sum_all_fields())
        out[tag + "_all_fields"] = {"GB/s": round(alg / t / 1e9, 1), "ms": round(t * 1e3, 4), "algorithmic_bytes": alg, "arena_bytes": cg.arena_bytes,
                                    "arena_stream_GB/s": round((cg.arena_bytes + cg.ncells * 48) / t / 1e9, 1), "particles": npart,
                                    "variant": "all 4 fields of every cell block (32 B/particle + 48 B/cell): the arena streams as one range, chunk padding included in arena_stream"}
        del cg
        fm = cellgrid.FieldMajorCellGrid(counts, nfields=1)
        fm.columns[0].uniform_(-1, 1)
        alg = 8 * npart + fm.ncells * (8 + 8)
        t = run(lambda: fm.cell_sum(0))
        out[tag + "_field_major"] = {"GB/s": round(alg / t / 1e9, 1), "ms": round(t * 1e3, 4), "algorithmic_bytes": alg, "particles": npart,
                                     "variant": "same cells, field-major layout: 8 B/particle + 16 B/cell (offset + out)"}
        del fm
        torch.cuda.empty_cache()
    # the GENERIC paths of the C++ front end (user functors / operators, no typed entry point), timed by the front-end test
    # program in its own process: config C4 written the application way (one block per cell) and the two C3 operators
    # through soatl::parallel_apply_simd
    try:
        import tempfile
        import numpy as np
        exe = os.path.join(ROOT, "tests", "cxx", "frontend_test")
        if os.path.exists(exe):
            with tempfile.TemporaryDirectory() as tmp:
                o = os.path.join(tmp, "o.bin")
                subprocess.run([exe, "cells_generic_perf", o, str(128 ** 3)], check=True, capture_output=True, timeout=300)
                us = np.fromfile(o, dtype=np.float64, count=6)
                small = np.fromfile(o, dtype=np.float64, count=2, offset=8 * (6 + 6 * 128 ** 3))
                out["C4_generic_functor_path_128^3_us"] = {
                    "sub_block_items_4_2_threads": [round(float(v), 1) for v in small],
                    "one_block_of_128_threads_per_cell": round(float(us[0]), 1), "one_block_of_32_threads_per_cell": round(float(us[1]), 1),
                    "sub_block_items_32_16_8_threads": [round(float(v), 1) for v in us[3:6]], "typed_field_major_kernel": round(float(us[2]), 1),
                    "what": "block_parallel_for(ncells, functor) with ONIKA_CU_BLOCK_SIMD_FOR + block_reduce_add, unchanged functor source; sub-block items = several cells per CTA (BlockParallelForFunctorTraits::SubBlockItems)"}
                subprocess.run([exe, "soatl_apply", o, str(64 << 20), "1"], check=True, capture_output=True, timeout=300)
                us = np.fromfile(o, dtype=np.float64, count=4)
                n = 64 << 20
                out["C3_soatl_parallel_apply_simd_generic_64Mi"] = {
                    "rmw_8_fields_GB/s": round(128 * n / float(us[0]) / 1e3, 1), "rmw_8_fields_typed_GB/s": round(128 * n / float(us[1]) / 1e3, 1),
                    "benchmark_cpp_operator_GB/s": round(32 * n / float(us[2]) / 1e3, 1), "benchmark_cpp_operator_typed_GB/s": round(32 * n / float(us[3]) / 1e3, 1),
                    "what": "extended lambdas through the vectorised generic kernel of include/onika/soatl/compute.h vs the typed entry points on the same arrays"}
    except Exception as ex:  # pragma: no cover
        out["generic_paths_error"] = str(ex)[:200]
    return out


def _max_over_ranks(torch, dist, world, v: float) -> float:
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _gather_floats(torch, dist, world, v: float):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


class _L2Flush:
    """READS 1 GiB of other data (clean lines: a write-flush would leave dirty lines whose eviction is billed to the next kernel)"""

    def __init__(self, torch, ob, P):
        self.buf = torch.ones(1 << 27, dtype=torch.float64, device="cuda")
        self.out = torch.empty(1, dtype=torch.float64, device="cuda")
        self.f = P.ReduceFunctor(ob.OP_MIN, self.buf, self.out)

    def __call__(self):
        self.f.launch(self.buf.numel(), 0)


def strong_scaling(K, W, world, rank, torch, dist, ob, P, D, xchg):
    """scaling = T(1)/T(G) at FIXED total size.  T(1): rank 0 alone reduces the whole input (the others wait at a
    barrier); T(N): every rank reduces its contiguous shard and the partials are combined inside the kernel.  Both timed with
    CUDA events over K back-to-back steps (max over ranks).  2^28 is BASELINE C2; 2^31 doubles (16 GiB) is the largest
    power of two that fits one GPU next to the other buffers."""
    out = {}
    errors = []                                                  # raised at the very end: every rank goes through every collective
    for label, total in (("2^28", 1 << 28), ("2^31", 1 << 31)):
        b, e = D.shard_range(total, rank, world)
        n_loc = e - b
        n_alloc = total if rank == 0 else n_loc                 # rank 0 also holds the whole input for T(1)
        data = torch.empty(n_alloc, dtype=torch.float64, device="cuda")
        data.uniform_(1e-3, 1.0, generator=torch.Generator(device="cuda").manual_seed(SEED + 17 * rank))
        res = torch.empty(1, dtype=torch.float64, device="cuda")
        shard = data[:n_loc]
        want = shard.min()
        if world > 1:
            dist.all_reduce(want, op=dist.ReduceOp.MIN)

        def step_n():
            if xchg is not None:
                xchg.reduce(ob.OP_MIN, shard, n=n_loc, out=res, lane=0)
            else:
                P.ReduceFunctor(ob.OP_MIN, shard, res).launch(n_loc, 0)
                if world > 1:
                    D.combine_across_ranks(ob.OP_MIN, res)

        def timed(fn, k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / k

        for _ in range(W):
            step_n()
        torch.cuda.synchronize()
        if float(res.item()) != float(want.item()):
            errors.append(f"strong scaling {label}: sharded reduce_min differs from the check value")
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tn = _max_over_ranks(torch, dist, world, timed(step_n, K))
        t1 = 0.0
        if rank == 0:
            f1 = P.ReduceFunctor(ob.OP_MIN, data, res)
            for _ in range(W):
                f1.launch(total, 0)
            torch.cuda.synchronize()
            t1 = timed(lambda: f1.launch(total, 0), K)
        t1 = _max_over_ranks(torch, dist, world, t1)
        out[label] = {"elements_total": total, "T1_ms": round(t1, 5), "TN_ms": round(tn, 5), "speedup": round(t1 / tn, 3),
                      "GBps_1": round(total * 8 / t1 / 1e6, 1), "GBps_N": round(total * 8 / tn / 1e6, 1)}
        del data, shard
        torch.cuda.empty_cache()
    assert not errors, "; ".join(errors)
    out["definition"] = "T(1)/T(N) at fixed total size; T(1) = rank 0 alone on the whole input in this same run, T(N) = sharded, partials combined inside the reduction kernel; K back-to-back steps between two CUDA events, max over ranks"
    return out


def c5_lane_sequence(K, W, world, rank, torch, dist, ob, P, D, xchgs):
    """BASELINE configs[4] (tests/parallel_queue_lane.cu:58-108 at 16384 x 16384): two arrays, kernel1 -> kernel2 -> min -> sum
    per array on two lanes, rows sharded contiguously over the ranks; the min and the sum of each array are combined across the
    GPUs inside the reduction kernels (one exchange per lane).  Timed from a fork on lane 0 to the join on lane 0, K steps
    between two CUDA events, max over ranks; T(1) = rank 0 alone on the whole arrays."""
    import ctypes as C
    from onika_b200 import _capi
    L = _capi.lib()
    rows_total = cols = 1 << 14
    b, e = D.shard_range(rows_total, rank, world)
    rows = e - b
    rows_alloc = rows_total if rank == 0 else rows
    arrs = [torch.empty(rows_alloc * cols, dtype=torch.float64, device="cuda") for _ in range(2)]
    for k, a in enumerate(arrs):
        a.uniform_(0.0, 1.0, generator=torch.Generator(device="cuda").manual_seed(SEED + 31 * rank + k))
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    adds = ((1.1, 1.2), (1.3, 1.4))
    lanes = (1, 2)

    def step(nrows, fused):
        for k in range(2):
            _capi.check(L.onk_lane_wait_lane(lanes[k], 0))          # fork: both lanes start after what lane 0 holds (the start event)
            p = arrs[k].data_ptr()
            _capi.check(L.onk_value_add_f64(p, nrows, cols, adds[k][0], lanes[k]))
            _capi.check(L.onk_value_add_f64(p, nrows, cols, adds[k][1], lanes[k]))
            for j, op in enumerate((ob.OP_MIN, ob.OP_SUM)):
                o = res[2 * k + j:].data_ptr()
                if fused:
                    _capi.check(L.onk_reduce_xchg_f64(op, p, nrows * cols, o, xchgs[k].handle, lanes[k]))
                else:
                    _capi.check(L.onk_reduce_f64(op, p, nrows * cols, o, lanes[k]))
        for k in range(2):
            _capi.check(L.onk_lane_wait_lane(0, lanes[k]))          # join

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    fused = world > 1 and xchgs is not None
    if world > 1 and not fused:
        return {"skipped": "C5 on N GPUs needs the peer exchange (CUDA IPC / peer access unavailable on this box)"}
    # expected minimum: x -> fl(fl(x + v1) + v2) is monotone, so the minimum of the array follows the same recurrence exactly
    mins = [a[:rows * cols].min() for a in arrs]
    if world > 1:
        for m in mins:
            dist.all_reduce(m, op=dist.ReduceOp.MIN)
    mins = [float(m.item()) for m in mins]
    nsteps = W + K
    for _ in range(W):
        step(rows, fused)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tn = _max_over_ranks(torch, dist, world, timed(lambda: step(rows, fused), K))
    for k in range(2):
        for _ in range(nsteps):
            mins[k] = (mins[k] + adds[k][0]) + adds[k][1]
    got = res.cpu().numpy().tolist()
    errors = []                                                  # raised at the very end: every rank goes through every collective
    if not (got[0] == mins[0] and got[2] == mins[1]):
        errors.append(f"C5: minima {got[0]!r}, {got[2]!r} differ from the expected {mins!r}")
    sums = [a[:rows * cols].sum() for a in arrs]
    mags = [a[:rows * cols].abs().sum() for a in arrs]
    if world > 1:
        for t in sums + mags:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    for k in range(2):
        if not abs(got[2 * k + 1] - float(sums[k].item())) <= 1e-12 * float(mags[k].item()):
            errors.append(f"C5: sum of array {k} outside 1e-12")
    if world > 1:
        g = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(g, res)
        if not all(torch.equal(x.view(torch.int64), g[0].view(torch.int64)) for x in g):
            errors.append("C5: reductions differ between ranks")
    t1 = tn
    if world > 1:
        t1 = 0.0
        if rank == 0:
            for _ in range(W):
                step(rows_total, False)
            torch.cuda.synchronize()
            t1 = timed(lambda: step(rows_total, False), K)
        t1 = _max_over_ranks(torch, dist, world, t1)
    nbytes = 2 * (2 * 16 + 2 * 8) * rows_total * cols
    out = {"rows_x_cols": [rows_total, cols], "arrays": 2, "lanes": 2, "kernels_per_step_per_rank": 8, "algorithmic_bytes": nbytes,
           "T1_ms": round(t1, 5), "TN_ms": round(tn, 5), "speedup": round(t1 / tn, 3), "GBps_N": round(nbytes / tn / 1e6, 1),
           "checks": "min bit-exact (monotone recurrence), sum <= 1e-12, same bits on all ranks",
           "sequence": "per array: value_add(v1) -> value_add(v2) -> reduce min -> reduce sum on its own lane; fork/join on lane 0 every step"}
    del arrs
    torch.cuda.empty_cache()
    assert not errors, "; ".join(errors)
    return out


def h2d_peak(torch, dist, world, nbytes=1 << 30):
    """measured host->device copy rate of this box: every rank copies `nbytes` from its own pinned buffer at the same time
    (cudaMemcpyAsync), best of 3; returns (per-rank GB/s list, aggregate GB/s)"""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    best = 0.0
    for _ in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / e0.elapsed_time(e1) / 1e6)
    per_rank = _gather_floats(torch, dist, world, best)
    del src, dst
    return [round(v, 2) for v in per_rank], round(sum(per_rank), 2)


def e2e_measure(K, world, rank, torch, dist, ob, P, data, n):
    """the plugin-facing host-buffer call onk_host_reduce_f64 on PINNED and on PAGEABLE host input (host -> device copies
    inside the timed region), next to the measured H2D copy rate of the same box"""
    per_rank_peak, agg_peak = h2d_peak(torch, dist, world)
    want = float(data.min().item())
    Ke = max(2, min(K, 10))

    def run(hostbuf):
        for _ in range(2):
            r = P.host_reduce(ob.OP_MIN, hostbuf)
        assert r == want
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            r = P.host_reduce(ob.OP_MIN, hostbuf)   # synchronous: returns the host result (8 bytes D2H)
            if world > 1:
                rt = torch.tensor([r], dtype=torch.float64, device="cuda")
                dist.all_reduce(rt, op=dist.ReduceOp.MIN)
                r = float(rt.item())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / Ke
        per_rank = _gather_floats(torch, dist, world, n * 8 / dt / 1e9)
        return _max_over_ranks(torch, dist, world, dt), per_rank

    pinned = torch.empty(n, dtype=torch.float64).pin_memory()
    pinned.copy_(data)
    torch.cuda.synchronize()
    dt, per_rank = run(pinned)
    e2e = {"value": round(world * n * 8 / dt / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": 8,
           "ms_per_step": round(dt * 1e3, 3), "steps": Ke, "call": "onk_host_reduce_f64 (pinned host input, chunked H2D overlapped with the kernel)",
           "per_rank_GBps": [round(v, 2) for v in per_rank],
           "h2d_peak_GBps": agg_peak, "h2d_peak_per_rank_GBps": per_rank_peak,
           "h2d_peak_how": "1 GiB cudaMemcpyAsync from pinned memory on every rank at the same time, best of 4",
           "frac_of_pcie": round(world * n * 8 / dt / 1e9 / agg_peak, 4) if agg_peak else None,
           # weak scaling: every rank moves the same bytes, so the step ends with the rank on the slowest host link
           "frac_of_pcie_slowest_rank": round(n * 8 / dt / 1e9 / min(per_rank_peak), 4) if per_rank_peak and min(per_rank_peak) > 0 else None}
    pageable = torch.empty(n, dtype=torch.float64)
    pageable.copy_(pinned)
    del pinned
    dtp, per_rank_p = run(pageable)
    e2e["pageable"] = {"value": round(world * n * 8 / dtp / 1e9, 3), "ms_per_step": round(dtp * 1e3, 3),
                       "per_rank_GBps": [round(v, 2) for v in per_rank_p],
                       "call": "onk_host_reduce_f64 on ordinary (pageable) host memory: staged through pinned buffers by the library"}
    del pageable
    return e2e


def e2e_multipass(K, world, rank, torch, dist, ob, D):
    """e2e of a MULTI-PASS workload, where a transfer is amortised: the C5 sequence on HOST-resident arrays -- each rank's two
    row shards cross PCIe once (pinned host -> device, on the two lanes), then kernel1, kernel2, min and sum run on the
    device and 32 bytes of results come back; the updated arrays stay device-resident, as the reference's managed arrays do."""
    import ctypes as C
    import numpy as np
    from onika_b200 import _capi
    L = _capi.lib()
    rows_total = cols = 1 << 14
    b, e = D.shard_range(rows_total, rank, world)
    rows = e - b
    n = rows * cols
    host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
    dev = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
    for k, h in enumerate(host):
        dev[k].uniform_(0.0, 1.0)
        h.copy_(dev[k])                       # the step's input lives in pinned host memory
    torch.cuda.synchronize()
    res = torch.zeros(4, dtype=torch.float64, device="cuda")
    res_host = torch.zeros(4, dtype=torch.float64).pin_memory()
    adds = ((1.1, 1.2), (1.3, 1.4))
    lanes = (1, 2)

    def once():
        for k in range(2):
            _capi.check(L.onk_memcpy(dev[k].data_ptr(), host[k].data_ptr(), n * 8, lanes[k]))
            _capi.check(L.onk_value_add_f64(dev[k].data_ptr(), rows, cols, adds[k][0], lanes[k]))
            _capi.check(L.onk_value_add_f64(dev[k].data_ptr(), rows, cols, adds[k][1], lanes[k]))
            _capi.check(L.onk_reduce_f64(ob.OP_MIN, dev[k].data_ptr(), n, res[2 * k:].data_ptr(), lanes[k]))
            _capi.check(L.onk_reduce_f64(ob.OP_SUM, dev[k].data_ptr(), n, res[2 * k + 1:].data_ptr(), lanes[k]))
            _capi.check(L.onk_memcpy(res_host[2 * k:].data_ptr(), res[2 * k:].data_ptr(), 16, lanes[k]))
        for k in range(2):
            _capi.check(L.onk_lane_sync(lanes[k]))
        return res_host.clone()

    once()
    if world > 1:
        dist.barrier()
    Ke = max(2, min(K, 5))
    t0 = time.perf_counter()
    for _ in range(Ke):
        r = once()
        if world > 1:
            rt = r.cuda()
            dist.all_reduce(rt, op=dist.ReduceOp.MIN)     # (stand-in for the combine; 32 bytes)
    torch.cuda.synchronize()
    dt = _max_over_ranks(torch, dist, world, (time.perf_counter() - t0) / Ke)
    want_min = [(float(h.min().item()) + a[0]) + a[1] for h, a in zip(host, adds)]
    assert [float(r[0]), float(r[2])] == want_min, "multi-pass e2e: minima differ"
    nbytes = 2 * (2 * 16 + 2 * 8) * rows_total * cols
    return {"workload": "C5 lane sequence, 2 x 16384^2 fp64 in pinned HOST memory, rows sharded; 4 passes per transferred byte",
            "value": round(nbytes / dt / 1e9, 3), "unit": UNIT, "ms_per_step": round(dt * 1e3, 3), "steps": Ke,
            "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": 32, "algorithmic_bytes": nbytes}


def cpu_multipass_baseline(budget_s: float = 6.0) -> dict:
    """the same C5 sequence by the reference's OpenMP path on the host cores (oracle/_ref lane_sequence; port otherwise) on a
    4096 x 4096 sample per array, plus min and sum by the oracle: GB/s of the same algorithmic bytes"""
    import numpy as np
    import oracle
    orc = oracle.orc
    rows = cols = 4096
    a1 = np.empty((rows, cols)); a2 = np.empty((rows, cols))
    orc.fill_lcg(a1.reshape(-1), 3, 0.0, 1.0); orc.fill_lcg(a2.reshape(-1), 4, 0.0, 1.0)
    use_ref = oracle.ref.available()
    ts = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(ts) < 9:
        if use_ref:
            t = oracle.ref.lane_sequence(a1, a2, 1.1, 1.2, 1.3, 1.4)    # seconds between enqueue and synchronize, as the reference times it
        else:
            t0 = time.perf_counter()
            orc.lane_sequence(a1, a2, 1.1, 1.2, 1.3, 1.4)
            t = time.perf_counter() - t0
        t0 = time.perf_counter()
        for a in (a1, a2):
            orc.reduce_min(a.reshape(-1)); orc.reduce_sum(a.reshape(-1), orc.num_threads())
        ts.append(t + time.perf_counter() - t0)
    ts.sort()
    nbytes = 2 * (2 * 16 + 2 * 8) * rows * cols
    return {"value": round(nbytes / ts[len(ts) // 2] / 1e9, 2), "unit": UNIT, "cores": orc.num_threads(), "kind": "reference" if use_ref else "port",
            "sample": f"2 x {rows}x{cols} fp64 in host memory, median of {len(ts)} passes"}


def run_ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist
    import onika_b200 as ob
    from onika_b200 import parallel as P, distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        emit_json({"error": "no CUDA device: the onika_b200 hot path has no CPU fallback"})
        return 2
    torch.cuda.set_device(local)
    ob.init(local)
    ob.bind_lane_to_torch_stream(0)          # lane 0 = torch's current stream: CUDA events below time our kernels
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong = args.scaling == "strong"
    n_total = N_ELEMS * (1 if strong else world)
    b0, e0 = D.shard_range(n_total, rank, world)
    n = e0 - b0                               # weak: 2^28 per GPU; strong: 2^28 / N per GPU
    # synthetic shard of this rank: uniform(1e-3, 1), seeded per rank; NaN-free, no signed-zero mix
    g = torch.Generator(device="cuda").manual_seed(SEED + rank)
    data = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * (1.0 - 1e-3) + 1e-3
    result = torch.empty(1, dtype=torch.float64, device="cuda")
    functor = P.ReduceFunctor(ob.OP_MIN, data, result)

    # N > 1: the path's only exchange is one fp64 partial per rank.  Default: fused into the reduction kernel over NVLink
    # peer memory (onk_reduce_xchg_f64).  --exchange nccl: separate NCCL all-reduce(MIN) after the kernel.
    xchg = None
    xchg_lanes = None
    xchg_transport = None
    if world > 1 and args.exchange in ("fused", "fused-host"):
        # device memory over NVLink first; where CUDA IPC / peer access is refused, the same exchange through pinned host
        # memory in POSIX shared memory; NCCL only if neither can be set up.  Every rank must take the same path.
        for transport in (("device", "host") if args.exchange == "fused" else ("host",)):
            made = []
            try:
                made = [D.PeerExchange(transport=transport) for _ in range(3)]      # headline + one exchange per lane of the C5 sequence
                ok = 1
            except Exception as ex:
                print(f"[bench] rank {rank}: fused exchange over {transport} memory unavailable ({str(ex)[:120]})", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                xchg, xchg_lanes, xchg_transport = made[0], (made[1], made[2]), transport
                break
            for m in made:
                m.close()

    flush = _L2Flush(torch, ob, P) if n * 8 < (512 << 20) else None   # strong scaling at large N: the shard may fit L2

    def step():
        if xchg is not None:
            xchg.reduce(ob.OP_MIN, data, n=n, out=result, lane=0)   # ONE kernel: local reduce + cross-GPU combine
        else:
            functor.launch(n, 0)              # one pass of the hot path over this rank's shard
            if world > 1:
                D.combine_across_ranks(ob.OP_MIN, result)

    K, W = max(1, args.steps), max(3, args.warmup)
    # the clock sampler starts BEFORE the warm-up (its own start-up needs ~0.15 s of host time): nothing may idle the GPU
    # between the warm-up steps and the timed region, or the first timed steps run while the clocks ramp back up
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    step()
    torch.cuda.synchronize()
    expect = float(data.min().item())
    if world > 1:
        te = torch.tensor([expect], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MIN)
        expect = float(te.item())
    assert float(result.item()) == expect, "reduce_min result differs from the check value"
    if world > 1:
        dist.barrier()
    # clock spin-up: the GPU idled while the host prepared (sampler start, check); ~30 ms of the same kernel bring the clocks
    # back up before the W warm-up steps proper (untimed, same on every rank; skipped under a profiler)
    SPINUP = 0 if args.profile_mode else 100
    for _ in range(SPINUP):
        step()
    for _ in range(W):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ob.launch_count()
    t_wall0 = time.time()
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()                              # exactly K steps between the two events, nothing else in the stream
        for k in range(K):
            step()
        ev1.record()
        torch.cuda.synchronize()
        total_ms = ev0.elapsed_time(ev1)
        flush_launches = 0
    else:
        # small shards: L2 flushed (by a read of 1 GiB) before every timed step, one event pair per step
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            flush()
            evs[k][0].record(); step(); evs[k][1].record()
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        flush_launches = K
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    launches = ob.launch_count() - launches0 - flush_launches
    if world > 1:
        total_ms = _max_over_ranks(torch, dist, world, total_ms)
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl.item())
    ms_per_step = total_ms / K
    value = n_total * 8 / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: device time of the kernel alone, measured live ----
    if world == 1 and flush is None:
        # the timed region above IS K launches of this kernel and nothing else: same events, same number
        kernel_ms = ms_per_step
    else:
        # N > 1: a step is the kernel plus the cross-GPU combine; time the local kernel alone the same way
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        k0.record()
        for k in range(K):
            functor.launch(n, 0)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms = k0.elapsed_time(k1) / K
    # per-launch spread, from a separate short pass with an event between launches (the events cost ~1 % themselves)
    Ks = min(K, 50)
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(Ks + 1)]
    kev[0].record()
    for k in range(Ks):
        functor.launch(n, 0)
        kev[k + 1].record()
    torch.cuda.synchronize()
    per_step = sorted(kev[k].elapsed_time(kev[k + 1]) for k in range(Ks))
    # clocks: sampled from the start of the timed region to the end of the kernel-only passes above (the same kernel under
    # the same load: more NVML samples than the few milliseconds of the K timed steps alone would give)
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    peak, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if peaks and "hbm_gbs" in peaks \
        else (6650.0, "fallback (B200_PROFILING.md)")
    achieved = n * 8 / (kernel_ms * 1e-3) / 1e9
    traffic = read_json(os.path.join(ROOT, "profiles", "roofline_traffic.json"))
    big = n >= 12288 * 8192                    # onk_reduce.cu: the TMA-staged body serves inputs of 768 MiB and more
    roofline = {"bound": "hbm", "kernel": "onk::reduce_f64_tma_kernel<MIN>" if big else "onk::reduce_f64_kernel<MIN>", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "frac_of_nominal_8TBs": round(achieved / 8000.0, 4), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": n * 8, "kernel_ms": round(kernel_ms, 5),
                "traffic": (traffic or {}).get("reduce_f64_kernel_dram_bytes_per_launch") if n == N_ELEMS else None,
                "traffic_source": "static: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this size (profiles/roofline_traffic.json), not measured in this run"}

    # ---- the headline kernel against the ORACLE on this rank's shard (outside every timed region): min and max bit for bit,
    # sum within 1e-12; at N > 1 the fused result must equal the minimum of the ranks' oracle minima ----
    shard_check = None
    if not args.profile_mode:
        try:
            import oracle
            try:
                ncpu = len(os.sched_getaffinity(0))
            except AttributeError:
                ncpu = os.cpu_count() or 1
            oracle.orc.set_num_threads(max(1, ncpu // world))
            host = data.cpu().numpy()
            got = {}
            for name, op in (("min", ob.OP_MIN), ("max", ob.OP_MAX), ("sum", ob.OP_SUM)):
                P.ReduceFunctor(op, data, result).launch(n, 0)
                got[name] = float(result.item())
            omin, omax, osum = oracle.orc.reduce_min(host), oracle.orc.reduce_max(host), oracle.orc.reduce_sum_ld(host)
            ok = got["min"] == omin and got["max"] == omax and abs(got["sum"] - osum) <= 1e-12 * float(np.abs(host).sum())
            gmin = omin
            if world > 1:
                t = torch.tensor([omin], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                gmin = float(t.item())
            ok = ok and expect == gmin
            shard_check = "ok" if ok else f"FAILED on rank {rank}: kernel {got} vs oracle min {omin!r} max {omax!r} sum {osum!r}, fused {expect!r} vs {gmin!r}"
            del host
        except Exception as ex:  # pragma: no cover
            shard_check = f"FAILED on rank {rank}: {str(ex)[:200]}"

    # ---- e2e: the plugin-facing host-buffer call (host -> device copies inside the timed region) ----
    e2e = None
    e2e_mp = None
    if not args.profile_mode:
        try:
            e2e = e2e_measure(K, world, rank, torch, dist, ob, P, data, n)
        except Exception as ex:  # pragma: no cover
            e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}
    del data, functor
    torch.cuda.empty_cache()

    # ---- fixed-size scaling, the C5 sequence on N GPUs, multi-GPU parity (all outside the headline's timed region) ----
    strong_obj = c5 = checks = None
    failures = []

    def section(fn):
        # a failed check is recorded and the run goes on: every rank walks through the same collectives, then all fail together
        try:
            return fn()
        except AssertionError as ex:
            failures.append(str(ex)[:300])
            print(f"[bench] CHECK FAILED on rank {rank}: {ex}", file=sys.stderr)
            return {"failed": str(ex)[:300]}

    if not args.profile_mode and not args.no_extras:
        if world > 1:
            strong_obj = section(lambda: strong_scaling(min(K, 50), 3, world, rank, torch, dist, ob, P, D, xchg))
        c5 = section(lambda: c5_lane_sequence(min(K, 10), 3, world, rank, torch, dist, ob, P, D, xchg_lanes))
        e2e_mp = section(lambda: e2e_multipass(K, world, rank, torch, dist, ob, D))
        if world > 1 and xchg is not None:
            from tests import _multigpu_checks
            checks = _multigpu_checks.run_checks(rank, world)        # never raises: {name: "ok" | message}
            failures += [f"{k}: {v}" for k, v in checks.items() if v != "ok"]
    if shard_check is not None:
        checks = dict(checks or {})
        checks["headline_shard_min_max_sum_vs_oracle"] = shard_check
        if shard_check != "ok":
            failures.append(shard_check)
    failed = "; ".join(failures) if failures else None
    if world > 1:
        bad = torch.tensor([1 if failed else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()) and not failed:
            failed = "a check failed on another rank"

    cpu_baseline = None
    kernels = None
    if rank == 0 and not args.profile_mode:
        if world == 1 and not args.no_extras:
            try:
                kernels = secondary_kernels(ob, torch)
            except Exception as ex:  # pragma: no cover
                kernels = {"error": str(ex)[:200]}
        if world == 1:
            try:
                cpu_baseline = _cpu_baseline_object(cpu_reduce_min_baseline(args.cpu_seconds))
                if kernels is not None and "error" not in kernels:
                    kernels["cpu_reference_same_box"] = cpu_secondary_baselines()
                if e2e_mp is not None:
                    e2e_mp["cpu_baseline"] = cpu_multipass_baseline()
            except Exception as ex:  # pragma: no cover
                cpu_baseline = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    if rank == 0:
        method = {"sharding": (f"element range x{world}; one fp64 partial per rank and step, combined " + (("inside the reduction kernel over NVLink peer memory (rank order)" if xchg_transport == "device" else "inside the reduction kernel through pinned host memory (rank order; no peer access on this box)") if xchg is not None else "by an NCCL all-reduce(min)")) if world > 1 else "single GPU",
                  "timing": "CUDA events on the launching stream, max over ranks",
                  "before_the_timed_region": f"{SPINUP} untimed launches of the same step (clock spin-up after host-side preparation), then the {W} warm-up steps, no host work in between"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.scaling), "method": method,
            "elements_per_s": round(value * 1e9 / 8, 1),
            "frac_of_nominal_8TBs_per_gpu": round(value / world / 8000.0, 4),
            "kernel_launch_ms_min_median_max": [round(per_step[0], 5), round(per_step[len(per_step) // 2], 5), round(per_step[-1], 5)],
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }
        if strong_obj is not None:
            line["strong"] = strong_obj
        if c5 is not None:
            line["C5"] = c5
        if e2e_mp is not None:
            line["e2e_multipass"] = e2e_mp
        if checks is not None:
            line["checks"] = checks
        if failed:
            line["checks_failed"] = failed
        if kernels is not None:
            line["kernels"] = kernels
        emit_json(line)
    if xchg is not None:
        torch.cuda.synchronize()
        dist.barrier()                            # no peer is still storing into this rank's buffers
        xchg.close(); xchg_lanes[0].close(); xchg_lanes[1].close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 3 if failed else 0


class _OnlyJsonOnStdout:
    """Libraries (NCCL's version banner, for one) write to fd 1.  The contract is ONE JSON line on stdout: fd 1 is pointed
    at stderr for the whole run and the line is written to the saved descriptor at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str) -> None:
        os.write(self.saved, (text.rstrip("\n") + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


_STDOUT = None


def emit_json(obj) -> None:
    if _STDOUT is not None:
        _STDOUT.emit(json.dumps(obj))
    else:
        print(json.dumps(obj))


def main() -> int:
    global _STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: 2^28 doubles per GPU (default, the driver's efficiency arithmetic); strong: 2^28 doubles in total")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="wall budget per CPU-baseline variant")
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused-host", "nccl"], help="N>1: cross-GPU combine inside the kernel (NVLink peer stores; fused-host: through pinned host memory) or a separate NCCL all-reduce")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary per-config kernel timings")
    ap.add_argument("--profile-mode", action="store_true", help="timed region + roofline only (for ncu runs): no e2e, no CPU baseline, no extras")
    args = ap.parse_args()
    with _OnlyJsonOnStdout() as guard:
        _STDOUT = guard
        try:
            if args.impl == "reference":
                return run_reference_arm(args)
            return run_ours(args)
        finally:
            _STDOUT = None


if __name__ == "__main__":
    sys.exit(main())
