#!/usr/bin/env python
"""Generate tests/golden/*.json from the REAL reference (oracle/_ref/libonika_ref.so, built from /root/reference).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The fixtures are small, committed, and travel to the GPU box, where /root/reference does not exist.
Floats are stored as C99 hex strings (float.hex()) so that comparison is bit-exact.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

oracle.build(ref=True)
ref = oracle.ref


from tests._inputs import hx, lcg_uniform  # noqa: E402


def main():
    g = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libonika_ref.so (unmodified reference, OpenMP host path)",
         "reference_version": "onika 1.0.4 (CMakeLists.txt:41)", "float_format": "float.hex()"}

    # ---- SOATL layouts (field_arrays.h / packed_field_arrays.h / alloc_strategy.h) ----
    seqs = [[17, 33, 10, 0, 100, 40], [1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 31, 32, 33, 64, 65, 0],
            [1000, 999, 501, 500, 499, 250, 125, 126, 1, 0], [10000, 0, 10000, 20001, 5]]
    lay = []
    for cfg in ref.SOATL_CONFIGS:
        for seq in seqs:
            cap, by, offs = ref.soatl_layout(cfg, seq)
            lay.append(dict(cfg=cfg, **ref.SOATL_CONFIGS[cfg], sizes=seq, capacity=cap, storage_bytes=by, offsets=offs))
    g["soatl_layout"] = lay

    # ---- reduce min: the reference's own fixture (ramp i/N, answer 0.0) + hashed inputs, N = 2^20 ----
    ramp = ref.reduce_min_init_1m()
    m, _ = ref.reduce_min_chunks(ramp)
    red = [dict(kind="ramp", n=1 << 20, min=float(m).hex(), first3=hx(ramp[:3]), last=float(ramp[-1]).hex())]
    for seed in (1, 2, 1234):
        d = lcg_uniform(1 << 20, seed)
        m, _ = ref.reduce_min_chunks(d)
        red.append(dict(kind="lcg", seed=seed, n=1 << 20, min=float(m).hex(), probe=hx(d[[0, 1, 777, (1 << 20) - 1]])))
    d = lcg_uniform(4 << 20, 99)
    d[3_000_001] = 1e-9
    m, _ = ref.reduce_min_chunks(d)
    red.append(dict(kind="lcg_planted", seed=99, n=4 << 20, planted_index=3_000_001, planted_value=float(1e-9).hex(), min=float(m).hex()))
    g["reduce_min"] = red

    # ---- axpy through the reference parallel_for ----
    n = 1003
    x = lcg_uniform(n, 5, -1.0, 1.0)
    y = lcg_uniform(n, 6, -1.0, 1.0)
    y_out = y.copy()
    ref.axpy(y_out, x, 0.5)
    g["axpy"] = dict(n=n, a=0.5, x=hx(x), y=hx(y), y_out=hx(y_out))

    # ---- memset ----
    d = np.zeros(77)
    ref.memset(d, 3.25)
    g["memset"] = dict(n=77, v=3.25, out=hx(d))

    # ---- value add (1D rows and 2D space) ----
    a = lcg_uniform(7 * 13, 7).reshape(7, 13)
    a1 = a.copy()
    ref.value_add(a1, 1.1, ndims=1)
    a2 = a.copy()
    ref.value_add(a2, 1.1, ndims=2)
    g["value_add"] = dict(rows=7, cols=13, v=1.1, a=hx(a), out_1d=hx(a1), out_2d=hx(a2))

    # ---- 3D coverage: tutorial ranges (parallel_for_3d_benchmark.cu:96-125) ----
    N = 16
    m = np.zeros(N * N * N)
    ref.parallel_for_3d(m, N, (2, 3, 3), (N - 1, N - 3, N - 3), block=False)
    mb = np.zeros(4 * 4 * 4)
    ref.parallel_for_3d(mb, 4, (0, 0, 0), (4, 4, 4), block=True)
    g["parallel_for_3d"] = dict(n=N, start=[2, 3, 3], end=[N - 1, N - 3, N - 3], nonzero=[int(i) for i in np.nonzero(m)[0]],
                                max=float(m.max()), block_n=4, block_all_ones=bool((mb == 1.0).all()))

    # ---- lane sequence (tests/parallel_queue_lane.cu hybrid-lane / delayed / sequence; pfor and task modes) ----
    a1 = lcg_uniform(64, 8).reshape(8, 8)
    a2 = lcg_uniform(64, 9).reshape(8, 8)
    lanes = []
    for mode in (0, 1, 2):
        for tasks in (False, True):
            b1, b2 = a1.copy(), a2.copy()
            ref.lane_sequence(b1, b2, 1.1, 1.2, 1.3, 1.4, mode=mode, use_tasks=tasks)
            lanes.append(dict(mode=mode, tasks=tasks, a1_out=hx(b1), a2_out=hx(b2)))
    g["lane_sequence"] = dict(rows=8, cols=8, v=[1.1, 1.2, 1.3, 1.4], a1=hx(a1), a2=hx(a2), runs=lanes)

    # ---- SOATL benchmark kernel (tests/benchmark.cpp:96-104), n not a multiple of the chunk ----
    n = 53
    rx, ry, rz = lcg_uniform(n, 10, 0.0, 1.0), lcg_uniform(n, 11, 0.0, 1.0), lcg_uniform(n, 12, 0.0, 1.0)
    e, _ = ref.soatl_dist(rx, ry, rz)
    g["soatl_dist"] = dict(n=n, rx=hx(rx), ry=hx(ry), rz=hx(rz), e=hx(e))

    # ---- SOATL 8-field RMW ----
    n = 37
    cols = [lcg_uniform(n, 20 + k, 0.0, 1.0) for k in range(8)]
    c = [0.125 * (k + 1) for k in range(8)]
    out64 = [col.copy() for col in cols]
    ref.soatl_rmw8(out64, 1.0000001, c, align=64)
    out256 = [col.copy() for col in cols]
    ref.soatl_rmw8(out256, 1.0000001, c, align=256)
    g["soatl_rmw8"] = dict(n=n, s=1.0000001, c=c, cols=[hx(col) for col in cols], out=[hx(col) for col in out64],
                           same_for_align_256=all(np.array_equal(p, q) for p, q in zip(out64, out256)))

    # ---- SOATL serialisation to the <1,1> wire layout ----
    n = 13
    e = lcg_uniform(n, 30)
    rx, ry, rz = lcg_uniform(n, 31), lcg_uniform(n, 32), lcg_uniform(n, 33)
    atype = (lcg_uniform(n, 34, 0, 50)).astype(np.uint8)
    mid = (lcg_uniform(n, 35, 0, 500)).astype(np.int32)
    packed = ref.soatl_serialize(e, atype, rx, mid, ry, rz)
    g["soatl_serialize"] = dict(n=n, e=hx(e), atype=[int(v) for v in atype], rx=hx(rx), mid=[int(v) for v in mid], ry=hx(ry), rz=hx(rz),
                                packed_hex=packed.tobytes().hex())

    # ---- cell grid ----
    counts = np.array([0, 1, 15, 16, 17, 31, 32, 33, 64, 5, 0, 0, 16, 16, 2, 48, 7, 9, 23, 40], dtype=np.uint32)
    ec = lcg_uniform(int(counts.sum()), 40, -1.0, 1.0)
    sums, total, nbytes, _ = ref.cell_sum(counts, ec, schedule=2)
    g["cell_sum"] = dict(counts=[int(v) for v in counts], e=hx(ec), cell_sum=hx(sums), total=float(total).hex(), allocated_bytes=nbytes)

    # ---- id-based redistribution: index tables of communication_scheme_from_ids, run on in-process ranks ----
    from tests._inputs import xs_random_case, xs_simple_case
    g["xs_data_move"] = []
    for kind, args in (("simple", (3, 0, 20, 3)), ("simple", (3, 0, 107, 3)), ("simple", (7, 1, 111, 77)),
                       ("random", (2, 3, 97, 50)), ("random", (8, 1976, 110, 1000))):
        id_min, id_max, before, after = (xs_simple_case if kind == "simple" else xs_random_case)(*args)
        sch, _, _ = ref.xs_data_move(id_min, id_max, before, after)
        g["xs_data_move"].append(dict(kind=kind, args=list(args), send_indices=sch.send_indices.tolist(), recv_indices=sch.recv_indices.tolist(),
                                      send_count=sch.send_count.tolist(), recv_displ=sch.recv_displ.tolist()))

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(g, f, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "reference_vectors.json"), os.path.getsize(os.path.join(HERE, "reference_vectors.json")), "bytes")


if __name__ == "__main__":
    main()
