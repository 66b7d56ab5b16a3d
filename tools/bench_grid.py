#!/usr/bin/env python
"""Micro-benchmark of the gridding SpMM pair of a workload (development tool, GPU only).

    python tools/bench_grid.py [--workload cfg3] [--coils 16] [--reps 5]

Times, with CUDA events on the launching stream, the pieces of Backend.ccsrmm for the
Kaiser-Bessel matrix G' (forward: grid -> samples) and its stored conjugate transpose
(adjoint: samples -> grid): interleave / gather / deinterleave.  Prints one JSON line.
Kept short so that it can run under `ncu --set full -k regex:csrmm_il`.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import bench
    from indigo_b200 import B200Backend
    from indigo_b200.sense import gridding_matrix_device

    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--coils", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    C = args.coils or wl["C"]
    B = B200Backend(0)
    lib, s = B._lib, B._stream
    t0 = time.time()
    G, oN, _, _ = gridding_matrix_device(B, wl["N"], bench.make_traj(wl["traj"]), wl["oversamp"])
    B.barrier(); t_g = time.time() - t0
    t0 = time.time()
    tp, ti, tv = G._stored_adjoint()
    B.barrier(); t_t = time.time() - t0
    m, k = G.shape
    nnz = int(G.values.size)
    xil = torch.rand(k * C * 2, device="cuda", dtype=torch.float32)
    kil = torch.rand(m * C * 2, device="cuda", dtype=torch.float32)
    xcm = torch.rand(k * C * 2, device="cuda", dtype=torch.float32)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(max(args.reps, 0)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)) if ts else 0.0

    out = {"workload": args.workload, "coils": C, "m": m, "k": k, "nnz": nnz, "build_G_s": round(t_g, 2),
           "build_GH_s": round(t_t, 2)}
    import ctypes
    out["fwd_gather_ms"] = timed(lambda: lib.ccsrmm_il(s, m, k, C, nnz, 1.0, 0.0, G.values.ptr, G.colInds.ptr,
                                                       G.rowPtrs.ptr, xil.data_ptr(), C, kil.data_ptr(), C, None, 0))
    for rpg in (1, 2, 8):
        out["fwd_gather_rpg%d_ms" % rpg] = timed(lambda: lib.ccsrmm_il(
            s, m, k, C, nnz, 1.0, 0.0, G.values.ptr, G.colInds.ptr, G.rowPtrs.ptr, xil.data_ptr(), C, kil.data_ptr(), C,
            None, rpg))
    out["adj_gather_ms"] = timed(lambda: lib.ccsrmm_il(s, k, m, C, nnz, 1.0, 0.0, tv.ptr, ti.ptr, tp.ptr,
                                                       kil.data_ptr(), C, xil.data_ptr(), C, None, 0))
    ref = xil.clone()
    # stored adjoint with rows in tile-major order of the grid
    for tile in ((4, 4, 4), (8, 4, 4), (8, 8, 8)):
        grid3 = (ctypes.c_int64 * 3)(*oN); tile3 = (ctypes.c_int64 * 3)(*tile)
        padded = ctypes.c_int64()
        lib.grid_tile_rank(s, grid3, tile3, None, None, ctypes.byref(padded))
        kp = padded.value
        colrank = torch.empty(k, dtype=torch.int32, device="cuda")
        rowmap = torch.empty(kp, dtype=torch.int32, device="cuda")
        lib.grid_tile_rank(s, grid3, tile3, colrank.data_ptr(), rowmap.data_ptr(), ctypes.byref(padded))
        t_ptr = torch.empty(kp + 1, dtype=torch.int32, device="cuda")
        t_ind = torch.empty(nnz, dtype=torch.int32, device="cuda")
        t_val = torch.empty(nnz * 2, dtype=torch.float32, device="cuda")
        work = torch.empty(kp + 1, dtype=torch.int32, device="cuda")
        t0 = time.time()
        lib.csr_transpose_conj(s, m, kp, nnz, G.values.ptr, G.colInds.ptr, G.rowPtrs.ptr, t_val.data_ptr(),
                               t_ind.data_ptr(), t_ptr.data_ptr(), work.data_ptr(), colrank.data_ptr())
        torch.cuda.synchronize()
        tag = "x".join(str(v) for v in tile)
        out["build_GH_tiled_%s_s" % tag] = round(time.time() - t0, 2)
        rows_per_tile = tile[0] * tile[1] * tile[2]
        for rpg in sorted({max(1, rows_per_tile // 16), 1}):
            xil.zero_()
            out["adj_tiled_%s_rpg%d_ms" % (tag, rpg)] = timed(lambda: lib.ccsrmm_il(
                s, kp, m, C, nnz, 1.0, 0.0, t_val.data_ptr(), t_ind.data_ptr(), t_ptr.data_ptr(), kil.data_ptr(), C,
                xil.data_ptr(), C, rowmap.data_ptr(), rpg))
            out["adj_tiled_%s_relerr" % tag] = float((xil - ref).norm() / ref.norm())
        del colrank, rowmap, t_ptr, t_ind, t_val, work
    out["interleave_grid_ms"] = timed(lambda: lib.interleave(s, k, C, xcm.data_ptr(), k, xil.data_ptr(), C))
    out["deinterleave_grid_ms"] = timed(lambda: lib.deinterleave(s, k, C, xil.data_ptr(), C, 0.0, 0.0, xcm.data_ptr(), k))
    alg = nnz * 12 + (m + 1) * 4 + 8 * C * (k + m)
    out["alg_GB"] = alg / 1e9
    out["fwd_gather_GBs"] = alg / out["fwd_gather_ms"] / 1e6
    out["adj_gather_GBs"] = (alg + (k - m) * 4) / out["adj_gather_ms"] / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
