#!/usr/bin/env python
"""Micro-benchmark of the gridding SpMM pair of a workload (development tool, GPU only).

    python tools/bench_grid.py [--workload cfg3] [--coils 16] [--reps 5]

Times, with CUDA events on the launching stream, the interleaved gathers for the Kaiser-Bessel
matrix G' (forward: grid -> samples) and its stored conjugate transpose in tile-major row order
(adjoint: samples -> grid), complex and real-packed entries, with and without the long-row split.
Prints one JSON line.  Short enough to run under `ncu --set full -k regex:csrmm_il`.
"""
import argparse
import ctypes
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
    ap.add_argument("--thresh", type=int, default=512)
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    C = args.coils or wl["C"]
    B = B200Backend(0)
    lib, s = B._lib, B._stream
    G, oN, _, _ = gridding_matrix_device(B, wl["N"], bench.make_traj(wl["traj"]), wl["oversamp"])
    m, k = G.shape
    nnz = int(G.values.size)
    torch.manual_seed(1)
    xil = torch.randn(k * C * 2, device="cuda", dtype=torch.float32)
    kil = torch.randn(m * C * 2, device="cuda", dtype=torch.float32)

    def timed(fn):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(max(args.reps, 0)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return round(float(np.median(ts)), 3) if ts else 0.0

    out = {"workload": args.workload, "coils": C, "m": m, "k": k, "nnz": nnz}
    # stored adjoint, tile-major rows
    tile = (4, 4, 4)
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
    out["build_GH_s"] = round(time.time() - t0, 2)
    lens = (t_ptr[1:] - t_ptr[:-1])
    out["GH_row_len_max"] = int(lens.max().item())
    out["GH_rows_gt_thresh"] = int((lens > args.thresh).sum().item())
    out["GH_nnz_in_long_rows"] = int(lens[lens > args.thresh].sum().item())
    cnt = ctypes.c_int()
    lib.csr_long_rows(s, kp, t_ptr.data_ptr(), args.thresh, None, 0, ctypes.byref(cnt))
    nlong = cnt.value
    longrows = torch.empty(max(nlong, 1), dtype=torch.int32, device="cuda")
    lib.csr_long_rows(s, kp, t_ptr.data_ptr(), args.thresh, longrows.data_ptr(), nlong, ctypes.byref(cnt))
    hmax = (ctypes.c_float * 2)()
    g_pk = torch.zeros(nnz + 2, dtype=torch.int64, device="cuda")
    t_pk = torch.zeros(nnz + 2, dtype=torch.int64, device="cuda")
    lib.csr_pack_real(s, nnz, G.values.ptr, G.colInds.ptr, g_pk.data_ptr(), hmax)
    lib.csr_pack_real(s, nnz, t_val.data_ptr(), t_ind.data_ptr(), t_pk.data_ptr(), hmax)

    fwd_c = lambda rpg: lib.ccsrmm_il(s, m, k, C, nnz, 1.0, 0.0, G.values.ptr, G.colInds.ptr, G.rowPtrs.ptr,
                                      xil.data_ptr(), C, kil.data_ptr(), C, None, rpg, None, 0, 0)
    fwd_r = lambda rpg: lib.ccsrmm_ilr(s, m, k, C, nnz, 1.0, 0.0, g_pk.data_ptr(), G.rowPtrs.ptr,
                                       xil.data_ptr(), C, kil.data_ptr(), C, None, rpg, None, 0, 0)
    adj_c = lambda nl: lib.ccsrmm_il(s, kp, m, C, nnz, 1.0, 0.0, t_val.data_ptr(), t_ind.data_ptr(), t_ptr.data_ptr(),
                                     kil.data_ptr(), C, xil.data_ptr(), C, rowmap.data_ptr(), 1,
                                     longrows.data_ptr(), nl, args.thresh)
    adj_r = lambda nl: lib.ccsrmm_ilr(s, kp, m, C, nnz, 1.0, 0.0, t_pk.data_ptr(), t_ptr.data_ptr(),
                                      kil.data_ptr(), C, xil.data_ptr(), C, rowmap.data_ptr(), 1,
                                      longrows.data_ptr(), nl, args.thresh)
    out["fwd_complex_ms"] = timed(lambda: fwd_c(0))
    out["fwd_packed_ms"] = timed(lambda: fwd_r(0))
    kil.normal_()
    out["adj_complex_nosplit_ms"] = timed(lambda: adj_c(0))
    ref = xil.clone()
    out["adj_complex_split_ms"] = timed(lambda: adj_c(nlong))
    out["adj_complex_split_relerr"] = float((xil - ref).norm() / ref.norm())
    out["adj_packed_nosplit_ms"] = timed(lambda: adj_r(0))
    out["adj_packed_split_ms"] = timed(lambda: adj_r(nlong))
    out["adj_packed_split_relerr"] = float((xil - ref).norm() / ref.norm())
    # shared-memory staged variants (rows_per_group = -4 / -8)
    kref = torch.empty_like(kil); xsave = xil.clone()
    xil.normal_()
    lib.ccsrmm_ilr(s, m, k, C, nnz, 1.0, 0.0, g_pk.data_ptr(), G.rowPtrs.ptr, xil.data_ptr(), C, kref.data_ptr(), C, None, 0, None, 0, 0)
    for u in (4, 41):
        out["fwd_staged_u%d_ms" % u] = timed(lambda: lib.ccsrmm_ilr(
            s, m, k, C, nnz, 1.0, 0.0, g_pk.data_ptr(), G.rowPtrs.ptr, xil.data_ptr(), C, kil.data_ptr(), C, None, -u, None, 0, 0))
        out["fwd_staged_u%d_relerr" % u] = float((kil - kref).norm() / kref.norm())
    kil.normal_()
    xref = torch.empty_like(xil)
    lib.ccsrmm_ilr(s, kp, m, C, nnz, 1.0, 0.0, t_pk.data_ptr(), t_ptr.data_ptr(), kil.data_ptr(), C, xref.data_ptr(), C,
                   rowmap.data_ptr(), 1, longrows.data_ptr(), nlong, args.thresh)
    for u in (4, 41):
        xil.zero_()
        out["adj_staged_u%d_ms" % u] = timed(lambda: lib.ccsrmm_ilr(
            s, kp, m, C, nnz, 1.0, 0.0, t_pk.data_ptr(), t_ptr.data_ptr(), kil.data_ptr(), C, xil.data_ptr(), C,
            rowmap.data_ptr(), -u, longrows.data_ptr(), nlong, args.thresh))
        out["adj_staged_u%d_relerr" % u] = float((xil - xref).norm() / xref.norm())
    # forward with the samples sorted by grid tile (rows permuted at setup, outputs scattered through rowmap)
    xil.normal_()
    lib.ccsrmm_ilr(s, m, k, C, nnz, 1.0, 0.0, g_pk.data_ptr(), G.rowPtrs.ptr, xil.data_ptr(), C, kref.data_ptr(), C, None, 0, None, 0, 0)
    for stile in ((4, 4, 4), (8, 4, 4), (8, 8, 8)):
        st3 = (ctypes.c_int64 * 3)(*stile)
        lib.grid_tile_rank(s, grid3, st3, None, None, ctypes.byref(padded))
        kp2 = padded.value
        cr2 = torch.empty(k, dtype=torch.int32, device="cuda"); rm2 = torch.empty(kp2, dtype=torch.int32, device="cuda")
        lib.grid_tile_rank(s, grid3, st3, cr2.data_ptr(), rm2.data_ptr(), ctypes.byref(padded))
        del rm2
        g_ptr2 = torch.empty(m + 1, dtype=torch.int32, device="cuda")
        g_pk2 = torch.zeros(nnz + 2, dtype=torch.int64, device="cuda")
        g_map = torch.empty(m, dtype=torch.int32, device="cuda")
        t0 = time.time()
        lib.csr_permute_rows(s, m, nnz, G.rowPtrs.ptr, g_pk.data_ptr(), cr2.data_ptr(), kp2, g_ptr2.data_ptr(),
                             g_pk2.data_ptr(), g_map.data_ptr())
        torch.cuda.synchronize()
        tag = "x".join(str(v) for v in stile)
        out["sort_rows_%s_s" % tag] = round(time.time() - t0, 3)
        for mode in (-41, -4):
            kil.zero_()
            out["fwd_sorted_%s_%s_ms" % (tag, "vc2" if mode == -4 else "vc1")] = timed(lambda: lib.ccsrmm_ilr(
                s, m, k, C, nnz, 1.0, 0.0, g_pk2.data_ptr(), g_ptr2.data_ptr(), xil.data_ptr(), C, kil.data_ptr(), C,
                g_map.data_ptr(), mode, None, 0, 0))
            out["fwd_sorted_%s_relerr" % tag] = float((kil - kref).norm() / kref.norm())
        del cr2, g_ptr2, g_pk2, g_map
    alg = nnz * 12 + (m + 1) * 4 + 8 * C * (k + m)
    out["alg_GB"] = round(alg / 1e9, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
