#!/usr/bin/env python
"""BASELINE.json configs[1]: random complex64 CSR SpMM sweep (examples/spmm.py style) through the public
Backend interface (csr_matrix.forward / .adjoint), 1M x 1M, 16-64 nnz/row, 1-32 right-hand sides,
against the HBM roofline.  Prints a markdown table; GPU only.

    python tools/bench_spmm.py [--rows 1000000] [--reps 5] > profiles/rNN_spmm_sweep.md
"""
import argparse
import json
import os
import sys

import numpy as np
import scipy.sparse as spp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import bench
    from indigo_b200 import B200Backend, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    peak, src = bench.peaks()
    B = B200Backend(0)
    L2 = torch.cuda.get_device_properties(0).L2_cache_size
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clock = 1.965e9
    C64 = np.dtype("complex64")
    rows = cols = args.rows
    print("# cfg2: random complex64 CSR SpMM sweep, %d x %d, through Backend.csr_matrix.forward/.adjoint" % (rows, cols))
    print()
    print("Algorithmic bytes (SURVEY.md 8d): 12*nnz + 4*(rows+1) + 8*ncols*(rows+cols).  Peak %.1f GB/s, %s." % (peak, src))
    print("Multi-column products take the coil-interleaved path (interleave -> gather -> deinterleave, 3 launches);")
    print("adjoints use the stored conjugate transpose (built once on the device, not timed).  Median of %d, CUDA events." % args.reps)
    print()
    print("`gather floor`: what uniformly random columns cost on this memory system, which the one-pass figure of 8(d) leaves")
    print("out: every stored entry pulls ceil(8*ncols/32) 32-byte sectors of X; X (8*ncols*cols bytes) exceeds the %d MB L2" % (L2 >> 20))
    print("from 16 columns on, so a fraction 1 - L2/|X| of those sectors comes from DRAM; and every gather is one L1")
    print("wavefront per 128 bytes (1 wavefront/cycle/SM).  floor = max(8(d) bytes / peak, matrix + gathered DRAM sectors")
    print("/ peak, wavefronts / (SMs * clock)); `vs floor` = floor / measured.")
    print()
    print("| nnz/row | ncols | fwd ms | fwd GB/s | fwd frac | adj ms | adj GB/s | adj frac | gather floor ms | fwd vs floor |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")

    def timed(fn):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    for r in (16, 32, 64):
        rs = np.random.RandomState(r)
        ptr, ind, val = synth.random_csr(rs, rows, cols, r)
        A = spp.csr_matrix((val, ind, ptr), shape=(rows, cols))
        A.sort_indices()
        Ad = B.csr_matrix(B, A)
        for n in (1, 2, 4, 8, 16, 32):
            x = B.copy_array(synth.rand64c(rs, cols, n)); y = B.zero_array((rows, n), C64)
            xt = B.copy_array(synth.rand64c(rs, rows, n)); yt = B.zero_array((cols, n), C64)
            tf = timed(lambda: Ad.forward(y, x))
            ta = timed(lambda: Ad.adjoint(yt, xt))
            alg = 12 * A.nnz + 4 * (rows + 1) + 8 * n * (rows + cols)
            xbytes = 8.0 * n * cols
            sect = -(-8 * n // 32) * 32
            miss = max(0.0, 1.0 - L2 / xbytes)
            dram = 12.0 * A.nnz + A.nnz * sect * miss + xbytes * min(1.0, 1.0 - miss + 1e-9) + 8.0 * n * rows
            wavefronts = A.nnz * (-(-8 * n // 128))
            floor = max(alg / peak / 1e6, dram / peak / 1e6, wavefronts / (sms * clock) * 1e3)
            print("| %d | %d | %.3f | %.0f | %.2f | %.3f | %.0f | %.2f | %.3f | %.2f |" % (
                r, n, tf, alg / tf / 1e6, alg / tf / 1e6 / peak, ta, alg / ta / 1e6, alg / ta / 1e6 / peak, floor, floor / tf))
            sys.stdout.flush()
            del x, y, xt, yt
        del Ad


if __name__ == "__main__":
    main()
