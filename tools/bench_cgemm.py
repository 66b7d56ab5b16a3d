#!/usr/bin/env python
"""BASELINE.json configs[4] (cfg5), the coil-compression step: DenseMatrix cgemm of a 12 x 48 matrix on the
coil-fastest k-space of 128 kz-planes x 48 spirals x 2048 samples (n = 12 582 912 columns), forward
(compression) and adjoint (expansion), through Backend.cgemm.  tcgen05 3xTF32 kernel against the mma.sync 3xTF32 kernel, the SIMT
kernel and the HBM roofline (algorithmic bytes 8*(m*k + k*n + m*n)).  GPU only.

    python tools/bench_cgemm.py [--n 12582912] [--reps 10] > profiles/rNN_cgemm_cfg5.md
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import bench
    from indigo_b200 import B200Backend, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128 * 48 * 2048)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--modes", default="0,3,1", help="cgemm_mode values to time: 0 mma.sync (default path), 3 tcgen05, 1 SIMT")
    args = ap.parse_args()
    peak, src = bench.peaks()
    B = B200Backend(0)
    C64 = np.dtype("complex64")
    m, k, n = 12, 48, args.n
    rs = np.random.RandomState(5)
    U = np.linalg.svd(synth.rand64c(rs, k, 64).astype(np.complex128), full_matrices=False)[0][:, :m]
    Md = B.copy_array(np.asfortranarray(U.conj().T.astype(C64)))
    # k-space columns: one random block tiled on the device (host generation of 4.8 GB is not the point here)
    blk = 1 << 16
    xb = torch.from_numpy(np.ascontiguousarray(synth.rand64c(rs, k, blk).T).view(np.float32).ravel()).cuda()   # column-major floats
    xd = B.zero_array((k, n), C64)
    flat = xd._arr._keep.view(torch.float32)        # device arrays are owned by a torch uint8 tensor
    reps_full, rem = divmod(2 * k * n, xb.numel())
    flat[:reps_full * xb.numel()].view(reps_full, -1).copy_(xb.unsqueeze(0).expand(reps_full, -1))
    if rem:
        flat[reps_full * xb.numel():2 * k * n].copy_(xb[:rem])
    yd = B.zero_array((m, n), C64)
    zd = B.zero_array((k, n), C64)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    alg = 8.0 * (m * k + k * n + m * n)
    flops = 8.0 * m * k * n
    print("# cfg5 coil compression: Backend.cgemm, M %d x %d, %d coil-fastest columns (complex64)" % (m, k, n))
    print()
    print("Algorithmic bytes 8(mk + kn + mn) = %.2f GB per call, %.1f GFLOP (complex64-equivalent; the tensor-core" % (alg / 1e9, flops / 1e9))
    print("kernel issues 3x that as TF32 MMAs).  Peak %.1f GB/s, %s.  Median of %d, CUDA events, operands (%.1f GB)" % (peak, src, args.reps, alg / 1e9))
    print("far larger than L2.")
    print()
    print("| product | kernel | ms | GB/s | frac of HBM peak | TFLOP/s (complex64-equivalent) |")
    print("|---|---|---:|---:|---:|---:|")
    for label, fn in (("Y(12 x n) = M X  (compression)", lambda: B.cgemm(yd, Md, xd, 1.0, 0.0, forward=True)),
                      ("Z(48 x n) = M^H Y (expansion)", lambda: B.cgemm(zd, Md, yd, 1.0, 0.0, forward=False))):
        for mode, name in ((0, "mma.sync, 3xTF32 (cgemm_tc_kernel; default)"), (3, "tcgen05 + TMEM, warp-specialised, 3xTF32 (cgemm_t5ws_kernel)"), (1, "SIMT fp32 (cgemm_kernel)")):
            if str(mode) not in args.modes.split(","):
                continue
            B._lib.cgemm_mode(mode)
            t = timed(fn)
            print("| %s | %s | %.3f | %.0f | %.2f | %.1f |" % (label, name, t, alg / t / 1e6, alg / t / 1e6 / peak, flops / t / 1e9))
            sys.stdout.flush()
    B._lib.cgemm_mode(0)
    # accuracy of the timed configuration against fp64 on the first block of columns
    B.cgemm(yd, Md, xd, 1.0, 0.0, forward=True)
    got = yd.to_host()[:, :4096]
    xh = xb.cpu().numpy().view(np.complex64).reshape((blk, k)).T[:, :4096]
    want = U.conj().T.astype(C64).astype(np.complex128) @ xh.astype(np.complex128)
    print()
    print("rel-L2 of the tensor-core product against fp64: %.2e" % (np.linalg.norm(got - want) / np.linalg.norm(want)))


if __name__ == "__main__":
    main()
