"""Development probe (not the bench): times individual kernels at BASELINE shapes with
CUDA events and prints achieved algorithmic GB/s.  Run on the GPU box."""
import ctypes
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from indigo_b200 import B200Backend, synth        # noqa: E402

C64 = np.dtype('complex64')
B = B200Backend(0)
lib = B._lib
PEAK = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
MODE = sys.argv[1] if len(sys.argv) > 1 else "all"


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def report(name, nbytes, t):
    print("%-44s %9.3f ms  %8.1f GB/s  %5.1f%% of %.0f" % (name, t[0] * 1e3, nbytes / t[0] / 1e9, 100 * nbytes / t[0] / 1e9 / PEAK, PEAK), flush=True)


def dev(shape, dtype=C64):
    return B.empty_array(shape, dtype)


def fft_probe(shape):
    x = dev(shape); y = dev(shape)
    lib.memset0(B._stream, x.ptr, int(x.nbytes))
    n = int(np.prod(shape))
    report("fftn %s" % (shape,), 16 * n, timeit(lambda: B.fftn(y, x)))
    report("ifftn %s (in place)" % (shape,), 16 * n, timeit(lambda: B.ifftn(y, y)))
    del x, y


def blas_probe(n):
    x = dev((n,)); y = dev((n,))
    lib.memset0(B._stream, x.ptr, int(x.nbytes)); lib.memset0(B._stream, y.ptr, int(y.nbytes))
    report("axpby n=%d (beta!=0)" % n, 24 * n, timeit(lambda: B.axpby(0.5, y, 1.5, x)))
    report("axpby n=%d (beta=0)" % n, 16 * n, timeit(lambda: B.axpby(0, y, 1.5, x)))
    report("dot n=%d (incl. host sync)" % n, 16 * n, timeit(lambda: B.dot(x, y)))
    report("norm2 n=%d (incl. host sync)" % n, 8 * n, timeit(lambda: B.norm2(x)))


def stencil_csr(grid, nspokes, nread, seed=0):
    """Synthetic gridding-like CSR on the device: 125 taps (5x5x5, wrapped) per sample along radial spokes."""
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    d = torch.randn(nspokes, 3, device='cuda', generator=g, dtype=torch.float64)
    d = d / d.norm(dim=1, keepdim=True)
    r = (torch.arange(nread, device='cuda', dtype=torch.float64) - nread // 2) / nread
    pos = (r[None, :, None] * d[:, None, :]).reshape(-1, 3) * torch.tensor(grid, device='cuda') + torch.tensor([s // 2 for s in grid], device='cuda')
    base = torch.ceil(pos - 3).to(torch.int64)
    M = base.shape[0]
    o = torch.arange(5, device='cuda')
    ix = (base[:, 0:1] + o) % grid[0]; iy = (base[:, 1:2] + o) % grid[1]; iz = (base[:, 2:3] + o) % grid[2]
    ix, iy, iz = ix.sort(dim=1).values, iy.sort(dim=1).values, iz.sort(dim=1).values
    cols = (ix[:, None, None, :] + grid[0] * (iy[:, None, :, None] + grid[1] * iz[:, :, None, None])).reshape(M, 125)
    indptr = (torch.arange(M + 1, device='cuda', dtype=torch.int64) * 125).to(torch.int32)
    vals = torch.rand(M * 125, 2, device='cuda', generator=g, dtype=torch.float32)
    return M, int(np.prod(grid)), indptr, cols.reshape(-1).to(torch.int32), vals


def wrap(t, shape, dtype):
    """dndarray view over a torch tensor (probe plumbing)."""
    from indigo_b200.backend import DevPtr
    return B.dndarray(B, shape, np.dtype(dtype), own=False, data=DevPtr(t.data_ptr(), keep=t))


def csr_probe(grid, nspokes, nread, ncols):
    M, Kc, indptr, cols, vals = stencil_csr(grid, nspokes, nread)
    nnz = M * 125
    ptr_d, col_d, val_d = wrap(indptr, (M + 1,), np.int32), wrap(cols, (nnz,), np.int32), wrap(vals, (nnz,), C64)
    X = dev((Kc, ncols)); Y = dev((M, ncols))
    lib.memset0(B._stream, X.ptr, int(X.nbytes))
    alg = nnz * 12 + (M + 1) * 4 + 8 * ncols * (Kc + M)
    tag = "grid %s M=%d nnz=%.0fM ncols=%d" % (grid, M, nnz / 1e6, ncols)
    report("ccsrmm gather  " + tag, alg, timeit(lambda: B.ccsrmm(Y, (M, Kc), col_d, ptr_d, val_d, X, 1, 0, False, True)))
    report("ccsrmm atomic scatter " + tag, alg, timeit(lambda: B.ccsrmm(X, (M, Kc), col_d, ptr_d, val_d, Y, 1, 0, True, False), reps=3, warm=1))
    # stored adjoint
    t0 = time.time()
    t_ptr = dev((Kc + 1,), np.int32); t_col = dev((nnz,), np.int32); t_val = dev((nnz,)); work = dev((Kc + 1,), np.int32)
    lib.csr_transpose_conj(B._stream, M, Kc, nnz, val_d.ptr, col_d.ptr, ptr_d.ptr, t_val.ptr, t_col.ptr, t_ptr.ptr, work.ptr)
    print("   device transpose took %.2f s" % (time.time() - t0))
    alg_t = nnz * 12 + (Kc + 1) * 4 + 8 * ncols * (Kc + M)
    report("ccsrmm stored-adjoint gather " + tag, alg_t, timeit(lambda: B.ccsrmm(X, (Kc, M), t_col, t_ptr, t_val, Y, 1, 0, False, True)))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "peak", PEAK)
    if MODE == "fft":          # short run for ncu
        fft_probe((416, 416, 416, 4))
        sys.exit(0)
    if MODE == "csr":
        csr_probe((416, 416, 416), 2048, 416, 16)
        sys.exit(0)
    blas_probe(8998912)
    blas_probe(1 << 26)
    fft_probe((512, 512, 2, 8))
    fft_probe((416, 16 * 416 * 416))          # axis-0 pass alone
    fft_probe((416, 416, 416 * 16))           # axis 0 + axis 1
    fft_probe((416, 416, 416, 16))
    fft_probe((512, 512, 256, 12))
    csr_probe((512, 512, 2), 402, 512, 8)
    csr_probe((416, 416, 416), 2048, 416, 16)
    csr_probe((416, 416, 416), 2048, 416, 2)
    print("launches", lib.launch_count())
