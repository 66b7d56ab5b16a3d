"""
TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference Backend
primitives on the SENSE-NUFFT hot path.  See oracle/README.md.  Nothing under
indigo_b200/ may import this module; tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg use it as the checker / timed CPU baseline.

Parity: PINNED against outputs of the unmodified reference NumpyBackend
(tests/golden/*.npz, produced by tests/golden/make_golden.py).

Conventions (the reference's): complex64, column-major; 2-D operands are
(rows, ncols) views that may have a leading dimension larger than `rows`
(any numpy view with strides (8, 8*ld) works); every function writes its
result into `y` in place and returns None, like the Backend methods do.
"""
import ctypes
import os

import numpy as np
import scipy.sparse as spp

C64 = np.dtype("complex64")


def _H(A):
    return A.conjugate().transpose()


# ---------------------------------------------------------------------------
# BLAS-1                                      indigo/backends/np.py:53-74
# ---------------------------------------------------------------------------
def axpby(beta, y, alpha, x):
    """y = beta*y + alpha*x.  np.py:53-58 (note: reads y even when beta == 0)."""
    y[...] = beta * y + alpha * x.reshape(y.shape, order="F")


def dot(x, y):
    """Re(x^H y) as a host float.  np.py:60-64."""
    return np.vdot(x, y).real


def norm2(x):
    """||x||_2 SQUARED as a host float.  np.py:66-69."""
    return np.linalg.norm(x) ** 2


def scale(x, alpha):
    """x *= alpha.  np.py:71-74."""
    x *= alpha


# ---------------------------------------------------------------------------
# dense                                       indigo/backends/np.py:76-97
# ---------------------------------------------------------------------------
def cgemm(y, M, x, alpha=1, beta=0, forward=True, left=True):
    """Y = alpha*op(M)*X + beta*Y (left) or alpha*X*op(M) + beta*Y.  np.py:76-87."""
    if not forward:
        M = np.conj(M.T)
    if left:
        X = x.reshape((M.shape[1], -1), order="F")
        Y = y.reshape((M.shape[0], -1), order="F")
        Y[...] = alpha * (M @ X) + beta * Y
    else:
        X = x.reshape((-1, M.shape[0]), order="F")
        Y = y.reshape((-1, M.shape[1]), order="F")
        Y[...] = alpha * (X @ M) + beta * Y


def csymm(y, M, x, alpha, beta, left=True):
    """Real-symmetric M; np.py:89-90 forwards to cgemm."""
    cgemm(y, M, x, alpha, beta, forward=True, left=left)


def onemm(y, x, alpha, beta):
    """Y = beta*Y + alpha*ones(M,K)*X.  np.py:95-97."""
    y[...] = beta * y + alpha * np.broadcast_to(x.sum(axis=0, keepdims=True), y.shape)


def fmax(val, arr):
    """arr = max(arr, val) on real and imaginary parts.  np.py:141-145."""
    arr[...] = np.maximum(arr.real, val) + 1j * np.maximum(arr.imag, val)


# ---------------------------------------------------------------------------
# FFT                                         indigo/backends/np.py:102-115
# ---------------------------------------------------------------------------
def fftn(y, x):
    """Unscaled forward C2C over all but the last axis of x (d0[,d1[,d2]], batch)."""
    axes = tuple(range(x.ndim - 1))
    y[...] = np.fft.fftn(x, axes=axes)


def ifftn(y, x):
    """UNSCALED inverse: numpy's ifftn times prod(dims).  np.py:109-115."""
    axes = tuple(range(x.ndim - 1))
    y[...] = np.fft.ifftn(x, axes=axes) * np.prod(x.shape[:-1])


# ---------------------------------------------------------------------------
# sparse                                      indigo/backends/np.py:120-136
# ---------------------------------------------------------------------------
def ccsrmm(y, A_shape, A_indx, A_ptr, A_vals, x, alpha=1, beta=0, adjoint=False, exwrite=False):
    """Y = alpha*op(A)*X + beta*Y, op = id or conjugate transpose.  np.py:120-127.

    `exwrite` is a hint only.  Y is read even when beta == 0 (run on initialised Y)."""
    A = spp.csr_matrix((A_vals, A_indx, A_ptr), shape=A_shape)
    if adjoint:
        y[...] = alpha * (_H(A) @ x) + beta * y
    else:
        y[...] = alpha * (A @ x) + beta * y


def cdiamm(y, shape, offsets, data, x, alpha=1.0, beta=0.0, adjoint=True):
    """DIA SpMM; `data` is the device layout (K x noffsets), i.e. scipy's data.T
    (backend.py:610).  np.py:129-136."""
    A = spp.dia_matrix((data.T, offsets), shape=shape)
    if adjoint:
        y[...] = alpha * (_H(A) @ x) + beta * y
    else:
        y[...] = alpha * (A @ x) + beta * y


def csr_inspect(A):
    """(row_frac, col_frac, exwrite) as backend.py:556-563 derives them from
    _customcpu.inspect (_customcpu.c:179-215)."""
    A = A.tocsr()
    per_col = np.bincount(A.indices, minlength=A.shape[1])
    nzrows = int(np.count_nonzero(np.diff(A.indptr)))
    nzcols = int(np.count_nonzero(per_col))
    exw = int(per_col.max(initial=0) <= 1)
    return nzrows / A.shape[0], nzcols / A.shape[1], exw


# ---------------------------------------------------------------------------
# CG                                          indigo/backends/backend.py:639-689
# ---------------------------------------------------------------------------
def cg(apply_A, b, x0, lamda=0.0, tol=1e-10, maxiter=100, iterates=None, allreduce=None):
    """Conjugate gradient on (A + lamda*I) x = b, the exact update order of
    Backend.cg.  `apply_A(out, inp)` overwrites `out` with A*inp.  Returns x;
    if `iterates` is a list, a copy of x is appended after every iteration.
    `allreduce` stands for team.allreduce in pdot/pnorm2 (backend.py:469-479)."""
    red = allreduce if allreduce is not None else (lambda v: v)
    x = np.array(x0, dtype=C64, order="F", copy=True)
    r = np.array(b, dtype=C64, order="F", copy=True)
    Ap = x.copy(order="F")
    apply_A(Ap, x)
    axpby(1, r, -1, Ap)
    axpby(1, r, -lamda, x)
    p = r.copy(order="F")
    rr = red(norm2(r))
    r0 = rr
    for _ in range(maxiter):
        apply_A(Ap, p)
        axpby(1, Ap, lamda, p)
        alpha = rr / red(dot(p, Ap))
        axpby(1, x, alpha, p)
        axpby(1, r, -alpha, Ap)
        r2 = red(norm2(r))
        beta = r2 / rr
        scale(p, beta)
        axpby(1, p, 1, r)
        rr = r2
        if iterates is not None:
            iterates.append(x.copy())
        if np.sqrt(rr / r0) < tol:
            break
    return x


# ---------------------------------------------------------------------------
# C restatement / compiled reference, loaded lazily (tests + cpu_baseline only)
# ---------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))


def load_oracle_c():
    """ctypes handle on liboracle_c.so (oracle_c.c), or None if not built."""
    path = os.path.join(_HERE, "liboracle_c.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    i64, f32, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    lib.oracle_csr_inspect.argtypes = [i64, i64, vp, vp, vp]
    lib.oracle_ccsrmm.argtypes = [ctypes.c_int, i64, i64, i64, f32, f32, vp, vp, vp, vp, i64, f32, f32, vp, i64]
    lib.oracle_onemm.argtypes = [i64, i64, i64, f32, f32, vp, i64, f32, f32, vp, i64]
    lib.oracle_fmax.argtypes = [i64, f32, vp]
    for fn in (lib.oracle_csr_inspect, lib.oracle_ccsrmm, lib.oracle_onemm, lib.oracle_fmax):
        fn.restype = None
    return lib


def load_ref_customcpu():
    """The reference's own _customcpu extension compiled by `make -C oracle ref`
    (oracle/_ref/), or None.  Exposes csrmm/onemm/max/inspect exactly as
    indigo/backends/_customcpu.c:249-255 registers them."""
    import importlib.util
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(ref_dir):
        return None
    so = sorted(f for f in os.listdir(ref_dir) if f.startswith("_customcpu") and f.endswith(".so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("_customcpu", os.path.join(ref_dir, so[0]))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
