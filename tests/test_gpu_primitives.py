"""
GPU parity tests: every Backend primitive of the hot path, executed by
B200Backend (-> C ABI -> sm_100a kernels), against the numpy oracle on the same
seeded inputs and against the golden vectors produced by the unmodified
reference.  Cases follow indigo/backends/test_backends.py (array moves :13-151,
FFT :153-180, CSR :183-243, BLAS-1 :259-330, cgemm :333-361, max :427-438,
DIA :441-477, csymm :493-523).  Tolerances: bit-exact for byte/index moves,
rel-L2 <= 1e-5 (usually ~1e-7) for complex64 arithmetic.
"""
import os
from itertools import product

import numpy as np
import pytest
import scipy.sparse as spp

from indigo_b200 import synth
from oracle import np_oracle as K

pytestmark = pytest.mark.gpu
C64 = np.dtype('complex64')


@pytest.fixture(scope="module")
def B():
    from indigo_b200 import B200Backend
    return B200Backend(0)


@pytest.fixture(params=["direct", "interleaved"])
def il(request, B):
    """Runs a CSR test twice: with the column-major gather kernels and with the
    coil-interleaved path (csrmm_il.cu) forced for every multi-column product."""
    old = B.il_min_work
    B.il_min_work = 0 if request.param == "interleaved" else (1 << 62)
    yield request.param
    B.il_min_work = old


def relerr(a, b):
    a = np.asarray(a).ravel(order='F'); b = np.asarray(b).ravel(order='F')
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# --------------------------------------------------------------------------- arrays
@pytest.mark.parametrize("n", [0, 1, 4, 8, 129, 100003])
def test_array_roundtrip_bit_exact(B, n):
    rs = np.random.RandomState(n)
    a = synth.rand64c(rs, n)
    d = B.copy_array(a)
    np.testing.assert_array_equal(d.to_host(), a)
    d2 = d.copy(); d._zero()
    np.testing.assert_array_equal(d2.to_host(), a)
    np.testing.assert_array_equal(d.to_host(), np.zeros_like(a))
    d3 = B.zero_array(a.shape, a.dtype); d3[:] = d2
    np.testing.assert_array_equal(d3.to_host(), a)
    assert isinstance(d._arr, __import__('ctypes').c_ulong) and d._arr.value != 0


def test_array_errors(B):
    a = np.random.rand(8).astype(C64)
    with pytest.raises(ValueError):
        B.zero_array((9,), a.dtype).copy_from(a)
    with pytest.raises(TypeError):
        B.zero_array(a.shape, np.complex128).copy_from(a)
    d = B.copy_array(synth.rand64c(np.random.RandomState(0), 8, 4))
    with pytest.raises(AssertionError):
        d[2:6, :].reshape((8, 2))


@pytest.mark.parametrize("s", [-2, -1, 1, 2])
def test_array_slice_1d(B, s):
    a = np.arange(10)
    d = B.copy_array(a)
    np.testing.assert_array_equal(d[:s].to_host(), a[:s])
    np.testing.assert_array_equal(d[s:].to_host(), a[s:])


@pytest.mark.parametrize("M,N,xb,xe,yb,ye", list(product([6, 7], [8, 9], [0, 2], [4, 5], [0, 1], [6, 7])))
def test_array_slice_2d(B, M, N, xb, xe, yb, ye):
    a = synth.rand64c(np.random.RandomState(M * N), M, N)
    d = B.copy_array(a)
    sub = d[xb:xe, yb:ye]
    np.testing.assert_array_equal(sub.to_host(), a[xb:xe, yb:ye])
    sub2 = B.zeros_like(sub); sub2.copy(sub)
    np.testing.assert_array_equal(sub2.to_host(), a[xb:xe, yb:ye])
    with d.on_host() as h:
        h += 1
    np.testing.assert_array_equal(d.to_host(), a + 1)


# --------------------------------------------------------------------------- BLAS-1
@pytest.mark.parametrize("n,alpha,beta", list(product([1, 10, 23, 129, 100001], [0.0, 1.0, -2.1 + 3j], [0.0, 0.5, 1.0, 1.5 - 1j])))
def test_axpby(B, n, alpha, beta):
    rs = np.random.RandomState(n)
    x, y = synth.rand64c(rs, n), synth.rand64c(rs, n)
    xd, yd = B.copy_array(x), B.copy_array(y)
    B.axpby(beta, yd, alpha, xd)
    want = y.copy(); K.axpby(beta, want, alpha, x)
    np.testing.assert_allclose(yd.to_host(), want, atol=1e-6)


def test_axpby_beta_zero_never_reads_y(B):
    x = synth.rand64c(np.random.RandomState(1), 1000)
    y = np.full(1000, np.nan + 1j * np.nan, dtype=C64)
    xd, yd = B.copy_array(x), B.copy_array(y)
    B.axpby(0, yd, 2.0, xd)
    np.testing.assert_allclose(yd.to_host(), 2 * x, rtol=1e-6)
    B.scale(yd, 0); assert not np.isnan(yd.to_host()).any()


def test_axpby_unaligned_views(B):
    rs = np.random.RandomState(2)
    x, y = synth.rand64c(rs, 1001), synth.rand64c(rs, 1003)
    xd, yd = B.copy_array(x), B.copy_array(y)
    for xo, yo in [(1, 1), (1, 2), (0, 3), (2, 0)]:
        yd.copy_from(y)
        B.axpby(0.5, yd[yo:yo + 900], 1.5 - 1j, xd[xo:xo + 900])
        want = y.copy(); want[yo:yo + 900] = 0.5 * y[yo:yo + 900] + (1.5 - 1j) * x[xo:xo + 900]
        np.testing.assert_allclose(yd.to_host(), want, atol=1e-6)


@pytest.mark.parametrize("n", [1, 10, 23, 129, 144, 1 << 20])
def test_dot_norm2_scale(B, n):
    rs = np.random.RandomState(n)
    x, y = synth.rand64c(rs, n), synth.rand64c(rs, n)
    xd, yd = B.copy_array(x), B.copy_array(y)
    x128, y128 = x.astype(np.complex128), y.astype(np.complex128)
    assert abs(B.dot(xd, yd) - np.vdot(x128, y128).real) <= 1e-6 * n
    assert abs(B.norm2(xd) - np.linalg.norm(x128) ** 2) <= 1e-6 * n
    assert abs(B.dot(xd, yd) - K.dot(x, y)) <= 2e-5 * max(1, abs(K.dot(x, y)))
    assert B.dot(xd, yd) == B.dot(xd, yd)                      # bit-reproducible reductions
    B.scale(xd, 1.1 - 2j)
    np.testing.assert_allclose(xd.to_host(), x * np.complex64(1.1 - 2j), rtol=1e-5)


# --------------------------------------------------------------------------- CSR
def _rand_csr(rs, m, n, density):
    A = (spp.random(m, n, density=density, format='csr', random_state=rs, dtype=np.float32)
         + 1j * spp.random(m, n, density=density, format='csr', random_state=rs, dtype=np.float32)).astype(C64).tocsr()
    A.sort_indices()
    return A


@pytest.mark.parametrize("Kc", [2, 3, 8, 16, 17, 32])
@pytest.mark.parametrize("alpha,beta", [(1, 0), (0.5 - 2j, 1.5)])
def test_csr_matrix_real_values_packed(B, Kc, alpha, beta, monkeypatch):
    """Real-valued matrices (gridding matrices on MRI grids) take the packed 8-byte-entry gather for
    multi-column products, forward and through the stored adjoint, with long rows split off."""
    monkeypatch.setattr(B, "il_min_work", 0)
    monkeypatch.setattr(B.csr_matrix, "long_thresh", 16)
    rs = np.random.RandomState(77 + Kc)
    M, N = 300, 211
    A = spp.random(M, N, density=0.08, format='csr', random_state=rs, dtype=np.float32).tolil()
    A[5, :] = rs.rand(N).astype(np.float32)                  # a long row and (transposed) a long column
    A[:, 7] = rs.rand(M, 1).astype(np.float32)
    A = A.tocsr().astype(C64); A.sort_indices()
    Ad = B.csr_matrix(B, A)
    x, y0 = synth.rand64c(rs, N, Kc), synth.rand64c(rs, M, Kc)
    yd = B.copy_array(y0)
    Ad.forward(yd, B.copy_array(x), alpha=alpha, beta=beta)
    assert Ad._packed('fwd') is not None and Ad._packed('fwd')['nlong'] > 0
    np.testing.assert_allclose(yd.to_host(), alpha * (A @ x) + beta * y0, atol=2e-4, rtol=1e-5)
    x, y0 = synth.rand64c(rs, M, Kc), synth.rand64c(rs, N, Kc)
    yd = B.copy_array(y0)
    Ad.adjoint(yd, B.copy_array(x), alpha=alpha, beta=beta)
    assert Ad._packed('adj') is not None and Ad._packed('adj')['nlong'] > 0
    np.testing.assert_allclose(yd.to_host(), alpha * (A.conj().T @ x) + beta * y0, atol=2e-4, rtol=1e-5)


@pytest.mark.parametrize("M,N,Kc,density", list(product([23, 45], [45, 23], [1, 8, 9, 17], [0.01, 0.1, 0.5])))
def test_csr_matrix(B, M, N, Kc, density, il):
    rs = np.random.RandomState(M * N + Kc)
    A = _rand_csr(rs, M, N, density)
    Ad = B.csr_matrix(B, A)
    np.testing.assert_array_equal(Ad.rowPtrs.to_host(), A.indptr)
    np.testing.assert_array_equal(Ad.colInds.to_host(), A.indices)
    rf, cf, exw = K.csr_inspect(A)
    assert (Ad._row_frac, Ad._col_frac, Ad._exwrite) == (rf, cf, exw)
    x = synth.rand64c(rs, N, Kc)
    yd = B.zero_array((M, Kc), C64)
    Ad.forward(yd, B.copy_array(x))
    np.testing.assert_allclose(yd.to_host(), A @ x, atol=1e-5)
    x = synth.rand64c(rs, M, Kc)
    for stored in (True, False):                 # stored-adjoint gather and atomic scatter
        B.stored_adjoints = stored
        yd = B.zero_array((N, Kc), C64)
        Ad.adjoint(yd, B.copy_array(x))
        np.testing.assert_allclose(yd.to_host(), A.conj().T @ x, atol=1e-5)
    B.stored_adjoints = True


@pytest.mark.parametrize("M,N,Kc,alpha,beta", list(product([23, 45], [1, 8, 9, 17], [18, 19], [0.0, 0.5, 1.5], [0.0, 1.0, 1.5])))
def test_exw_csr_matrix(B, M, N, Kc, alpha, beta, il):
    rs = np.random.RandomState(M + N + Kc)
    counts = rs.randint(0, 2, Kc)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    ind = rs.randint(0, M, counts.sum())
    A = spp.csr_matrix((synth.rand64c(rs, ind.size, order='C'), ind, ptr), shape=(Kc, M)).T.tocsr()
    A.sort_indices()
    Ad = B.csr_matrix(B, A)
    assert Ad._exwrite == 1
    x, y = synth.rand64c(rs, Kc, N), synth.rand64c(rs, M, N)
    yd = B.copy_array(y)
    Ad.forward(yd, B.copy_array(x), alpha=alpha, beta=beta)
    np.testing.assert_allclose(yd.to_host(), beta * y + alpha * (A @ x), atol=1e-5)
    x, y = synth.rand64c(rs, M, N), synth.rand64c(rs, Kc, N)
    yd = B.copy_array(y)
    Ad.adjoint(yd, B.copy_array(x), alpha=alpha, beta=beta)
    np.testing.assert_allclose(yd.to_host(), beta * y + alpha * (A.conj().T @ x), atol=1e-5)


def test_csr_golden_leading_dims(B, golden_dir, il):
    g = np.load(os.path.join(golden_dir, "primitives.npz"))
    m, k = (int(v) for v in g["csr_shape"])
    A = spp.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(m, k))
    Ad = B.csr_matrix(B, A)
    alpha, beta = complex(g["csr_alpha"]), complex(g["csr_beta"])
    xd, yd = B.copy_array(g["csr_fwd_xbig"]), B.copy_array(g["csr_fwd_ybig"])
    Ad.forward(yd[2:2 + m, :], xd[3:3 + k, :], alpha=alpha, beta=beta)
    assert relerr(yd.to_host(), g["csr_fwd_out"]) < 1e-6
    for stored in (True, False):
        B.stored_adjoints = stored
        xd, yd = B.copy_array(g["csr_adj_xbig"]), B.copy_array(g["csr_adj_ybig"])
        Ad.adjoint(yd[3:3 + k, :], xd[2:2 + m, :], alpha=alpha, beta=beta)
        assert relerr(yd.to_host(), g["csr_adj_out"]) < 1e-6
    B.stored_adjoints = True
    E = spp.csr_matrix((g["exw_data"], g["exw_indices"], g["exw_indptr"]), shape=tuple(g["exw_shape"]))
    Ed = B.csr_matrix(B, E)
    yd = B.copy_array(g["exw_y"])
    Ed.adjoint(yd, B.copy_array(g["exw_x"]), alpha=0.5, beta=1.5)
    assert relerr(yd.to_host(), g["exw_adj_out"]) < 1e-6


def test_csr_beta_zero_ignores_nan_in_y(B, il):
    rs = np.random.RandomState(4)
    A = _rand_csr(rs, 200, 300, 0.05)
    Ad = B.csr_matrix(B, A)
    x = synth.rand64c(rs, 300, 5)
    yd = B.copy_array(np.full((200, 5), np.nan, dtype=C64, order='F'))
    Ad.forward(yd, B.copy_array(x))
    assert relerr(yd.to_host(), A @ x) < 1e-6
    x = synth.rand64c(rs, 200, 5)
    for stored in (True, False):
        B.stored_adjoints = stored
        yd = B.copy_array(np.full((300, 5), np.nan, dtype=C64, order='F'))
        Ad.adjoint(yd, B.copy_array(x))
        assert relerr(yd.to_host(), A.conj().T @ x) < 1e-6
    B.stored_adjoints = True


@pytest.mark.parametrize("nnz_row,ncols", [(16, 1), (32, 8), (64, 3)])
def test_csr_sweep_shape_adjointness(B, nnz_row, ncols, il):
    """cfg2-style matrix (reduced rows) -- size-independent property <Ax,y> = <x,A^H y>
    plus a seeded oracle check on a row block."""
    rs = np.random.RandomState(nnz_row)
    rows = cols = 200000
    ptr, ind, val = synth.random_csr(rs, rows, cols, nnz_row)
    A = spp.csr_matrix((val, ind, ptr), shape=(rows, cols)); A.sort_indices()
    Ad = B.csr_matrix(B, A)
    x, y = synth.rand64c(rs, cols, ncols), synth.rand64c(rs, rows, ncols)
    xd, yd = B.copy_array(x), B.copy_array(y)
    Ax, AHy = B.zero_array((rows, ncols), C64), B.zero_array((cols, ncols), C64)
    Ad.forward(Ax, xd); Ad.adjoint(AHy, yd)
    lhs = np.vdot(Ax.to_host().astype(np.complex128), y.astype(np.complex128))
    rhs = np.vdot(x.astype(np.complex128), AHy.to_host().astype(np.complex128))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    assert relerr(Ax.to_host()[:5000], (A[:5000] @ x)) < 1e-6
    B.stored_adjoints = False
    AHy2 = B.zero_array((cols, ncols), C64); Ad.adjoint(AHy2, yd)
    B.stored_adjoints = True
    assert relerr(AHy2.to_host(), AHy.to_host()) < 1e-5


# --------------------------------------------------------------------------- FFT
@pytest.mark.parametrize("shape", [(23, 24, 25, 1), (24, 25, 23, 2), (25, 23, 24, 4), (24, 24, 24, 8), (22, 3), (23, 1), (24, 22, 2),
                                   (16, 13, 7, 3), (128, 96, 2), (416, 8, 4, 2), (8, 416, 3, 1), (4, 6, 416, 2), (512, 512, 2, 2), (1000, 3), (17 * 4, 5, 2)])
def test_fft_vs_oracle(B, shape):
    rs = np.random.RandomState(sum(shape))
    v = synth.rand64c(rs, *shape)
    want = np.zeros_like(v, order='F'); K.fftn(want, v)
    vd, ud = B.copy_array(v), B.zero_array(v.shape, C64)
    B.fftn(ud, vd)
    assert relerr(ud.to_host(), want) < 2e-6
    K.ifftn(want, v)
    B.ifftn(ud, vd)
    assert relerr(ud.to_host(), want) < 2e-6
    # round trip / n, and in place
    B.fftn(ud, vd); B.ifftn(ud, ud)
    np.testing.assert_allclose(ud.to_host() / np.prod(shape[:-1]), v, atol=2e-6)


@pytest.mark.parametrize("tag", ["fft3", "fft2", "fft1", "fft3b"])
def test_fft_golden(B, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "primitives.npz"))
    v = g[tag + "_in"]
    vd, ud = B.copy_array(v), B.zero_array(v.shape, C64)
    B.fftn(ud, vd); assert relerr(ud.to_host(), g[tag + "_fwd"]) < 2e-6
    B.ifftn(ud, vd); assert relerr(ud.to_host(), g[tag + "_inv"]) < 2e-6


def test_fft_full_size_axis_roundtrip(B):
    """One 416^3 coil volume (cfg3 grid): unitary round trip and Parseval."""
    rs = np.random.RandomState(416)
    shp = (416, 416, 416, 1)
    v = synth.rand64c(rs, *shp)
    vd, ud = B.copy_array(v), B.zero_array(shp, C64)
    B.fftn(ud, vd)
    n = float(np.prod(shp[:-1]))
    assert abs(B.norm2(ud) / (n * B.norm2(vd)) - 1) < 1e-5
    B.ifftn(ud, ud)
    B.axpby(1.0 / n, ud, -1.0, vd)
    assert np.sqrt(B.norm2(ud) / B.norm2(vd)) < 2e-6


# --------------------------------------------------------------------------- dense, one, dia, max
@pytest.mark.parametrize("m,n,k,alpha,beta,forward", list(product([10, 23, 129, 144], [10, 144], [23, 129], [1, 0.5 + 0.5j, 0.0], [0, 0.5], [True, False])))
def test_cgemm(B, m, n, k, alpha, beta, forward):
    rs = np.random.RandomState(m * n + k)
    y, M, x = synth.rand64c(rs, m, n), synth.rand64c(rs, m, k), synth.rand64c(rs, k, n)
    if not forward:
        x, y = y, x
    want = y.copy(order='F'); K.cgemm(want, M, x, alpha, beta, forward=forward)
    yd = B.copy_array(y)
    B.cgemm(yd, B.copy_array(M), B.copy_array(x), alpha, beta, forward=forward)
    assert relerr(yd.to_host(), want) < 1e-5 or np.allclose(yd.to_host(), want, atol=1e-4)


@pytest.fixture
def simt_gemm(B):
    """Switch handle for the two cgemm kernels: simt_gemm(True) forces the SIMT kernel."""
    def force(on):
        B._lib.cgemm_mode({False: 0, True: 1, "tcgen05": 3}[on])
    yield force
    B._lib.cgemm_mode(0)


# tall-skinny products (coil compression): the tensor-core 3xTF32 kernel.  Shapes walk the
# n-tile instantiations (m up to 64), k groups with a half-empty tail (k % 8 != 0), ragged column counts.
TC_SHAPES = [(12, 48, 5000), (48, 12, 5000), (6, 24, 1000), (1, 2, 33), (64, 6, 777), (17, 10, 4099), (33, 22, 100),
             (4, 96, 257), (24, 14, 64), (3, 192, 48),
             # tcgen05 kernel: one k chunk (2k <= 48) / two chunks, every accumulator width (N = 32 ... 128)
             (12, 32, 300), (5, 8, 129), (64, 48, 1000), (33, 4, 128), (20, 40, 2500), (48, 24, 1 << 15)]


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("alpha,beta", [(1, 0), (0.5 + 0.5j, 0.5 - 0.25j)])
@pytest.mark.parametrize("forward", [True, False])
@pytest.mark.parametrize("path", ["auto", "tcgen05"])
def test_cgemm_tensor_core(B, simt_gemm, shape, alpha, beta, forward, path):
    """Y = alpha*op(M)*X + beta*Y on the tensor cores against the fp64 product: three TF32 MMAs on (hi, lo)
    splits must stay at complex64 accuracy (plain TF32 would sit near 1e-3), and agree with the SIMT kernel."""
    m, k, n = shape
    rs = np.random.RandomState(m * 1000 + k)
    M = synth.rand64c(rs, m, k) if forward else synth.rand64c(rs, k, m)
    x, y = synth.rand64c(rs, k, n), synth.rand64c(rs, m, n)
    opM = M if forward else M.conj().T
    want = alpha * (opM.astype(np.complex128) @ x.astype(np.complex128)) + beta * y.astype(np.complex128)
    Md, xd = B.copy_array(M), B.copy_array(x)
    yd = B.copy_array(y)
    simt_gemm("tcgen05" if path == "tcgen05" else False)       # auto: mma.sync kernel; tcgen05: that kernel where it applies
    B.cgemm(yd, Md, xd, alpha, beta, forward=forward)
    got = yd.to_host()
    # the tensor cores truncate when they align the addends of the fp32 accumulator: a chain of 3*k/4 MMAs on
    # all-positive data (rand64c) drifts by ~2^-24 per MMA; 2e-6 up to cfg5's k = 48, the 1e-5 parity bar beyond
    tol = 2e-6 if k <= 64 else 1e-5
    assert relerr(got, want) < tol, relerr(got, want)
    simt_gemm(True)
    ys = B.copy_array(y)
    B.cgemm(ys, Md, xd, alpha, beta, forward=forward)
    assert relerr(ys.to_host(), want) < 2e-6
    assert relerr(got, ys.to_host()) < tol


def test_cgemm_tensor_core_leading_dims(B):
    """Raw C-ABI call with ldx > k and ldy > m (arena slices): padding rows of Y stay untouched."""
    m, k, n, ldx, ldy = 12, 48, 1500, 52, 14
    rs = np.random.RandomState(5)
    M, xf, yf = synth.rand64c(rs, m, k), synth.rand64c(rs, ldx, n), synth.rand64c(rs, ldy, n)
    Md, xd, yd = B.copy_array(M), B.copy_array(xf), B.copy_array(yf)
    B._lib.cgemm(B._stream, 0, m, n, k, 2.0, -1.0, Md.ptr, m, xd.ptr, ldx, 0.0, 0.0, yd.ptr, ldy)
    got = yd.to_host()
    want = (2 - 1j) * (M.astype(np.complex128) @ xf[:k].astype(np.complex128))
    assert relerr(got[:m], want) < 2e-6
    np.testing.assert_array_equal(got[m:], yf[m:])


def test_cgemm_cfg5_shape_at_size(B, simt_gemm):
    """12 x 48 compression matrix on 2^20 coil-fastest columns (cfg5 has 12.6 M): tensor-core kernel against the
    SIMT kernel on the device (norm of the difference), against fp64 on a slice, and compression followed by
    expansion with an orthonormal basis is a projection (idempotent): size-independent check."""
    m, k, n = 12, 48, 1 << 20
    rs = np.random.RandomState(48)
    U = np.linalg.svd(synth.rand64c(rs, k, 64).astype(np.complex128), full_matrices=False)[0][:, :m]
    M = np.asfortranarray(U.conj().T.astype(C64))
    x = synth.rand64c(rs, k, n)
    Md, xd = B.copy_array(M), B.copy_array(x)
    y1, y2 = B.zero_array((m, n), C64), B.zero_array((m, n), C64)
    B.cgemm(y1, Md, xd, 1.0, 0.0, forward=True)
    simt_gemm(True); B.cgemm(y2, Md, xd, 1.0, 0.0, forward=True); simt_gemm(False)
    ref2 = B.norm2(y2)
    B.axpby(1.0, y2, -1.0, y1)
    assert np.sqrt(B.norm2(y2) / ref2) < 2e-6
    cols = slice(n - 4099, n)
    got = y1.to_host()[:, cols]
    assert relerr(got, M.astype(np.complex128) @ x[:, cols].astype(np.complex128)) < 2e-6
    # P = M^H M is a projector: M (M^H (M x)) = M x
    back, again = B.zero_array((k, n), C64), B.zero_array((m, n), C64)
    B.cgemm(back, Md, y1, 1.0, 0.0, forward=False)
    B.cgemm(again, Md, back, 1.0, 0.0, forward=True)
    ref1 = B.norm2(y1)
    B.axpby(1.0, again, -1.0, y1)
    assert np.sqrt(B.norm2(again) / ref1) < 5e-6


@pytest.mark.parametrize("m,k,alpha,beta,left", list(product([2, 5, 6, 40], [1, 3], [0.0, 1.5], [0.0, 0.5], [True, False])))
def test_csymm(B, m, k, alpha, beta, left):
    rs = np.random.RandomState(m + k)
    S = synth.rand64c(rs, m, m); S = np.asfortranarray(S + S.T); S.imag = 0
    x, y = (synth.rand64c(rs, m, k), synth.rand64c(rs, m, k)) if left else (synth.rand64c(rs, k, m), synth.rand64c(rs, k, m))
    want = alpha * (S @ x) + beta * y if left else alpha * (x @ S) + beta * y
    yd = B.copy_array(y)
    B.csymm(yd, B.copy_array(S), B.copy_array(x), alpha, beta, left)
    np.testing.assert_allclose(yd.to_host(), want, atol=1e-4)


def test_onemm_dia_max_golden(B, golden_dir):
    g = np.load(os.path.join(golden_dir, "primitives.npz"))
    yd = B.copy_array(g["one_y"]); B.onemm(yd, B.copy_array(g["one_x"]), 1.5 - 1j, 0.5)
    assert relerr(yd.to_host(), g["one_out"]) < 1e-6
    D = spp.dia_matrix((g["dia_data"], g["dia_offsets"]), shape=tuple(g["dia_shape"]))
    Dd = B.dia_matrix(B, D)
    yd = B.copy_array(g["dia_y"]); Dd.forward(yd, B.copy_array(g["dia_x"]), alpha=1.5, beta=0.5)
    assert relerr(yd.to_host(), g["dia_fwd"]) < 1e-6
    xd = B.copy_array(g["dia_x"]); Dd.adjoint(xd, B.copy_array(g["dia_y"]), alpha=0.5, beta=1.5)
    assert relerr(xd.to_host(), g["dia_adj"]) < 1e-6
    ad = B.copy_array(g["max_in"]); B.max(0.1, ad)
    np.testing.assert_array_equal(ad.to_host(), g["max_out"])
    yd = B.copy_array(g["gemm_y"]); B.cgemm(yd, B.copy_array(g["gemm_M"]), B.copy_array(g["gemm_x"]), 0.5 + 0.5j, 0.5, True)
    assert relerr(yd.to_host(), g["gemm_fwd"]) < 1e-6
    xd = B.copy_array(g["b1_x"]); yd = B.copy_array(g["b1_y"]); B.axpby(0.5 + 1.5j, yd, -2.1 + 3j, xd)
    assert relerr(yd.to_host(), g["b1_axpby"]) < 1e-6
    assert abs(B.dot(xd, B.copy_array(g["b1_y"])) - float(g["b1_dot"])) < 1e-4
    assert abs(B.norm2(xd) - float(g["b1_nrm2"])) < 1e-4


@pytest.mark.parametrize("M,Kd,N,alpha,beta,noff", list(product([23, 45], [45, 23], [1, 9], [0.5, 1.5], [0.0, 1.5], [1, 4])))
def test_dia_matrix(B, M, Kd, N, alpha, beta, noff):
    rs = np.random.RandomState(M * Kd + N + noff)
    offs = np.array(sorted(set(rs.randint(-Kd, M + Kd, size=noff))))
    data = (rs.rand(offs.size, Kd) + 1j * rs.rand(offs.size, Kd)).astype(C64)
    A = spp.dia_matrix((data, offs), shape=(M, Kd))
    Ad = B.dia_matrix(B, A)
    x, y = synth.rand64c(rs, Kd, N), synth.rand64c(rs, M, N)
    yd = B.copy_array(y); Ad.forward(yd, B.copy_array(x), alpha=alpha, beta=beta)
    np.testing.assert_allclose(yd.to_host(), beta * y + alpha * (A @ x), atol=1e-5)
    xd = B.copy_array(x); Ad.adjoint(xd, B.copy_array(y), alpha=alpha, beta=beta)
    np.testing.assert_allclose(xd.to_host(), beta * x + alpha * (A.conj().T @ y), atol=1e-5)


@pytest.mark.parametrize("val,N", list(product([-1.5, 0, 0.5, 1.5], [4, 5, 1000])))
def test_max(B, val, N):
    a = synth.rand64c(np.random.RandomState(N), N)
    ad = B.copy_array(a); B.max(val, ad)
    got = ad.to_host()
    np.testing.assert_array_equal(got.real, np.maximum(a.real, np.float32(val)))
    np.testing.assert_array_equal(got.imag, np.maximum(a.imag, np.float32(val)))


def test_only_complex64(B):
    """csr_matrix.forward/adjoint refuse anything but complex64 operands (backend.py:571-573, test_backends.py:480-490)."""
    A0 = _rand_csr(np.random.RandomState(0), 22, 33, 0.5)
    x = B.copy_array(synth.rand64c(np.random.RandomState(0), 33, 4).astype(np.complex128))
    y = B.copy_array(synth.rand64c(np.random.RandomState(0), 22, 4).astype(np.complex128))
    Ad = B.csr_matrix(B, A0)
    with pytest.raises(AssertionError):
        Ad.forward(y, x)
    with pytest.raises(AssertionError):
        Ad.adjoint(x, y)
