"""
Pins oracle/ against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only.
"""
import hashlib
import os

import numpy as np
import pytest
import scipy.sparse as spp

from oracle import np_oracle as K
from oracle import sense
from indigo_b200 import synth

C64 = np.dtype("complex64")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def relerr(a, b):
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --------------------------------------------------------------------------- primitives
def test_ccsrmm_forward_adjoint_with_leading_dims(golden_dir):
    g = _load(golden_dir, "primitives")
    m, k = g["csr_shape"]
    alpha, beta = complex(g["csr_alpha"]), complex(g["csr_beta"])
    y = g["csr_fwd_ybig"].copy(order="F"); x = g["csr_fwd_xbig"]
    K.ccsrmm(y[2:2 + m, :], (m, k), g["csr_indices"], g["csr_indptr"], g["csr_data"], x[3:3 + k, :], alpha, beta)
    np.testing.assert_array_equal(y, g["csr_fwd_out"])
    y = g["csr_adj_ybig"].copy(order="F"); x = g["csr_adj_xbig"]
    K.ccsrmm(y[3:3 + k, :], (m, k), g["csr_indices"], g["csr_indptr"], g["csr_data"], x[2:2 + m, :], alpha, beta, adjoint=True)
    np.testing.assert_array_equal(y, g["csr_adj_out"])
    A = spp.csr_matrix((g["csr_data"], g["csr_indices"], g["csr_indptr"]), shape=(m, k))
    np.testing.assert_allclose(K.csr_inspect(A), g["csr_inspect"])


def test_exwrite_adjoint(golden_dir):
    g = _load(golden_dir, "primitives")
    shp = tuple(g["exw_shape"])
    y = g["exw_y"].copy(order="F")
    K.ccsrmm(y, shp, g["exw_indices"], g["exw_indptr"], g["exw_data"], g["exw_x"], 0.5, 1.5, adjoint=True, exwrite=True)
    np.testing.assert_array_equal(y, g["exw_adj_out"])
    A = spp.csr_matrix((g["exw_data"], g["exw_indices"], g["exw_indptr"]), shape=shp)
    assert K.csr_inspect(A)[2] == int(g["exw_flag"]) == 1


@pytest.mark.parametrize("tag", ["fft3", "fft2", "fft1", "fft3b"])
def test_fft(golden_dir, tag):
    g = _load(golden_dir, "primitives")
    x = g[tag + "_in"]
    y = np.zeros_like(x, order="F")
    K.fftn(y, x); np.testing.assert_array_equal(y, g[tag + "_fwd"])
    K.ifftn(y, x); np.testing.assert_array_equal(y, g[tag + "_inv"])


def test_blas1_dense_misc(golden_dir):
    g = _load(golden_dir, "primitives")
    x, y = g["b1_x"].copy(), g["b1_y"].copy()
    K.axpby(0.5 + 1.5j, y, -2.1 + 3j, x); np.testing.assert_array_equal(y, g["b1_axpby"])
    assert K.dot(x, g["b1_y"]) == g["b1_dot"]
    assert K.norm2(x) == g["b1_nrm2"]
    K.scale(x, 1.1 - 2j); np.testing.assert_array_equal(x, g["b1_scale"])
    y = g["gemm_y"].copy(order="F")
    K.cgemm(y, g["gemm_M"], g["gemm_x"], 0.5 + 0.5j, 0.5, forward=True); np.testing.assert_array_equal(y, g["gemm_fwd"])
    x = g["gemm_x"].copy(order="F")
    K.cgemm(x, g["gemm_M"], g["gemm_y"], 1.0, 0.5, forward=False); np.testing.assert_array_equal(x, g["gemm_adj"])
    y = g["symm_yl"].copy(order="F"); K.csymm(y, g["symm_M"], g["symm_xl"], 1.5, 0.5, True)
    np.testing.assert_array_equal(y, g["symm_left"])
    y = g["symm_yr"].copy(order="F"); K.csymm(y, g["symm_M"], g["symm_xr"], 1.5, 0.5, False)
    np.testing.assert_array_equal(y, g["symm_right"])
    y = g["one_y"].copy(order="F"); K.onemm(y, g["one_x"], 1.5 - 1j, 0.5); np.testing.assert_array_equal(y, g["one_out"])
    shp = tuple(g["dia_shape"]); dev_data = np.asfortranarray(g["dia_data"].T)
    y = g["dia_y"].copy(order="F"); K.cdiamm(y, shp, g["dia_offsets"], dev_data, g["dia_x"], 1.5, 0.5, adjoint=False)
    np.testing.assert_array_equal(y, g["dia_fwd"])
    x = g["dia_x"].copy(order="F"); K.cdiamm(x, shp, g["dia_offsets"], dev_data, g["dia_y"], 0.5, 1.5, adjoint=True)
    np.testing.assert_array_equal(x, g["dia_adj"])
    a = g["max_in"].copy(); K.fmax(0.1, a); np.testing.assert_array_equal(a, g["max_out"])


# --------------------------------------------------------------------------- C restatement vs numpy oracle vs compiled reference
def test_c_restatement_and_ref_customcpu(golden_dir):
    lib = K.load_oracle_c()
    assert lib is not None, "run `make -C oracle`"
    g = _load(golden_dir, "primitives")
    m, k = (int(v) for v in g["csr_shape"])
    ind, ptr, val = (np.ascontiguousarray(g[n]) for n in ("csr_indices", "csr_indptr", "csr_data"))
    out = np.zeros(3, dtype=np.int64)
    lib.oracle_csr_inspect(m, k, ind.ctypes.data, ptr.ctypes.data, out.ctypes.data)
    A = spp.csr_matrix((val, ind, ptr), shape=(m, k))
    rf, cf, exw = K.csr_inspect(A)
    assert (out[0] / m, out[1] / k, out[2]) == (rf, cf, exw)
    ref = K.load_ref_customcpu()
    if ref is not None:
        assert tuple(ref.inspect(m, k, ind, ptr)) == tuple(int(v) for v in out)
    alpha, beta = complex(g["csr_alpha"]), complex(g["csr_beta"])
    for adjoint in (False, True):
        xb = g["csr_adj_xbig" if adjoint else "csr_fwd_xbig"]; yb = g["csr_adj_ybig" if adjoint else "csr_fwd_ybig"]
        xo, yo = (2, 3) if adjoint else (3, 2)
        y = yb.copy(order="F"); x = np.asfortranarray(xb)
        n = x.shape[1]
        lib.oracle_ccsrmm(int(adjoint), m, n, k, alpha.real, alpha.imag, val.ctypes.data, ind.ctypes.data, ptr.ctypes.data,
                          x.ctypes.data + 8 * xo, x.shape[0], beta.real, beta.imag, y.ctypes.data + 8 * yo, y.shape[0])
        want = g["csr_adj_out" if adjoint else "csr_fwd_out"]
        assert relerr(y, want) < 2e-7
        if ref is not None and n > 1:
            # the reference's own OpenMP kernel (_customcpu.c:14-114); n==1 forward would need MKL
            yr = yb.copy(order="F")
            xs = np.asfortranarray(x[xo:xo + (m if adjoint else k), :]); ys = np.asfortranarray(yr[yo:yo + (k if adjoint else m), :])
            ref.csrmm(adjoint, m, n, k, alpha, val, ind, ptr, xs, xs.shape[0], beta, ys, ys.shape[0], False)
            assert relerr(ys, want[yo:yo + ys.shape[0], :]) < 2e-7


# --------------------------------------------------------------------------- SENSE operator construction
@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_sense_matrices_bit_identical(golden_dir, name):
    g = _load(golden_dir, name)
    op = sense.SenseOperator(tuple(g["N"]), g["coord"], g["maps"], float(g["oversamp"]))
    for ours, tag in ((op.G, "G"), (op.PH, "P")):
        assert ours.shape == tuple(g[tag + "_shape"])
        np.testing.assert_array_equal(ours.indptr, g[tag + "_indptr"])
        np.testing.assert_array_equal(ours.indices, g[tag + "_indices"])
        np.testing.assert_array_equal(ours.data, g[tag + "_data"])
        assert ours.indices.dtype == np.int32
    assert K.csr_inspect(op.G)[2] == int(g["G_exwrite"])
    assert K.csr_inspect(op.PH)[2] == int(g["P_exwrite"])


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_sense_applies(golden_dir, name):
    g = _load(golden_dir, name)
    op = sense.SenseOperator(tuple(g["N"]), g["coord"], g["maps"], float(g["oversamp"]))
    assert relerr(op.forward(g["x"]), g["Ax"].reshape(op.M, op.C, order="F")) < 1e-6
    assert relerr(op.adjoint(g["y"]), g["AHy"]) < 1e-6
    assert relerr(op.normal(g["x"]), g["AHAx"]) < 1e-6


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_sense_cg_iterates(golden_dir, name):
    g = _load(golden_dir, name)
    op = sense.SenseOperator(tuple(g["N"]), g["coord"], g["maps"], float(g["oversamp"]), weights=g["cg_w"])
    np.testing.assert_array_equal(sense.sqrt_dcf(g["coord"]), g["cg_w"])
    its = []
    K.cg(op.normal_into, g["cg_b"], np.zeros_like(g["cg_b"]), lamda=float(g["cg_lamda"]),
         tol=0.0, maxiter=len(g["cg_iterates"]), iterates=its)
    for k, (mine, ref) in enumerate(zip(its, g["cg_iterates"])):
        assert relerr(mine, ref) < 1e-5, (k, relerr(mine, ref))


def test_sense_cfg1_structure_digest(golden_dir):
    """Full config 1 (256x256x1, 8 coils, 402 spokes x 512): CSR structure by hash."""
    g = _load(golden_dir, "sense_cfg1_digest")
    N, C = tuple(int(v) for v in g["N"]), int(g["C"])
    rs = np.random.RandomState(int(g["seed"]))
    maps = synth.unit_rss_maps(rs, N, C)
    op = sense.SenseOperator(N, synth.radial_2d(402, 512), maps, float(g["oversamp"]))
    assert op.G.nnz == int(g["G_nnz"]) and op.PH.nnz == int(g["P_nnz"])
    assert sha(op.G.indptr) == str(g["G_indptr_sha"]) and sha(op.G.indices) == str(g["G_indices_sha"])
    assert sha(op.PH.indptr) == str(g["P_indptr_sha"]) and sha(op.PH.indices) == str(g["P_indices_sha"])
    sub = slice(None, None, 997)
    np.testing.assert_array_equal(op.G.data[sub], g["G_data_sub"])
    np.testing.assert_array_equal(op.PH.data[sub], g["P_data_sub"])
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, op.M * C, 1)
    Ax = op.forward(x)
    assert relerr(Ax.ravel(order="F")[sub], g["Ax_sub"]) < 1e-6
    assert abs(np.linalg.norm(Ax) / float(g["Ax_norm"]) - 1) < 1e-6
    assert relerr(op.adjoint(y).ravel(order="F")[sub], g["AHy_sub"]) < 1e-6


def test_direct64_matches_the_oracle():
    """oracle/direct64.py (float64 single-output evaluation used for the full-size GPU checks) against the pinned
    oracle operator: A x at chosen samples, A^H y at chosen voxels for sparse y; with and without row weights,
    including samples that sit exactly on a grid line (six taps) and wrap around the grid edge."""
    from oracle import direct64
    from indigo_b200 import synth
    rs = np.random.RandomState(17)
    N, C = (12, 10, 8), 3
    coord = synth.random_3d(rs, 120)[:, :, 0]
    coord[:, 0] = (0.0, 0.25, -0.5)                       # on-grid in every axis, and the wrapped lower edge
    coord[:, 1] = (0.499, -0.499, 0.0)
    maps = synth.unit_rss_maps(rs, N, C)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    for w in (None, (0.5 + rs.rand(coord.shape[1])).astype(np.float32)):
        op = sense.SenseOperator(N, coord, maps, 2.0, weights=w)
        pick = np.array([0, 1, 5, 17, 63, 119])
        want = op.forward(x).reshape((op.M, C), order='F')[pick]
        got = direct64.forward_at_samples(N, coord[:, pick], maps, x, 2.0, weights=None if w is None else w[pick])
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6
        y = np.zeros((op.M, C), dtype=np.complex64, order='F')
        y[pick] = synth.rand64c(rs, pick.size, C)
        full = op.adjoint(np.asfortranarray(y.reshape((-1, 1), order='F'))).reshape(N, order='F')
        vox = np.array([[0, 0, 0], [11, 9, 7], [6, 5, 4], [3, 8, 1], [10, 0, 6]])
        got = direct64.adjoint_at_voxels(N, coord[:, pick], y[pick], maps, vox, 2.0, weights=None if w is None else w[pick])
        want = full[vox[:, 0], vox[:, 1], vox[:, 2]]
        assert np.linalg.norm(got - want) / np.linalg.norm(want) < 2e-6
