"""
The host-side mirror of the reference interface (indigo_b200/host) evaluated by
the numpy oracle must reproduce what the unmodified reference produced
(tests/golden/*.npz): bit-identical CSR structure and values of G' and P^H, the
same six backend calls per A^H A apply, the same applies and CG iterates.
CPU only.
"""
import os

import numpy as np
import pytest
import scipy.sparse as spp

from indigo_b200 import synth
from indigo_b200.sense import sense_operator, normal_operator, sqrt_dcf
from indigo_b200.host import operators as op
from np_host_backend import NpHostBackend

C64 = np.dtype('complex64')


def relerr(a, b):
    a = np.asarray(a).ravel(order='F'); b = np.asarray(b).ravel(order='F')
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def leaves(A):
    found = []

    def walk(n):
        if isinstance(n, op.SpMatrix):
            found.append(n)
        for c in getattr(n, '_children', []):
            walk(c)
    walk(A)
    G = [n for n in found if 'interp' in n._name][0]
    P = [n for n in found if 'zpad' in n._name][0]
    return G, P


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_mirror_builds_reference_matrices(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B = NpHostBackend()
    A = sense_operator(B, tuple(int(v) for v in g["N"]), g["coord"], g["maps"], float(g["oversamp"]))
    G, P = leaves(A)
    assert G._name == 'interp*mod*scale'
    for node, tag in ((G, "G"), (P, "P")):
        d = node._get_or_create_device_matrix()
        assert d.shape == tuple(g[tag + "_shape"])
        np.testing.assert_array_equal(d.rowPtrs._arr, g[tag + "_indptr"])
        np.testing.assert_array_equal(d.colInds._arr, g[tag + "_indices"])
        np.testing.assert_array_equal(d.values._arr, g[tag + "_data"])
        assert d._exwrite == int(g[tag + "_exwrite"])


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_mirror_applies_and_call_sequence(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B = NpHostBackend()
    A = sense_operator(B, tuple(int(v) for v in g["N"]), g["coord"], g["maps"], float(g["oversamp"]))
    B._scratch._arr[...] = 0
    assert relerr(A * g["x"], g["Ax"]) < 1e-6
    assert relerr(A.H * g["y"], g["AHy"]) < 1e-6
    AHA = normal_operator(A)
    B.calls.clear()
    assert relerr(AHA * g["x"], g["AHAx"]) < 1e-6
    seq = [(n, kw.get('adjoint'), kw.get('exwrite')) for n, kw in B.calls]
    # exactly six backend calls
    assert seq == [('ccsrmm', True, 1), ('fftn', None, None), ('ccsrmm', False, 1),
                   ('ccsrmm', True, 0), ('ifftn', None, None), ('ccsrmm', False, 1)]
    C = int(g["C"]); on = int(g["G_shape"][1])
    assert B.calls[2][1]['ldx'] == on and B.calls[2][1]['x'] == (on, C)


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_mirror_cg_iterates(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    B = NpHostBackend()
    np.testing.assert_array_equal(sqrt_dcf(g["coord"]), g["cg_w"])
    A = sense_operator(B, tuple(int(v) for v in g["N"]), g["coord"], g["maps"], float(g["oversamp"]), weights=g["cg_w"])
    B._scratch._arr[...] = 0
    AHA = normal_operator(A)
    for k, ref in enumerate(g["cg_iterates"], start=1):
        x = np.zeros_like(g["cg_b"], order='F')
        B.cg(AHA, g["cg_b"], x, lamda=float(g["cg_lamda"]), maxiter=k, tol=0.0)
        assert relerr(x, ref) < 1e-5, (k, relerr(x, ref))


def test_operator_algebra_small():
    """SpMatrix / Product / Kron / VStack / HStack / BlockDiag / Scale / Sum / One /
    DenseMatrix forward and adjoint against scipy (the checks of reference
    test_operators.py:12-224,471-642 on one seeded instance each)."""
    rs = np.random.RandomState(3)
    B = NpHostBackend()

    def rmat(m, n):
        return (spp.random(m, n, density=0.3, random_state=rs, format='csr')
                + 1j * spp.random(m, n, density=0.3, random_state=rs, format='csr')).astype(C64)

    A0, A1 = rmat(22, 33), rmat(33, 11)
    x = synth.rand64c(rs, 11, 4)
    P = B.SpMatrix(A0) * B.SpMatrix(A1)
    np.testing.assert_allclose(P * x, A0 @ (A1 @ x), rtol=1e-4)
    y = synth.rand64c(rs, 22, 4)
    np.testing.assert_allclose(P.H * y, (A0 @ A1).conj().T @ y, rtol=1e-4)
    K4 = B.KronI(4, B.SpMatrix(A1))
    xx = synth.rand64c(rs, 11 * 4, 2)
    np.testing.assert_allclose(K4 * xx, spp.kron(spp.eye(4), A1) @ xx, rtol=1e-4)
    V = B.VStack([B.SpMatrix(A1), B.SpMatrix(rmat(5, 11))])
    Vm = spp.vstack([c._matrix for c in V.children])
    np.testing.assert_allclose(V * x, Vm @ x, rtol=1e-4)
    yv = synth.rand64c(rs, 38, 3)
    np.testing.assert_allclose(V.H * yv, Vm.conj().T @ yv, rtol=1e-4)
    Hs = B.HStack([B.SpMatrix(A0), B.SpMatrix(rmat(22, 7))])
    Hm = spp.hstack([c._matrix for c in Hs.children])
    xh = synth.rand64c(rs, 40, 3)
    np.testing.assert_allclose(Hs * xh, Hm @ xh, rtol=1e-4)
    np.testing.assert_allclose(Hs.H * y[:, :3], Hm.conj().T @ y[:, :3], rtol=1e-4)
    Bd = B.BlockDiag([B.SpMatrix(A0), B.SpMatrix(A1)])
    Bm = spp.block_diag([A0, A1])
    xb = synth.rand64c(rs, 44, 2)
    np.testing.assert_allclose(Bd * xb, Bm @ xb, rtol=1e-4)
    S = (2 - 1j) * B.SpMatrix(A1) + B.SpMatrix(A1)
    np.testing.assert_allclose(S * x, (3 - 1j) * (A1 @ x), rtol=1e-4)
    ys = synth.rand64c(rs, 33, 1)
    np.testing.assert_allclose(S.H * ys, (3 + 1j) * (A1.conj().T @ ys), rtol=1e-4)
    O = B.One((5, 11))
    np.testing.assert_allclose(O * x, np.ones((5, 11)) @ x, rtol=1e-5)
    D = synth.rand64c(rs, 6, 11)
    Dm = B.DenseMatrix(D)
    np.testing.assert_allclose(Dm * x, D @ x, rtol=1e-4)
    yd = synth.rand64c(rs, 6, 2)
    np.testing.assert_allclose(Dm.H * yd, D.conj().T @ yd, rtol=1e-4)
    # general Kron with a real-symmetric left factor (reference test_operators.py:593-642)
    Ssym = synth.rand64c(rs, 3, 3); Ssym = np.asfortranarray(Ssym + Ssym.T); Ssym.imag = 0
    Kg = B.Kron(B.DenseMatrix(Ssym), B.SpMatrix(A1))
    xk = synth.rand64c(rs, 3 * 11, 1)
    np.testing.assert_allclose(Kg * xk, spp.kron(Ssym, A1) @ xk, rtol=1e-4)
    # FFTc against fftshift(fftn(ifftshift)) (reference test_operators.py:337-367)
    shp = (8, 6, 4)
    Fc = B.FFTc(shp, dtype=C64)
    v = synth.rand64c(rs, *shp)
    got = (Fc * v.reshape(-1, 1, order='F')).reshape(shp, order='F')
    want = np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(v), norm='ortho'))
    np.testing.assert_allclose(got, want, atol=1e-5)
    # memusage is what Optimize uses to size the arena (reference test_analyses.py:12-43)
    assert P.memusage(ncols=4) == A0.data.nbytes + A1.data.nbytes + (22 + 1 + 33 + 1) * 4 + (A0.nnz + A1.nnz) * 4 + 33 * 4 * 8
