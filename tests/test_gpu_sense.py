"""
GPU parity tests of the north-star path: the -O3 SENSE-NUFFT tree -- built by the REFERENCE's own
builders (backend.py:287-448), rewritten by its own Transform machinery and walked by its own
operators.py, on a B200Backend derived from the reference's Backend (tests/refenv.py) -- against
(a) golden vectors from the unmodified reference NumpyBackend and (b) the numpy oracle on the same
seeded inputs; then the same six calls issued directly by the standalone build
(sense_operator_device).  Bar (BASELINE.json): CSR structure bit-identical; rel-L2 <= 1e-5 on
A x, A^H y, A^H A x and CG iterates.
"""
import hashlib
import os

import numpy as np
import pytest

from indigo_b200 import synth
from indigo_b200.sense import sense_operator, normal_operator, sqrt_dcf
from oracle import sense as osense, np_oracle as K
from refenv import b200_reference_backend

pytestmark = pytest.mark.gpu
C64 = np.dtype('complex64')
TOL = 1e-5


@pytest.fixture(scope="module")
def B():
    """B200Backend on the reference's Backend base: trees are the reference's own."""
    return b200_reference_backend(0)


@pytest.fixture(scope="module")
def SB():
    """Standalone build (no reference package involved)."""
    from indigo_b200 import B200Backend
    return B200Backend(0)


def relerr(a, b):
    a = np.asarray(a).ravel(order='F'); b = np.asarray(b).ravel(order='F')
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def leaves(A):
    """Device matrices (G', P^H) of the operator: SpMatrix leaves of a reference tree, or the operands of
    the standalone six-call operator."""
    if hasattr(A, 'G') and hasattr(A, 'P'):
        return A.G, A.P
    found = []

    def walk(n):
        if type(n).__name__ == 'SpMatrix':
            found.append(n)
        for c in getattr(n, '_children', []):
            walk(c)
    walk(A)
    return ([n for n in found if 'interp' in n._name][0]._get_or_create_device_matrix(),
            [n for n in found if 'zpad' in n._name][0]._get_or_create_device_matrix())


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_sense_against_reference_golden(B, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    N = tuple(int(v) for v in g["N"])
    A = sense_operator(B, N, g["coord"], g["maps"], float(g["oversamp"]))
    Gd, Pd = leaves(A)
    for d, tag in ((Gd, "G"), (Pd, "P")):          # device-resident CSR, read back: bit-identical
        np.testing.assert_array_equal(d.rowPtrs.to_host(), g[tag + "_indptr"])
        np.testing.assert_array_equal(d.colInds.to_host(), g[tag + "_indices"])
        np.testing.assert_array_equal(d.values.to_host(), g[tag + "_data"])
        assert d._exwrite == int(g[tag + "_exwrite"])
    assert relerr(A * g["x"], g["Ax"]) < TOL
    assert relerr(A.H * g["y"], g["AHy"]) < TOL
    AHA = normal_operator(A)
    n0 = B._lib.launch_count()
    y_d = B.zero_array((AHA.shape[0], 1), C64); x_d = B.copy_array(g["x"])
    n0 = B._lib.launch_count()
    AHA.eval(y_d, x_d)
    assert relerr(y_d.to_host(), g["AHAx"]) < TOL
    assert 6 <= B._lib.launch_count() - n0 <= 16       # six backend calls, a few kernels each


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_cg_iterates_against_reference_golden(B, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    N = tuple(int(v) for v in g["N"])
    A = sense_operator(B, N, g["coord"], g["maps"], float(g["oversamp"]), weights=g["cg_w"])
    AHA = normal_operator(A)
    its = []
    x = np.zeros_like(g["cg_b"], order='F')
    B.cg(AHA, g["cg_b"], x, lamda=float(g["cg_lamda"]), maxiter=len(g["cg_iterates"]), tol=0.0, iterates=its)
    for k, (mine, ref) in enumerate(zip(its, g["cg_iterates"])):
        assert relerr(mine, ref) < TOL, (k, relerr(mine, ref))
    assert relerr(x, g["cg_iterates"][-1]) < TOL
    # the reference's own Backend.cg (backend.py:639-689, host scalars) through the same kernels agrees too
    x2 = np.zeros_like(g["cg_b"], order='F')
    super(type(B), B).cg(AHA, g["cg_b"], x2, lamda=float(g["cg_lamda"]), maxiter=len(g["cg_iterates"]), tol=0.0)
    assert relerr(x2, g["cg_iterates"][-1]) < TOL


def test_cfg1_full_size(B, golden_dir):
    """Config 1 at full size: 256x256x1 image, 8 coils, 402 spokes x 512 samples, grid 512x512x2."""
    g = np.load(os.path.join(golden_dir, "sense_cfg1_digest.npz"))
    N, C = tuple(int(v) for v in g["N"]), int(g["C"])
    rs = np.random.RandomState(int(g["seed"]))
    maps = synth.unit_rss_maps(rs, N, C)
    coord = synth.radial_2d(402, 512)
    A = sense_operator(B, N, coord, maps, float(g["oversamp"]))
    Gd, Pd = leaves(A)
    assert Gd.nnz == int(g["G_nnz"]) and Pd.nnz == int(g["P_nnz"])
    assert sha(Gd.rowPtrs.to_host()) == str(g["G_indptr_sha"]) and sha(Gd.colInds.to_host()) == str(g["G_indices_sha"])
    assert sha(Pd.rowPtrs.to_host()) == str(g["P_indptr_sha"]) and sha(Pd.colInds.to_host()) == str(g["P_indices_sha"])
    sub = slice(None, None, 997)
    np.testing.assert_array_equal(Gd.values.to_host()[sub], g["G_data_sub"])
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, coord.shape[1] * coord.shape[2] * C, 1)
    Ax = A * x
    assert relerr(Ax.ravel(order='F')[sub], g["Ax_sub"]) < TOL
    assert abs(np.linalg.norm(Ax) / float(g["Ax_norm"]) - 1) < TOL
    AHy = A.H * y
    assert relerr(AHy.ravel(order='F')[sub], g["AHy_sub"]) < TOL
    AHAx = normal_operator(A) * x
    assert relerr(AHAx.ravel(order='F')[sub], g["AHAx_sub"]) < TOL
    assert abs(np.linalg.norm(AHAx) / float(g["AHAx_norm"]) - 1) < TOL
    # adjointness at full size (size-independent property)
    lhs = np.vdot(Ax.astype(np.complex128), y.reshape(Ax.shape, order='F').astype(np.complex128))
    rhs = np.vdot(x.astype(np.complex128), AHy.astype(np.complex128))
    assert abs(lhs - rhs) / abs(lhs) < TOL


def test_reduced_cfg3_against_oracle(B):
    """cfg3 geometry (3-D kooshball, 2x oversampling, 16 coils) at a size the numpy oracle
    finishes in seconds: 26^3 image -> 52^3 grid (52 = 13*4, the radix mix of 416)."""
    rs = np.random.RandomState(33)
    N, C = (26, 26, 26), 16
    coord = synth.kooshball_3d(nspokes=96, nread=52)
    maps = synth.unit_rss_maps(rs, N, C)
    w = sqrt_dcf(coord)
    A = sense_operator(B, N, coord, maps, 2.0, weights=w)
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    Gd, Pd = leaves(A)
    np.testing.assert_array_equal(Gd.colInds.to_host(), ref.G.indices)
    np.testing.assert_array_equal(Gd.rowPtrs.to_host(), ref.G.indptr)
    np.testing.assert_array_equal(Pd.colInds.to_host(), ref.PH.indices)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A * x, ref.forward(x)) < TOL
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    AHA = normal_operator(A)
    want = ref.normal(x)
    assert relerr(AHA * x, want) < TOL
    b = (want / np.abs(want).max()).astype(C64)
    nrm = osense.spectral_norm(ref)

    def run(lam):
        mine, theirs = [], []
        B.cg(AHA, b, np.zeros_like(b, order='F'), lamda=lam, maxiter=50, tol=0.0, iterates=mine)
        K.cg(ref.normal_into, b, np.zeros_like(b), lamda=lam, tol=0.0, maxiter=50, iterates=theirs)
        return mine, theirs

    # (1) well-conditioned protocol: every one of the 50 iterates within 1e-5 of the oracle's.
    # lamda = 0.05*||A^H A||_2 keeps the complex64 oracle itself within 1.3e-6 of the fp64
    # recurrence (measured, DESIGN.md "CG parity protocol").
    mine, theirs = run(0.05 * nrm)
    worst = max(relerr(m, t) for m, t in zip(mine, theirs))
    assert worst < TOL, worst
    # (2) SURVEY 8(d)'s lamda = 1e-2*||A^H A||_2: here the complex64 oracle drifts 2e-4 from the
    # fp64 recurrence mid-run, so 1e-5 iterate-by-iterate is below the oracle's own noise floor;
    # the meaningful bar is "no further from the fp64 truth than the oracle is".
    lam = 1e-2 * nrm
    mine, theirs = run(lam)
    truth = osense.SenseTruth64(ref).cg(b, lam, 50)
    for k, (m, t, tr) in enumerate(zip(mine, theirs, truth)):
        assert relerr(m, tr) <= max(TOL, 3 * relerr(t, tr)), (k, relerr(m, tr), relerr(t, tr))
    assert relerr(mine[-1], truth[-1]) < TOL


@pytest.mark.parametrize("name", ["sense_small", "sense_odd"])
def test_device_built_operator_matches_reference(SB, golden_dir, name):
    B = SB
    """G' and P^H built by CUDA kernels from (coord, maps): CSR structure bit-identical to the
    reference's, values bit-identical (no aliasing taps in these grids), applies within 1e-5."""
    from indigo_b200.sense import sense_operator_device
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    N = tuple(int(v) for v in g["N"])
    A = sense_operator_device(B, N, g["coord"], g["maps"], float(g["oversamp"]))
    Gd, Pd = leaves(A)
    for d, tag in ((Gd, "G"), (Pd, "P")):
        assert d.shape == tuple(g[tag + "_shape"])
        np.testing.assert_array_equal(d.rowPtrs.to_host(), g[tag + "_indptr"])
        np.testing.assert_array_equal(d.colInds.to_host(), g[tag + "_indices"])
        np.testing.assert_array_equal(d.values.to_host(), g[tag + "_data"])
        assert d._exwrite == int(g[tag + "_exwrite"])
    assert relerr(A * g["x"], g["Ax"]) < TOL
    assert relerr(A.H * g["y"], g["AHy"]) < TOL
    assert relerr(normal_operator(A) * g["x"], g["AHAx"]) < TOL
    Aw = sense_operator_device(B, N, g["coord"], g["maps"], float(g["oversamp"]), weights=g["cg_w"])
    x = np.zeros_like(g["cg_b"], order='F')
    B.cg(normal_operator(Aw), g["cg_b"], x, lamda=float(g["cg_lamda"]), maxiter=len(g["cg_iterates"]), tol=0.0)
    assert relerr(x, g["cg_iterates"][-1]) < TOL


def test_device_built_cfg1_structure(SB, golden_dir):
    B = SB
    """Config 1 has a 2-point z axis: 5-6 taps alias onto 2 cells and are merged."""
    from indigo_b200.sense import sense_operator_device
    g = np.load(os.path.join(golden_dir, "sense_cfg1_digest.npz"))
    N, C = tuple(int(v) for v in g["N"]), int(g["C"])
    rs = np.random.RandomState(int(g["seed"]))
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_device(B, N, synth.radial_2d(402, 512), maps, float(g["oversamp"]))
    Gd, Pd = leaves(A)
    assert Gd.nnz == int(g["G_nnz"]) and Pd.nnz == int(g["P_nnz"])
    assert sha(Gd.rowPtrs.to_host()) == str(g["G_indptr_sha"]) and sha(Gd.colInds.to_host()) == str(g["G_indices_sha"])
    assert sha(Pd.rowPtrs.to_host()) == str(g["P_indptr_sha"]) and sha(Pd.colInds.to_host()) == str(g["P_indices_sha"])
    sub = slice(None, None, 997)
    assert relerr(Gd.values.to_host()[sub], g["G_data_sub"]) < 1e-6
    np.testing.assert_array_equal(Pd.values.to_host()[sub], g["P_data_sub"])
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    assert relerr((normal_operator(A) * x).ravel(order='F')[sub], g["AHAx_sub"]) < TOL
