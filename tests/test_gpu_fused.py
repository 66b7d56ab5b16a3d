"""
GPU parity tests of the fused SENSE-NUFFT path (indigo_b200/fused.py: pruned, coil-interleaved
FFT passes with the coil maps / apodisation / zero-padding folded in, interleaved gridding
gathers in both directions) against the numpy oracle on the same seeded operators, and against
the unfused six-call tree at a size the oracle cannot reach.  Tolerance: rel-L2 <= 1e-5 in
complex64 (BASELINE.json north_star) on A x, A^H y, A^H A x and on every CG iterate.
"""
import numpy as np
import pytest

from indigo_b200 import synth
from indigo_b200.sense import sense_operator_device, normal_operator, sqrt_dcf
from indigo_b200.fused import sense_operator_fused
from oracle import np_oracle as K
from oracle import sense as osense

pytestmark = pytest.mark.gpu
C64 = np.dtype('complex64')
TOL = 1e-5


@pytest.fixture(scope="module")
def B():
    from indigo_b200 import B200Backend
    return B200Backend(0)


def relerr(a, b):
    a = np.asarray(a).ravel(order='F'); b = np.asarray(b).ravel(order='F')
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


CASES = [((16, 16, 16), 4, "random", True), ((16, 26, 16), 3, "random", False), ((26, 26, 26), 16, "koosh", True),
         ((16, 16, 26), 20, "random", True), ((32, 16, 16), 1, "koosh", False),
         ((16, 26, 16), 2, "random", True), ((26, 16, 16), 8, "koosh", False), ((16, 26, 16), 4, "random", False)]


def _setup(N, C, traj, weighted, seed=0):
    rs = np.random.RandomState(seed + C)
    coord = synth.random_3d(rs, 700) if traj == "random" else synth.kooshball_3d(nspokes=96, nread=2 * max(N))
    maps = synth.unit_rss_maps(rs, N, C)
    w = sqrt_dcf(coord) if weighted else None
    return rs, coord, maps, w


@pytest.mark.parametrize("mode", ["separable", "separable-runs", "separable-nowindows", "separable-unsorted", "real-packed",
                                  "complex"])
@pytest.mark.parametrize("N,C,traj,weighted", CASES)
def test_fused_against_oracle(B, N, C, traj, weighted, mode, monkeypatch):
    from indigo_b200 import fused
    real = mode != "complex"
    monkeypatch.setattr(fused.SenseDevice, "allow_real", real)
    monkeypatch.setattr(fused.SenseDevice, "allow_separable", mode.startswith("separable"))
    monkeypatch.setattr(fused.SenseDevice, "allow_windows", mode != "separable-nowindows")
    monkeypatch.setattr(fused.SenseDevice, "allow_runs", mode.startswith("separable"))
    monkeypatch.setattr(fused.SenseDevice, "allow_tiles", mode != "separable-runs")
    monkeypatch.setattr(fused.SenseDevice, "allow_sorted_ksp", mode != "separable-unsorted")
    monkeypatch.setattr(fused.SenseDevice, "window_min_saving", 0.0)
    rs, coord, maps, w = _setup(N, C, traj, weighted)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    assert A._dev.real == real
    assert (A._dev.kb is not None) == mode.startswith("separable")
    assert A._dev.ksp_sorted == (mode in ("separable", "separable-runs", "separable-nowindows") and C % 2 == 0)
    if mode.startswith("separable") and C % 2 == 0:
        assert (A._dev.runs is not None) == (mode == "separable-runs") and (A._dev.tiles is None) == (mode == "separable-runs")
    if mode == "separable-nowindows":
        assert A._dev.win is None
    elif C % 2 == 0 and (C % 16 == 0 or 16 % C == 0):
        assert A._dev.win is not None and 0.0 < A._dev.support_fraction <= 1.0
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    nvox = int(np.prod(N))
    x = synth.rand64c(rs, nvox, 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert A.shape == (ref.M * C, nvox)
    assert relerr(A * x, ref.forward(x)) < TOL
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    AHA = normal_operator(A)
    assert relerr(AHA * x, ref.normal(x)) < TOL
    # the generic Product of the two fused halves gives the same answer as the fused normal node
    assert relerr((A.H * A) * x, ref.normal(x)) < TOL
    # alpha / beta semantics of Operator.eval, and beta == 0 must never read y
    y0 = synth.rand64c(rs, nvox, 1)
    yd = B.copy_array(y0)
    AHA.eval(yd, B.copy_array(x), alpha=0.5 - 1j, beta=1.5)
    assert relerr(yd.to_host(), (0.5 - 1j) * ref.normal(x) + 1.5 * y0) < TOL
    yd = B.copy_array(np.full((nvox, 1), np.nan, dtype=C64, order='F'))
    AHA.eval(yd, B.copy_array(x))
    assert relerr(yd.to_host(), ref.normal(x)) < TOL
    kd = B.copy_array(np.full((ref.M * C, 1), np.nan, dtype=C64, order='F'))
    A.eval(kd, B.copy_array(x), alpha=2.0)
    assert relerr(kd.to_host(), 2.0 * ref.forward(x)) < TOL


@pytest.mark.parametrize("runs", [True, False], ids=["x-runs", "rows"])
@pytest.mark.parametrize("C", [16, 4, 6])
def test_fused_long_rows(B, C, runs, monkeypatch):
    """Rows / runs of the stored adjoint above the length threshold go to the one-CTA-per-row kernel (the
    k-space centre of a radial trajectory): thresholds forced low so that small problems exercise it."""
    from indigo_b200 import fused
    monkeypatch.setattr(fused.SenseDevice, "allow_runs", runs)
    monkeypatch.setattr(fused.SenseDevice, "allow_tiles", False)
    monkeypatch.setattr(fused.SenseDevice, "long_thresh", 6)
    monkeypatch.setattr(fused.SenseDevice, "run_long_thresh", 24)
    monkeypatch.setattr(fused.SenseDevice, "window_min_saving", 0.0)
    N = (16, 16, 16)
    rs = np.random.RandomState(11 + C)
    coord = synth.kooshball_3d(nspokes=128, nread=32)
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_fused(B, N, coord, maps, 2.0)
    d = A._dev
    assert (d.runs is not None) == runs
    assert (d.runs['nseg'] if runs else d.nlong) > 0
    ref = osense.SenseOperator(N, coord, maps, 2.0)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL


@pytest.mark.parametrize("shape,lanes", [((4, 4), 0), ((4, 4), 4), ((2, 2), 0), ((2, 2), 2), ((2, 1), 0), ((2, 1), 2),
                                         ((1, 1), 0)], ids=lambda v: "x".join(str(i) for i in v) if isinstance(v, tuple) else "l%d" % v)
@pytest.mark.parametrize("segb", [64, 2], ids=["whole-blocks", "split-blocks"])
@pytest.mark.parametrize("N,C,traj,weighted", [((16, 16, 16), 2, "koosh", False), ((16, 26, 16), 4, "random", True),
                                               ((26, 16, 16), 6, "koosh", True), ((16, 16, 26), 8, "koosh", False),
                                               ((16, 16, 16), 16, "koosh", True), ((16, 16, 16), 20, "koosh", False)])
def test_fused_block_gather(B, N, C, traj, weighted, segb, shape, lanes, monkeypatch):
    """Matrix-free adjoint gridding on block entries (csrc/kbblocks.cu): every block shape and lane geometry, whole
    blocks and blocks cut into work items with the ordered fold (segment length forced low so that the dense k-space
    centre of a small kooshball splits), coil counts that need one and two chunks of 16 columns."""
    from indigo_b200 import fused
    monkeypatch.setattr(fused.SenseDevice, "block_shape", shape)
    monkeypatch.setattr(fused.SenseDevice, "tiles_seg_batches", segb)
    monkeypatch.setattr(fused.SenseDevice, "tiles_lanes", lanes)
    monkeypatch.setattr(fused.SenseDevice, "window_min_saving", 0.0)
    rs, coord, maps, w = _setup(N, C, traj, weighted)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    d = A._dev
    assert d.tiles is not None and d.tiles['shape'] == shape and d.runs is None and d.kb is not None and d.ksp_sorted
    if segb == 2 and traj == "koosh":
        assert d.tiles['nsplit'] > 0
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL
    # same sums as the x-run formulation on the same operator
    monkeypatch.setattr(fused.SenseDevice, "allow_tiles", False)
    A2 = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    assert A2._dev.tiles is None and A2._dev.runs is not None
    assert relerr(A.H * y, A2.H * y) < 2e-6
    # deterministic: two applies give identical bits
    np.testing.assert_array_equal(A.H * y, A.H * y)


@pytest.mark.parametrize("N,C,traj,weighted", [((16, 16, 16), 2, "koosh", False), ((26, 26, 26), 16, "koosh", True),
                                               ((16, 26, 16), 4, "random", True), ((16, 16, 1), 8, "radial2d", False)])
def test_matrix_free_setup_matches_stored(B, N, C, traj, weighted, monkeypatch):
    """The matrix-free construction (sample order from the coordinates, support windows and block entries from the
    separable records: no CSR matrix, no stored adjoint) gives the same windows, row map and operator as the
    construction on stored matrices."""
    from indigo_b200 import fused
    monkeypatch.setattr(fused.SenseDevice, "window_min_saving", 0.0)
    if traj == "radial2d":
        rs = np.random.RandomState(7)
        coord, w = synth.radial_2d(nspokes=24, nread=32), None
        maps = synth.unit_rss_maps(rs, N, C)
    else:
        rs, coord, maps, w = _setup(N, C, traj, weighted)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    d = A._dev
    assert d.G is None and d.t_ptr is None and d.tiles is not None and d.kb is not None
    monkeypatch.setattr(fused.SenseDevice, "matrix_free_setup", False)
    A2 = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    d2 = A2._dev
    assert d2.tiles is not None and d2.nnz == d.nnz and d2.M == d.M
    assert (d.win is None) == (d2.win is None)
    if d.win is not None:
        np.testing.assert_array_equal(d.win.to_host(), d2.win.to_host())
        assert d.support_fraction == d2.support_fraction
    np.testing.assert_array_equal(d.rowmap.to_host(), d2.rowmap.to_host())
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A * x, ref.forward(x)) < TOL and relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(A * x, A2 * x) < 2e-6 and relerr(A.H * y, A2.H * y) < 2e-6
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL


def test_fused_cg_iterates(B):
    """50 CG iterates of the well-conditioned protocol (sqrt-DCF rows, lamda = 0.05 ||A^H A||) on the
    reduced cfg3 geometry, fused operator vs the numpy oracle."""
    N, C = (26, 26, 26), 16
    rs, coord, maps, w = _setup(N, C, "koosh", True, seed=3)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    AHA = normal_operator(A)
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    want = ref.normal(x)
    b = (want / np.abs(want).max()).astype(C64)
    lam = 0.05 * osense.spectral_norm(ref)
    mine, theirs = [], []
    B.cg(AHA, b, np.zeros_like(b, order='F'), lamda=lam, maxiter=50, tol=0.0, iterates=mine)
    K.cg(ref.normal_into, b, np.zeros_like(b), lamda=lam, tol=0.0, maxiter=50, iterates=theirs)
    worst = max(relerr(m, t) for m, t in zip(mine, theirs))
    assert worst < TOL, worst


def test_cg_graph_replay_matches_eager(B):
    """Whole-iteration CUDA graph (apply + fused solver kernels, one launch per iteration) gives the iterates of the
    eager loop bit for bit: same kernels, same order, deterministic reductions."""
    N, C = (16, 16, 16), 4
    rs, coord, maps, w = _setup(N, C, "koosh", True, seed=9)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    AHA = normal_operator(A)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    b = AHA * x
    b = (b / np.abs(b).max()).astype(C64)
    xe, xg = np.zeros_like(b, order='F'), np.zeros_like(b, order='F')
    B.cg(AHA, b, xe, lamda=0.3, maxiter=12, tol=0.0)
    n0 = B._lib.launch_count()
    B.cg(AHA, b, xg, lamda=0.3, maxiter=12, tol=0.0, graph=True)
    np.testing.assert_array_equal(xe, xg)
    assert np.linalg.norm(xe) > 0
    # the six-call recipe (generic ccsrmm / fftn kernels, side-stream fork/join) is capturable too
    Au = sense_operator_device(B, N, coord, maps, 2.0, weights=w)
    AHAu = normal_operator(Au)
    xe2, xg2 = np.zeros_like(b, order='F'), np.zeros_like(b, order='F')
    B.cg(AHAu, b, xe2, lamda=0.3, maxiter=8, tol=0.0)
    B.cg(AHAu, b, xg2, lamda=0.3, maxiter=8, tol=0.0, graph=True)
    np.testing.assert_array_equal(xe2, xg2)


def test_reduced_cfg4_32_coils_cg(B):
    """BASELINE config 4 at reduced size: 32 coils on one GPU (the fused recipe's maximum), sqrt-DCF rows,
    lamda passed through cg(lamda=) (well-conditioned protocol of DESIGN.md section 5, 0.05 ||A^H A||):
    A^H A x and all 50 CG iterates within 1e-5 of the oracle."""
    N, C = (26, 26, 26), 32
    rs, coord, maps, w = _setup(N, C, "koosh", True, seed=5)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    assert A._dev.tiles is not None and A._dev.kb is not None
    AHA = normal_operator(A)
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    want = ref.normal(x)
    assert relerr(AHA * x, want) < TOL
    b = (want / np.abs(want).max()).astype(C64)
    lam = 0.05 * osense.spectral_norm(ref)
    mine, theirs = [], []
    B.cg(AHA, b, np.zeros_like(b, order='F'), lamda=lam, maxiter=50, tol=0.0, iterates=mine)
    K.cg(ref.normal_into, b, np.zeros_like(b), lamda=lam, tol=0.0, maxiter=50, iterates=theirs)
    worst = max(relerr(m, t) for m, t in zip(mine, theirs))
    assert worst < TOL, worst


def test_reduced_cfg5_spirals_with_coil_compression(B):
    """BASELINE config 5 at reduced size: stack-of-spirals trajectory, multi-coil k-space compressed by a
    DenseMatrix (cgemm, coil-fastest data: tall-skinny GEMM) to virtual coils, then the fused NUFFT of the
    virtual coils.  Every stage against the numpy oracle."""
    N, Cv, Cfull = (16, 16, 16), 6, 24
    rs = np.random.RandomState(21)
    coord = synth.stack_of_spirals(nz=16, nleaves=6, nread=96, turns=4.0)
    ns = int(np.prod(coord.shape[1:]))
    # coil compression matrix: first Cv left singular vectors of a calibration block
    calib = synth.rand64c(rs, Cfull, 64)
    U = np.linalg.svd(calib.astype(np.complex128), full_matrices=False)[0][:, :Cv]
    Mc = np.asfortranarray(U.conj().T.astype(C64))                       # Cv x Cfull
    ksp = synth.rand64c(rs, Cfull, ns)                                   # coil-fastest k-space
    yd = B.zero_array((Cv, ns), C64)
    B.cgemm(yd, B.copy_array(Mc), B.copy_array(ksp), 1.0, 0.0, forward=True)
    comp = yd.to_host()
    assert relerr(comp, Mc.astype(np.complex128) @ ksp.astype(np.complex128)) < TOL
    # virtual-coil operator: maps compressed the same way, unit RSS
    maps = synth.unit_rss_maps(rs, N, Cv)
    A = sense_operator_fused(B, N, coord, maps, 2.0)
    ref = osense.SenseOperator(N, coord, maps, 2.0)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    assert relerr(A * x, ref.forward(x)) < TOL
    # adjoint applied to the compressed data (sample-fastest per coil, as Operator.eval expects)
    y = np.asfortranarray(comp.T.reshape((-1, 1), order='F'))
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL


def test_fused_matches_six_call_tree_at_size(B):
    """64^3 image, 128^3 grid, 8 coils, 20k samples: beyond the oracle's reach in a test, so the fused
    path is compared with the (oracle-checked) six-call tree on the same device-built matrices, and the
    size-independent adjointness property <A x, y> = <x, A^H y> is checked."""
    N, C = (64, 64, 64), 8
    rs = np.random.RandomState(11)
    coord = synth.kooshball_3d(nspokes=160, nread=128)
    maps = synth.unit_rss_maps(rs, N, C)
    Af = sense_operator_fused(B, N, coord, maps, 2.0)
    Au = sense_operator_device(B, N, coord, maps, 2.0)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, Af.shape[0], 1)
    Ax, AHy = Af * x, Af.H * y
    assert relerr(Ax, Au * x) < 2e-6
    assert relerr(AHy, Au.H * y) < 2e-6
    assert relerr(normal_operator(Af) * x, normal_operator(Au) * x) < 2e-6
    lhs = np.vdot(Ax.astype(np.complex128), y.astype(np.complex128))
    rhs = np.vdot(x.astype(np.complex128), AHy.astype(np.complex128))
    assert abs(lhs - rhs) / abs(lhs) < TOL


@pytest.mark.parametrize("C", [8, 2])
def test_fused_two_dimensional_problem(B, C):
    """2-D problems are carried as N0 x N1 x 1 images (the reference's NUFFT is 3-D only, backend.py:404-405): the
    oversampled grid has a z axis of two points, served by the direct tiny-axis pass; the Kaiser-Bessel taps alias
    onto the two planes: the separable records hold the summed weights per plane (the reference's COO -> CSR conversion
    sums the duplicates), so both gridding steps stay matrix-free."""
    N = (16, 16, 1)
    rs = np.random.RandomState(40 + C)
    coord = synth.radial_2d(nspokes=24, nread=32)
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_fused(B, N, coord, maps, 2.0)
    assert A._dev.oN == (32, 32, 2) and A._dev.kb is not None and A._dev.tiles is not None
    ref = osense.SenseOperator(N, coord, maps, 2.0)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A * x, ref.forward(x)) < TOL
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL


def test_cfg1_full_size_fused(B, golden_dir):
    """BASELINE configs[0] at full size through the fused recipe, against the digest of the unmodified reference's
    NumpyBackend (tests/golden/make_golden.py): 256 x 256 x 1 image, grid 512 x 512 x 2, 8 coils, 402 spokes x 512."""
    import os
    g = np.load(os.path.join(golden_dir, "sense_cfg1_digest.npz"))
    N, C = tuple(int(v) for v in g["N"]), int(g["C"])
    rs = np.random.RandomState(int(g["seed"]))
    maps = synth.unit_rss_maps(rs, N, C)
    coord = synth.radial_2d(402, 512)
    A = sense_operator_fused(B, N, coord, maps, float(g["oversamp"]))
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, coord.shape[1] * coord.shape[2] * C, 1)
    sub = slice(None, None, 997)
    Ax = A * x
    assert relerr(Ax.ravel(order='F')[sub], g["Ax_sub"]) < TOL
    assert abs(np.linalg.norm(Ax) / float(g["Ax_norm"]) - 1) < TOL
    AHy = A.H * y
    assert relerr(AHy.ravel(order='F')[sub], g["AHy_sub"]) < TOL
    AHAx = normal_operator(A) * x
    assert relerr(AHAx.ravel(order='F')[sub], g["AHAx_sub"]) < TOL
    assert abs(np.linalg.norm(AHAx) / float(g["AHAx_norm"]) - 1) < TOL


def test_fused_refuses_unsupported_grid(B):
    N, C = (11, 12, 13), 2                       # 22 x 24 x 26: no specialised passes
    rs, coord, maps, w = _setup(N, C, "random", False)
    with pytest.raises(RuntimeError):
        sense_operator_fused(B, N, coord, maps, 2.0)


def test_fuse_transform_swaps_the_pics_tree():
    """The tree of examples/pics.py:92-95 built by the reference's own builders (B.NUFFT, B.KronI, B.VStack, B.Diag) and
    handed to the fusion Transform comes back as the fused node and matches the oracle; a grid without
    specialised passes keeps its tree and still evaluates (six-call path)."""
    from indigo_b200.fused import fuse_transform
    from indigo_b200.sense import sense_operator
    from refenv import b200_reference_backend
    B = b200_reference_backend(0)
    N, C = (16, 16, 16), 4
    rs, coord, maps, w = _setup(N, C, "koosh", True)
    A = sense_operator(B, N, coord, maps, 2.0, weights=w, recipe=[fuse_transform(B)])
    assert type(A).__name__ == "FusedSenseNUFFT"
    ref = osense.SenseOperator(N, coord, maps, 2.0, weights=w)
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, ref.M * C, 1)
    assert relerr(A * x, ref.forward(x)) < TOL
    assert relerr(A.H * y, ref.adjoint(y)) < TOL
    assert relerr(normal_operator(A) * x, ref.normal(x)) < TOL
    N2 = (11, 12, 13)                                # 22 x 24 x 26 grid: no fused plan
    rs2, coord2, maps2, _ = _setup(N2, 2, "random", False)
    A2 = sense_operator(B, N2, coord2, maps2, 2.0, recipe=[fuse_transform(B)])
    assert type(A2).__name__ == "Product"
    ref2 = osense.SenseOperator(N2, coord2, maps2, 2.0)
    x2 = synth.rand64c(rs2, int(np.prod(N2)), 1)
    assert relerr(A2 * x2, ref2.forward(x2)) < TOL
