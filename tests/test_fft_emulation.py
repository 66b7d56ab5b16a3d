"""
The FFT tile logic and planner that the device kernels are compiled from
(indigo_b200/csrc/fft_core.cuh, fft_plan.hpp) run here on the CPU through the
emulation harness tests/csrc/fft_emul.cu -- same index arithmetic, butterflies,
stage sequencing and diagonal fusion -- against numpy.  CPU only.
"""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "csrc", "libfft_emul.so")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build_test_helpers()
    lib = ctypes.CDLL(SO)
    lib.emul_last_error.restype = ctypes.c_char_p

    def run(x, direction=-1, din=None, dout=None, cin=0, cout=0):
        dims = np.array(x.shape[:-1], dtype=np.int64)
        xf = np.asfortranarray(x.astype(np.complex64))
        y = np.zeros_like(xf, order='F')
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
        rc = lib.emul_fft(len(dims), p(dims), ctypes.c_int64(x.shape[-1]), p(y), p(xf), direction,
                          p(din), cin, p(dout), cout, None)
        assert rc == 0, lib.emul_last_error()
        return y
    return run


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


SHAPES = [(2, 3), (3, 2), (5, 2), (7, 2), (11, 1), (13, 2), (16, 2), (22, 4), (23, 2), (24, 3), (25, 2), (128, 2),
          (416, 2), (512, 1), (17 * 4, 2), (289, 1), (24, 22, 3), (23, 24, 25, 2), (16, 13, 7, 3), (416, 4, 3, 2),
          (3, 5, 416, 1), (1, 24, 2), (24, 1, 2), (1, 1, 3), (34, 38, 2, 3),
          # strided axes with whole 16-line tiles and a specialised length: packed two-line passes (fft_pk.cuh)
          (32, 64, 2), (16, 416, 1), (32, 32, 52, 2), (16, 208, 32, 1), (48, 128, 2)]


@pytest.mark.parametrize("shape", SHAPES)
def test_forward_and_unscaled_inverse(emul, shape):
    rs = np.random.RandomState(sum(shape))
    x = (rs.rand(*shape) + 1j * rs.rand(*shape)).astype(np.complex64)
    ax = tuple(range(len(shape) - 1))
    x128 = x.astype(np.complex128)
    assert rel(emul(x, -1), np.fft.fftn(x128, axes=ax)) < 3e-7
    assert rel(emul(x, +1), np.fft.ifftn(x128, axes=ax) * np.prod(shape[:-1])) < 3e-7


def test_fused_diagonals(emul):
    rs = np.random.RandomState(5)
    shp = (24, 22, 6, 3)
    x = (rs.rand(*shp) + 1j * rs.rand(*shp)).astype(np.complex64)
    d1 = np.asfortranarray((rs.rand(*shp[:-1]) + 1j * rs.rand(*shp[:-1])).astype(np.complex64))
    d2 = np.asfortranarray((rs.rand(*shp[:-1]) + 1j * rs.rand(*shp[:-1])).astype(np.complex64))
    x128 = x.astype(np.complex128)
    want = np.conj(d2)[..., None] * np.fft.fftn(d1[..., None] * x128, axes=(0, 1, 2))
    assert rel(emul(x, -1, d1, d2, 0, 1), want) < 3e-7
    want = d2[..., None] * np.fft.ifftn(np.conj(d1)[..., None] * x128, axes=(0, 1, 2)) * np.prod(shp[:-1])
    assert rel(emul(x, +1, d1, d2, 1, 0), want) < 3e-7


# --------------------------------------------------------------------------- fused SENSE transforms
@pytest.fixture(scope="module")
def emul_sense():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build_test_helpers()
    lib = ctypes.CDLL(SO)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    i64 = lambda v: np.array(v, dtype=np.int64)

    def run(which, N, oN, C, grid, img, pf, alpha=1.0, beta=0.0):
        a, b = complex(alpha), complex(beta)
        n, on = i64(N), i64(oN)
        lib.emul_pk_used()
        rc = lib.emul_sense(p(n), p(on), ctypes.c_int64(C), which, p(grid), p(img), p(pf),
                            ctypes.c_float(a.real), ctypes.c_float(a.imag), ctypes.c_float(b.real), ctypes.c_float(b.imag))
        assert rc == 0
        # the packed two-lines-per-thread bodies (fft_pk.cuh) serve all three passes when the coil count is
        # even, the strided passes alone otherwise (unless disabled through IB200_FFT_NOPK)
        if os.environ.get("IB200_FFT_NOPK") is None:
            assert lib.emul_pk_used() == (3 if C % 2 == 0 else 2)
    return run


def _crand(rs, *shape):
    return (rs.rand(*shape) + 1j * rs.rand(*shape)).astype(np.complex64)


@pytest.mark.parametrize("N,oN,C", [((16, 16, 16), (32, 32, 32), 16), ((16, 26, 16), (32, 52, 32), 4),
                                    ((13, 20, 16), (32, 32, 52), 20), ((32, 16, 26), (64, 32, 52), 3),
                                    ((16, 13, 8), (32, 32, 32), 2), ((16, 15, 4), (32, 32, 32), 8),
                                    ((16, 16, 5), (32, 32, 32), 1),
                                    # every radix of the specialised passes: (16,8) (3,8,8) (13,16) (5,8,8) (7,8,8) (16,16) (13,8,4)
                                    ((64, 16, 16), (128, 32, 32), 2), ((96, 16, 8), (192, 32, 32), 2),
                                    ((16, 104, 8), (32, 208, 32), 2), ((16, 8, 160), (32, 32, 320), 2),
                                    ((224, 16, 8), (448, 32, 32), 2), ((16, 128, 8), (32, 256, 32), 4),
                                    ((8, 16, 208), (32, 32, 416), 6)])
def test_fused_sense_expand_and_combine(emul_sense, N, oN, C):
    """grid[z][y][x][c] = FFT3(zpad(pf*img)) and img = alpha*sum_c conj(pf)*crop(IFFT3_unscaled(grid)) + beta*img,
    against numpy on the same data (zero-pad placement of Backend.Zpad 'center', backend.py:371-387)."""
    rs = np.random.RandomState(sum(N) + C)
    img = _crand(rs, *N)                                   # indexed [x, y, z]
    pf = _crand(rs, *N, C)                                 # [x, y, z, c]
    off = [o // 2 - n // 2 for n, o in zip(N, oN)]
    sl = tuple(slice(o, o + n) for o, n in zip(off, N))
    pad = np.zeros(tuple(oN) + (C,), dtype=np.complex128)
    pad[sl] = (pf * img[..., None]).astype(np.complex128)
    want = np.fft.fftn(pad, axes=(0, 1, 2))                # [x, y, z, c]
    # device layouts: image x fastest; pf [voxel][coil]; grid [z][y][x][c]
    img_d = np.ascontiguousarray(img.transpose(2, 1, 0))
    pf_d = np.ascontiguousarray(pf.transpose(2, 1, 0, 3))
    grid = np.full((oN[2], oN[1], oN[0], C), np.nan + 0j, dtype=np.complex64)      # every point must be written
    emul_sense(0, N, oN, C, grid, img_d, pf_d)
    got = grid.transpose(2, 1, 0, 3)
    assert rel(got, want) < 5e-7
    # inverse + combine on an arbitrary grid
    g = _crand(rs, *oN, C)
    grid = np.ascontiguousarray(g.transpose(2, 1, 0, 3))
    inv = np.fft.ifftn(g.astype(np.complex128), axes=(0, 1, 2)) * np.prod(oN)
    y0 = _crand(rs, *N)
    alpha, beta = 0.5 - 0.25j, 1.5 + 0.5j
    want = alpha * (np.conj(pf) * inv[sl]).sum(axis=3) + beta * y0
    y_d = np.ascontiguousarray(y0.transpose(2, 1, 0))
    emul_sense(1, N, oN, C, grid, y_d, pf_d, alpha, beta)
    assert rel(y_d.transpose(2, 1, 0), want) < 5e-7
    y_d = np.full(N[::-1], np.nan + 0j, dtype=np.complex64)                        # beta == 0 never reads img
    grid = np.ascontiguousarray(g.transpose(2, 1, 0, 3))
    emul_sense(1, N, oN, C, grid, y_d, pf_d, 1.0, 0.0)
    assert rel(y_d.transpose(2, 1, 0), (np.conj(pf) * inv[sl]).sum(axis=3)) < 5e-7


@pytest.mark.parametrize("N,oN,C,bx", [((16, 16, 16), (32, 32, 32), 16, 4), ((16, 26, 16), (32, 52, 32), 4, 4),
                                       ((16, 16, 13), (32, 32, 52), 2, 8), ((16, 16, 16), (32, 32, 32), 32, 4)])
def test_fused_sense_support_windows(emul_sense, N, oN, C, bx):
    """Per-column k-space support windows of the z passes (csrc/fft_pk.cuh): the forward transform writes the
    grid only inside the windows (and there it equals the full transform), the inverse transform reads it only
    inside them (NaNs outside must not matter) and treats the rest as zero."""
    if os.environ.get("IB200_FFT_NOPK") or os.environ.get("IB200_FFT_NOPKP"):
        pytest.skip("windows need the persistent packed passes")
    lib = ctypes.CDLL(SO)
    rs = np.random.RandomState(3 + C)
    nbx, nby = oN[0] // bx, oN[1] // 4
    lo = rs.randint(0, oN[2] // 2, size=(nby, nbx)) // 4 * 4
    hi = np.minimum(oN[2], lo + rs.randint(0, oN[2], size=(nby, nbx)) // 4 * 4)
    hi[0, 0] = lo[0, 0]                                     # at least one empty block
    win = np.zeros((oN[1], oN[0], 2), dtype=np.int32)       # [y][x] -> (lo, hi)
    for y in range(oN[1]):
        for x in range(oN[0]):
            l, h = lo[y // 4, x // bx], hi[y // 4, x // bx]
            win[y, x] = (l, h) if h > l else (0, 0)
    inside = np.zeros(tuple(oN), dtype=bool)                # [x, y, z]
    for y in range(oN[1]):
        for x in range(oN[0]):
            inside[x, y, win[y, x, 0]:win[y, x, 1]] = True
    img = _crand(rs, *N); pf = _crand(rs, *N, C)
    off = [o // 2 - n // 2 for n, o in zip(N, oN)]
    sl = tuple(slice(o, o + n) for o, n in zip(off, N))
    pad = np.zeros(tuple(oN) + (C,), dtype=np.complex128)
    pad[sl] = (pf * img[..., None]).astype(np.complex128)
    want = np.fft.fftn(pad, axes=(0, 1, 2))
    img_d = np.ascontiguousarray(img.transpose(2, 1, 0)); pf_d = np.ascontiguousarray(pf.transpose(2, 1, 0, 3))
    lib.emul_sense_set_windows(win.ctypes.data_as(ctypes.c_void_p), C)
    try:
        grid = np.full((oN[2], oN[1], oN[0], C), np.nan + 0j, dtype=np.complex64)
        emul_sense(0, N, oN, C, grid, img_d, pf_d)
        got = grid.transpose(2, 1, 0, 3)
        assert rel(got[inside], want[inside]) < 5e-7
        # outside the windows nothing of the LAST pass was written: planes the earlier passes never touch stay NaN
        zout = np.ones(oN[2], dtype=bool); zout[off[2]:off[2] + N[2]] = False
        assert np.isnan(got[:, :, zout][~inside[:, :, zout]]).all()
        # inverse: NaN outside the windows must be ignored (treated as zero)
        g = _crand(rs, *oN, C)
        gz = np.where(inside[..., None], g, 0).astype(np.complex128)
        gn = np.where(inside[..., None], g, np.nan + 0j).astype(np.complex64)
        grid = np.ascontiguousarray(gn.transpose(2, 1, 0, 3))
        inv = np.fft.ifftn(gz, axes=(0, 1, 2)) * np.prod(oN)
        y_d = np.full(N[::-1], np.nan + 0j, dtype=np.complex64)
        emul_sense(1, N, oN, C, grid, y_d, pf_d, 1.0, 0.0)
        assert rel(y_d.transpose(2, 1, 0), (np.conj(pf) * inv[sl]).sum(axis=3)) < 5e-7
    finally:
        lib.emul_sense_set_windows(None, 1)
