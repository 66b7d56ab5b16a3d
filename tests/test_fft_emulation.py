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
          (3, 5, 416, 1), (1, 24, 2), (24, 1, 2), (1, 1, 3), (34, 38, 2, 3)]


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
