"""
The C-ABI library must load without a GPU and export exactly the symbols that
include/indigo_b200.h declares (the drop-in boundary).  Host-only entry points
(FFT planning) are exercised; nothing here launches a kernel.  CPU only.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from indigo_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(REPO, "include", "indigo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ib200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib._dll, name), "libindigo_b200.so does not export %s" % name
    # and the Python binding table covers the header one to one
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.version() >= 100


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Library(str(tmp_path / "nope.so"))


def test_backend_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from indigo_b200 import B200Backend
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Backend()


@pytest.mark.parametrize("dims,expect", [((416, 416, 416), [13, 8, 4]), ((512, 512, 2), [8, 8, 8]),
                                         ((23,), [23]), ((24,), [3, 8]), ((25,), [5, 5]), ((22,), [11, 2]),
                                         ((256,), [16, 16]), ((208,), [13, 16]), ((17 * 19,), [17, 19])])
def test_fft_plan_factorisation(dims, expect):
    lib = _lib.load()
    plan = ctypes.c_void_p()
    arr = (ctypes.c_int64 * len(dims))(*dims)
    lib.fft_plan_create(ctypes.byref(plan), len(dims), arr, 4)
    rad = (ctypes.c_int * 12)()
    n = lib.fft_plan_describe(plan, 0, rad, 12)
    assert list(rad[:n]) == expect
    assert int(np.prod(rad[:n])) == dims[0]
    lib.fft_plan_destroy(plan)


def test_bad_arguments_return_status_not_crash():
    lib = _lib.load()
    plan = ctypes.c_void_p()
    arr = (ctypes.c_int64 * 1)(0)
    with pytest.raises(RuntimeError, match="unsupported length"):
        lib.fft_plan_create(ctypes.byref(plan), 1, arr, 1)
    with pytest.raises(RuntimeError, match="ndim"):
        lib.fft_plan_create(ctypes.byref(plan), 4, arr, 1)
