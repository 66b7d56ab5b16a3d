"""
Test helper: the UNMODIFIED reference package (`indigo`) for tests that run the reference's own
operators / transforms / solvers on the B200 backend, or compare against its NumpyBackend.

Where it comes from: /root/reference in the build container, oracle/_ref/reference_pkg.zip (packed
by `make -C oracle ref`, unpacked to a temp directory) on the GPU box; loaded through the
non-invasive compatibility shim tests/golden/refshim.py.  Tests that need it call `reference()`,
which skips the test when neither source exists.
"""
import os
import sys

import pytest

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_HERE, "golden") not in sys.path:
    sys.path.insert(0, os.path.join(_HERE, "golden"))

import refshim  # noqa: E402


def reference():
    """The imported, shimmed reference package; skips the calling test when it is unavailable."""
    if not refshim.have_reference():
        pytest.skip("reference package not available (run `make -C oracle ref` where /root/reference exists)")
    return refshim.load_reference()


def numpy_backend():
    """The reference's own NumpyBackend (the parity oracle of BASELINE.json's north star)."""
    reference()
    from indigo.backends import get_backend
    return get_backend('numpy')


_b200_cls = {}


def b200_reference_backend(device_id=0):
    """B200Backend built on the reference's `indigo.backends.backend.Backend` and registered with
    `indigo.backends.get_backend('b200')`: the drop-in configuration of INTEGRATION.md."""
    reference()
    import indigo_b200
    import indigo.backends
    cls = indigo_b200.register()
    assert indigo.backends._b200_class is cls
    return indigo.backends.get_backend('b200', device_id=device_id)


def pics_recipe(level=3):
    """The -O`level` recipe of examples/pics.py:104-191, exec'd from the reference's own text (the script
    defines its Transform classes inline and cannot be imported without running a reconstruction)."""
    reference()
    root = refshim.reference_root()
    src = open(os.path.join(root, "examples", "pics.py")).read()
    seg = src[src.index("import scipy.sparse as spp"):src.index("recipe = []")]
    ns = {}
    exec(compile(seg, "examples/pics.py[104:177]", "exec"), ns)
    steps = []
    if level >= 1:
        steps += [ns[k] for k in ("MakeRightLeaning", "AssocSpMatrices", "DistKroniOverFFT", "MakeRightLeaning")]
    if level >= 2:
        steps += [ns["MriRealize"]]
    if level >= 3:
        steps += [ns["MriGoodAdjoints"]]
    return steps
