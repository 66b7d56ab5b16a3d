"""
pytest plugin used ONLY to run the reference's own test modules (indigo/backends/test_backends.py,
indigo/test_operators.py) on the B200 backend:

    INDIGO_TEST_BACKENDS=b200 python -m pytest -p refsuite_plugin <reference>/indigo/backends/test_backends.py

Those modules capture `BACKENDS = available_backends()` at import time (test_backends.py:11), so the
backend has to be registered before collection: pytest_configure loads the reference through the
compatibility shim and calls indigo_b200.register().  IB200_REFSUITE_STRIDE=k keeps every k-th
collected test (deterministic sub-sample for the quick run inside tests/test_gpu_reference.py;
tools/run_reference_suites.py runs everything).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(_HERE), _HERE, os.path.join(_HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    import refshim
    refshim.load_reference()
    import indigo_b200
    indigo_b200.register()


def pytest_collection_modifyitems(config, items):
    stride = int(os.environ.get("IB200_REFSUITE_STRIDE", "1"))
    offset = int(os.environ.get("IB200_REFSUITE_OFFSET", "0"))
    # reference tests that cannot pass on any machine without Intel MKL (they ask for the mkl / customcpu backends by name)
    drop = ("test_get_backend[mkl]", "test_get_backend[customcpu]")
    keep = [it for i, it in enumerate(items) if i % stride == offset % stride and not it.nodeid.endswith(drop)]
    gone = [it for it in items if it not in keep]
    if gone:
        config.hook.pytest_deselected(items=gone)
        items[:] = keep
