"""
Glue for machines where the reference package `indigo` is importable: registers
B200Backend with the reference's hard-coded backend lookup
(indigo/backends/__init__.py:6-64 has no plugin registry) and rebuilds the
-O1..-O3 recipe out of the reference's own Transform family.

INTEGRATION.md shows the two-line patch a maintainer would apply upstream
instead of this monkey-patch.
"""
import functools
import os


def register():
    """Makes `indigo.backends.get_backend('b200', device_id=0)` work and appends the
    backend to `available_backends()` when 'b200' is in INDIGO_TEST_BACKENDS.
    Must run before the reference's test modules import (they capture BACKENDS at
    import time, test_backends.py:11).  Returns the backend class."""
    import indigo.backends as ib
    from indigo.backends.backend import Backend
    from .backend import make_backend_class

    cls = getattr(ib, '_b200_class', None)
    if cls is not None:
        return cls
    cls = make_backend_class(Backend, name="B200Backend")
    orig_get, orig_avail = ib.get_backend, ib.available_backends

    @functools.wraps(orig_get)
    def get_backend(name, **init):
        if name == 'b200':
            return cls(**init)
        return orig_get(name, **init)

    @functools.wraps(orig_avail)
    def available_backends():
        found = orig_avail()
        if 'b200' in os.environ.get("INDIGO_TEST_BACKENDS", "np,mkl,cuda,customcpu,customgpu,b200"):
            found.append(cls)
        return found

    ib.get_backend, ib.available_backends, ib._b200_class = get_backend, available_backends, cls
    return cls


def reference_sense_recipe(level=3):
    """examples/pics.py:104-191 rebuilt on the reference's Transform base class (the
    script defines these classes inline, so they cannot be imported)."""
    from indigo.transforms import Transform
    from indigo.operators import Product, UnscaledFFT, SpMatrix, VStack, Eye, Kron
    from .host import rewrites as mine

    def port(cls_name):
        # same visit_* bodies, bound to the reference's node classes
        src = getattr(mine, cls_name)
        ns = dict(Product=Product, UnscaledFFT=UnscaledFFT, SpMatrix=SpMatrix, VStack=VStack, Eye=Eye, Kron=Kron)
        body = {}
        for k, fn in vars(src).items():
            if k.startswith('visit_'):
                g = dict(fn.__globals__); g.update(ns)
                body[k] = type(fn)(fn.__code__, g, fn.__name__, fn.__defaults__, fn.__closure__)
        return type(cls_name, (Transform,), body)

    names = []
    if level >= 1:
        names += ['MakeRightLeaning', 'AssocSpMatrices', 'DistKroniOverFFT', 'MakeRightLeaning']
    if level >= 2:
        names += ['MriRealize']
    if level >= 3:
        names += ['MriGoodAdjoints']
    cache = {}
    return [cache.setdefault(n, port(n)) for n in names]
