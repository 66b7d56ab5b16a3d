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


_recipe_cache = {}


def reference_sense_recipe(level=3):
    """The -O`level` recipe of examples/pics.py:179-191 as a list of Transform classes of the
    reference's own family (`indigo.transforms.Transform`, nodes of `indigo.operators`).

    The script defines its recipe classes inline (pics.py:104-177), so they cannot be imported;
    these are restatements of what each step does to the SENSE tree
        KronI(C, G * (mod * (scale*FFT) * mod) * zpad * apod) * VStack_c Diag(maps_c)
    written against the reference's node API (`left/right/children`, `realize()`, `has()`).
    tests/test_gpu_reference.py runs this list and the script's own classes (exec'd from the
    reference's text) on the same tree and checks that both end in the same six Backend calls."""
    if level in _recipe_cache:
        return list(_recipe_cache[level])
    from indigo.transforms import Transform
    from indigo.operators import Product, UnscaledFFT, SpMatrix, VStack, Eye, Kron

    class MakeRightLeaning(Transform):
        """(A*B)*C -> A*(B*C) all the way down (pics.py:152-161)."""

        def visit_Product(self, node):
            lhs, rhs = self.visit(node.left), self.visit(node.right)
            if not isinstance(lhs, Product):
                return lhs * rhs
            return self.visit(lhs.left * (lhs.right * rhs))

    class AssocSpMatrices(Transform):
        """S*(A*rest) -> (S*A)*rest unless A is the FFT: gathers the sparse factors on either
        side of the transform (pics.py:138-150)."""

        def visit_Product(self, node):
            lhs, rhs = self.visit(node.left), self.visit(node.right)
            inner = getattr(rhs, 'children', None) if isinstance(rhs, Product) else None
            if inner and isinstance(lhs, SpMatrix) and not isinstance(inner[0], UnscaledFFT):
                return (lhs * inner[0]) * inner[1]
            return lhs * rhs

    class DistKroniOverFFT(Transform):
        """Kron(I, A*B) -> Kron(I,A)*Kron(I,B) where the subtree holds the FFT (pics.py:128-136)."""

        def visit_Kron(self, node):
            eye, body = node.children
            if isinstance(eye, Eye) and isinstance(body, Product) and node.has(UnscaledFFT):
                make = node._backend.Kron
                return self.visit(make(eye, body.left) * make(eye, body.right))
            return node

    class MriRealize(Transform):
        """Multiplies adjacent sparse factors on the host: G' = interp*mod*scale and
        P = kron(I_C, mod*zpad*apod) * vstack(maps) (pics.py:111-126)."""

        def visit_VStack(self, node):
            return node.realize()

        def visit_Product(self, node):
            lhs, rhs = node.children
            if isinstance(lhs, Kron) and isinstance(rhs, VStack):
                return node.realize()
            node = self.generic_visit(node)
            if all(isinstance(c, SpMatrix) for c in node.children):
                return node.realize()
            return node

    class MriGoodAdjoints(Transform):
        """Keeps the zero-pad side as its stored adjoint: the coil combine becomes a row gather
        and the expand an exclusive-write scatter (pics.py:104-109)."""

        def visit_SpMatrix(self, node):
            return node.H.realize().H if 'zpad' in node._name else node

    steps = []
    if level >= 1:
        steps += [MakeRightLeaning, AssocSpMatrices, DistKroniOverFFT, MakeRightLeaning]
    if level >= 2:
        steps += [MriRealize]
    if level >= 3:
        steps += [MriGoodAdjoints]
    _recipe_cache[level] = steps
    return list(steps)
