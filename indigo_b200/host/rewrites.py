"""
Tree rewrites: host-side mirror of indigo/transforms.py plus the MRI recipe
that examples/pics.py defines inline (-O1..-O3), packaged so that the SENSE
tree can be rewritten on a box where the reference package is absent.

`Optimize` also reserves the scratch arena (transforms.py:62-78): every
Product/Kron temporary is then a bump-allocated slice of one device buffer.
"""
import logging

import numpy as np
import scipy.sparse as spp

from .optree import (CompositeOperator, Product, Eye, Kron, VStack, SpMatrix, UnscaledFFT)

log = logging.getLogger(__name__)


class Transform(object):
    """Rebuilds a tree bottom-up; `visit_<ClassName>` hooks return the replacement
    node.  reference: transforms.py:19-39."""

    def visit(self, node):
        hook = getattr(self, "visit_%s" % type(node).__name__, None)
        return hook(node) if hook else self.generic_visit(node)

    def generic_visit(self, node):
        if isinstance(node, CompositeOperator):
            node._adopt([self.visit(c) for c in node._children])
        return node


class Visitor(object):
    """Read-only post-order walk.  reference: transforms.py:41-59."""

    def visit(self, node):
        self.generic_visit(node)
        hook = getattr(self, "visit_%s" % type(node).__name__, None)
        if hook:
            hook(node)

    def generic_visit(self, node):
        if isinstance(node, CompositeOperator):
            for c in node._children:
                self.visit(c)


class Optimize(Transform):
    """Runs a recipe (list of Transform classes) then reserves the arena."""

    def __init__(self, recipe):
        self._recipe = recipe or []

    def visit(self, node):
        for Step in self._recipe:
            log.info("running optimization step: %s", Step.__name__)
            node = Step().visit(node)
        b = node._backend
        b._scratch = b.empty_array((node.memusage() // node.dtype.itemsize,), node.dtype)
        b._scratch_pos = 0
        return node


class RealizeMatrices(Transform):
    """Collapses composites whose children are all SpMatrix into one SpMatrix using
    scipy on the host.  reference: transforms.py:81-175."""

    def _leafs(self, node):
        node = self.generic_visit(node)
        return node, all(isinstance(c, SpMatrix) for c in node._children)

    def visit_Product(self, node):
        node, ok = self._leafs(node)
        if not ok:
            return node
        l, r = node._children
        return SpMatrix(node._backend, l._matrix @ r._matrix, name="{}*{}".format(l._name, r._name))

    def _stack(self, node, fn):
        node, ok = self._leafs(node)
        if not ok:
            return node
        first = node._children[0]
        return SpMatrix(node._backend, fn([c._matrix for c in node._children], dtype=first.dtype),
                        name="{}+".format(first._name))

    def visit_VStack(self, node):
        return self._stack(node, spp.vstack)

    def visit_HStack(self, node):
        return self._stack(node, spp.hstack)

    def visit_BlockDiag(self, node):
        return self._stack(node, spp.block_diag)

    def visit_Kron(self, node):
        node = self.generic_visit(node)
        L, R = node.children
        if isinstance(L, Eye):
            L = L.realize()
        if isinstance(L, SpMatrix) and isinstance(R, SpMatrix):
            return SpMatrix(node._backend, spp.kron(L._matrix, R._matrix),
                            name="({}(x){})".format(L._name, R._name))
        return node

    def visit_Adjoint(self, node):
        node = self.generic_visit(node)
        c = node.child
        if isinstance(c, SpMatrix):
            return SpMatrix(node._backend, c._matrix.conjugate().transpose(), name="{}.H".format(c._name))
        return node

    def visit_Eye(self, node):
        return SpMatrix(node._backend, spp.eye(node.shape[0], dtype=node.dtype), name=node._name)

    def visit_Scale(self, node):
        node = self.generic_visit(node)
        if isinstance(node.child, SpMatrix):
            return SpMatrix(node._backend, node.child._matrix * node._val, name=node._name)
        return node

    def visit_One(self, node):
        return SpMatrix(node._backend, spp.csr_matrix(np.ones(node.shape, dtype=node.dtype)), name=node._name)


class DistributeKroniOverProd(Transform):
    """Kron(I, A*B) -> Kron(I, A) * Kron(I, B).  reference: transforms.py:178-188."""

    def visit_Kron(self, node):
        node = self.generic_visit(node)
        L, R = node.children
        if isinstance(L, Eye) and isinstance(R, Product):
            b = node._backend
            return self.visit(b.Kron(L, R.left) * b.Kron(L, R.right))
        return node


class DistributeAdjointOverProd(Transform):
    """(A*B)^H -> B^H * A^H.  reference: transforms.py:191-199."""

    def visit_Adjoint(self, node):
        node = self.generic_visit(node)
        if isinstance(node.child, Product):
            l, r = node.child.children
            return r.H * l.H
        return node


class MakeRightLeaning(Transform):
    """(A*B)*C -> A*(B*C), recursively.  reference: transforms.py:229-237 (pics.py:152-161
    visits the children first; both give the same fully right-leaning chain)."""

    def visit_Product(self, node):
        l, r = self.visit(node.left), self.visit(node.right)
        if isinstance(l, Product):
            return self.visit(l.left * (l.right * r))
        return l * r


class GroupRightLeaningProducts(Transform):
    """reference: transforms.py:240-249."""

    def visit_Product(self, node):
        node = self.generic_visit(node)
        if isinstance(node, Product):
            l, r = node.children
            if isinstance(r, Product) and isinstance(r.left, SpMatrix):
                node = (l * r.left) * r.right
        return node


# ---------------------------------------------------------------------------
# The MRI recipe of examples/pics.py:104-191 (defined inline in that script)
# ---------------------------------------------------------------------------
class AssocSpMatrices(Transform):
    """S*(A*rest) -> (S*A)*rest unless A is the FFT: groups the sparse factors on
    each side of the FFT.  reference: pics.py:138-150."""

    def visit_Product(self, node):
        l, r = self.visit(node.left), self.visit(node.right)
        if isinstance(r, Product) and isinstance(l, SpMatrix) and not isinstance(r.left, UnscaledFFT):
            return (l * r.left) * r.right
        return l * r


class DistKroniOverFFT(Transform):
    """Kron(I, A*B) -> Kron(I,A)*Kron(I,B) when the subtree holds an FFT.  reference: pics.py:128-136."""

    def visit_Kron(self, node):
        L, R = node.children
        if isinstance(L, Eye) and isinstance(R, Product) and node.has(UnscaledFFT):
            b = node._backend
            return self.visit(b.Kron(L, R.left) * b.Kron(L, R.right))
        return node


class MriRealize(Transform):
    """Multiplies adjacent sparse factors on the host: gives G' = interp*mod*scale and
    P = kron(I_C, mod*zpad*apod) * vstack(maps).  reference: pics.py:111-126."""

    def visit_VStack(self, node):
        return node.realize()

    def visit_Product(self, node):
        l, r = node.children
        if isinstance(r, VStack) and isinstance(l, Kron):
            return node.realize()
        node = self.generic_visit(node)
        l, r = node.children
        if isinstance(l, SpMatrix) and isinstance(r, SpMatrix):
            return node.realize()
        return node


class MriGoodAdjoints(Transform):
    """Stores the zero-pad side as its adjoint so that the coil combine is a row gather
    and the expand an exclusive-write scatter.  reference: pics.py:104-109."""

    def visit_SpMatrix(self, node):
        return node.H.realize().H if 'zpad' in node._name else node


def sense_recipe(level=3):
    """Recipe lists of examples/pics.py:179-191 (-O0 .. -O3; -O4 is a no-op there)."""
    recipe = []
    if level >= 1:
        recipe += [MakeRightLeaning, AssocSpMatrices, DistKroniOverFFT, MakeRightLeaning]
    if level >= 2:
        recipe += [MriRealize]
    if level >= 3:
        recipe += [MriGoodAdjoints]
    return recipe
