"""
Tree analyses: host-side mirror of indigo/analyses.py (scratch-arena size
estimate and tree search).  `Memusage` must agree with the reference byte for
byte because Optimize sizes the arena with it (reference test_analyses.py:12-43).
"""
from contextlib import contextmanager

import numpy as np

from .rewrites import Visitor


class Memusage(Visitor):
    """Peak bytes of matrices + nested Product temporaries + FFT workspaces for an
    evaluation with `ncols` right-hand sides.  reference: analyses.py:10-71."""

    def measure(self, node, ncols=1):
        self._seen = set()
        self._stack = [0]           # [0] accumulates resident matrices, the rest are live temporaries
        self._cols = [ncols]
        self._peak = 0
        self.visit(node)
        return self._peak

    @contextmanager
    def _holding(self, nbytes):
        self._stack.append(nbytes)
        yield
        self._stack.pop()

    def generic_visit(self, node):
        self._peak = max(self._peak, sum(self._stack))
        super().generic_visit(node)

    def _temporary(self, node):
        with self._holding(node._mem_usage(np.prod(self._cols))):
            self.generic_visit(node)

    visit_Product = _temporary
    visit_UnscaledFFT = _temporary

    def _resident(self, node, nbytes):
        if id(node) not in self._seen:
            self._seen.add(id(node))
            self._stack[0] += nbytes

    def visit_DenseMatrix(self, node):
        self._resident(node, node._matrix.nbytes)

    def visit_SpMatrix(self, node):
        i32 = np.dtype('int32').itemsize
        self._resident(node, node.nnz * node.dtype.itemsize + (node.shape[0] + 1) * i32 + node.nnz * i32)


class TreeHasOp(Visitor):
    """reference: analyses.py:74-86."""

    def __init__(self, op_classes):
        self._op_classes = op_classes

    def search(self, node):
        self._found = False
        self.visit(node)
        return self._found

    def visit(self, node):
        if isinstance(node, self._op_classes):
            self._found = True
        self.generic_visit(node)
