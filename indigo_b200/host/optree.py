"""
Host-side mirror of indigo's operator-tree IR (reference: indigo/operators.py).

The B200 backend is a drop-in below this layer: when the reference package is
importable its own `indigo.operators` is used unchanged (see
indigo_b200.register()); this mirror exists because the reference is a Python
package that does not travel to the GPU box, and the SENSE-NUFFT tree still has
to be built and walked there.  Same class names, constructor arguments, `eval`
contract (`y = alpha*op(A)*x + beta*y`, column-major 2-D views, `forward` /
`left` flags) and error behaviour as the reference, so trees written against
either are interchangeable and the parity tests read like the reference's.

Evaluation is a recursive descent that ends in Backend primitives:
  SpMatrix -> csr_matrix.forward/adjoint -> ccsrmm      (operators.py:242-263)
  UnscaledFFT -> fftn/ifftn                              (operators.py:311-338)
  Eye -> axpby, One -> onemm, DenseMatrix -> cgemm/csymm (operators.py:291-302,350-354,595-601)
  Product/Kron borrow temporaries from Backend.scratch   (operators.py:382,524)
"""
import io
import logging

import numpy as np
import scipy.sparse as spp

log = logging.getLogger(__name__)
_C64 = np.dtype("complex64")


def _unsupported_right(op):
    return NotImplementedError("Right-multiplication not implemented for {}.".format(type(op).__name__))


class Operator(object):
    """Linear operator node.  reference: operators.py:14-114."""

    def __init__(self, backend, name='', alpha=1, batch=None):
        self._backend, self._name, self._batch = backend, name, batch

    # -- evaluation ---------------------------------------------------------
    def eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        """y = alpha*A*x + beta*y (left) or alpha*x*A + beta*y; A is adjointed when
        forward is False.  x and y are viewed as 2-D column-major blocks."""
        rows, cols = self.shape if forward else self.shape[::-1]
        if left:
            x, y = x.reshape((cols, -1)), y.reshape((rows, -1))
            assert x.shape[1] == y.shape[1], "Dimension mismatch"
        else:
            x, y = x.reshape((-1, rows)), y.reshape((-1, cols))
            assert x.shape[0] == y.shape[0], "Dimension mismatch"
        self._eval(y, x, alpha=alpha, beta=beta, forward=forward, left=left)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        raise NotImplementedError()

    @property
    def shape(self):
        raise NotImplementedError()

    @property
    def dtype(self):
        raise NotImplementedError()

    # -- algebra --------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Operator):
            return Product(self._backend, self, other)
        if isinstance(other, np.ndarray):                       # convenience path, operators.py:49-54
            x = other.reshape((self.shape[1], -1), order='F')
            x_d = self._backend.copy_array(x)
            y_d = self._backend.zero_array((self.shape[0], x.shape[1]), dtype=other.dtype)
            self.eval(y_d, x_d)
            return y_d.to_host()
        if isinstance(other, (int, float, complex)):
            return Scale(self._backend, other, self)
        raise ValueError("Cannot multiply Operator by %s" % type(other))

    def __rmul__(self, other):
        if isinstance(other, (int, float, complex)):
            return self * other
        raise ValueError("Cannot right-multiply Operator by %s" % type(other))

    def __add__(self, other):
        if isinstance(other, (int, float, complex)):
            other = other * self._backend.Eye(self.shape[1])
        if isinstance(other, Operator):
            return Sum(self._backend, self, other)
        raise ValueError("Cannot right-add Operator by %s" % type(other))

    __radd__ = __add__

    def __sub__(self, other):
        return self + Scale(self._backend, -1, other)

    @property
    def H(self):
        return Adjoint(self._backend, self, name=self._name + ".H")

    # -- introspection ----------------------------------------------------------
    def dump(self):
        with io.StringIO() as f:
            self._dump(file=f, indent=0)
            return f.getvalue()

    def _dump(self, file, indent=0):
        print('{}{}, {}, {}, {} MB, {}'.format('|   ' * indent, self._name or 'noname', type(self).__name__,
                                               self.shape, self._mem_usage(ncols=1) / 1e6, self.dtype), file=file)

    def optimize(self, recipe=None):
        from .rewrites import Optimize
        return Optimize(recipe).visit(self)

    def memusage(self, ncols=1):
        from .treeinfo import Memusage
        return Memusage().measure(self, ncols)

    def _mem_usage(self, ncols):
        return 0

    def has(self, *op_classes):
        from .treeinfo import TreeHasOp
        return TreeHasOp(op_classes).search(self)


class CompositeOperator(Operator):
    """reference: operators.py:117-145."""

    def __init__(self, backend, *children, **kwargs):
        super().__init__(backend, **kwargs)
        self._adopt(children)

    def _adopt(self, children):
        self._children = children

    @property
    def children(self):
        return self._children

    @property
    def child(self):
        assert len(self._children) == 1
        return self._children[0]

    @property
    def dtype(self):
        return self._children[0].dtype

    def _dump(self, file, indent=0):
        super()._dump(file, indent)
        for c in self._children:
            c._dump(file, indent + 1)

    def realize(self):
        from .rewrites import RealizeMatrices
        return RealizeMatrices().visit(self)


class BinaryOperator(CompositeOperator):
    @property
    def left(self):
        return self._children[0]

    @property
    def right(self):
        return self._children[1]


class MatrixFreeOperator(CompositeOperator):
    def __init__(self, backend, shape, *args, dtype=_C64, **kwargs):
        super().__init__(backend, *args, **kwargs)
        self._shape, self._dtype = shape, dtype

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return self._dtype


class Adjoint(CompositeOperator):
    """A^H: flips `forward` on the way down.  reference: operators.py:173-190."""

    def __init__(self, backend, child, *args, **kwargs):
        super().__init__(backend, child, *args, **kwargs)

    @property
    def shape(self):
        return tuple(reversed(self.child.shape))

    @property
    def H(self):
        return self.child

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        self.child.eval(y, x, alpha, beta, forward=not forward, left=left)


class SpMatrix(Operator):
    """Leaf holding a scipy sparse matrix; the device copy is created on first
    use (CSR with sorted indices, or DIA).  reference: operators.py:193-263."""

    def __init__(self, backend, M, **kwargs):
        super().__init__(backend, **kwargs)
        assert isinstance(M, (spp.spmatrix, spp.sparray))
        self._matrix, self._matrix_d = M, None
        self._allow_exwrite, self._use_dia = True, False

    @property
    def dtype(self):
        return self._matrix.dtype

    @property
    def shape(self):
        return self._matrix.shape

    @property
    def nnz(self):
        return self._matrix.nnz

    def _mem_usage(self, ncols=1):
        return self._matrix.data.nbytes

    def _get_or_create_device_matrix(self):
        if self._matrix_d is None:
            self._matrix = self._matrix.astype(np.complex64)
            assert self._matrix.dtype == _C64, 'Indigo only supports single precision complex numbers for now.'
            if self._use_dia:
                self._matrix_d = self._backend.dia_matrix(self._backend, self._matrix.todia(), self._name)
            else:
                M = self._matrix.tocsr()
                M.sort_indices()
                self._matrix_d = self._backend.csr_matrix(self._backend, M, self._name)
                if not self._allow_exwrite:
                    self._matrix_d._exwrite = False
        return self._matrix_d

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        M = self._get_or_create_device_matrix()
        (M.forward if forward else M.adjoint)(y, x, alpha=alpha, beta=beta)


class DenseMatrix(Operator):
    """Leaf holding a dense complex64 matrix.  reference: operators.py:266-302."""

    def __init__(self, backend, M, **kwargs):
        super().__init__(backend, **kwargs)
        assert isinstance(M, np.ndarray)
        M = np.require(M, requirements='F')
        assert M.dtype == _C64
        assert M.ndim == 2
        self._matrix, self._matrix_d = M, None
        self._real_symmetric = M.shape[0] == M.shape[1] and np.allclose(M.imag, 0) and np.allclose(M, M.T)

    @property
    def dtype(self):
        return self._matrix.dtype

    @property
    def shape(self):
        return self._matrix.shape

    def _get_or_create_device_matrix(self):
        if self._matrix_d is None:
            self._matrix_d = self._backend.copy_array(self._matrix)
        return self._matrix_d

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left and not self._real_symmetric:
            raise NotImplementedError("Right-multiplication not implemented for non-real-symmetric {}."
                                      .format(type(self).__name__))
        M_d = self._get_or_create_device_matrix()
        if self._real_symmetric:
            self._backend.csymm(y, M_d, x, alpha=alpha, beta=beta, left=left)
        else:
            self._backend.cgemm(y, M_d, x, alpha=alpha, beta=beta, forward=forward)


class UnscaledFFT(MatrixFreeOperator):
    """Batched unscaled N-D DFT over `ft_shape`.  reference: operators.py:305-343."""

    def __init__(self, backend, ft_shape, forward=True, **kwargs):
        self._ft_shape = tuple(ft_shape)
        n = int(np.prod(self._ft_shape))
        super().__init__(backend, shape=(n, n), **kwargs)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        assert alpha == 1, "FFT expected alpha == 1, got %s" % alpha
        assert beta == 0, "FFT expected beta == 0, got %s" % beta
        X = x.reshape(self._ft_shape + (x.shape[1],))
        Y = y.reshape(self._ft_shape + (x.shape[1],))
        (self._backend.fftn if forward else self._backend.ifftn)(Y, X)

    def _mem_usage(self, ncols):
        ncols = min(ncols, self._batch or ncols)
        return self._backend._fft_workspace_size(self._ft_shape + (ncols,))


class Eye(MatrixFreeOperator):
    """Matrix-free identity: an axpby.  reference: operators.py:346-354."""

    def __init__(self, backend, n, **kwargs):
        super().__init__(backend, shape=(n, n), **kwargs)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        self._backend.axpby(beta, y, alpha, x)


class One(MatrixFreeOperator):
    """Matrix of ones.  reference: operators.py:594-601."""

    def _eval(self, y, x, alpha=1, beta=0, forward=None, left=True):
        if not left:
            raise _unsupported_right(self)
        self._backend.onemm(y, x, alpha, beta)


class Kron(BinaryOperator):
    """A (x) B.  With A = Eye the right factor is applied to the (n, C) reshape --
    the only data-parallel axis of the library (coils).  reference: operators.py:357-390."""

    @property
    def shape(self):
        return (int(np.prod([c.shape[0] for c in self._children])),
                int(np.prod([c.shape[1] for c in self._children])))

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        L, R = self.children
        if isinstance(L, Eye):
            return R.eval(y, x, alpha=alpha, beta=beta, forward=forward, left=left)
        if isinstance(R, Eye):
            return L.eval(y, x, alpha=alpha, beta=beta, forward=forward, left=not left)
        L_shape = L.shape if forward else L.shape[::-1]
        R_shape = R.shape if forward else R.shape[::-1]
        x = x.reshape((-1, L_shape[0]))
        y = y.reshape((-1, L_shape[1]))
        with self._backend.scratch(shape=(x.shape[0], L_shape[1])) as tmp:
            # X * op(L)^T on the right, then op(R) on the left (operators.py:383-390)
            L.eval(tmp, x, alpha=alpha, beta=0, forward=False, left=not left)
            tmp = tmp.reshape((R_shape[1], -1))
            R.eval(y, tmp, alpha=1, beta=beta, forward=forward, left=left)


class BlockDiag(CompositeOperator):
    """reference: operators.py:393-412."""

    @property
    def shape(self):
        return (sum(c.shape[0] for c in self._children), sum(c.shape[1] for c in self._children))

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        ho = wo = 0
        for C in self._children:
            h, w = C.shape if forward else C.shape[::-1]
            C.eval(y[ho:ho + h, :], x[wo:wo + w, :], alpha=alpha, beta=beta, forward=forward, left=left)
            ho, wo = ho + h, wo + w


class VStack(CompositeOperator):
    """reference: operators.py:415-455."""

    @property
    def shape(self):
        return (sum(c.shape[0] for c in self._children), self._children[-1].shape[1])

    def _adopt(self, children):
        widths = [c.shape[1] for c in children]
        if len(set(widths)) > 1:
            raise ValueError("Mismatched widths in VStack: attempting to stack {}".format(
                list(zip(widths, [c._name for c in children]))))
        super()._adopt(children)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        off = 0
        if forward:
            for C in self._children:
                h = C.shape[0]
                C.eval(y[off:off + h, :], x, alpha=alpha, beta=beta, forward=True, left=left)
                off += h
        else:
            self._backend.scale(y, beta)            # then accumulate every block (operators.py:440-447)
            for C in self._children:
                h = C.shape[0]
                C.eval(y, x[off:off + h, :], alpha=alpha, beta=1, forward=False, left=left)
                off += h


class HStack(CompositeOperator):
    """reference: operators.py:458-498."""

    @property
    def shape(self):
        return (self._children[-1].shape[0], sum(c.shape[1] for c in self._children))

    def _adopt(self, children):
        heights = [c.shape[0] for c in children]
        if len(set(heights)) > 1:
            raise ValueError("Mismatched heights in HStack: attempting to stack {}".format(
                list(zip(heights, [c._name for c in children]))))
        super()._adopt(children)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        off = 0
        if forward:
            self._backend.scale(y, beta)
            for C in self._children:
                w = C.shape[1]
                C.eval(y, x[off:off + w, :], alpha=alpha, beta=1, forward=True, left=left)
                off += w
        else:
            for C in self._children:
                w = C.shape[1]
                C.eval(y[off:off + w, :], x, alpha=alpha, beta=beta, forward=False, left=left)
                off += w


class Product(BinaryOperator):
    """L*R through one scratch temporary.  reference: operators.py:501-535."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._name = "{}*{}".format(self.left._name, self.right._name)

    @property
    def shape(self):
        return (self.left.shape[0], self.right.shape[1])

    def _adopt(self, children):
        L, R = children
        if L.shape[1] != R.shape[0]:
            raise ValueError("Mismatched shapes in Product: attempting {} x {} ({} x {})".format(
                L.shape, R.shape, L._name, R._name))
        super()._adopt(children)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        L, R = self._children
        first, second = (R, L) if forward else (L, R)
        with self._backend.scratch(shape=(R.shape[0], x.shape[1])) as tmp:
            first.eval(tmp, x, alpha=alpha, beta=0, forward=forward, left=left)
            second.eval(y, tmp, alpha=1, beta=beta, forward=forward, left=left)

    def _mem_usage(self, ncols):
        ncols = min(ncols, self._batch or ncols)
        return self._children[1].shape[0] * ncols * self.dtype.itemsize


class Sum(BinaryOperator):
    """L + R: the right term overwrites/accumulates first.  reference: operators.py:538-570."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._name = "{}+{}".format(self.left._name, self.right._name)

    @property
    def shape(self):
        return self.left.shape

    def _adopt(self, children):
        L, R = children
        if L.shape != R.shape:
            raise ValueError("Mismatched shapes in Sum: attempting {} + {} ({} + {})".format(
                L.shape, R.shape, L._name, R._name))
        super()._adopt(children)

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        L, R = self._children
        R.eval(y, x, alpha=alpha, beta=beta, forward=forward, left=left)
        L.eval(y, x, alpha=alpha, beta=1.0, forward=forward, left=left)


class Scale(CompositeOperator):
    """v*A (conjugated on the adjoint).  reference: operators.py:573-591."""

    def __init__(self, backend, v, child, **kwargs):
        super().__init__(backend, child, **kwargs)
        self._name = "%s*{}".format(child._name)
        self._val = v

    @property
    def shape(self):
        return self.child.shape

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        if not left:
            raise _unsupported_right(self)
        a = alpha * (self._val if forward else np.conj(self._val))
        self.child.eval(y, x, alpha=a, beta=beta, forward=forward, left=left)
