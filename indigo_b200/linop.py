"""
The three operator classes of the standalone build (no reference package importable).

With `indigo` installed the fused SENSE nodes derive from `indigo.operators.Operator` and live
inside the reference's own trees; none of this is used.  Without it, a user still needs to write
    A = sense_operator_fused(B, ...);  y = A * x;  x = A.H * y;  B.cg(A.H * A, b, x0)
so this module supplies just that much: a base class with the calling convention of the
reference's `Operator.eval` (operators.py:20-36: `eval(y, x, alpha, beta, forward, left)` computes
`y = alpha * op(A) x + beta * y` on device arrays viewed as (rows, -1) matrices), an adjoint
wrapper and a two-factor composition.  There is no tree IR, no rewriting and no scratch arena here.
"""
import numpy as np


class Operator(object):
    """Linear map between column-major device vectors.  Subclasses provide `shape`, `dtype` and
    `_eval(y, x, alpha, beta, forward, left)`."""

    def __init__(self, backend, name=''):
        self._backend = backend
        self._name = name

    shape = property(lambda self: self._shape())
    dtype = property(lambda self: np.dtype('complex64'))

    def _shape(self):
        raise NotImplementedError()

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        raise NotImplementedError()

    def eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        rows, cols = self.shape if forward else self.shape[::-1]
        if not left:
            raise NotImplementedError("the standalone operators only multiply from the left")
        xm, ym = x.reshape((cols, -1)), y.reshape((rows, -1))
        assert xm.shape[1] == ym.shape[1], "Dimension mismatch"
        self._eval(ym, xm, alpha=alpha, beta=beta, forward=forward, left=True)

    @property
    def H(self):
        return Adjoint(self._backend, self, name=self._name + ".H")

    def __mul__(self, other):
        if isinstance(other, Operator):
            return Product(self._backend, self, other)
        if isinstance(other, np.ndarray):                      # host vector(s): upload, apply, download
            B = self._backend
            x = B.copy_array(np.asfortranarray(other.reshape((self.shape[1], -1), order='F')))
            y = B.zero_array((self.shape[0], x.shape[1]), dtype=other.dtype)
            self.eval(y, x)
            return y.to_host()
        raise ValueError("Cannot multiply Operator by %s" % type(other))


class Adjoint(Operator):
    def __init__(self, backend, child, name=''):
        Operator.__init__(self, backend, name=name)
        self.child = child

    def _shape(self):
        return self.child.shape[::-1]

    @property
    def H(self):
        return self.child

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        self.child.eval(y, x, alpha=alpha, beta=beta, forward=not forward, left=left)


class Product(Operator):
    """left * right; the intermediate vector is allocated on first use and kept."""

    def __init__(self, backend, left, right, name=''):
        Operator.__init__(self, backend, name=name or "%s*%s" % (left._name, right._name))
        if left.shape[1] != right.shape[0]:
            raise ValueError("shape mismatch in product: %r * %r" % (left.shape, right.shape))
        self.left, self.right = left, right
        self._tmp = {}

    def _shape(self):
        return (self.left.shape[0], self.right.shape[1])

    def _between(self, ncols):
        t = self._tmp.get(ncols)
        if t is None:
            t = self._tmp[ncols] = self._backend.zero_array((self.left.shape[1], ncols), np.dtype('complex64'))
        return t

    def _eval(self, y, x, alpha=1, beta=0, forward=True, left=True):
        first, second = (self.right, self.left) if forward else (self.left, self.right)
        t = self._between(int(x.shape[1]))
        first.eval(t, x, alpha=1, beta=0, forward=forward)
        second.eval(y, t, alpha=alpha, beta=beta, forward=forward)
