"""
Host-side mirror of indigo's abstract `Backend` (reference:
indigo/backends/backend.py): the column-major device array with leading
dimensions, the scratch arena, the operator builders (NUFFT, FFTc, Zpad, Interp,
Diag ...), the CSR/DIA device-matrix holders and the CG / APGD drivers.

`indigo_b200.B200Backend` derives from this class when the reference package is
not importable (the GPU box) and from the reference's own `Backend` when it is
(indigo_b200.register()); the two bases expose the same names, argument
meanings and error behaviour, which tests/test_host_mirror.py checks against
reference-generated golden vectors.
"""
import logging
from contextlib import contextmanager

import numpy as np
import scipy.sparse as spp

from . import optree as op

log = logging.getLogger(__name__)
_C64 = np.dtype('complex64')


class DeviceArrayBase(object):
    """N-d column-major array in device memory.  reference: backend.py:22-220.

    shape/dtype/_leading_dim/_own/_arr carry the reference's meaning: `_arr` is
    the memory handle, views share it, `_leading_dim` is the row pitch (in
    elements) of the allocation a 2-D view was cut from."""
    _memory = dict()

    def __init__(self, backend, shape, dtype, ld=None, own=True, data=None, name=''):
        assert isinstance(shape, (tuple, list))
        self.dtype, self.shape = dtype, shape
        self._backend = backend
        self._leading_dim = ld or shape[0]
        self._own = own
        if data is None:
            self._arr = self._malloc(shape, dtype)
            self._memory[id(self._arr)] = (name, shape, dtype)
        else:
            self._arr = data

    def reshape(self, new_shape):
        """View with a new shape.  Growing the row count needs contiguous columns;
        shrinking it re-bases the leading dimension (backend.py:59-89)."""
        old_shape = self.shape
        if -1 in new_shape:
            at = new_shape.index(-1)
            known = -int(np.prod(new_shape))
            fill = self.size // known
            assert known * fill == self.size, \
                "Cannot reshape {} into {}. (size mismatch)".format(old_shape, new_shape)
            new_shape = tuple(new_shape[:at]) + (fill,) + tuple(new_shape[at + 1:])
        if new_shape[0] > old_shape[0]:
            assert old_shape[0] == self._leading_dim, "Cannot stack non-contiguous columns."
        assert np.prod(new_shape) == self.size
        ld = new_shape[0] if new_shape[0] < old_shape[0] else self._leading_dim
        return self._backend.dndarray(self._backend, new_shape, dtype=self.dtype, ld=ld, own=False, data=self._arr)

    size = property(lambda self: np.prod(self.shape))
    itemsize = property(lambda self: self.dtype.itemsize)
    nbytes = property(lambda self: self.size * np.dtype(self.dtype).itemsize)
    ndim = property(lambda self: len(self.shape))
    contiguous = property(lambda self: self.ndim == 1 or self._leading_dim == self.shape[0])

    def _check_host(self, arr, need_f):
        assert isinstance(arr, np.ndarray)
        if self.size != arr.size:
            raise ValueError("size mismatch, expected {} got {}".format(self.shape, arr.shape))
        if self.dtype != arr.dtype:
            raise TypeError("dtype mismatch, expected {} got {}".format(self.dtype, arr.dtype))
        if need_f and not arr.flags['F_CONTIGUOUS']:
            raise TypeError("order mismatch, expected 'F' got {}".format(arr.flags['F_CONTIGUOUS']))

    def copy_from(self, arr):
        self._check_host(arr, need_f=True)
        self._copy_from(arr)

    def copy_to(self, arr):
        self._check_host(arr, need_f=False)
        self._copy_to(arr)

    def to_host(self):
        arr = np.ndarray(self.shape, self.dtype, order='F')
        self.copy_to(arr)
        return arr

    @contextmanager
    def on_host(self):
        arr_h = self.to_host()
        yield arr_h
        self.copy_from(arr_h)

    def copy(self, other=None, name=''):
        """copy(other): self <- other;  copy(): returns a fresh duplicate (backend.py:151-159)."""
        if other:
            assert isinstance(other, self._backend.dndarray)
            self._copy(other)
            return None
        dup = self._backend.zero_array(self.shape, self.dtype, name=name)
        dup._copy(self)
        return dup

    @classmethod
    def to_device(cls, backend, arr, name=''):
        arr_f = np.require(arr, requirements='F')
        d_arr = cls(backend, arr.shape, arr.dtype, name=name)
        d_arr.copy_from(arr_f)
        return d_arr

    def __del__(self):
        if getattr(self, '_own', False) and hasattr(self, '_arr'):
            self._memory.pop(id(self._arr), None)
            self._free()

    def __setitem__(self, slc, other):
        assert not (slc.start or slc.stop), "dndarray setitem cant slice"
        self._copy(other)

    # hooks a concrete backend provides (backend.py:180-220)
    def __getitem__(self, slc):
        raise NotImplementedError()

    def _copy_from(self, arr):
        raise NotImplementedError()

    def _copy_to(self, arr):
        raise NotImplementedError()

    def _copy(self, arr):
        raise NotImplementedError()

    def _malloc(self, shape, dtype):
        raise NotImplementedError()

    def _free(self):
        raise NotImplementedError()

    def _zero(self):
        raise NotImplementedError()

    @staticmethod
    def from_param(obj):
        raise NotImplementedError()


class HostBackend(object):
    """Abstract backend.  reference: backend.py:12-736."""

    dndarray = DeviceArrayBase
    ops = op                       # operator module the builders instantiate

    def __init__(self, device_id=0):
        pass

    # ------------------------------------------------------------------ arrays
    def copy_array(self, arr, name=''):
        return self.dndarray.to_device(self, arr, name=name)

    def empty_array(self, shape, dtype, name=''):
        return self.dndarray(self, shape, dtype, name=name)

    def zero_array(self, shape, dtype, name=''):
        d_arr = self.empty_array(shape, dtype, name=name)
        d_arr._zero()
        return d_arr

    def zeros_like(self, other, name=''):
        return self.zero_array(other.shape, other.dtype, name=name)

    def rand_array(self, shape, dtype=_C64, name=''):
        x = np.random.random(shape) + 1j * np.random.random(shape)
        return self.copy_array(np.require(x, dtype=_C64, requirements='F'), name=name)

    def get_max_threads(self):
        return 1

    def barrier(self):
        pass

    def mem_usage(self):
        total = 0
        for name, shape, dtype in self.dndarray._memory.values():
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            total += n
            if n > 1e6:
                log.info("  %40s: % 3.0f MB, %20s, %15s", name, n / 1e6, shape, dtype)
        return total

    @contextmanager
    def scratch(self, shape=None, nbytes=None):
        """Temporary complex64 block: a bump-allocated slice of the arena reserved by
        Optimize, else a fresh zeroed array (backend.py:262-281)."""
        assert not (shape is not None and nbytes is not None), \
            "Specify either shape or nbytes to backend.scratch()."
        if nbytes is not None:
            shape = (nbytes // _C64.itemsize,)
        size = int(np.prod(shape))
        if hasattr(self, '_scratch'):
            pos, total = self._scratch_pos, self._scratch.size
            assert pos + size <= total, \
                "Not enough scratch memory (wanted %d MB, but only have %d MB available of %d MB total)." % (
                    size / 1e6, (total - pos) / 1e6, total / 1e6)
            self._scratch_pos += size
            try:
                yield self._scratch[pos:pos + size].reshape(shape)
            finally:
                self._scratch_pos -= size
        else:
            yield self.zero_array(shape, dtype=np.complex64)

    # ------------------------------------------------------------------ builders (backend.py:287-448)
    def SpMatrix(self, M, **kwargs):
        assert isinstance(M, (spp.spmatrix, spp.sparray))
        return self.ops.SpMatrix(self, M, **kwargs)

    def DenseMatrix(self, M, **kwargs):
        assert isinstance(M, np.ndarray) and M.ndim == 2
        return self.ops.DenseMatrix(self, M, **kwargs)

    def Diag(self, v, **kwargs):
        v = np.require(v, requirements='F')
        if v.ndim > 1:
            v = v.flatten(order='A')
        dtype = kwargs.get('dtype', _C64)
        return self.SpMatrix(spp.diags(v, offsets=0).astype(dtype), **kwargs)

    def Adjoint(self, A, **kwargs):
        return self.ops.Adjoint(self, A, **kwargs)

    def KronI(self, c, B, **kwargs):
        return self.ops.Kron(self, self.Eye(c), B, **kwargs)

    def Kron(self, A, B, **kwargs):
        return self.ops.Kron(self, A, B, **kwargs)

    def BlockDiag(self, Ms, **kwargs):
        return self.ops.BlockDiag(self, *Ms, **kwargs)

    def VStack(self, Ms, **kwargs):
        return self.ops.VStack(self, *Ms, **kwargs)

    def HStack(self, Ms, **kwargs):
        return self.ops.HStack(self, *Ms, **kwargs)

    def UnscaledFFT(self, shape, dtype, **kwargs):
        return self.ops.UnscaledFFT(self, shape, dtype, **kwargs)

    def Eye(self, n, dtype=_C64, **kwargs):
        return self.ops.Eye(self, n, dtype=dtype, **kwargs)

    def One(self, shape, dtype=_C64, **kwargs):
        return self.ops.One(self, shape, dtype=dtype, **kwargs)

    def FFT(self, shape, dtype, **kwargs):
        """Unitary FFT = Diag(1/sqrt(n)) * UnscaledFFT (backend.py:347-353)."""
        n = np.prod(shape)
        S = self.Diag(np.ones(n, order='F', dtype=dtype) / np.sqrt(n), name='scale')
        return S * self.UnscaledFFT(shape, dtype, **kwargs)

    def FFTc(self, ft_shape, dtype, normalize=True, **kwargs):
        """Centred FFT = Mod * FFT * Mod with the phase ramp of backend.py:355-369."""
        grid = np.mgrid[tuple(slice(d) for d in ft_shape)]
        ramp = 0
        for i, n in enumerate(ft_shape):
            c = n // 2
            ramp = ramp + (grid[i] - c / 2.0) * (c / n)
        M = self.Diag(np.exp(1j * 2.0 * np.pi * ramp).astype(dtype), name='mod')
        F = (self.FFT if normalize else self.UnscaledFFT)(ft_shape, dtype=dtype, **kwargs)
        return M * F * M

    def Zpad(self, M, N, mode='center', dtype=_C64, **kwargs):
        """Selection matrix embedding an N-shaped block into an M-shaped array (backend.py:371-387)."""
        if mode == 'center':
            cut = tuple(slice(m // 2 + int(np.ceil(-n / 2)), m // 2 + int(np.ceil(n / 2))) for m, n in zip(M, N))
        elif mode == 'edge':
            cut = tuple(slice(n) for n in N)
        else:
            cut = ()
        lin = np.arange(int(np.prod(M)), dtype=int).reshape(M, order='F')
        rows = lin[cut].flatten(order='F')
        cols = np.arange(rows.size)
        mat = spp.coo_matrix((np.ones_like(cols), (rows, cols)), shape=(np.prod(M), np.prod(N)), dtype=dtype)
        return self.SpMatrix(mat, **kwargs)

    def Crop(self, M, N, dtype=_C64, **kwargs):
        return self.Zpad(N, M, dtype=dtype, **kwargs).H

    def Interp(self, N, coord, width, table, dtype=_C64, **kwargs):
        """Kaiser-Bessel gridding matrix, samples x grid (backend.py:392-401)."""
        assert len(N) == 3
        ndim, npts = coord.shape[0], int(np.prod(coord.shape[1:]))
        from .noncart import interp_mat
        M = interp_mat(npts, N, width, table, coord.reshape((ndim, -1), order='F'), 1).astype(dtype)
        return self.SpMatrix(M, **kwargs)

    def NUFFT(self, M, N, coord, width=3, n=128, oversamp=None, dtype=_C64, **kwargs):
        """G * FFTc * Zpad * Diag(rolloff): image N -> samples M (backend.py:403-442)."""
        assert len(M) == 3 and len(N) == 3
        assert M[1:] == coord.shape[1:]
        if isinstance(oversamp, tuple):
            omin = min(oversamp)
        else:
            omin, oversamp = oversamp, (oversamp,) * 3
        from scipy.signal.windows import kaiser
        from .noncart import rolloff3
        oN = tuple(int(n_ * o) for n_, o in zip(N, oversamp))
        Z = self.Zpad(oN, N, dtype=dtype, name='zpad')
        F = self.FFTc(oN, dtype=dtype, name='fft')
        beta = np.pi * np.sqrt(((width * 2. / omin) * (omin - 0.5)) ** 2 - 0.8)
        kb = kaiser(2 * n + 1, beta)[n:]
        G = self.Interp(oN, coord, width, kb, dtype=np.float32, name='interp')
        R = self.Diag(rolloff3(omin, width, beta, N), name='apod')
        return G * F * Z * R

    def Convolution(self, kernel, normalize=True, name='noname'):
        F = self.FFTc(kernel.shape, name='%s.convF' % name, normalize=normalize, dtype=np.complex64)
        K = self.Diag(F * kernel, name='%s.convK' % name)
        return F.H * K * F

    # ------------------------------------------------------------------ primitives (backend.py:453-533)
    def _abstract(self, *a, **k):
        raise NotImplementedError()

    axpby = dot = norm2 = scale = cgemm = csymm = fftn = ifftn = ccsrmm = cdiamm = onemm = max = _abstract

    def _fft_workspace_size(self, x_shape):
        return 0

    def pdot(self, x, y, comm):
        v = self.dot(x, y)
        return comm.allreduce(v) if comm is not None else v

    def pnorm2(self, x, comm):
        v = self.norm2(x)
        return comm.allreduce(v) if comm is not None else v

    # ------------------------------------------------------------------ device matrices
    class csr_matrix(object):
        """Device CSR holder.  reference: backend.py:535-596."""
        _index_base = 0

        def __init__(self, backend, A, name='mat'):
            if not isinstance(A, spp.csr_matrix):
                A = A.tocsr()
            A = self._type_correct(A)
            self._backend = backend
            self.rowPtrs = backend.copy_array(A.indptr + self._index_base, name=name + ".rowPtrs")
            self.colInds = backend.copy_array(A.indices + self._index_base, name=name + ".colInds")
            self.values = backend.copy_array(A.data, name=name + ".data")
            self.shape, self.dtype = A.shape, A.dtype
            self._inspect(A, name)

        def _inspect(self, A, name):
            per_col = np.bincount(A.indices, minlength=A.shape[1]) if A.nnz else np.zeros(A.shape[1], dtype=int)
            self._row_frac = np.count_nonzero(np.diff(A.indptr)) / A.shape[0]
            self._col_frac = np.count_nonzero(per_col) / A.shape[1]
            self._exwrite = int(per_col.max(initial=0) <= 1)

        def _check(self, y, x):
            assert x.dtype == _C64, "Bad dtype: expected compelx64, got %s" % x.dtype
            assert y.dtype == _C64, "Bad dtype: expected compelx64, got %s" % y.dtype
            assert self.values.dtype == _C64

        def forward(self, y, x, alpha=1, beta=0):
            self._check(y, x)
            self._backend.ccsrmm(y, self.shape, self.colInds, self.rowPtrs, self.values, x,
                                 alpha=alpha, beta=beta, adjoint=False, exwrite=True)

        def adjoint(self, y, x, alpha=1, beta=0):
            self._check(y, x)
            self._backend.ccsrmm(y, self.shape, self.colInds, self.rowPtrs, self.values, x,
                                 alpha=alpha, beta=beta, adjoint=True, exwrite=self._exwrite)

        nbytes = property(lambda self: self.rowPtrs.nbytes + self.colInds.nbytes + self.values.nbytes)
        nnz = property(lambda self: self.values.size)

        def _type_correct(self, A):
            return A.astype(np.complex64)

    class dia_matrix(object):
        """Device DIA holder; data is stored transposed, (ncols x noffsets).  reference: backend.py:599-633."""

        def __init__(self, backend, A, name='mat'):
            assert isinstance(A, spp.dia_matrix)
            A = A.astype(np.complex64)
            self._backend = backend
            self.data = backend.copy_array(A.data.T, name=name + ".data")
            self.offsets = backend.copy_array(A.offsets, name=name + ".data")
            self.shape, self.dtype = A.shape, A.dtype
            self._row_frac = self._col_frac = 1

        def forward(self, y, x, alpha=1, beta=0):
            self._backend.cdiamm(y, self.shape, self.offsets, self.data, x, alpha=alpha, beta=beta, adjoint=False)

        def adjoint(self, y, x, alpha=1, beta=0):
            self._backend.cdiamm(y, self.shape, self.offsets, self.data, x, alpha=alpha, beta=beta, adjoint=True)

        nbytes = property(lambda self: self.offsets.nbytes + self.data.nbytes)
        nnz = property(lambda self: self.data.size)

    # ------------------------------------------------------------------ solvers
    def cg(self, A, b_h, x_h, lamda=0.0, tol=1e-10, maxiter=100, team=None):
        """Conjugate gradient on (A + lamda I) x = b; x_h holds the start and receives the
        result.  Update order of backend.py:639-689."""
        x, b = self.copy_array(x_h, name='x'), self.copy_array(b_h, name='b')
        Ap = x.copy()
        r = b
        A.eval(Ap, x)
        self.axpby(1, r, -1, Ap)
        self.axpby(1, r, -lamda, x)
        p = r.copy(name='p')
        rr = self.pnorm2(r, team)
        r0 = rr
        for it in range(maxiter):
            A.eval(Ap, p)
            self.axpby(1, Ap, lamda, p)
            alpha = rr / self.pdot(p, Ap, team)
            self.axpby(1, x, alpha, p)
            self.axpby(1, r, -alpha, Ap)
            r2 = self.pnorm2(r, team)
            beta = r2 / rr
            self.scale(p, beta)
            self.axpby(1, p, 1, r)
            rr = r2
            resid = np.sqrt(rr / r0)
            log.info("iter %d, residual %g", it, resid.real)
            if resid < tol:
                log.info("cg reached tolerance")
                break
        else:
            log.info("cg reached maxiter")
        x.copy_to(x_h)

    def apgd(self, gradf, proxg, alpha, x_h, maxiter=100, team=None):
        """Accelerated proximal gradient descent (FISTA momentum).  backend.py:691-732."""
        x_k = self.copy_array(x_h)
        y_k, y_k1, x_k1, gf = x_k.copy(), x_k.copy(), x_k.copy(), x_k.copy()
        t_k = 1
        for it in range(1, maxiter + 1):
            gradf(gf, y_k)
            self.axpby(1, x_k, -alpha, gf)
            proxg(x_k, alpha)
            t_k1 = (1.0 + np.sqrt(1.0 + 4.0 * t_k ** 2)) / 2.0
            t_ratio = (t_k - 1) / t_k1
            self.axpby(0, y_k1, 1 + t_ratio, x_k)
            self.axpby(1, y_k1, -t_ratio, x_k1)
            x_k1.copy(x_k)
            y_k.copy(y_k1)
            log.info("iter %d", it)
        x_k.copy_to(x_h)
