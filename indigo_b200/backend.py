"""
B200Backend: indigo's `Backend` interface executed by hand-written sm_100a
kernels through the C ABI of libindigo_b200.so (include/indigo_b200.h).

Drop-in boundary.  The class is produced by
`make_backend_class(Base)`:
  * Base = indigo.backends.backend.Backend when the reference package is
    importable (indigo_b200.register() then exposes it as get_backend('b200')),
    so trees built by the reference's own operators.py / transforms.py run
    unchanged -- tests/test_gpu_reference.py does exactly that, and runs the
    reference's own test_backends.py / test_operators.py on this backend;
  * Base = indigo_b200.standalone.StandaloneBase on machines without the
    reference (the GPU box): arrays, primitives, cg, the fused SENSE node and the
    direct six-call operator, no operator tree.
Every method below replaces one abstract method of the reference
(backend.py:453-533) or one hook of its device array (backend.py:180-220) and
ends in exactly one C-ABI call; PyTorch only owns the device buffers and the
current stream.  There is no CPU fallback: constructing the backend without a
CUDA device or without the built library raises RuntimeError.
"""
import ctypes
import logging

import numpy as np
import scipy.sparse as spp

from . import _lib
from .standalone import StandaloneBase

log = logging.getLogger(__name__)
_C64 = np.dtype('complex64')


class DevPtr(ctypes.c_ulong):
    """Device address as the `ctypes.c_ulong` the reference expects in
    `dndarray._arr` (operators.py:322-332 inspects `.value`), carrying a
    reference to the torch tensor that owns the memory so views keep it alive."""

    def __init__(self, value=0, keep=None):
        ctypes.c_ulong.__init__(self, value)
        self._keep = keep


def _re_im(v):
    v = complex(v)
    return float(v.real), float(v.imag)


def make_backend_class(Base, name="B200Backend"):
    """Builds the backend class on top of `Base`: the reference's `indigo.backends.backend.Backend`, or
    indigo_b200.standalone.StandaloneBase when the reference is not installed."""

    class B200Array(Base.dndarray):
        """Column-major device array; pointer model of the reference's CUDA array
        (cuda.py:126-208): slicing is pointer arithmetic keeping the parent's
        leading dimension, 2-D copies are pitched."""

        # -- memory ----------------------------------------------------------
        def _malloc(self, shape, dtype):
            b = self._backend
            nbytes = max(int(np.prod(shape)) * np.dtype(dtype).itemsize, 16)     # never a null handle
            t = b._torch.empty(nbytes, dtype=b._torch.uint8, device=b._device)   # 512-byte aligned
            return DevPtr(t.data_ptr(), keep=t)

        def _free(self):
            self._arr._keep = None

        def _zero(self):
            b = self._backend
            if self.ndim == 2 and not self.contiguous and self.shape[1] > 1:
                z = np.zeros(self.shape, dtype=self.dtype, order='F')
                self._copy_from(z)
            else:
                b._lib.memset0(b._stream, self._arr.value, int(self.nbytes))

        def _pitched(self):
            """(pitch_bytes, width_bytes, height) of this array's memory."""
            it = np.dtype(self.dtype).itemsize
            if self.ndim == 2:
                return self._leading_dim * it, self.shape[0] * it, self.shape[1]
            return int(self.nbytes), int(self.nbytes), 1

        def _copy_from(self, arr):
            assert arr.flags['F_CONTIGUOUS']
            b = self._backend
            pitch, width, height = self._pitched()
            if self.size:
                b._lib.copy2d(b._stream, self._arr.value, pitch, arr.ctypes.data, width, width, height, 1)
                if not b._pinned(arr):
                    b._lib.stream_sync(b._stream)       # pageable source: safe to reuse on return

        def _copy_to(self, arr):
            b = self._backend
            dst = arr if arr.flags['F_CONTIGUOUS'] else np.empty(self.shape, self.dtype, order='F')
            pitch, width, height = self._pitched()
            if self.size:
                b._lib.copy2d(b._stream, dst.ctypes.data, width, self._arr.value, pitch, width, height, 2)
                b._lib.stream_sync(b._stream)
            if dst is not arr:
                arr[...] = dst.reshape(arr.shape, order='F')

        def _copy(self, other):
            b = self._backend
            dp, width, height = self._pitched()
            if other.ndim == 2 and self.ndim == 2:
                sp = other._leading_dim * np.dtype(other.dtype).itemsize
            else:
                assert self.contiguous and other.contiguous
                sp, dp, width, height = int(self.nbytes), int(self.nbytes), int(self.nbytes), 1
            if self.size:
                b._lib.copy2d(b._stream, self._arr.value, dp, other._arr.value, sp, width, height, 0)

        # -- views -----------------------------------------------------------
        def __getitem__(self, slc):
            if not isinstance(slc, tuple):
                slc = (slc,)
            start, shape = [], []
            for s, n in zip(slc, self.shape):
                if isinstance(s, (int, np.integer)):
                    s = slice(s, s + 1)
                lo = 0 if s.start is None else s.start
                hi = n if s.stop is None else s.stop
                if lo < 0: lo += n
                if hi < 0: hi += n
                hi = max(min(hi, n), lo)
                start.append(lo); shape.append(hi - lo)
            for n in self.shape[len(slc):]:
                start.append(0); shape.append(n)
            strides, acc = [], 1
            for d, n in enumerate(self.shape):
                strides.append(acc)
                acc *= self._leading_dim if (d == 0 and self.ndim == 2) else n
            off = sum(a * s for a, s in zip(start, strides)) * np.dtype(self.dtype).itemsize
            ptr = DevPtr(self._arr.value + off, keep=getattr(self._arr, '_keep', None))
            return self._backend.dndarray(self._backend, tuple(shape), self.dtype,
                                          ld=self._leading_dim, own=False, data=ptr)

        @staticmethod
        def from_param(obj):
            if not isinstance(obj, Base.dndarray):
                raise ctypes.ArgumentError('{} is not a dndarray'.format(type(obj)))
            return obj._arr

        # -- helpers used by the primitives -----------------------------------
        @property
        def ptr(self):
            return self._arr.value

        @property
        def ld(self):
            return int(self._leading_dim)

        def _columns(self):
            """(ptr, length) chunks that are contiguous in memory."""
            it = np.dtype(self.dtype).itemsize
            if self.ndim != 2 or self.contiguous or self.shape[1] == 1:
                return [(self.ptr, int(self.size))]
            return [(self.ptr + c * self.ld * it, int(self.shape[0])) for c in range(self.shape[1])]

    class B200Csr(object):
        """Device CSR holder of Backend.csr_matrix (backend.py:535-596) with a device-side inspector and, for
        matrices whose adjoint is a scatter with collisions, a stored conjugate transpose so that A^H x is a
        gather (replaces the atomic path of _customcpu.c:49-79 / cusparse's transpose mode).
        User-visible rowPtrs/colInds/values stay bit-identical to the reference's, index dtype included:
        scipy switches to int64 when a dimension or nnz reaches 2^31 (cfg4's P on one GPU has 2.3 G columns);
        such a matrix is additionally held as column blocks of fewer than `max_block_cols` columns with
        int32 local indices, and a product is the sum / concatenation of the block products."""
        _index_base = 0
        max_block_cols = (1 << 31) - 1024

        def __init__(self, backend, A, name='mat'):
            if not isinstance(A, spp.csr_matrix):
                A = A.tocsr()
            A = self._type_correct(A)
            self._backend = backend
            self._name = name
            self.shape, self.dtype = A.shape, A.dtype
            self._adj = None
            self._blocks = None
            wide = A.indices.dtype != np.int32 or A.indptr.dtype != np.int32 or A.shape[1] > self.max_block_cols
            if wide:
                if A.nnz >= (1 << 31):
                    raise ValueError("b200 backend holds at most 2^31-1 stored entries per matrix (matrix %s has %d); "
                                     "split the operator (SURVEY.md 8a, cfg4 note)" % (name, A.nnz))
                self.rowPtrs = backend.copy_array(A.indptr, name=name + ".rowPtrs")
                self.colInds = backend.copy_array(A.indices, name=name + ".colInds")
                self.values = backend.copy_array(A.data, name=name + ".data")
                self._split_columns(A)
                return
            self.rowPtrs = backend.copy_array(A.indptr, name=name + ".rowPtrs")
            self.colInds = backend.copy_array(A.indices, name=name + ".colInds")
            self.values = backend.copy_array(A.data, name=name + ".data")
            self._inspect_device()

        def _type_correct(self, A):
            return A.astype(np.complex64)

        nbytes = property(lambda self: self.rowPtrs.nbytes + self.colInds.nbytes + self.values.nbytes)
        nnz = property(lambda self: self.values.size)

        def _split_columns(self, A):
            """Column blocks [c0, c1) of a matrix with 64-bit indices, each a B200Csr of its own with int32 indices
            (`A x = sum_b A_b x[c0:c1]`, `(A^H y)[c0:c1] = A_b^H y`); the inspector's figures are combined."""
            k, step = A.shape[1], int(self.max_block_cols)
            self._blocks = []
            nzcols, exw, nzrows = 0, 1, np.zeros(A.shape[0], dtype=bool)
            for c0 in range(0, k, step):
                c1 = min(k, c0 + step)
                sub = A[:, c0:c1].tocsr()
                sub.sort_indices()
                sub = spp.csr_matrix((sub.data, sub.indices.astype(np.int32), sub.indptr.astype(np.int32)),
                                     shape=sub.shape)
                blk = type(self)(self._backend, sub, name="%s[:, %d:%d]" % (self._name, c0, c1))
                self._blocks.append((c0, c1, blk))
                nzcols += int(round(blk._col_frac * (c1 - c0)))
                exw &= blk._exwrite
                nzrows |= np.diff(sub.indptr) > 0
            self._row_frac = float(nzrows.sum()) / A.shape[0] if A.shape[0] else 1.0
            self._col_frac = nzcols / k if k else 1.0
            self._exwrite = int(exw)

        @classmethod
        def from_device(cls, backend, shape, rowPtrs, colInds, values, name='mat'):
            """Wraps CSR arrays that already live on the device (built by ib200_kb_fill /
            ib200_sense_ph_fill) and runs the same device inspector."""
            self = cls.__new__(cls)
            self._backend, self._name = backend, name
            self.rowPtrs, self.colInds, self.values = rowPtrs, colInds, values
            self.shape, self.dtype = tuple(int(v) for v in shape), _C64
            self._adj = None
            self._blocks = None
            self._inspect_device()
            return self

        def _inspect_device(self):
            backend, (m, k) = self._backend, self.shape
            out = (ctypes.c_int64 * 4)()
            work = backend.empty_array((max(k, 1),), np.dtype('int32'), name=self._name + ".inspect")
            backend._lib.csr_inspect(backend._stream, m, k, self.colInds.ptr, self.rowPtrs.ptr, work.ptr, out)
            self._row_frac = out[0] / m if m else 1.0
            self._col_frac = out[1] / k if k else 1.0
            self._exwrite = int(out[2])
            self._max_col_count = int(out[3])
            log.debug("matrix %s: %d%% nonzero rows, %d%% nonzero cols, exwrite=%d", self._name,
                      100 * self._row_frac, 100 * self._col_frac, self._exwrite)

        def _use_stored_adjoint(self):
            # 'auto': always.  With the coil-interleaved gather (csrmm_il.cu) the transposed matrix
            # costs one full line per stored entry, against one L2 atomic per entry AND coil for the
            # scatter (measured 44 ms vs 23 ms forward at cfg3 before the re-layout, profiles/r01_s1_*).
            mode = self._backend.stored_adjoints
            return True if mode == 'auto' else bool(mode)

        def _stored_adjoint(self):
            if self._adj is None:
                b = self._backend
                (m, k), nnz = self.shape, int(self.values.size)
                t_ptr = b.empty_array((k + 1,), np.dtype('int32'), name=self._name + ".H.rowPtrs")
                t_ind = b.empty_array((max(nnz, 1),), np.dtype('int32'), name=self._name + ".H.colInds")
                t_val = b.empty_array((max(nnz, 1),), _C64, name=self._name + ".H.data")
                work = b.empty_array((k + 1,), np.dtype('int32'))
                b._lib.csr_transpose_conj(b._stream, m, k, nnz, self.values.ptr, self.colInds.ptr, self.rowPtrs.ptr,
                                          t_val.ptr, t_ind.ptr, t_ptr.ptr, work.ptr, None)
                self._adj = (t_ptr, t_ind, t_val)
            return self._adj

        # -- multi-column products of real-valued matrices ------------------------------
        # Gridding matrices are real on the grids MRI uses (the centring phase is +-1).  For them the
        # interleaved gather runs on packed (int32 column, float weight) entries staged through shared
        # memory (ib200_ccsrmm_ilr, 8 bytes per entry instead of 12, half the multiplies), and rows
        # longer than `long_thresh` entries get a whole CTA each.  Decided once per matrix.
        long_thresh = 512

        def _packed(self, which):
            """{'pk', 'ptr', 'long', 'nlong'} for which = 'fwd' (this matrix) or 'adj' (stored adjoint), or None
            when the values are not real / the matrix is empty."""
            cache = self.__dict__.setdefault('_pk_cache', {})
            if which in cache:
                return cache[which]
            b = self._backend
            if which == 'fwd':
                rows, ptr, ind, val = self.shape[0], self.rowPtrs, self.colInds, self.values
            else:
                rows = self.shape[1]
                ptr, ind, val = self._stored_adjoint()
            nnz = int(self.values.size)
            info = None
            if nnz > 0 and rows > 0 and b.packed_real:
                pk = b.zero_array((nnz + 2,), np.dtype('int64'), name=self._name + ".packed." + which)
                hmax = (ctypes.c_float * 2)()
                b._lib.csr_pack_real(b._stream, nnz, val.ptr, ind.ptr, pk.ptr, hmax)
                if hmax[1] <= 1e-8 * hmax[0]:
                    cnt = ctypes.c_int()
                    b._lib.csr_long_rows(b._stream, rows, ptr.ptr, self.long_thresh, None, 0, ctypes.byref(cnt))
                    nlong, lr = int(cnt.value), None
                    if nlong:
                        lr = b.empty_array((nlong,), np.dtype('int32'), name=self._name + ".longrows." + which)
                        b._lib.csr_long_rows(b._stream, rows, ptr.ptr, self.long_thresh, lr.ptr, nlong, ctypes.byref(cnt))
                    info = dict(pk=pk, ptr=ptr, long=lr, nlong=nlong)
            cache[which] = info
            return info

        def _packed_product(self, which, y, x, alpha, beta):
            """Y = alpha * op(A) X + beta * Y through the packed real-weight gather; False when not applicable."""
            b = self._backend
            ncols = int(x.shape[1]) if x.ndim == 2 else 1
            m, k = self.shape if which == 'fwd' else self.shape[::-1]
            nnz = int(self.values.size)
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            if not (2 <= ncols <= 32 and m > 0 and k > 0 and nnz * ncols >= b.il_min_work and (ar != 0.0 or ai != 0.0)):
                return False
            info = self._packed(which)
            if info is None:
                return False
            b.ccsrmm_packed(y, (m, k), nnz, info, self.long_thresh, x, alpha, beta)
            return True

        def forward(self, y, x, alpha=1, beta=0):
            assert x.dtype == _C64, "Bad dtype: expected compelx64, got %s" % x.dtype
            assert y.dtype == _C64, "Bad dtype: expected compelx64, got %s" % y.dtype
            assert self.values.dtype == _C64
            if self._blocks is not None:
                for i, (c0, c1, blk) in enumerate(self._blocks):
                    blk.forward(y, x[c0:c1, :], alpha=alpha, beta=beta if i == 0 else 1)
                return
            if self._packed_product('fwd', y, x, alpha, beta):
                return
            self._backend.ccsrmm(y, self.shape, self.colInds, self.rowPtrs, self.values, x, alpha=alpha, beta=beta,
                                 adjoint=False, exwrite=True)

        def adjoint(self, y, x, alpha=1, beta=0):
            assert x.dtype == _C64, "Bad dtype: expected compelx64, got %s" % x.dtype
            assert y.dtype == _C64, "Bad dtype: expected compelx64, got %s" % y.dtype
            assert self.values.dtype == _C64
            b = self._backend
            if self._blocks is not None:
                for c0, c1, blk in self._blocks:
                    blk.adjoint(y[c0:c1, :], x, alpha=alpha, beta=beta)
                return
            if self._exwrite or not self._use_stored_adjoint():
                return b.ccsrmm(y, self.shape, self.colInds, self.rowPtrs, self.values, x,
                                alpha=alpha, beta=beta, adjoint=True, exwrite=self._exwrite)
            if self._packed_product('adj', y, x, alpha, beta):
                return
            t_ptr, t_ind, t_val = self._stored_adjoint()
            b.ccsrmm(y, self.shape[::-1], t_ind, t_ptr, t_val, x, alpha=alpha, beta=beta, adjoint=False, exwrite=True)

    class B200Backend(Base):
        dndarray = B200Array
        csr_matrix = B200Csr
        stored_adjoints = 'auto'        # True / False / 'auto': keep A^H in CSR for non-exclusive-write matrices
        il_min_work = 1 << 14           # nnz*ncols from which multi-column products take the interleaved path
        packed_real = True              # real-valued matrices: packed 8-byte entries for multi-column products

        def __init__(self, device_id=0, lib=None):
            super().__init__(device_id)
            import torch
            self._torch = torch
            if not torch.cuda.is_available():
                raise RuntimeError("B200Backend needs a CUDA device (torch.cuda.is_available() is False); "
                                   "there is no CPU fallback")
            self._lib = lib if lib is not None else _lib.load()
            self._device = torch.device('cuda', int(device_id))
            torch.cuda.set_device(self._device)
            self._plans = {}
            self._cg_scal = None
            self._il_bufs = {}

        def __del__(self):
            try:
                for plan in self._plans.values():
                    self._lib.fft_plan_destroy(plan)
                self._plans.clear()
            except Exception:
                pass

        # ------------------------------------------------------------ plumbing
        @property
        def _stream(self):
            """Current torch stream of this backend's device.  Every primitive fetches it first, so this is
            also where the device is made current: the C layer resolves per-device state (workspaces, SM
            count, plans' twiddles) through cudaGetDevice, and a second backend on another GPU or a user's
            torch.cuda.set_device must not redirect this one's launches."""
            cuda = self._torch.cuda
            if cuda.current_device() != self._device.index:
                cuda.set_device(self._device)
            return cuda.current_stream(self._device).cuda_stream

        def _pinned(self, arr):
            return getattr(arr, '_b200_pinned', False)

        def pinned_array(self, shape, dtype=_C64):
            """Page-locked host ndarray (column-major) for asynchronous H2D/D2H copies."""
            n = int(np.prod(shape))
            t = self._torch.empty(max(n, 1) * np.dtype(dtype).itemsize, dtype=self._torch.uint8).pin_memory()
            arr = t.numpy().view(dtype)[:n].reshape(shape, order='F')

            class _Pinned(np.ndarray):
                pass
            out = arr.view(_Pinned)
            out._b200_pinned = True
            out._b200_keep = t
            return out

        def mapped_array(self, pinned):
            """Device-addressable view of a page-locked host array from `pinned_array` (no copy): under unified
            virtual addressing the kernels read / write the host buffer over PCIe directly, so that an operator
            evaluated on such views moves its input and output inside its first and last kernel instead of
            through separate copies.  Only for arrays that are touched once per evaluation (operator inputs and
            outputs); everything else belongs in HBM."""
            if not self._pinned(pinned):
                raise ValueError("mapped_array needs an array from pinned_array()")
            if not pinned.flags['F_CONTIGUOUS']:
                raise ValueError("mapped_array needs a column-major array")
            return self.dndarray(self, pinned.shape, pinned.dtype, own=False,
                                 data=DevPtr(pinned.ctypes.data, keep=pinned), name='mapped')

        if hasattr(Base, 'NUFFT'):
            def NUFFT(self, M, N, coord, width=3, n=128, oversamp=None, dtype=_C64, **kwargs):
                """Backend.NUFFT (backend.py:393-450) unchanged; the returned product remembers its arguments so
                that indigo_b200.fused.fuse_transform can swap the SENSE tree it ends up in for the fused node."""
                from .fused import tag_nufft
                op = super().NUFFT(M, N, coord, width=width, n=n, oversamp=oversamp, dtype=dtype, **kwargs)
                return tag_nufft(op, N, coord, width, n, oversamp)

        def barrier(self):
            self._lib.stream_sync(self._stream)

        def get_max_threads(self):
            return 1

        # ------------------------------------------------------------ BLAS-1  (backend.py:453-467)
        def axpby(self, beta, y, alpha, x):
            """y = beta*y + alpha*x"""
            assert isinstance(x, Base.dndarray) and isinstance(y, Base.dndarray)
            (br, bi), (ar, ai) = _re_im(beta), _re_im(alpha)
            ys, xs = y._columns(), x._columns()
            if len(ys) != len(xs):                 # one side contiguous, the other a pitched view
                n = max(len(ys), len(xs))
                ys = ys if len(ys) == n else [(ys[0][0] + 8 * i * (ys[0][1] // n), ys[0][1] // n) for i in range(n)]
                xs = xs if len(xs) == n else [(xs[0][0] + 8 * i * (xs[0][1] // n), xs[0][1] // n) for i in range(n)]
            for (yp, n), (xp, _) in zip(ys, xs):
                self._lib.caxpby(self._stream, n, br, bi, yp, ar, ai, xp)

        def scale(self, x, alpha):
            """x *= alpha (alpha == 0 is a memset, never a read: SURVEY appendix A)"""
            ar, ai = _re_im(alpha)
            for p, n in x._columns():
                self._lib.cscal(self._stream, n, ar, ai, p)

        def dot(self, x, y):
            """Re(x^H y) as a host float (np.py:60-64)."""
            re, im = ctypes.c_double(), ctypes.c_double()
            total = 0.0
            for (xp, n), (yp, _) in zip(x._columns(), y._columns()):
                self._lib.cdotc(self._stream, n, xp, yp, ctypes.byref(re), ctypes.byref(im))
                total += re.value
            return total

        def norm2(self, x):
            """||x||^2 (squared, np.py:66-69) as a host float."""
            out = ctypes.c_double()
            total = 0.0
            for p, n in x._columns():
                self._lib.scnrm2sq(self._stream, n, p, ctypes.byref(out))
                total += out.value
            return total

        # ------------------------------------------------------------ dense  (backend.py:481-491)
        def cgemm(self, y, M, x, alpha, beta, forward):
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            (m, n), k = y.shape, x.shape[0]
            self._lib.cgemm(self._stream, 0 if forward else 1, m, n, k, ar, ai, M.ptr, M.ld, x.ptr, x.ld,
                            br, bi, y.ptr, y.ld)

        def csymm(self, y, M, x, alpha, beta, left=True):
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            m, n = y.shape
            self._lib.csymm(self._stream, 1 if left else 0, m, n, ar, ai, M.ptr, M.ld, x.ptr, x.ld,
                            br, bi, y.ptr, y.ld)

        # ------------------------------------------------------------ FFT  (backend.py:497-512)
        def _fft_plan(self, shape):
            shape = tuple(int(s) for s in shape)
            plan = self._plans.get(shape)
            if plan is None:
                dims = (ctypes.c_int64 * (len(shape) - 1))(*shape[:-1])
                plan = ctypes.c_void_p()
                self._lib.fft_plan_create(ctypes.byref(plan), len(shape) - 1, dims, shape[-1])
                self._plans[shape] = plan
            return plan

        def _fft_workspace_size(self, x_shape):
            return 0

        def fftn(self, y, x):
            """Unscaled forward FFT over all but the last axis."""
            self._lib.fft_exec(self._fft_plan(x.shape), self._stream, y.ptr, x.ptr, -1)

        def ifftn(self, y, x):
            """Unscaled inverse FFT (numpy's ifftn times prod(shape), np.py:109-115)."""
            self._lib.fft_exec(self._fft_plan(x.shape), self._stream, y.ptr, x.ptr, +1)

        # ------------------------------------------------------------ sparse  (backend.py:514-533)
        def _il_scratch(self, which, nelem):
            """Backend-owned buffers for the coil-interleaved copies of X and Y (grown on demand,
            never shrunk; they are not part of indigo's scratch arena)."""
            buf = self._il_bufs.get(which)
            if buf is None or buf.numel() < nelem * 8:
                self._il_bufs[which] = None
                buf = self._torch.empty(max(nelem, 1) * 8, dtype=self._torch.uint8, device=self._device)
                self._il_bufs[which] = buf
            return buf.data_ptr()

        def ccsrmm(self, y, A_shape, A_indx, A_ptr, A_vals, x, alpha=1, beta=0, adjoint=False, exwrite=False):
            if A_indx.dtype != np.int32 or A_ptr.dtype != np.int32:
                raise ValueError("b200 ccsrmm needs int32 indices, got %s" % A_indx.dtype)
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            m, k = (int(v) for v in A_shape)
            ncols, nnz = int(x.shape[1]), int(A_vals.size)
            if (not adjoint and 2 <= ncols <= 32 and m > 0 and k > 0 and nnz * ncols >= self.il_min_work
                    and (ar != 0.0 or ai != 0.0)):
                # multi-coil gather: transpose X to [row][coil], gather whole coil segments, transpose back
                s = self._stream
                xil = self._il_scratch('x', k * ncols)
                yil = self._il_scratch('y', m * ncols)
                self._lib.interleave(s, k, ncols, x.ptr, x.ld, xil, ncols)
                self._lib.ccsrmm_il(s, m, k, ncols, nnz, ar, ai, A_vals.ptr, A_indx.ptr, A_ptr.ptr, xil, ncols, yil, ncols,
                                    None, 0, None, 0, 0)
                self._lib.deinterleave(s, m, ncols, yil, ncols, br, bi, y.ptr, y.ld)
                return
            self._lib.ccsrmm(self._stream, 1 if adjoint else 0, 1 if exwrite else 0, m, k, ncols,
                             nnz, ar, ai, A_vals.ptr, A_indx.ptr, A_ptr.ptr, x.ptr, x.ld,
                             br, bi, y.ptr, y.ld)

        def ccsrmm_packed(self, y, A_shape, nnz, info, long_thresh, x, alpha=1, beta=0):
            """Y = alpha*A*X + beta*Y for a real-valued A held as packed (int32 column, float weight) entries
            (B200Csr._packed): interleave -> staged packed gather (+ one CTA per long row) -> deinterleave."""
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            m, k = (int(v) for v in A_shape)
            ncols, s = int(x.shape[1]), self._stream
            xil = self._il_scratch('x', k * ncols)
            yil = self._il_scratch('y', m * ncols)
            self._lib.interleave(s, k, ncols, x.ptr, x.ld, xil, ncols)
            self._lib.ccsrmm_ilr(s, m, k, ncols, nnz, ar, ai, info['pk'].ptr, info['ptr'].ptr, xil, ncols, yil, ncols,
                                 None, -4, info['long'].ptr if info['nlong'] else None, info['nlong'], long_thresh)
            self._lib.deinterleave(s, m, ncols, yil, ncols, br, bi, y.ptr, y.ld)

        def cdiamm(self, y, shape, offsets, data, x, alpha=1.0, beta=0.0, adjoint=True):
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            m, k = (int(v) for v in shape)
            if offsets.dtype != np.int32:
                raise ValueError("b200 cdiamm needs int32 offsets, got %s" % offsets.dtype)
            # data is (diagonal length x noffsets), column-major: its own row count is the diagonal length
            # scipy chose (not necessarily k) and its leading dimension the pitch between diagonals
            self._lib.cdiamm(self._stream, 1 if adjoint else 0, m, k, int(x.shape[1]), int(offsets.size),
                             offsets.ptr, data.ptr, int(data.shape[0]), data.ld, ar, ai, x.ptr, x.ld, br, bi, y.ptr, y.ld)

        def onemm(self, y, x, alpha=1, beta=0):
            (ar, ai), (br, bi) = _re_im(alpha), _re_im(beta)
            (k, n), m = x.shape, y.shape[0]
            self._lib.onemm(self._stream, m, n, k, ar, ai, x.ptr, x.ld, br, bi, y.ptr, y.ld)

        def max(self, val, arr):
            """arr = max(arr, val) on real and imaginary parts (np.py:141-145)."""
            for p, n in arr._columns():
                self._lib.fmax(self._stream, 2 * n, float(val), p)

        class dia_matrix(object):
            """Device DIA holder of Backend.dia_matrix (backend.py:599-633): `data` is scipy's dia.data
            transposed, (diagonal length x noffsets) column-major, whatever length scipy chose."""

            def __init__(self, backend, A, name='mat'):
                assert isinstance(A, spp.dia_matrix)
                A = A.astype(np.complex64)
                self._backend = backend
                self.data = backend.copy_array(np.asfortranarray(A.data.T), name=name + ".data")
                self.offsets = backend.copy_array(np.ascontiguousarray(A.offsets, dtype=np.int32), name=name + ".offsets")
                self.shape, self.dtype = A.shape, A.dtype
                self._row_frac = self._col_frac = 1

            nbytes = property(lambda self: self.offsets.nbytes + self.data.nbytes)
            nnz = property(lambda self: self.data.size)

            def forward(self, y, x, alpha=1, beta=0):
                self._backend.cdiamm(y, self.shape, self.offsets, self.data, x, alpha=alpha, beta=beta, adjoint=False)

            def adjoint(self, y, x, alpha=1, beta=0):
                self._backend.cdiamm(y, self.shape, self.offsets, self.data, x, alpha=alpha, beta=beta, adjoint=True)

        # ------------------------------------------------------------ solvers
        def pdot(self, x, y, comm):
            v = self.dot(x, y)
            if comm is not None and not getattr(comm, 'replicated_vectors', False):
                v = comm.allreduce(v)
            return v

        def pnorm2(self, x, comm):
            v = self.norm2(x)
            if comm is not None and not getattr(comm, 'replicated_vectors', False):
                v = comm.allreduce(v)
            return v

        def cg(self, A, b_h, x_h, lamda=0.0, tol=1e-10, maxiter=100, team=None, iterates=None, graph=False):
            """Conjugate gradient on (A + lamda I) x = b with the update order of
            backend.py:639-689, but with the five BLAS-1 passes and two blocking scalar
            read-backs of one iteration fused into three kernels whose scalars stay on the
            device (ib200_cdotc_dev / ib200_cg_xr / ib200_cg_p).  With a coil-sharded
            operator, `team.allreduce_array` sums the partial A*p over ranks (NCCL) and
            the replicated vectors make every rank's scalars bit-identical, so no scalar
            all-reduce is needed.

            graph=True (needs tol <= 0 and no `iterates`): one iteration -- the apply, the collective and the
            three solver kernels -- is captured in a CUDA graph after a first eager iteration and replayed
            maxiter - 1 times, one launch per iteration (the launch-bound regime, cfg1)."""
            sums_image = team is not None and hasattr(team, 'allreduce_array')
            replicated = team is None or bool(getattr(team, 'replicated_vectors', False))
            if team is not None and not (sums_image and replicated):
                # a reference-style team (only `allreduce(scalar)`) or partial vectors: every scalar has to go
                # through pdot / pnorm2 exactly as backend.py:661-677 does, on the host
                return self._cg_host_scalars(A, b_h, x_h, lamda, tol, maxiter, team, iterates)
            lib, s = self._lib, self._stream
            x, b = self.copy_array(x_h, name='x'), self.copy_array(b_h, name='b')
            n = int(x.size)
            Ap = x.copy()
            r = b

            def apply(out, inp):
                A.eval(out, inp)
                if sums_image:
                    team.allreduce_array(out)

            apply(Ap, x)
            self.axpby(1, r, -1, Ap)
            self.axpby(1, r, -lamda, x)
            p = r.copy(name='p')
            if self._cg_scal is None:
                self._cg_scal = self._torch.zeros(4, dtype=self._torch.float64, device=self._device)
            scal = self._cg_scal
            sp = scal.data_ptr()
            lib.scnrm2sq_dev(s, n, r.ptr, sp)                         # scal[0] = rr
            r0 = None
            if tol > 0:
                r0 = float(scal[0].item())

            def iteration():
                st = self._stream
                apply(Ap, p)
                if lamda != 0:
                    self.axpby(1, Ap, lamda, p)
                lib.cdotc_dev(st, n, p.ptr, Ap.ptr, sp + 8)           # scal[1..2] = p^H Ap
                lib.cg_xr(st, n, x.ptr, r.ptr, p.ptr, Ap.ptr, sp)     # x, r, scal[3] = ||r||^2
                lib.cg_p(st, n, p.ptr, r.ptr, sp)                     # p, scal[0] = scal[3]

            if graph and tol <= 0 and iterates is None and maxiter > 2:
                torch = self._torch
                iteration()                                            # eager: first-use allocations and attributes
                side = torch.cuda.Stream(device=self._device)
                side.wait_stream(torch.cuda.current_stream(self._device))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        iteration()
                torch.cuda.current_stream(self._device).wait_stream(side)
                for it in range(maxiter - 1):                          # capturing records the iteration, it does not run it
                    g.replay()
                log.info("cg reached maxiter (graph replay)")
                x.copy_to(x_h)
                return
            for it in range(maxiter):
                iteration()
                if iterates is not None:
                    iterates.append(x.to_host())
                if tol > 0:
                    resid = np.sqrt(float(scal[0].item()) / r0)
                    log.info("iter %d, residual %g", it, resid)
                    if resid < tol:
                        log.info("cg reached tolerance")
                        break
            else:
                log.info("cg reached maxiter")
            x.copy_to(x_h)

        def _cg_host_scalars(self, A, b_h, x_h, lamda, tol, maxiter, team, iterates=None):
            """CG with the scalar protocol of the reference (backend.py:639-689): ||r||^2 and p^H A p are host
            floats obtained through pnorm2 / pdot, which sum them over `team` unless its vectors are replicated.
            Used whenever the device-resident-scalar loop of cg() cannot honour the team's contract."""
            x, r = self.copy_array(x_h, name='x'), self.copy_array(b_h, name='b')
            Ap = x.copy()
            sums_image = hasattr(team, 'allreduce_array') and bool(getattr(team, 'replicated_vectors', False))

            def apply(out, inp):
                A.eval(out, inp)
                if sums_image:
                    team.allreduce_array(out)

            apply(Ap, x)
            self.axpby(1, r, -1, Ap)
            self.axpby(1, r, -lamda, x)
            p = r.copy(name='p')
            rr = r0 = self.pnorm2(r, team)
            for it in range(maxiter):
                apply(Ap, p)
                self.axpby(1, Ap, lamda, p)
                step = rr / self.pdot(p, Ap, team)
                self.axpby(1, x, step, p)
                self.axpby(1, r, -step, Ap)
                rr_new = self.pnorm2(r, team)
                self.scale(p, rr_new / rr)
                self.axpby(1, p, 1, r)
                rr = rr_new
                if iterates is not None:
                    iterates.append(x.to_host())
                resid = np.sqrt(rr / r0) if r0 else 0.0
                log.info("iter %d, residual %g", it, resid)
                if resid < tol:
                    log.info("cg reached tolerance")
                    break
            else:
                log.info("cg reached maxiter")
            x.copy_to(x_h)

    B200Backend.__name__ = name
    B200Backend.__qualname__ = name
    return B200Backend


B200Backend = make_backend_class(StandaloneBase)
