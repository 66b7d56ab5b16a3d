"""
Standalone base of B200Backend: what the backend needs from its base class when the
reference package `indigo` is NOT importable (the GPU box; users who only want the fused
SENSE operator and the primitives).

With `indigo` installed, `indigo_b200.register()` builds the backend on the reference's own
`indigo.backends.backend.Backend` instead, and everything here is unused: arrays, builders,
the scratch arena, `apgd` and the operator tree then all come from the unmodified reference.

This module is written against the interface contract in SURVEY.md section 8(b)/(a8), not against
the reference's source: it provides ONLY
  * `DeviceArray`   column-major device array with the fields the C-ABI wrappers use
                    (`shape`, `dtype`, `_leading_dim`, `_own`, `_arr`) and the view / transfer
                    protocol of the reference's `dndarray` (backend.py:22-220), so that the
                    same `B200Array` hooks (`_malloc`, `_copy_from`, ...) serve both bases;
  * `StandaloneBase` array factories (names of backend.py:222-241) and nothing else: no operator
                    builders, no tree, no arena.  Operators of the standalone build are the three
                    classes of indigo_b200/linop.py.
"""
from contextlib import contextmanager

import numpy as np


def _complete_shape(shape, size):
    """Resolves one `-1` entry of `shape` against an element count."""
    shape = tuple(int(s) for s in shape)
    free = [i for i, s in enumerate(shape) if s == -1]
    if not free:
        return shape
    if len(free) > 1:
        raise ValueError("at most one dimension may be -1, got %r" % (shape,))
    rest = 1
    for i, s in enumerate(shape):
        if i != free[0]:
            rest *= s
    if rest == 0 or size % rest:
        raise AssertionError("Cannot reshape {} elements into {}. (size mismatch)".format(size, shape))
    return shape[:free[0]] + (size // rest,) + shape[free[0] + 1:]


class DeviceArray(object):
    """Column-major array in device memory.

    `_arr` is the memory handle (for B200Array a `ctypes.c_ulong` subclass holding the device
    address and a reference to the owning torch tensor); views made by `reshape` / slicing share
    it and never free it (`_own` False).  `_leading_dim` is the distance, in elements, between two
    columns of a 2-D view: slices of a larger allocation keep the parent's pitch, which is what
    the `ld*` arguments of the C ABI receive."""

    def __init__(self, backend, shape, dtype, ld=None, own=True, data=None, name=''):
        if not isinstance(shape, (tuple, list)):
            raise AssertionError("shape must be a tuple or list, got %r" % (shape,))
        self._backend = backend
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self._leading_dim = int(ld) if ld else (self.shape[0] if self.shape else 1)
        self._name = name
        self._own = bool(own) and data is None
        self._arr = self._malloc(self.shape, self.dtype) if data is None else data

    # ---- geometry -----------------------------------------------------------------------
    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= int(s)
        return n

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def contiguous(self):
        return self.ndim < 2 or self._leading_dim == self.shape[0]

    def _view(self, shape, ld):
        return self._backend.dndarray(self._backend, tuple(shape), dtype=self.dtype, ld=ld, own=False, data=self._arr)

    def reshape(self, new_shape):
        """View of the same memory under another shape (backend.py:59-89 semantics): more rows than
        before means columns get stacked, which needs them to be adjacent in memory; fewer rows
        re-bases the pitch on the new row count; otherwise the pitch is inherited."""
        new_shape = _complete_shape(new_shape, self.size)
        count = 1
        for s in new_shape:
            count *= s
        assert count == self.size, "Cannot reshape {} into {}. (size mismatch)".format(self.shape, new_shape)
        rows_now, rows_new = self.shape[0], new_shape[0]
        if rows_new > rows_now:
            assert self._leading_dim == rows_now, "Cannot stack non-contiguous columns."
        return self._view(new_shape, rows_new if rows_new < rows_now else self._leading_dim)

    # ---- transfers ----------------------------------------------------------------------
    def _match_host(self, arr, column_major):
        assert isinstance(arr, np.ndarray)
        if arr.size != self.size:
            raise ValueError("size mismatch, expected {} got {}".format(self.shape, arr.shape))
        if arr.dtype != self.dtype:
            raise TypeError("dtype mismatch, expected {} got {}".format(self.dtype, arr.dtype))
        if column_major and not arr.flags['F_CONTIGUOUS']:
            raise TypeError("order mismatch, expected 'F' got {}".format(arr.flags['F_CONTIGUOUS']))

    def copy_from(self, arr):
        """host -> device; the host array must be column-major with this dtype and element count."""
        self._match_host(arr, column_major=True)
        self._copy_from(arr)

    def copy_to(self, arr):
        """device -> host."""
        self._match_host(arr, column_major=False)
        self._copy_to(arr)

    def to_host(self):
        out = np.empty(self.shape, dtype=self.dtype, order='F')
        self.copy_to(out)
        return out

    @contextmanager
    def on_host(self):
        """`with d.on_host() as h:` edit a host copy; it is written back on exit."""
        staged = self.to_host()
        yield staged
        self.copy_from(staged)

    def copy(self, other=None, name=''):
        """`a.copy(b)` overwrites a with b on the device; `a.copy()` returns a new array equal to a."""
        if other is not None and other is not False:
            assert isinstance(other, self._backend.dndarray)
            self._copy(other)
            return None
        twin = self._backend.zero_array(self.shape, self.dtype, name=name)
        twin._copy(self)
        return twin

    @classmethod
    def to_device(cls, backend, arr, name=''):
        host = np.asfortranarray(arr)
        out = cls(backend, arr.shape, arr.dtype, name=name)
        out.copy_from(host)
        return out

    def __setitem__(self, slc, other):
        assert not (slc.start or slc.stop), "dndarray setitem cant slice"
        self._copy(other)

    def __del__(self):
        if getattr(self, '_own', False) and getattr(self, '_arr', None) is not None:
            self._free()

    # ---- supplied by the concrete array (B200Array in backend.py) --------------------------
    def _missing(self, *a, **k):
        raise NotImplementedError("%s does not implement this array hook" % type(self).__name__)

    __getitem__ = _malloc = _free = _zero = _copy = _copy_from = _copy_to = _missing
    from_param = staticmethod(_missing)


class StandaloneBase(object):
    """Array factories of the Backend interface (backend.py:222-241) for the standalone build."""

    dndarray = DeviceArray
    ops = None                    # set to indigo_b200.linop below

    def __init__(self, device_id=0):
        self.device_id = int(device_id)

    def empty_array(self, shape, dtype, name=''):
        return self.dndarray(self, shape, dtype, name=name)

    def zero_array(self, shape, dtype, name=''):
        out = self.empty_array(shape, dtype, name=name)
        out._zero()
        return out

    def zeros_like(self, other, name=''):
        return self.zero_array(other.shape, other.dtype, name=name)

    def copy_array(self, arr, name=''):
        return self.dndarray.to_device(self, arr, name=name)

    def rand_array(self, shape, dtype=np.dtype('complex64'), name=''):
        draw = np.random.random(shape) + 1j * np.random.random(shape)
        return self.copy_array(np.asfortranarray(draw.astype(np.complex64)), name=name)

    def get_max_threads(self):
        return 1

    def barrier(self):
        pass


from . import linop as _linop      # noqa: E402  (linop has no dependency on this module)
StandaloneBase.ops = _linop
