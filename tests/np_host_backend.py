"""
TEST-ONLY backend: indigo_b200's host mirror (HostBackend) executed by the numpy
oracle.  It exists to check the mirror (operator tree, rewrites, builders, CG
driver) against the reference-generated golden vectors on a machine without a
GPU.  It is the checker, never a fallback: it lives under tests/ and nothing in
indigo_b200/ can reach it.
"""
import numpy as np

from indigo_b200.host import HostBackend, DeviceArrayBase
from oracle import np_oracle as K


class NpArray(DeviceArrayBase):
    def _malloc(self, shape, dtype):
        return np.zeros(shape, dtype, order='F')

    def _free(self):
        pass

    def _zero(self):
        self._arr[...] = 0

    def _view(self):
        return self._arr.reshape(self.shape, order='F') if self._arr.shape != tuple(self.shape) else self._arr

    def _copy_from(self, arr):
        self._view()[...] = arr.reshape(self.shape, order='F')

    def _copy_to(self, arr):
        arr[...] = self._view().reshape(arr.shape, order='F')

    def _copy(self, other):
        self._view()[...] = other._view().reshape(self.shape, order='F')

    def reshape(self, new_shape):
        out = super().reshape(new_shape)
        out._arr = self._view().reshape(out.shape, order='F')      # numpy keeps the strides of pitched views
        return out

    def __getitem__(self, slc):
        d = self._view()[slc]
        return self._backend.dndarray(self._backend, d.shape, d.dtype, ld=self._leading_dim, own=False, data=d)


class NpHostBackend(HostBackend):
    dndarray = NpArray

    def __init__(self, device_id=0):
        super().__init__(device_id)
        self.calls = []

    def _log(self, name, **kw):
        self.calls.append((name, kw))

    def axpby(self, beta, y, alpha, x):
        self._log('axpby'); K.axpby(beta, y._view(), alpha, x._view())

    def dot(self, x, y):
        return K.dot(x._view(), y._view())

    def norm2(self, x):
        return K.norm2(x._view())

    def scale(self, x, alpha):
        K.scale(x._view(), alpha)

    def cgemm(self, y, M, x, alpha, beta, forward):
        K.cgemm(y._view(), M._view(), x._view(), alpha, beta, forward=forward)

    def csymm(self, y, M, x, alpha, beta, left=True):
        K.csymm(y._view(), M._view(), x._view(), alpha, beta, left=left)

    def onemm(self, y, x, alpha, beta):
        K.onemm(y._view(), x._view(), alpha, beta)

    def fftn(self, y, x):
        self._log('fftn', shape=tuple(x.shape)); K.fftn(y._view(), x._view())

    def ifftn(self, y, x):
        self._log('ifftn', shape=tuple(x.shape)); K.ifftn(y._view(), x._view())

    def ccsrmm(self, y, A_shape, A_indx, A_ptr, A_vals, x, alpha=1, beta=0, adjoint=False, exwrite=False):
        self._log('ccsrmm', A_shape=tuple(A_shape), nnz=int(A_vals.size), x=tuple(x.shape), y=tuple(y.shape),
                  ldx=int(x._leading_dim), ldy=int(y._leading_dim), alpha=alpha, beta=beta,
                  adjoint=bool(adjoint), exwrite=int(exwrite))
        K.ccsrmm(y._view(), A_shape, A_indx._view(), A_ptr._view(), A_vals._view(), x._view(), alpha, beta,
                 adjoint=adjoint, exwrite=exwrite)

    def cdiamm(self, y, shape, offsets, data, x, alpha=1.0, beta=0.0, adjoint=True):
        K.cdiamm(y._view(), shape, offsets._view(), data._view(), x._view(), alpha, beta, adjoint=adjoint)

    def max(self, val, arr):
        K.fmax(val, arr._view())
