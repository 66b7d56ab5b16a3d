"""
TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of how the reference builds
the non-Cartesian SENSE operator and of the `-O3` tree its own rewrites turn it
into.  See oracle/README.md.  Nothing under indigo_b200/
may import this module.

Parity: PINNED.  tests/test_oracle.py compares the matrices built here with the
ones the unmodified reference builds (tests/golden/sense_small.npz: CSR
structure AND values bit-identical) and the applies / CG iterates with the
reference NumpyBackend's.

Reference construction being restated
  A    = KronI(C, NUFFT) * VStack_c Diag(maps_c)          examples/pics.py:92-95
  NUFFT= G * FFTc * Zpad * Diag(rolloff3)                 backend.py:403-442
  FFTc = Mod * (Scale * UnscaledFFT) * Mod                backend.py:347-369
  -O2  : G' = G @ (Mod @ Scale) ; P = kron(I_C,(Mod @ Zpad) @ Apod) @ vstack(maps)
                                                          pics.py:111-126, transforms.py:86-146
  -O3  : P is stored as its adjoint P^H                    pics.py:104-109
  apply: ccsrmm(P^H,adj) -> fftn -> ccsrmm(G') -> ccsrmm(G',adj) -> ifftn -> ccsrmm(P^H)
"""
import numpy as np
import scipy.sparse as spp
from scipy.signal.windows import kaiser

from . import np_oracle as K

C64 = np.dtype("complex64")


# ---------------------------------------------------------------------------
# pieces of Backend.NUFFT
# ---------------------------------------------------------------------------
def oversampled_shape(N, oversamp):
    """backend.py:416-430: scalar oversamp is broadcast; oN_i = int(N_i*os_i)."""
    if isinstance(oversamp, tuple):
        omin, os3 = min(oversamp), tuple(oversamp)
    else:
        omin, os3 = oversamp, (oversamp,) * 3
    oN = tuple(int(n * o) for n, o in zip(N, os3))
    return oN, omin


def kb_table(width, omin, n=128):
    """Kaiser-Bessel shape parameter and half-window lookup table.  backend.py:435-436."""
    beta = np.pi * np.sqrt(((width * 2.0 / omin) * (omin - 0.5)) ** 2 - 0.8)
    return beta, kaiser(2 * n + 1, beta)[n:]


def flat_coord(coord):
    """(3, d1, d2, ...) -> (3, npts) in COLUMN-major sample order, as Backend.Interp
    flattens it (backend.py:396: coord.reshape((ndim,-1), order='F'))."""
    coord = np.asarray(coord)
    return coord.reshape((coord.shape[0], -1), order="F")


def _table_lookup(table, x):
    """interp.py:9-15 (lin_interp), vectorised; x >= 1 gives 0."""
    n = len(table)
    inside = x < 1
    xs = np.where(inside, x, 0.0) * (n - 1)
    idx = xs.astype(np.int64)
    frac = xs - idx
    val = (1.0 - frac) * table[idx] + frac * table[np.minimum(idx + 1, n - 1)]
    return np.where(inside, val, 0.0)


def interp_coo(oN, coord, width, table, chunk=1 << 16):
    """Gridding matrix in the reference's COO emission order.  interp.py:19-60.

    coord is (3, npts), in cycles/FOV in [-1/2, 1/2).  Taps per axis are
    range(ceil(pos-width), floor(pos+width)) (5 off-grid, 6 on-grid for width 3),
    wrapped modulo the grid; weights multiply as (wz*wy)*wx in float64."""
    coord = flat_coord(coord).astype(np.float64)
    m = coord.shape[1]
    T = int(2 * width + 1)                     # upper bound on taps per axis (interp.py:21)
    rows, cols, kers = [], [], []
    a = np.arange(T)
    for lo in range(0, m, chunk):
        hi = min(m, lo + chunk)
        taps, wts, cnt = [], [], []
        for d in range(3):
            pos = oN[d] * coord[d, lo:hi] + (oN[d] // 2)
            start = np.ceil(pos - width).astype(np.int64)
            end = np.floor(pos + width).astype(np.int64)
            t = start[:, None] + a[None, :]
            taps.append(t)
            wts.append(_table_lookup(table, np.abs(t - pos[:, None]) / width))
            cnt.append((end - start)[:, None] > a[None, :])
        valid = cnt[2][:, :, None, None] & cnt[1][:, None, :, None] & cnt[0][:, None, None, :]
        wz = wts[2][:, :, None, None]
        wy = wz * wts[1][:, None, :, None]
        w = wy * wts[0][:, None, None, :]
        jz = (taps[2] % oN[2]) * (oN[1] * oN[0])
        jy = (taps[1] % oN[1])[:, None, :] * oN[0] + jz[:, :, None]
        j = (taps[0] % oN[0])[:, None, None, :] + jy[:, :, :, None]
        i = np.broadcast_to(np.arange(lo, hi)[:, None, None, None], valid.shape)
        rows.append(i[valid]); cols.append(j[valid]); kers.append(np.broadcast_to(w, valid.shape)[valid])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(kers)


def interp_matrix(oN, coord, width, table):
    """Backend.Interp with dtype=float32 as NUFFT calls it.  backend.py:392-401,437; interp.py:63-80."""
    coord = flat_coord(coord)
    r, c, k = interp_coo(oN, coord, width, table)
    M = spp.coo_matrix((k, (r, c)), shape=(coord.shape[1], int(np.prod(oN))))
    return M.astype(np.float32)


def _ftkb(beta, x):
    """noncart.py:5-14."""
    a = np.sqrt(beta ** 2 - (np.pi * x) ** 2)
    out = np.empty(a.shape, dtype=a.dtype)
    out[a == 0.0] = 1.0
    b = a[a != 0.0]
    out[a != 0.0] = np.sinh(b) / b
    return out


def rolloff3(oversamp, width, beta, N):
    """Kaiser-Bessel apodisation correction.  noncart.py:17-23."""
    x, y, z = np.mgrid[:N[0], :N[1], :N[2]]
    s = width * 2.0 / oversamp
    return _ftkb(beta, 0.0) ** 3 / (_ftkb(beta, (x - N[0] // 2) / N[0] * s) *
                                    _ftkb(beta, (y - N[1] // 2) / N[1] * s) *
                                    _ftkb(beta, (z - N[2] // 2) / N[2] * s))


def diag_matrix(v, dtype=C64):
    """Backend.Diag: column-major flatten, DIA matrix cast to dtype.  backend.py:298-305."""
    v = np.require(v, requirements="F")
    if v.ndim > 1:
        v = v.flatten(order="A")
    return spp.diags(v, offsets=0).astype(dtype)


def zpad_matrix(big, small, dtype=C64):
    """Centred zero-pad selection matrix (prod(big) x prod(small)).  backend.py:371-387."""
    slc = tuple(slice(m // 2 + int(np.ceil(-n / 2)), m // 2 + int(np.ceil(n / 2)))
                for m, n in zip(big, small))
    lin = np.arange(int(np.prod(big)), dtype=int).reshape(big, order="F")
    rows = lin[slc].flatten(order="F")
    cols = np.arange(rows.size)
    return spp.coo_matrix((np.ones_like(cols), (rows, cols)),
                          shape=(int(np.prod(big)), int(np.prod(small))), dtype=dtype)


def fftc_mod(shape, dtype=C64):
    """Centring phase ramp of FFTc.  backend.py:357-363."""
    idx = np.mgrid[tuple(slice(d) for d in shape)]
    ph = 0
    for i, n in enumerate(shape):
        c = n // 2
        ph = ph + (idx[i] - c / 2.0) * (c / n)
    return np.exp(1j * 2.0 * np.pi * ph).astype(dtype)


def _device_csr(M):
    """What SpMatrix._get_or_create_device_matrix hands to csr_matrix:
    complex64, CSR, sorted indices.  operators.py:222-234."""
    M = M.astype(np.complex64).tocsr()
    M.sort_indices()
    return M


# ---------------------------------------------------------------------------
# the -O3 SENSE operator
# ---------------------------------------------------------------------------
class SenseOperator:
    """A = KronI(C, NUFFT) * VStack(Diag(maps_c)) after the reference's -O3 recipe.

    Attributes: G (M x oN csr c64), PH (N x C*oN csr c64, the stored adjoint of
    P), N, oN, C, M.  maps is (N0,N1,N2,C); coord is (3, npts...)."""

    def __init__(self, N, coord, maps, oversamp=2.0, width=3, n=128, weights=None):
        N = tuple(int(v) for v in N)
        self.N, self.C = N, int(maps.shape[3])
        self.oN, omin = oversampled_shape(N, oversamp)
        beta, table = kb_table(width, omin, n)
        self.beta = beta
        coord = flat_coord(coord)
        self.M = coord.shape[1]
        on = int(np.prod(self.oN))

        G = interp_matrix(self.oN, coord, width, table)                    # float32 COO
        mod = diag_matrix(fftc_mod(self.oN))                                # 'mod'
        scl = diag_matrix(np.ones(on, order="F", dtype=C64) / np.sqrt(on))  # 'scale' (backend.py:349-351)
        Z = zpad_matrix(self.oN, N)                                         # 'zpad'
        R = diag_matrix(rolloff3(omin, width, beta, N))                     # 'apod'
        if weights is not None:
            # optional sqrt-DCF row weighting, Diag(w) * G (test_compat.py:185 style)
            G = diag_matrix(np.asarray(weights).reshape(-1)) @ G
        # -O2 realisation order (pics.py:111-126 through transforms.py:86-96)
        Gp = G @ (mod @ scl)                                                # 'interp*mod*scale' (right-leaning)
        MZR = (mod @ Z) @ R                                                 # 'mod*zpad*apod'
        S = spp.vstack([diag_matrix(maps[:, :, :, c:c + 1]) for c in range(self.C)], dtype=C64)
        P = spp.kron(spp.eye(self.C, dtype=C64), MZR) @ S
        self.G = _device_csr(Gp)
        self.PH = _device_csr(P.conjugate().transpose())                    # -O3: stored adjoint
        self.shape = (self.M * self.C, int(np.prod(N)))

    # --- the six backend calls of SURVEY.md section 3.1 ---------------------
    def _expand(self, x):
        on = int(np.prod(self.oN))
        t = np.zeros((self.PH.shape[1], 1), dtype=C64, order="F")
        K.ccsrmm(t, self.PH.shape, self.PH.indices, self.PH.indptr, self.PH.data,
                 x.reshape(-1, 1, order="F"), 1, 0, adjoint=True, exwrite=True)
        return t.reshape((on, self.C), order="F")

    def _combine(self, g):
        y = np.zeros((self.PH.shape[0], 1), dtype=C64, order="F")
        K.ccsrmm(y, self.PH.shape, self.PH.indices, self.PH.indptr, self.PH.data,
                 g.reshape(-1, 1, order="F"), 1, 0, adjoint=False)
        return y

    def _fft(self, g, inverse=False):
        X = g.reshape(self.oN + (self.C,), order="F")
        Y = np.zeros_like(X, order="F")
        (K.ifftn if inverse else K.fftn)(Y, X)
        return Y.reshape((-1, self.C), order="F")

    def forward(self, x):
        """A x : image (prod(N),) -> k-space (M, C)."""
        g = self._fft(self._expand(np.asarray(x, dtype=C64)))
        y = np.zeros((self.M, self.C), dtype=C64, order="F")
        K.ccsrmm(y, self.G.shape, self.G.indices, self.G.indptr, self.G.data, g, 1, 0, adjoint=False)
        return y

    def adjoint(self, y):
        """A^H y : k-space (M, C) -> image (prod(N), 1)."""
        y = np.asarray(y, dtype=C64).reshape((self.M, self.C), order="F")
        g = np.zeros((self.G.shape[1], self.C), dtype=C64, order="F")
        K.ccsrmm(g, self.G.shape, self.G.indices, self.G.indptr, self.G.data, y, 1, 0, adjoint=True)
        return self._combine(self._fft(g, inverse=True))

    def normal(self, x):
        """A^H A x."""
        return self.adjoint(self.forward(x))

    def normal_into(self, out, inp):
        out[...] = self.normal(inp).reshape(out.shape, order="F")


class SenseTruth64:
    """complex128 evaluation of the same -O3 tree (same CSR matrices promoted to
    double, numpy FFT in double): the arbiter when the complex64 oracle and the
    CUDA path disagree, and the yardstick for the oracle's own rounding noise
    (SURVEY.md section 8c "oracle hygiene")."""

    def __init__(self, op):
        self.op = op
        self.G = op.G.astype(np.complex128)
        self.PH = op.PH.astype(np.complex128)

    def normal(self, x):
        op, C = self.op, self.op.C
        t = (self.PH.conjugate().transpose() @ np.asarray(x, dtype=np.complex128).reshape(-1, 1))
        t = np.fft.fftn(t.reshape(op.oN + (C,), order="F"), axes=(0, 1, 2)).reshape((-1, C), order="F")
        g = self.G.conjugate().transpose() @ (self.G @ t)
        g = np.fft.ifftn(g.reshape(op.oN + (C,), order="F"), axes=(0, 1, 2)) * np.prod(op.oN)
        return self.PH @ g.reshape((-1, 1), order="F")

    def cg(self, b, lamda, maxiter):
        """Backend.cg's recurrence (backend.py:666-679) in double; returns all iterates."""
        x = np.zeros(np.shape(b), dtype=np.complex128)
        r = np.asarray(b, dtype=np.complex128).copy()
        p = r.copy()
        rr = np.vdot(r, r).real
        out = []
        for _ in range(maxiter):
            Ap = self.normal(p).reshape(p.shape) + lamda * p
            a = rr / np.vdot(p, Ap).real
            x = x + a * p
            r = r - a * Ap
            r2 = np.vdot(r, r).real
            p = r + (r2 / rr) * p
            rr = r2
            out.append(x.copy())
        return out


def spectral_norm(op, iters=15, seed=0):
    """Power-iteration estimate of ||A^H A||_2 (used to scale lamda in the CG protocol)."""
    rs = np.random.RandomState(seed)
    v = (rs.rand(op.shape[1], 1) + 1j * rs.rand(op.shape[1], 1)).astype(C64)
    nv = 1.0
    for _ in range(iters):
        v = op.normal(v)
        nv = float(np.linalg.norm(v))
        v = (v / nv).astype(C64)
    return nv


def sqrt_dcf(coord):
    """sqrt(|k|) row weights of the well-conditioned CG protocol."""
    c = flat_coord(coord).astype(np.float64)
    return np.sqrt(np.sqrt((c ** 2).sum(axis=0))).astype(np.float32)
