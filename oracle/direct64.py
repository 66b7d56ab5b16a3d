"""
TEST INFRASTRUCTURE ONLY -- float64 evaluation of single outputs of the SENSE-NUFFT operator,
for checking the CUDA path at BASELINE.json's full sizes, where neither the numpy oracle nor the
reference can hold the operator (853 M stored entries, 416^3 x 16 grid).  Nothing under
indigo_b200/ may import this module.

What is evaluated is the reference's operator itself (not the ideal NUDFT it approximates):
    A = KronI(C, G * FFTc * Zpad * Diag(apod)) * VStack_c Diag(maps_c)      examples/pics.py:92-95
with G from interp.py:19-60 (Kaiser-Bessel table lookup, taps ceil(p-w) .. floor(p+w)-1 wrapped
modulo the grid), FFTc = Mod * (Scale * UnscaledFFT) * Mod (backend.py:347-369) and the centred
zero-pad of backend.py:371-387.  Every factor is a product over the three axes, so one k-space
sample of A x is
    k[s, c] = rw_s / sqrt(prod oN) * sum_{x,y,z} e0_s[x] e1_s[y] e2_s[z] * apod*maps_c*img [x, y, z]
    e_d,s[v] = sum_taps w_d(tap) * mod_d[kappa_tap] * mod_d[v + off_d] * exp(-2 pi i kappa_tap (v + off_d) / oN_d)
and one voxel of A^H y, for y supported on a few samples, the conjugate-transposed sum.  All
arithmetic in float64 / complex128; the reference rounds its matrix entries to float32, which is
the ~1e-7 difference the 1e-5 tolerance absorbs.

Pinned by tests/test_oracle.py::test_direct64_matches_the_oracle (against oracle/sense.py, which is
pinned to the reference's golden vectors).
"""
import numpy as np

from . import sense as S


def _axis_terms(oN, N, coord, width, table):
    """Per axis d and sample s: e_d,s[v] for v in range(N_d) (complex128, shape (S, N_d))."""
    coord = np.asarray(coord, dtype=np.float64).reshape(3, -1)
    out = []
    T = int(2 * width + 1)
    a = np.arange(T)
    for d in range(3):
        n, c = oN[d], oN[d] // 2
        off = n // 2 + int(np.ceil(-N[d] / 2))                                 # backend.py:379-381
        pos = n * coord[d] + c                                                 # interp.py:27
        start = np.ceil(pos - width).astype(np.int64)
        end = np.floor(pos + width).astype(np.int64)
        t = start[:, None] + a[None, :]
        w = S._table_lookup(np.asarray(table, dtype=np.float64), np.abs(t - pos[:, None]) / width)
        w = np.where((end - start)[:, None] > a[None, :], w, 0.0)
        kap = t % n                                                            # grid index of each tap
        mod = np.exp(1j * 2.0 * np.pi * ((np.arange(n) - c / 2.0) * (c / n)))  # backend.py:357-363, one axis
        g = np.arange(N[d]) + off                                              # grid position of image voxel v
        ph = np.exp(-2j * np.pi * (kap[:, :, None] * g[None, None, :] % n) / n)
        e = (w[:, :, None] * mod[kap][:, :, None] * ph).sum(axis=1) * mod[g][None, :]
        out.append(e)
    return out


def forward_at_samples(N, coord, maps, img, oversamp=2.0, width=3, n=128, weights=None):
    """(A x)[s, c] for the samples `coord` (3, S): complex128 array (S, C)."""
    N = tuple(int(v) for v in N)
    oN, omin = S.oversampled_shape(N, oversamp)
    beta, table = S.kb_table(width, omin, n)
    e0, e1, e2 = _axis_terms(oN, N, coord, width, table)
    apod = S.rolloff3(omin, width, beta, N)
    u = (apod[..., None] * maps.astype(np.complex128)) * img.reshape(N, order='F').astype(np.complex128)[..., None]
    t = np.tensordot(e0, u, axes=(1, 0))                  # (S, N1, N2, C)
    t = np.einsum('sy,syzc->szc', e1, t)
    k = np.einsum('sz,szc->sc', e2, t) / np.sqrt(float(np.prod(oN)))
    if weights is not None:
        k = k * np.asarray(weights, dtype=np.float64).reshape(-1, 1)
    return k


def adjoint_at_voxels(N, coord, ksp, maps, voxels, oversamp=2.0, width=3, n=128, weights=None):
    """(A^H y)[v] for y that is zero except at the samples `coord` (3, S) where it equals ksp (S, C);
    voxels: (V, 3) integer image coordinates.  Returns complex128 (V,)."""
    N = tuple(int(v) for v in N)
    oN, omin = S.oversampled_shape(N, oversamp)
    beta, table = S.kb_table(width, omin, n)
    e0, e1, e2 = _axis_terms(oN, N, coord, width, table)
    voxels = np.asarray(voxels, dtype=np.int64)
    y = np.asarray(ksp, dtype=np.complex128)
    if weights is not None:
        y = y * np.asarray(weights, dtype=np.float64).reshape(-1, 1)
    # conj of the forward kernel, sample by voxel
    kern = np.conj(e0[:, voxels[:, 0]] * e1[:, voxels[:, 1]] * e2[:, voxels[:, 2]])          # (S, V)
    z = np.tensordot(kern, y, axes=(0, 0)) / np.sqrt(float(np.prod(oN)))                      # (V, C)
    apod = S.rolloff3(omin, width, beta, N)
    ix = (voxels[:, 0], voxels[:, 1], voxels[:, 2])
    pf = apod[ix][:, None] * maps.astype(np.complex128)[ix]                                   # (V, C)
    return (np.conj(pf) * z).sum(axis=1)
