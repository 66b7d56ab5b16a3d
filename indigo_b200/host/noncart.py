"""
Host-side construction of the non-Cartesian pieces: Kaiser-Bessel gridding
matrix and apodisation (mirror of indigo/interp.py and indigo/noncart.py, which
need numba and numexpr; this version is plain vectorised numpy and produces the
same COO triplets in the same order, verified bit for bit against the
reference-generated golden vectors in tests/test_host_mirror.py).
"""
import numpy as np
import scipy.sparse as spp


def _kb_weight(table, dist):
    """Linear interpolation into the half-window table; 0 at and beyond the window
    edge (interp.py:9-15)."""
    n = table.shape[0]
    hit = dist < 1
    u = np.where(hit, dist, 0.0) * (n - 1)
    i0 = u.astype(np.int64)
    f = u - i0
    w = (1.0 - f) * table[i0] + f * table[np.minimum(i0 + 1, n - 1)]
    return np.where(hit, w, 0.0)


def interp_triplets(N, width, table, coord, block=1 << 15):
    """(row, col, weight) of the gridding matrix for sample positions coord (3, m) in
    cycles/FOV: per axis the taps ceil(p-width) .. floor(p+width)-1 around
    p = N*k + N//2, wrapped modulo N; weight = wz*wy*wx in float64; emitted
    sample-major, then z, y, x (interp.py:19-60)."""
    table = np.asarray(table, dtype=np.float64)
    coord = np.asarray(coord, dtype=np.float64)
    m = coord.shape[1]
    ntap = int(2 * width + 1)
    k = np.arange(ntap)
    out_r, out_c, out_w = [], [], []
    for lo in range(0, m, block):
        hi = min(lo + block, m)
        idx, wgt, use = [], [], []
        for d in range(3):
            p = N[d] * coord[d, lo:hi] + (N[d] // 2)
            first = np.ceil(p - width).astype(np.int64)
            count = np.floor(p + width).astype(np.int64) - first
            t = first[:, None] + k
            idx.append(t % N[d])
            wgt.append(_kb_weight(table, np.abs(t - p[:, None]) / width))
            use.append(k[None, :] < count[:, None])
        keep = use[2][:, :, None, None] & use[1][:, None, :, None] & use[0][:, None, None, :]
        w = (wgt[2][:, :, None, None] * wgt[1][:, None, :, None]) * wgt[0][:, None, None, :]
        c = idx[0][:, None, None, :] + ((idx[1] * N[0])[:, None, :, None] + (idx[2] * (N[1] * N[0]))[:, :, None, None])
        r = np.broadcast_to(np.arange(lo, hi)[:, None, None, None], keep.shape)
        out_r.append(r[keep]); out_c.append(np.broadcast_to(c, keep.shape)[keep]); out_w.append(w[keep])
    return np.concatenate(out_r), np.concatenate(out_c), np.concatenate(out_w)


def interp_mat(m, N, width, table, coord, backend=None):
    """COO gridding matrix, m x prod(N) (interp.py:63-80; only the 3-D variant exists there)."""
    if coord.shape[0] != 3:
        raise ValueError('Number of dimensions can only be 3, got %r' % (coord.shape[0],))
    r, c, w = interp_triplets(tuple(int(n) for n in N), width, table, coord)
    return spp.coo_matrix((w, (r, c)), shape=(m, int(np.prod(N))))


def ftkb(beta, x):
    """Fourier transform of the Kaiser-Bessel window, sinh(a)/a with a = sqrt(beta^2 - (pi x)^2)
    (noncart.py:5-14)."""
    a = np.sqrt(beta ** 2 - (np.pi * x) ** 2)
    y = np.ones(a.shape, dtype=a.dtype)
    nz = a != 0.0
    y[nz] = np.sinh(a[nz]) / a[nz]
    return y


def rolloff3(oversamp, width, beta, N):
    """Apodisation correction on the N0 x N1 x N2 image grid (noncart.py:17-23)."""
    g = np.mgrid[:N[0], :N[1], :N[2]]
    den = 1.0
    for d in range(3):
        den = den * ftkb(beta, (g[d] - N[d] // 2) / N[d] * width * 2.0 / oversamp)
    return ftkb(beta, 0.0) ** 3 / den
