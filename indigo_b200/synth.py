"""
Seeded synthetic inputs of BASELINE.json's configurations.
Pure numpy, no device code: shared by tests, bench.py (both arms) and smoke().

Values follow indigo.util.rand64c (indigo/util.py:9-19): U[0,1) + i*U[0,1),
real part drawn first, each cast to float32 -- but from an explicit
RandomState so every run is reproducible (the reference never seeds).
"""
import numpy as np

C64 = np.dtype("complex64")


def rand64c(rs, *shape, order="F"):
    re = rs.rand(*shape).astype(np.float32)
    im = rs.rand(*shape).astype(np.float32)
    arr = (re + 1j * im).astype(np.complex64)
    return np.asfortranarray(arr) if order == "F" else arr


def unit_rss_maps(rs, N, C):
    """Coil sensitivities (N0,N1,N2,C), rand64c normalised to unit root-sum-of-squares."""
    m = rand64c(rs, *N, C)
    rss = np.sqrt((np.abs(m) ** 2).sum(axis=3, keepdims=True))
    return np.asfortranarray((m / rss).astype(np.complex64))


def radial_2d(nspokes=402, nread=512):
    """cfg1: in-plane radial, k = r*(cos t, sin t, 0), r in [-1/2,1/2), t_j = pi*j/nspokes.
    Returns coord (3, nread, nspokes) float64 (BART layout: ksp dims (1, nread, nspokes))."""
    r = (np.arange(nread) - nread // 2) / nread
    t = np.pi * np.arange(nspokes) / nspokes
    c = np.zeros((3, nread, nspokes))
    c[0] = r[:, None] * np.cos(t)[None, :]
    c[1] = r[:, None] * np.sin(t)[None, :]
    return c


def kooshball_3d(nspokes=16384, nread=416):
    """cfg3/4: 3-D radial with Fibonacci (golden-angle) spoke directions.
    Returns coord (3, nread, nspokes) float64."""
    j = np.arange(nspokes) + 0.5
    z = 1.0 - 2.0 * j / nspokes
    phi = np.pi * (1.0 + 5.0 ** 0.5) * j
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([s * np.cos(phi), s * np.sin(phi), z])          # (3, nspokes)
    r = (np.arange(nread) - nread // 2) / nread                   # [-1/2, 1/2)
    return r[None, :, None] * d[:, None, :]


def stack_of_spirals(nz=128, nleaves=48, nread=2048, turns=16.0):
    """cfg5: Archimedean spiral interleaves in-plane, one stack per kz plane.
    Returns coord (3, nread, nleaves*nz) float64."""
    t = np.arange(nread) / nread
    rad = 0.5 * t
    c = np.zeros((3, nread, nleaves, nz))
    for l in range(nleaves):
        ang = 2.0 * np.pi * (turns * t + l / nleaves)
        c[0, :, l, :] = (rad * np.cos(ang))[:, None]
        c[1, :, l, :] = (rad * np.sin(ang))[:, None]
    c[2] = ((np.arange(nz) - nz // 2) / nz)[None, None, :]
    return c.reshape(3, nread, nleaves * nz)


def random_3d(rs, npts):
    """Uniform random sample positions in [-1/2, 1/2)^3, (3, npts, 1)."""
    return (rs.rand(3, npts, 1) - 0.5)


def random_csr(rs, rows, cols, nnz_per_row):
    """cfg2: `nnz_per_row` distinct uniformly-random sorted columns per row, rand64c values.
    Returns (indptr int32, indices int32, data complex64)."""
    r = int(nnz_per_row)
    # distinct columns per row: sample with a stride trick (jittered stratified draw)
    width = cols // r
    base = (np.arange(r, dtype=np.int64) * width)[None, :]
    idx = base + (rs.rand(rows, r) * width).astype(np.int64)
    indptr = (np.arange(rows + 1, dtype=np.int64) * r).astype(np.int32)
    data = rand64c(rs, rows * r, order="C")
    return indptr, idx.astype(np.int32).reshape(-1), data
