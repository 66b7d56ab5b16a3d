"""
Host-side Kaiser-Bessel helpers of the fused SENSE recipe: the apodisation correction that is
folded into `pf` (indigo_b200/fused.py).  Plain vectorised numpy (the reference's noncart.py:5-23
needs numexpr); the arithmetic order follows the reference so that `pf` matches the product of
its `apod`, `zpad` and `mod` diagonals to the last bit (tests/test_oracle.py pins the oracle's copy
of the same formula against reference-generated vectors).
"""
import numpy as np


def ftkb(beta, x):
    """Fourier transform of the Kaiser-Bessel window: sinh(a)/a with a = sqrt(beta^2 - (pi x)^2)."""
    a = np.sqrt(beta ** 2 - (np.pi * x) ** 2)
    y = np.ones(a.shape, dtype=a.dtype)
    nz = a != 0.0
    y[nz] = np.sinh(a[nz]) / a[nz]
    return y


def rolloff3(oversamp, width, beta, N):
    """Apodisation correction on the N0 x N1 x N2 image grid."""
    g = np.mgrid[:N[0], :N[1], :N[2]]
    den = 1.0
    for d in range(3):
        den = den * ftkb(beta, (g[d] - N[d] // 2) / N[d] * width * 2.0 / oversamp)
    return ftkb(beta, 0.0) ** 3 / den
