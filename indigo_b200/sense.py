"""
Construction of the non-Cartesian SENSE operator exactly as examples/pics.py
does it (pics.py:92-95 and its -O1..-O3 recipe, pics.py:179-193), for any
backend object that offers the reference's builder interface (B200Backend, the
reference's own backends, or the test backends).

    A = KronI(C, NUFFT(M, N, coord)) * VStack_c Diag(maps_c)

After the -O3 recipe one `A.H * A` evaluation is the six backend calls of
SURVEY.md section 3.1:
    ccsrmm(P^H, adjoint) -> fftn -> ccsrmm(G') -> ccsrmm(G', adjoint) -> ifftn -> ccsrmm(P^H)
"""
import numpy as np

_C64 = np.dtype('complex64')


def sense_operator(B, N, coord, maps, oversamp=2.0, level=3, weights=None, recipe=None, width=3, n=128):
    """Returns the optimised forward operator A (image -> multi-coil k-space).

    N      image shape (N0, N1, N2)
    coord  (3, nread, nspokes...) sample positions in cycles/FOV, [-1/2, 1/2)
    maps   (N0, N1, N2, C) coil sensitivities
    weights optional per-sample row weights (sqrt density compensation), applied as
            Diag(w) * NUFFT like test_compat.py:185
    recipe list of Transform classes; default = the reference's -O`level` recipe
           taken from the same module family as B's operators."""
    coord = np.asarray(coord)
    Mshape = (1,) + tuple(coord.shape[1:])           # BART layout: READ dim of k-space is 1 (pics.py:60-63)
    F1 = B.NUFFT(Mshape, tuple(N), coord, width=width, n=n, oversamp=oversamp, dtype=_C64)
    if weights is not None:
        F1 = B.Diag(np.asarray(weights), name='dcf') * F1
    C = maps.shape[3]
    F = B.KronI(C, F1)
    S = B.VStack([B.Diag(maps[:, :, :, c:c + 1]) for c in range(C)], name='maps')
    A = F * S
    A._name = 'SENSE1'
    if recipe is None:
        recipe = default_recipe(B, level)
    return A.optimize(recipe)


def default_recipe(B, level=3):
    """The -O`level` recipe built from the Transform family matching B's operators."""
    ops = getattr(B, 'ops', None)
    if ops is not None and ops.__name__.startswith('indigo_b200'):
        from .host.rewrites import sense_recipe
        return sense_recipe(level)
    from .refcompat import reference_sense_recipe
    return reference_sense_recipe(level)


def normal_operator(A):
    """A^H A (the north-star apply).  lamda is passed to cg(lamda=...), never folded
    in as a Sum: under coil sharding a Sum would add it once per rank."""
    AHA = A.H * A
    AHA._name = 'SENSE'
    return AHA


def sqrt_dcf(coord):
    """sqrt(|k|) row weights of the well-conditioned CG protocol."""
    c = np.asarray(coord, dtype=np.float64)
    c = c.reshape((3, -1), order='F')
    return np.sqrt(np.sqrt((c ** 2).sum(axis=0))).astype(np.float32)
