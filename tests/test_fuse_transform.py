"""
Host-side test of the fusion Transform (indigo_b200.fused.fuse_transform): the tree of
examples/pics.py:92-95, built by the REFERENCE's own builders on its NumpyBackend, is recognised, the arguments of the fused node are
recovered bit-exactly from it (trajectory from the NUFFT tag, maps and row weights from the diagonal matrices),
and any other tree is left alone.  The fused node itself needs a GPU (tests/test_gpu_fused.py); here its
constructor is replaced by a recorder.
"""
import numpy as np
import scipy.sparse as spp

from indigo_b200 import synth
from indigo_b200.fused import fuse_transform, match_sense_tree, tag_nufft
from indigo_b200.sense import sqrt_dcf
from refenv import numpy_backend as NpHostBackend

C64 = np.dtype('complex64')


def _tree(B, N, coord, maps, weights=None, tag=True):
    M = (1,) + tuple(coord.shape[1:])
    F1 = B.NUFFT(M, tuple(N), coord, width=3, n=128, oversamp=2.0, dtype=C64)
    if tag:
        tag_nufft(F1, N, coord, 3, 128, 2.0)
    if weights is not None:
        F1 = B.Diag(np.asarray(weights), name='dcf') * F1
    C = maps.shape[3]
    A = B.KronI(C, F1) * B.VStack([B.Diag(maps[:, :, :, c:c + 1]) for c in range(C)], name='maps')
    A._name = 'SENSE1'
    return A


def _setup(weighted):
    rs = np.random.RandomState(4)
    N, C = (6, 8, 4), 3
    coord = synth.random_3d(rs, 50)
    maps = synth.unit_rss_maps(rs, N, C)
    return N, C, coord, maps, (sqrt_dcf(coord) if weighted else None)


def test_sense_tree_is_recognised_and_arguments_recovered():
    for weighted in (False, True):
        N, C, coord, maps, w = _setup(weighted)
        B = NpHostBackend()
        A = _tree(B, N, coord, maps, w)
        hit = match_sense_tree(A)
        assert hit is not None
        assert hit['N'] == N and hit['oversamp'] == 2.0 and hit['width'] == 3 and hit['n'] == 128
        np.testing.assert_array_equal(hit['coord'], coord)
        np.testing.assert_array_equal(hit['maps'], maps.astype(C64))
        if weighted:
            np.testing.assert_array_equal(hit['weights'], np.asarray(w, dtype=np.float32).reshape(-1))
        else:
            assert hit['weights'] is None


def test_transform_swaps_only_the_sense_product():
    N, C, coord, maps, w = _setup(True)
    B = NpHostBackend()
    T = fuse_transform(B)
    seen = []

    def recorder(backend, **kw):
        seen.append(kw)
        rows = int(np.prod(kw['coord'].shape[1:])) * kw['maps'].shape[3]
        return B.SpMatrix(spp.csr_matrix((rows, int(np.prod(kw['N']))), dtype=C64), name='stand-in')

    T.build = staticmethod(recorder)
    A = _tree(B, N, coord, maps, w)
    out = T().visit(A)
    assert len(seen) == 1 and out._name == 'SENSE1.fused'
    # inside a larger tree the SENSE product is replaced in place
    big = B.Diag(np.ones(A.shape[0], dtype=C64), name='mask') * _tree(B, N, coord, maps, w)
    out = T().visit(big)
    assert len(seen) == 2 and type(out).__name__ == 'Product' and out.children[1]._name == 'SENSE1.fused'
    # untagged NUFFT (trajectory unknown), a grid without a fused plan, and unrelated trees stay as they are
    plain = _tree(B, N, coord, maps, w, tag=False)
    assert match_sense_tree(plain) is None and type(T().visit(plain)).__name__ == 'Product' and len(seen) == 2

    def refuse(backend, **kw):
        raise RuntimeError("no specialised passes for this grid")

    T.build = staticmethod(refuse)
    kept = T().visit(_tree(B, N, coord, maps, w))
    assert type(kept).__name__ == 'Product' and type(kept.children[0]).__name__ == 'Kron'
    other = B.Diag(np.ones(5, dtype=C64)) * B.Diag(np.ones(5, dtype=C64))
    assert match_sense_tree(other) is None
