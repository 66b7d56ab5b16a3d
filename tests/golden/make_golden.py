"""
Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, NumpyBackend) on seeded inputs.  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

The reference holds no golden vectors of its own, so
these files are the pin for oracle/ (tests/test_oracle.py) and, through it,
for the CUDA path.  The fixtures are committed; this script cannot run on the
GPU box (no /root/reference there).

The `-O3` recipe classes are taken from the reference's examples/pics.py by
exec'ing that file's own text between two markers -- nothing is copied here.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import scipy.sparse as spp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

from refshim import load_reference, REFERENCE  # noqa: E402
from indigo_b200 import synth                  # noqa: E402

indigo = load_reference()
from indigo.backends import get_backend        # noqa: E402

C64 = np.dtype("complex64")


def pics_recipe():
    src = open(os.path.join(REFERENCE, "examples", "pics.py")).read()
    seg = src[src.index("import scipy.sparse as spp"):src.index("recipe = []")]
    ns = {}
    exec(seg, ns)
    return [ns[k] for k in ("MakeRightLeaning", "AssocSpMatrices", "DistKroniOverFFT",
                            "MakeRightLeaning", "MriRealize", "MriGoodAdjoints")]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def F(a):
    return np.asfortranarray(a)


# ---------------------------------------------------------------------------
def primitives():
    B = get_backend("numpy")
    rs = np.random.RandomState(1234)
    out = {}

    # ---- ccsrmm: generic matrix, fwd + adjoint, alpha/beta, ld > rows -------
    m, k, n = 23, 45, 9
    A = (spp.random(m, k, density=0.2, format="csr", random_state=rs, dtype=np.float32)
         + 1j * spp.random(m, k, density=0.2, format="csr", random_state=rs, dtype=np.float32)).astype(C64).tocsr()
    A.sort_indices()
    A_d = B.csr_matrix(B, A)
    alpha, beta = 1.5 - 0.5j, 0.5 + 0.25j
    xbig, ybig = synth.rand64c(rs, k + 7, n), synth.rand64c(rs, m + 5, n)
    xd, yd = B.copy_array(xbig), B.copy_array(ybig)
    A_d.forward(yd[2:2 + m, :], xd[3:3 + k, :], alpha=alpha, beta=beta)
    out.update(csr_indptr=A.indptr, csr_indices=A.indices, csr_data=A.data, csr_shape=np.array(A.shape),
               csr_alpha=np.array(alpha), csr_beta=np.array(beta),
               csr_fwd_xbig=xbig, csr_fwd_ybig=ybig, csr_fwd_out=yd.to_host())
    xbig2, ybig2 = synth.rand64c(rs, m + 5, n), synth.rand64c(rs, k + 7, n)
    xd, yd = B.copy_array(xbig2), B.copy_array(ybig2)
    A_d.adjoint(yd[3:3 + k, :], xd[2:2 + m, :], alpha=alpha, beta=beta)
    out.update(csr_adj_xbig=xbig2, csr_adj_ybig=ybig2, csr_adj_out=yd.to_host(),
               csr_inspect=np.array([A_d._row_frac, A_d._col_frac, A_d._exwrite], dtype=np.float64))

    # ---- exclusive-write matrix (test_backends.py:213-243 construction) ----
    Kx, Mx = 19, 23
    counts = rs.randint(0, 2, Kx)
    ptr = np.concatenate([[0], np.cumsum(counts)])
    ind = rs.randint(0, Mx, counts.sum())
    vals = synth.rand64c(rs, ind.size, order="C")
    E = spp.csr_matrix((vals, ind, ptr), shape=(Kx, Mx)).T.tocsr()
    E.sort_indices()
    E_d = B.csr_matrix(B, E)
    x, y = synth.rand64c(rs, Mx, 8), synth.rand64c(rs, Kx, 8)
    xd, yd = B.copy_array(x), B.copy_array(y)
    E_d.adjoint(yd, xd, alpha=0.5, beta=1.5)
    out.update(exw_indptr=E.indptr, exw_indices=E.indices, exw_data=E.data, exw_shape=np.array(E.shape),
               exw_x=x, exw_y=y, exw_adj_out=yd.to_host(), exw_flag=np.array(E_d._exwrite))

    # ---- FFTs (sizes of test_backends.py:154 / test_operators.py:228) ------
    for tag, shp in (("fft3", (23, 24, 25, 2)), ("fft2", (24, 22, 3)), ("fft1", (22, 4)), ("fft3b", (16, 13, 7, 3))):
        v = synth.rand64c(rs, *shp)
        vd, ud = B.copy_array(v), B.copy_array(v)
        B.fftn(ud, vd)
        fwd = ud.to_host()
        B.ifftn(ud, vd)
        out.update({tag + "_in": v, tag + "_fwd": fwd, tag + "_inv": ud.to_host()})

    # ---- BLAS-1 -------------------------------------------------------------
    n1 = 129
    x, y = synth.rand64c(rs, n1), synth.rand64c(rs, n1)
    xd, yd = B.copy_array(x), B.copy_array(y)
    B.axpby(0.5 + 1.5j, yd, -2.1 + 3j, xd)
    out.update(b1_x=x, b1_y=y, b1_axpby=yd.to_host(), b1_dot=np.array(B.dot(xd, B.copy_array(y))),
               b1_nrm2=np.array(B.norm2(xd)))
    B.scale(xd, 1.1 - 2j)
    out.update(b1_scale=xd.to_host())

    # ---- dense --------------------------------------------------------------
    mm, nn, kk = 23, 10, 129
    Md, x, y = synth.rand64c(rs, mm, kk), synth.rand64c(rs, kk, nn), synth.rand64c(rs, mm, nn)
    yd = B.copy_array(y)
    B.cgemm(yd, B.copy_array(Md), B.copy_array(x), 0.5 + 0.5j, 0.5, forward=True)
    out.update(gemm_M=Md, gemm_x=x, gemm_y=y, gemm_fwd=yd.to_host())
    xd = B.copy_array(x)
    B.cgemm(xd, B.copy_array(Md), B.copy_array(y), 1.0, 0.5, forward=False)
    out.update(gemm_adj=xd.to_host())
    S = synth.rand64c(rs, 6, 6); S = F(S + S.T); S.imag = 0
    xl, yl = synth.rand64c(rs, 6, 3), synth.rand64c(rs, 6, 3)
    xr, yr = synth.rand64c(rs, 3, 6), synth.rand64c(rs, 3, 6)
    yd = B.copy_array(yl); B.csymm(yd, B.copy_array(S), B.copy_array(xl), 1.5, 0.5, True)
    out.update(symm_M=S, symm_xl=xl, symm_yl=yl, symm_left=yd.to_host())
    yd = B.copy_array(yr); B.csymm(yd, B.copy_array(S), B.copy_array(xr), 1.5, 0.5, False)
    out.update(symm_xr=xr, symm_yr=yr, symm_right=yd.to_host())

    # ---- onemm / cdiamm / max ----------------------------------------------
    x, y = synth.rand64c(rs, 11, 5), synth.rand64c(rs, 7, 5)
    yd = B.copy_array(y); B.onemm(yd, B.copy_array(x), 1.5 - 1j, 0.5)
    out.update(one_x=x, one_y=y, one_out=yd.to_host())
    Md_, Kd_, Nd_ = 23, 45, 9
    offs = np.array(sorted(set(rs.randint(-Kd_, Md_ + Kd_, size=4))))
    data = (rs.rand(offs.size, Kd_) + 1j * rs.rand(offs.size, Kd_)).astype(C64)
    D = spp.dia_matrix((data, offs), shape=(Md_, Kd_))
    D_d = B.dia_matrix(B, D)
    x, y = synth.rand64c(rs, Kd_, Nd_), synth.rand64c(rs, Md_, Nd_)
    yd = B.copy_array(y); D_d.forward(yd, B.copy_array(x), alpha=1.5, beta=0.5)
    out.update(dia_offsets=offs, dia_data=data, dia_shape=np.array(D.shape), dia_x=x, dia_y=y, dia_fwd=yd.to_host())
    xd = B.copy_array(x); D_d.adjoint(xd, B.copy_array(y), alpha=0.5, beta=1.5)
    out.update(dia_adj=xd.to_host())
    a = (synth.rand64c(rs, 37) - (0.5 + 0.5j)).astype(C64)
    ad = B.copy_array(a); B.max(0.1, ad)
    out.update(max_in=a, max_out=ad.to_host())

    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **out)
    print("primitives.npz:", len(out), "arrays")


# ---------------------------------------------------------------------------
def build_sense(B, N, coord, maps, oversamp, weights=None):
    """The reference construction of examples/pics.py:92-95 + its -O3 recipe."""
    Mshape = (1,) + coord.shape[1:]
    F1 = B.NUFFT(Mshape, N, coord, oversamp=oversamp, dtype=C64)
    if weights is not None:
        F1 = B.Diag(weights, name="dcf") * F1
    C = maps.shape[3]
    Fk = B.KronI(C, F1)
    S = B.VStack([B.Diag(maps[:, :, :, c:c + 1]) for c in range(C)], name="maps")
    A = Fk * S
    A = A.optimize(pics_recipe())
    B._scratch._arr[:] = 0          # oracle hygiene (SURVEY.md landmine 3)
    return A


def leaf_matrices(A):
    from indigo.operators import SpMatrix
    found = []

    def walk(n):
        if isinstance(n, SpMatrix):
            found.append(n)
        for c in getattr(n, "_children", []):
            walk(c)
    walk(A)
    G = [n for n in found if "interp" in n._name][0]
    P = [n for n in found if "zpad" in n._name][0]
    return G._get_or_create_device_matrix(), P._get_or_create_device_matrix()


def sense_case(tag, N, C, coord, oversamp, seed, cg_iters=8, full=True):
    B = get_backend("numpy")
    rs = np.random.RandomState(seed)
    maps = synth.unit_rss_maps(rs, N, C)
    A = build_sense(B, N, coord, maps, oversamp)
    Gd, Pd = leaf_matrices(A)
    npts = int(np.prod(coord.shape[1:]))
    x = synth.rand64c(rs, int(np.prod(N)), 1)
    y = synth.rand64c(rs, npts * C, 1)
    Ax = A * x
    AHy = A.H * y
    AHA = A.H * A
    AHAx = AHA * x
    out = dict(N=np.array(N), C=np.array(C), oversamp=np.array(oversamp), seed=np.array(seed),
               G_shape=np.array(Gd.shape), P_shape=np.array(Pd.shape),
               G_nnz=np.array(Gd.values._arr.size), P_nnz=np.array(Pd.values._arr.size),
               G_exwrite=np.array(Gd._exwrite), P_exwrite=np.array(Pd._exwrite))
    if full:
        out.update(coord=coord, maps=maps, x=x, y=y, Ax=Ax, AHy=AHy, AHAx=AHAx,
                   G_indptr=Gd.rowPtrs._arr, G_indices=Gd.colInds._arr, G_data=Gd.values._arr,
                   P_indptr=Pd.rowPtrs._arr, P_indices=Pd.colInds._arr, P_data=Pd.values._arr)
        # well-conditioned CG protocol: sqrt-DCF rows, lamda via cg(lamda=)
        w = np.sqrt(np.sqrt((coord.reshape((3, -1), order='F') ** 2).sum(axis=0))).astype(np.float32)
        Bw = get_backend("numpy")
        Aw = build_sense(Bw, N, coord, maps, oversamp, weights=w)
        AwHAw = Aw.H * Aw
        xt = synth.rand64c(rs, int(np.prod(N)), 1)
        b = AwHAw * xt
        b = F((b / np.abs(b).max()).astype(C64))
        lam = 1e-2 * float(np.abs(np.vdot(xt, AwHAw * xt)) / np.vdot(xt, xt).real)
        its = []
        for k in range(1, cg_iters + 1):
            xk = np.zeros_like(b, order="F")
            Bw.cg(AwHAw, b, xk, lamda=lam, maxiter=k, tol=0.0)
            its.append(xk.copy())
        out.update(cg_w=w, cg_b=b, cg_lamda=np.array(lam), cg_iterates=np.stack(its))
    else:
        sub = slice(None, None, 997)
        out.update(G_indptr_sha=np.array(sha(Gd.rowPtrs._arr)), G_indices_sha=np.array(sha(Gd.colInds._arr)),
                   P_indptr_sha=np.array(sha(Pd.rowPtrs._arr)), P_indices_sha=np.array(sha(Pd.colInds._arr)),
                   G_data_sub=Gd.values._arr[sub], P_data_sub=Pd.values._arr[sub],
                   Ax_sub=Ax.ravel(order="F")[sub], AHy_sub=AHy.ravel(order="F")[sub],
                   AHAx_sub=AHAx.ravel(order="F")[sub],
                   Ax_norm=np.array(np.linalg.norm(Ax)), AHy_norm=np.array(np.linalg.norm(AHy)),
                   AHAx_norm=np.array(np.linalg.norm(AHAx)))
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print(tag, "G", Gd.shape, "nnz", Gd.values._arr.size, "P", Pd.shape, "nnz", Pd.values._arr.size)


def main():
    primitives()
    rs = np.random.RandomState(7)
    # small, even grid; a few samples exactly on grid points (6-tap rows) and at the wrap-around edge
    coord = synth.random_3d(rs, 200)
    coord[:, :6, 0] = np.array([[0.0, 0.25, -0.5, 0.125, -0.5, 0.0],
                                [0.0, -0.25, -0.5, 0.3, 0.49, 0.1],
                                [0.0, 0.0, -0.5, -0.2, -0.49, 0.25]])
    sense_case("sense_small", (12, 10, 6), 3, coord, 2.0, seed=11)
    # odd oversampled grid (16,13,7): complex centring phases, generic-radix FFT sizes
    sense_case("sense_odd", (11, 9, 5), 2, synth.random_3d(np.random.RandomState(8), 150), 1.5, seed=12, cg_iters=4)
    if os.environ.get("GOLDEN_CFG1", "1") == "1":
        sense_case("sense_cfg1_digest", (256, 256, 1), 8, synth.radial_2d(402, 512), 2.0, seed=13, full=False)


if __name__ == "__main__":
    main()
