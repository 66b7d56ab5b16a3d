"""
CPU emulation of the index arithmetic of the tensor-core cgemm kernel (indigo_b200/csrc/gemm.cu,
cgemm_tc_kernel): the complex product as a real product on interleaved floats, the permuted reduction
order (one 16-byte load per lane feeding two m16n8k8 k-steps, 8-float tail step), the per-lane fragment
table of alpha*op(M)' with and without the paired output assignment, and the accumulator -> Y mapping,
lane by lane against numpy in float64 (so only the indexing is under test; the TF32 splitting is checked
on the GPU by tests/test_gpu_primitives.py::test_cgemm_tensor_core).  Replaces Backend.cgemm
(indigo/backends/backend.py:481-485; numpy semantics indigo/backends/np.py:76-87).
"""
import numpy as np
import pytest


def emulate_tc(M, X, alpha, forward, pair):
    opM = M if forward else M.conj().T
    m, k = opM.shape
    n = X.shape[1]
    K2, KG = 2 * k, (2 * k + 15) // 16
    NT = -(-m // 4)
    if pair and NT % 2:
        NT += 1
    Xf = np.zeros((n, K2))                                   # column j of X as interleaved (re, im) floats
    Xf[:, 0::2], Xf[:, 1::2] = X.real.T, X.imag.T
    W = alpha * opM

    def table(s, nt, lane):                                  # (b0, b1) of k-step s, n-tile nt
        G, n8, t = s >> 1, lane >> 2, lane & 3
        o = 2 * (8 * (nt >> 1) + 2 * (n8 >> 1) + (nt & 1)) + (n8 & 1) if pair else nt * 8 + n8
        i = o >> 1
        if K2 - 16 * G == 8:
            l = k if (s & 1) else (16 * G + 2 * t) >> 1
        else:
            l = (16 * G + 4 * t + 2 * (s & 1)) >> 1
        if i < m and l < k:
            w = W[i, l]
            return (w.imag, w.real) if (o & 1) else (w.real, -w.imag)
        return (0.0, 0.0)

    Y = np.zeros((m, n), dtype=np.complex128)
    for tile in range(-(-n // 16)):
        acc = np.zeros((NT, 16, 8))                          # accumulator tiles: 16 columns of X x 8 real outputs
        for G in range(KG):
            tail = K2 - 16 * G == 8
            for h in range(1 if tail else 2):
                A, B = np.zeros((16, 8)), np.zeros((NT, 8, 8))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for half, row in ((0, g), (1, g + 8)):
                        j = tile * 16 + row
                        if tail:
                            v = list(Xf[j, 16 * G + 2 * t:16 * G + 2 * t + 2]) + [0.0, 0.0] if j < n else [0.0] * 4
                        else:
                            ok = 16 * G + 4 * t + 3 < K2 and j < n
                            v = list(Xf[j, 16 * G + 4 * t:16 * G + 4 * t + 4]) if ok else [0.0] * 4
                        A[row, t], A[row, t + 4] = v[2 * h], v[2 * h + 1]
                    for nt in range(NT):
                        B[nt, t, g], B[nt, t + 4, g] = table(2 * G + h, nt, lane)
                for nt in range(NT):
                    acc[nt] += A @ B[nt]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for row in (g, g + 8):
                j = tile * 16 + row
                if j >= n:
                    continue
                if pair:
                    for nt in range(0, NT, 2):
                        i = 4 * nt + 2 * t
                        if i < m:
                            Y[i, j] = acc[nt, row, 2 * t] + 1j * acc[nt, row, 2 * t + 1]
                        if i + 1 < m:
                            Y[i + 1, j] = acc[nt + 1, row, 2 * t] + 1j * acc[nt + 1, row, 2 * t + 1]
                else:
                    for nt in range(NT):
                        i = nt * 4 + t
                        if i < m:
                            Y[i, j] = acc[nt, row, 2 * t] + 1j * acc[nt, row, 2 * t + 1]
    return Y


@pytest.mark.parametrize("m,k,n", [(12, 48, 32), (48, 12, 21), (5, 2, 16), (3, 10, 40), (8, 4, 17), (6, 22, 16), (1, 14, 5)])
@pytest.mark.parametrize("forward", [True, False])
@pytest.mark.parametrize("pair", [False, True])
def test_tc_index_arithmetic(m, k, n, forward, pair):
    rs = np.random.RandomState(m * 100 + k)
    M = rs.randn(m, k) + 1j * rs.randn(m, k) if forward else rs.randn(k, m) + 1j * rs.randn(k, m)
    X = rs.randn(k, n) + 1j * rs.randn(k, n)
    alpha = 0.5 - 1.5j
    want = alpha * ((M if forward else M.conj().T) @ X)
    got = emulate_tc(M, X, alpha, forward, pair)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_split_tf32_bit_trick():
    """hi = (bits + 0x1000) & ~0x1fff rounds to 10 mantissa bits (half away from zero); lo = x - hi is exact in
    fp32 and, truncated to 10 mantissa bits by the tensor cores, leaves |x - hi - lo'| <= 2^-21 |x|."""
    rs = np.random.RandomState(3)
    x = (rs.randn(100000) * np.exp(rs.uniform(-20, 20, 100000))).astype(np.float32)
    bits = x.view(np.uint32)
    hi = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (x - hi).astype(np.float32)
    assert np.all(np.abs(x - hi) <= np.abs(x) * 2.0 ** -11 * (1 + 1e-6))
    assert np.all(x.astype(np.float64) - hi.astype(np.float64) == lo.astype(np.float64))        # exact difference
    lo_t = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)                           # operand as the MMA reads it
    err = np.abs(x.astype(np.float64) - hi.astype(np.float64) - lo_t.astype(np.float64))
    assert np.all(err <= np.abs(x.astype(np.float64)) * 2.0 ** -21)


# ---------------------------------------------------------------------------------------------------------------
# tcgen05 kernel (cgemm_t5ws_kernel): shared-memory operand layout and descriptor arithmetic.  The model of the
# hardware side is the canonical K-major no-swizzle layout of a shared-memory matrix descriptor (8-row x 16-byte
# core matrices; SBO bytes between 8-row groups, LBO bytes between the two 16-byte k slices of one K = 8 MMA), which
# is what the kernel's descriptors (start, LBO, SBO = 128) claim; GPU parity tests confirm the claim on hardware.
def _mma_read(plane, start, lbo, sbo, rows):
    """rows x 8 operand an MMA fetches through descriptor (start, lbo, sbo); plane is a float32 array, bytes / 4."""
    out = np.zeros((rows, 8))
    for r in range(rows):
        for kk in range(8):
            out[r, kk] = plane[(start + (r // 8) * sbo + (r % 8) * 16 + (kk // 4) * lbo + (kk % 4) * 4) // 4]
    return out


def _split(x):
    hi = np.round(x * 2.0 ** 10 / 2.0 ** np.floor(np.log2(np.maximum(np.abs(x), 1e-300)))) \
        * 2.0 ** np.floor(np.log2(np.maximum(np.abs(x), 1e-300))) / 2.0 ** 10
    return hi, x - hi


@pytest.mark.parametrize("m,k", [(12, 48), (48, 12), (5, 8), (20, 40)])
def test_t5_operand_layout(m, k):
    rs = np.random.RandomState(k)
    ROWS, pad = 128, 16
    K2 = 2 * k
    NCH = 1 if K2 <= 48 else 2
    KCH, N = K2 // NCH, (2 * m + 31) // 32 * 32
    CPR, lbo_a, lbo_b = KCH // 4, ROWS * 16 + pad, 2 * N * 16
    M = rs.randn(m, k) + 1j * rs.randn(m, k)
    X = rs.randn(k, ROWS) + 1j * rs.randn(k, ROWS)
    alpha = 1.25 + 0.5j
    Xf = np.zeros((ROWS, K2)); Xf[:, 0::2], Xf[:, 1::2] = X.real.T, X.imag.T
    # stacked planes of alpha*M' as the kernel writes them
    bcat = np.zeros((K2 // 4) * lbo_b // 4)
    for o in range(N):
        for kf in range(K2):
            i, l, v = o >> 1, kf >> 1, 0.0
            if i < m:
                w = alpha * M[i, l]
                v = (w.real if kf & 1 else w.imag) if o & 1 else (-w.imag if kf & 1 else w.real)
            h, lo = _split(np.array([v]))
            off = (kf >> 2) * lbo_b + o * 16 + (kf & 3) * 4
            bcat[off // 4], bcat[(off + N * 16) // 4] = h[0], lo[0]
    D = np.zeros((ROWS, 2 * N))                                # TMEM accumulator: lane = column of X
    for ch in range(NCH):
        a_hi, a_lo = np.zeros(CPR * lbo_a // 4), np.zeros(CPR * lbo_a // 4)
        for q in range(ROWS * CPR):                            # converter pieces: q -> (r, slice)
            r, sl = q // CPR, q % CPR
            h, lo = _split(Xf[r, ch * KCH + 4 * sl: ch * KCH + 4 * sl + 4])
            off = (sl * lbo_a + r * 16) // 4
            a_hi[off:off + 4], a_lo[off:off + 4] = h, lo
        for ks in range(KCH // 8):
            ao, bo = 2 * ks * lbo_a, (ch * CPR + 2 * ks) * lbo_b
            Ah, Al = _mma_read(a_hi, ao, lbo_a, 128, ROWS), _mma_read(a_lo, ao, lbo_a, 128, ROWS)
            Bc = _mma_read(bcat, bo, lbo_b, 128, 2 * N)
            D += Ah @ Bc.T                                     # 128 x 2N: hi*hi | hi*lo
            D[:, :N] += Al @ Bc[:N].T                          # 128 x N : lo*hi
    Yf = D[:, :N] + D[:, N:]
    got = (Yf[:, 0:2 * m:2] + 1j * Yf[:, 1:2 * m:2]).T
    want = alpha * (M @ X)
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-5
    assert np.all(Yf[:, 2 * m:] == 0)
