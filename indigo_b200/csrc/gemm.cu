// Dense complex64 GEMM for DenseMatrix operators (coil compression).
// Interfaces replaced: Backend.cgemm (backend.py:481-485) and Backend.csymm
// (backend.py:487-491); numpy semantics np.py:76-90; the reference GPU path is
// cublasCgemm/cublasCsymm (cuda.py:314-366).
//
// This file is the SIMT fp32 implementation: exact complex64 arithmetic with
// fp32 accumulation, shared-memory tiled, any shape / leading dimension /
// alignment.  One kernel serves op(M) in {M, M^H} on the left and the
// real-symmetric right-multiply through generic element strides.
#include "common.cuh"

namespace ib200 {

static const int BM = 32, BN = 32, BK = 16;

// C[i,j] = alpha * sum_l A(i,l) * B(l,j) + beta * C[i,j]
//   A(i,l) = A[i*sa_i + l*sa_l]  (conjugated if conjA),  B(l,j) = B[l*sb_l + j*sb_j]
__global__ void __launch_bounds__(256) cgemm_kernel(int64_t m, int64_t n, int64_t k, c64 alpha,
                                                    const c64 *__restrict__ A, int64_t sa_i, int64_t sa_l, int conjA,
                                                    const c64 *__restrict__ B, int64_t sb_l, int64_t sb_j, c64 beta,
                                                    int beta_zero, c64 *__restrict__ C, int64_t ldc) {
    __shared__ c64 As[BK][BM + 1];
    __shared__ c64 Bs[BK][BN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;       // 16 x 16 threads, 2 x 2 outputs each
    const int64_t i0 = (int64_t)blockIdx.x * BM, j0 = (int64_t)blockIdx.y * BN;
    c64 acc[2][2] = {{mk(0, 0), mk(0, 0)}, {mk(0, 0), mk(0, 0)}};
    for (int64_t l0 = 0; l0 < k; l0 += BK) {
        for (int e = threadIdx.x; e < BK * BM; e += 256) {
            const int ii = e % BM, ll = e / BM;
            const int64_t i = i0 + ii, l = l0 + ll;
            c64 v = mk(0.f, 0.f);
            if (i < m && l < k) { v = __ldg(A + i * sa_i + l * sa_l); if (conjA) v.y = -v.y; }
            As[ll][ii] = v;
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            const int ll = e % BK, jj = e / BK;
            const int64_t l = l0 + ll, j = j0 + jj;
            c64 v = mk(0.f, 0.f);
            if (l < k && j < n) v = __ldg(B + l * sb_l + j * sb_j);
            Bs[ll][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ll = 0; ll < BK; ++ll) {
            const c64 a0 = As[ll][tx], a1 = As[ll][tx + 16];
            const c64 b0 = Bs[ll][ty], b1 = Bs[ll][ty + 16];
            acc[0][0] = cfma(a0, b0, acc[0][0]); acc[0][1] = cfma(a0, b1, acc[0][1]);
            acc[1][0] = cfma(a1, b0, acc[1][0]); acc[1][1] = cfma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
            if (i < m && j < n) {
                c64 *cp = C + i + j * ldc;
                c64 r = cmul(alpha, acc[a][b]);
                if (!beta_zero) r = cfma(beta, *cp, r);
                *cp = r;
            }
        }
}

static int run_gemm(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i,
                    int64_t sa_l, int conjA, const c64 *B, int64_t sb_l, int64_t sb_j, c64 beta, c64 *C, int64_t ldc) {
    if (m == 0 || n == 0) return 0;
    const int64_t gx = ceil_div(m, BM), gy = ceil_div(n, BN);
    IB200_REQUIRE(gy <= 65535, "cgemm: too many column tiles");
    cgemm_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_l, sb_j,
                                                                 beta, (beta.x == 0.f && beta.y == 0.f) ? 1 : 0, C, ldc);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_cgemm(void *stream, int conjtrans, int64_t m, int64_t n, int64_t k, float ar, float ai, const void *M,
                int64_t ldm, const void *X, int64_t ldx, float br, float bi, void *Y, int64_t ldy) {
    IB200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative dimension");
    IB200_REQUIRE((M && X) || k == 0 || m == 0 || n == 0, "null pointer");
    IB200_REQUIRE(Y || m == 0 || n == 0, "null Y");
    const c64 *Mp = (const c64 *)M;
    if (!conjtrans)   // M is m x k
        return run_gemm(as_stream(stream), m, n, k, mk(ar, ai), Mp, 1, ldm, 0, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
    // M is k x m, op(M)(i,l) = conj(M[l + i*ldm])
    return run_gemm(as_stream(stream), m, n, k, mk(ar, ai), Mp, ldm, 1, 1, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
}

int ib200_csymm(void *stream, int left, int64_t m, int64_t n, float ar, float ai, const void *M, int64_t ldm,
                const void *X, int64_t ldx, float br, float bi, void *Y, int64_t ldy) {
    IB200_REQUIRE(m >= 0 && n >= 0, "negative dimension");
    if (m == 0 || n == 0) return 0;
    IB200_REQUIRE(M && X && Y, "null pointer");
    if (left)     // Y(m x n) = M(m x m) X(m x n)
        return run_gemm(as_stream(stream), m, n, m, mk(ar, ai), (const c64 *)M, 1, ldm, 0, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
    // Y(m x n) = X(m x n) M(n x n)
    return run_gemm(as_stream(stream), m, n, n, mk(ar, ai), (const c64 *)X, 1, ldx, 0, (const c64 *)M, 1, ldm, mk(br, bi), (c64 *)Y, ldy);
}

}  // extern "C"
