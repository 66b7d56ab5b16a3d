// Dense complex64 GEMM for DenseMatrix operators (coil compression).
// Interfaces replaced: Backend.cgemm (backend.py:481-485) and Backend.csymm
// (backend.py:487-491); numpy semantics np.py:76-90; the reference GPU path is
// cublasCgemm/cublasCsymm (cuda.py:314-366).
//
// Two kernels:
//  * cgemm_tc_kernel: the tensor-core path for the shape the SENSE path has
//    (SURVEY 8 row a6, cfg5 coil compression: op(M) is 12 x 48 or 48 x 12, X has
//    millions of coil-fastest columns).  The complex product is the real product
//    Y'(2m x n) = M'(2m x 2k) X'(2k x n) on the interleaved (re, im) floats; the
//    columns of X are the rows of the m16n8k8 A operand and go from global memory
//    straight into the fragment registers (one 16-byte load per lane serves two
//    k-steps, the k order inside the sum being free); alpha*op(M)' is expanded
//    once per CTA into a per-lane fragment table in shared memory.  Every product
//    is three TF32 MMAs on the (hi, lo) splits of both operands (lo*hi + hi*lo +
//    hi*hi, fp32 accumulate): complex64-level accuracy (1e-5 parity bar; plain
//    TF32 gives 1e-3).  HBM-bound by construction: 8(k+m) bytes per column.
//  * cgemm_kernel: SIMT fp32 tiles, any shape / leading dimension / alignment;
//    serves op(M) in {M, M^H} on the left and the real-symmetric right-multiply
//    through generic element strides.
#include <cstdlib>

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
    const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
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

// ---------------------------------------------------------------- tensor-core path
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

static const int TC_WARPS = 8;        // warps per CTA, one 16-column tile each per step
static const int TC_KCHUNK = 6;       // 16-float k groups held in registers at a time (6 = 48 complex rows of X)

// Real-form index conventions.  Output row o = 2i + p (p = 0 real, 1 imaginary part of Y[i, j]);
// reduction index kk = 2l + q over the interleaved floats of column j of X.  With W = alpha*op(M)[i, l]:
//   M'[2i][2l] = Re W, M'[2i][2l+1] = -Im W, M'[2i+1][2l] = Im W, M'[2i+1][2l+1] = Re W.
// Lane (g = lane/4, t = lane%4) of k group G holds the floats 16G + 4t .. 4t+3 of columns g and g+8:
// .x/.y feed k-step 2G (fragment slots t and t+4), .z/.w feed k-step 2G+1, and the fragment table
// uses the same assignment, so the permutation of the reduction order cancels.
// One 16-column tile of X as fragment registers: KC k groups for columns j0 = 16*tile + g and j0 + 8.
template <int KC>
__device__ __forceinline__ void tc_load_tile(float4 (&a0)[KC], float4 (&a1)[KC], const float *__restrict__ X, int64_t ldx2,
                                             int64_t tile, int64_t n, int G0, int K2, int g, int t) {
    const int64_t j0 = tile * 16 + g, j1 = j0 + 8;
    const float *x0 = X + j0 * ldx2 + 4 * t, *x1 = X + j1 * ldx2 + 4 * t;
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        const int G = G0 + u;
        const bool ok = 16 * G + 4 * t + 3 < K2;                  // K2 % 4 == 0: a lane's four floats are all in or all out
        a0[u] = (ok && j0 < n) ? ldg_stream4(x0 + 16 * G) : make_float4(0.f, 0.f, 0.f, 0.f);
        a1[u] = (ok && j1 < n) ? ldg_stream4(x1 + 16 * G) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// acc += X-tile fragments (k groups G0 .. G0+KC-1) times the fragment table: three TF32 MMAs per product.
template <int NT, int KC>
__device__ __forceinline__ void tc_mma_tile(float (&acc)[NT][4], const float4 (&a0)[KC], const float4 (&a1)[KC],
                                            const float4 *bfrag, int G0, int KG, int lane) {
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        const int G = G0 + u;
        if (G < KG) {                                             // warp-uniform
            const float v[2][4] = {{a0[u].x, a1[u].x, a0[u].y, a1[u].y}, {a0[u].z, a1[u].z, a0[u].w, a1[u].w}};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    ah[r] = to_tf32(v[h][r]);
                    al[r] = to_tf32(v[h][r] - __uint_as_float(ah[r]));
                }
                const float4 *bp = bfrag + (size_t)((2 * G + h) * NT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float4 b = bp[nt * 32];
                    mma_tf32(acc[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                    mma_tf32(acc[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                    mma_tf32(acc[nt], ah, __float_as_uint(b.x), __float_as_uint(b.y));
                }
            }
        }
    }
}

// accumulator (row g | g+8, columns 2t, 2t+1 of n-tile nt) = (re, im) of Y[nt*4 + t, j0 | j1]
template <int NT>
__device__ __forceinline__ void tc_store_tile(const float (&acc)[NT][4], float *__restrict__ Y, int64_t ldy2, int64_t tile,
                                              int64_t n, int m, c64 beta, int beta_zero, int g, int t) {
    const int64_t j0 = tile * 16 + g, j1 = j0 + 8;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int i = nt * 4 + t;
        if (i < m) {
            if (j0 < n) {
                c64 *yp = reinterpret_cast<c64 *>(Y + j0 * ldy2) + i;
                c64 r = mk(acc[nt][0], acc[nt][1]);
                if (!beta_zero) r = cfma(beta, *yp, r);
                __stcs(yp, r);
            }
            if (j1 < n) {
                c64 *yp = reinterpret_cast<c64 *>(Y + j1 * ldy2) + i;
                c64 r = mk(acc[nt][2], acc[nt][3]);
                if (!beta_zero) r = cfma(beta, *yp, r);
                __stcs(yp, r);
            }
        }
    }
}

// NT n-tiles of 8 real outputs (4 rows of Y); KC k groups in registers at a time; PIPE: when all of k fits
// one chunk, the loads of a warp's next tile are issued before the MMAs of the current one.
template <int NT, int KC, bool PIPE>
__global__ void __launch_bounds__(TC_WARPS * 32)
cgemm_tc_kernel(int m, int k, int64_t n, c64 alpha, const c64 *__restrict__ Mp, int64_t sa_i, int64_t sa_l, int conjA,
                const float *__restrict__ X, int64_t ldx2, c64 beta, int beta_zero, float *__restrict__ Y, int64_t ldy2) {
    extern __shared__ float4 bfrag[];                     // [k-step][n-tile][lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
    const int K2 = 2 * k, KG = (K2 + 15) / 16;
    for (int e = threadIdx.x; e < KG * 2 * NT * 32; e += blockDim.x) {
        const int lane = e & 31, nt = (e >> 5) % NT, s = (e >> 5) / NT;
        const int o = nt * 8 + (lane >> 2), i = o >> 1;
        const int l = (16 * (s >> 1) + 4 * (lane & 3) + 2 * (s & 1)) >> 1;
        float b0 = 0.f, b1 = 0.f;
        if (i < m && l < k) {
            c64 w = __ldg(Mp + i * sa_i + l * sa_l);
            if (conjA) w.y = -w.y;
            w = cmul(alpha, w);
            if (o & 1) { b0 = w.y; b1 = w.x; } else { b0 = w.x; b1 = -w.y; }
        }
        const float h0 = __uint_as_float(to_tf32(b0)), h1 = __uint_as_float(to_tf32(b1));
        bfrag[e] = make_float4(h0, h1, __uint_as_float(to_tf32(b0 - h0)), __uint_as_float(to_tf32(b1 - h1)));
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int64_t ntiles = (n + 15) / 16, stride = (int64_t)gridDim.x * TC_WARPS;
    int64_t tile = (int64_t)blockIdx.x * TC_WARPS + warp;
    if (PIPE && KG <= KC) {
        float4 c0[KC], c1[KC];
        if (tile < ntiles) tc_load_tile<KC>(c0, c1, X, ldx2, tile, n, 0, K2, g, t);
        for (; tile < ntiles; tile += stride) {
            float4 n0[KC], n1[KC];
            tc_load_tile<KC>(n0, n1, X, ldx2, tile + stride, n, 0, K2, g, t);     // past the end: all lanes load nothing
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            tc_mma_tile<NT, KC>(acc, c0, c1, bfrag, 0, KG, lane);
            tc_store_tile<NT>(acc, Y, ldy2, tile, n, m, beta, beta_zero, g, t);
#pragma unroll
            for (int u = 0; u < KC; ++u) { c0[u] = n0[u]; c1[u] = n1[u]; }
        }
        return;
    }
    for (; tile < ntiles; tile += stride) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        for (int G0 = 0; G0 < KG; G0 += KC) {
            float4 a0[KC], a1[KC];
            tc_load_tile<KC>(a0, a1, X, ldx2, tile, n, G0, K2, g, t);
            tc_mma_tile<NT, KC>(acc, a0, a1, bfrag, G0, KG, lane);
        }
        tc_store_tile<NT>(acc, Y, ldy2, tile, n, m, beta, beta_zero, g, t);
    }
}

static int g_gemm_mode = 0;           // 0 = automatic, 1 = SIMT only (tests compare the two paths)

static int g_gemm_pipe = -1;          // -1 = automatic (short k only), 0 / 1 forced (IB200_CGEMM_PIPE, tools/)

template <int NT, int KC, bool PIPE>
static int launch_tc2(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i, int64_t sa_l,
                      int conjA, const c64 *B, int64_t ldb, c64 beta, c64 *C, int64_t ldc) {
    const int KG = (int)((2 * k + 15) / 16);
    const size_t smem = (size_t)KG * 2 * NT * 32 * sizeof(float4);
    const int64_t tiles = ceil_div(n, 16);
    int64_t grid = ceil_div(tiles, TC_WARPS);
    const int64_t cap = (int64_t)sm_count() * 4;          // persistent CTAs: the fragment table is built once per CTA
    if (grid > cap) grid = cap;
    cgemm_tc_kernel<NT, KC, PIPE><<<(unsigned)grid, TC_WARPS * 32, smem, s>>>(
        (int)m, (int)k, n, alpha, A, sa_i, sa_l, conjA, (const float *)B, 2 * ldb, beta,
        (beta.x == 0.f && beta.y == 0.f) ? 1 : 0, (float *)C, 2 * ldc);
    IB200_LAUNCH_CHECK();
    return 0;
}

template <int NT>
static int launch_tc(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i, int64_t sa_l,
                     int conjA, const c64 *B, int64_t ldb, c64 beta, c64 *C, int64_t ldc) {
    if (g_gemm_pipe < 0) {
        const char *e = getenv("IB200_CGEMM_PIPE");
        g_gemm_pipe = e ? (atoi(e) ? 1 : 0) : 2;
    }
    const int KG = (int)((2 * k + 15) / 16);
    if (KG <= 2) {                                        // short columns (expansion, k = 12): little to load per tile
        if (g_gemm_pipe != 0) return launch_tc2<NT, 2, true>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
        return launch_tc2<NT, 2, false>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
    }
    if (g_gemm_pipe == 1) return launch_tc2<NT, TC_KCHUNK, true>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
    return launch_tc2<NT, TC_KCHUNK, false>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
}

// C(m x n) = alpha * op(A)(m x k) * B(k x n) + beta * C with B's columns contiguous: tall-skinny tensor-core path
static bool tc_applicable(int64_t m, int64_t n, int64_t k, const c64 *B, int64_t sb_l, int64_t sb_j) {
    if (g_gemm_mode == 1 || sb_l != 1) return false;
    if (m < 1 || m > 64 || k < 2 || (k & 1) || n < 32) return false;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (sb_j & 1)) return false;           // 16-byte loads of X columns
    const int64_t KG = (2 * k + 15) / 16, NT = ceil_div(m, 4);
    return KG * NT <= 48;                                                            // fragment table <= 48 KB
}

static int run_gemm(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i,
                    int64_t sa_l, int conjA, const c64 *B, int64_t sb_l, int64_t sb_j, c64 beta, c64 *C, int64_t ldc) {
    if (m == 0 || n == 0) return 0;
    if (tc_applicable(m, n, k, B, sb_l, sb_j)) {
        const int64_t NT = ceil_div(m, 4);
#define IB200_TC(N_) return launch_tc<N_>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_j, beta, C, ldc)
        if (NT <= 1) IB200_TC(1);
        if (NT <= 2) IB200_TC(2);
        if (NT <= 3) IB200_TC(3);
        if (NT <= 4) IB200_TC(4);
        if (NT <= 6) IB200_TC(6);
        if (NT <= 8) IB200_TC(8);
        if (NT <= 12) IB200_TC(12);
        IB200_TC(16);
#undef IB200_TC
    }
    const int64_t gx = ceil_div(n, BN), gy = ceil_div(m, BM);          // column tiles on x: n may be millions
    IB200_REQUIRE(gy <= 65535 && gx <= 2147483647LL, "cgemm: too many tiles");
    cgemm_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_l, sb_j,
                                                                 beta, (beta.x == 0.f && beta.y == 0.f) ? 1 : 0, C, ldc);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_cgemm_mode(int mode) {
    IB200_REQUIRE(mode == 0 || mode == 1, "mode is 0 (automatic) or 1 (SIMT only)");
    g_gemm_mode = mode;
    return 0;
}

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
