// onemm / cdiamm / max: the small operators the Backend interface requires
// beyond the SENSE hot path (used by indigo's phase-space and MPI examples).
// Interfaces replaced: Backend.onemm (backend.py:528-533), Backend.cdiamm
// (backend.py:521-526), Backend.max (backend.py:734-736); numpy semantics
// np.py:95-97,129-145.  The reference's sm_61 kernels (_customgpu.cu:7-143)
// are a thread-per-row loop with a shared-memory tree; here onemm is a
// warp-shuffle block reduction and the DIA product is one thread per output
// element with the diagonals as the inner loop (coalesced along rows).
#include "common.cuh"

namespace ib200 {

__global__ void __launch_bounds__(256) onemm_kernel(int64_t m, int64_t k, c64 alpha, const c64 *__restrict__ X,
                                                    int64_t ldx, c64 beta, int beta_zero, c64 *__restrict__ Y,
                                                    int64_t ldy) {
    __shared__ float sh[2][8];
    __shared__ c64 total;
    const c64 *x = X + (int64_t)blockIdx.x * ldx;
    c64 *y = Y + (int64_t)blockIdx.x * ldy;
    float re = 0.f, im = 0.f;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) { const c64 v = __ldg(x + i); re += v.x; im += v.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = re; sh[1][w] = im; }
    __syncthreads();
    if (w == 0) {
        re = lane < 8 ? sh[0][lane] : 0.f; im = lane < 8 ? sh[1][lane] : 0.f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
        if (lane == 0) total = cmul(alpha, mk(re, im));
    }
    __syncthreads();
    const c64 t = total;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) y[i] = beta_zero ? t : cfma(beta, y[i], t);
}

// one thread per output element; blockIdx.y = right-hand-side column
template <bool ADJ>
__global__ void __launch_bounds__(256) diamm_kernel(int64_t yrows, int64_t xrows, int64_t dcols, int64_t dpitch, int noff,
                                                    const int32_t *__restrict__ offsets, const c64 *__restrict__ data,
                                                    c64 alpha, const c64 *__restrict__ X, int64_t ldx, c64 beta,
                                                    int beta_zero, c64 *__restrict__ Y, int64_t ldy) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= yrows) return;
    const c64 *x = X + (int64_t)blockIdx.y * ldx;
    c64 *y = Y + (int64_t)blockIdx.y * ldy;
    c64 acc = mk(0.f, 0.f);
    for (int d = 0; d < noff; ++d) {
        const int64_t off = __ldg(offsets + d);
        if (!ADJ) {
            const int64_t j = r + off;                  // column of A; data is indexed by column
            if (j >= 0 && j < xrows && j < dcols) acc = cfma(__ldg(data + j + (int64_t)d * dpitch), __ldg(x + j), acc);
        } else {
            const int64_t i = r - off;                  // row of A; r is the column
            if (i >= 0 && i < xrows && r < dcols) acc = cfmac(__ldg(data + r + (int64_t)d * dpitch), __ldg(x + i), acc);
        }
    }
    c64 out = cmul(alpha, acc);
    if (!beta_zero) out = cfma(beta, y[r], out);
    y[r] = out;
}

__global__ void __launch_bounds__(256) fmax_kernel(int64_t n, float val, float *__restrict__ a) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) a[i] = fmaxf(a[i], val);
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_onemm(void *stream, int64_t m, int64_t ncols, int64_t k, float ar, float ai, const void *X, int64_t ldx,
                float br, float bi, void *Y, int64_t ldy) {
    IB200_REQUIRE(m >= 0 && ncols >= 0 && k >= 0, "negative dimension");
    if (m == 0 || ncols == 0) return 0;
    IB200_REQUIRE(Y && (X || k == 0), "null pointer");
    IB200_REQUIRE(ncols < (1LL << 31), "too many columns");
    onemm_kernel<<<(unsigned)ncols, 256, 0, as_stream(stream)>>>(m, k, mk(ar, ai), (const c64 *)X, ldx, mk(br, bi),
                                                                (br == 0.f && bi == 0.f) ? 1 : 0, (c64 *)Y, ldy);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_cdiamm(void *stream, int adjoint, int64_t m, int64_t k, int64_t ncols, int64_t noffsets,
                 const int32_t *offsets, const void *data, int64_t data_cols, int64_t data_pitch, float ar, float ai,
                 const void *X, int64_t ldx, float br, float bi, void *Y, int64_t ldy) {
    IB200_REQUIRE(m >= 0 && k >= 0 && ncols >= 0 && noffsets >= 0, "negative dimension");
    IB200_REQUIRE(data_cols >= 0 && data_pitch >= data_cols, "diagonal pitch smaller than the diagonal length");
    const int64_t yrows = adjoint ? k : m, xrows = adjoint ? m : k;
    if (yrows == 0 || ncols == 0) return 0;
    IB200_REQUIRE(Y && X && (noffsets == 0 || (offsets && data)), "null pointer");
    IB200_REQUIRE(ncols <= 65535, "more than 65535 right-hand sides");
    const dim3 grid((unsigned)ceil_div(yrows, 256), (unsigned)ncols);
    const int b0 = (br == 0.f && bi == 0.f) ? 1 : 0;
    if (adjoint)
        diamm_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(yrows, xrows, data_cols, data_pitch, (int)noffsets, offsets, (const c64 *)data,
                                                               mk(ar, ai), (const c64 *)X, ldx, mk(br, bi), b0, (c64 *)Y, ldy);
    else
        diamm_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(yrows, xrows, data_cols, data_pitch, (int)noffsets, offsets, (const c64 *)data,
                                                                mk(ar, ai), (const c64 *)X, ldx, mk(br, bi), b0, (c64 *)Y, ldy);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_fmax(void *stream, int64_t nfloats, float val, void *arr) {
    IB200_REQUIRE(nfloats >= 0, "negative length");
    if (nfloats == 0) return 0;
    IB200_REQUIRE(arr, "null pointer");
    int64_t g = ceil_div(nfloats, 256 * 4); const int64_t cap = (int64_t)sm_count() * 8; if (g > cap) g = cap;
    fmax_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(nfloats, val, (float *)arr);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
