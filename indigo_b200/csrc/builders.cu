// Device-side construction of the two sparse factors of the -O3 SENSE tree
// (SURVEY.md section 8f rank 2): the Kaiser-Bessel gridding matrix
// G' = interp * (mod * scale) and the stored adjoint P^H of
// P = kron(I_C, mod*zpad*apod) * vstack(maps).
//
// Host reference being replaced: indigo/interp.py:19-80 (numba COO emission,
// 24 B per slot, 343 slots reserved per sample) + scipy COO->CSR + the -O2
// scipy products (indigo/transforms.py:86-96, examples/pics.py:111-126), which
// take minutes and tens of GB of host memory at 416^3 / 6.8 M samples.  Here
// each row is produced by one thread in two passes (count, exclusive scan,
// fill) and lands directly in the CSR arrays the backend's csr_matrix holds.
//
// Parity: column indices and row pointers are bit-identical to the reference's
// (rows hold the taps whose weight is non-zero, wrapped modulo the grid,
// duplicates merged, sorted ascending); values reproduce the reference's
// float64 weight arithmetic without FMA contraction and are bit-identical
// whenever no taps alias (grid extent >= taps), else equal to rounding.
#include "common.cuh"
#include "kb.cuh"

namespace ib200 {

static const int kMaxTaps = 10;      // 2*width+1 <= 9 for width <= 4

struct AxisTaps {
    int n;                 // distinct wrapped indices with non-zero weight
    int idx[kMaxTaps];     // ascending
    double w[kMaxTaps];
};

// taps of one axis: range(ceil(pos-width), floor(pos+width)), interp.py:27-37
__device__ __forceinline__ void axis_taps(double coord, int N, double width, const double *__restrict__ table,
                                          int ntab, AxisTaps &a) {
    const double pos = __dadd_rn(__dmul_rn((double)N, coord), (double)(N / 2));
    const int start = (int)ceil(__dsub_rn(pos, width));
    const int end = (int)floor(__dadd_rn(pos, width));
    a.n = 0;
    for (int t = start; t < end && a.n < kMaxTaps; ++t) {
        const double w = table ? kb_lookup(table, ntab, fabs(__dsub_rn((double)t, pos)) / width)
                               : ((fabs(__dsub_rn((double)t, pos)) / width) >= 1.0 ? 0.0 : 1.0);
        int j = t % N; if (j < 0) j += N;
        // insert sorted, merging aliases (tiny grids only)
        int k = 0;
        while (k < a.n && a.idx[k] < j) ++k;
        if (k < a.n && a.idx[k] == j) { a.w[k] += w; continue; }
        for (int s = a.n; s > k; --s) { a.idx[s] = a.idx[s - 1]; a.w[s] = a.w[s - 1]; }
        a.idx[k] = j; a.w[k] = w; a.n++;
    }
    // drop exact zeros (scipy's csr_matmat keeps only non-zero sums)
    int o = 0;
    for (int k = 0; k < a.n; ++k)
        if (a.w[k] != 0.0) { a.idx[o] = a.idx[k]; a.w[o] = a.w[k]; ++o; }
    a.n = o;
}

__global__ void __launch_bounds__(128) kb_count_kernel(int64_t m, const double *__restrict__ coord, int N0, int N1,
                                                       int N2, double width, int32_t *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    AxisTaps a;
    int c = 1;
    axis_taps(coord[3 * i + 0], N0, width, nullptr, 0, a); c *= a.n;
    axis_taps(coord[3 * i + 1], N1, width, nullptr, 0, a); c *= a.n;
    axis_taps(coord[3 * i + 2], N2, width, nullptr, 0, a); c *= a.n;
    counts[i] = c;
}

__global__ void __launch_bounds__(128) kb_fill_kernel(int64_t m, const double *__restrict__ coord, int N0, int N1,
                                                      int N2, double width, const double *__restrict__ table, int ntab,
                                                      const float *__restrict__ rowweight,
                                                      const c64 *__restrict__ colscale,
                                                      const int32_t *__restrict__ rowptr,
                                                      int32_t *__restrict__ colind, c64 *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    AxisTaps ax, ay, az;
    axis_taps(coord[3 * i + 0], N0, width, table, ntab, ax);
    axis_taps(coord[3 * i + 1], N1, width, table, ntab, ay);
    axis_taps(coord[3 * i + 2], N2, width, table, ntab, az);
    int64_t p = rowptr[i];
    const bool hasw = rowweight != nullptr;
    const float rw = hasw ? rowweight[i] : 1.f;
    for (int z = 0; z < az.n; ++z)
        for (int y = 0; y < ay.n; ++y) {
            const double wzy = __dmul_rn(az.w[z], ay.w[y]);                 // wy = wz * li(y), interp.py:47
            const int64_t rowbase = ((int64_t)az.idx[z] * N1 + ay.idx[y]) * N0;
            for (int x = 0; x < ax.n; ++x) {
                float g = (float)__dmul_rn(wzy, ax.w[x]);                   // w = wy * li(x) -> float32
                if (hasw) g = __fmul_rn(rw, g);
                const int64_t col = rowbase + ax.idx[x];
                const c64 s = colscale ? colscale[col] : mk(1.f, 0.f);
                colind[p] = (int32_t)col;
                vals[p] = mk(__fmul_rn(g, s.x), __fmul_rn(g, s.y));
                ++p;
            }
        }
}

// P^H rows: voxel r holds one entry per coil at column c*oN + zp[r], value conj(q[r]*maps[r,c])
__device__ __forceinline__ c64 cmul_exact(c64 a, c64 b) {      // no FMA contraction: matches scipy's complex<float>
    return mk(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

__global__ void __launch_bounds__(256) ph_count_kernel(int64_t n, int ncoils, const c64 *__restrict__ maps,
                                                       const c64 *__restrict__ q, int32_t *__restrict__ counts) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const c64 qr = q[r];
    int c = 0;
    for (int k = 0; k < ncoils; ++k) {
        const c64 v = cmul_exact(qr, maps[r + (int64_t)k * n]);
        c += (v.x != 0.f || v.y != 0.f);
    }
    counts[r] = c;
}

__global__ void __launch_bounds__(256) ph_fill_kernel(int64_t n, int ncoils, int64_t ogrid,
                                                      const c64 *__restrict__ maps, const c64 *__restrict__ q,
                                                      const int32_t *__restrict__ zp, const int32_t *__restrict__ rowptr,
                                                      int32_t *__restrict__ colind, c64 *__restrict__ vals) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const c64 qr = q[r];
    int64_t p = rowptr[r];
    const int64_t z = zp[r];
    for (int k = 0; k < ncoils; ++k) {
        const c64 v = cmul_exact(qr, maps[r + (int64_t)k * n]);
        if (v.x != 0.f || v.y != 0.f) {
            colind[p] = (int32_t)((int64_t)k * ogrid + z);
            vals[p] = mk(v.x, -v.y);
            ++p;
        }
    }
}

int exclusive_scan_public(cudaStream_t s, int64_t n, const int32_t *in, int32_t *out);   // csrmm.cu

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_exclusive_scan_i32(void *stream, int64_t n, const int32_t *in, int32_t *out) {
    IB200_REQUIRE(n >= 0 && out && (in || n == 0), "bad arguments");
    return exclusive_scan_public(as_stream(stream), n, in, out);
}

int ib200_kb_count(void *stream, int64_t m, const double *coord, const int64_t grid[3], double width,
                   int32_t *counts) {
    IB200_REQUIRE(m >= 0 && grid, "bad arguments");
    if (m == 0) return 0;
    IB200_REQUIRE(coord && counts, "null pointer");
    IB200_REQUIRE(width > 0 && 2 * width + 1 <= kMaxTaps, "kernel width out of range");
    IB200_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0 && grid[0] * grid[1] * grid[2] < (1LL << 31),
                  "grid must be positive and hold fewer than 2^31 points");
    kb_count_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, as_stream(stream)>>>(m, coord, (int)grid[0], (int)grid[1],
                                                                              (int)grid[2], width, counts);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_kb_fill(void *stream, int64_t m, const double *coord, const int64_t grid[3], double width,
                  const double *table, int ntable, const float *rowweight, const void *colscale,
                  const int32_t *rowptr, int32_t *colind, void *vals) {
    IB200_REQUIRE(m >= 0 && grid, "bad arguments");
    if (m == 0) return 0;
    IB200_REQUIRE(coord && table && rowptr && colind && vals, "null pointer");
    IB200_REQUIRE(ntable >= 2, "table too short");
    IB200_REQUIRE(width > 0 && 2 * width + 1 <= kMaxTaps, "kernel width out of range");
    kb_fill_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, as_stream(stream)>>>(
        m, coord, (int)grid[0], (int)grid[1], (int)grid[2], width, table, ntable, rowweight, (const c64 *)colscale,
        rowptr, colind, (c64 *)vals);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_sense_ph_count(void *stream, int64_t nvox, int ncoils, const void *maps, const void *q, int32_t *counts) {
    IB200_REQUIRE(nvox >= 0 && ncoils >= 0, "bad arguments");
    if (nvox == 0) return 0;
    IB200_REQUIRE(maps && q && counts, "null pointer");
    ph_count_kernel<<<(unsigned)ceil_div(nvox, 256), 256, 0, as_stream(stream)>>>(nvox, ncoils, (const c64 *)maps,
                                                                                (const c64 *)q, counts);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_sense_ph_fill(void *stream, int64_t nvox, int ncoils, int64_t ogrid, const void *maps, const void *q,
                        const int32_t *zp, const int32_t *rowptr, int32_t *colind, void *vals) {
    IB200_REQUIRE(nvox >= 0 && ncoils >= 0, "bad arguments");
    if (nvox == 0) return 0;
    IB200_REQUIRE(maps && q && zp && rowptr && colind && vals, "null pointer");
    IB200_REQUIRE((int64_t)ncoils * ogrid < (1LL << 31), "coils x grid must stay below 2^31 (int32 columns)");
    ph_fill_kernel<<<(unsigned)ceil_div(nvox, 256), 256, 0, as_stream(stream)>>>(
        nvox, ncoils, ogrid, (const c64 *)maps, (const c64 *)q, zp, rowptr, colind, (c64 *)vals);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
