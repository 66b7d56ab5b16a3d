// Separable Kaiser-Bessel gridding gather: the forward interpolation grid -> samples of the fused
// SENSE recipe without a stored matrix.
//
// Reference being replaced: the product ccsrmm(G') of the -O3 tree, where
// G' = interp * mod * scale is the CSR matrix that indigo/interp.py:19-80 emits tap by tap:
// value(sample, tap) = li(x) * li(y) * li(z) (interp.py:44-59) times the centring phase and the
// 1/sqrt(prod oN) scale of backend.py:349-364, which the -O2 recipe folds in (transforms.py:86-96).
// Every factor is a product over the three axes, so a row of G' is the outer product of three
// short vectors.  A stored entry costs 8-12 bytes of HBM traffic and one extra load per tap; a
// record of 3 x 6 weights + 3 base indices costs 96 bytes per SAMPLE instead of 1000-1500 bytes,
// and the tap addresses follow from the base indices by adds ("structured KB
// layout", the re-layout that keeps coil sharding from re-reading a 10 GB matrix on every GPU).
//
// Parity: weights are the reference's float64 table interpolation (kb.cuh, same arithmetic as the
// CSR builder) rounded to float32 per axis, times the per-axis real centring factors; the product
// of the three differs from the reference's float32 value by a few ulp (tests/test_gpu_fused.py
// compares against the CSR product and the numpy oracle).  Only real-valued G' (grid extents that
// make the centring phase +-1) and kernels of at most 6 taps per axis are served; anything else
// stays on the stored-matrix path.  On an axis shorter than the footprint (2-D problems are carried
// with a two-point z axis) the taps alias; the record then holds the summed weight of every point
// of the axis, where the reference's COO -> CSR conversion sums the duplicate entries.
#include "common.cuh"
#include "kb.cuh"
#include "pk2.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace ib200 {

// record r describes sample perm[r] (perm = NULL: identity)
__global__ void __launch_bounds__(128) kb_records_kernel(int64_t m, const double *__restrict__ coord, int N0, int N1,
                                                         int N2, double width, const double *__restrict__ table,
                                                         int ntab, const float *__restrict__ rowweight,
                                                         const float *__restrict__ f0, const float *__restrict__ f1,
                                                         const float *__restrict__ f2,
                                                         const int32_t *__restrict__ perm, int out_is_record,
                                                         KbRecord *__restrict__ rec, int *flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int64_t i = perm ? (int64_t)perm[r] : r;
    KbRecord q;
    int cnt[3], first[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int N = d == 0 ? N0 : d == 1 ? N1 : N2;
        const float *f = d == 0 ? f0 : d == 1 ? f1 : f2;
        float *w = d == 0 ? q.wx : d == 1 ? q.wy : q.wz;
        // taps range(ceil(pos - width), floor(pos + width)), interp.py:27-37
        const double pos = __dadd_rn(__dmul_rn((double)N, coord[3 * i + d]), (double)(N / 2));
        const int start = (int)ceil(__dsub_rn(pos, width));
        const int end = (int)floor(__dadd_rn(pos, width));
        int n = end - start;
        if (n < 0 || n > kKbTaps) { atomicOr(flag, 1); n = n < 0 ? 0 : kKbTaps; }
        int j = start % N; if (j < 0) j += N;
        first[d] = j;
        // an axis shorter than the footprint (the two-point z axis of a 2-D problem): the taps alias onto its N points,
        // tap t lands on point (first + t) mod N; the record holds the summed weights of the N points
        const int np = n > N ? N : n;
        cnt[d] = np;
#pragma unroll
        for (int t = 0; t < kKbTaps; ++t) w[t] = 0.f;
#pragma unroll
        for (int t = 0; t < kKbTaps; ++t) {
            if (t < n) {
                const double wv = kb_lookup(table, ntab, fabs(__dsub_rn((double)(start + t), pos)) / width);
                const float v = __fmul_rn((float)wv, f[j]);
                if (n > N) w[t % N] += v; else w[t] = v;
            }
            if (++j >= N) j = 0;
        }
    }
    if (rowweight) {
        const float rw = rowweight[i];
#pragma unroll
        for (int t = 0; t < kKbTaps; ++t) q.wz[t] = __fmul_rn(rw, q.wz[t]);
    }
    q.ix0 = first[0]; q.iy0 = first[1]; q.iz0 = first[2];
    q.out = out_is_record ? (int32_t)r : (int32_t)i;
    q.ntaps = cnt[0] | (cnt[1] << 8) | (cnt[2] << 16);
    q.pad = 0;
    rec[r] = q;
}

// (A packed-FFMA2 form of this row sum with incrementally stepped row pointers was measured in round 2,
// round 2, session 2: 14 % fewer warp instructions but IPC 2.0 -> 1.5 and 5.0 -> 5.9 ms at cfg3 -- the scalar
// form below lets ptxas hoist the next row's loads over this row's FMAs within the 64-register budget.)
// VC coils per lane: one 8-byte or one 16-byte load per tap
template <int VC> struct KbVec;
template <> struct KbVec<1> {
    float x[1], y[1];
    __device__ __forceinline__ void load(const char *p) { const float2 v = __ldg(reinterpret_cast<const float2 *>(p)); x[0] = v.x; y[0] = v.y; }
};
template <> struct KbVec<2> {
    float x[2], y[2];
    __device__ __forceinline__ void load(const char *p) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p)); x[0] = v.x; y[0] = v.y; x[1] = v.z; y[1] = v.w;
    }
};

// one (z, y) row of taps: weighted sum over the x taps, folded into the accumulator with wjk = wz*wy.
// DENSE: the x taps are consecutive grid points (no wrap-around in this warp) and the pitch of a
// grid point is the compile-time constant PITCH, so the tap addresses are immediates.
template <int VC, bool DENSE, int PITCH>
__device__ __forceinline__ void kb_row(const char *rp, const uint32_t (&xo)[kKbTaps], const float (&wx)[kKbTaps],
                                       bool sixth, float wjk, float (&ax)[VC], float (&ay)[VC]) {
    KbVec<VC> q[kKbTaps];
    const char *r0 = rp + xo[0];
#pragma unroll
    for (int t = 0; t < kKbTaps - 1; ++t) q[t].load(DENSE ? r0 + t * PITCH : rp + xo[t]);
    float tx[VC], ty[VC];
#pragma unroll
    for (int v = 0; v < VC; ++v) { tx[v] = wx[0] * q[0].x[v]; ty[v] = wx[0] * q[0].y[v]; }
#pragma unroll
    for (int t = 1; t < kKbTaps - 1; ++t)
#pragma unroll
        for (int v = 0; v < VC; ++v) { tx[v] = fmaf(wx[t], q[t].x[v], tx[v]); ty[v] = fmaf(wx[t], q[t].y[v], ty[v]); }
    if (sixth) {                                                     // warp-uniform, rare
        q[kKbTaps - 1].load(DENSE ? r0 + (kKbTaps - 1) * PITCH : rp + xo[kKbTaps - 1]);
#pragma unroll
        for (int v = 0; v < VC; ++v) {
            tx[v] = fmaf(wx[kKbTaps - 1], q[kKbTaps - 1].x[v], tx[v]);
            ty[v] = fmaf(wx[kKbTaps - 1], q[kKbTaps - 1].y[v], ty[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < VC; ++v) { ax[v] = fmaf(wjk, tx[v], ax[v]); ay[v] = fmaf(wjk, ty[v], ay[v]); }
}

// Y[out(r)*ypitch + c] = alpha * sum_{k,j,i} wz[k] wy[j] wx[i] * grid[iz_k][iy_j][ix_i][c]
//
// CL lanes per sample (VC coils each), 32/CL samples per warp; a warp walks `iters` batches of
// consecutive records, i.e. neighbouring samples of one grid tile when the records were sorted by
// tile.  The x taps of one (z, y) row are issued together (5-6 independent loads per lane), their
// weighted sum is folded into the accumulator with the row's wz*wy.  Trip counts are the maximum
// over the samples of the warp (5 taps per axis except for samples that sit exactly on a grid
// line, whose sixth tap has weight zero), so control flow is warp-uniform.
template <int CL, int VC>
__global__ void __launch_bounds__(256, 4) kb_gather_kernel(int64_t m, int C, c64 alpha, const KbRecord *__restrict__ rec,
                                                           const c64 *__restrict__ grid, uint32_t pitch, int n0, int n1,
                                                           int n2, c64 *__restrict__ Y, int64_t ypitch, int iters) {
    constexpr int SPW = 32 / CL;
    constexpr int PITCH = CL * VC * 8;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const int sub = lane & (CL - 1), slot = lane / CL;
    const int coil = sub * VC;
    const bool coil_ok = coil < C;
    const char *gb = reinterpret_cast<const char *>(grid + (coil_ok ? coil : 0));
    const uint64_t rowbytes = (uint64_t)n0 * pitch;
    const bool dense_pitch = pitch == (uint32_t)PITCH;
    // the 8 warps of the CTA sweep its records together (32 consecutive samples per step): what is in L1
    // at any time is the footprint of one or two neighbouring tiles, not of 8 unrelated ones
    const int64_t s0 = (int64_t)blockIdx.x * (int64_t)(8 * SPW * iters) + warp * SPW;
    for (int it = 0; it < iters; ++it) {
        const int64_t rbase = s0 + (int64_t)it * (8 * SPW);
        if (rbase >= m) break;                                       // warp-uniform
        const int64_t r = rbase + slot;
        if (sub == 0 && r + 8 * SPW < m) {                           // next step's record: two lines, pulled into L2 now
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + r + 8 * SPW));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(rec + r + 8 * SPW) + 64));
        }
        float wx[kKbTaps], wy[kKbTaps], wz[kKbTaps];
        int ix0 = 0, iy0 = 0, iz0 = 0, out = -1, nt = 0;
        if (r < m) {
            const float4 *p = reinterpret_cast<const float4 *>(rec + r);
            const float4 a = __ldcs(p), b = __ldcs(p + 1), c = __ldcs(p + 2), d = __ldcs(p + 3), e = __ldcs(p + 4);
            const int4 g = __ldcs(reinterpret_cast<const int4 *>(p + 5));
            wx[0] = a.x; wx[1] = a.y; wx[2] = a.z; wx[3] = a.w; wx[4] = b.x; wx[5] = b.y;
            wy[0] = b.z; wy[1] = b.w; wy[2] = c.x; wy[3] = c.y; wy[4] = c.z; wy[5] = c.w;
            wz[0] = d.x; wz[1] = d.y; wz[2] = d.z; wz[3] = d.w; wz[4] = e.x; wz[5] = e.y;
            ix0 = __float_as_int(e.z); iy0 = __float_as_int(e.w); iz0 = g.x; out = g.y; nt = g.z;
        } else {
#pragma unroll
            for (int t = 0; t < kKbTaps; ++t) { wx[t] = 0.f; wy[t] = 0.f; wz[t] = 0.f; }
        }
        const int nxm = __reduce_max_sync(FULL, nt & 255), nym = __reduce_max_sync(FULL, (nt >> 8) & 255),
                  nzm = __reduce_max_sync(FULL, (nt >> 16) & 255);
        const bool sixth = nxm > kKbTaps - 1;
        const bool dense = dense_pitch && __all_sync(FULL, ix0 + kKbTaps <= n0);
        uint32_t xo[kKbTaps];
        {
            int j = ix0;
#pragma unroll
            for (int t = 0; t < kKbTaps; ++t) { xo[t] = (uint32_t)j * pitch; if (++j >= n0) j = 0; }
        }
        float ax[VC], ay[VC];
#pragma unroll
        for (int v = 0; v < VC; ++v) { ax[v] = 0.f; ay[v] = 0.f; }
        int iz = iz0;
        for (int k = 0; k < nzm; ++k) {
            const float wk = wz[0];
#pragma unroll
            for (int t = 0; t < kKbTaps - 1; ++t) wz[t] = wz[t + 1];
            wz[kKbTaps - 1] = 0.f;
            const uint64_t zrow = (uint64_t)iz * (uint64_t)n1;
            int iy = iy0;
            if (dense) {
#pragma unroll
                for (int j = 0; j < kKbTaps; ++j) {
                    if (j < nym) {                                   // warp-uniform
                        kb_row<VC, true, PITCH>(gb + (zrow + (uint64_t)iy) * rowbytes, xo, wx, sixth, wk * wy[j], ax, ay);
                        if (++iy >= n1) iy = 0;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kKbTaps; ++j) {
                    if (j < nym) {
                        kb_row<VC, false, PITCH>(gb + (zrow + (uint64_t)iy) * rowbytes, xo, wx, sixth, wk * wy[j], ax, ay);
                        if (++iy >= n1) iy = 0;
                    }
                }
            }
            if (++iz >= n2) iz = 0;
        }
        if (out >= 0 && coil_ok) {
            c64 *yp = Y + (int64_t)out * ypitch + coil;
            if (VC == 2) {
                const c64 o0 = cmul(alpha, mk(ax[0], ay[0])), o1 = cmul(alpha, mk(ax[VC - 1], ay[VC - 1]));
                __stcs(reinterpret_cast<float4 *>(yp), make_float4(o0.x, o0.y, o1.x, o1.y));
            } else {
                __stcs(yp, cmul(alpha, mk(ax[0], ay[0])));
            }
        }
    }
}

template <int CL, int VC>
static int launch_kb_gather(cudaStream_t s, int64_t m, int C, c64 alpha, const KbRecord *rec, const c64 *grid,
                            int64_t xpitch, const int64_t n[3], c64 *Y, int64_t ypitch) {
    constexpr int SPW = 32 / CL;
    // ~32 samples per warp: long enough to amortise the launch of a CTA, short enough for >= 20 waves
    int iters = 32 / SPW; if (iters < 1) iters = 1;
    if (m < (1 << 20)) iters = 1;                                    // small problems: more, shorter CTAs (latency-bound)
    const int64_t per_cta = (int64_t)8 * SPW * iters;
    const int64_t blocks = ceil_div(m, per_cta);
    IB200_REQUIRE(blocks < (1LL << 31), "too many samples for one launch");
    kb_gather_kernel<CL, VC><<<(unsigned)blocks, 256, 0, s>>>(m, C, alpha, rec, grid, (uint32_t)(xpitch * sizeof(c64)),
                                                              (int)n[0], (int)n[1], (int)n[2], Y, ypitch, iters);
    IB200_LAUNCH_CHECK();
    return 0;
}

// ---- k-space support windows ---------------------------------------------------------------------------
// A trajectory touches only part of the oversampled grid (a radial "kooshball" the inscribed sphere,
// 52 % of the cube; a stack of spirals a cylinder).  Grid points outside the support are never read
// by G' and are exactly zero in G'^H k, so the last forward FFT pass need not write them, the adjoint
// gather need not store zeros for them and the first inverse pass need not read them.  Per block of
// bx x by grid columns the support is summarised as one interval [lo, hi) along z (hull over the
// block, rounded to multiples of tz so that whole tiles of the stored adjoint fall inside or outside).
__global__ void __launch_bounds__(256) support_init_kernel(int n, int32_t *lo, int32_t *hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { lo[i] = 0x7fffffff; hi[i] = 0; }
}

__global__ void __launch_bounds__(256) support_mark_kernel(int64_t kp, const int32_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ rowmap, int n0, int n1, int n2,
                                                           int bx, int by, int tz, int nbx, int32_t *lo, int32_t *hi) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= kp) return;
    if (rowptr[r + 1] == rowptr[r]) return;
    const int64_t g = rowmap[r];
    if (g < 0) return;
    const int x = (int)(g % n0), y = (int)((g / n0) % n1), z = (int)(g / ((int64_t)n0 * n1));
    const int b = (y / by) * nbx + x / bx;
    const int zl = (z / tz) * tz;
    int zh = zl + tz; if (zh > n2) zh = n2;
    atomicMin(lo + b, zl);
    atomicMax(hi + b, zh);
}

__global__ void __launch_bounds__(256) support_expand_kernel(int n0, int n1, int bx, int by, int nbx,
                                                             const int32_t *__restrict__ lo,
                                                             const int32_t *__restrict__ hi, int32_t *__restrict__ win,
                                                             unsigned long long *inside) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (p < n0 * n1) {
        const int x = p % n0, y = p / n0;
        const int b = (y / by) * nbx + x / bx;
        int l = lo[b], h = hi[b];
        if (h <= l) { l = 0; h = 0; }
        win[2 * p] = l; win[2 * p + 1] = h;
        mine = (unsigned long long)(h - l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(inside, mine);
}

__global__ void __launch_bounds__(256) support_rowmap_kernel(int64_t kp, const int32_t *__restrict__ rowmap, int n0, int n1,
                                                             const int32_t *__restrict__ win,
                                                             int32_t *__restrict__ rowmap_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= kp) return;
    const int64_t g = rowmap[r];
    int32_t out = -1;
    if (g >= 0) {
        const int64_t p = g % ((int64_t)n0 * n1);
        const int z = (int)(g / ((int64_t)n0 * n1));
        if (z >= win[2 * p] && z < win[2 * p + 1]) out = (int32_t)g;
    }
    rowmap_out[r] = out;
}

// ---- matrix-free construction: sample order and support windows without the CSR matrix ---------------------
// key of a sample = two-level tile rank (tile_rank2_kernel, csrmm_il.cu) of its first tap, computed from the
// coordinates with the arithmetic of kb_records_kernel; rows[i] = i
__global__ void __launch_bounds__(256) kb_base_key_kernel(int64_t m, const double *__restrict__ coord, int N0, int N1,
                                                          int N2, double width, int t0, int t1, int t2, int s0, int s1,
                                                          int s2, int32_t *__restrict__ keys, int32_t *__restrict__ rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int N = d == 0 ? N0 : d == 1 ? N1 : N2;
        const double pos = __dadd_rn(__dmul_rn((double)N, coord[3 * i + d]), (double)(N / 2));
        int j = (int)ceil(__dsub_rn(pos, width)) % N; if (j < 0) j += N;
        c[d] = j;
    }
    const int nt0 = (N0 + t0 - 1) / t0, nt1 = (N1 + t1 - 1) / t1;
    const int ns0 = (nt0 + s0 - 1) / s0, ns1 = (nt1 + s1 - 1) / s1;
    const int tvol = t0 * t1 * t2, svol = s0 * s1 * s2;
    const int tx = c[0] / t0, ty = c[1] / t1, tz = c[2] / t2;
    const int64_t super = ((int64_t)(tz / s2) * ns1 + ty / s1) * ns0 + tx / s0;
    const int intile = ((c[2] % t2) * t1 + (c[1] % t1)) * t0 + (c[0] % t0);
    const int insuper = ((tz % s2) * s1 + (ty % s1)) * s0 + (tx % s0);
    keys[i] = (int32_t)((super * svol + insuper) * tvol + intile);
    rows[i] = (int32_t)i;
}

// hull along z of the grid points every record touches, per block of bx x by grid columns (what support_mark_kernel
// derives from the rows of the stored adjoint)
__global__ void __launch_bounds__(256) support_mark_records_kernel(int64_t m, const KbRecord *__restrict__ rec, int n0,
                                                                   int n1, int n2, int bx, int by, int tz, int nbx,
                                                                   int32_t *lo, int32_t *hi) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const KbRecord *q = rec + r;
    const int nt = q->ntaps, nx = nt & 255, ny = (nt >> 8) & 255, nz = (nt >> 16) & 255;
    if (nx == 0 || ny == 0 || nz == 0) return;
    int zl = 0x7fffffff, zh = 0, z = q->iz0;
    for (int k = 0; k < nz; ++k) {
        const int l = (z / tz) * tz; int h = l + tz; if (h > n2) h = n2;
        zl = l < zl ? l : zl; zh = h > zh ? h : zh;
        if (++z >= n2) z = 0;
    }
    int y = q->iy0;
    for (int j = 0; j < ny; ++j) {
        int x = q->ix0, last = -1;
        for (int i = 0; i < nx; ++i) {
            const int b = (y / by) * nbx + x / bx;
            if (b != last) { atomicMin(lo + b, zl); atomicMax(hi + b, zh); last = b; }
            if (++x >= n0) x = 0;
        }
        if (++y >= n1) y = 0;
    }
}

static int kb_pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_kb_record_bytes(void) { return (int)sizeof(KbRecord); }

int ib200_kb_records(void *stream, int64_t m, const double *coord, const int64_t grid[3], double width,
                     const double *table, int ntable, const float *rowweight, const float *f0, const float *f1,
                     const float *f2, const int32_t *perm, int out_is_record, void *records, int *host_flag) {
    IB200_REQUIRE(m >= 0 && grid && host_flag, "bad arguments");
    *host_flag = 0;
    if (m == 0) return 0;
    IB200_REQUIRE(coord && table && f0 && f1 && f2 && records, "null pointer");
    IB200_REQUIRE(ntable >= 2, "table too short");
    IB200_REQUIRE(width > 0, "kernel width must be positive");
    IB200_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0 && grid[0] * grid[1] * grid[2] < (1LL << 31),
                  "grid must be positive and hold fewer than 2^31 points");
    cudaStream_t s = as_stream(stream);
    int *flag = nullptr;
    IB200_TRY(cudaMalloc(&flag, sizeof(int)));
    cudaMemsetAsync(flag, 0, sizeof(int), s);
    kb_records_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, coord, (int)grid[0], (int)grid[1], (int)grid[2], width,
                                                                table, ntable, rowweight, f0, f1, f2, perm, out_is_record,
                                                                (KbRecord *)records, flag);
    count_launch();
    cudaMemcpyAsync(host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(flag);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_grid_support_windows(void *stream, const int64_t grid[3], int64_t kp, const int32_t *rowptr,
                               const int32_t *rowmap, const int64_t block[3], int32_t *win, int32_t *rowmap_out,
                               int64_t *host_inside) {
    IB200_REQUIRE(grid && block && host_inside, "null pointer");
    *host_inside = 0;
    IB200_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0 && grid[0] * grid[1] < (1LL << 30), "bad grid");
    IB200_REQUIRE(block[0] > 0 && block[1] > 0 && block[2] > 0, "bad block");
    IB200_REQUIRE(kp >= 0 && kp < (1LL << 31), "bad row count");
    IB200_REQUIRE(rowptr && rowmap && win && rowmap_out, "null pointer");
    const int n0 = (int)grid[0], n1 = (int)grid[1], n2 = (int)grid[2];
    const int bx = (int)block[0], by = (int)block[1], tz = (int)block[2];
    const int nbx = (n0 + bx - 1) / bx, nby = (n1 + by - 1) / by, nb = nbx * nby;
    cudaStream_t s = as_stream(stream);
    int32_t *lo = nullptr;
    unsigned long long *inside = nullptr;
    IB200_TRY(cudaMalloc(&lo, (size_t)nb * 2 * sizeof(int32_t) + 16));
    int32_t *hi = lo + nb;
    cudaError_t e = cudaMalloc(&inside, sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(lo); IB200_TRY(e); }
    cudaMemsetAsync(inside, 0, sizeof(unsigned long long), s);
    support_init_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, s>>>(nb, lo, hi);
    if (kp > 0)
        support_mark_kernel<<<(unsigned)ceil_div(kp, 256), 256, 0, s>>>(kp, rowptr, rowmap, n0, n1, n2, bx, by, tz, nbx, lo, hi);
    support_expand_kernel<<<(unsigned)ceil_div((int64_t)n0 * n1, 256), 256, 0, s>>>(n0, n1, bx, by, nbx, lo, hi, win, inside);
    if (kp > 0)
        support_rowmap_kernel<<<(unsigned)ceil_div(kp, 256), 256, 0, s>>>(kp, rowmap, n0, n1, win, rowmap_out);
    count_launch(4);
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, inside, sizeof(h), cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    cudaFree(lo); cudaFree(inside);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    *host_inside = (int64_t)h;
    return 0;
}

int ib200_kb_sample_order(void *stream, int64_t m, const double *coord, const int64_t grid[3], double width,
                          const int64_t tile[3], const int64_t super[3], int32_t *perm) {
    IB200_REQUIRE(m >= 0 && m < (1LL << 31) && grid && tile && super, "bad arguments");
    if (m == 0) return 0;
    IB200_REQUIRE(coord && perm && width > 0, "null pointer");
    int64_t nranks = 1;
    for (int d = 0; d < 3; ++d) {
        IB200_REQUIRE(grid[d] > 0 && tile[d] > 0 && tile[d] <= 64 && super[d] > 0 && super[d] <= 64, "bad grid / tile extent");
        nranks *= ceil_div(ceil_div(grid[d], tile[d]), super[d]) * super[d] * tile[d];
    }
    IB200_REQUIRE(nranks < (1LL << 31), "padded grid must hold fewer than 2^31 points");
    cudaStream_t s = as_stream(stream);
    int32_t *buf = nullptr;
    IB200_TRY(cudaMalloc(&buf, (size_t)m * sizeof(int32_t) * 4));
    int32_t *keys_a = buf, *keys_b = buf + m, *rows_a = buf + 2 * m, *rows_b = buf + 3 * m;
    kb_base_key_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(m, coord, (int)grid[0], (int)grid[1], (int)grid[2], width,
                                                                 (int)tile[0], (int)tile[1], (int)tile[2], (int)super[0],
                                                                 (int)super[1], (int)super[2], keys_a, rows_a);
    count_launch();
    int end_bit = 1; while ((1LL << end_bit) <= nranks) ++end_bit;
    cub::DoubleBuffer<int32_t> dk(keys_a, keys_b), dv(rows_a, rows_b);
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)m, 0, end_bit, s);
    void *tmp = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)m, 0, end_bit, s);
    if (e == cudaSuccess) {
        count_launch(end_bit / 8 + 2);
        e = cudaMemcpyAsync(perm, dv.Current(), (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToDevice, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tmp); cudaFree(buf);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_kb_support_windows(void *stream, int64_t m, const void *records, const int64_t grid[3], int64_t kp,
                             const int32_t *rowmap, const int64_t block[3], int32_t *win, int32_t *rowmap_out,
                             int64_t *host_inside) {
    IB200_REQUIRE(grid && block && host_inside, "null pointer");
    *host_inside = 0;
    IB200_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0 && grid[0] * grid[1] < (1LL << 30), "bad grid");
    IB200_REQUIRE(block[0] > 0 && block[1] > 0 && block[2] > 0, "bad block");
    IB200_REQUIRE(kp >= 0 && kp < (1LL << 31) && m >= 0, "bad row count");
    IB200_REQUIRE(rowmap && win && rowmap_out && (records || m == 0), "null pointer");
    const int n0 = (int)grid[0], n1 = (int)grid[1], n2 = (int)grid[2];
    const int bx = (int)block[0], by = (int)block[1], tz = (int)block[2];
    const int nbx = (n0 + bx - 1) / bx, nby = (n1 + by - 1) / by, nb = nbx * nby;
    cudaStream_t s = as_stream(stream);
    int32_t *lo = nullptr;
    unsigned long long *inside = nullptr;
    IB200_TRY(cudaMalloc(&lo, (size_t)nb * 2 * sizeof(int32_t) + 16));
    int32_t *hi = lo + nb;
    cudaError_t e = cudaMalloc(&inside, sizeof(unsigned long long));
    if (e != cudaSuccess) { cudaFree(lo); IB200_TRY(e); }
    cudaMemsetAsync(inside, 0, sizeof(unsigned long long), s);
    support_init_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, s>>>(nb, lo, hi);
    if (m > 0)
        support_mark_records_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(m, (const KbRecord *)records, n0, n1, n2, bx, by,
                                                                              tz, nbx, lo, hi);
    support_expand_kernel<<<(unsigned)ceil_div((int64_t)n0 * n1, 256), 256, 0, s>>>(n0, n1, bx, by, nbx, lo, hi, win, inside);
    if (kp > 0)
        support_rowmap_kernel<<<(unsigned)ceil_div(kp, 256), 256, 0, s>>>(kp, rowmap, n0, n1, win, rowmap_out);
    count_launch(4);
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, inside, sizeof(h), cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    cudaFree(lo); cudaFree(inside);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    *host_inside = (int64_t)h;
    return 0;
}

int ib200_kb_gather(void *stream, int64_t m, int64_t ncols, float ar, float ai, const void *records,
                    const void *grid_il, int64_t xpitch, const int64_t grid[3], void *Yil, int64_t ypitch) {
    IB200_RANGE("ib200_kb_gather");
    IB200_REQUIRE(m >= 0 && ncols >= 0 && grid, "bad arguments");
    if (m == 0 || ncols == 0) return 0;
    IB200_REQUIRE(records && grid_il && Yil, "null pointer");
    IB200_REQUIRE(ncols <= 64, "separable gather serves at most 64 columns per call");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols, "pitch smaller than the column count");
    IB200_REQUIRE(grid[0] > 0 && grid[1] > 0 && grid[2] > 0, "bad grid");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) * grid[0] < (1LL << 32), "grid row too long");
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    const bool vec2 = ncols % 2 == 0 && xpitch % 2 == 0 && ypitch % 2 == 0 && ((uintptr_t)grid_il % 16) == 0 &&
                      ((uintptr_t)Yil % 16) == 0;
    const int CL = kb_pow2_ceil(vec2 ? ncols / 2 : ncols);
    IB200_REQUIRE(CL <= 32, "too many columns for one lane group");
#define IB200_KB_CASE(cl)                                                                                            \
    case cl:                                                                                                         \
        return vec2 ? launch_kb_gather<cl, 2>(s, m, (int)ncols, alpha, (const KbRecord *)records, (const c64 *)grid_il, \
                                              xpitch, grid, (c64 *)Yil, ypitch)                                      \
                    : launch_kb_gather<cl, 1>(s, m, (int)ncols, alpha, (const KbRecord *)records, (const c64 *)grid_il, \
                                              xpitch, grid, (c64 *)Yil, ypitch)
    switch (CL) {
        IB200_KB_CASE(1); IB200_KB_CASE(2); IB200_KB_CASE(4); IB200_KB_CASE(8); IB200_KB_CASE(16); IB200_KB_CASE(32);
    }
#undef IB200_KB_CASE
    set_error("internal: no separable gather for CL=%d", CL);
    return IB200_E_UNSUPPORTED;
}

}  // extern "C"
