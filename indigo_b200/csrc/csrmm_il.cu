// Coil-interleaved CSR SpMM: the multi-column fast path of Backend.ccsrmm.
//
// Interface served: Backend.ccsrmm (indigo/backends/backend.py:514-519) with
// 2..32 right-hand-side columns (coils).  The interface hands X and Y over
// column-major, coil = slowest axis (operators.py:27-31), so a CSR entry needs
// one 8-byte gather per coil, `ld*8` bytes apart: at 125 entries per gridding
// row and 16 coils every warp-wide load touches ~13 sectors in ~7 lines for 256
// useful bytes and the kernel is bound by L1/L2 request throughput, not HBM
// (measured: 873 GB/s algorithmic, 13 % of peak, profiles/r01_s1_*).
//
// Re-layout (north star: "multi-coil right-hand sides are read with vectorised,
// coalesced loads"): X is transposed once into Xil[row][coil] (one pass at copy
// bandwidth), after which the C coils of a grid point are ONE contiguous
// C*8-byte segment -- a full 128-byte line for 16 coils.  A group of GL lanes
// owns a matrix row: (index, value) pairs are fetched GL at a time with one
// coalesced load and handed round by shuffles, lane (s, c) accumulates coil c
// of every (GL/CL)-th entry, and the partial sums of the GL/CL slots are folded
// with xor-shuffles.  The same kernel serves the forward gridding (rows =
// samples, 125 entries) and the adjoint through the stored conjugate transpose
// (rows = grid points, 0..2400 entries), so no atomics are needed anywhere.
#include "common.cuh"

#include <cstring>

namespace ib200 {

// ---------------------------------------------------------------------------
// X(rows x C, column-major, ld)  ->  Xil[r*pitch + c]
static const int kTrRows = 64;

template <bool POW2>
__global__ void __launch_bounds__(256) interleave_kernel(int64_t rows, int C, int log2C, const c64 *__restrict__ X,
                                                         int64_t ldx, c64 *__restrict__ Xil, int64_t pitch,
                                                         const int32_t *__restrict__ perm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *tile = reinterpret_cast<c64 *>(smem_raw);                 // [C][kTrRows + 1]
    const int64_t r0 = (int64_t)blockIdx.x * kTrRows;
    const int nr = rows - r0 < kTrRows ? (int)(rows - r0) : kTrRows;
    const int r = threadIdx.x & (kTrRows - 1);
    if (r < nr)
        for (int c = threadIdx.x / kTrRows; c < C; c += 256 / kTrRows)
            tile[c * (kTrRows + 1) + r] = __ldg(X + r0 + r + (int64_t)c * ldx);
    __syncthreads();
    const int n = nr * C;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int rr = POW2 ? (i >> log2C) : (i / C);
        const int c = POW2 ? (i & (C - 1)) : (i - rr * C);
        const int64_t dst = perm ? (int64_t)__ldg(perm + r0 + rr) : r0 + rr;
        Xil[dst * pitch + c] = tile[c * (kTrRows + 1) + rr];
    }
}

// Y(rows x C, column-major, ld) = Yil^T + beta * Y      (beta == 0: Y is never read)
template <bool POW2>
__global__ void __launch_bounds__(256) deinterleave_kernel(int64_t rows, int C, int log2C,
                                                           const c64 *__restrict__ Yil, int64_t pitch, c64 beta,
                                                           int beta_zero, c64 *__restrict__ Y, int64_t ldy,
                                                           const int32_t *__restrict__ perm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *tile = reinterpret_cast<c64 *>(smem_raw);
    const int64_t r0 = (int64_t)blockIdx.x * kTrRows;
    const int nr = rows - r0 < kTrRows ? (int)(rows - r0) : kTrRows;
    const int n = nr * C;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int rr = POW2 ? (i >> log2C) : (i / C);
        const int c = POW2 ? (i & (C - 1)) : (i - rr * C);
        const int64_t src = perm ? (int64_t)__ldg(perm + r0 + rr) : r0 + rr;
        tile[c * (kTrRows + 1) + rr] = __ldg(Yil + src * pitch + c);
    }
    __syncthreads();
    const int r = threadIdx.x & (kTrRows - 1);
    if (r < nr)
        for (int c = threadIdx.x / kTrRows; c < C; c += 256 / kTrRows) {
            c64 *yp = Y + r0 + r + (int64_t)c * ldy;
            c64 v = tile[c * (kTrRows + 1) + r];
            if (!beta_zero) v = cfma(beta, *yp, v);
            *yp = v;
        }
}

// ---------------------------------------------------------------------------
// Yil[out(row)*ypitch + c] = alpha * sum_p vals[p] * Xil[colind[p]*xpitch + c]
//
// GL lanes per row, CL lanes per entry (one per coil), NP = GL/CL entries in flight per step.
// A CTA owns GPB*rpg consecutive rows (GPB = 256/GL groups, rpg rows per group, interleaved so
// that the groups always work on neighbouring rows): consecutive samples of a readout -- or the
// grid points of one tile when the stored adjoint was built in tile-major order -- share most of
// their operand lines, which then hit in L1.  Control flow is warp-uniform (trip counts are the
// maximum over the groups of a warp, short rows are predicated off), so groups of one warp never
// serialise.  Matrix entries are streamed with evict-first loads: they are used once, while the
// operand rows are re-used from L2.  `rowmap` (optional) sends row r to output row rowmap[r]
// (negative: padding row, nothing is written).
template <int U>
struct IlLoads {
    c64 x[U];
    float vx[U], vy[U];
};

template <int GL, int CL>
__global__ void __launch_bounds__(256) csrmm_il_kernel(int64_t m, int C, c64 alpha, const c64 *__restrict__ vals,
                                                       const int32_t *__restrict__ colind,
                                                       const int32_t *__restrict__ rowptr,
                                                       const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                       c64 *__restrict__ Yil, int64_t ypitch,
                                                       const int32_t *__restrict__ rowmap, int rpg, int long_thresh) {
    constexpr int NP = GL / CL, GPB = 256 / GL;
    constexpr int U = CL >= 4 ? 4 : CL;                             // loads issued back to back
    constexpr unsigned FULL = 0xffffffffu;
    const int gl = (int)(threadIdx.x & (GL - 1));                   // lane within the group
    const int coil = gl & (CL - 1);
    const int slot = gl / CL;
    const int group = (int)(threadIdx.x / GL);
    const char *xb = reinterpret_cast<const char *>(Xil + (coil < C ? coil : 0));
    const int64_t row0 = (int64_t)blockIdx.x * ((int64_t)GPB * rpg) + group;
    for (int i = 0; i < rpg; ++i) {
        const int64_t row = row0 + (int64_t)i * GPB;
        int p0 = 0, len = 0;
        if (row < m) { p0 = __ldg(rowptr + row); len = __ldg(rowptr + row + 1) - p0; }
        const bool is_long = len > long_thresh;                     // left to csrmm_il_long_kernel
        if (is_long) len = 0;
        int maxlen = len;
#pragma unroll
        for (int o = GL; o < 32; o <<= 1) { const int t = __shfl_xor_sync(FULL, maxlen, o); maxlen = t > maxlen ? t : maxlen; }
        c64 acc = mk(0.f, 0.f);
        for (int base = 0; base < maxlen; base += GL) {
            const int left = len - base;                            // <= 0 once this group's row is done
            int myc = 0;
            c64 myv = mk(0.f, 0.f);
            if (gl < left) { myc = __ldcs(colind + p0 + base + gl); myv = __ldcs(vals + p0 + base + gl); }
            int minleft = left, maxleft = left;
#pragma unroll
            for (int o = GL; o < 32; o <<= 1) {
                const int a = __shfl_xor_sync(FULL, minleft, o), b = __shfl_xor_sync(FULL, maxleft, o);
                minleft = a < minleft ? a : minleft; maxleft = b > maxleft ? b : maxleft;
            }
            if (minleft >= GL) {                                    // every group of the warp has a full batch
#pragma unroll
                for (int t0 = 0; t0 < CL; t0 += U) {
                    IlLoads<U> q;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int src = (t0 + u) * NP + slot;
                        const unsigned c = (unsigned)__shfl_sync(FULL, myc, src, GL);
                        q.vx[u] = __shfl_sync(FULL, myv.x, src, GL);
                        q.vy[u] = __shfl_sync(FULL, myv.y, src, GL);
                        q.x[u] = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c * xpitch_bytes));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) acc = cfma(mk(q.vx[u], q.vy[u]), q.x[u], acc);
                }
            } else {
                // ragged batch: lanes past the end of their row point at the row's first entry with
                // weight 0, so every step is an unpredicated load of an address this row reads anyway
                { const int c0 = __shfl_sync(FULL, myc, 0, GL); if (gl >= left) myc = c0; }
                const int nsteps = maxleft >= GL ? CL : (maxleft + NP - 1) / NP;      // warp-uniform, <= CL
                for (int t0 = 0; t0 < nsteps; t0 += U) {
                    IlLoads<U> q;
                    unsigned cc[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int src = (t0 + u) * NP + slot;       // < GL because CL is a multiple of U
                        cc[u] = (unsigned)__shfl_sync(FULL, myc, src, GL);
                        q.vx[u] = __shfl_sync(FULL, myv.x, src, GL);
                        q.vy[u] = __shfl_sync(FULL, myv.y, src, GL);
                    }
                    if (left > 0) {                                 // uniform per group
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            q.x[u] = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)cc[u] * xpitch_bytes));
#pragma unroll
                        for (int u = 0; u < U; ++u) acc = cfma(mk(q.vx[u], q.vy[u]), q.x[u], acc);
                    }
                }
            }
        }
#pragma unroll
        for (int o = CL; o < GL; o <<= 1) {
            acc.x += __shfl_xor_sync(FULL, acc.x, o, GL);
            acc.y += __shfl_xor_sync(FULL, acc.y, o, GL);
        }
        if (row < m && slot == 0 && coil < C && !is_long) {
            const int64_t out = rowmap ? (int64_t)__ldg(rowmap + row) : row;
            if (out >= 0) __stcs(Yil + out * ypitch + coil, cmul(alpha, acc));
        }
    }
}

// ---------------------------------------------------------------------------
// Real-weight variant: entries are packed (column, weight) pairs of 8 bytes.
// Gridding matrices are real up to the centring phase, which is +-1 on grids whose extents are
// multiples of four, so two thirds of the matrix bytes and half of the multiplies suffice.
// Every lane fetches its entry itself: the CL lanes of an entry read the same 8 bytes (one
// broadcast request), which removes the batch load + shuffle hand-out of the complex kernel and
// all of its per-row bookkeeping -- what matters for the stored adjoint, whose rows hold 12
// entries on average.  Entries of one row are consumed four at a time (independent loads).
struct __align__(8) PackedEntry { int32_t col; float w; };

__device__ __forceinline__ PackedEntry ld_entry(const PackedEntry *p) {
    const int2 v = __ldg(reinterpret_cast<const int2 *>(p));
    PackedEntry e; e.col = v.x; e.w = __int_as_float(v.y);
    return e;
}

template <int GL, int CL>
__global__ void __launch_bounds__(256) csrmm_ilr_kernel(int64_t m, int C, c64 alpha,
                                                        const PackedEntry *__restrict__ ent,
                                                        const int32_t *__restrict__ rowptr,
                                                        const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                        c64 *__restrict__ Yil, int64_t ypitch,
                                                        const int32_t *__restrict__ rowmap, int rpg, int long_thresh) {
    constexpr int NP = GL / CL, GPB = 256 / GL;
    constexpr unsigned FULL = 0xffffffffu;
    const int gl = (int)(threadIdx.x & (GL - 1));
    const int coil = gl & (CL - 1);
    const int slot = gl / CL;
    const int group = (int)(threadIdx.x / GL);
    const char *xb = reinterpret_cast<const char *>(Xil + (coil < C ? coil : 0));
    const int64_t row0 = (int64_t)blockIdx.x * ((int64_t)GPB * rpg) + group;
    for (int i = 0; i < rpg; ++i) {
        const int64_t row = row0 + (int64_t)i * GPB;
        int p = 0, p1 = 0;
        if (row < m) { p = __ldg(rowptr + row); p1 = __ldg(rowptr + row + 1); }
        const bool is_long = p1 - p > long_thresh;                  // left to csrmm_il_long_kernel
        if (is_long) p1 = p;
        p += slot;
        float ax = 0.f, ay = 0.f, bx = 0.f, by = 0.f;
        for (; p + 3 * NP < p1; p += 4 * NP) {
            const PackedEntry e0 = ld_entry(ent + p), e1 = ld_entry(ent + p + NP), e2 = ld_entry(ent + p + 2 * NP),
                              e3 = ld_entry(ent + p + 3 * NP);
            const c64 x0 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)(uint32_t)e0.col * xpitch_bytes));
            const c64 x1 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)(uint32_t)e1.col * xpitch_bytes));
            const c64 x2 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)(uint32_t)e2.col * xpitch_bytes));
            const c64 x3 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)(uint32_t)e3.col * xpitch_bytes));
            ax = fmaf(e0.w, x0.x, ax); ay = fmaf(e0.w, x0.y, ay);
            bx = fmaf(e1.w, x1.x, bx); by = fmaf(e1.w, x1.y, by);
            ax = fmaf(e2.w, x2.x, ax); ay = fmaf(e2.w, x2.y, ay);
            bx = fmaf(e3.w, x3.x, bx); by = fmaf(e3.w, x3.y, by);
        }
        for (; p < p1; p += NP) {
            const PackedEntry e0 = ld_entry(ent + p);
            const c64 x0 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)(uint32_t)e0.col * xpitch_bytes));
            ax = fmaf(e0.w, x0.x, ax); ay = fmaf(e0.w, x0.y, ay);
        }
        c64 acc = mk(ax + bx, ay + by);
        if (NP > 1) {
            __syncwarp();
#pragma unroll
            for (int o = CL; o < GL; o <<= 1) {
                acc.x += __shfl_xor_sync(FULL, acc.x, o, GL);
                acc.y += __shfl_xor_sync(FULL, acc.y, o, GL);
            }
        }
        if (row < m && slot == 0 && coil < C && !is_long) {
            const int64_t out = rowmap ? (int64_t)__ldg(rowmap + row) : row;
            if (out >= 0) __stcs(Yil + out * ypitch + coil, cmul(alpha, acc));
        }
    }
}

template <int GL, int CL>
static int launch_ilr(cudaStream_t s, int64_t m, int C, c64 alpha, const PackedEntry *ent, const int32_t *rowptr,
                      const c64 *Xil, int64_t xpitch, c64 *Yil, int64_t ypitch, const int32_t *rowmap, int rpg,
                      int long_thresh) {
    const int64_t rows_per_cta = (int64_t)(256 / GL) * rpg;
    const int64_t blocks = ceil_div(m, rows_per_cta);
    IB200_REQUIRE(blocks < (1LL << 31), "matrix too large for one launch");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    csrmm_ilr_kernel<GL, CL><<<(unsigned)blocks, 256, 0, s>>>(m, C, alpha, ent, rowptr, Xil,
                                                              (uint32_t)(xpitch * sizeof(c64)), Yil, ypitch, rowmap, rpg,
                                                              long_thresh);
    IB200_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// Long rows.  The stored adjoint of a radial trajectory has a handful of enormous rows: every
// spoke passes through the k-space centre, so the central grid points of cfg3 collect ~100 000
// entries each while the mean is 12.  One 16-lane group walking such a row alone would take as
// long as the rest of the matrix; rows above `long_thresh` entries are therefore skipped by the
// kernels above and each gets a whole CTA here (256/CL entries in flight per step, shared-memory
// fold of the partial sums, plain store).
template <int CL, bool PACKED>
__global__ void __launch_bounds__(256) csrmm_il_long_kernel(const int32_t *__restrict__ longrows, int C, c64 alpha,
                                                            const PackedEntry *__restrict__ ent,
                                                            const c64 *__restrict__ vals,
                                                            const int32_t *__restrict__ colind,
                                                            const int32_t *__restrict__ rowptr,
                                                            const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                            c64 *__restrict__ Yil, int64_t ypitch,
                                                            const int32_t *__restrict__ rowmap) {
    constexpr int NS = 256 / CL;                                    // entries in flight per step
    __shared__ c64 part[256];
    const int64_t row = longrows[blockIdx.x];
    const int coil = (int)(threadIdx.x & (CL - 1));
    const int slot = (int)(threadIdx.x / CL);
    const char *xb = reinterpret_cast<const char *>(Xil + (coil < C ? coil : 0));
    int p = __ldg(rowptr + row) + slot;
    const int p1 = __ldg(rowptr + row + 1);
    c64 a0 = mk(0.f, 0.f), a1 = mk(0.f, 0.f);
    auto fetch = [&](int q, unsigned &col, c64 &v) {
        if (PACKED) { const PackedEntry e = ld_entry(ent + q); col = (unsigned)e.col; v = mk(e.w, 0.f); }
        else { col = (unsigned)__ldg(colind + q); v = __ldg(vals + q); }
    };
    for (; p + 3 * NS < p1; p += 4 * NS) {
        unsigned c0, c1, c2, c3; c64 v0, v1, v2, v3;
        fetch(p, c0, v0); fetch(p + NS, c1, v1); fetch(p + 2 * NS, c2, v2); fetch(p + 3 * NS, c3, v3);
        const c64 x0 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c0 * xpitch_bytes));
        const c64 x1 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c1 * xpitch_bytes));
        const c64 x2 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c2 * xpitch_bytes));
        const c64 x3 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c3 * xpitch_bytes));
        if (PACKED) {
            a0.x = fmaf(v0.x, x0.x, a0.x); a0.y = fmaf(v0.x, x0.y, a0.y);
            a1.x = fmaf(v1.x, x1.x, a1.x); a1.y = fmaf(v1.x, x1.y, a1.y);
            a0.x = fmaf(v2.x, x2.x, a0.x); a0.y = fmaf(v2.x, x2.y, a0.y);
            a1.x = fmaf(v3.x, x3.x, a1.x); a1.y = fmaf(v3.x, x3.y, a1.y);
        } else {
            a0 = cfma(v0, x0, a0); a1 = cfma(v1, x1, a1); a0 = cfma(v2, x2, a0); a1 = cfma(v3, x3, a1);
        }
    }
    for (; p < p1; p += NS) {
        unsigned c0; c64 v0;
        fetch(p, c0, v0);
        const c64 x0 = __ldg(reinterpret_cast<const c64 *>(xb + (uint64_t)c0 * xpitch_bytes));
        a0 = cfma(v0, x0, a0);
    }
    part[threadIdx.x] = cadd(a0, a1);
    __syncthreads();
    for (int h = NS / 2; h > 0; h >>= 1) {                          // fold the slots (tree order: deterministic)
        if (slot < h) part[threadIdx.x] = cadd(part[threadIdx.x], part[threadIdx.x + h * CL]);
        __syncthreads();
    }
    if (slot == 0 && coil < C) {
        const int64_t out = rowmap ? (int64_t)__ldg(rowmap + row) : row;
        if (out >= 0) Yil[out * ypitch + coil] = cmul(alpha, part[threadIdx.x]);
    }
}

__global__ void __launch_bounds__(256) long_rows_kernel(int64_t m, const int32_t *__restrict__ rowptr, int thresh,
                                                        int32_t *__restrict__ list, int capacity, int *count) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    if (rowptr[r + 1] - rowptr[r] > thresh) {
        const int at = atomicAdd(count, 1);
        if (at < capacity) list[at] = (int32_t)r;
    }
}

// The long-row kernel is latency bound (one CTA per row, few CTAs) and independent of the main
// gather: it runs on the device's side stream (core.cu), forked from and joined back into the caller's
// stream with events, so that both kernels share the SMs instead of running back to back.
template <bool PACKED>
static int launch_long(cudaStream_t s, int CL, int nlong, const int32_t *longrows, int C, c64 alpha, const void *ent,
                       const c64 *vals, const int32_t *colind, const int32_t *rowptr, const c64 *Xil, int64_t xpitch,
                       c64 *Yil, int64_t ypitch, const int32_t *rowmap) {
    if (nlong <= 0) return 0;
    IB200_REQUIRE(longrows != nullptr, "long-row list missing");
    const uint32_t pb = (uint32_t)(xpitch * sizeof(c64));
    const PackedEntry *e = (const PackedEntry *)ent;
#define IB200_LONG(cl) case cl: csrmm_il_long_kernel<cl, PACKED><<<(unsigned)nlong, 256, 0, s>>>(longrows, C, alpha, e, vals, colind, rowptr, Xil, pb, Yil, ypitch, rowmap); break
    switch (CL) { IB200_LONG(1); IB200_LONG(2); IB200_LONG(4); IB200_LONG(8); IB200_LONG(16); IB200_LONG(32);
                  default: set_error("internal: bad CL"); return IB200_E_UNSUPPORTED; }
#undef IB200_LONG
    IB200_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// Staged variant of the real-weight gather (north star: "column indices staged through shared
// memory").  A CTA owns R = GPB*RPG consecutive rows, i.e. one contiguous range of packed entries;
// the range is streamed into shared memory with 16-byte cp.async copies (perfectly coalesced, no
// register staging), after which an entry costs one LDS instead of a dependent global load: the
// operand gathers of a row are then all independent and are issued U at a time.  Rows spanning
// two chunks keep their partial sums in registers.  The packed array must be readable two
// entries past its end (16-byte granules).
static const int kStageChunk = 4096;                               // entries per chunk (32 KB)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// VC = coils per lane (1: 8-byte operand loads, 2: 16-byte loads of two neighbouring coils, which
// halves the instructions per entry and doubles the rows a warp serves; needs an even coil count).
template <int VC> struct CoilVec;
template <> struct CoilVec<1> {
    float x[1], y[1];
    __device__ __forceinline__ void load(const char *p) { const c64 v = __ldg(reinterpret_cast<const c64 *>(p)); x[0] = v.x; y[0] = v.y; }
};
template <> struct CoilVec<2> {
    float x[2], y[2];
    __device__ __forceinline__ void load(const char *p) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p)); x[0] = v.x; y[0] = v.y; x[1] = v.z; y[1] = v.w;
    }
};

template <int GL, int CL, int RPG, int U, int VC>
__global__ void __launch_bounds__(256) csrmm_ils_kernel(int64_t m, int C, c64 alpha,
                                                        const PackedEntry *__restrict__ ent,
                                                        const int32_t *__restrict__ rowptr,
                                                        const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                        c64 *__restrict__ Yil, int64_t ypitch,
                                                        const int32_t *__restrict__ rowmap, int long_thresh) {
    constexpr int NP = GL / CL, GPB = 256 / GL, R = GPB * RPG;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) PackedEntry sent[kStageChunk + 2];
    __shared__ int srp[R + 1];
    const int gl = (int)(threadIdx.x & (GL - 1));
    const int coil = (gl & (CL - 1)) * VC;                          // first coil of this lane
    const int slot = gl / CL;
    const int group = (int)(threadIdx.x / GL);
    const char *xb = reinterpret_cast<const char *>(Xil + (coil < C ? coil : 0));
    const int64_t r0 = (int64_t)blockIdx.x * R;
    const int nr = m - r0 < R ? (int)(m - r0) : R;
    for (int i = threadIdx.x; i <= nr; i += 256) srp[i] = __ldg(rowptr + r0 + i);
    __syncthreads();
    const int E0 = srp[0], E1 = srp[nr];
    if (E0 == E1) {                                                 // no entries at all (grid tiles outside the sampled region)
        for (int i = threadIdx.x; i < nr * C; i += 256) {
            const int row = i / C, cc = i - row * C;
            const int64_t gr = r0 + row;
            const int64_t out = rowmap ? (int64_t)__ldg(rowmap + gr) : gr;
            if (out >= 0) __stcs(Yil + out * ypitch + cc, mk(0.f, 0.f));
        }
        return;
    }
    float ax[RPG][VC], ay[RPG][VC];
#pragma unroll
    for (int i = 0; i < RPG; ++i)
#pragma unroll
        for (int v = 0; v < VC; ++v) { ax[i][v] = 0.f; ay[i][v] = 0.f; }
    for (int c0 = E0 & ~1; c0 < E1; c0 += kStageChunk) {
        const int cend = c0 + kStageChunk < E1 ? c0 + kStageChunk : E1;        // entries [c0, cend) are usable
        for (int q = c0 + 2 * (int)threadIdx.x; q < cend; q += 512) cp_async16(sent + (q - c0), ent + q);
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RPG; ++i) {
            const int row = group + i * GPB;
            if (row < nr) {
                int a = srp[row], b = srp[row + 1];
                if (b - a > long_thresh) b = a;                                 // left to csrmm_il_long_kernel
                a = a > c0 ? a : c0; b = b < cend ? b : cend;
                const PackedEntry *se = sent - c0;
                int p = a + slot;
                for (; p + (U - 1) * NP < b; p += U * NP) {
                    PackedEntry e[U];
                    CoilVec<VC> x[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) e[u] = se[p + u * NP];
#pragma unroll
                    for (int u = 0; u < U; ++u) x[u].load(xb + (uint64_t)(uint32_t)e[u].col * xpitch_bytes);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int v = 0; v < VC; ++v) {
                            ax[i][v] = fmaf(e[u].w, x[u].x[v], ax[i][v]); ay[i][v] = fmaf(e[u].w, x[u].y[v], ay[i][v]);
                        }
                }
                for (; p < b; p += NP) {
                    const PackedEntry e0 = se[p];
                    CoilVec<VC> x0;
                    x0.load(xb + (uint64_t)(uint32_t)e0.col * xpitch_bytes);
#pragma unroll
                    for (int v = 0; v < VC; ++v) { ax[i][v] = fmaf(e0.w, x0.x[v], ax[i][v]); ay[i][v] = fmaf(e0.w, x0.y[v], ay[i][v]); }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < RPG; ++i) {
        const int row = group + i * GPB;
        c64 acc[VC];
#pragma unroll
        for (int v = 0; v < VC; ++v) {
            acc[v] = mk(ax[i][v], ay[i][v]);
            if (NP > 1) {
#pragma unroll
                for (int o = CL; o < GL; o <<= 1) {
                    acc[v].x += __shfl_xor_sync(FULL, acc[v].x, o, GL);
                    acc[v].y += __shfl_xor_sync(FULL, acc[v].y, o, GL);
                }
            }
        }
        if (row < nr && slot == 0 && coil < C && srp[row + 1] - srp[row] <= long_thresh) {
            const int64_t gr = r0 + row;
            const int64_t out = rowmap ? (int64_t)__ldg(rowmap + gr) : gr;
            if (out >= 0) {
                c64 *yp = Yil + out * ypitch + coil;
                if (VC == 2) {
                    const c64 o0 = cmul(alpha, acc[0]), o1 = cmul(alpha, acc[VC - 1]);
                    __stcs(reinterpret_cast<float4 *>(yp), make_float4(o0.x, o0.y, o1.x, o1.y));
                } else {
                    __stcs(yp, cmul(alpha, acc[0]));
                }
            }
        }
    }
}

template <int GL, int CL, int RPG, int U, int VC>
static int launch_ils(cudaStream_t s, int64_t m, int C, c64 alpha, const PackedEntry *ent, const int32_t *rowptr,
                      const c64 *Xil, int64_t xpitch, c64 *Yil, int64_t ypitch, const int32_t *rowmap, int long_thresh) {
    const int64_t rows_per_cta = (int64_t)(256 / GL) * RPG;
    const int64_t blocks = ceil_div(m, rows_per_cta);
    IB200_REQUIRE(blocks < (1LL << 31), "matrix too large for one launch");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    csrmm_ils_kernel<GL, CL, RPG, U, VC><<<(unsigned)blocks, 256, 0, s>>>(m, C, alpha, ent, rowptr, Xil,
                                                                          (uint32_t)(xpitch * sizeof(c64)), Yil, ypitch,
                                                                          rowmap, long_thresh);
    IB200_LAUNCH_CHECK();
    return 0;
}

// packed[p] = (colind[p], Re vals[p]);  stats[0] = max |Re|, stats[1] = max |Im| (as float bit patterns, >= 0)
__global__ void __launch_bounds__(256) pack_real_kernel(int64_t nnz, const c64 *__restrict__ vals,
                                                        const int32_t *__restrict__ colind,
                                                        PackedEntry *__restrict__ packed, unsigned *stats) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    float mre = 0.f, mim = 0.f;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += nth) {
        const c64 v = vals[p];
        PackedEntry e; e.col = colind[p]; e.w = v.x;
        packed[p] = e;
        mre = fmaxf(mre, fabsf(v.x)); mim = fmaxf(mim, fabsf(v.y));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mre = fmaxf(mre, __shfl_xor_sync(0xffffffffu, mre, o));
        mim = fmaxf(mim, __shfl_xor_sync(0xffffffffu, mim, o));
    }
    if ((threadIdx.x & 31) == 0) {                                // non-negative floats order like unsigned ints
        atomicMax(stats, __float_as_uint(mre));
        atomicMax(stats + 1, __float_as_uint(mim));
    }
}

template <int GL, int CL>
static int launch_il(cudaStream_t s, int64_t m, int C, c64 alpha, const c64 *vals, const int32_t *colind,
                     const int32_t *rowptr, const c64 *Xil, int64_t xpitch, c64 *Yil, int64_t ypitch,
                     const int32_t *rowmap, int rpg, int long_thresh) {
    const int64_t rows_per_cta = (int64_t)(256 / GL) * rpg;
    const int64_t blocks = ceil_div(m, rows_per_cta);
    IB200_REQUIRE(blocks < (1LL << 31), "matrix too large for one launch");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    csrmm_il_kernel<GL, CL><<<(unsigned)blocks, 256, 0, s>>>(m, C, alpha, vals, colind, rowptr, Xil,
                                                             (uint32_t)(xpitch * sizeof(c64)), Yil, ypitch, rowmap, rpg,
                                                             long_thresh);
    IB200_LAUNCH_CHECK();
    return 0;
}

// padded tile-major rank of every point of a 3-D grid (x fastest) and its inverse
__global__ void __launch_bounds__(256) tile_rank_kernel(int n0, int n1, int n2, int t0, int t1, int t2,
                                                        int32_t *__restrict__ colrank, int32_t *__restrict__ rowmap) {
    const int nt0 = (n0 + t0 - 1) / t0, nt1 = (n1 + t1 - 1) / t1, nt2 = (n2 + t2 - 1) / t2;
    const int64_t padded = (int64_t)nt0 * nt1 * nt2 * t0 * t1 * t2;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int tvol = t0 * t1 * t2;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < padded; r += nth) {
        const int64_t tile = r / tvol;
        const int w = (int)(r - tile * tvol);
        const int wx = w % t0, wy = (w / t0) % t1, wz = w / (t0 * t1);
        const int tx = (int)(tile % nt0), ty = (int)((tile / nt0) % nt1), tz = (int)(tile / ((int64_t)nt0 * nt1));
        const int x = tx * t0 + wx, y = ty * t1 + wy, z = tz * t2 + wz;
        if (x < n0 && y < n1 && z < n2) {
            const int64_t g = ((int64_t)z * n1 + y) * n0 + x;
            rowmap[r] = (int32_t)g;
            colrank[g] = (int32_t)r;
        } else {
            rowmap[r] = -1;
        }
    }
}

// two-level variant: tiles of t0 x t1 x t2 points grouped into super-tiles of s0 x s1 x s2 tiles; rank =
// super-tile major, then tile major inside the super-tile, then point inside the tile.  Used as the sort
// key of the SAMPLES: the CTAs in flight at any time (a few thousand consecutive samples each) then
// cover a compact block of the grid whose lines stay in L2, instead of a 13-plane slab of the whole grid.
__global__ void __launch_bounds__(256) tile_rank2_kernel(int n0, int n1, int n2, int t0, int t1, int t2, int s0, int s1,
                                                         int s2, int32_t *__restrict__ colrank) {
    const int64_t total = (int64_t)n0 * n1 * n2;
    const int nt0 = (n0 + t0 - 1) / t0, nt1 = (n1 + t1 - 1) / t1;
    const int ns0 = (nt0 + s0 - 1) / s0, ns1 = (nt1 + s1 - 1) / s1;
    const int tvol = t0 * t1 * t2, svol = s0 * s1 * s2;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += nth) {
        const int x = (int)(g % n0), y = (int)((g / n0) % n1), z = (int)(g / ((int64_t)n0 * n1));
        const int tx = x / t0, ty = y / t1, tz = z / t2;
        const int sx = tx / s0, sy = ty / s1, sz = tz / s2;
        const int64_t super = ((int64_t)sz * ns1 + sy) * ns0 + sx;
        const int intile = ((z % t2) * t1 + (y % t1)) * t0 + (x % t0);
        const int insuper = ((tz % s2) * s1 + (ty % s1)) * s0 + (tx % s0);
        colrank[g] = (int32_t)((super * svol + insuper) * tvol + intile);
    }
}

__global__ void __launch_bounds__(256) invert_perm_kernel(int64_t n, const int32_t *__restrict__ perm, int32_t *__restrict__ inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (int32_t)i;
}

static int pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }
static int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

int ilr_main(cudaStream_t s, int staged, int CL, double avg, int rows_per_group, int64_t m, int64_t ncols, c64 alpha,
             const void *packed, const int32_t *rowptr, const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch,
             const int32_t *rowmap, int long_thresh);

}  // namespace ib200

using namespace ib200;

extern "C" {

static int interleave_impl(void *stream, int64_t rows, int64_t ncols, const void *X, int64_t ldx, void *Xil, int64_t pitch,
                           const int32_t *perm) {
    IB200_REQUIRE(rows >= 0 && ncols >= 0, "negative dimension");
    if (rows == 0 || ncols == 0) return 0;
    IB200_REQUIRE(X && Xil, "null pointer");
    IB200_REQUIRE(ncols <= 1024 && pitch >= ncols && (ncols == 1 || ldx >= rows), "bad column count / pitch / ld");
    const int C = (int)ncols;
    const size_t smem = (size_t)C * (kTrRows + 1) * sizeof(c64);
    IB200_REQUIRE((int64_t)smem <= 48 * 1024, "too many columns for the transposing tile (max 94)");
    const int64_t blocks = ceil_div(rows, kTrRows);
    IB200_REQUIRE(blocks < (1LL << 31), "too many rows for one launch");
    const bool p2 = (C & (C - 1)) == 0;
    if (p2) interleave_kernel<true><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(rows, C, ilog2(C), (const c64 *)X, ldx, (c64 *)Xil, pitch, perm);
    else    interleave_kernel<false><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(rows, C, 0, (const c64 *)X, ldx, (c64 *)Xil, pitch, perm);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_interleave(void *stream, int64_t rows, int64_t ncols, const void *X, int64_t ldx, void *Xil, int64_t pitch) {
    IB200_RANGE("ib200_interleave");
    return interleave_impl(stream, rows, ncols, X, ldx, Xil, pitch, nullptr);
}

int ib200_interleave_rows(void *stream, int64_t rows, int64_t ncols, const void *X, int64_t ldx, void *Xil, int64_t pitch,
                          const int32_t *perm) {
    return interleave_impl(stream, rows, ncols, X, ldx, Xil, pitch, perm);
}

static int deinterleave_impl(void *stream, int64_t rows, int64_t ncols, const void *Yil, int64_t pitch, float br, float bi,
                             void *Y, int64_t ldy, const int32_t *perm) {
    IB200_REQUIRE(rows >= 0 && ncols >= 0, "negative dimension");
    if (rows == 0 || ncols == 0) return 0;
    IB200_REQUIRE(Y && Yil, "null pointer");
    IB200_REQUIRE(ncols <= 1024 && pitch >= ncols && (ncols == 1 || ldy >= rows), "bad column count / pitch / ld");
    const int C = (int)ncols;
    const size_t smem = (size_t)C * (kTrRows + 1) * sizeof(c64);
    IB200_REQUIRE((int64_t)smem <= 48 * 1024, "too many columns for the transposing tile (max 94)");
    const int64_t blocks = ceil_div(rows, kTrRows);
    IB200_REQUIRE(blocks < (1LL << 31), "too many rows for one launch");
    const int b0 = (br == 0.f && bi == 0.f) ? 1 : 0;
    const bool p2 = (C & (C - 1)) == 0;
    if (p2) deinterleave_kernel<true><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(rows, C, ilog2(C), (const c64 *)Yil, pitch, mk(br, bi), b0, (c64 *)Y, ldy, perm);
    else    deinterleave_kernel<false><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(rows, C, 0, (const c64 *)Yil, pitch, mk(br, bi), b0, (c64 *)Y, ldy, perm);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_deinterleave(void *stream, int64_t rows, int64_t ncols, const void *Yil, int64_t pitch, float br, float bi,
                       void *Y, int64_t ldy) {
    IB200_RANGE("ib200_deinterleave");
    return deinterleave_impl(stream, rows, ncols, Yil, pitch, br, bi, Y, ldy, nullptr);
}

int ib200_deinterleave_rows(void *stream, int64_t rows, int64_t ncols, const void *Yil, int64_t pitch, float br, float bi,
                            void *Y, int64_t ldy, const int32_t *perm) {
    return deinterleave_impl(stream, rows, ncols, Yil, pitch, br, bi, Y, ldy, perm);
}

int ib200_invert_perm(void *stream, int64_t n, const int32_t *perm, int32_t *inv) {
    IB200_REQUIRE(n >= 0 && n < (1LL << 31), "bad length");
    if (n == 0) return 0;
    IB200_REQUIRE(perm && inv, "null pointer");
    invert_perm_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(n, perm, inv);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_ccsrmm_il(void *stream, int64_t m, int64_t k, int64_t ncols, int64_t nnz, float ar, float ai,
                    const void *vals, const int32_t *colind, const int32_t *rowptr, const void *Xil, int64_t xpitch,
                    void *Yil, int64_t ypitch, const int32_t *rowmap, int rows_per_group,
                    const int32_t *longrows, int nlong, int long_thresh) {
    IB200_RANGE("ib200_ccsrmm_il");
    IB200_REQUIRE(m >= 0 && k >= 0 && ncols >= 0 && nnz >= 0, "negative dimension");
    IB200_REQUIRE(m < (1LL << 31) && k < (1LL << 31) && nnz < (1LL << 31), "int32 CSR indices: dimensions must be < 2^31");
    IB200_REQUIRE(ncols <= 32, "interleaved SpMM serves at most 32 columns per call");
    if (m == 0 || ncols == 0) return 0;
    if (nlong <= 0 || !longrows) { nlong = 0; long_thresh = 0x7fffffff; }
    IB200_REQUIRE(rowptr && Yil && (nnz == 0 || (vals && colind && Xil)), "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols, "pitch smaller than the column count");
    IB200_REQUIRE(rows_per_group >= 0 && rows_per_group <= 4096, "rows_per_group out of range");
    const int CL = pow2_ceil(ncols);
    // lanes per row: at least one per coil, more when rows are long enough to feed them
    const double avg = (double)nnz / (double)m;
    int GL = CL;
    while (GL < 32 && avg >= 1.5 * GL) GL <<= 1;
    if (CL == 16 && GL == 32 && avg < 64) GL = 16;
    // rows per group: 0 = automatic (long rows: 32 consecutive rows per CTA; short rows: one pass)
    int rpg = rows_per_group;
    if (rpg == 0) rpg = avg >= 32 ? (32 * GL) / 256 : 1;
    if (rpg < 1) rpg = 1;
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    int rc = launch_long<false>(s, CL, nlong, longrows, (int)ncols, alpha, nullptr, (const c64 *)vals, colind, rowptr,
                                (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap);
    if (rc) return rc;
#define IB200_IL_CASE(gl, cl) \
    case (gl) * 100 + (cl): return launch_il<gl, cl>(s, m, (int)ncols, alpha, (const c64 *)vals, colind, rowptr, (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap, rpg, long_thresh)
    switch (GL * 100 + CL) {
        IB200_IL_CASE(1, 1);
        IB200_IL_CASE(2, 1); IB200_IL_CASE(2, 2);
        IB200_IL_CASE(4, 1); IB200_IL_CASE(4, 2); IB200_IL_CASE(4, 4);
        IB200_IL_CASE(8, 1); IB200_IL_CASE(8, 2); IB200_IL_CASE(8, 4); IB200_IL_CASE(8, 8);
        IB200_IL_CASE(16, 1); IB200_IL_CASE(16, 2); IB200_IL_CASE(16, 4); IB200_IL_CASE(16, 8); IB200_IL_CASE(16, 16);
        IB200_IL_CASE(32, 1); IB200_IL_CASE(32, 2); IB200_IL_CASE(32, 4); IB200_IL_CASE(32, 8); IB200_IL_CASE(32, 16);
        IB200_IL_CASE(32, 32);
    }
#undef IB200_IL_CASE
    set_error("internal: no interleaved kernel for GL=%d CL=%d", GL, CL);
    return IB200_E_UNSUPPORTED;
}

int ib200_grid_tile_rank(void *stream, const int64_t grid[3], const int64_t tile[3], int32_t *colrank, int32_t *rowmap,
                         int64_t *padded_rows) {
    IB200_REQUIRE(grid && tile && padded_rows, "null pointer");
    int64_t padded = 1, plain = 1;
    for (int d = 0; d < 3; ++d) {
        IB200_REQUIRE(grid[d] > 0 && tile[d] > 0 && tile[d] <= 64, "bad grid / tile extent");
        padded *= ceil_div(grid[d], tile[d]) * tile[d];
        plain *= grid[d];
    }
    IB200_REQUIRE(padded < (1LL << 31), "padded grid must hold fewer than 2^31 points");
    *padded_rows = padded;
    if (!colrank && !rowmap) return 0;                           // size query
    IB200_REQUIRE(colrank && rowmap, "null pointer");
    int64_t g = ceil_div(padded, 256); const int64_t cap = (int64_t)sm_count() * 16; if (g > cap) g = cap;
    tile_rank_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>((int)grid[0], (int)grid[1], (int)grid[2], (int)tile[0],
                                                                (int)tile[1], (int)tile[2], colrank, rowmap);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_grid_tile_rank2(void *stream, const int64_t grid[3], const int64_t tile[3], const int64_t super[3],
                          int32_t *colrank, int64_t *nranks) {
    IB200_REQUIRE(grid && tile && super && nranks, "null pointer");
    int64_t padded = 1;
    for (int d = 0; d < 3; ++d) {
        IB200_REQUIRE(grid[d] > 0 && tile[d] > 0 && tile[d] <= 64 && super[d] > 0 && super[d] <= 64, "bad grid / tile extent");
        const int64_t nt = ceil_div(grid[d], tile[d]);
        padded *= ceil_div(nt, super[d]) * super[d] * tile[d];
    }
    IB200_REQUIRE(padded < (1LL << 31), "padded grid must hold fewer than 2^31 points");
    *nranks = padded;
    if (!colrank) return 0;
    const int64_t total = grid[0] * grid[1] * grid[2];
    int64_t g = ceil_div(total, 256); const int64_t cap = (int64_t)sm_count() * 16; if (g > cap) g = cap;
    tile_rank2_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>((int)grid[0], (int)grid[1], (int)grid[2], (int)tile[0],
                                                                 (int)tile[1], (int)tile[2], (int)super[0], (int)super[1],
                                                                 (int)super[2], colrank);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_csr_long_rows(void *stream, int64_t m, const int32_t *rowptr, int thresh, int32_t *list, int capacity,
                        int *host_count) {
    IB200_REQUIRE(m >= 0 && host_count && thresh >= 0 && capacity >= 0, "bad arguments");
    *host_count = 0;
    if (m == 0) return 0;
    IB200_REQUIRE(rowptr && (list || capacity == 0), "null pointer");
    cudaStream_t s = as_stream(stream);
    int *cnt = nullptr;
    IB200_TRY(cudaMalloc(&cnt, sizeof(int)));
    cudaMemsetAsync(cnt, 0, sizeof(int), s);
    long_rows_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(m, rowptr, thresh, list, capacity, cnt);
    count_launch();
    cudaMemcpyAsync(host_count, cnt, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(cnt);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_csr_pack_real(void *stream, int64_t nnz, const void *vals, const int32_t *colind, void *packed,
                        float host_max[2]) {
    IB200_REQUIRE(nnz >= 0 && host_max, "bad arguments");
    host_max[0] = host_max[1] = 0.f;
    if (nnz == 0) return 0;
    IB200_REQUIRE(vals && colind && packed, "null pointer");
    cudaStream_t s = as_stream(stream);
    unsigned *stats = nullptr;
    IB200_TRY(cudaMalloc(&stats, 2 * sizeof(unsigned)));
    cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned), s);
    int64_t g = ceil_div(nnz, 256 * 4); const int64_t cap = (int64_t)sm_count() * 16; if (g > cap) g = cap;
    pack_real_kernel<<<(unsigned)g, 256, 0, s>>>(nnz, (const c64 *)vals, colind, (PackedEntry *)packed, stats);
    count_launch();
    unsigned h[2] = {0, 0};
    cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(stats);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    memcpy(host_max, h, sizeof(h));
    return 0;
}

int ib200_ccsrmm_ilr(void *stream, int64_t m, int64_t k, int64_t ncols, int64_t nnz, float ar, float ai,
                     const void *packed, const int32_t *rowptr, const void *Xil, int64_t xpitch,
                     void *Yil, int64_t ypitch, const int32_t *rowmap, int rows_per_group,
                     const int32_t *longrows, int nlong, int long_thresh) {
    IB200_RANGE("ib200_ccsrmm_ilr");
    IB200_REQUIRE(m >= 0 && k >= 0 && ncols >= 0 && nnz >= 0, "negative dimension");
    IB200_REQUIRE(m < (1LL << 31) && k < (1LL << 31) && nnz < (1LL << 31), "int32 CSR indices: dimensions must be < 2^31");
    IB200_REQUIRE(ncols <= 32, "interleaved SpMM serves at most 32 columns per call");
    if (m == 0 || ncols == 0) return 0;
    if (nlong <= 0 || !longrows) { nlong = 0; long_thresh = 0x7fffffff; }
    IB200_REQUIRE(rowptr && Yil && (nnz == 0 || (packed && Xil)), "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols, "pitch smaller than the column count");
    // rows_per_group < 0 selects the shared-memory staged kernel (-4; -41 forces one coil per lane);
    // the packed array must then be readable two entries past its end
    const int staged = rows_per_group < 0 ? -rows_per_group : 0;
    if (staged) rows_per_group = 0;
    IB200_REQUIRE(rows_per_group >= 0 && rows_per_group <= 4096, "rows_per_group out of range");
    const int CL = pow2_ceil(ncols);
    const double avg = (double)nnz / (double)m;
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    int rc = 0;
    if (nlong > 0) {
        cudaStream_t side = nullptr;
        rc = side_stream_begin(s, &side);
        if (rc) return rc;
        rc = launch_long<true>(side, CL, nlong, longrows, (int)ncols, alpha, packed, nullptr, nullptr, rowptr,
                               (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap);
        if (rc) { side_stream_end(s); return rc; }
        rc = ilr_main(s, staged, CL, avg, rows_per_group, m, ncols, alpha, packed, rowptr, Xil, xpitch, Yil, ypitch, rowmap,
                      long_thresh);
        const int rc2 = side_stream_end(s);
        return rc ? rc : rc2;
    }
    return ilr_main(s, staged, CL, avg, rows_per_group, m, ncols, alpha, packed, rowptr, Xil, xpitch, Yil, ypitch, rowmap,
                    long_thresh);
}

}  // extern "C"

namespace ib200 {
int ilr_main(cudaStream_t s, int staged, int CL, double avg, int rows_per_group, int64_t m, int64_t ncols, c64 alpha,
             const void *packed, const int32_t *rowptr, const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch,
             const int32_t *rowmap, int long_thresh) {
    int GL = CL;
    while (GL < 32 && avg >= 4.0 * GL) GL <<= 1;                 // slots only pay when each gets >= 4 entries
    int rpg = rows_per_group;
    if (rpg == 0) rpg = avg >= 32 ? (32 * GL) / 256 : 1;
    if (rpg < 1) rpg = 1;
    if (staged) {                                                  // shared-memory staged entries, 4 rows per group
        // two coils per lane (16-byte operand loads) whenever the layout allows it
        const bool vec2 = staged != 41 && ncols % 2 == 0 && xpitch % 2 == 0 && ypitch % 2 == 0 &&
                          ((uintptr_t)Xil % 16) == 0 && ((uintptr_t)Yil % 16) == 0;
        const int CLv = vec2 ? pow2_ceil(ncols / 2) : CL;
        int GLv = CLv;
        while (GLv < 32 && avg >= 4.0 * GLv) GLv <<= 1;
#define IB200_ILS_CASE(gl, cl) \
    case (gl) * 100 + (cl): return vec2 \
        ? launch_ils<gl, cl, 4, 4, 2>(s, m, (int)ncols, alpha, (const PackedEntry *)packed, rowptr, (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap, long_thresh) \
        : launch_ils<gl, cl, 4, 4, 1>(s, m, (int)ncols, alpha, (const PackedEntry *)packed, rowptr, (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap, long_thresh)
        switch (GLv * 100 + CLv) {
            IB200_ILS_CASE(1, 1); IB200_ILS_CASE(2, 1); IB200_ILS_CASE(4, 1); IB200_ILS_CASE(8, 1); IB200_ILS_CASE(16, 1); IB200_ILS_CASE(32, 1);
            IB200_ILS_CASE(2, 2); IB200_ILS_CASE(4, 2); IB200_ILS_CASE(4, 4); IB200_ILS_CASE(8, 2); IB200_ILS_CASE(8, 4);
            IB200_ILS_CASE(8, 8); IB200_ILS_CASE(16, 2); IB200_ILS_CASE(16, 4); IB200_ILS_CASE(16, 8); IB200_ILS_CASE(16, 16);
            IB200_ILS_CASE(32, 2); IB200_ILS_CASE(32, 4); IB200_ILS_CASE(32, 8); IB200_ILS_CASE(32, 16); IB200_ILS_CASE(32, 32);
        }
#undef IB200_ILS_CASE
    }
#define IB200_ILR_CASE(gl, cl) \
    case (gl) * 100 + (cl): return launch_ilr<gl, cl>(s, m, (int)ncols, alpha, (const PackedEntry *)packed, rowptr, (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap, rpg, long_thresh)
    switch (GL * 100 + CL) {
        IB200_ILR_CASE(1, 1);
        IB200_ILR_CASE(2, 1); IB200_ILR_CASE(2, 2);
        IB200_ILR_CASE(4, 1); IB200_ILR_CASE(4, 2); IB200_ILR_CASE(4, 4);
        IB200_ILR_CASE(8, 1); IB200_ILR_CASE(8, 2); IB200_ILR_CASE(8, 4); IB200_ILR_CASE(8, 8);
        IB200_ILR_CASE(16, 1); IB200_ILR_CASE(16, 2); IB200_ILR_CASE(16, 4); IB200_ILR_CASE(16, 8); IB200_ILR_CASE(16, 16);
        IB200_ILR_CASE(32, 1); IB200_ILR_CASE(32, 2); IB200_ILR_CASE(32, 4); IB200_ILR_CASE(32, 8); IB200_ILR_CASE(32, 16);
        IB200_ILR_CASE(32, 32);
    }
#undef IB200_ILR_CASE
    set_error("internal: no interleaved kernel for GL=%d CL=%d", GL, CL);
    return IB200_E_UNSUPPORTED;
}

}  // namespace ib200
