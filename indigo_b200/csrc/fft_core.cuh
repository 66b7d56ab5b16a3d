// Stockham auto-sort FFT building blocks shared by the device kernels (fft.cu)
// and by the host-side emulation harness (tests/csrc/fft_emul.cu), which runs
// the very same index arithmetic and butterflies thread by thread on the CPU so
// that the kernel logic can be checked in the GPU-less build container.
//
// One "item" = one radix-R butterfly of one line of the tile.  For stage s with
// radix R, p = product of the radices of earlier stages, t = n / R:
//     k      = b mod p
//     u[q]   = in[b + q*t] * w_n^(q*k*n/(p*R))          q = 0..R-1
//     u      = DFT_R(u)                                  (forward sign)
//     out[(b-k)*R + k + q*p] = u[q]
// Tile layout in shared memory is [position][line] with the line pitch padded
// to L+1 complex words: a half-warp of 16 consecutive lines touches 16
// consecutive 8-byte words, so every stage is bank-conflict free for any radix.
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
#define IB_HD __host__ __device__ __forceinline__
#else
#define IB_HD inline
#endif

namespace ib200 {

IB_HD c64 h_mk(float a, float b) { c64 r; r.x = a; r.y = b; return r; }
IB_HD c64 h_add(c64 a, c64 b) { return h_mk(a.x + b.x, a.y + b.y); }
IB_HD c64 h_sub(c64 a, c64 b) { return h_mk(a.x - b.x, a.y - b.y); }
IB_HD c64 h_mul(c64 a, c64 b) { return h_mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
IB_HD c64 h_mulc(c64 a, c64 b) { return h_mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }   // a*conj(b)
IB_HD c64 h_mi(c64 a) { return h_mk(a.y, -a.x); }     // * (-i)
IB_HD c64 h_pi(c64 a) { return h_mk(-a.y, a.x); }     // * (+i)
IB_HD c64 h_swap(c64 a) { return h_mk(a.y, a.x); }

template <int R> struct Trig;
template <> struct Trig<3> {
    IB_HD static float c(int j) { const float t[] = {-5.000000000e-01f}; return t[j]; }
    IB_HD static float s(int j) { const float t[] = {8.660254038e-01f}; return t[j]; }
};
template <> struct Trig<5> {
    IB_HD static float c(int j) { const float t[] = {3.090169944e-01f, -8.090169944e-01f}; return t[j]; }
    IB_HD static float s(int j) { const float t[] = {9.510565163e-01f, 5.877852523e-01f}; return t[j]; }
};
template <> struct Trig<7> {
    IB_HD static float c(int j) { const float t[] = {6.234898019e-01f, -2.225209340e-01f, -9.009688679e-01f}; return t[j]; }
    IB_HD static float s(int j) { const float t[] = {7.818314825e-01f, 9.749279122e-01f, 4.338837391e-01f}; return t[j]; }
};
template <> struct Trig<11> {
    IB_HD static float c(int j) { const float t[] = {8.412535328e-01f, 4.154150130e-01f, -1.423148383e-01f, -6.548607339e-01f, -9.594929736e-01f}; return t[j]; }
    IB_HD static float s(int j) { const float t[] = {5.406408175e-01f, 9.096319954e-01f, 9.898214419e-01f, 7.557495744e-01f, 2.817325568e-01f}; return t[j]; }
};
template <> struct Trig<13> {
    IB_HD static float c(int j) { const float t[] = {8.854560257e-01f, 5.680647467e-01f, 1.205366803e-01f, -3.546048870e-01f, -7.485107482e-01f, -9.709418174e-01f}; return t[j]; }
    IB_HD static float s(int j) { const float t[] = {4.647231720e-01f, 8.229838659e-01f, 9.927088741e-01f, 9.350162427e-01f, 6.631226582e-01f, 2.393156643e-01f}; return t[j]; }
};

// ---- forward DFTs of a register-resident vector, natural order in and out ---
template <int R> struct Dft {
    // odd prime R: pair x_j with x_{R-j}; (R-1)/2 cosine sums and sine sums
    IB_HD static void run(c64 (&u)[R]) {
        constexpr int H = (R - 1) / 2;
        c64 a[H], b[H];
#pragma unroll
        for (int j = 0; j < H; ++j) { a[j] = h_add(u[j + 1], u[R - 1 - j]); b[j] = h_sub(u[j + 1], u[R - 1 - j]); }
        const c64 x0 = u[0];
        c64 s0 = x0;
#pragma unroll
        for (int j = 0; j < H; ++j) s0 = h_add(s0, a[j]);
        u[0] = s0;
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            c64 ck = x0, sk = h_mk(0.f, 0.f);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const int m = (j * k) % R;
                const float cv = m <= H ? Trig<R>::c(m - 1) : Trig<R>::c(R - m - 1);
                const float sv = m <= H ? Trig<R>::s(m - 1) : -Trig<R>::s(R - m - 1);
                ck.x += cv * a[j - 1].x; ck.y += cv * a[j - 1].y;
                sk.x += sv * b[j - 1].x; sk.y += sv * b[j - 1].y;
            }
            u[k] = h_mk(ck.x + sk.y, ck.y - sk.x);          // ck - i*sk
            u[R - k] = h_mk(ck.x - sk.y, ck.y + sk.x);      // ck + i*sk
        }
    }
};

template <> struct Dft<2> {
    IB_HD static void run(c64 (&u)[2]) { const c64 a = u[0], b = u[1]; u[0] = h_add(a, b); u[1] = h_sub(a, b); }
};

template <> struct Dft<4> {
    IB_HD static void run(c64 (&u)[4]) {
        const c64 t0 = h_add(u[0], u[2]), t1 = h_sub(u[0], u[2]);
        const c64 t2 = h_add(u[1], u[3]), t3 = h_mi(h_sub(u[1], u[3]));
        u[0] = h_add(t0, t2); u[1] = h_add(t1, t3); u[2] = h_sub(t0, t2); u[3] = h_sub(t1, t3);
    }
};

template <> struct Dft<8> {
    IB_HD static void run(c64 (&u)[8]) {
        c64 e[4] = {u[0], u[2], u[4], u[6]}, o[4] = {u[1], u[3], u[5], u[7]};
        Dft<4>::run(e); Dft<4>::run(o);
        const float h = 0.70710678118654752f;
        o[1] = h_mk(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * (1-i)/sqrt2
        o[2] = h_mi(o[2]);                                               // * -i
        o[3] = h_mk(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * (-1-i)/sqrt2
#pragma unroll
        for (int k = 0; k < 4; ++k) { u[k] = h_add(e[k], o[k]); u[k + 4] = h_sub(e[k], o[k]); }
    }
};

template <> struct Dft<16> {
    IB_HD static void run(c64 (&u)[16]) {
        // 16 = 4 x 4: X[k1 + 4*k2] = sum_n2 w16^(n2*k1) * (sum_n1 x[4*n1+n2] w4^(n1*k1)) * w4^(n2*k2)
        c64 y[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            c64 t[4] = {u[n2], u[4 + n2], u[8 + n2], u[12 + n2]};
            Dft<4>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) y[n2][k1] = t[k1];
        }
        const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
        // w16^m = cos(2 pi m/16) - i sin(2 pi m/16)
        y[1][1] = h_mul(y[1][1], h_mk(c1, -s1));
        y[1][2] = h_mul(y[1][2], h_mk(h, -h));
        y[1][3] = h_mul(y[1][3], h_mk(s1, -c1));
        y[2][1] = h_mul(y[2][1], h_mk(h, -h));
        y[2][2] = h_mi(y[2][2]);
        y[2][3] = h_mul(y[2][3], h_mk(-h, -h));
        y[3][1] = h_mul(y[3][1], h_mk(s1, -c1));
        y[3][2] = h_mul(y[3][2], h_mk(-h, -h));
        y[3][3] = h_mul(y[3][3], h_mk(-c1, s1));       // w16^9 = cos(9pi/8) - i sin(9pi/8)
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            c64 t[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
            Dft<4>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) u[k1 + 4 * k2] = t[k2];
        }
    }
};

// ---- tile context -----------------------------------------------------------
struct FftCtx {
    const c64 *gin;        // global input, already offset to this tile's first element
    c64 *gout;             // global output, same offset
    int64_t gstride_j;     // elements between consecutive positions of a line
    int64_t gstride_l;     // elements between consecutive lines of the tile
    const c64 *tw;         // n twiddles exp(-2 pi i j / n)
    const c64 *din;        // optional diagonal on load (offset like gin, or null)
    const c64 *dout;       // optional diagonal on store
    int n, L, log2L, LP, nl;
    int swap_in, swap_out, conj_in, conj_out;
    int in0, in1;          // positions outside [in0, in1) read as zero without touching memory (pruned input)
    int out0, out1;        // positions outside [out0, out1) are not stored (pruned output)
    // two-level line addressing: line l sits at (l & lmask)*gstride_l + (l >> lshift)*gstride_l2
    // (tiles made of a few coils of several neighbouring rows, fft_il.cuh); default: one level
    int lmask = 0x7fffffff, lshift = 31;
    int64_t gstride_l2 = 0;
};

IB_HD c64 fft_gload(const FftCtx &c, int l, int j) {
    if (j < c.in0 || j >= c.in1) return h_mk(0.f, 0.f);
    const int64_t off = (int64_t)(l & c.lmask) * c.gstride_l + (int64_t)(l >> c.lshift) * c.gstride_l2 + (int64_t)j * c.gstride_j;
    c64 v = c.gin[off];
    if (c.din) { const c64 d = c.din[off]; v = c.conj_in ? h_mulc(v, d) : h_mul(v, d); }
    return c.swap_in ? h_swap(v) : v;
}

IB_HD void fft_gstore(const FftCtx &c, int l, int j, c64 v) {
    if (j < c.out0 || j >= c.out1) return;
    const int64_t off = (int64_t)(l & c.lmask) * c.gstride_l + (int64_t)(l >> c.lshift) * c.gstride_l2 + (int64_t)j * c.gstride_j;
    if (c.swap_out) v = h_swap(v);
    if (c.dout) { const c64 d = c.dout[off]; v = c.conj_out ? h_mulc(v, d) : h_mul(v, d); }
    c.gout[off] = v;
}

// One butterfly.  idx in [0, L*(n/R)); lines are the fast index.
template <int R, bool SRC_G, bool DST_G>
IB_HD void fft_stage_item(const FftCtx &c, const c64 *sin_, c64 *sout, int p, int idx) {
    const int t = c.n / R;
    const int l = idx & (c.L - 1);
    const int b = idx >> c.log2L;
    if (l >= c.nl) return;
    const int k = p == 1 ? 0 : b % p;
    c64 u[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int pos = b + s * t;
        u[s] = SRC_G ? fft_gload(c, l, pos) : sin_[(size_t)pos * c.LP + l];
    }
    if (p > 1) {
        const int tstep = k * (c.n / (p * R));
#pragma unroll
        for (int s = 1; s < R; ++s) u[s] = h_mul(u[s], c.tw[s * tstep]);
    }
    Dft<R>::run(u);
    const int o0 = (b - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        const int pos = o0 + q * p;
        if (DST_G) fft_gstore(c, l, pos, u[q]);
        else sout[(size_t)pos * c.LP + l] = u[q];
    }
}

// Generic (runtime) radix: one OUTPUT element per item, idx in [0, L*n); source
// always shared memory.  O(r) work per output, meant for the odd prime factors
// > 13 that MRI grids occasionally have (17, 19, 23 ...).
template <bool DST_G>
IB_HD void fft_stage_item_generic(const FftCtx &c, const c64 *sin_, c64 *sout, int r, int p, int idx) {
    const int t = c.n / r;
    const int l = idx & (c.L - 1);
    const int e = idx >> c.log2L;
    if (l >= c.nl) return;
    const int b = e % t, q = e / t;
    const int k = b % p;
    const int step = k * (c.n / (p * r)) + q * t;      // < n
    c64 acc = h_mk(0.f, 0.f);
    int ex = 0;
    for (int s = 0; s < r; ++s) {
        const c64 v = sin_[(size_t)(b + s * t) * c.LP + l];
        const c64 w = c.tw[ex];
        acc.x += v.x * w.x - v.y * w.y;
        acc.y += v.x * w.y + v.y * w.x;
        ex += step;
        if (ex >= c.n) ex -= c.n;
    }
    const int pos = (b - k) * r + k + q * p;
    if (DST_G) fft_gstore(c, l, pos, acc);
    else sout[(size_t)pos * c.LP + l] = acc;
}

// tile <-> shared copies.  j_fast selects the thread order that is coalesced in
// global memory: positions fastest for contiguous lines (axis 0), lines fastest
// for strided axes.
IB_HD void fft_load_item(const FftCtx &c, c64 *buf, bool j_fast, int idx) {
    int l, j;
    if (j_fast) { j = idx % c.n; l = idx / c.n; } else { l = idx & (c.L - 1); j = idx >> c.log2L; }
    if (l >= c.nl) return;
    buf[(size_t)j * c.LP + l] = fft_gload(c, l, j);
}

IB_HD void fft_store_item(const FftCtx &c, const c64 *buf, bool j_fast, int idx) {
    int l, j;
    if (j_fast) { j = idx % c.n; l = idx / c.n; } else { l = idx & (c.L - 1); j = idx >> c.log2L; }
    if (l >= c.nl) return;
    fft_gstore(c, l, j, buf[(size_t)j * c.LP + l]);
}

// ---- plan description shared by host and device ------------------------------
static const int kMaxStages = 12;
struct FftStages {
    int nst;
    int radix[kMaxStages];       // specialised: 2,3,4,5,7,8,11,13,16; anything else = generic prime
};

IB_HD bool fft_radix_is_special(int r) {
    return r == 2 || r == 3 || r == 4 || r == 5 || r == 7 || r == 8 || r == 11 || r == 13 || r == 16;
}


// ---- one pass over one tile: shared by the device kernel and the host emulation
struct FftKernelArgs {
    const c64 *x; c64 *y;
    const c64 *tw;
    const c64 *din; const c64 *dout;
    int64_t inner, outer;           // element stride of the axis / number of outer slabs (lines for axis 0)
    int64_t plane;                  // prod(dims): period of the diagonals
    int n, L, log2L;
    int swap_in, swap_out, conj_in, conj_out;
    int load_first, store_last;
    int in0, in1, out0, out1;       // input / output windows along the transformed axis (see FftCtx)
    int64_t outer_stride;           // elements between consecutive outer slabs (n*inner when dense)
    FftStages st;
};

#ifdef __CUDA_ARCH__
#define IB_SYNC() __syncthreads()
#else
#define IB_SYNC() ((void)0)
#endif

template <int R>
IB_HD void fft_run_stage_special(const FftCtx &c, const c64 *src, c64 *dst, int p, bool src_g, bool dst_g, int tid, int nt) {
    const int items = c.L * (c.n / R);
    if (src_g && dst_g)      { for (int i = tid; i < items; i += nt) fft_stage_item<R, true, true>(c, src, dst, p, i); }
    else if (src_g)          { for (int i = tid; i < items; i += nt) fft_stage_item<R, true, false>(c, src, dst, p, i); }
    else if (dst_g)          { for (int i = tid; i < items; i += nt) fft_stage_item<R, false, true>(c, src, dst, p, i); }
    else                     { for (int i = tid; i < items; i += nt) fft_stage_item<R, false, false>(c, src, dst, p, i); }
}

IB_HD void fft_run_stage(const FftCtx &c, int r, const c64 *src, c64 *dst, int p, bool src_g, bool dst_g, int tid, int nt) {
    switch (r) {
        case 2:  fft_run_stage_special<2>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 3:  fft_run_stage_special<3>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 4:  fft_run_stage_special<4>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 5:  fft_run_stage_special<5>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 7:  fft_run_stage_special<7>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 8:  fft_run_stage_special<8>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 11: fft_run_stage_special<11>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 13: fft_run_stage_special<13>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        case 16: fft_run_stage_special<16>(c, src, dst, p, src_g, dst_g, tid, nt); break;
        default: {
            const int items = c.L * c.n;
            if (dst_g) { for (int i = tid; i < items; i += nt) fft_stage_item_generic<true>(c, src, dst, r, p, i); }
            else       { for (int i = tid; i < items; i += nt) fft_stage_item_generic<false>(c, src, dst, r, p, i); }
        }
    }
}

// AXIS0: lines are contiguous (inner == 1), tile = L consecutive lines.
// (tid, nt) = (threadIdx.x, blockDim.x) on the device, (0, 1) in the emulation.
template <bool AXIS0>
IB_HD void fft_pass_body(const FftKernelArgs &a, c64 *bufA, int64_t block, int tid, int nt) {
    FftCtx c;
    c.n = a.n; c.L = a.L; c.log2L = a.log2L; c.LP = a.L + 1;
    c.tw = a.tw;
    c.swap_in = a.swap_in; c.swap_out = a.swap_out; c.conj_in = a.conj_in; c.conj_out = a.conj_out;
    c.in0 = a.in0; c.in1 = a.in1; c.out0 = a.out0; c.out1 = a.out1;
    c64 *bufB = bufA + (size_t)a.n * c.LP;

    int64_t base;
    if (AXIS0) {
        const int64_t line0 = block * a.L;
        const int64_t left = a.outer - line0;
        c.nl = left < a.L ? (int)left : a.L;
        base = line0 * a.n;
        c.gstride_j = 1; c.gstride_l = a.n;
    } else {
        const int64_t tiles = (a.inner + a.L - 1) / a.L;
        const int64_t o = block / tiles, ts = block % tiles;
        const int64_t s0 = ts * a.L;
        const int64_t left = a.inner - s0;
        c.nl = left < a.L ? (int)left : a.L;
        base = o * a.outer_stride + s0;
        c.gstride_j = a.inner; c.gstride_l = 1;
    }
    c.gin = a.x + base; c.gout = a.y + base;
    // Diagonals repeat with period `plane`.  A tile never straddles a batch item
    // on strided axes; on axis 0 the host sizes L to divide the lines of one
    // item whenever a diagonal is present, so the offset is tile-uniform.
    const int64_t dbase = a.plane > 0 ? base % a.plane : 0;
    c.din = a.din ? a.din + dbase : nullptr;
    c.dout = a.dout ? a.dout + dbase : nullptr;

    const int nst = a.st.nst;
    const c64 *src = nullptr;
    c64 *dst = bufA;
    bool src_g = true;
    const int copy_items = AXIS0 ? c.nl * c.n : c.L * c.n;
    if (a.load_first) {
        for (int i = tid; i < copy_items; i += nt) fft_load_item(c, bufA, AXIS0, i);
        IB_SYNC();
        src = bufA; dst = bufB; src_g = false;
    }
    int p = 1;
    for (int s = 0; s < nst; ++s) {
        const int r = a.st.radix[s];
        const bool dst_g = (s == nst - 1) && !a.store_last;
        fft_run_stage(c, r, src, dst, p, src_g, dst_g, tid, nt);
        p *= r;
        if (!dst_g) {
            IB_SYNC();
            src = dst; dst = (dst == bufA) ? bufB : bufA; src_g = false;
        }
    }
    if (a.store_last)
        for (int i = tid; i < copy_items; i += nt) fft_store_item(c, src, AXIS0, i);
}


// ============================================================================
// Compile-time specialised passes for the grid sizes MRI reconstructions use.
// Same Stockham recurrence as above, but n, the radix sequence, p and t are
// template constants (no integer division or runtime radix dispatch), twiddle
// powers come from one table load per butterfly by log-depth multiplication,
// tiles are always L = 16 lines, and on axis 0 the first and last stage talk to
// global memory directly with positions as the fast thread index (coalesced)
// instead of transposing through an extra shared-memory round trip.
// ============================================================================
static const int kSpecL = 16, kSpecLP = 17;

// Lean tile context of the fused interleaved passes (fft_il.cuh): always 16 full lines of stride 1,
// no diagonals, re/im swaps fixed at compile time, windows as one unsigned compare, 32-bit position
// stride (one IMAD.WIDE per access instead of the generic 64-bit index arithmetic of FftCtx).
template <bool SWAP_IN, bool SWAP_OUT>
struct IlCtx {
    const c64 *gin;
    c64 *gout;
    const c64 *tw;
    unsigned pstride;                  // elements between consecutive positions of a line
    int in0; unsigned inlen;           // input window  [in0, in0 + inlen)
    int out0; unsigned outlen;         // output window [out0, out0 + outlen)
    static constexpr int nl = kSpecL;
};

template <bool SI, bool SO>
IB_HD c64 fft_gload(const IlCtx<SI, SO> &c, int l, int j) {
    if ((unsigned)(j - c.in0) >= c.inlen) return h_mk(0.f, 0.f);
    const c64 v = c.gin[(uint64_t)(unsigned)j * c.pstride + (unsigned)l];
    return SI ? h_swap(v) : v;
}

template <bool SI, bool SO>
IB_HD void fft_gstore(const IlCtx<SI, SO> &c, int l, int j, c64 v) {
    if ((unsigned)(j - c.out0) >= c.outlen) return;
    c.gout[(uint64_t)(unsigned)j * c.pstride + (unsigned)l] = SO ? h_swap(v) : v;
}

// powers w^1 .. w^(R-1) with multiplication depth log2(R)
template <int R>
IB_HD void twiddle_powers(c64 w, c64 (&pw)[R]) {
    pw[0] = h_mk(1.f, 0.f);
    if (R > 1) pw[1] = w;
#pragma unroll
    for (int s = 2; s < R; ++s) pw[s] = h_mul(pw[s / 2], pw[s - s / 2]);
}

// shared-memory address of (position, line); ROT_R > 0 rotates the line index by
// pos / ROT_R so that a position-fast writer with an even radix stays conflict free
template <int ROT_R>
IB_HD int spec_addr(int pos, int l) {
    if (ROT_R > 0) return pos * kSpecLP + ((l + pos / ROT_R) & (kSpecL - 1));
    return pos * kSpecLP + l;
}

// lines-fast stage (strided axes: every stage; axis 0: the middle stage)
template <int N, int R, int P, bool SRC_G, bool DST_G, int ROT_IN, int ROT_OUT, class CTX>
IB_HD void spec_stage_lfast(const CTX &c, const c64 *sin_, c64 *sout, int tid, int nt) {
    constexpr int T = N / R, ITEMS = kSpecL * T, STEP = N / (P * R);
    for (int idx = tid; idx < ITEMS; idx += nt) {
        const int l = idx & (kSpecL - 1), b = idx >> 4;
        if (l >= c.nl) continue;
        const int k = P == 1 ? 0 : b % P;
        c64 u[R];
#pragma unroll
        for (int s = 0; s < R; ++s) {
            const int pos = b + s * T;
            u[s] = SRC_G ? fft_gload(c, l, pos) : sin_[spec_addr<ROT_IN>(pos, l)];
        }
        if (P > 1) {
            c64 pw[R];
            twiddle_powers<R>(c.tw[k * STEP], pw);
#pragma unroll
            for (int s = 1; s < R; ++s) u[s] = h_mul(u[s], pw[s]);
        }
        Dft<R>::run(u);
        const int o0 = (b - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int pos = o0 + q * P;
            if (DST_G) fft_gstore(c, l, pos, u[q]);
            else sout[spec_addr<ROT_OUT>(pos, l)] = u[q];
        }
    }
}

// axis 0, first stage (P = 1): global -> shared, positions fast
template <int N, int R, int ROT_OUT>
IB_HD void spec_stage_first_bfast(const FftCtx &c, c64 *sout, int tid, int nt) {
    constexpr int T = N / R, ITEMS = kSpecL * T;
    for (int idx = tid; idx < ITEMS; idx += nt) {
        const int b = idx % T, l = idx / T;
        if (l >= c.nl) continue;
        c64 u[R];
#pragma unroll
        for (int s = 0; s < R; ++s) u[s] = fft_gload(c, l, b + s * T);
        Dft<R>::run(u);
#pragma unroll
        for (int q = 0; q < R; ++q) sout[spec_addr<ROT_OUT>(b * R + q, l)] = u[q];
    }
}

// axis 0, last stage (P = N / R): shared -> global, positions fast
template <int N, int R, int ROT_IN>
IB_HD void spec_stage_last_bfast(const FftCtx &c, const c64 *sin_, int tid, int nt) {
    constexpr int T = N / R, ITEMS = kSpecL * T;
    for (int idx = tid; idx < ITEMS; idx += nt) {
        const int b = idx % T, l = idx / T;
        if (l >= c.nl) continue;
        c64 u[R], pw[R];
#pragma unroll
        for (int s = 0; s < R; ++s) u[s] = sin_[spec_addr<ROT_IN>(b + s * T, l)];
        twiddle_powers<R>(c.tw[b], pw);                 // k = b, step = 1
#pragma unroll
        for (int s = 1; s < R; ++s) u[s] = h_mul(u[s], pw[s]);
        Dft<R>::run(u);
#pragma unroll
        for (int q = 0; q < R; ++q) fft_gstore(c, l, b + q * T, u[q]);
    }
}

template <int N, int R0, int R1, int R2, bool AXIS0>
IB_HD void fft_pass_body_spec(const FftKernelArgs &a, c64 *bufA, int64_t block, int tid, int nt) {
    constexpr bool THREE = R2 > 1;
    FftCtx c;
    c.n = N; c.L = kSpecL; c.log2L = 4; c.LP = kSpecLP;
    c.tw = a.tw;
    c.swap_in = a.swap_in; c.swap_out = a.swap_out; c.conj_in = a.conj_in; c.conj_out = a.conj_out;
    c.in0 = a.in0; c.in1 = a.in1; c.out0 = a.out0; c.out1 = a.out1;
    c64 *bufB = bufA + (size_t)N * kSpecLP;
    int64_t base;
    if (AXIS0) {
        const int64_t line0 = block * kSpecL, left = a.outer - line0;
        c.nl = left < kSpecL ? (int)left : kSpecL;
        base = line0 * N;
        c.gstride_j = 1; c.gstride_l = N;
    } else {
        const int64_t tiles = (a.inner + kSpecL - 1) / kSpecL;
        const int64_t o = block / tiles, s0 = (block % tiles) * kSpecL, left = a.inner - s0;
        c.nl = left < kSpecL ? (int)left : kSpecL;
        base = o * a.outer_stride + s0;
        c.gstride_j = a.inner; c.gstride_l = 1;
    }
    c.gin = a.x + base; c.gout = a.y + base;
    const int64_t dbase = a.plane > 0 ? base % a.plane : 0;
    c.din = a.din ? a.din + dbase : nullptr;
    c.dout = a.dout ? a.dout + dbase : nullptr;

    if (AXIS0) {
        constexpr int ROT = (R0 % 2 == 0) ? R0 : 0;
        spec_stage_first_bfast<N, R0, ROT>(c, bufA, tid, nt);
        IB_SYNC();
        if (THREE) {
            spec_stage_lfast<N, R1, R0, false, false, ROT, 0>(c, bufA, bufB, tid, nt);
            IB_SYNC();
            spec_stage_last_bfast<N, THREE ? R2 : R1, 0>(c, bufB, tid, nt);
        } else {
            spec_stage_last_bfast<N, R1, ROT>(c, bufA, tid, nt);
        }
    } else {
        spec_stage_lfast<N, R0, 1, true, false, 0, 0>(c, nullptr, bufA, tid, nt);
        IB_SYNC();
        if (THREE) {
            spec_stage_lfast<N, R1, R0, false, false, 0, 0>(c, bufA, bufB, tid, nt);
            IB_SYNC();
            spec_stage_lfast<N, THREE ? R2 : R1, R0 * R1, false, true, 0, 0>(c, bufB, nullptr, tid, nt);
        } else {
            spec_stage_lfast<N, R1, R0, false, true, 0, 0>(c, bufA, nullptr, tid, nt);
        }
    }
}

// sizes with a specialised pass: X(n, r0, r1, r2)   (r2 == 1: two stages).
// The radix order must be the planner's (fft_factorize).
#define IB200_FFT_SPEC_LIST(X) \
    X(32, 8, 4, 1) X(52, 13, 4, 1) X(64, 8, 8, 1) X(104, 13, 8, 1) X(128, 16, 8, 1) X(192, 3, 8, 8) \
    X(208, 13, 16, 1) X(256, 16, 16, 1) X(320, 5, 8, 8) X(384, 3, 16, 8) X(416, 13, 8, 4) X(448, 7, 8, 8) \
    X(512, 8, 8, 8) X(640, 5, 16, 8) X(768, 3, 16, 16) X(832, 13, 8, 8) X(1024, 16, 8, 8)

IB_HD bool fft_spec_matches(const FftKernelArgs &a, int n, int r0, int r1, int r2) {
    const int nst = r2 > 1 ? 3 : 2;
    if (a.n != n || a.st.nst != nst) return false;
    if (a.st.radix[0] != r0 || a.st.radix[1] != r1) return false;
    return nst == 2 || a.st.radix[2] == r2;
}

}  // namespace ib200
