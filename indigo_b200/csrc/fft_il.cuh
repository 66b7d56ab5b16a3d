// Fused SENSE passes on the coil-interleaved oversampled grid
//     grid[z][y][x][c]      (c fastest; x, y, z over the oversampled extents n0, n1, n2)
// shared by the device kernels (fft.cu) and the CPU emulation (tests/csrc/fft_emul.cu).
//
// Reference calls being fused:
//   ccsrmm(P^H, adjoint)  +  fftn          ->  sense_expand_body  + two windowed strided passes
//   ifftn  +  ccsrmm(P^H)                  ->  two windowed strided passes + sense_combine_body
// with P = kron(I_C, mod*zpad*apod) * vstack(maps) (examples/pics.py:111-126): the coil maps,
// apodisation and centring phase are one dense factor pf[voxel][coil], zero-padding is the input
// window of each pass (rows outside the image are never read, written or transformed before the
// pass that creates them) and the crop is the output window.
//
// In this layout every axis is a "strided" axis whose 16 neighbouring lines are the 16 coils
// (or 16 consecutive (x, c) pairs) of one 128-byte segment, so all three passes use the
// lines-fast Stockham stages of fft_core.cuh.
#pragma once
#include "fft_core.cuh"

namespace ib200 {

struct SenseFftArgs {
    const c64 *img;        // expand: image, N0*N1*N2, x fastest
    c64 *img_out;          // combine: image to update
    const c64 *pf;         // [voxel][coil] = q[voxel] * maps[voxel, coil]
    c64 *grid;             // interleaved oversampled grid, n0*n1*n2*C
    const c64 *tw;         // n0 twiddles
    int N0, N1, N2;
    int n0, n1, n2;
    int off0, off1, off2;  // position of the image inside the grid (Zpad 'center', backend.py:371-387)
    int C;
    c64 alpha, beta;
    int beta_zero;
};

// Tile geometry of the x passes: 16 lines = YY neighbouring image rows x CT coils.  With 16 or more
// coils a tile is one row and the coils go in chunks of 16; with 2, 4 or 8 coils per GPU (coil
// sharding) a tile takes 8, 4 or 2 rows so that all 16 lines stay busy.
struct XTile { int CT, YY, shift; };
IB_HD XTile sense_x_tile(int C) {
    XTile t;
    t.CT = C < kSpecL ? C : kSpecL; t.YY = 1; t.shift = 31;
    if (C == 1 || C == 2 || C == 4 || C == 8) {
        t.YY = kSpecL / C;
        t.shift = C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : 3;
    }
    return t;
}
IB_HD int64_t sense_x_blocks(int N1, int N2, int C) {
    const XTile t = sense_x_tile(C);
    return (int64_t)((N1 + t.YY - 1) / t.YY) * N2;
}

// x pass of the forward transform: grid[z+off2][y+off1][:][c] = FFT_x( zpad( img[:, y, z] * pf[., c] ) )
// one CTA per group of YY image rows (y .. y+YY-1, z); coils in chunks of 16 lines when C > 16.
template <int N, int R0, int R1, int R2>
IB_HD void sense_expand_body(const SenseFftArgs &a, c64 *bufA, int64_t block, int tid, int nt) {
    constexpr bool THREE = R2 > 1;
    c64 *bufB = bufA + (size_t)N * kSpecLP;
    const XTile t = sense_x_tile(a.C);
    const int ygroups = (a.N1 + t.YY - 1) / t.YY;
    const int y = (int)(block % ygroups) * t.YY, z = (int)(block / ygroups);
    const int rows = a.N1 - y < t.YY ? a.N1 - y : t.YY;
    const int64_t vox0 = ((int64_t)z * a.N1 + y) * a.N0;
    FftCtx c;
    c.n = N; c.L = kSpecL; c.log2L = 4; c.LP = kSpecLP; c.tw = a.tw;
    c.swap_in = c.swap_out = c.conj_in = c.conj_out = 0;
    c.din = c.dout = nullptr;
    c.in0 = 0; c.in1 = N; c.out0 = 0; c.out1 = N;
    c.gstride_j = a.C; c.gstride_l = 1;
    if (t.YY > 1) { c.lmask = t.CT - 1; c.lshift = t.shift; c.gstride_l2 = (int64_t)a.n0 * a.C; }
    c.gin = nullptr;
    const int64_t row = ((int64_t)(z + a.off2) * a.n1 + (y + a.off1)) * (int64_t)a.n0 * a.C;
    for (int c0 = 0; c0 < a.C; c0 += kSpecL) {
        c.nl = t.YY > 1 ? rows * t.CT : (a.C - c0 < kSpecL ? a.C - c0 : kSpecL);
        c.gout = a.grid + row + c0;
        for (int idx = tid; idx < N * kSpecL; idx += nt) {
            const int l = idx & (kSpecL - 1), pos = idx >> 4;
            if (l >= c.nl) continue;
            const int j = pos - a.off0;
            c64 v = h_mk(0.f, 0.f);
            if (j >= 0 && j < a.N0) {
                const int64_t vox = vox0 + (int64_t)(l >> c.lshift) * a.N0 + j;
                v = h_mul(a.img[vox], a.pf[vox * a.C + c0 + (l & c.lmask)]);
            }
            bufA[spec_addr<0>(pos, l)] = v;
        }
        IB_SYNC();
        spec_stage_lfast<N, R0, 1, false, false, 0, 0>(c, bufA, bufB, tid, nt);
        IB_SYNC();
        if (THREE) {
            spec_stage_lfast<N, R1, R0, false, false, 0, 0>(c, bufB, bufA, tid, nt);
            IB_SYNC();
            spec_stage_lfast<N, THREE ? R2 : R1, R0 * R1, false, true, 0, 0>(c, bufA, nullptr, tid, nt);
        } else {
            spec_stage_lfast<N, R1, R0, false, true, 0, 0>(c, bufB, nullptr, tid, nt);
        }
        IB_SYNC();
    }
}

// x pass of the inverse transform with the coil combination:
//   img_out[:, y, z] = alpha * sum_c conj(pf[., c]) * crop( IFFT_x( grid[z+off2][y+off1][:][c] ) ) + beta * img_out
// The re/im swap that turns the forward butterflies into the inverse transform was applied on the
// load of the first inverse pass (z); this is the last pass, so it swaps back before the product.
// `acc` is YY*N0 complex words of shared memory.
template <int N, int R0, int R1, int R2>
IB_HD void sense_combine_body(const SenseFftArgs &a, c64 *bufA, c64 *acc, int64_t block, int tid, int nt) {
    constexpr bool THREE = R2 > 1;
    c64 *bufB = bufA + (size_t)N * kSpecLP;
    const XTile t = sense_x_tile(a.C);
    const int ygroups = (a.N1 + t.YY - 1) / t.YY;
    const int y = (int)(block % ygroups) * t.YY, z = (int)(block / ygroups);
    const int rows = a.N1 - y < t.YY ? a.N1 - y : t.YY;
    const int64_t vox0 = ((int64_t)z * a.N1 + y) * a.N0;
    FftCtx c;
    c.n = N; c.L = kSpecL; c.log2L = 4; c.LP = kSpecLP; c.tw = a.tw;
    c.swap_in = c.swap_out = c.conj_in = c.conj_out = 0;
    c.din = c.dout = nullptr;
    c.in0 = 0; c.in1 = N; c.out0 = 0; c.out1 = N;
    c.gstride_j = a.C; c.gstride_l = 1;
    if (t.YY > 1) { c.lmask = t.CT - 1; c.lshift = t.shift; c.gstride_l2 = (int64_t)a.n0 * a.C; }
    c.gout = nullptr;
    const int64_t row = ((int64_t)(z + a.off2) * a.n1 + (y + a.off1)) * (int64_t)a.n0 * a.C;
    for (int c0 = 0; c0 < a.C; c0 += kSpecL) {
        c.nl = t.YY > 1 ? rows * t.CT : (a.C - c0 < kSpecL ? a.C - c0 : kSpecL);
        const int ncl = t.YY > 1 ? t.CT : c.nl;                   // coils of one row in this tile
        c.gin = a.grid + row + c0;
        spec_stage_lfast<N, R0, 1, true, false, 0, 0>(c, nullptr, bufA, tid, nt);
        IB_SYNC();
        spec_stage_lfast<N, R1, R0, false, false, 0, 0>(c, bufA, bufB, tid, nt);
        IB_SYNC();
        c64 *res = bufB;
        if (THREE) {
            spec_stage_lfast<N, THREE ? R2 : R1, R0 * R1, false, false, 0, 0>(c, bufB, bufA, tid, nt);
            IB_SYNC();
            res = bufA;
        }
        // multiply the cropped positions by conj(pf) in place (coil index fastest: coalesced pf reads) ...
        for (int idx = tid; idx < a.N0 * kSpecL; idx += nt) {
            const int l = idx & (kSpecL - 1), j = idx >> 4;
            if (l >= c.nl) continue;
            const int at = spec_addr<0>(j + a.off0, l);
            const int64_t vox = vox0 + (int64_t)(l >> c.lshift) * a.N0 + j;
            res[at] = h_mulc(h_swap(res[at]), a.pf[vox * a.C + c0 + (l & c.lmask)]);
        }
        IB_SYNC();
        // ... and fold the coils of each (row, position)
        for (int i = tid; i < rows * a.N0; i += nt) {
            const int yy = i / a.N0, j = i - yy * a.N0;
            c64 s = c0 == 0 ? h_mk(0.f, 0.f) : acc[i];
            for (int l = 0; l < ncl; ++l) s = h_add(s, res[spec_addr<0>(j + a.off0, yy * ncl + l)]);
            acc[i] = s;
        }
        IB_SYNC();
    }
    for (int i = tid; i < rows * a.N0; i += nt) {
        c64 v = h_mul(a.alpha, acc[i]);
        if (!a.beta_zero) v = h_add(v, h_mul(a.beta, a.img_out[vox0 + i]));
        a.img_out[vox0 + i] = v;
    }
}

// Windowed strided pass over 16-line tiles of the interleaved grid (y and z passes of both
// directions).  Same Stockham stages as fft_pass_body_spec<..., AXIS0 = false>, specialised through
// IlCtx: full tiles only (inner % 16 == 0), position stride < 2^32 elements.
struct IlPassArgs {
    c64 *x;                            // first element of slab 0, transformed in place
    const c64 *tw;
    int64_t inner, outer, outer_stride;
    unsigned pstride;
    int in0, in1, out0, out1;
    // optional per-tile windows (k-space support of the trajectory, fft_pk.cuh): win[2*p], win[2*p+1] =
    // [lo, hi) along the transformed axis for the grid point p = (first line of the tile) / win_div.
    // win_mode 1: output window (forward pass; tiles with an empty window are skipped altogether);
    // win_mode 2: input window (inverse pass; tiles with an empty window store zeros).  Persistent
    // packed kernels only.
    const int32_t *win = nullptr;
    int win_mode = 0, win_div = 1;
};

template <int N, int R0, int R1, int R2, bool SWAP_IN, bool SWAP_OUT>
IB_HD void fft_il_pass_body(const IlPassArgs &a, c64 *bufA, int64_t block, int tid, int nt) {
    constexpr bool THREE = R2 > 1;
    c64 *bufB = bufA + (size_t)N * kSpecLP;
    const int64_t tiles = a.inner / kSpecL;
    const int64_t o = block / tiles, s0 = (block % tiles) * kSpecL;
    IlCtx<SWAP_IN, SWAP_OUT> c;
    c.gin = a.x + o * a.outer_stride + s0; c.gout = a.x + o * a.outer_stride + s0;
    c.tw = a.tw; c.pstride = a.pstride;
    c.in0 = a.in0; c.inlen = (unsigned)(a.in1 - a.in0); c.out0 = a.out0; c.outlen = (unsigned)(a.out1 - a.out0);
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

}  // namespace ib200
