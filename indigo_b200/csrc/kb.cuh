// Kaiser-Bessel table lookup shared by the CSR builder (builders.cu) and the separable-weight
// record builder (kbgrid.cu).
#pragma once
#include "common.cuh"

namespace ib200 {

// separable-weight record of one sample (kbgrid.cu builds them; kbgrid.cu and kbtiles.cu read them)
static const int kKbTaps = 6;

struct __align__(16) KbRecord {
    float wx[kKbTaps], wy[kKbTaps], wz[kKbTaps];   // per-axis weights (row weight and scale folded into wz)
    int32_t ix0, iy0, iz0;                         // first tap per axis, wrapped into [0, n)
    int32_t out;                                   // output row (original sample index)
    int32_t ntaps;                                 // nx | ny << 8 | nz << 16
    int32_t pad;
};
static_assert(sizeof(KbRecord) == 96, "KbRecord must be 6 x 16 bytes");

// interp.py:9-15 (lin_interp) with every operation individually rounded
__device__ __forceinline__ double kb_lookup(const double *__restrict__ table, int ntab, double x) {
    if (x >= 1.0) return 0.0;
    const double xs = __dmul_rn(x, (double)(ntab - 1));
    const int i = (int)xs;
    const double frac = __dsub_rn(xs, (double)i);
    return __dadd_rn(__dmul_rn(__dsub_rn(1.0, frac), table[i]), __dmul_rn(frac, table[i + 1]));
}

}  // namespace ib200
