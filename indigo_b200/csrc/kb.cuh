// Kaiser-Bessel table lookup shared by the CSR builder (builders.cu) and the separable-weight
// record builder (kbgrid.cu).
#pragma once
#include "common.cuh"

namespace ib200 {

// interp.py:9-15 (lin_interp) with every operation individually rounded
__device__ __forceinline__ double kb_lookup(const double *__restrict__ table, int ntab, double x) {
    if (x >= 1.0) return 0.0;
    const double xs = __dmul_rn(x, (double)(ntab - 1));
    const int i = (int)xs;
    const double frac = __dsub_rn(xs, (double)i);
    return __dadd_rn(__dmul_rn(__dsub_rn(1.0, frac), table[i]), __dmul_rn(frac, table[i + 1]));
}

}  // namespace ib200
