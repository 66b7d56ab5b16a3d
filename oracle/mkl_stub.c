/*
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * One-function stand-in for MKL's `mkl_ccsrmv`, the only external symbol that
 * the reference's indigo/backends/_customcpu.c needs at link time
 * (_customcpu.c:12, called only on the `N==1 && !transA` branch, :43-47).
 * MKL is not in this image, so the symbol is provided here with the same
 * documented semantics (y = alpha*op(A)*x + beta*y, general matrix, 0-based
 * "C" indexing as requested by the "G NC" descriptor the reference passes).
 * It is written from the MKL Sparse BLAS level-2 documentation, not from any
 * reference source.
 */
#include <complex.h>

void mkl_ccsrmv(const char *transa, const int *m, const int *k,
                const float complex *alpha, const char *matdescra,
                const float complex *val, const int *indx,
                const int *pntrb, const int *pntre,
                const float complex *x, const float complex *beta,
                float complex *y)
{
    (void)matdescra;
    const int rows = *m, cols = *k;
    const int base = pntrb[0];
    const char t = transa[0];
    if (t == 'N' || t == 'n') {
        for (int r = 0; r < rows; ++r) {
            float complex acc = 0.0f;
            for (int p = pntrb[r] - base; p < pntre[r] - base; ++p)
                acc += val[p] * x[indx[p]];
            y[r] = (*alpha) * acc + (*beta) * y[r];
        }
    } else {
        const int conj = (t == 'C' || t == 'c');
        for (int c = 0; c < cols; ++c)
            y[c] = (*beta) * y[c];
        for (int r = 0; r < rows; ++r) {
            const float complex ax = (*alpha) * x[r];
            for (int p = pntrb[r] - base; p < pntre[r] - base; ++p) {
                const float complex v = conj ? conjf(val[p]) : val[p];
                y[indx[p]] += v * ax;
            }
        }
    }
}
