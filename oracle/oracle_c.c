/*
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under indigo_b200/
 * may link, load or call this file.  It is the plain-C restatement of the
 * integer / scalar parts of the reference's hot path, used by tests/ (and by
 * bench.py's cpu_baseline leg) as the checker.
 *
 * Parity: pinned.  tests/test_oracle.py checks every function here against
 * (a) golden vectors produced by the unmodified reference NumpyBackend
 * (tests/golden/make_golden.py) and (b) oracle/_ref/_customcpu, the
 * reference's own _customcpu.c compiled where it lies.
 */
#include <complex.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Follows indigo/backends/_customcpu.c:179-215 (py_inspect): count rows and
 * columns that hold at least one stored entry, and report whether every column
 * holds at most one (the "exclusive write" property the adjoint scatter uses).
 * out = {nzrows, nzcols, exwrite}.  64-bit counters, 32-bit indices like the
 * reference (which reads them as unsigned int, _customcpu.c:188-189). */
void oracle_csr_inspect(int64_t m, int64_t k, const int32_t *colind,
                        const int32_t *rowptr, int64_t out[3])
{
    int32_t *per_col = (int32_t *)calloc((size_t)(k > 0 ? k : 1), sizeof(int32_t));
    int64_t nzrows = 0, nzcols = 0, exw = 1;
    for (int64_t r = 0; r < m; ++r) {
        if (rowptr[r + 1] > rowptr[r]) nzrows++;
        for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) per_col[colind[p]]++;
    }
    for (int64_t c = 0; c < k; ++c) {
        if (per_col[c] > 0) nzcols++;
        if (per_col[c] > 1) exw = 0;
    }
    free(per_col);
    out[0] = nzrows; out[1] = nzcols; out[2] = exw;
}

/* Y = alpha*op(A)*X + beta*Y with op = identity or conjugate transpose.
 * Semantics of Backend.ccsrmm (indigo/backends/backend.py:514-519) as the
 * numpy backend evaluates it (indigo/backends/np.py:120-127) and as
 * _customcpu.c:14-114 implements it: column-major X (ldx) and Y (ldy),
 * 0-based CSR with int32 indices, complex64 data, fp32 accumulation in row
 * order.  beta == 0 overwrites Y without reading it (the reference computes
 * beta*Y and is therefore NaN-unsafe on an uninitialised arena, SURVEY.md
 * section 0 landmine 3; the oracle is always run on initialised Y). */
void oracle_ccsrmm(int adjoint, int64_t m, int64_t n, int64_t k,
                   float alpha_re, float alpha_im,
                   const float complex *vals, const int32_t *colind,
                   const int32_t *rowptr,
                   const float complex *X, int64_t ldx,
                   float beta_re, float beta_im,
                   float complex *Y, int64_t ldy)
{
    const float complex alpha = alpha_re + I * alpha_im;
    const float complex beta = beta_re + I * beta_im;
    const int beta_zero = (beta_re == 0.0f && beta_im == 0.0f);
    if (!adjoint) {
        #pragma omp parallel for schedule(static)
        for (int64_t r = 0; r < m; ++r) {
            for (int64_t c = 0; c < n; ++c) {
                float complex acc = 0.0f;
                for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p)
                    acc += vals[p] * X[colind[p] + c * ldx];
                float complex *y = &Y[r + c * ldy];
                *y = beta_zero ? alpha * acc : alpha * acc + beta * (*y);
            }
        }
    } else {
        #pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < n; ++c) {
            float complex *y = &Y[c * ldy];
            for (int64_t j = 0; j < k; ++j) y[j] = beta_zero ? 0.0f : beta * y[j];
            for (int64_t r = 0; r < m; ++r) {
                const float complex ax = alpha * X[r + c * ldx];
                for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p)
                    y[colind[p]] += conjf(vals[p]) * ax;
            }
        }
    }
}

/* Follows custom_onemm, _customcpu.c:117-134: column sums of X broadcast down
 * the m rows of Y.  X is (k x n), Y is (m x n). */
void oracle_onemm(int64_t m, int64_t n, int64_t k,
                  float alpha_re, float alpha_im, const float complex *X, int64_t ldx,
                  float beta_re, float beta_im, float complex *Y, int64_t ldy)
{
    const float complex alpha = alpha_re + I * alpha_im;
    const float complex beta = beta_re + I * beta_im;
    for (int64_t c = 0; c < n; ++c) {
        float complex acc = 0.0f;
        for (int64_t j = 0; j < k; ++j) acc += X[j + c * ldx];
        for (int64_t r = 0; r < m; ++r)
            Y[r + c * ldy] = beta * Y[r + c * ldy] + alpha * acc;
    }
}

/* Follows c_max, _customcpu.c:242-246: elementwise max over the float view. */
void oracle_fmax(int64_t nfloats, float val, float *arr)
{
    for (int64_t i = 0; i < nfloats; ++i)
        arr[i] = arr[i] > val ? arr[i] : val;
}
