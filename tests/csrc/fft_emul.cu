// TEST INFRASTRUCTURE ONLY: runs the library's FFT tile logic (fft_core.cuh)
// and planner (fft_plan.hpp) on the CPU, one emulated CTA at a time with a
// single emulated thread, so that index arithmetic, radix butterflies, stage
// sequencing and diagonal fusion can be verified in the GPU-less container.
// The product never links this file.
#include "../../indigo_b200/csrc/fft_plan.hpp"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace ib200 {
static char g_err[512];
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
void count_launch(int) {}
int sm_count() { return 148; }
int64_t smem_optin() { return 232448; }
}  // namespace ib200

using namespace ib200;

extern "C" const char *emul_last_error() { return g_err; }

extern "C" int emul_fft(int ndim, const int64_t *dims, int64_t batch, float *y, const float *x, int direction,
                        const float *din, int conj_in, const float *dout, int conj_out, int *tile_L_out) {
    const bool use_spec = getenv("IB200_FFT_GENERIC") == nullptr;
    FftPlanData pl;
    int rc = fft_plan_init(&pl, ndim, dims, batch);
    if (rc) return rc;
    std::vector<std::vector<c64>> tw(3);
    for (int a = 0; a < ndim; ++a) {
        if (pl.ax[a].n > 1) { fft_make_twiddles(pl.ax[a].n, tw[a]); pl.ax[a].tw_dev = tw[a].data(); }
    }
    int pass = 0;
    auto launch = [&](bool axis0, int64_t blocks, size_t smem, const FftKernelArgs &k) -> int {
        std::vector<c64> sm(smem / sizeof(c64) + 1);
        if (tile_L_out) tile_L_out[pass] = k.L;
        ++pass;
        bool spec = false;
        if (use_spec && (axis0 ? k.outer >= kSpecL : k.inner >= kSpecL)) {
#define EMUL_SPEC(n, r0, r1, r2)                                                                   \
            if (!spec && fft_spec_matches(k, n, r0, r1, r2)) {                                     \
                spec = true;                                                                       \
                std::vector<c64> sp((size_t)2 * n * kSpecLP + 1);                                  \
                const int64_t nb = axis0 ? ceil_div(k.outer, kSpecL) : ceil_div(k.inner, kSpecL) * k.outer; \
                for (int64_t b = 0; b < nb; ++b) {                                                 \
                    if (axis0) fft_pass_body_spec<n, r0, r1, r2, true>(k, sp.data(), b, 0, 1);     \
                    else       fft_pass_body_spec<n, r0, r1, r2, false>(k, sp.data(), b, 0, 1);    \
                }                                                                                  \
            }
            IB200_FFT_SPEC_LIST(EMUL_SPEC)
#undef EMUL_SPEC
        }
        if (spec) { if (tile_L_out) tile_L_out[pass - 1] = -16; return 0; }
        for (int64_t b = 0; b < blocks; ++b) {
            if (axis0) fft_pass_body<true>(k, sm.data(), b, 0, 1);
            else       fft_pass_body<false>(k, sm.data(), b, 0, 1);
        }
        return 0;
    };
    bool copy_only = false;
    rc = fft_exec_passes(&pl, (c64 *)y, (const c64 *)x, direction, (const c64 *)din, conj_in, (const c64 *)dout,
                         conj_out, smem_optin(), launch, &copy_only);
    if (rc) return rc;
    if (copy_only && x != y) {
        int64_t total = batch; for (int a = 0; a < ndim; ++a) total *= dims[a];
        for (int64_t i = 0; i < 2 * total; ++i) y[i] = x[i];
    }
    return 0;
}
