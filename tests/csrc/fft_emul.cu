// TEST INFRASTRUCTURE ONLY: runs the library's FFT tile logic (fft_core.cuh)
// and planner (fft_plan.hpp) on the CPU, one emulated CTA at a time with a
// single emulated thread, so that index arithmetic, radix butterflies, stage
// sequencing and diagonal fusion can be verified in the GPU-less container.
// The product never links this file.
#include "../../indigo_b200/csrc/fft_plan.hpp"
#include "../../indigo_b200/csrc/fft_il.cuh"
#include "../../indigo_b200/csrc/fft_pk.cuh"

#include <cmath>
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

static int g_pk_used = 0;
extern "C" const char *emul_last_error() { return g_err; }
extern "C" int emul_pk_used() { const int v = g_pk_used; g_pk_used = 0; return v; }

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
        // same routing as try_spec (fft.cu): in-place strided passes over whole tiles use the packed bodies
        if (use_spec && !axis0 && k.inner >= kSpecL && k.x == k.y && !k.din && !k.dout && k.inner % kSpecL == 0 &&
            !(k.swap_in && k.swap_out) && getenv("IB200_FFT_NOPK") == nullptr && (k.outer_stride & 1) == 0 &&
            ((uintptr_t)k.y & 15) == 0) {
            IlPassArgs a;
            a.x = k.y; a.tw = k.tw; a.inner = k.inner; a.outer = k.outer; a.outer_stride = k.outer_stride;
            a.pstride = (unsigned)k.inner; a.in0 = 0; a.in1 = k.n; a.out0 = 0; a.out1 = k.n;
#define EMUL_PKI(n, r0, r1, r2)                                                                    \
            if (!spec && fft_spec_matches(k, n, r0, r1, r2)) {                                     \
                spec = true; ++g_pk_used;                                                          \
                std::vector<float> sp(2 * pk_buf_floats(n) + 4);                                   \
                std::vector<c64> raw((size_t)n * kSpecL + 2);                                      \
                const bool pkp = getenv("IB200_FFT_NOPKP") == nullptr && pkp_mid_pairs(n, r1, r2, 256) <= 16; \
                const int64_t nb = (k.inner / kSpecL) * k.outer;                                   \
                for (int64_t b = 0; b < nb; ++b) {                                                 \
                    if (pkp) {                                                                     \
                        pkp_prefetch(a, b, raw.data(), 0, 1);                                      \
                        if (k.swap_in)       fft_pkp_tile_body<n, r0, r1, r2, true, false, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                        else if (k.swap_out) fft_pkp_tile_body<n, r0, r1, r2, false, true, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                        else                 fft_pkp_tile_body<n, r0, r1, r2, false, false, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                    } else {                                                                       \
                        if (k.swap_in)       fft_pk_pass_body<n, r0, r1, r2, true, false, 0>(a, sp.data(), b, 0, 1);  \
                        else if (k.swap_out) fft_pk_pass_body<n, r0, r1, r2, false, true, 0>(a, sp.data(), b, 0, 1);  \
                        else                 fft_pk_pass_body<n, r0, r1, r2, false, false, 0>(a, sp.data(), b, 0, 1); \
                    }                                                                              \
                }                                                                                  \
            }
            IB200_FFT_SPEC_LIST(EMUL_PKI)
#undef EMUL_PKI
            if (spec) { if (tile_L_out) tile_L_out[pass - 1] = -16; return 0; }
        }
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


// ---- fused SENSE transforms on the interleaved grid (fft_il.cuh + windowed strided passes) ----
// Mirrors ib200_sense_expand_fft / ib200_sense_ifft_combine of fft.cu pass by pass.
static const int32_t *g_win = nullptr;       // optional support windows of the z passes (emul_sense_set_windows)
static int g_win_div = 1;
extern "C" void emul_sense_set_windows(const int32_t *win, int div) { g_win = win; g_win_div = div; }

static int emul_strided(const AxisPlan &ax, c64 *base, int64_t inner, int64_t outer, int64_t outer_stride, int in0,
                        int in1, int out0, int out1, int swap_in, int swap_out, int win_mode = 0) {
    FftKernelArgs k;
    k.x = base; k.y = base; k.tw = ax.tw_dev; k.din = k.dout = nullptr; k.conj_in = k.conj_out = 0;
    k.plane = 0; k.n = ax.n; k.L = kSpecL; k.log2L = 4; k.swap_in = swap_in; k.swap_out = swap_out;
    k.load_first = 0; k.store_last = 0; k.st = ax.st;
    k.in0 = in0; k.in1 = in1; k.out0 = out0; k.out1 = out1;
    k.inner = inner; k.outer = outer; k.outer_stride = outer_stride;
    bool done = false;
    if (inner % kSpecL == 0 && getenv("IB200_FFT_IL_GENERIC") == nullptr) {      // same choice as sense_strided_pass (fft.cu)
        IlPassArgs a;
        a.x = base; a.tw = ax.tw_dev; a.inner = inner; a.outer = outer; a.outer_stride = outer_stride;
        a.pstride = (unsigned)inner; a.in0 = in0; a.in1 = in1; a.out0 = out0; a.out1 = out1;
        if (win_mode && g_win) { a.win = g_win; a.win_mode = win_mode; a.win_div = g_win_div; }
        if (getenv("IB200_FFT_NOPK") == nullptr && (outer_stride & 1) == 0) {   // same choice as try_pk_pass (fft.cu)
#define EMUL_PK(n, r0, r1, r2)                                                                     \
            if (!done && fft_spec_matches(k, n, r0, r1, r2)) {                                     \
                done = true; ++g_pk_used;                                                          \
                std::vector<float> sp(2 * pk_buf_floats(n) + 4);                                   \
                const int64_t nb = (inner / kSpecL) * outer;                                       \
                const bool pkp = getenv("IB200_FFT_NOPKP") == nullptr && pkp_mid_pairs(n, r1, r2, 256) <= 16; \
                std::vector<c64> raw((size_t)n * kSpecL + 2);                                      \
                for (int64_t b = 0; b < nb; ++b) {                                                 \
                    if (pkp) {      /* persistent form: prefetch of tile b, then the tile body (next = none) */ \
                        if (pkp_next_active(a, b, 1, b + 1, 0, 1) < 0) continue;                   \
                        for (auto &v : raw) v = mk(NAN, NAN);                                      \
                        pkp_prefetch(a, b, raw.data(), 0, 1);                                      \
                        if (swap_in)       fft_pkp_tile_body<n, r0, r1, r2, true, false, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                        else if (swap_out) fft_pkp_tile_body<n, r0, r1, r2, false, true, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                        else               fft_pkp_tile_body<n, r0, r1, r2, false, false, 0>(a, b, -1, sp.data(), raw.data(), 0, 1); \
                        continue;                                                                  \
                    }                                                                              \
                    if (swap_in)       fft_pk_pass_body<n, r0, r1, r2, true, false, 0>(a, sp.data(), b, 0, 1);  \
                    else if (swap_out) fft_pk_pass_body<n, r0, r1, r2, false, true, 0>(a, sp.data(), b, 0, 1);  \
                    else               fft_pk_pass_body<n, r0, r1, r2, false, false, 0>(a, sp.data(), b, 0, 1); \
                }                                                                                  \
            }
            IB200_FFT_SPEC_LIST(EMUL_PK)
#undef EMUL_PK
            if (done) return 0;
        }
#define EMUL_IL(n, r0, r1, r2)                                                                     \
        if (!done && fft_spec_matches(k, n, r0, r1, r2)) {                                         \
            done = true;                                                                           \
            std::vector<c64> sp((size_t)2 * n * kSpecLP + 1);                                      \
            const int64_t nb = (inner / kSpecL) * outer;                                           \
            for (int64_t b = 0; b < nb; ++b) {                                                     \
                if (swap_in)       fft_il_pass_body<n, r0, r1, r2, true, false>(a, sp.data(), b, 0, 1);  \
                else if (swap_out) fft_il_pass_body<n, r0, r1, r2, false, true>(a, sp.data(), b, 0, 1);  \
                else               fft_il_pass_body<n, r0, r1, r2, false, false>(a, sp.data(), b, 0, 1); \
            }                                                                                      \
        }
        IB200_FFT_SPEC_LIST(EMUL_IL)
#undef EMUL_IL
        if (done) return 0;
    }
#define EMUL_SP(n, r0, r1, r2)                                                                     \
    if (!done && fft_spec_matches(k, n, r0, r1, r2)) {                                             \
        done = true;                                                                               \
        std::vector<c64> sp((size_t)2 * n * kSpecLP + 1);                                          \
        const int64_t nb = ceil_div(k.inner, kSpecL) * k.outer;                                    \
        for (int64_t b = 0; b < nb; ++b) fft_pass_body_spec<n, r0, r1, r2, false>(k, sp.data(), b, 0, 1); \
    }
    IB200_FFT_SPEC_LIST(EMUL_SP)
#undef EMUL_SP
    return done ? 0 : IB200_E_UNSUPPORTED;
}

extern "C" int emul_sense(const int64_t *N, const int64_t *oN, int64_t C, int which, float *grid, float *img,
                          const float *pf, float ar, float ai, float br, float bi) {
    FftPlanData pl;
    int rc = fft_plan_init(&pl, 3, oN, C);
    if (rc) return rc;
    std::vector<std::vector<c64>> tw(3);
    for (int a = 0; a < 3; ++a) { fft_make_twiddles(pl.ax[a].n, tw[a]); pl.ax[a].tw_dev = tw[a].data(); }
    int64_t off[3];
    for (int d = 0; d < 3; ++d) off[d] = oN[d] / 2 - N[d] / 2;
    SenseFftArgs a;
    a.N0 = (int)N[0]; a.N1 = (int)N[1]; a.N2 = (int)N[2]; a.n0 = (int)oN[0]; a.n1 = (int)oN[1]; a.n2 = (int)oN[2];
    a.off0 = (int)off[0]; a.off1 = (int)off[1]; a.off2 = (int)off[2]; a.C = (int)C; a.tw = pl.ax[0].tw_dev;
    a.img = (const c64 *)img; a.img_out = (c64 *)img; a.pf = (const c64 *)pf; a.grid = (c64 *)grid;
    a.alpha = mk(ar, ai); a.beta = mk(br, bi); a.beta_zero = (br == 0.f && bi == 0.f) ? 1 : 0;
    const int64_t sy = oN[0] * C, sz = sy * oN[1];
    FftKernelArgs k0; k0.n = a.n0; k0.st = pl.ax[0].st;
    bool done = false;
    const bool pk = getenv("IB200_FFT_NOPK") == nullptr && (C % 2) == 0;          // same choice as launch_sense_x_pk
    if (which == 0) {            // expand + forward FFT
#define EMUL_XP(n, r0, r1, r2)                                                                     \
        if (pk && !done && fft_spec_matches(k0, n, r0, r1, r2)) {                                  \
            done = true; ++g_pk_used;                                                              \
            std::vector<float> sp(2 * pk_buf_floats(n) + 4);                                       \
            for (int64_t b = 0; b < sense_x_blocks((int)N[1], (int)N[2], (int)C); ++b) sense_expand_pk_body<n, r0, r1, r2, 0>(a, sp.data(), b, 0, 1); \
        }
        IB200_FFT_SPEC_LIST(EMUL_XP)
#undef EMUL_XP
#define EMUL_X(n, r0, r1, r2)                                                                      \
        if (!done && fft_spec_matches(k0, n, r0, r1, r2)) {                                        \
            done = true;                                                                           \
            std::vector<c64> sp((size_t)2 * n * kSpecLP + 1);                                      \
            for (int64_t b = 0; b < sense_x_blocks((int)N[1], (int)N[2], (int)C); ++b) sense_expand_body<n, r0, r1, r2>(a, sp.data(), b, 0, 1); \
        }
        IB200_FFT_SPEC_LIST(EMUL_X)
#undef EMUL_X
        if (!done) return IB200_E_UNSUPPORTED;
        rc = emul_strided(pl.ax[1], (c64 *)grid + off[2] * sz, sy, N[2], sz, (int)off[1], (int)(off[1] + N[1]), 0, (int)oN[1], 0, 0);
        if (rc) return rc;
        return emul_strided(pl.ax[2], (c64 *)grid, sz, 1, sz * oN[2], (int)off[2], (int)(off[2] + N[2]), 0, (int)oN[2], 0, 0, 1);
    }
    // inverse FFT + combine
    rc = emul_strided(pl.ax[2], (c64 *)grid, sz, 1, sz * oN[2], 0, (int)oN[2], (int)off[2], (int)(off[2] + N[2]), 1, 0, 2);
    if (rc) return rc;
    rc = emul_strided(pl.ax[1], (c64 *)grid + off[2] * sz, sy, N[2], sz, 0, (int)oN[1], (int)off[1], (int)(off[1] + N[1]), 0, 0);
    if (rc) return rc;
#define EMUL_CP(n, r0, r1, r2)                                                                     \
    if (pk && !done && fft_spec_matches(k0, n, r0, r1, r2)) {                                      \
        done = true; ++g_pk_used;                                                                  \
        std::vector<float> sp(2 * pk_buf_floats(n) + 4);                                           \
        std::vector<c64> acc((size_t)N[0] * kSpecL + 1);                                           \
        for (int64_t b = 0; b < sense_x_blocks((int)N[1], (int)N[2], (int)C); ++b) sense_combine_pk_body<n, r0, r1, r2, 0>(a, sp.data(), acc.data(), b, 0, 1); \
    }
    IB200_FFT_SPEC_LIST(EMUL_CP)
#undef EMUL_CP
#define EMUL_C(n, r0, r1, r2)                                                                      \
    if (!done && fft_spec_matches(k0, n, r0, r1, r2)) {                                            \
        done = true;                                                                               \
        std::vector<c64> sp((size_t)2 * n * kSpecLP + 1), acc((size_t)N[0] * kSpecL + 1);                   \
        for (int64_t b = 0; b < sense_x_blocks((int)N[1], (int)N[2], (int)C); ++b) sense_combine_body<n, r0, r1, r2>(a, sp.data(), acc.data(), b, 0, 1); \
    }
    IB200_FFT_SPEC_LIST(EMUL_C)
#undef EMUL_C
    return done ? 0 : IB200_E_UNSUPPORTED;
}
