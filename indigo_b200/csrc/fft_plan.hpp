// Host-side FFT planning shared by the library (fft.cu) and the CPU emulation
// harness (tests/csrc/fft_emul.cu): radix selection, tile geometry and the
// per-axis pass sequence.  The pass "launcher" is a template parameter, so the
// emulation runs exactly the passes the device would.
#pragma once
#include "fft_core.cuh"

#include <cmath>
#include <vector>

namespace ib200 {

struct AxisPlan {
    int n = 1;
    FftStages st{};
    bool has_generic = false;
    c64 *tw_dev = nullptr;       // n twiddles, device (or host in the emulation)
};

struct FftPlanData {
    int ndim = 0;
    int64_t dims[3] = {1, 1, 1};
    int64_t batch = 1;
    AxisPlan ax[3];
    int dev = -1;
};

inline void fft_split_pow2(int e, std::vector<int> &out) {
    if (e <= 0) return;
    const int nst = (e + 3) / 4;                 // radices up to 16
    const int base = e / nst, rem = e % nst;
    for (int i = 0; i < nst; ++i) out.push_back(1 << (base + (i < rem ? 1 : 0)));
}

// Stage order: odd specialised radices (largest first), powers of two, then
// generic primes (they read shared memory, so they must not come first unless
// they are alone).
inline int fft_factorize(int n, FftStages *st, bool *has_generic) {
    std::vector<int> odd, generic, pow2, all;
    int m = n, e2 = 0;
    while (m % 2 == 0) { m /= 2; ++e2; }
    const int small[5] = {13, 11, 7, 5, 3};
    for (int f : small) while (m % f == 0) { odd.push_back(f); m /= f; }
    for (int f = 17; (int64_t)f * f <= m; f += 2) while (m % f == 0) { generic.push_back(f); m /= f; }
    if (m > 1) generic.push_back(m);
    fft_split_pow2(e2, pow2);
    for (int r : odd) all.push_back(r);
    for (int r : pow2) all.push_back(r);
    for (int r : generic) all.push_back(r);
    if ((int)all.size() > kMaxStages) return IB200_E_UNSUPPORTED;
    st->nst = (int)all.size();
    for (int i = 0; i < st->nst; ++i) st->radix[i] = all[i];
    *has_generic = !generic.empty();
    return 0;
}

inline int fft_plan_init(FftPlanData *p, int ndim, const int64_t *dims, int64_t batch) {
    p->ndim = ndim; p->batch = batch;
    for (int a = 0; a < ndim; ++a) {
        if (dims[a] < 1 || dims[a] >= (1LL << 24)) {
            set_error("fft: axis %d has unsupported length %lld", a, (long long)dims[a]);
            return IB200_E_INVALID;
        }
        p->dims[a] = dims[a];
        p->ax[a].n = (int)dims[a];
        if (dims[a] > 1) {
            int rc = fft_factorize((int)dims[a], &p->ax[a].st, &p->ax[a].has_generic);
            if (rc) { set_error("fft: cannot factor axis length %lld", (long long)dims[a]); return rc; }
        }
    }
    return 0;
}

inline void fft_make_twiddles(int n, std::vector<c64> &tw) {
    tw.resize((size_t)n);
    for (int j = 0; j < n; ++j) {
        const double ang = -2.0 * M_PI * (double)j / (double)n;
        tw[j] = mk((float)std::cos(ang), (float)std::sin(ang));
    }
}

inline int fft_ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

struct PassGeom { int L; int nbuf; size_t smem; bool load_first, store_last; };

inline int fft_plan_pass(const AxisPlan &ax, bool axis0, int64_t inner, int64_t lines_per_item, bool diag,
                         int64_t smem_limit, PassGeom *g) {
    const int nst = ax.st.nst;
    // Axis 0 transposes through shared memory on the way in and out (coalesced
    // both ways); a lone stage or a generic first radix also stages its input.
    const bool load_first = axis0 || nst == 1 || !fft_radix_is_special(ax.st.radix[0]);
    const bool store_last = axis0;
    const int inter = (load_first ? 1 : 0) + (nst - 1) + (store_last ? 1 : 0);
    const int nbuf = inter >= 2 ? 2 : 1;
    const int64_t hard = smem_limit - 1024;
    const int64_t half = (smem_limit - 2048) / 2;              // two CTAs per SM when possible
    auto fits = [&](int L, int64_t lim) { return (int64_t)nbuf * ax.n * (L + 1) * (int64_t)sizeof(c64) <= lim; };
    int L = 16;
    while (L < 256 && (int64_t)L * ax.n < 4096) L *= 2;      // short axes: more lines per tile
    if (!axis0) {                                              // no point exceeding the inner extent
        int cap = 1;
        while (cap < inner && cap < 256) cap *= 2;
        if (L > cap) L = cap;
    }
    while (L > 16 && !fits(L, half)) L /= 2;
    while (L > 1 && !fits(L, hard)) L /= 2;
    if (!fits(L, hard)) {
        set_error("fft: axis length %d does not fit a shared-memory tile", ax.n);
        return IB200_E_UNSUPPORTED;
    }
    if (axis0 && diag)                                         // keep a tile inside one batch item
        while (L > 1 && (lines_per_item % L) != 0) L /= 2;
    g->L = L; g->nbuf = nbuf; g->load_first = load_first; g->store_last = store_last;
    g->smem = (size_t)nbuf * ax.n * (L + 1) * sizeof(c64);
    return 0;
}

// Launcher: int operator()(bool axis0, int64_t blocks, size_t smem, const FftKernelArgs&)
template <class Launcher>
int fft_exec_passes(const FftPlanData *pl, c64 *y, const c64 *x, int direction, const c64 *din, int conj_in,
                    const c64 *dout, int conj_out, int64_t smem_limit, Launcher &&launch, bool *copied_only) {
    IB200_REQUIRE(direction == IB200_FFT_FORWARD || direction == IB200_FFT_INVERSE, "direction must be -1 or +1");
    int64_t plane = 1;
    for (int a = 0; a < pl->ndim; ++a) plane *= pl->dims[a];
    const int64_t total = plane * pl->batch;
    *copied_only = false;
    if (total == 0) return 0;
    IB200_REQUIRE(x && y, "null pointer");
    int axes[3], npass = 0;
    for (int a = 0; a < pl->ndim; ++a) if (pl->dims[a] > 1) axes[npass++] = a;
    if (npass == 0) {
        IB200_REQUIRE(!din && !dout, "diagonal fusion needs at least one axis longer than 1");
        *copied_only = true;
        return 0;
    }
    const bool inv = direction == IB200_FFT_INVERSE;
    for (int i = 0; i < npass; ++i) {
        const int a = axes[i];
        const AxisPlan &ax = pl->ax[a];
        int64_t inner = 1;
        for (int b = 0; b < a; ++b) inner *= pl->dims[b];
        const int64_t outer = total / (inner * ax.n);
        const bool axis0 = inner == 1;
        const bool first = i == 0, last = i == npass - 1;
        PassGeom g;
        int rc = fft_plan_pass(ax, axis0, inner, plane / ax.n, (first && din) || (last && dout), smem_limit, &g);
        if (rc) return rc;
        FftKernelArgs k;
        k.x = first ? x : y; k.y = y; k.tw = ax.tw_dev;
        k.din = first ? din : nullptr; k.dout = last ? dout : nullptr;
        k.conj_in = conj_in; k.conj_out = conj_out;
        k.inner = inner; k.outer = outer; k.plane = plane;
        k.n = ax.n; k.L = g.L; k.log2L = fft_ilog2(g.L);
        k.swap_in = (inv && first) ? 1 : 0; k.swap_out = (inv && last) ? 1 : 0;
        k.load_first = g.load_first; k.store_last = g.store_last;
        k.in0 = 0; k.in1 = ax.n; k.out0 = 0; k.out1 = ax.n;
        k.outer_stride = (int64_t)ax.n * inner;
        k.st = ax.st;
        const int64_t blocks = axis0 ? ceil_div(outer, g.L) : ceil_div(inner, g.L) * outer;
        IB200_REQUIRE(blocks < (1LL << 31), "fft: too many tiles for one launch");
        rc = launch(axis0, blocks, g.smem, k);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace ib200
