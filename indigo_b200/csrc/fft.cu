// Batched 1-D/2-D/3-D complex64 FFT (unscaled forward, unscaled inverse).
//
// Interface replaced: Backend.fftn / ifftn (indigo/backends/backend.py:497-509),
// semantics of the numpy backend (np.py:102-115); the reference GPU path is a
// cuFFT plan cache (cuda.py:470-498).  Hand-written instead:
//   * one pass per axis, each a shared-memory Stockham auto-sort over a tile of
//     L lines x n points; on strided axes the first radix stage reads global
//     memory directly and the last one writes it directly, so a pass moves every
//     element through HBM exactly once in each direction (16 B per point);
//   * strided axes (1, 2) take tiles of L=16 neighbouring lines so that every
//     global access is a 128-byte segment; axis 0 takes L consecutive lines
//     (one contiguous chunk) and transposes through shared memory;
//   * mixed radix 2,3,4,5,7,8,11,13,16 in registers (416 = 13*8*4), generic
//     butterfly for other primes; the inverse reuses the forward butterflies by
//     swapping re/im on load and store; optional diagonal multiply fused into
//     the first load / last store (coil maps, apodisation, centring phase).
// Tile logic lives in fft_core.cuh, planning in fft_plan.hpp (both shared with
// the CPU emulation harness used by the GPU-less tests).
#include "fft_plan.hpp"

#include <cstdlib>
#include <new>

namespace ib200 {

static const int kFftThreads = 256;

template <bool AXIS0>
__global__ void __launch_bounds__(kFftThreads) fft_pass_kernel(const FftKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fft_pass_body<AXIS0>(a, reinterpret_cast<c64 *>(smem_raw), (int64_t)blockIdx.x, (int)threadIdx.x, (int)blockDim.x);
}

template <int N, int R0, int R1, int R2, bool AXIS0, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS > 256 ? 2 : 2)) fft_spec_kernel(const FftKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fft_pass_body_spec<N, R0, R1, R2, AXIS0>(a, reinterpret_cast<c64 *>(smem_raw), (int64_t)blockIdx.x,
                                             (int)threadIdx.x, (int)blockDim.x);
}

template <int N, int R0, int R1, int R2, bool AXIS0, int THREADS>
static int launch_spec_t(cudaStream_t s, const FftKernelArgs &k) {
    static bool attr_done[64] = {false};
    const size_t smem = (size_t)(R2 > 1 ? 2 : 1) * N * kSpecLP * sizeof(c64);
    int dev = 0;
    IB200_TRY(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        IB200_TRY(cudaFuncSetAttribute(fft_spec_kernel<N, R0, R1, R2, AXIS0, THREADS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[dev & 63] = true;
    }
    const int64_t blocks = AXIS0 ? ceil_div(k.outer, kSpecL) : ceil_div(k.inner, kSpecL) * k.outer;
    IB200_REQUIRE(blocks < (1LL << 31), "fft: too many tiles for one launch");
    fft_spec_kernel<N, R0, R1, R2, AXIS0, THREADS><<<(unsigned)blocks, THREADS, smem, s>>>(k);
    IB200_LAUNCH_CHECK();
    return 0;
}

template <int N, int R0, int R1, int R2, bool AXIS0>
static int launch_spec(cudaStream_t s, const FftKernelArgs &k) {
    static const int threads = getenv("IB200_FFT_THREADS") ? atoi(getenv("IB200_FFT_THREADS")) : 256;
    if (threads == 512) return launch_spec_t<N, R0, R1, R2, AXIS0, 512>(s, k);
    if (threads == 128) return launch_spec_t<N, R0, R1, R2, AXIS0, 128>(s, k);
    return launch_spec_t<N, R0, R1, R2, AXIS0, 256>(s, k);
}

// returns 1 if a specialised kernel was launched, 0 if none applies, <0 / >0 on error (offset by 1000)
static int try_spec(cudaStream_t s, bool axis0, const FftKernelArgs &k) {
    if (!axis0 && k.inner < kSpecL) return 0;
    if (axis0 && k.outer < kSpecL) return 0;
#define IB200_TRY_SPEC(n, r0, r1, r2)                                                        \
    if (fft_spec_matches(k, n, r0, r1, r2)) {                                                \
        const int rc = axis0 ? launch_spec<n, r0, r1, r2, true>(s, k) : launch_spec<n, r0, r1, r2, false>(s, k); \
        return rc == 0 ? 1 : (rc > 0 ? rc + 1000 : rc);                                      \
    }
    IB200_FFT_SPEC_LIST(IB200_TRY_SPEC)
#undef IB200_TRY_SPEC
    return 0;
}

}  // namespace ib200

struct ib200_fft_plan_s {
    ib200::FftPlanData d;
};

namespace ib200 {

static int ensure_device_state(FftPlanData *pl) {
    int dev = 0;
    IB200_TRY(cudaGetDevice(&dev));
    if (pl->dev == dev) return 0;
    IB200_REQUIRE(pl->dev == -1, "fft plan used on a different device than it was first executed on");
    for (int a = 0; a < pl->ndim; ++a) {
        AxisPlan &ax = pl->ax[a];
        if (ax.n <= 1) continue;
        std::vector<c64> tw;
        fft_make_twiddles(ax.n, tw);
        IB200_TRY(cudaMalloc(&ax.tw_dev, sizeof(c64) * (size_t)ax.n));
        IB200_TRY(cudaMemcpy(ax.tw_dev, tw.data(), sizeof(c64) * (size_t)ax.n, cudaMemcpyHostToDevice));
    }
    IB200_TRY(cudaFuncSetAttribute(fft_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin()));
    IB200_TRY(cudaFuncSetAttribute(fft_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin()));
    pl->dev = dev;
    return 0;
}

static int exec_impl(FftPlanData *pl, cudaStream_t s, c64 *y, const c64 *x, int direction, const c64 *din,
                     int conj_in, const c64 *dout, int conj_out) {
    int64_t total = pl->batch;
    for (int a = 0; a < pl->ndim; ++a) total *= pl->dims[a];
    if (total == 0) return 0;
    int rc = ensure_device_state(pl);
    if (rc) return rc;
    bool copy_only = false;
    static const bool use_spec = getenv("IB200_FFT_GENERIC") == nullptr;
    auto launch = [&](bool axis0, int64_t blocks, size_t smem, const FftKernelArgs &k) -> int {
        if (use_spec) {
            const int sp = try_spec(s, axis0, k);
            if (sp == 1) return 0;
            if (sp != 0) return sp > 1000 ? sp - 1000 : sp;
        }
        if (axis0) fft_pass_kernel<true><<<(unsigned)blocks, kFftThreads, smem, s>>>(k);
        else       fft_pass_kernel<false><<<(unsigned)blocks, kFftThreads, smem, s>>>(k);
        IB200_LAUNCH_CHECK();
        return 0;
    };
    rc = fft_exec_passes(pl, y, x, direction, din, conj_in, dout, conj_out, smem_optin(), launch, &copy_only);
    if (rc) return rc;
    if (copy_only && x != y)
        IB200_TRY(cudaMemcpyAsync(y, x, (size_t)total * sizeof(c64), cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_fft_plan_create(ib200_fft_plan *plan, int ndim, const int64_t *dims, int64_t batch) {
    IB200_REQUIRE(plan && dims, "null pointer");
    IB200_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1, 2 or 3");
    IB200_REQUIRE(batch >= 0, "negative batch");
    ib200_fft_plan_s *p = new (std::nothrow) ib200_fft_plan_s();
    if (!p) { set_error("out of host memory"); return IB200_E_NOMEM; }
    int rc = fft_plan_init(&p->d, ndim, dims, batch);
    if (rc) { delete p; return rc; }
    *plan = p;
    return 0;
}

int ib200_fft_plan_destroy(ib200_fft_plan plan) {
    if (!plan) return 0;
    for (int a = 0; a < 3; ++a)
        if (plan->d.ax[a].tw_dev) cudaFree(plan->d.ax[a].tw_dev);
    delete plan;
    return 0;
}

int ib200_fft_plan_describe(ib200_fft_plan plan, int axis, int *radices, int max) {
    IB200_REQUIRE(plan && axis >= 0 && axis < plan->d.ndim, "bad plan/axis");
    const FftStages &st = plan->d.ax[axis].st;
    if (plan->d.ax[axis].n <= 1) return 0;
    for (int i = 0; i < st.nst && i < max; ++i) radices[i] = st.radix[i];
    return st.nst;
}

int ib200_fft_exec(ib200_fft_plan plan, void *stream, void *y, const void *x, int direction) {
    IB200_REQUIRE(plan, "null plan");
    return exec_impl(&plan->d, as_stream(stream), (c64 *)y, (const c64 *)x, direction, nullptr, 0, nullptr, 0);
}

int ib200_fft_exec_diag(ib200_fft_plan plan, void *stream, void *y, const void *x, int direction, const void *d_in,
                        int conj_in, const void *d_out, int conj_out) {
    IB200_REQUIRE(plan, "null plan");
    return exec_impl(&plan->d, as_stream(stream), (c64 *)y, (const c64 *)x, direction, (const c64 *)d_in, conj_in,
                     (const c64 *)d_out, conj_out);
}

}  // extern "C"
