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

#include <new>

namespace ib200 {

static const int kFftThreads = 256;

template <bool AXIS0>
__global__ void __launch_bounds__(kFftThreads) fft_pass_kernel(const FftKernelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    fft_pass_body<AXIS0>(a, reinterpret_cast<c64 *>(smem_raw), (int64_t)blockIdx.x, (int)threadIdx.x, (int)blockDim.x);
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
    auto launch = [&](bool axis0, int64_t blocks, size_t smem, const FftKernelArgs &k) -> int {
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
