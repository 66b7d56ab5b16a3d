// Error plumbing, device queries and the array-movement entry points
// (Backend.dndarray._copy_from/_copy_to/_copy/_zero; reference model:
// indigo/backends/cuda.py:127-181, pitched cudaMemcpy2D copies).
#include "common.cuh"

#include <atomic>
#include <cstring>

namespace ib200 {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DevCache { int dev = -1; int sms = 0; int64_t smem = 0; };
static thread_local DevCache g_dev;

static void refresh_dev() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (dev == g_dev.dev) return;
    int sms = 0, smem = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_dev.dev = dev; g_dev.sms = sms; g_dev.smem = smem;
}

// Side stream per device (streams and events are created once per device and kept for the life of the
// process): small latency-bound kernels that are independent of a call's main kernel are forked onto it
// and joined back with events, so that both share the SMs instead of running back to back.  The
// fork/join pattern is capturable in a CUDA graph.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream g_side[64];

int side_stream_begin(cudaStream_t main, cudaStream_t *side) {
    int dev = 0;
    IB200_TRY(cudaGetDevice(&dev));
    SideStream &q = g_side[dev & 63];
    if (!q.s) {
        IB200_TRY(cudaStreamCreateWithFlags(&q.s, cudaStreamNonBlocking));
        IB200_TRY(cudaEventCreateWithFlags(&q.fork, cudaEventDisableTiming));
        IB200_TRY(cudaEventCreateWithFlags(&q.join, cudaEventDisableTiming));
    }
    IB200_TRY(cudaEventRecord(q.fork, main));
    IB200_TRY(cudaStreamWaitEvent(q.s, q.fork, 0));
    *side = q.s;
    return 0;
}

int side_stream_end(cudaStream_t main) {
    int dev = 0;
    IB200_TRY(cudaGetDevice(&dev));
    SideStream &q = g_side[dev & 63];
    IB200_REQUIRE(q.s != nullptr, "side stream was never opened on this device");
    IB200_TRY(cudaEventRecord(q.join, q.s));
    IB200_TRY(cudaStreamWaitEvent(main, q.join, 0));
    return 0;
}

int sm_count() { refresh_dev(); return g_dev.sms > 0 ? g_dev.sms : 148; }
int64_t smem_optin() { refresh_dev(); return g_dev.smem > 0 ? g_dev.smem : 232448; }

}  // namespace ib200

using namespace ib200;

extern "C" {

const char *ib200_last_error(void) { return g_err; }

int ib200_version(void) { return 100; }

int ib200_device_info(int dev, int *sm_count_out, int64_t *smem_optin_out, int64_t *l2_bytes, int64_t *mem_bytes) {
    cudaDeviceProp p;
    IB200_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    if (smem_optin_out) *smem_optin_out = (int64_t)p.sharedMemPerBlockOptin;
    if (l2_bytes) *l2_bytes = (int64_t)p.l2CacheSize;
    if (mem_bytes) *mem_bytes = (int64_t)p.totalGlobalMem;
    return 0;
}

int64_t ib200_launch_count(void) { return g_launches.load(); }
void ib200_launch_count_reset(void) { g_launches.store(0); }

int ib200_copy2d(void *stream, void *dst, int64_t dpitch, const void *src, int64_t spitch,
                 int64_t width, int64_t height, int kind) {
    IB200_REQUIRE(kind >= 0 && kind <= 2, "kind must be 0 (D2D), 1 (H2D) or 2 (D2H)");
    IB200_REQUIRE(width >= 0 && height >= 0, "negative extent");
    if (width == 0 || height == 0) return 0;
    IB200_REQUIRE(dst && src, "null pointer");
    IB200_REQUIRE(dpitch >= width && spitch >= width, "pitch smaller than row width");
    cudaMemcpyKind k = kind == 0 ? cudaMemcpyDeviceToDevice : kind == 1 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (height == 1 || (dpitch == width && spitch == width)) {
        IB200_TRY(cudaMemcpyAsync(dst, src, (size_t)(width * height), k, as_stream(stream)));
    } else {
        IB200_TRY(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, k,
                                    as_stream(stream)));
    }
    return 0;
}

int ib200_memset0(void *stream, void *dst, int64_t nbytes) {
    IB200_REQUIRE(nbytes >= 0, "negative size");
    if (nbytes == 0) return 0;
    IB200_REQUIRE(dst, "null pointer");
    IB200_TRY(cudaMemsetAsync(dst, 0, (size_t)nbytes, as_stream(stream)));
    return 0;
}

int ib200_stream_sync(void *stream) {
    IB200_TRY(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

}  // extern "C"
