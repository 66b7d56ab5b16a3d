// BLAS-1 on complex64 vectors: axpby / scal / dotc / nrm2^2 and the fused CG
// updates.  Replaces the reference's two-pass cublasCscal+cublasCaxpy and the
// blocking cublasCdotc / cublasScnrm2 calls (indigo/backends/cuda.py:239-302);
// semantics follow the numpy backend (indigo/backends/np.py:53-74).
//
// All kernels are pure HBM streams: 128-bit loads when both operands share a
// 16-byte phase, grid sized to a multiple of the SM count, fp64 accumulation
// with a fixed two-level reduction order so results are bit-reproducible (every
// rank of a coil-sharded run computes identical CG scalars).
#include "common.cuh"

namespace ib200 {

static const int kThreads = 256;
static const int kMaxBlocks = 148 * 8;          // partial slots per reduction
static const int kSlots = 8;                    // reductions that may be in flight

struct ReduceWs {
    double *partials = nullptr;                 // kSlots * kMaxBlocks * 2 doubles
    unsigned *tickets = nullptr;                // kSlots counters
    double *result = nullptr;                   // kSlots * 2 doubles (device)
    double *host = nullptr;                     // kSlots * 2 doubles (pinned)
    int next = 0;
};
static ReduceWs g_ws[64];

static int get_ws(ReduceWs **out) {
    int dev = 0;
    IB200_TRY(cudaGetDevice(&dev));
    IB200_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
    ReduceWs &w = g_ws[dev];
    if (!w.partials) {
        IB200_TRY(cudaMalloc(&w.partials, sizeof(double) * 2 * kMaxBlocks * kSlots));
        IB200_TRY(cudaMalloc(&w.tickets, sizeof(unsigned) * kSlots));
        IB200_TRY(cudaMemset(w.tickets, 0, sizeof(unsigned) * kSlots));
        IB200_TRY(cudaMalloc(&w.result, sizeof(double) * 2 * kSlots));
        IB200_TRY(cudaMallocHost(&w.host, sizeof(double) * 2 * kSlots));
    }
    *out = &w;
    return 0;
}

static int grid_for(int64_t items) {
    int64_t blocks = ceil_div(items, (int64_t)kThreads * 4);
    int64_t cap = (int64_t)sm_count() * 8;
    if (cap > kMaxBlocks) cap = kMaxBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------
// axpby: MODE 0: y = a*x ; 1: y = b*y + a*x ; 2: y = b*y ; 3: y = y + a*x (b == 1)
template <int MODE>
__device__ __forceinline__ c64 axpby_one(c64 a, c64 b, c64 x, c64 y) {
    if (MODE == 0) return cmul(a, x);
    if (MODE == 1) return cfma(a, x, cmul(b, y));
    if (MODE == 2) return cmul(b, y);
    return cfma(a, x, y);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) axpby_kernel(int64_t n, c64 b, c64 *__restrict__ y, c64 a,
                                                         const c64 *__restrict__ x, int head) {
    // `head` (0/1) elements are peeled so that the bulk is 16-byte aligned.
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0 && head) {
        c64 xv = MODE == 2 ? mk(0, 0) : x[0];
        c64 yv = MODE == 0 ? mk(0, 0) : y[0];
        y[0] = axpby_one<MODE>(a, b, xv, yv);
    }
    const int64_t pairs = (n - head) >> 1;
    float4 *y4 = reinterpret_cast<float4 *>(y + head);
    const float4 *x4 = reinterpret_cast<const float4 *>(x + head);
    for (int64_t i = tid; i < pairs; i += nth) {
        float4 xv = MODE == 2 ? make_float4(0, 0, 0, 0) : __ldg(x4 + i);
        float4 yv = MODE == 0 ? make_float4(0, 0, 0, 0) : y4[i];
        c64 r0 = axpby_one<MODE>(a, b, mk(xv.x, xv.y), mk(yv.x, yv.y));
        c64 r1 = axpby_one<MODE>(a, b, mk(xv.z, xv.w), mk(yv.z, yv.w));
        y4[i] = make_float4(r0.x, r0.y, r1.x, r1.y);
    }
    if (tid == 0 && ((n - head) & 1)) {
        const int64_t j = n - 1;
        c64 xv = MODE == 2 ? mk(0, 0) : x[j];
        c64 yv = MODE == 0 ? mk(0, 0) : y[j];
        y[j] = axpby_one<MODE>(a, b, xv, yv);
    }
}

// operands with different 16-byte phases: 64-bit accesses
template <int MODE>
__global__ void __launch_bounds__(kThreads) axpby_kernel_unaligned(int64_t n, c64 b, c64 *__restrict__ y, c64 a,
                                                                   const c64 *__restrict__ x) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nth) {
        c64 xv = MODE == 2 ? mk(0, 0) : __ldg(x + i);
        c64 yv = MODE == 0 ? mk(0, 0) : y[i];
        y[i] = axpby_one<MODE>(a, b, xv, yv);
    }
}

template <int MODE>
static int launch_axpby(cudaStream_t s, int64_t n, c64 b, c64 *y, c64 a, const c64 *x) {
    const uintptr_t py = (uintptr_t)y, px = MODE == 2 ? (uintptr_t)y : (uintptr_t)x;
    const int grid = grid_for(n);
    if ((py & 15) == (px & 15)) {
        int head = (py & 15) ? 1 : 0;
        if (head > n) head = (int)n;
        axpby_kernel<MODE><<<grid, kThreads, 0, s>>>(n, b, y, a, x, head);
    } else {
        axpby_kernel_unaligned<MODE><<<grid, kThreads, 0, s>>>(n, b, y, a, x);
    }
    IB200_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------
// reductions.  Each block writes one double2 partial; the last block to finish
// (ticket) folds the partials in index order and stores the result.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level sum of (a, b); valid in thread 0.
__device__ __forceinline__ void block_sum2(double &a, double &b, double *sh /* 2*8 */) {
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[w] = a; sh[8 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        b = lane < (blockDim.x >> 5) ? sh[8 + lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
    }
    __syncthreads();
}

// Finishes a grid-wide reduction.  Returns true in thread 0 of the last block
// with the totals in (a, b).
__device__ __forceinline__ bool grid_finish(double &a, double &b, double *partials, unsigned *ticket, double *sh) {
    __shared__ bool is_last;
    block_sum2(a, b, sh);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = a;
        partials[2 * blockIdx.x + 1] = b;
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double sa = 0.0, sb = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {   // fixed order per thread
        sa += __ldcg(partials + 2 * i);
        sb += __ldcg(partials + 2 * i + 1);
    }
    block_sum2(sa, sb, sh);
    a = sa; b = sb;
    if (threadIdx.x == 0) *ticket = 0u;          // re-arm for the next use of this slot
    return threadIdx.x == 0;
}

__global__ void __launch_bounds__(kThreads) dotc_kernel(int64_t n, const c64 *__restrict__ x, const c64 *__restrict__ y,
                                                        double *partials, unsigned *ticket, double *out) {
    __shared__ double sh[16];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double re = 0.0, im = 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        const c64 a = __ldg(x + i), b = __ldg(y + i);
        re += (double)a.x * b.x + (double)a.y * b.y;
        im += (double)a.x * b.y - (double)a.y * b.x;
    }
    if (grid_finish(re, im, partials, ticket, sh)) { out[0] = re; out[1] = im; }
}

__global__ void __launch_bounds__(kThreads) nrm2sq_kernel(int64_t n, const c64 *__restrict__ x, double *partials,
                                                          unsigned *ticket, double *out) {
    __shared__ double sh[16];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double s = 0.0, z = 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        const c64 a = __ldg(x + i);
        s += (double)a.x * a.x + (double)a.y * a.y;
    }
    if (grid_finish(s, z, partials, ticket, sh)) out[0] = s;
}

// x += alpha p ; r -= alpha Ap ; scal[3] = ||r||^2, alpha = scal[0]/scal[1]
__global__ void __launch_bounds__(kThreads) cg_xr_kernel(int64_t n, c64 *__restrict__ x, c64 *__restrict__ r,
                                                         const c64 *__restrict__ p, const c64 *__restrict__ Ap,
                                                         double *scal, double *partials, unsigned *ticket) {
    __shared__ double sh[16];
    const float alpha = (float)(scal[0] / scal[1]);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    double s = 0.0, z = 0.0;
    for (int64_t i = tid; i < n; i += nth) {
        const c64 pv = __ldg(p + i), av = __ldg(Ap + i);
        c64 xv = x[i], rv = r[i];
        xv.x = fmaf(alpha, pv.x, xv.x); xv.y = fmaf(alpha, pv.y, xv.y);
        rv.x = fmaf(-alpha, av.x, rv.x); rv.y = fmaf(-alpha, av.y, rv.y);
        x[i] = xv; r[i] = rv;
        s += (double)rv.x * rv.x + (double)rv.y * rv.y;
    }
    if (grid_finish(s, z, partials, ticket, sh)) scal[3] = s;
}

// p = beta p + r, beta = scal[3]/scal[0]; afterwards scal[0] = scal[3]
__global__ void __launch_bounds__(kThreads) cg_p_kernel(int64_t n, c64 *__restrict__ p, const c64 *__restrict__ r,
                                                        double *scal, unsigned *ticket) {
    const double rr = scal[0], r2 = scal[3];
    const float beta = (float)(r2 / rr);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nth) {
        const c64 rv = __ldg(r + i);
        c64 pv = p[i];
        pv.x = fmaf(beta, pv.x, rv.x); pv.y = fmaf(beta, pv.y, rv.y);
        p[i] = pv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {        // every block has read scal[] by now
            scal[0] = r2;
            *ticket = 0u;
        }
    }
}

static int next_slot(ReduceWs *w) {
    int s = w->next;
    w->next = (w->next + 1) % kSlots;
    return s;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_caxpby(void *stream, int64_t n, float br, float bi, void *y, float ar, float ai, const void *x) {
    IB200_RANGE("ib200_caxpby");
    IB200_REQUIRE(n >= 0, "negative length");
    if (n == 0) return 0;
    IB200_REQUIRE(y != nullptr, "null y");
    cudaStream_t s = as_stream(stream);
    const c64 a = mk(ar, ai), b = mk(br, bi);
    const bool a0 = (ar == 0.f && ai == 0.f), b0 = (br == 0.f && bi == 0.f), b1 = (br == 1.f && bi == 0.f);
    if (a0 && b0) {
        IB200_TRY(cudaMemsetAsync(y, 0, (size_t)n * sizeof(c64), s));
        return 0;
    }
    if (a0 && b1) return 0;
    if (!a0) IB200_REQUIRE(x != nullptr, "null x");
    if (a0) return launch_axpby<2>(s, n, b, (c64 *)y, a, (const c64 *)y);
    if (b0) return launch_axpby<0>(s, n, b, (c64 *)y, a, (const c64 *)x);
    if (b1) return launch_axpby<3>(s, n, b, (c64 *)y, a, (const c64 *)x);
    return launch_axpby<1>(s, n, b, (c64 *)y, a, (const c64 *)x);
}

int ib200_cscal(void *stream, int64_t n, float ar, float ai, void *x) {
    return ib200_caxpby(stream, n, ar, ai, x, 0.f, 0.f, nullptr);
}

int ib200_cdotc_dev(void *stream, int64_t n, const void *x, const void *y, double *dev_out2) {
    IB200_RANGE("ib200_cdotc_dev");
    IB200_REQUIRE(n >= 0 && dev_out2, "bad arguments");
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = next_slot(w);
    dotc_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(
        n, (const c64 *)x, (const c64 *)y, w->partials + (size_t)slot * 2 * kMaxBlocks, w->tickets + slot, dev_out2);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_scnrm2sq_dev(void *stream, int64_t n, const void *x, double *dev_out1) {
    IB200_REQUIRE(n >= 0 && dev_out1, "bad arguments");
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = next_slot(w);
    nrm2sq_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(
        n, (const c64 *)x, w->partials + (size_t)slot * 2 * kMaxBlocks, w->tickets + slot, dev_out1);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_cdotc(void *stream, int64_t n, const void *x, const void *y, double *host_re, double *host_im) {
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = w->next;      // the _dev call below takes the same slot
    rc = ib200_cdotc_dev(stream, n, x, y, w->result + 2 * slot); if (rc) return rc;
    IB200_TRY(cudaMemcpyAsync(w->host + 2 * slot, w->result + 2 * slot, 2 * sizeof(double), cudaMemcpyDeviceToHost,
                              as_stream(stream)));
    IB200_TRY(cudaStreamSynchronize(as_stream(stream)));
    if (host_re) *host_re = w->host[2 * slot];
    if (host_im) *host_im = w->host[2 * slot + 1];
    return 0;
}

int ib200_scnrm2sq(void *stream, int64_t n, const void *x, double *host_out) {
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = w->next;
    rc = ib200_scnrm2sq_dev(stream, n, x, w->result + 2 * slot); if (rc) return rc;
    IB200_TRY(cudaMemcpyAsync(w->host + 2 * slot, w->result + 2 * slot, sizeof(double), cudaMemcpyDeviceToHost,
                              as_stream(stream)));
    IB200_TRY(cudaStreamSynchronize(as_stream(stream)));
    if (host_out) *host_out = w->host[2 * slot];
    return 0;
}

int ib200_cg_xr(void *stream, int64_t n, void *x, void *r, const void *p, const void *Ap, double *scal) {
    IB200_RANGE("ib200_cg_xr");
    IB200_REQUIRE(n >= 0 && x && r && p && Ap && scal, "bad arguments");
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = next_slot(w);
    cg_xr_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(
        n, (c64 *)x, (c64 *)r, (const c64 *)p, (const c64 *)Ap, scal,
        w->partials + (size_t)slot * 2 * kMaxBlocks, w->tickets + slot);
    IB200_LAUNCH_CHECK();
    return 0;
}

int ib200_cg_p(void *stream, int64_t n, void *p, const void *r, double *scal) {
    IB200_RANGE("ib200_cg_p");
    IB200_REQUIRE(n >= 0 && p && r && scal, "bad arguments");
    ReduceWs *w; int rc = get_ws(&w); if (rc) return rc;
    const int slot = next_slot(w);
    cg_p_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(n, (c64 *)p, (const c64 *)r, scal, w->tickets + slot);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
