// Adjoint gridding gather on x-runs of the stored adjoint (fused SENSE recipe: ccsrmm(G', adjoint)).
//
// The stored adjoint of the gridding matrix (csrmm_il.cu) has one row per grid point and twelve
// entries per row on average: the row-per-lane-group gather spends four instructions on overhead for
// every useful one (6 warp instructions per stored entry, profiles/r01_s6_state_cfg3.md) and gathers
// the same k-space sample once for every grid point it touches.  A Kaiser-Bessel footprint covers five
// consecutive grid points along x, so the four x-neighbours that form one row of a 4x4x4 tile share
// most of their samples.  Here the four rows of such a run are merged into one list of
//     (sample, w0, w1, w2, w3)          20 bytes per run entry, 2.5 stored entries on average
// (wi = weight of the sample at the i-th point of the run, 0 when it does not reach it).  A lane group
// walks one run: one 16-byte gather of the sample's coils serves all four points, the products are
// packed FFMA2s with the weight as a broadcast scalar, and a run is 36 entries long on average inside
// the sampled region instead of 12, so the per-row overhead is paid a quarter as often.  Lists are
// padded to multiples of four entries (weight 0) so that the loop has no tail.
//
// Runs longer than `long_thresh` entries (k-space centre of radial trajectories) are left out; their
// rows are listed for csrmm_il_long_kernel, which gives each of them a whole CTA.
#include "common.cuh"
#include "pk2.cuh"

namespace ib200 {

static const int kRun = 4;            // grid points per run = x extent of the adjoint's tiles
static const int kRunPad = 4;         // run lists are padded to multiples of this many entries

struct __align__(8) RunPacked { int32_t col; float w; };

// number of distinct samples in the four rows of run r (rows hold ascending sample indices)
__device__ __forceinline__ int run_merge(const int32_t *__restrict__ rowptr, const RunPacked *__restrict__ ent, int64_t r,
                                         int32_t *ids, float4 *w4) {
    int p[kRun], e[kRun];
#pragma unroll
    for (int i = 0; i < kRun; ++i) { p[i] = rowptr[kRun * r + i]; e[i] = rowptr[kRun * r + i + 1]; }
    int n = 0;
    while (true) {
        int m = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kRun; ++i)
            if (p[i] < e[i]) { const int c = ent[p[i]].col; m = c < m ? c : m; }
        if (m == 0x7fffffff) break;
        float w[kRun];
#pragma unroll
        for (int i = 0; i < kRun; ++i) {
            w[i] = 0.f;
            if (p[i] < e[i] && ent[p[i]].col == m) { w[i] = ent[p[i]].w; ++p[i]; }
        }
        if (ids) { ids[n] = m; w4[n] = make_float4(w[0], w[1], w[2], w[3]); }
        ++n;
    }
    return n;
}

__device__ __forceinline__ bool run_is_long(const int32_t *__restrict__ rowptr, int64_t r, int long_thresh) {
    // the merged list holds at most the sum of the four rows: a cheap, conservative bound
    return rowptr[kRun * r + kRun] - rowptr[kRun * r] > long_thresh;
}

__global__ void __launch_bounds__(128) run_count_kernel(int64_t nruns, const int32_t *__restrict__ rowptr,
                                                        const RunPacked *__restrict__ ent, int long_thresh,
                                                        int32_t *__restrict__ counts, int *nlongrows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    if (run_is_long(rowptr, r, long_thresh)) {
        counts[r] = 0;
        int c = 0;
#pragma unroll
        for (int i = 0; i < kRun; ++i) c += rowptr[kRun * r + i + 1] > rowptr[kRun * r + i];
        atomicAdd(nlongrows, c);
        return;
    }
    const int n = run_merge(rowptr, ent, r, nullptr, nullptr);
    counts[r] = (n + kRunPad - 1) / kRunPad * kRunPad;
}

__global__ void __launch_bounds__(128) run_fill_kernel(int64_t nruns, const int32_t *__restrict__ rowptr,
                                                       const RunPacked *__restrict__ ent, int long_thresh,
                                                       const int32_t *__restrict__ run_ptr, int32_t *__restrict__ ids,
                                                       float4 *__restrict__ w4, int32_t *__restrict__ longrows,
                                                       int capacity, int *nlongrows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    if (run_is_long(rowptr, r, long_thresh)) {
#pragma unroll
        for (int i = 0; i < kRun; ++i)
            if (rowptr[kRun * r + i + 1] > rowptr[kRun * r + i]) {
                const int at = atomicAdd(nlongrows, 1);
                if (at < capacity) longrows[at] = (int32_t)(kRun * r + i);
            }
        return;
    }
    const int a = run_ptr[r], b = run_ptr[r + 1];
    const int n = run_merge(rowptr, ent, r, ids + a, w4 + a);
    for (int q = a + n; q < b; ++q) { ids[q] = n ? ids[a + n - 1] : 0; w4[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
}

// ---------------------------------------------------------------------------------------------------
struct RunBatch { int4 id; float4 w[kRunPad]; };

__device__ __forceinline__ void run_load_batch(RunBatch &e, const int32_t *__restrict__ ids, const float4 *__restrict__ w4, int p) {
    e.id = __ldcs(reinterpret_cast<const int4 *>(ids + p));
#pragma unroll
    for (int u = 0; u < kRunPad; ++u) e.w[u] = __ldcs(w4 + p + u);
}

// Yil[rowmap[4*run + i]][c] = alpha * sum_e w_i(e) * Xil[id(e)][c]        (rowmap < 0: nothing stored)
// CL lanes per run, two coils per lane (one 16-byte gather per run entry and lane).
template <int CL>
__global__ void __launch_bounds__(256) csrmm_runs_kernel(int64_t nruns, int C, c64 alpha,
                                                         const int32_t *__restrict__ run_ptr,
                                                         const int32_t *__restrict__ ids, const float4 *__restrict__ w4,
                                                         const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                         c64 *__restrict__ Yil, int64_t ypitch,
                                                         const int32_t *__restrict__ rowmap,
                                                         const int32_t *__restrict__ rowptr, int long_thresh, int rpg) {
    constexpr int GPB = 256 / CL;
    const int gl = (int)(threadIdx.x & (CL - 1)), group = (int)(threadIdx.x / CL);
    const int coil = 2 * gl;
    const bool coil_ok = coil < C;
    const char *xb = reinterpret_cast<const char *>(Xil + (coil_ok ? coil : 0));
    const int64_t run0 = (int64_t)blockIdx.x * ((int64_t)GPB * rpg) + group;
    {
        // The run lists of this CTA are one contiguous range that is read exactly once, a batch at a time
        // with one batch of look-ahead: pull the whole range into L2 up front so that the batch loads pay
        // an L2 hit instead of a DRAM round trip each.
        const int64_t first = (int64_t)blockIdx.x * ((int64_t)GPB * rpg);
        const int64_t last = first + (int64_t)GPB * rpg < nruns ? first + (int64_t)GPB * rpg : nruns;
        const int e0 = __ldg(run_ptr + first), e1 = __ldg(run_ptr + last);
        for (int q = e0 + 8 * (int)threadIdx.x; q < e1; q += 8 * 256)        // 128 bytes of weights, 32 of ids
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w4 + q));
        for (int q = e0 + 32 * (int)threadIdx.x; q < e1; q += 32 * 256)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ids + q));
    }
    for (int i = 0; i < rpg; ++i) {
        const int64_t run = run0 + (int64_t)i * GPB;
        if (run >= nruns) break;
        const int a = __ldg(run_ptr + run), b = __ldg(run_ptr + run + 1);
        pk2 acc[kRun][2];
#pragma unroll
        for (int t = 0; t < kRun; ++t) { acc[t][0] = p_make(0.f, 0.f); acc[t][1] = p_make(0.f, 0.f); }
        if (a < b) {
            RunBatch e;
            run_load_batch(e, ids, w4, a);
            for (int p = a; p < b; p += kRunPad) {
                float4 x[kRunPad];
                const int idv[kRunPad] = {e.id.x, e.id.y, e.id.z, e.id.w};
#pragma unroll
                for (int u = 0; u < kRunPad; ++u)
                    x[u] = __ldg(reinterpret_cast<const float4 *>(xb + (uint64_t)(uint32_t)idv[u] * xpitch_bytes));
                RunBatch nx;
                const bool more = p + kRunPad < b;
                if (more) run_load_batch(nx, ids, w4, p + kRunPad);
#pragma unroll
                for (int u = 0; u < kRunPad; ++u) {
                    const pk2 x0 = p_make(x[u].x, x[u].y), x1 = p_make(x[u].z, x[u].w);
                    const float wv[kRun] = {e.w[u].x, e.w[u].y, e.w[u].z, e.w[u].w};
#pragma unroll
                    for (int t = 0; t < kRun; ++t) {
                        acc[t][0] = p_fma(p_bc(wv[t]), x0, acc[t][0]);
                        acc[t][1] = p_fma(p_bc(wv[t]), x1, acc[t][1]);
                    }
                }
                if (more) e = nx;
            }
        }
        // rows of a long run belong to csrmm_il_long_kernel, except its empty rows, which nobody else visits
        const bool is_long = a == b && run_is_long(rowptr, run, long_thresh);
        if (coil_ok) {
#pragma unroll
            for (int t = 0; t < kRun; ++t) {
                if (is_long && __ldg(rowptr + kRun * run + t + 1) > __ldg(rowptr + kRun * run + t)) continue;
                const int64_t out = (int64_t)__ldg(rowmap + kRun * run + t);
                if (out >= 0) {
                    const c64 o0 = cmul(alpha, mk(p_lo(acc[t][0]), p_hi(acc[t][0])));
                    const c64 o1 = cmul(alpha, mk(p_lo(acc[t][1]), p_hi(acc[t][1])));
                    __stcs(reinterpret_cast<float4 *>(Yil + out * ypitch + coil), make_float4(o0.x, o0.y, o1.x, o1.y));
                }
            }
        }
    }
}

int exclusive_scan_public(cudaStream_t s, int64_t n, const int32_t *in, int32_t *out);   // csrmm.cu
int launch_long_packed(cudaStream_t s, int CL, int nlong, const int32_t *longrows, int C, c64 alpha, const void *ent,
                       const int32_t *rowptr, const c64 *Xil, int64_t xpitch, c64 *Yil, int64_t ypitch,
                       const int32_t *rowmap);                                          // csrmm_il.cu

static int runs_pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_csr_runs_count(void *stream, int64_t kp, const int32_t *rowptr, const void *packed, int long_thresh,
                         int32_t *run_ptr, int64_t *host_entries, int *host_longrows) {
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    IB200_REQUIRE(host_entries && host_longrows && long_thresh >= 0, "bad arguments");
    *host_entries = 0; *host_longrows = 0;
    if (kp == 0) return 0;
    IB200_REQUIRE(rowptr && packed && run_ptr, "null pointer");
    const int64_t nruns = kp / kRun;
    cudaStream_t s = as_stream(stream);
    int32_t *counts = nullptr;
    IB200_TRY(cudaMalloc(&counts, (size_t)(nruns + 1) * sizeof(int32_t) + sizeof(int)));
    int *nlong = reinterpret_cast<int *>(counts + nruns + 1);
    cudaMemsetAsync(nlong, 0, sizeof(int), s);
    run_count_kernel<<<(unsigned)ceil_div(nruns, 128), 128, 0, s>>>(nruns, rowptr, (const RunPacked *)packed, long_thresh,
                                                                   counts, nlong);
    count_launch();
    int rc = exclusive_scan_public(s, nruns, counts, run_ptr);
    int32_t total = 0;
    int hl = 0;
    cudaError_t e = cudaSuccess;
    if (!rc) {
        cudaMemcpyAsync(&total, run_ptr + nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(&hl, nlong, sizeof(int), cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    cudaFree(counts);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    *host_entries = total; *host_longrows = hl;
    return 0;
}

int ib200_csr_runs_fill(void *stream, int64_t kp, const int32_t *rowptr, const void *packed, int long_thresh,
                        const int32_t *run_ptr, int32_t *ids, void *w4, int32_t *longrows, int capacity) {
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    if (kp == 0) return 0;
    IB200_REQUIRE(rowptr && packed && run_ptr && ids && w4 && (longrows || capacity == 0), "null pointer");
    IB200_REQUIRE(((uintptr_t)ids & 15) == 0 && ((uintptr_t)w4 & 15) == 0, "run arrays must be 16-byte aligned");
    const int64_t nruns = kp / kRun;
    cudaStream_t s = as_stream(stream);
    int *nlong = nullptr;
    IB200_TRY(cudaMalloc(&nlong, sizeof(int)));
    cudaMemsetAsync(nlong, 0, sizeof(int), s);
    run_fill_kernel<<<(unsigned)ceil_div(nruns, 128), 128, 0, s>>>(nruns, rowptr, (const RunPacked *)packed, long_thresh,
                                                                  run_ptr, ids, (float4 *)w4, longrows, capacity, nlong);
    count_launch();
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(nlong);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_ccsrmm_runs(void *stream, int64_t kp, int64_t ncols, float ar, float ai, const int32_t *run_ptr,
                      const int32_t *ids, const void *w4, const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch,
                      const int32_t *rowmap, const int32_t *rowptr, const void *packed, const int32_t *longrows,
                      int nlong, int long_thresh) {
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    if (kp == 0 || ncols == 0) return 0;
    IB200_REQUIRE(ncols > 0 && ncols <= 64 && ncols % 2 == 0, "run gather serves an even number of at most 64 columns");
    IB200_REQUIRE(run_ptr && ids && w4 && Xil && Yil && rowmap && rowptr, "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols && xpitch % 2 == 0 && ypitch % 2 == 0, "bad pitch");
    IB200_REQUIRE(((uintptr_t)Xil & 15) == 0 && ((uintptr_t)Yil & 15) == 0, "operands must be 16-byte aligned");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    const int64_t nruns = kp / kRun;
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    if (nlong > 0) {
        IB200_REQUIRE(longrows && packed, "long-row list / packed entries missing");
        const int rc = launch_long_packed(s, runs_pow2_ceil(ncols), nlong, longrows, (int)ncols, alpha, packed, rowptr,
                                          (const c64 *)Xil, xpitch, (c64 *)Yil, ypitch, rowmap);
        if (rc) return rc;
    }
    const int CL = runs_pow2_ceil(ncols / 2);
    const int rpg = 4;
    const int64_t per_cta = (int64_t)(256 / CL) * rpg;
    const int64_t blocks = ceil_div(nruns, per_cta);
    IB200_REQUIRE(blocks < (1LL << 31), "too many runs for one launch");
#define IB200_RUNS_CASE(cl)                                                                                            \
    case cl:                                                                                                           \
        csrmm_runs_kernel<cl><<<(unsigned)blocks, 256, 0, s>>>(nruns, (int)ncols, alpha, run_ptr, ids, (const float4 *)w4, \
                                                               (const c64 *)Xil, (uint32_t)(xpitch * sizeof(c64)), (c64 *)Yil, \
                                                               ypitch, rowmap, rowptr, long_thresh, rpg);             \
        break
    switch (CL) {
        IB200_RUNS_CASE(1); IB200_RUNS_CASE(2); IB200_RUNS_CASE(4); IB200_RUNS_CASE(8); IB200_RUNS_CASE(16); IB200_RUNS_CASE(32);
        default: set_error("internal: no run gather for CL=%d", CL); return IB200_E_UNSUPPORTED;
    }
#undef IB200_RUNS_CASE
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
