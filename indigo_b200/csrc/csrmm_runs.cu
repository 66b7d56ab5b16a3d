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
// Runs longer than `seg_len` entries (the k-space centre of a radial trajectory collects up to ~10^5
// samples in one run) are cut into segments of seg_len entries: every segment is walked by its own lane
// group into a scratch row of partial sums, and a last small kernel adds the partial sums of each such
// run in segment order, so the result does not depend on scheduling.
#include "common.cuh"
#include "pk2.cuh"
#include <cstdlib>

namespace ib200 {

static const int kRun = 4;            // grid points per run = x extent of the adjoint's tiles
static const int kRunPad = 4;         // run lists are padded to multiples of this many entries

struct __align__(8) RunPacked { int32_t col; float w; };

// number of distinct samples in the four rows of run r (rows hold ascending sample indices)
__device__ __forceinline__ int run_merge(const int32_t *__restrict__ rowptr, const RunPacked *__restrict__ ent, int64_t r,
                                         int32_t *ids, float4 *w4, const int32_t *__restrict__ colmap = nullptr) {
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
        if (ids) { ids[n] = colmap ? colmap[m] : m; w4[n] = make_float4(w[0], w[1], w[2], w[3]); }
        ++n;
    }
    return n;
}

__global__ void __launch_bounds__(128) run_count_kernel(int64_t nruns, const int32_t *__restrict__ rowptr,
                                                        const RunPacked *__restrict__ ent, int seg_len,
                                                        int32_t *__restrict__ counts, int *totals) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const int n = run_merge(rowptr, ent, r, nullptr, nullptr);
    const int padded = (n + kRunPad - 1) / kRunPad * kRunPad;
    counts[r] = padded;
    if (padded > seg_len) { atomicAdd(totals, (padded + seg_len - 1) / seg_len); atomicAdd(totals + 1, 1); }
}

// lists of one run; runs longer than seg_len also get their segment descriptors {run, first entry, end
// entry, 0} and one split descriptor {run, first segment, segments, 0}
__global__ void __launch_bounds__(128) run_fill_kernel(int64_t nruns, const int32_t *__restrict__ rowptr,
                                                       const RunPacked *__restrict__ ent, int seg_len,
                                                       const int32_t *__restrict__ run_ptr, int32_t *__restrict__ ids,
                                                       float4 *__restrict__ w4, int4 *__restrict__ seg_desc,
                                                       int4 *__restrict__ split_desc, int *cursors,
                                                       const int32_t *__restrict__ colmap) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const int a = run_ptr[r], b = run_ptr[r + 1];
    const int n = run_merge(rowptr, ent, r, ids + a, w4 + a, colmap);
    for (int q = a + n; q < b; ++q) { ids[q] = n ? ids[a + n - 1] : 0; w4[q] = make_float4(0.f, 0.f, 0.f, 0.f); }
    if (b - a > seg_len) {
        const int k = (b - a + seg_len - 1) / seg_len;
        const int s0 = atomicAdd(cursors, k);
        const int at = atomicAdd(cursors + 1, 1);
        for (int j = 0; j < k; ++j) {
            const int sa = a + j * seg_len;
            seg_desc[s0 + j] = make_int4((int)r, sa, sa + seg_len < b ? sa + seg_len : b, 0);
        }
        split_desc[at] = make_int4((int)r, s0, k, 0);
    }
}

// ---------------------------------------------------------------------------------------------------
// Run entries reach the lane group through a small ring in shared memory that cp.async fills kRingDepth
// batches ahead (a batch = 4 ids + 4 x 4 weights = 80 bytes): the stream of entries costs no registers
// while it is in flight, and by the time a batch is needed it is a shared-memory read instead of an
// L2/DRAM round trip in front of the dependent gathers (ncu: 22 % of the stall samples sat on the first
// use of a register-prefetched batch, profiles/r01_s7_*).
static const int kRingDepth = 4;
static const int kBatchBytes = 16 + 16 * kRunPad;

__device__ __forceinline__ void run_issue_batch(unsigned char *slot, const int32_t *__restrict__ ids,
                                                const float4 *__restrict__ w4, int p, int gl, int CL) {
    for (int ch = gl; ch < 1 + kRunPad; ch += CL) {
        const void *src = ch == 0 ? (const void *)(ids + p) : (const void *)(w4 + p + ch - 1);
        const unsigned d = (unsigned)__cvta_generic_to_shared(slot + 16 * ch);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
    }
}

// acc[t][0..1] += sum over the run entries [a, b) of w_(p0+t) * X[id]  (two coils per lane, PPL of the run's four
// points per lane starting at p0, b - a a multiple of 4).  ring: kRingDepth * kBatchBytes bytes of shared memory
// owned by this lane group of GS lanes (gl = lane within the group); gmask: its lanes.
template <int PPL>
__device__ __forceinline__ void run_walk(int a, int b, const int32_t *__restrict__ ids, const float4 *__restrict__ w4,
                                         const char *xb, uint32_t xpitch_bytes, pk2 (&acc)[PPL][2],
                                         unsigned char *ring, int gl, int GS, int p0, unsigned gmask) {
    const int nb = (b - a) / kRunPad;
#pragma unroll
    for (int k = 0; k < kRingDepth; ++k) {
        if (k < nb) run_issue_batch(ring + k * kBatchBytes, ids, w4, a + k * kRunPad, gl, GS);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    int slot = 0;
    for (int k = 0; k < nb; ++k) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kRingDepth - 1) : "memory");
        __syncwarp(gmask);
        const unsigned char *sl = ring + slot * kBatchBytes;
        const int4 id = *reinterpret_cast<const int4 *>(sl);
        float w[kRunPad][PPL];
#pragma unroll
        for (int u = 0; u < kRunPad; ++u) {
            const unsigned char *wp = sl + 16 + 16 * u + 4 * p0;
            if (PPL == 4) { const float4 v = *reinterpret_cast<const float4 *>(wp); w[u][0] = v.x; w[u][1 % PPL] = v.y; w[u][2 % PPL] = v.z; w[u][3 % PPL] = v.w; }
            else if (PPL == 2) { const float2 v = *reinterpret_cast<const float2 *>(wp); w[u][0] = v.x; w[u][1 % PPL] = v.y; }
            else w[u][0] = *reinterpret_cast<const float *>(wp);
        }
        float4 x[kRunPad];
        const int idv[kRunPad] = {id.x, id.y, id.z, id.w};
#pragma unroll
        for (int u = 0; u < kRunPad; ++u)
            x[u] = __ldg(reinterpret_cast<const float4 *>(xb + (uint64_t)(uint32_t)idv[u] * xpitch_bytes));
        __syncwarp(gmask);                                           // every lane has read the slot
        if (k + kRingDepth < nb) run_issue_batch(ring + slot * kBatchBytes, ids, w4, a + (k + kRingDepth) * kRunPad, gl, GS);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
#pragma unroll
        for (int u = 0; u < kRunPad; ++u) {
            const pk2 x0 = p_make(x[u].x, x[u].y), x1 = p_make(x[u].z, x[u].w);
#pragma unroll
            for (int t = 0; t < PPL; ++t) {
                acc[t][0] = p_fma(p_bc(w[u][t]), x0, acc[t][0]);
                acc[t][1] = p_fma(p_bc(w[u][t]), x1, acc[t][1]);
            }
        }
        if (++slot == kRingDepth) slot = 0;
    }
}

// output rows of this lane's points: needed last, loaded first (one 4/8/16-byte load)
template <int PPL>
__device__ __forceinline__ void run_rows(const int32_t *__restrict__ rowmap, int64_t run, int p0, int (&rv)[PPL]) {
    const int32_t *p = rowmap + kRun * run + p0;
    if (PPL == 4) { const int4 v = __ldg(reinterpret_cast<const int4 *>(p)); rv[0] = v.x; rv[1 % PPL] = v.y; rv[2 % PPL] = v.z; rv[3 % PPL] = v.w; }
    else if (PPL == 2) { const int2 v = __ldg(reinterpret_cast<const int2 *>(p)); rv[0] = v.x; rv[1 % PPL] = v.y; }
    else rv[0] = __ldg(p);
}

template <int PPL>
__device__ __forceinline__ void run_store(const pk2 (&acc)[PPL][2], c64 alpha, const int (&rv)[PPL],
                                          c64 *__restrict__ Yil, int64_t ypitch, int coil) {
#pragma unroll
    for (int t = 0; t < PPL; ++t) {
        const int64_t out = (int64_t)rv[t];
        if (out >= 0) {
            const c64 o0 = cmul(alpha, mk(p_lo(acc[t][0]), p_hi(acc[t][0])));
            const c64 o1 = cmul(alpha, mk(p_lo(acc[t][1]), p_hi(acc[t][1])));
            __stcs(reinterpret_cast<float4 *>(Yil + out * ypitch + coil), make_float4(o0.x, o0.y, o1.x, o1.y));
        }
    }
}

// lane geometry shared by the three kernels: a group of GS = CL*PL lanes serves one run; lane (cl, pl) of
// the group holds coils 2*cl, 2*cl+1 of the points pl*PPL .. pl*PPL + PPL-1 (PPL = 4/PL).  PL > 1 is used
// when there are fewer than 8 coils (coil-sharded operators): the lanes a wide coil vector would fill
// are spent on the four points of the run instead, so that a warp still serves 8 runs, not 32.
template <int CL, int PL>
struct RunLanes {
    static constexpr int GS = CL * PL, GPB = 256 / GS, PPL = kRun / PL;
    int gl, group, coil, p0;
    unsigned gmask;
    unsigned char *ring;
    __device__ __forceinline__ RunLanes(unsigned char *ring_all) {
        gl = (int)(threadIdx.x & (GS - 1)); group = (int)(threadIdx.x / GS);
        coil = 2 * (gl & (CL - 1)); p0 = (gl / CL) * PPL;
        gmask = GS >= 32 ? 0xffffffffu : (((1u << GS) - 1u) << (((int)(threadIdx.x & 31) / GS) * GS));
        ring = ring_all + (size_t)group * (kRingDepth * kBatchBytes);
    }
};

// Yil[rowmap[4*run + i]][c] = alpha * sum_e w_i(e) * Xil[id(e)][c]        (rowmap < 0: nothing stored)
// Runs longer than seg_len are left to the segment kernels below.
template <int CL, int PL>
__global__ void __launch_bounds__(256) csrmm_runs_kernel(int64_t nruns, int C, c64 alpha,
                                                         const int32_t *__restrict__ run_ptr,
                                                         const int32_t *__restrict__ ids, const float4 *__restrict__ w4,
                                                         const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                         c64 *__restrict__ Yil, int64_t ypitch,
                                                         const int32_t *__restrict__ rowmap, int seg_len, int rpg) {
    extern __shared__ __align__(16) unsigned char ring_all[];
    using L = RunLanes<CL, PL>;
    const L ln(ring_all);
    const bool coil_ok = ln.coil < C;
    const char *xb = reinterpret_cast<const char *>(Xil + (coil_ok ? ln.coil : 0));
    const int64_t run0 = (int64_t)blockIdx.x * ((int64_t)L::GPB * rpg) + ln.group;
    {
        // The run lists of this CTA are one contiguous range that is read exactly once: pull it into L2 up
        // front so that the ring's cp.async batches pay an L2 hit instead of a DRAM round trip each.
        const int64_t first = (int64_t)blockIdx.x * ((int64_t)L::GPB * rpg);
        const int64_t last = first + (int64_t)L::GPB * rpg < nruns ? first + (int64_t)L::GPB * rpg : nruns;
        const int e0 = __ldg(run_ptr + first);
        int e1 = __ldg(run_ptr + last);
        if (e1 - e0 > 64 * 1024) e1 = e0 + 64 * 1024;                // split runs: their segments prefetch for themselves
        for (int q = e0 + 8 * (int)threadIdx.x; q < e1; q += 8 * 256)        // 128 bytes of weights, 32 of ids
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w4 + q));
        for (int q = e0 + 32 * (int)threadIdx.x; q < e1; q += 32 * 256)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ids + q));
    }
    for (int i = 0; i < rpg; ++i) {
        const int64_t run = run0 + (int64_t)i * L::GPB;
        if (run >= nruns) break;
        const int a = __ldg(run_ptr + run), b = __ldg(run_ptr + run + 1);
        if (b - a > seg_len) continue;
        int rv[L::PPL];
        run_rows<L::PPL>(rowmap, run, ln.p0, rv);
        pk2 acc[L::PPL][2];
#pragma unroll
        for (int t = 0; t < L::PPL; ++t) { acc[t][0] = p_make(0.f, 0.f); acc[t][1] = p_make(0.f, 0.f); }
        if (a < b) run_walk<L::PPL>(a, b, ids, w4, xb, xpitch_bytes, acc, ln.ring, ln.gl, L::GS, ln.p0, ln.gmask);
        if (coil_ok) run_store<L::PPL>(acc, alpha, rv, Yil, ypitch, ln.coil);
    }
}

// one lane group per segment of a split run: partial sums into scratch[seg][t][coil]
template <int CL, int PL>
__global__ void __launch_bounds__(256) csrmm_runs_seg_kernel(int nseg, int C, const int4 *__restrict__ seg_desc,
                                                             const int32_t *__restrict__ ids,
                                                             const float4 *__restrict__ w4, const c64 *__restrict__ Xil,
                                                             uint32_t xpitch_bytes, c64 *__restrict__ scratch, int cpitch) {
    extern __shared__ __align__(16) unsigned char ring_all[];
    using L = RunLanes<CL, PL>;
    const L ln(ring_all);
    const int sidx = blockIdx.x * L::GPB + ln.group;
    if (sidx >= nseg) return;
    const int4 d = __ldg(seg_desc + sidx);
    const char *xb = reinterpret_cast<const char *>(Xil + (ln.coil < C ? ln.coil : 0));
    pk2 acc[L::PPL][2];
#pragma unroll
    for (int t = 0; t < L::PPL; ++t) { acc[t][0] = p_make(0.f, 0.f); acc[t][1] = p_make(0.f, 0.f); }
    run_walk<L::PPL>(d.y, d.z, ids, w4, xb, xpitch_bytes, acc, ln.ring, ln.gl, L::GS, ln.p0, ln.gmask);
    if (ln.coil < C) {
#pragma unroll
        for (int t = 0; t < L::PPL; ++t)
            *reinterpret_cast<float4 *>(scratch + ((int64_t)sidx * kRun + ln.p0 + t) * cpitch + ln.coil) =
                make_float4(p_lo(acc[t][0]), p_hi(acc[t][0]), p_lo(acc[t][1]), p_hi(acc[t][1]));
    }
}

// one lane group per split run: partial sums added in segment order, then stored like any other run
template <int CL, int PL>
__global__ void __launch_bounds__(256) csrmm_runs_fold_kernel(int nsplit, int C, c64 alpha, const int4 *__restrict__ split_desc,
                                                              const c64 *__restrict__ scratch, int cpitch,
                                                              c64 *__restrict__ Yil, int64_t ypitch,
                                                              const int32_t *__restrict__ rowmap) {
    using L = RunLanes<CL, PL>;
    const L ln(nullptr);
    const int idx = blockIdx.x * L::GPB + ln.group;
    if (idx >= nsplit || ln.coil >= C) return;
    const int4 d = __ldg(split_desc + idx);
    pk2 acc[L::PPL][2];
#pragma unroll
    for (int t = 0; t < L::PPL; ++t) { acc[t][0] = p_make(0.f, 0.f); acc[t][1] = p_make(0.f, 0.f); }
    for (int sgm = d.y; sgm < d.y + d.z; ++sgm) {
#pragma unroll
        for (int t = 0; t < L::PPL; ++t) {
            const float4 v = *reinterpret_cast<const float4 *>(scratch + ((int64_t)sgm * kRun + ln.p0 + t) * cpitch + ln.coil);
            acc[t][0] = p_add(acc[t][0], p_make(v.x, v.y));
            acc[t][1] = p_add(acc[t][1], p_make(v.z, v.w));
        }
    }
    int rv[L::PPL];
    run_rows<L::PPL>(rowmap, d.x, ln.p0, rv);
    run_store<L::PPL>(acc, alpha, rv, Yil, ypitch, ln.coil);
}

int exclusive_scan_public(cudaStream_t s, int64_t n, const int32_t *in, int32_t *out);   // csrmm.cu

static int runs_pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_csr_runs_count(void *stream, int64_t kp, const int32_t *rowptr, const void *packed, int seg_len,
                         int32_t *run_ptr, int64_t *host_entries, int *host_segments, int *host_split) {
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    IB200_REQUIRE(host_entries && host_segments && host_split, "null pointer");
    IB200_REQUIRE(seg_len >= kRunPad && seg_len % kRunPad == 0, "segment length must be a positive multiple of 4");
    *host_entries = 0; *host_segments = 0; *host_split = 0;
    if (kp == 0) return 0;
    IB200_REQUIRE(rowptr && packed && run_ptr, "null pointer");
    const int64_t nruns = kp / kRun;
    cudaStream_t s = as_stream(stream);
    int32_t *counts = nullptr;
    IB200_TRY(cudaMalloc(&counts, (size_t)(nruns + 1) * sizeof(int32_t) + 2 * sizeof(int)));
    int *totals = reinterpret_cast<int *>(counts + nruns + 1);
    cudaMemsetAsync(totals, 0, 2 * sizeof(int), s);
    run_count_kernel<<<(unsigned)ceil_div(nruns, 128), 128, 0, s>>>(nruns, rowptr, (const RunPacked *)packed, seg_len, counts,
                                                                   totals);
    count_launch();
    int rc = exclusive_scan_public(s, nruns, counts, run_ptr);
    int32_t total = 0;
    int ht[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    if (!rc) {
        cudaMemcpyAsync(&total, run_ptr + nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(ht, totals, sizeof(ht), cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    cudaFree(counts);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    IB200_REQUIRE(total >= 0, "run lists exceed 2^31 entries");
    *host_entries = total; *host_segments = ht[0]; *host_split = ht[1];
    return 0;
}

int ib200_csr_runs_fill(void *stream, int64_t kp, const int32_t *rowptr, const void *packed, int seg_len,
                        const int32_t *run_ptr, int32_t *ids, void *w4, int32_t *seg_desc, int32_t *split_desc,
                        const int32_t *colmap) {
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    IB200_REQUIRE(seg_len >= kRunPad && seg_len % kRunPad == 0, "segment length must be a positive multiple of 4");
    if (kp == 0) return 0;
    IB200_REQUIRE(rowptr && packed && run_ptr && ids && w4 && seg_desc && split_desc, "null pointer");
    IB200_REQUIRE(((uintptr_t)ids & 15) == 0 && ((uintptr_t)w4 & 15) == 0 && ((uintptr_t)seg_desc & 15) == 0 &&
                  ((uintptr_t)split_desc & 15) == 0, "run arrays must be 16-byte aligned");
    const int64_t nruns = kp / kRun;
    cudaStream_t s = as_stream(stream);
    int *cursors = nullptr;
    IB200_TRY(cudaMalloc(&cursors, 2 * sizeof(int)));
    cudaMemsetAsync(cursors, 0, 2 * sizeof(int), s);
    run_fill_kernel<<<(unsigned)ceil_div(nruns, 128), 128, 0, s>>>(nruns, rowptr, (const RunPacked *)packed, seg_len, run_ptr,
                                                                  ids, (float4 *)w4, (int4 *)seg_desc, (int4 *)split_desc,
                                                                  cursors, colmap);
    count_launch();
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(cursors);
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_ccsrmm_runs(void *stream, int64_t kp, int64_t ncols, float ar, float ai, const int32_t *run_ptr,
                      const int32_t *ids, const void *w4, const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch,
                      const int32_t *rowmap, int seg_len, const int32_t *seg_desc, int nseg, const int32_t *split_desc,
                      int nsplit, void *scratch) {
    IB200_RANGE("ib200_ccsrmm_runs");
    IB200_REQUIRE(kp >= 0 && kp % kRun == 0 && kp < (1LL << 31), "row count must be a multiple of 4");
    if (kp == 0 || ncols == 0) return 0;
    IB200_REQUIRE(ncols > 0 && ncols <= 64 && ncols % 2 == 0, "run gather serves an even number of at most 64 columns");
    IB200_REQUIRE(run_ptr && ids && w4 && Xil && Yil && rowmap, "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols && xpitch % 2 == 0 && ypitch % 2 == 0, "bad pitch");
    IB200_REQUIRE(((uintptr_t)Xil & 15) == 0 && ((uintptr_t)Yil & 15) == 0, "operands must be 16-byte aligned");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    IB200_REQUIRE(seg_len >= kRunPad && nseg >= 0 && nsplit >= 0, "bad segment arguments");
    IB200_REQUIRE(nseg == 0 || (seg_desc && split_desc && scratch && ((uintptr_t)scratch & 15) == 0), "segment arrays missing");
    const int64_t nruns = kp / kRun;
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    const int CL = runs_pow2_ceil(ncols / 2);
    int PL = CL >= 4 ? 1 : kRun / CL;                               // lanes of a group split the run's points when coils are few
    if (const char *e = getenv("IB200_RUNS_PL")) {                   // tuning knob (tools/): points-split factor 1, 2 or 4
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4) && CL * v <= 32 && (CL < 4 || v == 1)) PL = v;
    }
    const int rpg = 4;
    const int GPB = 256 / (CL * PL);
    const int64_t blocks = ceil_div(nruns, (int64_t)GPB * rpg);
    IB200_REQUIRE(blocks < (1LL << 31), "too many runs for one launch");
    const int cpitch = 2 * CL;                                       // scratch row: 2*CL complex words per point
    const size_t ring_bytes = (size_t)GPB * kRingDepth * kBatchBytes;
    const uint32_t pb = (uint32_t)(xpitch * sizeof(c64));
#define IB200_RUNS_CASE(cl, pl)                                                                                        \
    case (cl) * 8 + (pl):                                                                                              \
        if (ring_bytes > 48 * 1024) {                                                                                  \
            IB200_TRY(cudaFuncSetAttribute(csrmm_runs_kernel<cl, pl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes)); \
            IB200_TRY(cudaFuncSetAttribute(csrmm_runs_seg_kernel<cl, pl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes)); \
        }                                                                                                              \
        if (nseg > 0) {                                                                                                \
            /* few, long, latency-bound lane groups: on the side stream, sharing the SMs with the main kernel */      \
            rc = side_stream_begin(s, &side);                                                                          \
            if (rc) return rc;                                                                                         \
            csrmm_runs_seg_kernel<cl, pl><<<(unsigned)ceil_div(nseg, GPB), 256, ring_bytes, side>>>(nseg, (int)ncols, (const int4 *)seg_desc, ids, \
                                                                                   (const float4 *)w4, (const c64 *)Xil, pb,  \
                                                                                   (c64 *)scratch, cpitch);           \
            count_launch();                                                                                            \
        }                                                                                                              \
        csrmm_runs_kernel<cl, pl><<<(unsigned)blocks, 256, ring_bytes, s>>>(nruns, (int)ncols, alpha, run_ptr, ids, (const float4 *)w4, \
                                                               (const c64 *)Xil, pb, (c64 *)Yil, ypitch, rowmap, seg_len, rpg); \
        if (nseg > 0) { rc = side_stream_end(s); if (rc) return rc; }                                                  \
        if (nsplit > 0) {                                                                                              \
            count_launch();                                                                                            \
            csrmm_runs_fold_kernel<cl, pl><<<(unsigned)ceil_div(nsplit, GPB), 256, 0, s>>>(nsplit, (int)ncols, alpha,  \
                                                                                      (const int4 *)split_desc, (const c64 *)scratch, \
                                                                                      cpitch, (c64 *)Yil, ypitch, rowmap); \
        }                                                                                                              \
        break
    int rc = 0;
    cudaStream_t side = nullptr;
    switch (CL * 8 + PL) {
        IB200_RUNS_CASE(1, 4); IB200_RUNS_CASE(1, 2); IB200_RUNS_CASE(1, 1); IB200_RUNS_CASE(2, 2); IB200_RUNS_CASE(2, 1);
        IB200_RUNS_CASE(4, 1); IB200_RUNS_CASE(8, 1); IB200_RUNS_CASE(16, 1);
        IB200_RUNS_CASE(32, 1);
        default: set_error("internal: no run gather for CL=%d PL=%d", CL, PL); return IB200_E_UNSUPPORTED;
    }
#undef IB200_RUNS_CASE
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
