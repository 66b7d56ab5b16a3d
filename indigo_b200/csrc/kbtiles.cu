// Adjoint gridding on tile blocks: ccsrmm(G', adjoint) of the fused SENSE recipe for operators with FEW
// coils (the per-GPU shard of a coil-sharded operator: 2 coils at 8 GPUs, 4 at 4 GPUs).
//
// Reference being replaced: the adjoint product of the gridding matrix that indigo/interp.py:19-80 emits
// (SpMatrix._eval with forward=False, operators.py:322-332 -> Backend.ccsrmm, backend.py:560-596).
//
// Why a second formulation next to csrmm_runs.cu: the x-run lists cost 20 bytes and one gather per
// (sample, 4 grid points) pair -- 54 pairs per sample, 7.4 GB at cfg3 -- and that stream does not shrink
// when the coils are sharded, so at 2 coils per GPU it is the whole cost of the step (2.8 of 5.6 ms,
// profiles/r02_s1_coils2.md).  A Kaiser-Bessel footprint of 5 taps per axis meets only 2 x 2 x 2 tiles of
// 4 x 4 x 4 grid points, and inside a tile its weights are still an outer product.  One entry
//     (sample, wx[4], wy[4], wz[4])              52 bytes, serves the 64 points of a tile
// replaces 6.75 run entries (135 bytes, 6.75 gathers): 8 entries per sample, 2.8 GB at cfg3.  The price is
// arithmetic on zero weights (64 multiply-adds per entry for 15.6 useful ones on average), which is cheap
// exactly when the coils are few.  Entries are built straight from the separable records of the forward
// gather (kbgrid.cu), so forward and adjoint use bit-identical weight factors; no stored adjoint is read.
//
// Layout: entries of a tile are consecutive, ordered by sample, in batches of four:
//     batch = ids[4] | wx[4][4] | wy[4][4] | wz[4][4]         208 bytes, 16-byte aligned
// (lists padded with zero-weight entries to whole batches).  A work item is (tile, batch range, slot): tiles
// with more than seg_batches batches (k-space centre of a radial trajectory) are cut into several items
// that write partial sums into scratch[slot]; a fold kernel adds them in segment order, so the result does
// not depend on scheduling.  Tiles without entries but with rows inside the support windows get one empty
// item: every grid point inside the windows is overwritten on every apply.
#include "common.cuh"
#include "kb.cuh"
#include "pk2.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>

namespace ib200 {

int exclusive_scan_public(cudaStream_t s, int64_t n, const int32_t *in, int32_t *out);   // csrmm.cu

static const int kTE = 4;                            // tile edge
static const int kTV = kTE * kTE * kTE;              // points per tile
static const int kTB = 4;                            // entries per batch
static const int kTBatchBytes = 16 + 3 * 16 * kTB;   // 208
static const int kTRing = 4;                         // batches in flight per lane group

// ---- setup: which tiles a sample meets, and with which weights -------------------------------------------
struct AxisTiles { int n; int tc[3]; float w[3][kTE]; };

// taps j0, j0+1, ... (mod N) of one axis, sorted into the tiles they fall into; tiles whose four weights are all
// zero (sixth tap of an on-grid sample) are dropped
__device__ __forceinline__ void axis_tiles(int j0, int cnt, int N, const float *w6, AxisTiles &a) {
    a.n = 0;
    int j = j0;
    for (int t = 0; t < kKbTaps; ++t) {
        if (t < cnt) {
            const int tc = j / kTE, pos = j % kTE;
            int k = -1;
            for (int q = 0; q < a.n; ++q) if (a.tc[q] == tc) k = q;
            if (k < 0 && a.n < 3) { k = a.n++; a.tc[k] = tc; for (int i = 0; i < kTE; ++i) a.w[k][i] = 0.f; }
            if (k >= 0) a.w[k][pos] += w6[t];
            if (++j >= N) j = 0;
        }
    }
    int m = 0;
    for (int q = 0; q < a.n; ++q) {
        bool any = false;
        for (int i = 0; i < kTE; ++i) any = any || a.w[q][i] != 0.f;
        if (any) { if (m != q) { a.tc[m] = a.tc[q]; for (int i = 0; i < kTE; ++i) a.w[m][i] = a.w[q][i]; } ++m; }
    }
    a.n = m;
}

__device__ __forceinline__ void record_tiles(const KbRecord &q, int n0, int n1, int n2, AxisTiles &ax, AxisTiles &ay,
                                             AxisTiles &az) {
    axis_tiles(q.ix0, q.ntaps & 255, n0, q.wx, ax);
    axis_tiles(q.iy0, (q.ntaps >> 8) & 255, n1, q.wy, ay);
    axis_tiles(q.iz0, (q.ntaps >> 16) & 255, n2, q.wz, az);
}

// work items of a tile with nb batches: segments of seg_batches batches, lengthened for the densest tiles so that no
// tile has more than kTMaxSeg of them (the fold of a tile's partial sums is a serial chain over its segments)
static const int kTMaxSeg = 48;
__host__ __device__ __forceinline__ int tile_seg_len(int nb, int seg_batches) {
    const int cap = (nb + kTMaxSeg - 1) / kTMaxSeg;
    return cap > seg_batches ? cap : seg_batches;
}

// pairs per record (rcnt) and per tile (tcnt)
__global__ void __launch_bounds__(128) tile_pairs_count_kernel(int64_t m, const KbRecord *__restrict__ rec, int n0, int n1,
                                                               int n2, int nt0, int nt1, int32_t *__restrict__ rcnt,
                                                               int32_t *tcnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const KbRecord q = rec[r];
    AxisTiles ax, ay, az;
    record_tiles(q, n0, n1, n2, ax, ay, az);
    if (rcnt) rcnt[r] = ax.n * ay.n * az.n;
    if (tcnt)
        for (int k = 0; k < az.n; ++k)
            for (int j = 0; j < ay.n; ++j)
                for (int i = 0; i < ax.n; ++i) atomicAdd(tcnt + ((int64_t)az.tc[k] * nt1 + ay.tc[j]) * nt0 + ax.tc[i], 1);
}

__global__ void __launch_bounds__(128) tile_pairs_emit_kernel(int64_t m, const KbRecord *__restrict__ rec, int n0, int n1,
                                                              int n2, int nt0, int nt1, const int32_t *__restrict__ rpos,
                                                              int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const KbRecord q = rec[r];
    AxisTiles ax, ay, az;
    record_tiles(q, n0, n1, n2, ax, ay, az);
    int64_t at = rpos[r];
    for (int k = 0; k < az.n; ++k)
        for (int j = 0; j < ay.n; ++j)
            for (int i = 0; i < ax.n; ++i) {
                keys[at] = (int32_t)(((int64_t)az.tc[k] * nt1 + ay.tc[j]) * nt0 + ax.tc[i]);
                vals[at] = (int32_t)r;
                ++at;
            }
}

// batches and work items of every tile; totals[0] += split tiles, totals[1] += segments of split tiles
__global__ void __launch_bounds__(256) tile_sizes_kernel(int64_t ntiles, const int32_t *__restrict__ tcnt,
                                                         const int32_t *__restrict__ rowmap, int seg_batches,
                                                         int32_t *__restrict__ nbatch, int32_t *__restrict__ nwork,
                                                         int *totals) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int nb = (tcnt[t] + kTB - 1) / kTB;
    bool rows = false;
    const int4 *rm = reinterpret_cast<const int4 *>(rowmap + t * kTV);
    for (int i = 0; i < kTV / 4; ++i) { const int4 v = __ldg(rm + i); rows = rows || v.x >= 0 || v.y >= 0 || v.z >= 0 || v.w >= 0; }
    int nw = 0;
    if (rows) { const int sl = tile_seg_len(nb, seg_batches); nw = (nb + sl - 1) / sl; if (nw < 1) nw = 1; }
    nbatch[t] = rows ? nb : 0;
    nwork[t] = nw;
    if (nw > 1) { atomicAdd(totals, 1); atomicAdd(totals + 1, nw); }
}

// first sorted pair of every tile that has pairs
__global__ void __launch_bounds__(256) tile_starts_kernel(int64_t npairs, const int32_t *__restrict__ keys,
                                                          int32_t *__restrict__ tstart) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    if (i == 0 || keys[i] != keys[i - 1]) tstart[keys[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(128) tile_fill_kernel(int64_t npairs, const int32_t *__restrict__ keys,
                                                        const int32_t *__restrict__ vals,
                                                        const int32_t *__restrict__ tstart,
                                                        const KbRecord *__restrict__ rec, int n0, int n1, int n2, int nt0,
                                                        int nt1, const int32_t *__restrict__ bptr,
                                                        unsigned char *__restrict__ ent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int t = keys[i];
    if (bptr[t + 1] == bptr[t]) return;                              // tile without rows inside the windows
    const KbRecord q = rec[vals[i]];
    AxisTiles ax, ay, az;
    record_tiles(q, n0, n1, n2, ax, ay, az);
    const int tx = t % nt0, ty = (t / nt0) % nt1, tz = t / (nt0 * nt1);
    float wx[kTE] = {0.f, 0.f, 0.f, 0.f}, wy[kTE] = {0.f, 0.f, 0.f, 0.f}, wz[kTE] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ax.n; ++k) if (ax.tc[k] == tx) for (int p = 0; p < kTE; ++p) wx[p] = ax.w[k][p];
    for (int k = 0; k < ay.n; ++k) if (ay.tc[k] == ty) for (int p = 0; p < kTE; ++p) wy[p] = ay.w[k][p];
    for (int k = 0; k < az.n; ++k) if (az.tc[k] == tz) for (int p = 0; p < kTE; ++p) wz[p] = az.w[k][p];
    const int kk = (int)(i - tstart[t]);
    unsigned char *b = ent + ((int64_t)bptr[t] + kk / kTB) * kTBatchBytes;
    const int u = kk % kTB;
    reinterpret_cast<int32_t *>(b)[u] = q.out;
    *reinterpret_cast<float4 *>(b + 16 + 16 * u) = make_float4(wx[0], wx[1], wx[2], wx[3]);
    *reinterpret_cast<float4 *>(b + 16 + 16 * kTB + 16 * u) = make_float4(wy[0], wy[1], wy[2], wy[3]);
    *reinterpret_cast<float4 *>(b + 16 + 32 * kTB + 16 * u) = make_float4(wz[0], wz[1], wz[2], wz[3]);
    if (i + 1 == npairs || keys[i + 1] != t) {                       // last entry of the tile pads its batch
        for (int v = u + 1; v < kTB; ++v) {
            reinterpret_cast<int32_t *>(b)[v] = q.out;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(b + 16 + 16 * v) = z;
            *reinterpret_cast<float4 *>(b + 16 + 16 * kTB + 16 * v) = z;
            *reinterpret_cast<float4 *>(b + 16 + 32 * kTB + 16 * v) = z;
        }
    }
}

// work items {tile, first batch, end batch, scratch slot or -1} and split descriptors {tile, first slot, items, 0}
__global__ void __launch_bounds__(256) tile_work_kernel(int64_t ntiles, const int32_t *__restrict__ bptr,
                                                        const int32_t *__restrict__ wptr, int seg_batches,
                                                        int4 *__restrict__ work, int4 *__restrict__ split, int *cursors) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const int w0 = wptr[t], nw = wptr[t + 1] - w0;
    if (nw == 0) return;
    const int b0 = bptr[t], b1 = bptr[t + 1];
    if (nw == 1) { work[w0] = make_int4((int)t, b0, b1, -1); return; }
    const int s0 = atomicAdd(cursors, nw);
    const int at = atomicAdd(cursors + 1, 1);
    const int sl = tile_seg_len(b1 - b0, seg_batches);
    for (int j = 0; j < nw; ++j) {
        const int a = b0 + j * sl;
        work[w0 + j] = make_int4((int)t, a, a + sl < b1 ? a + sl : b1, s0 + j);
    }
    split[at] = make_int4((int)t, s0, nw, 0);
}

// ---- apply ---------------------------------------------------------------------------------------------
// Lane geometry: a group of GS = CL * PLN lanes serves one work item.  Lane (cl, pl) holds coils 2cl, 2cl+1
// of the x-rows (y = pl % 4, z = z0 .. z0 + ZPL-1) of the tile, z0 = (pl / 4) * ZPL, ZPL = 16 / PLN:
// 4 * ZPL points, two packed accumulators each.  Few point lanes (PLN = 4: 16 points per lane) amortise the
// shared-memory reads of an entry over 32 packed multiply-adds: the LSU issues one instruction per 1.8 cycles
// per SM, and with 16 point lanes it, not the arithmetic, bounds the kernel (ncu, profiles/r02_s6_tiles.md).
//
// Everything that comes from global memory arrives through cp.async into a ring of kTRing slots per lane
// group, a slot = one batch of entries (208 bytes) + the k-space rows of its four samples (4 x 16*CL bytes):
//   iteration k:  wait until batch k's rows and batch k+2's entries have landed
//                 issue the gathers of batch k+2 (ids are in its slot)             |  one commit group
//                 consume batch k from shared memory                               |  per iteration
//                 re-fill slot k with the entries of batch k+4                     |
// so no register is held across a global-memory latency and no lane ever waits on one.  The loop count is the
// maximum over the groups of a warp (idle groups skip the body), which keeps every barrier a full-warp one.
template <int CL>
struct TileRing {
    static constexpr int XB = 16 * CL;                         // bytes of one sample's coils
    static constexpr int SLOT = kTBatchBytes + kTB * XB;
    static constexpr int BYTES = kTRing * SLOT + 16;           // +16: consecutive groups start 20 banks apart
};

template <int CL, int PLN>
struct TileLanes {
    static constexpr int GS = CL * PLN, GPB = 256 / GS, ZPL = 16 / PLN;
    int gl, group, cl, y, z0;
    unsigned char *ring;
    __device__ __forceinline__ TileLanes(unsigned char *ring_all) {
        gl = (int)(threadIdx.x & (GS - 1)); group = (int)(threadIdx.x / GS);
        cl = gl & (CL - 1);
        const int pl = gl / CL;
        y = pl & 3; z0 = (pl >> 2) * ZPL;
        ring = ring_all + (size_t)group * TileRing<CL>::BYTES;
    }
};

__device__ __forceinline__ void tile_cp16(unsigned char *dst, const void *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}

template <int GS>
__device__ __forceinline__ void tile_issue_stream(unsigned char *slot, const unsigned char *src, int gl) {
#pragma unroll
    for (int c0 = 0; c0 < kTBatchBytes / 16; c0 += GS) {
        const int c = c0 + gl;
        if (c < kTBatchBytes / 16) tile_cp16(slot + 16 * c, src + 16 * c);
    }
}

// k-space rows of the four samples of the batch in `slot` (its ids have landed): 16 bytes per lane
template <int CL, int GS>
__device__ __forceinline__ void tile_issue_gather(unsigned char *slot, const char *xb, uint32_t xpitch_bytes, int C, int gl) {
#pragma unroll
    for (int c0 = 0; c0 < kTB * CL; c0 += GS) {
        const int c = c0 + gl;
        const int u = c / CL, part = c % CL;
        if (c < kTB * CL && 2 * part < C) {
            const uint32_t id = reinterpret_cast<const uint32_t *>(slot)[u];
            tile_cp16(slot + kTBatchBytes + u * (16 * CL) + 16 * part, xb + (uint64_t)id * xpitch_bytes + 16 * part);
        }
    }
}

template <int CL, int ZPL>
__device__ __forceinline__ void tile_consume(const unsigned char *sl, int cl, int y, int z0, pk2 (&acc)[ZPL][kTE][2]) {
#pragma unroll
    for (int u = 0; u < kTB; ++u) {
        const float4 wx = *reinterpret_cast<const float4 *>(sl + 16 + 16 * u);
        const float wy = *reinterpret_cast<const float *>(sl + 16 + 16 * kTB + 16 * u + 4 * y);
        float wz[ZPL];
        const unsigned char *zp = sl + 16 + 32 * kTB + 16 * u + 4 * z0;
        if (ZPL == 4) { const float4 v = *reinterpret_cast<const float4 *>(zp); wz[0] = v.x; wz[1 % ZPL] = v.y; wz[2 % ZPL] = v.z; wz[3 % ZPL] = v.w; }
        else if (ZPL == 2) { const float2 v = *reinterpret_cast<const float2 *>(zp); wz[0] = v.x; wz[1 % ZPL] = v.y; }
        else wz[0] = *reinterpret_cast<const float *>(zp);
        const float4 xv = *reinterpret_cast<const float4 *>(sl + kTBatchBytes + u * (16 * CL) + 16 * cl);
        const float wxv[kTE] = {wx.x, wx.y, wx.z, wx.w};
        const pk2 x0 = p_make(xv.x, xv.y), x1 = p_make(xv.z, xv.w);
        if (ZPL >= 4) {
            // 16 points per lane: scale the sample by the four x weights once, then one multiply-add per point
            pk2 t0[kTE], t1[kTE];
#pragma unroll
            for (int j = 0; j < kTE; ++j) { t0[j] = p_scale(wxv[j], x0); t1[j] = p_scale(wxv[j], x1); }
#pragma unroll
            for (int q = 0; q < ZPL; ++q) {
                const float wzy = wz[q] * wy;
#pragma unroll
                for (int j = 0; j < kTE; ++j) {
                    acc[q][j][0] = p_fma(p_bc(wzy), t0[j], acc[q][j][0]);
                    acc[q][j][1] = p_fma(p_bc(wzy), t1[j], acc[q][j][1]);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < ZPL; ++q) {
                const float wzy = wz[q] * wy;
#pragma unroll
                for (int j = 0; j < kTE; ++j) {
                    const float w = wzy * wxv[j];
                    acc[q][j][0] = p_fma(p_bc(w), x0, acc[q][j][0]);
                    acc[q][j][1] = p_fma(p_bc(w), x1, acc[q][j][1]);
                }
            }
        }
    }
}

// acc[q][j][0..1] += sum over the entries of batches [b0, b0 + nb) of wz[z0+q] * wy[y] * wx[j] * X[id]; nbmax = the
// largest nb among the groups of this warp
template <int CL, int PLN>
__device__ __forceinline__ void tile_walk(int b0, int nb, int nbmax, const unsigned char *__restrict__ ent, const char *xb,
                                          uint32_t xpitch_bytes, int C, pk2 (&acc)[16 / PLN][kTE][2],
                                          const TileLanes<CL, PLN> &ln) {
    constexpr int GS = CL * PLN, SLOT = TileRing<CL>::SLOT;
    const unsigned char *src = ent + (int64_t)b0 * kTBatchBytes;
    unsigned char *ring = ln.ring;
#pragma unroll
    for (int k = 0; k < kTRing; ++k)
        if (k < nb) tile_issue_stream<GS>(ring + k * SLOT, src + (int64_t)k * kTBatchBytes, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (0 < nb) tile_issue_gather<CL, GS>(ring, xb, xpitch_bytes, C, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (1 < nb) tile_issue_gather<CL, GS>(ring + SLOT, xb, xpitch_bytes, C, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    int slot = 0;
    for (int k = 0; k < nbmax; ++k) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        unsigned char *sl = ring + slot * SLOT;
        if (k + 2 < nb) tile_issue_gather<CL, GS>(ring + ((slot + 2) & (kTRing - 1)) * SLOT, xb, xpitch_bytes, C, ln.gl);
        if (k < nb) tile_consume<CL, 16 / PLN>(sl, ln.cl, ln.y, ln.z0, acc);
        __syncwarp();                                                // every lane has read the slot
        if (k + kTRing < nb) tile_issue_stream<GS>(sl, src + (int64_t)(k + kTRing) * kTBatchBytes, ln.gl);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        slot = (slot + 1) & (kTRing - 1);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
static_assert((kTRing & (kTRing - 1)) == 0, "ring depth must be a power of two");

template <int ZPL>
__device__ __forceinline__ void tile_store(const pk2 (&acc)[ZPL][kTE][2], c64 alpha, int tile, int y, int z0,
                                           const int32_t *__restrict__ rowmap, c64 *__restrict__ Yil, int64_t ypitch,
                                           int coil) {
#pragma unroll
    for (int q = 0; q < ZPL; ++q) {
        const int4 rv4 = __ldg(reinterpret_cast<const int4 *>(rowmap + (int64_t)tile * kTV + ((z0 + q) * kTE + y) * kTE));
        const int rv[kTE] = {rv4.x, rv4.y, rv4.z, rv4.w};
#pragma unroll
        for (int j = 0; j < kTE; ++j) {
            const int64_t out = (int64_t)rv[j];
            if (out >= 0) {
                const c64 o0 = cmul(alpha, mk(p_lo(acc[q][j][0]), p_hi(acc[q][j][0])));
                const c64 o1 = cmul(alpha, mk(p_lo(acc[q][j][1]), p_hi(acc[q][j][1])));
                __stcs(reinterpret_cast<float4 *>(Yil + out * ypitch + coil), make_float4(o0.x, o0.y, o1.x, o1.y));
            }
        }
    }
}

// Yil[rowmap[64*tile + p]][c] = alpha * sum_e wz_e[pz] wy_e[py] wx_e[px] * Xil[id_e][c]   (items with slot < 0)
// scratch[slot][p][c]         =         the same sum over the item's batches                  (items of split tiles)
// for the C <= 2*CL columns starting at Xil / Yil / scratch (the host loops over chunks of 16 columns)
template <int CL, int PLN>
__global__ void __launch_bounds__(256) kb_tiles_kernel(int nwork, int C, c64 alpha, const int4 *__restrict__ work,
                                                       const unsigned char *__restrict__ ent, const c64 *__restrict__ Xil,
                                                       uint32_t xpitch_bytes, c64 *__restrict__ Yil, int64_t ypitch,
                                                       const int32_t *__restrict__ rowmap, c64 *__restrict__ scratch,
                                                       int cpitch) {
    typedef TileLanes<CL, PLN> L;
    extern __shared__ __align__(16) unsigned char tile_ring[];
    const L ln(tile_ring);
    const int idx = blockIdx.x * L::GPB + ln.group;
    const bool live = idx < nwork;
    int4 d = make_int4(0, 0, 0, -1);
    if (live) d = __ldg(work + idx);
    const int nb = d.z - d.y;
    const int nbmax = __reduce_max_sync(0xffffffffu, nb);
    const int coil = 2 * ln.cl;
    pk2 acc[L::ZPL][kTE][2];
#pragma unroll
    for (int q = 0; q < L::ZPL; ++q)
#pragma unroll
        for (int j = 0; j < kTE; ++j) { acc[q][j][0] = p_make(0.f, 0.f); acc[q][j][1] = p_make(0.f, 0.f); }
    tile_walk<CL, PLN>(d.y, nb, nbmax, ent, reinterpret_cast<const char *>(Xil), xpitch_bytes, C, acc, ln);
    if (!live || coil >= C) return;
    if (d.w < 0) {
        tile_store<L::ZPL>(acc, alpha, d.x, ln.y, ln.z0, rowmap, Yil, ypitch, coil);
    } else {
#pragma unroll
        for (int q = 0; q < L::ZPL; ++q)
#pragma unroll
            for (int j = 0; j < kTE; ++j) {
                const int p = ((ln.z0 + q) * kTE + ln.y) * kTE + j;
                *reinterpret_cast<float4 *>(scratch + ((int64_t)d.w * kTV + p) * cpitch + coil) =
                    make_float4(p_lo(acc[q][j][0]), p_hi(acc[q][j][0]), p_lo(acc[q][j][1]), p_hi(acc[q][j][1]));
            }
    }
}

// split tiles: one CTA per tile; thread (sl, point, coil lane) adds the partial sums of segments sl, sl + NS, ...
// (four independent chains for memory-level parallelism), the NS segment lanes are then added in order through
// shared memory: a fixed summation tree, independent of scheduling.  With 8 coil lanes (16 columns) the 512
// (point, coil lane) pairs of a tile take two rounds of the 256 threads.
template <int CL>
__global__ void __launch_bounds__(256) kb_tiles_fold_kernel(int nsplit, int C, c64 alpha, const int4 *__restrict__ split,
                                                            const c64 *__restrict__ scratch, int cpitch,
                                                            c64 *__restrict__ Yil, int64_t ypitch,
                                                            const int32_t *__restrict__ rowmap) {
    constexpr int PAIRS = kTV * CL;
    constexpr int NS = PAIRS >= 256 ? 1 : 256 / PAIRS, ROUNDS = PAIRS > 256 ? PAIRS / 256 : 1;
    __shared__ float4 part[NS > 1 ? (NS - 1) * PAIRS : 1];
    const int4 d = __ldg(split + blockIdx.x);
    const int64_t seg_stride = (int64_t)kTV * cpitch;
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int item = (int)threadIdx.x % (PAIRS < 256 ? PAIRS : 256) + 256 * rd;
        const int sl = PAIRS < 256 ? (int)threadIdx.x / PAIRS : 0;
        const int cl = item % CL, p = item / CL;
        const int coil = 2 * cl;
        pk2 a0[4], a1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0[u] = p_make(0.f, 0.f); a1[u] = p_make(0.f, 0.f); }
        const c64 *base = scratch + ((int64_t)d.y * kTV + p) * cpitch + coil;
        for (int g = sl; g < d.z; g += 4 * NS) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int sg = g + u * NS;
                if (sg < d.z) {
                    const float4 v = __ldcs(reinterpret_cast<const float4 *>(base + sg * seg_stride));
                    a0[u] = p_add(a0[u], p_make(v.x, v.y));
                    a1[u] = p_add(a1[u], p_make(v.z, v.w));
                }
            }
        }
        pk2 s0 = p_add(p_add(a0[0], a0[1]), p_add(a0[2], a0[3])), s1 = p_add(p_add(a1[0], a1[1]), p_add(a1[2], a1[3]));
        if (NS > 1) {
            if (sl > 0) part[(sl - 1) * PAIRS + item] = make_float4(p_lo(s0), p_hi(s0), p_lo(s1), p_hi(s1));
            __syncthreads();
            if (sl == 0) {
#pragma unroll
                for (int q = 1; q < NS; ++q) {
                    const float4 v = part[(q - 1) * PAIRS + item];
                    s0 = p_add(s0, p_make(v.x, v.y)); s1 = p_add(s1, p_make(v.z, v.w));
                }
            }
        }
        if (sl == 0 && coil < C) {
            const int64_t out = (int64_t)__ldg(rowmap + (int64_t)d.x * kTV + p);
            if (out >= 0) {
                const c64 o0 = cmul(alpha, mk(p_lo(s0), p_hi(s0))), o1 = cmul(alpha, mk(p_lo(s1), p_hi(s1)));
                __stcs(reinterpret_cast<float4 *>(Yil + out * ypitch + coil), make_float4(o0.x, o0.y, o1.x, o1.y));
            }
        }
    }
}

static int tiles_pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

static bool tiles_grid_ok(const int64_t grid[3], int64_t *ntiles, int nt[3]) {
    if (!grid) return false;
    int64_t n = 1;
    for (int d = 0; d < 3; ++d) {
        if (grid[d] <= 0 || grid[d] >= (1LL << 30)) return false;
        nt[d] = (int)ceil_div(grid[d], kTE);
        n *= nt[d];
    }
    *ntiles = n;
    return n * kTV < (1LL << 31);
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_kb_tiles_batch_bytes(void) { return kTBatchBytes; }

int ib200_kb_tiles_count(void *stream, int64_t m, const void *records, const int64_t grid[3], const int32_t *rowmap,
                         int seg_batches, int32_t *bptr, int32_t *wptr, int64_t *host_totals) {
    int64_t ntiles = 0;
    int nt[3];
    IB200_REQUIRE(tiles_grid_ok(grid, &ntiles, nt), "bad grid");
    IB200_REQUIRE(m >= 0 && m < (1LL << 31) && seg_batches >= 1 && host_totals, "bad arguments");
    IB200_REQUIRE(rowmap && bptr && wptr && (records || m == 0), "null pointer");
    for (int i = 0; i < 5; ++i) host_totals[i] = 0;
    cudaStream_t s = as_stream(stream);
    int32_t *buf = nullptr;
    IB200_TRY(cudaMalloc(&buf, (size_t)(3 * ntiles + 4) * sizeof(int32_t)));
    int32_t *tcnt = buf, *nbatch = buf + ntiles, *nwork = buf + 2 * ntiles;
    int *totals = reinterpret_cast<int *>(buf + 3 * ntiles);
    cudaMemsetAsync(buf, 0, (size_t)(3 * ntiles + 4) * sizeof(int32_t), s);
    if (m > 0) {
        tile_pairs_count_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, (const KbRecord *)records, (int)grid[0], (int)grid[1],
                                                                          (int)grid[2], nt[0], nt[1], nullptr, tcnt);
        count_launch();
    }
    tile_sizes_kernel<<<(unsigned)ceil_div(ntiles, 256), 256, 0, s>>>(ntiles, tcnt, rowmap, seg_batches, nbatch, nwork, totals);
    count_launch();
    int rc = exclusive_scan_public(s, ntiles, nbatch, bptr);
    if (!rc) rc = exclusive_scan_public(s, ntiles, nwork, wptr);
    int32_t tb = 0, tw = 0;
    int ht[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    if (!rc) {
        cudaMemcpyAsync(&tb, bptr + ntiles, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(&tw, wptr + ntiles, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(ht, totals, sizeof(ht), cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    cudaFree(buf);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    IB200_REQUIRE(tb >= 0 && tw >= 0, "tile lists exceed 2^31 batches");
    host_totals[0] = ntiles; host_totals[1] = tb; host_totals[2] = tw; host_totals[3] = ht[0]; host_totals[4] = ht[1];
    return 0;
}

int ib200_kb_tiles_fill(void *stream, int64_t m, const void *records, const int64_t grid[3], int seg_batches,
                        const int32_t *bptr, const int32_t *wptr, void *entries, int32_t *work, int32_t *split) {
    int64_t ntiles = 0;
    int nt[3];
    IB200_REQUIRE(tiles_grid_ok(grid, &ntiles, nt), "bad grid");
    IB200_REQUIRE(m >= 0 && m < (1LL << 31) && seg_batches >= 1, "bad arguments");
    IB200_REQUIRE(bptr && wptr && entries && work && split && (records || m == 0), "null pointer");
    IB200_REQUIRE(((uintptr_t)entries & 15) == 0 && ((uintptr_t)work & 15) == 0 && ((uintptr_t)split & 15) == 0,
                  "tile arrays must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const KbRecord *rec = (const KbRecord *)records;
    const int n0 = (int)grid[0], n1 = (int)grid[1], n2 = (int)grid[2];
    int rc = 0;
    cudaError_t e = cudaSuccess;
    int32_t *rpos = nullptr, *pairs = nullptr, *tstart = nullptr;
    void *tmp = nullptr;
    int *cursors = nullptr;
    int32_t npairs = 0;
    if (m > 0) {
        e = cudaMalloc(&rpos, (size_t)(m + 1) * sizeof(int32_t));
        if (e == cudaSuccess) {
            tile_pairs_count_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, rec, n0, n1, n2, nt[0], nt[1], rpos, nullptr);
            count_launch();
            rc = exclusive_scan_public(s, m, rpos, rpos);
            if (!rc) {
                cudaMemcpyAsync(&npairs, rpos + m, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                e = cudaStreamSynchronize(s);
            }
        }
        if (!rc && e == cudaSuccess && npairs < 0) { set_error("tile lists exceed 2^31 entries"); rc = IB200_E_INVALID; }
    }
    if (!rc && e == cudaSuccess && npairs > 0) {
        e = cudaMalloc(&pairs, (size_t)npairs * 4 * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc(&tstart, (size_t)ntiles * sizeof(int32_t));
        if (e == cudaSuccess) {
            int32_t *ka = pairs, *kb = pairs + npairs, *va = pairs + 2 * (int64_t)npairs, *vb = pairs + 3 * (int64_t)npairs;
            tile_pairs_emit_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, rec, n0, n1, n2, nt[0], nt[1], rpos, ka, va);
            count_launch();
            int end_bit = 1; while ((1LL << end_bit) < ntiles) ++end_bit;
            cub::DoubleBuffer<int32_t> dk(ka, kb), dv(va, vb);
            size_t tmp_bytes = 0;
            e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)npairs, 0, end_bit, s);
            if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
            if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)npairs, 0, end_bit, s);
            if (e == cudaSuccess) {
                count_launch(end_bit / 8 + 2);
                tile_starts_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, s>>>(npairs, dk.Current(), tstart);
                tile_fill_kernel<<<(unsigned)ceil_div(npairs, 128), 128, 0, s>>>(npairs, dk.Current(), dv.Current(), tstart, rec, n0,
                                                                             n1, n2, nt[0], nt[1], bptr,
                                                                             (unsigned char *)entries);
                count_launch(2);
            }
        }
    }
    if (!rc && e == cudaSuccess) e = cudaMalloc(&cursors, 2 * sizeof(int));
    if (!rc && e == cudaSuccess) {
        cudaMemsetAsync(cursors, 0, 2 * sizeof(int), s);
        tile_work_kernel<<<(unsigned)ceil_div(ntiles, 256), 256, 0, s>>>(ntiles, bptr, wptr, seg_batches, (int4 *)work, (int4 *)split,
                                                                       cursors);
        count_launch();
        e = cudaStreamSynchronize(s);
    }
    cudaFree(rpos); cudaFree(pairs); cudaFree(tstart); cudaFree(tmp); cudaFree(cursors);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_kb_tiles_apply(void *stream, int64_t ncols, float ar, float ai, int nwork, const int32_t *work, const void *entries,
                         const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch, const int32_t *rowmap, int nsplit,
                         const int32_t *split, void *scratch, int lanes) {
    IB200_RANGE("ib200_kb_tiles_apply");
    IB200_REQUIRE(nwork >= 0 && nsplit >= 0 && ncols >= 0, "bad arguments");
    if (nwork == 0 || ncols == 0) return 0;
    IB200_REQUIRE(ncols <= 64 && ncols % 2 == 0, "tile gather serves an even number of at most 64 columns");
    IB200_REQUIRE(work && entries && Xil && Yil && rowmap, "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols && xpitch % 2 == 0 && ypitch % 2 == 0, "bad pitch");
    IB200_REQUIRE(((uintptr_t)Xil & 15) == 0 && ((uintptr_t)Yil & 15) == 0 && ((uintptr_t)entries & 15) == 0,
                  "operands must be 16-byte aligned");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    IB200_REQUIRE(nsplit == 0 || (split && scratch && ((uintptr_t)scratch & 15) == 0), "split arrays missing");
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    int want = lanes == 4 || lanes == 8 || lanes == 16 ? lanes : 4;  // point lanes per item
    if (const char *e = getenv("IB200_TILES_PLN")) {                 // tuning knob (tools/)
        const int v = atoi(e);
        if (v == 4 || v == 8 || v == 16) want = v;
    }
    const uint32_t pb = (uint32_t)(xpitch * sizeof(c64));
    // chunks of at most 16 columns: 8 coil lanes x 4 point lanes fill a warp
    for (int64_t c0 = 0; c0 < ncols; c0 += 16) {
        const int cc = (int)(ncols - c0 < 16 ? ncols - c0 : 16);
        const int CL = tiles_pow2_ceil(cc / 2);
        int PLN = want;
        while (CL * PLN > 32) PLN /= 2;
        const int GS = CL * PLN, GPB = 256 / GS;
        const int cpitch = 2 * CL;
        const c64 *X = (const c64 *)Xil + c0;
        c64 *Y = (c64 *)Yil + c0;
#define IB200_TILES_CASE(cl, pln)                                                                                      \
    case (cl) * 32 + (pln): {                                                                                          \
        const size_t ring_bytes = (size_t)GPB * TileRing<cl>::BYTES;                                                   \
        if (ring_bytes > 48 * 1024)                                                                                    \
            IB200_TRY(cudaFuncSetAttribute(kb_tiles_kernel<cl, pln>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes)); \
        kb_tiles_kernel<cl, pln><<<(unsigned)ceil_div(nwork, GPB), 256, ring_bytes, s>>>(nwork, cc, alpha, (const int4 *)work, \
                                                                (const unsigned char *)entries, X, pb, Y, ypitch, rowmap, \
                                                                (c64 *)scratch, cpitch);                               \
        if (nsplit > 0) {                                                                                              \
            count_launch();                                                                                            \
            kb_tiles_fold_kernel<cl><<<(unsigned)nsplit, 256, 0, s>>>(nsplit, cc, alpha, (const int4 *)split,          \
                                                                      (const c64 *)scratch, cpitch, Y, ypitch, rowmap); \
        }                                                                                                              \
    } break
        switch (CL * 32 + PLN) {
            IB200_TILES_CASE(1, 4); IB200_TILES_CASE(1, 8); IB200_TILES_CASE(1, 16);
            IB200_TILES_CASE(2, 4); IB200_TILES_CASE(2, 8); IB200_TILES_CASE(2, 16);
            IB200_TILES_CASE(4, 4); IB200_TILES_CASE(4, 8);
            IB200_TILES_CASE(8, 4);
            default: set_error("internal: no tile gather for CL=%d PLN=%d", CL, PLN); return IB200_E_UNSUPPORTED;
        }
#undef IB200_TILES_CASE
        IB200_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
