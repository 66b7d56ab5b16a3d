// Kernels of the block form of the adjoint gridding (see kbblocks.cu for the formulation and the setup side).
// Templates only: kbblocks_s*.cu instantiate one block shape each so that the shapes compile in parallel.
#pragma once
#include "common.cuh"
#include "pk2.cuh"
#include <cstdlib>

namespace ib200 {

static const int kTE = 4;                            // tile edge = x extent of every block
static const int kTV = kTE * kTE * kTE;              // points per tile
static const int kTB = 4;                            // entries per batch
static const int kTRing = 4;                         // batches in flight per lane group

// block shape and the layout of a batch that follows from it
struct BlockShape {
    int by, bz;
    __host__ __device__ int nsub() const { return (kTE / by) * (kTE / bz); }       // blocks per tile
    __host__ __device__ int points() const { return kTE * by * bz; }
    __host__ __device__ bool has_wy() const { return by * bz > 1; }                // else everything is folded into wx
    __host__ __device__ bool has_wz() const { return bz > 1; }                     // else wz is folded into wy
    __host__ __device__ int wy_off() const { return 16 + 16 * kTB; }
    __host__ __device__ int wz_off() const { return wy_off() + (has_wy() ? 4 * by * kTB : 0); }
    __host__ __device__ int batch_bytes() const { return wz_off() + (has_wz() ? 4 * bz * kTB : 0); }
};
template <int BY, int BZ>
struct BlockLayout {
    static constexpr int NSUB = (kTE / BY) * (kTE / BZ), PV = kTE * BY * BZ;
    static constexpr bool HAS_WY = BY * BZ > 1, HAS_WZ = BZ > 1;
    static constexpr int WY_OFF = 16 + 16 * kTB, WZ_OFF = WY_OFF + (HAS_WY ? 4 * BY * kTB : 0);
    static constexpr int BATCH = WZ_OFF + (HAS_WZ ? 4 * BZ * kTB : 0);
};

// first row (in the tile-major row order) of the x-row rr = zz*BY + yy of a block
__host__ __device__ __forceinline__ int64_t block_row(int block, int rr, int by, int bz) {
    const int py = kTE / by, nsub = py * (kTE / bz);
    const int tile = block / nsub, sub = block % nsub;
    const int y = (sub % py) * by + rr % by, z = (sub / py) * bz + rr / by;
    return (int64_t)tile * kTV + (z * kTE + y) * kTE;
}

// ---- apply ---------------------------------------------------------------------------------------------
// Lane geometry: a group of GS = CL * PLN lanes serves one work item.  Lane (cl, pl) holds coils 2cl, 2cl+1 of
// RPL = BY*BZ / PLN consecutive x-rows rr = zz*BY + yy of the block: NY = min(RPL, BY) values of yy starting at
// y0, NZ = RPL / NY values of zz starting at z0; 4 * RPL points, two packed accumulators each.  Few point lanes
// amortise the shared-memory reads of an entry over more multiply-adds: the LSU issues one instruction per 1.8
// cycles per SM, and with many point lanes it, not the arithmetic, bounds the kernel (profiles/r02_blocks.md).
//
// Everything that comes from global memory arrives through cp.async into a ring of kTRing slots per lane
// group, a slot = one batch of entries + the k-space rows of its four samples (4 x 16*CL bytes):
//   iteration k:  wait until batch k's rows and batch k+2's entries have landed
//                 issue the gathers of batch k+2 (ids are in its slot)             |  one commit group
//                 consume batch k from shared memory                               |  per iteration
//                 re-fill slot k with the entries of batch k+4                     |
// so no register is held across a global-memory latency and no lane ever waits on one.  The loop count is the
// maximum over the groups of a warp (idle groups skip the body), which keeps every barrier a full-warp one.
template <int CL, int BY, int BZ>
struct BlockRing {
    static constexpr int XB = 16 * CL;                         // bytes of one sample's coils
    static constexpr int SLOT = BlockLayout<BY, BZ>::BATCH + kTB * XB;
    static constexpr int BYTES = kTRing * SLOT + ((kTRing * SLOT / 4) % 32 == 20 ? 0 : 16);   // groups start >= 4 banks apart
};

template <int CL, int PLN, int BY, int BZ>
struct BlockLanes {
    static constexpr int GS = CL * PLN, GPB = 256 / GS, RPL = BY * BZ / PLN;
    static constexpr int NY = RPL < BY ? RPL : BY, NZ = RPL / NY;
    static_assert(RPL >= 1 && RPL * PLN == BY * BZ && GS <= 32, "bad lane geometry");
    int gl, group, cl, y0, z0;
    unsigned char *ring;
    __device__ __forceinline__ BlockLanes(unsigned char *ring_all) {
        gl = (int)(threadIdx.x & (GS - 1)); group = (int)(threadIdx.x / GS);
        cl = gl & (CL - 1);
        const int r0 = (gl / CL) * RPL;
        y0 = r0 % BY; z0 = r0 / BY;
        ring = ring_all + (size_t)group * BlockRing<CL, BY, BZ>::BYTES;
    }
};

__device__ __forceinline__ void block_cp16(unsigned char *dst, const void *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}

template <int GS, int BATCH>
__device__ __forceinline__ void block_issue_stream(unsigned char *slot, const unsigned char *src, int gl) {
#pragma unroll
    for (int c0 = 0; c0 < BATCH / 16; c0 += GS) {
        const int c = c0 + gl;
        if (c < BATCH / 16) block_cp16(slot + 16 * c, src + 16 * c);
    }
}

// k-space rows of the four samples of the batch in `slot` (its ids have landed): 16 bytes per lane
template <int CL, int GS, int BATCH>
__device__ __forceinline__ void block_issue_gather(unsigned char *slot, const char *xb, uint32_t xpitch_bytes, int C, int gl) {
#pragma unroll
    for (int c0 = 0; c0 < kTB * CL; c0 += GS) {
        const int c = c0 + gl;
        const int u = c / CL, part = c % CL;
        if (c < kTB * CL && 2 * part < C) {
            const uint32_t id = reinterpret_cast<const uint32_t *>(slot)[u];
            block_cp16(slot + BATCH + u * (16 * CL) + 16 * part, xb + (uint64_t)id * xpitch_bytes + 16 * part);
        }
    }
}

template <int NV>
__device__ __forceinline__ void block_lds(const unsigned char *p, float (&w)[NV]) {
    if (NV == 4) { const float4 v = *reinterpret_cast<const float4 *>(p); w[0] = v.x; w[1 % NV] = v.y; w[2 % NV] = v.z; w[3 % NV] = v.w; }
    else if (NV == 2) { const float2 v = *reinterpret_cast<const float2 *>(p); w[0] = v.x; w[1 % NV] = v.y; }
    else w[0] = *reinterpret_cast<const float *>(p);
}

template <int CL, int PLN, int BY, int BZ>
__device__ __forceinline__ void block_consume(const unsigned char *sl, const BlockLanes<CL, PLN, BY, BZ> &ln,
                                              pk2 (&acc)[BY * BZ / PLN][kTE][2]) {
    typedef BlockLayout<BY, BZ> LY;
    typedef BlockLanes<CL, PLN, BY, BZ> L;
#pragma unroll
    for (int u = 0; u < kTB; ++u) {
        const float4 wx = *reinterpret_cast<const float4 *>(sl + 16 + 16 * u);
        float wy[L::NY], wz[L::NZ];
        if (LY::HAS_WY) block_lds<L::NY>(sl + LY::WY_OFF + 4 * (BY * u + ln.y0), wy); else wy[0] = 1.f;
        if (LY::HAS_WZ) block_lds<L::NZ>(sl + LY::WZ_OFF + 4 * (BZ * u + ln.z0), wz); else wz[0] = 1.f;
        const float4 xv = *reinterpret_cast<const float4 *>(sl + LY::BATCH + u * (16 * CL) + 16 * ln.cl);
        const float wxv[kTE] = {wx.x, wx.y, wx.z, wx.w};
        const pk2 x0 = p_make(xv.x, xv.y), x1 = p_make(xv.z, xv.w);
        if (L::RPL >= 4) {
            // many rows per lane: scale the sample by the four x weights once, then one multiply-add per point
            pk2 t0[kTE], t1[kTE];
#pragma unroll
            for (int j = 0; j < kTE; ++j) { t0[j] = p_scale(wxv[j], x0); t1[j] = p_scale(wxv[j], x1); }
#pragma unroll
            for (int iz = 0; iz < L::NZ; ++iz)
#pragma unroll
                for (int iy = 0; iy < L::NY; ++iy) {
                    const float wr = !LY::HAS_WY ? 1.f : (LY::HAS_WZ ? wz[iz] * wy[iy] : wy[iy]);
#pragma unroll
                    for (int j = 0; j < kTE; ++j) {
                        acc[iz * L::NY + iy][j][0] = p_fma(p_bc(wr), t0[j], acc[iz * L::NY + iy][j][0]);
                        acc[iz * L::NY + iy][j][1] = p_fma(p_bc(wr), t1[j], acc[iz * L::NY + iy][j][1]);
                    }
                }
        } else {
#pragma unroll
            for (int iz = 0; iz < L::NZ; ++iz)
#pragma unroll
                for (int iy = 0; iy < L::NY; ++iy) {
                    const float wr = !LY::HAS_WY ? 1.f : (LY::HAS_WZ ? wz[iz] * wy[iy] : wy[iy]);
#pragma unroll
                    for (int j = 0; j < kTE; ++j) {
                        const float w = LY::HAS_WY ? wr * wxv[j] : wxv[j];
                        acc[iz * L::NY + iy][j][0] = p_fma(p_bc(w), x0, acc[iz * L::NY + iy][j][0]);
                        acc[iz * L::NY + iy][j][1] = p_fma(p_bc(w), x1, acc[iz * L::NY + iy][j][1]);
                    }
                }
        }
    }
}

// acc += sum over the entries of batches [b0, b0 + nb) of wz wy wx * X[id]; nbmax = the largest nb among the groups
// of this warp
template <int CL, int PLN, int BY, int BZ>
__device__ __forceinline__ void block_walk(int b0, int nb, int nbmax, const unsigned char *__restrict__ ent, const char *xb,
                                           uint32_t xpitch_bytes, int C, pk2 (&acc)[BY * BZ / PLN][kTE][2],
                                           const BlockLanes<CL, PLN, BY, BZ> &ln) {
    constexpr int GS = CL * PLN, SLOT = BlockRing<CL, BY, BZ>::SLOT, BATCH = BlockLayout<BY, BZ>::BATCH;
    const unsigned char *src = ent + (int64_t)b0 * BATCH;
    unsigned char *ring = ln.ring;
#pragma unroll
    for (int k = 0; k < kTRing; ++k)
        if (k < nb) block_issue_stream<GS, BATCH>(ring + k * SLOT, src + (int64_t)k * BATCH, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (0 < nb) block_issue_gather<CL, GS, BATCH>(ring, xb, xpitch_bytes, C, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (1 < nb) block_issue_gather<CL, GS, BATCH>(ring + SLOT, xb, xpitch_bytes, C, ln.gl);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    int slot = 0;
    for (int k = 0; k < nbmax; ++k) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        unsigned char *sl = ring + slot * SLOT;
        if (k + 2 < nb) block_issue_gather<CL, GS, BATCH>(ring + ((slot + 2) & (kTRing - 1)) * SLOT, xb, xpitch_bytes, C, ln.gl);
        if (k < nb) block_consume<CL, PLN, BY, BZ>(sl, ln, acc);
        __syncwarp();                                                // every lane has read the slot
        if (k + kTRing < nb) block_issue_stream<GS, BATCH>(sl, src + (int64_t)(k + kTRing) * BATCH, ln.gl);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        slot = (slot + 1) & (kTRing - 1);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
static_assert((kTRing & (kTRing - 1)) == 0, "ring depth must be a power of two");

// Yil[rowmap[row(block, p)]][c] = alpha * sum_e wz_e wy_e wx_e * Xil[id_e][c]          (items with slot < 0)
// scratch[slot][p][c]           =         the same sum over the item's batches            (items of split blocks)
// for the C <= 2*CL columns starting at Xil / Yil / scratch (the host loops over chunks of 16 columns)
template <int CL, int PLN, int BY, int BZ>
__global__ void __launch_bounds__(256) kb_blocks_kernel(int nwork, int C, c64 alpha, const int4 *__restrict__ work,
                                                        const unsigned char *__restrict__ ent, const c64 *__restrict__ Xil,
                                                        uint32_t xpitch_bytes, c64 *__restrict__ Yil, int64_t ypitch,
                                                        const int32_t *__restrict__ rowmap, c64 *__restrict__ scratch,
                                                        int cpitch) {
    typedef BlockLanes<CL, PLN, BY, BZ> L;
    extern __shared__ __align__(16) unsigned char block_ring[];
    const L ln(block_ring);
    const int idx = blockIdx.x * L::GPB + ln.group;
    const bool live = idx < nwork;
    int4 d = make_int4(0, 0, 0, -1);
    if (live) d = __ldg(work + idx);
    const int nb = d.z - d.y;
    const int nbmax = __reduce_max_sync(0xffffffffu, nb);
    const int coil = 2 * ln.cl;
    pk2 acc[L::RPL][kTE][2];
#pragma unroll
    for (int q = 0; q < L::RPL; ++q)
#pragma unroll
        for (int j = 0; j < kTE; ++j) { acc[q][j][0] = p_make(0.f, 0.f); acc[q][j][1] = p_make(0.f, 0.f); }
    block_walk<CL, PLN, BY, BZ>(d.y, nb, nbmax, ent, reinterpret_cast<const char *>(Xil), xpitch_bytes, C, acc, ln);
    if (!live || coil >= C) return;
#pragma unroll
    for (int iz = 0; iz < L::NZ; ++iz)
#pragma unroll
        for (int iy = 0; iy < L::NY; ++iy) {
            const int q = iz * L::NY + iy, rr = (ln.z0 + iz) * BY + ln.y0 + iy;
            if (d.w < 0) {
                const int4 rv4 = __ldg(reinterpret_cast<const int4 *>(rowmap + block_row(d.x, rr, BY, BZ)));
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
            } else {
#pragma unroll
                for (int j = 0; j < kTE; ++j)
                    *reinterpret_cast<float4 *>(scratch + ((int64_t)d.w * (kTE * BY * BZ) + rr * kTE + j) * cpitch + coil) =
                        make_float4(p_lo(acc[q][j][0]), p_hi(acc[q][j][0]), p_lo(acc[q][j][1]), p_hi(acc[q][j][1]));
            }
        }
}

// split blocks: one CTA per block; thread (sl, point, coil lane) adds the partial sums of segments sl, sl + NS, ...
// (four independent chains for memory-level parallelism), the NS segment lanes are then added in order through
// shared memory: a fixed summation tree, independent of scheduling.  More than 256 (point, coil lane) pairs take
// several rounds of the 256 threads.
template <int CL, int BY, int BZ>
__global__ void __launch_bounds__(256) kb_blocks_fold_kernel(int nsplit, int C, c64 alpha, const int4 *__restrict__ split,
                                                             const c64 *__restrict__ scratch, int cpitch,
                                                             c64 *__restrict__ Yil, int64_t ypitch,
                                                             const int32_t *__restrict__ rowmap) {
    constexpr int PV = kTE * BY * BZ, PAIRS = PV * CL;
    constexpr int NS = PAIRS >= 256 ? 1 : 256 / PAIRS, ROUNDS = PAIRS > 256 ? PAIRS / 256 : 1;
    __shared__ float4 part[NS > 1 ? (NS - 1) * PAIRS : 1];
    const int4 d = __ldg(split + blockIdx.x);
    const int64_t seg_stride = (int64_t)PV * cpitch;
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
        const int item = (int)threadIdx.x % (PAIRS < 256 ? PAIRS : 256) + 256 * rd;
        const int sl = PAIRS < 256 ? (int)threadIdx.x / PAIRS : 0;
        const int cl = item % CL, p = item / CL;
        const int coil = 2 * cl;
        pk2 a0[4], a1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { a0[u] = p_make(0.f, 0.f); a1[u] = p_make(0.f, 0.f); }
        const c64 *base = scratch + ((int64_t)d.y * PV + p) * cpitch + coil;
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
            const int64_t out = (int64_t)__ldg(rowmap + block_row(d.x, p / kTE, BY, BZ) + p % kTE);
            if (out >= 0) {
                const c64 o0 = cmul(alpha, mk(p_lo(s0), p_hi(s0))), o1 = cmul(alpha, mk(p_lo(s1), p_hi(s1)));
                __stcs(reinterpret_cast<float4 *>(Yil + out * ypitch + coil), make_float4(o0.x, o0.y, o1.x, o1.y));
            }
        }
    }
}

template <int CL, int PLN, int BY, int BZ>
int launch_blocks(cudaStream_t s, int nwork, int cc, c64 alpha, const int32_t *work, const void *entries, const c64 *X,
                         uint32_t pb, c64 *Y, int64_t ypitch, const int32_t *rowmap, int nsplit, const int32_t *split,
                         void *scratch) {
    typedef BlockLanes<CL, PLN, BY, BZ> L;
    const size_t ring_bytes = (size_t)L::GPB * BlockRing<CL, BY, BZ>::BYTES;
    const int cpitch = 2 * CL;
    if (ring_bytes > 48 * 1024)
        IB200_TRY(cudaFuncSetAttribute(kb_blocks_kernel<CL, PLN, BY, BZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    kb_blocks_kernel<CL, PLN, BY, BZ><<<(unsigned)ceil_div(nwork, L::GPB), 256, ring_bytes, s>>>(
        nwork, cc, alpha, (const int4 *)work, (const unsigned char *)entries, X, pb, Y, ypitch, rowmap, (c64 *)scratch, cpitch);
    IB200_LAUNCH_CHECK();
    if (nsplit > 0) {
        kb_blocks_fold_kernel<CL, BY, BZ><<<(unsigned)nsplit, 256, 0, s>>>(nsplit, cc, alpha, (const int4 *)split,
                                                                          (const c64 *)scratch, cpitch, Y, ypitch, rowmap);
        IB200_LAUNCH_CHECK();
    }
    return 0;
}

// point lanes: the wanted number, halved until a group fits a warp, capped by the rows of the block
template <int CL, int BY, int BZ>
int dispatch_lanes(int want, cudaStream_t s, int nwork, int cc, c64 alpha, const int32_t *work, const void *entries,
                          const c64 *X, uint32_t pb, c64 *Y, int64_t ypitch, const int32_t *rowmap, int nsplit,
                          const int32_t *split, void *scratch) {
    constexpr int ROWS = BY * BZ;
    int pln = want;
    while (pln > ROWS || CL * pln > 32) pln /= 2;
    if (pln < 1) pln = 1;
    if (ROWS == 16 && pln < 4) pln = 4;                              // 64 points per lane would not fit the registers
#define IB200_LANES_CASE(p)                                                                                            \
    case p:                                                                                                            \
        if constexpr (p <= ROWS && CL * p <= 32 && ROWS / p <= 4)                                                      \
            return launch_blocks<CL, p, BY, BZ>(s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch); \
        break
    switch (pln) { IB200_LANES_CASE(1); IB200_LANES_CASE(2); IB200_LANES_CASE(4); IB200_LANES_CASE(8); }
#undef IB200_LANES_CASE
    set_error("internal: no block gather for CL=%d lanes=%d block 4x%dx%d", CL, pln, BY, BZ);
    return IB200_E_UNSUPPORTED;
}

template <int BY, int BZ>
int dispatch_coils(int CL, int want, cudaStream_t s, int nwork, int cc, c64 alpha, const int32_t *work,
                          const void *entries, const c64 *X, uint32_t pb, c64 *Y, int64_t ypitch, const int32_t *rowmap,
                          int nsplit, const int32_t *split, void *scratch) {
    switch (CL) {
        case 1: return dispatch_lanes<1, BY, BZ>(want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 2: return dispatch_lanes<2, BY, BZ>(want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 4: return dispatch_lanes<4, BY, BZ>(want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 8: return dispatch_lanes<8, BY, BZ>(want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
    }
    set_error("internal: no block gather for CL=%d", CL);
    return IB200_E_UNSUPPORTED;
}


#define IB200_BLOCKS_INSTANTIATE(BY_, BZ_)                                                                              \
    template int dispatch_coils<BY_, BZ_>(int, int, cudaStream_t, int, int, c64, const int32_t *, const void *, const c64 *, \
                                          uint32_t, c64 *, int64_t, const int32_t *, int, const int32_t *, void *)

}  // namespace ib200
