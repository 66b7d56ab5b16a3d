// Tensor-core form of the block adjoint gridding for whole 4 x 4 x 4 tiles (kbblocks.cu, shape (4,4)).
//
// With many coils the tile formulation spends its time on fp32 multiply-adds with zero weights (64 per entry and
// coil for 15.6 useful ones): 6.7 ms at 16 coils against 4.9 ms for 4 x 2 x 2 blocks.  The same sum is a small
// real matrix product per tile,
//     D[64 points][2C floats] = W^T[64][E entries] * X[E][2C],      W[e][p] = wz_e[pz] wy_e[py] wx_e[px],
// whose zero-weight work costs nothing on the tensor cores.  A warp owns one work item and walks its entries eight
// at a time (one k-step of mma.sync.m16n8k8, two batches of the entry stream):
//     A fragments  the lane's 16 weights of the k-step are formed in registers from the separable factors in
//                  shared memory (8 loads, 20 multiplies) -- the 64 x 8 weight matrix is never stored anywhere;
//     B fragments  the gathered k-space rows, read from the ring with a row stride that spreads the four rows of a
//                  fragment over the banks;
//     3xTF32       every operand is split into hi + lo (three instructions, gemm.cu) and the product is
//                  hi*hi + hi*lo + lo*hi with fp32 accumulation: 2^-21 relative error per product, the same
//                  accuracy class as the fp32 FFMA path (tests compare both with the oracle at 1e-5).
// Entry stream, gathers and work items are those of kbblocks.cuh: entries and k-space rows arrive through cp.async
// into a ring per warp, two k-steps ahead; block lists are built with an even number of batches.
#include "kbblocks.cuh"

namespace ib200 {

__device__ __forceinline__ void mma_split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_m16n8k8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ring slot = one k-step: two batches of the 4x4x4 entry stream + the k-space rows of its eight samples
template <int NT>
struct MmaRing {
    static constexpr int C = 4 * NT;                            // columns (coils) of this chunk
    static constexpr int ROWB = 8 * C;                          // bytes of one k-space row
    static constexpr int XS = ROWB % 64 == 32 ? ROWB : ROWB + 32;   // row stride: rows tig = 0..3 land 8 banks apart
    static constexpr int BATCH = BlockLayout<4, 4>::BATCH;      // 208
    static constexpr int STEP = 2 * BATCH;                      // bytes of the entry stream per k-step
    static constexpr int SLOT = STEP + 2 * kTB * XS;
    static constexpr int BYTES = kTRing * SLOT;
};

template <int NT>
__device__ __forceinline__ void mma_issue_stream(unsigned char *slot, const unsigned char *src, int lane) {
    if (lane < MmaRing<NT>::STEP / 16) block_cp16(slot + 16 * lane, src + 16 * lane);
}

template <int NT>
__device__ __forceinline__ void mma_issue_gather(unsigned char *slot, const char *xb, uint32_t xpitch_bytes, int lane) {
    typedef MmaRing<NT> R;
    constexpr int CPR = R::ROWB / 16, TOTAL = 2 * kTB * CPR;    // 16-byte chunks per row / per k-step
#pragma unroll
    for (int c0 = 0; c0 < TOTAL; c0 += 32) {
        const int c = c0 + lane;
        if (c < TOTAL) {
            const int e = c / CPR, part = c % CPR;
            const uint32_t id = reinterpret_cast<const uint32_t *>(slot + (e / kTB) * R::BATCH)[e % kTB];
            block_cp16(slot + R::STEP + e * R::XS + 16 * part, xb + (uint64_t)id * xpitch_bytes + 16 * part);
        }
    }
}

// one k-step: acc[mt][nt] += W^T(16 points of plane mt, 8 entries) * X(8 entries, 8 columns of n-tile nt)
template <int NT>
__device__ __forceinline__ void mma_consume(const unsigned char *sl, int g, int tig, float (&acc)[4][NT][4]) {
    typedef MmaRing<NT> R;
    // B fragments: rows tig (batch 0) and tig + 4 (batch 1), column 8 nt + g
    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const float x0 = *reinterpret_cast<const float *>(sl + R::STEP + tig * R::XS + 4 * (8 * nt + g));
        const float x1 = *reinterpret_cast<const float *>(sl + R::STEP + (tig + kTB) * R::XS + 4 * (8 * nt + g));
        mma_split_tf32(x0, bh[nt][0], bl[nt][0]);
        mma_split_tf32(x1, bh[nt][1], bl[nt][1]);
    }
    // separable factors of this lane's weights: entries tig / tig + 4, points x = g % 4, y = g / 4 and y + 2
    const int x = g & 3, y = g >> 2;
    float p[2][2];
    float4 wz[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const unsigned char *b = sl + h * R::BATCH;
        const float wx = *reinterpret_cast<const float *>(b + 16 + 16 * tig + 4 * x);
        const float wy0 = *reinterpret_cast<const float *>(b + 16 + 16 * kTB + 16 * tig + 4 * y);
        const float wy1 = *reinterpret_cast<const float *>(b + 16 + 16 * kTB + 16 * tig + 4 * (y + 2));
        wz[h] = *reinterpret_cast<const float4 *>(b + 16 + 32 * kTB + 16 * tig);
        p[h][0] = wy0 * wx; p[h][1] = wy1 * wx;
    }
    const float wzv[2][4] = {{wz[0].x, wz[0].y, wz[0].z, wz[0].w}, {wz[1].x, wz[1].y, wz[1].z, wz[1].w}};
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        // a0: (row g, entry tig)  a1: (row g + 8, entry tig)  a2: (row g, entry tig + 4)  a3: (row g + 8, entry tig + 4)
        uint32_t ah[4], al[4];
        mma_split_tf32(wzv[0][mt] * p[0][0], ah[0], al[0]);
        mma_split_tf32(wzv[0][mt] * p[0][1], ah[1], al[1]);
        mma_split_tf32(wzv[1][mt] * p[1][0], ah[2], al[2]);
        mma_split_tf32(wzv[1][mt] * p[1][1], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            mma_m16n8k8(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
            mma_m16n8k8(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
            mma_m16n8k8(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
        }
    }
}

// same contract as kb_blocks_kernel<CL, PLN, 4, 4> for C = 4 NT columns; one warp per work item, work items hold an
// even number of batches
template <int NT>
__global__ void __launch_bounds__(256, 2) kb_tiles_mma_kernel(int nwork, c64 alpha, const int4 *__restrict__ work,
                                                              const unsigned char *__restrict__ ent,
                                                              const c64 *__restrict__ Xil, uint32_t xpitch_bytes,
                                                              c64 *__restrict__ Yil, int64_t ypitch,
                                                              const int32_t *__restrict__ rowmap, c64 *__restrict__ scratch,
                                                              int cpitch) {
    typedef MmaRing<NT> R;
    extern __shared__ __align__(16) unsigned char mma_ring[];
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const int g = lane >> 2, tig = lane & 3;
    const int idx = blockIdx.x * 8 + warp;
    if (idx >= nwork) return;                                        // whole warp
    const int4 d = __ldg(work + idx);
    unsigned char *ring = mma_ring + (size_t)warp * R::BYTES;
    const char *xb = reinterpret_cast<const char *>(Xil);
    const int ns = (d.z - d.y) / 2;                                  // k-steps
    const unsigned char *src = ent + (int64_t)d.y * R::BATCH;
    float acc[4][NT][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int k = 0; k < kTRing; ++k)
        if (k < ns) mma_issue_stream<NT>(ring + k * R::SLOT, src + (int64_t)k * R::STEP, lane);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncwarp();
    if (0 < ns) mma_issue_gather<NT>(ring, xb, xpitch_bytes, lane);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (1 < ns) mma_issue_gather<NT>(ring + R::SLOT, xb, xpitch_bytes, lane);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    int slot = 0;
    for (int k = 0; k < ns; ++k) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        unsigned char *sl = ring + slot * R::SLOT;
        if (k + 2 < ns) mma_issue_gather<NT>(ring + ((slot + 2) & (kTRing - 1)) * R::SLOT, xb, xpitch_bytes, lane);
        mma_consume<NT>(sl, g, tig, acc);
        __syncwarp();                                                // every lane has read the slot
        if (k + kTRing < ns) mma_issue_stream<NT>(sl, src + (int64_t)(k + kTRing) * R::STEP, lane);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        slot = (slot + 1) & (kTRing - 1);
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    // c0, c1: point (plane mt, row g), coil 4 nt + tig;  c2, c3: row g + 8
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = mt * 16 + g + 8 * h;
            if (d.w < 0) {
                const int64_t out = (int64_t)__ldg(rowmap + (int64_t)d.x * kTV + p);
                if (out >= 0) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        __stcs(Yil + out * ypitch + 4 * nt + tig, cmul(alpha, mk(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1])));
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    scratch[((int64_t)d.w * kTV + p) * cpitch + 4 * nt + tig] = mk(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
            }
        }
}

template <int NT>
static int launch_tiles_mma(cudaStream_t s, int nwork, c64 alpha, const int32_t *work, const void *entries, const c64 *X,
                            uint32_t pb, c64 *Y, int64_t ypitch, const int32_t *rowmap, int nsplit, const int32_t *split,
                            void *scratch) {
    constexpr int CL = NT <= 1 ? 2 : (NT == 2 ? 4 : 8);             // pow2ceil(C / 2): layout of scratch / the fold kernel
    const size_t ring_bytes = (size_t)8 * MmaRing<NT>::BYTES;
    const int cpitch = 2 * CL;
    if (ring_bytes > 48 * 1024)
        IB200_TRY(cudaFuncSetAttribute(kb_tiles_mma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
    kb_tiles_mma_kernel<NT><<<(unsigned)ceil_div(nwork, 8), 256, ring_bytes, s>>>(
        nwork, alpha, (const int4 *)work, (const unsigned char *)entries, X, pb, Y, ypitch, rowmap, (c64 *)scratch, cpitch);
    IB200_LAUNCH_CHECK();
    if (nsplit > 0) {
        kb_blocks_fold_kernel<CL, 4, 4><<<(unsigned)nsplit, 256, 0, s>>>(nsplit, 4 * NT, alpha, (const int4 *)split,
                                                                        (const c64 *)scratch, cpitch, Y, ypitch, rowmap);
        IB200_LAUNCH_CHECK();
    }
    return 0;
}

// cc = 4, 8, 12 or 16 columns
int dispatch_tiles_mma(int cc, cudaStream_t s, int nwork, c64 alpha, const int32_t *work, const void *entries, const c64 *X,
                       uint32_t pb, c64 *Y, int64_t ypitch, const int32_t *rowmap, int nsplit, const int32_t *split,
                       void *scratch) {
    switch (cc) {
        case 4: return launch_tiles_mma<1>(s, nwork, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 8: return launch_tiles_mma<2>(s, nwork, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 12: return launch_tiles_mma<3>(s, nwork, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        case 16: return launch_tiles_mma<4>(s, nwork, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
    }
    set_error("internal: tensor-core tile gather serves 4, 8, 12 or 16 columns, not %d", cc);
    return IB200_E_UNSUPPORTED;
}

}  // namespace ib200
