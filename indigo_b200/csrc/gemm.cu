// Dense complex64 GEMM for DenseMatrix operators (coil compression).
// Interfaces replaced: Backend.cgemm (backend.py:481-485) and Backend.csymm
// (backend.py:487-491); numpy semantics np.py:76-90; the reference GPU path is
// cublasCgemm/cublasCsymm (cuda.py:314-366).
//
// Three kernels:
//  * cgemm_tc_kernel (default for tall-skinny products): the tensor-core path for the shape the SENSE path
//    has (SURVEY 8 row a6, cfg5 coil compression: op(M) is 12 x 48 or 48 x 12, X has millions of
//    coil-fastest columns).  The complex product is the real product Y'(2m x n) = M'(2m x 2k) X'(2k x n) on
//    the interleaved (re, im) floats; the columns of X are the rows of the m16n8k8 A operand and go from
//    global memory straight into the fragment registers (one 16-byte load per lane serves two k-steps, the
//    k order inside the sum being free); alpha*op(M)' is expanded once per CTA into a per-lane fragment
//    table in shared memory.  Every product is three TF32 MMAs on the (hi, lo) splits of both operands
//    (lo*hi + hi*lo + hi*hi, fp32 accumulate): complex64-level accuracy (6e-7 measured; 1e-5 parity bar;
//    plain TF32 gives 1e-3).  8(k+m) bytes per column; measured 81 % (12 x 48) / 65 % (48 x 12) of the HBM
//    copy peak on 12.6 M columns (profiles/r01_s9_cgemm_cfg5.md).
//  * cgemm_t5ws_kernel (cgemm_mode 3): the same product on tcgen05.mma with the accumulator in TMEM: one
//    persistent warp-specialised CTA per SM -- two groups of 8 converter warps (global -> registers -> TF32
//    split -> operand ring in the canonical no-swizzle K-major shared-memory layout), one MMA warp (two
//    MMAs per k-step on stacked (hi | lo) planes of op(M)', tcgen05.commit -> mbarriers), four epilogue warps
//    (tcgen05.ld -> global), two operand stages and two accumulator stages.  Parity-green; 1.29 ms on the
//    12 x 48 product (71 % of the copy peak) against 1.13 ms for the mma.sync kernel, 2.3 ms on 48 x 12 (its
//    four epilogue warps write 16-byte pieces of 384-byte columns): an MMA with 24 - 96 real outputs is far
//    too narrow to amortise the operand staging through shared memory that tcgen05 requires, while mma.sync
//    takes X straight from global memory into fragments.  Kept selectable for the wider products
//    (more virtual coils) where the balance turns.
//  * cgemm_kernel: SIMT fp32 tiles, any shape / leading dimension / alignment;
//    serves op(M) in {M, M^H} on the left and the real-symmetric right-multiply
//    through generic element strides.
#include <cstdlib>

#include "common.cuh"

namespace ib200 {

static const int BM = 32, BN = 32, BK = 16;

// C[i,j] = alpha * sum_l A(i,l) * B(l,j) + beta * C[i,j]
//   A(i,l) = A[i*sa_i + l*sa_l]  (conjugated if conjA),  B(l,j) = B[l*sb_l + j*sb_j]
__global__ void __launch_bounds__(256) cgemm_kernel(int64_t m, int64_t n, int64_t k, c64 alpha,
                                                    const c64 *__restrict__ A, int64_t sa_i, int64_t sa_l, int conjA,
                                                    const c64 *__restrict__ B, int64_t sb_l, int64_t sb_j, c64 beta,
                                                    int beta_zero, c64 *__restrict__ C, int64_t ldc) {
    __shared__ c64 As[BK][BM + 1];
    __shared__ c64 Bs[BK][BN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;       // 16 x 16 threads, 2 x 2 outputs each
    const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
    c64 acc[2][2] = {{mk(0, 0), mk(0, 0)}, {mk(0, 0), mk(0, 0)}};
    for (int64_t l0 = 0; l0 < k; l0 += BK) {
        for (int e = threadIdx.x; e < BK * BM; e += 256) {
            const int ii = e % BM, ll = e / BM;
            const int64_t i = i0 + ii, l = l0 + ll;
            c64 v = mk(0.f, 0.f);
            if (i < m && l < k) { v = __ldg(A + i * sa_i + l * sa_l); if (conjA) v.y = -v.y; }
            As[ll][ii] = v;
        }
        for (int e = threadIdx.x; e < BK * BN; e += 256) {
            const int ll = e % BK, jj = e / BK;
            const int64_t l = l0 + ll, j = j0 + jj;
            c64 v = mk(0.f, 0.f);
            if (l < k && j < n) v = __ldg(B + l * sb_l + j * sb_j);
            Bs[ll][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ll = 0; ll < BK; ++ll) {
            const c64 a0 = As[ll][tx], a1 = As[ll][tx + 16];
            const c64 b0 = Bs[ll][ty], b1 = Bs[ll][ty + 16];
            acc[0][0] = cfma(a0, b0, acc[0][0]); acc[0][1] = cfma(a0, b1, acc[0][1]);
            acc[1][0] = cfma(a1, b0, acc[1][0]); acc[1][1] = cfma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int64_t i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
            if (i < m && j < n) {
                c64 *cp = C + i + j * ldc;
                c64 r = cmul(alpha, acc[a][b]);
                if (!beta_zero) r = cfma(beta, *cp, r);
                *cp = r;
            }
        }
}

// ---------------------------------------------------------------- tensor-core path
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// (hi, lo) TF32 split of an fp32 operand in three instructions: hi = x rounded to 10 mantissa bits by an integer
// add-and-mask on the bit pattern (round half away from zero, what cvt.rna does in nine instructions with its
// NaN handling), lo = x - hi exactly; the tensor cores ignore the 13 low mantissa bits of a TF32 operand, so lo
// needs no rounding of its own (|x - hi - tf32(lo)| <= 2^-22 |x|).
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
// same with a 256-byte L2 prefetch hint: a k chunk reads half of each 384-byte column, the hint brings the
// other half into L2 while the DRAM page is open
__device__ __forceinline__ float4 ldg_stream4_pf(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

static const int TC_WARPS = 8;        // warps per CTA, one 16-column tile each per step
static const int TC_KCHUNK = 6;       // 16-float k groups held in registers at a time (6 = 48 complex rows of X)

// Real-form index conventions.  Output row o = 2i + p (p = 0 real, 1 imaginary part of Y[i, j]);
// reduction index kk = 2l + q over the interleaved floats of column j of X.  With W = alpha*op(M)[i, l]:
//   M'[2i][2l] = Re W, M'[2i][2l+1] = -Im W, M'[2i+1][2l] = Im W, M'[2i+1][2l+1] = Re W.
// Lane (g = lane/4, t = lane%4) of k group G holds the floats 16G + 4t .. 4t+3 of columns g and g+8:
// .x/.y feed k-step 2G (fragment slots t and t+4), .z/.w feed k-step 2G+1, and the fragment table
// uses the same assignment, so the permutation of the reduction order cancels.  A last group of only 8
// floats (k % 8 == 4) is a single k-step on floats 16G + 2t, 2t+1.
// One 16-column tile of X as fragment registers: KC k groups for columns j0 = 16*tile + g and j0 + 8.
template <int KC>
__device__ __forceinline__ void tc_load_tile(float4 (&a0)[KC], float4 (&a1)[KC], const float *__restrict__ X, int64_t ldx2,
                                             int64_t tile, int64_t n, int G0, int K2, int g, int t) {
    const int64_t j0 = tile * 16 + g, j1 = j0 + 8;
    const float *x0 = X + j0 * ldx2, *x1 = X + j1 * ldx2;
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        const int G = G0 + u;
        if (K2 - 16 * G == 8) {                                   // tail of 8 floats: one k-step, two floats per lane
            const float2 p0 = j0 < n ? ldg_stream2(x0 + 16 * G + 2 * t) : make_float2(0.f, 0.f);
            const float2 p1 = j1 < n ? ldg_stream2(x1 + 16 * G + 2 * t) : make_float2(0.f, 0.f);
            a0[u] = make_float4(p0.x, p0.y, 0.f, 0.f);
            a1[u] = make_float4(p1.x, p1.y, 0.f, 0.f);
        } else {
            const bool ok = 16 * G + 4 * t + 3 < K2;              // K2 % 4 == 0: a lane's four floats are all in or all out
            a0[u] = (ok && j0 < n) ? ldg_stream4(x0 + 16 * G + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
            a1[u] = (ok && j1 < n) ? ldg_stream4(x1 + 16 * G + 4 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// acc += X-tile fragments (k groups G0 .. G0+KC-1) times the fragment table: three TF32 MMAs per product.
template <int NT, int KC>
__device__ __forceinline__ void tc_mma_tile(float (&acc)[NT][4], const float4 (&a0)[KC], const float4 (&a1)[KC],
                                            const float4 *bfrag, int G0, int KG, int K2, int lane) {
#pragma unroll
    for (int u = 0; u < KC; ++u) {
        const int G = G0 + u;
        if (G < KG) {                                             // warp-uniform
            const float v[2][4] = {{a0[u].x, a1[u].x, a0[u].y, a1[u].y}, {a0[u].z, a1[u].z, a0[u].w, a1[u].w}};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 1 && K2 - 16 * G == 8) break;            // tail group: a single k-step
                uint32_t ah[4], al[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) split_tf32(v[h][r], ah[r], al[r]);
                const float4 *bp = bfrag + (size_t)((2 * G + h) * NT) * 32 + lane;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float4 b = bp[nt * 32];
                    mma_tf32(acc[nt], al, __float_as_uint(b.x), __float_as_uint(b.y));
                    mma_tf32(acc[nt], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                    mma_tf32(acc[nt], ah, __float_as_uint(b.x), __float_as_uint(b.y));
                }
            }
        }
    }
}

// accumulator (row g | g+8, columns 2t, 2t+1 of n-tile nt) = (re, im) of Y[nt*4 + t, j0 | j1]; with `pair`
// (even NT, 16-byte aligned Y columns) the fragment table assigns the outputs so that n-tiles 2p and 2p+1
// hold Y[8p + 2t] and Y[8p + 2t + 1]: one 16-byte store per lane, 64 contiguous bytes per column and instruction
template <int NT>
__device__ __forceinline__ void tc_store_tile(const float (&acc)[NT][4], float *__restrict__ Y, int64_t ldy2, int64_t tile,
                                              int64_t n, int m, c64 beta, int beta_zero, int pair, int g, int t) {
    const int64_t j0 = tile * 16 + g, j1 = j0 + 8;
    if (pair) {
#pragma unroll
        for (int nt = 0; nt + 1 < NT; nt += 2) {
            const int i = 4 * nt + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t j = h ? j1 : j0;
                if (j < n && i < m) {
                    float *yp = Y + j * ldy2 + 2 * i;
                    c64 r0 = mk(acc[nt][2 * h], acc[nt][2 * h + 1]), r1 = mk(acc[nt + 1][2 * h], acc[nt + 1][2 * h + 1]);
                    if (i + 1 < m) {
                        if (!beta_zero) {
                            const float4 o = *reinterpret_cast<const float4 *>(yp);
                            r0 = cfma(beta, mk(o.x, o.y), r0); r1 = cfma(beta, mk(o.z, o.w), r1);
                        }
                        __stcs(reinterpret_cast<float4 *>(yp), make_float4(r0.x, r0.y, r1.x, r1.y));
                    } else {
                        if (!beta_zero) r0 = cfma(beta, *reinterpret_cast<const c64 *>(yp), r0);
                        __stcs(reinterpret_cast<c64 *>(yp), r0);
                    }
                }
            }
        }
        return;
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int i = nt * 4 + t;
        if (i < m) {
            if (j0 < n) {
                c64 *yp = reinterpret_cast<c64 *>(Y + j0 * ldy2) + i;
                c64 r = mk(acc[nt][0], acc[nt][1]);
                if (!beta_zero) r = cfma(beta, *yp, r);
                __stcs(yp, r);
            }
            if (j1 < n) {
                c64 *yp = reinterpret_cast<c64 *>(Y + j1 * ldy2) + i;
                c64 r = mk(acc[nt][2], acc[nt][3]);
                if (!beta_zero) r = cfma(beta, *yp, r);
                __stcs(yp, r);
            }
        }
    }
}

// NT n-tiles of 8 real outputs (4 rows of Y); KC k groups in registers at a time; PIPE: when all of k fits
// one chunk, the loads of a warp's next tile are issued before the MMAs of the current one.
template <int NT, int KC, bool PIPE>
__global__ void __launch_bounds__(TC_WARPS * 32)
cgemm_tc_kernel(int m, int k, int64_t n, c64 alpha, const c64 *__restrict__ Mp, int64_t sa_i, int64_t sa_l, int conjA,
                const float *__restrict__ X, int64_t ldx2, c64 beta, int beta_zero, float *__restrict__ Y, int64_t ldy2, int pair) {
    extern __shared__ float4 bfrag[];                     // [k-step][n-tile][lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
    const int K2 = 2 * k, KG = (K2 + 15) / 16;
    for (int e = threadIdx.x; e < KG * 2 * NT * 32; e += blockDim.x) {
        const int lane = e & 31, nt = (e >> 5) % NT, s = (e >> 5) / NT;
        const int n8 = lane >> 2;
        const int o = pair ? 2 * (8 * (nt >> 1) + 2 * (n8 >> 1) + (nt & 1)) + (n8 & 1) : nt * 8 + n8, i = o >> 1;
        const int G = s >> 1;
        const int l = (K2 - 16 * G == 8) ? ((s & 1) ? k : (16 * G + 2 * (lane & 3)) >> 1)       // tail of 8 floats: one k-step
                                         : (16 * G + 4 * (lane & 3) + 2 * (s & 1)) >> 1;
        float b0 = 0.f, b1 = 0.f;
        if (i < m && l < k) {
            c64 w = __ldg(Mp + i * sa_i + l * sa_l);
            if (conjA) w.y = -w.y;
            w = cmul(alpha, w);
            if (o & 1) { b0 = w.y; b1 = w.x; } else { b0 = w.x; b1 = -w.y; }
        }
        const float h0 = __uint_as_float(to_tf32(b0)), h1 = __uint_as_float(to_tf32(b1));
        bfrag[e] = make_float4(h0, h1, __uint_as_float(to_tf32(b0 - h0)), __uint_as_float(to_tf32(b1 - h1)));
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int64_t ntiles = (n + 15) / 16, stride = (int64_t)gridDim.x * TC_WARPS;
    int64_t tile = (int64_t)blockIdx.x * TC_WARPS + warp;
    if (PIPE && KG <= KC) {
        float4 c0[KC], c1[KC];
        if (tile < ntiles) tc_load_tile<KC>(c0, c1, X, ldx2, tile, n, 0, K2, g, t);
        for (; tile < ntiles; tile += stride) {
            float4 n0[KC], n1[KC];
            tc_load_tile<KC>(n0, n1, X, ldx2, tile + stride, n, 0, K2, g, t);     // past the end: all lanes load nothing
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            tc_mma_tile<NT, KC>(acc, c0, c1, bfrag, 0, KG, K2, lane);
            tc_store_tile<NT>(acc, Y, ldy2, tile, n, m, beta, beta_zero, pair, g, t);
#pragma unroll
            for (int u = 0; u < KC; ++u) { c0[u] = n0[u]; c1[u] = n1[u]; }
        }
        return;
    }
    for (; tile < ntiles; tile += stride) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        for (int G0 = 0; G0 < KG; G0 += KC) {
            float4 a0[KC], a1[KC];
            tc_load_tile<KC>(a0, a1, X, ldx2, tile, n, G0, K2, g, t);
            tc_mma_tile<NT, KC>(acc, a0, a1, bfrag, G0, KG, K2, lane);
        }
        tc_store_tile<NT>(acc, Y, ldy2, tile, n, m, beta, beta_zero, pair, g, t);
    }
}

// ---------------------------------------------------------------- tcgen05 path
// The same real-form product on the 5th-generation tensor cores: D(128 columns of X  x  N real outputs) in
// TMEM, A = 128 columns of X (K-major: the interleaved floats of a column are the reduction index), B =
// alpha*op(M)' (N x 2k, K-major), both in shared memory in the canonical no-swizzle K-major layout (8-row x
// 16-byte core matrices; SBO = 128 bytes between row groups, LBO between 16-byte k slices), hi and lo TF32
// planes of each; per 8-float k-step one elected thread issues lo*hi + hi*lo + hi*hi (kind::tf32, fp32
// accumulate).  The CUDA cores only move data: global -> registers (one k chunk of <= 48 floats per column,
// two chunks in flight per CTA) -> split -> shared, and TMEM -> registers -> global for the result.
namespace t5 {

static const int ROWS = 128, THREADS = 256, RPT = 6, KCH_MAX = 48;   // THREADS: converter threads per group, RPT 16-byte pieces each per chunk

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, version 1 (sm_100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
        if (spin > (1u << 22)) __trap();                   // a lost arrival becomes a launch error, not a hang
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Ctx {
    const float *X; int64_t ldx2, n, ntiles;
    int KCH, CPR;                      // floats / 16-byte slices per column per k chunk
    uint32_t lbo_a, lbo_b;
    unsigned char *a_hi, *a_lo, *b_cat;   // b_cat: rows 0..N-1 = hi plane of alpha*op(M)', rows N..2N-1 = lo plane
    uint32_t tmem, idesc_cat, idesc_hi;   // MMA shapes 128 x 2N and 128 x N
    int64_t goff[RPT];                 // this thread's pieces: float offset r*ldx + 4*slice inside a tile chunk,
    uint32_t soff[RPT];                // byte offset slice*LBO + r*16 inside an operand plane (0xFFFFFFFF: no piece),
    int rrow[RPT];                     // and the column-in-tile r (for the ragged last tile)
};

__device__ __forceinline__ void load_chunk(float4 (&rb)[RPT], const Ctx &c, int64_t tile, int ch) {
    const float *base = c.X + tile * ROWS * c.ldx2 + ch * c.KCH;
    const int64_t left = c.n - tile * ROWS;                       // columns of X from this tile on (<= 0: past the end)
    if (tile < c.ntiles && left >= ROWS) {                        // full tile: no per-piece bounds
#pragma unroll
        for (int it = 0; it < RPT; ++it)
            rb[it] = c.soff[it] != 0xFFFFFFFFu ? ldg_stream4_pf(base + c.goff[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
        for (int it = 0; it < RPT; ++it)
            rb[it] = (c.soff[it] != 0xFFFFFFFFu && tile < c.ntiles && c.rrow[it] < left) ? ldg_stream4_pf(base + c.goff[it])
                                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ---- warp-specialised version: converter warps (global -> registers -> split -> operand ring), one MMA warp,
// epilogue warps (TMEM -> global); two operand stages and two accumulator stages decouple the three roles, so that
// loads, MMAs and result stores of different tiles overlap inside one persistent CTA per SM.
static const int WS_GROUPS = 2, WS_CONV = WS_GROUPS * THREADS, WS_EPI = 128, WS_THREADS = WS_CONV + WS_EPI + 32;

// position in the ring of operand stages: stage index and phase parity, advanced without divisions
struct Ring {
    int s, ns; uint32_t ph;
    __device__ __forceinline__ void advance(int d) { s += d; if (s >= ns) { s -= ns; ph ^= 1u; } }
};

static const int WS_MAX_STAGES = 4;
struct WsBars { uint64_t full[WS_MAX_STAGES], empty[WS_MAX_STAGES], tfull[2], tempty[2]; };

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

// converter: one k chunk from registers into operand stage (q & 1)
__device__ __forceinline__ void ws_convert(const float4 (&rb)[RPT], const Ctx &c, WsBars &bars, uint32_t stage_bytes, const Ring &ring) {
    const int s = ring.s;
    mbar_wait(&bars.empty[s], ring.ph ^ 1u);                           // the MMAs that read this stage ns steps ago are done
    unsigned char *hi = c.a_hi + (size_t)s * stage_bytes, *lo = c.a_lo + (size_t)s * stage_bytes;
#pragma unroll
    for (int it = 0; it < RPT; ++it) {
        if (c.soff[it] != 0xFFFFFFFFu) {
            const float4 v = rb[it];
            uint4 h, l;
            split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<uint4 *>(hi + c.soff[it]) = h;
            *reinterpret_cast<uint4 *>(lo + c.soff[it]) = l;
        }
    }
    fence_async_smem();
    mbar_arrive(&bars.full[s]);
}

template <int NCH>
__global__ void __launch_bounds__(WS_THREADS, 1)
cgemm_t5ws_kernel(int m, int k, int N, int tmem_cols, int lbo_pad, int ns, int64_t n, c64 alpha, const c64 *__restrict__ Mp, int64_t sa_i,
                  int64_t sa_l, int conjA, const float *__restrict__ X, int64_t ldx2, c64 beta, int beta_zero,
                  float *__restrict__ Y, int64_t ldy2) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) WsBars bars;
    __shared__ uint32_t tmem_slot;
    Ctx c;
    const int K2 = 2 * k;
    c.X = X; c.ldx2 = ldx2; c.n = n; c.ntiles = (n + ROWS - 1) / ROWS;
    c.KCH = K2 / NCH; c.CPR = c.KCH / 4;
    c.lbo_a = ROWS * 16 + lbo_pad; c.lbo_b = (uint32_t)(2 * N) * 16;
    const uint32_t stage_bytes = 2u * (uint32_t)c.CPR * c.lbo_a;
    c.a_hi = smem; c.a_lo = c.a_hi + (size_t)c.CPR * c.lbo_a;
    c.b_cat = smem + (size_t)ns * stage_bytes;
    c.idesc_hi = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
    c.idesc_cat = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
#pragma unroll
    for (int it = 0; it < RPT; ++it) {
        const int q = it * THREADS + (threadIdx.x & (THREADS - 1)), r = q / c.CPR, sl = q - r * c.CPR;
        c.rrow[it] = r;
        c.goff[it] = (int64_t)r * ldx2 + 4 * sl;
        c.soff[it] = q < ROWS * c.CPR ? (uint32_t)sl * c.lbo_a + (uint32_t)r * 16 : 0xFFFFFFFFu;
    }
    for (int e = threadIdx.x; e < N * K2; e += WS_THREADS) {
        const int kf = e % K2, o = e / K2, i = o >> 1, l = kf >> 1;
        float v = 0.f;
        if (i < m) {
            c64 w = __ldg(Mp + i * sa_i + l * sa_l);
            if (conjA) w.y = -w.y;
            w = cmul(alpha, w);
            v = (o & 1) ? ((kf & 1) ? w.x : w.y) : ((kf & 1) ? -w.y : w.x);
        }
        uint32_t vh, vl;
        split_tf32(v, vh, vl);
        const uint32_t off = (uint32_t)(kf >> 2) * c.lbo_b + (uint32_t)o * 16 + (uint32_t)(kf & 3) * 4;
        *reinterpret_cast<uint32_t *>(c.b_cat + off) = vh;
        *reinterpret_cast<uint32_t *>(c.b_cat + off + (uint32_t)N * 16) = vl;
    }
    fence_async_smem();
    if (threadIdx.x == 0) {
        for (int s = 0; s < WS_MAX_STAGES; ++s) { mbar_init(&bars.full[s], THREADS); mbar_init(&bars.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.tfull[s], 1); mbar_init(&bars.tempty[s], WS_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    c.tmem = tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntl = (int64_t)blockIdx.x < c.ntiles ? (c.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA
    const int64_t nsteps = ntl * NCH;
    if (warp < WS_CONV / 32) {
        // ---------------- converters
        // two groups of 8 warps take alternate chunks, so that the wait -> split -> fence -> arrive chain of one chunk
        // overlaps the other group's; two chunks of loads in flight per group (96 KB per SM)
        const int grp = warp / (THREADS / 32);
        float4 rbA[RPT], rbB[RPT];
#define IB200_WS_LOAD(RB, Q) load_chunk(RB, c, (int64_t)blockIdx.x + ((Q) / NCH) * (int64_t)gridDim.x, (int)((Q) % NCH))
        int64_t q = grp;
        IB200_WS_LOAD(rbA, q); IB200_WS_LOAD(rbB, q + 2);
        Ring ring; ring.s = grp; ring.ns = ns; ring.ph = 0;
        for (; q < nsteps; q += 4) {
            ws_convert(rbA, c, bars, stage_bytes, ring); ring.advance(2);
            IB200_WS_LOAD(rbA, q + 4);
            if (q + 2 < nsteps) { ws_convert(rbB, c, bars, stage_bytes, ring); ring.advance(2); IB200_WS_LOAD(rbB, q + 6); }
        }
#undef IB200_WS_LOAD
    } else if (warp < (WS_CONV + WS_EPI) / 32) {
        // ---------------- epilogue: TMEM (lane = column of X, column = real output | + N: the hi*lo half) -> Y
        const int quarter = warp & 3;
        for (int64_t i = 0; i < ntl; ++i) {
            const int a = (int)(i & 1);
            const int64_t tile = (int64_t)blockIdx.x + i * (int64_t)gridDim.x, col = tile * ROWS + quarter * 32 + lane;
            mbar_wait(&bars.tfull[a], (uint32_t)((i >> 1) & 1));
            fence_after();
            for (int cb = 0; cb < 2 * m; cb += 16) {
                uint32_t v[16], w[16];
                const uint32_t taddr = c.tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * 2 * N + cb);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr));
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                               "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                             : "r"(taddr + (uint32_t)N));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb + 16 >= 2 * m) {                                // last block read: the accumulator stage is free again
                    fence_before();
                    mbar_arrive(&bars.tempty[a]);
                }
                if (col < n) {
                    float *yp = Y + col * ldy2 + cb;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ii = cb / 2 + 2 * j;
                        c64 r0 = mk(__uint_as_float(v[4 * j]) + __uint_as_float(w[4 * j]), __uint_as_float(v[4 * j + 1]) + __uint_as_float(w[4 * j + 1]));
                        c64 r1 = mk(__uint_as_float(v[4 * j + 2]) + __uint_as_float(w[4 * j + 2]), __uint_as_float(v[4 * j + 3]) + __uint_as_float(w[4 * j + 3]));
                        if (ii + 1 < m) {
                            if (!beta_zero) {
                                const float4 o = *reinterpret_cast<const float4 *>(yp + 4 * j);
                                r0 = cfma(beta, mk(o.x, o.y), r0); r1 = cfma(beta, mk(o.z, o.w), r1);
                            }
                            __stcs(reinterpret_cast<float4 *>(yp + 4 * j), make_float4(r0.x, r0.y, r1.x, r1.y));
                        } else if (ii < m) {
                            if (!beta_zero) r0 = cfma(beta, *reinterpret_cast<const c64 *>(yp + 4 * j), r0);
                            __stcs(reinterpret_cast<c64 *>(yp + 4 * j), r0);
                        }
                    }
                }
            }
        }
    } else {
        // ---------------- MMA warp: lane 0 issues, the warp waits together
        const uint32_t ah0 = smem_u32(c.a_hi), al0 = smem_u32(c.a_lo), bc = smem_u32(c.b_cat);
        Ring ring; ring.s = 0; ring.ns = ns; ring.ph = 0;
        for (int64_t i = 0; i < ntl; ++i) {
            const int a = (int)(i & 1);
            mbar_wait(&bars.tempty[a], (uint32_t)((i >> 1) & 1) ^ 1u);  // epilogue has drained this accumulator stage
            for (int ch = 0; ch < NCH; ++ch) {
                const int s = ring.s;
                mbar_wait(&bars.full[s], ring.ph);
                fence_after();
                if (lane == 0) {
                    const uint32_t tm = c.tmem + (uint32_t)(a * 2 * N);
                    for (int ks = 0; ks < c.KCH / 8; ++ks) {
                        const uint32_t ao = (uint32_t)s * stage_bytes + (uint32_t)(2 * ks) * c.lbo_a;
                        const uint32_t bo = (uint32_t)(ch * c.CPR + 2 * ks) * c.lbo_b;
                        const uint64_t dah = make_desc(ah0 + ao, c.lbo_a, 128), dal = make_desc(al0 + ao, c.lbo_a, 128);
                        const uint64_t db = make_desc(bc + bo, c.lbo_b, 128);
                        mma_ss(tm, dah, db, c.idesc_cat, (ch == 0 && ks == 0) ? 0u : 1u);
                        mma_ss(tm, dal, db, c.idesc_hi, 1u);
                    }
                    commit(&bars.empty[s]);                            // operand stage free when these MMAs have read it
                    if (ch == NCH - 1) commit(&bars.tfull[a]);         // accumulator complete
                }
                __syncwarp();
                ring.advance(1);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(c.tmem), "r"(tmem_cols) : "memory");
}

}  // namespace t5

static int g_t5_pad = -1;             // IB200_T5_PAD: bytes added to the k-slice stride of the X operand planes (bank spread)

// ns (hi, lo) operand stages + stacked op(M)' planes
static size_t t5_smem(int64_t m, int64_t k, int pad, int ns) {
    const int64_t K2 = 2 * k, nch = K2 <= t5::KCH_MAX ? 1 : 2, cpr = K2 / nch / 4, N = (2 * m + 31) / 32 * 32;
    return (size_t)(ns * 2 * cpr * (t5::ROWS * 16 + pad) + 2 * (K2 / 4) * N * 16);
}

// tcgen05 path: op(M) with <= 64 rows, k a multiple of 4 (8 when two chunks are needed) up to 48, X and Y columns 16-byte aligned
static bool t5_applicable(int64_t m, int64_t n, int64_t k, const c64 *B, int64_t sb_l, int64_t sb_j, const c64 *C, int64_t ldc) {
    if (sb_l != 1 || m < 1 || m > 64 || k < 4 || k > 48 || (k & 3) || n < 128) return false;
    if (2 * k > t5::KCH_MAX && (k & 7)) return false;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (sb_j & 1) || (reinterpret_cast<uintptr_t>(C) & 15) || (ldc & 1)) return false;
    return t5_smem(m, k, 128, 2) <= 200 * 1024;
}

static int launch_t5(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i, int64_t sa_l,
                     int conjA, const c64 *B, int64_t ldb, c64 beta, c64 *C, int64_t ldc) {
    if (g_t5_pad < 0) {
        const char *e = getenv("IB200_T5_PAD");
        g_t5_pad = e ? atoi(e) : 16;
        IB200_REQUIRE(g_t5_pad >= 0 && g_t5_pad <= 128 && g_t5_pad % 16 == 0, "IB200_T5_PAD is a multiple of 16 up to 128");
    }
    const int N = (int)((2 * m + 31) / 32 * 32);
    const int nch = 2 * k <= t5::KCH_MAX ? 1 : 2;
    static int ns_env = -1;
    if (ns_env < 0) { const char *e = getenv("IB200_T5_STAGES"); ns_env = e ? atoi(e) : 0; }
    int ns = 2;                                            // operand stages: two measured fastest (12 x 48: 1.29 ms against 1.51 ms with three)
    if (ns_env > 2 && ns_env <= t5::WS_MAX_STAGES && t5_smem(m, k, g_t5_pad, ns_env) <= 200 * 1024) ns = ns_env;
    const size_t smem = t5_smem(m, k, g_t5_pad, ns);
    const int need = 4 * N, tmem_cols = need <= 128 ? 128 : need <= 256 ? 256 : 512;    // two accumulator stages of 2N columns
    auto kern = nch == 1 ? t5::cgemm_t5ws_kernel<1> : t5::cgemm_t5ws_kernel<2>;
    static bool attr_ws[2] = {false, false};
    if (!attr_ws[nch - 1]) {
        IB200_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        attr_ws[nch - 1] = true;
    }
    int64_t grid = ceil_div(n, t5::ROWS);
    if (grid > sm_count()) grid = sm_count();                                           // one persistent CTA per SM
    kern<<<(unsigned)grid, t5::WS_THREADS, smem, s>>>((int)m, (int)k, N, tmem_cols, g_t5_pad, ns, n, alpha, A, sa_i, sa_l, conjA,
                                                      (const float *)B, 2 * ldb, beta, (beta.x == 0.f && beta.y == 0.f) ? 1 : 0,
                                                      (float *)C, 2 * ldc);
    IB200_LAUNCH_CHECK();
    return 0;
}

static int g_gemm_mode = 0;           // 0 = automatic (mma.sync kernel where it applies, else SIMT), 1 = SIMT only,
                                      // 3 = tcgen05 kernel where it applies (slower today, see the header): tests compare the paths

static int g_gemm_pipe = -1;          // -1 = automatic (short k only), 0 / 1 forced (IB200_CGEMM_PIPE, tools/)

template <int NT, int KC, bool PIPE>
static int launch_tc2(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i, int64_t sa_l,
                      int conjA, const c64 *B, int64_t ldb, c64 beta, c64 *C, int64_t ldc) {
    const int KG = (int)((2 * k + 15) / 16);
    const size_t smem = (size_t)KG * 2 * NT * 32 * sizeof(float4);
    const int64_t tiles = ceil_div(n, 16);
    int64_t grid = ceil_div(tiles, TC_WARPS);
    const int64_t cap = (int64_t)sm_count() * 4;          // persistent CTAs: the fragment table is built once per CTA
    if (grid > cap) grid = cap;
    cgemm_tc_kernel<NT, KC, PIPE><<<(unsigned)grid, TC_WARPS * 32, smem, s>>>(
        (int)m, (int)k, n, alpha, A, sa_i, sa_l, conjA, (const float *)B, 2 * ldb, beta,
        (beta.x == 0.f && beta.y == 0.f) ? 1 : 0, (float *)C, 2 * ldc,
        (NT % 2 == 0 && !(reinterpret_cast<uintptr_t>(C) & 15) && !(ldc & 1) && !getenv("IB200_CGEMM_NOPAIR")) ? 1 : 0);
    IB200_LAUNCH_CHECK();
    return 0;
}

template <int NT>
static int launch_tc(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i, int64_t sa_l,
                     int conjA, const c64 *B, int64_t ldb, c64 beta, c64 *C, int64_t ldc) {
    if (g_gemm_pipe < 0) {
        const char *e = getenv("IB200_CGEMM_PIPE");
        g_gemm_pipe = e ? (atoi(e) ? 1 : 0) : 2;
    }
    const int KG = (int)((2 * k + 15) / 16);
    if (KG <= 2) {                                        // short columns (expansion, k = 12): little to load per tile
        if (g_gemm_pipe != 0) return launch_tc2<NT, 2, true>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
        return launch_tc2<NT, 2, false>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
    }
    if (g_gemm_pipe == 1) return launch_tc2<NT, TC_KCHUNK, true>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
    return launch_tc2<NT, TC_KCHUNK, false>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, ldb, beta, C, ldc);
}

// C(m x n) = alpha * op(A)(m x k) * B(k x n) + beta * C with B's columns contiguous: tall-skinny tensor-core path
static bool tc_applicable(int64_t m, int64_t n, int64_t k, const c64 *B, int64_t sb_l, int64_t sb_j) {
    if (g_gemm_mode == 1 || sb_l != 1) return false;
    if (m < 1 || m > 64 || k < 2 || (k & 1) || n < 32) return false;
    if ((reinterpret_cast<uintptr_t>(B) & 15) || (sb_j & 1)) return false;           // 16-byte loads of X columns
    const int64_t KG = (2 * k + 15) / 16, NT = ceil_div(m, 4);
    return KG * NT <= 48;                                                            // fragment table <= 48 KB
}

static int run_gemm(cudaStream_t s, int64_t m, int64_t n, int64_t k, c64 alpha, const c64 *A, int64_t sa_i,
                    int64_t sa_l, int conjA, const c64 *B, int64_t sb_l, int64_t sb_j, c64 beta, c64 *C, int64_t ldc) {
    if (m == 0 || n == 0) return 0;
    if (g_gemm_mode == 3 && t5_applicable(m, n, k, B, sb_l, sb_j, C, ldc))
        return launch_t5(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_j, beta, C, ldc);
    if (tc_applicable(m, n, k, B, sb_l, sb_j)) {
        const int64_t NT = ceil_div(m, 4);
#define IB200_TC(N_) return launch_tc<N_>(s, m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_j, beta, C, ldc)
        if (NT <= 1) IB200_TC(1);
        if (NT <= 2) IB200_TC(2);
        if (NT <= 3) IB200_TC(3);
        if (NT <= 4) IB200_TC(4);
        if (NT <= 6) IB200_TC(6);
        if (NT <= 8) IB200_TC(8);
        if (NT <= 12) IB200_TC(12);
        IB200_TC(16);
#undef IB200_TC
    }
    const int64_t gx = ceil_div(n, BN), gy = ceil_div(m, BM);          // column tiles on x: n may be millions
    IB200_REQUIRE(gy <= 65535 && gx <= 2147483647LL, "cgemm: too many tiles");
    cgemm_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, s>>>(m, n, k, alpha, A, sa_i, sa_l, conjA, B, sb_l, sb_j,
                                                                 beta, (beta.x == 0.f && beta.y == 0.f) ? 1 : 0, C, ldc);
    IB200_LAUNCH_CHECK();
    return 0;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_cgemm_mode(int mode) {
    IB200_REQUIRE(mode == 0 || mode == 1 || mode == 3, "mode is 0 (automatic), 1 (SIMT only) or 3 (tcgen05 where it applies)");
    g_gemm_mode = mode;
    return 0;
}

int ib200_cgemm(void *stream, int conjtrans, int64_t m, int64_t n, int64_t k, float ar, float ai, const void *M,
                int64_t ldm, const void *X, int64_t ldx, float br, float bi, void *Y, int64_t ldy) {
    IB200_RANGE("ib200_cgemm");
    IB200_REQUIRE(m >= 0 && n >= 0 && k >= 0, "negative dimension");
    IB200_REQUIRE((M && X) || k == 0 || m == 0 || n == 0, "null pointer");
    IB200_REQUIRE(Y || m == 0 || n == 0, "null Y");
    const c64 *Mp = (const c64 *)M;
    if (!conjtrans)   // M is m x k
        return run_gemm(as_stream(stream), m, n, k, mk(ar, ai), Mp, 1, ldm, 0, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
    // M is k x m, op(M)(i,l) = conj(M[l + i*ldm])
    return run_gemm(as_stream(stream), m, n, k, mk(ar, ai), Mp, ldm, 1, 1, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
}

int ib200_csymm(void *stream, int left, int64_t m, int64_t n, float ar, float ai, const void *M, int64_t ldm,
                const void *X, int64_t ldx, float br, float bi, void *Y, int64_t ldy) {
    IB200_REQUIRE(m >= 0 && n >= 0, "negative dimension");
    if (m == 0 || n == 0) return 0;
    IB200_REQUIRE(M && X && Y, "null pointer");
    if (left)     // Y(m x n) = M(m x m) X(m x n)
        return run_gemm(as_stream(stream), m, n, m, mk(ar, ai), (const c64 *)M, 1, ldm, 0, (const c64 *)X, 1, ldx, mk(br, bi), (c64 *)Y, ldy);
    // Y(m x n) = X(m x n) M(n x n)
    return run_gemm(as_stream(stream), m, n, n, mk(ar, ai), (const c64 *)X, 1, ldx, 0, (const c64 *)M, 1, ldm, mk(br, bi), (c64 *)Y, ldy);
}

}  // extern "C"
