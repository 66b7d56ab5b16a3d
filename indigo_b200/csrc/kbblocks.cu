// Adjoint gridding on blocks of grid points: ccsrmm(G', adjoint) of the fused SENSE recipe, matrix-free.
//
// Reference being replaced: the adjoint product of the gridding matrix that indigo/interp.py:19-80 emits
// (SpMatrix._eval with forward=False, operators.py:322-332 -> Backend.ccsrmm, backend.py:560-596).
//
// The x-run lists of csrmm_runs.cu cost 20 bytes and one gather per (sample, 4 grid points) pair -- 54 pairs
// per sample, 7.4 GB at cfg3 -- and that stream does not shrink when the coils are sharded: at 2 coils per GPU
// it is the whole cost of the step (2.8 of 5.6 ms).  A Kaiser-Bessel footprint is an outer product of three
// short weight vectors, and it stays one inside any box of grid points.  Here the grid is cut into blocks of
// 4 x BY x BZ points (BY, BZ in {1, 2, 4}, inside the 4 x 4 x 4 tiles of the tile-major row order) and every
// (sample, block) pair that meets is one entry
//     (sample, wx[4], wy[BY], wz[BZ])
// which serves the 4*BY*BZ points of the block with ONE gather of the sample's coils:
//     4 x 4 x 4   8 entries / sample, 52 B each   2.8 GB at cfg3   64 multiply-adds per entry and coil (15.6 useful)
//     4 x 2 x 2  18 entries / sample, 36 B each   4.4 GB           16  (6.9 useful)
//     4 x 2 x 1  30 entries / sample, 28 B each   5.7 GB            8  (4.2 useful)      [wz folded into wy]
//     4 x 1 x 1  50 entries / sample, 20 B each   6.8 GB            4  (2.5 useful)      [one weight vector]
// Large blocks trade arithmetic on zero weights for less stream, fewer gathers and fewer instructions per
// sample: right when the coils are few (coil-sharded operators); small blocks when they are many.  Entries are
// built straight from the separable records of the forward gather (kbgrid.cu): forward and adjoint use
// bit-identical weight factors, and no stored adjoint matrix is read.
//
// Layout: entries of a block are consecutive, ordered by record, in batches of four:
//     batch = ids[4] | wx[4][4] | wy[4][BY] | wz[4][BZ]         16-byte aligned, zero-weight padding
// A work item is (block, batch range, slot): blocks with more than seg_batches batches (k-space centre of a
// radial trajectory) are cut into several items that write partial sums into scratch[slot]; a fold kernel adds
// them in a fixed order, so the result does not depend on scheduling.  Blocks without entries but with rows
// inside the support windows get one empty item: every grid point inside the windows is overwritten on every
// apply.
#include "kbblocks.cuh"
#include "kb.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cstdlib>

namespace ib200 {

int exclusive_scan_public(cudaStream_t s, int64_t n, const int32_t *in, int32_t *out);   // csrmm.cu

// ---- setup: which blocks a sample meets, and with which weights -------------------------------------------
struct AxisBlocks { int n; int bc[kKbTaps]; float w[kKbTaps][kTE]; };

// taps j0, j0+1, ... (mod N) of one axis, sorted into the blocks of extent E they fall into; blocks whose weights
// are all zero (sixth tap of an on-grid sample) are dropped
__device__ __forceinline__ void axis_blocks(int j0, int cnt, int N, int E, const float *w6, AxisBlocks &a) {
    a.n = 0;
    int j = j0;
    for (int t = 0; t < kKbTaps; ++t) {
        if (t < cnt) {
            const int bc = j / E, pos = j % E;
            int k = -1;
            for (int q = 0; q < a.n; ++q) if (a.bc[q] == bc) k = q;
            if (k < 0) { k = a.n++; a.bc[k] = bc; for (int i = 0; i < kTE; ++i) a.w[k][i] = 0.f; }
            a.w[k][pos] += w6[t];
            if (++j >= N) j = 0;
        }
    }
    int m = 0;
    for (int q = 0; q < a.n; ++q) {
        bool any = false;
        for (int i = 0; i < kTE; ++i) any = any || a.w[q][i] != 0.f;
        if (any) { if (m != q) { a.bc[m] = a.bc[q]; for (int i = 0; i < kTE; ++i) a.w[m][i] = a.w[q][i]; } ++m; }
    }
    a.n = m;
}

__device__ __forceinline__ void record_blocks(const KbRecord &q, int n0, int n1, int n2, BlockShape sh, AxisBlocks &ax,
                                              AxisBlocks &ay, AxisBlocks &az) {
    axis_blocks(q.ix0, q.ntaps & 255, n0, kTE, q.wx, ax);
    axis_blocks(q.iy0, (q.ntaps >> 8) & 255, n1, sh.by, q.wy, ay);
    axis_blocks(q.iz0, (q.ntaps >> 16) & 255, n2, sh.bz, q.wz, az);
}

// block number of block coordinates (bx, cy, cz) (cy, cz in units of BY, BZ points)
__device__ __forceinline__ int block_id(int bx, int cy, int cz, BlockShape sh, int nt0, int nt1) {
    const int py = kTE / sh.by, pz = kTE / sh.bz;
    const int64_t tile = ((int64_t)(cz / pz) * nt1 + cy / py) * nt0 + bx;
    return (int)(tile * sh.nsub() + (cz % pz) * py + cy % py);
}

// work items of a block with nb batches: segments of seg_batches batches, lengthened for the densest blocks so that
// no block has more than kTMaxSeg of them
static const int kTMaxSeg = 48;
__host__ __device__ __forceinline__ int block_seg_len(int nb, int seg_batches) {
    const int cap = (nb + kTMaxSeg - 1) / kTMaxSeg;
    return cap > seg_batches ? cap : seg_batches;
}

// pairs per record (rcnt) and per block (bcnt)
__global__ void __launch_bounds__(128) block_pairs_count_kernel(int64_t m, const KbRecord *__restrict__ rec, int n0, int n1,
                                                                int n2, BlockShape sh, int nt0, int nt1,
                                                                int32_t *__restrict__ rcnt, int32_t *bcnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const KbRecord q = rec[r];
    AxisBlocks ax, ay, az;
    record_blocks(q, n0, n1, n2, sh, ax, ay, az);
    if (rcnt) rcnt[r] = ax.n * ay.n * az.n;
    if (bcnt)
        for (int k = 0; k < az.n; ++k)
            for (int j = 0; j < ay.n; ++j)
                for (int i = 0; i < ax.n; ++i) atomicAdd(bcnt + block_id(ax.bc[i], ay.bc[j], az.bc[k], sh, nt0, nt1), 1);
}

__global__ void __launch_bounds__(128) block_pairs_emit_kernel(int64_t m, const KbRecord *__restrict__ rec, int n0, int n1,
                                                               int n2, BlockShape sh, int nt0, int nt1,
                                                               const int32_t *__restrict__ rpos, int32_t *__restrict__ keys,
                                                               int32_t *__restrict__ vals) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const KbRecord q = rec[r];
    AxisBlocks ax, ay, az;
    record_blocks(q, n0, n1, n2, sh, ax, ay, az);
    int64_t at = rpos[r];
    for (int k = 0; k < az.n; ++k)
        for (int j = 0; j < ay.n; ++j)
            for (int i = 0; i < ax.n; ++i) {
                keys[at] = block_id(ax.bc[i], ay.bc[j], az.bc[k], sh, nt0, nt1);
                vals[at] = (int32_t)r;
                ++at;
            }
}

// batches and work items of every block; totals[0] += split blocks, totals[1] += work items of split blocks
__global__ void __launch_bounds__(256) block_sizes_kernel(int64_t nblocks, const int32_t *__restrict__ bcnt,
                                                          const int32_t *__restrict__ rowmap, BlockShape sh, int seg_batches,
                                                          int32_t *__restrict__ nbatch, int32_t *__restrict__ nwork,
                                                          int *totals) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int nb = (bcnt[b] + kTB - 1) / kTB;
    bool rows = false;
    for (int rr = 0; rr < sh.by * sh.bz; ++rr) {
        const int4 v = __ldg(reinterpret_cast<const int4 *>(rowmap + block_row((int)b, rr, sh.by, sh.bz)));
        rows = rows || v.x >= 0 || v.y >= 0 || v.z >= 0 || v.w >= 0;
    }
    int nw = 0;
    if (rows) { const int sl = block_seg_len(nb, seg_batches); nw = (nb + sl - 1) / sl; if (nw < 1) nw = 1; }
    nbatch[b] = rows ? nb : 0;
    nwork[b] = nw;
    if (nw > 1) { atomicAdd(totals, 1); atomicAdd(totals + 1, nw); }
}

// first sorted pair of every block that has pairs
__global__ void __launch_bounds__(256) block_starts_kernel(int64_t npairs, const int32_t *__restrict__ keys,
                                                           int32_t *__restrict__ bstart) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    if (i == 0 || keys[i] != keys[i - 1]) bstart[keys[i]] = (int32_t)i;
}

__global__ void __launch_bounds__(128) block_fill_kernel(int64_t npairs, const int32_t *__restrict__ keys,
                                                         const int32_t *__restrict__ vals,
                                                         const int32_t *__restrict__ bstart,
                                                         const KbRecord *__restrict__ rec, int n0, int n1, int n2,
                                                         BlockShape sh, int nt0, int nt1, const int32_t *__restrict__ bptr,
                                                         unsigned char *__restrict__ ent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const int b = keys[i];
    if (bptr[b + 1] == bptr[b]) return;                              // block without rows inside the windows
    const KbRecord q = rec[vals[i]];
    AxisBlocks ax, ay, az;
    record_blocks(q, n0, n1, n2, sh, ax, ay, az);
    const int py = kTE / sh.by, pz = kTE / sh.bz, nsub = py * pz;
    const int tile = b / nsub, sub = b % nsub;
    const int tx = tile % nt0, ty = (tile / nt0) % nt1, tz = tile / (nt0 * nt1);
    const int cy = ty * py + sub % py, cz = tz * pz + sub / py;
    float wx[kTE] = {0.f, 0.f, 0.f, 0.f}, wy[kTE] = {0.f, 0.f, 0.f, 0.f}, wz[kTE] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ax.n; ++k) if (ax.bc[k] == tx) for (int p = 0; p < kTE; ++p) wx[p] = ax.w[k][p];
    for (int k = 0; k < ay.n; ++k) if (ay.bc[k] == cy) for (int p = 0; p < kTE; ++p) wy[p] = ay.w[k][p];
    for (int k = 0; k < az.n; ++k) if (az.bc[k] == cz) for (int p = 0; p < kTE; ++p) wz[p] = az.w[k][p];
    if (!sh.has_wz()) for (int p = 0; p < kTE; ++p) wy[p] = __fmul_rn(wz[0], wy[p]);          // BZ = 1: wz folded into wy
    if (!sh.has_wy()) for (int p = 0; p < kTE; ++p) wx[p] = __fmul_rn(wy[0], wx[p]);          // one row: all in wx
    const int kk = (int)(i - bstart[b]);
    unsigned char *base = ent + ((int64_t)bptr[b] + kk / kTB) * sh.batch_bytes();
    const int u = kk % kTB;
    const bool last = i + 1 == npairs || keys[i + 1] != b;           // last entry of the block pads its batch
    for (int v = u; v < (last ? kTB : u + 1); ++v) {
        const float s = v == u ? 1.f : 0.f;
        reinterpret_cast<int32_t *>(base)[v] = q.out;
        *reinterpret_cast<float4 *>(base + 16 + 16 * v) = make_float4(s * wx[0], s * wx[1], s * wx[2], s * wx[3]);
        if (sh.has_wy()) for (int p = 0; p < sh.by; ++p) reinterpret_cast<float *>(base + sh.wy_off())[v * sh.by + p] = s * wy[p];
        if (sh.has_wz()) for (int p = 0; p < sh.bz; ++p) reinterpret_cast<float *>(base + sh.wz_off())[v * sh.bz + p] = s * wz[p];
    }
}

// work items {block, first batch, end batch, scratch slot or -1} and split descriptors {block, first slot, items, 0}
__global__ void __launch_bounds__(256) block_work_kernel(int64_t nblocks, const int32_t *__restrict__ bptr,
                                                         const int32_t *__restrict__ wptr, int seg_batches,
                                                         int4 *__restrict__ work, int4 *__restrict__ split, int *cursors) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int w0 = wptr[b], nw = wptr[b + 1] - w0;
    if (nw == 0) return;
    const int b0 = bptr[b], b1 = bptr[b + 1];
    if (nw == 1) { work[w0] = make_int4((int)b, b0, b1, -1); return; }
    const int s0 = atomicAdd(cursors, nw);
    const int at = atomicAdd(cursors + 1, 1);
    const int sl = block_seg_len(b1 - b0, seg_batches);
    for (int j = 0; j < nw; ++j) {
        const int a = b0 + j * sl;
        work[w0 + j] = make_int4((int)b, a, a + sl < b1 ? a + sl : b1, s0 + j);
    }
    split[at] = make_int4((int)b, s0, nw, 0);
}

static int blocks_pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

static bool blocks_shape_ok(int by, int bz) {
    return (by == 4 && bz == 4) || (by == 2 && bz == 2) || (by == 2 && bz == 1) || (by == 1 && bz == 1);
}

static bool blocks_grid_ok(const int64_t grid[3], int by, int bz, int64_t *nblocks, int nt[3]) {
    if (!grid || !blocks_shape_ok(by, bz)) return false;
    int64_t n = 1;
    for (int d = 0; d < 3; ++d) {
        if (grid[d] <= 0 || grid[d] >= (1LL << 30)) return false;
        nt[d] = (int)ceil_div(grid[d], kTE);
        n *= nt[d];
    }
    const BlockShape sh{by, bz};
    *nblocks = n * sh.nsub();
    return n * kTV < (1LL << 31);
}


extern template int dispatch_coils<4, 4>(int, int, cudaStream_t, int, int, c64, const int32_t *, const void *, const c64 *, uint32_t, c64 *, int64_t, const int32_t *, int, const int32_t *, void *);
extern template int dispatch_coils<2, 2>(int, int, cudaStream_t, int, int, c64, const int32_t *, const void *, const c64 *, uint32_t, c64 *, int64_t, const int32_t *, int, const int32_t *, void *);
extern template int dispatch_coils<2, 1>(int, int, cudaStream_t, int, int, c64, const int32_t *, const void *, const c64 *, uint32_t, c64 *, int64_t, const int32_t *, int, const int32_t *, void *);
extern template int dispatch_coils<1, 1>(int, int, cudaStream_t, int, int, c64, const int32_t *, const void *, const c64 *, uint32_t, c64 *, int64_t, const int32_t *, int, const int32_t *, void *);

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_kb_blocks_batch_bytes(int by, int bz) {
    if (!blocks_shape_ok(by, bz)) return -1;
    const BlockShape sh{by, bz};
    return sh.batch_bytes();
}

int ib200_kb_blocks_count(void *stream, int64_t m, const void *records, const int64_t grid[3], int by, int bz,
                          const int32_t *rowmap, int seg_batches, int32_t *bptr, int32_t *wptr, int64_t *host_totals) {
    int64_t nblocks = 0;
    int nt[3];
    IB200_REQUIRE(blocks_grid_ok(grid, by, bz, &nblocks, nt), "bad grid or block shape");
    IB200_REQUIRE(nblocks < (1LL << 31), "too many blocks");
    IB200_REQUIRE(m >= 0 && m < (1LL << 31) && seg_batches >= 1 && host_totals, "bad arguments");
    IB200_REQUIRE(rowmap && bptr && wptr && (records || m == 0), "null pointer");
    for (int i = 0; i < 5; ++i) host_totals[i] = 0;
    const BlockShape sh{by, bz};
    cudaStream_t s = as_stream(stream);
    int32_t *buf = nullptr;
    IB200_TRY(cudaMalloc(&buf, (size_t)(3 * nblocks + 4) * sizeof(int32_t)));
    int32_t *bcnt = buf, *nbatch = buf + nblocks, *nwork = buf + 2 * nblocks;
    int *totals = reinterpret_cast<int *>(buf + 3 * nblocks);
    cudaMemsetAsync(buf, 0, (size_t)(3 * nblocks + 4) * sizeof(int32_t), s);
    if (m > 0) {
        block_pairs_count_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, (const KbRecord *)records, (int)grid[0], (int)grid[1],
                                                                           (int)grid[2], sh, nt[0], nt[1], nullptr, bcnt);
        count_launch();
    }
    block_sizes_kernel<<<(unsigned)ceil_div(nblocks, 256), 256, 0, s>>>(nblocks, bcnt, rowmap, sh, seg_batches, nbatch, nwork, totals);
    count_launch();
    int rc = exclusive_scan_public(s, nblocks, nbatch, bptr);
    if (!rc) rc = exclusive_scan_public(s, nblocks, nwork, wptr);
    int32_t tb = 0, tw = 0;
    int ht[2] = {0, 0};
    cudaError_t e = cudaSuccess;
    if (!rc) {
        cudaMemcpyAsync(&tb, bptr + nblocks, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(&tw, wptr + nblocks, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(ht, totals, sizeof(ht), cudaMemcpyDeviceToHost, s);
        e = cudaStreamSynchronize(s);
    }
    cudaFree(buf);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    IB200_REQUIRE(tb >= 0 && tw >= 0, "block lists exceed 2^31 batches");
    host_totals[0] = nblocks; host_totals[1] = tb; host_totals[2] = tw; host_totals[3] = ht[0]; host_totals[4] = ht[1];
    return 0;
}

int ib200_kb_blocks_fill(void *stream, int64_t m, const void *records, const int64_t grid[3], int by, int bz, int seg_batches,
                         const int32_t *bptr, const int32_t *wptr, void *entries, int32_t *work, int32_t *split) {
    int64_t nblocks = 0;
    int nt[3];
    IB200_REQUIRE(blocks_grid_ok(grid, by, bz, &nblocks, nt), "bad grid or block shape");
    IB200_REQUIRE(m >= 0 && m < (1LL << 31) && seg_batches >= 1, "bad arguments");
    IB200_REQUIRE(bptr && wptr && entries && work && split && (records || m == 0), "null pointer");
    IB200_REQUIRE(((uintptr_t)entries & 15) == 0 && ((uintptr_t)work & 15) == 0 && ((uintptr_t)split & 15) == 0,
                  "block arrays must be 16-byte aligned");
    const BlockShape sh{by, bz};
    cudaStream_t s = as_stream(stream);
    const KbRecord *rec = (const KbRecord *)records;
    const int n0 = (int)grid[0], n1 = (int)grid[1], n2 = (int)grid[2];
    int rc = 0;
    cudaError_t e = cudaSuccess;
    int32_t *rpos = nullptr, *pairs = nullptr, *bstart = nullptr;
    void *tmp = nullptr;
    int *cursors = nullptr;
    int32_t npairs = 0;
    if (m > 0) {
        e = cudaMalloc(&rpos, (size_t)(m + 1) * sizeof(int32_t));
        if (e == cudaSuccess) {
            block_pairs_count_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, rec, n0, n1, n2, sh, nt[0], nt[1], rpos, nullptr);
            count_launch();
            rc = exclusive_scan_public(s, m, rpos, rpos);
            if (!rc) {
                cudaMemcpyAsync(&npairs, rpos + m, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                e = cudaStreamSynchronize(s);
            }
        }
        if (!rc && e == cudaSuccess && npairs < 0) { set_error("block lists exceed 2^31 entries"); rc = IB200_E_INVALID; }
    }
    if (!rc && e == cudaSuccess && npairs > 0) {
        e = cudaMalloc(&pairs, (size_t)npairs * 4 * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc(&bstart, (size_t)nblocks * sizeof(int32_t));
        if (e == cudaSuccess) {
            int32_t *ka = pairs, *kb = pairs + npairs, *va = pairs + 2 * (int64_t)npairs, *vb = pairs + 3 * (int64_t)npairs;
            block_pairs_emit_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(m, rec, n0, n1, n2, sh, nt[0], nt[1], rpos, ka, va);
            count_launch();
            int end_bit = 1; while ((1LL << end_bit) < nblocks) ++end_bit;
            cub::DoubleBuffer<int32_t> dk(ka, kb), dv(va, vb);
            size_t tmp_bytes = 0;
            e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)npairs, 0, end_bit, s);
            if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
            if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)npairs, 0, end_bit, s);
            if (e == cudaSuccess) {
                count_launch(end_bit / 8 + 2);
                block_starts_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, s>>>(npairs, dk.Current(), bstart);
                block_fill_kernel<<<(unsigned)ceil_div(npairs, 128), 128, 0, s>>>(npairs, dk.Current(), dv.Current(), bstart, rec,
                                                                              n0, n1, n2, sh, nt[0], nt[1], bptr,
                                                                              (unsigned char *)entries);
                count_launch(2);
            }
        }
    }
    if (!rc && e == cudaSuccess) e = cudaMalloc(&cursors, 2 * sizeof(int));
    if (!rc && e == cudaSuccess) {
        cudaMemsetAsync(cursors, 0, 2 * sizeof(int), s);
        block_work_kernel<<<(unsigned)ceil_div(nblocks, 256), 256, 0, s>>>(nblocks, bptr, wptr, seg_batches, (int4 *)work,
                                                                         (int4 *)split, cursors);
        count_launch();
        e = cudaStreamSynchronize(s);
    }
    cudaFree(rpos); cudaFree(pairs); cudaFree(bstart); cudaFree(tmp); cudaFree(cursors);
    if (rc) return rc;
    IB200_TRY(e);
    IB200_TRY(cudaGetLastError());
    return 0;
}

int ib200_kb_blocks_apply(void *stream, int64_t ncols, int by, int bz, float ar, float ai, int nwork, const int32_t *work,
                          const void *entries, const void *Xil, int64_t xpitch, void *Yil, int64_t ypitch,
                          const int32_t *rowmap, int nsplit, const int32_t *split, void *scratch, int lanes) {
    IB200_RANGE("ib200_kb_blocks_apply");
    IB200_REQUIRE(nwork >= 0 && nsplit >= 0 && ncols >= 0, "bad arguments");
    IB200_REQUIRE(blocks_shape_ok(by, bz), "block shape must be 4x4x4, 4x2x2, 4x2x1 or 4x1x1");
    if (nwork == 0 || ncols == 0) return 0;
    IB200_REQUIRE(ncols <= 64 && ncols % 2 == 0, "block gather serves an even number of at most 64 columns");
    IB200_REQUIRE(work && entries && Xil && Yil && rowmap, "null pointer");
    IB200_REQUIRE(xpitch >= ncols && ypitch >= ncols && xpitch % 2 == 0 && ypitch % 2 == 0, "bad pitch");
    IB200_REQUIRE(((uintptr_t)Xil & 15) == 0 && ((uintptr_t)Yil & 15) == 0 && ((uintptr_t)entries & 15) == 0,
                  "operands must be 16-byte aligned");
    IB200_REQUIRE(xpitch * (int64_t)sizeof(c64) < (1LL << 32), "operand pitch too large");
    IB200_REQUIRE(nsplit == 0 || (split && scratch && ((uintptr_t)scratch & 15) == 0), "split arrays missing");
    const c64 alpha = mk(ar, ai);
    cudaStream_t s = as_stream(stream);
    int want = lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8 || lanes == 16 ? lanes : (by * bz >= 16 ? 8 : (by * bz >= 4 ? 2 : 1));
    if (const char *e = getenv("IB200_BLOCKS_LANES")) {              // tuning knob (tools/)
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) want = v;
    }
    const uint32_t pb = (uint32_t)(xpitch * sizeof(c64));
    // chunks of at most 16 columns (8 coil lanes)
    for (int64_t c0 = 0; c0 < ncols; c0 += 16) {
        const int cc = (int)(ncols - c0 < 16 ? ncols - c0 : 16);
        const int CL = blocks_pow2_ceil(cc / 2);
        const c64 *X = (const c64 *)Xil + c0;
        c64 *Y = (c64 *)Yil + c0;
        int rc = IB200_E_UNSUPPORTED;
        if (by == 4 && bz == 4) rc = dispatch_coils<4, 4>(CL, want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        else if (by == 2 && bz == 2) rc = dispatch_coils<2, 2>(CL, want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        else if (by == 2 && bz == 1) rc = dispatch_coils<2, 1>(CL, want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        else if (by == 1 && bz == 1) rc = dispatch_coils<1, 1>(CL, want, s, nwork, cc, alpha, work, entries, X, pb, Y, ypitch, rowmap, nsplit, split, scratch);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
