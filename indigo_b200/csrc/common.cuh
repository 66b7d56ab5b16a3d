// Shared helpers for libindigo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include <nvtx3/nvToolsExt.h>

#include "../../include/indigo_b200.h"

namespace ib200 {

// ---- error plumbing: nothing throws across the C ABI -----------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define IB200_TRY(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            ::ib200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                               cudaGetErrorString(_e));                              \
            return (int)_e;                                                          \
        }                                                                            \
    } while (0)

#define IB200_REQUIRE(cond, msg)                                                     \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::ib200::set_error("%s:%d: invalid argument: %s (%s)", __FILE__,        \
                               __LINE__, msg, #cond);                                \
            return IB200_E_INVALID;                                                  \
        }                                                                            \
    } while (0)

// call after every kernel launch
#define IB200_LAUNCH_CHECK()                                                         \
    do {                                                                             \
        ::ib200::count_launch();                                                     \
        IB200_TRY(cudaGetLastError());                                               \
    } while (0)

// NVTX range around a C-ABI call (ranges per backend call / fused step, visible in
// Nsight Systems and in ncu's --nvtx filters; header-only NVTX3, a no-op without an attached tool)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define IB200_RANGE(name) ::ib200::NvtxRange _ib200_range(name)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();             // SMs of the current device (cached)
// fork a per-device side stream off `main` / join it back (core.cu); every begin must be followed by an end
int side_stream_begin(cudaStream_t main, cudaStream_t *side);
int side_stream_end(cudaStream_t main);
int64_t smem_optin();       // max opt-in dynamic shared memory per block

// ---- complex64 arithmetic on float2 ----------------------------------------
typedef float2 c64;

__host__ __device__ __forceinline__ c64 mk(float re, float im) { return make_float2(re, im); }
__device__ __forceinline__ c64 cadd(c64 a, c64 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ c64 csub(c64 a, c64 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ c64 cmul(c64 a, c64 b) {
    return mk(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ c64 cmulc(c64 a, c64 b) {
    return mk(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
// acc + a*b
__device__ __forceinline__ c64 cfma(c64 a, c64 b, c64 acc) {
    acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
    return acc;
}
// acc + conj(a)*b
__device__ __forceinline__ c64 cfmac(c64 a, c64 b, c64 acc) {
    acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(-a.y, b.x, acc.y);
    return acc;
}
__device__ __forceinline__ c64 cconj(c64 a) { return mk(a.x, -a.y); }
__device__ __forceinline__ c64 cscale(float s, c64 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ c64 cswap(c64 a) { return mk(a.y, a.x); }
// multiply by -i (forward) : (x, y) -> (y, -x)
__device__ __forceinline__ c64 cmul_mi(c64 a) { return mk(a.y, -a.x); }
// multiply by +i
__device__ __forceinline__ c64 cmul_pi(c64 a) { return mk(-a.y, a.x); }

__device__ __forceinline__ c64 ldg(const c64 *p) { return __ldg(p); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ib200
