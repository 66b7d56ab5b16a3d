// Packed pair of fp32 values in one 64-bit register: add / mul / fma on both halves cost ONE issue slot
// on sm_100a (FADD2 / FMUL2 / FFMA2; tools/micro/ffma2_bench.cu).  ptxas folds broadcast scalars
// (p_bc), immediates and negated addends into operand modifiers.  Plain struct on the host so that the
// CPU emulation of the FFT tile code (tests/csrc/fft_emul.cu) runs the same templates.
#pragma once
#include "common.cuh"

#ifndef IB_HD
#ifdef __CUDACC__
#define IB_HD __host__ __device__ __forceinline__
#else
#define IB_HD inline
#endif
#endif

namespace ib200 {

// ---- packed pair of fp32 ------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
struct pk2 { unsigned long long v; };
__device__ __forceinline__ pk2 p_make(float a, float b) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float p_lo(pk2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return a; }
__device__ __forceinline__ float p_hi(pk2 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); return b; }
__device__ __forceinline__ pk2 p_add(pk2 a, pk2 b) { pk2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pk2 p_sub(pk2 a, pk2 b) { pk2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pk2 p_mul(pk2 a, pk2 b) { pk2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pk2 p_fma(pk2 a, pk2 b, pk2 c) { pk2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct pk2 { float a, b; };
inline pk2 p_make(float a, float b) { pk2 r; r.a = a; r.b = b; return r; }
inline float p_lo(pk2 p) { return p.a; }
inline float p_hi(pk2 p) { return p.b; }
inline pk2 p_add(pk2 a, pk2 b) { return p_make(a.a + b.a, a.b + b.b); }
inline pk2 p_sub(pk2 a, pk2 b) { return p_make(a.a - b.a, a.b - b.b); }
inline pk2 p_mul(pk2 a, pk2 b) { return p_make(a.a * b.a, a.b * b.b); }
inline pk2 p_fma(pk2 a, pk2 b, pk2 c) { return p_make(a.a * b.a + c.a, a.b * b.b + c.b); }
#endif
IB_HD pk2 p_bc(float c) { return p_make(c, c); }                    // broadcast: an operand modifier / immediate in SASS
IB_HD pk2 p_fmac(float c, pk2 a, pk2 acc) { return p_fma(p_bc(c), a, acc); }   // acc + c*a
IB_HD pk2 p_scale(float c, pk2 a) { return p_mul(p_bc(c), a); }


}  // namespace ib200
