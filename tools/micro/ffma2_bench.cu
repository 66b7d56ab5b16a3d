// Micro-benchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) and FADD vs FADD2 on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma2_bench tools/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float a[16]; u64 p[8];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); p[i] = *reinterpret_cast<u64 *>(&t); }
    float2 sv = make_float2(s, s * 0.5f); u64 sp = *reinterpret_cast<u64 *>(&sv);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.25f);          // 16 FFMA (imm-free form: 3 regs)
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sp, sp);             // 8 FFMA2 = 16 lane-FMAs
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + s;                       // 16 FADD
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], sp);                  // 8 FADD2
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2 *>(&p[i]); r += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name, float *out) {
    const int iters = 20000, blocks = 148 * 2, threads = 1024;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0001f);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(out, iters, 1.0001f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double lane_ops = (double)blocks * threads * iters * 16;
    printf("%-6s %.3f ms  %.1f G lane-ops/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, lane_ops / ms / 1e6, lane_ops / ms / 1e6 / 148 / 1.965);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 2 * 1024 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out);
    return 0;
}
