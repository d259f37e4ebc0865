// racinglmpc_b200/csrc/probe.cuh — fp64 micro-benchmarks of the device the solver runs on.
//
// The FTOCP kernel is bound by fp64 issue / dependent-chain latency, not by HBM, and the driver-written
// MEASURED_PEAKS.json only holds HBM and bf16 numbers.  lmpc_probe_fp64() measures the denominators the QP kernel's
// roofline needs on the box it runs on: DFMA and DMMA (mma.sync.m8n8k4.f64) throughput, and the latencies of the
// operations that make up the kernel's dependent chain (DFMA, DMMA, LDS.64, SHFL, rsqrt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lmpc {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) probe_dfma_tput(double* out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// 4 independent DMMA accumulator chains per warp
__global__ void __launch_bounds__(256) probe_dmma_tput(double* out, int iters, double seed) {
    double c00 = seed, c01 = 0, c10 = seed, c11 = 0, c20 = seed, c21 = 0, c30 = seed, c31 = 0;
    const double a = 1e-3 * (threadIdx.x & 3), b = 1e-3 * (threadIdx.x >> 2);
    for (int i = 0; i < iters; ++i) {
        dmma884(c00, c01, a, b, c00, c01);
        dmma884(c10, c11, a, b, c10, c11);
        dmma884(c20, c21, a, b, c20, c21);
        dmma884(c30, c31, a, b, c30, c31);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((c00 + c01) + (c10 + c11)) + ((c20 + c21) + (c30 + c31));
}

// single-warp dependent chains, cycles per operation in lat[0..5]:
//   0 DFMA, 1 DMMA (accumulator chain), 2 DMMA (result feeds the A operand), 3 LDS.64 (pointer chase), 4 SHFL.64 (2 x SHFL.32), 5 rsqrt(double)
__global__ void __launch_bounds__(32) probe_latency(double* lat, double* sink, int iters) {
    __shared__ int chase[64];
    __shared__ double sval[64];
    const int lane = threadIdx.x;
    chase[lane] = (lane + 7) & 31; chase[lane + 32] = (lane + 5) & 31;
    sval[lane] = 1.0 + 1e-9 * lane; sval[lane + 32] = 1.0;
    __syncwarp();
    double acc = 1.0 + 1e-6 * lane;
    long long t0, t1;
    // DFMA
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc = fma(acc, 0.9999999, 1e-12);
    t1 = clock64();
    if (lane == 0) lat[0] = (double)(t1 - t0) / iters;
    // DMMA accumulator chain
    double c0 = acc, c1 = 0.5 * acc;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma884(c0, c1, 1e-3, 1e-3, c0, c1);
    t1 = clock64();
    if (lane == 0) lat[1] = (double)(t1 - t0) / iters;
    // DMMA result -> A operand
    double a = 1e-3 * acc;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { double d0, d1; dmma884(d0, d1, a, 1e-3, 1e-9, 1e-9); a = d0; }
    t1 = clock64();
    if (lane == 0) lat[2] = (double)(t1 - t0) / iters;
    // LDS pointer chase
    int p = lane;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) p = chase[p];
    t1 = clock64();
    if (lane == 0) lat[3] = (double)(t1 - t0) / iters;
    // SHFL.64
    double s = acc;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) s = __shfl_xor_sync(0xffffffffu, s, 1);
    t1 = clock64();
    if (lane == 0) lat[4] = (double)(t1 - t0) / iters;
    // rsqrt(double)
    double r = 1.0 + acc;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) r = rsqrt(r) + 1.0;
    t1 = clock64();
    if (lane == 0) lat[5] = (double)(t1 - t0) / iters;
    // LDS.64 value chain (load -> use as index is the chase above; this one is load -> fma -> address)
    sink[lane] = acc + c0 + c1 + a + (double)p + s + r + sval[p];
}

}  // namespace lmpc
