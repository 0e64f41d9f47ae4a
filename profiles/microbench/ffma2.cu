// microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8], b[8];
    unsigned long long p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = a[i] * 0.5f; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(b[i]); }
    unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = fmaf(a[i], s, b[i]); b[i] = fmaf(b[i], s, a[i]); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, p[(i + 1) & 7]); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, p[(i + 3) & 7]); }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] = a[i] + b[i]; b[i] = b[i] + a[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = add2(p[i], p[(i + 1) & 7]); }
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = add2(p[i], p[(i + 3) & 7]); }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i] + b[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 4, 512>>>(d, iters, 0.999f);
            if (mode == 1) k<1><<<148 * 4, 512>>>(d, iters, 0.999f);
            if (mode == 2) k<2><<<148 * 4, 512>>>(d, iters, 0.999f);
            if (mode == 3) k<3><<<148 * 4, 512>>>(d, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double inst = (double)148 * 4 * 512 * iters * 16;   // thread-instructions
            double lanes = inst * (mode & 1 ? 2 : 1);
            printf("mode %d (%s): %.3f ms  %.2f Tinst/s  %.2f T lane-ops/s\n", mode, mode==0?"FFMA":mode==1?"FFMA2":mode==2?"FADD":"FADD2", ms, inst / ms / 1e9, lanes / ms / 1e9);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
