// FFMA2 throughput by operand pattern on B200: is the interpolation kernel's inner loop (scalar-broadcast sample x weight
// pair + accumulator pair, ~60 live 64-bit registers) issued at the 2 cycles per warp instruction of the plain chain test?
//   A: kernel-like  acc[mm][q] = ffma2(bcast(w[mm + t].x|.y), c[t][q], acc)   (28 accumulators, 14 samples, 16 weight pairs)
//   B: the same with the sample pre-duplicated into a register pair (no broadcast modifier)
//   C: plain chains  v = ffma2(v, a, b) with constant a, b (profiles/microbench/ffma2.cu)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int K = 8, R = 7;
template <int MODE> __global__ void __launch_bounds__(128, 4) kern(const float2* in, float* out, int iters) {
    float2 w[R + K - 1];
    float2 c[K][2];
    for (int i = 0; i < R + K - 1; ++i) w[i] = in[threadIdx.x + 32 * i];
    for (int t = 0; t < K; ++t) { c[t][0] = in[threadIdx.x + 7 * t]; c[t][1] = in[threadIdx.x + 11 * t + 3]; }
    float s = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
            float2 are0 = make_float2(0.f, 0.f), aim0 = are0, are1 = are0, aim1 = are0;
#pragma unroll
            for (int t = 0; t < K; ++t) {
                if (MODE == 0) {
                    are0 = __ffma2_rn(make_float2(w[mm + t].x, w[mm + t].x), c[t][0], are0);
                    aim0 = __ffma2_rn(make_float2(w[mm + t].y, w[mm + t].y), c[t][0], aim0);
                    are1 = __ffma2_rn(make_float2(w[mm + t].x, w[mm + t].x), c[t][1], are1);
                    aim1 = __ffma2_rn(make_float2(w[mm + t].y, w[mm + t].y), c[t][1], aim1);
                } else if (MODE == 1) {   // pair x pair: (re, im) of one output per accumulator, weight broadcast... as full pairs
                    are0 = __ffma2_rn(w[mm + t], c[t][0], are0);
                    aim0 = __ffma2_rn(w[mm + t], c[t][1], aim0);
                    are1 = __ffma2_rn(w[(mm + t + 1) % (R + K - 1)], c[t][0], are1);
                    aim1 = __ffma2_rn(w[(mm + t + 1) % (R + K - 1)], c[t][1], aim1);
                } else {
                    are0 = __ffma2_rn(are0, c[0][0], c[0][1]);
                    aim0 = __ffma2_rn(aim0, c[0][0], c[0][1]);
                    are1 = __ffma2_rn(are1, c[0][0], c[0][1]);
                    aim1 = __ffma2_rn(aim1, c[0][0], c[0][1]);
                }
            }
            const float2 p0 = __ffma2_rn(aim0, aim0, __fmul2_rn(are0, are0)), p1 = __ffma2_rn(aim1, aim1, __fmul2_rn(are1, are1));
            s += p0.x + p0.y + p1.x + p1.y;
        }
        w[it & 7].x += s * 1e-30f;   // keep the loop body from being hoisted
    }
    if (s == 123.456f) out[threadIdx.x] = s;
}
int main() {
    float2* in; float* out;
    cudaMalloc(&in, 1 << 20); cudaMemset(in, 0, 1 << 20); cudaMalloc(&out, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int wps = 1; wps <= 4; wps *= 2)   // resident warps per scheduler (128-thread CTAs, 128 registers: at most 4)
    for (int mode = 0; mode < 2; ++mode) {
        const int grid = 148 * wps;
        float best = 1e9;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) kern<0><<<grid, 128>>>(in, out, iters);
            else if (mode == 1) kern<1><<<grid, 128>>>(in, out, iters);
            else kern<2><<<grid, 128>>>(in, out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        const double n = (double)grid * 128 * iters * R * (4.0 * K + 4);
        printf("warps/scheduler %d mode %d: %.3f ms, %.2f T packed thread-instr/s (plain-chain peak 18.25)\n", wps, mode, best, n / best / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
