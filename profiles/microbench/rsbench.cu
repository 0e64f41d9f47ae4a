// A/B timing of the interpolation kernel variants (nw_resample.cuh) on synthetic decimated rows.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../../ninwavelets_b200/csrc -o rsbench rsbench.cu
#include <cstdio>
#ifndef MINCTA
#define MINCTA 5
#endif
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "nw_common.h"
#include "nw_pk.cuh"
#include "nw_kernels.cuh"
#include "nw_resample.cuh"
namespace nw {
template <typename T> NW_D void rs_fma1(cx<T>& acc, cx<T> w, T c) { rs_fma(acc, w, c); }
#include "rs_v1_body.inc"
}
using namespace nw;
template <int K, int R, int PQ>
__global__ void __launch_bounds__(128, 5) k_v1(const __grid_constant__ ResampleParams<float> P) {
    extern __shared__ __align__(16) char sm[];
    resample_vec1_body<float, K, R, PQ, OUT_POWER>(P, sm, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <int K, int R, int PQ>
__global__ void __launch_bounds__(128, MINCTA) k_v2(const __grid_constant__ ResampleParams<float> P) {
    extern __shared__ __align__(16) char sm[];
    resample_vec_body<float, K, R, PQ, OUT_POWER>(P, sm, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    return best;
}
template <int K, int R, int PQ> void run(int D, int WR, int WP, long long N, int rows, float* out, cx<float>* y, float* coefq) {
    const int M = (int)(N / D);
    ResampleParams<float> P; memset(&P, 0, sizeof(P));
    P.y = y; P.ystride = M; P.out = out; P.N = N; P.M = M; P.D = D; P.coefq = coefq; P.F = rows; P.F_out = rows; P.row0 = 0;
    const double outs = (double)rows * N;
    {   // v1: CTA of 128 threads, C = 128 R
        const int C = 128 * R; const unsigned tiles = (M + C - 1) / C;
        const size_t ys = ((size_t)C + K + (C + K) / R + 2) * 8, tile = (size_t)128 * (R * D + PQ) * 4, smem = ys > tile ? ys : tile;
        P.dRD = make_fastdiv(R * D / PQ);
        cudaFuncSetAttribute(k_v1<K, R, PQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float ms = timeit([&] { k_v1<K, R, PQ><<<dim3(tiles, rows), 128, smem>>>(P); });
        printf("D=%2d K=%d R=%2d PQ=%d v1            smem %6zu : %7.3f ms  %5.2f ps/out  %6.0f GB/s\n", D, K, R, PQ, smem, ms, ms * 1e9 / outs, outs * 4 / ms / 1e6);
    }
    {
        const size_t CG = 32 * R, ys = (CG + K + (CG + K) / R + 2) * 8, tile = (size_t)32 * (R * D + PQ) * 4;
        const size_t gb = ((ys > tile ? ys : tile) + 15) / 16 * 16, smem = gb * WR;
        const unsigned items = (unsigned)((M + CG - 1) / CG);
        P.WR = WR; P.WP = WP; P.RS = (int)gb; P.dRD = make_fastdiv(R * D / PQ); P.dGT = make_fastdiv(items); P.nrows = rows;
        cudaFuncSetAttribute(k_v2<K, R, PQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int cps : {4, 5, 6}) {
            const long long ctas = ((long long)items * rows + WR - 1) / WR;
            const unsigned grid = (unsigned)std::min<long long>(ctas, 148LL * cps);
            float ms = timeit([&] { k_v2<K, R, PQ><<<grid, 128, smem>>>(P); });
            printf("D=%2d K=%d R=%2d PQ=%d v2 WR=%d WP=%d smem %6zu grid 148x%d: %7.3f ms  %5.2f ps/out  %6.0f GB/s\n", D, K, R, PQ, WR, WP, smem, cps, ms, ms * 1e9 / outs, outs * 4 / ms / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}
int main() {
    const long long N = 600000;
    const int rows = 640;
    float* out; cudaMalloc(&out, (size_t)rows * N * 4);
    cx<float>* y; cudaMalloc(&y, (size_t)rows * (N / 2) * 8);
    cudaMemset(y, 0, (size_t)rows * (N / 2) * 8);
    std::vector<float> c(64 * 16, 0.125f);
    float* coefq; cudaMalloc(&coefq, c.size() * 4); cudaMemcpy(coefq, c.data(), c.size() * 4, cudaMemcpyHostToDevice);
    run<8, 8, 4>(8, 4, 1, N, rows, out, y, coefq);
    run<8, 8, 4>(8, 2, 2, N, rows, out, y, coefq);
    run<8, 16, 4>(8, 2, 2, N, rows, out, y, coefq);
    run<6, 8, 4>(8, 4, 1, N, rows, out, y, coefq);
    run<8, 8, 4>(32, 1, 4, N, rows, out, y, coefq);
    run<8, 4, 4>(32, 2, 2, N, rows, out, y, coefq);
    run<8, 8, 4>(16, 2, 2, N, rows, out, y, coefq);
    run<8, 4, 4>(12, 4, 1, N, rows, out, y, coefq);
    run<8, 8, 2>(6, 4, 1, N, rows, out, y, coefq);
    run<8, 16, 4>(4, 4, 1, N, rows, out, y, coefq);
    run<8, 8, 4>(4, 4, 1, N, rows, out, y, coefq);
    return 0;
}
