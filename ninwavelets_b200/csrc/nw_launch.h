// Launch interface between the translation units of libnwcwt.so.  The kernels are
// instantiated per (precision, kernel group) in their own .cu files so that ptxas
// runs on them in parallel; nwcwt.cu (C ABI, plans) only sees these functions.
#pragma once
#include <cuda_runtime.h>
#include "nw_kernels.cuh"
#include "nw_kernels2.cuh"
#include "nw_kernels3.cuh"
#include "nw_kernels4.cuh"
#include "nw_resample.cuh"

namespace nw {
template <typename T> cudaError_t prepare_short();
template <typename T> cudaError_t prepare_passA();
template <typename T> cudaError_t prepare_passB();
template <typename T>
cudaError_t launch_short(const ShortParams<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s);
template <typename T>
cudaError_t launch_passA(int dir, const LongParams<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s);
template <typename T>
cudaError_t launch_passB(int dir, const LongParams<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s);
// fast long-row path (nw_kernels2.cuh); CFG = compiled launch shape (nw_plan.h CFG2_*), sp = StaticPlan id
// (0 = run-time plan); has_static_plan says whether the (pass, sp) kernel exists in shape CFG (pass 0 = A, 1 = B)
template <typename T, int CFG> cudaError_t prepare_long2();
template <typename T, int CFG> bool has_static_plan(int pass, int sp);
template <typename T, int CFG>
cudaError_t launch_passA2(int sp, const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s);
template <typename T, int CFG>
cudaError_t launch_passB2(int sp, const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s);
// fused short-row kernel on the packed engine (nw_kernels3.cuh); sp = StaticPlan id of the inverse transform
template <typename T> cudaError_t prepare_short2();
template <typename T> bool has_static_short2(int sp);
template <typename T>
cudaError_t launch_short2(int sp, const Short2Params<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s);
// the same kernel with the epoch reduction fused in (nw_kernels3.cuh: short2_epochs_body); kind 0 = mean power, 1 = ITC
template <typename T>
cudaError_t launch_short2_epochs(int sp, int kind, const Short2Params<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s);
// fused short-row kernel with resampled rows (nw_kernels4.cuh): abs / power output
template <typename T> cudaError_t prepare_short3();
template <typename T>
cudaError_t launch_short3(const Short3Params<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s);
// interpolation kernel of the resampled rows (nw_resample.cuh); K = taps, mode = OUT_ABS / OUT_POWER
template <typename T> cudaError_t prepare_resample();
template <typename T> bool has_resample(int K);
template <typename T>
cudaError_t launch_resample(int K, int mode, const ResampleParams<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s);
// vector interpolation kernel (nw_resample.cuh: resample_vec_body): even K, PQ outputs per access, run length R
static const int RSV_THREADS = 128;
template <typename T, int PQ, int MODE> cudaError_t prepare_resample_vec();
template <typename T, int PQ, int MODE> bool has_resample_vec(int K);
template <typename T, int PQ, int MODE>
cudaError_t launch_resample_vec(int K, int R, const ResampleParams<T>& P, dim3 grid, size_t smem, cudaStream_t s);
// direct interpolation kernel (nw_resample.cuh: resample_dir_body): no shared memory, run length 8 (K <= 8) or 4
template <typename T, int PQ, int MODE>
cudaError_t launch_resample_dir(int K, const ResampleParams<T>& P, dim3 grid, cudaStream_t s);
}  // namespace nw
