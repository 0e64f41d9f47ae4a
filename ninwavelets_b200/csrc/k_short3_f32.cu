// Fused short-row kernel with resampled rows (nw_kernels4.cuh), fp32.
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {
template <typename T, int MODE>
__global__ void __launch_bounds__(512) nwcwt_short3_kernel(const __grid_constant__ Short3Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    short3_body<T, MODE>(P, nw_smem, blockIdx.x, threadIdx.x, blockDim.x);
}
template <> cudaError_t prepare_short3<float>() {
    const int v = (int)SMEM_MAX;
    cudaError_t e = cudaFuncSetAttribute(nwcwt_short3_kernel<float, OUT_ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_short3_kernel<float, OUT_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}
template <>
cudaError_t launch_short3<float>(const Short3Params<float>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.out_mode == OUT_POWER) nwcwt_short3_kernel<float, OUT_POWER><<<grid, nthr, smem, s>>>(P);
    else if (P.out_mode == OUT_ABS) nwcwt_short3_kernel<float, OUT_ABS><<<grid, nthr, smem, s>>>(P);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
// fp64 rows stay on nw_kernels3.cuh (the planner builds no short-row groups for them)
template <> cudaError_t prepare_short3<double>() { return cudaSuccess; }
template <>
cudaError_t launch_short3<double>(const Short3Params<double>&, unsigned, int, size_t, cudaStream_t) { return cudaErrorInvalidValue; }
}  // namespace nw
