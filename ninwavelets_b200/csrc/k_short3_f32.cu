// Fused short-row kernel with resampled rows (nw_kernels4.cuh), fp32.
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {
template <typename T>
__global__ void __launch_bounds__(256, 2) nwcwt_short3_kernel(const __grid_constant__ Short3Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    short3_body<T>(P, nw_smem, blockIdx.x, threadIdx.x, blockDim.x);
}
template <> cudaError_t prepare_short3<float>() {
    const int v = (int)SMEM_MAX;
    return cudaFuncSetAttribute(nwcwt_short3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}
template <>
cudaError_t launch_short3<float>(const Short3Params<float>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.out_mode != OUT_POWER && P.out_mode != OUT_ABS) return cudaErrorInvalidValue;
    nwcwt_short3_kernel<float><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
// fp64 rows stay on nw_kernels3.cuh (the planner builds no short-row groups for them)
template <> cudaError_t prepare_short3<double>() { return cudaSuccess; }
template <>
cudaError_t launch_short3<double>(const Short3Params<double>&, unsigned, int, size_t, cudaStream_t) { return cudaErrorInvalidValue; }
}  // namespace nw
