// Fused short-row kernel (see nw_kernels.cuh: short_body), instantiated for NW_REAL.
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {
template <typename T>
__global__ void __launch_bounds__(512) nwcwt_short_kernel(const __grid_constant__ ShortParams<T> P) {
    extern __shared__ __align__(16) char nw_smem[];
    short_body<T>(P, nw_smem, blockIdx.x, threadIdx.x, blockDim.x);
}
template <> cudaError_t prepare_short<NW_REAL>() {
    return cudaFuncSetAttribute(nwcwt_short_kernel<NW_REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
}
template <>
cudaError_t launch_short<NW_REAL>(const ShortParams<NW_REAL>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    nwcwt_short_kernel<NW_REAL><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
}  // namespace nw
