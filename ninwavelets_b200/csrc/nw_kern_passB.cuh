// Long rows, pass B (bulk tile load, row transforms, epilogue), instantiated for NW_REAL.
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {
template <typename T, int DIR>
__global__ void __launch_bounds__(512) nwcwt_passB_kernel(const __grid_constant__ LongParams<T> P) {
    extern __shared__ __align__(16) char nw_smem[];
    passB_body<T, DIR>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <> cudaError_t prepare_passB<NW_REAL>() {
    cudaError_t e = cudaFuncSetAttribute(nwcwt_passB_kernel<NW_REAL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_passB_kernel<NW_REAL, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
}
template <>
cudaError_t launch_passB<NW_REAL>(int dir, const LongParams<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (dir > 0) nwcwt_passB_kernel<NW_REAL, 1><<<grid, nthr, smem, s>>>(P);
    else nwcwt_passB_kernel<NW_REAL, -1><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
}  // namespace nw
