// Interpolation kernels of the resampled rows (nw_resample.cuh), instantiated for NW_REAL and the tap counts in
// NW_RS_TAPS (X-macro list).
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {

template <typename T, int K, int MODE>
__global__ void __launch_bounds__(384) nwcwt_resample_kernel(const __grid_constant__ ResampleParams<T> P) {
    extern __shared__ __align__(16) char nw_smem[];
    resample_body<T, K, ResampleRun<T>::R, MODE>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

#define NW_RS_PREP(k) \
    { cudaError_t e = cudaFuncSetAttribute(nwcwt_resample_kernel<NW_REAL, k, OUT_POWER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); if (e != cudaSuccess) return e; \
      e = cudaFuncSetAttribute(nwcwt_resample_kernel<NW_REAL, k, OUT_ABS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); if (e != cudaSuccess) return e; }
#define NW_RS_HAS(k) if (K == k) return true;
#define NW_RS_RUN(k) case k: \
    if (mode == OUT_POWER) nwcwt_resample_kernel<NW_REAL, k, OUT_POWER><<<grid, nthr, smem, s>>>(P); \
    else nwcwt_resample_kernel<NW_REAL, k, OUT_ABS><<<grid, nthr, smem, s>>>(P); \
    return cudaGetLastError();

template <> cudaError_t prepare_resample<NW_REAL>() {
    NW_RS_TAPS(NW_RS_PREP)
    return cudaSuccess;
}
template <> bool has_resample<NW_REAL>(int K) {
    NW_RS_TAPS(NW_RS_HAS)
    return false;
}
template <>
cudaError_t launch_resample<NW_REAL>(int K, int mode, const ResampleParams<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    switch (K) {
        NW_RS_TAPS(NW_RS_RUN)
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace nw
