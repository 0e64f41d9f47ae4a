// Vector interpolation kernels of the resampled rows (nw_resample.cuh: resample_vec_body), instantiated for NW_REAL,
// NW_RSV_PQ outputs per access, output mode NW_RSV_MODE and the even tap counts in NW_RSV_TAPS (X-macro list), each for
// the run lengths 8 and 4.
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {

template <typename T, int K, int R, int PQ, int MODE>
__global__ void __launch_bounds__(RSV_THREADS, rsv_min_ctas(K, R, PQ)) nwcwt_resample_vec_kernel(const __grid_constant__ ResampleParams<T> P) {
    extern __shared__ __align__(16) char nw_smem[];
    resample_vec_body<T, K, R, PQ, MODE>(P, nw_smem, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
}

#ifndef RSD_MINCTAS
#define RSD_MINCTAS 4
#endif
template <typename T, int K, int R, int PQ, int MODE>
__global__ void __launch_bounds__(RSV_THREADS, RSD_MINCTAS) nwcwt_resample_dir_kernel(const __grid_constant__ ResampleParams<T> P) {
    __shared__ __align__(16) cx<T> strip[(RSV_THREADS / 32) * RsDirGeo<K, R>::DEPTH * RsDirGeo<K, R>::SLOTS];
    resample_dir_body<T, K, R, PQ, MODE>(P, (char*)strip, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
}

#define NW_RSD_RUN(k) case k: nwcwt_resample_dir_kernel<NW_REAL, k, rs_dir_run(k), NW_RSV_PQ, NW_RSV_MODE><<<grid, RSV_THREADS, 0, s>>>(P); return cudaGetLastError();
#define NW_RSV_K(k, r) nwcwt_resample_vec_kernel<NW_REAL, k, r, NW_RSV_PQ, NW_RSV_MODE>
#define NW_RSV_PREP1(k, r) \
    { cudaError_t e = cudaFuncSetAttribute(NW_RSV_K(k, r), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); if (e != cudaSuccess) return e; }
#define NW_RSV_PREP(k) NW_RSV_PREP1(k, 8) NW_RSV_PREP1(k, 4)
#define NW_RSV_HAS(k) if (K == k) return true;
#define NW_RSV_RUN1(k, r) case r: NW_RSV_K(k, r)<<<grid, RSV_THREADS, smem, s>>>(P); return cudaGetLastError();
#define NW_RSV_RUN(k) case k: \
    switch (R) { NW_RSV_RUN1(k, 8) NW_RSV_RUN1(k, 4) default: return cudaErrorInvalidValue; }

template <> cudaError_t prepare_resample_vec<NW_REAL, NW_RSV_PQ, NW_RSV_MODE>() {
    NW_RSV_TAPS(NW_RSV_PREP)
    return cudaSuccess;
}
template <> bool has_resample_vec<NW_REAL, NW_RSV_PQ, NW_RSV_MODE>(int K) {
    NW_RSV_TAPS(NW_RSV_HAS)
    return false;
}
template <>
cudaError_t launch_resample_vec<NW_REAL, NW_RSV_PQ, NW_RSV_MODE>(int K, int R, const ResampleParams<NW_REAL>& P, dim3 grid, size_t smem, cudaStream_t s) {
    switch (K) {
        NW_RSV_TAPS(NW_RSV_RUN)
        default: return cudaErrorInvalidValue;
    }
}
template <>
cudaError_t launch_resample_dir<NW_REAL, NW_RSV_PQ, NW_RSV_MODE>(int K, const ResampleParams<NW_REAL>& P, dim3 grid, cudaStream_t s) {
    switch (K) {
        NW_RSV_TAPS(NW_RSD_RUN)
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace nw
