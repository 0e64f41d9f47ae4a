// Fused short-row kernel on the packed engine (nw_kernels3.cuh), instantiated for NW_REAL.
#include "nw_launch.h"
#include "nw_plan.h"

#ifndef NW_SP_S
#define NW_SP_S(X)
#endif

namespace nw {
template <typename T, int MODE, int SP>
__global__ void __maxnreg__(NW_S2_MAXREG) nwcwt_short2_kernel(const __grid_constant__ Short2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    short2_body<T, MODE, SP>(P, nw_smem, blockIdx.x, threadIdx.x, blockDim.x);
}
template <typename T, int KIND, int SP>
__global__ void __maxnreg__(NW_S2_MAXREG) nwcwt_short2_epochs_kernel(const __grid_constant__ Short2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    short2_epochs_body<T, KIND, SP>(P, nw_smem, blockIdx.x, threadIdx.x, blockDim.x);
}
template <typename T, int SP> static cudaError_t prepS() {
    const int v = (int)SMEM_MAX;
    cudaError_t e = cudaFuncSetAttribute(nwcwt_short2_kernel<T, OUT_CWT, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_short2_kernel<T, OUT_ABS, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_short2_kernel<T, OUT_POWER, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_short2_epochs_kernel<T, 0, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_short2_epochs_kernel<T, 1, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}
template <typename T, int SP> static cudaError_t runSE(int kind, const Short2Params<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    if (kind == 0) nwcwt_short2_epochs_kernel<T, 0, SP><<<grid, nthr, smem, s>>>(P);
    else nwcwt_short2_epochs_kernel<T, 1, SP><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
template <typename T, int SP> static cudaError_t runS(const Short2Params<T>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.out_mode == OUT_POWER) nwcwt_short2_kernel<T, OUT_POWER, SP><<<grid, nthr, smem, s>>>(P);
    else if (P.out_mode == OUT_ABS) nwcwt_short2_kernel<T, OUT_ABS, SP><<<grid, nthr, smem, s>>>(P);
    else nwcwt_short2_kernel<T, OUT_CWT, SP><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
#define NW_PREP_S(id) { cudaError_t e = prepS<NW_REAL, id>(); if (e != cudaSuccess) return e; }
#define NW_HAS_S(id) if (sp == id) return true;
#define NW_RUN_S(id) case id: return runS<NW_REAL, id>(P, grid, nthr, smem, s);
#define NW_RUN_SE(id) case id: return runSE<NW_REAL, id>(kind, P, grid, nthr, smem, s);
template <> cudaError_t prepare_short2<NW_REAL>() {
    NW_PREP_S(0) NW_SP_S(NW_PREP_S)
    return cudaSuccess;
}
template <> bool has_static_short2<NW_REAL>(int sp) {
    if (sp == 0) return true;
    NW_SP_S(NW_HAS_S)
    return false;
}
template <>
cudaError_t launch_short2<NW_REAL>(int sp, const Short2Params<NW_REAL>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    switch (sp) {
        NW_SP_S(NW_RUN_S)
        default: return runS<NW_REAL, 0>(P, grid, nthr, smem, s);
    }
}
template <>
cudaError_t launch_short2_epochs<NW_REAL>(int sp, int kind, const Short2Params<NW_REAL>& P, unsigned grid, int nthr, size_t smem, cudaStream_t s) {
    switch (sp) {
        NW_SP_S(NW_RUN_SE)
        default: return runSE<NW_REAL, 0>(kind, P, grid, nthr, smem, s);
    }
}
}  // namespace nw
