// Fast long-row kernels (packed in-place engine; nw_kernels2.cuh), instantiated for NW_REAL and
// the launch shapes NW_CFG_LIST (nw_plan.h: CFG2_MAXTHR / CFG2_MINCTA give the register budget).
#include "nw_launch.h"
#include "nw_plan.h"

namespace nw {
template <int CFG> struct Cfg2;
// registers per thread so that (threads x CTAs) of nw_plan.h's CFG2 table stay resident
template <> struct Cfg2<0> { static const int maxreg = NW_CFG0_MAXREG; };   // 256 x 3 (fp64: 256 x 2)
template <> struct Cfg2<1> { static const int maxreg = 96; };               // 224 x 3
template <> struct Cfg2<2> { static const int maxreg = 96; };               // 128 x 5
template <> struct Cfg2<3> { static const int maxreg = 128; };              //  64 x 8

template <typename T, int CFG>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passA2_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passA2_body<T>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <typename T, int MODE, int CFG>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passB2_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passB2_body<T, MODE>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

template <typename T, int CFG> static cudaError_t prepare_cfg() {
    const int v = (int)SMEM_MAX;
    cudaError_t e = cudaFuncSetAttribute(nwcwt_passA2_kernel<T, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_CWT, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_ABS, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_POWER, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}
template <typename T, int CFG> static cudaError_t launchA_cfg(const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    nwcwt_passA2_kernel<T, CFG><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
template <typename T, int CFG> static cudaError_t launchB_cfg(const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.out_mode == OUT_POWER) nwcwt_passB2_kernel<T, OUT_POWER, CFG><<<grid, nthr, smem, s>>>(P);
    else if (P.out_mode == OUT_ABS) nwcwt_passB2_kernel<T, OUT_ABS, CFG><<<grid, nthr, smem, s>>>(P);
    else nwcwt_passB2_kernel<T, OUT_CWT, CFG><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}

#define NW_CFG_CASE(c, call) case c: return call
template <> cudaError_t prepare_long2<NW_REAL, NW_CFG>() { return prepare_cfg<NW_REAL, NW_CFG>(); }
template <>
cudaError_t launch_passA2<NW_REAL, NW_CFG>(const Long2Params<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    return launchA_cfg<NW_REAL, NW_CFG>(P, grid, nthr, smem, s);
}
template <>
cudaError_t launch_passB2<NW_REAL, NW_CFG>(const Long2Params<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    return launchB_cfg<NW_REAL, NW_CFG>(P, grid, nthr, smem, s);
}
}  // namespace nw
