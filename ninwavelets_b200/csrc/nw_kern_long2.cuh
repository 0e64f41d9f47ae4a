// Fast long-row kernels (packed in-place engine; nw_kernels2.cuh), instantiated for NW_REAL, one launch
// shape NW_CFG (nw_plan.h: CFG2_*) and the compile-time plans listed in NW_SP_A / NW_SP_B (X-macro lists of
// StaticPlan ids; the run-time plan, id 0, is always there).
#include "nw_launch.h"
#include "nw_plan.h"

#ifndef NW_SP_A
#define NW_SP_A(X)
#endif
#ifndef NW_SP_B
#define NW_SP_B(X)
#endif

namespace nw {
template <int CFG> struct Cfg2;
// registers per thread so that (threads x CTAs) of nw_plan.h's CFG2 table stay resident
template <> struct Cfg2<0> { static const int maxreg = NW_CFG0_MAXREG; };   // 256 x 3 (fp64: 168 registers, 128 x 3)
template <> struct Cfg2<1> { static const int maxreg = 96; };               // 224 x 3
template <> struct Cfg2<2> { static const int maxreg = 96; };               // 128 x 5
#ifndef NW_CFG3_MAXREG
#define NW_CFG3_MAXREG 96
#endif
template <> struct Cfg2<3> { static const int maxreg = NW_CFG3_MAXREG; };   //  64 x 10 (128 registers, 64 x 8: cfg2 6.82 vs 6.78 ms)

template <typename T, int CFG, int SP>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passA2_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passA2_body<T, SP, false>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <typename T, int CFG, int SP>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passA2n_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passA2_body<T, SP, true>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <typename T, int CFG>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passA2f_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passA2f_body<T>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <typename T, int CFG>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passB2f_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passB2_body<T, OUT_CWT, 0, -1>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}
template <typename T, int MODE, int CFG, int SP>
__global__ void __maxnreg__(Cfg2<CFG>::maxreg) nwcwt_passB2_kernel(const __grid_constant__ Long2Params<T> P) {
    extern __shared__ __align__(32) char nw_smem[];
    passB2_body<T, MODE, SP>(P, nw_smem, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

template <typename T, int CFG, int SP> static cudaError_t prepA() {
    cudaError_t e = cudaFuncSetAttribute(nwcwt_passA2n_kernel<T, CFG, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_passA2_kernel<T, CFG, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
}
template <typename T, int CFG, int SP> static cudaError_t prepB() {
    const int v = (int)SMEM_MAX;
    cudaError_t e = cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_CWT, CFG, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_ABS, CFG, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(nwcwt_passB2_kernel<T, OUT_POWER, CFG, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, v);
}
template <typename T, int CFG, int SP> static cudaError_t runA(const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.narrow) nwcwt_passA2n_kernel<T, CFG, SP><<<grid, nthr, smem, s>>>(P);
    else nwcwt_passA2_kernel<T, CFG, SP><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}
template <typename T, int CFG, int SP> static cudaError_t runB(const Long2Params<T>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (P.out_mode == OUT_POWER) nwcwt_passB2_kernel<T, OUT_POWER, CFG, SP><<<grid, nthr, smem, s>>>(P);
    else if (P.out_mode == OUT_ABS) nwcwt_passB2_kernel<T, OUT_ABS, CFG, SP><<<grid, nthr, smem, s>>>(P);
    else nwcwt_passB2_kernel<T, OUT_CWT, CFG, SP><<<grid, nthr, smem, s>>>(P);
    return cudaGetLastError();
}

#define NW_PREP_A(id) { cudaError_t e = prepA<NW_REAL, NW_CFG, id>(); if (e != cudaSuccess) return e; }
#define NW_PREP_B(id) { cudaError_t e = prepB<NW_REAL, NW_CFG, id>(); if (e != cudaSuccess) return e; }
#define NW_HAS(id) if (sp == id) return true;
#define NW_RUN_A(id) case id: return runA<NW_REAL, NW_CFG, id>(P, grid, nthr, smem, s);
#define NW_RUN_B(id) case id: return runB<NW_REAL, NW_CFG, id>(P, grid, nthr, smem, s);

template <> cudaError_t prepare_long2<NW_REAL, NW_CFG>() {
    { cudaError_t e = cudaFuncSetAttribute(nwcwt_passA2f_kernel<NW_REAL, NW_CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); if (e != cudaSuccess) return e; }
    { cudaError_t e = cudaFuncSetAttribute(nwcwt_passB2f_kernel<NW_REAL, NW_CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); if (e != cudaSuccess) return e; }
    NW_PREP_A(0) NW_SP_A(NW_PREP_A)
    NW_PREP_B(0) NW_SP_B(NW_PREP_B)
    return cudaSuccess;
}
template <> bool has_static_plan<NW_REAL, NW_CFG>(int pass, int sp) {
    if (sp == 0) return true;
    if (pass == 0) { NW_SP_A(NW_HAS) } else { NW_SP_B(NW_HAS) }
    return false;
}
template <>
cudaError_t launch_passA2<NW_REAL, NW_CFG>(int sp, const Long2Params<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (sp == -2) {   // forward transform
        nwcwt_passA2f_kernel<NW_REAL, NW_CFG><<<grid, nthr, smem, s>>>(P);
        return cudaGetLastError();
    }
    switch (sp) {
        NW_SP_A(NW_RUN_A)
        default: return runA<NW_REAL, NW_CFG, 0>(P, grid, nthr, smem, s);
    }
}
template <>
cudaError_t launch_passB2<NW_REAL, NW_CFG>(int sp, const Long2Params<NW_REAL>& P, dim3 grid, int nthr, size_t smem, cudaStream_t s) {
    if (sp == -2) {   // forward transform
        nwcwt_passB2f_kernel<NW_REAL, NW_CFG><<<grid, nthr, smem, s>>>(P);
        return cudaGetLastError();
    }
    switch (sp) {
        NW_SP_B(NW_RUN_B)
        default: return runB<NW_REAL, NW_CFG, 0>(P, grid, nthr, smem, s);
    }
}
}  // namespace nw
