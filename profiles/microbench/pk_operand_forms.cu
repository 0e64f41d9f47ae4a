#include <cuda_runtime.h>
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__global__ void k1(const float2* in, float2* out) {   // sub via negated operand
    float2 a = in[threadIdx.x], b = in[threadIdx.x + 32];
    out[threadIdx.x] = __fadd2_rn(a, neg2(b));
}
__global__ void k2(const float2* in, float2* out) {   // fma with negated multiplicand
    float2 a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
    out[threadIdx.x] = __ffma2_rn(neg2(a), b, c);
}
__global__ void k3(const float2* in, float2* out) {   // fma with negated addend
    float2 a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
    out[threadIdx.x] = __ffma2_rn(a, b, neg2(c));
}
__global__ void k4(const float2* in, const float* tw, float2* out) {   // broadcast scalar twiddle
    float2 a = in[threadIdx.x], b = in[threadIdx.x + 32];
    float wr = tw[threadIdx.x], wi = tw[threadIdx.x + 32];
    float2 WR = make_float2(wr, wr), WI = make_float2(wi, wi);
    float2 re = __ffma2_rn(a, WR, __fmul2_rn(neg2(b), WI));
    float2 im = __ffma2_rn(a, WI, __fmul2_rn(b, WR));
    out[threadIdx.x] = re; out[threadIdx.x + 32] = im;
}
__global__ void k5(const float4* in, float2* out) {   // plain float2 operators: does nvcc pack by itself?
    float4 a = in[threadIdx.x], b = in[threadIdx.x + 32];
    out[threadIdx.x] = make_float2(a.x + b.x, a.y + b.y);
    out[threadIdx.x + 32] = make_float2(a.z * b.z + a.x, a.w * b.w + a.y);
}
