// Two-lane ("packed") real and complex values.
//
// The fast transform engine (nw_fft2.cuh) runs TWO interleaved sequences per
// thread: lane 0 and lane 1 of every value belong to two different sequences
// that share the butterfly's twiddles.  For float the two lanes sit in one
// 64-bit register pair and every add / mul / fma is one FADD2 / FMUL2 / FFMA2
// (sm_100a packed fp32: one issue slot for two lanes, sign flips and scalar
// broadcasts are operand modifiers, see profiles/r01/ffma2_microbench.md);
// for double the lanes are two scalars.  The host instantiation (tests/emul
// only) uses plain scalar arithmetic.
#pragma once
#include "nw_common.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

namespace nw {

template <typename T> struct pk;

#if defined(__CUDA_ARCH__)
template <> struct pk<float> {
    float2 v;
};
NW_D pk<float> pk_make(float a, float b) { pk<float> r; r.v = make_float2(a, b); return r; }
NW_D pk<float> pk_bcast(float a) { pk<float> r; r.v = make_float2(a, a); return r; }
NW_D float pk_lo(pk<float> a) { return a.v.x; }
NW_D float pk_hi(pk<float> a) { return a.v.y; }
NW_D pk<float> operator-(pk<float> a) { pk<float> r; r.v = make_float2(-a.v.x, -a.v.y); return r; }
NW_D pk<float> operator+(pk<float> a, pk<float> b) { pk<float> r; r.v = __fadd2_rn(a.v, b.v); return r; }
NW_D pk<float> operator-(pk<float> a, pk<float> b) { pk<float> r; r.v = __fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)); return r; }
NW_D pk<float> operator*(pk<float> a, pk<float> b) { pk<float> r; r.v = __fmul2_rn(a.v, b.v); return r; }
NW_D pk<float> operator*(pk<float> a, float s) { pk<float> r; r.v = __fmul2_rn(a.v, make_float2(s, s)); return r; }
// a * b + c
NW_D pk<float> pk_fma(pk<float> a, pk<float> b, pk<float> c) { pk<float> r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
NW_D pk<float> pk_fma(pk<float> a, float s, pk<float> c) { pk<float> r; r.v = __ffma2_rn(a.v, make_float2(s, s), c.v); return r; }
// c - a * b
NW_D pk<float> pk_fnma(pk<float> a, pk<float> b, pk<float> c) {
    pk<float> r; r.v = __ffma2_rn(make_float2(-a.v.x, -a.v.y), b.v, c.v); return r;
}
NW_D pk<float> pk_fnma(pk<float> a, float s, pk<float> c) {
    pk<float> r; r.v = __ffma2_rn(make_float2(-a.v.x, -a.v.y), make_float2(s, s), c.v); return r;
}
#else
template <> struct pk<float> {
    float a, b;
};
NW_HD pk<float> pk_make(float a, float b) { pk<float> r; r.a = a; r.b = b; return r; }
NW_HD pk<float> pk_bcast(float a) { return pk_make(a, a); }
NW_HD float pk_lo(pk<float> a) { return a.a; }
NW_HD float pk_hi(pk<float> a) { return a.b; }
NW_HD pk<float> operator-(pk<float> a) { return pk_make(-a.a, -a.b); }
NW_HD pk<float> operator+(pk<float> a, pk<float> b) { return pk_make(a.a + b.a, a.b + b.b); }
NW_HD pk<float> operator-(pk<float> a, pk<float> b) { return pk_make(a.a - b.a, a.b - b.b); }
NW_HD pk<float> operator*(pk<float> a, pk<float> b) { return pk_make(a.a * b.a, a.b * b.b); }
NW_HD pk<float> operator*(pk<float> a, float s) { return pk_make(a.a * s, a.b * s); }
NW_HD pk<float> pk_fma(pk<float> a, pk<float> b, pk<float> c) { return pk_make(fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)); }
NW_HD pk<float> pk_fma(pk<float> a, float s, pk<float> c) { return pk_make(fmaf(a.a, s, c.a), fmaf(a.b, s, c.b)); }
NW_HD pk<float> pk_fnma(pk<float> a, pk<float> b, pk<float> c) { return pk_make(fmaf(-a.a, b.a, c.a), fmaf(-a.b, b.b, c.b)); }
NW_HD pk<float> pk_fnma(pk<float> a, float s, pk<float> c) { return pk_make(fmaf(-a.a, s, c.a), fmaf(-a.b, s, c.b)); }
#endif

template <> struct pk<double> {
    double a, b;
};
NW_HD pk<double> pk_make(double a, double b) { pk<double> r; r.a = a; r.b = b; return r; }
NW_HD pk<double> pk_bcast(double a) { return pk_make(a, a); }
NW_HD double pk_lo(pk<double> a) { return a.a; }
NW_HD double pk_hi(pk<double> a) { return a.b; }
NW_HD pk<double> operator-(pk<double> a) { return pk_make(-a.a, -a.b); }
NW_HD pk<double> operator+(pk<double> a, pk<double> b) { return pk_make(a.a + b.a, a.b + b.b); }
NW_HD pk<double> operator-(pk<double> a, pk<double> b) { return pk_make(a.a - b.a, a.b - b.b); }
NW_HD pk<double> operator*(pk<double> a, pk<double> b) { return pk_make(a.a * b.a, a.b * b.b); }
NW_HD pk<double> operator*(pk<double> a, double s) { return pk_make(a.a * s, a.b * s); }
NW_HD pk<double> pk_fma(pk<double> a, pk<double> b, pk<double> c) { return pk_make(fma(a.a, b.a, c.a), fma(a.b, b.b, c.b)); }
NW_HD pk<double> pk_fma(pk<double> a, double s, pk<double> c) { return pk_make(fma(a.a, s, c.a), fma(a.b, s, c.b)); }
NW_HD pk<double> pk_fnma(pk<double> a, pk<double> b, pk<double> c) { return pk_make(fma(-a.a, b.a, c.a), fma(-a.b, b.b, c.b)); }
NW_HD pk<double> pk_fnma(pk<double> a, double s, pk<double> c) { return pk_make(fma(-a.a, s, c.a), fma(-a.b, s, c.b)); }

// ---- two complex numbers, lane-wise (structure of arrays) ----------------------------
template <typename T> struct alignas(4 * sizeof(T)) cx2 {
    pk<T> re, im;
};
template <typename T> NW_HD cx2<T> mk2(pk<T> re, pk<T> im) { cx2<T> r; r.re = re; r.im = im; return r; }
template <typename T> NW_HD cx2<T> mk2(cx<T> l0, cx<T> l1) { return mk2<T>(pk_make(l0.x, l1.x), pk_make(l0.y, l1.y)); }
template <typename T> NW_HD cx2<T> zero2() { return mk2<T>(pk_bcast((T)0), pk_bcast((T)0)); }
template <typename T> NW_HD cx<T> lane0(cx2<T> v) { return mk<T>(pk_lo(v.re), pk_lo(v.im)); }
template <typename T> NW_HD cx<T> lane1(cx2<T> v) { return mk<T>(pk_hi(v.re), pk_hi(v.im)); }
template <typename T> NW_HD cx2<T> operator+(cx2<T> a, cx2<T> b) { return mk2<T>(a.re + b.re, a.im + b.im); }
template <typename T> NW_HD cx2<T> operator-(cx2<T> a, cx2<T> b) { return mk2<T>(a.re - b.re, a.im - b.im); }
// a + DIR*i*b  and  a - DIR*i*b   (i*b = (-b.im, b.re))
template <int DIR, typename T> NW_HD cx2<T> add_rot(cx2<T> a, cx2<T> b) {
    return DIR > 0 ? mk2<T>(a.re - b.im, a.im + b.re) : mk2<T>(a.re + b.im, a.im - b.re);
}
template <int DIR, typename T> NW_HD cx2<T> sub_rot(cx2<T> a, cx2<T> b) { return add_rot<-DIR, T>(a, b); }
// DIR*i*v
template <int DIR, typename T> NW_HD cx2<T> rot2(cx2<T> v) { return DIR > 0 ? mk2<T>(-v.im, v.re) : mk2<T>(v.im, -v.re); }
// both lanes times one scalar complex w (a twiddle shared by the two sequences)
template <typename T> NW_HD cx2<T> cmul_s(cx2<T> v, cx<T> w) {
    return mk2<T>(pk_fnma(v.im, w.y, v.re * w.x), pk_fma(v.im, w.x, v.re * w.y));
}
// lane-wise times a packed complex (different factor per lane)
template <typename T> NW_HD cx2<T> cmul_p(cx2<T> v, cx2<T> w) {
    return mk2<T>(pk_fnma(v.im, w.im, v.re * w.re), pk_fma(v.im, w.re, v.re * w.im));
}
template <typename T> NW_HD cx2<T> scale2(cx2<T> v, T s) { return mk2<T>(v.re * s, v.im * s); }
// v * (c + DIR*i*s) with compile-time-known c, s
template <int DIR, typename T> NW_HD cx2<T> cmul_k(cx2<T> v, T c, T s) {
    return DIR > 0 ? mk2<T>(pk_fnma(v.im, s, v.re * c), pk_fma(v.re, s, v.im * c))
                   : mk2<T>(pk_fma(v.im, s, v.re * c), pk_fnma(v.re, s, v.im * c));
}

// a 4-value unit holding two plain complex numbers {re0, im0, re1, im1} -> lane-packed
template <typename T> NW_HD cx2<T> raw_to_packed(cx2<T> u) {
    return mk2<T>(pk_make(pk_lo(u.re), pk_lo(u.im)), pk_make(pk_hi(u.re), pk_hi(u.im)));
}

}  // namespace nw
