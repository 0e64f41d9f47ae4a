// Radix butterflies of the packed engine: an R-point DFT on R two-lane complex
// values held in registers, natural order in and out.
//   DIR = +1: kernel e^{+2 pi i n k / R} (inverse transform), DIR = -1: forward.
// 2, 3, 4, 5 are written out; 6, 10, 12, 15, 20 are prime-factor (Good-Thomas)
// compositions - no internal twiddles, the index maps are compile-time register
// renamings; 8 and 16 are Cooley-Tukey compositions with constant twiddles.
// Operation counts (packed instructions per butterfly): 2:4  3:12  4:16  5:36
// 6:36  8:56  10:92  12:84  15:168  16:160.
#pragma once
#include "nw_pk.cuh"

namespace nw {

template <typename T, int R, int DIR> struct B2;

template <typename T, int DIR> struct B2<T, 1, DIR> {
    static NW_HD void run(cx2<T>*) {}
};

template <typename T, int DIR> struct B2<T, 2, DIR> {
    static NW_HD void run(cx2<T>* v) {
        const cx2<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};

template <typename T, int DIR> struct B2<T, 3, DIR> {
    static NW_HD void run(cx2<T>* v) {
        const T s = (T)(DIR * 0.86602540378443864676372317075294);  // DIR * sin(2pi/3)
        const cx2<T> t1 = v[1] + v[2], t2 = v[1] - v[2];
        const cx2<T> m = mk2<T>(pk_fnma(t1.re, (T)0.5, v[0].re), pk_fnma(t1.im, (T)0.5, v[0].im));
        v[0] = v[0] + t1;
        // m +- i*s*t2
        v[1] = mk2<T>(pk_fnma(t2.im, s, m.re), pk_fma(t2.re, s, m.im));
        v[2] = mk2<T>(pk_fma(t2.im, s, m.re), pk_fnma(t2.re, s, m.im));
    }
};

template <typename T, int DIR> struct B2<T, 4, DIR> {
    static NW_HD void run(cx2<T>* v) {
        const cx2<T> t0 = v[0] + v[2], t1 = v[0] - v[2];
        const cx2<T> t2 = v[1] + v[3], t3 = v[1] - v[3];
        v[0] = t0 + t2;
        v[2] = t0 - t2;
        v[1] = add_rot<DIR, T>(t1, t3);
        v[3] = sub_rot<DIR, T>(t1, t3);
    }
};

template <typename T, int DIR> struct B2<T, 5, DIR> {
    static NW_HD void run(cx2<T>* v) {
        const T c1 = (T)0.30901699437494742410229341718282;    // cos(2pi/5)
        const T c2 = (T)-0.80901699437494742410229341718282;   // cos(4pi/5)
        const T s1 = (T)(DIR * 0.95105651629515357211643933337938);    // DIR * sin(2pi/5)
        const T s2 = (T)(DIR * 0.58778525229247312916870595463907);    // DIR * sin(4pi/5)
        const cx2<T> a1 = v[1] + v[4], b1 = v[1] - v[4];
        const cx2<T> a2 = v[2] + v[3], b2 = v[2] - v[3];
        const cx2<T> m1 = mk2<T>(pk_fma(a2.re, c2, pk_fma(a1.re, c1, v[0].re)), pk_fma(a2.im, c2, pk_fma(a1.im, c1, v[0].im)));
        const cx2<T> m2 = mk2<T>(pk_fma(a2.re, c1, pk_fma(a1.re, c2, v[0].re)), pk_fma(a2.im, c1, pk_fma(a1.im, c2, v[0].im)));
        // n1 = s1 b1 + s2 b2,  n2 = s2 b1 - s1 b2   (already carry DIR)
        const cx2<T> n1 = mk2<T>(pk_fma(b2.re, s2, b1.re * s1), pk_fma(b2.im, s2, b1.im * s1));
        const cx2<T> n2 = mk2<T>(pk_fnma(b2.re, s1, b1.re * s2), pk_fnma(b2.im, s1, b1.im * s2));
        v[0] = v[0] + a1 + a2;
        // m + i*n, m - i*n
        v[1] = mk2<T>(m1.re - n1.im, m1.im + n1.re);
        v[4] = mk2<T>(m1.re + n1.im, m1.im - n1.re);
        v[2] = mk2<T>(m2.re - n2.im, m2.im + n2.re);
        v[3] = mk2<T>(m2.re + n2.im, m2.im - n2.re);
    }
};

// ---- prime-factor composition R = RA * RB, gcd(RA, RB) = 1 --------------------------
//   input  n = (RB n1 + RA n2) mod R,   output k with k = k1 (mod RA), k = k2 (mod RB)
NW_HD constexpr int pfa_out_index(int RA, int RB, int k1, int k2) {
    for (int k = 0; k < RA * RB; ++k)
        if (k % RA == k1 && k % RB == k2) return k;
    return -1;
}

template <typename T, int RA, int RB, int DIR> struct PfaB2 {
    static NW_HD void run(cx2<T>* v) {
        constexpr int R = RA * RB;
        cx2<T> u[R];   // u[k1 * RB + n2]
#pragma unroll
        for (int n2 = 0; n2 < RB; ++n2) {
            cx2<T> t[RA];
#pragma unroll
            for (int n1 = 0; n1 < RA; ++n1) t[n1] = v[(RB * n1 + RA * n2) % R];
            B2<T, RA, DIR>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < RA; ++k1) u[k1 * RB + n2] = t[k1];
        }
#pragma unroll
        for (int k1 = 0; k1 < RA; ++k1) {
            cx2<T> t[RB];
#pragma unroll
            for (int n2 = 0; n2 < RB; ++n2) t[n2] = u[k1 * RB + n2];
            B2<T, RB, DIR>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < RB; ++k2) v[pfa_out_index(RA, RB, k1, k2)] = t[k2];
        }
    }
};

template <typename T, int DIR> struct B2<T, 6, DIR> : PfaB2<T, 2, 3, DIR> {};
template <typename T, int DIR> struct B2<T, 10, DIR> : PfaB2<T, 2, 5, DIR> {};
template <typename T, int DIR> struct B2<T, 12, DIR> : PfaB2<T, 3, 4, DIR> {};
template <typename T, int DIR> struct B2<T, 15, DIR> : PfaB2<T, 3, 5, DIR> {};
template <typename T, int DIR> struct B2<T, 20, DIR> : PfaB2<T, 4, 5, DIR> {};

// ---- Cooley-Tukey compositions with constant twiddles ---------------------------------
template <typename T, int DIR> struct B2<T, 8, DIR> {
    static NW_HD void run(cx2<T>* v) {
        const T h = (T)0.70710678118654752440084436210485;
        cx2<T> e[4] = {v[0], v[2], v[4], v[6]};
        cx2<T> o[4] = {v[1], v[3], v[5], v[7]};
        B2<T, 4, DIR>::run(e);
        B2<T, 4, DIR>::run(o);
        const cx2<T> w1 = cmul_k<DIR, T>(o[1], h, h);    // o1 * e^{DIR i pi/4}
        const cx2<T> w3 = cmul_k<DIR, T>(o[3], -h, h);   // o3 * e^{DIR i 3pi/4}
        v[0] = e[0] + o[0];
        v[4] = e[0] - o[0];
        v[1] = e[1] + w1;
        v[5] = e[1] - w1;
        v[2] = add_rot<DIR, T>(e[2], o[2]);
        v[6] = sub_rot<DIR, T>(e[2], o[2]);
        v[3] = e[3] + w3;
        v[7] = e[3] - w3;
    }
};

template <typename T, int DIR> struct B2<T, 16, DIR> {
    static NW_HD void run(cx2<T>* v) {
        // n = 4 n1 + n2, k = k1 + 4 k2
        const T c1 = (T)0.92387953251128675612818318939679;  // cos(pi/8)
        const T s1 = (T)0.38268343236508977172845998403040;  // sin(pi/8)
        const T h = (T)0.70710678118654752440084436210485;
        cx2<T> u[4][4];   // u[k1][n2]
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            cx2<T> t[4] = {v[n2], v[4 + n2], v[8 + n2], v[12 + n2]};
            B2<T, 4, DIR>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) u[k1][n2] = t[k1];
        }
        // u[k1][n2] *= w16^{DIR k1 n2}
        u[1][1] = cmul_k<DIR, T>(u[1][1], c1, s1);
        u[1][2] = cmul_k<DIR, T>(u[1][2], h, h);
        u[1][3] = cmul_k<DIR, T>(u[1][3], s1, c1);
        u[2][1] = cmul_k<DIR, T>(u[2][1], h, h);
        u[2][2] = rot2<DIR, T>(u[2][2]);
        u[2][3] = cmul_k<DIR, T>(u[2][3], -h, h);
        u[3][1] = cmul_k<DIR, T>(u[3][1], s1, c1);
        u[3][2] = cmul_k<DIR, T>(u[3][2], -h, h);
        u[3][3] = cmul_k<DIR, T>(u[3][3], -c1, -s1);   // w16^9
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            cx2<T> t[4] = {u[k1][0], u[k1][1], u[k1][2], u[k1][3]};
            B2<T, 4, DIR>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) v[k1 + 4 * k2] = t[k2];
        }
    }
};

// ---- generic Cooley-Tukey composition R = RA * RB with compile-time twiddles ----------------
//   n = RB n1 + n2,  k = k1 + RA k2:   X[k] = sum_n2 w_RB^{n2 k2} [ w_R^{n2 k1} sum_n1 x[n] w_RA^{n1 k1} ]
// cos / sin of 2 pi j / R evaluated at compile time (Taylor series on the first octant).
constexpr double ct_pi = 3.14159265358979323846264338327950288;
constexpr double ct_cos_small(double x) {   // |x| <= pi/4
    double t = 1, sum = 1;
    for (int i = 1; i <= 12; ++i) { t *= -x * x / ((2 * i - 1) * (2 * i)); sum += t; }
    return sum;
}
constexpr double ct_sin_small(double x) {
    double t = x, sum = x;
    for (int i = 1; i <= 12; ++i) { t *= -x * x / ((2 * i) * (2 * i + 1)); sum += t; }
    return sum;
}
// cos(2 pi j / R), j in [0, R)
constexpr double ct_cos(int j, int R) {
    j %= R;
    if (2 * j > R) j = R - j;                              // cos is even around pi
    if (4 * j > R) return -ct_cos(R - 2 * j, 2 * R);       // cos(x) = -cos(pi - x); (R - 2j)/(2R) turns
    if (8 * j > R) return ct_sin_small(2 * ct_pi * (R - 4 * j) / (4.0 * R));   // cos(x) = sin(pi/2 - x)
    return ct_cos_small(2 * ct_pi * j / R);
}
constexpr double ct_sin(int j, int R) {   // sin(x) = cos(x - pi/2) = cos(2 pi (4j - R) / (4R))
    return ct_cos(((4 * j - R) % (4 * R) + 4 * R) % (4 * R), 4 * R);
}
template <int R> struct CtTab {
    double c[R], s[R];
    constexpr CtTab() : c(), s() {
        for (int j = 0; j < R; ++j) { c[j] = ct_cos(j, R); s[j] = ct_sin(j, R); }
    }
};

template <typename T, int RA, int RB, int DIR> struct CtB2 {
    static NW_HD void run(cx2<T>* v) {
        constexpr int R = RA * RB;
        constexpr CtTab<R> tab{};
        cx2<T> u[R];   // u[k1 * RB + n2]
#pragma unroll
        for (int n2 = 0; n2 < RB; ++n2) {
            cx2<T> t[RA];
#pragma unroll
            for (int n1 = 0; n1 < RA; ++n1) t[n1] = v[RB * n1 + n2];
            B2<T, RA, DIR>::run(t);
#pragma unroll
            for (int k1 = 0; k1 < RA; ++k1) {
                const int j = (k1 * n2) % R;
                if (j == 0) u[k1 * RB + n2] = t[k1];
                else if (4 * j == R) u[k1 * RB + n2] = rot2<DIR, T>(t[k1]);
                else u[k1 * RB + n2] = cmul_k<DIR, T>(t[k1], (T)tab.c[j], (T)tab.s[j]);
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < RA; ++k1) {
            cx2<T> t[RB];
#pragma unroll
            for (int n2 = 0; n2 < RB; ++n2) t[n2] = u[k1 * RB + n2];
            B2<T, RB, DIR>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < RB; ++k2) v[k1 + RA * k2] = t[k2];
        }
    }
};
template <typename T, int DIR> struct B2<T, 25, DIR> : CtB2<T, 5, 5, DIR> {};
template <typename T, int DIR> struct B2<T, 32, DIR> : CtB2<T, 4, 8, DIR> {};
template <typename T, int DIR> struct B2<T, 24, DIR> : PfaB2<T, 3, 8, DIR> {};
template <typename T, int DIR> struct B2<T, 30, DIR> : PfaB2<T, 5, 6, DIR> {};

}  // namespace nw
