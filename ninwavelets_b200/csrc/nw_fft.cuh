// Batched, interleaved shared-memory Stockham FFT engine.
//
// A CTA transforms TT = 2^tsh independent length-P sequences at once.  Element
// p of sequence t lives at buf[p * pitch + t]: the batch index is the fastest
// one, so with the "t-fastest" thread mapping every butterfly access of a warp
// is a contiguous run of shared memory regardless of radix or stride (no bank
// conflicts, one twiddle shared by the TT lanes of a butterfly).  The last
// stage may instead use the "b-fastest" mapping so that a warp produces
// consecutive output samples of one sequence (coalesced global stores); that
// needs an odd pitch (TT+1) to stay conflict free.
//
// The first stage reads through a Src functor and the last stage writes
// through a Dst functor, which is how spectrum generation, the twiddle of the
// four-step split and the |z|^2 epilogue are fused into the transform without
// extra passes over shared memory.
//
// DIR = +1: inverse kernel e^{+2 pi i k n / P} (unnormalised), DIR = -1: forward.
// Twiddle table tw[j] = e^{+2 pi i j / P}, j in [0, P); forward conjugates it.
#pragma once
#include "nw_common.h"

namespace nw {

// ---- radix butterflies (in registers) ---------------------------------------
template <typename T, int R, int DIR> struct Bfly;

template <typename T, int DIR> struct Bfly<T, 2, DIR> {
    static NW_HD void run(cx<T>* v) {
        cx<T> a = v[0], b = v[1];
        v[0] = a + b;
        v[1] = a - b;
    }
};

template <typename T, int DIR> struct Bfly<T, 3, DIR> {
    static NW_HD void run(cx<T>* v) {
        const T s = (T)0.86602540378443864676372317075294;  // sin(2pi/3)
        cx<T> t1 = v[1] + v[2];
        cx<T> t2 = v[1] - v[2];
        cx<T> m = mk<T>(v[0].x - (T)0.5 * t1.x, v[0].y - (T)0.5 * t1.y);
        cx<T> n = scale(rot<DIR>(t2), s);
        v[0] = v[0] + t1;
        v[1] = m + n;
        v[2] = m - n;
    }
};

template <typename T, int DIR> struct Bfly<T, 4, DIR> {
    static NW_HD void run(cx<T>* v) {
        cx<T> t0 = v[0] + v[2], t1 = v[0] - v[2];
        cx<T> t2 = v[1] + v[3], t3 = rot<DIR>(v[1] - v[3]);
        v[0] = t0 + t2;
        v[2] = t0 - t2;
        v[1] = t1 + t3;
        v[3] = t1 - t3;
    }
};

template <typename T, int DIR> struct Bfly<T, 5, DIR> {
    static NW_HD void run(cx<T>* v) {
        const T c1 = (T)0.30901699437494742410229341718282;    // cos(2pi/5)
        const T c2 = (T)-0.80901699437494742410229341718282;   // cos(4pi/5)
        const T s1 = (T)0.95105651629515357211643933337938;    // sin(2pi/5)
        const T s2 = (T)0.58778525229247312916870595463907;    // sin(4pi/5)
        cx<T> a1 = v[1] + v[4], b1 = v[1] - v[4];
        cx<T> a2 = v[2] + v[3], b2 = v[2] - v[3];
        cx<T> m1 = mk<T>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        cx<T> m2 = mk<T>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        cx<T> n1 = rot<DIR>(mk<T>(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
        cx<T> n2 = rot<DIR>(mk<T>(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
        v[0] = v[0] + a1 + a2;
        v[1] = m1 + n1;
        v[4] = m1 - n1;
        v[2] = m2 + n2;
        v[3] = m2 - n2;
    }
};

template <typename T, int DIR> struct Bfly<T, 8, DIR> {
    static NW_HD void run(cx<T>* v) {
        const T h = (T)0.70710678118654752440084436210485;
        cx<T> e[4] = {v[0], v[2], v[4], v[6]};
        cx<T> o[4] = {v[1], v[3], v[5], v[7]};
        Bfly<T, 4, DIR>::run(e);
        Bfly<T, 4, DIR>::run(o);
        cx<T> w1 = scale(o[1] + rot<DIR>(o[1]), h);   // o1 * e^{DIR i pi/4}
        cx<T> w2 = rot<DIR>(o[2]);                    // o2 * e^{DIR i pi/2}
        cx<T> w3 = scale(rot<DIR>(o[3]) - o[3], h);   // o3 * e^{DIR i 3pi/4}
        v[0] = e[0] + o[0];
        v[4] = e[0] - o[0];
        v[1] = e[1] + w1;
        v[5] = e[1] - w1;
        v[2] = e[2] + w2;
        v[6] = e[2] - w2;
        v[3] = e[3] + w3;
        v[7] = e[3] - w3;
    }
};

template <typename T, int DIR> struct Bfly<T, 16, DIR> {
    static NW_HD void run(cx<T>* v) {
        // 16 = 4 x 4 Cooley-Tukey in registers: v[4*a + b], a,b in [0,4)
        const T c1 = (T)0.92387953251128675612818318939679;  // cos(pi/8)
        const T s1 = (T)0.38268343236508977172845998403040;  // sin(pi/8)
        const T h = (T)0.70710678118654752440084436210485;
        cx<T> col[4][4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            cx<T> t[4] = {v[b], v[4 + b], v[8 + b], v[12 + b]};
            Bfly<T, 4, DIR>::run(t);
#pragma unroll
            for (int q = 0; q < 4; ++q) col[b][q] = t[q];
        }
        // twiddle col[b][q] by w16^{b*q}, w16 = e^{DIR 2 pi i/16}
        const cx<T> w1 = mk<T>(c1, DIR * s1), w2 = mk<T>(h, DIR * h), w3 = mk<T>(s1, DIR * c1);
        col[1][1] = cmul(col[1][1], w1);
        col[1][2] = cmul(col[1][2], w2);
        col[1][3] = cmul(col[1][3], w3);
        col[2][1] = cmul(col[2][1], w2);
        col[2][2] = rot<DIR>(col[2][2]);
        col[2][3] = cmul(col[2][3], mk<T>(-h, DIR * h));
        col[3][1] = cmul(col[3][1], w3);
        col[3][2] = cmul(col[3][2], mk<T>(-h, DIR * h));
        col[3][3] = cmul(col[3][3], mk<T>(-c1, -DIR * s1));  // w16^9
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cx<T> t[4] = {col[0][q], col[1][q], col[2][q], col[3][q]};
            Bfly<T, 4, DIR>::run(t);
#pragma unroll
            for (int p = 0; p < 4; ++p) v[q + 4 * p] = t[p];
        }
    }
};

// ---- shared-memory views ------------------------------------------------------
template <typename T> struct SmemSrc {
    const cx<T>* buf;
    int pitch;
    NW_HD cx<T> load(int p, int t) const { return buf[p * pitch + t]; }
};
template <typename T> struct SmemDst {
    cx<T>* buf;
    int pitch;
    NW_HD void store(int p, int t, cx<T> v) const { buf[p * pitch + t] = v; }
};

template <typename T> NW_HD cx<T> ldtw(const cx<T>* NW_RESTRICT tw, int idx) {
#if defined(__CUDA_ARCH__)
    return tw[idx];
#else
    return tw[idx];
#endif
}

// ---- one Stockham stage --------------------------------------------------------
// MAP 0: thread index -> (butterfly b, sequence t) with t fastest.
// MAP 1: b fastest (coalesced outputs along the transform axis).
template <typename T, int R, int DIR, int MAP, class Src, class Dst>
NW_HD void fft_stage(const FftStages& st, int s, int tsh, const cx<T>* NW_RESTRICT tw, const Src& src,
                     const Dst& dst, int tid, int nthr) {
    const int P = st.P;
    const int PR = P / R;
    const int ns = st.ns[s];
    const int tws = P / (ns * R);
    const uint32_t nwork = (uint32_t)PR << tsh;
    for (uint32_t lin = tid; lin < nwork; lin += nthr) {
        int b, t;
        if (MAP == 0) {
            b = (int)(lin >> tsh);
            t = (int)(lin & ((1u << tsh) - 1u));
        } else {
            t = (int)fd_div(lin, st.div_pr[s]);
            b = (int)lin - t * PR;
        }
        const int k = (ns == 1) ? 0 : b - (int)fd_div((uint32_t)b, st.div_ns[s]) * ns;
        cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src.load(b + r * PR, t);
        if (ns != 1) {
            const int base = k * tws;
#pragma unroll
            for (int r = 1; r < R; ++r) {
                cx<T> w = ldtw(tw, r * base);
                v[r] = DIR > 0 ? cmul(v[r], w) : cmulc(v[r], w);
            }
        }
        Bfly<T, R, DIR>::run(v);
        const int j0 = (b - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst.store(j0 + r * ns, t, v[r]);
    }
}

// Arbitrary (prime) radix: O(R^2) per butterfly, inputs parked in local memory.
template <typename T, int DIR, int MAP, class Src, class Dst>
NW_HD void fft_stage_generic(const FftStages& st, int s, int tsh, const cx<T>* NW_RESTRICT tw, const Src& src,
                             const Dst& dst, int tid, int nthr) {
    const int P = st.P;
    const int R = st.radix[s];
    const int PR = P / R;
    const int ns = st.ns[s];
    const int tws = P / (ns * R);
    const uint32_t nwork = (uint32_t)PR << tsh;
    for (uint32_t lin = tid; lin < nwork; lin += nthr) {
        int b, t;
        if (MAP == 0) {
            b = (int)(lin >> tsh);
            t = (int)(lin & ((1u << tsh) - 1u));
        } else {
            t = (int)fd_div(lin, st.div_pr[s]);
            b = (int)lin - t * PR;
        }
        const int k = (ns == 1) ? 0 : b - (int)fd_div((uint32_t)b, st.div_ns[s]) * ns;
        cx<T> v[MAX_GENERIC_RADIX];
        for (int r = 0; r < R; ++r) {
            cx<T> a = src.load(b + r * PR, t);
            if (ns != 1 && r) {
                cx<T> w = ldtw(tw, r * k * tws);
                a = DIR > 0 ? cmul(a, w) : cmulc(a, w);
            }
            v[r] = a;
        }
        const int j0 = (b - k) * R + k;
        for (int q = 0; q < R; ++q) {
            cx<T> acc = v[0];
            int e = 0;  // (r*q) mod R
            for (int r = 1; r < R; ++r) {
                e += q;
                if (e >= R) e -= R;
                cx<T> w = ldtw(tw, e * PR);
                acc = acc + (DIR > 0 ? cmul(v[r], w) : cmulc(v[r], w));
            }
            dst.store(j0 + q * ns, t, acc);
        }
    }
}

template <typename T, int DIR, int MAP, class Src, class Dst>
NW_HD void fft_stage_any(const FftStages& st, int s, int tsh, const cx<T>* NW_RESTRICT tw, const Src& src,
                         const Dst& dst, int tid, int nthr) {
    switch (st.radix[s]) {
        case 2: fft_stage<T, 2, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        case 3: fft_stage<T, 3, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        case 4: fft_stage<T, 4, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        case 5: fft_stage<T, 5, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        case 8: fft_stage<T, 8, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        case 16: fft_stage<T, 16, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
        default: fft_stage_generic<T, DIR, MAP>(st, s, tsh, tw, src, dst, tid, nthr); break;
    }
}

// ---- whole transform ----------------------------------------------------------
// Runs all stages; intermediate data ping-pongs between bufA and bufB (each
// P * pitch elements).  Ends with a barrier, so dst's target is visible to the
// CTA on return.  LASTMAP selects the thread mapping of the final stage.
template <typename T, int DIR, int LASTMAP, class Src, class Dst>
NW_HD void fft_run(const FftStages& st, int tsh, int pitch, cx<T>* bufA, cx<T>* bufB,
                   const cx<T>* NW_RESTRICT tw, const Src& src, const Dst& dst, int tid, int nthr) {
    const int n = st.nst;
    cx<T>* cur = bufA;
    cx<T>* oth = bufB;
    for (int s = 0; s < n; ++s) {
        const bool first = (s == 0), last = (s == n - 1);
        if (first && last) {
            fft_stage_any<T, DIR, LASTMAP>(st, s, tsh, tw, src, dst, tid, nthr);
        } else if (first) {
            SmemDst<T> d{cur, pitch};
            fft_stage_any<T, DIR, 0>(st, s, tsh, tw, src, d, tid, nthr);
        } else if (last) {
            SmemSrc<T> a{cur, pitch};
            fft_stage_any<T, DIR, LASTMAP>(st, s, tsh, tw, a, dst, tid, nthr);
        } else {
            SmemSrc<T> a{cur, pitch};
            SmemDst<T> d{oth, pitch};
            fft_stage_any<T, DIR, 0>(st, s, tsh, tw, a, d, tid, nthr);
            cx<T>* tmp = cur; cur = oth; oth = tmp;
        }
        NW_SYNC();
    }
}

}  // namespace nw
