// Kernel bodies of the fast long-row path (packed in-place engine, nw_fft2.cuh).
//
// Four-step split N = N1 * N2 (k = N2 k1 + k2, n = n1 + N1 n2), as in nw_kernels.cuh:
//   passA2_body  columns: for TA = 2*TPA consecutive k2 of one (signal, frequency) row, spectrum
//                generation x signal spectrum (base.py:236-248, 404) gathered straight into a
//                decimation-in-time N1-point transform; natural-order results times w_N^{k2 n1}
//                go to the L2-resident intermediate Tm in TB-blocked layout.
//   passB2_body  rows: one bulk (TMA) copy brings a tile of TB = 2*TPB consecutive n1 x all k2 (already in the
//                engine's lane-packed units) into shared memory, a decimation-in-frequency N2-point transform runs in place, and
//                the |z|^2 / |z| / z epilogue (base.py:425, 443) stores straight from registers -
//                each n2 gives TB consecutive output samples (32-byte sectors for fp32, TB = 8).
// Both bodies compile for the host as well (tests/emul steps them block by block).
#pragma once
#include "nw_common.h"
#include "nw_fft2.cuh"
#include "nw_family.cuh"
#include "nw_kernels.cuh"

namespace nw {

template <typename T>
struct Long2Params {
    long long N;        // transform length N1 * N2 (a resampled group's plan: the decimated length M)
    long long xstride;  // elements between the spectra of consecutive signals in X (the data length Nd)
    int N1, N2, F;
    int tpshA;          // pass A: 2 << tpshA columns per CTA
    int tpshB;          // pass B: 2 << tpshB rows (n1) per CTA
    Fft2Plan stA, stB;
    const cx<T>* twA;   // [N1]
    const cx<T>* twB;   // [N2]
    const cx<T>* twH;   // w_N^{j << lb}
    const cx<T>* twL;   // w_N^{j}, j < 1 << lb
    int lb;
    const T* signal;    // forward transform: real signals of this launch, [nsig][N]
    const cx<T>* X;     // spectra of the signals of this group, [nsig][xstride]
    cx<T>* Tm;          // intermediate ring, [rows][tm_stride]
    long long tm_stride;
    void* out;          // output of row 0 of the group
    // frequency subsets (resampled groups, nw_resample.cuh): launch row (signal si, local frequency fi) writes output
    // row si * F_out + fmap[fi]; fmap == nullptr: row si * F + fi
    const int* fmap;
    int F_out;
    // resampled groups: eq[|j|] = D / H(j), the inverse of the interpolation kernel's frequency response at the
    // (signed) transform bin j - applied to the spectrum so that the interpolated row is exact in the pass band
    const T* eq;
    int narrow;         // launch the narrow-band variant of pass A (host side only)
    const int* ditpos;  // [N1] fft2_dit_pos(stA, k1) tabulated (nullptr: computed per element)
    fastdiv dstepA;     // narrow-band pass A: x / (N1 / radix of the first pass)
    int row0;           // first row (signal-major: row = signal * F + frequency) of this launch
    int out_mode;
    SpecParams<T> sp;
};

// Compile-time plans of the hot lengths (nw_fft2.cuh: fft2_dif_static / fft2_dit_static).  0 = run-time plan.
template <int ID> struct StaticPlan { static const int P = 0, R0 = 1, R1 = 1, R2 = 1, TPS = 0; };
template <> struct StaticPlan<1> { static const int P = 1000, R0 = 10, R1 = 10, R2 = 10, TPS = 2; };
template <> struct StaticPlan<2> { static const int P = 600, R0 = 12, R1 = 10, R2 = 5, TPS = 2; };
template <> struct StaticPlan<3> { static const int P = 1024, R0 = 16, R1 = 16, R2 = 4, TPS = 2; };
template <> struct StaticPlan<4> { static const int P = 256, R0 = 16, R1 = 16, R2 = 1, TPS = 2; };
template <> struct StaticPlan<5> { static const int P = 512, R0 = 16, R1 = 8, R2 = 4, TPS = 2; };
template <> struct StaticPlan<6> { static const int P = 1500, R0 = 10, R1 = 10, R2 = 15, TPS = 1; };   // short rows, cfg3
template <> struct StaticPlan<7> { static const int P = 300, R0 = 12, R1 = 5, R2 = 5, TPS = 3; };      // short rows, cfg1
// fp64 tiles hold half as many columns
template <> struct StaticPlan<12> { static const int P = 1000, R0 = 10, R1 = 10, R2 = 10, TPS = 1; };
template <> struct StaticPlan<13> { static const int P = 1024, R0 = 16, R1 = 16, R2 = 4, TPS = 1; };
// long rows (2^22 ... 2^24)
template <> struct StaticPlan<14> { static const int P = 2048, R0 = 16, R1 = 16, R2 = 8, TPS = 1; };
template <> struct StaticPlan<15> { static const int P = 4096, R0 = 16, R1 = 16, R2 = 16, TPS = 0; };
template <> struct StaticPlan<16> { static const int P = 4096, R0 = 16, R1 = 16, R2 = 16, TPS = 1; };
// decimated lengths of resampled rows (N = 600 000 / D splits into 100, 120, 150 x 125 ... 1000): two-pass column
// transforms and three-pass row transforms of the short factors
template <> struct StaticPlan<20> { static const int P = 100, R0 = 10, R1 = 10, R2 = 1, TPS = 2; };
template <> struct StaticPlan<21> { static const int P = 120, R0 = 12, R1 = 10, R2 = 1, TPS = 2; };
template <> struct StaticPlan<22> { static const int P = 150, R0 = 10, R1 = 15, R2 = 1, TPS = 2; };
template <> struct StaticPlan<23> { static const int P = 500, R0 = 10, R1 = 10, R2 = 5, TPS = 2; };
template <> struct StaticPlan<24> { static const int P = 250, R0 = 10, R1 = 5, R2 = 5, TPS = 2; };
template <> struct StaticPlan<25> { static const int P = 125, R0 = 5, R1 = 5, R2 = 5, TPS = 2; };
template <> struct StaticPlan<26> { static const int P = 200, R0 = 10, R1 = 10, R2 = 2, TPS = 2; };
static const int N_STATIC_PLANS = 26;
template <int ID> NW_HD bool static_plan_matches(const Fft2Plan& st, int tpsh) {
    typedef StaticPlan<ID> S;
    if (st.P != S::P || tpsh != S::TPS) return false;
    const int n = S::R2 > 1 ? 3 : 2;
    if (st.nst != n || st.radix[0] != S::R0 || st.radix[1] != S::R1) return false;
    return n == 2 || st.radix[2] == S::R2;
}
inline int static_plan_id(const Fft2Plan& st, int tpsh) {
    if (static_plan_matches<1>(st, tpsh)) return 1;
    if (static_plan_matches<2>(st, tpsh)) return 2;
    if (static_plan_matches<3>(st, tpsh)) return 3;
    if (static_plan_matches<4>(st, tpsh)) return 4;
    if (static_plan_matches<5>(st, tpsh)) return 5;
    if (static_plan_matches<6>(st, tpsh)) return 6;
    if (static_plan_matches<7>(st, tpsh)) return 7;
    if (static_plan_matches<12>(st, tpsh)) return 12;
    if (static_plan_matches<13>(st, tpsh)) return 13;
    if (static_plan_matches<14>(st, tpsh)) return 14;
    if (static_plan_matches<15>(st, tpsh)) return 15;
    if (static_plan_matches<16>(st, tpsh)) return 16;
    if (static_plan_matches<20>(st, tpsh)) return 20;
    if (static_plan_matches<21>(st, tpsh)) return 21;
    if (static_plan_matches<22>(st, tpsh)) return 22;
    if (static_plan_matches<23>(st, tpsh)) return 23;
    if (static_plan_matches<24>(st, tpsh)) return 24;
    if (static_plan_matches<25>(st, tpsh)) return 25;
    if (static_plan_matches<26>(st, tpsh)) return 26;
    return 0;
}

template <typename T, int DIR = 1> NW_HD cx<T> big_twiddle2(const Long2Params<T>& P, int m) {
    const cx<T> a = P.twH[m >> P.lb];
    const cx<T> b = P.twL[m & ((1 << P.lb) - 1)];
    const cx<T> w = cmul(a, b);
    return DIR > 0 ? w : mk<T>(w.x, -w.y);
}

// ---- pass A --------------------------------------------------------------------------------
// natural-order results n1 = base + q * step of columns k2, k2 + 1, times w_N^{k2 n1}: the factor runs
// as a geometric sequence in q (ratio w_N^{k2 step}), two table look-ups per lane and butterfly
template <typename T, int DIR = 1> struct TmDst2 {
    const Long2Params<T>* P;
    cx<T>* tm;
    int c;
    struct Ctx {
        cx2<T> cur, g;
        cx<T>* col;
        uint32_t n1, step;
        bool valid, two;
    };
    NW_HD Ctx begin(int base, int step, int tp) const {
        Ctx x;
        const int N2 = P->N2;
        const int k2 = c + 2 * tp;
        x.valid = k2 < N2;
        x.two = k2 + 1 < N2;
        const int k2a = x.valid ? k2 : 0, k2b = x.two ? k2 + 1 : k2a;
        x.cur = mk2<T>(big_twiddle2<T, DIR>(*P, k2a * base), big_twiddle2<T, DIR>(*P, k2b * base));
        x.g = mk2<T>(big_twiddle2<T, DIR>(*P, k2a * step), big_twiddle2<T, DIR>(*P, k2b * step));
        x.col = tm + ((size_t)k2a << (P->tpshB + 1));
        x.n1 = (uint32_t)base;
        x.step = (uint32_t)step;
        return x;
    }
    template <int R> NW_HD void store_all(const Ctx& x, const cx2<T>* v) const {
        if (!x.valid) return;
        const int shB = P->tpshB + 1;
        const uint32_t mask = (1u << shB) - 1u;
        const uint32_t blk = (uint32_t)P->N2 << shB;        // elements per block of 2^shB rows
        // two interleaved recurrences (even / odd q, ratio g^2) halve the dependency chain
        cx2<T> cur[2] = {x.cur, cmul_p(x.cur, x.g)};
        const cx2<T> g2 = cmul_p(x.g, x.g);
        uint32_t n1 = x.n1;
#pragma unroll
        for (int q = 0; q < R; ++q, n1 += x.step) {
            const cx2<T> y = cmul_p(v[q], cur[q & 1]);
            if (q + 2 < R) cur[q & 1] = cmul_p(cur[q & 1], g2);
            // Tm holds plain complex values, TB consecutive rows n1 of one column contiguous: 8-byte stores that fill
            // whole sectors; pass B repacks pairs of rows into lanes in its first pass
            const size_t e = (size_t)((n1 >> shB) * blk + (n1 & mask));          // complex index within the column pair's rows
            cx<T>* oc = x.col + e;
            oc[0] = lane0(y);
            if (x.two) oc[(size_t)1 << shB] = lane1(y);
        }
    }
};

template <typename T> NW_HD size_t passA2_smem_bytes(int N1, int tpsh) { return ((size_t)N1 << tpsh) * sizeof(cx2<T>); }
template <typename T> NW_HD size_t passB2_smem_bytes(int N2, int tpsh) { return ((size_t)N2 << tpsh) * sizeof(cx2<T>) + 16; }

NW_HD int nw_floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }   // b > 0
NW_HD int nw_ceil_div(int a, int b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }  // b > 0

// W_f(k) X[k] for the two columns k2, k2 + 1 of row k1s (signed: the band [rec.lo, rec.hi) is given in signed transform
// bins j = k1s N2 + k2, which a resampled group's plan places around bin 0, i.e. wrapped around the ends of its
// M-point spectrum; the data bin is j + rec.shift), times the group's equaliser
template <typename T>
NW_HD cx2<T> passA2_bins(const Long2Params<T>& P, const FreqRec& rec, int fi, const cx<T>* X, int k1s, int k2) {
    const int j = k1s * P.N2 + k2;
    cx<T> a = mk<T>((T)0, (T)0), b = a;
    if (P.sp.wtab) {   // tabulated weights (equaliser folded in)
        const T* wt = P.sp.wtab + rec.woff - rec.lo;
        if (k2 < P.N2 && j >= rec.lo && j < rec.hi) a = scale(X[j + rec.shift], wt[j]);
        if (k2 + 1 < P.N2 && j + 1 >= rec.lo && j + 1 < rec.hi) b = scale(X[j + 1 + rec.shift], wt[j + 1]);
        return mk2<T>(a, b);
    }
    if (k2 < P.N2 && j >= rec.lo && j < rec.hi) {
        a = spec_times<T>(P.sp, rec, fi, j + rec.shift, X[j + rec.shift]);
        if (P.eq) a = scale(a, P.eq[j < 0 ? -j : j]);
    }
    if (k2 + 1 < P.N2 && j + 1 >= rec.lo && j + 1 < rec.hi) {
        b = spec_times<T>(P.sp, rec, fi, j + 1 + rec.shift, X[j + 1 + rec.shift]);
        if (P.eq) b = scale(b, P.eq[j + 1 < 0 ? -(j + 1) : j + 1]);
    }
    return mk2<T>(a, b);
}

// The gather of a tile with tabulated weights, four slots per thread and trip, every load (slot position, two bins of X,
// two weights) issued before the first use: one global-memory round trip per four slots instead of one per slot (ncu r02:
// long-scoreboard stalls dominated pass A at two resident warps per CTA; cfg2 7.31 -> 6.82 ms per step).
//   WIDE:  element i is slot row k1 = i >> tpsh, which holds the row of the window [k1lo, k1lo + N1) congruent to it
//   !WIDE: element i is signed row k1s = k1lo + (i >> tpsh) of the band's own rows, slot row k1s mod N1
template <typename T, bool WIDE>
NW_HD void passA2_gather4(const Long2Params<T>& P, const FreqRec& rec, const cx<T>* X, cx2<T>* buf, int total, int k1lo, int c,
                          int tid, int nthr) {
    const int tpsh = P.tpshA, TP = 1 << tpsh, N1 = P.N1, N2 = P.N2;
    const T* wt = P.sp.wtab + rec.woff - rec.lo;
    const cx<T>* Xs = X + rec.shift;
    const cx<T> z0 = mk<T>((T)0, (T)0);
    for (int i0 = tid; i0 < total; i0 += 4 * nthr) {
        cx<T> xa[4], xb[4];
        T wa[4], wb[4];
        int pos[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nthr;
            const bool v = i < total;
            const int k2 = c + 2 * (i & (TP - 1));
            int k1, k1s;
            if (WIDE) {
                k1 = i >> tpsh;
                int u = k1 - k1lo;             // |k1lo| <= N1: u in (-N1, 2 N1)
                if (u < 0) u += N1;
                if (u >= N1) u -= N1;
                k1s = k1lo + u;
            } else {
                k1s = k1lo + (i >> tpsh);
                k1 = k1s % N1;
                if (k1 < 0) k1 += N1;
            }
            const int j = k1s * N2 + k2;
            const bool oa = v && k2 < N2 && j >= rec.lo && j < rec.hi;
            const bool ob = v && k2 + 1 < N2 && j + 1 >= rec.lo && j + 1 < rec.hi;
            pos[q] = !v ? 0 : P.ditpos ? P.ditpos[k1] : fft2_dit_pos(P.stA, k1);
            xa[q] = oa ? Xs[j] : z0;
            wa[q] = oa ? wt[j] : (T)0;
            xb[q] = ob ? Xs[j + 1] : z0;
            wb[q] = ob ? wt[j + 1] : (T)0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nthr;
            if (i < total) buf[((size_t)pos[q] << tpsh) + (i & (TP - 1))] = mk2<T>(scale(xa[q], wa[q]), scale(xb[q], wb[q]));
        }
    }
}

// The tile's input - spectrum x signal spectrum on the non-zero band only - is gathered into the
// transform's shared-memory slots by a compact loop (one evaluation per in-band bin, nothing unrolled
// around the formula), the rest of the tile is zero.
// NARROW: the planner guarantees that every frequency's band touches at most N1 / R_last rows (HostPlan::narrowA).
template <typename T, int SP, bool NARROW = false>
NW_HD void passA2_body(const Long2Params<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    cx2<T>* buf = (cx2<T>*)smem;
    const int tpsh = P.tpshA, TP = 1 << tpsh;
    const int c = bx << (tpsh + 1);
    const int gr = P.row0 + by;
    const int si = gr / P.F, fi = gr - si * P.F;
    const int N1 = P.N1, N2 = P.N2;
    const FreqRec rec = P.sp.rec[fi];
    const cx<T>* X = P.X + (size_t)si * (size_t)P.xstride;
    // signed rows k1s of this tile that hold a non-zero bin j = k1s N2 + k2, k2 in [c, c + 2 TP); the slot is k1s mod N1
    const int clast = (c + 2 * TP < N2 ? c + 2 * TP : N2) - 1;
    const int k1lo = nw_ceil_div(rec.lo - clast, N2);
    int nk1 = nw_floor_div(rec.hi - 1 - c, N2) - k1lo + 1;
    if (nk1 > N1) nk1 = N1;
    if (rec.hi <= rec.lo || nk1 < 0) nk1 = 0;
    TmDst2<T> dst{&P, P.Tm + (size_t)by * P.tm_stride, c};
    typedef StaticPlan<SP> S;
    if constexpr (NARROW) {
        // Narrow band (it touches at most N1 / R rows): every butterfly of the first pass - inputs k1 = rev + q * step -
        // has at most ONE non-zero input, so the pass needs no zero fill, no gather and no loads: evaluate that one
        // product y = W_f(k) X[k] (if any) and write its R outputs  y * (w_R^q)^j,  j = 0..R-1.
        const int rl = P.stA.radix[P.stA.nst - 1], step = N1 / rl;
        if (P.sp.wtab) {
            // tabulated weights: four products per thread and trip, their loads (two bins of X, two weights, the twiddle)
            // issued before the first use (as passA2_gather4)
            const T* wt = P.sp.wtab + rec.woff - rec.lo;
            const cx<T>* Xs = X + rec.shift;
            const cx<T> z0 = mk<T>((T)0, (T)0);
            const int total = step << tpsh;
            for (int i0 = tid; i0 < total; i0 += 4 * nthr) {
                cx<T> xa[4], xb[4], wq[4];
                T wa[4], wb[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + q * nthr;
                    const int k2 = c + 2 * (i & (TP - 1));
                    const int rev = fft2_rev(P.stA, i >> tpsh);
                    const uint32_t d = (uint32_t)(rev - k1lo + N1);
                    const int u = (int)(d - fd_div(d, P.dstepA) * (uint32_t)step);
                    const bool ok = i < total && u < nk1;
                    const int k1s = k1lo + u;
                    int k1 = k1s % N1;
                    if (k1 < 0) k1 += N1;
                    const int j = k1s * N2 + k2;
                    const bool oa = ok && k2 < N2 && j >= rec.lo && j < rec.hi;
                    const bool ob = ok && k2 + 1 < N2 && j + 1 >= rec.lo && j + 1 < rec.hi;
                    xa[q] = oa ? Xs[j] : z0;
                    wa[q] = oa ? wt[j] : (T)0;
                    xb[q] = ob ? Xs[j + 1] : z0;
                    wb[q] = ob ? wt[j + 1] : (T)0;
                    wq[q] = ok ? P.twA[k1 - rev] : mk<T>((T)1, (T)0);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + q * nthr;
                    if (i >= total) break;
                    cx2<T>* e = buf + (((size_t)(i >> tpsh) * rl) << tpsh) + (i & (TP - 1));
                    cx2<T> y = mk2<T>(scale(xa[q], wa[q]), scale(xb[q], wb[q]));
                    e[0] = y;
#pragma unroll 4
                    for (int j = 1; j < rl; ++j) {
                        y = cmul_s(y, wq[q]);
                        e[(size_t)j << tpsh] = y;
                    }
                }
            }
        } else
        for (int i = tid; i < (step << tpsh); i += nthr) {
            const int tp = i & (TP - 1);
            const int blk = i >> tpsh;
            const int rev = fft2_rev(P.stA, blk);
            // the one row k1s = k1lo + u of the window with k1s = rev (mod step)
            const uint32_t d = (uint32_t)(rev - k1lo + N1);                 // > 0: |k1lo| <= N1
            const int u = (int)(d - fd_div(d, P.dstepA) * (uint32_t)step);
            const bool ok = u < nk1;
            cx2<T>* e = buf + (((size_t)blk * rl) << tpsh) + tp;
            cx2<T> y = zero2<T>();
            cx<T> wq = mk<T>((T)1, (T)0);
            if (ok) {
                const int k1s = k1lo + u;
                int k1 = k1s % N1;
                if (k1 < 0) k1 += N1;
                y = passA2_bins<T>(P, rec, fi, X, k1s, c + 2 * tp);
                wq = P.twA[k1 - rev];                                       // q * step, q = k1 / step
            }
            e[0] = y;
#pragma unroll 4
            for (int j = 1; j < rl; ++j) {
                y = cmul_s(y, wq);
                e[(size_t)j << tpsh] = y;
            }
        }
        if constexpr (SP == 0) fft2_dit<T, +1, FromBuf, TmDst2<T>, true>(P.stA, seq_pow2(tpsh), P.twA, buf, FromBuf(), dst, tid, nthr);
        else fft2_dit_static<T, +1, S::TPS, (SP ? S::P : 4), (SP ? S::R0 : 2), (SP ? S::R1 : 2), S::R2, TmDst2<T>, true>(P.twA, buf, dst, tid, nthr);
        return;
    }
    if (2 * nk1 >= N1) {
        // wide band (the decimated transforms of resampled rows): one sweep over the whole tile - row k1 holds the bin of
        // the window [k1lo, k1lo + N1) that is congruent to it, or zero (passA2_bins checks the band)
        const int total = N1 << tpsh;
        if (P.sp.wtab) {
            passA2_gather4<T, true>(P, rec, X, buf, total, k1lo, c, tid, nthr);
        } else {
            for (int i = tid; i < total; i += nthr) {
                const int tp = i & (TP - 1);
                const int k1 = i >> tpsh;
                int u = k1 - k1lo;             // |k1lo| <= N1: u in (-N1, 2 N1)
                if (u < 0) u += N1;
                if (u >= N1) u -= N1;
                const int pos = P.ditpos ? P.ditpos[k1] : fft2_dit_pos(P.stA, k1);
                buf[((size_t)pos << tpsh) + tp] = passA2_bins<T>(P, rec, fi, X, k1lo + u, c + 2 * tp);
            }
        }
        NW_SYNC();
        if constexpr (SP == 0) fft2_dit<T, +1>(P.stA, tpsh, P.twA, buf, FromBuf(), dst, tid, nthr);
        else fft2_dit_static<T, +1, S::TPS, (SP ? S::P : 4), (SP ? S::R0 : 2), (SP ? S::R1 : 2), S::R2>(P.twA, buf, dst, tid, nthr);
        return;
    }
    const cx2<T> z = zero2<T>();
    for (int i = tid; i < (N1 << tpsh); i += nthr) buf[i] = z;
    NW_SYNC();
    if (P.sp.wtab) {
        passA2_gather4<T, false>(P, rec, X, buf, nk1 << tpsh, k1lo, c, tid, nthr);
    } else {
        for (int i = tid; i < (nk1 << tpsh); i += nthr) {
            const int tp = i & (TP - 1);
            const int k1s = k1lo + (i >> tpsh);
            int k1 = k1s % N1;
            if (k1 < 0) k1 += N1;
            const int pos = P.ditpos ? P.ditpos[k1] : fft2_dit_pos(P.stA, k1);
            buf[((size_t)pos << tpsh) + tp] = passA2_bins<T>(P, rec, fi, X, k1s, c + 2 * tp);
        }
    }
    NW_SYNC();
    if constexpr (SP == 0) fft2_dit<T, +1>(P.stA, tpsh, P.twA, buf, FromBuf(), dst, tid, nthr);
    else fft2_dit_static<T, +1, S::TPS, (SP ? S::P : 4), (SP ? S::R0 : 2), (SP ? S::R1 : 2), S::R2>(P.twA, buf, dst, tid, nthr);
}

// Forward transform, pass A (scipy.fftpack.fft of the signal, base.py:399): the same column pass with the
// conjugate kernel; input = the real signal itself (every bin), row index = signal.
template <typename T>
NW_HD void passA2f_body(const Long2Params<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    cx2<T>* buf = (cx2<T>*)smem;
    const int tpsh = P.tpshA, TP = 1 << tpsh;
    const int c = bx << (tpsh + 1);
    const int N1 = P.N1, N2 = P.N2;
    const T* x = P.signal + (size_t)(P.row0 + by) * (size_t)P.N;
    const int total = N1 << tpsh;
    for (int i0 = tid; i0 < total; i0 += 4 * nthr) {   // four slots per trip, the loads issued first (as passA2_gather4)
        T a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nthr;
            const int k2 = c + 2 * (i & (TP - 1));
            const size_t k = (size_t)(i >> tpsh) * N2 + k2;
            a[q] = (i < total && k2 < N2) ? x[k] : (T)0;
            b[q] = (i < total && k2 + 1 < N2) ? x[k + 1] : (T)0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * nthr;
            if (i < total)
                buf[((size_t)fft2_dit_pos(P.stA, i >> tpsh) << tpsh) + (i & (TP - 1))] = mk2<T>(pk_make(a[q], b[q]), pk_bcast((T)0));
        }
    }
    NW_SYNC();
    TmDst2<T, -1> dst{&P, P.Tm + (size_t)by * P.tm_stride, c};
    fft2_dit<T, -1>(P.stA, tpsh, P.twA, buf, FromBuf(), dst, tid, nthr);
}

// ---- pass B --------------------------------------------------------------------------------
// results n2 = base + q * step of rows n1 = r0 + 2 tp, + 1  ->  out[n1 + N1 n2]
template <typename T, int MODE> struct LongOutDst2 {
    void* out;       // row base
    int N1, r0;
    bool vec_ok;     // N1 even: pairs of consecutive samples are 2*sizeof(T) aligned
    struct Ctx {
        size_t idx, dstep;
        bool valid, two;
    };
    NW_HD Ctx begin(int base, int step, int tp) const {
        Ctx x;
        const int n1 = r0 + 2 * tp;
        x.valid = n1 < N1;
        x.two = n1 + 1 < N1;
        x.idx = (size_t)n1 + (size_t)N1 * (size_t)base;
        x.dstep = (size_t)N1 * (size_t)step;
        return x;
    }
    template <int R> NW_HD void store_all(const Ctx& x, const cx2<T>* v) const {
        if (!x.valid) return;
        const bool two = x.two;
        const size_t idx = x.idx, dstep = x.dstep;
        if (MODE == OUT_CWT) {
            cx<T>* o = (cx<T>*)out + idx;
#pragma unroll
            for (int q = 0; q < R; ++q, o += dstep) {
                st_stream(o, lane0(v[q]));
                if (two) st_stream(o + 1, lane1(v[q]));
            }
            return;
        }
        T* o = (T*)out + idx;
        if (two && vec_ok) {
#pragma unroll
            for (int q = 0; q < R; ++q, o += dstep) {
                if (MODE == OUT_POWER) {
                    const pk<T> p = pk_fma(v[q].im, v[q].im, v[q].re * v[q].re);
                    st_stream((cx<T>*)o, mk<T>(pk_lo(p), pk_hi(p)));   // one 2*sizeof(T) store
                } else {
                    st_stream((cx<T>*)o, mk<T>(nw_hypot(pk_lo(v[q].re), pk_lo(v[q].im)), nw_hypot(pk_hi(v[q].re), pk_hi(v[q].im))));
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q, o += dstep) {
                st_stream(o, real_out<T>(MODE, lane0(v[q])));
                if (two) st_stream(o + 1, real_out<T>(MODE, lane1(v[q])));
            }
        }
    }
};

// DIR = -1: forward transform, pass B: P.out is the spectrum buffer [nsig][N] (MODE = OUT_CWT)
template <typename T, int MODE, int SP, int DIR = 1>
NW_HD void passB2_body(const Long2Params<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    const int shB = P.tpshB + 1;
    const int TB = 1 << shB;
    cx2<T>* buf = (cx2<T>*)smem;
    const size_t tile_elems = (size_t)P.N2 << shB;   // complex values
    const cx<T>* tile = P.Tm + (size_t)by * P.tm_stride + (size_t)bx * tile_elems;
#if defined(__CUDA_ARCH__)
    const size_t bytes = tile_elems * sizeof(cx<T>);
    uint64_t* bar = (uint64_t*)((char*)smem + bytes);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)bytes);
        const size_t CH = 32768;
        for (size_t off = 0; off < bytes; off += CH) {
            const size_t n = bytes - off < CH ? bytes - off : CH;
            bulk_g2s((char*)smem + off, (const char*)tile + off, (uint32_t)n, bar);
        }
    }
    mbar_wait(bar, 0);
#else
    for (size_t i = tid; i < tile_elems; i += nthr) ((cx<T>*)smem)[i] = tile[i];
    NW_SYNC();
#endif
    size_t orow = (size_t)(P.row0 + by);
    if (P.fmap) {   // frequency subset: launch row (si, fi) -> output row si * F_out + fmap[fi]
        const int gr = P.row0 + by, si = gr / P.F;
        orow = (size_t)si * (size_t)P.F_out + (size_t)P.fmap[gr - si * P.F];
    }
    const size_t esz = (MODE == OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
    LongOutDst2<T, MODE> dst{(char*)P.out + orow * (size_t)P.N * esz, P.N1, bx * TB, (P.N1 & 1) == 0};
    typedef StaticPlan<SP> S;
    constexpr bool RAWT = true;   // Tm holds plain complex values; the first pass repacks pairs of rows into lanes
    if constexpr (SP == 0) fft2_dif<T, DIR, RAWT>(P.stB, P.tpshB, P.twB, buf, dst, tid, nthr);
    else fft2_dif_static<T, DIR, RAWT, S::TPS, (SP ? S::P : 4), (SP ? S::R0 : 2), (SP ? S::R1 : 2), S::R2>(P.twB, buf, dst, tid, nthr);
}

}  // namespace nw
