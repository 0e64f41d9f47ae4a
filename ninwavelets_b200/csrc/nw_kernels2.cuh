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
    long long N;
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
    const cx<T>* X;     // spectra of the signals of this group, [nsig][N]
    cx<T>* Tm;          // intermediate ring, [rows][tm_stride]
    long long tm_stride;
    void* out;          // output of row 0 of the group
    const PrunePlan* pplans;   // pruned pass A: plans indexed by FreqRec::pad_
    int skew, skew_mod; // start-up stagger of the first wave: CTA waits ((linear block id / n_sm) % skew_mod) * skew cycles
    int n_sm;
    int narrow;         // launch the narrow-band variant of pass A (host side only)
    fastdiv dstepA;     // narrow-band pass A: x / (N1 / radix of the first pass)
    int tm_mod;         // timing experiment only: rows share Tm slots (by % tm_mod); 0 = off
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
// two-pass plans with radix 25-32 butterflies (launch shapes 4 and 5, NW_BIG_RADIX translation units)
template <> struct StaticPlan<8> { static const int P = 1024, R0 = 32, R1 = 32, R2 = 1, TPS = 2; };
template <> struct StaticPlan<9> { static const int P = 960, R0 = 32, R1 = 30, R2 = 1, TPS = 2; };
template <> struct StaticPlan<10> { static const int P = 800, R0 = 32, R1 = 25, R2 = 1, TPS = 2; };
template <> struct StaticPlan<11> { static const int P = 625, R0 = 25, R1 = 25, R2 = 1, TPS = 2; };
// fp64 tiles hold half as many columns
template <> struct StaticPlan<12> { static const int P = 1000, R0 = 10, R1 = 10, R2 = 10, TPS = 1; };
template <> struct StaticPlan<13> { static const int P = 1024, R0 = 16, R1 = 16, R2 = 4, TPS = 1; };
// long rows (2^22 ... 2^24)
template <> struct StaticPlan<14> { static const int P = 2048, R0 = 16, R1 = 16, R2 = 8, TPS = 1; };
template <> struct StaticPlan<15> { static const int P = 4096, R0 = 16, R1 = 16, R2 = 16, TPS = 0; };
template <> struct StaticPlan<16> { static const int P = 4096, R0 = 16, R1 = 16, R2 = 16, TPS = 1; };
static const int N_STATIC_PLANS = 16;
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
    if (static_plan_matches<8>(st, tpsh)) return 8;
    if (static_plan_matches<9>(st, tpsh)) return 9;
    if (static_plan_matches<10>(st, tpsh)) return 10;
    if (static_plan_matches<11>(st, tpsh)) return 11;
    if (static_plan_matches<12>(st, tpsh)) return 12;
    if (static_plan_matches<13>(st, tpsh)) return 13;
    if (static_plan_matches<14>(st, tpsh)) return 14;
    if (static_plan_matches<15>(st, tpsh)) return 15;
    if (static_plan_matches<16>(st, tpsh)) return 16;
    return 0;
}

// CTAs of one launch start together and, doing identical work, would run their load / butterfly / store phases in
// lockstep, leaving the FP32 pipe idle while all of them load and the LSU idle while all of them compute.  A one-time
// stagger of the first wave (later CTAs inherit it from the slot they take over) de-phases the CTAs of an SM.
template <typename T> NW_HD void first_wave_stagger(const Long2Params<T>& P, int bx, int by, int gx) {
#if defined(__CUDA_ARCH__)
    if (P.skew > 0) {
        const unsigned lin = (unsigned)by * (unsigned)gx + (unsigned)bx;
        if (lin < (unsigned)(P.n_sm * P.skew_mod)) {
            const long long wait = (long long)((lin / (unsigned)P.n_sm) % (unsigned)P.skew_mod) * P.skew;
            const long long t0 = clock64();
            while (clock64() - t0 < wait) {}
        }
    }
#else
    (void)P; (void)bx; (void)by; (void)gx;
#endif
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
    int n1a;          // pruned transform: sequence t = tp * n1a + a computes rows n1 = a + n1a * b; else 1
    fastdiv dn1a;
    struct Ctx {
        cx2<T> cur, g;
        cx<T>* col;
        uint32_t n1, step;
        bool valid, two;
    };
    NW_HD Ctx begin(int base, int step, int t) const {
        Ctx x;
        const int N2 = P->N2;
        int tp = t, a = 0;
        if (n1a > 1) { tp = (int)fd_div((uint32_t)t, dn1a); a = t - tp * n1a; }
        const int k2 = c + 2 * tp;
        x.valid = k2 < N2;
        x.two = k2 + 1 < N2;
        const int k2a = x.valid ? k2 : 0, k2b = x.two ? k2 + 1 : k2a;
        const int n1 = a + n1a * base, dn1 = n1a * step;
#if defined(NW_KNOCKOUT) && NW_KNOCKOUT == 4   /* timing experiment: no twiddle-table look-ups */
        x.cur = mk2<T>(mk<T>((T)1, (T)(k2a * n1)), mk<T>((T)1, (T)0));
        x.g = mk2<T>(mk<T>((T)1, (T)dn1), mk<T>((T)1, (T)0));
#else
        x.cur = mk2<T>(big_twiddle2<T, DIR>(*P, k2a * n1), big_twiddle2<T, DIR>(*P, k2b * n1));
        x.g = mk2<T>(big_twiddle2<T, DIR>(*P, k2a * dn1), big_twiddle2<T, DIR>(*P, k2b * dn1));
#endif
        x.col = tm + ((size_t)k2a << (P->tpshB + 1));
        x.n1 = (uint32_t)n1;
        x.step = (uint32_t)dn1;
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
            // Tm holds plain complex values, TB consecutive rows n1 of one column contiguous.  (-DNW_TM_PACKED: lane-packed
            // units {re(n1), re(n1+1), im(n1), im(n1+1)} written as four scalars - no repacking in pass B but twice the
            // store sectors; measured 2.6 % slower per cfg2 step.)
            const size_t e = (size_t)((n1 >> shB) * blk + (n1 & mask));          // complex index within the column pair's rows
            T* o = (T*)x.col + (((e >> 1) << 2) + (n1 & 1u));
#if defined(NW_KNOCKOUT) && NW_KNOCKOUT == 3   /* timing experiment: no Tm stores */
            if (pk_lo(y.re) != (T)123.456) continue;
#endif
#if !defined(NW_TM_PACKED)   /* plain complex Tm: 8-byte stores that fill whole sectors; pass B repacks in its first pass */
            cx<T>* oc = x.col + e;
            oc[0] = lane0(y);
            if (x.two) oc[(size_t)1 << shB] = lane1(y);
#else
            o[0] = pk_lo(y.re);
            o[2] = pk_lo(y.im);
            if (x.two) {
                o[(size_t)2 << shB] = pk_hi(y.re);
                o[((size_t)2 << shB) + 2] = pk_hi(y.im);
            }
#endif
        }
    }
};

template <typename T> NW_HD size_t passA2_smem_bytes(int N1, int tpsh) { return ((size_t)N1 << tpsh) * sizeof(cx2<T>); }
template <typename T> NW_HD size_t passB2_smem_bytes(int N2, int tpsh) { return ((size_t)N2 << tpsh) * sizeof(cx2<T>) + 16; }

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
    const cx<T>* X = P.X + (size_t)si * (size_t)P.N;
    // rows k1 of this tile that hold a non-zero bin k = k1 N2 + k2, k2 in [c, c + 2 TP)
    const int clast = (c + 2 * TP < N2 ? c + 2 * TP : N2) - 1;
    int k1lo = rec.lo > clast ? (rec.lo - clast + N2 - 1) / N2 : 0;
    int k1hi = rec.hi - 1 >= c ? (rec.hi - 1 - c) / N2 : -1;
    if (k1hi > N1 - 1) k1hi = N1 - 1;
    const int nk1 = k1hi - k1lo + 1;
    first_wave_stagger<T>(P, bx, by, (N2 + 2 * TP - 1) / (2 * TP));
    TmDst2<T> dst{&P, P.Tm + (size_t)(P.tm_mod > 0 ? by % P.tm_mod : by) * P.tm_stride, c, 1, fastdiv{1, 0}};
    typedef StaticPlan<SP> S;
    if constexpr (NARROW) {
        // Narrow band (it touches at most N1 / R rows): every butterfly of the first pass - inputs k1 = rev + q * step -
        // has at most ONE non-zero input, so the pass needs no zero fill, no gather and no loads: evaluate that one
        // product y = W_f(k) X[k] (if any) and write its R outputs  y * (w_R^q)^j,  j = 0..R-1.
        const int rl = P.stA.radix[P.stA.nst - 1], step = N1 / rl;
        for (int i = tid; i < (step << tpsh); i += nthr) {
            const int tp = i & (TP - 1);
            const int blk = i >> tpsh;
            const int rev = fft2_rev(P.stA, blk);
            const int q = rev >= k1lo ? 0 : (int)fd_div((uint32_t)(k1lo - rev + step - 1), P.dstepA);
            const int k1 = rev + q * step;
            const int k2 = c + 2 * tp;
            cx2<T>* e = buf + (((size_t)blk * rl) << tpsh) + tp;
            const bool ok = nk1 > 0 && q < rl && k1 <= k1hi;
            cx2<T> y = zero2<T>();
            cx<T> wq = mk<T>((T)1, (T)0);
            if (ok) {
                const int k = k1 * N2 + k2;
                cx<T> a = mk<T>((T)0, (T)0), b = a;
                if (k2 < N2 && k >= rec.lo && k < rec.hi) a = spec_times<T>(P.sp, rec, fi, k, X[k]);
                if (k2 + 1 < N2 && k + 1 >= rec.lo && k + 1 < rec.hi) b = spec_times<T>(P.sp, rec, fi, k + 1, X[k + 1]);
                y = mk2<T>(a, b);
                wq = P.twA[q * step];
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
    const cx2<T> z = zero2<T>();
    for (int i = tid; i < (N1 << tpsh); i += nthr) buf[i] = z;
    NW_SYNC();
    for (int i = tid; i < (nk1 << tpsh); i += nthr) {
        const int tp = i & (TP - 1);
        const int k1 = k1lo + (i >> tpsh);
        const int k2 = c + 2 * tp;
        const int k = k1 * N2 + k2;
        cx<T> a = mk<T>((T)0, (T)0), b = a;
        if (k2 < N2 && k >= rec.lo && k < rec.hi) a = spec_times<T>(P.sp, rec, fi, k, X[k]);
        if (k2 + 1 < N2 && k + 1 >= rec.lo && k + 1 < rec.hi) b = spec_times<T>(P.sp, rec, fi, k + 1, X[k + 1]);
        buf[((size_t)fft2_dit_pos(P.stA, k1) << tpsh) + tp] = mk2<T>(a, b);
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
    for (int i = tid; i < (N1 << tpsh); i += nthr) {
        const int tp = i & (TP - 1);
        const int k1 = i >> tpsh;
        const int k2 = c + 2 * tp;
        const size_t k = (size_t)k1 * N2 + k2;
        const T a = k2 < N2 ? x[k] : (T)0, b = k2 + 1 < N2 ? x[k + 1] : (T)0;
        buf[((size_t)fft2_dit_pos(P.stA, k1) << tpsh) + tp] = mk2<T>(pk_make(a, b), pk_bcast((T)0));
    }
    NW_SYNC();
    TmDst2<T, -1> dst{&P, P.Tm + (size_t)by * P.tm_stride, c, 1, fastdiv{1, 0}};
    fft2_dit<T, -1>(P.stA, tpsh, P.twA, buf, FromBuf(), dst, tid, nthr);
}

// Pruned pass A.  A band that touches C <= n1b rows k1 gives every column a window of <= n1b consecutive k1;
// with n1 = a + n1a b (n1a = N1 / n1b):   A[a + n1a b] = sum_j U_a[j] w_n1b^{j b},   U_a[k1 mod n1b] = Y[k1] w_N1^{k1 a}
// - n1a transforms of length n1b per column instead of one of length N1 (log2 n1b levels instead of log2 N1).
// The tile interleaves (column pair, phase a) sequences, a fastest, so natural-order results of neighbouring
// lanes are consecutive rows n1 of the same column.
template <typename T>
NW_HD void passA2p_body(const Long2Params<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    cx2<T>* buf = (cx2<T>*)smem;
    const int tpsh = P.tpshA, TP = 1 << tpsh;
    const int c = bx << (tpsh + 1);
    const int gr = P.row0 + by;
    const int si = gr / P.F, fi = gr - si * P.F;
    const int N1 = P.N1, N2 = P.N2;
    const FreqRec rec = P.sp.rec[fi];
    const PrunePlan& pp = P.pplans[rec.pad_];
    const int n1a = pp.n1a, n1b = pp.st.P, nseq = pp.nseq;
    const cx<T>* X = P.X + (size_t)si * (size_t)P.N;
    const int clast = (c + 2 * TP < N2 ? c + 2 * TP : N2) - 1;
    int k1lo = rec.lo > clast ? (rec.lo - clast + N2 - 1) / N2 : 0;
    int k1hi = rec.hi - 1 >= c ? (rec.hi - 1 - c) / N2 : -1;
    if (k1hi > N1 - 1) k1hi = N1 - 1;
    const int nk1 = k1hi - k1lo + 1;   // <= n1b by construction of the plan
    const cx2<T> z = zero2<T>();
    for (int i = tid; i < (N1 << tpsh); i += nthr) buf[i] = z;
    NW_SYNC();
    // one thread per in-band (row k1, column pair): evaluate Y once, write its n1a phase-shifted copies
    // U_a = Y w_N1^{k1 a}; the table index k1 a mod N1 is kept incrementally
    for (int i = tid; i < (nk1 << tpsh); i += nthr) {
        const int tp = i & (TP - 1);
        const int k1 = k1lo + (i >> tpsh);
        const int k2 = c + 2 * tp;
        const int k = k1 * N2 + k2;
        cx<T> a = mk<T>((T)0, (T)0), b = a;
        if (k2 < N2 && k >= rec.lo && k < rec.hi) a = spec_times<T>(P.sp, rec, fi, k, X[k]);
        if (k2 + 1 < N2 && k + 1 >= rec.lo && k + 1 < rec.hi) b = spec_times<T>(P.sp, rec, fi, k + 1, X[k + 1]);
        const cx2<T> y = mk2<T>(a, b);
        const int j = k1 - (int)fd_div((uint32_t)k1, pp.dn1b) * n1b;
        cx2<T>* slot = buf + (size_t)fft2_dit_pos(pp.st, j) * nseq + tp * n1a;
        slot[0] = y;
        int idx = 0;
#pragma unroll 2
        for (int q = 1; q < n1a; ++q) {
            idx += k1;
            if (idx >= N1) idx -= N1;
            slot[q] = cmul_s(y, P.twA[idx]);
        }
    }
    NW_SYNC();
    TmDst2<T> dst{&P, P.Tm + (size_t)by * P.tm_stride, c, n1a, pp.dn1a};
    SeqDesc sq;
    sq.tpsh = -1;
    sq.nseq = nseq;
    sq.twscale = n1a;
    sq.d = pp.dseq;
    fft2_dit<T, +1>(pp.st, sq, P.twA, buf, FromBuf(), dst, tid, nthr);
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
    const cx<T>* tile = P.Tm + (size_t)(P.tm_mod > 0 ? by % P.tm_mod : by) * P.tm_stride + (size_t)bx * tile_elems;
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
    first_wave_stagger<T>(P, bx, by, (P.N1 + TB - 1) / TB);
#else
    for (size_t i = tid; i < tile_elems; i += nthr) ((cx<T>*)smem)[i] = tile[i];
    NW_SYNC();
#endif
    const int gr = P.row0 + by;
    const size_t esz = (MODE == OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
    LongOutDst2<T, MODE> dst{(char*)P.out + (size_t)gr * (size_t)P.N * esz, P.N1, bx * TB, (P.N1 & 1) == 0};
    typedef StaticPlan<SP> S;
#if !defined(NW_TM_PACKED)
    constexpr bool RAWT = true;
#else
    constexpr bool RAWT = false;
#endif
    if constexpr (SP == 0) fft2_dif<T, DIR, RAWT>(P.stB, P.tpshB, P.twB, buf, dst, tid, nthr);
    else fft2_dif_static<T, DIR, RAWT, S::TPS, (SP ? S::P : 4), (SP ? S::R0 : 2), (SP ? S::R1 : 2), S::R2>(P.twB, buf, dst, tid, nthr);
}

}  // namespace nw
