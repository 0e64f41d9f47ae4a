// Host-side planner: everything the reference's make_fft_wavelets (base.py:258-279)
// and _setup_trans_shape (base.py:173-194) decide per (family, freqs, N), restated
// for a device that evaluates spectra on the fly: grid geometry, per-frequency
// parameter records, non-zero bands, radix plans, the four-step split and tile
// sizes.  Pure C++ (no CUDA calls) so the CPU test-suite can exercise it through
// the C ABI without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <algorithm>

#include "nw_common.h"
#include "nw_family.cuh"

namespace nw {

static const size_t SMEM_MAX = 227 * 1024;       // opt-in dynamic shared memory per CTA (sm_100)
static const size_t SMEM_HALF = 113 * 1024;      // two CTAs per SM

struct HostPlan {
    // description
    int device = 0, dtype = 0, family = 0, interpolate = 0;
    long long N = 0;
    int F = 0;
    double sfreq = 0, p0 = 0, p1 = 0, p2 = 0, prune_eps = 0;
    std::vector<double> freqs, aux;
    std::vector<double> table;   // complex interleaved [F][table_len]
    long long table_len = 0;
    std::vector<long long> table_lens;   // per frequency, <= table_len
    // grid geometry (base.py:173-194, 239-246, 75-82, 107-123)
    double df = 0;               // 1 / L
    long long n_eval = 0;        // evaluated grid bins
    int grid_off = 0;            // centre-pad offset
    long long cut = 0;           // bins >= cut are zero (interpolate), else N
    std::vector<FreqRec> rec;
    long long band_bins = 0;
    // execution plan
    int path = 0;                // 0 short, 1 long
    FftStages st{}, stA{}, stB{};
    int tsh = 0, pitch = 1;      // short
    int N1 = 0, N2 = 0, tshA = 0, pitchA = 1, tshB = 0, lb = 0;
    size_t smem_short = 0, smem_A = 0, smem_B = 0;
    int nthr_short = 256, nthr_long = 512;
    long long tm_stride = 0;
    int ring = 1;                // rows (and signals) in flight on the long path
};

inline size_t cx_size(int dtype) { return dtype == 0 ? 8 : 16; }

// ---- radix plan ----------------------------------------------------------------
inline bool factorise(long long P, FftStages& st, std::string& err) {
    std::vector<int> rad;
    long long m = P;
    int e2 = 0;
    while (m % 2 == 0) { m /= 2; ++e2; }
    while (m % 5 == 0) { m /= 5; rad.push_back(5); }
    while (m % 3 == 0) { m /= 3; rad.push_back(3); }
    for (long long p = 7; p * p <= m; p += 2)
        while (m % p == 0) {
            if (p > MAX_GENERIC_RADIX) { err = "length has a prime factor > 64"; return false; }
            rad.push_back((int)p);
            m /= p;
        }
    if (m > 1) {
        if (m > MAX_GENERIC_RADIX) { err = "length has a prime factor > 64"; return false; }
        rad.push_back((int)m);
    }
    if (e2 > 0) {  // balanced split of 2^e2 into radices <= 16
        int ns = (e2 + 3) / 4, base = e2 / ns, rem = e2 % ns;
        for (int i = 0; i < ns; ++i) rad.push_back(1 << (base + (i < rem ? 1 : 0)));
    }
    if (rad.empty()) rad.push_back(1);
    std::sort(rad.begin(), rad.end(), [](int a, int b) { return a > b; });
    if (P == 1) { err = "length 1"; return false; }
    if ((int)rad.size() > MAX_STAGES) { err = "too many radix stages"; return false; }
    st.P = (int)P;
    st.nst = (int)rad.size();
    int ns = 1;
    for (int i = 0; i < st.nst; ++i) {
        st.radix[i] = rad[i];
        st.ns[i] = ns;
        st.div_ns[i] = make_fastdiv((uint32_t)ns);
        st.div_pr[i] = make_fastdiv((uint32_t)(P / rad[i]));
        ns *= rad[i];
    }
    return true;
}

inline bool smooth_enough(long long P) {
    FftStages st;
    std::string e;
    return P >= 2 && P <= (1 << 15) && factorise(P, st, e);
}

// ---- grid geometry -----------------------------------------------------------------
// numpy.arange(0, total, one) has ceil(total / one) elements.
inline long long arange_len(double total, double one) {
    double q = total / one;
    long long n = (long long)ceil(q);
    return n < 0 ? 0 : n;
}

inline void plan_geometry(HostPlan& hp) {
    const double L = (double)hp.N / hp.sfreq;   // wave.shape[0] / self.sfreq, base.py:395
    const double one = 1 / L;                    // base.py:192
    hp.df = one;
    hp.cut = hp.N;
    hp.grid_off = 0;
    if (hp.family == FAM_TABLE) {
        // pad_to, base.py:75-82: truncate, or centre with the short half in front
        hp.n_eval = hp.table_len;
        if (hp.interpolate) hp.cut = hp.N / 2;   // interpolate_alias, base.py:121
        return;
    }
    if (hp.interpolate) {
        // base.py:240-242: formula on arange(0, sfreq/L*(L/2), 1/L), as many zeros appended,
        // pad_to N, then interpolate_alias keeps bins < int(N/2) (base.py:121-123, 276)
        const double total = hp.sfreq / L * (L / 2);
        const long long half = arange_len(total, one);
        long long len = 2 * half;
        hp.grid_off = len >= hp.N ? 0 : (int)((hp.N - len) / 2);
        hp.n_eval = half;
        hp.cut = hp.N / 2;
    } else {
        const double total = hp.sfreq / L * L;   // base.py:193, 244-245
        const long long len = arange_len(total, one);
        hp.grid_off = len >= hp.N ? 0 : (int)((hp.N - len) / 2);
        hp.n_eval = len;
    }
}

// ---- bands ---------------------------------------------------------------------------
inline double morse_logw(double x, double b, double r) { return b * log(x) + (b / r) * (1.0 - pow(x, r)); }

// Grid-index interval [glo, ghi) outside of which |W| < eps * peak.
inline void analytic_band(const HostPlan& hp, int fi, double eps, long long& glo, long long& ghi) {
    const long long nev = hp.n_eval;
    glo = 0;
    ghi = nev;
    const double f = hp.freqs[fi];
    if (hp.family == FAM_SHANNON) {
        // wavelets.py:256-262: grid value <= 1.0; the grid is i * one in fp64
        long long k = (long long)floor(1.0 / hp.df);
        while (k + 1 < nev && (double)(k + 1) * hp.df <= 1.0) ++k;
        while (k >= 0 && (double)k * hp.df > 1.0) --k;
        ghi = std::min(nev, k + 1);
        return;
    }
    if (!(f > 0) || !(eps > 0)) {
        if (hp.family == FAM_MORSE && f > 0) glo = std::min(nev, 1LL);  // heaviside(0,0) = 0
        return;
    }
    // eps is relative to the largest spectrum value ON THE EVALUATED GRID: when the analytic peak
    // lies beyond the grid (analysis frequency above the last bin) only a tail is present and the
    // threshold has to follow it down.  logpk = log(max on grid / analytic peak) <= 0.
    double xlo = 0, xhi = INFINITY;
    if (hp.family == FAM_MORSE) {
        const double b = hp.p0, r = hp.p1;
        if (!(b > 0) || !(r > 0) || nev < 2) { glo = std::min(nev, 1LL); return; }
        const double s = f / hp.df;   // x = g*df/f  ->  g = x * f / df
        const double x_first = 1.0 / s, x_last = (double)(nev - 1) / s;
        double logpk = 0.0;
        if (x_last < 1.0) logpk = morse_logw(x_last, b, r);
        else if (x_first > 1.0) logpk = morse_logw(x_first, b, r);
        const double le = log(eps) + logpk;
        if (!(le > -1e300)) { glo = std::min(nev, 1LL); return; }
        // log(W/2) peaks at x = 1 with value 0 and is monotone on either side
        double a = 0.0, c = 1.0;
        for (int i = 0; i < 200; ++i) { double m = 0.5 * (a + c); (morse_logw(m, b, r) < le ? a : c) = m; }
        xlo = a;
        a = 1.0; c = 2.0;
        while (morse_logw(c, b, r) > le && c < 1e300) c *= 2;
        for (int i = 0; i < 200; ++i) { double m = 0.5 * (a + c); (morse_logw(m, b, r) < le ? c : a) = m; }
        xhi = c;
        glo = (long long)std::min(floor(xlo * s) - 1, (double)nev);
        ghi = (long long)std::min(ceil(xhi * s) + 2, (double)nev);
        glo = std::max(glo, 1LL);
    } else {  // Morlet: |W|/A <= exp(-(s-x)^2/2) + k exp(-x^2/2)
        const double sg = hp.p0, kap = fabs(hp.p2);
        const double s = f / (hp.df * hp.aux[fi]);   // x = g*df/f*pk
        const double x_last = (double)(nev - 1) / s;
        double logpk = 0.0;
        if (x_last < sg) logpk = -0.5 * (sg - x_last) * (sg - x_last);
        const double lt = log(0.5 * eps) + logpk;   // log of the per-term threshold
        if (!(lt > -1e300)) return;
        const double d1 = sqrt(-2.0 * lt);
        xlo = sg - d1;
        xhi = sg + d1;
        if (kap > 0 && log(kap) > lt) {
            const double d2 = sqrt(-2.0 * (lt - log(kap)));
            xlo = std::min(xlo, -d2);
            xhi = std::max(xhi, d2);
        }
        glo = (long long)std::max(-1.0, std::min(floor(xlo * s) - 1, (double)nev));
        ghi = (long long)std::max(-1.0, std::min(ceil(xhi * s) + 2, (double)nev));
    }
    glo = std::max(0LL, std::min(glo, nev));
    ghi = std::max(glo, std::min(ghi, nev));
}

inline void plan_bands(HostPlan& hp) {
    hp.rec.resize(hp.F);
    hp.band_bins = 0;
    for (int i = 0; i < hp.F; ++i) {
        FreqRec r;
        r.freq = hp.freqs[i];
        r.aux = hp.family == FAM_MORLET ? hp.aux[i] : 1.0;
        r.kx = hp.family == FAM_MORLET ? hp.df / r.freq * r.aux : hp.df / r.freq;
        long long lo, hi;
        r.toff = 0;
        r.pad_ = 0;
        if (hp.family == FAM_TABLE) {
            // pad_to, base.py:75-82: truncate, or centre with the short half in front
            const long long m = hp.table_lens.empty() ? hp.table_len : hp.table_lens[i];
            r.toff = m >= hp.N ? 0 : (int)((hp.N - m) / 2);
            lo = r.toff;
            hi = r.toff + std::min(m, hp.N);
        } else {
            long long glo, ghi;
            analytic_band(hp, i, hp.prune_eps, glo, ghi);
            lo = glo + hp.grid_off;
            hi = ghi + hp.grid_off;
        }
        lo = std::max(0LL, std::min(lo, hp.N));
        hi = std::max(lo, std::min(std::min(hi, hp.N), hp.cut));
        lo = std::min(lo, hi);
        r.lo = (int)lo;
        r.hi = (int)hi;
        hp.rec[i] = r;
        hp.band_bins += hi - lo;
    }
}

// ---- execution shape -----------------------------------------------------------------
inline int ilog2_floor(long long v) { int l = 0; while ((1LL << (l + 1)) <= v) ++l; return l; }

inline bool plan_shape(HostPlan& hp, std::string& err, bool force_long = false) {
    const size_t cs = cx_size(hp.dtype);
    const long long N = hp.N;
    // short path: largest interleave that fits (prefer two CTAs per SM)
    FftStages st;
    std::string e;
    if (!force_long && N <= (1 << 15) && factorise(N, st, e)) {
        int fcap = 1;
        while (fcap < hp.F && fcap < 8) fcap <<= 1;
        const size_t caps[2] = {SMEM_HALF, SMEM_MAX};
        for (int c = 0; c < 2; ++c)
            for (int tt = fcap; tt >= 1; tt >>= 1) {
                if (c == 0 && tt < 4 && fcap >= 4) continue;   // do not trade interleave for occupancy below 4
                const int pitch = tt > 1 ? tt + 1 : 1;
                const size_t bytes = (size_t)N * (1 + 2 * (size_t)pitch) * cs + (size_t)tt * (sizeof(FreqRec) + 16);
                if (bytes <= caps[c]) {
                    hp.path = 0;
                    hp.st = st;
                    hp.tsh = ilog2_floor(tt);
                    hp.pitch = pitch;
                    hp.smem_short = bytes;
                    long long work = N * tt / 8;
                    int nt = (int)std::min<long long>(512, std::max<long long>(64, (work + 31) / 32 * 32));
                    hp.nthr_short = nt;
                    return true;
                }
            }
    }
    // long path: N = N1 * N2
    long long best1 = 0;
    int bestTA = 0, bestTB = 0;
    double bestScore = -1e300;
    for (long long d = 2; d * d <= N; ++d) {
        if (N % d) continue;
        const long long cand[2] = {d, N / d};
        for (int w = 0; w < 2; ++w) {
            const long long n1 = cand[w], n2 = N / n1;
            if (w == 1 && n1 == n2) continue;
            if (!smooth_enough(n1) || !smooth_enough(n2)) continue;
            int ta = 0, tb = 0;
            for (int t = 16; t >= 1; t >>= 1)
                if (!ta && 2 * (size_t)n1 * (t + 1) * cs <= SMEM_MAX) ta = t;
            for (int t = 16; t >= 1; t >>= 1)
                if (!tb && 2 * (size_t)n2 * t * cs + 16 <= SMEM_MAX) tb = t;
            if (!ta || !tb) continue;
            ta = std::min(ta, 8);
            tb = std::min(tb, 8);
            // prefer tiles whose smem allows two CTAs per SM when that keeps T >= 8
            double score = 100.0 * std::min(ta, 8) + 100.0 * std::min(tb, 8);
            if (n1 % tb == 0) score += 50.0;
            if (n2 % ta == 0) score += 25.0;
            score -= 10.0 * fabs(log2((double)n1 / (double)n2));
            if (score > bestScore) { bestScore = score; best1 = n1; bestTA = ta; bestTB = tb; }
        }
    }
    if (!best1) { err = "no usable two-pass split for this length (prime factor > 64 or length > 2^30)"; return false; }
    hp.path = 1;
    hp.N1 = (int)best1;
    hp.N2 = (int)(N / best1);
    if (!factorise(hp.N1, hp.stA, err) || !factorise(hp.N2, hp.stB, err)) return false;
    hp.tshA = ilog2_floor(bestTA);
    hp.pitchA = bestTA + 1;
    hp.tshB = ilog2_floor(bestTB);
    hp.smem_A = 2 * (size_t)hp.N1 * hp.pitchA * cs;
    hp.smem_B = 2 * (size_t)hp.N2 * bestTB * cs + 16;
    hp.lb = (ilog2_floor(N) + 2) / 2;
    const long long nblk = (hp.N1 + bestTB - 1) / bestTB;
    hp.tm_stride = nblk * hp.N2 * bestTB;
    const size_t slot = (size_t)hp.tm_stride * cs;
    long long ring = (long long)((48u << 20) / slot);
    hp.ring = (int)std::max<long long>(1, std::min<long long>(ring, 64));
    return true;
}

}  // namespace nw
