// Host-side planner: everything the reference's make_fft_wavelets (base.py:258-279)
// and _setup_trans_shape (base.py:173-194) decide per (family, freqs, N), restated
// for a device that evaluates spectra on the fly: grid geometry, per-frequency
// parameter records, non-zero bands, radix plans, the four-step split and tile
// sizes.  Pure C++ (no CUDA calls) so the CPU test-suite can exercise it through
// the C ABI without a GPU.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <algorithm>

#include <memory>

#include "nw_common.h"
#include "nw_family.cuh"

namespace nw {

struct HostPlan;

// A group of frequencies whose rows are computed at a reduced rate and interpolated ("resampled rows"):
// the band of every row fits M = N / D bins with room to spare, so the inverse transform runs at length M on the band
// moved to bin 0 (FreqRec::shift) and nw_resample.cuh interpolates the M samples to the N outputs with a K-tap
// Kaiser-Bessel kernel whose pass-band response is divided out of the spectrum beforehand (eq).  D == 1: the group's rows
// take the exact length-N transform.
struct MrGroup {
    int D = 1, K = 0;
    double beta = 0, err = 0;          // kernel shape; worst-case relative error of an interpolated row, any input
    std::vector<int> fidx;             // plan frequency index of each of the group's rows
    std::shared_ptr<HostPlan> sub;     // plan of length N / D for those frequencies
    std::vector<double> coef;          // [D][K] interpolation weights, phase-major
    std::vector<int> t0;               // [D] offset of the first tap of each phase
    std::vector<double> eq;            // D / H(j), j = 0 .. eq.size() - 1
};

static const size_t SMEM_MAX = 227 * 1024;       // opt-in dynamic shared memory per CTA (sm_100)
static const size_t SMEM_HALF = 113 * 1024;      // two CTAs per SM

struct HostPlan {
    // description
    int device = 0, dtype = 0, family = 0, interpolate = 0;
    long long N = 0;             // transform length
    long long Nd = 0;            // data length: frequency grid, spectra pitch, 1/Nd normalisation (0: = N; a resampled group's plan: N * D)
    int F = 0;
    double sfreq = 0, p0 = 0, p1 = 0, p2 = 0, prune_eps = 0;
    int resample = -1;           // < 0: library default (on), 0: exact transforms only
    double resample_tol = 0;     // worst-case relative error allowed to the interpolation (0: default per dtype)
    std::vector<MrGroup> groups; // empty: every row takes the exact transform
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
    // fused short-row kernel on the packed engine (nw_kernels3.cuh); 0 = not available (generic kernel is used)
    int short2 = 0, tpshS = 0, nthrS2 = 256;
    Fft2Plan stS{};
    size_t smem_S2 = 0;
    // the same with resampled rows (nw_kernels4.cuh; groups hold the per-decimation sub-plans): 0 = not used
    int short3 = 0, tpshS3 = 0, nthrS3 = 256, yslotsS3 = 0, tpitchS3 = 0;
    size_t smem_S3 = 0;
    // fast long path (packed in-place engine, nw_kernels2.cuh); 0 = not available for this N
    int fast = 0;
    int N1f = 0, N2f = 0, tpshA = 0, tpshB = 0;
    Fft2Plan stA2{}, stB2{};
    int nthrA2 = 256, nthrB2 = 256;
    int cfgA = 0, cfgB = 0;      // compiled launch shape (CFG2_*) of each pass
    int generic_ok = 1;          // long path: the generic two-pass kernels also have a plan for this N
    int narrowA = 0;             // every band touches <= N1 / R_last rows: pass A's first pass needs no loads (nw_kernels2.cuh)
    std::vector<char> narrow_ok; // the same per frequency: a launch group whose rows all qualify takes the narrow kernel
    size_t smem_A2 = 0, smem_B2 = 0;
    long long tm_stride2 = 0;
    int ring2 = 1;               // rows per pass-A / pass-B launch pair
    long long data_len() const { return Nd > 0 ? Nd : N; }
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

// ---- radix plan of the packed in-place engine (nw_fft2.cuh) -----------------------------------
// P = 2^a 3^b 5^c is covered by radices {16,15,12,10,8,6,5,4,3,2}: fewest stages first (every stage is
// one round trip of the tile through shared memory), then least butterfly arithmetic per point.
// An odd radix, when there is one, goes last: the stride-1 pass then reads R-element groups whose
// shared-memory footprint is conflict free for two butterflies per quarter warp.
inline void fill_plan2(long long P, const std::vector<int>& rad, Fft2Plan& out) {
    out.P = (int)P;
    out.nst = (int)rad.size();
    int ns = 1;
    for (int i = 0; i < out.nst; ++i) {
        out.radix[i] = rad[i];
        out.ns[i] = ns;
        const int L = (int)(P / ns);
        out.div_q[i] = make_fastdiv((uint32_t)(L / rad[i]));
        out.div_r[i] = make_fastdiv((uint32_t)rad[i]);
        ns *= rad[i];
    }
}

inline int env_int(const char* name, int dflt);

inline bool plan_packed(long long P, Fft2Plan& out) {
    if (P < 2 || P > (1 << 16)) return false;
    int e[3] = {0, 0, 0};
    long long m = P;
    while (m % 2 == 0) { m /= 2; ++e[0]; }
    while (m % 3 == 0) { m /= 3; ++e[1]; }
    while (m % 5 == 0) { m /= 5; ++e[2]; }
    if (m != 1) return false;
    static const int RAD[10] = {16, 15, 12, 10, 8, 6, 5, 4, 3, 2};
    static const int EXP[10][3] = {{4,0,0},{0,1,1},{2,1,0},{1,0,1},{3,0,0},{1,1,0},{0,0,1},{2,0,0},{0,1,0},{1,0,0}};
    // packed ops per point of each butterfly plus ~3.5 for the twiddle multiply of a non-final stage
    static const double COST[10] = {10.0, 11.2, 7.0, 9.2, 7.0, 6.0, 7.2, 4.0, 4.0, 2.0};
    struct Best { int stages; double cost; std::vector<int> rad; };
    const int A = e[0] + 1, B = e[1] + 1, C = e[2] + 1;
    std::vector<Best> best((size_t)A * B * C);
    std::vector<char> done((size_t)A * B * C, 0);
    auto idx = [&](int a, int b, int c) { return ((size_t)a * B + b) * C + c; };
    best[idx(0, 0, 0)] = Best{0, 0.0, {}};
    done[idx(0, 0, 0)] = 1;
    for (int a = 0; a < A; ++a)
        for (int b = 0; b < B; ++b)
            for (int c = 0; c < C; ++c) {
                if (a == 0 && b == 0 && c == 0) continue;
                Best bb{1 << 20, 1e300, {}};
                for (int i = 0; i < 10; ++i) {
                    const int pa = a - EXP[i][0], pb = b - EXP[i][1], pc = c - EXP[i][2];
                    if (pa < 0 || pb < 0 || pc < 0 || !done[idx(pa, pb, pc)]) continue;
                    const Best& prev = best[idx(pa, pb, pc)];
                    const int st = prev.stages + 1;
                    const double cost = prev.cost + COST[i] + 3.5;
                    if (st < bb.stages || (st == bb.stages && cost < bb.cost - 1e-9)) {
                        bb.stages = st;
                        bb.cost = cost;
                        bb.rad = prev.rad;
                        bb.rad.push_back(RAD[i]);
                    }
                }
                if (bb.stages < (1 << 20)) { best[idx(a, b, c)] = bb; done[idx(a, b, c)] = 1; }
            }
    const Best& r = best[idx(e[0], e[1], e[2])];
    if (!done[idx(e[0], e[1], e[2])] || r.stages > MAX_STAGES2) return false;
    std::vector<int> rad = r.rad;
    std::sort(rad.begin(), rad.end(), [](int x, int y) { return x > y; });
    for (size_t i = 0; i < rad.size(); ++i)
        if (rad[i] & 1) { std::swap(rad[i], rad.back()); break; }   // one odd radix last
    fill_plan2(P, rad, out);
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
    const long long N = hp.data_len();
    const double L = (double)N / hp.sfreq;      // wave.shape[0] / self.sfreq, base.py:395
    const double one = 1 / L;                    // base.py:192
    hp.df = one;
    hp.cut = N;
    hp.grid_off = 0;
    if (hp.family == FAM_TABLE) {
        // pad_to, base.py:75-82: truncate, or centre with the short half in front
        hp.n_eval = hp.table_len;
        if (hp.interpolate) hp.cut = N / 2;   // interpolate_alias, base.py:121
        return;
    }
    if (hp.interpolate) {
        // base.py:240-242: formula on arange(0, sfreq/L*(L/2), 1/L), as many zeros appended,
        // pad_to N, then interpolate_alias keeps bins < int(N/2) (base.py:121-123, 276)
        const double total = hp.sfreq / L * (L / 2);
        const long long half = arange_len(total, one);
        long long len = 2 * half;
        hp.grid_off = len >= N ? 0 : (int)((N - len) / 2);
        hp.n_eval = half;
        hp.cut = N / 2;
    } else {
        const double total = hp.sfreq / L * L;   // base.py:193, 244-245
        const long long len = arange_len(total, one);
        hp.grid_off = len >= N ? 0 : (int)((N - len) / 2);
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
    const long long N = hp.data_len();
    hp.rec.resize(hp.F);
    hp.band_bins = 0;
    for (int i = 0; i < hp.F; ++i) {
        FreqRec r;
        r.freq = hp.freqs[i];
        r.aux = hp.family == FAM_MORLET ? hp.aux[i] : 1.0;
        r.kx = hp.family == FAM_MORLET ? hp.df / r.freq * r.aux : hp.df / r.freq;
        long long lo, hi;
        r.toff = 0;
        r.shift = 0;
        r.woff = 0;
        if (hp.family == FAM_TABLE) {
            // pad_to, base.py:75-82: truncate, or centre with the short half in front
            const long long m = hp.table_lens.empty() ? hp.table_len : hp.table_lens[i];
            r.toff = m >= N ? 0 : (int)((N - m) / 2);
            lo = r.toff;
            hi = r.toff + std::min(m, N);
            // interpolate_alias is applied to the table at ITS length before pad_to (base.py:276, 121-123)
            if (hp.interpolate) hi = std::min(hi, r.toff + m / 2);
        } else {
            long long glo, ghi;
            analytic_band(hp, i, hp.prune_eps, glo, ghi);
            lo = glo + hp.grid_off;
            hi = ghi + hp.grid_off;
        }
        lo = std::max(0LL, std::min(lo, N));
        hi = std::max(lo, std::min(std::min(hi, N), hp.cut));
        lo = std::min(lo, hi);
        r.lo = (int)lo;
        r.hi = (int)hi;
        hp.rec[i] = r;
        hp.band_bins += hi - lo;
    }
}

// ---- execution shape -----------------------------------------------------------------
inline int ilog2_floor(long long v) { int l = 0; while ((1LL << (l + 1)) <= v) ++l; return l; }

// Launch shapes the fast kernels are compiled for: {max threads, min resident CTAs} -> register cap.
// A tile's thread count and shape are chosen together: every stage has (P / R_s) << tpsh butterflies;
// take the multiple of 32 that wastes the fewest thread slots over all stages, weighted by how many
// threads the shape keeps resident per SM (shared memory and registers), with a small bonus for the
// shape with more registers per thread.
static const int N_CFG2 = 4;
static const int CFG2_MAXTHR[N_CFG2] = {256, 224, 128, 64};
static const int CFG2_MINCTA[N_CFG2] = {3, 3, 5, 10};
static const int CFG2_MAXREG[N_CFG2] = {80, 96, 96, 96};   // __maxnreg__ of the compiled kernels
inline int cfg2_regcap(int c) { return CFG2_MAXREG[c]; }

inline void pick_threads2(const Fft2Plan& st, int tpsh, size_t smem, int ncfg, int& nthr, int& cfg) {
    // Measured on B200 (profiles/r01/shape_sweep.md): with room for >= 5 tiles per SM, 128-thread CTAs at
    // 96 registers win (more independent CTAs to overlap load, butterfly and store phases); with 3-4 tiles
    // per SM the most threads win over more registers; tiny tiles take 64-thread CTAs.
    const int by_smem = (int)std::min<size_t>(16, (SMEM_MAX + 1024) / (smem + 1024));
    long long maxnb = 0;
    for (int s = 0; s < st.nst; ++s) maxnb = std::max(maxnb, (long long)(st.P / st.radix[s]) << tpsh);
    cfg = 0;
    if (ncfg > 1) {
        if (maxnb <= 64 && by_smem >= 8) cfg = 3;
        else if (by_smem >= 5 || maxnb <= 128) cfg = 2;
    }
    const int maxthr = CFG2_MAXTHR[cfg];
    if (cfg == 0) {   // fill the shape: more warps beat a better fit of butterflies to threads here
        nthr = (int)std::min<long long>(maxthr, (maxnb + 31) / 32 * 32);
        // a lone CTA per SM with at least two butterflies per thread left: 512 threads (measured 2^26: 349 -> 324 ms)
        if (by_smem == 1 && maxnb >= 1024) nthr = 512;
        return;
    }
    const int lo = std::max(32, (int)std::min<long long>(maxthr * 3 / 4, (maxnb + 31) / 32 * 32) / 32 * 32);
    double best = -1;
    nthr = maxthr;
    for (int nt = maxthr; nt >= lo; nt -= 32) {   // larger first: ties go to more threads
        double work = 0, slots = 0;
        for (int s = 0; s < st.nst; ++s) {
            const long long nb = (long long)(st.P / st.radix[s]) << tpsh;
            work += (double)nb * st.radix[s];
            slots += (double)((nb + nt - 1) / nt) * nt * st.radix[s];
        }
        const double e = work / slots * (0.5 + 0.5 * nt / maxthr);
        if (e > best + 1e-9) { best = e; nthr = nt; }
    }
}

inline int env_int(const char* name, int dflt);

// Narrow-band first pass of pass A (nw_kernels2.cuh): allowed for a frequency whose band touches at most
// N1 / R_last rows of any column tile.
inline long long floor_div_ll(long long a, long long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
inline void plan_narrow(HostPlan& hp) {
    hp.narrowA = 0;
    hp.narrow_ok.clear();
    if (hp.fast && hp.F > 0 && hp.stA2.nst >= 2 && !env_int("NWCWT_NO_NARROW", 0)) {
        const int step = hp.N1f / hp.stA2.radix[hp.stA2.nst - 1];
        int worst = 0;
        hp.narrow_ok.assign((size_t)hp.F, 0);
        for (int i = 0; i < hp.F; ++i) {
            const FreqRec& r = hp.rec[i];
            // rows a tile of 2 << tpshA columns can touch: one more than the band's own row span
            const int C = r.hi > r.lo ? (int)(floor_div_ll(r.hi - 1, hp.N2f) - floor_div_ll(r.lo, hp.N2f)) + 2 : 0;
            worst = std::max(worst, C);
            hp.narrow_ok[(size_t)i] = C <= step;
        }
        hp.narrowA = worst <= step;
    }
}

// May the launch group of rows [r0, r0 + g) (row = signal * F + frequency) use the narrow-band first pass of pass A?
inline bool group_narrow(const HostPlan& hp, long long r0, int g) {
    if (hp.narrow_ok.empty()) return false;
    if (hp.narrowA) return true;
    if (g >= hp.F) return false;
    for (int i = 0; i < g; ++i)
        if (!hp.narrow_ok[(size_t)((r0 + i) % hp.F)]) return false;
    return true;
}

inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// Fast long path: N = N1 * N2 with both factors 2-3-5 smooth; tiles sized for three CTAs per SM.
inline void plan_shape_fast(HostPlan& hp) {
    const size_t cs = cx_size(hp.dtype);
    const long long N = hp.N;
    hp.fast = 0;
    double bestScore = -1e300;
    auto ctas = [&](size_t bytes) { return bytes <= 75 * 1024 ? 3 : bytes <= SMEM_HALF ? 2 : bytes <= SMEM_MAX ? 1 : 0; };
    for (long long d = 2; d * d <= N; ++d) {
        if (N % d) continue;
        const long long cand[2] = {d, N / d};
        for (int w = 0; w < 2; ++w) {
            const long long n1 = cand[w], n2 = N / n1;
            if (w == 1 && n1 == n2) continue;
            if (n1 > 32768 || n2 > 32768) continue;
            if (env_int("NWCWT_SPLIT_N1", 0) > 0 && n1 != env_int("NWCWT_SPLIT_N1", 0)) continue;   // tuning override
            Fft2Plan a, b;
            if (!plan_packed(n1, a) || !plan_packed(n2, b)) continue;
            const int fa = env_int("NWCWT_TPSH_A", -1), fb = env_int("NWCWT_TPSH_B", -1);   // tuning overrides
            for (int ta = 4; ta >= 0; --ta)
                for (int tb = 4; tb >= 0; --tb) {
                    if ((fa >= 0 && ta != fa) || (fb >= 0 && tb != fb)) continue;
                    // 8 and 16 lane pairs per tile only for short transforms (the decimated lengths of resampled rows split
                    // into factors of 75 .. 250): tiles of up to 2048 two-lane values, so that a CTA has enough butterflies
                    // per stage to fill its warps
                    if ((ta > 2 && ((size_t)n1 << ta) > 2048) || (tb > 2 && ((size_t)n2 << tb) > 2048)) continue;
                    if ((ta > 2 || tb > 2) && !env_int("NWCWT_BIG_TILES", 0)) continue;   // measured slower on cfg2's decimated lengths (profiles/r02)
                    const size_t ba = ((size_t)n1 << ta) * 2 * cs, bb = ((size_t)n2 << tb) * 2 * cs + 16;
                    const int ca = ctas(ba), cb = ctas(bb);
                    if (!ca || !cb) continue;
                    // lane pairs per tile: the step from one pair to two is worth more than from two to four
                    // (measured: fp64 2^20, 1024 x 1024 with two pairs each beats 2048 x 512 with one and four)
                    static const double TPS_SCORE[5] = {0.0, 120.0, 200.0, 260.0, 300.0};
                    // resident CTAs matter more to pass A, the run length of the intermediate more to pass B (measured: 2^24 fp32
                    // and 2^22 fp64 are fastest with one pair / three CTAs in pass A and two pairs / one CTA in pass B)
                    double score = TPS_SCORE[ta] + TPS_SCORE[tb] + 80.0 * ca + 55.0 * cb - 40.0 * (a.nst + b.nst);
                    for (int i = 0; i < a.nst; ++i) if (a.radix[i] >= 15) score -= 15.0;   // register pressure (measured: cfg5 2^18 sweep)
                    for (int i = 0; i < b.nst; ++i) if (b.radix[i] >= 15) score -= 15.0;
                    if (n1 & 1) score -= 30.0;                 // unaligned output pairs
                    if (n2 >= n1) score += 5.0;
                    // pass A writes the intermediate in runs of (2 << tb) values: below 64 bytes the stores waste sectors
                    for (size_t run = ((size_t)2 << tb) * cs; run < 64; run *= 2) score -= 20.0;
                    score -= 4.0 * fabs(log2((double)n1 / (double)n2));
                    if (score > bestScore) {
                        bestScore = score;
                        hp.fast = 1;
                        hp.N1f = (int)n1; hp.N2f = (int)n2; hp.tpshA = ta; hp.tpshB = tb;
                        hp.stA2 = a; hp.stB2 = b;
                        hp.smem_A2 = ba; hp.smem_B2 = bb;
                    }
                }
        }
    }
    if (!hp.fast) return;
    const int ncfg = hp.dtype == 0 ? 4 : 1;   // fp64 kernels exist in shape 0 only
    pick_threads2(hp.stA2, hp.tpshA, hp.smem_A2, ncfg, hp.nthrA2, hp.cfgA);
    pick_threads2(hp.stB2, hp.tpshB, hp.smem_B2, ncfg, hp.nthrB2, hp.cfgB);
    if (hp.dtype == 1) {   // fp64 kernels: 168 registers, three 128-thread CTAs per SM (measured: +6 % cfg2, +20 % 2^20 over 256 x 2 at 128)
        // tiles that leave room for only two or one CTA per SM take 192 / 384 threads (the register file holds 390 x 168)
        auto cap64 = [](size_t smem, const Fft2Plan& st, int tpsh) {
            long long maxnb = 0;
            for (int s = 0; s < st.nst; ++s) maxnb = std::max(maxnb, (long long)(st.P / st.radix[s]) << tpsh);
            const int fit = (int)((SMEM_MAX + 1024) / (smem + 1024));
            const int cap = fit >= 3 ? 128 : fit == 2 ? 192 : 384;
            if (fit == 1) return cap;   // a lone CTA: more warps than butterflies still pays (tile load, epilogue; measured 2^24)
            return (int)std::min<long long>(cap, (maxnb + 31) / 32 * 32);
        };
        hp.nthrA2 = cap64(hp.smem_A2, hp.stA2, hp.tpshA);
        hp.nthrB2 = cap64(hp.smem_B2, hp.stB2, hp.tpshB);
    }
    // tuning overrides, validated against what the kernels are compiled for: a launch shape that exists, and a whole
    // number of warps within the shape's launch bound (fp64: 512, the bound of its single shape)
    if (hp.dtype == 0) {   // launch shapes exist in fp32 only
        const int ca = env_int("NWCWT_CFG_A", hp.cfgA), cb = env_int("NWCWT_CFG_B", hp.cfgB);
        if (ca >= 0 && ca < N_CFG2) hp.cfgA = ca;
        if (cb >= 0 && cb < N_CFG2) hp.cfgB = cb;
    }
    auto clamp_thr = [&](int v, int cfg) {
        const int cap = hp.dtype == 0 ? (cfg == 0 ? 512 : CFG2_MAXTHR[cfg]) : 512;
        v = (v + 31) / 32 * 32;
        return v < 32 ? 32 : v > cap ? cap : v;
    };
    hp.nthrA2 = clamp_thr(env_int("NWCWT_NTHR_A", hp.nthrA2), hp.cfgA);
    hp.nthrB2 = clamp_thr(env_int("NWCWT_NTHR_B", hp.nthrB2), hp.cfgB);
    const int TB = 2 << hp.tpshB;
    const long long nblk = (hp.N1f + TB - 1) / TB;
    hp.tm_stride2 = nblk * hp.N2f * TB;
    const size_t slot = (size_t)hp.tm_stride2 * cs;
    size_t ring_mb = 48;                                  // per stream slot; two slots are in flight
    if (const char* e = getenv("NWCWT_RING_MB")) ring_mb = (size_t)std::max(1, atoi(e));
    long long ring = (long long)((ring_mb << 20) / slot);
    hp.ring2 = (int)std::max<long long>(1, std::min<long long>(ring, 256));
    if (env_int("NWCWT_RING_ROWS", 0) > 0) hp.ring2 = std::min(hp.ring2, env_int("NWCWT_RING_ROWS", 0));   // tests: small launch groups
}

// Packed-engine short kernel: N 2-3-5 smooth, N (1 + NF) two-lane complex values of shared memory with NF = 1, 2, 4
// or 8 frequencies per pass - as many as keep three CTAs per SM (75 KB), at least one.
inline void plan_shape_short2(HostPlan& hp) {
    hp.short2 = 0;
    if (hp.N < 8 || hp.N > 16384 || env_int("NWCWT_NO_SHORT2", 0)) return;
    Fft2Plan st;
    if (!plan_packed(hp.N, st)) return;
    const size_t c2 = 2 * cx_size(hp.dtype);
    int fcap = 0;
    while ((1 << fcap) < hp.F && fcap < 3) ++fcap;
    const int forced = env_int("NWCWT_TPSH_S", -1);
    int pick = -1;
    for (int t = fcap; t >= 0 && pick < 0; --t)
        if ((size_t)hp.N * (1 + ((size_t)1 << t)) * c2 + 1280 <= 75 * 1024) pick = t;
    if (pick < 0 && (size_t)hp.N * 2 * c2 + 1280 <= SMEM_MAX) pick = 0;
    if (forced >= 0 && (size_t)hp.N * (1 + ((size_t)1 << forced)) * c2 + 1280 <= SMEM_MAX) pick = forced;
    if (pick < 0) return;
    hp.short2 = 1;
    hp.tpshS = pick;
    hp.stS = st;
    hp.smem_S2 = (size_t)hp.N * (1 + ((size_t)1 << pick)) * c2 + 1280;
    long long maxnb = 0;
    for (int s = 0; s < st.nst; ++s) maxnb = std::max(maxnb, (long long)(st.P / st.radix[s]) << pick);
    hp.nthrS2 = (int)std::min<long long>(256, std::max<long long>(64, (maxnb + 31) / 32 * 32));
    {   // tuning override: a whole number of warps, at most the kernel's launch bound
        int v = (env_int("NWCWT_NTHR_S", hp.nthrS2) + 31) / 32 * 32;
        hp.nthrS2 = v < 32 ? 32 : v > 256 ? 256 : v;
    }
}

inline void plan_multirate(HostPlan& hp, bool shortrows);
inline void plan_shape_short3(HostPlan& hp);

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
                    plan_shape_short2(hp);
                    if (hp.short2 && env_int("NWCWT_SHORT3", 1)) {   // band-limited rows at decimated lengths (nw_kernels4.cuh); NWCWT_SHORT3=0: exact kernel only
                        plan_multirate(hp, true);
                        plan_shape_short3(hp);
                    }
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
            score -= 4.0 * fabs(log2((double)n1 / (double)n2));
            if (score > bestScore) { bestScore = score; best1 = n1; bestTA = ta; bestTB = tb; }
        }
    }
    hp.path = 1;
    hp.lb = (ilog2_floor(N) + 2) / 2;
    plan_shape_fast(hp);
    if (!best1) {
        if (!hp.fast) { err = "no usable two-pass split for this length (prime factor > 64 or length > 2^30)"; return false; }
        // only the packed kernels can take this length (tiles too large for the generic engine's two buffers)
        hp.generic_ok = 0;
        hp.ring = (int)std::max<long long>(1, std::min<long long>(64, (256LL << 20) / (N * (long long)cs)));
        plan_narrow(hp);
        plan_multirate(hp, false);
        return true;
    }
    hp.N1 = (int)best1;
    hp.N2 = (int)(N / best1);
    if (!factorise(hp.N1, hp.stA, err) || !factorise(hp.N2, hp.stB, err)) return false;
    hp.tshA = ilog2_floor(bestTA);
    hp.pitchA = bestTA + 1;
    hp.tshB = ilog2_floor(bestTB);
    hp.smem_A = 2 * (size_t)hp.N1 * hp.pitchA * cs;
    hp.smem_B = 2 * (size_t)hp.N2 * bestTB * cs + 16;
    const long long nblk = (hp.N1 + bestTB - 1) / bestTB;
    hp.tm_stride = nblk * hp.N2 * bestTB;
    const size_t slot = (size_t)hp.tm_stride * cs;
    long long ring = (long long)((48u << 20) / slot);
    hp.ring = (int)std::max<long long>(1, std::min<long long>(ring, 64));
    plan_narrow(hp);
    plan_multirate(hp, false);
    return true;
}


// ---- resampled rows ---------------------------------------------------------------------------------------
// A row whose band [lo, hi) holds B bins is band limited: its N samples are determined by M >= B of them.  With
// M = N / D the decimated row  y[m] = z(m D) e^{-2 pi i kc m / M}  (kc = band centre) is the M-point inverse transform
// of the band moved to bin 0, and  |z(m D + p)| = | sum_t phi(p / D - (t0_p + t)) y[m + t0_p + t] |  up to the images
// of the band that the interpolation kernel phi lets through.  phi is a Kaiser-Bessel window of K taps; its own
// frequency response H is divided out of the spectrum (MrGroup::eq), so the pass band is exact and the error of a row
// is bounded - for ANY input - by  max_j sqrt(sum_{r != 0} |H(j + r M)|^2) / |H(j)|  over the band, which the planner
// evaluates for the kernel it picks (MrGroup::err) and keeps below resample_tol.

inline double bessel_i0(double x) {   // power series; x <= ~60 here
    const double q = 0.25 * x * x;
    double t = 1.0, s = 1.0;
    for (int k = 1; k < 500; ++k) {
        t *= q / ((double)k * (double)k);
        s += t;
        if (t < 1e-18 * s) break;
    }
    return s;
}

// phi(tau), |tau| <= K / 2: (I0(beta sqrt(1 - (2 tau / K)^2)) - 1) / (I0(beta) - 1), zero at the ends of its support
inline double kb_phi(double tau, int K, double beta, double inv_i0m1) {
    const double u = 2.0 * tau / (double)K, a = 1.0 - u * u;
    if (!(a > 0.0)) return 0.0;
    return (bessel_i0(beta * sqrt(a)) - 1.0) * inv_i0m1;
}

// H(j) = sum_i phi(i / D) cos(2 pi j i / (M D)) for the listed bins j (signed, any size)
inline void kb_response(int D, int K, double beta, long long M, const std::vector<long long>& bins, std::vector<double>& H) {
    const long long half = (long long)K * D / 2;
    const double inv = 1.0 / (bessel_i0(beta) - 1.0);
    std::vector<double> h((size_t)half + 1);
    for (long long i = 0; i <= half; ++i) h[(size_t)i] = kb_phi((double)i / (double)D, K, beta, inv);
    const long long Nf = M * (long long)D;
    const long double tp = 6.283185307179586476925286766559005768L;
    H.resize(bins.size());
    for (size_t b = 0; b < bins.size(); ++b) {
        long long j = bins[b] % Nf;
        if (j < 0) j += Nf;
        const double th = (double)(tp * (long double)j / (long double)Nf);
        // cos(i th) by the Chebyshev recurrence, restarted from an exact value every 64 terms
        double acc = h[0];
        for (long long i0 = 1; i0 <= half; i0 += 64) {
            const long long i1 = std::min(half, i0 + 63);
            double c0 = cos((double)(i0 - 1) * th), c1 = cos((double)i0 * th);
            const double k2 = 2.0 * cos(th);
            for (long long i = i0; i <= i1; ++i) {
                acc += 2.0 * h[(size_t)i] * c1;
                const double c2 = k2 * c1 - c0;
                c0 = c1;
                c1 = c2;
            }
        }
        H[b] = acc;
    }
}

// ---- error model ---------------------------------------------------------------------------------------------
// A component of a row at transform bin j (offset from the band centre) reaches the output with its images at
// j + r M, r = 1 .. D - 1, attenuated by H(j + r M) / H(j).  The planner bounds, for every row,
//     max_j  |W_f(j)| / max|W_f|  *  sqrt(sum_r H(j + r M)^2) / |H(j)|   <=  resample_tol:
// for ANY input, the interpolation error of the row is at most resample_tol times (largest spectrum value of the row's
// wavelet) x (sum of the input's spectral magnitudes) - the scale the rounding error of the exact fp32 / fp64 transform
// itself lives on.  Bins whose weight is below resample_tol / 8 are dropped from a resampled row's band.

// alias gain A(j) and relative pass-band response Hr(j) = |H(j) / H(0)| of the kernel (D, K, beta) at NA + 1 offsets
// j = i * (M / 2) / NA
static const int KB_NA = 64;
struct KbCurve { std::vector<double> A, Hr; };
inline void kb_alias_curve(int D, int K, double beta, long long M, KbCurve& c) {
    std::vector<long long> bins;
    bins.reserve((size_t)(KB_NA + 1) * D);
    for (int i = 0; i <= KB_NA; ++i) {
        const long long j = (long long)((double)(M / 2) * (double)i / (double)KB_NA);
        for (int r = 0; r < D; ++r) bins.push_back(j + (long long)r * M);
    }
    std::vector<double> H;
    kb_response(D, K, beta, M, bins, H);
    c.A.resize(KB_NA + 1);
    c.Hr.resize(KB_NA + 1);
    for (int i = 0; i <= KB_NA; ++i) {
        double al = 0.0;
        for (int r = 1; r < D; ++r) al += H[(size_t)i * D + r] * H[(size_t)i * D + r];
        const double main = fabs(H[(size_t)i * D]);
        c.A[(size_t)i] = main > 0 ? sqrt(al) / main : 1e300;
        c.Hr[(size_t)i] = fabs(H[0]) > 0 ? main / fabs(H[0]) : 0.0;
    }
}
// upper envelope of A (lower envelope of Hr) on the cell that holds offset |j| (both are smooth on the scale of a cell)
inline void kb_curve_at(const KbCurve& c, long long M, long long j, double& a, double& hr) {
    if (j < 0) j = -j;
    const double u = (double)j / (double)(M / 2) * (double)KB_NA;
    if (!(u < (double)KB_NA)) { a = 1e300; hr = 0.0; return; }   // at or beyond M / 2: the image is as close as the component itself
    const int i = (int)u;
    a = std::max(c.A[(size_t)i], c.A[(size_t)i + 1]);
    hr = std::min(c.Hr[(size_t)i], c.Hr[(size_t)i + 1]);
}

// magnitude (any positive scale per frequency) of frequency fi's spectrum at data bin k
inline double spec_mag_host(const HostPlan& hp, int fi, long long k) {
    if (hp.family == FAM_TABLE) {
        const long long m = hp.table_lens.empty() ? hp.table_len : hp.table_lens[(size_t)fi];
        const long long idx = k - hp.rec[(size_t)fi].toff;
        if (idx < 0 || idx >= m) return 0.0;
        const double* t = &hp.table[2 * ((size_t)fi * (size_t)hp.table_len + (size_t)idx)];
        return hypot(t[0], t[1]);
    }
    const double g = (double)(k - hp.grid_off) * hp.df, f = hp.freqs[(size_t)fi];
    if (hp.family == FAM_MORSE) {
        const double x = g / f;
        return x > 0 ? exp(morse_logw(x, hp.p0, hp.p1)) : 0.0;
    }
    if (hp.family == FAM_MORLET) {
        const double x = g / f * hp.aux[(size_t)fi], d = hp.p0 - x;
        return fabs(exp(-0.5 * d * d) - hp.p2 * exp(-0.5 * x * x));
    }
    return 1.0;   // Shannon: flat inside the band
}

// Weight profile of one row: cells [pos[s], pos[s + 1]) of its band with an upper bound w[s] of |W| / max|W| on the cell,
// and the resample band [rlo, rhi) outside of which the weight is below eps_rs.
struct RowProfile {
    std::vector<long long> pos;
    std::vector<double> w;      // upper bound of |W| / max|W| on the cell
    std::vector<double> w2;     // lower bound of the cell's mean (|W| / max|W|)^2
    long long rlo = 0, rhi = 0;
};
inline void row_profile(const HostPlan& hp, int fi, double eps_rs, RowProfile& rp) {
    const FreqRec& r = hp.rec[(size_t)fi];
    const long long lo = r.lo, hi = r.hi, B = hi - lo;
    rp.pos.clear(); rp.w.clear(); rp.w2.clear();
    rp.rlo = rp.rhi = lo;
    if (B <= 0) return;
    const int NS = (int)std::min<long long>(B, 384);
    std::vector<double> mag((size_t)NS + 1);
    rp.pos.resize((size_t)NS + 1);
    for (int s = 0; s <= NS; ++s) rp.pos[(size_t)s] = lo + (long long)((double)B * (double)s / (double)NS);
    rp.pos[(size_t)NS] = hi;
    rp.w.assign((size_t)NS, 0.0);
    rp.w2.assign((size_t)NS, 0.0);
    double peak = 0.0;
    if (hp.family == FAM_TABLE) {   // arbitrary shape: true maximum of every cell
        for (int s = 0; s < NS; ++s) {
            double m = 0.0, q = 0.0;
            for (long long k = rp.pos[(size_t)s]; k < rp.pos[(size_t)s + 1]; ++k) {
                const double v = spec_mag_host(hp, fi, k);
                m = std::max(m, v);
                q += v * v;
            }
            rp.w[(size_t)s] = m;
            rp.w2[(size_t)s] = q / (double)std::max<long long>(1, rp.pos[(size_t)s + 1] - rp.pos[(size_t)s]);
            peak = std::max(peak, m);
        }
    } else {                        // smooth, single-peaked (Morlet: two-term) spectra: cell ends, the maximum refined by bisection
        for (int s = 0; s <= NS; ++s) mag[(size_t)s] = spec_mag_host(hp, fi, std::min(rp.pos[(size_t)s], hi - 1));
        int sm = 0;
        for (int s = 0; s <= NS; ++s) if (mag[(size_t)s] > mag[(size_t)sm]) sm = s;
        long long a = rp.pos[(size_t)std::max(0, sm - 1)], c = std::min(hi - 1, rp.pos[(size_t)std::min(NS, sm + 1)]);
        peak = mag[(size_t)sm];
        while (c - a > 2) {   // ternary search on the unimodal piece
            const long long m1 = a + (c - a) / 3, m2 = c - (c - a) / 3;
            const double v1 = spec_mag_host(hp, fi, m1), v2 = spec_mag_host(hp, fi, m2);
            peak = std::max(peak, std::max(v1, v2));
            if (v1 < v2) a = m1; else c = m2;
        }
        for (int s = 0; s < NS; ++s) {
            rp.w[(size_t)s] = std::max(mag[(size_t)s], mag[(size_t)s + 1]);
            const double lo2 = std::min(mag[(size_t)s], mag[(size_t)s + 1]);
            rp.w2[(size_t)s] = lo2 * lo2;
        }
        for (int s = std::max(0, sm - 1); s < std::min(NS, sm + 1); ++s) rp.w[(size_t)s] = peak;
    }
    if (!(peak > 0.0)) { rp.pos.clear(); rp.w.clear(); rp.w2.clear(); return; }
    for (double& v : rp.w) v /= peak;
    for (double& v : rp.w2) v /= peak * peak;
    int s0 = 0, s1 = NS;
    while (s0 < s1 && rp.w[(size_t)s0] < eps_rs) ++s0;
    while (s1 > s0 && rp.w[(size_t)s1 - 1] < eps_rs) --s1;
    rp.rlo = rp.pos[(size_t)s0];
    rp.rhi = rp.pos[(size_t)s1];
}
// Error of a row for the alias curve A of a kernel at decimated length M (the row's band is centred): the larger of
//   * the weighted worst case  max_j w(j) A(j)  - any input, relative to the wavelet's peak gain x the input's spectral mass;
//   * the relative L2 error for an input with a flat spectrum,  sqrt(sum_j w(j)^2 A(j)^2 / sum_j w(j)^2)  - what a
//     per-row relative L2 comparison of a broadband signal sees (it exceeds the first when the spectrum has a long
//     low-level tail, e.g. the |re| + i |im| tables of MexicanHat).
//   * and the equaliser must not lift any part of the spectrum above RS_MAX_LIFT times its peak: the rounding noise of the
//     decimated transform is relative to the LARGEST equalised component, and the interpolation passes it at full gain in
//     the middle of the band (a spectrum with its weight at the band edges, like MexicanHat's tables, would otherwise
//     get its noise amplified by 1 / Hr(edge)).
static const double RS_MAX_LIFT = 3.0;
inline double row_error(const RowProfile& rp, const KbCurve& c, long long M) {
    const long long kc = rp.rlo + (rp.rhi - rp.rlo) / 2;
    double e = 0.0, num = 0.0, den = 0.0;
    for (size_t s = 0; s + 1 < rp.pos.size(); ++s) {
        if (rp.pos[s + 1] <= rp.rlo || rp.pos[s] >= rp.rhi) continue;
        const long long j0 = rp.pos[s] - kc, j1 = rp.pos[s + 1] - 1 - kc;
        const long long j = std::max(j0 < 0 ? -j0 : j0, j1 < 0 ? -j1 : j1);
        double a, hr;
        kb_curve_at(c, M, j, a, hr);
        if (rp.w[s] > RS_MAX_LIFT * hr) return 1e300;
        const double wa = rp.w[s] * a, cells = (double)(rp.pos[s + 1] - rp.pos[s]);
        e = std::max(e, wa);
        num += wa * wa * cells;
        den += rp.w2[s] * cells;
    }
    const double l2 = den > 0 ? sqrt(num / den) : 1e300;
    return std::max(e, l2);
}

inline void plan_shape_fast(HostPlan& hp);

// Build the plan of a frequency subset at length M = N / D (D == 1: the exact transform of that subset).  bands: for
// D > 1 the resample band [lo, hi) in data bins of every row of the subset.
inline std::shared_ptr<HostPlan> make_sub_plan(const HostPlan& hp, const std::vector<int>& fidx, int D,
                                               const std::vector<std::pair<long long, long long>>& bands) {
    std::shared_ptr<HostPlan> sp = std::make_shared<HostPlan>();
    HostPlan& s = *sp;
    s.device = hp.device; s.dtype = hp.dtype; s.family = hp.family; s.interpolate = hp.interpolate;
    s.N = hp.N / D;
    s.Nd = hp.N;
    s.F = (int)fidx.size();
    s.sfreq = hp.sfreq; s.p0 = hp.p0; s.p1 = hp.p1; s.p2 = hp.p2; s.prune_eps = hp.prune_eps;
    s.resample = 0;
    for (int i : fidx) {
        s.freqs.push_back(hp.freqs[(size_t)i]);
        if (!hp.aux.empty()) s.aux.push_back(hp.aux[(size_t)i]);
        if (!hp.table_lens.empty()) s.table_lens.push_back(hp.table_lens[(size_t)i]);
    }
    if (hp.family == FAM_TABLE) {
        s.table_len = hp.table_len;
        s.table.reserve(2 * fidx.size() * (size_t)hp.table_len);
        for (int i : fidx)
            s.table.insert(s.table.end(), hp.table.begin() + 2 * (size_t)i * (size_t)hp.table_len,
                           hp.table.begin() + 2 * ((size_t)i + 1) * (size_t)hp.table_len);
    }
    plan_geometry(s);
    plan_bands(s);
    if (D > 1) {
        s.band_bins = 0;
        for (size_t q = 0; q < s.rec.size(); ++q) {   // the resample band, centred on transform bin 0
            FreqRec& r = s.rec[q];
            r.lo = (int)std::max<long long>(r.lo, bands[q].first);
            r.hi = (int)std::max<long long>(r.lo, std::min<long long>(r.hi, bands[q].second));
            s.band_bins += r.hi - r.lo;
            const int kc = r.lo + (r.hi - r.lo) / 2;
            r.shift = kc;
            r.lo -= kc;
            r.hi -= kc;
        }
    }
    s.path = 1;
    s.lb = (ilog2_floor(s.N) + 2) / 2;
    s.generic_ok = 0;
    plan_shape_fast(s);
    if (!s.fast) return nullptr;
    plan_narrow(s);
    return sp;
}

// Weight table of a long-row plan of an analytic family (SpecParams::wtab): every frequency's band evaluated once, with
// the very formulas the kernels would use per bin (SpecEval<T>), times norm and the group's equaliser eq[|j|] (or null).
// Sets FreqRec::woff; false (and no table) when the bands hold more than max_bytes.
template <typename T>
inline bool build_weight_table(HostPlan& hp, const double* eq, size_t max_bytes, std::vector<T>& tab) {
    tab.clear();
    if (hp.family == FAM_TABLE || hp.F <= 0) return false;
    long long total = 0;
    for (FreqRec& r : hp.rec) { r.woff = total; total += std::max(0, r.hi - r.lo); }
    if (total <= 0 || (size_t)total * sizeof(T) > max_bytes) return false;
    tab.resize((size_t)total);
    SpecParams<T> sp;
    memset(&sp, 0, sizeof(sp));
    sp.family = hp.family; sp.grid_off = hp.grid_off; sp.df = hp.df; sp.p0 = hp.p0; sp.p1 = hp.p1;
    sp.p2 = hp.family == FAM_MORSE ? hp.p0 / hp.p1 : hp.p2;
    sp.norm = (T)(1.0 / (double)hp.data_len());
    for (const FreqRec& r : hp.rec)
        for (int j = r.lo; j < r.hi; ++j) {
            T w = SpecEval<T>::real(sp, r, j + r.shift);
            if (eq) w *= (T)eq[j < 0 ? -j : j];
            tab[(size_t)(r.woff + (j - r.lo))] = w;
        }
    return true;
}

// cost model, picoseconds per OUTPUT sample on B200 (profiles/r02): engine = the packed two-pass transform per point of
// its own length; interpolation = fir0 + fir_tap K with the vector kernel (even D), firs0 + firs_tap K with the scalar one
struct MrCost { double engine, fir0, fir_tap, firs0, firs_tap; };
inline MrCost mr_cost(int dtype) {
    MrCost c = dtype == 0 ? MrCost{8.0, 0.55, 0.035, 1.0, 0.18} : MrCost{20.0, 1.6, 0.3, 1.6, 0.3};
    // tuning overrides (thousandths of a picosecond): NWCWT_COST_ENGINE, NWCWT_COST_FIR0, NWCWT_COST_TAP
    if (env_int("NWCWT_COST_ENGINE", 0) > 0) c.engine = 0.001 * env_int("NWCWT_COST_ENGINE", 0);
    if (env_int("NWCWT_COST_FIR0", 0) > 0) c.fir0 = 0.001 * env_int("NWCWT_COST_FIR0", 0);
    if (env_int("NWCWT_COST_TAP", 0) > 0) c.fir_tap = 0.001 * env_int("NWCWT_COST_TAP", 0);
    return c;
}
// the interpolation kernels a decimation can use: vector kernel (nw_resample.cuh: resample_vec_body) for even D in fp32
inline bool resample_is_vec(int dtype, int D) { return dtype == 0 && (D & 1) == 0; }

inline std::shared_ptr<HostPlan> make_sub_plan_short(const HostPlan& hp, const std::vector<int>& fidx, int D,
                                                     const std::vector<std::pair<long long, long long>>& bands);

// short == false: long rows on the packed two-pass kernels (decimated rows go through the interpolation kernels of
// nw_resample.cuh).  short == true: rows that fit one CTA (nw_kernels4.cuh): any decimation whose length has a packed plan.
inline void plan_multirate(HostPlan& hp, bool shortrows) {
    hp.groups.clear();
    if (hp.F <= 0 || hp.resample == 0 || hp.Nd > 0) return;
    if (!shortrows && (hp.path != 1 || !hp.fast)) return;
    if (shortrows && (hp.path != 0 || !hp.short2 || hp.dtype != 0 || hp.family == FAM_TABLE)) return;
    if (env_int("NWCWT_NO_RESAMPLE", 0)) return;
    const long long N = hp.N;
    const double tol = hp.resample_tol > 0 ? hp.resample_tol : (hp.dtype == 0 ? 1e-6 : 5e-14);
    const double eps_rs = tol / 8;
    // tap counts the kernels are compiled for (even): 4 .. 12 (fp32 vector kernel), 4 .. 16 (fp32 scalar), 8 .. 24 (fp64)
    const int kmin = hp.dtype == 0 ? 4 : 8, kmax_vec = 12, kmax = hp.dtype == 0 ? 16 : 24;
    // shortest decimated length: below it the engine's tiles are too small to pay (tests lower it to reach the path on
    // lengths the host emulation can afford)
    long long MMIN = env_int("NWCWT_RESAMPLE_MMIN", 16384);
    if (MMIN < 64) MMIN = 64;
    if (shortrows) MMIN = 48;
    MrCost cm = mr_cost(hp.dtype);
    if (shortrows) cm = MrCost{3.0, 0.9, 0.04, 0.9, 0.04};   // measured on cfg3 (profiles/r02): exact row 3.9 ps per point
    // candidate decimations: divisors D of N with a fast plan at N / D
    std::vector<int> cands;
    for (int D = 2; D <= 64; ++D) {
        if (N % D || N / D < MMIN) continue;
        if (shortrows) {
            Fft2Plan t;
            const int pq = (D % 4 == 0 && N % 4 == 0) ? 4 : (D % 2 == 0 && N % 2 == 0) ? 2 : 1;
            if (D / pq <= 32 && plan_packed(N / D, t)) cands.push_back(D);
            continue;
        }
        HostPlan t;
        t.dtype = hp.dtype; t.N = N / D; t.F = 1;
        plan_shape_fast(t);
        if (t.fast) cands.push_back(D);
    }
    if (cands.empty()) return;
    std::vector<RowProfile> prof((size_t)hp.F);
    for (int i = 0; i < hp.F; ++i) row_profile(hp, i, eps_rs, prof[(size_t)i]);
    // alias curves per (D, K, beta), built on demand; beta = fb pi K on a grid of fb
    static const int NB = 9;
    auto fb_of = [](int b) { return 0.52 + 0.06 * b; };
    const double PI = 3.14159265358979323846;
    struct Curve { bool have = false; KbCurve A; };
    std::vector<Curve> curves(65 * 13 * NB);
    auto curve = [&](int D, int K, int b) -> const KbCurve& {
        Curve& c = curves[((size_t)D * 13 + (size_t)(K / 2)) * NB + (size_t)b];
        if (!c.have) { kb_alias_curve(D, K, fb_of(b) * PI * K, N / D, c.A); c.have = true; }
        return c.A;
    };
    // smallest even K (and its best beta) for which every listed row meets the tolerance at decimation D
    auto design = [&](int D, const std::vector<int>& rows, int& K, int& bsel, double& err) -> bool {
        const long long M = N / D;
        const int kcap = (shortrows || resample_is_vec(hp.dtype, D)) ? std::max(kmax_vec, kmin) : kmax;
        for (int k = kmin; k <= kcap; k += 2) {
            double best = 1e300;
            int bb = 0;
            for (int b = 0; b < NB; ++b) {
                const KbCurve& A = curve(D, k, b);
                double e = 0.0;
                for (int i : rows) {
                    e = std::max(e, row_error(prof[(size_t)i], A, M));
                    if (e > best) break;
                }
                if (e < best) { best = e; bb = b; }
            }
            if (best <= tol) { K = k; bsel = bb; err = best; return true; }
        }
        return false;
    };
    // per frequency: the cheapest decimation whose kernel meets the tolerance
    std::vector<int> pickD((size_t)hp.F, 1);
    for (int i = 0; i < hp.F; ++i) {
        const RowProfile& rp = prof[(size_t)i];
        const long long B = rp.rhi - rp.rlo;
        if (rp.pos.empty() || B <= 0) { pickD[(size_t)i] = cands.back(); continue; }   // empty band: zero row, cheapest group
        double best = cm.engine * 0.92;   // switch only for a clear gain
        for (int D : cands) {
            const long long M = N / D;
            if (B + 2 > M) continue;
            int K, b; double err;
            if (!design(D, std::vector<int>{i}, K, b, err)) continue;
            const bool vec = shortrows || resample_is_vec(hp.dtype, D);
            const double cost = cm.engine / D + (vec ? cm.fir0 + cm.fir_tap * K : cm.firs0 + cm.firs_tap * K);
            if (cost < best) { best = cost; pickD[(size_t)i] = D; }
        }
    }
    bool any = false;
    for (int d : pickD) any = any || d > 1;
    if (!any) return;
    // groups by decimation, exact rows (D = 1) last
    std::vector<int> Ds(cands);
    std::sort(Ds.begin(), Ds.end(), [](int a, int b) { return a > b; });
    Ds.push_back(1);
    for (size_t di = 0; di < Ds.size(); ++di) {
        const int D = Ds[di];
        MrGroup g;
        g.D = D;
        for (int i = 0; i < hp.F; ++i) if (pickD[(size_t)i] == D) g.fidx.push_back(i);
        if (g.fidx.empty()) continue;
        std::vector<std::pair<long long, long long>> bands;
        for (int i : g.fidx) bands.push_back(std::make_pair(prof[(size_t)i].rlo, prof[(size_t)i].rhi));
        g.sub = shortrows ? make_sub_plan_short(hp, g.fidx, D, bands) : make_sub_plan(hp, g.fidx, D, bands);
        bool ok = (bool)g.sub;
        if (ok && D > 1) {
            int bsel = 0;
            ok = design(D, g.fidx, g.K, bsel, g.err);
            if (ok) {
                g.beta = fb_of(bsel) * PI * g.K;
                const double inv = 1.0 / (bessel_i0(g.beta) - 1.0);
                g.coef.resize((size_t)D * g.K);
                g.t0.resize((size_t)D);
                for (int p = 0; p < D; ++p) {
                    const double x = (double)p / (double)D;
                    const int t0 = (int)floor(x - 0.5 * g.K) + 1;
                    g.t0[(size_t)p] = t0;
                    for (int t = 0; t < g.K; ++t) g.coef[(size_t)p * g.K + t] = kb_phi(x - (double)(t0 + t), g.K, g.beta, inv);
                }
                long long jmax = 1;
                for (const FreqRec& r : g.sub->rec) jmax = std::max<long long>(jmax, std::max<long long>(-(long long)r.lo, (long long)r.hi));
                std::vector<long long> bins((size_t)jmax + 1);
                for (long long j = 0; j <= jmax; ++j) bins[(size_t)j] = j;
                std::vector<double> H;
                kb_response(D, g.K, g.beta, N / D, bins, H);
                g.eq.resize(H.size());
                for (size_t j = 0; j < H.size(); ++j) g.eq[j] = (double)D / H[j];
            }
        }
        if (!ok) {   // fall back to the exact transform for these rows: merge into the D = 1 group (built last)
            if (D == 1) { hp.groups.clear(); return; }
            for (int i : g.fidx) pickD[(size_t)i] = 1;
            continue;
        }
        hp.groups.push_back(std::move(g));
    }
    if (hp.groups.size() == 1 && hp.groups[0].D == 1) hp.groups.clear();   // nothing resampled after all
}

// Short rows: the plan of a frequency subset at length M = N / D (D == 1: exact rows): bands (centred on bin 0 for D > 1)
// and the packed M-point plan.
inline std::shared_ptr<HostPlan> make_sub_plan_short(const HostPlan& hp, const std::vector<int>& fidx, int D,
                                                     const std::vector<std::pair<long long, long long>>& bands) {
    std::shared_ptr<HostPlan> sp = std::make_shared<HostPlan>();
    HostPlan& s = *sp;
    s.device = hp.device; s.dtype = hp.dtype; s.family = hp.family; s.interpolate = hp.interpolate;
    s.N = hp.N / D;
    s.Nd = hp.N;
    s.F = (int)fidx.size();
    s.sfreq = hp.sfreq; s.p0 = hp.p0; s.p1 = hp.p1; s.p2 = hp.p2; s.prune_eps = hp.prune_eps;
    s.resample = 0;
    for (int i : fidx) {
        s.freqs.push_back(hp.freqs[(size_t)i]);
        if (!hp.aux.empty()) s.aux.push_back(hp.aux[(size_t)i]);
    }
    plan_geometry(s);
    plan_bands(s);
    if (D > 1) {
        s.band_bins = 0;
        for (size_t q = 0; q < s.rec.size(); ++q) {   // the resample band, centred on transform bin 0
            FreqRec& r = s.rec[q];
            r.lo = (int)std::max<long long>(r.lo, bands[q].first);
            r.hi = (int)std::max<long long>(r.lo, std::min<long long>(r.hi, bands[q].second));
            s.band_bins += r.hi - r.lo;
            const int kc = r.lo + (r.hi - r.lo) / 2;
            r.shift = kc;
            r.lo -= kc;
            r.hi -= kc;
        }
    }
    s.path = 0;
    if (!plan_packed(s.N, s.stS)) return nullptr;
    return sp;
}

// Shape of the resampled short-row kernel (nw_kernels4.cuh) once the groups exist.
inline void plan_shape_short3(HostPlan& hp) {
    hp.short3 = 0;
    if (hp.groups.empty()) return;
    for (const MrGroup& g : hp.groups) {
        if (!g.sub) { hp.groups.clear(); return; }
        if (g.D > 1 && (g.K & 1 || g.K < 4 || g.K > 12)) { hp.groups.clear(); return; }
        for (int t : g.t0) if (t != 1 - g.K / 2) { hp.groups.clear(); return; }
        const int pq = (g.D % 4 == 0 && hp.N % 4 == 0) ? 4 : (g.D % 2 == 0 && hp.N % 2 == 0) ? 2 : 1;
        if (g.D > 1 && g.D / pq > 32) { hp.groups.clear(); return; }   // lanes per output sample m
    }
    const size_t bytes = 2 * (size_t)hp.N * 2 * cx_size(hp.dtype) + 512 + 64 * 12 * (cx_size(hp.dtype) / 2);   // short3_smem_bytes
    if (bytes > SMEM_MAX) { hp.groups.clear(); return; }
    hp.short3 = 1;
    hp.smem_S3 = bytes;
    int v = (env_int("NWCWT_NTHR_S3", 256) + 31) / 32 * 32;
    hp.nthrS3 = v < 64 ? 64 : v > 512 ? 512 : v;
}

}  // namespace nw
