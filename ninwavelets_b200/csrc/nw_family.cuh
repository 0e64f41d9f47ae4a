// Analytic wavelet spectra evaluated in registers, one frequency bin at a time.
//
// Follows the reference's numpy formulas:
//   Morse    wavelets.py:65-74    W = 2 H(x) x^b exp((b/r)(1 - x^r)),  x = f_k / freq
//   Morlet   wavelets.py:132-136  W = c pi^-1/4 (exp(-(s-x)^2/2) - k exp(-x^2/2)),
//                                 x = f_k / freq * peak_freq(freq)   (wavelets.py:143-144)
//   Shannon  wavelets.py:256-262  W = 1 where f_k <= 1.0 Hz (a pure band, no arithmetic)
//   Table    base.py:250-255 (MexicanHat / Haar / user formulas): complex table per
//            frequency, placed on the bin axis by pad_to (base.py:75-82).
// f_k = k * (1/L) is the DFT grid of base.py:173-194.  Bins outside the plan's
// per-frequency band [lo, hi) are zero (exactly zero in the reference, or below
// prune_eps * peak) and are never evaluated.
//
// fp64 mode keeps numpy's operation order (divide by freq, pow, exp).  fp32
// mode forms x in double (k can exceed 2^24) and evaluates Morse in the log
// domain, 2 exp(b ln x + (b/r)(1 - x^r)), which cannot overflow.
#pragma once
#include "nw_common.h"

namespace nw {

// Per-frequency record prepared by the host planner (all doubles evaluated with
// the same operation order as the reference).
struct FreqRec {
    int lo, hi;     // band of non-zero bins, in (signed) transform bins: data bin = transform bin + shift
    int toff;       // TABLE: data bin of table element 0 (pad_to front offset of this row)
    int shift;      // 0 except in the plan of a resampled group (nw_resample.cuh), whose M-point spectrum holds the
                    // band around bin 0 (negative bins wrap to the top end)
    double freq;    // analysis frequency (Hz)
    double aux;     // Morlet: peak_freq(freq)
    double kx;      // fp32 path: x = (k - grid_off) * kx   (df/freq [* peak])
    long long woff; // long rows: index of the band's first bin in the plan's weight table (SpecParams::wtab)
};

template <typename T>
struct SpecParams {
    int family;
    int grid_off;        // bin k maps to grid index k - grid_off (centre padding of pad_to)
    double df;           // grid step 1/L
    double p0, p1, p2;   // Morse: b, r, b/r      Morlet: sigma, c*pi^-1/4, k
    T norm;              // 1/N of the inverse transform, folded into the spectrum
    const FreqRec* rec;  // [F] device
    const cx<T>* table;  // TABLE: [F][table_len]
    long long table_len; // row pitch of the table
    // long rows, analytic families: the band of every frequency evaluated ONCE per plan (all signals share it) -
    // wtab[rec.woff + (j - rec.lo)] = W_f at transform bin j, times norm and, in a resampled group, the equaliser.
    // nullptr: evaluate in registers (bands too large to tabulate).
    const T* wtab;
};

NW_HD float nw_exp(float x) { return expf(x); }
NW_HD double nw_exp(double x) { return exp(x); }
NW_HD float nw_log(float x) { return logf(x); }
NW_HD double nw_log(double x) { return log(x); }

// Real-valued analytic spectrum at bin k (caller guarantees lo <= k < hi).
template <typename T> struct SpecEval;

template <> struct SpecEval<double> {
    static NW_HD double real(const SpecParams<double>& sp, const FreqRec& r, int k) {
        const double g = (double)(k - sp.grid_off) * sp.df;
        if (sp.family == FAM_MORSE) {
            const double x = g / r.freq;
            if (!(x > 0.0)) return 0.0;  // heaviside(x, x) at x == 0
            return sp.norm * (2.0 * (pow(x, sp.p0) * exp(sp.p2 * (1.0 - pow(x, sp.p1)))));
        } else if (sp.family == FAM_MORLET) {
            const double x = g / r.freq * r.aux;
            const double d = sp.p0 - x;
            return sp.norm * (sp.p1 * (exp(-(d * d) / 2) - sp.p2 * exp(-(x * x) / 2)));
        }
        return sp.norm;  // Shannon: inside the band the spectrum is 1
    }
};

template <> struct SpecEval<float> {
    static NW_HD float real(const SpecParams<float>& sp, const FreqRec& r, int k) {
        const double xd = (double)(k - sp.grid_off) * r.kx;
        if (sp.family == FAM_MORSE) {
            const float x = (float)xd;
            if (!(x > 0.0f)) return 0.0f;
            const float lx = logf(x);
            const float e = (float)sp.p0 * lx + (float)sp.p2 * (1.0f - expf((float)sp.p1 * lx));
            return sp.norm * 2.0f * expf(e);
        } else if (sp.family == FAM_MORLET) {
            const float d = (float)(sp.p0 - xd);
            const float x = (float)xd;
            return sp.norm * ((float)sp.p1 * (expf(-0.5f * d * d) - (float)sp.p2 * expf(-0.5f * x * x)));
        }
        return sp.norm;
    }
};

// Spectrum times signal spectrum at bin k for frequency record r (index fi).
template <typename T>
NW_HD cx<T> spec_times(const SpecParams<T>& sp, const FreqRec& r, int fi, int k, cx<T> xk) {
    if (sp.family == FAM_TABLE) {
        const cx<T> w = sp.table[(long long)fi * sp.table_len + (k - r.toff)];
        return scale(cmul(w, xk), sp.norm);
    }
    const T w = SpecEval<T>::real(sp, r, k);
    return mk<T>(w * xk.x, w * xk.y);
}

}  // namespace nw
