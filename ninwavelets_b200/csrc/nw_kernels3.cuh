// Fused short-row kernel on the packed in-place engine (nw_fft2.cuh): rows that fit one CTA.
//
// One CTA owns a PAIR of signals (the two lanes of every packed value) and a share of the frequencies:
//   1. forward transform of both signals at once (scipy.fftpack.fft, base.py:399): decimation in time straight
//      from global memory, natural-order spectrum pair Xs[k] kept in shared memory for the whole CTA;
//   2. per group of NF = 2^tpsh frequencies: zero the tile, gather  W_f(k) * Xs[k]  for the in-band bins of each
//      frequency into the transform's slots (one spectrum evaluation serves both signals; base.py:236-248, 404),
//      N-point in-place inverse transform of the NF interleaved lane-pair sequences (base.py:406);
//   3. the last pass leaves z (cwt) or |z|^2 / |z| (base.py:425, 443) IN PLACE in natural order; after a barrier
//      each of the 2 NF rows gets its Baseline statistics from shared memory (one warp per row, fp64
//      accumulation; base.py:46-68) and is written with coalesced, vectorised stores.
// Shared memory: N (1 + NF) two-lane complex values (fp32, N = 1500, NF = 2: 72 KB -> three CTAs per SM).
#pragma once
#include "nw_common.h"
#include "nw_fft2.cuh"
#include "nw_family.cuh"
#include "nw_kernels.cuh"
#include "nw_kernels2.cuh"

namespace nw {

template <typename T>
struct Short2Params {
    const T* signals;   // [S][N] real
    void* out;          // [S][F][N] T (abs/power) or cx<T> (cwt)
    int N, F, S;
    int tpsh;           // NF = 1 << tpsh frequencies per pass
    int out_mode, bl_mode, bl_lo, bl_hi;
    int fsplit;         // CTAs per signal pair (frequency groups are dealt round-robin)
    Fft2Plan st;
    const cx<T>* tw;    // [N]
    SpecParams<T> sp;
    // epoch reductions (short2_epochs_body): signals are [n_channels][n_epochs][N]; acc is the complex accumulator of
    // the inter-trial coherence, [n_channels][F][N]
    int n_epochs;
    cx<T>* acc;
};

template <typename T> NW_HD size_t short2_smem_bytes(int N, int tpsh) {
    return (size_t)N * (1 + ((size_t)1 << tpsh)) * sizeof(cx2<T>) + 160 * sizeof(double);
}

// forward: two real signals as the two lanes
template <typename T> struct RealPairSrc {
    const T* x0;
    const T* x1;   // == x0 when the pair has one signal
    template <int R> NW_HD void load_all(int base, int step, int, cx2<T>* v) const {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int n = base + q * step;
            v[q] = mk2<T>(pk_make(x0[n], x1[n]), pk_bcast((T)0));
        }
    }
};
template <typename T> struct SpectrumPairDst {
    cx2<T>* Xs;
    struct Ctx { int base, step; };
    NW_HD Ctx begin(int base, int step, int) const { return Ctx{base, step}; }
    template <int R> NW_HD void store_all(const Ctx& x, const cx2<T>* v) const {
#pragma unroll
        for (int q = 0; q < R; ++q) Xs[x.base + q * x.step] = v[q];
    }
};

// inverse, last pass: result n = base + q * step of sequence t stays in its slot buf[n * NF + t]
template <typename T, int MODE> struct InPlaceOutDst {
    cx2<T>* buf;
    int tpsh;
    struct Ctx { cx2<T>* p; int stride; };
    NW_HD Ctx begin(int base, int step, int t) const { return Ctx{buf + ((size_t)base << tpsh) + t, step << tpsh}; }
    template <int R> NW_HD void store_all(const Ctx& x, const cx2<T>* v) const {
#pragma unroll
        for (int q = 0; q < R; ++q) {
            if (MODE == OUT_CWT) {
                x.p[(size_t)q * x.stride] = v[q];
            } else {
                pk<T> p = pk_fma(v[q].im, v[q].im, v[q].re * v[q].re);
                if (MODE == OUT_ABS) p = pk_make(nw_sqrt(pk_lo(p)), nw_sqrt(pk_hi(p)));
                x.p[(size_t)q * x.stride].re = p;
            }
        }
    }
};

template <typename T, int MODE, int SP>
NW_HD void short2_body(const Short2Params<T>& P, char* smem, int bx, int tid, int nthr) {
    const int N = P.N, tpsh = P.tpsh, NF = 1 << tpsh;
    cx2<T>* Xs = (cx2<T>*)smem;
    cx2<T>* buf = Xs + N;
    double* rstat = (double*)(buf + ((size_t)N << tpsh));   // [2 NF][2]
    const int pair = bx / P.fsplit, part = bx - pair * P.fsplit;
    const int s0 = 2 * pair;
    const bool has1 = s0 + 1 < P.S;

    // ---- forward transform of the signal pair ------------------------------------------------------
    {
        RealPairSrc<T> src{P.signals + (size_t)s0 * N, P.signals + (size_t)(has1 ? s0 + 1 : s0) * N};
        SpectrumPairDst<T> dst{Xs};
        fft2_dit<T, -1>(P.st, 0, P.tw, buf, src, dst, tid, nthr);
    }
    NW_SYNC();

    typedef StaticPlan<SP> SPL;
    const int ngroups = (P.F + NF - 1) >> tpsh;
    const size_t esz = (MODE == OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
    // the tile starts as zeros; every pass re-zeroes the slots it reads on the way out
    const cx2<T> z = zero2<T>();
    for (int i = tid; i < (N << tpsh); i += nthr) buf[i] = z;
    NW_SYNC();
    for (int g = part; g < ngroups; g += P.fsplit) {
        const int f0 = g << tpsh;
        const int nvalid = (P.F - f0 < NF) ? (P.F - f0) : NF;
        // ---- the in-band products into the zero tile ------------------------------------------------
        for (int t = 0; t < nvalid; ++t) {
            const int fi = f0 + t;
            const FreqRec rec = P.sp.rec[fi];
            for (int k = rec.lo + tid; k < rec.hi; k += nthr) {
                const cx2<T> x = Xs[k];
                cx2<T> y;
                if (P.sp.family == FAM_TABLE) {
                    const cx<T> w = scale(P.sp.table[(long long)fi * P.sp.table_len + (k - rec.toff)], P.sp.norm);
                    y = cmul_s(x, w);
                } else {
                    const T w = SpecEval<T>::real(P.sp, rec, k);
                    y = mk2<T>(x.re * w, x.im * w);
                }
                buf[((size_t)fft2_dit_pos(P.st, k) << tpsh) + t] = y;
            }
        }
        NW_SYNC();
        // ---- inverse transform, result left in place ------------------------------------------------
        InPlaceOutDst<T, MODE> dst{buf, tpsh};
        if constexpr (SP == 0) fft2_dit<T, +1>(P.st, tpsh, P.tw, buf, FromBuf(), dst, tid, nthr);
        else fft2_dit_static<T, +1, SPL::TPS, (SP ? SPL::P : 4), (SP ? SPL::R0 : 2), (SP ? SPL::R1 : 2), SPL::R2>(P.tw, buf, dst, tid, nthr);
        NW_SYNC();
        // ---- rows out: row r = 2 t + lane  ->  out[s0 + lane][f0 + t][:] ------------------------------
        // Baseline as y = (x + b) * a [, log10(y) * c]: b, a, c per row from the window statistics
        //   mean: x - m   ratio: x / m   percent: (x - m) / m   zscore: (x - m) / sd   log: log10(x / m)   zlog: log10(x / m) / sd
        const int nrows = 2 * nvalid;
        const bool blon = MODE != OUT_CWT && P.bl_mode != BL_NONE;
        const bool uselog = P.bl_mode == BL_LOG || P.bl_mode == BL_ZLOG;
        if (blon) {
            // window statistics of every row (np.mean / np.std, ddof = 0; base.py:49-50, 65): sums of d = x - x[lo]
            // and d^2 in fp64, the window of a row split over the warps the CTA has per row
            const int cnt = P.bl_hi - P.bl_lo;
#if defined(__CUDA_ARCH__)
            const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
            const int wpr = nwarp / nrows > 0 ? (nwarp / nrows > 4 ? 4 : nwarp / nrows) : 1;   // warps per row
            for (int job = warp; job < nrows * wpr; job += nwarp) {
                const int r = job / wpr, part = job - r * wpr;
                const int t = r >> 1, l = r & 1;
                double s = 0.0, q = 0.0;
                if (cnt > 0) {
                    const pk<T> p0 = buf[((size_t)P.bl_lo << tpsh) + t].re;
                    const double x0 = (double)(l ? pk_hi(p0) : pk_lo(p0));
                    for (int i = P.bl_lo + part * 32 + lane; i < P.bl_hi; i += 32 * wpr) {
                        const pk<T> p = buf[((size_t)i << tpsh) + t].re;
                        const double d = (double)(l ? pk_hi(p) : pk_lo(p)) - x0;
                        s += d;
                        q += d * d;
                    }
                }
                for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
                if (lane == 0) { rstat[32 + 2 * (r * 4 + part)] = s; rstat[32 + 2 * (r * 4 + part) + 1] = q; }
            }
            NW_SYNC();
            if (tid < nrows) {
                const int r = tid, t = r >> 1, l = r & 1;
                double s = 0.0, q = 0.0;
                for (int part = 0; part < wpr; ++part) { s += rstat[32 + 2 * (r * 4 + part)]; q += rstat[32 + 2 * (r * 4 + part) + 1]; }
                double m = nan(""), sd = nan("");
                if (cnt > 0) {
                    const pk<T> p0 = buf[((size_t)P.bl_lo << tpsh) + t].re;
                    const double x0 = (double)(l ? pk_hi(p0) : pk_lo(p0));
                    const double ms = s / cnt;
                    double var = q / cnt - ms * ms;
                    if (var < 0.0) var = 0.0;
                    m = x0 + ms;
                    sd = sqrt(var);
                }
                rstat[2 * r] = m;
                rstat[2 * r + 1] = sd;
            }
#else
            for (int r = tid; r < nrows; r += nthr) {
                const int t = r >> 1, l = r & 1;
                double s = 0.0, q = 0.0, x0 = 0.0;
                if (cnt > 0) { const pk<T> p0 = buf[((size_t)P.bl_lo << tpsh) + t].re; x0 = (double)(l ? pk_hi(p0) : pk_lo(p0)); }
                for (int i = P.bl_lo; i < P.bl_hi; ++i) {
                    const pk<T> p = buf[((size_t)i << tpsh) + t].re;
                    const double d = (double)(l ? pk_hi(p) : pk_lo(p)) - x0;
                    s += d;
                    q += d * d;
                }
                const double ms = cnt > 0 ? s / cnt : 0.0;
                double var = cnt > 0 ? q / cnt - ms * ms : 0.0;
                if (var < 0.0) var = 0.0;
                rstat[2 * r] = cnt > 0 ? x0 + ms : nan("");
                rstat[2 * r + 1] = cnt > 0 ? sqrt(var) : nan("");
            }
#endif
            NW_SYNC();
        }
        // one thread per sample n, all rows of the group: lanes read consecutive slots (a 2-way conflict at
        // most) and every store instruction of a warp covers 32 consecutive samples of one row
#pragma unroll 1
        for (int t = 0; t < nvalid; ++t) {
            T b0 = (T)0, a0 = (T)1, c0 = (T)1, b1 = (T)0, a1 = (T)1, c1 = (T)1;
            if (blon) {
                const T m0 = (T)rstat[4 * t], sd0 = (T)rstat[4 * t + 1], m1 = (T)rstat[4 * t + 2], sd1 = (T)rstat[4 * t + 3];
                switch (P.bl_mode) {
                    case BL_MEAN: b0 = -m0; b1 = -m1; break;
                    case BL_RATIO: a0 = (T)1 / m0; a1 = (T)1 / m1; break;
                    case BL_PERCENT: b0 = -m0; b1 = -m1; a0 = (T)1 / m0; a1 = (T)1 / m1; break;
                    case BL_ZSCORE: b0 = -m0; b1 = -m1; a0 = (T)1 / sd0; a1 = (T)1 / sd1; break;
                    case BL_LOG: a0 = (T)1 / m0; a1 = (T)1 / m1; break;
                    case BL_ZLOG: a0 = (T)1 / m0; a1 = (T)1 / m1; c0 = (T)1 / sd0; c1 = (T)1 / sd1; break;
                    default: break;
                }
            }
            char* row0 = (char*)P.out + ((size_t)s0 * P.F + (f0 + t)) * (size_t)N * esz;
            char* row1 = row0 + (size_t)P.F * (size_t)N * esz;
            for (int n = tid; n < N; n += nthr) {
                const cx2<T> v = buf[((size_t)n << tpsh) + t];
                buf[((size_t)n << tpsh) + t] = z;   // the next pass gathers into a zero tile
                if (MODE == OUT_CWT) {
                    st_stream((cx<T>*)row0 + n, lane0(v));
                    if (has1) st_stream((cx<T>*)row1 + n, lane1(v));
                } else {
                    T y0 = pk_lo(v.re), y1 = pk_hi(v.re);
                    if (blon) {
                        y0 = (y0 + b0) * a0;
                        y1 = (y1 + b1) * a1;
                        if (uselog) { y0 = nw_log10(y0) * c0; y1 = nw_log10(y1) * c1; }
                    }
                    st_stream((T*)row0 + n, y0);
                    if (has1) st_stream((T*)row1 + n, y1);
                }
            }
        }
        NW_SYNC();
    }
}

// ---- epoch reductions fused into the transform (mneutils.py:53-55 mean power over epochs, :68-71 inter-trial coherence) --
// One CTA owns a CHANNEL and a share of the frequency groups and loops over the channel's epochs two at a time (the two
// lanes): forward transform of the epoch pair, then for each of its frequency groups the inverse transform as in
// short2_body, whose rows are not written out but accumulated into the channel's (F, N) result - every (frequency, sample)
// of it is owned by exactly one thread of one CTA, so the accumulation is a plain read-add-write of L2-resident lines and
// the (E, F, N) array of per-epoch rows is never materialised.
//   KIND 0: out[c][f][n] = mean_e |z_e|^2        KIND 1: out[c][f][n] = | mean_e z_e / |z_e| |  (complex sums in P.acc)
template <typename T, int KIND, int SP>
NW_HD void short2_epochs_body(const Short2Params<T>& P, char* smem, int bx, int tid, int nthr) {
    const int N = P.N, tpsh = P.tpsh, NF = 1 << tpsh, E = P.n_epochs;
    cx2<T>* Xs = (cx2<T>*)smem;
    cx2<T>* buf = Xs + N;
    const int ch = bx / P.fsplit, part = bx - ch * P.fsplit;
    typedef StaticPlan<SP> SPL;
    const int ngroups = (P.F + NF - 1) >> tpsh;
    const cx2<T> z = zero2<T>();
    const T inv_e = (T)1 / (T)E;
    constexpr int MODE = KIND == 0 ? OUT_POWER : OUT_CWT;
    for (int i = tid; i < (N << tpsh); i += nthr) buf[i] = z;
    for (int e0 = 0; e0 < E; e0 += 2) {
        const bool has1 = e0 + 1 < E, first = e0 == 0, last = e0 + 2 >= E;
        NW_SYNC();   // the previous pair's rows are consumed (and the tile is zero) before Xs is overwritten
        {
            const T* x0 = P.signals + ((size_t)ch * E + e0) * (size_t)N;
            RealPairSrc<T> src{x0, has1 ? x0 + N : x0};
            SpectrumPairDst<T> dst{Xs};
            fft2_dit<T, -1>(P.st, 0, P.tw, buf, src, dst, tid, nthr);   // first pass reads global memory, writes buf slots it later re-reads
        }
        NW_SYNC();
        // the forward transform used the tile as its work space: zero it again for the gathers
        for (int i = tid; i < (N << tpsh); i += nthr) buf[i] = z;
        NW_SYNC();
        for (int g = part; g < ngroups; g += P.fsplit) {
            const int f0 = g << tpsh;
            const int nvalid = (P.F - f0 < NF) ? (P.F - f0) : NF;
            for (int t = 0; t < nvalid; ++t) {
                const int fi = f0 + t;
                const FreqRec rec = P.sp.rec[fi];
                for (int k = rec.lo + tid; k < rec.hi; k += nthr) {
                    const cx2<T> x = Xs[k];
                    cx2<T> y;
                    if (P.sp.family == FAM_TABLE) {
                        const cx<T> w = scale(P.sp.table[(long long)fi * P.sp.table_len + (k - rec.toff)], P.sp.norm);
                        y = cmul_s(x, w);
                    } else {
                        const T w = SpecEval<T>::real(P.sp, rec, k);
                        y = mk2<T>(x.re * w, x.im * w);
                    }
                    buf[((size_t)fft2_dit_pos(P.st, k) << tpsh) + t] = y;
                }
            }
            NW_SYNC();
            InPlaceOutDst<T, MODE> dst{buf, tpsh};
            if constexpr (SP == 0) fft2_dit<T, +1>(P.st, tpsh, P.tw, buf, FromBuf(), dst, tid, nthr);
            else fft2_dit_static<T, +1, SPL::TPS, (SP ? SPL::P : 4), (SP ? SPL::R0 : 2), (SP ? SPL::R1 : 2), SPL::R2>(P.tw, buf, dst, tid, nthr);
            NW_SYNC();
#pragma unroll 1
            for (int t = 0; t < nvalid; ++t) {
                const size_t row = ((size_t)ch * P.F + (f0 + t)) * (size_t)N;
                for (int n = tid; n < N; n += nthr) {
                    const cx2<T> v = buf[((size_t)n << tpsh) + t];
                    buf[((size_t)n << tpsh) + t] = z;   // the next pass gathers into a zero tile
                    if (KIND == 0) {
                        T* o = (T*)P.out + row + n;
                        T a = pk_lo(v.re) + (has1 ? pk_hi(v.re) : (T)0);
                        if (!first) a += *o;
                        *o = last ? a * inv_e : a;
                    } else {
                        const cx<T> u0 = lane0(v), u1 = lane1(v);
                        const T r0 = (T)1 / nw_hypot(u0.x, u0.y), r1 = has1 ? (T)1 / nw_hypot(u1.x, u1.y) : (T)0;
                        cx<T> a = mk<T>(u0.x * r0 + (has1 ? u1.x * r1 : (T)0), u0.y * r0 + (has1 ? u1.y * r1 : (T)0));
                        cx<T>* ac = P.acc + row + n;
                        if (!first) a = a + *ac;
                        if (last) ((T*)P.out)[row + n] = nw_hypot(a.x * inv_e, a.y * inv_e);
                        else *ac = a;
                    }
                }
            }
            NW_SYNC();
        }
    }
}

}  // namespace nw
