// Fused short-row kernel with resampled rows (r02): rows that fit one CTA, band-limited frequencies computed at a
// decimated length and interpolated (DESIGN.md "resampled rows"; nw_plan.h: plan_multirate_short).
//
// One CTA owns a PAIR of signals (the two lanes of every packed value) and a share of the work units:
//   1. forward transform of both signals (scipy.fftpack.fft, base.py:399) as in nw_kernels3.cuh: the natural-order
//      spectrum pair Xs[k] stays in shared memory for the whole CTA;
//   2. a unit = NF = 2^tpsh frequencies of one resample group (decimation D, M = N / D): zero the M-point tile, gather
//      W_f(k) Xs[k] (weights from the group's table: 1/N, equaliser folded in; base.py:236-248, 404) of each band, moved
//      to bin 0, into the decimation-in-time slots; in-place M-point inverse transform of the NF interleaved lane-pair
//      sequences (base.py:406) - y[m] = z(m D) up to a unit-modulus factor;
//   3. interpolation: |z(m D + p)|^2 = |sum_t coef[p][t] y[m + 1 - K/2 + t]|^2 for both signals at once (the signals are
//      the lanes: K packed FMAs for the real parts, K for the imaginary parts, the weight a broadcast scalar) into the
//      unit's output tile tile[t][m DP + p] (DP = D | 1: the stores of consecutive m are bank-conflict free);
//      D == 1 (rows the planner keeps exact): the tile is |y|^2 itself;
//   4. Baseline statistics per row from the tile (fp64, base.py:46-68) and the thread-per-sample epilogue of
//      nw_kernels3.cuh: every store of a warp covers 32 consecutive samples of one row.
// Shared memory: Xs N + tile max(N, Mmax NF) two-lane complex values + NF (N + Mmax) two-lane reals.
#pragma once
#include "nw_common.h"
#include "nw_fft2.cuh"
#include "nw_family.cuh"
#include "nw_kernels.cuh"
#include "nw_kernels2.cuh"
#include "nw_kernels3.cuh"

namespace nw {

template <typename T>
struct Short3Group {
    int M, D, K, F;          // decimated length, decimation (1: exact rows), taps, frequencies of the group
    int DP;                  // tile pitch per m: D | 1
    int unit0, nunits;       // this group's units (NF frequencies each): [unit0, unit0 + nunits)
    int PCH;                 // phase chunks of 4 per m
    Fft2Plan st;             // M-point plan
    fastdiv dD, dM, dPer;    // n / D, x / M, x / (M PCH)
    const cx<T>* tw;         // [M]
    const FreqRec* rec;      // [F] bands centred on transform bin 0 (FreqRec::shift), woff into wtab
    const T* wtab;           // weights: W_f / N (x equaliser)
    const T* coef;           // [D][K]
    const int* fmap;         // [F] -> plan frequency index
};

template <typename T>
struct Short3Params {
    const T* signals;   // [S][N] real
    void* out;          // [S][F_out][N] T
    int N, F_out, S;
    int tpsh;           // NF = 1 << tpsh frequencies per unit
    int out_mode, bl_mode, bl_lo, bl_hi;
    int fsplit;         // CTAs per signal pair (units are dealt round-robin)
    Fft2Plan st;        // N-point plan (forward transform)
    const cx<T>* tw;    // [N]
    const Short3Group<T>* groups;
    int ngroups, nunits;
    int yslots;         // two-lane complex slots of the transform tile: max(N, Mmax << tpsh)
    int tpitch;         // two-lane reals per frequency of the output tile
};

template <typename T> NW_HD size_t short3_smem_bytes(int N, int yslots, int tpitch, int tpsh) {
    return ((size_t)N + (size_t)yslots) * sizeof(cx2<T>) + ((size_t)tpitch << tpsh) * sizeof(pk<T>) + 160 * sizeof(double);
}

// interpolation of the unit's rows: tasks (frequency t, phase chunk pc of 4, sample m), m fastest
template <typename T, int K, int MODE>
NW_HD void short3_interp(const Short3Group<T>& g, const cx2<T>* ybuf, pk<T>* tile, int tpsh, int nvalid, int tpitch, int tid,
                         int nthr) {
    const int M = g.M, D = g.D;
    const uint32_t per = (uint32_t)M * (uint32_t)g.PCH, total = per * (uint32_t)nvalid;
    for (uint32_t i = tid; i < total; i += nthr) {
        const uint32_t t = fd_div(i, g.dPer), r = i - t * per;
        const uint32_t pc = fd_div(r, g.dM);
        const int m = (int)(r - pc * (uint32_t)M);
        cx2<T> w[K];
        {
            int mi = m + 1 - K / 2;
            if (mi < 0) mi += M;
#pragma unroll
            for (int o = 0; o < K; ++o) {
                w[o] = ybuf[((size_t)mi << tpsh) + t];
                if (++mi == M) mi = 0;
            }
        }
        const int p0 = (int)pc * 4, p1 = p0 + 4 < D ? p0 + 4 : D;
        pk<T>* dst = tile + (size_t)t * tpitch + (size_t)m * g.DP;
#pragma unroll 1
        for (int p = p0; p < p1; ++p) {
            const T* c = g.coef + (size_t)p * K;
            pk<T> are = pk_bcast((T)0), aim = are;
#pragma unroll
            for (int o = 0; o < K; ++o) {
                const T cv = c[o];
                are = pk_fma(w[o].re, cv, are);
                aim = pk_fma(w[o].im, cv, aim);
            }
            pk<T> v = pk_fma(aim, aim, are * are);
            if (MODE == OUT_ABS) v = pk_make(nw_sqrt(pk_lo(v)), nw_sqrt(pk_hi(v)));
            dst[p] = v;
        }
    }
}

template <typename T, int MODE>
NW_HD void short3_body(const Short3Params<T>& P, char* smem, int bx, int tid, int nthr) {
    const int N = P.N, tpsh = P.tpsh, NF = 1 << tpsh;
    cx2<T>* Xs = (cx2<T>*)smem;
    cx2<T>* buf = Xs + N;
    pk<T>* tile = (pk<T>*)(buf + P.yslots);
    double* rstat = (double*)(tile + ((size_t)P.tpitch << tpsh));   // [2 NF][2] + partial sums
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

    const cx2<T> z = zero2<T>();
    int gi = 0;
    for (int u = part; u < P.nunits; u += P.fsplit) {
        while (u >= P.groups[gi].unit0 + P.groups[gi].nunits) ++gi;
        const Short3Group<T>& g = P.groups[gi];
        const int M = g.M, D = g.D;
        const int f0 = (u - g.unit0) << tpsh;
        const int nvalid = (g.F - f0 < NF) ? (g.F - f0) : NF;
        // ---- the in-band products, band centre at bin 0, into the zero tile -------------------------
        for (int i = tid; i < (M << tpsh); i += nthr) buf[i] = z;
        NW_SYNC();
        for (int t = 0; t < nvalid; ++t) {
            const FreqRec rec = g.rec[f0 + t];
            const T* wt = g.wtab + rec.woff - rec.lo;
            for (int j = rec.lo + tid; j < rec.hi; j += nthr) {
                const cx2<T> x = Xs[j + rec.shift];
                const T w = wt[j];
                const int jm = j < 0 ? j + M : j;
                buf[((size_t)fft2_dit_pos(g.st, jm) << tpsh) + t] = mk2<T>(x.re * w, x.im * w);
            }
        }
        NW_SYNC();
        // ---- M-point inverse transform, complex result left in place ---------------------------------
        {
            InPlaceOutDst<T, OUT_CWT> dst{buf, tpsh};
            fft2_dit<T, +1>(g.st, tpsh, g.tw, buf, FromBuf(), dst, tid, nthr);
        }
        NW_SYNC();
        // ---- interpolation to the N samples of every row: tile[t][m DP + p] --------------------------
        if (D == 1) {
            for (int i = tid; i < (N << tpsh); i += nthr) {
                const int t = i & (NF - 1), n = i >> tpsh;
                if (t < nvalid) {
                    const cx2<T> v = buf[i];
                    pk<T> p = pk_fma(v.im, v.im, v.re * v.re);
                    if (MODE == OUT_ABS) p = pk_make(nw_sqrt(pk_lo(p)), nw_sqrt(pk_hi(p)));
                    tile[(size_t)t * P.tpitch + n] = p;
                }
            }
        } else {
            switch (g.K) {
                case 4: short3_interp<T, 4, MODE>(g, buf, tile, tpsh, nvalid, P.tpitch, tid, nthr); break;
                case 6: short3_interp<T, 6, MODE>(g, buf, tile, tpsh, nvalid, P.tpitch, tid, nthr); break;
                case 8: short3_interp<T, 8, MODE>(g, buf, tile, tpsh, nvalid, P.tpitch, tid, nthr); break;
                case 10: short3_interp<T, 10, MODE>(g, buf, tile, tpsh, nvalid, P.tpitch, tid, nthr); break;
                default: short3_interp<T, 12, MODE>(g, buf, tile, tpsh, nvalid, P.tpitch, tid, nthr); break;
            }
        }
        NW_SYNC();
        // ---- rows out: row r = 2 t + lane  ->  out[s0 + lane][fmap[f0 + t]][:] --------------------------
        // sample n of row t sits at tile[t][(n / D) DP + n % D]
        const int DP = g.DP;
        const int nrows = 2 * nvalid;
        const bool blon = P.bl_mode != BL_NONE;
        const bool uselog = P.bl_mode == BL_LOG || P.bl_mode == BL_ZLOG;
        auto slot = [&](int n) -> int {
            const int m = (int)fd_div((uint32_t)n, g.dD);
            return m * DP + (n - m * D);
        };
        if (blon) {
            // window statistics of every row (np.mean / np.std, ddof = 0; base.py:49-50, 65): sums of d = x - x[lo]
            // and d^2 in fp64, the window of a row split over the warps the CTA has per row
            const int cnt = P.bl_hi - P.bl_lo;
#if defined(__CUDA_ARCH__)
            const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
            const int wpr = nwarp / nrows > 0 ? (nwarp / nrows > 4 ? 4 : nwarp / nrows) : 1;   // warps per row
            for (int job = warp; job < nrows * wpr; job += nwarp) {
                const int r = job / wpr, wpart = job - r * wpr;
                const int t = r >> 1, l = r & 1;
                const pk<T>* row = tile + (size_t)t * P.tpitch;
                double s = 0.0, q = 0.0;
                if (cnt > 0) {
                    const pk<T> q0 = row[slot(P.bl_lo)];
                    const double x0 = (double)(l ? pk_hi(q0) : pk_lo(q0));
                    for (int i = P.bl_lo + wpart * 32 + lane; i < P.bl_hi; i += 32 * wpr) {
                        const pk<T> pv = row[slot(i)];
                        const double d = (double)(l ? pk_hi(pv) : pk_lo(pv)) - x0;
                        s += d;
                        q += d * d;
                    }
                }
                for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
                if (lane == 0) { rstat[32 + 2 * (r * 4 + wpart)] = s; rstat[32 + 2 * (r * 4 + wpart) + 1] = q; }
            }
            NW_SYNC();
            if (tid < nrows) {
                const int r = tid, t = r >> 1, l = r & 1;
                const pk<T>* row = tile + (size_t)t * P.tpitch;
                double s = 0.0, q = 0.0;
                for (int wpart = 0; wpart < wpr; ++wpart) { s += rstat[32 + 2 * (r * 4 + wpart)]; q += rstat[32 + 2 * (r * 4 + wpart) + 1]; }
                double m = nan(""), sd = nan("");
                if (cnt > 0) {
                    const pk<T> q0 = row[slot(P.bl_lo)];
                    const double x0 = (double)(l ? pk_hi(q0) : pk_lo(q0));
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
                const pk<T>* row = tile + (size_t)t * P.tpitch;
                double s = 0.0, q = 0.0, x0 = 0.0;
                if (cnt > 0) { const pk<T> q0 = row[slot(P.bl_lo)]; x0 = (double)(l ? pk_hi(q0) : pk_lo(q0)); }
                for (int i = P.bl_lo; i < P.bl_hi; ++i) {
                    const pk<T> pv = row[slot(i)];
                    const double d = (double)(l ? pk_hi(pv) : pk_lo(pv)) - x0;
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
        // Baseline as y = (x + b) * a [, log10(y) * c]; one thread per sample n, all rows of the unit
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
            const pk<T>* row = tile + (size_t)t * P.tpitch;
            T* row0 = (T*)P.out + ((size_t)s0 * P.F_out + g.fmap[f0 + t]) * (size_t)N;
            T* row1 = row0 + (size_t)P.F_out * (size_t)N;
            for (int n = tid; n < N; n += nthr) {
                const pk<T> v = row[slot(n)];
                T y0 = pk_lo(v), y1 = pk_hi(v);
                if (blon) {
                    y0 = (y0 + b0) * a0;
                    y1 = (y1 + b1) * a1;
                    if (uselog) { y0 = nw_log10(y0) * c0; y1 = nw_log10(y1) * c1; }
                }
                st_stream(row0 + n, y0);
                if (has1) st_stream(row1 + n, y1);
            }
        }
        NW_SYNC();
    }
}

}  // namespace nw
