// Fused short-row kernel with resampled rows (r02): rows that fit one CTA, band-limited frequencies computed at a
// decimated length and interpolated (DESIGN.md "resampled rows"; nw_plan.h: plan_multirate(short)).
//
// One CTA owns a PAIR of signals (the two lanes of every packed value) and a share of the work units:
//   1. forward transform of both signals (scipy.fftpack.fft, base.py:399) as in nw_kernels3.cuh: the natural-order
//      spectrum pair Xs[k] stays in shared memory for the whole CTA;
//   2. a unit = NF = 2^tpsh frequencies of one resample group (decimation D, M = N / D; NF as many as fit the tile): zero
//      the tile, gather W_f(k) Xs[k] (weights from the group's table: 1/N and the equaliser folded in; base.py:236-248,
//      404) of each band, moved to bin 0, into the decimation-in-time slots; in-place M-point inverse transform of the NF
//      interleaved lane-pair sequences (base.py:406) - y[m] = z(m D) up to a unit-modulus factor;
//   3. interpolation straight to global memory, as in nw_resample.cuh's direct kernel: lane = (phase group g of PQ
//      neighbouring phases, run j of R consecutive m); the group's PQ K weights of the lane stay in registers for the whole
//      unit, the R + K - 1 window samples (both signals: the two lanes of a packed value) come from the tile;
//      |z(m D + p)|^2 = |sum_t coef[p][t] y[m + 1 - K/2 + t]|^2 costs K packed FMAs for the real parts of both signals and K
//      for the imaginary parts; the Baseline epilogue (x + b) a [, log10] is applied in registers and the PQ outputs of each
//      signal leave as one 4 PQ-byte streaming store - the G = D / PQ lanes of a run cover the D consecutive outputs of m;
//   4. Baseline (base.py:46-68): the window statistics need the window's samples first, so the same interpolation runs over
//      the window's items only (200 of 1500 samples on config 3), accumulating sum and sum of squares in fp64 per row.
// Shared memory: 2 N two-lane complex values (spectrum pair + tile / forward work space) + the rows' statistics.
#pragma once
#include <string.h>
#include "nw_common.h"
#include "nw_fft2.cuh"
#include "nw_family.cuh"
#include "nw_kernels.cuh"
#include "nw_kernels2.cuh"
#include "nw_kernels3.cuh"
#include "nw_resample.cuh"

namespace nw {

template <typename T>
struct Short3Group {
    int M, D, K, F;          // decimated length, decimation (1: exact rows), taps, frequencies of the group
    int tpsh;                // NF = 1 << tpsh frequencies per unit
    int PQ, G, MW;           // outputs per store, lanes per m, runs per warp (nw_resample.cuh: direct kernel)
    int unit0, nunits;       // this group's units: [unit0, unit0 + nunits)
    int items;               // warp-items per row: ceil(M / (MW R)); D == 1: ceil(N / (32 PQ))
    Fft2Plan st;             // M-point plan
    fastdiv dG, dItems;      // lane / G, x / items
    const cx<T>* tw;         // [M]
    const FreqRec* rec;      // [F] bands centred on transform bin 0 (FreqRec::shift), woff into wtab
    const T* wtab;           // weights: W_f / N (x equaliser)
    const T* coefq;          // [D / PQ][K][PQ]
    const int* fmap;         // [F] -> plan frequency index
    const int* ditpos;       // [M] fft2_dit_pos(st, k): slot of spectrum bin k in the decimation-in-time tile
};

template <typename T>
struct Short3Params {
    const T* signals;   // [S][N] real
    void* out;          // [S][F_out][N] T
    int N, F_out, S;
    int out_mode, bl_mode, bl_lo, bl_hi;
    int fsplit;         // CTAs per signal pair (units are dealt round-robin)
    Fft2Plan st;        // N-point plan (forward transform)
    const cx<T>* tw;    // [N]
    const Short3Group<T>* groups;
    int ngroups, nunits;
};

// run length by outputs per store: the window holds (R + K - 1) two-lane complex values = 4 (R + K - 1) registers, a lane
// produces 2 R PQ outputs per item (24, 20, 20)
template <int PQ> struct S3Run { static const int R = PQ == 4 ? 3 : PQ == 2 ? 5 : 10; };
inline int s3_run(int PQ) { return PQ == 4 ? 3 : PQ == 2 ? 5 : 10; }
constexpr int S3_MAXROWS = 16; // rows of a unit: 2 signals x 8 frequencies

// host side: launch geometry of a group inside the kernel (shared by nwcwt.cu and tests/emul)
template <typename T>
inline bool short3_fill_group(Short3Group<T>& g, long long N, long long M, int D, int K, int F) {
    memset(&g, 0, sizeof(g));
    g.M = (int)M; g.D = D; g.K = D > 1 ? K : 0; g.F = F;
    int nf = 1;
    while (nf < 8 && (long long)(2 * nf) * M <= N) nf *= 2;
    g.tpsh = nf == 8 ? 3 : nf == 4 ? 2 : nf == 2 ? 1 : 0;
    g.PQ = (D % 4 == 0 && N % 4 == 0) ? 4 : (D % 2 == 0 && N % 2 == 0) ? 2 : 1;
    if (D == 1) g.PQ = N % 4 == 0 ? 4 : N % 2 == 0 ? 2 : 1;
    g.G = D > 1 ? D / g.PQ : 1;
    if (g.G > 32) return false;
    g.MW = 32 / g.G;
    g.items = D > 1 ? (int)((M + (long long)g.MW * s3_run(g.PQ) - 1) / ((long long)g.MW * s3_run(g.PQ))) : 0;
    g.dG = make_fastdiv((uint32_t)g.G);
    g.dItems = make_fastdiv((uint32_t)(g.items > 0 ? g.items : 1));
    g.nunits = (F + nf - 1) / nf;
    return true;
}

template <typename T> NW_HD size_t short3_smem_bytes(int N) {
    return 2 * (size_t)N * sizeof(cx2<T>) + (size_t)(4 * S3_MAXROWS) * sizeof(double) + 64 * 12 * sizeof(T);
}

// Baseline epilogue of one row pair as y = v * a + ba [, log10(y) * c]  (ba = b * a)
template <typename T> struct S3Epi { pk<T> a, ba, c; };

#if defined(__CUDA_ARCH__)
NW_D void s3_acc(double* rs, double s0, double q0, double s1, double q1) {   // warp reduction, then one atomic per value
    for (int o = 16; o; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(rs, s0); atomicAdd(rs + 1, q0); atomicAdd(rs + 2, s1); atomicAdd(rs + 3, q1); }
}
#else
inline void s3_acc(double* rs, double s0, double q0, double s1, double q1) { rs[0] += s0; rs[1] += q0; rs[2] += s1; rs[3] += q1; }
#endif

template <typename T, int PQ> NW_HD void s3_store(T* p, const T* v) {
    RsVec<T, PQ> r;
#pragma unroll
    for (int q = 0; q < PQ; ++q) r.v[q] = v[q];
    rs_st_out(p, r);
}
template <typename T> NW_HD void s3_store1(T* p, const T* v) { st_stream(p, v[0]); }

// Interpolation of the unit's rows.  STATS: accumulate the window statistics into rstat[4 t ..] instead of storing.
// Items (t, piece of MW R consecutive m) are dealt to the CTA's warps round-robin.
template <typename T, int K, int PQ, bool STATS>
NW_HD void short3_interp(const Short3Params<T>& P, const Short3Group<T>& g, const cx2<T>* buf, double* rstat, const S3Epi<T>* epi,
                         const T* coefs, int nvalid, int f0, int s0, bool has1, int tid, int nthr) {
    constexpr bool stats = STATS;
    constexpr int R = S3Run<PQ>::R;
    const int M = g.M, D = g.D, N = P.N, tpsh = g.tpsh;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int j = (int)fd_div((uint32_t)lane, g.dG), gq = lane - j * g.G;
    const bool act = j < g.MW;
    T c[K][PQ];
    {
        const T* cq = coefs + (size_t)(act ? gq : 0) * (K * PQ);   // the group's weights, staged in shared memory
#pragma unroll
        for (int t = 0; t < K; ++t)
#pragma unroll
            for (int q = 0; q < PQ; ++q) c[t][q] = cq[t * PQ + q];
    }
    const int CW = g.MW * R;
    // window of the statistics: items that hold a sample of [bl_lo, bl_hi)
    int it_lo = 0, it_hi = g.items;
    if (stats) {
        it_lo = (P.bl_lo / D) / CW;
        it_hi = P.bl_hi > P.bl_lo ? ((P.bl_hi - 1) / D) / CW + 1 : it_lo;
    }
    const int nit = it_hi - it_lo;
    const bool uselog = P.bl_mode == BL_LOG || P.bl_mode == BL_ZLOG;
    const bool isabs = P.out_mode == OUT_ABS;
    for (int w = warp; w < nit * nvalid; w += nwarp) {
        const int t = w / nit, it = it_lo + (w - t * nit);
        const int m0 = it * CW + (act ? j : 0) * R;
        cx2<T> win[R + K - 1];
        {
            int mi = m0 + 1 - K / 2;
            if (mi < 0) mi += M;
            if (mi >= M) mi -= M;   // runs past the end of the row (their outputs are not stored)
            if (mi >= M) mi = 0;
#pragma unroll
            for (int o = 0; o < R + K - 1; ++o) {
                win[o] = buf[((size_t)mi << tpsh) + t];
                if (++mi == M) mi = 0;
            }
        }
        double s0a = 0, q0a = 0, s1a = 0, q1a = 0;
        T* row0 = (T*)P.out + ((size_t)s0 * P.F_out + g.fmap[f0 + t]) * (size_t)N;
        T* row1 = row0 + (size_t)P.F_out * (size_t)N;
        const S3Epi<T> e = epi[t];
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
            T o0[PQ], o1[PQ];
#pragma unroll
            for (int q = 0; q < PQ; ++q) {
                pk<T> are = pk_bcast((T)0), aim = are;
#pragma unroll
                for (int o = 0; o < K; ++o) {
                    are = pk_fma(win[mm + o].re, c[o][q], are);
                    aim = pk_fma(win[mm + o].im, c[o][q], aim);
                }
                pk<T> v = pk_fma(aim, aim, are * are);
                if (isabs) v = pk_make(nw_sqrt(pk_lo(v)), nw_sqrt(pk_hi(v)));
                if (!stats) {
                    v = pk_fma(v, e.a, e.ba);
                    if (uselog) v = pk_make(nw_log10(pk_lo(v)), nw_log10(pk_hi(v))) * e.c;
                }
                o0[q] = pk_lo(v);
                o1[q] = pk_hi(v);
            }
            const int m = m0 + mm;
            const int n = m * D + gq * PQ;
            if (!act || m >= M) continue;
            if (stats) {
#pragma unroll
                for (int q = 0; q < PQ; ++q)
                    if (n + q >= P.bl_lo && n + q < P.bl_hi) {
                        const double d0 = (double)o0[q], d1 = (double)o1[q];
                        s0a += d0; q0a += d0 * d0; s1a += d1; q1a += d1 * d1;
                    }
            } else {
                if (PQ == 1) { s3_store1<T>(row0 + n, o0); if (has1) s3_store1<T>(row1 + n, o1); }
                else { s3_store<T, (PQ > 1 ? PQ : 2)>(row0 + n, o0); if (has1) s3_store<T, (PQ > 1 ? PQ : 2)>(row1 + n, o1); }
            }
        }
        if (stats) s3_acc(rstat + 4 * t, s0a, q0a, s1a, q1a);
    }
}

// exact rows (D == 1): |y|^2 of the tile itself, PQ consecutive samples per lane
template <typename T, int PQ, bool STATS>
NW_HD void short3_exact(const Short3Params<T>& P, const Short3Group<T>& g, const cx2<T>* buf, double* rstat, const S3Epi<T>* epi,
                        int nvalid, int f0, int s0, bool has1, int tid, int nthr) {
    constexpr bool stats = STATS;
    const int N = P.N, tpsh = g.tpsh;
    const bool uselog = P.bl_mode == BL_LOG || P.bl_mode == BL_ZLOG;
    const bool isabs = P.out_mode == OUT_ABS;
    for (int t = 0; t < nvalid; ++t) {
        T* row0 = (T*)P.out + ((size_t)s0 * P.F_out + g.fmap[f0 + t]) * (size_t)N;
        T* row1 = row0 + (size_t)P.F_out * (size_t)N;
        const S3Epi<T> e = epi[t];
        double s0a = 0, q0a = 0, s1a = 0, q1a = 0;
        const int lo = stats ? P.bl_lo / PQ * PQ : 0, hi = stats ? P.bl_hi : N;
        for (int n = lo + tid * PQ; n < hi; n += nthr * PQ) {
            T o0[PQ], o1[PQ];
#pragma unroll
            for (int q = 0; q < PQ; ++q) {
                const cx2<T> y = buf[((size_t)(n + q) << tpsh) + t];
                pk<T> v = pk_fma(y.im, y.im, y.re * y.re);
                if (isabs) v = pk_make(nw_sqrt(pk_lo(v)), nw_sqrt(pk_hi(v)));
                if (!stats) {
                    v = pk_fma(v, e.a, e.ba);
                    if (uselog) v = pk_make(nw_log10(pk_lo(v)), nw_log10(pk_hi(v))) * e.c;
                }
                o0[q] = pk_lo(v);
                o1[q] = pk_hi(v);
            }
            if (stats) {
#pragma unroll
                for (int q = 0; q < PQ; ++q)
                    if (n + q >= P.bl_lo && n + q < P.bl_hi) {
                        const double d0 = (double)o0[q], d1 = (double)o1[q];
                        s0a += d0; q0a += d0 * d0; s1a += d1; q1a += d1 * d1;
                    }
            } else {
                if (PQ == 1) { s3_store1<T>(row0 + n, o0); if (has1) s3_store1<T>(row1 + n, o1); }
                else { s3_store<T, (PQ > 1 ? PQ : 2)>(row0 + n, o0); if (has1) s3_store<T, (PQ > 1 ? PQ : 2)>(row1 + n, o1); }
            }
        }
        if (stats) s3_acc(rstat + 4 * t, s0a, q0a, s1a, q1a);   // every thread of the CTA takes part (warp shuffles)
    }
}

template <typename T, bool STATS>
NW_HD void short3_rows(const Short3Params<T>& P, const Short3Group<T>& g, const cx2<T>* buf, double* rstat, const S3Epi<T>* epi,
                       const T* coefs, int nvalid, int f0, int s0, bool has1, int tid, int nthr) {
#define NW_S3_CALL(k, pq) short3_interp<T, k, pq, STATS>(P, g, buf, rstat, epi, coefs, nvalid, f0, s0, has1, tid, nthr)
#define NW_S3_K(pq) \
    switch (g.K) { case 4: NW_S3_CALL(4, pq); break; case 6: NW_S3_CALL(6, pq); break; case 8: NW_S3_CALL(8, pq); break; \
                   case 10: NW_S3_CALL(10, pq); break; default: NW_S3_CALL(12, pq); break; }
    if (g.D == 1) {
        if (g.PQ == 4) short3_exact<T, 4, STATS>(P, g, buf, rstat, epi, nvalid, f0, s0, has1, tid, nthr);
        else if (g.PQ == 2) short3_exact<T, 2, STATS>(P, g, buf, rstat, epi, nvalid, f0, s0, has1, tid, nthr);
        else short3_exact<T, 1, STATS>(P, g, buf, rstat, epi, nvalid, f0, s0, has1, tid, nthr);
        return;
    }
    if (g.PQ == 4) { NW_S3_K(4) }
    else if (g.PQ == 2) { NW_S3_K(2) }
    else { NW_S3_K(1) }
#undef NW_S3_K
#undef NW_S3_CALL
}

template <typename T>
NW_HD void short3_body(const Short3Params<T>& P, char* smem, int bx, int tid, int nthr) {
    const int N = P.N;
    cx2<T>* Xs = (cx2<T>*)smem;
    cx2<T>* buf = Xs + N;
    double* rstat = (double*)(buf + N);                     // [NF][2 signals][sum, sum of squares]
    S3Epi<T>* epi = (S3Epi<T>*)(rstat + 4 * (S3_MAXROWS / 2));
    T* coefs = (T*)(rstat + 4 * S3_MAXROWS);              // [D / PQ][K][PQ] of the current group (<= 64 x 12 values)
    int staged = -1;
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
    const bool blon = P.bl_mode != BL_NONE;
    int gi = 0;
    for (int u = part; u < P.nunits; u += P.fsplit) {
        while (u >= P.groups[gi].unit0 + P.groups[gi].nunits) ++gi;
        const Short3Group<T>& g = P.groups[gi];
        const int M = g.M, tpsh = g.tpsh, NF = 1 << tpsh;
        const int f0 = (u - g.unit0) << tpsh;
        const int nvalid = (g.F - f0 < NF) ? (g.F - f0) : NF;
        // ---- the in-band products, band centre at bin 0, into the zero tile -------------------------
        if (staged != gi) {                                 // the previous unit's readers passed its trailing barrier
            for (int i = tid; i < g.D * g.K; i += nthr) coefs[i] = g.coefq[i];
            staged = gi;
        }
        if (tid < 4 * NF) rstat[tid] = 0.0;
        if (tid < NF) { S3Epi<T> e; e.a = pk_bcast((T)1); e.ba = pk_bcast((T)0); e.c = pk_bcast((T)1); epi[tid] = e; }
        // every slot of the tile is written once: the band's product (bin jb = k or k - M, whichever lies in the band)
        // or zero
        for (int t = 0; t < NF; ++t) {
            FreqRec rec;
            rec.lo = rec.hi = 0; rec.shift = 0; rec.woff = 0;
            if (t < nvalid) rec = g.rec[f0 + t];
            const T* wt = g.wtab + rec.woff - rec.lo;
            for (int k = tid; k < M; k += nthr) {
                int jb = k;
                if (jb >= rec.hi) jb -= M;
                cx2<T> v = z;
                if (jb >= rec.lo && jb < rec.hi) {
                    const cx2<T> x = Xs[jb + rec.shift];
                    const T w = wt[jb];
                    v = mk2<T>(x.re * w, x.im * w);
                }
                buf[((size_t)g.ditpos[k] << tpsh) + t] = v;
            }
        }
        NW_SYNC();
        // ---- M-point inverse transform, complex result left in place ---------------------------------
        {
            InPlaceOutDst<T, OUT_CWT> dst{buf, tpsh};
            fft2_dit<T, +1>(g.st, tpsh, g.tw, buf, FromBuf(), dst, tid, nthr);
        }
        NW_SYNC();
        // ---- Baseline: statistics of the window of every row, then y = (x + b) a [, log10(y) c] ----------
        if (blon) {
            short3_rows<T, true>(P, g, buf, rstat, epi, coefs, nvalid, f0, s0, has1, tid, nthr);
            NW_SYNC();
            if (tid < 2 * nvalid) {
                const int t = tid >> 1, l = tid & 1;
                const int cnt = P.bl_hi - P.bl_lo;
                double m = nan(""), sd = nan("");
                if (cnt > 0) {
                    const double ms = rstat[4 * t + 2 * l] / cnt;
                    double var = rstat[4 * t + 2 * l + 1] / cnt - ms * ms;
                    if (var < 0.0) var = 0.0;
                    m = ms;
                    sd = sqrt(var);
                }
                T b = (T)0, a = (T)1, c = (T)1;
                switch (P.bl_mode) {
                    case BL_MEAN: b = (T)-m; break;
                    case BL_RATIO: a = (T)(1.0 / m); break;
                    case BL_PERCENT: b = (T)-m; a = (T)(1.0 / m); break;
                    case BL_ZSCORE: b = (T)-m; a = (T)(1.0 / sd); break;
                    case BL_LOG: a = (T)(1.0 / m); break;
                    case BL_ZLOG: a = (T)(1.0 / m); c = (T)(1.0 / sd); break;
                    default: break;
                }
                T* ea = (T*)&epi[t].a; T* eb = (T*)&epi[t].ba; T* ec = (T*)&epi[t].c;
                ea[l] = a; eb[l] = b * a; ec[l] = c;
            }
            NW_SYNC();
        }
        short3_rows<T, false>(P, g, buf, rstat, epi, coefs, nvalid, f0, s0, has1, tid, nthr);
        NW_SYNC();
    }
}

}  // namespace nw
