// Kernel bodies of the CWT path.  Each body is written against (block index,
// tid, nthr, shared-memory base) so the same code is launched as an sm_100a
// kernel (nwcwt.cu) and stepped block-by-block on the host by tests/emul.
//
//   short_body   N fits one CTA: forward FFT of the signal (base.py:399), then per
//                group of TT analysis frequencies: spectrum generation (base.py:
//                236-248 + wavelets.py formulas) x signal spectrum -> inverse FFT
//                (base.py:406) -> |z|^2 / |z| / z epilogue (base.py:425,443), with
//                the optional Baseline epilogue (base.py:46-68).
//   passA_body   long rows, first half of the four-step split N = N1*N2:
//                N1-point transforms down the columns + twiddle, written to the
//                L2-resident intermediate in TB-blocked layout.
//   passB_body   second half: bulk (TMA) tile load, N2-point transforms, epilogue.
//   baseline_rows_body   Baseline epilogue for long rows (second sweep over a row).
#pragma once
#include "nw_common.h"
#include "nw_fft.cuh"
#include "nw_family.cuh"

namespace nw {

// ------------------------------------------------------------------------------
// output conversion
// ------------------------------------------------------------------------------
NW_HD float nw_sqrt(float x) { return sqrtf(x); }
NW_HD double nw_sqrt(double x) { return sqrt(x); }
NW_HD float nw_log10(float x) { return log10f(x); }
NW_HD double nw_log10(double x) { return log10(x); }
NW_HD float nw_hypot(float a, float b) { return hypotf(a, b); }
NW_HD double nw_hypot(double a, double b) { return hypot(a, b); }

template <typename T> NW_HD T real_out(int mode, cx<T> v) {
    if (mode == OUT_POWER) {
        // reference: np.abs(z) ** 2 (base.py:425,443); |z|^2 directly differs by <= 1 ulp
        return v.x * v.x + v.y * v.y;
    }
    return nw_hypot(v.x, v.y);
}

template <typename T> NW_HD T baseline_apply(int mode, T v, T m, T sd) {
    switch (mode) {
        case BL_MEAN: return v - m;                       // base.py:52-53
        case BL_RATIO: return v / m;                      // base.py:55-56
        case BL_PERCENT: return (v - m) / m;              // base.py:58-59
        case BL_LOG: return nw_log10(v / m);              // base.py:61-62
        case BL_ZSCORE: return (v - m) / sd;              // base.py:64-65
        case BL_ZLOG: return nw_log10(v / m) / sd;        // base.py:67-68
        default: return v;
    }
}

// mean and population std (ddof=0, np.std) of row[lo:hi); one warp per call on the
// device, serial on the host emulation.  Accumulates in double.
template <typename T>
NW_HD void window_stats(const T* row, int lo, int hi, int lane, int nlanes, double* mean, double* sd) {
    double s = 0.0;
    for (int i = lo + lane; i < hi; i += nlanes) s += (double)row[i];
#if defined(__CUDA_ARCH__)
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#endif
    const int cnt = hi - lo;
    const double m = cnt > 0 ? s / cnt : nan("");
    double q = 0.0;
    for (int i = lo + lane; i < hi; i += nlanes) {
        const double d = (double)row[i] - m;
        q += d * d;
    }
#if defined(__CUDA_ARCH__)
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
#endif
    *mean = m;
    *sd = cnt > 0 ? sqrt(q / cnt) : nan("");
}

// ------------------------------------------------------------------------------
// short rows
// ------------------------------------------------------------------------------
template <typename T>
struct ShortParams {
    const T* signals;   // [S][N] real
    void* out;          // [S][F][N] T (abs/power) or cx<T> (cwt)
    cx<T>* spectra;     // optional [S][N]: forward-only mode when F == 0
    int N, F, S;
    int tsh, pitch;     // TT = 1 << tsh frequencies interleaved; pitch = TT + 1 (or 1)
    int out_mode, bl_mode, bl_lo, bl_hi;
    int fsplit;         // CTAs per signal (frequency groups are dealt round-robin)
    FftStages st;
    const cx<T>* tw;    // [N]
    SpecParams<T> sp;
};

template <typename T> struct RealSrc {
    const T* x;
    NW_HD cx<T> load(int p, int) const { return mk<T>(x[p], (T)0); }
};

template <typename T> struct GlobalCxDst {
    cx<T>* y;
    NW_HD void store(int p, int, cx<T> v) const { y[p] = v; }
};

template <typename T> struct ShortSpecSrc {
    const SpecParams<T>* sp;
    const cx<T>* X;       // shared memory, [N]
    const FreqRec* rec;   // shared memory, [TT] records of the current group
    int f0, F;
    NW_HD cx<T> load(int p, int t) const {
        const FreqRec& r = rec[t];
        if (p < r.lo || p >= r.hi) return mk<T>((T)0, (T)0);
        return spec_times<T>(*sp, r, f0 + t, p, X[p]);
    }
};

template <typename T> struct ShortGlobalDst {
    void* out;   // row base of (signal, frequency f0)
    int N, nvalid, mode;
    NW_HD void store(int p, int t, cx<T> v) const {
        if (t >= nvalid) return;
        const size_t idx = (size_t)t * (size_t)N + (size_t)p;
        if (mode == OUT_CWT) ((cx<T>*)out)[idx] = v;
        else ((T*)out)[idx] = real_out<T>(mode, v);
    }
};

template <typename T> struct ShortStageDst {
    T* stag;     // shared memory [TT][NP]
    int NP, mode;
    NW_HD void store(int p, int t, cx<T> v) const { stag[t * NP + p] = real_out<T>(mode, v); }
};

template <typename T> NW_HD size_t short_smem_bytes(int N, int pitch, int TT) {
    size_t cxs = (size_t)N * (1 + 2 * (size_t)pitch) * sizeof(cx<T>);
    return cxs + (size_t)TT * (sizeof(FreqRec) + 2 * sizeof(double));
}

template <typename T>
NW_HD void short_body(const ShortParams<T>& P, char* smem, int bx, int tid, int nthr) {
    const int N = P.N, TT = 1 << P.tsh;
    cx<T>* Xs = (cx<T>*)smem;
    cx<T>* bufA = Xs + N;
    cx<T>* bufB = bufA + (size_t)N * P.pitch;
    FreqRec* rec = (FreqRec*)(bufB + (size_t)N * P.pitch);
    double* rstat = (double*)(rec + TT);

    const int s = bx / P.fsplit, part = bx - s * P.fsplit;

    // forward transform of the signal (scipy.fftpack.fft, base.py:399)
    {
        RealSrc<T> src{P.signals + (size_t)s * N};
        if (P.F == 0) {
            GlobalCxDst<T> dst{P.spectra + (size_t)s * N};
            fft_run<T, -1, 1>(P.st, 0, 1, bufA, bufB, P.tw, src, dst, tid, nthr);
            return;
        }
        SmemDst<T> dst{Xs, 1};
        fft_run<T, -1, 0>(P.st, 0, 1, bufA, bufB, P.tw, src, dst, tid, nthr);
    }

    // which ping-pong buffer the last stage does NOT read from
    cx<T>* freebuf = (P.st.nst <= 1) ? bufA : (((P.st.nst - 2) & 1) ? bufA : bufB);
    const int NP = (N + 3) & ~3;
    const int ngroups = (P.F + TT - 1) / TT;
    for (int g = part; g < ngroups; g += P.fsplit) {
        const int f0 = g * TT;
        const int nvalid = (P.F - f0 < TT) ? (P.F - f0) : TT;
        NW_SYNC();
        for (int t = tid; t < TT; t += nthr) {
            if (t < nvalid) rec[t] = P.sp.rec[f0 + t];
            else { rec[t].lo = 0; rec[t].hi = 0; rec[t].toff = 0; rec[t].freq = 1; rec[t].aux = 1; rec[t].kx = 0; }
        }
        NW_SYNC();
        ShortSpecSrc<T> src{&P.sp, Xs, rec, f0, P.F};
        const size_t esz = (P.out_mode == OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
        char* rowbase = (char*)P.out + ((size_t)s * P.F + f0) * (size_t)N * esz;
        if (P.bl_mode == BL_NONE) {
            ShortGlobalDst<T> dst{rowbase, N, nvalid, P.out_mode};
            fft_run<T, +1, 1>(P.st, P.tsh, P.pitch, bufA, bufB, P.tw, src, dst, tid, nthr);
        } else {
            T* stag = (T*)freebuf;
            ShortStageDst<T> dst{stag, NP, P.out_mode};
            fft_run<T, +1, 1>(P.st, P.tsh, P.pitch, bufA, bufB, P.tw, src, dst, tid, nthr);
            // Baseline statistics over [bl_lo, bl_hi) of every row (base.py:49-50, 65)
#if defined(__CUDA_ARCH__)
            const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
            for (int t = warp; t < nvalid; t += nwarp) {
                double m, sd;
                window_stats<T>(stag + t * NP, P.bl_lo, P.bl_hi, lane, 32, &m, &sd);
                if (lane == 0) { rstat[2 * t] = m; rstat[2 * t + 1] = sd; }
            }
#else
            for (int t = tid; t < nvalid; t += nthr)
                window_stats<T>(stag + t * NP, P.bl_lo, P.bl_hi, 0, 1, &rstat[2 * t], &rstat[2 * t + 1]);
#endif
            NW_SYNC();
            T* orow = (T*)rowbase;
            for (int t = 0; t < nvalid; ++t) {
                const T m = (T)rstat[2 * t], sd = (T)rstat[2 * t + 1];
                for (int p = tid; p < N; p += nthr)
                    orow[(size_t)t * N + p] = baseline_apply<T>(P.bl_mode, stag[t * NP + p], m, sd);
            }
        }
    }
}

// ------------------------------------------------------------------------------
// long rows: four-step split  N = N1 * N2,  k = N2*k1 + k2,  n = n1 + N1*n2
//   A[n1][k2] = sum_k1 Y[N2 k1 + k2] w_N1^{k1 n1}
//   Tm[n1][k2] = A[n1][k2] w_N^{k2 n1}
//   z[n1 + N1 n2] = sum_k2 Tm[n1][k2] w_N2^{k2 n2}
// Tm is stored TB-blocked: element (n1, k2) at ((n1/TB)*N2 + k2)*TB + n1%TB, so a
// pass-B tile (TB consecutive n1, all k2) is one contiguous chunk already in the
// interleaved shared-memory layout of the engine -> a single bulk copy.
// ------------------------------------------------------------------------------
template <typename T>
struct LongParams {
    long long N;
    int N1, N2;
    int tshA, pitchA;   // pass A: TA = 1<<tshA columns per CTA, pitch TA+1
    int tshB;           // pass B: TB = 1<<tshB rows per CTA, dense pitch TB
    FftStages stA, stB;
    const cx<T>* twA;   // [N1]
    const cx<T>* twB;   // [N2]
    const cx<T>* twH;   // w_N^{j << lb}
    const cx<T>* twL;   // w_N^{j}, j < 1<<lb
    int lb;
    // data
    const T* signal;        // forward pass A: one real signal [N]
    const cx<T>* X;         // inverse pass A: spectrum of the signal [N]
    cx<T>* Xout;            // forward pass B: spectrum out [N]
    cx<T>* Tm;              // intermediate ring, [rows][ceil(N1/TB)*N2*TB]
    long long tm_stride;    // elements per ring slot
    void* out;              // inverse pass B: row 0 of this launch
    int f0;                 // first frequency index of this launch (rows = blockIdx.y)
    int out_mode;
    SpecParams<T> sp;
};

template <typename T, int DIR> NW_HD cx<T> big_twiddle(const LongParams<T>& P, int m) {
    const cx<T> a = P.twH[m >> P.lb];
    const cx<T> b = P.twL[m & ((1 << P.lb) - 1)];
    const cx<T> w = cmul(a, b);
    return DIR > 0 ? w : mk<T>(w.x, -w.y);
}

template <typename T> struct LongRealSrc {   // forward pass A
    const T* x;
    int N2, c;
    NW_HD cx<T> load(int p, int t) const {
        const int k2 = c + t;
        if (k2 >= N2) return mk<T>((T)0, (T)0);
        return mk<T>(x[(size_t)p * N2 + k2], (T)0);
    }
};

template <typename T> struct LongSpecSrc {   // inverse pass A
    const SpecParams<T>* sp;
    const cx<T>* X;
    FreqRec r;
    int fi, N2, c;
    NW_HD cx<T> load(int p, int t) const {
        const int k2 = c + t;
        const long long k = (long long)p * N2 + k2;
        if (k2 >= N2 || k < r.lo || k >= r.hi) return mk<T>((T)0, (T)0);
        return spec_times<T>(*sp, r, fi, (int)k, X[k]);
    }
};

template <typename T, int DIR> struct LongTmDst {   // pass A epilogue: twiddle + blocked store
    const LongParams<T>* P;
    cx<T>* tm;
    int c;
    NW_HD void store(int n1, int t, cx<T> v) const {
        const int k2 = c + t;
        if (k2 >= P->N2) return;
        const cx<T> w = big_twiddle<T, DIR>(*P, k2 * n1);
        const int tb = 1 << P->tshB;
        const size_t idx = (((size_t)(n1 >> P->tshB) * P->N2 + k2) << P->tshB) + (n1 & (tb - 1));
        tm[idx] = cmul(v, w);
    }
};

template <typename T> struct LongOutDst {   // inverse pass B epilogue
    void* out;
    int N1, r0, mode;
    NW_HD void store(int n2, int t, cx<T> v) const {
        const int n1 = r0 + t;
        if (n1 >= N1) return;
        const size_t idx = (size_t)n1 + (size_t)N1 * n2;
        if (mode == OUT_CWT) ((cx<T>*)out)[idx] = v;
        else ((T*)out)[idx] = real_out<T>(mode, v);
    }
};

template <typename T> struct LongSpectrumDst {   // forward pass B epilogue
    cx<T>* X;
    int N1, r0;
    NW_HD void store(int k2, int t, cx<T> v) const {
        const int k1 = r0 + t;
        if (k1 >= N1) return;
        X[(size_t)k1 + (size_t)N1 * k2] = v;
    }
};

template <typename T> NW_HD size_t passA_smem_bytes(int N1, int pitchA) {
    return 2 * (size_t)N1 * pitchA * sizeof(cx<T>);
}
template <typename T> NW_HD size_t passB_smem_bytes(int N2, int TB) {
    return 2 * (size_t)N2 * TB * sizeof(cx<T>) + 16;
}

// DIR=-1: forward (signal -> Tm);  DIR=+1: inverse (spectrum generation -> Tm)
template <typename T, int DIR>
NW_HD void passA_body(const LongParams<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    cx<T>* bufA = (cx<T>*)smem;
    cx<T>* bufB = bufA + (size_t)P.N1 * P.pitchA;
    const int c = bx << P.tshA;
    cx<T>* tm = P.Tm + (size_t)by * P.tm_stride;
    LongTmDst<T, DIR> dst{&P, tm, c};
    if (DIR < 0) {
        LongRealSrc<T> src{P.signal + (size_t)by * (size_t)P.N, P.N2, c};
        fft_run<T, DIR, 1>(P.stA, P.tshA, P.pitchA, bufA, bufB, P.twA, src, dst, tid, nthr);
    } else {
        const int fi = P.f0 + by;
        LongSpecSrc<T> src{&P.sp, P.X, P.sp.rec[fi], fi, P.N2, c};
        fft_run<T, DIR, 1>(P.stA, P.tshA, P.pitchA, bufA, bufB, P.twA, src, dst, tid, nthr);
    }
}

#if defined(__CUDA_ARCH__)
// ---- 1-D bulk asynchronous copy (TMA) global -> shared, completion on an mbarrier
NW_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
NW_D void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NW_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
NW_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
NW_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

template <typename T, int DIR>
NW_HD void passB_body(const LongParams<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    const int TB = 1 << P.tshB;
    cx<T>* bufA = (cx<T>*)smem;
    cx<T>* bufB = bufA + (size_t)P.N2 * TB;
    const size_t tile_elems = (size_t)P.N2 * TB;
    const cx<T>* tile = P.Tm + (size_t)by * P.tm_stride + (size_t)bx * tile_elems;
#if defined(__CUDA_ARCH__)
    const size_t bytes = tile_elems * sizeof(cx<T>);
    if ((bytes & 15) == 0) {   // bulk copies move multiples of 16 bytes between 16-byte aligned addresses
        uint64_t* bar = (uint64_t*)(bufB + tile_elems);
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)bytes);
            const size_t CH = 32768;
            for (size_t off = 0; off < bytes; off += CH) {
                const size_t n = bytes - off < CH ? bytes - off : CH;
                bulk_g2s((char*)bufB + off, (const char*)tile + off, (uint32_t)n, bar);
            }
        }
        mbar_wait(bar, 0);
    } else {
        for (size_t i = tid; i < tile_elems; i += nthr) bufB[i] = tile[i];
        __syncthreads();
    }
#else
    for (size_t i = tid; i < tile_elems; i += nthr) bufB[i] = tile[i];
#endif
    SmemSrc<T> src{bufB, TB};
    const int r0 = bx << P.tshB;
    if (DIR < 0) {
        LongSpectrumDst<T> dst{P.Xout + (size_t)by * (size_t)P.N, P.N1, r0};
        fft_run<T, DIR, 0>(P.stB, P.tshB, TB, bufA, bufB, P.twB, src, dst, tid, nthr);
    } else {
        const size_t esz = (P.out_mode == OUT_CWT) ? sizeof(cx<T>) : sizeof(T);
        LongOutDst<T> dst{(char*)P.out + (size_t)by * (size_t)P.N * esz, P.N1, r0, P.out_mode};
        fft_run<T, DIR, 0>(P.stB, P.tshB, TB, bufA, bufB, P.twB, src, dst, tid, nthr);
    }
}

// ------------------------------------------------------------------------------
// Baseline epilogue for rows that were written by pass B (one CTA per row).
// ------------------------------------------------------------------------------
template <typename T>
NW_HD void baseline_rows_body(T* rows, long long N, int mode, int lo, int hi, double* sh, int bx, int tid, int nthr) {
    T* row = rows + (size_t)bx * (size_t)N;
#if defined(__CUDA_ARCH__)
    if (tid < 32) {
        double m, sd;
        window_stats<T>(row, lo, hi, tid, 32, &m, &sd);
        if (tid == 0) { sh[0] = m; sh[1] = sd; }
    }
#else
    if (tid == 0) window_stats<T>(row, lo, hi, 0, 1, &sh[0], &sh[1]);
#endif
    NW_SYNC();
    const T m = (T)sh[0], sd = (T)sh[1];
    for (long long i = tid; i < N; i += nthr) row[i] = baseline_apply<T>(mode, row[i], m, sd);
}

}  // namespace nw
