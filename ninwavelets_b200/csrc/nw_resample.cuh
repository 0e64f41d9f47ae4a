// Interpolation kernel of the resampled rows (nw_plan.h: MrGroup; DESIGN.md "resampled rows").
//
// A group's rows arrive as M = N / D complex samples  y[m] = z(m D) e^{-2 pi i kc m / M}  (the M-point inverse
// transform of the band moved to bin 0, pass band pre-divided by the kernel's response) and leave as the N real
// outputs of base.py:443 / :425:
//     out[m D + p] = | sum_{t < K} coef[p][t] * y[(m + t0[p] + t) mod M] |  (^2 for power)
// i.e. a polyphase FIR with D phases of K taps.  The modulation e^{2 pi i kc n / N} that separates y from z has unit
// modulus and drops out of |z|.
//
// Work split.  A CTA owns C = 32 * WR * R consecutive m of one row: WR "run-warps", each lane of which owns a run of
// R consecutive m and keeps the R + K samples its outputs need IN REGISTERS for the whole kernel; WP "phase-warps"
// per run-warp deal the D phases between them.  The inner loop is K packed FMAs (re, im in one FFMA2) per output with
// compile-time register indices - no shared-memory traffic except K / 2 broadcast coefficient loads per phase and
// one 4-byte store per output into the staging tile, from which the CTA's C * D outputs (one contiguous piece of the
// row) go to global memory in fully coalesced streaming stores.
//   shared memory:  ys[WR][R + K][32]   staged samples, transposed so that a warp's window loads are conflict free;
//                   tile[32 WR][RS]     results (aliases ys: all windows are in registers before the first store);
//                                       RS = R D (+1 if even) keeps the per-lane stores conflict free.
#pragma once
#include "nw_common.h"
#include "nw_pk.cuh"
#include "nw_kernels.cuh"

namespace nw {

template <typename T>
struct ResampleParams {
    const cx<T>* y;        // [rows of this launch][ystride]
    long long ystride;
    void* out;             // output base, real T [S][F_out][N]
    long long N;           // M * D
    int M, D;
    const T* coef;         // [D][K]
    const int* t0;         // [D] first tap offset of each phase
    int t0min;             // min over phases
    const int* fmap;       // group row fi -> plan frequency index
    int F, F_out;
    int row0;              // first group row (signal-major: row = signal * F + fi) of this launch
    int WR, WP;            // run-warps and phase-warps per run-warp: blockDim = 32 * WR * WP
    int RS;                // staging-tile pitch per run
    fastdiv dRD;           // x / (R * D)
};

template <typename T, int K, int R> struct ResampleGeo {
    static const int WN = R + K - 1 + (K & 1);   // window samples per run: odd K has phases whose taps start one sample later
    static NW_HD size_t smem_bytes(int WR, int D) {
        const int RS = (R * D) | 1;
        const size_t ys = (size_t)WR * (R + K) * 32 * sizeof(cx<T>);
        const size_t tile = (size_t)WR * 32 * RS * sizeof(T);
        return ys > tile ? ys : tile;
    }
};

// acc += w * c  (complex sample times real weight)
#if defined(__CUDA_ARCH__)
NW_D void rs_fma(cx<float>& acc, cx<float> w, float c) {
    const float2 r = __ffma2_rn(make_float2(w.x, w.y), make_float2(c, c), make_float2(acc.x, acc.y));
    acc.x = r.x;
    acc.y = r.y;
}
NW_D void rs_fma(cx<double>& acc, cx<double> w, double c) {
    acc.x = fma(w.x, c, acc.x);
    acc.y = fma(w.y, c, acc.y);
}
#else
template <typename T> inline void rs_fma(cx<T>& acc, cx<T> w, T c) {
    acc.x += w.x * c;
    acc.y += w.y * c;
}
#endif

// one phase: R outputs from the register window, taps starting DT samples into it
template <typename T, int K, int R, int MODE, int DT>
NW_HD void resample_phase(const cx<T>* w, const T* c, T* dst, int D) {
#pragma unroll
    for (int mm = 0; mm < R; ++mm) {
        cx<T> a0 = mk<T>((T)0, (T)0), a1 = a0;
#pragma unroll
        for (int t = 0; t + 1 < K; t += 2) {
            rs_fma(a0, w[mm + DT + t], c[t]);
            rs_fma(a1, w[mm + DT + t + 1], c[t + 1]);
        }
        if (K & 1) rs_fma(a0, w[mm + DT + K - 1], c[K - 1]);
        dst[mm * D] = real_out<T>(MODE, a0 + a1);
    }
}

template <typename T, int K, int R, int MODE>
NW_HD void resample_body(const ResampleParams<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    typedef ResampleGeo<T, K, R> G;
    const int D = P.D, M = P.M, WR = P.WR, WP = P.WP;
    const int C = 32 * WR * R;
    const long long m0 = (long long)bx * C;
    const int gr = P.row0 + by, si = gr / P.F, fi = gr - si * P.F;
    const size_t orow = (size_t)si * (size_t)P.F_out + (size_t)(P.fmap ? P.fmap[fi] : fi);
    T* out = (T*)P.out + orow * (size_t)P.N;
    const cx<T>* y = P.y + (size_t)by * (size_t)P.ystride;
    cx<T>* ys = (cx<T>*)smem;
    T* tile = (T*)smem;
    // stage the C + K samples the CTA's runs need; run jj holds samples [jj R, jj R + R + K) of the span
    {
        long long mi = (m0 + P.t0min + tid) % M;
        if (mi < 0) mi += M;
        const int adv = nthr % M;
        for (int i = tid; i < C + K; i += nthr) {
            const cx<T> v = y[mi];
            mi += adv;
            if (mi >= M) mi -= M;
            int jj = i / R, o = i - jj * R;
            for (; o < R + K && jj >= 0; --jj, o += R)
                if (jj < 32 * WR) ys[((size_t)(jj >> 5) * (R + K) + o) * 32 + (jj & 31)] = v;
        }
    }
    NW_SYNC();
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = warp / WP, wp = warp - wr * WP;
    cx<T> w[G::WN];
#pragma unroll
    for (int o = 0; o < G::WN; ++o) w[o] = ys[((size_t)wr * (R + K) + o) * 32 + lane];
    NW_SYNC();   // every window is in registers: the tile may overwrite the staged samples
    T* dst = tile + (size_t)(wr * 32 + lane) * P.RS;
    for (int p = wp; p < D; p += WP) {
        T c[K];
#pragma unroll
        for (int t = 0; t < K; ++t) c[t] = P.coef[p * K + t];
        if ((K & 1) && P.t0[p] != P.t0min) resample_phase<T, K, R, MODE, (K & 1)>(w, c, dst + p, D);
        else resample_phase<T, K, R, MODE, 0>(w, c, dst + p, D);
    }
    NW_SYNC();
    // the CTA's outputs n = m0 D + idx, idx < C D, are one contiguous piece of the row
    const uint32_t total = (uint32_t)C * (uint32_t)D;
    const uint32_t RD = (uint32_t)R * (uint32_t)D;
    const long long n0 = m0 * D;
    for (uint32_t idx = tid; idx < total; idx += nthr) {
        const long long n = n0 + idx;
        if (n >= P.N) break;
        const uint32_t jj = fd_div(idx, P.dRD);
        st_stream(out + n, tile[(size_t)jj * P.RS + (idx - jj * RD)]);
    }
}

// run length per lane (compile-time: the window lives in registers)
template <typename T> struct ResampleRun { static const int R = 12; };
template <> struct ResampleRun<double> { static const int R = 8; };

// host-side launch geometry of a group: the staging tile (32 WR R D values) stays near 48 KB; the phase-warps divide
// D evenly where they can, up to 12 warps per CTA
struct ResampleShape { int WR, WP, RS, C; size_t smem; };
template <typename T> inline ResampleShape resample_shape(int D, int K) {
    const int R = ResampleRun<T>::R;
    int wrmax = 8;
    while (wrmax > 1 && (size_t)wrmax * 32 * R * D * sizeof(T) > 49152) wrmax >>= 1;
    double best = -1;
    ResampleShape s{1, 1, 0, 0, 0};
    for (int wr = wrmax; wr >= 1; wr >>= 1)
        for (int wp = 1; wp <= D && wr * wp <= 12; ++wp) {
            const double eff = (double)D / (double)(((D + wp - 1) / wp) * wp);
            const double score = eff * (double)(wr * wp < 8 ? wr * wp : 8) / 8.0;
            if (score > best + 1e-9) { best = score; s.WR = wr; s.WP = wp; }
        }
    s.RS = (R * D) | 1;
    s.C = 32 * s.WR * R;
    const size_t ys = (size_t)s.WR * (R + K) * 32 * sizeof(cx<T>), tile = (size_t)s.WR * 32 * s.RS * sizeof(T);
    s.smem = ys > tile ? ys : tile;
    return s;
}

}  // namespace nw
