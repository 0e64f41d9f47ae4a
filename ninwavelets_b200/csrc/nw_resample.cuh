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
    fastdiv dRD;           // x / (R * D)   (vector kernel: x / (R * D / PQ))
    const T* coefq;        // vector kernel: weights regrouped [D / PQ][K][PQ]
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

// ---- vector kernel ----------------------------------------------------------------------------------------
// The same interpolation for even tap counts and even decimations, laid out for 16-byte (PQ = 4 outputs, D % 4 == 0) or
// 8-byte (PQ = 2, D % 2 == 0; fp64: 16 bytes) shared-memory and global accesses:
//   * a CTA of nthr threads owns C = nthr * R consecutive m of one row; every thread keeps the R + K - 1 samples its run
//     needs in registers and computes ALL D phases of its run, PQ neighbouring phases at a time (PQ K coefficients in
//     registers per phase group, loaded with warp-uniform vector loads), so one sample serves PQ outputs per register read;
//   * the CTA's samples are staged once, linearly, with one pad slot per R samples: staging stores and the per-thread
//     window loads (lane stride R + 1 slots) are both bank-conflict free;
//   * results go to a per-thread row of the staging tile (pitch R D + PQ words: conflict free for the PQ-wide stores) and
//     leave as the CTA's one contiguous piece of the output row in PQ-wide streaming stores, a warp covering
//     32 PQ consecutive samples per instruction.
// Per output: K packed FMAs, one |z|^2, and 3 / PQ memory instructions.
template <typename T, int PQ> struct RsVec { T v[PQ]; };
#if defined(__CUDA_ARCH__)
NW_D void rs_st_shared(float* p, const RsVec<float, 4>& r) { *(float4*)p = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]); }
NW_D void rs_st_shared(float* p, const RsVec<float, 2>& r) { *(float2*)p = make_float2(r.v[0], r.v[1]); }
NW_D void rs_st_shared(double* p, const RsVec<double, 2>& r) { *(double2*)p = make_double2(r.v[0], r.v[1]); }
NW_D void rs_st_shared(double* p, const RsVec<double, 4>& r) {
    *(double2*)p = make_double2(r.v[0], r.v[1]);
    *(double2*)(p + 2) = make_double2(r.v[2], r.v[3]);
}
NW_D void rs_copy_out(float* g, const float* s, RsVec<float, 4>*) { __stcs((float4*)g, *(const float4*)s); }
NW_D void rs_copy_out(float* g, const float* s, RsVec<float, 2>*) { __stcs((float2*)g, *(const float2*)s); }
NW_D void rs_copy_out(double* g, const double* s, RsVec<double, 2>*) { __stcs((double2*)g, *(const double2*)s); }
NW_D void rs_copy_out(double* g, const double* s, RsVec<double, 4>*) {
    __stcs((double2*)g, *(const double2*)s);
    __stcs((double2*)(g + 2), *(const double2*)(s + 2));
}
template <typename T, int PQ> NW_D RsVec<T, PQ> rs_ld_coef(const T* p);
template <> NW_D RsVec<float, 4> rs_ld_coef<float, 4>(const float* p) {
    const float4 v = __ldg((const float4*)p);
    RsVec<float, 4> r;
    r.v[0] = v.x; r.v[1] = v.y; r.v[2] = v.z; r.v[3] = v.w;
    return r;
}
template <> NW_D RsVec<float, 2> rs_ld_coef<float, 2>(const float* p) {
    const float2 v = __ldg((const float2*)p);
    RsVec<float, 2> r;
    r.v[0] = v.x; r.v[1] = v.y;
    return r;
}
template <> NW_D RsVec<double, 2> rs_ld_coef<double, 2>(const double* p) {
    const double2 v = __ldg((const double2*)p);
    RsVec<double, 2> r;
    r.v[0] = v.x; r.v[1] = v.y;
    return r;
}
template <> NW_D RsVec<double, 4> rs_ld_coef<double, 4>(const double* p) {
    const double2 a = __ldg((const double2*)p), b = __ldg((const double2*)p + 1);
    RsVec<double, 4> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    return r;
}
#else
template <typename T, int PQ> inline RsVec<T, PQ> rs_ld_coef(const T* p) { RsVec<T, PQ> r; for (int j = 0; j < PQ; ++j) r.v[j] = p[j]; return r; }
template <typename T, int PQ> inline void rs_st_shared(T* p, const RsVec<T, PQ>& r) { for (int j = 0; j < PQ; ++j) p[j] = r.v[j]; }
template <typename T, int PQ> inline void rs_copy_out(T* g, const T* s, RsVec<T, PQ>*) { for (int j = 0; j < PQ; ++j) g[j] = s[j]; }
#endif

template <typename T, int K, int R, int PQ> struct ResampleVecGeo {
    static const int WN = R + K - 1;                         // window samples per run (K even: one tap offset for all phases)
    static NW_HD int pitch(int D) { return R * D + PQ; }     // staging-tile words per thread
    static NW_HD size_t smem_bytes(int nthr, int D) {
        const size_t C = (size_t)nthr * R;
        const size_t ys = (C + K + (C + K) / R + 2) * sizeof(cx<T>);
        const size_t tile = (size_t)nthr * pitch(D) * sizeof(T);
        return ys > tile ? ys : tile;
    }
};

template <typename T, int K, int R, int PQ, int MODE>
NW_HD void resample_vec_body(const ResampleParams<T>& P, char* smem, int bx, int by, int tid, int nthr) {
    typedef ResampleVecGeo<T, K, R, PQ> G;
    static_assert((K & 1) == 0 && (R & (R - 1)) == 0, "even taps, power-of-two runs");
    const int D = P.D, M = P.M;
    const int C = nthr * R;
    const long long m0 = (long long)bx * C;
    const int gr = P.row0 + by, si = gr / P.F, fi = gr - si * P.F;
    const size_t orow = (size_t)si * (size_t)P.F_out + (size_t)(P.fmap ? P.fmap[fi] : fi);
    T* out = (T*)P.out + orow * (size_t)P.N;
    const cx<T>* y = P.y + (size_t)by * (size_t)P.ystride;
    cx<T>* ys = (cx<T>*)smem;
    T* tile = (T*)smem;
    // stage samples m0 + t0 + i, i < C + K - 1 (t0 = 1 - K / 2, indices mod M) at slot i + i / R
    {
        long long mi = (m0 + (1 - K / 2) + tid) % M;
        if (mi < 0) mi += M;
        const int adv = nthr % M;
        for (int i = tid; i < C + K - 1; i += nthr) {
            ys[i + i / R] = y[mi];
            mi += adv;
            if (mi >= M) mi -= M;
        }
    }
    NW_SYNC();
    cx<T> w[G::WN];
    {
        const cx<T>* yb = ys + (size_t)tid * (R + 1);
#pragma unroll
        for (int o = 0; o < G::WN; ++o) w[o] = yb[o + o / R];
    }
    NW_SYNC();   // every window is in registers: the tile may overwrite the staged samples
    const int RS = G::pitch(D);
    T* dst = tile + (size_t)tid * RS;
    const T* cq = P.coefq;
    for (int p0 = 0; p0 < D; p0 += PQ, cq += K * PQ) {
        RsVec<T, PQ> c[K];
#pragma unroll
        for (int t = 0; t < K; ++t) c[t] = rs_ld_coef<T, PQ>(cq + t * PQ);
#pragma unroll
        for (int mm = 0; mm < R; ++mm) {
            RsVec<T, PQ> r;
#pragma unroll
            for (int j = 0; j < PQ; ++j) {
                // one accumulator per output: the PQ R outputs of a phase group are independent chains
                cx<T> a = mk<T>((T)0, (T)0);
#pragma unroll
                for (int t = 0; t < K; ++t) rs_fma(a, w[mm + t], c[t].v[j]);
                r.v[j] = real_out<T>(MODE, a);
            }
            rs_st_shared(dst + mm * D + p0, r);
        }
    }
    NW_SYNC();
    // the CTA's outputs n = m0 D + PQ v, v < C D / PQ, are one contiguous piece of the row; thread row jj holds R D of them
    const uint32_t total = (uint32_t)C * (uint32_t)D / PQ;
    const uint32_t rowv = (uint32_t)R * (uint32_t)D / PQ;
    const long long n0 = m0 * D;
    for (uint32_t v = tid; v < total; v += nthr) {
        const long long n = n0 + (long long)v * PQ;
        if (n >= P.N) break;
        const uint32_t jj = fd_div(v, P.dRD);
        rs_copy_out(out + n, tile + (size_t)jj * RS + (size_t)(v - jj * rowv) * PQ, (RsVec<T, PQ>*)nullptr);
    }
}

// run length of the vector kernel by decimation: R D stays near 64 outputs per thread (tile <= 33 KB per 128 threads, so
// that six CTAs share an SM), window overhead (R + K - 1) / R is paid once per D outputs
inline int resample_vec_run(int D) { return D <= 4 ? 16 : D <= 8 ? 8 : D <= 16 ? 4 : 2; }

// host-side launch geometry of the vector kernel for a group (D, K) at decimated length M; false: not eligible
// (odd D or K, or a tap count / precision the vector kernels are not compiled for - the scalar kernel takes those)
struct ResampleVecShape { int R, PQ, nthr, C; unsigned tiles; size_t smem; fastdiv dRD; };
template <typename T> inline bool resample_vec_shape(int D, int K, long long M, ResampleVecShape& v) {
    if (sizeof(T) != 4 || (D & 1) || (K & 1) || K < 4 || K > 12) return false;
    v.R = resample_vec_run(D);
    v.PQ = (D & 3) ? 2 : 4;
    v.nthr = 128;
    v.C = v.nthr * v.R;
    v.tiles = (unsigned)((M + v.C - 1) / v.C);
    const size_t ys = ((size_t)v.C + K + ((size_t)v.C + K) / v.R + 2) * 2 * sizeof(T);
    const size_t tile = (size_t)v.nthr * ((size_t)v.R * D + v.PQ) * sizeof(T);
    v.smem = ys > tile ? ys : tile;
    v.dRD = make_fastdiv((uint32_t)(v.R * D / v.PQ));
    return true;
}
// weights regrouped for the vector kernel: coefq[(p / PQ) K + t][p % PQ] = coef[p][t]
template <typename T> inline void resample_coefq(const double* coef, int D, int K, int PQ, T* q) {
    for (int p = 0; p < D; ++p)
        for (int t = 0; t < K; ++t) q[((size_t)(p / PQ) * K + t) * PQ + (p % PQ)] = (T)coef[(size_t)p * K + t];
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
