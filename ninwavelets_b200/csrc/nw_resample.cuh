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
#include <string.h>
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
    int RS;                // staging-tile pitch per run   (vector kernel: bytes of a run-warp group's shared-memory region)
    fastdiv dRD;           // x / (R * D)   (vector kernel: x / (R * D / PQ))
    const T* coefq;        // vector kernel: weights regrouped [D / PQ][K][PQ]
    fastdiv dGT;           // vector kernel: x / (work items per row = ceil(M / (32 R)))
    int nrows;             // vector kernel: rows of this launch
    int G, MW;             // direct kernel: lanes per output sample m (phase groups D / PQ), samples m a warp handles side by side
    fastdiv dG, dF;        // direct kernel: lane / G, row / F
    int NSUB;              // direct kernel: sub-items (MW runs of R samples m side by side) per item
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
// 8-byte (PQ = 2) shared-memory and global accesses:
//   * a CTA of 128 threads = WR run-warp groups x WP phase-warps owns C = 32 WR R consecutive m of one row; the groups
//     are independent of each other (own barrier), so their load / arithmetic / store phases interleave.  A lane of a
//     run-warp owns a run of R consecutive m and keeps the R + K - 1 samples it needs in registers; the WP phase-warps of
//     a run-warp share the D / PQ phase groups (PQ neighbouring phases) between them, so one sample register serves
//     PQ D / WP outputs and the PQ K weights of a phase group (warp-uniform vector loads) serve 32 R PQ outputs;
//   * outputs are accumulated as PAIRS of neighbouring phases: (re, re') and (im, im') in one packed register each,
//     FFMA2(sample broadcast, weight pair, acc) - K packed FMAs and one packed |z|^2 per output;
//   * the CTA's samples are staged once, linearly, with one pad slot per R samples: staging stores and the per-lane
//     window loads (lane stride R + 1 slots) are both bank-conflict free;
//   * results go to a per-run row of the staging tile (pitch R D + PQ words: conflict free for the PQ-wide stores) and
//     leave as the CTA's one contiguous piece of the output row, a warp copying one run's R D samples at a time in
//     PQ-wide streaming stores (512 contiguous bytes per instruction).
template <typename T, int PQ> struct RsVec { T v[PQ]; };
#if defined(__CUDA_ARCH__)
NW_D void rs_st_shared(float* p, const RsVec<float, 4>& r) { *(float4*)p = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]); }
NW_D void rs_st_shared(float* p, const RsVec<float, 2>& r) { *(float2*)p = make_float2(r.v[0], r.v[1]); }
NW_D void rs_st_shared(double* p, const RsVec<double, 2>& r) { *(double2*)p = make_double2(r.v[0], r.v[1]); }
NW_D void rs_copy_out(float* g, const float* s, RsVec<float, 4>*) { __stcs((float4*)g, *(const float4*)s); }
NW_D void rs_copy_out(float* g, const float* s, RsVec<float, 2>*) { __stcs((float2*)g, *(const float2*)s); }
NW_D void rs_copy_out(double* g, const double* s, RsVec<double, 2>*) { __stcs((double2*)g, *(const double2*)s); }
NW_D cx<float> rs_ld_sample(const cx<float>* p) { const float2 v = __ldg((const float2*)p); return mk<float>(v.x, v.y); }
NW_D cx<double> rs_ld_sample(const cx<double>* p) { const double2 v = __ldg((const double2*)p); return mk<double>(v.x, v.y); }
NW_D void rs_st_out(float* g, const RsVec<float, 4>& r) { __stcs((float4*)g, make_float4(r.v[0], r.v[1], r.v[2], r.v[3])); }
NW_D void rs_st_out(float* g, const RsVec<float, 2>& r) { __stcs((float2*)g, make_float2(r.v[0], r.v[1])); }
NW_D void rs_st_out(double* g, const RsVec<double, 2>& r) { __stcs((double2*)g, make_double2(r.v[0], r.v[1])); }
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
// two neighbouring phases of one output sample m: acc over the K taps, then |z|^2 (or |z|) of both
template <int K, int MODE>
NW_D void rs_pair(const cx<float>* w, const float* c0, int cstride, float& o0, float& o1) {
    float2 are = make_float2(0.f, 0.f), aim = are;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        const float2 cp = make_float2(c0[t * cstride], c0[t * cstride + 1]);
        are = __ffma2_rn(make_float2(w[t].x, w[t].x), cp, are);
        aim = __ffma2_rn(make_float2(w[t].y, w[t].y), cp, aim);
    }
    if (MODE == OUT_POWER) {
        const float2 p = __ffma2_rn(aim, aim, __fmul2_rn(are, are));
        o0 = p.x;
        o1 = p.y;
    } else {
        o0 = nw_hypot(are.x, aim.x);
        o1 = nw_hypot(are.y, aim.y);
    }
}
template <int K, int MODE>
NW_D void rs_pair(const cx<double>* w, const double* c0, int cstride, double& o0, double& o1) {
    double r0 = 0, i0 = 0, r1 = 0, i1 = 0;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        const double ca = c0[t * cstride], cb = c0[t * cstride + 1];
        r0 = fma(w[t].x, ca, r0); i0 = fma(w[t].y, ca, i0);
        r1 = fma(w[t].x, cb, r1); i1 = fma(w[t].y, cb, i1);
    }
    o0 = real_out<double>(MODE, mk<double>(r0, i0));
    o1 = real_out<double>(MODE, mk<double>(r1, i1));
}
#else
template <typename T, int PQ> inline RsVec<T, PQ> rs_ld_coef(const T* p) { RsVec<T, PQ> r; for (int j = 0; j < PQ; ++j) r.v[j] = p[j]; return r; }
template <typename T, int PQ> inline void rs_st_shared(T* p, const RsVec<T, PQ>& r) { for (int j = 0; j < PQ; ++j) p[j] = r.v[j]; }
template <typename T, int PQ> inline void rs_copy_out(T* g, const T* s, RsVec<T, PQ>*) { for (int j = 0; j < PQ; ++j) g[j] = s[j]; }
template <typename T> inline cx<T> rs_ld_sample(const cx<T>* p) { return *p; }
template <typename T, int PQ> inline void rs_st_out(T* g, const RsVec<T, PQ>& r) { for (int j = 0; j < PQ; ++j) g[j] = r.v[j]; }
template <int K, int MODE, typename T>
inline void rs_pair(const cx<T>* w, const T* c0, int cstride, T& o0, T& o1) {
    T r0 = 0, i0 = 0, r1 = 0, i1 = 0;
    for (int t = 0; t < K; ++t) {
        const T ca = c0[t * cstride], cb = c0[t * cstride + 1];
        r0 += w[t].x * ca; i0 += w[t].y * ca;
        r1 += w[t].x * cb; i1 += w[t].y * cb;
    }
    o0 = real_out<T>(MODE, mk<T>(r0, i0));
    o1 = real_out<T>(MODE, mk<T>(r1, i1));
}
#endif

// barrier among the 32 WP threads of one run-warp group.  (WR, WP) is (4, 1), (2, 2) or (1, 4): a warp-level barrier, the
// hardware barriers 1 and 2 with 64 threads each, or the CTA barrier - compile-time barrier ids, so that a CTA reserves
// three hardware barriers and not all sixteen.  The host emulation steps whole CTAs.
#if defined(__CUDA_ARCH__)
NW_D void rs_group_sync(int wr, int WP, int WR) {
    if (WP == 1) __syncwarp();
    else if (WR == 1) __syncthreads();
    else if (wr == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else asm volatile("bar.sync 2, 64;" ::: "memory");
}
#else
inline void rs_group_sync(int, int, int) { NW_SYNC(); }
#endif

template <typename T, int K, int R, int PQ, int MODE>
NW_HD void resample_vec_body(const ResampleParams<T>& P, char* smem, int bx, int nbx, int tid, int nthr) {
    static_assert((K & 1) == 0 && (R & (R - 1)) == 0 && (PQ == 2 || PQ == 4), "even taps, power-of-two runs");
    const int D = P.D, M = P.M, WP = P.WP, WR = P.WR;
    const int warp = tid >> 5, lane = tid & 31;
    const int wr = warp / WP, wp = warp - wr * WP;
    // Every run-warp group (its WP warps) is independent of the others: own work items, own samples, own staging tile,
    // own barrier - the groups of an SM drift apart, so that one group's stores overlap another's arithmetic.  A group's
    // work items are (row, piece of CG = 32 R consecutive m); the CTAs are persistent and stride over the items, and the
    // samples of the NEXT item are fetched into registers while the current one is computed.
    const int gthr = 32 * WP, gtid = wp * 32 + lane;
    const int CG = 32 * R, NS = CG + K - 1;                             // m per item, samples an item needs
    constexpr int NPF = R + 1;                                          // >= ceil(NS / 32): prefetch registers per thread
    const uint32_t GT = P.dGT.d;                                        // items per row
    const uint32_t total = GT * (uint32_t)P.nrows;
    char* gsm = smem + (size_t)wr * (size_t)P.RS;                       // RS: bytes of a group's region
    cx<T>* ys = (cx<T>*)gsm;
    T* tile = (T*)gsm;
    const int pitch = R * D + PQ;
    const int G = D / PQ;
    const int rowv = R * D / PQ;
    cx<T> pf[NPF];
    // samples mg0 + t0 + i, i < NS (t0 = 1 - K / 2, indices mod M) of item `it`
    auto fetch = [&](uint32_t it) {
        const uint32_t by = it / GT, gt = it - by * GT;
        const cx<T>* y = P.y + (size_t)by * (size_t)P.ystride;
        long long base = (long long)gt * CG + (1 - K / 2);
        if (base < 0) base += M;
#pragma unroll
        for (int q = 0; q < NPF; ++q) {
            const int i = gtid + q * gthr;
            if (i < NS) {
                long long mi = base + i;
                if (NS <= M) { if (mi >= M) mi -= M; } else mi %= M;
                pf[q] = y[mi];
            }
        }
    };
    uint32_t cit = (uint32_t)bx;
    uint32_t it = cit * (uint32_t)WR + (uint32_t)wr;
    if (it < total) fetch(it);
    for (; cit * (uint32_t)WR < total; cit += (uint32_t)nbx) {
        it = cit * (uint32_t)WR + (uint32_t)wr;
        const bool valid = it < total;
        // stage the item's samples at slot i + i / R
#pragma unroll
        for (int q = 0; q < NPF; ++q) {
            const int i = gtid + q * gthr;
            if (valid && i < NS) ys[i + i / R] = pf[q];
        }
        rs_group_sync(wr, WP, WR);
        cx<T> w[R + K - 1];
        {
            const cx<T>* yb = ys + (size_t)lane * (R + 1);
#pragma unroll
            for (int o = 0; o < R + K - 1; ++o) w[o] = yb[o + o / R];
        }
        rs_group_sync(wr, WP, WR);   // every window is in registers: the tile may overwrite the staged samples
        {
            const uint32_t nit = (cit + (uint32_t)nbx) * (uint32_t)WR + (uint32_t)wr;
            if (nit < total) fetch(nit);   // in flight during the arithmetic below
        }
        T* dst = tile + (size_t)lane * pitch;
        for (int g = wp; g < G; g += WP) {
            const T* cq = P.coefq + (size_t)g * (K * PQ);
            RsVec<T, PQ> c[K];
#pragma unroll
            for (int t = 0; t < K; ++t) c[t] = rs_ld_coef<T, PQ>(cq + t * PQ);
#pragma unroll
            for (int mm = 0; mm < R; ++mm) {
                RsVec<T, PQ> r;
#ifdef RS_EXP_NOFMA
#pragma unroll
                for (int j = 0; j < PQ; ++j) r.v[j] = w[mm + (j & 1)].x * c[j & 1].v[j];
#else
#pragma unroll
                for (int j = 0; j < PQ; j += 2) rs_pair<K, MODE>(w + mm, &c[0].v[j], PQ, r.v[j], r.v[j + 1]);
#endif
                rs_st_shared(dst + mm * D + g * PQ, r);
            }
        }
        rs_group_sync(wr, WP, WR);
        // the item's outputs are one contiguous piece of the row starting at n0 = mg0 D; run jj holds R D of them
        if (valid) {
            const uint32_t by = it / GT, gt = it - by * GT;
            const int gr = P.row0 + (int)by, si = gr / P.F, fi = gr - si * P.F;
            const size_t orow = (size_t)si * (size_t)P.F_out + (size_t)(P.fmap ? P.fmap[fi] : fi);
            const long long n0 = (long long)gt * CG * D;
            T* o0 = (T*)P.out + orow * (size_t)P.N + n0;
            const uint32_t nv = (uint32_t)CG * (uint32_t)D / PQ;          // vectors of the item
            const long long left = (P.N - n0) / PQ;                          // vectors up to the end of the row
            const uint32_t lim = left < (long long)nv ? (uint32_t)left : nv;
#ifdef RS_EXP_NOST
            if (tile[gtid] == (T)123.456) o0[gtid] = tile[gtid + 1];
            if (true) {} else
#endif
            if (lim == nv) {
#pragma unroll 8
                for (uint32_t v = gtid; v < nv; v += gthr) {
                    const uint32_t jj = fd_div(v, P.dRD);
                    rs_copy_out(o0 + (size_t)v * PQ, tile + (size_t)jj * pitch + (size_t)(v - jj * rowv) * PQ, (RsVec<T, PQ>*)nullptr);
                }
            } else {   // the row ends inside this item (N is a multiple of D, hence of PQ)
                for (uint32_t v = gtid; v < lim; v += gthr) {
                    const uint32_t jj = fd_div(v, P.dRD);
                    rs_copy_out(o0 + (size_t)v * PQ, tile + (size_t)jj * pitch + (size_t)(v - jj * rowv) * PQ, (RsVec<T, PQ>*)nullptr);
                }
            }
        }
        rs_group_sync(wr, WP, WR);   // the tile is read: the next item's samples may overwrite it
    }
}

// ---- direct kernel ------------------------------------------------------------------------------------------
// The interpolation with its results leaving straight from registers in stores that are already coalesced: no output
// staging, no CTA barrier.
//   * lane = j G + g: phase group g (PQ neighbouring phases, G = D / PQ groups) of the run j of R consecutive m; the
//     MW = 32 / G runs of a warp are consecutive, so a warp-item is MW R consecutive m of one row = MW R D outputs.
//     A lane's phase group never changes: its PQ K weights are loaded ONCE per kernel and stay in registers.
//   * the MW R + K - 1 samples of an item - one contiguous piece of the decimated row - are copied by the warp into its
//     own shared-memory strip with 16-byte asynchronous copies (LDGSTS), double buffered: the copy of the next item is
//     in flight during the arithmetic of the current one.  R is ODD, so the per-lane window loads (lane stride R
//     samples = 8 R bytes) are bank-conflict free on the linear strip.  (Per-lane loads of the windows straight from
//     global memory cost one L1 sector access per run and sample and bound the first version of this kernel at 83 % L1
//     throughput; staging through registers and a padded strip doubled its instruction count: profiles/r02.)
//   * the inner loop is the K packed FMAs per output pair of resample_vec_body (rs_pair);
//   * the PQ outputs of (m, g) are one 16-byte (8-byte) streaming store; the G lanes of a run cover the D consecutive
//     outputs of m, so every store instruction of a warp writes MW whole pieces of 4 D bytes, and the R stores of a
//     lane fill its run's R D outputs back to back.
//   * warps are independent (persistent, striding over the items; one warp-level barrier per item); item -> (row,
//     piece) is carried incrementally, without divisions.
template <int K, int R> struct RsDirGeo {
    static const int SLOTS = 16 * R * 4 + K + 2 + (K & 1);  // strip of a warp, in samples: 4 sub-items of 16 runs (or 2 of 32), one alignment sample each side
    static const int DEPTH = 3;                             // strips per warp: the copies run DEPTH - 1 items ahead
};
#if defined(__CUDA_ARCH__)
#define NW_WARP_SYNC() __syncwarp()
NW_D void rs_cp16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
NW_D void rs_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> NW_D void rs_cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
#else
#define NW_WARP_SYNC() NW_SYNC()
inline void rs_cp16(void* dst, const void* src) { memcpy(dst, src, 16); }
inline void rs_cp_commit() {}
template <int N> inline void rs_cp_wait() {}
#endif
template <typename T, int K, int R, int PQ, int MODE>
NW_HD void resample_dir_body(const ResampleParams<T>& P, char* smem, int bx, int nbx, int tid, int nthr) {
    static_assert((K & 1) == 0 && (R & 1) == 1 && (PQ == 2 || PQ == 4), "even taps, odd runs");
    typedef RsDirGeo<K, R> GEO;
    const int M = P.M, D = P.D, G = P.G, MW = P.MW;
    const int lane = tid & 31, warp = tid >> 5, wpc = nthr >> 5;
    const int j = (int)fd_div((uint32_t)lane, P.dG), g = lane - j * G;
    const bool act = j < MW;                                // 32 - MW G idle lanes when G does not divide 32
    cx<T>* strip = (cx<T>*)smem + (size_t)warp * (GEO::DEPTH * GEO::SLOTS);
    RsVec<T, PQ> c[K];
    {
        const T* cq = P.coefq + (size_t)(act ? g : 0) * (K * PQ);
#pragma unroll
        for (int t = 0; t < K; ++t) c[t] = rs_ld_coef<T, PQ>(cq + t * PQ);
    }
    const int GT = (int)P.dGT.d;                            // items per row
    const int CW = MW * R;                                  // m per sub-item: the runs of the warp's lanes side by side
    const int NSUB = P.NSUB, CWI = NSUB * CW, NS = CWI + K - 1;   // sub-items, m and samples per item
    // every warp owns a CONTIGUOUS range of items (consecutive pieces of consecutive rows): pointers are set up once per
    // row and advance by constants - the item loop carries no divisions, no table look-ups and no 64-bit multiplies
    const uint32_t total = (uint32_t)GT * (uint32_t)P.nrows;
    const uint32_t nwarps = (uint32_t)nbx * (uint32_t)wpc;
    const uint32_t per = (total + nwarps - 1) / nwarps;     // trip count, uniform over the grid (host emulation: CTA barriers)
    const uint32_t i0 = ((uint32_t)bx * (uint32_t)wpc + (uint32_t)warp) * per;
    const int cnt = i0 >= total ? 0 : (int)(total - i0 < per ? total - i0 : per);   // items of this warp
    // an item whose samples lie inside the row, with one spare sample either side, is copied in aligned 16-byte pieces;
    // the first and the last one or two items of a row are copied sample by sample, wrapped
    const int gt_fast = (M - NS - K / 2) / CWI;             // gt <= gt_fast: gt CWI + 1 - K / 2 + NS + 1 <= M
    const int by0 = (int)(i0 / (uint32_t)GT), gt0 = (int)(i0 - (uint32_t)by0 * (uint32_t)GT);
    // fetch cursor: DEPTH - 1 items ahead of the consume cursor
    int pgt = gt0, pleft = cnt;
    const cx<T>* psrc = P.y + (size_t)by0 * (size_t)P.ystride + ((long long)gt0 * CWI + 1 - K / 2);   // sample b of the item
    cx<T>* const ring0 = strip;
    cx<T>* const ring1 = strip + GEO::DEPTH * GEO::SLOTS;
    cx<T>* pdst = ring0;
    auto issue = [&]() {
        if (pleft > 0) {
#ifdef RS_EXP_NOLOAD   // timing experiment: no sample loads
            if (pgt == -12345) {
#endif
            if (pgt >= 1 && pgt <= gt_fast) {
                const int a = (int)(((uintptr_t)psrc / sizeof(cx<T>)) & 1);
                const cx<T>* src = psrc - a + 2 * lane;
                cx<T>* dst = pdst + 2 * lane;
                const int n16 = (NS + a + 1) >> 1;
#pragma unroll 4
                for (int u = lane; u < n16; u += 32) rs_cp16(dst + 2 * (u - lane), src + 2 * (u - lane));
            } else {
                const cx<T>* y = psrc - ((long long)pgt * CWI + 1 - K / 2);   // row start
                const int b = pgt * CWI + 1 - K / 2;
                for (int i = lane; i < NS; i += 32) {
                    int mi = b + i;
                    if (mi < 0) mi += M;
                    if (mi >= M) mi -= M;                   // M >= NS (resample_dir_shape)
                    pdst[i] = y[mi];
                }
            }
#ifdef RS_EXP_NOLOAD
            }
#endif
        }
        rs_cp_commit();
        --pleft;
        psrc += CWI;
        if (++pgt == GT) { pgt = 0; psrc += P.ystride - (long long)GT * CWI; }
        pdst += GEO::SLOTS;
        if (pdst == ring1) pdst = ring0;
    };
#pragma unroll 1
    for (int d = 0; d < GEO::DEPTH - 1; ++d) issue();
    // consume cursor
    int gt = gt0, by = by0;
    const cx<T>* csrc = P.y + (size_t)by0 * (size_t)P.ystride + ((long long)gt0 * CWI + 1 - K / 2);
    const cx<T>* cstrip = ring0 + (act ? j : 0) * R;
    const int lane_off = (act ? j : 0) * R;
    T* o = nullptr;                                         // output of (first m of this lane's run in the item, phase group g)
    bool new_row = true;
    const size_t ostep = (size_t)CW * (size_t)D;           // outputs per sub-item
    for (uint32_t k = 0; k < per; ++k) {
        rs_cp_wait<GEO::DEPTH - 2>();                       // all but the DEPTH - 2 most recent copies have landed
        NW_WARP_SYNC();                                     // ... for every lane; and the strip read DEPTH - 1 items ago is free
        issue();
        if ((int)k >= cnt) continue;
        if (new_row) {                                      // output row through the group's frequency map
            const int gr = P.row0 + by, si = gr / P.F, fi = gr - si * P.F;
            o = (T*)P.out + ((size_t)si * (size_t)P.F_out + (size_t)(P.fmap ? P.fmap[fi] : fi)) * (size_t)P.N +
                ((size_t)gt * (size_t)CWI + (size_t)lane_off) * (size_t)D + (size_t)(g * PQ);
            new_row = false;
        }
        const bool fast = gt >= 1 && gt <= gt_fast;
        const int a0 = fast ? (int)(((uintptr_t)csrc / sizeof(cx<T>)) & 1) : 0;
        const cx<T>* yb = cstrip + a0;
        int mrun = gt * CWI + lane_off;                     // first m of this lane's run in the sub-item
        const bool full = gt < GT - 1;                      // every run of the item lies inside the row
        csrc += CWI;
        if (++gt == GT) { gt = 0; ++by; csrc += P.ystride - (long long)GT * CWI; new_row = true; }
        cstrip += GEO::SLOTS;
        if (cstrip >= ring1) cstrip -= GEO::DEPTH * GEO::SLOTS;
#pragma unroll 1
        for (int sub = 0; sub < NSUB; ++sub, yb += CW, mrun += CW, o += ostep) {
            if (!act) continue;
            cx<T> w[R + K - 1];
#pragma unroll
            for (int q = 0; q < R + K - 1; ++q) w[q] = yb[q];
            if (full) {                                     // straight-line: the chains of neighbouring m interleave
#pragma unroll
                for (int mm = 0; mm < R; ++mm) {
                    RsVec<T, PQ> r;
#pragma unroll
#ifdef RS_EXP_NOFMA   // timing experiment: one multiply per output instead of the K taps
                    for (int q = 0; q < PQ; ++q) r.v[q] = w[mm + (q & 1)].x * c[q & 1].v[q];
#else
                    for (int q = 0; q < PQ; q += 2) rs_pair<K, MODE>(w + mm, &c[0].v[q], PQ, r.v[q], r.v[q + 1]);
#endif
#ifdef RS_EXP_NOST    // timing experiment: results are computed but (almost) never stored
                    if (r.v[0] == (T)123.456) rs_st_out(o + (size_t)mm * (size_t)D, r);
#else
                    rs_st_out(o + (size_t)mm * (size_t)D, r);
#endif
                }
            } else {                                        // the row ends inside this item
                const int left = M - mrun;
#pragma unroll
                for (int mm = 0; mm < R; ++mm) {
                    RsVec<T, PQ> r;
#pragma unroll
                    for (int q = 0; q < PQ; q += 2) rs_pair<K, MODE>(w + mm, &c[0].v[q], PQ, r.v[q], r.v[q + 1]);
                    if (mm < left) rs_st_out(o + (size_t)mm * (size_t)D, r);
                }
            }
        }
    }
}

// host-side launch geometry of the direct kernel for a group (D, K) at decimated length M; false: not eligible
struct ResampleDirShape { int R, PQ, G, MW, NSUB, nthr, ctas_per_sm; unsigned items; size_t smem; };
constexpr int rs_dir_run(int K) { return K <= 8 ? 7 : 5; }   // odd: conflict-free window loads; window 2 (R + K - 1) + weights PQ K registers
template <typename T> inline bool resample_dir_shape(int D, int K, long long M, long long rows, int F, ResampleDirShape& v) {
    if (sizeof(T) != 4 || (D & 1) || (K & 1) || K < 4 || K > 12) return false;
    v.PQ = (D & 3) ? 2 : 4;
    v.G = D / v.PQ;
    if (v.G > 32) return false;
    v.MW = 32 / v.G;
    v.R = rs_dir_run(K);
    if (M < (long long)v.MW * v.R + K + 2) return false;   // an item's samples never wrap twice
    if ((rows + 1) * (long long)F >= (1LL << 32)) return false;   // fastdiv of the row index by F
    v.nthr = 128;
    v.ctas_per_sm = 4;
    const int slots = 16 * v.R * 4 + K + 2 + (K & 1);      // RsDirGeo::SLOTS
    v.smem = (size_t)(v.nthr / 32) * 3 * (size_t)slots * 2 * sizeof(T);   // RsDirGeo::DEPTH strips per warp
    v.NSUB = (slots - K - 2) / (v.MW * v.R);               // sub-items per item: as many as fill a strip
    if (v.NSUB > 8) v.NSUB = 8;
    while (v.NSUB > 1 && (long long)v.NSUB * v.MW * v.R + K + 2 > M) --v.NSUB;
    const long long cw = (long long)v.NSUB * v.MW * v.R;
    v.items = (unsigned)((M + cw - 1) / cw);
    if ((long long)v.items * rows >= (1LL << 31)) return false;
    return true;
}
inline unsigned resample_dir_grid(const ResampleDirShape& v, long long rows, int sms) {
    const long long wpc = v.nthr / 32;
    const long long ctas = ((long long)v.items * rows + wpc - 1) / wpc;
    const long long cap = (long long)sms * v.ctas_per_sm;
    return (unsigned)(ctas < cap ? ctas : cap);
}

// resident CTAs per SM the vector kernel's register budget is sized for (__launch_bounds__): window 2 (R + K - 1),
// weights PQ K, prefetched samples 2 (R + 1), accumulators and addresses ~32
constexpr int rsv_min_ctas(int K, int R, int PQ) {
    const int need = 2 * (R + K - 1) + PQ * K + 2 * (R + 1) + 32;
    return need <= 96 ? 5 : need <= 128 ? 4 : 3;
}

// host-side launch geometry of the vector kernel for a group (D, K) at decimated length M; false: not eligible
// (odd D or K, or a tap count / precision the vector kernels are not compiled for - the scalar kernel takes those).
// R, WR, WP: the most evenly shared phase groups first, then longer runs, then the smaller staging tile (<= 66 KB).
struct ResampleVecShape { int R, PQ, WR, WP, nthr, C, ctas_per_sm; unsigned items; size_t smem, gbytes; fastdiv dRD; };
template <typename T> inline bool resample_vec_shape(int D, int K, long long M, ResampleVecShape& v) {
    if (sizeof(T) != 4 || (D & 1) || (K & 1) || K < 4 || K > 12) return false;
    v.PQ = (D & 3) ? 2 : 4;
    v.nthr = 128;
    const int G = D / v.PQ;
    double best = -1;
    static const int RR[2] = {8, 4}, WW[3][2] = {{4, 1}, {2, 2}, {1, 4}};
    static const double RF[2] = {1.0, 0.95};   // runs of 4: more window and weight loads per output (runs of 16 spill)
    for (int ri = 0; ri < 2; ++ri)
        for (int wi = 0; wi < 3; ++wi) {
            const int R = RR[ri], WR = WW[wi][0], WP = WW[wi][1];
            const size_t tile = (size_t)32 * WR * ((size_t)R * D + v.PQ) * sizeof(T);
            const int fit = (int)((227 * 1024) / (tile + 1024));
            if (fit < 3) continue;
            const double eff = (double)G / (double)(((G + WP - 1) / WP) * WP);   // phase groups shared evenly by the phase-warps
            const int rc = rsv_min_ctas(K, R, v.PQ);
            const double occ = (double)(fit < rc ? fit : rc) / 5.0;                 // resident CTAs: shared memory, registers
            const double score = eff * occ * RF[ri] - 0.01 * (WP - 1);
            if (score > best + 1e-9) { best = score; v.R = R; v.WR = WR; v.WP = WP; }
        }
    if (best < 0) return false;
    v.C = 32 * v.WR * v.R;
    const size_t CG = (size_t)32 * v.R;
    v.items = (unsigned)((M + (long long)CG - 1) / (long long)CG);
    const size_t ys = (CG + K + (CG + K) / v.R + 2) * 2 * sizeof(T);
    const size_t tile = (size_t)32 * ((size_t)v.R * D + v.PQ) * sizeof(T);
    v.gbytes = ((ys > tile ? ys : tile) + 15) / 16 * 16;
    v.smem = v.gbytes * v.WR;
    const int fit = (int)((227 * 1024) / (v.smem + 1024));
    const int rc = rsv_min_ctas(K, v.R, v.PQ);
    v.ctas_per_sm = fit < rc ? fit : rc;
    v.dRD = make_fastdiv((uint32_t)(v.R * D / v.PQ));
    return true;
}
// persistent grid of the vector kernel for `rows` rows on a device with `sms` SMs
inline unsigned resample_vec_grid(const ResampleVecShape& v, long long rows, int sms) {
    const long long ctas = ((long long)v.items * rows + v.WR - 1) / v.WR;
    const long long cap = (long long)sms * v.ctas_per_sm;
    return (unsigned)(ctas < cap ? ctas : cap);
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
