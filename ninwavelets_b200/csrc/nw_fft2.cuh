// In-place, packed, mixed-radix shared-memory FFT engine (the fast path).
//
// A CTA transforms TT = 2*TP interleaved length-P sequences that live in ONE
// shared-memory buffer buf[p * TP + tp] of two-lane complex values (cx2: lanes
// = sequences 2*tp and 2*tp+1).  Every stage is an in-place Cooley-Tukey pass:
// a butterfly reads and writes the same R slots, so there is no ping-pong
// buffer and one barrier per stage; a thread takes as many butterflies as the
// loop gives it, so any thread count works.
//
//   fft2_dif   natural-order input, digit-reversed output (decimation in frequency)
//              stage s works in blocks of L_s = P / (R_0..R_{s-1}) with stride
//              L_s / R_s and multiplies output r by w_{L_s}^{n' r}.  The last
//              stage hands result index  rev(blk) + (P / R_last) * r  to Dst.
//   fft2_dit   digit-reversed input (gathered through Src by the first pass, which
//              runs the LAST radix), natural-order output: the transposed network.
//
// Twiddles: one table load per butterfly (w_{L_s}^{n'}), the other R-2 powers are
// formed in registers - shared-memory / L1 bandwidth, not arithmetic, is the scarce
// resource of these kernels ("what bounds the kernels").
//
// Src / Dst see a whole butterfly at a time so that they can hoist index arithmetic and run
// recurrences over r:   src.load_all<R>(base, step, tp, v)   fills v[r] with element base + r*step of
// sequences 2*tp, 2*tp+1;   ctx = dst.begin(base, step, tp)  is called before the butterfly's loads (so
// the epilogue can start its own table loads) and   dst.store_all<R>(ctx, v)   receives results
// base + r*step.
//
// DIR = +1: e^{+2 pi i n k / P} (inverse, unnormalised); DIR = -1: forward.
// tw[j] = e^{+2 pi i j / P}, j in [0, P).
#pragma once
#include "nw_common.h"
#include "nw_pk.cuh"
#include "nw_bfly2.cuh"

namespace nw {

static const int MAX_PACKED_RADIX = 16;

// digit reversal of a butterfly block index for the last DIF / first DIT pass:
// blk = ((k_0 R_1 + k_1) R_2 + ...) over radices 0..m-2  ->  k_0 + R_0 k_1 + R_0 R_1 k_2 + ...
NW_HD int fft2_rev(const Fft2Plan& st, int blk) {
    int rev = 0;
    // peel digits from the least significant (radix m-2) upwards
    for (int s = st.nst - 2; s >= 0; --s) {
        const int q = (int)fd_div((uint32_t)blk, st.div_r[s]);
        const int d = blk - q * st.radix[s];
        rev += d * st.ns[s];
        blk = q;
    }
    return rev;
}

// w^r for r = 1..R-1 from w (registers only)
template <typename T, int R> NW_HD void tw_powers(cx<T> w, cx<T>* p) {
    p[1] = w;
#pragma unroll
    for (int r = 2; r < R; ++r) p[r] = cmul(p[r / 2], p[r - r / 2]);
}

template <typename T, int DIR> NW_HD cx<T> tw_dir(cx<T> w) { return DIR > 0 ? w : mk<T>(w.x, -w.y); }

// w^r, r = 1..R-1, for w = tw[idx]: one table load, the powers by a product tree (depth log2 R).  A second table
// seed (w^(R/2)) would halve the rounding error of the powers but was measured to DOUBLE the kernel times on B200:
// its lane addresses are R/2 times further apart, and the kernels are bound by L1 / shared-memory wavefronts.
template <typename T, int R, int DIR> NW_HD void tw_powers_tab(const cx<T>* NW_RESTRICT tw, int idx, cx<T>* p) {
    p[1] = tw_dir<T, DIR>(tw[idx]);
#pragma unroll
    for (int r = 2; r < R; ++r) p[r] = cmul(p[r / 2], p[r - r / 2]);
}

// Geometry of one pass: P points, blocks of L = P / ns with Q = L / R butterflies each (= element
// stride), twiddle table stride tws = ns, 2^tpsh lane pairs interleaved.  GeoDyn reads a run-time
// plan; GeoStat is the same interface with everything a compile-time constant - used by the kernels
// specialised for the hot lengths, where the element offsets become immediates of the LDS/STS and the
// divisions fold away (about a third fewer instructions per butterfly than the run-time form).
// How many lane-pair sequences are interleaved in the buffer: 2^tpsh (tpsh >= 0), or any count with an
// exact-division helper (tpsh < 0; pruned transforms interleave columns x phases).  twscale multiplies the
// twiddle-table index: a length-P transform can use the table of any multiple of P.
struct SeqDesc {
    int tpsh, nseq, twscale;
    fastdiv d;
};
NW_HD SeqDesc seq_pow2(int tpsh) { SeqDesc q; q.tpsh = tpsh; q.nseq = 1 << tpsh; q.twscale = 1; q.d.d = 1; q.d.m = 0; return q; }
inline SeqDesc seq_any(int nseq, int twscale) { SeqDesc q; q.tpsh = -1; q.nseq = nseq; q.twscale = twscale; q.d = make_fastdiv((uint32_t)nseq); return q; }

struct GeoDyn {
    const Fft2Plan* st;
    int s, P, L, Q, tws;
    SeqDesc sq;
    NW_HD GeoDyn(const Fft2Plan& p, int stage, const SeqDesc& q, int R) : st(&p), s(stage), P(p.P), L(p.P / p.ns[stage]),
        Q(p.P / p.ns[stage] / R), tws(p.ns[stage] * q.twscale), sq(q) {}
    NW_HD int nseq() const { return sq.nseq; }
    NW_HD void split(uint32_t lin, int& t, int& bi) const {
        if (sq.tpsh >= 0) { t = (int)(lin & (uint32_t)(sq.nseq - 1)); bi = (int)(lin >> sq.tpsh); }
        else { bi = (int)fd_div(lin, sq.d); t = (int)lin - bi * sq.nseq; }
    }
    NW_HD int blk_of(int bi) const { return (int)fd_div((uint32_t)bi, st->div_q[s]); }
    NW_HD int rev(int blk) const { return fft2_rev(*st, blk); }
};
struct RevNone { static NW_HD int rev(int blk) { return blk; } };                    // <= 2 passes
template <int R0, int R1> struct Rev3 {                                                // 3 passes: blk = k0 R1 + k1
    static NW_HD int rev(int blk) { const int k0 = blk / R1; return k0 + R0 * (blk - k0 * R1); }
};
template <int PS, int NSS, int R, int TPS, class REV> struct GeoStat {
    static constexpr int P = PS, L = PS / NSS, Q = PS / NSS / R, tws = NSS, tpsh = TPS;
    NW_HD constexpr int nseq() const { return 1 << TPS; }
    NW_HD void split(uint32_t lin, int& t, int& bi) const { t = (int)(lin & ((1u << TPS) - 1u)); bi = (int)(lin >> TPS); }
    NW_HD int blk_of(int bi) const { return bi / Q; }
    NW_HD int rev(int blk) const { return REV::rev(blk); }
};

// ---- decimation in frequency ------------------------------------------------------------
template <typename T, int R, int DIR, bool LAST, bool RAW, class G, class Dst>
NW_HD void dif_body(const G& g, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Dst& dst, int tid, int nthr) {
    const int nseq = g.nseq();
    const uint32_t nwork = (uint32_t)(g.P / R) * (uint32_t)nseq;
#pragma unroll 1
    for (uint32_t lin = tid; lin < nwork; lin += nthr) {
        int tp, bi;
        g.split(lin, tp, bi);
        const int blk = LAST ? bi : g.blk_of(bi);   // bi / Q
        const int np = bi - blk * g.Q;
        cx2<T>* e = buf + ((size_t)blk * g.L + np) * nseq + tp;
        const int stride = g.Q * nseq;
        cx2<T> v[R];
        typename Dst::Ctx ctx;
        if (LAST) ctx = dst.begin(g.rev(blk), g.P / R, tp);   // issues the epilogue's own loads early
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = RAW ? raw_to_packed<T>(e[r * stride]) : e[r * stride];
        B2<T, R, DIR>::run(v);
        if (LAST) {
            dst.template store_all<R>(ctx, v);
        } else {
            cx<T> w[R];
            tw_powers_tab<T, R, DIR>(tw, np * g.tws, w);
            e[0] = v[0];
#pragma unroll
            for (int r = 1; r < R; ++r) e[r * stride] = cmul_s(v[r], w[r]);
        }
    }
}

template <typename T, int R, int DIR, bool LAST, bool RAW, class Dst>
NW_HD void dif_stage(const Fft2Plan& st, int s, const SeqDesc& tpsh, const cx<T>* NW_RESTRICT tw, cx2<T>* buf,
                     const Dst& dst, int tid, int nthr) {
    dif_body<T, R, DIR, LAST, RAW>(GeoDyn(st, s, tpsh, R), tw, buf, dst, tid, nthr);
}

template <typename T, int DIR, bool LAST, bool RAW, class Dst>
NW_HD void dif_stage_any(const Fft2Plan& st, int s, const SeqDesc& tpsh, const cx<T>* NW_RESTRICT tw, cx2<T>* buf,
                         const Dst& dst, int tid, int nthr) {
    switch (st.radix[s]) {
        case 2: dif_stage<T, 2, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 3: dif_stage<T, 3, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 4: dif_stage<T, 4, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 5: dif_stage<T, 5, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 6: dif_stage<T, 6, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 8: dif_stage<T, 8, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 10: dif_stage<T, 10, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 12: dif_stage<T, 12, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 15: dif_stage<T, 15, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        case 16: dif_stage<T, 16, DIR, LAST, RAW>(st, s, tpsh, tw, buf, dst, tid, nthr); break;
        default: break;
    }
}

// buf holds the input in natural order on entry (visible to the CTA); results go to dst.
// RAW0: the units of buf are two plain complex values {re0, im0, re1, im1} (a tile as it sits in
// global memory) instead of lane-packed {re0, re1, im0, im1}; the first pass repacks in registers.
template <typename T, int DIR, bool RAW0, class Dst>
NW_HD void fft2_dif(const Fft2Plan& st, int tpsh_, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Dst& dst, int tid,
                    int nthr) {
    const SeqDesc tpsh = seq_pow2(tpsh_);
    if (st.nst == 1) {
        dif_stage_any<T, DIR, true, RAW0>(st, 0, tpsh, tw, buf, dst, tid, nthr);
        return;
    }
    dif_stage_any<T, DIR, false, RAW0>(st, 0, tpsh, tw, buf, dst, tid, nthr);
    NW_SYNC();
#pragma unroll 1
    for (int s = 1; s < st.nst - 1; ++s) {
        dif_stage_any<T, DIR, false, false>(st, s, tpsh, tw, buf, dst, tid, nthr);
        NW_SYNC();
    }
    dif_stage_any<T, DIR, true, false>(st, st.nst - 1, tpsh, tw, buf, dst, tid, nthr);
}

// The same transform for a compile-time plan P = R0 R1 R2 (R2 = 1: two passes).
template <typename T, int DIR, bool RAW0, int TPS, int P, int R0, int R1, int R2, class Dst>
NW_HD void fft2_dif_static(const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Dst& dst, int tid, int nthr) {
    dif_body<T, R0, DIR, false, RAW0>(GeoStat<P, 1, R0, TPS, RevNone>(), tw, buf, dst, tid, nthr);
    NW_SYNC();
    if constexpr (R2 > 1) {
        dif_body<T, R1, DIR, false, false>(GeoStat<P, R0, R1, TPS, RevNone>(), tw, buf, dst, tid, nthr);
        NW_SYNC();
        dif_body<T, (R2 > 1 ? R2 : 2), DIR, true, false>(GeoStat<P, R0 * R1, (R2 > 1 ? R2 : 2), TPS, Rev3<R0, R1>>(), tw, buf, dst, tid, nthr);
    } else {
        dif_body<T, R1, DIR, true, false>(GeoStat<P, R0, R1, TPS, RevNone>(), tw, buf, dst, tid, nthr);
    }
}

// Src tag: the input already sits in buf at its decimation-in-time position (fft2_dit_pos)
struct FromBuf {};
template <class S> struct is_from_buf { static const bool value = false; };
template <> struct is_from_buf<FromBuf> { static const bool value = true; };

// slot of input element n for fft2_dit: butterfly blk = unrev(n mod step) of the first pass, input r = n / step
NW_HD int fft2_dit_pos(const Fft2Plan& st, int n) {
    const int rl = st.radix[st.nst - 1];
    const int step = st.P / rl;
    const int r = n / step;
    int x = n - r * step, blk = 0;
    for (int s = 0; s <= st.nst - 2; ++s) {
        const int q = (int)fd_div((uint32_t)x, st.div_r[s]);
        blk = blk * st.radix[s] + (x - q * st.radix[s]);
        x = q;
    }
    return blk * rl + r;
}

// ---- decimation in time ---------------------------------------------------------------------
// pass index q = 0 runs radix[nst-1] on contiguous groups, reading from Src at digit-reversed
// indices; the last pass runs radix[0] at stride P / R_0 and writes natural-order results to Dst.
template <typename T, int R, int DIR, bool FIRST, bool LAST, class G, class Src, class Dst>
NW_HD void dit_body(const G& g, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Src& src, const Dst& dst, int tid,
                    int nthr) {
    const int nseq = g.nseq();
    const uint32_t nwork = (uint32_t)(g.P / R) * (uint32_t)nseq;
#pragma unroll 1
    for (uint32_t lin = tid; lin < nwork; lin += nthr) {
        int tp, bi;
        g.split(lin, tp, bi);
        const int blk = FIRST ? bi : g.blk_of(bi);
        const int np = bi - blk * g.Q;
        cx2<T>* e = buf + ((size_t)blk * g.L + np) * nseq + tp;
        const int stride = g.Q * nseq;
        cx2<T> v[R];
        typename Dst::Ctx ctx;
        if (LAST) ctx = dst.begin(np, g.Q, tp);   // blk == 0, L == P; issues the epilogue's own loads early
        if (FIRST) {
            if constexpr (is_from_buf<Src>::value) {
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = e[r * stride];
            } else {
                src.template load_all<R>(g.rev(blk), g.P / R, tp, v);
            }
        } else {
            cx<T> w[R];
            tw_powers_tab<T, R, DIR>(tw, np * g.tws, w);
            v[0] = e[0];
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul_s(e[r * stride], w[r]);
        }
        B2<T, R, DIR>::run(v);
        if (LAST) {
            dst.template store_all<R>(ctx, v);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) e[r * stride] = v[r];
        }
    }
}

template <typename T, int R, int DIR, bool FIRST, bool LAST, class Src, class Dst>
NW_HD void dit_stage(const Fft2Plan& st, int s, const SeqDesc& tpsh, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Src& src,
                     const Dst& dst, int tid, int nthr) {
    dit_body<T, R, DIR, FIRST, LAST>(GeoDyn(st, s, tpsh, R), tw, buf, src, dst, tid, nthr);
}

template <typename T, int DIR, bool FIRST, bool LAST, class Src, class Dst>
NW_HD void dit_stage_any(const Fft2Plan& st, int s, const SeqDesc& tpsh, const cx<T>* NW_RESTRICT tw, cx2<T>* buf,
                         const Src& src, const Dst& dst, int tid, int nthr) {
    switch (st.radix[s]) {
        case 2: dit_stage<T, 2, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 3: dit_stage<T, 3, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 4: dit_stage<T, 4, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 5: dit_stage<T, 5, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 6: dit_stage<T, 6, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 8: dit_stage<T, 8, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 10: dit_stage<T, 10, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 12: dit_stage<T, 12, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 15: dit_stage<T, 15, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        case 16: dit_stage<T, 16, DIR, FIRST, LAST>(st, s, tpsh, tw, buf, src, dst, tid, nthr); break;
        default: break;
    }
}

// No barrier on entry or exit: the caller orders buf's reuse and dst's visibility.
// SKIP_FIRST: the caller has already run the first pass (radix[nst-1] on contiguous groups; nst >= 2) itself.
template <typename T, int DIR, class Src, class Dst, bool SKIP_FIRST = false>
NW_HD void fft2_dit(const Fft2Plan& st, const SeqDesc& sq, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Src& src,
                    const Dst& dst, int tid, int nthr) {
    const int m = st.nst;
    if (m == 1) {
        dit_stage_any<T, DIR, true, true>(st, 0, sq, tw, buf, src, dst, tid, nthr);
        return;
    }
    if (!SKIP_FIRST) dit_stage_any<T, DIR, true, false>(st, m - 1, sq, tw, buf, src, dst, tid, nthr);
    NW_SYNC();
#pragma unroll 1
    for (int s = m - 2; s >= 1; --s) {
        dit_stage_any<T, DIR, false, false>(st, s, sq, tw, buf, src, dst, tid, nthr);
        NW_SYNC();
    }
    dit_stage_any<T, DIR, false, true>(st, 0, sq, tw, buf, src, dst, tid, nthr);
}
template <typename T, int DIR, class Src, class Dst>
NW_HD void fft2_dit(const Fft2Plan& st, int tpsh, const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Src& src,
                    const Dst& dst, int tid, int nthr) {
    fft2_dit<T, DIR>(st, seq_pow2(tpsh), tw, buf, src, dst, tid, nthr);
}

// Compile-time plan P = R0 R1 R2 (radix order as in Fft2Plan: the first pass runs the LAST radix; R2 = 1:
// two passes).  Input already in buf at its fft2_dit_pos slots.
template <typename T, int DIR, int TPS, int P, int R0, int R1, int R2, class Dst, bool SKIP_FIRST = false>
NW_HD void fft2_dit_static(const cx<T>* NW_RESTRICT tw, cx2<T>* buf, const Dst& dst, int tid, int nthr) {
    const FromBuf src;
    if constexpr (R2 > 1) {
        if (!SKIP_FIRST) dit_body<T, (R2 > 1 ? R2 : 2), DIR, true, false>(GeoStat<P, R0 * R1, (R2 > 1 ? R2 : 2), TPS, RevNone>(), tw, buf, src, dst, tid, nthr);
        NW_SYNC();
        dit_body<T, R1, DIR, false, false>(GeoStat<P, R0, R1, TPS, RevNone>(), tw, buf, src, dst, tid, nthr);
    } else {
        if (!SKIP_FIRST) dit_body<T, R1, DIR, true, false>(GeoStat<P, R0, R1, TPS, RevNone>(), tw, buf, src, dst, tid, nthr);
    }
    NW_SYNC();
    dit_body<T, R0, DIR, false, true>(GeoStat<P, 1, R0, TPS, RevNone>(), tw, buf, src, dst, tid, nthr);
}

}  // namespace nw
