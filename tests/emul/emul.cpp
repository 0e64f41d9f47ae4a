// Host emulation of the kernel bodies (TEST INFRASTRUCTURE, never shipped, never
// linked into libnwcwt.so).  The CUDA kernel bodies in ninwavelets_b200/csrc are
// written against (block, tid, nthr, smem); here each block is stepped by a single
// "thread" on the CPU so the index algebra, the radix plans, the band logic and the
// epilogues can be checked against the oracle in the GPU-less authoring container.
// It proves nothing about races or barriers - the GPU parity tests do that.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/nwcwt.h"
#include "../../ninwavelets_b200/csrc/nw_common.h"
#include "../../ninwavelets_b200/csrc/nw_fft.cuh"
#include "../../ninwavelets_b200/csrc/nw_family.cuh"
#include "../../ninwavelets_b200/csrc/nw_kernels.cuh"
#include "../../ninwavelets_b200/csrc/nw_kernels2.cuh"
#include "../../ninwavelets_b200/csrc/nw_kernels3.cuh"
#include "../../ninwavelets_b200/csrc/nw_resample.cuh"
#include "../../ninwavelets_b200/csrc/nw_kernels4.cuh"
#include "../../ninwavelets_b200/csrc/nw_plan.h"

#include <ucontext.h>
#include <functional>

using namespace nw;

// ---- a block of `nthr` emulated threads stepped as fibers; NW_SYNC() is a real barrier -----------
// Every thread runs until its next barrier (or its end), round robin: a missing or misplaced
// barrier shows up as a wrong result, exactly as thread 0 racing ahead would on the device.
namespace {
struct Fibers {
    static const size_t STACK = 256 * 1024;
    ucontext_t main_ctx;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    std::vector<char> stacks;
    const std::function<void(int)>* body = nullptr;
    int cur = 0;
    static Fibers& get() { static Fibers f; return f; }
    static void sync_hook() { Fibers& f = get(); swapcontext(&f.ctx[f.cur], &f.main_ctx); }
    static void tramp() {
        Fibers& f = get();
        (*f.body)(f.cur);
        f.done[f.cur] = 1;
        swapcontext(&f.ctx[f.cur], &f.main_ctx);
    }
    void run(int nthr, const std::function<void(int)>& fn) {
        if (nthr <= 1) { host_sync_hook() = nullptr; fn(0); return; }
        body = &fn;
        ctx.assign(nthr, ucontext_t());
        done.assign(nthr, 0);
        if (stacks.size() < (size_t)nthr * STACK) stacks.resize((size_t)nthr * STACK);
        for (int t = 0; t < nthr; ++t) {
            getcontext(&ctx[t]);
            ctx[t].uc_stack.ss_sp = stacks.data() + (size_t)t * STACK;
            ctx[t].uc_stack.ss_size = STACK;
            ctx[t].uc_link = &main_ctx;
            makecontext(&ctx[t], (void (*)())tramp, 0);
        }
        host_sync_hook() = sync_hook;
        for (bool any = true; any;) {
            any = false;
            for (int t = 0; t < nthr; ++t)
                if (!done[t]) { cur = t; swapcontext(&main_ctx, &ctx[t]); any = any || !done[t]; }
        }
        host_sync_hook() = nullptr;
    }
};
int g_mode = 0;   // bit 1: generic kernels only; bit 2: one stepping thread per block (no fibers)
}  // namespace

template <typename T>
static void fill_tw(std::vector<cx<T>>& v, long long count, long long P, long long step) {
    v.resize((size_t)count);
    const long double tp = 6.283185307179586476925286766559005768L;
    for (long long j = 0; j < count; ++j) {
        const long long m = (j * step) % P;
        const long double a = tp * (long double)m / (long double)P;
        v[(size_t)j].x = (T)cosl(a);
        v[(size_t)j].y = (T)sinl(a);
    }
}

// SpecParams of a plan (the main one or a group's sub-plan)
template <typename T>
static SpecParams<T> make_sp(const HostPlan& hp, const std::vector<cx<T>>& table) {
    SpecParams<T> sp;
    sp.family = hp.family; sp.grid_off = hp.grid_off; sp.df = hp.df; sp.p0 = hp.p0; sp.p1 = hp.p1;
    sp.p2 = hp.family == FAM_MORSE ? hp.p0 / hp.p1 : hp.p2;
    sp.norm = (T)(1.0 / (double)hp.data_len());
    sp.rec = hp.rec.data(); sp.table = table.data(); sp.table_len = hp.table_len;
    sp.wtab = nullptr;
    return sp;
}
template <typename T>
static void make_table(const HostPlan& hp, std::vector<cx<T>>& table) {
    table.clear();
    if (hp.family != FAM_TABLE) return;
    table.resize((size_t)hp.F * hp.table_len);
    for (size_t i = 0; i < table.size(); ++i) { table[i].x = (T)hp.table[2 * i]; table[i].y = (T)hp.table[2 * i + 1]; }
}

template <typename T, int K>
static void resample_launch(const ResampleParams<T>& R, int mode, char* smp, int tiles, int g, int nt) {
    for (int y = 0; y < g; ++y) for (int x = 0; x < tiles; ++x)
        Fibers::get().run(nt, [&](int t) {
            if (mode == OUT_POWER) resample_body<T, K, ResampleRun<T>::R, OUT_POWER>(R, smp, x, y, t, nt);
            else resample_body<T, K, ResampleRun<T>::R, OUT_ABS>(R, smp, x, y, t, nt);
        });
}

template <typename T, int K, int PQ>
static void resample_vec_launch(ResampleParams<T> R, int mode, char* smp, const ResampleVecShape& v, int g) {
    R.nrows = g;
    const int nbx = (int)std::min<unsigned>(resample_vec_grid(v, g, 1), 3u);   // a small persistent grid: every CTA loops
    for (int x = 0; x < nbx; ++x)
        Fibers::get().run(v.nthr, [&](int t) {
#define RV(r) case r: if (mode == OUT_POWER) resample_vec_body<T, K, r, PQ, OUT_POWER>(R, smp, x, nbx, t, v.nthr); \
                      else resample_vec_body<T, K, r, PQ, OUT_ABS>(R, smp, x, nbx, t, v.nthr); break;
            switch (v.R) { RV(8) RV(4) }
#undef RV
        });
}

template <typename T, int K, int PQ>
static void resample_dir_launch(ResampleParams<T> R, int mode, const ResampleDirShape& v, int g) {
    R.nrows = g;
    const int nbx = (int)std::min<unsigned>(resample_dir_grid(v, g, 1), 3u);   // a small persistent grid: every warp loops
    constexpr int RR = rs_dir_run(K);
    std::vector<char> strip(v.smem + 64);
    char* sp = (char*)(((uintptr_t)strip.data() + 31) & ~(uintptr_t)31);
    for (int x = 0; x < nbx; ++x)
        Fibers::get().run(v.nthr, [&](int t) {   // warp-level synchronisation is emulated by CTA barriers
            if (mode == OUT_POWER) resample_dir_body<T, K, RR, PQ, OUT_POWER>(R, sp, x, nbx, t, v.nthr);
            else resample_dir_body<T, K, RR, PQ, OUT_ABS>(R, sp, x, nbx, t, v.nthr);
        });
}

// Inverse transforms of gs signals x the frequencies of plan eh (the main plan, or the sub-plan of group mg) on the
// packed kernels, followed by the interpolation kernel for a resampled group: mirrors nwcwt.cu: inverse_rows.
template <typename T>
static int inverse_rows(const HostPlan& hp, const HostPlan& eh_in, const MrGroup* mg, const cx<T>* X, void* out_s0, int gs, int output) {
    HostPlan eh = eh_in;   // the weight table sets FreqRec::woff
    std::vector<T> wtab;
    const bool have_wtab = !(g_mode & 256) && build_weight_table<T>(eh, mg && mg->D > 1 ? mg->eq.data() : nullptr, (size_t)512 << 20, wtab);
    std::vector<cx<T>> table, twA2, twB2, twH, twL;
    make_table<T>(eh, table);
    fill_tw<T>(twA2, eh.N1f, eh.N1f, 1);
    fill_tw<T>(twB2, eh.N2f, eh.N2f, 1);
    const long long nL = 1LL << eh.lb, nH = (eh.N + nL - 1) / nL;
    fill_tw<T>(twL, nL, eh.N, 1);
    fill_tw<T>(twH, nH, eh.N, nL);
    const int D = mg ? mg->D : 1;
    Long2Params<T> Q;
    memset(&Q, 0, sizeof(Q));
    Q.N = eh.N; Q.xstride = hp.N; Q.N1 = eh.N1f; Q.N2 = eh.N2f; Q.F = eh.F; Q.tpshA = eh.tpshA; Q.tpshB = eh.tpshB;
    Q.stA = eh.stA2; Q.stB = eh.stB2; Q.twA = twA2.data(); Q.twB = twB2.data(); Q.twH = twH.data(); Q.twL = twL.data();
    Q.lb = eh.lb; Q.tm_stride = eh.tm_stride2; Q.sp = make_sp<T>(eh, table);
    if (have_wtab) Q.sp.wtab = wtab.data();
    std::vector<int> ditpos((size_t)eh.N1f);
    for (int k1 = 0; k1 < eh.N1f; ++k1) ditpos[(size_t)k1] = fft2_dit_pos(eh.stA2, k1);
    if (!(g_mode & 256)) Q.ditpos = ditpos.data();
    const int ring2 = eh.ring2 < 3 ? eh.ring2 : 3;
    std::vector<cx<T>> Tm2((size_t)ring2 * eh.tm_stride2), Y;
    Q.X = X; Q.Tm = Tm2.data();
    std::vector<T> eq, coef;
    ResampleParams<T> R;
    ResampleShape shp{1, 1, 0, 0, 0};
    ResampleVecShape vshp;
    ResampleDirShape dshp;
    bool vec = false, dir = false;
    std::vector<T> coefq;
    memset(&R, 0, sizeof(R));
    if (D > 1) {
        vec = resample_vec_shape<T>(D, mg->K, eh.N, vshp);
        dir = vec && !(g_mode & 512) && resample_dir_shape<T>(D, mg->K, eh.N, (long long)gs * eh.F, eh.F, dshp) && dshp.PQ == vshp.PQ;
        eq.assign(mg->eq.begin(), mg->eq.end());
        coef.assign(mg->coef.begin(), mg->coef.end());
        Y.resize((size_t)ring2 * eh.N);
        shp = resample_shape<T>(D, mg->K);
        R.ystride = eh.N; R.out = out_s0; R.N = hp.N; R.M = (int)eh.N; R.D = D; R.coef = coef.data(); R.t0 = mg->t0.data();
        R.t0min = *std::min_element(mg->t0.begin(), mg->t0.end());
        R.fmap = mg->fidx.data(); R.F = eh.F; R.F_out = hp.F; R.WR = shp.WR; R.WP = shp.WP; R.RS = shp.RS;
        R.dRD = make_fastdiv((uint32_t)(ResampleRun<T>::R * D));
        if (vec) {
            coefq.resize(coef.size());
            resample_coefq<T>(mg->coef.data(), D, mg->K, vshp.PQ, coefq.data());
            R.coefq = coefq.data();
            R.WR = vshp.WR; R.WP = vshp.WP; R.RS = (int)vshp.gbytes; R.dRD = vshp.dRD; R.dGT = make_fastdiv(vshp.items); R.dGT.d = vshp.items;
            if (dir) { R.G = dshp.G; R.MW = dshp.MW; R.NSUB = dshp.NSUB; R.dG = make_fastdiv((uint32_t)dshp.G); R.dF = make_fastdiv((uint32_t)eh.F); R.dGT = make_fastdiv(dshp.items); R.dGT.d = dshp.items; }
        }
        Q.eq = eq.data();
        Q.out_mode = OUT_CWT;
    } else {
        Q.out = out_s0; Q.out_mode = output;
        if (mg) { Q.fmap = mg->fidx.data(); Q.F_out = hp.F; }
    }
    const int tA = (eh.N2f + (2 << eh.tpshA) - 1) / (2 << eh.tpshA), tB = (eh.N1f + (2 << eh.tpshB) - 1) / (2 << eh.tpshB);
    std::vector<char> sm2(std::max(std::max(std::max(eh.smem_A2, eh.smem_B2), shp.smem), vec ? vshp.smem : (size_t)0) + 64);
    char* smp = (char*)(((uintptr_t)sm2.data() + 31) & ~(uintptr_t)31);
    const int ntA = (g_mode & 4) ? 1 : eh.nthrA2, ntB = (g_mode & 4) ? 1 : eh.nthrB2;
    const long long rows = (long long)gs * eh.F;
    if (eh.stA2.nst >= 1) Q.dstepA = make_fastdiv((uint32_t)std::max(1, eh.N1f / eh.stA2.radix[eh.stA2.nst - 1]));
    const int spA = (g_mode & 8) ? 0 : static_plan_id(eh.stA2, eh.tpshA), spB = (g_mode & 8) ? 0 : static_plan_id(eh.stB2, eh.tpshB);
    const int omode = D > 1 ? OUT_CWT : output;
    for (long long r0 = 0; r0 < rows; r0 += ring2) {
        const int g = (int)std::min<long long>(ring2, rows - r0);
        Q.row0 = (int)r0;
        if (D > 1) Q.out = (char*)Y.data() - (size_t)r0 * (size_t)eh.N * sizeof(cx<T>);
        for (int y = 0; y < g; ++y) for (int x = 0; x < tA; ++x)
            Fibers::get().run(ntA, [&](int t) {
                if (group_narrow(eh, r0, g) && !(g_mode & 64)) {
                    if (spA == 2) passA2_body<T, 2, true>(Q, smp, x, y, t, ntA);
                    else passA2_body<T, 0, true>(Q, smp, x, y, t, ntA);
                } else {
                    if (spA == 2) passA2_body<T, 2>(Q, smp, x, y, t, ntA);
                    else if (spA == 4) passA2_body<T, 4>(Q, smp, x, y, t, ntA);
                    else if (spA == 20) passA2_body<T, 20>(Q, smp, x, y, t, ntA);
                    else if (spA == 21) passA2_body<T, 21>(Q, smp, x, y, t, ntA);
                    else if (spA == 22) passA2_body<T, 22>(Q, smp, x, y, t, ntA);
                    else passA2_body<T, 0>(Q, smp, x, y, t, ntA);
                }
            });
        for (int y = 0; y < g; ++y) for (int x = 0; x < tB; ++x)
            Fibers::get().run(ntB, [&](int t) {
                if (omode == OUT_POWER) { if (spB == 1) passB2_body<T, OUT_POWER, 1>(Q, smp, x, y, t, ntB); else if (spB == 4) passB2_body<T, OUT_POWER, 4>(Q, smp, x, y, t, ntB); else passB2_body<T, OUT_POWER, 0>(Q, smp, x, y, t, ntB); }
                else if (omode == OUT_ABS) passB2_body<T, OUT_ABS, 0>(Q, smp, x, y, t, ntB);
                else {
                    if (spB == 1) passB2_body<T, OUT_CWT, 1>(Q, smp, x, y, t, ntB);
                    else if (spB == 23) passB2_body<T, OUT_CWT, 23>(Q, smp, x, y, t, ntB);
                    else if (spB == 24) passB2_body<T, OUT_CWT, 24>(Q, smp, x, y, t, ntB);
                    else if (spB == 25) passB2_body<T, OUT_CWT, 25>(Q, smp, x, y, t, ntB);
                    else if (spB == 26) passB2_body<T, OUT_CWT, 26>(Q, smp, x, y, t, ntB);
                    else passB2_body<T, OUT_CWT, 0>(Q, smp, x, y, t, ntB);
                }
            });
        if (D > 1) {
            R.y = Y.data(); R.row0 = (int)r0;
            const int tiles = (int)((eh.N + shp.C - 1) / shp.C), nt = 32 * shp.WR * shp.WP;
            if (dir) {
                switch (mg->K * 10 + dshp.PQ) {
#define RD_CASE(k) case k * 10 + 4: resample_dir_launch<T, k, 4>(R, output, dshp, g); break; \
                   case k * 10 + 2: resample_dir_launch<T, k, 2>(R, output, dshp, g); break;
                    RD_CASE(4) RD_CASE(6) RD_CASE(8) RD_CASE(10) RD_CASE(12)
#undef RD_CASE
                    default: return -2;
                }
            } else if (vec) {
                switch (mg->K * 10 + vshp.PQ) {
#define RV_CASE(k) case k * 10 + 4: resample_vec_launch<T, k, 4>(R, output, smp, vshp, g); break; \
                   case k * 10 + 2: resample_vec_launch<T, k, 2>(R, output, smp, vshp, g); break;
                    RV_CASE(4) RV_CASE(6) RV_CASE(8) RV_CASE(10) RV_CASE(12)
#undef RV_CASE
                    default: return -2;
                }
            } else switch (mg->K) {
#define RS_CASE(k) case k: resample_launch<T, k>(R, output, smp, tiles, g, nt); break;
                RS_CASE(4) RS_CASE(5) RS_CASE(6) RS_CASE(7) RS_CASE(8) RS_CASE(9) RS_CASE(10) RS_CASE(11) RS_CASE(12) RS_CASE(13)
                RS_CASE(14) RS_CASE(15) RS_CASE(16) RS_CASE(18) RS_CASE(20) RS_CASE(22) RS_CASE(24)
#undef RS_CASE
                default: return -2;
            }
        }
    }
    return 0;
}

template <typename T>
static int run(const HostPlan& hp, const void* signals, void* out, long long S, int output, int bl, long long blo,
               long long bhi) {
    std::vector<cx<T>> table;
    make_table<T>(hp, table);
    SpecParams<T> sp = make_sp<T>(hp, table);
    const size_t esz = output == OUT_CWT ? sizeof(cx<T>) : sizeof(T);
    if (hp.path == 0 && hp.short3 && output != OUT_CWT && !(g_mode & 2) && !(g_mode & 128)) {
        // resampled short rows (nw_kernels4.cuh): mirrors nwcwt.cu: ensure_device_short3 / launch_short3_t
        std::vector<cx<T>> tw;
        fill_tw<T>(tw, hp.N, hp.N, 1);
        const size_t ng = hp.groups.size();
        std::vector<Short3Group<T>> gs(ng);
        std::vector<HostPlan> subs(ng);
        std::vector<std::vector<T>> wtabs(ng), coefs(ng);
        std::vector<std::vector<cx<T>>> tws(ng);
        std::vector<std::vector<int>> poss(ng);
        int unit = 0;
        for (size_t gi = 0; gi < ng; ++gi) {
            const MrGroup& mg = hp.groups[gi];
            subs[gi] = *mg.sub;
            if (!build_weight_table<T>(subs[gi], mg.D > 1 ? mg.eq.data() : nullptr, (size_t)256 << 20, wtabs[gi])) return -2;
            Short3Group<T>& g = gs[gi];
            if (!short3_fill_group<T>(g, hp.N, subs[gi].N, mg.D, mg.K, subs[gi].F)) return -2;
            g.unit0 = unit; unit += g.nunits;
            g.st = subs[gi].stS;
            fill_tw<T>(tws[gi], g.M, g.M, 1);
            coefs[gi].assign(mg.coef.size() + 4, (T)0);
            if (mg.D > 1) resample_coefq<T>(mg.coef.data(), mg.D, mg.K, g.PQ, coefs[gi].data());
            g.tw = tws[gi].data(); g.rec = subs[gi].rec.data(); g.wtab = wtabs[gi].data(); g.coefq = coefs[gi].data();
            g.fmap = mg.fidx.data();
            poss[gi].resize((size_t)g.M);
            for (int k = 0; k < g.M; ++k) poss[gi][(size_t)k] = fft2_dit_pos(g.st, k);
            g.ditpos = poss[gi].data();
        }
        Short3Params<T> P;
        memset(&P, 0, sizeof(P));
        P.signals = (const T*)signals; P.out = out; P.N = (int)hp.N; P.F_out = hp.F; P.S = (int)S;
        P.out_mode = output; P.bl_mode = bl; P.bl_lo = (int)blo; P.bl_hi = (int)bhi; P.st = hp.stS; P.tw = tw.data();
        P.groups = gs.data(); P.ngroups = (int)ng; P.nunits = unit;
        P.fsplit = unit < 3 ? unit : 3;
        std::vector<char> sm(hp.smem_S3 + 64);
        char* smp = (char*)(((uintptr_t)sm.data() + 31) & ~(uintptr_t)31);
        const int nt = (g_mode & 4) ? 1 : hp.nthrS3;
        const long long nblk = ((S + 1) / 2) * P.fsplit;
        for (long long b = 0; b < nblk; ++b)
            Fibers::get().run(nt, [&](int t) {
                short3_body<T>(P, smp, (int)b, t, nt);
            });
        return 0;
    }
    if (hp.path == 0 && hp.short2 && !(g_mode & 2)) {
        std::vector<cx<T>> tw;
        fill_tw<T>(tw, hp.N, hp.N, 1);
        Short2Params<T> P;
        memset(&P, 0, sizeof(P));
        P.signals = (const T*)signals; P.out = out; P.N = (int)hp.N; P.F = hp.F; P.S = (int)S; P.tpsh = hp.tpshS;
        P.out_mode = output; P.bl_mode = bl; P.bl_lo = (int)blo; P.bl_hi = (int)bhi; P.st = hp.stS; P.tw = tw.data(); P.sp = sp;
        const int ngroups = (hp.F + (1 << hp.tpshS) - 1) >> hp.tpshS;
        P.fsplit = ngroups < 3 ? ngroups : 3;
        std::vector<char> sm(hp.smem_S2 + 64);
        char* smp = (char*)(((uintptr_t)sm.data() + 31) & ~(uintptr_t)31);
        const int nt = (g_mode & 4) ? 1 : hp.nthrS2;
        const int spS = (g_mode & 8) ? 0 : static_plan_id(hp.stS, hp.tpshS);
        const long long nblk = ((S + 1) / 2) * P.fsplit;
        for (long long b = 0; b < nblk; ++b)
            Fibers::get().run(nt, [&](int t) {
                if (output == OUT_POWER) { if (spS == 6) short2_body<T, OUT_POWER, 6>(P, smp, (int)b, t, nt); else if (spS == 7) short2_body<T, OUT_POWER, 7>(P, smp, (int)b, t, nt); else short2_body<T, OUT_POWER, 0>(P, smp, (int)b, t, nt); }
                else if (output == OUT_ABS) short2_body<T, OUT_ABS, 0>(P, smp, (int)b, t, nt);
                else { if (spS == 6) short2_body<T, OUT_CWT, 6>(P, smp, (int)b, t, nt); else short2_body<T, OUT_CWT, 0>(P, smp, (int)b, t, nt); }
            });
        return 0;
    }
    if (hp.path == 0) {
        std::vector<cx<T>> tw;
        fill_tw<T>(tw, hp.N, hp.N, 1);
        ShortParams<T> P;
        memset(&P, 0, sizeof(P));
        P.signals = (const T*)signals; P.out = out; P.N = (int)hp.N; P.F = hp.F; P.S = (int)S;
        P.tsh = hp.tsh; P.pitch = hp.pitch; P.out_mode = output; P.bl_mode = bl; P.bl_lo = (int)blo; P.bl_hi = (int)bhi;
        P.st = hp.st; P.tw = tw.data(); P.sp = sp;
        const int TT = 1 << hp.tsh, ngroups = (hp.F + TT - 1) / TT;
        P.fsplit = ngroups < 3 ? ngroups : 3;
        std::vector<char> smem(hp.smem_short + 64);
        for (long long b = 0; b < S * P.fsplit; ++b) short_body<T>(P, smem.data(), (int)b, 0, 1);
        return 0;
    }
    std::vector<cx<T>> twA, twB, twH, twL;
    if (hp.generic_ok) {
        fill_tw<T>(twA, hp.N1, hp.N1, 1);
        fill_tw<T>(twB, hp.N2, hp.N2, 1);
    }
    const long long nL = 1LL << hp.lb, nH = (hp.N + nL - 1) / nL;
    fill_tw<T>(twL, nL, hp.N, 1);
    fill_tw<T>(twH, nH, hp.N, nL);
    LongParams<T> P;
    memset(&P, 0, sizeof(P));
    P.N = hp.N; P.N1 = hp.N1; P.N2 = hp.N2; P.tshA = hp.tshA; P.pitchA = hp.pitchA; P.tshB = hp.tshB;
    P.stA = hp.stA; P.stB = hp.stB; P.twA = twA.data(); P.twB = twB.data(); P.twH = twH.data(); P.twL = twL.data();
    P.lb = hp.lb; P.tm_stride = hp.tm_stride; P.sp = sp;
    const int ring = hp.ring < 3 ? hp.ring : 3;
    std::vector<cx<T>> X((size_t)ring * hp.N), Tm((size_t)ring * hp.tm_stride);
    P.Tm = Tm.data();
    const int TA = 1 << hp.tshA, TB = 1 << hp.tshB;
    const int tilesA = hp.generic_ok ? (hp.N2 + TA - 1) / TA : 0, tilesB = hp.generic_ok ? (hp.N1 + TB - 1) / TB : 0;
    std::vector<char> smem(std::max(hp.smem_A, hp.smem_B) + 64);
    for (long long s0 = 0; s0 < S; s0 += ring) {
        const int gs = (int)std::min<long long>(ring, S - s0);
        if (hp.fast && !(g_mode & 2) && !(g_mode & 32)) {
            std::vector<cx<T>> twA2, twB2;
            fill_tw<T>(twA2, hp.N1f, hp.N1f, 1);
            fill_tw<T>(twB2, hp.N2f, hp.N2f, 1);
            Long2Params<T> Q;
            memset(&Q, 0, sizeof(Q));
            Q.N = hp.N; Q.xstride = hp.N; Q.N1 = hp.N1f; Q.N2 = hp.N2f; Q.F = hp.F; Q.tpshA = hp.tpshA; Q.tpshB = hp.tpshB;
            Q.stA = hp.stA2; Q.stB = hp.stB2; Q.twA = twA2.data(); Q.twB = twB2.data(); Q.twH = twH.data(); Q.twL = twL.data();
            Q.lb = hp.lb; Q.tm_stride = hp.tm_stride2; Q.out_mode = OUT_CWT;
            std::vector<cx<T>> Tmf((size_t)hp.tm_stride2);
            Q.Tm = Tmf.data();
            const int tA = (hp.N2f + (2 << hp.tpshA) - 1) / (2 << hp.tpshA), tB = (hp.N1f + (2 << hp.tpshB) - 1) / (2 << hp.tpshB);
            std::vector<char> sm2(std::max(hp.smem_A2, hp.smem_B2) + 64);
            char* smp = (char*)(((uintptr_t)sm2.data() + 31) & ~(uintptr_t)31);
            const int ntA = (g_mode & 4) ? 1 : hp.nthrA2, ntB = (g_mode & 4) ? 1 : hp.nthrB2;
            for (int y = 0; y < gs; ++y) {
                Q.signal = (const T*)signals + (size_t)(s0 + y) * hp.N;
                Q.out = X.data() + (size_t)y * hp.N;
                Q.row0 = 0;
                for (int x = 0; x < tA; ++x) Fibers::get().run(ntA, [&](int t) { passA2f_body<T>(Q, smp, x, 0, t, ntA); });
                for (int x = 0; x < tB; ++x) Fibers::get().run(ntB, [&](int t) { passB2_body<T, OUT_CWT, 0, -1>(Q, smp, x, 0, t, ntB); });
            }
        } else {
        P.signal = (const T*)signals + (size_t)s0 * hp.N;
        P.Xout = X.data();
        for (int y = 0; y < gs; ++y) for (int x = 0; x < tilesA; ++x) passA_body<T, -1>(P, smem.data(), x, y, 0, 1);
        for (int y = 0; y < gs; ++y) for (int x = 0; x < tilesB; ++x) passB_body<T, -1>(P, smem.data(), x, y, 0, 1);
        }
        if (hp.fast && !(g_mode & 2)) {
            void* out_s0 = (char*)out + (size_t)s0 * hp.F * (size_t)hp.N * esz;
            const long long rows = (long long)gs * hp.F;
            int rc = 0;
            if (!hp.groups.empty() && output != OUT_CWT && !(g_mode & 128)) {
                for (const MrGroup& mg : hp.groups)
                    if ((rc = inverse_rows<T>(hp, *mg.sub, &mg, X.data(), out_s0, gs, output))) return rc;
            } else if ((rc = inverse_rows<T>(hp, hp, nullptr, X.data(), out_s0, gs, output))) return rc;
            if (bl != BL_NONE) {
                double sh[2];
                for (long long r = 0; r < rows; ++r) baseline_rows_body<T>((T*)out_s0, hp.N, bl, (int)blo, (int)bhi, sh, (int)r, 0, 1);
            }
            continue;
        }
        for (int si = 0; si < gs; ++si) {
            P.X = X.data() + (size_t)si * hp.N;
            char* out_s = (char*)out + (size_t)(s0 + si) * hp.F * (size_t)hp.N * esz;
            for (int f0 = 0; f0 < hp.F; f0 += ring) {
                const int g = std::min(ring, hp.F - f0);
                P.f0 = f0; P.out = out_s + (size_t)f0 * hp.N * esz; P.out_mode = output;
                for (int y = 0; y < g; ++y) for (int x = 0; x < tilesA; ++x) passA_body<T, 1>(P, smem.data(), x, y, 0, 1);
                for (int y = 0; y < g; ++y) for (int x = 0; x < tilesB; ++x) passB_body<T, 1>(P, smem.data(), x, y, 0, 1);
            }
            if (bl != BL_NONE) {
                double sh[2];
                for (int f = 0; f < hp.F; ++f) baseline_rows_body<T>((T*)out_s, hp.N, bl, (int)blo, (int)bhi, sh, f, 0, 1);
            }
        }
    }
    return 0;
}

extern "C" int emul_transform(const nwcwt_plan_desc* d, const void* signals, void* out, long long S, int output,
                              int bl, long long blo, long long bhi, int force_long, char* errbuf, int errlen) {
    HostPlan hp;
    hp.device = 0; hp.dtype = d->dtype; hp.family = d->family; hp.interpolate = d->interpolate ? 1 : 0;
    hp.N = d->n; hp.F = d->n_freqs; hp.sfreq = d->sfreq; hp.p0 = d->p0; hp.p1 = d->p1; hp.p2 = d->p2;
    hp.prune_eps = d->prune_eps < 0 ? (d->dtype == 0 ? 1e-12 : 1e-24) : d->prune_eps;
    hp.resample = d->resample >= 0 ? 1 : 0;
    hp.resample_tol = d->resample_tol;
    hp.freqs.assign(d->freqs, d->freqs + d->n_freqs);
    if (d->family == FAM_MORLET) hp.aux.assign(d->aux, d->aux + d->n_freqs);
    if (d->family == FAM_TABLE) {
        hp.table_len = d->table_len;
        hp.table.assign(d->table, d->table + 2 * (size_t)d->n_freqs * (size_t)d->table_len);
        if (d->table_lens) hp.table_lens.assign(d->table_lens, d->table_lens + d->n_freqs);
    }
    plan_geometry(hp);
    plan_bands(hp);
    std::string err;
    g_mode = force_long;
    if (!plan_shape(hp, err, (force_long & 1) != 0)) {
        strncpy(errbuf, err.c_str(), errlen - 1);
        return -2;
    }
    if (bhi > hp.N) bhi = hp.N;
    if (blo > bhi) blo = bhi;
    return d->dtype == 0 ? run<float>(hp, signals, out, S, output, bl, blo, bhi)
                         : run<double>(hp, signals, out, S, output, bl, blo, bhi);
}
