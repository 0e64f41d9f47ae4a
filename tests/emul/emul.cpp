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

template <typename T>
static int run(const HostPlan& hp, const void* signals, void* out, long long S, int output, int bl, long long blo,
               long long bhi) {
    std::vector<cx<T>> table;
    if (hp.family == FAM_TABLE) {
        table.resize((size_t)hp.F * hp.table_len);
        for (size_t i = 0; i < table.size(); ++i) { table[i].x = (T)hp.table[2 * i]; table[i].y = (T)hp.table[2 * i + 1]; }
    }
    SpecParams<T> sp;
    sp.family = hp.family; sp.grid_off = hp.grid_off; sp.df = hp.df; sp.p0 = hp.p0; sp.p1 = hp.p1;
    sp.p2 = hp.family == FAM_MORSE ? hp.p0 / hp.p1 : hp.p2;
    sp.norm = (T)(1.0 / (double)hp.N);
    sp.rec = hp.rec.data(); sp.table = table.data(); sp.table_len = hp.table_len;
    const size_t esz = output == OUT_CWT ? sizeof(cx<T>) : sizeof(T);
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
            Q.N = hp.N; Q.N1 = hp.N1f; Q.N2 = hp.N2f; Q.F = hp.F; Q.tpshA = hp.tpshA; Q.tpshB = hp.tpshB;
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
            std::vector<cx<T>> twA2, twB2;
            fill_tw<T>(twA2, hp.N1f, hp.N1f, 1);
            fill_tw<T>(twB2, hp.N2f, hp.N2f, 1);
            Long2Params<T> Q;
            memset(&Q, 0, sizeof(Q));
            Q.N = hp.N; Q.N1 = hp.N1f; Q.N2 = hp.N2f; Q.F = hp.F; Q.tpshA = hp.tpshA; Q.tpshB = hp.tpshB;
            Q.stA = hp.stA2; Q.stB = hp.stB2; Q.twA = twA2.data(); Q.twB = twB2.data(); Q.twH = twH.data(); Q.twL = twL.data();
            Q.lb = hp.lb; Q.tm_stride = hp.tm_stride2; Q.sp = sp;
            const int ring2 = hp.ring2 < 3 ? hp.ring2 : 3;
            std::vector<cx<T>> Tm2((size_t)ring2 * hp.tm_stride2);
            Q.X = X.data(); Q.Tm = Tm2.data();
            Q.out = (char*)out + (size_t)s0 * hp.F * (size_t)hp.N * esz; Q.out_mode = output;
            const int tA = (hp.N2f + (2 << hp.tpshA) - 1) / (2 << hp.tpshA), tB = (hp.N1f + (2 << hp.tpshB) - 1) / (2 << hp.tpshB);
            std::vector<char> sm2(std::max(hp.smem_A2, hp.smem_B2) + 64);
            char* smp = (char*)(((uintptr_t)sm2.data() + 31) & ~(uintptr_t)31);
            const int ntA = (g_mode & 4) ? 1 : hp.nthrA2, ntB = (g_mode & 4) ? 1 : hp.nthrB2;
            const long long rows = (long long)gs * hp.F;
            Q.pplans = hp.pplans.data();
            if (hp.stA2.nst >= 1) Q.dstepA = make_fastdiv((uint32_t)std::max(1, hp.N1f / hp.stA2.radix[hp.stA2.nst - 1]));
            const int spA = (hp.pruneA && !(g_mode & 16)) ? -1 : (g_mode & 8) ? 0 : static_plan_id(hp.stA2, hp.tpshA), spB = (g_mode & 8) ? 0 : static_plan_id(hp.stB2, hp.tpshB);
            for (long long r0 = 0; r0 < rows; r0 += ring2) {
                const int g = (int)std::min<long long>(ring2, rows - r0);
                Q.row0 = (int)r0;
                for (int y = 0; y < g; ++y) for (int x = 0; x < tA; ++x)
                    Fibers::get().run(ntA, [&](int t) {
                        if (spA == -1) passA2p_body<T>(Q, smp, x, y, t, ntA);
                        else if (group_narrow(hp, r0, g) && !(g_mode & 64)) {
                            if (spA == 2) passA2_body<T, 2, true>(Q, smp, x, y, t, ntA);
                            else passA2_body<T, 0, true>(Q, smp, x, y, t, ntA);
                        } else {
                            if (spA == 2) passA2_body<T, 2>(Q, smp, x, y, t, ntA);
                            else if (spA == 4) passA2_body<T, 4>(Q, smp, x, y, t, ntA);
                            else if (spA == 11) passA2_body<T, 11>(Q, smp, x, y, t, ntA);
                            else passA2_body<T, 0>(Q, smp, x, y, t, ntA);
                        }
                    });
                for (int y = 0; y < g; ++y) for (int x = 0; x < tB; ++x)
                    Fibers::get().run(ntB, [&](int t) {
                        if (output == OUT_POWER) { if (spB == 1) passB2_body<T, OUT_POWER, 1>(Q, smp, x, y, t, ntB); else if (spB == 4) passB2_body<T, OUT_POWER, 4>(Q, smp, x, y, t, ntB); else if (spB == 8) passB2_body<T, OUT_POWER, 8>(Q, smp, x, y, t, ntB); else if (spB == 9) passB2_body<T, OUT_POWER, 9>(Q, smp, x, y, t, ntB); else if (spB == 10) passB2_body<T, OUT_POWER, 10>(Q, smp, x, y, t, ntB); else passB2_body<T, OUT_POWER, 0>(Q, smp, x, y, t, ntB); }
                        else if (output == OUT_ABS) passB2_body<T, OUT_ABS, 0>(Q, smp, x, y, t, ntB);
                        else { if (spB == 1) passB2_body<T, OUT_CWT, 1>(Q, smp, x, y, t, ntB); else passB2_body<T, OUT_CWT, 0>(Q, smp, x, y, t, ntB); }
                    });
            }
            if (bl != BL_NONE) {
                double sh[2];
                for (long long r = 0; r < rows; ++r) baseline_rows_body<T>((T*)Q.out, hp.N, bl, (int)blo, (int)bhi, sh, (int)r, 0, 1);
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
    setenv("NWCWT_PRUNE", (force_long & 16) ? "0" : "1", 1);   // the emulation exercises the pruned pass A by default
    if (!plan_shape(hp, err, (force_long & 1) != 0)) {
        strncpy(errbuf, err.c_str(), errlen - 1);
        return -2;
    }
    if (bhi > hp.N) bhi = hp.N;
    if (blo > bhi) blo = bhi;
    return d->dtype == 0 ? run<float>(hp, signals, out, S, output, bl, blo, bhi)
                         : run<double>(hp, signals, out, S, output, bl, blo, bhi);
}
