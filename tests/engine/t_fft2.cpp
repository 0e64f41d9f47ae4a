#include <stdio.h>
#include <complex>
#include <vector>
#include "../../ninwavelets_b200/csrc/nw_fft2.cuh"
#include "../../ninwavelets_b200/csrc/nw_plan.h"
using namespace nw;
typedef std::complex<double> cd;
template <typename T> struct VecSrc { const cd* x; int P; int TT;   // x[t*P + p]
    template <int R> void load_all(int base, int step, int tp, cx2<T>* v) const { for (int q = 0; q < R; ++q) v[q] = load(base + q * step, tp); }
    cx2<T> load(int p, int tp) const { cd a = x[(2*tp)*P + p], b = x[(2*tp+1)*P + p]; return mk2<T>(mk<T>((T)a.real(), (T)a.imag()), mk<T>((T)b.real(), (T)b.imag())); } };
template <typename T> struct VecDst { cd* y; int P;
    struct Ctx { int base, step, tp; };
    Ctx begin(int base, int step, int tp) const { return Ctx{base, step, tp}; }
    template <int R> void store_all(const Ctx& x, const cx2<T>* v) const { for (int q = 0; q < R; ++q) store(x.base + q * x.step, x.tp, v[q]); }
    void store(int p, int tp, cx2<T> v) const { y[(2*tp)*P + p] = cd(lane0(v).x, lane0(v).y); y[(2*tp+1)*P+p] = cd(lane1(v).x, lane1(v).y); } };
template <typename T, int DIR> double run(int P, int tpsh, bool dit, int nthr) {
    Fft2Plan st; if (!plan_packed(P, st)) return -1;
    int TP = 1 << tpsh, TT = 2 * TP;
    std::vector<cd> x(TT * P), y(TT * P), ref(TT * P);
    for (int i = 0; i < TT * P; ++i) x[i] = cd(sin(i * 0.37 + 1), cos(i * 0.11 + i % 7));
    std::vector<cx<T>> tw(P);
    for (int j = 0; j < P; ++j) tw[j] = mk<T>((T)cos(2 * M_PI * j / P), (T)sin(2 * M_PI * j / P));
    std::vector<cx2<T>> buf(P * TP);
    VecSrc<T> src{x.data(), P, TT}; VecDst<T> dst{y.data(), P};
    if (dit) {
        // emulate nthr threads stage by stage is not possible from outside; run with one thread (barriers are no-ops)
        fft2_dit<T, DIR>(st, tpsh, tw.data(), buf.data(), src, dst, 0, 1);
    } else {
        for (int p = 0; p < P; ++p) for (int tp = 0; tp < TP; ++tp) buf[p * TP + tp] = src.load(p, tp);
        fft2_dif<T, DIR, false>(st, tpsh, tw.data(), buf.data(), dst, 0, 1);
    }
    double err = 0, mag = 0;
    for (int t = 0; t < TT; ++t) for (int k = 0; k < P; k += (P > 512 ? 37 : 1)) {
        cd a = 0; for (int n = 0; n < P; ++n) a += x[t * P + n] * std::polar(1.0, DIR * 2 * M_PI * ((long long)n * k % P) / P);
        err = std::max(err, std::abs(a - y[t * P + k])); mag = std::max(mag, std::abs(a));
    }
    return err / mag;
}
int main() {
    int Ps[] = {2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 24, 25, 30, 60, 100, 120, 150, 250, 256, 300, 375, 500, 600, 750, 1000, 1024, 1500, 4096, 8192};
    for (int P : Ps) {
        Fft2Plan st; plan_packed(P, st);
        printf("P=%5d radices", P); for (int i = 0; i < st.nst; ++i) printf(" %d", st.radix[i]);
        printf("  dif f32 %.1e f64 %.1e (fwd %.1e) | dit f32 %.1e f64 %.1e (fwd %.1e)\n", run<float, 1>(P, 1, false, 1), run<double, 1>(P, 0, false, 1), run<double, -1>(P, 2, false, 1),
               run<float, 1>(P, 1, true, 1), run<double, 1>(P, 0, true, 1), run<double, -1>(P, 2, true, 1));
    }
}
