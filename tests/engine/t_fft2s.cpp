#include <stdio.h>
#include <complex>
#include <vector>
#include "../../ninwavelets_b200/csrc/nw_fft2.cuh"
#include "../../ninwavelets_b200/csrc/nw_plan.h"
using namespace nw;
typedef std::complex<double> cd;
template <typename T> struct VecDst { cd* y; int P;
    struct Ctx { int base, step, tp; };
    Ctx begin(int base, int step, int tp) const { return Ctx{base, step, tp}; }
    template <int R> void store_all(const Ctx& x, const cx2<T>* v) const { for (int q = 0; q < R; ++q) store(x.base + q * x.step, x.tp, v[q]); }
    void store(int p, int tp, cx2<T> v) const { y[(2*tp)*P + p] = cd(lane0(v).x, lane0(v).y); y[(2*tp+1)*P+p] = cd(lane1(v).x, lane1(v).y); } };
template <typename T, int TPS, int P, int R0, int R1, int R2> void run(bool dit) {
    Fft2Plan st; plan_packed(P, st);
    printf("P=%d plan:", P); for (int i = 0; i < st.nst; ++i) printf(" %d", st.radix[i]); printf(" static: %d %d %d  ", R0, R1, R2);
    const int TP = 1 << TPS, TT = 2 * TP;
    std::vector<cd> x(TT * P), y(TT * P);
    for (int i = 0; i < TT * P; ++i) x[i] = cd(sin(i * 0.37 + 1), cos(i * 0.11 + i % 7));
    std::vector<cx<T>> tw(P);
    for (int j = 0; j < P; ++j) tw[j] = mk<T>((T)cos(2 * M_PI * j / P), (T)sin(2 * M_PI * j / P));
    std::vector<cx2<T>> buf(P * TP);
    VecDst<T> dst{y.data(), P};
    for (int p = 0; p < P; ++p) for (int tp = 0; tp < TP; ++tp) {
        cd a = x[(2*tp)*P + p], b = x[(2*tp+1)*P + p];
        cx2<T> v = mk2<T>(mk<T>((T)a.real(), (T)a.imag()), mk<T>((T)b.real(), (T)b.imag()));
        buf[(dit ? fft2_dit_pos(st, p) : p) * TP + tp] = v;
    }
    if (dit) fft2_dit_static<T, 1, TPS, P, R0, R1, R2>(tw.data(), buf.data(), dst, 0, 1);
    else fft2_dif_static<T, 1, false, TPS, P, R0, R1, R2>(tw.data(), buf.data(), dst, 0, 1);
    double err = 0, mag = 0;
    for (int t = 0; t < TT; ++t) for (int k = 0; k < P; k += 7) {
        cd a = 0; for (int n = 0; n < P; ++n) a += x[t * P + n] * std::polar(1.0, 2 * M_PI * ((long long)n * k % P) / P);
        err = std::max(err, std::abs(a - y[t * P + k])); mag = std::max(mag, std::abs(a));
    }
    printf("%s err %.2e\n", dit ? "dit" : "dif", err / mag);
}
int main() {
    run<float, 2, 1000, 10, 10, 10>(false); run<double, 1, 1000, 10, 10, 10>(true);
    run<float, 2, 600, 12, 10, 5>(true); run<double, 2, 600, 12, 10, 5>(false);
    run<float, 2, 256, 16, 16, 1>(true); run<float, 2, 256, 16, 16, 1>(false);
    run<float, 2, 1024, 16, 16, 4>(true); run<float, 2, 1024, 16, 16, 4>(false);
}
