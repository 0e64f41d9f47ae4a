#include <stdio.h>
#include <complex>
#include <vector>
#include "../../ninwavelets_b200/csrc/nw_bfly2.cuh"
using namespace nw;
template <typename T, int R, int DIR> double test() {
    cx2<T> v[R];
    std::complex<double> in0[R], in1[R];
    for (int i = 0; i < R; ++i) {
        in0[i] = {sin(i * 1.3 + 0.2), cos(i * 0.7 + 1.1)};
        in1[i] = {cos(i * 2.1 + 0.5), sin(i * 0.3 - 0.4)};
        v[i] = mk2<T>(mk<T>((T)in0[i].real(), (T)in0[i].imag()), mk<T>((T)in1[i].real(), (T)in1[i].imag()));
    }
    B2<T, R, DIR>::run(v);
    double err = 0;
    for (int k = 0; k < R; ++k) {
        std::complex<double> a0 = 0, a1 = 0;
        for (int n = 0; n < R; ++n) {
            std::complex<double> w = std::polar(1.0, DIR * 2 * M_PI * n * k / R);
            a0 += in0[n] * w; a1 += in1[n] * w;
        }
        err = std::max(err, std::abs(a0 - std::complex<double>(lane0(v[k]).x, lane0(v[k]).y)));
        err = std::max(err, std::abs(a1 - std::complex<double>(lane1(v[k]).x, lane1(v[k]).y)));
    }
    return err;
}
#define TST(R) printf("R=%2d  f32 %+d: %.2e %.2e   f64: %.2e %.2e\n", R, 1, test<float, R, 1>(), test<float, R, -1>(), test<double, R, 1>(), test<double, R, -1>());
int main() { TST(2) TST(3) TST(4) TST(5) TST(6) TST(8) TST(10) TST(12) TST(15) TST(16) TST(20) TST(24) TST(25) TST(30) TST(32) return 0; }
