// Host harness for csrc/regression_terms.cuh: evaluates the per-element terms of a regression op on the CPU.
//   usage: reg_terms_host <op> <param> <eps> <f32|f64>   then lines "pred target" on stdin -> the op's terms, one line each
// Built and driven by tests/test_reg_terms_host.py (nvcc, host code only; no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../regression_terms.cuh"

template <typename F, bool kTweedie>
static void run(int op, double param, double eps) {
    double p, t;
    const int k = mb200::reg_num_sums(op);
    while (std::scanf("%lf %lf", &p, &t) == 2) {
        F out[mb200::kRegMaxK] = {0, 0, 0, 0};
        mb200::reg_terms<F, kTweedie>(op, (F)p, (F)t, (F)param, (F)eps, out);
        for (int i = 0; i < k; ++i) std::printf(i ? " %.17g" : "%.17g", (double)out[i]);
        std::printf("\n");
    }
}

int main(int argc, char** argv) {
    if (argc != 5) return 2;
    const int op = std::atoi(argv[1]);
    const double param = std::atof(argv[2]), eps = std::atof(argv[3]);
    const bool f64 = std::strcmp(argv[4], "f64") == 0;
    if (op == mb200::REG_TWEEDIE) {
        f64 ? run<double, true>(op, param, eps) : run<float, true>(op, param, eps);
    } else {
        f64 ? run<double, false>(op, param, eps) : run<float, false>(op, param, eps);
    }
    return 0;
}
