// Standalone C-ABI smoke test (no Python, no torch): exercises K9 op 10 (Tweedie deviance) and its neighbours through
// libmetrics_b200.so on a real GPU and compares with a host evaluation in double precision.  Starts in about a second, so
// it fits in whatever GPU time is left.    build:  nvcc -O2 -std=c++17 -I../../include tools/abi_smoke.cu -o build/abi_smoke
//                                                         -L../_lib -lmetrics_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../_lib'
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "metrics_b200.h"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

static double tweedie(double p, double t, double power) {
    if (power == 1.0) return 2 * ((t == 0 ? 0.0 : t * std::log(t / p)) + p - t);
    if (power == 2.0) return 2 * (std::log(p / t) + t / p - 1);
    const double a = 1 - power, b = 2 - power;
    return 2 * (std::pow(std::fmax(t, 0.0), b) / (a * b) - t * std::pow(p, a) / a + std::pow(p, b) / b);
}

template <typename T>
static int run_case(const char* name, int dtype, long long n, long long d, int op, double power, double rtol, bool inject) {
    std::vector<T> hp(n * d), ht(n * d);
    unsigned s = 12345u + (unsigned)(n * 31 + d * 7 + op);
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (double)(1u << 24); };
    for (long long i = 0; i < n * d; ++i) hp[i] = (T)(rnd() * 4 + 0.05), ht[i] = (T)(rnd() * 4 + 0.05);
    long long bad_p = 0, neg_t = 0, zero_t = 0;
    if (inject) {  // out-of-domain elements for the census (the deviance sum becomes NaN / inf: only the counts are checked)
        for (long long i = 0; i < n * d; i += 1001) hp[i] = (T)0, ++bad_p;
        for (long long i = 7; i < n * d; i += 1003) ht[i] = (T)-1, ++neg_t;
        for (long long i = 13; i < n * d; i += 1009) ht[i] = (T)0, ++zero_t;
    }
    T *dp, *dt;
    CK(cudaMalloc(&dp, sizeof(T) * n * d));
    CK(cudaMalloc(&dt, sizeof(T) * n * d));
    CK(cudaMemcpy(dp, hp.data(), sizeof(T) * n * d, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dt, ht.data(), sizeof(T) * n * d, cudaMemcpyHostToDevice));
    const int k = mb200_regression_num_sums(op);
    const long long scratch_n = mb200_regression_scratch_doubles(n, d, op);
    double *dout, *dscratch;
    CK(cudaMalloc(&dout, sizeof(double) * k * d));
    CK(cudaMalloc(&dscratch, sizeof(double) * scratch_n));
    const int rc = mb200_regression_sums(dp, dt, dtype, n, d, op, power, 0.0, dout, dscratch, nullptr);
    if (rc != 0) {
        std::printf("FAIL %s: rc=%d (%s)\n", name, rc, mb200_last_error());
        return 1;
    }
    CK(cudaDeviceSynchronize());
    std::vector<double> out(k * d);
    CK(cudaMemcpy(out.data(), dout, sizeof(double) * k * d, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (long long c = 0; c < d; ++c) {
        double want = 0, wp = 0, wn = 0, wz = 0;
        for (long long r = 0; r < n; ++r) {
            const double p = (double)hp[r * d + c], t = (double)ht[r * d + c];
            want += op == 10 ? tweedie(p, t, power) : (p - t) * (p - t);
            wp += p <= 0, wn += t < 0, wz += t == 0;
        }
        const double got = out[0 * d + c];
        if (!inject && !(std::fabs(got - want) <= rtol * std::fabs(want))) ++bad;
        if (op == 10 && (out[1 * d + c] != wp || out[2 * d + c] != wn || out[3 * d + c] != wz)) ++bad;
        if (c == 0)
            std::printf("%s %s: sum got %.10g want %.10g (rel %.2e)%s census got [%g %g %g] want [%g %g %g]\n", bad ? "FAIL" : "PASS",
                        name, got, want, std::fabs(got - want) / std::fabs(want), inject ? " (not compared: injected)" : "",
                        op == 10 ? out[1 * d] : 0.0, op == 10 ? out[2 * d] : 0.0, op == 10 ? out[3 * d] : 0.0, wp, wn, wz);
    }
    cudaFree(dp), cudaFree(dt), cudaFree(dout), cudaFree(dscratch);
    return bad ? 1 : 0;
}

int main() {
    std::printf("abi %d\n", mb200_abi_version());
    int fails = 0;
    fails += run_case<float>("mse f32 flat (control)", MB200_F32, 1 << 20, 1, 0, 0.0, 2e-6, false);
    fails += run_case<float>("tweedie p=1.5 f32 flat", MB200_F32, 1 << 20, 1, 10, 1.5, 2e-5, false);
    fails += run_case<float>("tweedie p=1 f32 flat", MB200_F32, (1 << 20) + 3, 1, 10, 1.0, 2e-5, false);
    fails += run_case<float>("tweedie p=2 f32 flat", MB200_F32, 4097, 1, 10, 2.0, 2e-5, false);
    fails += run_case<float>("tweedie p=-1.5 f32 flat", MB200_F32, 70001, 1, 10, -1.5, 2e-5, false);
    fails += run_case<double>("tweedie p=3 f64 flat", MB200_F64, 1 << 18, 1, 10, 3.0, 1e-11, false);
    fails += run_case<double>("tweedie p=1.5 f64 flat", MB200_F64, 33333, 1, 10, 1.5, 1e-11, false);
    fails += run_case<float>("tweedie p=2 f32 [n,3] columns", MB200_F32, 50000, 3, 10, 2.0, 2e-5, false);
    fails += run_case<double>("tweedie p=1 f64 [n,300] columns", MB200_F64, 2000, 300, 10, 1.0, 1e-11, false);
    fails += run_case<float>("tweedie census f32", MB200_F32, 1 << 20, 1, 10, 2.0, 0, true);
    fails += run_case<double>("tweedie census f64 [n,5]", MB200_F64, 100000, 5, 10, 1.5, 0, true);
    std::printf("%s: %d failing case(s), %llu kernel launches\n", fails ? "ABI_SMOKE_FAIL" : "ABI_SMOKE_OK", fails,
                (unsigned long long)mb200_launch_count());
    return fails ? 1 : 0;
}
