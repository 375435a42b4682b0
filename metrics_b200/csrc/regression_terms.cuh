// Per-element terms of the regression running sums (K9).  Shared by csrc/regression.cu (device) and by
// csrc/tools/reg_terms_host.cu, a host harness that lets the CPU test-suite check the very same formulas against the
// reference's values without a GPU (tests/test_reg_terms_host.py).
#pragma once
#include <cmath>

namespace mb200 {

enum RegOp { REG_MSE = 0, REG_MAE, REG_MAPE, REG_SMAPE, REG_WMAPE, REG_MSLE, REG_LOGCOSH, REG_MINKOWSKI, REG_R2, REG_EXPVAR,
             REG_TWEEDIE, REG_LAST = REG_TWEEDIE };
constexpr int kRegMaxK = 4;

__host__ __device__ inline int reg_num_sums(int op) {
    switch (op) {
        case REG_WMAPE: return 2;
        case REG_R2: return 3;
        case REG_EXPVAR: return 4;
        case REG_TWEEDIE: return 4;
        default: return 1;
    }
}

// kTweedie selects the Tweedie-deviance terms at compile time: they live in their own kernel instantiations, so the code
// (and register allocation) of the ten running-sum ops above the line is exactly what it was before op 10 existed.
template <typename F, bool kTweedie = false>
__host__ __device__ __forceinline__ void reg_terms(int op, F p, F t, F param, F eps, F (&out)[kRegMaxK]) {
    if constexpr (kTweedie) {
        // Tweedie deviance of power `param` (functional/regression/tweedie_deviance.py:44-78; power 0 is REG_MSE, powers
        // in (0, 1) are rejected by the caller).  The reference validates the domain with separate `torch.any` passes
        // (:51, :59, :65-75); here the same pass counts the offending elements instead:
        // out[1] = #(preds <= 0), out[2] = #(target < 0), out[3] = #(target == 0).
        F dev;
        if (param == (F)1) {  // Poisson: 2 (t log(t/p) + p - t), with 0 log 0 = 0
            const F xlogy = (t == (F)0) ? (F)0 : t * log(t / p);
            dev = (F)2 * (xlogy + p - t);
        } else if (param == (F)2) {  // Gamma
            dev = (F)2 * (log(p / t) + t / p - (F)1);
        } else {
            const F a = (F)1 - param, b = (F)2 - param;
            dev = (F)2 * (pow(fmax(t, (F)0), b) / (a * b) - t * pow(p, a) / a + pow(p, b) / b);
        }
        out[0] = dev;
        out[1] = p <= (F)0 ? (F)1 : (F)0;
        out[2] = t < (F)0 ? (F)1 : (F)0;
        out[3] = t == (F)0 ? (F)1 : (F)0;
        return;
    }
    const F d = p - t;
    switch (op) {
        case REG_MSE: out[0] = d * d; break;
        case REG_MAE: out[0] = fabs(d); break;
        case REG_MAPE: out[0] = fabs(d) / fmax(fabs(t), eps); break;
        case REG_SMAPE: out[0] = fabs(d) / fmax(fabs(t) + fabs(p), eps); break;  // the factor 2 is applied to the sum
        case REG_WMAPE: out[0] = fabs(d), out[1] = fabs(t); break;
        case REG_MSLE: {
            const F l = log1p(p) - log1p(t);
            out[0] = l * l;
            break;
        }
        case REG_LOGCOSH: out[0] = log((exp(d) + exp(-d)) / (F)2); break;
        case REG_MINKOWSKI: out[0] = pow(fabs(d), param); break;
        case REG_R2: {
            const F r = t - p;
            out[0] = t * t, out[1] = t, out[2] = r * r;
            break;
        }
        case REG_EXPVAR: {
            const F r = t - p;
            out[0] = r, out[1] = r * r, out[2] = t, out[3] = t * t;
            break;
        }
    }
}

}  // namespace mb200
