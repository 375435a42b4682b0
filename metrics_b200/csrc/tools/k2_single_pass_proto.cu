// ROUND-2 PROTOTYPE (not part of the library): binary stat counts in ONE pass over the scores.
//
// K2 today reads the scores twice (805 MB of algorithmic traffic become ~1.07 GB at 2^26 elements, 0.58 of the copy peak):
// the reference's `normalize_logits_if_needed` is a batch-global vote — "is ANY score outside [0, 1]?" — that has to be
// known before the first threshold decision.  This prototype evaluates BOTH outcomes of the vote per element (prediction
// from the raw score, prediction from its sigmoid), keeps eight counters instead of four, and lets a one-thread epilogue pick
// the four that the vote selects.  Scores and labels are read once: 12 B / element.
//
// Self-checking: compares with a host evaluation and with the library's two-pass `mb200_binary_stat_counts`, for a
// probabilities batch and a logits batch, then times both.     build: see `make tools` recipe for abi_smoke (same flags)
#include <cmath>
#include <cstdint>
#include <cstdio>
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

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// acc: [0..3] tp fp tn fn if the batch is probabilities, [4..7] if it is logits, [8] vote
__global__ void __launch_bounds__(256) dual_hypothesis_counts(const float* __restrict__ scores, const long long* __restrict__ target,
                                                              long long n, float threshold, unsigned long long* __restrict__ acc) {
    unsigned c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned outside = 0;
    auto consume = [&](float x, long long t) {
        outside |= (x < 0.0f) | (x > 1.0f);
        if (t != 0 && t != 1) return;  // (the library also flags these; the prototype only skips them)
        const int raw = x > threshold;
        const int sig = (1.0f / (1.0f + expf(-x))) > threshold;
        // index: 0 tp, 1 fp, 2 tn, 3 fn  ==  (pred ? (t ? 0 : 1) : (t ? 3 : 2))
        c[raw ? (t ? 0 : 1) : (t ? 3 : 2)]++;
        c[4 + (sig ? (t ? 0 : 1) : (t ? 3 : 2))]++;
    };
    const long long nvec = n / 4;  // 4 scores (16 B) + 4 labels (2 x 16 B) per step; both arrays are cudaMalloc-aligned
    const uint4* s4 = reinterpret_cast<const uint4*>(scores);
    const uint4* t4 = reinterpret_cast<const uint4*>(target);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 sv = ld_stream(s4 + v), ta = ld_stream(t4 + 2 * v), tb = ld_stream(t4 + 2 * v + 1);
        const float* x = reinterpret_cast<const float*>(&sv);
        const long long* la = reinterpret_cast<const long long*>(&ta);
        const long long* lb = reinterpret_cast<const long long*>(&tb);
        consume(x[0], la[0]), consume(x[1], la[1]), consume(x[2], lb[0]), consume(x[3], lb[1]);
    }
    for (long long i = nvec * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) consume(scores[i], target[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned w = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(acc + k, (unsigned long long)w);
    }
    if (__any_sync(0xffffffffu, outside) && (threadIdx.x & 31) == 0) atomicOr(acc + 8, 1ull);
}

__global__ void select_hypothesis(unsigned long long* acc, long long* counts) {
    const int base = acc[8] ? 4 : 0;
    for (int k = 0; k < 4; ++k) counts[k] += (long long)acc[base + k];
    for (int k = 0; k < 9; ++k) acc[k] = 0;  // self-cleaning, like the library's workspaces
}

static void host_counts(const std::vector<float>& x, const std::vector<long long>& t, float thr, long long out[4]) {
    bool logits = false;
    for (float v : x) logits |= (v < 0.0f) | (v > 1.0f);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (size_t i = 0; i < x.size(); ++i) {
        const float v = logits ? 1.0f / (1.0f + std::exp(-x[i])) : x[i];
        const int p = v > thr;
        out[p ? (t[i] ? 0 : 1) : (t[i] ? 3 : 2)]++;
    }
}

int main() {
    const long long n = 1ll << 26;
    const float thr = 0.5f;
    std::vector<float> hx(n);
    std::vector<long long> ht(n);
    float* dx;
    long long *dt, *dcounts, *dcounts_lib;
    unsigned long long* dacc;
    unsigned *dflag, *derr;
    CK(cudaMalloc(&dx, n * sizeof(float)));
    CK(cudaMalloc(&dt, n * sizeof(long long)));
    CK(cudaMalloc(&dcounts, 4 * sizeof(long long)));
    CK(cudaMalloc(&dcounts_lib, 4 * sizeof(long long)));
    CK(cudaMalloc(&dacc, 9 * sizeof(unsigned long long)));
    CK(cudaMalloc(&dflag, sizeof(unsigned)));
    CK(cudaMalloc(&derr, sizeof(unsigned)));
    CK(cudaMemset(dacc, 0, 9 * sizeof(unsigned long long)));
    const int grid = 148 * 8;
    int fails = 0;
    for (int logits = 0; logits < 2; ++logits) {
        unsigned s = 99u + logits;
        for (long long i = 0; i < n; ++i) {
            s = s * 1664525u + 1013904223u;
            const float u = (float)(s >> 8) / (float)(1u << 24);
            hx[i] = logits ? (u - 0.5f) * 8.0f : u;
            ht[i] = (s >> 3) & 1;
        }
        CK(cudaMemcpy(dx, hx.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dt, ht.data(), n * sizeof(long long), cudaMemcpyHostToDevice));
        long long want[4], got[4], lib[4];
        host_counts(hx, ht, thr, want);
        CK(cudaMemset(dcounts, 0, 4 * sizeof(long long)));
        CK(cudaMemset(dcounts_lib, 0, 4 * sizeof(long long)));
        CK(cudaMemset(derr, 0, sizeof(unsigned)));
        dual_hypothesis_counts<<<grid, 256>>>(dx, dt, n, thr, dacc);
        select_hypothesis<<<1, 1>>>(dacc, dcounts);
        const int rc = mb200_binary_stat_counts(dx, MB200_F32, dt, MB200_I64, n, 1, 1, (double)thr, 0, 0, 0, reinterpret_cast<int64_t*>(dcounts_lib), dflag, derr, nullptr);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got, dcounts, sizeof(got), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(lib, dcounts_lib, sizeof(lib), cudaMemcpyDeviceToHost));
        bool ok = rc == 0;
        for (int k = 0; k < 4; ++k) ok = ok && got[k] == want[k] && lib[k] == want[k];
        std::printf("%s %s: single-pass [%lld %lld %lld %lld] library [%lld %lld %lld %lld] host [%lld %lld %lld %lld]\n", ok ? "PASS" : "FAIL",
                    logits ? "logits" : "probabilities", got[0], got[1], got[2], got[3], lib[0], lib[1], lib[2], lib[3], want[0], want[1],
                    want[2], want[3]);
        fails += !ok;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        float ms_new = 0, ms_lib = 0;
        for (int rep = 0; rep < 23; ++rep) {  // 3 warm-up + 20 timed, per variant; 805 MB per pass >> L2
            if (rep == 3) cudaEventRecord(e0);
            dual_hypothesis_counts<<<grid, 256>>>(dx, dt, n, thr, dacc);
            select_hypothesis<<<1, 1>>>(dacc, dcounts);
        }
        cudaEventRecord(e1), cudaEventSynchronize(e1), cudaEventElapsedTime(&ms_new, e0, e1);
        for (int rep = 0; rep < 23; ++rep) {
            if (rep == 3) cudaEventRecord(e0);
            mb200_binary_stat_counts(dx, MB200_F32, dt, MB200_I64, n, 1, 1, (double)thr, 0, 0, 0, reinterpret_cast<int64_t*>(dcounts_lib), dflag, derr, nullptr);
        }
        cudaEventRecord(e1), cudaEventSynchronize(e1), cudaEventElapsedTime(&ms_lib, e0, e1);
        std::printf("     time per call: single-pass %.1f us (%.0f GB/s of 12 B/elem)   library two-pass %.1f us\n", ms_new / 20 * 1e3,
                    12.0 * n / (ms_new / 20 * 1e-3) / 1e9, ms_lib / 20 * 1e3);
    }
    std::printf("%s\n", fails ? "K2_PROTO_FAIL" : "K2_PROTO_OK");
    return fails ? 1 : 0;
}
