// Microbenchmark: cost of finding the lanes that hold the same 8-bit digit (the inner step of radix ranking) with
// MATCH.ANY vs 8 ballots, on sm_100a.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o match_bench match_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kItems = 16;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned peers_match(unsigned digit) { return __match_any_sync(kFull, digit); }
__device__ __forceinline__ unsigned peers_ballot(unsigned digit) {
    unsigned peers = kFull;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (digit >> b) & 1u;
        const unsigned m = __ballot_sync(kFull, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

template <int kMode>
__global__ void __launch_bounds__(512, 2) bench(unsigned* out, int iters, unsigned seed) {
    unsigned key[kItems];
    unsigned x = seed ^ (blockIdx.x * 512u + threadIdx.x) * 2654435761u;
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        x = x * 1664525u + 1013904223u;
        key[i] = x;
    }
    unsigned acc = 0;
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const unsigned digit = (key[i] >> ((it & 3) * 8)) & 255u;
            const unsigned peers = kMode == 0 ? peers_match(digit) : peers_ballot(digit);
            acc += __popc(peers & lt) + (__ffs(peers) - 1);
            key[i] += acc;  // keep the chain honest
        }
    }
    out[blockIdx.x * 512 + threadIdx.x] = acc;
}

int main() {
    unsigned* out;
    cudaMalloc(&out, 296 * 512 * sizeof(unsigned));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int iters = 200;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) bench<0><<<296, 512>>>(out, iters, 1234u);
            else bench<1><<<296, 512>>>(out, iters, 1234u);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            const double ops = 296.0 * 512 * kItems * iters;  // per-thread digit lookups
            printf("%s rep %d: %.3f ms  %.1f G lookups/s  (%.2f cycles per warp-level op per SM-subpartition at 1.9 GHz)\n",
                   mode == 0 ? "MATCH.ANY" : "8xBALLOT ", rep, ms, ops / ms / 1e6,
                   ms * 1e-3 * 1.9e9 / (ops / 32 / (148.0 * 4)));
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
