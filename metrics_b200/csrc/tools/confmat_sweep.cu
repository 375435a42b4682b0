// Standalone sweep of launch/load policies for the K1 row kernel (confusion-matrix sink) on the cfg2 shape.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o confmat_sweep confmat_sweep.cu ../core.cu
// It is a measurement tool, not part of the shipped library.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../argmax_core.cuh"
#include "../common.cuh"

using namespace mb200;

template <int kLoad>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 r;
    if (kLoad == 0)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (kLoad == 1)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (kLoad == 2) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    }
    else
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int kLoad>
struct Loader {
    const uint4* base;
    __device__ __forceinline__ uint4 operator()(int vi) const { return ld16<kLoad>(base + vi); }
};

// ---- pure streaming probe: what can a read-only pass over the logits reach at all? ----
template <int kLoad, int kUnroll>
__global__ void __launch_bounds__(256) probe_kernel(const uint4* __restrict__ p, size_t nvec, unsigned* out) {
    unsigned acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = ld16<kLoad>(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < nvec; i += stride) {
        uint4 v = ld16<kLoad>(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// ---- row kernel variants ----
template <int kLoad, int kThreads, int kMinBlocks, bool kPipe>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
row_kernel(const __nv_bfloat16* __restrict__ preds, const long long* __restrict__ target, int n, int C, long long* confmat) {
    using T = __nv_bfloat16;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int nwarps = gridDim.x * wpb;
    const int nvec = (C * 2) >> 4;
    int r = blockIdx.x * wpb + (threadIdx.x >> 5);
    if (r >= n) return;
    if (!kPipe) {
        long long t_next = __ldg(target + r);
        for (; r < n; r += nwarps) {
            const long long t = t_next;
            if (r + nwarps < n) t_next = __ldg(target + r + nwarps);
            Chunk c;
            load_chunk(c, Loader<kLoad>{reinterpret_cast<const uint4*>(preds + (size_t)r * C)}, 0, nvec, lane);
            unsigned ckey, ccol;
            chunk_argmax<T>(c, 0, lane, ckey, ccol);
            if (lane == 0) red_add_u64(confmat + t * C + ccol, 1ull);
        }
    } else {
        long long t = __ldg(target + r);
        long long t_ahead = (r + nwarps < n) ? __ldg(target + r + nwarps) : 0;
        Chunk buf;
        load_chunk(buf, Loader<kLoad>{reinterpret_cast<const uint4*>(preds + (size_t)r * C)}, 0, nvec, lane);
        while (true) {
            const int nr = r + nwarps;
            const long long nt = t_ahead;
            const bool has_next = nr < n;
            Chunk nbuf;
            if (has_next) {
                if (nr + nwarps < n) t_ahead = __ldg(target + nr + nwarps);
                load_chunk(nbuf, Loader<kLoad>{reinterpret_cast<const uint4*>(preds + (size_t)nr * C)}, 0, nvec, lane);
            }
            unsigned ckey, ccol;
            chunk_argmax<T>(buf, 0, lane, ckey, ccol);
            if (lane == 0) red_add_u64(confmat + t * C + ccol, 1ull);
            if (!has_next) break;
            buf = nbuf;
            r = nr;
            t = nt;
        }
    }
}

static const int N = 65536, C = 1000, NBUF = 4;
static __nv_bfloat16* d_logits[NBUF];
static long long* d_target[NBUF];
static long long* d_confmat;

template <typename F>
static float time_it(F&& launch, int iters = 400) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 20; ++i) launch(i % NBUF);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) launch(i % NBUF);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("   CUDA error: %s\n", cudaGetErrorString(e));
    return ms / iters * 1000.f;  // us
}

template <int kLoad, int kThreads, int kMinBlocks, bool kPipe>
static void run_variant(const char* name) {
    auto kern = row_kernel<kLoad, kThreads, kMinBlocks, kPipe>;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int mult = per_sm; mult >= 1 && mult >= per_sm - 2; --mult) {
        const int grid = sms * mult;
        float us = time_it([&](int b) { kern<<<grid, kThreads>>>(d_logits[b], d_target[b], N, C, d_confmat); });
        const double gbs = 132120576.0 / (us * 1e-6) / 1e9;
        printf("%-34s regs=%3d blk/SM=%d(of %d) grid=%5d  %7.2f us  %7.1f GB/s\n", name, fa.numRegs, mult, per_sm, grid, us, gbs);
    }
}

int main() {
    const size_t elems = (size_t)N * C;
    std::vector<unsigned short> h(elems);
    std::vector<long long> ht(N);
    for (int b = 0; b < NBUF; ++b) {
        unsigned s = 12345u + b;
        for (size_t i = 0; i < elems; ++i) {
            s = s * 1664525u + 1013904223u;
            // random bf16 in roughly [-4, 4): sign + exponent 0x3f..0x40 range
            h[i] = (unsigned short)(((s >> 16) & 0x8000u) | (0x3e00u + ((s >> 8) & 0x3ffu)));
        }
        for (int i = 0; i < N; ++i) {
            s = s * 1664525u + 1013904223u;
            ht[i] = (s >> 8) % C;
        }
        cudaMalloc(&d_logits[b], elems * 2);
        cudaMalloc(&d_target[b], N * 8);
        cudaMemcpy(d_logits[b], h.data(), elems * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(d_target[b], ht.data(), N * 8, cudaMemcpyHostToDevice);
    }
    cudaMalloc(&d_confmat, (size_t)C * C * 8);
    cudaMemset(d_confmat, 0, (size_t)C * C * 8);
    unsigned* d_out;
    cudaMalloc(&d_out, 4);

    printf("== streaming probes (read-only pass over the 131 MB logits; GB/s on 131072000 B) ==\n");
    const size_t nvec = elems * 2 / 16;
    for (int mult : {4, 8}) {
        float us;
        us = time_it([&](int b) { probe_kernel<0, 4><<<148 * mult, 256>>>((const uint4*)d_logits[b], nvec, d_out); });
        printf("probe nc.noalloc   unroll4 grid=148x%d  %7.2f us %7.1f GB/s\n", mult, us, 131072000.0 / us / 1e3);
        us = time_it([&](int b) { probe_kernel<0, 8><<<148 * mult, 256>>>((const uint4*)d_logits[b], nvec, d_out); });
        printf("probe nc.noalloc   unroll8 grid=148x%d  %7.2f us %7.1f GB/s\n", mult, us, 131072000.0 / us / 1e3);
        us = time_it([&](int b) { probe_kernel<1, 8><<<148 * mult, 256>>>((const uint4*)d_logits[b], nvec, d_out); });
        printf("probe +L2::256B    unroll8 grid=148x%d  %7.2f us %7.1f GB/s\n", mult, us, 131072000.0 / us / 1e3);
        us = time_it([&](int b) { probe_kernel<2, 8><<<148 * mult, 256>>>((const uint4*)d_logits[b], nvec, d_out); });
        printf("probe +evict_first unroll8 grid=148x%d  %7.2f us %7.1f GB/s\n", mult, us, 131072000.0 / us / 1e3);
        us = time_it([&](int b) { probe_kernel<3, 8><<<148 * mult, 256>>>((const uint4*)d_logits[b], nvec, d_out); });
        printf("probe plain ld     unroll8 grid=148x%d  %7.2f us %7.1f GB/s\n", mult, us, 131072000.0 / us / 1e3);
    }
    {
        // device-to-device copy of the same buffer for reference (read + write bytes)
        void* tmp;
        cudaMalloc(&tmp, elems * 2);
        float us = time_it([&](int b) { cudaMemcpyAsync(tmp, d_logits[b], elems * 2, cudaMemcpyDeviceToDevice); }, 100);
        printf("cudaMemcpy D2D 131 MB: %7.2f us  %7.1f GB/s (read+write)\n", us, 2 * 131072000.0 / us / 1e3);
        cudaFree(tmp);
    }
    printf("== row kernel variants (GB/s on 132120576 algorithmic B) ==\n");
    run_variant<0, 256, 1, false>("nopipe ld0 t256");
    run_variant<0, 256, 8, false>("nopipe ld0 t256 minblk8");
    run_variant<0, 128, 1, false>("nopipe ld0 t128");
    run_variant<0, 512, 1, false>("nopipe ld0 t512");
    run_variant<1, 256, 1, false>("nopipe L2::256B t256");
    run_variant<2, 256, 1, false>("nopipe evict_first t256");
    run_variant<3, 256, 1, false>("nopipe plain ld t256");
    run_variant<0, 256, 1, true>("pipe ld0 t256");
    run_variant<0, 128, 1, true>("pipe ld0 t128");
    run_variant<1, 256, 1, true>("pipe L2::256B t256");
    run_variant<0, 256, 5, true>("pipe ld0 t256 minblk5");
    return 0;
}
