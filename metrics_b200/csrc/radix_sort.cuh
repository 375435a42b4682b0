// Segmented LSD radix sort building blocks (8-bit digits), shared by the curve and detection kernels.
// KeyT in {u32, u64}, ValT any trivially copyable payload.  grid = (tiles_per_segment, segments).
#pragma once
#include "common.cuh"

namespace mb200 {

// =====================================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA

// (A) per-tile digit histogram -> tile_hist[seg][digit][tile]
template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const KeyT* __restrict__ keys, int n, int tiles,
                                                                  int shift, unsigned* __restrict__ tile_hist) {
    __shared__ unsigned hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int seg = blockIdx.y, tile = blockIdx.x;
    const KeyT* __restrict__ k = keys + (size_t)seg * n;
    const int base = tile * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = base + i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(unsigned)(k[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    tile_hist[((size_t)seg * 256 + threadIdx.x) * tiles + tile] = hist[threadIdx.x];
}

// (B) per (segment, digit): exclusive scan over tiles in place; digit totals -> digit_total[seg][digit]
static __global__ void __launch_bounds__(256) radix_scan_kernel(unsigned* __restrict__ tile_hist, int tiles,
                                                         unsigned* __restrict__ digit_total) {
    __shared__ unsigned warp_sums[8];
    __shared__ unsigned carry_s;
    const int seg = blockIdx.y, digit = blockIdx.x;
    unsigned* __restrict__ h = tile_hist + ((size_t)seg * 256 + digit) * tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 256) {
        const int idx = base + threadIdx.x;
        const unsigned v = idx < tiles ? h[idx] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_sums[w];
        const unsigned carry = carry_s;
        if (idx < tiles) h[idx] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) digit_total[seg * 256 + digit] = carry_s;
}

// (C) stable scatter.  Warp-striped arrangement: warp w owns keys [tile_base + w*512, +512), item i of lane l is
// element i*32 + l of that range, so (i, l) order == memory order.  Ranks come from MATCH.ANY groups + per-warp
// running digit counters in shared memory.
template <typename KeyT, typename ValT>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const KeyT* __restrict__ keys_in,
                                                                     const ValT* __restrict__ labels_in,
                                                                     KeyT* __restrict__ keys_out,
                                                                     ValT* __restrict__ labels_out, int n,
                                                                     int tiles, int shift,
                                                                     const unsigned* __restrict__ tile_hist,
                                                                     const unsigned* __restrict__ digit_total) {
    __shared__ unsigned warp_hist[8][256];
    __shared__ unsigned digit_base[256];
    __shared__ unsigned scan_tmp[8];
    const int seg = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t seg_off = (size_t)seg * n;
    const KeyT* __restrict__ kin = keys_in + seg_off;
    const ValT* __restrict__ lin = labels_in + seg_off;
    for (int i = threadIdx.x; i < 8 * 256; i += kSortThreads) (&warp_hist[0][0])[i] = 0;

    // global base of every digit for this tile: exclusive scan of the digit totals + this tile's exclusive offset
    {
        const unsigned tot = digit_total[seg * 256 + threadIdx.x];
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) scan_tmp[warp] = incl;
        __syncthreads();
        unsigned woff = 0;
        for (int w = 0; w < warp; ++w) woff += scan_tmp[w];
        digit_base[threadIdx.x] = woff + incl - tot + tile_hist[((size_t)seg * 256 + threadIdx.x) * tiles + tile];
    }
    __syncthreads();

    const int wbase = tile * kSortTile + warp * (kSortItems * 32);
    KeyT key[kSortItems];
    ValT lab[kSortItems];
    unsigned short rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? kin[idx] : (KeyT)0;
        lab[i] = valid ? lin[idx] : (ValT)0;
    }
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        const bool valid = idx < n;
        const unsigned digit = valid ? ((unsigned)(key[i] >> shift) & 255u) : (0x100u + lane);  // invalid lanes: unique groups
        const unsigned peers = __match_any_sync(kFull, digit);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (valid && lane == leader) {
            base = warp_hist[warp][digit];
            warp_hist[warp][digit] = base + __popc(peers);
        }
        base = __shfl_sync(kFull, base, leader);
        rank[i] = (unsigned short)(base + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over the 8 warps for every digit
    {
        unsigned off = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned c = warp_hist[w][threadIdx.x];
            warp_hist[w][threadIdx.x] = off;
            off += c;
        }
    }
    __syncthreads();
    KeyT* __restrict__ kout = keys_out + seg_off;
    ValT* __restrict__ lout = labels_out + seg_off;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        if (idx < n) {
            const unsigned digit = (unsigned)(key[i] >> shift) & 255u;
            const unsigned dst = digit_base[digit] + warp_hist[warp][digit] + rank[i];
            kout[dst] = key[i];
            lout[dst] = lab[i];
        }
    }
}


// Host helper: `key_bytes` passes over ping-pong buffers; returns 0/1 = which buffer pair holds the result.
template <typename KeyT, typename ValT>
static inline int radix_sort_passes(KeyT* keys_a, ValT* vals_a, KeyT* keys_b, ValT* vals_b, int n, int segments,
                                    int key_bytes, unsigned* tile_hist, unsigned* digit_total, cudaStream_t st,
                                    void (*on_launch)()) {
    const int tiles = (n + kSortTile - 1) / kSortTile;
    const dim3 tgrid((unsigned)tiles, (unsigned)segments);
    KeyT *kin = keys_a, *kout = keys_b;
    ValT *vin = vals_a, *vout = vals_b;
    for (int pass = 0; pass < key_bytes; ++pass) {
        const int shift = 8 * pass;
        radix_hist_kernel<KeyT><<<tgrid, kSortThreads, 0, st>>>(kin, n, tiles, shift, tile_hist);
        radix_scan_kernel<<<dim3(256, (unsigned)segments), 256, 0, st>>>(tile_hist, tiles, digit_total);
        radix_scatter_kernel<KeyT, ValT><<<tgrid, kSortThreads, 0, st>>>(kin, vin, kout, vout, n, tiles, shift,
                                                                          tile_hist, digit_total);
        if (on_launch) on_launch(), on_launch(), on_launch();
        KeyT* tk = kin;
        kin = kout;
        kout = tk;
        ValT* tv = vin;
        vin = vout;
        vout = tv;
    }
    return (key_bytes & 1);
}

}  // namespace mb200
