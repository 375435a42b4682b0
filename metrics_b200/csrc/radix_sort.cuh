// Segmented LSD radix sort (8-bit digits) for sm_100a — "one sweep" per digit.
//
//   digit histograms   ONE pass over the keys counts every digit of every radix pass (order-independent), a tiny scan
//                      turns them into per-(segment, pass) exclusive digit offsets;
//   per radix pass     ONE kernel: CTAs take tiles in ticket order, rank their 8192 keys (warp-striped layout,
//                      ballot-built digit groups + per-warp digit counters in shared memory), publish the tile's digit counts,
//                      obtain the counts of all earlier tiles of the segment by decoupled look-back on a status word
//                      per (tile, digit) [flag:2 | count:30], reorder keys+payload by digit in shared memory and write
//                      digit runs to their final positions with coalesced stores.
// A pass therefore reads and writes every record exactly once (5 B/record for the curve keys); no per-pass histogram or
// scan kernels.  The look-back spin is bounded: if it ever expires the kernel raises MB200_FLAG_SPIN_TIMEOUT and exits
// instead of hanging the GPU.
// KeyT in {u32, u64}, ValT any trivially copyable payload.  Segments have equal length n (< 2^30).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace mb200 {

constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 8192 keys per CTA
constexpr unsigned kStatAgg = 1u << 30;               // tile aggregate published
constexpr unsigned kStatPrefix = 2u << 30;            // inclusive prefix published
constexpr unsigned kStatMask = (1u << 30) - 1u;

// ---- digit histograms of all passes in one read ----------------------------------------------------------------------
// hist layout: [segment][pass][256]
template <typename KeyT>
__global__ void __launch_bounds__(256) radix_digit_hist_kernel(const KeyT* __restrict__ keys, int n, int key_bytes,
                                                               unsigned* __restrict__ hist) {
    extern __shared__ unsigned sh_hist[];  // key_bytes * 256
    for (int i = threadIdx.x; i < key_bytes * 256; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const int seg = blockIdx.y;
    const KeyT* __restrict__ k = keys + (size_t)seg * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const KeyT key = k[i];
        for (int p = 0; p < key_bytes; ++p) atomicAdd(&sh_hist[p * 256 + ((unsigned)(key >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < key_bytes * 256; i += blockDim.x) {
        const unsigned v = sh_hist[i];
        if (v) atomicAdd(&hist[(size_t)seg * key_bytes * 256 + i], v);
    }
}

// in-place exclusive scan of each 256-bin histogram: grid = segments * passes, block = 256
static __global__ void __launch_bounds__(256) radix_digit_scan_kernel(unsigned* __restrict__ hist) {
    __shared__ unsigned ws[8];
    unsigned* h = hist + (size_t)blockIdx.x * 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = h[threadIdx.x];
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < warp; ++w) woff += ws[w];
    h[threadIdx.x] = woff + incl - v;
}

// ---- one radix pass ------------------------------------------------------------------------------------------------------
// kMode: 0 = (key, value) pairs; 1 = keys only; 2 = keys only in, and the pass SPLITS every key on the way out into
// (key >> 1, key & 1) — the last pass of a sort whose keys carry a one-bit payload in bit 0 (curve.cu: the label).
template <typename KeyT, typename ValT, int kMode = 0>
struct SortSmem {
    KeyT keys[kSortTile];
    ValT vals[kMode == 0 ? kSortTile : 1];
    unsigned warp_hist[kSortWarps][256];
    unsigned digit_base[256];  // global position of the tile's first key of each digit
    unsigned tile_off[256];    // position of the digit's run inside the tile
    unsigned scan_tmp[8];
    unsigned tile_index;
};

template <typename KeyT, typename ValT, int kMode = 0>
__global__ void __launch_bounds__(kSortThreads, sizeof(KeyT) == 4 ? 2 : 1) radix_onesweep_kernel(
    const KeyT* __restrict__ keys_in, const ValT* __restrict__ vals_in, KeyT* __restrict__ keys_out,
    ValT* __restrict__ vals_out, int n, int tiles_per_seg, int shift, int pass, int key_bytes,
    const unsigned* __restrict__ digit_offsets /*[seg][pass][256] exclusive*/, unsigned* __restrict__ status /*[tiles][256]*/,
    unsigned* __restrict__ ticket, unsigned* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    SortSmem<KeyT, ValT, kMode>& sm = *reinterpret_cast<SortSmem<KeyT, ValT, kMode>*>(sort_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sm.tile_index = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&sm.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const unsigned tile_global = sm.tile_index;
    const int seg = (int)(tile_global / (unsigned)tiles_per_seg);
    const int tile = (int)(tile_global % (unsigned)tiles_per_seg);
    const size_t seg_off = (size_t)seg * n;
    const KeyT* __restrict__ kin = keys_in + seg_off;
    const ValT* __restrict__ vin = vals_in + seg_off;
    const int tile_base = tile * kSortTile;
    const int tile_count = min(kSortTile, n - tile_base);

    // ---- load (warp-striped: item i of lane l is element i*32 + l of the warp's 512-key span) and rank ----
    const int wbase = tile_base + warp * (kSortItems * 32);
    KeyT key[kSortItems];
    ValT val[kSortItems];
    unsigned short rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? kin[idx] : (KeyT)0;
        if (kMode == 0) val[i] = valid ? vin[idx] : (ValT)0;
    }
    const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        const bool valid = idx < n;
        const unsigned digit = (unsigned)(key[i] >> shift) & 255u;
        // lanes holding the same digit: 8 ballots (one per digit bit).  MATCH.ANY does this in one instruction but
        // measures ~235 cycles per warp instruction per SM sub-partition on B200 vs ~100 for this sequence
        // (tools/match_bench.cu) — it was half of the pass time.
        unsigned peers = __ballot_sync(kFull, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (digit >> b) & 1u;
            const unsigned m = __ballot_sync(kFull, bit);
            peers &= bit ? m : ~m;
        }
        if (!valid) peers = 1u << lane;  // tail lanes: singleton groups that never touch the counters
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (valid && lane == leader) {
            base = sm.warp_hist[warp][digit];
            sm.warp_hist[warp][digit] = base + __popc(peers);
        }
        base = __shfl_sync(kFull, base, leader);
        rank[i] = (unsigned short)(base + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive scan over warps, publish, look back ----
    unsigned my_count = 0;
    if (threadIdx.x < 256) {
        const int d = threadIdx.x;
        unsigned off = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned c = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = off;
            off += c;
        }
        my_count = off;
        volatile unsigned* st = status + (size_t)tile_global * 256 + d;
        unsigned excl = 0;
        if (tile == 0) {
            *st = kStatPrefix | my_count;
        } else {
            *st = kStatAgg | my_count;
            // Decoupled look-back over the earlier tiles of the segment.  A window of kLookWindow predecessors is
            // fetched with independent volatile loads (one L2 round trip for the whole window) and consumed in order:
            // walking back one tile per round trip chains hundreds of L2 latencies across a wave of co-resident tiles.
            constexpr int kLookWindow = 8;
            const long long seg_first = (long long)tile_global - tile;  // the segment's tile 0 always holds a prefix
            long long j = (long long)tile_global - 1;
            unsigned spins = 0;
            bool done = false;
            while (!done) {
                unsigned w[kLookWindow];
#pragma unroll
                for (int k = 0; k < kLookWindow; ++k) {
                    const long long jj = j - k;
                    w[k] = jj >= seg_first ? *(volatile unsigned*)(status + (size_t)jj * 256 + d) : (2u << 30);  // == kStatPrefix | 0
                }
                int used = 0;
#pragma unroll
                for (int k = 0; k < kLookWindow; ++k) {
                    if (done || used != k) continue;  // stop at the first entry that is not ready yet
                    const unsigned flag = w[k] & ~kStatMask;
                    if (flag == 0u) continue;
                    excl += w[k] & kStatMask;
                    used = k + 1;
                    if (flag == kStatPrefix) done = true;
                }
                j -= used;
                if (!done && used < kLookWindow) {  // ran into a tile that has not published anything yet
                    if (++spins > (1u << 22)) {     // never hang the GPU: report and produce garbage instead
                        if (err) atomicOr(err, MB200_FLAG_SPIN_TIMEOUT);
                        break;
                    }
                    __nanosleep(20);
                }
            }
            *st = kStatPrefix | ((excl + my_count) & kStatMask);
        }
        sm.digit_base[d] = digit_offsets[((size_t)seg * key_bytes + pass) * 256 + d] + excl;
        // exclusive scan of the tile's digit counts over the digits
        unsigned incl = my_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.scan_tmp[warp] = incl;
        sm.tile_off[d] = incl - my_count;  // warp-local for now
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        unsigned woff = 0;
        for (int w = 0; w < warp; ++w) woff += sm.scan_tmp[w];
        const unsigned off = sm.tile_off[threadIdx.x] + woff;
        sm.tile_off[threadIdx.x] = off;
        sm.digit_base[threadIdx.x] -= off;  // -> (global position of the digit's run) - (its position in the tile): dst = base + i
    }
    __syncthreads();

    // ---- reorder by digit in shared memory ----
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int idx = wbase + i * 32 + lane;
        if (idx < n) {
            const unsigned digit = (unsigned)(key[i] >> shift) & 255u;
            const unsigned pos = sm.tile_off[digit] + sm.warp_hist[warp][digit] + rank[i];
            sm.keys[pos] = key[i];
            if (kMode == 0) sm.vals[pos] = val[i];
        }
    }
    __syncthreads();
    // ---- coalesced stores of the digit runs ----
    KeyT* __restrict__ kout = keys_out + seg_off;
    ValT* __restrict__ vout = vals_out + seg_off;
    for (int i = threadIdx.x; i < tile_count; i += kSortThreads) {
        const KeyT k = sm.keys[i];
        const unsigned digit = (unsigned)(k >> shift) & 255u;
        const unsigned dst = sm.digit_base[digit] + (unsigned)i;
        if (kMode == 2) {
            kout[dst] = k >> 1;
            vout[dst] = (ValT)(k & (KeyT)1);
        } else {
            kout[dst] = k;
            if (kMode == 0) vout[dst] = sm.vals[i];
        }
    }
}

// ---- host driver -----------------------------------------------------------------------------------------------------------
// scratch (unsigned words): hist [segments*key_bytes*256] | status [key_bytes][segments*tiles][256] | tickets [key_bytes]
static inline size_t radix_sort_scratch_words(long long n, long long segments, int key_bytes) {
    const long long tiles = (n + kSortTile - 1) / kSortTile;
    return (size_t)(segments * key_bytes * 256 + (long long)key_bytes * segments * tiles * 256 + key_bytes + 64);
}

// Sorts every segment by the low `key_bytes` bytes of the key.  Returns 0/1 = which buffer pair holds the result
// (a = 0, b = 1), or a negative MB200_ERR_* code.
template <typename KeyT, typename ValT>
static inline int radix_sort_passes(KeyT* keys_a, ValT* vals_a, KeyT* keys_b, ValT* vals_b, int n, int segments,
                                    int key_bytes, unsigned* scratch, unsigned* err_flag, cudaStream_t st,
                                    void (*on_launch)()) {
    const int tiles = (n + kSortTile - 1) / kSortTile;
    const size_t words = radix_sort_scratch_words(n, segments, key_bytes);
    if (cudaMemsetAsync(scratch, 0, words * sizeof(unsigned), st) != cudaSuccess) return MB200_ERR_CUDA;
    unsigned* hist = scratch;
    unsigned* status = hist + (size_t)segments * key_bytes * 256;
    unsigned* tickets = status + (size_t)key_bytes * segments * tiles * 256;

    int hgrid = (n + 256 * 16 - 1) / (256 * 16);
    const int hcap = std::max(1, (sm_count() * 8) / std::max(1, segments));
    if (hgrid > hcap) hgrid = hcap;
    if (hgrid < 1) hgrid = 1;
    radix_digit_hist_kernel<KeyT><<<dim3((unsigned)hgrid, (unsigned)segments), 256, (size_t)key_bytes * 256 * 4, st>>>(
        keys_a, n, key_bytes, hist);
    radix_digit_scan_kernel<<<(unsigned)(segments * key_bytes), 256, 0, st>>>(hist);
    if (on_launch) on_launch(), on_launch();

    auto kern = radix_onesweep_kernel<KeyT, ValT>;
    const size_t smem = sizeof(SortSmem<KeyT, ValT>);
    if (ensure_dynamic_smem(kern, (int)smem) != cudaSuccess) return MB200_ERR_CUDA;
    KeyT *kin = keys_a, *kout = keys_b;
    ValT *vin = vals_a, *vout = vals_b;
    for (int pass = 0; pass < key_bytes; ++pass) {
        kern<<<(unsigned)(segments * tiles), kSortThreads, smem, st>>>(
            kin, vin, kout, vout, n, tiles, 8 * pass, pass, key_bytes, hist,
            status + (size_t)pass * segments * tiles * 256, tickets + pass, err_flag);
        if (on_launch) on_launch();
        KeyT* tk = kin;
        kin = kout;
        kout = tk;
        ValT* tv = vin;
        vin = vout;
        vout = tv;
    }
    return (key_bytes & 1);
}

// Keys with a one-bit payload in bit 0 (the curve label): every pass moves 4-byte keys only, the LAST pass writes the split
// (key >> 1, key & 1) pair.  Pass p sorts by byte p of the composite key; the result lands in (keys_b, vals_b) for an odd
// number of passes, (keys_a, vals_a) for an even one — returned like radix_sort_passes.
// hist_done: the caller zeroed the scratch (radix_sort_zero_scratch) and its key-producing kernel already counted every digit
// into the histograms at the front of it (layout [segment][pass][256]) — the separate histogram read of the keys is skipped.
static inline int radix_sort_zero_scratch(unsigned* scratch, int n, int segments, int key_bytes, cudaStream_t st) {
    const size_t words = radix_sort_scratch_words(n, segments, key_bytes);
    return cudaMemsetAsync(scratch, 0, words * sizeof(unsigned), st) == cudaSuccess ? 0 : MB200_ERR_CUDA;
}
template <typename KeyT, typename ValT>
static inline int radix_sort_passes_bit0(KeyT* keys_a, ValT* vals_a, KeyT* keys_b, ValT* vals_b, int n, int segments,
                                         int key_bytes, unsigned* scratch, unsigned* err_flag, cudaStream_t st,
                                         void (*on_launch)(), bool hist_done = false) {
    const int tiles = (n + kSortTile - 1) / kSortTile;
    unsigned* hist = scratch;
    unsigned* status = hist + (size_t)segments * key_bytes * 256;
    unsigned* tickets = status + (size_t)key_bytes * segments * tiles * 256;
    if (!hist_done) {
        if (radix_sort_zero_scratch(scratch, n, segments, key_bytes, st)) return MB200_ERR_CUDA;
        int hgrid = (n + 256 * 16 - 1) / (256 * 16);
        const int hcap = std::max(1, (sm_count() * 8) / std::max(1, segments));
        if (hgrid > hcap) hgrid = hcap;
        if (hgrid < 1) hgrid = 1;
        radix_digit_hist_kernel<KeyT><<<dim3((unsigned)hgrid, (unsigned)segments), 256, (size_t)key_bytes * 256 * 4, st>>>(
            keys_a, n, key_bytes, hist);
        if (on_launch) on_launch();
    }
    radix_digit_scan_kernel<<<(unsigned)(segments * key_bytes), 256, 0, st>>>(hist);
    if (on_launch) on_launch();
    auto mid = radix_onesweep_kernel<KeyT, ValT, 1>;
    auto last = radix_onesweep_kernel<KeyT, ValT, 2>;
    const size_t smem = sizeof(SortSmem<KeyT, ValT, 1>);
    if (ensure_dynamic_smem(mid, (int)smem) != cudaSuccess || ensure_dynamic_smem(last, (int)smem) != cudaSuccess)
        return MB200_ERR_CUDA;
    KeyT *kin = keys_a, *kout = keys_b;
    ValT* vout = (key_bytes & 1) ? vals_b : vals_a;
    for (int pass = 0; pass < key_bytes; ++pass) {
        auto kern = pass == key_bytes - 1 ? last : mid;
        kern<<<(unsigned)(segments * tiles), kSortThreads, smem, st>>>(
            kin, nullptr, kout, vout, n, tiles, 8 * pass, pass, key_bytes, hist,
            status + (size_t)pass * segments * tiles * 256, tickets + pass, err_flag);
        if (on_launch) on_launch();
        KeyT* tk = kin;
        kin = kout;
        kout = tk;
    }
    return (key_bytes & 1);
}

}  // namespace mb200
