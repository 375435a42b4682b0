// K1 / K1b — fused row-argmax + confusion-matrix / stat-scores accumulation for sm_100a.
//
// Reference op chain replaced (src/torchmetrics/):
//   functional/classification/confusion_matrix.py:297-328  (argmax -> flatten -> ignore drop -> t*C+p -> bincount)
//   functional/classification/stat_scores.py:328-344, 424-448 (same chain, then diag / row / col sums)
//   utilities/data.py:178-206 (_bincount)  and the `state += ...` of the modular classes.
//
// Design: one warp owns one row of the [N, C] score matrix.  Each lane issues up to four
// independent 16-byte streaming loads (the whole 2000-byte bf16 row of the C=1000 config is one "chunk"),
// reduces its registers with NaN-propagating packed max (HMNMX2), the warp agrees on the row maximum with a
// single REDUX on an order-preserving integer key, and a second register-only pass finds the FIRST column
// holding that maximum (torch.argmax tie rule) — again one REDUX.  Lane 0 then commits one 64-bit RED to
// the L2-resident state.  No intermediate (argmax vector, t*C+p, C*C bins) ever touches HBM, so the
// algorithmic traffic is the logits read itself.
#include <stdlib.h>
#include <string.h>

#include "argmax_core.cuh"
#include "common.cuh"
#include "sinks.cuh"

namespace mb200 {

// =====================================================================================================
// Kernels
// =====================================================================================================
struct RowArgs {
    const void* preds;
    const void* target;
    int target_dtype;
    long long n_outer;
    int C;
    long long inner;
    int has_ignore;
    long long ignore_index;
    unsigned* err;
    int pdl_wait = 1;  // overlapped launches: wait for the previous grid's memory before the first input load
};

template <bool kI64>
__device__ __forceinline__ long long fetch_label(const RowArgs& a, long long idx) {
    if (kI64) return __ldg(reinterpret_cast<const long long*>(a.target) + idx);
    return load_label(a.target, a.target_dtype, idx);
}

// Decide whether a row with label t takes part; flags out-of-range labels.
__device__ __forceinline__ bool admit_label(const RowArgs& a, long long t, bool report) {
    if (a.has_ignore && t == a.ignore_index) return false;
    if ((unsigned long long)t >= (unsigned long long)a.C) {  // also catches negatives
        if (report && a.err) atomicOr(a.err, MB200_FLAG_TARGET_RANGE);
        return false;
    }
    return true;
}

constexpr int kRowThreads = 256;

// (1) aligned path: warp per row, 16-byte streaming vector loads from global memory.  The label of the warp's NEXT
// row is requested before the current row is reduced, so label latency never sits in front of the row loads.
// Measured (profiles/r01_confmat_sweep.txt): with >= 4 resident CTAs/SM this loop runs at the same rate as a pure
// read-only streaming probe over the same 131 MB, i.e. the reduction is completely hidden behind HBM; deeper software
// pipelining or a TMA/bulk-copy ring (1b) buys nothing on top of it.
template <typename T, typename Sink, bool kI64>
__global__ void __launch_bounds__(kRowThreads) rows_vec_kernel(RowArgs a, Sink sink) {
    // Let the next update's grid (launched with programmatic stream serialization, see launch_overlapped) start filling
    // SMs as soon as this grid's CTAs retire, instead of after the whole grid has drained and a launch gap has passed.
    if constexpr (Sink::kOverlapSafe) {
        asm volatile("griddepcontrol.launch_dependents;");
        // Overlapped launch: this CTA may be resident while the previous grid of the stream is still draining; that
        // grid's memory (it may be the producer of our inputs, triggering its dependents early) is only guaranteed
        // visible after griddepcontrol.wait, so no input is touched before it.  (Prefetching the first rows into L2
        // ahead of the wait was measured and is slower: 22.3 vs 20.9 us.)
        if (a.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int nwarps = gridDim.x * wpb;
    const int n = (int)a.n_outer;
    const int nvec = (int)(((long long)a.C * sizeof(T)) >> 4);
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    constexpr bool kLabels = Sink::kNeedsTarget;
    int r = blockIdx.x * wpb + (threadIdx.x >> 5);
    long long t_next = 0;
    if (kLabels && r < n) t_next = fetch_label<kI64>(a, r);
    for (; r < n; r += nwarps) {
        const long long t = t_next;
        if (kLabels && r + nwarps < n) t_next = fetch_label<kI64>(a, r + nwarps);
        if (kLabels && !admit_label(a, t, lane == 0)) continue;  // ignored rows are never read
        const GlobalVecLoader load{reinterpret_cast<const uint4*>(preds + (size_t)r * a.C)};
        const int p = warp_row_argmax_vec<T>(load, nvec, lane);
        if (lane == 0) sink.row(loc, r, t, p);
    }
    sink.finish(loc);
}

// (1b) aligned path, bulk-copy pipeline: one producer thread streams tiles of kRows whole rows into a shared-memory
// ring with `cp.async.bulk` (1-D TMA, completion counted on an mbarrier); kRows consumer warps each reduce one staged
// row per tile.  HBM requests stay in flight for the whole ring depth and cost no registers.
constexpr int kBulkMaxStages = 8;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

template <typename T, typename Sink, bool kI64, int kRows>
__global__ void __launch_bounds__((kRows + 1) * 32, 1) rows_bulk_kernel(RowArgs a, Sink sink, int stages) {
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    __shared__ __align__(8) unsigned long long full_bar[kBulkMaxStages];
    __shared__ __align__(8) unsigned long long empty_bar[kBulkMaxStages];
    typename Sink::Local loc;
    sink.init(loc);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned row_bytes = (unsigned)a.C * (unsigned)sizeof(T);
    const unsigned stage_bytes = row_bytes * kRows;
    const int nvec = (int)(row_bytes >> 4);
    const int n = (int)a.n_outer;
    const int n_tiles = (n + kRows - 1) / kRows;
    const unsigned char* __restrict__ gbase = reinterpret_cast<const unsigned char*>(a.preds);
    constexpr bool kLabels = Sink::kNeedsTarget;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kRows);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kRows) {
        // ===== producer: one elected lane keeps the ring full =====
        if (lane == 0) {
            int s = 0;
            unsigned phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&empty_bar[s], phase ^ 1u);  // first lap: passes immediately
                const int row0 = tile * kRows;
                const int rows = min(kRows, n - row0);
                const unsigned bytes = (unsigned)rows * row_bytes;
                mbar_expect_tx(&full_bar[s], bytes);
                bulk_g2s(bulk_smem + (size_t)s * stage_bytes, gbase + (size_t)row0 * row_bytes, bytes, &full_bar[s]);
                if (++s == stages) {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
    } else {
        // ===== consumers: warp w reduces row w of every tile; labels are fetched one tile ahead =====
        int s = 0;
        unsigned phase = 0;
        int tile = blockIdx.x;
        long long t_next = 0;
        if (kLabels && tile < n_tiles && tile * kRows + warp < n) t_next = fetch_label<kI64>(a, tile * kRows + warp);
        for (; tile < n_tiles; tile += gridDim.x) {
            const int r = tile * kRows + warp;
            const long long t = t_next;
            const int rn = (tile + (int)gridDim.x) * kRows + warp;
            if (kLabels && rn < n) t_next = fetch_label<kI64>(a, rn);
            mbar_wait(&full_bar[s], phase);
            int p = 0;
            const bool live = r < n && (!kLabels || admit_label(a, t, lane == 0));
            if (live) {
                const SharedVecLoader load{
                    reinterpret_cast<const uint4*>(bulk_smem + (size_t)s * stage_bytes + (size_t)warp * row_bytes)};
                p = warp_row_argmax_vec<T>(load, nvec, lane);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty_bar[s]);  // this warp's row of the stage now lives in registers / is reduced
                if (live) sink.row(loc, r, t, p);
            }
            if (++s == stages) {
                s = 0;
                phase ^= 1u;
            }
        }
    }
    sink.finish(loc);
}

// (2) warp per row, scalar loads
template <typename T, typename Sink, bool kI64>
__global__ void __launch_bounds__(kRowThreads) rows_scalar_kernel(RowArgs a, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long nwarps = (long long)gridDim.x * wpb;
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < a.n_outer; r += nwarps) {
        long long t = 0;
        if (Sink::kNeedsTarget) {
            t = fetch_label<kI64>(a, r);
            if (!admit_label(a, t, lane == 0)) continue;
        }
        const int p = warp_row_argmax_scalar<T>(preds + r * a.C, a.C, lane);
        if (lane == 0) sink.row(loc, r, t, p);
    }
    sink.finish(loc);
}

// (2b) top-k "refined" prediction (functional/classification/stat_scores.py:347-368 `_refine_preds_oh`): the effective
// label is the target when it is among the k best scores of the row, else the argmax.  Warp per row, scalar loads;
// "among the k best" = fewer than k columns beat the target's score (greater key, or equal key and lower index —
// the order a stable descending sort gives).
template <typename T, typename Sink, bool kI64>
__global__ void __launch_bounds__(kRowThreads) rows_topk_kernel(RowArgs a, Sink sink, int top_k) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long nwarps = (long long)gridDim.x * wpb;
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < a.n_outer; r += nwarps) {
        const long long t = fetch_label<kI64>(a, r);
        if (!admit_label(a, t, lane == 0)) continue;
        const T* __restrict__ row = preds + r * a.C;
        const unsigned long long kt = order_key<T>(row[t]);
        int beats = 0;
        for (int c = lane; c < a.C; c += kWarp) {
            const unsigned long long k = order_key<T>(row[c]);
            beats += (k > kt) || (k == kt && c < (int)t);
        }
        beats = __reduce_add_sync(kFull, beats);
        int p = (int)t;
        if (beats >= top_k) p = warp_row_argmax_scalar<T>(row, a.C, lane);
        if (lane == 0) sink.row(loc, r, t, p);
    }
    sink.finish(loc);
}

// (3) thread per (outer, inner) position, class dim strided by `inner` (also the tiny-C path with inner == 1)
template <typename T, typename Sink, bool kI64>
__global__ void __launch_bounds__(kRowThreads) rows_strided_kernel(RowArgs a, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const long long total = a.n_outer * a.inner;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        long long t = 0;
        if (Sink::kNeedsTarget) {
            t = fetch_label<kI64>(a, i);
            if (!admit_label(a, t, true)) continue;
        }
        const long long n = i / a.inner, x = i - n * a.inner;
        const int p = thread_argmax_strided<T>(preds + (n * a.C) * a.inner + x, a.C, a.inner);
        sink.row(loc, i, t, p);
    }
    sink.finish(loc);
}

// (4) integer label predictions: thread per sample
template <typename Sink, bool kI64>
__global__ void __launch_bounds__(kRowThreads) labels_kernel(RowArgs a, int preds_dtype, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const long long total = a.n_outer * a.inner;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const long long t = fetch_label<kI64>(a, i);
        if (!admit_label(a, t, true)) continue;
        const long long p = load_label(a.preds, preds_dtype, i);
        if ((unsigned long long)p >= (unsigned long long)a.C) {
            if (a.err) atomicOr(a.err, MB200_FLAG_PREDS_RANGE);
            continue;
        }
        sink.row(loc, i, t, (int)p);
    }
    sink.finish(loc);
}

// =====================================================================================================
// Host dispatch
// =====================================================================================================
extern void count_launch();

static inline int grid_for(long long work_items, int items_per_block, int max_blocks) {
    long long g = (work_items + items_per_block - 1) / items_per_block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// Path override for A/B measurements: MB200_ROWS_PATH=vec|bulk (default vec: it measured faster, see (1)).
static int rows_path_override() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("MB200_ROWS_PATH");
        cached = (e && e[0] == 'b') ? 2 : 1;
    }
    return cached;
}

// MB200_ROWS_OVERLAP: 0 = plain launches; 1 (default) = programmatic dependent launch, the kernel waits for the previous
// grid's completion before its first input load (always correct); 2 = no wait: consecutive updates overlap drain and
// ramp-up — only valid when the inputs were complete before the PREVIOUS kernel of the stream started (e.g. a replay of
// resident batches), because a foreign producer that triggers its dependents early would otherwise race with the loads.
static int rows_overlap_mode() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("MB200_ROWS_OVERLAP");
        cached = (e && e[0] >= '0' && e[0] <= '2') ? (e[0] - '0') : 1;
    }
    return cached;
}
static bool rows_overlap_enabled() { return rows_overlap_mode() != 0; }

// Launch with cudaLaunchAttributeProgrammaticStreamSerialization: the grid may begin while the previous kernel of the
// stream is still draining (that kernel opts in with griddepcontrol.launch_dependents).  Only used for launches whose
// sole shared data are commutative atomics on the state.
template <typename Kernel, typename... Args>
static cudaError_t launch_overlapped(Kernel kern, int grid, int threads, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// One-wave grid size of a kernel: resident CTAs per SM x SMs.  The occupancy query takes a driver lock and costs 1-2 us —
// a tenth of a small update's host time — so its answer is kept per (kernel, block size, dynamic smem, device).
template <typename Kernel>
static int resident_blocks(Kernel k, int threads, size_t smem) {
    struct Entry { const void* fn; int threads; size_t smem; int dev; int blocks; };
    static thread_local Entry cache[8] = {};
    static thread_local int next = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    const void* fn = reinterpret_cast<const void*>(k);
    for (const Entry& e : cache)
        if (e.fn == fn && e.threads == threads && e.smem == smem && e.dev == dev && e.blocks > 0) return e.blocks;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int blocks = per_sm * sm_count();
    if (dev >= 0) {
        cache[next] = Entry{fn, threads, smem, dev, blocks};
        next = (next + 1) & 7;
    }
    return blocks;
}

template <typename T, typename Sink, bool kI64, int kRows>
static int launch_bulk(const RowArgs& a, Sink sink, cudaStream_t st, bool& launched) {
    const size_t stage_bytes = (size_t)a.C * sizeof(T) * kRows;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > kBulkMaxStages) stages = kBulkMaxStages;
    launched = false;
    if (stages < 3 || a.n_outer < 4 * kRows) return 0;
    auto kern = rows_bulk_kernel<T, Sink, kI64, kRows>;
    MB200_CUDA_OK(ensure_dynamic_smem(kern, 227 * 1024 - 512));
    const long long tiles = (a.n_outer + kRows - 1) / kRows;
    int grid = sm_count();
    if (tiles < grid) grid = (int)tiles;
    kern<<<grid, (kRows + 1) * 32, (size_t)stages * stage_bytes, st>>>(a, sink, stages);
    launched = true;
    return 0;
}

template <typename T, typename Sink, bool kI64>
static int launch_rows(const RowArgs& a, Sink sink, size_t smem, cudaStream_t st) {
    const size_t row_bytes = (size_t)a.C * sizeof(T);
    const bool vec_ok = sizeof(T) <= 4 && a.inner == 1 && a.C >= 32 && (row_bytes % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.preds) & 15) == 0);
    if (a.inner == 1 && a.C >= 32) {
        if (vec_ok) {
            if constexpr (sizeof(T) <= 4) {
                const int ov = rows_path_override();
                bool launched = false;
                if (smem == 0 && ov == 2) {
                    if (int rc = launch_bulk<T, Sink, kI64, 16>(a, sink, st, launched)) return rc;
                }
                if (!launched) {
                    auto kern = rows_vec_kernel<T, Sink, kI64>;
                    const int grid = grid_for(a.n_outer, kRowThreads / 32, resident_blocks(kern, kRowThreads, smem));
                    if constexpr (Sink::kOverlapSafe) {
                        if (rows_overlap_enabled()) {
                            RowArgs ao = a;
                            ao.pdl_wait = rows_overlap_mode() == 1 || Sink::kMustWait;  // see sinks.cuh
                            MB200_CUDA_OK(launch_overlapped(kern, grid, kRowThreads, smem, st, ao, sink));
                        } else {
                            kern<<<grid, kRowThreads, smem, st>>>(a, sink);
                        }
                    } else {
                        kern<<<grid, kRowThreads, smem, st>>>(a, sink);
                    }
                }
            }
        } else {
            auto kern = rows_scalar_kernel<T, Sink, kI64>;
            const int grid = grid_for(a.n_outer, kRowThreads / 32, resident_blocks(kern, kRowThreads, smem));
            kern<<<grid, kRowThreads, smem, st>>>(a, sink);
        }
    } else {
        auto kern = rows_strided_kernel<T, Sink, kI64>;
        const int grid = grid_for(a.n_outer * a.inner, kRowThreads, resident_blocks(kern, kRowThreads, smem));
        kern<<<grid, kRowThreads, smem, st>>>(a, sink);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "row kernel launch");
}

template <typename Sink, bool kI64>
static int dispatch_rows_t(int preds_dtype, int preds_has_class_dim, const RowArgs& a, Sink sink, size_t smem,
                           cudaStream_t st) {
    if (!preds_has_class_dim) {
        if constexpr (Sink::kNeedsTarget) {
            MB200_REQUIRE(preds_dtype >= MB200_I64 && preds_dtype <= MB200_BOOL,
                          "label-format preds must have an integer dtype (got dtype tag %d)", preds_dtype);
            auto kern = labels_kernel<Sink, kI64>;
            const int grid = grid_for(a.n_outer * a.inner, kRowThreads, resident_blocks(kern, kRowThreads, smem));
            kern<<<grid, kRowThreads, smem, st>>>(a, preds_dtype, sink);
            count_launch();
            return check_cuda(cudaGetLastError(), "labels kernel launch");
        } else {
            set_error("argmax needs a class dimension");
            return MB200_ERR_INVALID;
        }
    }
    switch (preds_dtype) {
        case MB200_BF16: return launch_rows<__nv_bfloat16, Sink, kI64>(a, sink, smem, st);
        case MB200_F16: return launch_rows<__half, Sink, kI64>(a, sink, smem, st);
        case MB200_F32: return launch_rows<float, Sink, kI64>(a, sink, smem, st);
        case MB200_F64: return launch_rows<double, Sink, kI64>(a, sink, smem, st);
        default:
            set_error("preds with a class dimension must be floating point (got dtype tag %d)", preds_dtype);
            return MB200_ERR_INVALID;
    }
}

template <typename Sink>
static int dispatch_rows(int preds_dtype, int preds_has_class_dim, const RowArgs& a, Sink sink, size_t smem,
                         cudaStream_t st) {
    MB200_REQUIRE(a.n_outer * a.inner < (1ll << 31), "more than 2^31-1 rows per call are not supported (got %lld)",
                  (long long)(a.n_outer * a.inner));
    if (a.target_dtype == MB200_I64) return dispatch_rows_t<Sink, true>(preds_dtype, preds_has_class_dim, a, sink, smem, st);
    return dispatch_rows_t<Sink, false>(preds_dtype, preds_has_class_dim, a, sink, smem, st);
}

static int validate_common(const void* preds, const void* target, int target_dtype, int64_t n_outer,
                           int64_t num_classes, int64_t inner, bool need_target) {
    MB200_REQUIRE(n_outer >= 0 && inner >= 1, "negative sizes (n_outer=%lld inner=%lld)", (long long)n_outer,
                  (long long)inner);
    MB200_REQUIRE(num_classes >= 1 && num_classes <= (1ll << 24),
                  "num_classes must be in [1, 2^24] (got %lld)", (long long)num_classes);
    if (n_outer * inner > 0) {
        MB200_REQUIRE(preds != nullptr, "preds is NULL");
        if (need_target) MB200_REQUIRE(target != nullptr, "target is NULL");
    }
    if (need_target)
        MB200_REQUIRE(target_dtype >= MB200_I64 && target_dtype <= MB200_BOOL,
                      "target must have an integer dtype (got dtype tag %d)", target_dtype);
    return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_multiclass_confmat_update(const void* preds, int preds_dtype, int preds_has_class_dim,
                                               const void* target, int target_dtype, int64_t n_outer,
                                               int64_t num_classes, int64_t inner, int has_ignore_index,
                                               int64_t ignore_index, int64_t* confmat, uint32_t* err_flag,
                                               void* stream) {
    if (int rc = validate_common(preds, target, target_dtype, n_outer, num_classes, inner, true)) return rc;
    MB200_REQUIRE(confmat != nullptr, "confmat is NULL");
    MB200_REQUIRE(num_classes <= 46340, "num_classes^2 must fit in int32 indexing (got %lld)",
                  (long long)num_classes);
    if (n_outer * inner == 0) return 0;
    RowArgs a{preds, target, target_dtype, n_outer, (int)num_classes, inner, has_ignore_index,
              ignore_index, err_flag};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool priv = num_classes * num_classes <= 4096 && n_outer * inner >= 4096;
    if (priv) {
        ConfmatSink<true> s{reinterpret_cast<long long*>(confmat), (int)num_classes};
        return dispatch_rows(preds_dtype, preds_has_class_dim, a, s,
                             (size_t)(num_classes * num_classes) * sizeof(unsigned), st);
    }
    ConfmatSink<false> s{reinterpret_cast<long long*>(confmat), (int)num_classes};
    return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, 0, st);
}

extern "C" int mb200_multiclass_stat_scores_update(const void* preds, int preds_dtype, int preds_has_class_dim,
                                                   const void* target, int target_dtype, int64_t n_outer,
                                                   int64_t num_classes, int64_t inner, int has_ignore_index,
                                                   int64_t ignore_index, int micro, int64_t* tp, int64_t* fp,
                                                   int64_t* tn, int64_t* fn, int64_t* workspace,
                                                   uint32_t* err_flag, void* stream) {
    if (int rc = validate_common(preds, target, target_dtype, n_outer, num_classes, inner, true)) return rc;
    MB200_REQUIRE(tp && fp && tn && fn && workspace, "state / workspace pointer is NULL");
    if (n_outer * inner == 0) return 0;
    RowArgs a{preds, target, target_dtype, n_outer, (int)num_classes, inner, has_ignore_index,
              ignore_index, err_flag};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // Shared-memory privatisation pays when many rows hit few class bins (L2 atomics on a handful of addresses
    // serialise).  With hundreds of classes the ~2 REDs per row spread over 3C L2-resident words cost nothing, while
    // zero-filling and flushing 3C counters per CTA does (cfg2 shape: 37 us privatised vs the 22 us of the confmat kernel).
    const bool priv = !micro && num_classes <= 256 && n_outer * inner >= 4096 && n_outer * inner >= 16 * num_classes;
    if (priv) {
        StatsSink<true> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn,
                          (long long*)workspace, (int)num_classes, micro};
        return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, (size_t)(3 * num_classes) * sizeof(unsigned),
                             st);
    }
    // Large launches: rows only RED into the workspace, the fold follows as its own one-CTA kernel (see sinks.cuh kDeferFold);
    // small ones keep the single launch (an extra launch costs more host time than the fold tail costs device time).
    if (preds_has_class_dim && inner == 1 && n_outer * num_classes >= (1ll << 24) && rows_overlap_enabled()) {
        StatsSink<false, true> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn, (long long*)workspace,
                                 (int)num_classes, micro};
        if (int rc = dispatch_rows(preds_dtype, preds_has_class_dim, a, s, 0, st)) return rc;
        MB200_CUDA_OK(launch_overlapped(stats_fold_kernel, 1, 1024, 0, st, (long long*)tp, (long long*)fp, (long long*)tn,
                                        (long long*)fn, (long long*)workspace, (int)num_classes, micro));
        count_launch();
        return check_cuda(cudaGetLastError(), "stat-scores fold launch");
    }
    StatsSink<false> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn, (long long*)workspace,
                       (int)num_classes, micro};
    return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, 0, st);
}

template <typename Sink, bool kI64>
static int launch_topk_t(int preds_dtype, const RowArgs& a, Sink sink, size_t smem, int top_k, cudaStream_t st) {
#define MB200_TOPK(T)                                                                                        \
    {                                                                                                        \
        auto kern = rows_topk_kernel<T, Sink, kI64>;                                                         \
        const int grid = grid_for(a.n_outer, kRowThreads / 32, resident_blocks(kern, kRowThreads, smem));   \
        kern<<<grid, kRowThreads, smem, st>>>(a, sink, top_k);                                               \
    }
    switch (preds_dtype) {
        case MB200_BF16: MB200_TOPK(__nv_bfloat16) break;
        case MB200_F16: MB200_TOPK(__half) break;
        case MB200_F32: MB200_TOPK(float) break;
        case MB200_F64: MB200_TOPK(double) break;
        default: set_error("top-k needs floating scores (dtype tag %d)", preds_dtype); return MB200_ERR_INVALID;
    }
#undef MB200_TOPK
    count_launch();
    return check_cuda(cudaGetLastError(), "top-k kernel launch");
}

extern "C" int mb200_multiclass_stat_scores_topk_update(const void* preds, int preds_dtype, const void* target,
                                                        int target_dtype, int64_t n, int64_t num_classes, int64_t top_k,
                                                        int has_ignore_index, int64_t ignore_index, int64_t* tp,
                                                        int64_t* fp, int64_t* tn, int64_t* fn, int64_t* workspace,
                                                        uint32_t* err_flag, void* stream) {
    if (int rc = validate_common(preds, target, target_dtype, n, num_classes, 1, true)) return rc;
    MB200_REQUIRE(tp && fp && tn && fn && workspace, "state / workspace pointer is NULL");
    MB200_REQUIRE(top_k >= 1 && top_k <= num_classes, "top_k must be in [1, num_classes]");
    if (n == 0) return 0;
    RowArgs a{preds, target, target_dtype, n, (int)num_classes, 1, has_ignore_index, ignore_index, err_flag};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    StatsSink<false> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn, (long long*)workspace,
                       (int)num_classes, 0};
    if (target_dtype == MB200_I64) return launch_topk_t<StatsSink<false>, true>(preds_dtype, a, s, 0, (int)top_k, st);
    return launch_topk_t<StatsSink<false>, false>(preds_dtype, a, s, 0, (int)top_k, st);
}

extern "C" int mb200_multiclass_stat_scores_samplewise(const void* preds, int preds_dtype, int preds_has_class_dim,
                                                       const void* target, int target_dtype, int64_t n_outer,
                                                       int64_t num_classes, int64_t inner, int has_ignore_index,
                                                       int64_t ignore_index, int64_t* counts, int64_t* n_valid,
                                                       uint32_t* err_flag, void* stream) {
    if (int rc = validate_common(preds, target, target_dtype, n_outer, num_classes, inner, true)) return rc;
    MB200_REQUIRE(counts && n_valid, "NULL pointer");
    if (n_outer * inner == 0) return 0;
    RowArgs a{preds, target, target_dtype, n_outer, (int)num_classes, inner, has_ignore_index, ignore_index, err_flag};
    SamplewiseSink s{(long long*)counts, (long long*)n_valid, n_outer, inner, (int)num_classes};
    return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mb200_argmax_rows(const void* preds, int preds_dtype, int64_t n_outer, int64_t num_classes,
                                 int64_t inner, int64_t* out, void* stream) {
    if (int rc = validate_common(preds, nullptr, MB200_I64, n_outer, num_classes, inner, false)) return rc;
    MB200_REQUIRE(out != nullptr || n_outer * inner == 0, "out is NULL");
    if (n_outer * inner == 0) return 0;
    RowArgs a{preds, nullptr, MB200_I64, n_outer, (int)num_classes, inner, 0, 0, nullptr};
    ArgmaxOutSink s{reinterpret_cast<long long*>(out)};
    return dispatch_rows(preds_dtype, 1, a, s, 0, reinterpret_cast<cudaStream_t>(stream));
}
