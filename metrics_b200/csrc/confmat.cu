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
#include "common.cuh"

namespace mb200 {

// =====================================================================================================
// Per-dtype row traits for the vectorised path.  A "vector" is 16 bytes.
// =====================================================================================================
template <typename T>
struct RowTraits;

template <>
struct RowTraits<__nv_bfloat16> {
    static constexpr int EPV = 8;
    static constexpr unsigned kFill = 0xff80ff80u;  // two bf16 -inf
    using Acc = __nv_bfloat162;
    static __device__ __forceinline__ Acc as2(unsigned u) { return *reinterpret_cast<Acc*>(&u); }
    static __device__ __forceinline__ Acc acc_init() { return as2(kFill); }
    static __device__ __forceinline__ void accumulate(Acc& a, const uint4& v) {
        a = __hmax2_nan(a, __hmax2_nan(__hmax2_nan(as2(v.x), as2(v.y)), __hmax2_nan(as2(v.z), as2(v.w))));
    }
    static __device__ __forceinline__ unsigned lane_key(Acc a) {
        return f32_order_key(__bfloat162float(__hmax_nan(a.x, a.y)));
    }
    // Two words of four 0xFF/0x00 bytes each, byte order == column order inside the vector.
    static __device__ __forceinline__ void match_words(const uint4& v, unsigned rowkey, unsigned& w0,
                                                       unsigned& w1) {
        unsigned e0, e1, e2, e3;
        if (rowkey == 0xffffffffu) {  // NaN row (warp-uniform): a NaN is the only thing != itself
            e0 = ~__heq2_mask(as2(v.x), as2(v.x));
            e1 = ~__heq2_mask(as2(v.y), as2(v.y));
            e2 = ~__heq2_mask(as2(v.z), as2(v.z));
            e3 = ~__heq2_mask(as2(v.w), as2(v.w));
        } else {
            const Acc m = __float2bfloat162_rn(f32_from_order_key(rowkey));
            e0 = __heq2_mask(as2(v.x), m);
            e1 = __heq2_mask(as2(v.y), m);
            e2 = __heq2_mask(as2(v.z), m);
            e3 = __heq2_mask(as2(v.w), m);
        }
        w0 = __byte_perm(e0, e1, 0x6420);
        w1 = __byte_perm(e2, e3, 0x6420);
    }
    static __device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
};

template <>
struct RowTraits<__half> {
    static constexpr int EPV = 8;
    static constexpr unsigned kFill = 0xfc00fc00u;  // two f16 -inf
    using Acc = __half2;
    static __device__ __forceinline__ Acc as2(unsigned u) { return *reinterpret_cast<Acc*>(&u); }
    static __device__ __forceinline__ Acc acc_init() { return as2(kFill); }
    static __device__ __forceinline__ void accumulate(Acc& a, const uint4& v) {
        a = __hmax2_nan(a, __hmax2_nan(__hmax2_nan(as2(v.x), as2(v.y)), __hmax2_nan(as2(v.z), as2(v.w))));
    }
    static __device__ __forceinline__ unsigned lane_key(Acc a) {
        return f32_order_key(__half2float(__hmax_nan(__low2half(a), __high2half(a))));
    }
    static __device__ __forceinline__ void match_words(const uint4& v, unsigned rowkey, unsigned& w0,
                                                       unsigned& w1) {
        unsigned e0, e1, e2, e3;
        if (rowkey == 0xffffffffu) {
            e0 = ~__heq2_mask(as2(v.x), as2(v.x));
            e1 = ~__heq2_mask(as2(v.y), as2(v.y));
            e2 = ~__heq2_mask(as2(v.z), as2(v.z));
            e3 = ~__heq2_mask(as2(v.w), as2(v.w));
        } else {
            const Acc m = __float2half2_rn(f32_from_order_key(rowkey));
            e0 = __heq2_mask(as2(v.x), m);
            e1 = __heq2_mask(as2(v.y), m);
            e2 = __heq2_mask(as2(v.z), m);
            e3 = __heq2_mask(as2(v.w), m);
        }
        w0 = __byte_perm(e0, e1, 0x6420);
        w1 = __byte_perm(e2, e3, 0x6420);
    }
    static __device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
};

__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

template <>
struct RowTraits<float> {
    static constexpr int EPV = 4;
    static constexpr unsigned kFill = 0xff800000u;  // -inf
    using Acc = float;
    static __device__ __forceinline__ Acc acc_init() { return __uint_as_float(kFill); }
    static __device__ __forceinline__ void accumulate(Acc& a, const uint4& v) {
        a = fmax_nan(a, fmax_nan(fmax_nan(__uint_as_float(v.x), __uint_as_float(v.y)),
                                 fmax_nan(__uint_as_float(v.z), __uint_as_float(v.w))));
    }
    static __device__ __forceinline__ unsigned lane_key(Acc a) { return f32_order_key(a); }
    static __device__ __forceinline__ void match_words(const uint4& v, unsigned rowkey, unsigned& w0,
                                                       unsigned& w1) {
        const float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z),
                    x3 = __uint_as_float(v.w);
        unsigned w = 0;
        if (rowkey == 0xffffffffu) {
            w |= (x0 != x0) ? 0x000000ffu : 0u;
            w |= (x1 != x1) ? 0x0000ff00u : 0u;
            w |= (x2 != x2) ? 0x00ff0000u : 0u;
            w |= (x3 != x3) ? 0xff000000u : 0u;
        } else {
            const float m = f32_from_order_key(rowkey);
            w |= (x0 == m) ? 0x000000ffu : 0u;
            w |= (x1 == m) ? 0x0000ff00u : 0u;
            w |= (x2 == m) ? 0x00ff0000u : 0u;
            w |= (x3 == m) ? 0xff000000u : 0u;
        }
        w0 = w;
        w1 = 0;
    }
    static __device__ __forceinline__ float to_f32(float x) { return x; }
};

// Generic order key used by the scalar paths (any float dtype, incl. f64).
template <typename T>
__device__ __forceinline__ unsigned long long order_key(T x) {
    return (unsigned long long)f32_order_key(RowTraits<T>::to_f32(x));
}
template <>
__device__ __forceinline__ unsigned long long order_key<double>(double x) {
    return f64_order_key(x);
}

// =====================================================================================================
// Warp-per-row argmax, vectorised.  Requires: row base 16-byte aligned and C * sizeof(T) % 16 == 0.
// All lanes return the same column index.
// =====================================================================================================
constexpr int kVPL = 4;  // vectors per lane per chunk  -> chunk = 32 lanes * 4 * 16 B = 2 KiB of one row

template <typename T>
__device__ __forceinline__ int warp_row_argmax_vec(const T* __restrict__ row, int nvec, int lane) {
    using TR = RowTraits<T>;
    const uint4* __restrict__ rv = reinterpret_cast<const uint4*>(row);
    unsigned best_key = 0;
    int best_col = 0;
    bool have = false;
    for (int cv = 0; cv < nvec; cv += kVPL * kWarp) {
        uint4 v[kVPL];
#pragma unroll
        for (int j = 0; j < kVPL; ++j) {
            const int vi = cv + j * kWarp + lane;
            if (vi < nvec) {
                v[j] = ld_stream16(rv + vi);
            } else {
                v[j] = make_uint4(TR::kFill, TR::kFill, TR::kFill, TR::kFill);
            }
        }
        typename TR::Acc acc = TR::acc_init();
#pragma unroll
        for (int j = 0; j < kVPL; ++j) TR::accumulate(acc, v[j]);
        const unsigned ckey = __reduce_max_sync(kFull, TR::lane_key(acc));

        unsigned bw0 = 0, bw1 = 0;
        int bj = 0;
#pragma unroll
        for (int j = kVPL - 1; j >= 0; --j) {
            unsigned w0, w1;
            TR::match_words(v[j], ckey, w0, w1);
            if ((w0 | w1) != 0u) {
                bw0 = w0;
                bw1 = w1;
                bj = j;
            }
        }
        unsigned col = 0x7fffffffu;
        if ((bw0 | bw1) != 0u) {
            const int e = bw0 ? ((__ffs(bw0) - 1) >> 3) : (4 + ((__ffs(bw1) - 1) >> 3));
            col = (unsigned)((cv + bj * kWarp + lane) * TR::EPV + e);
        }
        const unsigned ccol = __reduce_min_sync(kFull, col);
        // strictly-greater keeps the earliest chunk on ties (and the first NaN chunk: all NaN keys are equal)
        if (!have || ckey > best_key) {
            best_key = ckey;
            best_col = (int)ccol;
            have = true;
        }
    }
    return best_col;
}

// Warp-per-row argmax, scalar loads (any alignment / any C).  All lanes return the same column.
template <typename T>
__device__ __forceinline__ int warp_row_argmax_scalar(const T* __restrict__ row, int C, int lane) {
    unsigned long long bk = 0;
    int bc = 0x7fffffff;
    for (int c = lane; c < C; c += kWarp) {
        const unsigned long long k = order_key<T>(row[c]);
        if (bc == 0x7fffffff || k > bk) {
            bk = k;
            bc = c;
        }
    }
    // lanes without any element (C < 32) carry key 0 / col INT_MAX and can never win against a real element
    unsigned long long mk = bk;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFull, mk, o);
        mk = other > mk ? other : mk;
    }
    const unsigned col = (bk == mk && bc != 0x7fffffff) ? (unsigned)bc : 0x7fffffffu;
    return (int)__reduce_min_sync(kFull, col);
}

// Thread-per-(outer, inner) argmax over a strided class dimension: element c at base[c * stride].
template <typename T>
__device__ __forceinline__ int thread_argmax_strided(const T* __restrict__ base, int C, long long stride) {
    unsigned long long bk = order_key<T>(base[0]);
    int bc = 0;
    for (int c = 1; c < C; ++c) {
        const unsigned long long k = order_key<T>(base[(long long)c * stride]);
        if (k > bk) {
            bk = k;
            bc = c;
        }
    }
    return bc;
}

// =====================================================================================================
// Sinks: what happens with one (target, prediction) pair.
// =====================================================================================================
struct ArgmaxOutSink {
    long long* out;
    struct Local {};
    __device__ __forceinline__ void block_init() {}
    __device__ __forceinline__ void init(Local&) {}
    __device__ __forceinline__ void row(Local&, long long idx, long long /*t*/, int p) { out[idx] = p; }
    __device__ __forceinline__ void finish(Local&) {}
    static constexpr bool kNeedsTarget = false;
};

// confmat[t, p] += 1 straight into the (L2-resident) state; optional shared-memory privatisation for tiny C.
template <bool kSmem>
struct ConfmatSink {
    long long* confmat;
    int C;
    struct Local {};
    static constexpr bool kNeedsTarget = true;
    __device__ __forceinline__ void block_init() {
        if (kSmem) {
            extern __shared__ unsigned sh_bins[];
            for (int i = threadIdx.x; i < C * C; i += blockDim.x) sh_bins[i] = 0;
            __syncthreads();
        }
    }
    __device__ __forceinline__ void init(Local&) {}
    __device__ __forceinline__ void row(Local&, long long, long long t, int p) {
        if (kSmem) {
            extern __shared__ unsigned sh_bins[];
            atomicAdd(&sh_bins[(int)t * C + p], 1u);
        } else {
            red_add_u64(confmat + t * C + p, 1ull);
        }
    }
    __device__ __forceinline__ void finish(Local&) {
        if (kSmem) {
            extern __shared__ unsigned sh_bins[];
            __syncthreads();
            for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
                const unsigned v = sh_bins[i];
                if (v) red_add_u64(confmat + i, v);
            }
        }
    }
};

// tp/fp/fn deltas go to a zeroed workspace; the last block to finish folds them (and tn) into the states and
// re-zeroes the workspace.  ws layout: [0,C) dtp | [C,2C) dfp | [2C,3C) dfn | [3C] n_valid | [3C+1] ticket.
// micro: ws[0] = #match, ws[1] = #mismatch.
template <bool kSmem>
struct StatsSink {
    long long *tp, *fp, *tn, *fn, *ws;
    int C;
    int micro;
    struct Local {
        unsigned n_valid, n_match;
    };
    static constexpr bool kNeedsTarget = true;
    __device__ __forceinline__ void block_init() {
        if (kSmem) {
            extern __shared__ unsigned sh_bins[];
            for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sh_bins[i] = 0;
            __syncthreads();
        }
    }
    __device__ __forceinline__ void init(Local& l) { l.n_valid = 0, l.n_match = 0; }
    __device__ __forceinline__ void row(Local& l, long long, long long t, int p) {
        l.n_valid++;
        if (micro) {
            l.n_match += ((long long)p == t);
            return;
        }
        if (kSmem) {
            extern __shared__ unsigned sh_bins[];
            if ((long long)p == t) {
                atomicAdd(&sh_bins[p], 1u);
            } else {
                atomicAdd(&sh_bins[C + p], 1u);
                atomicAdd(&sh_bins[2 * C + (int)t], 1u);
            }
        } else {
            if ((long long)p == t) {
                red_add_u64(ws + p, 1ull);
            } else {
                red_add_u64(ws + C + p, 1ull);
                red_add_u64(ws + 2 * C + t, 1ull);
            }
        }
    }
    __device__ __forceinline__ void finish(Local& l) {
        // per-warp totals -> one atomic per warp
        const unsigned nv = __reduce_add_sync(kFull, l.n_valid);
        const unsigned nm = __reduce_add_sync(kFull, l.n_match);
        if ((threadIdx.x & 31) == 0) {
            if (nv) red_add_u64(ws + 3 * C, nv);
            if (micro) {
                if (nm) red_add_u64(ws + 0, nm);
                if (nv - nm) red_add_u64(ws + 1, nv - nm);
            }
        }
        if (kSmem && !micro) {
            extern __shared__ unsigned sh_bins[];
            __syncthreads();
            for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
                const unsigned v = sh_bins[i];
                if (v) red_add_u64(ws + i, v);
            }
        }
        // ---- last-block fold -------------------------------------------------------------------
        __shared__ int is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long ticket =
                atomicAdd(reinterpret_cast<unsigned long long*>(ws + 3 * C + 1), 1ull);
            is_last = (ticket == (unsigned long long)gridDim.x - 1ull);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        long long* vws = ws;
        const long long n_valid = __ldcg(ws + 3 * C);
        if (micro) {
            if (threadIdx.x == 0) {
                const long long m = __ldcg(ws + 0), mm = __ldcg(ws + 1);
                tp[0] += m;
                fp[0] += mm;
                fn[0] += mm;
                tn[0] += (long long)C * n_valid - (m + 2 * mm);
                vws[0] = 0;
                vws[1] = 0;
            }
        } else {
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                const long long a = __ldcg(ws + c), b = __ldcg(ws + C + c), d = __ldcg(ws + 2 * C + c);
                if (a | b | d) {
                    tp[c] += a;
                    fp[c] += b;
                    fn[c] += d;
                    vws[c] = 0;
                    vws[C + c] = 0;
                    vws[2 * C + c] = 0;
                }
                tn[c] += n_valid - (a + b + d);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            vws[3 * C] = 0;
            vws[3 * C + 1] = 0;
        }
    }
};

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
};

__device__ __forceinline__ bool admit_target(const RowArgs& a, long long idx, long long& t, bool report) {
    t = load_label(a.target, a.target_dtype, idx);
    if (a.has_ignore && t == a.ignore_index) return false;
    if (t < 0 || t >= a.C) {
        if (report && a.err) atomicOr(a.err, MB200_FLAG_TARGET_RANGE);
        return false;
    }
    return true;
}

constexpr int kRowThreads = 256;

// (1) aligned fast path: warp per row, 16-byte vectors
template <typename T, typename Sink>
__global__ void __launch_bounds__(kRowThreads) rows_vec_kernel(RowArgs a, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const int lane = threadIdx.x & 31;
    const long long wpb = blockDim.x >> 5;
    const long long nwarps = (long long)gridDim.x * wpb;
    const int nvec = (int)(((long long)a.C * sizeof(T)) >> 4);
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < a.n_outer; r += nwarps) {
        long long t = 0;
        if (Sink::kNeedsTarget && !admit_target(a, r, t, lane == 0)) continue;  // ignored rows are never read
        const int p = warp_row_argmax_vec<T>(preds + r * a.C, nvec, lane);
        if (lane == 0) sink.row(loc, r, t, p);
    }
    sink.finish(loc);
}

// (2) warp per row, scalar loads
template <typename T, typename Sink>
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
        if (Sink::kNeedsTarget && !admit_target(a, r, t, lane == 0)) continue;
        const int p = warp_row_argmax_scalar<T>(preds + r * a.C, a.C, lane);
        if (lane == 0) sink.row(loc, r, t, p);
    }
    sink.finish(loc);
}

// (3) thread per (outer, inner) position, class dim strided by `inner` (also the tiny-C path with inner == 1)
template <typename T, typename Sink>
__global__ void __launch_bounds__(kRowThreads) rows_strided_kernel(RowArgs a, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const long long total = a.n_outer * a.inner;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const T* __restrict__ preds = reinterpret_cast<const T*>(a.preds);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        long long t = 0;
        if (Sink::kNeedsTarget && !admit_target(a, i, t, true)) continue;
        const long long n = i / a.inner, x = i - n * a.inner;
        const int p = thread_argmax_strided<T>(preds + (n * a.C) * a.inner + x, a.C, a.inner);
        sink.row(loc, i, t, p);
    }
    sink.finish(loc);
}

// (4) integer label predictions: thread per sample
template <typename Sink>
__global__ void __launch_bounds__(kRowThreads) labels_kernel(RowArgs a, int preds_dtype, Sink sink) {
    sink.block_init();
    typename Sink::Local loc;
    sink.init(loc);
    const long long total = a.n_outer * a.inner;
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        long long t = 0;
        if (!admit_target(a, i, t, true)) continue;
        const long long p = load_label(a.preds, preds_dtype, i);
        if (p < 0 || p >= a.C) {
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

static inline int grid_for(long long work_items, int items_per_block, int max_waves_blocks) {
    long long g = (work_items + items_per_block - 1) / items_per_block;
    if (g < 1) g = 1;
    if (g > max_waves_blocks) g = max_waves_blocks;
    return (int)g;
}

template <typename T, typename Sink>
static int launch_rows(const RowArgs& a, Sink sink, size_t smem, cudaStream_t st) {
    const int sms = sm_count();
    const int max_blocks = sms * 8;  // 8 CTAs of 256 threads = 64 resident warps per SM
    const size_t row_bytes = (size_t)a.C * sizeof(T);
    const bool vec_ok = sizeof(T) <= 4 && a.inner == 1 && a.C >= 32 && (row_bytes % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.preds) & 15) == 0);
    if (a.inner == 1 && a.C >= 32) {
        const int grid = grid_for(a.n_outer, kRowThreads / 32, max_blocks);
        if (vec_ok) {
            if constexpr (sizeof(T) <= 4) {
                rows_vec_kernel<T, Sink><<<grid, kRowThreads, smem, st>>>(a, sink);
            }
        } else {
            rows_scalar_kernel<T, Sink><<<grid, kRowThreads, smem, st>>>(a, sink);
        }
    } else {
        const int grid = grid_for(a.n_outer * a.inner, kRowThreads, max_blocks);
        rows_strided_kernel<T, Sink><<<grid, kRowThreads, smem, st>>>(a, sink);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "row kernel launch");
}

template <typename Sink>
static int dispatch_rows(int preds_dtype, int preds_has_class_dim, const RowArgs& a, Sink sink, size_t smem,
                         cudaStream_t st) {
    if (!preds_has_class_dim) {
        if constexpr (Sink::kNeedsTarget) {
            MB200_REQUIRE(preds_dtype >= MB200_I64 && preds_dtype <= MB200_BOOL,
                          "label-format preds must have an integer dtype (got dtype tag %d)", preds_dtype);
            const int grid = grid_for(a.n_outer * a.inner, kRowThreads, sm_count() * 8);
            labels_kernel<Sink><<<grid, kRowThreads, smem, st>>>(a, preds_dtype, sink);
            count_launch();
            return check_cuda(cudaGetLastError(), "labels kernel launch");
        } else {
            set_error("argmax needs a class dimension");
            return MB200_ERR_INVALID;
        }
    }
    switch (preds_dtype) {
        case MB200_BF16: return launch_rows<__nv_bfloat16, Sink>(a, sink, smem, st);
        case MB200_F16: return launch_rows<__half, Sink>(a, sink, smem, st);
        case MB200_F32: return launch_rows<float, Sink>(a, sink, smem, st);
        case MB200_F64: return launch_rows<double, Sink>(a, sink, smem, st);
        default:
            set_error("preds with a class dimension must be floating point (got dtype tag %d)", preds_dtype);
            return MB200_ERR_INVALID;
    }
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
    const bool priv = !micro && num_classes <= 2048 && n_outer * inner >= 4096;
    if (priv) {
        StatsSink<true> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn,
                          (long long*)workspace, (int)num_classes, micro};
        return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, (size_t)(3 * num_classes) * sizeof(unsigned),
                             st);
    }
    StatsSink<false> s{(long long*)tp, (long long*)fp, (long long*)tn, (long long*)fn, (long long*)workspace,
                       (int)num_classes, micro};
    return dispatch_rows(preds_dtype, preds_has_class_dim, a, s, 0, st);
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
