// aggregate_tma.cu -- stage 1 fast path: TMA-staged streaming head / step reduction.
//
// Same arithmetic as aggregate.cu (MaskHookLogger / BatchMaskHookLogger reducers,
// "Attention Guided Warping/attention_extraction/llava.py:94-132, 385-411"):
//     out[b,t] = mean_l mean_h  a[b,l,h,t] / (sum_t a[b,l,h,t] + 1e-12)
// for the benchmark layout, where the L*Hh rows of an image are contiguous in memory.
//
// HBM-bound: every attention element is read exactly once and nothing else moves.
//   * grid = (nsplit, B); a CTA owns a contiguous slab of rows of one image;
//   * one PRODUCER thread streams the slab through a ring of shared-memory stages with
//     cp.async.bulk (one TMA bulk copy of kRowsPerStage whole rows per stage, completion on a
//     `full` mbarrier, reuse gated by an `empty` mbarrier): no LSU instructions and no registers
//     carry bytes in flight, so the loads of the next stages overlap the arithmetic of this one;
//   * eight CONSUMER warps take one row of a stage each: the row is read from shared memory with
//     conflict-free 32-bit loads (lane owns words lane, lane+32, ...: the same token columns for
//     every row), summed with a warp-shuffle tree, and the normalised row is accumulated in
//     per-lane registers;
//   * warps are combined through shared memory in warp order, CTAs of an image through the
//     [B][nsplit][T] fp32 partial buffer summed in split order by the consumer kernel ->
//     bitwise deterministic.
#include "common.cuh"

namespace aw {
namespace {

constexpr int kConsumerWarps = 8;
constexpr int kRowsPerStage = kConsumerWarps;
constexpr int kAggStages = 4;
constexpr int kThreads = (kConsumerWarps + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src_gmem), "r"(bytes), "r"(bar)
        : "memory");
}

// Packed fp32x2 arithmetic (Blackwell FADD2 / FFMA2: two lanes per issue slot).
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n"
        " mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// A lane owns NP "pairs" of token columns per row.  16-bit rows: pair k is the 32-bit word
// lane + 32 k (columns 2 w, 2 w + 1); fp32 rows: pair k is the words lane + 64 k and lane + 64 k + 32.
// Either way a warp-wide load touches 32 consecutive words: conflict-free.
template <typename T>
struct PairTraits;
template <>
struct PairTraits<__nv_bfloat16> {
    static constexpr int kWordsPerPair = 1;
    __device__ static __forceinline__ float2 make(uint32_t w, uint32_t) {
        return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
};
template <>
struct PairTraits<__half> {
    static constexpr int kWordsPerPair = 1;
    __device__ static __forceinline__ float2 make(uint32_t w, uint32_t) {
        return __half22float2(*reinterpret_cast<const __half2*>(&w));
    }
};
template <>
struct PairTraits<float> {
    static constexpr int kWordsPerPair = 2;
    __device__ static __forceinline__ float2 make(uint32_t w0, uint32_t w1) {
        return make_float2(__uint_as_float(w0), __uint_as_float(w1));
    }
};

// NP = pairs per lane per row (compile time so the row lives in registers); EXACT: the row fills
// all NP pairs of all 32 lanes (no bounds checks in the loop).
template <typename T, int NP, bool EXACT>
__global__ void __launch_bounds__(kThreads, 4)
aggregate_rows_tma_kernel(const T* __restrict__ attn, int n_rows, int Tlen, int64_t sb, int rows_per_cta,
                          float eps, float* __restrict__ partial) {
    constexpr int WPP = PairTraits<T>::kWordsPerPair;
    extern __shared__ __align__(128) uint8_t smem[];
    const int row_bytes = Tlen * (int)sizeof(T);
    const int row_words = row_bytes >> 2;
    const int stage_bytes = kRowsPerStage * row_bytes;
    float* s_red = reinterpret_cast<float*>(smem + kAggStages * stage_bytes);       // [warps][Tlen]
    const uint32_t ring_s = smem_u32(smem);
    const uint32_t bars_s = smem_u32(s_red + kConsumerWarps * Tlen);                // full[], empty[]
    const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int r0 = split * rows_per_cta;
    const int r1 = min(r0 + rows_per_cta, n_rows);
    const int n_stages = (r1 - r0 + kRowsPerStage - 1) / kRowsPerStage;

    if (tid == 0) {
        for (int s = 0; s < kAggStages; ++s) {
            mbar_init(bars_s + 8u * s, 1);
            mbar_init(bars_s + 8u * (kAggStages + s), kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ---- producer: one thread streams the slab, kRowsPerStage rows per bulk copy -------------
        if (lane == 0) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(attn + (int64_t)b * sb) + (int64_t)r0 * row_bytes;
            for (int s = 0; s < n_stages; ++s) {
                const int st = s % kAggStages;
                mbar_wait(bars_s + 8u * (kAggStages + st), ((s / kAggStages) & 1) ^ 1);
                const int rows = min(kRowsPerStage, r1 - r0 - s * kRowsPerStage);
                const uint32_t bytes = (uint32_t)(rows * row_bytes);
                mbar_arrive_expect_tx(bars_s + 8u * st, bytes);
                bulk_g2s(ring_s + (uint32_t)(st * stage_bytes), src + (int64_t)s * stage_bytes, bytes,
                         bars_s + 8u * st);
            }
        }
        return;
    }

    // ---- consumers: warp w takes row w of every stage -------------------------------------------
    float2 acc[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int s = 0; s < n_stages; ++s) {
        const int st = s % kAggStages;
        mbar_wait(bars_s + 8u * st, (s / kAggStages) & 1);
        if (r0 + s * kRowsPerStage + warp < r1) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(smem + st * stage_bytes + warp * row_bytes);
            float2 f[NP];
            float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const int w0 = lane + 32 * WPP * k, w1 = w0 + 32;
                const uint32_t a0 = (EXACT || w0 < row_words) ? row[w0] : 0u;
                const uint32_t a1 = WPP == 2 ? ((EXACT || w1 < row_words) ? row[w1] : 0u) : 0u;
                f[k] = PairTraits<T>::make(a0, a1);
                sum2 = add2(sum2, f[k]);
            }
            const float sum = warp_sum(sum2.x + sum2.y);
            const float inv = 1.0f / (sum + eps);
            const float2 inv2 = make_float2(inv, inv);
#pragma unroll
            for (int k = 0; k < NP; ++k) acc[k] = fma2(f[k], inv2, acc[k]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars_s + 8u * (kAggStages + st));
    }

    // ---- combine the warps in warp order ---------------------------------------------------------
    float* mine = s_red + warp * Tlen;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const int w0 = lane + 32 * WPP * k, w1 = w0 + 32;
        if (WPP == 1) {
            if (EXACT || w0 < row_words) {
                mine[2 * w0] = acc[k].x;
                mine[2 * w0 + 1] = acc[k].y;
            }
        } else {
            if (EXACT || w0 < row_words) mine[w0] = acc[k].x;
            if (EXACT || w1 < row_words) mine[w1] = acc[k].y;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    float* out = partial + ((int64_t)b * nsplit + split) * Tlen;
    for (int t = tid; t < Tlen; t += kConsumerWarps * 32) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) v += s_red[w * Tlen + t];
        out[t] = v;
    }
}

template <typename T, int NP, bool EXACT>
int launch(const void* attn, int B, int n_rows, int Tlen, int64_t sb, int rows_per_cta, int nsplit, float eps,
           float* partial, cudaStream_t st) {
    const size_t smem = (size_t)kAggStages * kRowsPerStage * Tlen * sizeof(T) +
                        (size_t)kConsumerWarps * Tlen * sizeof(float) + 2 * kAggStages * sizeof(uint64_t);
    auto kern = aggregate_rows_tma_kernel<T, NP, EXACT>;
    // the opt-in is a per-device function attribute: the cache is keyed on (device, bytes)
    struct OptIn { int dev; size_t smem; };
    static thread_local OptIn set_for = {-1, 0};
    if (smem > 48 * 1024) {
        int dev = 0;
        AW_CUDA(cudaGetDevice(&dev));
        if (set_for.dev != dev || set_for.smem != smem) {
            AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            set_for = OptIn{dev, smem};
        }
    }
    kern<<<dim3(nsplit, B), kThreads, smem, st>>>(static_cast<const T*>(attn), n_rows, Tlen, sb, rows_per_cta,
                                                   eps, partial);
    return check_launch("aggregate_rows_tma_kernel");
}

template <typename T>
int dispatch(const void* attn, int B, int n_rows, int Tlen, int64_t sb, int rows_per_cta, int nsplit, float eps,
             float* partial, cudaStream_t st) {
    const int words_per_lane = 32 * PairTraits<T>::kWordsPerPair;      // row words one pair index covers
    const int row_words = Tlen * (int)sizeof(T) / 4;
    const int need = (row_words + words_per_lane - 1) / words_per_lane;
    if (need == 9 && row_words == 9 * words_per_lane)                   // LLaVA-1.5: 576 image tokens
        return launch<T, 9, true>(attn, B, n_rows, Tlen, sb, rows_per_cta, nsplit, eps, partial, st);
#define AW_CASE(NP) \
    if (need <= NP) return launch<T, NP, false>(attn, B, n_rows, Tlen, sb, rows_per_cta, nsplit, eps, partial, st);
    AW_CASE(2)
    AW_CASE(5)
    AW_CASE(9)
#undef AW_CASE
    return ATTWARP_ERR_UNSUPPORTED;
}

}  // namespace

// True when the TMA path applies: rows of an image contiguous, every row a whole number of
// 16-byte units at a 16-byte aligned address, no per-sample token offset, row fits the lanes.
bool aggregate_tma_applicable(const void* attn, int dtype, int L, int Hh, int T, int64_t sb, int64_t sl,
                              int64_t sh, const int32_t* tok_start) {
    const int es = dtype == ATTWARP_F32 ? 4 : 2;
    if (tok_start != nullptr || sh != T || (L > 1 && sl != (int64_t)Hh * T)) return false;
    if ((reinterpret_cast<uintptr_t>(attn) & 15) != 0 || ((int64_t)T * es) % 16 != 0 || (sb * es) % 16 != 0) return false;
    if ((T + 63) / 64 > 9) return false;                  // <= 9 fp32x2 accumulators per lane
    const size_t smem = (size_t)kAggStages * kRowsPerStage * T * es + (size_t)kConsumerWarps * T * 4 + 64;
    return smem <= 200 * 1024;
}

int launch_aggregate_tma(const void* attn, int dtype, int B, int L, int Hh, int T, int64_t sb, float eps,
                         float* partial, int nsplit, cudaStream_t st) {
    const int n_rows = L * Hh;
    const int rows_per_cta = (n_rows + nsplit - 1) / nsplit;
    switch (dtype) {
        case ATTWARP_BF16: return dispatch<__nv_bfloat16>(attn, B, n_rows, T, sb, rows_per_cta, nsplit, eps, partial, st);
        case ATTWARP_F16: return dispatch<__half>(attn, B, n_rows, T, sb, rows_per_cta, nsplit, eps, partial, st);
        case ATTWARP_F32: return dispatch<float>(attn, B, n_rows, T, sb, rows_per_cta, nsplit, eps, partial, st);
        default: return ATTWARP_ERR_INVALID_ARG;
    }
}

}  // namespace aw
