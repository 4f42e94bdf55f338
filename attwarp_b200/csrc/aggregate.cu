// aggregate.cu -- stage 1: head / step reduction of MLLM attention into a token map.
//
// Replaces MaskHookLogger._process_attention + finalize and the BatchMaskHookLogger pair
// ("Attention Guided Warping/attention_extraction/llava.py:94-132, 385-411"):
//     out[b,t] = mean_l mean_h  a[b,l,h,t] / (sum_t a[b,l,h,t] + 1e-12)
//
// HBM-bound streaming reduction: every attention element is read exactly once.
//   * grid = (nsplit, B): each CTA owns a contiguous slice of the L*Hh rows of one image;
//   * one warp per row: the row (T elements) is pulled with 128-bit streaming loads
//     (ld.global.nc.L1::no_allocate) and kept in registers, the row sum is a warp-shuffle tree,
//     the normalised row is added into per-lane register accumulators (a lane always owns the
//     same token columns), two rows in flight per warp for memory-level parallelism;
//   * warps of a CTA are combined through shared memory in warp order, CTAs of an image through
//     a [B][nsplit][T] fp32 partial buffer that the consumer (finalize kernel or the fused
//     maps-from-tokens kernel) sums in split order -> bitwise deterministic.
#include <stdlib.h>

#include "common.cuh"

namespace aw {
namespace {

constexpr int kAggThreads = 256;
constexpr int kAggWarps = kAggThreads / 32;

template <typename T>
struct VecTraits;
template <>
struct VecTraits<__nv_bfloat16> {
    static constexpr int kElems = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ float scalar(const __nv_bfloat16* p) {
        return __bfloat162float(*p);
    }
};
template <>
struct VecTraits<__half> {
    static constexpr int kElems = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    __device__ static __forceinline__ float scalar(const __half* p) { return __half2float(*p); }
};
template <>
struct VecTraits<float> {
    static constexpr int kElems = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
        f[0] = __uint_as_float(v.x);
        f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z);
        f[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ float scalar(const float* p) { return *p; }
};

// ---- fast path: rows are 16-byte aligned and T is a multiple of the vector width -------------
// NV = 128-bit vectors per lane per row (compile time so the row lives in registers).
template <typename T, int NV>
__global__ void __launch_bounds__(kAggThreads)
aggregate_rows_vec_kernel(const T* __restrict__ attn, int n_rows, int Hh, int Tlen, int64_t sb,
                          int64_t sl, int64_t sh, int rows_per_cta, float eps,
                          float* __restrict__ partial) {
    constexpr int VE = VecTraits<T>::kElems;
    extern __shared__ float s_red[];  // [kAggWarps][Tlen]
    const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nvec = Tlen / VE;
    const int row_beg = split * rows_per_cta;
    const int row_end = min(row_beg + rows_per_cta, n_rows);
    const T* base = attn + (int64_t)b * sb;

    float acc[NV * VE];
#pragma unroll
    for (int i = 0; i < NV * VE; ++i) acc[i] = 0.f;

    auto row_ptr = [&](int r) -> const T* {
        const int l = r / Hh, h = r - l * Hh;
        return base + (int64_t)l * sl + (int64_t)h * sh;
    };
    auto load_row = [&](const T* p, uint4* v) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int vi = lane + 32 * k;
            v[k] = vi < nvec ? ldg_stream_v4(reinterpret_cast<const uint4*>(p) + vi)
                             : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    auto consume_row = [&](const uint4* v) {
        float f[NV * VE];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            VecTraits<T>::unpack(v[k], f + k * VE);
#pragma unroll
            for (int e = 0; e < VE; ++e) s += f[k * VE + e];
        }
        s = warp_sum(s);
        const float inv = 1.0f / (s + eps);
#pragma unroll
        for (int i = 0; i < NV * VE; ++i) acc[i] = fmaf(f[i], inv, acc[i]);
    };

    // two rows in flight per warp
    int r = row_beg + wid;
    for (; r + kAggWarps < row_end; r += 2 * kAggWarps) {
        uint4 v0[NV], v1[NV];
        load_row(row_ptr(r), v0);
        load_row(row_ptr(r + kAggWarps), v1);
        consume_row(v0);
        consume_row(v1);
    }
    if (r < row_end) {
        uint4 v0[NV];
        load_row(row_ptr(r), v0);
        consume_row(v0);
    }

    // combine the warps in warp order
    float* mine = s_red + wid * Tlen;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int vi = lane + 32 * k;
        if (vi < nvec) {
#pragma unroll
            for (int e = 0; e < VE; ++e) mine[vi * VE + e] = acc[k * VE + e];
        }
    }
    __syncthreads();
    float* out = partial + ((int64_t)b * nsplit + split) * Tlen;
    for (int t = threadIdx.x; t < Tlen; t += kAggThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAggWarps; ++w) s += s_red[w * Tlen + t];
        out[t] = s;
    }
}

// ---- generic path: any alignment / per-sample token offsets (the live hook layout) -----------
template <typename T>
__global__ void __launch_bounds__(kAggThreads)
aggregate_rows_generic_kernel(const T* __restrict__ attn, int n_rows, int Hh, int Tlen,
                              int64_t sb, int64_t sl, int64_t sh,
                              const int32_t* __restrict__ tok_start, int rows_per_cta, float eps,
                              float* __restrict__ partial) {
    extern __shared__ float s_red[];  // [kAggWarps][Tlen]
    const int b = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row_beg = split * rows_per_cta;
    const int row_end = min(row_beg + rows_per_cta, n_rows);
    const T* base = attn + (int64_t)b * sb + (tok_start ? (int64_t)tok_start[b] : 0);
    float* mine = s_red + wid * Tlen;
    for (int t = lane; t < Tlen; t += 32) mine[t] = 0.f;
    for (int r = row_beg + wid; r < row_end; r += kAggWarps) {
        const int l = r / Hh, h = r - l * Hh;
        const T* p = base + (int64_t)l * sl + (int64_t)h * sh;
        float s = 0.f;
        for (int t = lane; t < Tlen; t += 32) s += VecTraits<T>::scalar(p + t);
        s = warp_sum(s);
        const float inv = 1.0f / (s + eps);
        for (int t = lane; t < Tlen; t += 32)
            mine[t] = fmaf(VecTraits<T>::scalar(p + t), inv, mine[t]);
    }
    __syncthreads();
    float* out = partial + ((int64_t)b * nsplit + split) * Tlen;
    for (int t = threadIdx.x; t < Tlen; t += kAggThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAggWarps; ++w) s += s_red[w * Tlen + t];
        out[t] = s;
    }
}

__global__ void aggregate_finalize_kernel(const float* __restrict__ partial, int nsplit, int Tlen,
                                          float scale, float* __restrict__ out, int accumulate,
                                          float out_scale) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Tlen) return;
    const float* p = partial + (int64_t)b * nsplit * Tlen + t;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += p[(int64_t)k * Tlen];
    const float mean = s * scale;
    float* o = out + (int64_t)b * Tlen + t;
    *o = accumulate ? fmaf(out_scale, mean, *o) : out_scale * mean;
}

template <typename T, int NV>
int launch_vec(const void* attn, int B, int n_rows, int Hh, int Tlen, int64_t sb, int64_t sl,
               int64_t sh, int rows_per_cta, int nsplit, float eps, float* partial,
               cudaStream_t st) {
    const size_t smem = (size_t)kAggWarps * Tlen * sizeof(float);
    auto kern = aggregate_rows_vec_kernel<T, NV>;
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(nsplit, B), kAggThreads, smem, st>>>(static_cast<const T*>(attn), n_rows, Hh, Tlen,
                                                     sb, sl, sh, rows_per_cta, eps, partial);
    return check_launch("aggregate_rows_vec_kernel");
}

template <typename T>
int dispatch(const void* attn, int B, int L, int Hh, int Tlen, int64_t sb, int64_t sl, int64_t sh,
             const int32_t* tok_start, float eps, float* partial, int nsplit, cudaStream_t st) {
    constexpr int VE = VecTraits<T>::kElems;
    const int n_rows = L * Hh;
    const int rows_per_cta = (n_rows + nsplit - 1) / nsplit;
    const bool aligned = tok_start == nullptr && (reinterpret_cast<uintptr_t>(attn) % 16 == 0) &&
                         Tlen % VE == 0 && sb % VE == 0 && sl % VE == 0 && sh % VE == 0;
    const int need = (Tlen / VE + 31) / 32;  // vectors per lane
    if (aligned && need <= 9) {
#define AW_AGG_CASE(NV)                                                                        \
    if (need <= NV)                                                                            \
        return launch_vec<T, NV>(attn, B, n_rows, Hh, Tlen, sb, sl, sh, rows_per_cta, nsplit,  \
                                 eps, partial, st);
        AW_AGG_CASE(1)
        AW_AGG_CASE(2)
        AW_AGG_CASE(3)
        AW_AGG_CASE(5)
        AW_AGG_CASE(9)
#undef AW_AGG_CASE
    }
    const size_t smem = (size_t)kAggWarps * Tlen * sizeof(float);
    if (smem > 200 * 1024)
        return fail(ATTWARP_ERR_UNSUPPORTED, "aggregate: T=%d exceeds the supported row length", Tlen);
    auto kern = aggregate_rows_generic_kernel<T>;
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(nsplit, B), kAggThreads, smem, st>>>(static_cast<const T*>(attn), n_rows, Hh, Tlen,
                                                     sb, sl, sh, tok_start, rows_per_cta, eps,
                                                     partial);
    return check_launch("aggregate_rows_generic_kernel");
}

}  // namespace

bool aggregate_tma_applicable(const void* attn, int dtype, int L, int Hh, int T, int64_t sb, int64_t sl,
                              int64_t sh, const int32_t* tok_start);
int launch_aggregate_tma(const void* attn, int dtype, int B, int L, int Hh, int T, int64_t sb, float eps,
                         float* partial, int nsplit, cudaStream_t st);

// Number of CTAs per image.  The kernel is a pure stream, so the best launch is ONE resident
// wave of equal, long-lived CTAs (pipeline fill and the per-CTA epilogue are paid once per CTA,
// and equal slabs finish together whatever SM they landed on): as many CTAs per image as still
// fit the resident set (4 per SM), never fewer than 2 rows per warp.
int aggregate_nsplit(int B, int L, int Hh) {
    static const int forced = [] {
        const char* e = getenv("ATTWARP_AGG_NSPLIT");
        return e ? atoi(e) : 0;
    }();
    const int n_rows = L * Hh;
    const int max_split = (n_rows + 2 * kAggWarps - 1) / (2 * kAggWarps);
    int nsplit = forced > 0 ? forced : (4 * sm_count() / sm_share()) / (B > 0 ? B : 1);
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    return nsplit;
}

int launch_aggregate_partial(const void* attn, int dtype, int B, int L, int Hh, int T, int64_t sb,
                             int64_t sl, int64_t sh, const int32_t* tok_start, float eps,
                             float* partial, int nsplit, cudaStream_t st) {
    if (aggregate_tma_applicable(attn, dtype, L, Hh, T, sb, sl, sh, tok_start)) {
        const int rc = launch_aggregate_tma(attn, dtype, B, L, Hh, T, sb, eps, partial, nsplit, st);
        if (rc != ATTWARP_ERR_UNSUPPORTED) return rc;
    }
    switch (dtype) {
        case ATTWARP_BF16:
            return dispatch<__nv_bfloat16>(attn, B, L, Hh, T, sb, sl, sh, tok_start, eps, partial, nsplit, st);
        case ATTWARP_F16:
            return dispatch<__half>(attn, B, L, Hh, T, sb, sl, sh, tok_start, eps, partial, nsplit, st);
        case ATTWARP_F32:
            return dispatch<float>(attn, B, L, Hh, T, sb, sl, sh, tok_start, eps, partial, nsplit, st);
        default:
            return fail(ATTWARP_ERR_INVALID_ARG, "aggregate: attention dtype must be bf16/f16/f32 (got %d)", dtype);
    }
}

int launch_aggregate_finalize(const float* partial, int B, int nsplit, int T, float scale,
                              float* out, int accumulate, float out_scale, cudaStream_t st) {
    aggregate_finalize_kernel<<<dim3((T + 127) / 128, B), 128, 0, st>>>(partial, nsplit, T, scale,
                                                                        out, accumulate, out_scale);
    return check_launch("aggregate_finalize_kernel");
}

}  // namespace aw
