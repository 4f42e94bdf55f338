// remap.cu -- stage 5: bilinear resample through separable maps with cv2.remap semantics.
//
// Replaces cv2.remap(img, meshgrid(map_x, map_y), INTER_LINEAR, BORDER_REPLICATE) as called at
// "Attention Guided Warping/new_method.py:268-271" and
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198".
// The per-pixel arithmetic is in warp_math.h (quantise_coord / bilinear_u8 / bilinear_f32).
//
// Kernels
//   remap_f32_rows_kernel float32 images, any layout / channel count: one thread per output column,
//                         walking a tile of rows with the tapped source rows carried in registers.
//   remap_direct_kernel   any dtype / layout / channel count; one thread per output pixel,
//                         taps gathered straight from global memory through L1/L2.  Baseline and
//                         fallback for shapes the streaming kernel (remap_stream.cu) does not cover.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace aw {
namespace {

constexpr int kDirectThreads = 256;

// One thread per output pixel (x fastest).  Every output row uses two source rows and the map is
// monotone, so neighbouring threads touch neighbouring (or identical) source pixels: the gathers
// coalesce into a few sectors per warp and hit L1 for the second tap.
template <typename T, bool HWC>
__global__ void __launch_bounds__(kDirectThreads)
remap_direct_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int H, int W, int Ho,
                    int Wo, const float* __restrict__ map_x, const float* __restrict__ map_y) {
    const int x = blockIdx.x * kDirectThreads + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= Wo) return;
    const int sx = quantise_coord(__ldg(map_x + (int64_t)b * Wo + x));
    const int sy = quantise_coord(__ldg(map_y + (int64_t)b * Ho + y));
    const int ax = sx & 31, ay = sy & 31;
    const int x0 = clampi(sx >> 5, 0, W - 1), x1 = clampi((sx >> 5) + 1, 0, W - 1);
    const int y0 = clampi(sy >> 5, 0, H - 1), y1 = clampi((sy >> 5) + 1, 0, H - 1);
    const T* img = src + (int64_t)b * C * H * W;
    T* out = dst + (int64_t)b * C * Ho * Wo;
    BilinearWeightsF32 wf;
    if (sizeof(T) == 4) wf = bilinear_weights_f32(ax, ay);
    for (int c = 0; c < C; ++c) {
        int64_t o00, o01, o10, o11, od;
        if (HWC) {
            o00 = ((int64_t)y0 * W + x0) * C + c;
            o01 = ((int64_t)y0 * W + x1) * C + c;
            o10 = ((int64_t)y1 * W + x0) * C + c;
            o11 = ((int64_t)y1 * W + x1) * C + c;
            od = ((int64_t)y * Wo + x) * C + c;
        } else {
            const int64_t plane = (int64_t)c * H * W;
            o00 = plane + (int64_t)y0 * W + x0;
            o01 = plane + (int64_t)y0 * W + x1;
            o10 = plane + (int64_t)y1 * W + x0;
            o11 = plane + (int64_t)y1 * W + x1;
            od = (int64_t)c * Ho * Wo + (int64_t)y * Wo + x;
        }
        if constexpr (sizeof(T) == 1) {
            out[od] = bilinear_u8(__ldg(img + o00), __ldg(img + o01), __ldg(img + o10),
                                  __ldg(img + o11), ax, ay);
        } else {
            out[od] = bilinear_f32(__ldg(img + o00), __ldg(img + o01), __ldg(img + o10),
                                   __ldg(img + o11), wf);
        }
    }
}

// ---- float32 images: one thread per output column (x channel), walking a tile of output rows --------
// Each thread fixes its two horizontal taps and weights once and then walks kRowsF32 output rows of
// its plane: the separable map makes every thread of the CTA tap the same two source rows, so the row
// addresses and the vertical weights come from a small shared table, and a source row that the next
// output row taps again (as its upper or lower row) stays in registers -- two loads per NEW source row
// and thread instead of four per output.  Loads of neighbouring threads are neighbouring floats
// (coalesced through L1), stores are fully coalesced.  Arithmetic is OpenCV's float path exactly:
// w = fl32(wy * wx), ((p00 w00 + p01 w01) + p10 w10) + p11 w11 without FMA contraction.
#ifndef AW_F32_ROWS
#define AW_F32_ROWS 32
#endif
#ifndef AW_F32_COLS
#define AW_F32_COLS 2
#endif
constexpr int kThreadsF32 = 256;
constexpr int kRowsF32 = AW_F32_ROWS;
constexpr int kColsF32 = AW_F32_COLS;      // output columns per thread (independent chains)

// ELEMS: floats per pixel that are interleaved in memory (HWC: C, planar: 1)
template <bool HWC>
__global__ void __launch_bounds__(kThreadsF32)
remap_f32_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int Ho,
                      int Wo, int map_div, const float* __restrict__ map_x, const float* __restrict__ map_y) {
    __shared__ int4 rowtab[kRowsF32];      // {ya, yb, bits(1 - fy), bits(fy)}
    const int E = HWC ? C : 1;             // interleaved floats per pixel
    const int plane = blockIdx.z;
    const int mrow = plane / map_div;
    const int y0 = blockIdx.y * kRowsF32;
    const int nrows = min(kRowsF32, Ho - y0);
    if (threadIdx.x < nrows) {
        const int sy = quantise_coord(__ldg(map_y + (int64_t)mrow * Ho + y0 + threadIdx.x));
        const int ay = sy & 31;
        const float fy = fmul_nofma((float)ay, 1.0f / 32.0f);
        rowtab[threadIdx.x] = make_int4(clampi(sy >> 5, 0, H - 1), clampi((sy >> 5) + 1, 0, H - 1),
                                        __float_as_int(fadd_nofma(1.0f, -fy)), __float_as_int(fy));
    }
    const int64_t src_pitch = (int64_t)W * E, dst_pitch = (int64_t)Wo * E;
    const float* sp = src + (int64_t)plane * H * src_pitch;
    float* dp = dst + (int64_t)plane * Ho * dst_pitch + (int64_t)y0 * dst_pitch;
    const int n_e = Wo * E;                // floats per output row
    int s0[kColsF32], s1[kColsF32], e_out[kColsF32];
    float wx0[kColsF32], wx1[kColsF32];
    bool valid[kColsF32];
#pragma unroll
    for (int j = 0; j < kColsF32; ++j) {
        const int e = (blockIdx.x * kColsF32 + j) * kThreadsF32 + threadIdx.x;
        valid[j] = e < n_e;
        e_out[j] = e;
        const int x = valid[j] ? e / E : 0, c = valid[j] ? e - x * E : 0;
        const int sx = quantise_coord(__ldg(map_x + (int64_t)mrow * Wo + x));
        const float fx = fmul_nofma((float)(sx & 31), 1.0f / 32.0f);
        wx0[j] = fadd_nofma(1.0f, -fx);
        wx1[j] = fx;
        s0[j] = clampi(sx >> 5, 0, W - 1) * E + c;
        s1[j] = clampi((sx >> 5) + 1, 0, W - 1) * E + c;
    }
    __syncthreads();
    // Source rows held in registers: pa (upper tap), pb (lower tap) and pc, the lower row of the NEXT output
    // row, requested one iteration ahead so that its latency overlaps this row's arithmetic (the taps of a row
    // are otherwise consumed right after they are requested and the kernel is bound by bytes in flight).
    // Offsets are 32-bit (the launcher checks that a plane has fewer than 2^31 elements).
    int cur_a = -1, cur_b = -1, cur_c = -1;
    float pa0[kColsF32], pa1[kColsF32], pb0[kColsF32], pb1[kColsF32], pc0[kColsF32], pc1[kColsF32];
    float* op[kColsF32];
#pragma unroll
    for (int j = 0; j < kColsF32; ++j) {
        pa0[j] = pa1[j] = pb0[j] = pb1[j] = pc0[j] = pc1[j] = 0.0f;
        op[j] = dp + e_out[j];
        if (!valid[j]) { s0[j] = 0; s1[j] = 0; }         // loads stay in range, only the store is predicated
    }
    const int pitch32 = (int)src_pitch;
    int4 rt = rowtab[0];
    for (int r = 0; r < nrows; ++r) {
        const int4 nx = rowtab[min(r + 1, nrows - 1)];
        if (rt.x != cur_a) {
            if (rt.x == cur_b) {
#pragma unroll
                for (int j = 0; j < kColsF32; ++j) { pa0[j] = pb0[j]; pa1[j] = pb1[j]; }
            } else {
                const int ro = rt.x * pitch32;
#pragma unroll
                for (int j = 0; j < kColsF32; ++j) { pa0[j] = __ldg(sp + (ro + s0[j])); pa1[j] = __ldg(sp + (ro + s1[j])); }
            }
            cur_a = rt.x;
        }
        if (rt.y != cur_b) {
            if (rt.y == cur_c) {
#pragma unroll
                for (int j = 0; j < kColsF32; ++j) { pb0[j] = pc0[j]; pb1[j] = pc1[j]; }
            } else {
                const int ro = rt.y * pitch32;
#pragma unroll
                for (int j = 0; j < kColsF32; ++j) { pb0[j] = __ldg(sp + (ro + s0[j])); pb1[j] = __ldg(sp + (ro + s1[j])); }
            }
            cur_b = rt.y;
        }
        if (nx.y != cur_b && nx.y != cur_c) {                // request the next row's lower tap now
            const int ro = nx.y * pitch32;
#pragma unroll
            for (int j = 0; j < kColsF32; ++j) { pc0[j] = __ldg(sp + (ro + s0[j])); pc1[j] = __ldg(sp + (ro + s1[j])); }
            cur_c = nx.y;
        }
        const float wy0 = __int_as_float(rt.z), wy1 = __int_as_float(rt.w);
#pragma unroll
        for (int j = 0; j < kColsF32; ++j) {
            BilinearWeightsF32 w;
            w.w00 = fmul_nofma(wy0, wx0[j]);
            w.w01 = fmul_nofma(wy0, wx1[j]);
            w.w10 = fmul_nofma(wy1, wx0[j]);
            w.w11 = fmul_nofma(wy1, wx1[j]);
            const float v = bilinear_f32(pa0[j], pa1[j], pb0[j], pb1[j], w);
            if (valid[j]) *op[j] = v;
            op[j] += dst_pitch;
        }
        rt = nx;
    }
}

int launch_f32_rows(const float* src, float* dst, int layout, int B, int C, int H, int W, int Ho, int Wo,
                    const float* map_x, const float* map_y, cudaStream_t st) {
    const bool hwc = layout == ATTWARP_LAYOUT_HWC;
    const int planes = hwc ? B : B * C;
    const int n_e = hwc ? Wo * C : Wo;
    const dim3 grid((n_e + kThreadsF32 * kColsF32 - 1) / (kThreadsF32 * kColsF32), (Ho + kRowsF32 - 1) / kRowsF32, planes);
    if (hwc)
        remap_f32_rows_kernel<true><<<grid, kThreadsF32, 0, st>>>(src, dst, C, H, W, Ho, Wo, 1, map_x, map_y);
    else
        remap_f32_rows_kernel<false><<<grid, kThreadsF32, 0, st>>>(src, dst, C, H, W, Ho, Wo, C, map_x, map_y);
    return check_launch("remap_f32_rows_kernel");
}

template <typename T>
int launch_direct(const void* src, void* dst, int layout, int B, int C, int H, int W, int Ho,
                  int Wo, const float* map_x, const float* map_y, cudaStream_t st) {
    const dim3 grid((Wo + kDirectThreads - 1) / kDirectThreads, Ho, B);
    if (layout == ATTWARP_LAYOUT_HWC)
        remap_direct_kernel<T, true><<<grid, kDirectThreads, 0, st>>>(
            static_cast<const T*>(src), static_cast<T*>(dst), C, H, W, Ho, Wo, map_x, map_y);
    else
        remap_direct_kernel<T, false><<<grid, kDirectThreads, 0, st>>>(
            static_cast<const T*>(src), static_cast<T*>(dst), C, H, W, Ho, Wo, map_x, map_y);
    return check_launch("remap_direct_kernel");
}

}  // namespace

int launch_remap_u8_stream(const void* src, void* dst, int n_img, int C, int H, int W, int Ho, int Wo,
                           const float* map_x, const float* map_y, int map_div, cudaStream_t st);

// ATTWARP_REMAP=direct forces the baseline gather kernel (A/B comparisons, debugging); default is the
// streaming kernel (remap_stream.cu).
static bool force_direct() {
    static const bool v = [] {
        const char* e = getenv("ATTWARP_REMAP");
        return e != nullptr && strcmp(e, "direct") == 0;
    }();
    return v;
}
static int launch_u8_fast(const void* src, void* dst, int n_img, int C, int H, int W, int Ho, int Wo,
                          const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    // 3-channel interleaved images (the benchmark format): four adjacent pixels per thread (remap_quad.cu);
    // grey / 4-channel / planar images: remap_stream.cu.  (A column-walking kernel for those formats -- four output
    // bytes per thread, taps read ahead with cp.async -- was measured in round 2 and dropped: correct, but no faster
    // than remap_stream.cu except for large 4-channel images; profiles/r06_walk_kernel.md.)
    if (C == 3 && map_div == 1 && remap_quad_enabled())
        return launch_remap_u8_quad(src, dst, n_img, H, W, Ho, Wo, map_x, map_y, st);
    return launch_remap_u8_stream(src, dst, n_img, C, H, W, Ho, Wo, map_x, map_y, map_div, st);
}

int launch_remap(const void* src, void* dst, int dtype, int layout, int B, int C, int H, int W,
                 int Ho, int Wo, const float* map_x, const float* map_y, cudaStream_t st) {
    if (Ho > 65535 || B > 65535)
        return fail(ATTWARP_ERR_UNSUPPORTED, "remap: Ho=%d / B=%d exceed the grid limits", Ho, B);
    if (dtype == ATTWARP_U8 && !force_direct() && H >= 2 && W >= 2) {
        // HWC with 1/3/4 interleaved channels, or planar = B*C single-channel images
        if (layout == ATTWARP_LAYOUT_HWC && (C == 1 || C == 3 || C == 4))
            return launch_u8_fast(src, dst, B, C, H, W, Ho, Wo, map_x, map_y, 1, st);
        if (layout == ATTWARP_LAYOUT_CHW && (int64_t)B * C <= 65535)
            return launch_u8_fast(src, dst, B * C, 1, H, W, Ho, Wo, map_x, map_y, C, st);
    }
    if (dtype == ATTWARP_U8)
        return launch_direct<uint8_t>(src, dst, layout, B, C, H, W, Ho, Wo, map_x, map_y, st);
    if (dtype == ATTWARP_F32) {
        const int64_t planes = layout == ATTWARP_LAYOUT_HWC ? B : (int64_t)B * C;
        if (!force_direct() && remap_f32_stream_enabled() && planes <= 0x7fffffff) {
            const bool hwc = layout == ATTWARP_LAYOUT_HWC;
            const int rc = launch_remap_f32_stream(static_cast<const float*>(src), static_cast<float*>(dst), (int)planes,
                                                   hwc ? C : 1, H, W, Ho, Wo, map_x, map_y, hwc ? 1 : C, st);
            if (rc != ATTWARP_ERR_UNSUPPORTED) return rc;
        }
        if (!force_direct() && planes <= 65535 && (int64_t)Wo * C < 0x7fffffff && (int64_t)H * W * C < 0x7fffffff)
            return launch_f32_rows(static_cast<const float*>(src), static_cast<float*>(dst), layout, B, C, H, W,
                                   Ho, Wo, map_x, map_y, st);
        return launch_direct<float>(src, dst, layout, B, C, H, W, Ho, Wo, map_x, map_y, st);
    }
    return fail(ATTWARP_ERR_INVALID_ARG, "remap: image dtype must be u8/f32 (got %d)", dtype);
}

}  // namespace aw
