// remap.cu -- stage 5: bilinear resample through separable maps with cv2.remap semantics.
//
// Replaces cv2.remap(img, meshgrid(map_x, map_y), INTER_LINEAR, BORDER_REPLICATE) as called at
// "Attention Guided Warping/new_method.py:268-271" and
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198".
// The per-pixel arithmetic is in warp_math.h (quantise_coord / bilinear_u8 / bilinear_f32).
//
// Kernels
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
    if (dtype == ATTWARP_F32)
        return launch_direct<float>(src, dst, layout, B, C, H, W, Ho, Wo, map_x, map_y, st);
    return fail(ATTWARP_ERR_INVALID_ARG, "remap: image dtype must be u8/f32 (got %d)", dtype);
}

}  // namespace aw
