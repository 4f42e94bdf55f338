// TEST-ONLY host harness: compiles attwarp_b200/csrc/warp_math.h with g++ so the scalar device
// arithmetic (coordinate quantisation, fixed-point bilinear, np.interp restatement, transforms)
// can be compared with the oracle on a machine without a GPU.  Never linked into the product.
#include <stdint.h>

#include "../../attwarp_b200/csrc/warp_math.h"

extern "C" {

void hc_quantise(const float* m, int n, int* s) {
    for (int i = 0; i < n; ++i) s[i] = aw::quantise_coord(m[i]);
}

// separable remap, HWC layout, uint8
void hc_remap_u8(const uint8_t* src, uint8_t* dst, int C, int H, int W, int Ho, int Wo,
                 const float* mx, const float* my) {
    for (int y = 0; y < Ho; ++y) {
        const int sy = aw::quantise_coord(my[y]), ay = sy & 31;
        const int y0 = aw::clampi(sy >> 5, 0, H - 1), y1 = aw::clampi((sy >> 5) + 1, 0, H - 1);
        for (int x = 0; x < Wo; ++x) {
            const int sx = aw::quantise_coord(mx[x]), ax = sx & 31;
            const int x0 = aw::clampi(sx >> 5, 0, W - 1), x1 = aw::clampi((sx >> 5) + 1, 0, W - 1);
            for (int c = 0; c < C; ++c)
                dst[((int64_t)y * Wo + x) * C + c] = aw::bilinear_u8(
                    src[((int64_t)y0 * W + x0) * C + c], src[((int64_t)y0 * W + x1) * C + c],
                    src[((int64_t)y1 * W + x0) * C + c], src[((int64_t)y1 * W + x1) * C + c], ax, ay);
        }
    }
}

void hc_remap_f32(const float* src, float* dst, int C, int H, int W, int Ho, int Wo,
                  const float* mx, const float* my) {
    for (int y = 0; y < Ho; ++y) {
        const int sy = aw::quantise_coord(my[y]), ay = sy & 31;
        const int y0 = aw::clampi(sy >> 5, 0, H - 1), y1 = aw::clampi((sy >> 5) + 1, 0, H - 1);
        for (int x = 0; x < Wo; ++x) {
            const int sx = aw::quantise_coord(mx[x]), ax = sx & 31;
            const int x0 = aw::clampi(sx >> 5, 0, W - 1), x1 = aw::clampi((sx >> 5) + 1, 0, W - 1);
            const aw::BilinearWeightsF32 w = aw::bilinear_weights_f32(ax, ay);
            for (int c = 0; c < C; ++c)
                dst[((int64_t)y * Wo + x) * C + c] = aw::bilinear_f32(
                    src[((int64_t)y0 * W + x0) * C + c], src[((int64_t)y0 * W + x1) * C + c],
                    src[((int64_t)y1 * W + x0) * C + c], src[((int64_t)y1 * W + x1) * C + c], w);
        }
    }
}

void hc_interp(const double* xp, int n, int n_out, double* out) {
    for (int j = 0; j < n_out; ++j) out[j] = aw::interp_index((double)j, xp, n);
}

void hc_transform(const double* x, int n, int t, double scale, double divisor, int inverse, double* out) {
    for (int i = 0; i < n; ++i)
        out[i] = inverse ? aw::transform_inv(x[i], t, scale, divisor) : aw::transform_fwd(x[i], t, scale, divisor);
}
}
