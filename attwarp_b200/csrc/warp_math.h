// warp_math.h -- scalar arithmetic shared by every kernel of the warp path.
//
// Everything here is __host__ __device__ so the exact same expressions can be compiled by g++
// into the test-only host harness (tests/hostcheck) and compared with the oracle on a machine
// without a GPU.  The product never runs these on the host.
//
// Reference semantics restated here (file:line under the reference root):
//   * cv2.remap(INTER_LINEAR, BORDER_REPLICATE) called at
//     "Attention Guided Warping/new_method.py:268-271" and
//     "model/marginalnet_full_dataset/checkpoint_utils.py:195-198"
//   * np.interp called at new_method.py:260-261 and checkpoint_utils.py:188-189
//   * the transform registry new_method.py:133-188
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AW_HD __host__ __device__ __forceinline__
#else
#define AW_HD inline
#endif

namespace aw {

constexpr double kBaseAttention = 1e-9;  // new_method.py:195
constexpr double kEpsilon = 1e-9;        // new_method.py:194

// ---- stage 5: coordinate quantisation -------------------------------------------------------
// OpenCV converts float maps to fixed point with s = cvRound(m * INTER_TAB_SIZE), INTER_TAB_SIZE
// = 32, round-half-to-even; integer pixel = s >> 5 (arithmetic), fraction = s & 31.
AW_HD int quantise_coord(float m) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(m * 32.0f);
#else
    return (int)lrintf(m * 32.0f);
#endif
}

AW_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// uint8 bilinear with OpenCV's 15-bit fixed-point weights.  The weight table entry for
// fraction (ax, ay) is exactly {(32-ax)(32-ay), ax(32-ay), (32-ax)ay, ax*ay} * 32 (sums to
// 2^15) and the result is (sum + 2^14) >> 15.  All products are exact integers, so the
// separable evaluation below is bit-identical: (32*v + 2^14) >> 15 == (v + 512) >> 10.
AW_HD int hblend_u8(int p0, int p1, int ax) { return (32 - ax) * p0 + ax * p1; }
AW_HD uint8_t vblend_u8(int h0, int h1, int ay) {
    return (uint8_t)(((32 - ay) * h0 + ay * h1 + 512) >> 10);
}
AW_HD uint8_t bilinear_u8(int p00, int p01, int p10, int p11, int ax, int ay) {
    return vblend_u8(hblend_u8(p00, p01, ax), hblend_u8(p10, p11, ax), ay);
}

// float32 bilinear: OpenCV's float table holds w = fl32(wy * wx) with wx in {1-fx, fx},
// fx = ax * (1/32); the pixel is ((p00*w00 + p01*w01) + p10*w10) + p11*w11 in float32 with no
// FMA contraction (verified bit-equal against cv2 4.13.0 by the oracle tests).
AW_HD float fmul_nofma(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
AW_HD float fadd_nofma(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
struct BilinearWeightsF32 {
    float w00, w01, w10, w11;
};
AW_HD BilinearWeightsF32 bilinear_weights_f32(int ax, int ay) {
    const float fx = fmul_nofma((float)ax, 1.0f / 32.0f);
    const float fy = fmul_nofma((float)ay, 1.0f / 32.0f);
    const float gx = fadd_nofma(1.0f, -fx);
    const float gy = fadd_nofma(1.0f, -fy);
    BilinearWeightsF32 w;
    w.w00 = fmul_nofma(gy, gx);
    w.w01 = fmul_nofma(gy, fx);
    w.w10 = fmul_nofma(fy, gx);
    w.w11 = fmul_nofma(fy, fx);
    return w;
}
AW_HD float bilinear_f32(float p00, float p01, float p10, float p11, const BilinearWeightsF32& w) {
    float acc = fadd_nofma(fmul_nofma(p00, w.w00), fmul_nofma(p01, w.w01));
    acc = fadd_nofma(acc, fmul_nofma(p10, w.w10));
    return fadd_nofma(acc, fmul_nofma(p11, w.w11));
}

// ---- stage 4: np.interp with fp = [0, 1, ..., n-1] ------------------------------------------
AW_HD double dmul_nofma(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
AW_HD double dadd_nofma(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
AW_HD double ddiv_exact(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// xp: n knots; returns the float64 interpolant at x (NumPy arr_interp semantics; bisection for
// "last index with xp[j] <= x").
AW_HD double interp_index(double x, const double* xp, int n) {
    if (x > xp[n - 1]) return (double)(n - 1);
    if (x < xp[0]) return 0.0;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (x >= xp[mid]) lo = mid + 1; else hi = mid;
    }
    const int j = lo - 1;
    if (j >= n - 1) return (double)(n - 1);
    const double xj = xp[j];
    if (xj == x) return (double)j;
    const double slope = ddiv_exact(1.0, dadd_nofma(xp[j + 1], -xj));
    double r = dadd_nofma(dmul_nofma(slope, dadd_nofma(x, -xj)), (double)j);
    if (r != r) {  // NumPy: "if we get nan in one direction, try the other"
        r = dadd_nofma(dmul_nofma(slope, dadd_nofma(x, -xp[j + 1])), (double)(j + 1));
    }
    return r;
}

// ---- stage 2b: attention transforms (new_method.py:133-188) ----------------------------------
enum Transform : int { T_IDENTITY = 0, T_SQUARE = 1, T_SQRT = 2, T_EXP = 3, T_LOG = 4 };

AW_HD double transform_fwd(double x, int t, double exp_scale, double exp_divisor) {
    switch (t) {
        case T_SQUARE: return x * x;
        case T_SQRT: return sqrt(x > 0.0 ? x : 0.0);
        case T_EXP: return exp(exp_scale * x) / exp_divisor;
        case T_LOG: return log(x + 1e-5);
        default: return x;
    }
}
// The same with the transform known at compile time (TR < 0: no transform at all): per-pixel loops are
// instantiated once per transform, so their bodies hold one transform instead of a five-way switch.
template <int TR>
AW_HD double transform_fwd_t(double x, double exp_scale, double exp_divisor) {
    return TR < 0 ? x : transform_fwd(x, TR, exp_scale, exp_divisor);
}
AW_HD double transform_inv(double x, int t, double exp_scale, double exp_divisor) {
    switch (t) {
        case T_SQUARE: return sqrt(x > 0.0 ? x : 0.0);
        case T_SQRT: return x * x;
        case T_EXP: {
            const double v = x * exp_divisor;
            return log(v > 1e-9 ? v : 1e-9) / exp_scale;
        }
        case T_LOG: return exp(x) - 1e-5;
        default: return x;
    }
}
// np.maximum(att.astype(float64), 0): NaN propagates through np.maximum.
AW_HD double clamp_nonneg(double a) { return (a != a) ? a : (a > 0.0 ? a : 0.0); }

}  // namespace aw
