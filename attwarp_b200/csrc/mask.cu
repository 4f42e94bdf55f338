// mask.cu -- the mask post-processing between stage 1 and stage 2b of the driver flow.
//
// Replaces, on the device, the mask half of blend_mask
// ("Attention Guided Warping/attention_extraction/llava.py:207-256", called at AGW/main.py:361 and
// AGW/main_batched.py:268):
//   revise_mask: normalize(min) -> z-score x coe -> sigmoid -> clamp -> k x k box filter (replicate pad)
//   ToPILImage : float -> uint8 by truncation of v * 255
//   PIL resize(image.size, LANCZOS) of the mode-'L' mask
// The resize restates Pillow's 8-bit two-pass resampler (src/libImaging/Resample.c) exactly:
// per output coordinate a window of taps with weights normalised in double and rounded to 22-bit
// fixed point (computed once per (in, out) size pair on the host, with the C library's sin like Pillow),
// int32 accumulation from 1 << 21, clip to uint8 after the horizontal pass and after the vertical pass.
// Everything here is a few hundred KB per batch: one CTA per image (revise) / per tile of output rows
// (resize, the horizontal pass of the token rows recomputed per CTA into shared memory).
#include <math.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace aw {
namespace {

constexpr int kMaskThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;      // Resample.c PRECISION_BITS
constexpr int kResizeRows = 32;                 // output rows per CTA

// ---- revise_mask + ToPILImage ---------------------------------------------------------------------
__global__ void __launch_bounds__(kMaskThreads)
revise_mask_kernel(const float* __restrict__ tok, int gh, int gw, int ksize, float coe,
                   float* __restrict__ revised, uint8_t* __restrict__ mask_u8) {
    extern __shared__ float sm_f[];
    __shared__ float red[32];
    __shared__ double redd[32];
    const int G = gh * gw;
    float* m = sm_f;                    // G
    const float* src = tok + (int64_t)blockIdx.x * G;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        const float v = src[i];
        m[i] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    mn = -block_max(-mn, red);
    const float range = mx - mn;
    float part = 0.f;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {          // normalize(mat, "min")
        const float v = (m[i] - mn) / range;
        m[i] = v;
        part += v;
    }
    const float mean = block_sum(part, red) / (float)G;
    double sq = 0.0;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {          // enhance: centre, unbiased std
        const float v = m[i] - mean;
        m[i] = v;
        sq += (double)v * (double)v;
    }
    const float sd = (float)sqrt(block_sum(sq, redd) / (double)(G - 1));
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        const float z = m[i] / sd * coe;
        const float s = 1.0f / (1.0f + expf(-z));
        m[i] = fminf(fmaxf(s, 0.f), 1.f);
    }
    __syncthreads();
    const int pad = (ksize - 1) / 2;
    const float w = 1.0f / (float)(ksize * ksize);
    for (int i = threadIdx.x; i < G; i += blockDim.x) {          // box filter, replicate padding
        const int y = i / gw, x = i - y * gw;
        float acc = 0.f;
        for (int dy = -pad; dy <= pad; ++dy)
            for (int dx = -pad; dx <= pad; ++dx)
                acc = fadd_nofma(acc, fmul_nofma(w, m[clampi(y + dy, 0, gh - 1) * gw + clampi(x + dx, 0, gw - 1)]));
        if (revised != nullptr) revised[(int64_t)blockIdx.x * G + i] = acc;
        if (mask_u8 != nullptr) mask_u8[(int64_t)blockIdx.x * G + i] = (uint8_t)(int)fmul_nofma(acc, 255.0f);   // .mul(255).byte()
    }
}

// ---- Pillow LANCZOS resize of mode-'L' images -------------------------------------------------------
// tables: bounds [out] = {first tap, taps}, weights [out][ks]
__global__ void __launch_bounds__(kMaskThreads)
resize_lanczos_u8_kernel(const uint8_t* __restrict__ src, int h, int w, int Ho, int Wo,
                         const int2* __restrict__ bx, const int* __restrict__ kx, int ksx,
                         const int2* __restrict__ by, const int* __restrict__ ky, int ksy,
                         uint8_t* __restrict__ dst) {
    extern __shared__ uint8_t sm_b[];
    uint8_t* in = sm_b;                          // h * w        source image
    uint8_t* tmp = sm_b + ((h * w + 15) & ~15);  // rows * Wo    horizontal pass of the rows this tile taps
    const int b = blockIdx.y;
    const int y0 = blockIdx.x * kResizeRows, y1 = min(y0 + kResizeRows, Ho);
    // source rows tapped by the output rows [y0, y1): windows move monotonically with y
    const int r0 = Ho != h ? by[y0].x : y0;
    const int r1 = Ho != h ? by[y1 - 1].x + by[y1 - 1].y : y1;
    const uint8_t* img = src + (int64_t)b * h * w;
    for (int i = threadIdx.x; i < (r1 - r0) * w; i += blockDim.x) in[i] = img[r0 * w + i];
    __syncthreads();
    for (int i = threadIdx.x; i < (r1 - r0) * Wo; i += blockDim.x) {     // horizontal pass
        const int r = i / Wo, x = i - r * Wo;
        if (Wo != w) {
            const int2 bd = bx[x];
            const int* k = kx + (int64_t)x * ksx;
            int acc = 1 << (kPrecisionBits - 1);
            for (int t = 0; t < bd.y; ++t) acc += (int)in[r * w + bd.x + t] * k[t];
            tmp[i] = (uint8_t)clampi(acc >> kPrecisionBits, 0, 255);
        } else {
            tmp[i] = in[i];
        }
    }
    __syncthreads();
    uint8_t* out = dst + (int64_t)b * Ho * Wo;
    for (int i = threadIdx.x; i < (y1 - y0) * Wo; i += blockDim.x) {     // vertical pass
        const int yy = i / Wo, x = i - yy * Wo;
        const int y = y0 + yy;
        if (Ho != h) {
            const int2 bd = by[y];
            const int* k = ky + (int64_t)y * ksy;
            int acc = 1 << (kPrecisionBits - 1);
            for (int t = 0; t < bd.y; ++t) acc += (int)tmp[(bd.x - r0 + t) * Wo + x] * k[t];
            out[(int64_t)y * Wo + x] = (uint8_t)clampi(acc >> kPrecisionBits, 0, 255);
        } else {
            out[(int64_t)y * Wo + x] = tmp[yy * Wo + x];
        }
    }
}

// ---- the up-scaling case (a token-grid mask to image size: <= 8 taps per axis) ----------------------
// Same arithmetic, organised for throughput: the CTA owns a tile of output rows of one image, computes the
// horizontal pass of the (few) source rows the tile taps into shared memory once, then every thread owns
// FOUR adjacent output columns and walks down its share of the tile's rows with the tapped rows of the
// horizontal pass held in registers as a sliding window (the window moves one source row every Ho / h
// output rows); an output row is 8 x 4 IMAD with coefficients broadcast from shared memory, clip, one
// 32-bit store.  Zero-padded coefficient rows make every output row an 8-tap row.
constexpr int kUpTaps = 8;        // most taps per axis (coefficient rows in shared memory are padded to 8)
__device__ __forceinline__ void unpack4(uint32_t w, int* v) {
    v[0] = (int)(w & 0xffu); v[1] = (int)((w >> 8) & 0xffu); v[2] = (int)((w >> 16) & 0xffu); v[3] = (int)(w >> 24);
}
__device__ __forceinline__ uint32_t clip8(int acc) { return (uint32_t)clampi(acc >> kPrecisionBits, 0, 255); }

//
// MARG (the fusion of llava.py:243-255 with new_method.py:215-216): the mask is not written at all;
// the kernel leaves the marginal sums of its bytes instead, in the partial layout of profiles.cu (P2b) -- column
// sums per (tile, row group) chunk in colpart[b][chunk][Wo], whole row sums in rowpart[b][0][Ho], each with the
// + 1e-9 per element of new_method.py:212 added as count x 1e-9 -- for maps_from_partials_kernel to finish.  All
// sums are exact integers (<= 255 x 65535), whatever the order the threads add them in.
// Measured (profiles/r03d_row_kernels.txt): this saves the B x H x W buffer, not time -- the resize is bound by the
// fma pipe, so the write it skips was free, and the sums cost about what the separate marginals kernel does
// (256 x 336^2: 53.5 us fused vs 52.5 us in two steps; 64 x 1344^2: 188 vs 167 us).
constexpr int kUpMaxThreads = 512;
template <int TAPS, bool MARG>     // taps per axis actually used (7 for pure up-scaling)
__global__ void __launch_bounds__(kUpMaxThreads)
resize_lanczos_up_kernel(const uint8_t* __restrict__ src, int h, int w, int Ho, int Wo,
                         const int2* __restrict__ bx, const int* __restrict__ kx, int ksx,
                         const int2* __restrict__ by, const int* __restrict__ ky, int ksy,
                         int rows_per_tile, int n_rowgroups, uint8_t* __restrict__ dst,
                         double* __restrict__ colpart, double* __restrict__ rowpart) {
    extern __shared__ __align__(16) uint8_t sm_b[];
    const int Wp = (Wo + 3) & ~3;                        // pitch of the horizontal pass
    const int b = blockIdx.y;
    const int y0 = blockIdx.x * rows_per_tile, y1 = min(y0 + rows_per_tile, Ho);
    const int r0 = by[y0].x;
    const int nr = by[y1 - 1].x + by[y1 - 1].y - r0;     // source rows the tile taps: [r0, r0 + nr)
    int4* coef = reinterpret_cast<int4*>(sm_b);          // [rows_per_tile][2] int4: 8 coefficients per output row
    int* ymin = reinterpret_cast<int*>(sm_b + (size_t)rows_per_tile * 32);      // [rows_per_tile], relative to r0
    uint8_t* in = sm_b + (size_t)rows_per_tile * 36;                            // [nr][w]
    uint8_t* tmp = in + (((size_t)h * w + 15) & ~(size_t)15);                   // [nr][Wp]
    const uint8_t* img = src + ((int64_t)b * h + r0) * w;
    for (int i = threadIdx.x; i < nr * w; i += blockDim.x) in[i] = img[i];
    // MARG: the row sums of the tile are added up in shared memory
    int* rsum = reinterpret_cast<int*>(tmp + (((size_t)h * Wp + 15) & ~(size_t)15));   // [rows_per_tile], MARG only
    if (MARG)
        for (int i = threadIdx.x; i < y1 - y0; i += blockDim.x) rsum[i] = 0;
    for (int i = threadIdx.x; i < (y1 - y0) * kUpTaps; i += blockDim.x) {
        const int yy = i >> 3, t = i & 7;
        reinterpret_cast<int*>(coef)[i] = t < ksy ? __ldg(ky + (int64_t)(y0 + yy) * ksy + t) : 0;
        if (t == 0) ymin[yy] = by[y0 + yy].x - r0;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < Wp; x += blockDim.x) {                      // horizontal pass, column x
        int k[TAPS], c0 = 0;
        if (x < Wo) c0 = bx[x].x;
#pragma unroll
        for (int t = 0; t < TAPS; ++t) k[t] = (x < Wo && t < ksx) ? __ldg(kx + (int64_t)x * ksx + t) : 0;
        for (int r = 0; r < nr; ++r) {
            int acc = 1 << (kPrecisionBits - 1);
#pragma unroll
            for (int t = 0; t < TAPS; ++t) acc += (int)in[r * w + min(c0 + t, w - 1)] * k[t];
            tmp[r * Wp + x] = (uint8_t)clip8(acc);
        }
    }
    __syncthreads();
    // vertical pass: thread = (row group, column group of 4)
    const int n_cg = Wp >> 2;
    const int rg = threadIdx.x / n_cg;
    // MARG: the lanes of this warp that work on the same row group (they add their row sums up with one REDUX)
    const unsigned peers = MARG ? __match_any_sync(0xffffffffu, rg) : 0u;
    if (!MARG && rg >= n_rowgroups) return;
    const int cg = threadIdx.x - rg * n_cg;
    const int rows = y1 - y0;
    const bool live = rg < n_rowgroups;
    const int ya = live ? (int)(((int64_t)rows * rg) / n_rowgroups) : 0;
    const int yb = live ? (int)(((int64_t)rows * (rg + 1)) / n_rowgroups) : 0;
    const bool vec = (Wo & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    uint8_t* out = dst + ((int64_t)b * Ho + y0) * Wo;
    uint32_t csum[4] = {0u, 0u, 0u, 0u};
    {
        const int g = cg;               // one column group per thread (the launcher sizes the CTA so)
        int win[TAPS][4];
        int base = -1000;
        for (int yy = ya; yy < yb; ++yy) {
            const int ym = ymin[yy];
            if (ym != base) {
                if (ym == base + 1) {
#pragma unroll
                    for (int t = 0; t < TAPS - 1; ++t)
#pragma unroll
                        for (int q = 0; q < 4; ++q) win[t][q] = win[t + 1][q];
                    unpack4(*reinterpret_cast<const uint32_t*>(tmp + min(ym + TAPS - 1, nr - 1) * Wp + 4 * g), win[TAPS - 1]);
                } else {
#pragma unroll
                    for (int t = 0; t < TAPS; ++t)
                        unpack4(*reinterpret_cast<const uint32_t*>(tmp + min(ym + t, nr - 1) * Wp + 4 * g), win[t]);
                }
                base = ym;
            }
            const int4 ca = coef[2 * yy], cb = coef[2 * yy + 1];
            const int c[kUpTaps] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
            int acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = 1 << (kPrecisionBits - 1);
#pragma unroll
            for (int t = 0; t < TAPS; ++t)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += win[t][q] * c[t];
            // clip8 x 4 + pack: cvt.pack saturates two int32 to uint8 and packs them under the previous pair
            uint32_t hi2, o;
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi2) : "r"(acc[3] >> kPrecisionBits), "r"(acc[2] >> kPrecisionBits), "r"(0));
            asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(o) : "r"(acc[1] >> kPrecisionBits), "r"(acc[0] >> kPrecisionBits), "r"(hi2));
            if (MARG) {
                // columns >= Wo of the last group are zero (zero horizontal coefficients), so they add nothing
                csum[0] += o & 0xffu; csum[1] += (o >> 8) & 0xffu; csum[2] += (o >> 16) & 0xffu; csum[3] += o >> 24;
                const uint32_t part = __reduce_add_sync(peers, __dp4a(o, 0x01010101u, 0u));
                if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&rsum[yy], (int)part);
                continue;
            }
            uint8_t* p = out + (int64_t)yy * Wo + 4 * g;
            if (vec) {
                *reinterpret_cast<uint32_t*>(p) = o;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * g + q < Wo) p[q] = (uint8_t)(o >> (8 * q));
            }
        }
    }
    if (MARG) {
        const int n_chunks = gridDim.x * n_rowgroups;
        if (live) {
            double* cp = colpart + ((int64_t)b * n_chunks + blockIdx.x * n_rowgroups + rg) * Wo;
            const double base = (double)(yb - ya) * kBaseAttention;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * cg + q < Wo) cp[4 * cg + q] = (double)csum[q] + base;
        }
        __syncthreads();
        const double row_base = (double)Wo * kBaseAttention;
        for (int i = threadIdx.x; i < rows; i += blockDim.x)
            rowpart[(int64_t)b * Ho + y0 + i] = (double)rsum[i] + row_base;
    }
}

// ---- coefficient tables (Resample.c precompute_coeffs + normalize_coeffs_8bpc, Lanczos, support 3) ----
double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
double lanczos_filter(double x) { return (-3.0 <= x && x < 3.0) ? sinc_filter(x) * sinc_filter(x / 3.0) : 0.0; }

struct CoeffTable {
    int2* bounds = nullptr;     // device
    int* weights = nullptr;     // device
    int ksize = 0;
};

int build_table(int in_size, int out_size, CoeffTable* t) {
    const double scale = (double)in_size / (double)out_size;
    const double fscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * fscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    std::vector<int2> bounds((size_t)out_size);
    std::vector<int> kk((size_t)out_size * ksize, 0);
    std::vector<double> w((size_t)ksize);
    const double ss = 1.0 / fscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[(size_t)x] = lanczos_filter((x + xmin - center + 0.5) * ss);
            ww += w[(size_t)x];
        }
        for (int x = 0; x < xmax; ++x) {
            const double v = ww != 0.0 ? w[(size_t)x] / ww : w[(size_t)x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
        }
        bounds[(size_t)xx] = make_int2(xmin, xmax);
    }
    AW_CUDA(cudaMalloc(&t->bounds, sizeof(int2) * (size_t)out_size));
    AW_CUDA(cudaMalloc(&t->weights, sizeof(int) * kk.size()));
    AW_CUDA(cudaMemcpy(t->bounds, bounds.data(), sizeof(int2) * (size_t)out_size, cudaMemcpyHostToDevice));
    AW_CUDA(cudaMemcpy(t->weights, kk.data(), sizeof(int) * kk.size(), cudaMemcpyHostToDevice));
    t->ksize = ksize;
    return ATTWARP_OK;
}

// Tables are constants of (device, in, out): built once (synchronous upload) and kept for the life of the
// process, so later calls on any stream only read them.
int get_table(int in_size, int out_size, CoeffTable* out) {
    static std::mutex mu;
    static std::map<std::pair<int, std::pair<int, int>>, CoeffTable> cache;
    int dev = 0;
    AW_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_pair(dev, std::make_pair(in_size, out_size));
    auto it = cache.find(key);
    if (it == cache.end()) {
        CoeffTable t;
        const int rc = build_table(in_size, out_size, &t);
        if (rc != ATTWARP_OK) return rc;
        it = cache.emplace(key, t).first;
    }
    *out = it->second;
    return ATTWARP_OK;
}

}  // namespace

int launch_revise_mask(const float* tok, int B, int gh, int gw, int ksize, float coe, float* revised,
                       uint8_t* mask_u8, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)gh * gw;
    if (smem > 160 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "revise_mask: token grid %dx%d too large", gh, gw);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(revise_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    revise_mask_kernel<<<B, kMaskThreads, smem, st>>>(tok, gh, gw, ksize, coe, revised, mask_u8);
    return check_launch("revise_mask_kernel");
}

// Launch geometry of the up-scaling kernel: tiles of `rows` output rows (enough CTAs to fill the GPU, as tall as
// possible: the horizontal pass is redone per tile), n_rg row groups of Wp / 4 threads each.
struct UpGeometry { bool ok; int n_tiles, rows, n_rg, threads; size_t smem; };
static UpGeometry up_geometry(int B, int h, int w, int Ho, int Wo) {
    UpGeometry g{};
    const int Wp = (Wo + 3) & ~3, n_cg = Wp / 4;
    if (n_cg > kUpMaxThreads) return g;
    int n_tiles = (2 * sm_count() + B - 1) / B;
    n_tiles = n_tiles < 1 ? 1 : n_tiles;
    if (n_tiles > (Ho + 31) / 32) n_tiles = (Ho + 31) / 32;
    g.rows = (Ho + n_tiles - 1) / n_tiles;
    g.n_tiles = (Ho + g.rows - 1) / g.rows;
    g.n_rg = kUpMaxThreads / n_cg;
    g.threads = (n_cg * g.n_rg + 31) & ~31;
    // coefficients + first taps | source rows | horizontal pass | row sums (MARG)
    g.smem = (size_t)g.rows * 36 + (((size_t)h * w + 15) & ~(size_t)15) + (((size_t)h * Wp + 15) & ~(size_t)15) +
             (size_t)g.rows * 4;
    g.ok = g.smem <= 200 * 1024;
    return g;
}

// Partial-sum chunks per image the fused kernel writes (0: this resize is not taken by the fused kernel).
int lanczos_marginals_chunks(int B, int h, int w, int H, int W) {
    if (H <= h || W <= w || h > 0xffff || w > 0xffff) return 0;
    // pure up-scaling: 3 * 2 + 1 = 7 taps per axis
    const UpGeometry g = up_geometry(B, h, w, H, W);
    return g.ok ? g.n_tiles * g.n_rg : 0;
}

// mask_u8 [B][h][w] (revise_mask's uint8 output) -> marginal partial sums of its LANCZOS resize to H x W, which is
// never written: colpart [B][chunks][W], rowpart [B][1][H].
int launch_lanczos_marginals(const uint8_t* src, int B, int h, int w, int H, int W, double* colpart, double* rowpart,
                             int* n_chunks, cudaStream_t st) {
    const UpGeometry ug = up_geometry(B, h, w, H, W);
    if (H <= h || W <= w || !ug.ok)
        return fail(ATTWARP_ERR_UNSUPPORTED, "lanczos_marginals: %dx%d -> %dx%d is not an up-scaling that fits shared memory", h, w, H, W);
    CoeffTable tx, ty;
    int rc = get_table(w, W, &tx);
    if (rc != ATTWARP_OK) return rc;
    rc = get_table(h, H, &ty);
    if (rc != ATTWARP_OK) return rc;
    if (tx.ksize > kUpTaps || ty.ksize > kUpTaps) return fail(ATTWARP_ERR_UNSUPPORTED, "lanczos_marginals: too many taps");
    auto kern = (tx.ksize <= 7 && ty.ksize <= 7) ? resize_lanczos_up_kernel<7, true> : resize_lanczos_up_kernel<8, true>;
    if (ug.smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ug.smem));
    kern<<<dim3(ug.n_tiles, B), ug.threads, ug.smem, st>>>(
        src, h, w, H, W, tx.bounds, tx.weights, tx.ksize, ty.bounds, ty.weights, ty.ksize, ug.rows, ug.n_rg, nullptr,
        colpart, rowpart);
    *n_chunks = ug.n_tiles * ug.n_rg;
    return check_launch("resize_lanczos_up_kernel<marginals>");
}

int launch_resize_lanczos_u8(const uint8_t* src, int B, int h, int w, int Ho, int Wo, uint8_t* dst, cudaStream_t st) {
    CoeffTable tx, ty;
    if (Wo != w) {
        const int rc = get_table(w, Wo, &tx);
        if (rc != ATTWARP_OK) return rc;
    }
    if (Ho != h) {
        const int rc = get_table(h, Ho, &ty);
        if (rc != ATTWARP_OK) return rc;
    }
    // up-scaling (or any resize with <= 8 taps per axis) of both axes: the register-window kernel
    const UpGeometry ug = up_geometry(B, h, w, Ho, Wo);
    if (Wo != w && Ho != h && tx.ksize <= kUpTaps && ty.ksize <= kUpTaps && ug.ok) {
        auto kern = (tx.ksize <= 7 && ty.ksize <= 7) ? resize_lanczos_up_kernel<7, false> : resize_lanczos_up_kernel<8, false>;
        if (ug.smem > 48 * 1024)
            AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ug.smem));
        kern<<<dim3(ug.n_tiles, B), ug.threads, ug.smem, st>>>(
            src, h, w, Ho, Wo, tx.bounds, tx.weights, tx.ksize, ty.bounds, ty.weights, ty.ksize, ug.rows, ug.n_rg, dst,
            nullptr, nullptr);
        return check_launch("resize_lanczos_up_kernel");
    }
    // shared memory: the source rows a tile taps (at most all of them) + their horizontal pass
    const size_t smem = (((size_t)h * w + 15) & ~(size_t)15) + (size_t)h * Wo;
    if (smem > 200 * 1024)
        return fail(ATTWARP_ERR_UNSUPPORTED, "resize_lanczos: %dx%d -> width %d does not fit shared memory", h, w, Wo);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(resize_lanczos_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resize_lanczos_u8_kernel<<<dim3((Ho + kResizeRows - 1) / kResizeRows, B), kMaskThreads, smem, st>>>(
        src, h, w, Ho, Wo, tx.bounds, tx.weights, tx.ksize, ty.bounds, ty.weights, ty.ksize, dst);
    return check_launch("resize_lanczos_u8_kernel");
}

}  // namespace aw
