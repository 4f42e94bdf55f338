// profiles.cu -- stages 2b-4: attention -> marginal profiles -> CDF knots -> inverse-CDF maps.
//
// Replaces, in float64 like the reference:
//   * warp_image_by_attention up to np.interp ("Attention Guided Warping/new_method.py:207-261")
//   * the knot/tie-break/np.interp part of warp_from_cdf_torch
//     ("model/marginalnet_full_dataset/checkpoint_utils.py:157-189")
//
// Work per image is O(H*W) reads for a materialised attention map (marginals_partial_kernel,
// many CTAs per image) and O(H+W) afterwards (one CTA per image: block scan in shared memory,
// knots in shared memory, one bisection per output coordinate).  The 2-D meshgrid of the
// reference (new_method.py:263-265) is never built: the maps stay separable.
#include <type_traits>

#include "common.cuh"

namespace aw {
namespace {

constexpr int kProfThreads = 256;
constexpr int kMargThreads = 256;
constexpr int kMargRows = 64;          // rows per CTA of the marginals kernel
constexpr int kMargColsPerThread = 4;  // columns per thread
constexpr int kMargCols = kMargThreads * kMargColsPerThread;

struct TransformArgs {
    int transform;
    int apply_inverse;
    double exp_scale;
    double exp_divisor;
};

// -------------------------------------------------------------------------------------------
// Shared tail: profiles (shared memory) -> knots -> inverse maps for one image.
// The CTA is two groups of kProfThreads threads: group 0 inverts the x axis while group 1 inverts
// the y axis (the two chains of scan -> knots -> search are independent and latency-bound).
//   prof_x[W], prof_y[H]: raw marginal sums of the biased map (new_method.py:215-216)
//   knots: 2 x (max(W,H)+1) doubles; red: 2 x 32 doubles; xchg: 4 doubles
// -------------------------------------------------------------------------------------------
constexpr int kTailRed = 32;       // doubles of reduction scratch per group
constexpr int kTailXchg = 4;       // totals exchanged between the groups

__device__ void invert_axis(const Group& g, double* prof, int n, double total, int n_out, double* knots,
                            double* red, float* __restrict__ out_map) {
    group_inclusive_scan(g, prof, n, red);                    // np.cumsum            :242,248
    for (int i = g.tid; i <= n; i += g.nt) {
        double k;
        if (i == 0) k = 0.0;
        else if (i == n) k = (double)n_out;                   // forced last knot      :254-255
        else k = dmul_nofma(ddiv_exact(prof[i - 1], total), (double)n_out);  //       :243-245
        knots[i] = k;
    }
    g.sync();
    for (int j = g.tid; j < n_out; j += g.nt)                 // np.interp + f32 cast :260-265
        out_map[j] = (float)interp_index((double)j, knots, n + 1);
}

// Called by all 2 * kProfThreads threads of the CTA.
__device__ void profiles_to_maps(double* prof_x, double* prof_y, int W, int H, int Wo, int Ho,
                                 const TransformArgs& ta, double* knots, double* red, double* xchg,
                                 float* __restrict__ map_x, float* __restrict__ map_y,
                                 int* __restrict__ fallback_flag) {
    const int axis = threadIdx.x >= kProfThreads ? 1 : 0;     // group 0: x (columns), group 1: y (rows)
    const Group g = {(int)threadIdx.x - axis * kProfThreads, kProfThreads, 1 + axis};
    double* prof = axis ? prof_y : prof_x;
    const int n = axis ? H : W, n_other = axis ? W : H, n_out = axis ? Ho : Wo;
    double* my_red = red + axis * kTailRed;
    double* my_knots = knots + axis * (max(W, H) + 1);

    // sum of the biased map (needed by the fallback's np.mean) = sum of the raw row sums
    if (axis == 1) {
        double part = 0.0;
        for (int i = g.tid; i < H; i += g.nt) part += prof_y[i];
        const double sum_biased = group_sum(g, part, my_red);
        if (g.tid == 0) xchg[2] = sum_biased;
    }
    if (ta.apply_inverse) {                                   // :219-226
        const double bias = kBaseAttention * (double)n_other;
        for (int i = g.tid; i < n; i += g.nt)
            prof[i] = transform_inv(prof[i] - bias, ta.transform, ta.exp_scale, ta.exp_divisor) + bias;
        g.sync();
    }
    double part = 0.0;
    for (int i = g.tid; i < n; i += g.nt) part += prof[i];
    double total = group_sum(g, part, my_red);                // :228-229
    if (g.tid == 0) xchg[axis] = total;
    __syncthreads();
    const bool fallback = (xchg[0] < kEpsilon) || (xchg[1] < kEpsilon);   // :231
    if (fallback) {                                           // :233-239
        for (int i = g.tid; i < n; i += g.nt) prof[i] = 1.0;
        const double mean = xchg[2] / ((double)H * (double)W);
        total = fmax((double)n * (mean * (double)n_other), kEpsilon);
        g.sync();
    }
    if (fallback_flag != nullptr && threadIdx.x == 0) *fallback_flag = fallback ? 1 : 0;
    invert_axis(g, prof, n, total, n_out, my_knots, my_red, axis ? map_y : map_x);
}

// -------------------------------------------------------------------------------------------
// (P1) maps from a token grid that is index-upsampled on the fly.
//   tok[b][gh*gw] given either final (nsplit==1, scale==1) or as stage-1 partials
//   [b][nsplit][gh*gw] to be summed in split order and scaled (fused stage-1 finalize).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * kProfThreads)
maps_from_tokens_kernel(const float* __restrict__ tok, int nsplit, float scale,
                        float* __restrict__ tok_out, int gh, int gw, int H, int W, int Wo, int Ho,
                        TransformArgs ta, float* __restrict__ map_x, float* __restrict__ map_y,
                        int* __restrict__ fallback_flags, const RaggedImage* __restrict__ imgs) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int G = gh * gw;
    if (imgs != nullptr) {                    // ragged batch: sizes and map rows per image
        H = imgs[b].H; W = imgs[b].W; Wo = imgs[b].Wo; Ho = imgs[b].Ho;
        map_x = const_cast<float*>(imgs[b].map_x);
        map_y = const_cast<float*>(imgs[b].map_y);
    } else {
        map_x += (int64_t)b * Wo;
        map_y += (int64_t)b * Ho;
    }
    double* red = smem;                       // 2 * kTailRed
    double* xchg = red + 2 * kTailRed;        // kTailXchg
    double* grid = xchg + kTailXchg;          // G   transformed + biased token values
    double* csum = grid + G;                  // gw  column sums over the full-res rows
    double* rsum = csum + gw;                 // gh
    double* prof_x = rsum + gh;               // W
    double* prof_y = prof_x + W;              // H
    double* knots = prof_y + H;               // 2 * (max(W,H)+1)

    const float* src = tok + (int64_t)b * nsplit * G;
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        float v;
        if (nsplit == 1 && scale == 1.0f) {
            v = src[i];
        } else {
            float s = 0.f;
            for (int k = 0; k < nsplit; ++k) s += src[(int64_t)k * G + i];
            v = s * scale;
        }
        if (tok_out != nullptr) tok_out[(int64_t)b * G + i] = v;
        const double a = clamp_nonneg((double)v);                                   // :207-208
        grid[i] = transform_fwd(a, ta.transform, ta.exp_scale, ta.exp_divisor) + kBaseAttention;  // :210-212
    }
    __syncthreads();
    // number of full-res rows (cols) that index-map to grid row i (col c): the index upsample
    // att[y][x] = tok[(y*gh)/H][(x*gw)/W] gives row i the rows [ceil(i*H/gh), ceil((i+1)*H/gh)).
    // (sizes fit 32 bits: H, W < 65536 and gh, gw <= 8192)
    double* cnt_y = knots;                    // gh   (knots are not live yet)
    double* cnt_x = knots + gh;               // gw
    for (int i = threadIdx.x; i < gh + gw; i += blockDim.x) {
        if (i < gh) {
            cnt_y[i] = (double)(((i + 1) * H + gh - 1) / gh - (i * H + gh - 1) / gh);
        } else {
            const int c = i - gh;
            cnt_x[c] = (double)(((c + 1) * W + gw - 1) / gw - (c * W + gw - 1) / gw);
        }
    }
    __syncthreads();
    if (threadIdx.x < kProfThreads) {                                       // first group: column sums
        for (int c = threadIdx.x; c < gw; c += kProfThreads) {
            double s = 0.0;
            for (int i = 0; i < gh; ++i) s += cnt_y[i] * grid[i * gw + c];
            csum[c] = s;
        }
    } else {                                                                // second group: row sums
        for (int i = threadIdx.x - kProfThreads; i < gh; i += kProfThreads) {
            double s = 0.0;
            for (int c = 0; c < gw; ++c) s += cnt_x[c] * grid[i * gw + c];
            rsum[i] = s;
        }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) prof_x[x] = csum[(x * gw) / W];
    for (int y = threadIdx.x; y < H; y += blockDim.x) prof_y[y] = rsum[(y * gh) / H];
    __syncthreads();
    profiles_to_maps(prof_x, prof_y, W, H, Wo, Ho, ta, knots, red, xchg, map_x, map_y,
                     fallback_flags ? fallback_flags + b : nullptr);
}

// -------------------------------------------------------------------------------------------
// (P2) marginals of a materialised attention map: one pass over [H][W].
//   grid = (col tiles, row chunks, B).  Each thread owns kMargColsPerThread adjacent columns and
//   walks the rows of its chunk: column sums stay in registers (written as per-chunk partials),
//   row sums are reduced per row with a warp-shuffle tree and combined across warps in shared
//   memory (per-column-tile partials).  Both partial sets are summed in fixed order by the
//   finish kernel -> deterministic.
//   TR >= 0: NumPy path (clamp, transform TR, + 1e-9)   TR == kClampOnly: gt_marginals (clamp only).
//   The transform is a template parameter: the pixel loops hold one transform instead of a five-way run-time switch
//   (9000 -> 1000-3700 SASS instructions per kernel; by itself that did not change the run time -- the float64
//   evaluation of the transform is the cost, see marginals_f32_rows_kernel -- but it lets each transform have its own
//   arithmetic).  uint8 maps (TR == kByteTable) look their 256 possible values up in a shared-memory table the CTA
//   fills with the run-time transform first.
// -------------------------------------------------------------------------------------------
constexpr int kClampOnly = -1;
constexpr int kByteTable = 8;
template <typename T, int TR>
__global__ void __launch_bounds__(kMargThreads)
marginals_partial_kernel(const T* __restrict__ att, int H, int W, TransformArgs ta,
                         double* __restrict__ colpart, double* __restrict__ rowpart,
                         int n_row_chunks, int n_col_tiles) {
    __shared__ double s_row[kMargRows][kMargThreads / 32];
    __shared__ double s_tab[TR == kByteTable ? 256 : 1];
    if (TR == kByteTable) {
        for (int v = threadIdx.x; v < 256; v += kMargThreads)
            s_tab[v] = transform_fwd((double)v, ta.transform, ta.exp_scale, ta.exp_divisor) + kBaseAttention;
        __syncthreads();
    }
    const int b = blockIdx.z, chunk = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x0 = tile * kMargCols + threadIdx.x * kMargColsPerThread;
    const int y0 = chunk * kMargRows, y1 = min(y0 + kMargRows, H);
    const T* img = att + (int64_t)b * H * W;

    double col[kMargColsPerThread];
#pragma unroll
    for (int k = 0; k < kMargColsPerThread; ++k) col[k] = 0.0;

    for (int y = y0; y < y1; ++y) {
        double rs = 0.0;
#pragma unroll
        for (int k = 0; k < kMargColsPerThread; ++k) {
            const int x = x0 + k;
            if (x < W) {
                double a;
                if (TR == kByteTable) {
                    a = s_tab[(int)__ldg(reinterpret_cast<const uint8_t*>(img + (int64_t)y * W + x))];
                } else {
                    a = clamp_nonneg(load_as_double<T>(img + (int64_t)y * W + x));
                    if (TR >= 0) a = transform_fwd_t<TR>(a, ta.exp_scale, ta.exp_divisor) + kBaseAttention;
                }
                col[k] += a;
                rs += a;
            }
        }
        rs = warp_sum(rs);
        if (lane == 0) s_row[y - y0][wid] = rs;
    }
    double* cp = colpart + ((int64_t)b * n_row_chunks + chunk) * W;
#pragma unroll
    for (int k = 0; k < kMargColsPerThread; ++k)
        if (x0 + k < W) cp[x0 + k] = col[k];
    __syncthreads();
    double* rp = rowpart + ((int64_t)b * n_col_tiles + tile) * H;
    for (int r = threadIdx.x; r < y1 - y0; r += kMargThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kMargThreads / 32; ++w) s += s_row[r][w];
        rp[y0 + r] = s;
    }
}

// (P2b) the driver case of (P2): uint8 attention map (the mask blend_mask returns, main.py:361), identity
// transform, rows that are multiples of 16 bytes.  HBM-bound as it should be: every lane reads 16 pixels
// per load (8-12 loads in flight per lane: rows x 512-column steps), a warp owns whole rows of its column tile
// (row sum = 4 dp4a per load + one REDUX per row), column sums are carried as packed 16-bit lanes (a warp
// adds at most 256 / 8 = 32 rows of <= 255 before they are widened) and combined across the CTA's warps
// in shared memory.  All sums are exact integers; the + 1e-9 per element (new_method.py:212) is added as
// count x 1e-9 when the partial is written (differs from the reference's term-by-term float64 sum by
// < 1e-15 relative).  Partial layout and finish kernel as (P2), with its own chunk / tile geometry.
constexpr int kU8TileCols = 1536, kU8MaxRows = 256;
// STEPS: 512-column steps a warp takes across its tile (tile_cols <= 512 * STEPS); U: rows in flight per warp
template <int STEPS, int U>
__global__ void __launch_bounds__(kMargThreads)
marginals_u8_identity_kernel(const uint8_t* __restrict__ att, int H, int W, int rows_per_cta,
                             double* __restrict__ colpart, double* __restrict__ rowpart, int n_row_chunks,
                             int n_col_tiles) {
    constexpr int kWarps = kMargThreads / 32;
    __shared__ uint32_t cw[kWarps][STEPS * 32 * 8];        // per warp: the packed accumulators, as they are
    const int b = blockIdx.z, chunk = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int y0 = chunk * rows_per_cta, y1 = min(y0 + rows_per_cta, H);
    const int xt = tile * kU8TileCols;                     // first column of the tile
    const int tile_cols = min(kU8TileCols, W - xt);
    const uint8_t* img = att + (int64_t)b * H * W + xt;

    uint32_t acc[STEPS][8];                                // [step][2 * word + odd]: two 16-bit column sums
#pragma unroll
    for (int s = 0; s < STEPS; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[s][k] = 0u;
    double* rp = rowpart + ((int64_t)b * n_col_tiles + tile) * H;
    const double row_base = (double)tile_cols * kBaseAttention;
    // Double-buffered: the loads of the next group of U rows are in flight while this group is summed (the
    // arithmetic of a group, ~30 instructions per load, takes about as long as the loads' latency).
    auto fetch = [&](int y, uint4 (&v)[U][STEPS]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool row_in = y + u * kWarps < y1;
            const uint8_t* r = img + (int64_t)(y + u * kWarps) * W;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const int x = s * 512 + lane * 16;
                v[u][s] = (row_in && x < tile_cols) ? __ldg(reinterpret_cast<const uint4*>(r + x))
                                                    : make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    uint4 v[U][STEPS], nx[U][STEPS];
    fetch(y0 + wid, v);
    for (int y = y0 + wid; y < y1; y += U * kWarps) {
        fetch(y + U * kWarps, nx);                         // rows past y1 load nothing
        uint32_t rs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            rs[u] = 0u;
#pragma unroll
            for (int s = 0; s < STEPS; ++s) {
                const uint32_t a[4] = {v[u][s].x, v[u][s].y, v[u][s].z, v[u][s].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    rs[u] = __dp4a(a[k], 0x01010101u, rs[u]);
                    acc[s][2 * k] += a[k] & 0x00ff00ffu;
                    acc[s][2 * k + 1] += (a[k] >> 8) & 0x00ff00ffu;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t t = __reduce_add_sync(0xffffffffu, rs[u]);
            if (lane == 0 && y + u * kWarps < y1) rp[y + u * kWarps] = (double)t + row_base;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int s = 0; s < STEPS; ++s) v[u][s] = nx[u][s];
    }
#pragma unroll
    for (int s = 0; s < STEPS; ++s)
#pragma unroll
        for (int k = 0; k < 8; ++k) cw[wid][(s * 32 + lane) * 8 + k] = acc[s][k];
    __syncthreads();
    double* cp = colpart + ((int64_t)b * n_row_chunks + chunk) * W + xt;
    const double base = (double)(y1 - y0) * kBaseAttention;
    for (int c = threadIdx.x; c < tile_cols; c += kMargThreads) {
        // column c sits in word 2k + (q & 1), half q >> 1 of its lane's step block (k = word, q = byte)
        const int blk = c >> 4, k = (c >> 2) & 3, q = c & 3;
        const int wi = blk * 8 + 2 * k + (q & 1), sh = (q >> 1) * 16;
        uint32_t t = 0u;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += (cw[w][wi] >> sh) & 0xffffu;
        cp[c] = (double)t + base;
    }
}

// Rows per CTA for (P2b) / (P2c): as many as possible (P2b: <= 256, the 16-bit lanes) without a ragged last wave:
// minimise  waves x (rows + the epilogue's worth of rows)  over multiples of 8.
static int rows_per_cta(int B, int H, int nct, int ctas_per_sm, int max_rows) {
    const int64_t slots = ctas_per_sm * (int64_t)sm_count();
    int best = kMargRows;
    int64_t best_cost = -1;
    for (int rows = kMargRows; rows <= max_rows; rows += 8) {      // >= kMargRows: the workspace is sized for those
        const int64_t ctas = (int64_t)B * nct * ((H + rows - 1) / rows);
        const int64_t cost = ((ctas + slots - 1) / slots) * (rows + 24);
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && rows > best)) { best_cost = cost; best = rows; }
    }
    return best;
}

// sqrt of a non-negative float32 as a float32 pair hi + lo with relative error < 2^-47 (the correctly rounded float64
// sqrt has 2^-53; summing a 1344-pixel row in float64 in another order already moves the sum by up to 2^-43):
// hardware rsqrt seed, one Newton step in float32 for hi (2^-23), the exact residual x - hi^2 by fma, and
// lo = residual / (2 hi) with a refined reciprocal (2^-47.5 worst case over [1e-28, 3e38) with the seed off by
// +-2 ulp, checked on the host).  0 -> (0, 0); tiny, infinite and NaN inputs take the float64 sqrt.
__device__ __forceinline__ void sqrt_pair(float x, float& hi, float& lo) {
    if (x >= 1.0e-27f && x < 3.0e38f) {              // the residual below stays a normal float32
        float y;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
        const float s = __fmul_rn(x, y);
        float h = __fmul_rn(0.5f, y);
        hi = __fmaf_rn(__fmaf_rn(-s, s, x), h, s);
        h = __fmaf_rn(__fadd_rn(h, h), __fmaf_rn(-hi, h, 0.5f), h);     // h -> 1 / (2 hi)
        lo = __fmul_rn(__fmaf_rn(-hi, hi, x), h);
    } else if (x == 0.f) {
        hi = 0.f;
        lo = 0.f;
    } else {
        const double r = sqrt((double)x);
        hi = (float)r;
        lo = hi < __int_as_float(0x7f800000) ? (float)(r - (double)hi) : 0.f;   // NaN: hi < inf is false
    }
}

// (P2c) float32 attention maps with rows that are multiples of 4 floats (gt_marginals' A_full,
// checkpoint_utils.py:43-51; float att maps of the NumPy path): the (P2b) organisation in float64.
// A warp owns whole rows of a 512-column tile (4 float4 loads per lane and row, two rows in flight), column
// sums stay in 16 float64 registers per lane, the row sum is one shuffle tree per ROW (the generic kernel
// pays one per 128 columns), warps are combined in shared memory in warp order -> deterministic.
constexpr int kF32TileCols = 512;
// T = uint8_t (TR == kByteTable): uint8 maps with a transform other than the identity (or rows the integer kernel
// (P2b) does not take): the same organisation, four pixels per 32-bit load, each pixel's transformed value looked up
// in a 256-entry float64 table the CTA fills first (the generic kernel ran these at 0.04-0.09 of the HBM rate).
template <typename T, int TR>
__global__ void __launch_bounds__(kMargThreads, 2)      // 128 registers: two CTAs per SM (136 registers ran one: 76 us vs 45)
marginals_f32_rows_kernel(const T* __restrict__ att, int H, int W, int rows_per_cta, TransformArgs ta,
                          double* __restrict__ colpart, double* __restrict__ rowpart, int n_row_chunks,
                          int n_col_tiles) {
    constexpr int kWarps = kMargThreads / 32, kSteps = kF32TileCols / 128, U = 2;
    constexpr bool kBytes = std::is_same<T, uint8_t>::value;
    static_assert(kBytes == (TR == kByteTable), "uint8 maps use the table, float32 maps a compiled transform");
    using Vec = typename std::conditional<kBytes, uint32_t, float4>::type;     // four pixels
    extern __shared__ double cwd[];                        // [kWarps][kF32TileCols]
    __shared__ double s_tab[kBytes ? 256 : 1];
    if (kBytes) {
        for (int v = threadIdx.x; v < 256; v += kMargThreads)
            s_tab[v] = transform_fwd((double)v, ta.transform, ta.exp_scale, ta.exp_divisor);
        __syncthreads();
    }
    const int b = blockIdx.z, chunk = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int y0 = chunk * rows_per_cta, y1 = min(y0 + rows_per_cta, H);
    const int xt = tile * kF32TileCols;
    const int tile_cols = min(kF32TileCols, W - xt);
    const T* img = att + (int64_t)b * H * W + xt;
    // float64 instructions are the scarce resource here (a float64 sqrt per pixel ran the kernel at 0.13 of the HBM
    // rate, the clamp + bias + two sums of the identity at 0.36).  For the transforms whose value is a float32 PAIR
    // (identity: x; sqrt: hi + lo to 2^-47, see sqrt_pair; x^2 as an exact pair measured slower than one float64
    // multiplication) the clamp and the transform
    // run in float32, the float64 sums take the converted hi part (one conversion + two adds per pixel) and the lo
    // parts are summed in float32 beside them (they are 2^-24 of the hi parts); the + 1e-9 of every pixel is added once
    // per row / column as count x 1e-9, like the uint8 kernel does.  exp and log keep their float64 evaluation.
    constexpr bool kPair = TR < 0 || TR == T_IDENTITY || TR == T_SQRT;
    constexpr bool kHasLo = TR == T_SQRT;
    double acc[kSteps][4];
    float acc_lo[kHasLo ? kSteps : 1][4];
#pragma unroll
    for (int s = 0; s < kSteps; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[s][k] = 0.0;
            if (kHasLo) acc_lo[s][k] = 0.f;
        }
    double* rp = rowpart + ((int64_t)b * n_col_tiles + tile) * H;
    // double-buffered like (P2b): the next two rows are in flight while these two are summed
    auto fetch = [&](int y, Vec (&v)[U][kSteps]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool row_in = y + u * kWarps < y1;
            const T* r = img + (int64_t)(y + u * kWarps) * W;
#pragma unroll
            for (int s = 0; s < kSteps; ++s) {
                const int x = s * 128 + lane * 4;
                v[u][s] = (row_in && x < tile_cols) ? __ldg(reinterpret_cast<const Vec*>(r + x)) : Vec();
            }
        }
    };
    Vec v[U][kSteps], nx[U][kSteps];
    fetch(y0 + wid, v);
    for (int y = y0 + wid; y < y1; y += U * kWarps) {
        fetch(y + U * kWarps, nx);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool row_in = y + u * kWarps < y1;
            double rs = 0.0;
            float rs_lo = 0.f;
#pragma unroll
            for (int s = 0; s < kSteps; ++s) {
                if (s * 128 >= tile_cols) continue;            // (uniform) a narrow tile's empty steps
                const bool in = row_in && s * 128 + lane * 4 < tile_cols;
                if constexpr (kBytes) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double a = s_tab[(v[u][s] >> (8 * k)) & 0xffu];
                        if (in) { acc[s][k] += a; rs += a; }
                    }
                } else {
                const float f[4] = {v[u][s].x, v[u][s].y, v[u][s].z, v[u][s].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (kPair) {
                        const float x = (f[k] > 0.f || f[k] != f[k]) ? f[k] : 0.f;   // np.maximum(a, 0): NaN stays
                        float hi = x, lo = 0.f;
                        if (TR == T_SQRT) sqrt_pair(x, hi, lo);
                        const double a = (double)hi;
                        if (in) {
                            acc[s][k] += a;
                            rs += a;
                            if (kHasLo) { acc_lo[s][k] += lo; rs_lo += lo; }
                        }
                    } else {
                        const double a = transform_fwd_t<TR>(clamp_nonneg((double)f[k]), ta.exp_scale, ta.exp_divisor);
                        if (in) { acc[s][k] += a; rs += a; }
                    }
                }
                }
            }
            if (kHasLo) rs += (double)rs_lo;
            rs = warp_sum(rs);
            if (lane == 0 && row_in) rp[y + u * kWarps] = TR >= 0 ? rs + (double)tile_cols * kBaseAttention : rs;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int s = 0; s < kSteps; ++s) v[u][s] = nx[u][s];
    }
#pragma unroll
    for (int s = 0; s < kSteps; ++s)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            cwd[wid * kF32TileCols + s * 128 + lane * 4 + k] = kHasLo ? acc[s][k] + (double)acc_lo[s][k] : acc[s][k];
    __syncthreads();
    double* cp = colpart + ((int64_t)b * n_row_chunks + chunk) * W + xt;
    const double col_base = TR >= 0 ? (double)(y1 - y0) * kBaseAttention : 0.0;
    for (int c = threadIdx.x; c < tile_cols; c += kMargThreads) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += cwd[w * kF32TileCols + c];
        cp[c] = t + col_base;
    }
}

// (P3) finish: sum the partials in fixed order, then the shared tail.
__global__ void __launch_bounds__(2 * kProfThreads)
maps_from_partials_kernel(const double* __restrict__ colpart, const double* __restrict__ rowpart,
                          int n_row_chunks, int n_col_tiles, int H, int W, int Wo, int Ho,
                          TransformArgs ta, float* __restrict__ map_x, float* __restrict__ map_y,
                          int* __restrict__ fallback_flags) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    double* red = smem;                       // 2 * kTailRed
    double* xchg = red + 2 * kTailRed;        // kTailXchg
    double* prof_x = xchg + kTailXchg;
    double* prof_y = prof_x + W;
    double* knots = prof_y + H;               // 2 * (max(W,H)+1)
    const double* cp = colpart + (int64_t)b * n_row_chunks * W;
    const double* rp = rowpart + (int64_t)b * n_col_tiles * H;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < n_row_chunks; ++k) s += cp[(int64_t)k * W + x];
        prof_x[x] = s;
    }
    for (int y = threadIdx.x; y < H; y += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < n_col_tiles; ++k) s += rp[(int64_t)k * H + y];
        prof_y[y] = s;
    }
    __syncthreads();
    profiles_to_maps(prof_x, prof_y, W, H, Wo, Ho, ta, knots, red, xchg, map_x + (int64_t)b * Wo,
                     map_y + (int64_t)b * Ho, fallback_flags ? fallback_flags + b : nullptr);
}

// gt_marginals finish (checkpoint_utils.py:43-51): px = mx / max(sum mx, 1e-6), float32 out.
__global__ void __launch_bounds__(kProfThreads)
gt_marginals_finish_kernel(const double* __restrict__ colpart, const double* __restrict__ rowpart,
                           int n_row_chunks, int n_col_tiles, int H, int W,
                           float* __restrict__ px, float* __restrict__ py) {
    __shared__ double red[kProfThreads];
    const int b = blockIdx.x, axis = blockIdx.y;
    const int n = axis == 0 ? W : H;
    const int nparts = axis == 0 ? n_row_chunks : n_col_tiles;
    const double* part = axis == 0 ? colpart + (int64_t)b * n_row_chunks * W
                                   : rowpart + (int64_t)b * n_col_tiles * H;
    float* out = axis == 0 ? px + (int64_t)b * W : py + (int64_t)b * H;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nparts; ++k) s += part[(int64_t)k * n + i];
        acc += (double)(float)s;                      // the reference holds mx in float32
    }
    const float denom = fmaxf((float)block_sum(acc, red), 1e-6f);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nparts; ++k) s += part[(int64_t)k * n + i];
        out[i] = (float)s / denom;
    }
}

// -------------------------------------------------------------------------------------------
// (P4) maps from CDFs (torch path, checkpoint_utils.py:157-189), one CTA per (image, axis).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kProfThreads)
maps_from_cdf_kernel(const float* __restrict__ Fx, const float* __restrict__ Fy, int H, int W,
                     int Wo, int Ho, float* __restrict__ map_x, float* __restrict__ map_y) {
    extern __shared__ double smem[];
    __shared__ int s_tie;
    const int b = blockIdx.x, axis = blockIdx.y;
    const int n = axis == 0 ? W : H, n_out = axis == 0 ? Wo : Ho;
    const float* F = axis == 0 ? Fx + (int64_t)b * W : Fy + (int64_t)b * H;
    float* out = axis == 0 ? map_x + (int64_t)b * Wo : map_y + (int64_t)b * Ho;
    double* knots = smem;  // n + 1
    if (threadIdx.x == 0) s_tie = 0;
    for (int i = threadIdx.x; i <= n; i += blockDim.x) {
        double k;
        if (i == 0) k = 0.0;                                       // concat([0.0], F) * out :171-175
        else if (i == n) k = (double)n_out;                        // forced last knot      :177-178
        else k = dmul_nofma((double)F[i - 1], (double)n_out);
        knots[i] = k;
    }
    __syncthreads();
    int tie = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tie |= (knots[i + 1] - knots[i] <= 0.0) ? 1 : 0;
    if (tie) atomicOr(&s_tie, 1);                                  // np.any(np.diff <= 0)  :181,183
    __syncthreads();
    if (s_tie) {
        // += (1e-4 / max(out,1)) * arange(n+1, dtype=float32): a float32 product    :182,184
        const float step = (float)(1e-4 / (double)max(n_out, 1));
        for (int i = threadIdx.x; i <= n; i += blockDim.x)
            knots[i] = dadd_nofma(knots[i], (double)fmul_nofma(step, (float)i));
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n_out; j += blockDim.x)          // np.interp :188-189, f32 :192-193
        out[j] = (float)interp_index((double)j, knots, n + 1);
}

size_t maps_smem_bytes(int extra_doubles, int H, int W) {
    return sizeof(double) * ((size_t)2 * kTailRed + kTailXchg + extra_doubles + W + H +
                             2 * ((size_t)max(W, H) + 1));
}

template <typename K>
int opt_in_smem(K kern, size_t smem, const char* what) {
    if (smem > 220 * 1024)
        return fail(ATTWARP_ERR_UNSUPPORTED, "%s: image axes too long for shared-memory knots (%zu B)", what, smem);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return ATTWARP_OK;
}

TransformArgs to_args(const attwarp_transform_params& tp) {
    TransformArgs ta;
    ta.transform = tp.transform;
    ta.apply_inverse = tp.apply_inverse;
    ta.exp_scale = tp.exp_scale;
    ta.exp_divisor = tp.exp_divisor;
    return ta;
}

}  // namespace

// Workspace geometry: the finest tiling any marginals kernel uses (64-row chunks, 512-column tiles); every
// kernel reports the chunk / tile counts it actually wrote and the finish kernels sum exactly those.
void marginals_geometry(int H, int W, int* n_row_chunks, int* n_col_tiles) {
    *n_row_chunks = (H + kMargRows - 1) / kMargRows;
    *n_col_tiles = (W + kF32TileCols - 1) / kF32TileCols;
}

int launch_maps_from_tokens(const float* tok, int nsplit, float scale, float* tok_out, int B,
                            int gh, int gw, int H, int W, int Wo, int Ho,
                            const attwarp_transform_params& tp, float* map_x, float* map_y,
                            int* fallback_flags, cudaStream_t st) {
    if (gh * gw > 8192) return fail(ATTWARP_ERR_UNSUPPORTED, "token grid %dx%d too large", gh, gw);
    const size_t smem = maps_smem_bytes(gh * gw + gw + gh, H, W);
    int rc = opt_in_smem(maps_from_tokens_kernel, smem, "maps_from_tokens");
    if (rc != ATTWARP_OK) return rc;
    maps_from_tokens_kernel<<<B, 2 * kProfThreads, smem, st>>>(tok, nsplit, scale, tok_out, gh, gw, H, W,
                                                           Wo, Ho, to_args(tp), map_x, map_y,
                                                           fallback_flags, nullptr);
    return check_launch("maps_from_tokens_kernel");
}

int launch_maps_from_tokens_ragged(const float* tok, int n, int gh, int gw, const RaggedImage* imgs,
                                   int max_h, int max_w, const attwarp_transform_params& tp,
                                   int* fallback_flags, cudaStream_t st) {
    if (gh * gw > 8192) return fail(ATTWARP_ERR_UNSUPPORTED, "token grid %dx%d too large", gh, gw);
    const size_t smem = maps_smem_bytes(gh * gw + gw + gh, max_h, max_w);
    int rc = opt_in_smem(maps_from_tokens_kernel, smem, "maps_from_tokens");
    if (rc != ATTWARP_OK) return rc;
    maps_from_tokens_kernel<<<n, 2 * kProfThreads, smem, st>>>(tok, 1, 1.0f, nullptr, gh, gw, 0, 0, 0, 0,
                                                           to_args(tp), nullptr, nullptr, fallback_flags, imgs);
    return check_launch("maps_from_tokens_kernel");
}

// One pass over the attention maps -> column / row partials.  Picks the kernel, launches it and reports the
// number of row chunks / column tiles it wrote (<= marginals_geometry's, which sizes the workspace).
template <typename T, int TR>
static int launch_marginals_t(const void* att, int B, int H, int W, const TransformArgs& ta,
                              double* colpart, double* rowpart, cudaStream_t st, int* nrc_out, int* nct_out) {
    const bool aligned16 = (reinterpret_cast<uintptr_t>(att) & 15) == 0;
    if constexpr (std::is_same<T, uint8_t>::value) if (ta.transform == T_IDENTITY && (W & 15) == 0 && aligned16) {
        static_assert(kU8MaxRows / (kMargThreads / 32) * 255 < 65536, "16-bit column lanes would overflow");
        const int nct = (W + kU8TileCols - 1) / kU8TileCols;
        const int cols = W < kU8TileCols ? W : kU8TileCols;
        auto kern = cols <= 512 ? marginals_u8_identity_kernel<1, 4>
                  : cols <= 1024 ? marginals_u8_identity_kernel<2, 2> : marginals_u8_identity_kernel<3, 2>;
        static thread_local int occ[3] = {0, 0, 0};
        int& o = occ[cols <= 512 ? 0 : cols <= 1024 ? 1 : 2];
        if (o == 0) {
            AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kMargThreads, 0));
            if (o < 1) o = 1;
        }
        const int rows = rows_per_cta(B, H, nct, o, kU8MaxRows);
        const int nrc = (H + rows - 1) / rows;
        kern<<<dim3(nct, nrc, B), kMargThreads, 0, st>>>(static_cast<const uint8_t*>(att), H, W, rows, colpart,
                                                      rowpart, nrc, nct);
        *nrc_out = nrc; *nct_out = nct;
        return check_launch("marginals_u8_identity_kernel");
    }
    const bool rows_ok = std::is_same<T, float>::value ? (W & 3) == 0 && aligned16
                                                       : (W & 3) == 0 && (reinterpret_cast<uintptr_t>(att) & 3) == 0;
    if constexpr (std::is_same<T, float>::value || std::is_same<T, uint8_t>::value) if (rows_ok) {
        const int nct = (W + kF32TileCols - 1) / kF32TileCols;
        auto kern = marginals_f32_rows_kernel<T, TR>;
        const size_t smem = sizeof(double) * (kMargThreads / 32) * kF32TileCols;
        static thread_local int occ = 0;          // per instantiation = per transform
        if (occ == 0) {
            AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kMargThreads, smem));
            if (occ < 1) occ = 1;
        }
        const int rows = rows_per_cta(B, H, nct, occ, 256);
        const int nrc = (H + rows - 1) / rows;
        kern<<<dim3(nct, nrc, B), kMargThreads, smem, st>>>(static_cast<const T*>(att), H, W, rows, ta, colpart,
                                                         rowpart, nrc, nct);
        *nrc_out = nrc; *nct_out = nct;
        return check_launch("marginals_f32_rows_kernel");
    }
    const int nrc = (H + kMargRows - 1) / kMargRows, nct = (W + kMargCols - 1) / kMargCols;
    marginals_partial_kernel<T, TR><<<dim3(nct, nrc, B), kMargThreads, 0, st>>>(
        static_cast<const T*>(att), H, W, ta, colpart, rowpart, nrc, nct);
    *nrc_out = nrc; *nct_out = nct;
    return check_launch("marginals_partial_kernel");
}

// MODE 0: NumPy path (the run-time transform picks the instantiation; unknown ids are the identity, like
// transform_fwd's default).  MODE 1: gt_marginals (clamp only).
template <typename T, int MODE>
static int launch_marginals(const void* att, int B, int H, int W, const TransformArgs& ta,
                            double* colpart, double* rowpart, cudaStream_t st, int* nrc_out, int* nct_out) {
    if constexpr (MODE == 1) {
        return launch_marginals_t<T, kClampOnly>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
    } else if constexpr (std::is_same<T, uint8_t>::value) {
        return launch_marginals_t<T, kByteTable>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
    } else switch (ta.transform) {
        case T_SQUARE: return launch_marginals_t<T, T_SQUARE>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
        case T_SQRT: return launch_marginals_t<T, T_SQRT>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
        case T_EXP: return launch_marginals_t<T, T_EXP>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
        case T_LOG: return launch_marginals_t<T, T_LOG>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
        default: return launch_marginals_t<T, T_IDENTITY>(att, B, H, W, ta, colpart, rowpart, st, nrc_out, nct_out);
    }
}

size_t maps_workspace_bytes_impl(int B, int H, int W) {
    int nrc, nct;
    marginals_geometry(H, W, &nrc, &nct);
    return sizeof(double) * (size_t)B * ((size_t)nrc * W + (size_t)nct * H);
}

int launch_maps_from_attention(const void* att, int att_dtype, int B, int H, int W, int Wo, int Ho,
                               const attwarp_transform_params& tp, void* ws, size_t ws_bytes,
                               float* map_x, float* map_y, int* fallback_flags, cudaStream_t st) {
    if (ws == nullptr || ws_bytes < maps_workspace_bytes_impl(B, H, W))
        return fail(ATTWARP_ERR_WORKSPACE, "maps_from_attention: workspace too small (%zu < %zu)",
                    ws_bytes, maps_workspace_bytes_impl(B, H, W));
    int nrc, nct;
    marginals_geometry(H, W, &nrc, &nct);
    double* colpart = static_cast<double*>(ws);
    double* rowpart = colpart + (size_t)B * nrc * W;
    const TransformArgs ta = to_args(tp);
    int rc;
    switch (att_dtype) {
        case ATTWARP_U8: rc = launch_marginals<uint8_t, 0>(att, B, H, W, ta, colpart, rowpart, st, &nrc, &nct); break;
        case ATTWARP_F32: rc = launch_marginals<float, 0>(att, B, H, W, ta, colpart, rowpart, st, &nrc, &nct); break;
        case ATTWARP_F64: rc = launch_marginals<double, 0>(att, B, H, W, ta, colpart, rowpart, st, &nrc, &nct); break;
        default: return fail(ATTWARP_ERR_INVALID_ARG, "attention map dtype must be u8/f32/f64 (got %d)", att_dtype);
    }
    if (rc != ATTWARP_OK) return rc;
    const size_t smem = maps_smem_bytes(0, H, W);
    rc = opt_in_smem(maps_from_partials_kernel, smem, "maps_from_attention");
    if (rc != ATTWARP_OK) return rc;
    maps_from_partials_kernel<<<B, 2 * kProfThreads, smem, st>>>(colpart, rowpart, nrc, nct, H, W, Wo, Ho,
                                                             ta, map_x, map_y, fallback_flags);
    return check_launch("maps_from_partials_kernel");
}

// (P2d) maps from the token-grid mask of the driver flow, its LANCZOS resize to image size never written
// (mask.cu, resize_lanczos_up_kernel<.., MARG>), identity transform only (sums of bytes stay exact integers).
int lanczos_marginals_chunks(int B, int h, int w, int H, int W, int* n_col_tiles);
int launch_lanczos_marginals(const uint8_t* src, int B, int h, int w, int H, int W, double* colpart, double* rowpart,
                             int* n_chunks, int* n_col_tiles, cudaStream_t st);
size_t maps_from_mask_workspace_bytes_impl(int B, int h, int w, int H, int W) {
    int nct = 1;
    const int chunks = lanczos_marginals_chunks(B, h, w, H, W, &nct);
    return chunks == 0 ? 0 : sizeof(double) * (size_t)B * ((size_t)chunks * W + (size_t)nct * H);
}
int launch_maps_from_mask(const uint8_t* mask, int B, int h, int w, int H, int W, int Wo, int Ho,
                          const attwarp_transform_params& tp, void* ws, size_t ws_bytes, float* map_x, float* map_y,
                          cudaStream_t st) {
    if (tp.transform != ATTWARP_T_IDENTITY)
        return fail(ATTWARP_ERR_UNSUPPORTED, "maps_from_mask: identity transform only (resize the mask and use maps_from_attention)");
    const size_t need = maps_from_mask_workspace_bytes_impl(B, h, w, H, W);
    if (need == 0) return fail(ATTWARP_ERR_UNSUPPORTED, "maps_from_mask: %dx%d -> %dx%d is not an up-scaling the fused kernel takes", h, w, H, W);
    if (ws == nullptr || ws_bytes < need) return fail(ATTWARP_ERR_WORKSPACE, "maps_from_mask: workspace too small (%zu < %zu)", ws_bytes, need);
    int nct = 1;
    int chunks = lanczos_marginals_chunks(B, h, w, H, W, &nct);
    double* colpart = static_cast<double*>(ws);
    double* rowpart = colpart + (size_t)B * chunks * W;
    int rc = launch_lanczos_marginals(mask, B, h, w, H, W, colpart, rowpart, &chunks, &nct, st);
    if (rc != ATTWARP_OK) return rc;
    const TransformArgs ta = to_args(tp);
    const size_t smem = maps_smem_bytes(0, H, W);
    rc = opt_in_smem(maps_from_partials_kernel, smem, "maps_from_mask");
    if (rc != ATTWARP_OK) return rc;
    maps_from_partials_kernel<<<B, 2 * kProfThreads, smem, st>>>(colpart, rowpart, chunks, nct, H, W, Wo, Ho, ta, map_x,
                                                             map_y, nullptr);
    return check_launch("maps_from_partials_kernel");
}

int launch_gt_marginals(const float* A, int B, int H, int W, void* ws, size_t ws_bytes, float* px,
                        float* py, cudaStream_t st) {
    if (ws == nullptr || ws_bytes < maps_workspace_bytes_impl(B, H, W))
        return fail(ATTWARP_ERR_WORKSPACE, "gt_marginals: workspace too small");
    int nrc, nct;
    marginals_geometry(H, W, &nrc, &nct);
    double* colpart = static_cast<double*>(ws);
    double* rowpart = colpart + (size_t)B * nrc * W;
    TransformArgs ta = {0, 0, 1.0, 1.0};
    int rc = launch_marginals<float, 1>(A, B, H, W, ta, colpart, rowpart, st, &nrc, &nct);
    if (rc != ATTWARP_OK) return rc;
    gt_marginals_finish_kernel<<<dim3(B, 2), kProfThreads, 0, st>>>(colpart, rowpart, nrc, nct, H, W, px, py);
    return check_launch("gt_marginals_finish_kernel");
}

int launch_maps_from_cdf(const float* Fx, const float* Fy, int B, int H, int W, int Wo, int Ho,
                         float* map_x, float* map_y, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)max(W, H) + 1);
    int rc = opt_in_smem(maps_from_cdf_kernel, smem, "maps_from_cdf");
    if (rc != ATTWARP_OK) return rc;
    maps_from_cdf_kernel<<<dim3(B, 2), kProfThreads, smem, st>>>(Fx, Fy, H, W, Wo, Ho, map_x, map_y);
    return check_launch("maps_from_cdf_kernel");
}

}  // namespace aw
