// torch_path.cu -- stages 2-3 of the torch path: small row-wise float32 kernels.
//
// Replaces (file:line under "model/marginalnet_full_dataset/"):
//   safe_softmax                model.py:8-14
//   mix_with_uniform            model.py:98-101
//   cdf_from_density            checkpoint_utils.py:30-41
//   upsample_pdf_right_inverse  checkpoint_utils.py:64-131  (as y * M^T, M precomputed)
//   F.adaptive_avg_pool2d       trainer.py:197
// These move O(B*N) bytes (a few hundred KB at the benchmark sizes); they are latency-bound, one
// CTA (or warp) per row, and exist so the path never leaves the device.
#include "common.cuh"

namespace aw {
namespace {

constexpr int kRowThreads = 128;

__device__ __forceinline__ float nan_inf_to_zero(float v) { return (isnan(v) || isinf(v)) ? 0.f : v; }

// One CTA per row.  safe_softmax: nan_to_num(0,0,0) -> subtract max -> softmax -> nan_to_num ->
// / max(sum, eps).
__global__ void __launch_bounds__(kRowThreads)
safe_softmax_kernel(const float* __restrict__ logits, int N, float eps, int mix, float c1, float c2,
                    float* __restrict__ out) {
    __shared__ float red[32];
    const float* row = logits + (int64_t)blockIdx.x * N;
    float* o = out + (int64_t)blockIdx.x * N;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, nan_inf_to_zero(row[i]));
    m = block_max(m, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += expf(nan_inf_to_zero(row[i]) - m);
    s = block_sum(s, red);
    float s2 = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float p = nan_inf_to_zero(expf(nan_inf_to_zero(row[i]) - m) / s);
        o[i] = p;
        s2 += p;
    }
    s2 = block_sum(s2, red);
    const float denom = fmaxf(s2, eps);
    // mix != 0: mix_with_uniform fused behind it (model.py:98-101), the same two roundings as the stand-alone kernel
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float q = o[i] / denom;
        o[i] = mix ? fadd_nofma(fmul_nofma(c1, q), c2) : q;
    }
}

__global__ void mix_with_uniform_kernel(const float* __restrict__ p, int64_t total, float c1,
                                        float c2, int copy_only, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    out[i] = copy_only ? p[i] : fadd_nofma(fmul_nofma(c1, p[i]), c2);
}

// cdf_from_density: clamp_min(0) (NaN survives), nan_to_num(0,0,0), / max(sum,1e-6), cumsum with
// float64 accumulation rounded to float32 per element (what torch.cumsum does on CPU), last = 1.
__global__ void __launch_bounds__(kRowThreads)
cdf_from_density_kernel(const float* __restrict__ p, int N, float* __restrict__ F) {
    extern __shared__ double sm[];
    double* red = sm;              // kRowThreads
    double* a = sm + kRowThreads;  // N
    const float* row = p + (int64_t)blockIdx.x * N;
    float* o = F + (int64_t)blockIdx.x * N;
    double part = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float v = row[i];
        v = isnan(v) ? v : fmaxf(v, 0.f);
        v = nan_inf_to_zero(v);
        a[i] = (double)v;
        part += (double)v;
    }
    const float denom = fmaxf((float)block_sum(part, red), 1e-6f);
    for (int i = threadIdx.x; i < N; i += blockDim.x) a[i] = (double)((float)a[i] / denom);
    __syncthreads();
    block_inclusive_scan(a, N, red);
    for (int i = threadIdx.x; i < N; i += blockDim.x) o[i] = (i == N - 1) ? 1.0f : (float)a[i];
}

// x[b][i] = sum_k M[i][k] * y[b][k]
__global__ void upsample_right_inverse_kernel(const float* __restrict__ y,
                                              const float* __restrict__ M, int L_out, int L_in,
                                              float* __restrict__ x) {
    extern __shared__ float ys[];
    const int b = blockIdx.y;
    for (int k = threadIdx.x; k < L_out; k += blockDim.x) ys[k] = y[(int64_t)b * L_out + k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L_in) return;
    const float* m = M + (int64_t)i * L_out;
    float acc = 0.f;
    for (int k = 0; k < L_out; ++k) acc = fmaf(__ldg(m + k), ys[k], acc);
    x[(int64_t)b * L_in + i] = acc;
}

// Fused PDF -> CDF for both axes of a batch (BASELINE configs[4]; trainer.py:212-218, 285-288):
//   mix_with_uniform -> upsample_pdf_right_inverse (x = M y) -> .clamp_min(0) -> cdf_from_density
// in one launch, grid (B, 2): blockIdx.y = 0 is the x axis (px -> Fx, length W), 1 the y axis.
// Each step is the arithmetic of the stand-alone kernels above, in the same order, so the result is
// bit-identical to running them one after the other; the upsampled PDF never leaves shared memory.
__global__ void __launch_bounds__(kRowThreads)
pdf_to_cdf_kernel(const float* __restrict__ px, const float* __restrict__ py, int Nx, int Ny, float c1x,
                  float c2x, float c1y, float c2y, int mix, const float* __restrict__ Mx,
                  const float* __restrict__ My, int W, int H, float* __restrict__ Fx, float* __restrict__ Fy) {
    extern __shared__ double sm[];
    const int axis = blockIdx.y;
    const int N = axis ? Ny : Nx, L = axis ? H : W;
    const float* y = (axis ? py : px) + (int64_t)blockIdx.x * N;
    const float* M = axis ? My : Mx;
    float* o = (axis ? Fy : Fx) + (int64_t)blockIdx.x * L;
    const float c1 = axis ? c1y : c1x, c2 = axis ? c2y : c2x;
    double* red = sm;                                  // kRowThreads
    double* a = sm + kRowThreads;                      // L
    float* ys = reinterpret_cast<float*>(a + L);       // N
    for (int k = threadIdx.x; k < N; k += blockDim.x)
        ys[k] = mix ? fadd_nofma(fmul_nofma(c1, y[k]), c2) : y[k];
    __syncthreads();
    double part = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float* m = M + (int64_t)i * N;
        float acc = 0.f;
        for (int k = 0; k < N; ++k) acc = fmaf(__ldg(m + k), ys[k], acc);
        float v = isnan(acc) ? acc : fmaxf(acc, 0.f);  // .clamp_min(0) (idempotent with the CDF's own clamp)
        v = nan_inf_to_zero(v);
        a[i] = (double)v;
        part += (double)v;
    }
    const float denom = fmaxf((float)block_sum(part, red), 1e-6f);
    for (int i = threadIdx.x; i < L; i += blockDim.x) a[i] = (double)((float)a[i] / denom);
    __syncthreads();
    block_inclusive_scan(a, L, red);
    for (int i = threadIdx.x; i < L; i += blockDim.x) o[i] = (i == L - 1) ? 1.0f : (float)a[i];
}

// _make_strictly_increasing (checkpoint_utils.py:17-28), one CTA per row:
//   nan_to_num(nan 0, +inf 1, -inf 0) -> cummax -> steps clamped to >= eps/N -> re-accumulated from the
//   first value (torch.cumsum semantics as in cdf_from_density) -> / max(last, 1e-6) -> clip [0,1] -> last = 1.
__global__ void __launch_bounds__(kRowThreads)
strictly_increasing_kernel(const float* __restrict__ F, int N, float min_step, float* __restrict__ out) {
    extern __shared__ double sm[];
    double* red = sm;                                   // kRowThreads
    double* a = sm + kRowThreads;                       // N
    float* nd = reinterpret_cast<float*>(a + N);        // N   running maximum
    __shared__ float chunk_max[kRowThreads];
    const float* row = F + (int64_t)blockIdx.x * N;
    float* o = out + (int64_t)blockIdx.x * N;
    const int per = (N + blockDim.x - 1) / blockDim.x;
    const int beg = min((int)threadIdx.x * per, N), end = min(beg + per, N);
    float m = -INFINITY;
    for (int i = beg; i < end; ++i) {                   // cummax inside the thread's chunk
        float v = row[i];
        v = isnan(v) ? 0.f : (isinf(v) ? (v > 0.f ? 1.f : 0.f) : v);
        m = fmaxf(m, v);
        nd[i] = m;
    }
    chunk_max[threadIdx.x] = m;
    __syncthreads();
    float before = -INFINITY;                           // maximum of all earlier chunks
    for (int t = 0; t < (int)threadIdx.x; ++t) before = fmaxf(before, chunk_max[t]);
    for (int i = beg; i < end; ++i) nd[i] = fmaxf(nd[i], before);
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x)   // a[i] = clamped step into element i (a[0] = 0)
        a[i] = i == 0 ? 0.0 : (double)fmaxf(fadd_nofma(nd[i], -nd[i - 1]), min_step);
    __syncthreads();
    block_inclusive_scan(a, N, red);
    const float first = nd[0];
    const float last = fmaxf(N > 1 ? fadd_nofma(first, (float)a[N - 1]) : first, 1e-6f);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float v = i == 0 ? first : fadd_nofma(first, (float)a[i]);
        o[i] = (i == N - 1) ? 1.0f : fminf(fmaxf(v / last, 0.f), 1.f);
    }
}

// F.interpolate(mode='linear', align_corners=True) on rows (checkpoint_utils.py:57-59).
__global__ void interp_linear_rows_kernel(const float* __restrict__ F, int N, int L, float scale,
                                          float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    const float* row = F + (int64_t)blockIdx.y * N;
    const float pos = fmul_nofma(scale, (float)j);
    const int i0 = min((int)pos, N - 1), i1 = min(i0 + 1, N - 1);
    const float lam1 = fadd_nofma(pos, -(float)i0), lam0 = fadd_nofma(1.0f, -lam1);
    out[(int64_t)blockIdx.y * L + j] = fadd_nofma(fmul_nofma(lam0, row[i0]), fmul_nofma(lam1, row[i1]));
}

// adaptive_avg_pool2d (trainer.py:197,433,465): window [floor(i*H/gh), ceil((i+1)*H/gh)) per axis.
// One CTA per (image, output row): the rows of the window are read once, coalesced (float4 per lane, four
// rows in flight), into float64 column sums; the gw window sums of the row are then taken from shared
// memory.  Rows shared by two windows (H % gh != 0) are read by two CTAs: + gh / H of traffic.
// MODE 0: plain pooling.  MODE 1: the trainer's prologue fused in (trainer.py:186-194): every value is
// clamp_min(0)-ed and, for samples whose sqrt_mask byte is set, replaced by its float32 square root
// (A_pos.sqrt() * m + A_pos * (1 - m) with m in {0, 1}) before it is pooled.
template <int MODE>
__device__ __forceinline__ float pool_value(float v, bool take_sqrt) {
    if (MODE == 0) return v;
    v = (v != v) ? v : fmaxf(v, 0.f);                  // clamp_min keeps NaN
    return take_sqrt ? sqrtf(v) : v;
}
template <int MODE>
__global__ void __launch_bounds__(128)
adaptive_avg_pool2d_kernel(const float* __restrict__ A, const uint8_t* __restrict__ sqrt_mask, int n_img, int H, int W,
                           int gh, int gw, float* __restrict__ out) {
    extern __shared__ double cs[];                         // [W] column sums over the window's rows
    // (image, output row) pairs are dealt round-robin to a grid sized for whole rounds: every CTA takes the same number
    for (int pair = blockIdx.x; pair < gh * n_img; pair += gridDim.x) {
    const int i = pair % gh, b = pair / gh;
    const int y0 = (int)(((int64_t)i * H) / gh), y1 = (int)(((int64_t)(i + 1) * H + gh - 1) / gh);
    const float* img = A + (int64_t)b * H * W;
    const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    const bool sq = MODE == 1 && sqrt_mask != nullptr && sqrt_mask[b] != 0;
    if (vec) {
        for (int x = threadIdx.x * 4; x < W; x += blockDim.x * 4) {
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            for (int y = y0; y < y1; y += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = y + u < y1 ? __ldg(reinterpret_cast<const float4*>(img + (int64_t)(y + u) * W + x))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (MODE == 1 && !(y + u < y1)) continue;          // (MODE 0 adds the zeros it loaded)
                    a[0] += (double)pool_value<MODE>(v[u].x, sq); a[1] += (double)pool_value<MODE>(v[u].y, sq);
                    a[2] += (double)pool_value<MODE>(v[u].z, sq); a[3] += (double)pool_value<MODE>(v[u].w, sq);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) cs[x + k] = a[k];
        }
    } else {
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            double a = 0.0;
            for (int y = y0; y < y1; ++y) a += (double)pool_value<MODE>(__ldg(img + (int64_t)y * W + x), sq);
            cs[x] = a;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < gw; j += blockDim.x) {
        const int x0 = (int)(((int64_t)j * W) / gw), x1 = (int)(((int64_t)(j + 1) * W + gw - 1) / gw);
        double t = 0.0;
        for (int x = x0; x < x1; ++x) t += cs[x];
        out[((int64_t)b * gh + i) * gw + j] = (float)(t / (double)((y1 - y0) * (x1 - x0)));
    }
    __syncthreads();                                       // cs is reused by the next pair
    }
}

// ---- backward passes of the three helpers the reference differentiates through (trainer.py:209-250:
// net -> safe_softmax -> mix_with_uniform -> upsample_pdf_right_inverse -> clamp/normalise -> L1) -------------
// safe_softmax backward, one CTA per row.  Forward: z' = nan_to_num(z); p = softmax(z'); s = sum p;
// q = p / max(s, eps).  Backward: g' = s >= eps ? (g - sum g q) / s : g / eps;  dz = p (g' - sum g' p);
// dz = 0 where z is not finite (nan_to_num has zero slope there).
__global__ void __launch_bounds__(kRowThreads)
safe_softmax_backward_kernel(const float* __restrict__ logits, const float* __restrict__ grad_out, int N, float eps,
                             float gscale, float* __restrict__ grad_logits) {
    __shared__ float red[32];
    const float* row = logits + (int64_t)blockIdx.x * N;
    const float* g = grad_out + (int64_t)blockIdx.x * N;
    float* o = grad_logits + (int64_t)blockIdx.x * N;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) m = fmaxf(m, nan_inf_to_zero(row[i]));
    m = block_max(m, red);
    float e = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) e += expf(nan_inf_to_zero(row[i]) - m);
    e = block_sum(e, red);
    float s = 0.f, gp = 0.f;                            // s = sum p, gp = sum g p
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float p = nan_inf_to_zero(expf(nan_inf_to_zero(row[i]) - m) / e);
        s += p;
        gp += g[i] * gscale * p;
    }
    s = block_sum(s, red);
    gp = block_sum(gp, red);
    const bool live = s >= eps;
    const float denom = fmaxf(s, eps);
    // g'_i = live ? (g_i - gp / denom) / denom : g_i / denom   (q = p / denom)
    const float shift = live ? gp / denom : 0.f;
    float gpp = 0.f;                                    // sum g' p
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float p = nan_inf_to_zero(expf(nan_inf_to_zero(row[i]) - m) / e);
        gpp += (g[i] * gscale - shift) / denom * p;
    }
    gpp = block_sum(gpp, red);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float z = row[i];
        const float p = nan_inf_to_zero(expf(nan_inf_to_zero(z) - m) / e);
        const float gi = (g[i] * gscale - shift) / denom;
        o[i] = (isnan(z) || isinf(z)) ? 0.f : p * (gi - gpp);
    }
}

// upsample_pdf_right_inverse backward: x = y M^T  =>  grad_y[b][k] = sum_i grad_x[b][i] M[i][k].
// One CTA per row b; a warp owns output bins k, k + warps, ... and its lanes stride over i.
__global__ void __launch_bounds__(kRowThreads)
upsample_right_inverse_backward_kernel(const float* __restrict__ gx, const float* __restrict__ M, int L_out, int L_in,
                                       float* __restrict__ gy) {
    extern __shared__ float gs[];                       // grad_x row
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < L_in; i += blockDim.x) gs[i] = gx[(int64_t)b * L_in + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = wid; k < L_out; k += nw) {
        float acc = 0.f;
        for (int i = lane; i < L_in; i += 32) acc = fmaf(gs[i], __ldg(M + (int64_t)i * L_out + k), acc);
        acc = warp_sum(acc);
        if (lane == 0) gy[(int64_t)b * L_out + k] = acc;
    }
}

}  // namespace

int launch_safe_softmax(const float* logits, int B, int N, float eps, float* out, cudaStream_t st) {
    safe_softmax_kernel<<<B, kRowThreads, 0, st>>>(logits, N, eps, 0, 1.f, 0.f, out);
    return check_launch("safe_softmax_kernel");
}

// safe_softmax followed by mix_with_uniform in one launch (MarginalNet.forward's last op + trainer.py:212-214)
int launch_safe_softmax_mix(const float* logits, int B, int N, float eps, float alpha, float* out, cudaStream_t st) {
    const float c1 = (float)(1.0 - (double)alpha), c2 = (float)((double)alpha / (double)N);
    safe_softmax_kernel<<<B, kRowThreads, 0, st>>>(logits, N, eps, alpha > 0.f, c1, c2, out);
    return check_launch("safe_softmax_kernel");
}

int launch_mix_with_uniform(const float* p, int B, int N, float alpha, float* out, cudaStream_t st) {
    const int64_t total = (int64_t)B * N;
    // (1 - alpha) and alpha / N are Python floats (float64) cast to the tensor dtype by torch
    const float c1 = (float)(1.0 - (double)alpha), c2 = (float)((double)alpha / (double)N);
    mix_with_uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p, total, c1, c2,
                                                                             alpha <= 0.f, out);
    return check_launch("mix_with_uniform_kernel");
}

// d/dp [(1 - alpha) p + alpha / N] = (1 - alpha); alpha <= 0 is the identity
int launch_mix_with_uniform_backward(const float* g, int B, int N, float alpha, float* gp, cudaStream_t st) {
    const int64_t total = (int64_t)B * N;
    const float c1 = (float)(1.0 - (double)alpha);
    mix_with_uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, total, c1, 0.f, alpha <= 0.f, gp);
    return check_launch("mix_with_uniform_kernel");
}

int launch_cdf_from_density(const float* p, int B, int N, float* F, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)kRowThreads + N);
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "cdf_from_density: N=%d too long", N);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(cdf_from_density_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cdf_from_density_kernel<<<B, kRowThreads, smem, st>>>(p, N, F);
    return check_launch("cdf_from_density_kernel");
}

int launch_pdf_to_cdf(const float* px, const float* py, int B, int Nx, int Ny, float alpha, const float* Mx,
                      const float* My, int W, int H, float* Fx, float* Fy, cudaStream_t st) {
    const int Lmax = W > H ? W : H, Nmax = Nx > Ny ? Nx : Ny;
    const size_t smem = sizeof(double) * ((size_t)kRowThreads + Lmax) + sizeof(float) * (size_t)Nmax;
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "pdf_to_cdf: length %d too long", Lmax);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(pdf_to_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (1 - alpha) and alpha / N are Python floats (float64) cast to the tensor dtype by torch
    const float c1 = (float)(1.0 - (double)alpha);
    pdf_to_cdf_kernel<<<dim3(B, 2), kRowThreads, smem, st>>>(
        px, py, Nx, Ny, c1, (float)((double)alpha / (double)Nx), c1, (float)((double)alpha / (double)Ny),
        alpha > 0.f ? 1 : 0, Mx, My, W, H, Fx, Fy);
    return check_launch("pdf_to_cdf_kernel");
}

int launch_strictly_increasing(const float* F, int B, int N, float eps, float* out, cudaStream_t st) {
    const size_t smem = sizeof(double) * ((size_t)kRowThreads + N) + sizeof(float) * (size_t)N;
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "make_strictly_increasing: N=%d too long", N);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(strictly_increasing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float min_step = (float)((double)eps / (double)(N > 1 ? N : 1));
    strictly_increasing_kernel<<<B, kRowThreads, smem, st>>>(F, N, min_step, out);
    return check_launch("strictly_increasing_kernel");
}

int launch_interp_linear_rows(const float* F, int B, int N, int L, float* out, cudaStream_t st) {
    const float scale = L > 1 ? (float)(N - 1) / (float)(L - 1) : 0.f;
    interp_linear_rows_kernel<<<dim3((L + 127) / 128, B), 128, 0, st>>>(F, N, L, scale, out);
    return check_launch("interp_linear_rows_kernel");
}

int launch_upsample_right_inverse(const float* y, const float* M, int B, int L_out, int L_in,
                                  float* x, cudaStream_t st) {
    upsample_right_inverse_kernel<<<dim3((L_in + 127) / 128, B), 128, sizeof(float) * L_out, st>>>(
        y, M, L_out, L_in, x);
    return check_launch("upsample_right_inverse_kernel");
}

int launch_safe_softmax_backward(const float* logits, const float* grad_out, int B, int N, float eps,
                                 float* grad_logits, cudaStream_t st) {
    safe_softmax_backward_kernel<<<B, kRowThreads, 0, st>>>(logits, grad_out, N, eps, 1.f, grad_logits);
    return check_launch("safe_softmax_backward_kernel");
}

// backward of the fused pair: d mix / d q = (1 - alpha) (alpha <= 0: identity), then the softmax backward
int launch_safe_softmax_mix_backward(const float* logits, const float* grad_out, int B, int N, float eps, float alpha,
                                     float* grad_logits, cudaStream_t st) {
    const float c1 = alpha > 0.f ? (float)(1.0 - (double)alpha) : 1.f;
    safe_softmax_backward_kernel<<<B, kRowThreads, 0, st>>>(logits, grad_out, N, eps, c1, grad_logits);
    return check_launch("safe_softmax_backward_kernel");
}

int launch_upsample_right_inverse_backward(const float* gx, const float* M, int B, int L_out, int L_in, float* gy,
                                           cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)L_in;
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "upsample_right_inverse_backward: L_in=%d too long", L_in);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(upsample_right_inverse_backward_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upsample_right_inverse_backward_kernel<<<B, kRowThreads, smem, st>>>(gx, M, L_out, L_in, gy);
    return check_launch("upsample_right_inverse_backward_kernel");
}

// sqrt_mask == nullptr: plain F.adaptive_avg_pool2d; else [B] bytes, the trainer's clamp + per-sample sqrt first
int launch_adaptive_avg_pool2d(const float* A, const uint8_t* sqrt_mask, int B, int H, int W, int gh, int gw,
                               float* out, cudaStream_t st) {
    if (gh > 65535 || B > 65535) return fail(ATTWARP_ERR_UNSUPPORTED, "adaptive_avg_pool2d: grid too large");
    const size_t smem = sizeof(double) * (size_t)W;
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "adaptive_avg_pool2d: W=%d too wide", W);
    auto kern = sqrt_mask ? adaptive_avg_pool2d_kernel<1> : adaptive_avg_pool2d_kernel<0>;
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one resident wave of CTAs, each with the same number of (image, output row) pairs where that divides evenly
    // (3072 pairs over 2368 slots would run a full round and a 30 % one: 0.63 of the HBM peak; 1536 CTAs x 2: see
    // profiles/r03*_row_kernels.txt)
    static thread_local size_t occ_smem[2] = {~(size_t)0, ~(size_t)0};
    static thread_local int occ_val[2] = {0, 0};
    const int v = sqrt_mask ? 1 : 0;
    if (occ_smem[v] != smem) {
        AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_val[v], kern, 128, smem));
        occ_smem[v] = smem;
    }
    const int occ = occ_val[v];
    const int64_t slots = (int64_t)(occ < 1 ? 1 : occ) * sm_count(), pairs = (int64_t)gh * B;
    const int64_t rounds = (pairs + slots - 1) / slots;
    const int grid = (int)((pairs + rounds - 1) / rounds);
    kern<<<grid, 128, smem, st>>>(A, sqrt_mask, B, H, W, gh, gw, out);
    return check_launch("adaptive_avg_pool2d_kernel");
}

}  // namespace aw
