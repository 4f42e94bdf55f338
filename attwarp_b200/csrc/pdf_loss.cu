// pdf_loss.cu -- the image-resolution PDF-L1 loss the reference trains MarginalNet with, forward and backward in
// one launch each ("model/marginalnet_full_dataset/trainer.py":217-250):
//
//   p_img = upsample_pdf_right_inverse(p_s, L).clamp_min(0);  p_img /= p_img.sum(1, keepdim).clamp_min(1e-6)
//   g_img = upsample_pdf_right_inverse(g,   L).clamp_min(0);  g_img /= g_img.sum(1, keepdim).clamp_min(1e-6)
//   L_pdf = F.l1_loss(px_img, gx_img) + F.l1_loss(py_img, gy_img)                      (mean over B * L each)
//
// The reference materialises four [B, L] tensors per axis (eight launches forward, autograd's chain backward); here
// one CTA per (sample, axis) keeps the up-sampled rows in shared memory.  Forward: per-row sums of |p_img - g_img|,
// the last CTA to finish adds them in row order (deterministic) and writes the scalar.  Backward: the rows are
// recomputed (24 multiply-adds per bin: cheaper than storing them) and
//   d L / d p_img_i = upstream * sign(p_img_i - g_img_i) / (B * L)
//   d L / d u_i     = (s_i - [sum > 1e-6] * sum_j s_j p_img_j) / max(sum, 1e-6)     (u = clamped up-sampled row)
//   d L / d v_i     = d L / d u_i where v_i >= 0 (torch's clamp_min passes the gradient at equality)
//   d L / d p_s_k   = sum_i d L / d v_i * M[i][k]
// The ground-truth side carries no gradient (gt_marginals of the data).
#include "common.cuh"

namespace aw {
namespace {

constexpr int kLossThreads = 256;

// up-sampled, clamped row in `row` (shared, L floats); returns its float64 sum to every thread
__device__ __forceinline__ double upsample_clamped(const float* __restrict__ y, int N, const float* __restrict__ M, int L,
                                                   float* ys, float* row, double* red) {
    for (int k = threadIdx.x; k < N; k += blockDim.x) ys[k] = y[k];
    __syncthreads();
    double part = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float* m = M + (int64_t)i * N;
        float acc = 0.f;
        for (int k = 0; k < N; ++k) acc = fmaf(__ldg(m + k), ys[k], acc);      // == upsample_right_inverse_kernel
        const float c = acc > 0.f ? acc : 0.f;                                   // clamp_min(0): NaN -> NaN like torch
        row[i] = (acc != acc) ? acc : c;
        part += (double)row[i];
    }
    return block_sum(part, red);
}

// grid (B, 2): blockIdx.y = 0 the x axis (length W), 1 the y axis (length H)
__global__ void __launch_bounds__(kLossThreads)
pdf_l1_loss_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ gx,
                   const float* __restrict__ gy, int Nx, int Ny, int Ngx, int Ngy, const float* __restrict__ Mx,
                   const float* __restrict__ My, const float* __restrict__ Mgx, const float* __restrict__ Mgy, int W,
                   int H, double* __restrict__ partial, unsigned* __restrict__ counter, float* __restrict__ loss) {
    extern __shared__ double sm[];
    const int axis = blockIdx.y, b = blockIdx.x, B = gridDim.x;
    const int N = axis ? Ny : Nx, Ng = axis ? Ngy : Ngx, L = axis ? H : W;
    double* red = sm;                                         // 32
    float* prow = reinterpret_cast<float*>(sm + 32);          // L
    float* grow = prow + L;                                   // L
    float* ys = grow + L;                                     // max(N, Ng)
    const double sp = upsample_clamped((axis ? py : px) + (int64_t)b * N, N, axis ? My : Mx, L, ys, prow, red);
    __syncthreads();
    const double sg = upsample_clamped((axis ? gy : gx) + (int64_t)b * Ng, Ng, axis ? Mgy : Mgx, L, ys, grow, red);
    // torch divides float32 rows by the float32 sum clamped at 1e-6
    const float dp = fmaxf((float)sp, 1e-6f), dg = fmaxf((float)sg, 1e-6f);
    double part = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) part += (double)fabsf(prow[i] / dp - grow[i] / dg);
    const double tot = block_sum(part, red);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[axis * B + b] = tot;
        __threadfence();
        last = atomicAdd(counter, 1u) == 2u * B - 1u;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA: row sums added in row order, one warp per axis
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid < 2) {
        double acc = 0.0;
        for (int r = lane; r < B; r += 32) acc += __ldcg(partial + wid * B + r);
        acc = warp_sum(acc);
        if (lane == 0) red[wid] = acc / ((double)B * (double)(wid ? H : W));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        loss[0] = (float)red[0] + (float)red[1];
        *counter = 0u;                                        // ready for the next launch
    }
}

__global__ void __launch_bounds__(kLossThreads)
pdf_l1_loss_backward_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ gx,
                            const float* __restrict__ gy, int Nx, int Ny, int Ngx, int Ngy,
                            const float* __restrict__ Mx, const float* __restrict__ My, const float* __restrict__ Mgx,
                            const float* __restrict__ Mgy, int W, int H, const float* __restrict__ upstream,
                            float* __restrict__ dpx, float* __restrict__ dpy) {
    extern __shared__ double sm[];
    const int axis = blockIdx.y, b = blockIdx.x, B = gridDim.x;
    const int N = axis ? Ny : Nx, Ng = axis ? Ngy : Ngx, L = axis ? H : W;
    const float* M = axis ? My : Mx;
    double* red = sm;
    float* prow = reinterpret_cast<float*>(sm + 32);
    float* grow = prow + L;
    float* ys = grow + L;
    // ground truth first: `ys` is left holding the predicted PDF, which the clamp test below re-reads
    const double sg = upsample_clamped((axis ? gy : gx) + (int64_t)b * Ng, Ng, axis ? Mgy : Mgx, L, ys, grow, red);
    __syncthreads();
    const double sp = upsample_clamped((axis ? py : px) + (int64_t)b * N, N, M, L, ys, prow, red);
    const float dp = fmaxf((float)sp, 1e-6f), dg = fmaxf((float)sg, 1e-6f);
    const bool sum_live = (float)sp >= 1e-6f;                 // clamp_min passes the gradient when sum >= 1e-6
    const float scale = __ldg(upstream) / ((float)B * (float)L);
    // s_i = d L / d p_img_i; dot = sum_i s_i * p_img_i
    double part = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float ph = prow[i] / dp, d = ph - grow[i] / dg;
        const float s = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
        grow[i] = s;                                          // the ground-truth row is not needed any more
        part += (double)s * (double)ph;
    }
    const float dot_all = (float)block_sum(part, red);
    const float dot = sum_live ? dot_all : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        // prow[i] > 0 <=> the un-clamped bin was > 0; == 0 covers both "clamped" and "exactly zero" -- torch passes
        // the gradient at exactly zero, so re-derive the sign from the raw bin only there
        float live = prow[i] > 0.f ? 1.f : 0.f;
        if (prow[i] == 0.f) {
            const float* m = M + (int64_t)i * N;
            float acc = 0.f;
            for (int k = 0; k < N; ++k) acc = fmaf(__ldg(m + k), ys[k], acc);
            live = acc >= 0.f ? 1.f : 0.f;
        }
        grow[i] = live * (grow[i] - dot) / dp;
    }
    __syncthreads();
    // d p_s[k] = sum_i dv_i M[i][k]: a warp owns bins k, k + warps, ...
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* out = (axis ? dpy : dpx) + (int64_t)b * N;
    for (int k = wid; k < N; k += nw) {
        float acc = 0.f;
        for (int i = lane; i < L; i += 32) acc = fmaf(grow[i], __ldg(M + (int64_t)i * N + k), acc);
        acc = warp_sum(acc);
        if (lane == 0) out[k] = acc;
    }
}

size_t loss_smem(int W, int H, int n_max) {
    const int L = W > H ? W : H;
    return sizeof(double) * 32 + sizeof(float) * ((size_t)2 * L + n_max);
}

}  // namespace

int launch_pdf_l1_loss(const float* px, const float* py, const float* gx, const float* gy, int B, int Nx, int Ny, int Ngx,
                       int Ngy, const float* Mx, const float* My, const float* Mgx, const float* Mgy, int W, int H,
                       void* workspace, float* loss, cudaStream_t st) {
    const int n_max = (Nx > Ny ? Nx : Ny) > (Ngx > Ngy ? Ngx : Ngy) ? (Nx > Ny ? Nx : Ny) : (Ngx > Ngy ? Ngx : Ngy);
    const size_t smem = loss_smem(W, H, n_max);
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "pdf_l1_loss: rows of %d / %d bins are too long", W, H);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(pdf_l1_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double* partial = static_cast<double*>(workspace);
    unsigned* counter = reinterpret_cast<unsigned*>(partial + 2 * (size_t)B);
    pdf_l1_loss_kernel<<<dim3(B, 2), kLossThreads, smem, st>>>(px, py, gx, gy, Nx, Ny, Ngx, Ngy, Mx, My, Mgx, Mgy, W, H,
                                                                partial, counter, loss);
    return check_launch("pdf_l1_loss_kernel");
}

int launch_pdf_l1_loss_backward(const float* px, const float* py, const float* gx, const float* gy, int B, int Nx, int Ny,
                                int Ngx, int Ngy, const float* Mx, const float* My, const float* Mgx, const float* Mgy,
                                int W, int H, const float* upstream, float* dpx, float* dpy, cudaStream_t st) {
    const int n_max = (Nx > Ny ? Nx : Ny) > (Ngx > Ngy ? Ngx : Ngy) ? (Nx > Ny ? Nx : Ny) : (Ngx > Ngy ? Ngx : Ngy);
    const size_t smem = loss_smem(W, H, n_max);
    if (smem > 200 * 1024) return fail(ATTWARP_ERR_UNSUPPORTED, "pdf_l1_loss: rows of %d / %d bins are too long", W, H);
    if (smem > 48 * 1024)
        AW_CUDA(cudaFuncSetAttribute(pdf_l1_loss_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pdf_l1_loss_backward_kernel<<<dim3(B, 2), kLossThreads, smem, st>>>(px, py, gx, gy, Nx, Ny, Ngx, Ngy, Mx, My, Mgx,
                                                                         Mgy, W, H, upstream, dpx, dpy);
    return check_launch("pdf_l1_loss_backward_kernel");
}

}  // namespace aw
