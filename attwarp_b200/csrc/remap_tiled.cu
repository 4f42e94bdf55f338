// remap_tiled.cu -- stage 5 for uint8 images: the shared-memory / TMA-bulk-copy resample kernel.
//
// Same arithmetic as remap_direct_kernel (cv2.remap INTER_LINEAR + BORDER_REPLICATE, see
// warp_math.h; reference call sites "Attention Guided Warping/new_method.py:268-271",
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198"), organised for the memory system:
//
//   * tile = R output rows x blockDim.x output columns of one image; one thread per output column;
//   * the source rows the tile needs (<= 2R, found from map_y) are staged in shared memory, one
//     cp.async.bulk (TMA bulk copy, global -> shared, 16-byte aligned span around the needed
//     columns) per row, completion on an mbarrier: no LSU instructions, no register staging,
//     every DRAM sector of the source is fetched once per tile;
//   * the warp is separable, so the horizontal blend of a source row is computed ONCE per
//     (source row, output column) -- a funnel-shifted 8-byte window and dp4a with byte-positioned
//     weights -- kept in registers, and reused by every output row that taps that source row
//     (thread walks down its column: "A/B" row registers, all control flow CTA-uniform);
//   * the vertical blend + rounding writes bytes into a shared-memory output tile laid out with
//     the same 16-byte phase as its global destination, which leaves through cp.async.bulk
//     (shared -> global) for the aligned interior and a few byte stores for ragged row ends.
//
// Tiles whose source footprint does not fit the arena (strong local minification) are split into
// fewer output rows per pass; if even one output row does not fit, that row is gathered straight
// from global memory (same arithmetic), so the kernel is total.
#include "common.cuh"

namespace aw {
namespace {

constexpr int kTileRows = 16;          // R: output rows per tile
constexpr int kMaxSlots = 2 * kTileRows;
constexpr int kMaxThreads = 384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

struct TileTables {
    int need_off[kMaxSlots];   // (global address of the slot's first needed byte) & 15
    int slot_a[kTileRows];     // per output row: slot of the upper / lower source row
    int slot_b[kTileRows];
    int w_a[kTileRows];        // weight of the upper row (lower row gets 32 - w_a)
    int out_off[kTileRows];    // (global address of the output row segment) & 15
    int n_rows;                // output rows in this pass (0 => the next row takes the direct path)
    int n_slots;
};

// Horizontal blend of the C channels of output column `x` on one staged source row.
//   row   : shared-memory address of the slot
//   a     : byte offset of the window (first tap's first channel) from the slot start
//   wA/wB : dp4a weight words, see setup below
template <int C>
__device__ __forceinline__ void hblend_row(const uint8_t* row, int a, const uint32_t* wA,
                                           const uint32_t* wB, uint32_t* h) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + (a & ~3));
    const int sh = (a & 3) << 3;
    const uint32_t lo = wp[0], mid = wp[1];
    const uint32_t A = __funnelshift_r(lo, mid, sh);
    if (C == 1) {
        h[0] = __dp4a(A, wA[0], 0u);
    } else {
        const uint32_t hi = wp[2];
        const uint32_t Bv = __funnelshift_r(mid, hi, sh);
        if (C == 3) {
            h[0] = __dp4a(A, wA[0], 0u);                       // b0 (byte 0) , b1 (byte 3)
            h[1] = __dp4a(Bv, wB[1], __dp4a(A, wA[1], 0u));    // g0 (byte 1) , g1 (byte 4)
            h[2] = __dp4a(Bv, wB[2], __dp4a(A, wA[2], 0u));    // r0 (byte 2) , r1 (byte 5)
        } else {                                               // C == 4: taps at byte k and k+4
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __dp4a(Bv, wB[k], __dp4a(A, wA[k], 0u));
        }
    }
}

template <int C>
__global__ void __launch_bounds__(kMaxThreads)
remap_u8_tiled_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int Ho,
                      int Wo, const float* __restrict__ map_x, const float* __restrict__ map_y,
                      int map_div, int arena_bytes, int out_pitch) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ TileTables tb;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int s_min, s_max;

    uint8_t* arena = smem;                              // staged source rows (+16 B slack)
    uint8_t* stage = smem + arena_bytes + 16;           // R x out_pitch output tile

    const int Wt = blockDim.x;
    const int img = blockIdx.z;
    const int mrow = img / map_div;                     // CHW planes share their image's maps
    const int y_tile0 = blockIdx.y * kTileRows;
    const int y_tile1 = min(y_tile0 + kTileRows, Ho);
    const int x_first = blockIdx.x * Wt;
    const int ncols = min(Wt, Wo - x_first);            // valid output columns of this tile
    const int xl = threadIdx.x;
    const bool xvalid = xl < ncols;
    const uint8_t* simg = src + (int64_t)img * H * W * C;
    uint8_t* dimg = dst + (int64_t)img * Ho * Wo * C;
    const float* my = map_y + (int64_t)mrow * Ho;

    // ---- per-column setup: base source pixel, weights ------------------------------------------
    int xb = 0, w0 = 32, w1 = 0;
    if (xvalid) {
        const int sx = quantise_coord(__ldg(map_x + (int64_t)mrow * Wo + x_first + xl));
        const int ix = sx >> 5, ax = sx & 31;
        if (W >= 2) {
            if (ix < 0) { xb = 0; w0 = 32; w1 = 0; }
            else if (ix >= W - 1) { xb = W - 2; w0 = 0; w1 = 32; }
            else { xb = ix; w0 = 32 - ax; w1 = ax; }
        }
    }
    if (threadIdx.x == 0) {
        s_min = 0x7fffffff;
        s_max = -1;
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    {
        int lo = xvalid ? xb : 0x7fffffff, hi = xvalid ? xb : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&s_min, lo);
            atomicMax(&s_max, hi);
        }
    }
    __syncthreads();
    const int c_lo = s_min;
    const int c_hi = min(s_max + 1, W - 1);
    const int row_bytes = (c_hi - c_lo + 1) * C;
    const int slot_pitch = ((row_bytes + 15 + 15) & ~15) + 16;   // alignment head + window over-read
    const int max_slots = min(arena_bytes / slot_pitch, kMaxSlots);
    const int wo = (xb - c_lo) * C;                     // window byte offset from the span start

    // dp4a weight words: tap0 of channel k sits at byte k of the 8-byte window, tap1 at byte k+C
    uint32_t wA[C], wB[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        wA[k] = (uint32_t)w0 << (8 * k);
        wB[k] = 0u;
        if (k + C < 4) wA[k] |= (uint32_t)w1 << (8 * (k + C));
        else wB[k] = (uint32_t)w1 << (8 * (k + C - 4));
    }

    uint32_t parity = 0;
    int y_cur = y_tile0;
    while (y_cur < y_tile1) {
        // ---- warp 0 plans the pass: lane i <-> output row y_cur + i ----------------------------
        // source rows [r_min, r_max] needed by the first n_rows output rows are staged as one
        // slot per row (slot = row - r_min); n_rows is the longest prefix whose span fits.
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            const int y = y_cur + lane;
            const bool live = y < y_tile1 && lane < kTileRows;
            int ra = 0, rb = 0, wa = 32;
            if (live) {
                const int sy = quantise_coord(__ldg(my + y));
                const int iy = sy >> 5, ay = sy & 31;
                if (H >= 2) {
                    if (iy < 0) { ra = 0; wa = 32; }
                    else if (iy >= H - 1) { ra = H - 2; wa = 0; }
                    else { ra = iy; wa = 32 - ay; }
                }
                rb = min(ra + 1, H - 1);
            }
            int pmin = live ? ra : 0x7fffffff, pmax = live ? rb : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {          // inclusive prefix min / max
                const int a = __shfl_up_sync(0xffffffffu, pmin, o);
                const int b = __shfl_up_sync(0xffffffffu, pmax, o);
                if (lane >= o) { pmin = min(pmin, a); pmax = max(pmax, b); }
            }
            const bool fits = live && (pmax - pmin + 1 <= max_slots);
            const unsigned bad = ~__ballot_sync(0xffffffffu, fits);
            const int n_rows = bad ? (__ffs(bad) - 1) : 32;                 // leading fitting rows
            const int last = max(n_rows - 1, 0);
            const int r_min = __shfl_sync(0xffffffffu, pmin, last);
            const int r_max = __shfl_sync(0xffffffffu, pmax, last);
            const int n_slots = n_rows > 0 ? (r_max - r_min + 1) : 0;
            if (lane < n_rows) {
                tb.slot_a[lane] = ra - r_min;
                tb.slot_b[lane] = rb - r_min;
                tb.w_a[lane] = wa;
                const uintptr_t gd = reinterpret_cast<uintptr_t>(dimg + ((int64_t)y * Wo + x_first) * C);
                tb.out_off[lane] = (int)(gd & 15);
            }
            unsigned tx = 0;
            for (int k = lane; k < n_slots; k += 32) {
                const uintptr_t g = reinterpret_cast<uintptr_t>(simg + ((int64_t)(r_min + k) * W + c_lo) * C);
                const int off = (int)(g & 15);
                tb.need_off[k] = off;
                tx += (unsigned)((off + row_bytes + 15) & ~15);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, o);
            if (lane == 0) {
                tb.n_rows = n_rows;
                tb.n_slots = n_slots;
                if (n_rows > 0) mbar_arrive_expect_tx(&mbar, tx);
            }
            __syncwarp();
            // ---- stage the needed source rows: one bulk copy per slot ---------------------------
            for (int k = lane; k < n_slots; k += 32) {
                const uint8_t* g = simg + ((int64_t)(r_min + k) * W + c_lo) * C;
                const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                bulk_g2s(arena + k * slot_pitch, g - off, (uint32_t)((off + row_bytes + 15) & ~15), &mbar);
            }
        }
        __syncthreads();
        const int n_rows = tb.n_rows;

        if (n_rows == 0) {
            // ---- direct path for one output row whose footprint exceeds the arena ---------------
            if (xvalid) {
                const int sy = quantise_coord(__ldg(my + y_cur));
                const int ay = sy & 31;
                const int y0 = clampi(sy >> 5, 0, H - 1), y1 = clampi((sy >> 5) + 1, 0, H - 1);
                const int sx = quantise_coord(__ldg(map_x + (int64_t)mrow * Wo + x_first + xl));
                const int ax = sx & 31;
                const int x0 = clampi(sx >> 5, 0, W - 1), x1 = clampi((sx >> 5) + 1, 0, W - 1);
                uint8_t* o = dimg + ((int64_t)y_cur * Wo + x_first + xl) * C;
#pragma unroll
                for (int k = 0; k < C; ++k)
                    o[k] = bilinear_u8(__ldg(simg + ((int64_t)y0 * W + x0) * C + k),
                                       __ldg(simg + ((int64_t)y0 * W + x1) * C + k),
                                       __ldg(simg + ((int64_t)y1 * W + x0) * C + k),
                                       __ldg(simg + ((int64_t)y1 * W + x1) * C + k), ax, ay);
            }
            y_cur += 1;
            __syncthreads();
            continue;
        }

        mbar_wait(&mbar, parity);
        parity ^= 1u;

        // ---- walk down the column --------------------------------------------------------------
        if (xvalid) {
            uint32_t hA[C], hB[C];
            int curA = -1, curB = -1;
            for (int i = 0; i < n_rows; ++i) {
                const int sa = tb.slot_a[i], sb = tb.slot_b[i], wa = tb.w_a[i];
                if (sa != curA) {
                    if (sa == curB) {
#pragma unroll
                        for (int k = 0; k < C; ++k) hA[k] = hB[k];
                    } else {
                        hblend_row<C>(arena + sa * slot_pitch, tb.need_off[sa] + wo, wA, wB, hA);
                    }
                    curA = sa;
                }
                if (sb != curB) {
                    if (sb == curA) {
#pragma unroll
                        for (int k = 0; k < C; ++k) hB[k] = hA[k];
                    } else {
                        hblend_row<C>(arena + sb * slot_pitch, tb.need_off[sb] + wo, wA, wB, hB);
                    }
                    curB = sb;
                }
                uint8_t* o = stage + i * out_pitch + tb.out_off[i] + xl * C;
                const uint32_t wb = 32u - (uint32_t)wa;
                if (C == 4) {
                    uint32_t pk = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        pk |= (((uint32_t)wa * hA[k] + wb * hB[k] + 512u) >> 10) << (8 * k);
                    if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                        *reinterpret_cast<uint32_t*>(o) = pk;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) o[k] = (uint8_t)(pk >> (8 * k));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < C; ++k)
                        o[k] = (uint8_t)(((uint32_t)wa * hA[k] + wb * hB[k] + 512u) >> 10);
                }
            }
        }
        fence_proxy_async();       // make the generic-proxy writes to `stage` visible to the bulk store
        __syncthreads();

        // ---- ship the output rows: bulk store for the 16-byte aligned interior, bytes for the ends
        const int len = ncols * C;
        if (threadIdx.x < 32) {
            for (int i = threadIdx.x; i < n_rows; i += 32) {
                uint8_t* g = dimg + ((int64_t)(y_cur + i) * Wo + x_first) * C;
                const int off = tb.out_off[i];
                const int head = (16 - off) & 15;
                const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                if (body > 0) bulk_s2g(g + head, stage + i * out_pitch + off + head, (uint32_t)body);
            }
            bulk_commit();
        }
        // ragged ends: <= 15 head bytes and <= 15 tail bytes per row, one thread per byte
        for (int t = (int)threadIdx.x - 32; t >= 0 && t < n_rows * 32; t += (int)blockDim.x - 32) {
            const int i = t >> 5, j = t & 31;
            uint8_t* g = dimg + ((int64_t)(y_cur + i) * Wo + x_first) * C;
            const int off = tb.out_off[i];
            const int head = min((16 - off) & 15, len);
            const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
            const uint8_t* s = stage + i * out_pitch + off;
            if (j < 16) {
                if (j < head) g[j] = s[j];
            } else {
                const int q = head + body + (j - 16);
                if (q < len) g[q] = s[q];
            }
        }
        if (threadIdx.x < 32) bulk_wait_read0();
        __syncthreads();
        y_cur += n_rows;
    }
}

template <int C>
int launch_tiled(const uint8_t* src, uint8_t* dst, int n_img, int H, int W, int Ho, int Wo,
                 const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    // one thread per output column; tiles as wide as possible up to kMaxThreads
    const int n_ct = (Wo + kMaxThreads - 1) / kMaxThreads;
    int threads = ((Wo + n_ct - 1) / n_ct + 31) & ~31;
    if (threads < 64) threads = 64;                      // warp 0 issues stores, the rest do row ends
    const int out_pitch = ((threads * C + 15 + 15) & ~15);
    // arena: room for R+2 rows at unit scale plus 60 % slack for local minification
    const int unit_pitch = (((threads + 1) * C + 30) & ~15) + 16;
    int arena = (kTileRows + 2) * unit_pitch * 8 / 5;
    arena = (arena + 127) & ~127;
    const size_t smem = (size_t)arena + 16 + (size_t)kTileRows * out_pitch;
    auto kern = remap_u8_tiled_kernel<C>;
    if (smem > 40 * 1024)  // static tables count against the 48 KB default too: opt in early
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(n_ct, (Ho + kTileRows - 1) / kTileRows, n_img);
    kern<<<grid, threads, smem, st>>>(src, dst, H, W, Ho, Wo, map_x, map_y, map_div, arena, out_pitch);
    return check_launch("remap_u8_tiled_kernel");
}

}  // namespace

// uint8 images, HWC with C in {1,3,4} or planar (n_img = B*C single-channel planes, map_div = C).
int launch_remap_u8_tiled(const void* src, void* dst, int n_img, int C, int H, int W, int Ho, int Wo,
                          const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    switch (C) {
        case 1: return launch_tiled<1>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 3: return launch_tiled<3>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 4: return launch_tiled<4>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        default: return fail(ATTWARP_ERR_UNSUPPORTED, "tiled remap supports C in {1,3,4} (got %d)", C);
    }
}

}  // namespace aw
