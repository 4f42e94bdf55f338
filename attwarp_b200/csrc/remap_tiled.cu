// remap_tiled.cu -- stage 5 for uint8 images: persistent, warp-specialised resample kernel.
//
// Same arithmetic as remap_direct_kernel (cv2.remap INTER_LINEAR + BORDER_REPLICATE, see
// warp_math.h; reference call sites "Attention Guided Warping/new_method.py:268-271",
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198"), organised for the memory system
// and for the issue slots (a gather-resample of 3-byte pixels is instruction-bound long before it
// is HBM-bound, so the inner loop is counted in instructions per output pixel):
//
//   * work unit ("tile") = R output rows x Wt output columns of one image; the grid is persistent
//     (a few CTAs per SM), every CTA walks a contiguous range of tiles;
//   * one PRODUCER warp per CTA plans passes and moves source rows: the distinct source rows the
//     next output rows tap are ranked into compact "slots" (rows no output row taps are neither
//     copied nor blended) and fetched with one cp.async.bulk (TMA bulk copy, global -> shared,
//     16-byte aligned span around the needed columns) per slot into a ring of stages; completion
//     is signalled on a `full` mbarrier, reuse is gated by an `empty` mbarrier -- the copy of pass
//     p+1 overlaps the arithmetic of pass p, no LSU instructions or registers carry source bytes;
//   * CONSUMER warps hold one output column per thread.  The warp is separable, so the horizontal
//     blend of a source row is computed ONCE per (slot, output column) -- a funnel-shifted 8-byte
//     window and dp4a with byte-positioned weights -- and packed with the previous slot's blend as
//     the two 16-bit halves of one register; every output row whose taps are that slot pair is
//     then ONE dp2a (vertical blend + rounding constant) and a shift per channel.  The sweep runs
//     over slots in source order; all control flow is CTA-uniform and table-driven;
//   * results go to one of two shared-memory output tiles laid out with the same 16-byte phase as
//     their global destination and leave through cp.async.bulk (shared -> global) for the aligned
//     interior plus a few byte stores for ragged row ends; the store of pass p overlaps pass p+1.
//
// Maps need not be monotone: a pass whose rows are not in non-decreasing source order takes a
// per-row path (two fresh horizontal blends per output row, same arithmetic).  Passes whose
// source footprint does not fit a stage are split into fewer output rows; if even one output row
// does not fit, that row is gathered straight from global memory, so the kernel is total.
#include <stdlib.h>

#include "common.cuh"

namespace aw {
namespace {

constexpr int kStages = 2;             // source-row stages in flight per CTA
constexpr int kMaxThreads = 384;       // consumer threads (= output columns per tile)

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src_gmem), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
// Keeps a loop-invariant value in a register (stops the compiler from rematerialising it).
__device__ __forceinline__ uint32_t pin(uint32_t v) {
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return v;
}
// Shared memory is addressed as byte offsets from the one dynamic array below: the compiler then
// emits LDS/STS with register+immediate addressing and knows which loads are CTA-uniform (so
// the table-driven loops stay free of divergence bookkeeping).
extern __shared__ __align__(128) uint8_t smem[];
__device__ __forceinline__ uint32_t ld32(int off) { return *reinterpret_cast<const uint32_t*>(smem + off); }
__device__ __forceinline__ uint4 ld128(int off) { return *reinterpret_cast<const uint4*>(smem + off); }
__device__ __forceinline__ void st32(int off, uint32_t v) { *reinterpret_cast<uint32_t*>(smem + off) = v; }
__device__ __forceinline__ void st128(int off, uint4 v) { *reinterpret_cast<uint4*>(smem + off) = v; }
__device__ __forceinline__ void st8(int off, uint32_t v) { smem[off] = (uint8_t)v; }

// ---- per-stage pass table (byte offsets), written by the producer warp, read by the consumers ----
//   +0   uint4 {n_rows, n_slots, flags, c_lo}    n_rows 0: row y0 takes the direct path; -1: stop
//                                                flags bit0: rows in non-decreasing source order
//                                                      bit1: first pass of a new (image, strip)
//   +16  uint4 {img, x_first, y0, slot_pitch}
//   +32  uint4 {phase of slot 0 (global address & 15), 0, 0, 0}
//   +48  uint4 row[R + 1]:  x = dp2a weight word  wa | (32 - wa) << 8  (upper row, lower row)
//                           y = byte offset of the row inside the output tile
//                               (i * out_pitch + (dst address & 15))
//                           z = slot of the LOWER tap; entry n_rows is a sentinel (z = 0xffffffff)
//   +48 + 16 (R + 1)  uint32 slot_base[2 R]:  slot * slot_pitch + (global address of its first byte & 15)
constexpr int kTabRows = 48;
template <int R> constexpr int tab_slots() { return kTabRows + 16 * (R + 1); }
template <int R> constexpr int tab_bytes() { return tab_slots<R>() + 4 * 2 * R; }
constexpr uint32_t kFlagMonotone = 1u, kFlagNewStrip = 2u;

// Horizontal blend of the C channels of one output column on one staged source row.
//   wp    : byte offset (multiple of 4) of the word holding the window's first byte
//   sh    : 8 * (byte offset of the window inside that word); only the low 5 bits are used
//   wA/wB : dp4a weight words: tap0 of channel k at byte k of the 8-byte window, tap1 at k + C
template <int C>
__device__ __forceinline__ void hblend_row(int wp, uint32_t sh, const uint32_t* wA, const uint32_t* wB,
                                           uint32_t* h) {
    const uint32_t lo = ld32(wp), mid = ld32(wp + 4);
    const uint32_t A = __funnelshift_r(lo, mid, sh);
    if (C == 1) {
        h[0] = __dp4a(A, wA[0], 0u);
    } else {
        const uint32_t hi = ld32(wp + 8);
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

// Vertical blend + cv2 rounding of one output pixel from packed (upper | lower << 16) row blends.
template <int C>
__device__ __forceinline__ void vblend_store(const uint32_t* PQ, uint32_t wy, int o) {
    if (C == 4) {
        uint32_t pk = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) pk |= (__dp2a_lo(PQ[k], wy, 512u) >> 10) << (8 * k);
        if ((o & 3) == 0) {
            st32(o, pk);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) st8(o + k, pk >> (8 * k));
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) st8(o + k, __dp2a_lo(PQ[k], wy, 512u) >> 10);
    }
}

__device__ __forceinline__ uint64_t shfl_up64(uint64_t v, int o) {
    const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, o);
    const uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
    const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
// base source column and tap weights of one output column (border replicate folded into weights)
__device__ __forceinline__ void column_taps(float m, int W, int& xb, int& w0, int& w1) {
    const int sx = quantise_coord(m);
    const int ix = sx >> 5, ax = sx & 31;
    if (ix < 0) { xb = 0; w0 = 32; w1 = 0; }
    else if (ix >= W - 1) { xb = W - 2; w0 = 0; w1 = 32; }
    else { xb = ix; w0 = 32 - ax; w1 = ax; }
}


// ---- the monotone sweep, hand-scheduled in PTX ----------------------------------------------------
// The table-driven loops below branch on values loaded from shared memory.  They are CTA-uniform,
// but the compiler cannot prove it and wraps every branch in BSSY/BSYNC reconvergence bookkeeping
// and rebuilds the shared-window base per iteration (a quarter of the loop).  Writing the loop in
// PTX with `bra.uni` and explicit shared-space addresses removes all of that.
//   cur/arena : shared-space address of the window (U: word holding its first byte in slot 0;
//               !U: start of the staged span + this column's window offset)
//   sh        : U only, 8 * (window byte offset inside its word)
//   sp        : !U only, shared-space address of slot_base[0]
//   rp        : shared-space address of row[0];  ocol: of this column inside the output tile
#define AW_SWEEP_ROWCTL_BEGIN                                   \
    "setp.ne.u32 q, ez, s;\n"                                   \
    "@q bra.uni NEXT;\n"                                        \
    "ROW:\n"
#define AW_SWEEP_ROWCTL_END                                     \
    "add.u32 rp, rp, 16;\n"                                     \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"                \
    "setp.eq.u32 q, ez, s;\n"                                   \
    "@q bra.uni ROW;\n"                                         \
    "NEXT:\n"                                                   \
    "add.s32 s, s, 1;\n"                                        \
    "setp.lt.s32 p, s, %0;\n"
#define AW_SWEEP_WINDOW_U                                       \
    "ld.shared.b32 lo, [cur];\n"                                \
    "ld.shared.b32 mid, [cur+4];\n"
#define AW_SWEEP_WINDOW_T                                       \
    "ld.shared.b32 t, [sp];\n"                                  \
    "add.u32 t, t, %1;\n"                                       \
    "and.b32 cur, t, 0xfffffffc;\n"                             \
    "shl.b32 sh, t, 3;\n"                                       \
    "add.u32 sp, sp, 4;\n"                                      \
    "ld.shared.b32 lo, [cur];\n"                                \
    "ld.shared.b32 mid, [cur+4];\n"

template <bool U>
__device__ __forceinline__ void sweep_c3(int n_slots, uint32_t cur_or_arena, uint32_t sh_or_sp, uint32_t pitch,
                                         uint32_t rp, uint32_t ocol, const uint32_t* wA, const uint32_t* wB) {
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, hi, A, B, h0, h1, h2, t, P0, P1, P2, r0, r1, r2, o, cur, rp, ex, ey, ez, ew;\n"
            "mov.b32 P0, 0;\n mov.b32 P1, 0;\n mov.b32 P2, 0;\n mov.b32 s, 0;\n"
            "mov.b32 cur, %1;\n mov.b32 rp, %4;\n"
            "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"
            "setp.lt.s32 p, s, %0;\n"
            "@!p bra.uni DONE;\n"
            "SLOT:\n" AW_SWEEP_WINDOW_U
            "ld.shared.b32 hi, [cur+8];\n"
            "shf.r.wrap.b32 A, lo, mid, %2;\n"
            "shf.r.wrap.b32 B, mid, hi, %2;\n"
            "dp4a.u32.u32 h0, A, %6, 0;\n"
            "dp4a.u32.u32 t, A, %7, 0;\n"
            "dp4a.u32.u32 h1, B, %9, t;\n"
            "dp4a.u32.u32 t, A, %8, 0;\n"
            "dp4a.u32.u32 h2, B, %10, t;\n"
            "prmt.b32 P0, P0, h0, 0x5432;\n"
            "prmt.b32 P1, P1, h1, 0x5432;\n"
            "prmt.b32 P2, P2, h2, 0x5432;\n"
            "add.u32 cur, cur, %3;\n" AW_SWEEP_ROWCTL_BEGIN
            "dp2a.lo.u32.u32 r0, P0, ex, 512;\n"
            "dp2a.lo.u32.u32 r1, P1, ex, 512;\n"
            "dp2a.lo.u32.u32 r2, P2, ex, 512;\n"
            "add.u32 o, ey, %5;\n"
            "shr.u32 r0, r0, 10;\n shr.u32 r1, r1, 10;\n shr.u32 r2, r2, 10;\n"
            "st.shared.u8 [o], r0;\n st.shared.u8 [o+1], r1;\n st.shared.u8 [o+2], r2;\n" AW_SWEEP_ROWCTL_END
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n" ::"r"(n_slots),
            "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]), "r"(wA[1]), "r"(wA[2]),
            "r"(wB[1]), "r"(wB[2])
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, hi, A, B, h0, h1, h2, t, P0, P1, P2, r0, r1, r2, o, cur, sh, sp, rp, ex, ey, ez, ew;\n"
            "mov.b32 P0, 0;\n mov.b32 P1, 0;\n mov.b32 P2, 0;\n mov.b32 s, 0;\n"
            "mov.b32 sp, %2;\n mov.b32 rp, %4;\n"
            "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"
            "setp.lt.s32 p, s, %0;\n"
            "@!p bra.uni DONE;\n"
            "SLOT:\n" AW_SWEEP_WINDOW_T
            "ld.shared.b32 hi, [cur+8];\n"
            "shf.r.wrap.b32 A, lo, mid, sh;\n"
            "shf.r.wrap.b32 B, mid, hi, sh;\n"
            "dp4a.u32.u32 h0, A, %6, 0;\n"
            "dp4a.u32.u32 t, A, %7, 0;\n"
            "dp4a.u32.u32 h1, B, %9, t;\n"
            "dp4a.u32.u32 t, A, %8, 0;\n"
            "dp4a.u32.u32 h2, B, %10, t;\n"
            "prmt.b32 P0, P0, h0, 0x5432;\n"
            "prmt.b32 P1, P1, h1, 0x5432;\n"
            "prmt.b32 P2, P2, h2, 0x5432;\n" AW_SWEEP_ROWCTL_BEGIN
            "dp2a.lo.u32.u32 r0, P0, ex, 512;\n"
            "dp2a.lo.u32.u32 r1, P1, ex, 512;\n"
            "dp2a.lo.u32.u32 r2, P2, ex, 512;\n"
            "add.u32 o, ey, %5;\n"
            "shr.u32 r0, r0, 10;\n shr.u32 r1, r1, 10;\n shr.u32 r2, r2, 10;\n"
            "st.shared.u8 [o], r0;\n st.shared.u8 [o+1], r1;\n st.shared.u8 [o+2], r2;\n" AW_SWEEP_ROWCTL_END
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n" ::"r"(n_slots),
            "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]), "r"(wA[1]), "r"(wA[2]),
            "r"(wB[1]), "r"(wB[2])
            : "memory");
    }
}

template <bool U>
__device__ __forceinline__ void sweep_c1(int n_slots, uint32_t cur_or_arena, uint32_t sh_or_sp, uint32_t pitch,
                                         uint32_t rp, uint32_t ocol, const uint32_t* wA) {
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, A, h0, t, P0, r0, o, cur, rp, ex, ey, ez, ew;\n"
            "mov.b32 P0, 0;\n mov.b32 s, 0;\n"
            "mov.b32 cur, %1;\n mov.b32 rp, %4;\n"
            "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"
            "setp.lt.s32 p, s, %0;\n"
            "@!p bra.uni DONE;\n"
            "SLOT:\n" AW_SWEEP_WINDOW_U
            "shf.r.wrap.b32 A, lo, mid, %2;\n"
            "dp4a.u32.u32 h0, A, %6, 0;\n"
            "prmt.b32 P0, P0, h0, 0x5432;\n"
            "add.u32 cur, cur, %3;\n" AW_SWEEP_ROWCTL_BEGIN
            "dp2a.lo.u32.u32 r0, P0, ex, 512;\n"
            "add.u32 o, ey, %5;\n"
            "shr.u32 r0, r0, 10;\n"
            "st.shared.u8 [o], r0;\n" AW_SWEEP_ROWCTL_END
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n" ::"r"(n_slots),
            "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0])
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, A, h0, t, P0, r0, o, cur, sh, sp, rp, ex, ey, ez, ew;\n"
            "mov.b32 P0, 0;\n mov.b32 s, 0;\n"
            "mov.b32 sp, %2;\n mov.b32 rp, %4;\n"
            "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"
            "setp.lt.s32 p, s, %0;\n"
            "@!p bra.uni DONE;\n"
            "SLOT:\n" AW_SWEEP_WINDOW_T
            "shf.r.wrap.b32 A, lo, mid, sh;\n"
            "dp4a.u32.u32 h0, A, %6, 0;\n"
            "prmt.b32 P0, P0, h0, 0x5432;\n" AW_SWEEP_ROWCTL_BEGIN
            "dp2a.lo.u32.u32 r0, P0, ex, 512;\n"
            "add.u32 o, ey, %5;\n"
            "shr.u32 r0, r0, 10;\n"
            "st.shared.u8 [o], r0;\n" AW_SWEEP_ROWCTL_END
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n" ::"r"(n_slots),
            "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0])
            : "memory");
    }
}

struct RemapArgs {
    const uint8_t* src;
    uint8_t* dst;
    const float* map_x;
    const float* map_y;
    int H, W, Ho, Wo;
    int map_div;             // CHW planes share their image's maps
    int n_strips, n_rowtiles, total_tiles;
    int stage_bytes;         // bytes of one source-row stage (multiple of 128)
    int out_pitch;           // bytes per row of an output tile
};

// Requires H >= 2 and W >= 2 (the launcher routes degenerate images to the direct kernel).
// blockDim.x = Wt consumer threads + 32 producer threads.
// U ("uniform phase"): W*C is a multiple of 16, so every staged row starts at the same 16-byte
// phase and slot k sits at k * slot_pitch + phase -- the sweep advances by one add per slot.
template <int C, int R, bool U>
__global__ void __launch_bounds__(kMaxThreads + 32, R <= 8 ? 4 : (R <= 12 ? 3 : 2))
remap_u8_tiled_kernel(const RemapArgs a) {
    // layout: [kStages source stages][2 output tiles][kStages pass tables][mbarriers]
    const int Wt = (int)blockDim.x - 32;
    const int tid = threadIdx.x;
    const int out_bytes = R * a.out_pitch;
    const int out_off0 = kStages * a.stage_bytes;
    const int tab_off0 = out_off0 + 2 * out_bytes;
    const int bar_off0 = tab_off0 + kStages * tab_bytes<R>();       // full[kStages], empty[kStages]
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t bars_s = smem_s + (uint32_t)bar_off0;
    const int H = a.H, W = a.W, Ho = a.Ho, Wo = a.Wo;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bars_s + 8u * s, 1);
            mbar_init(bars_s + 8u * (kStages + s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // contiguous, balanced range of tiles for this CTA
    const int t0 = (int)(((int64_t)a.total_tiles * blockIdx.x) / gridDim.x);
    const int t1 = (int)(((int64_t)a.total_tiles * (blockIdx.x + 1)) / gridDim.x);

    // warp-uniform role split (the shuffle tells the compiler the branch does not diverge)
    const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (warp_idx == (Wt >> 5)) {
        // =========================== producer warp =============================================
        const int lane = tid & 31;
        int it = 0;
        int cur_key = -1, c_lo = 0, row_bytes = 0, slot_pitch = 16, max_slots = 0;
        for (int t = t0; t < t1; ++t) {
            const int rt = t % a.n_rowtiles;
            const int q = t / a.n_rowtiles;
            const int strip = q % a.n_strips, img = q / a.n_strips;
            const int mrow = img / a.map_div;
            const int x_first = strip * Wt;
            const int ncols = min(Wt, Wo - x_first);
            const int y_tile1 = min((rt + 1) * R, Ho);
            const uint8_t* simg = a.src + (int64_t)img * H * W * C;
            uint8_t* dimg = a.dst + (int64_t)img * Ho * Wo * C;
            const float* my = a.map_y + (int64_t)mrow * Ho;
            uint32_t strip_flag = 0u;
            if (q != cur_key) {                              // source column span of this strip
                cur_key = q;
                strip_flag = kFlagNewStrip;
                int lo = 0x7fffffff, hi = -1;
                const float* mx = a.map_x + (int64_t)mrow * Wo + x_first;
                for (int x = lane; x < ncols; x += 32) {
                    int xb, w0, w1;
                    column_taps(__ldg(mx + x), W, xb, w0, w1);
                    lo = min(lo, xb);
                    hi = max(hi, xb);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                c_lo = lo;
                const int c_hi = min(hi + 1, W - 1);
                row_bytes = (c_hi - c_lo + 1) * C;
                slot_pitch = ((row_bytes + 15 + 15) & ~15) + 16;   // alignment head + window over-read
                max_slots = min(a.stage_bytes / slot_pitch, 2 * R);
            }
            int y_cur = rt * R;
            while (y_cur < y_tile1) {
                const int st = it % kStages;
                const int tab = tab_off0 + st * tab_bytes<R>();
                const uint32_t full_s = bars_s + 8u * st, empty_s = bars_s + 8u * (kStages + st);
                const uint32_t stage_s = smem_s + (uint32_t)(st * a.stage_bytes);
                // ---- plan: lane i <-> output row y_cur + i ------------------------------------
                const int y = y_cur + lane;
                const bool live = y < y_tile1 && lane < R;
                int ra = 0x3fffffff, wa = 32;           // upper source row (lower = ra + 1), its weight
                if (live) {
                    const int sy = quantise_coord(__ldg(my + y));
                    const int iy = sy >> 5, ay = sy & 31;
                    if (iy < 0) { ra = 0; wa = 32; }
                    else if (iy >= H - 1) { ra = H - 2; wa = 0; }
                    else { ra = iy; wa = 32 - ay; }
                }
                int r_base = ra;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r_base = min(r_base, __shfl_xor_sync(0xffffffffu, r_base, o));
                const int rel = ra - r_base;
                const bool in_range = live && rel < 63;
                uint64_t used = in_range ? (3ull << (rel & 63)) : 0ull;   // rows ra and ra + 1
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {                        // inclusive prefix OR
                    const uint64_t u = shfl_up64(used, o);
                    if (lane >= o) used |= u;
                }
                const bool fits = in_range && __popcll(used) <= max_slots;
                const unsigned bad = ~__ballot_sync(0xffffffffu, fits);
                const int n_rows = bad ? (__ffs(bad) - 1) : 32;           // leading fitting rows
                const uint64_t mask = shfl64(used, max(n_rows - 1, 0));
                const int n_slots = n_rows > 0 ? __popcll(mask) : 0;
                const int prev_ra = __shfl_up_sync(0xffffffffu, ra, 1);
                const unsigned desc = __ballot_sync(0xffffffffu, lane > 0 && lane < n_rows && ra < prev_ra);
                // Each lane stages the source rows it is the FIRST to tap (rows ra and ra + 1): its
                // slot ranks follow from the mask, no search for the k-th set bit is needed.
                uint64_t before = shfl_up64(used, 1);                           // rows tapped by earlier lanes
                if (lane == 0) before = 0ull;
                const int slot_a = __popcll(mask & ((1ull << (rel & 63)) - 1ull));  // rank of row ra
                const uint8_t* g[2];
                int off[2], slot[2];
                unsigned tx = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    g[j] = nullptr;
                    off[j] = 0;
                    slot[j] = slot_a + j;
                    if (lane < n_rows && !((before >> ((rel + j) & 63)) & 1ull)) {
                        const uint8_t* p = simg + ((int64_t)(ra + j) * W + c_lo) * C;
                        off[j] = (int)(reinterpret_cast<uintptr_t>(p) & 15);
                        g[j] = p - off[j];
                        tx += (unsigned)((off[j] + row_bytes + 15) & ~15);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, o);
                const int off_first = __shfl_sync(0xffffffffu, off[0], 0);       // phase of slot 0

                mbar_wait(empty_s, ((it / kStages) & 1) ^ 1);          // stage + table free again
                if (lane < n_rows) {
                    const int slot_b = slot_a + 1;                                       // rank of row ra + 1
                    const uintptr_t gd = reinterpret_cast<uintptr_t>(dimg + ((int64_t)y * Wo + x_first) * C);
                    st128(tab + kTabRows + 16 * lane,
                          make_uint4((uint32_t)wa | ((uint32_t)(32 - wa) << 8),
                                     (uint32_t)(lane * a.out_pitch) + (uint32_t)(gd & 15), (uint32_t)slot_b, 0u));
                } else if (lane == n_rows) {
                    st128(tab + kTabRows + 16 * lane, make_uint4(0u, 0u, 0xffffffffu, 0u));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (g[j] != nullptr)
                        st32(tab + tab_slots<R>() + 4 * slot[j], (uint32_t)(slot[j] * slot_pitch + off[j]));
                if (lane == 0) {
                    st128(tab, make_uint4((uint32_t)n_rows, (uint32_t)n_slots,
                                          (desc == 0u ? kFlagMonotone : 0u) | strip_flag, (uint32_t)c_lo));
                    st128(tab + 16, make_uint4((uint32_t)img, (uint32_t)x_first, (uint32_t)y_cur,
                                               (uint32_t)slot_pitch));
                    st128(tab + 32, make_uint4((uint32_t)off_first, 0u, 0u, 0u));
                }
                strip_flag = 0u;
                __syncwarp();
                if (lane == 0) {
                    if (n_rows > 0) mbar_arrive_expect_tx(full_s, tx);
                    else mbar_arrive(full_s);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (g[j] != nullptr)
                        bulk_g2s(stage_s + (uint32_t)(slot[j] * slot_pitch), g[j],
                                 (uint32_t)((off[j] + row_bytes + 15) & ~15), full_s);
                y_cur += max(n_rows, 1);
                ++it;
            }
        }
        // terminator
        {
            const int st = it % kStages;
            mbar_wait(bars_s + 8u * (kStages + st), ((it / kStages) & 1) ^ 1);
            if (lane == 0) {
                st128(tab_off0 + st * tab_bytes<R>(), make_uint4(0xffffffffu, 0u, 0u, 0u));
                mbar_arrive(bars_s + 8u * st);
            }
        }
        return;
    }

    // =============================== consumer warps ==============================================
    const int xl = tid;
    int wo = 0;                       // byte offset of this column's window inside a staged row span
    bool xvalid = false;
    uint32_t wA[C], wB[C];
#pragma unroll
    for (int k = 0; k < C; ++k) wA[k] = wB[k] = 0u;
    const int out_col = xl * C;
    const bool rows_aligned = ((Wo * C) & 15) == 0;

    for (int it = 0;; ++it) {
        const int st = it % kStages;
        const int tab = tab_off0 + st * tab_bytes<R>();
        mbar_wait(bars_s + 8u * st, (it / kStages) & 1);
        const uint4 h0 = ld128(tab);
        const int n_rows = (int)h0.x;
        if (n_rows < 0) break;
        const uint4 h1 = ld128(tab + 16);
        const int img = (int)h1.x, x_first = (int)h1.y, y0 = (int)h1.z;
        const int ncols = min(Wt, Wo - x_first);
        const int mrow = img / a.map_div;
        uint8_t* dimg = a.dst + (int64_t)img * Ho * Wo * C;
        if (h0.z & kFlagNewStrip) {                          // new strip: per-column taps and weights
            xvalid = xl < ncols;
            int w0 = 32, w1 = 0, xb = (int)h0.w;
            if (xvalid) column_taps(__ldg(a.map_x + (int64_t)mrow * Wo + x_first + xl), W, xb, w0, w1);
            wo = (xb - (int)h0.w) * C;
            // dp4a weight words: tap0 of channel k at byte k of the 8-byte window, tap1 at byte k+C
#pragma unroll
            for (int k = 0; k < C; ++k) {
                wA[k] = (uint32_t)w0 << (8 * k);
                wB[k] = 0u;
                if (k + C < 4) wA[k] |= (uint32_t)w1 << (8 * (k + C));
                else wB[k] = (uint32_t)w1 << (8 * (k + C - 4));
            }
        }
        const int obuf = out_off0 + (it & 1) * out_bytes;

        if (n_rows == 0) {
            // ---- direct path for one output row whose footprint exceeds a stage -----------------
            if (xvalid) {
                const uint8_t* simg = a.src + (int64_t)img * H * W * C;
                const int sy = quantise_coord(__ldg(a.map_y + (int64_t)mrow * Ho + y0));
                const int ay = sy & 31;
                const int ya = clampi(sy >> 5, 0, H - 1), yb = clampi((sy >> 5) + 1, 0, H - 1);
                const int sx = quantise_coord(__ldg(a.map_x + (int64_t)mrow * Wo + x_first + xl));
                const int ax = sx & 31;
                const int x0 = clampi(sx >> 5, 0, W - 1), x1 = clampi((sx >> 5) + 1, 0, W - 1);
                uint8_t* o = dimg + ((int64_t)y0 * Wo + x_first + xl) * C;
#pragma unroll
                for (int k = 0; k < C; ++k)
                    o[k] = bilinear_u8(__ldg(simg + ((int64_t)ya * W + x0) * C + k),
                                       __ldg(simg + ((int64_t)ya * W + x1) * C + k),
                                       __ldg(simg + ((int64_t)yb * W + x0) * C + k),
                                       __ldg(simg + ((int64_t)yb * W + x1) * C + k), ax, ay);
            }
        } else if (xvalid) {
            const int arena = st * a.stage_bytes + wo;
            const int ocol = obuf + out_col;
            const int slot_tab = tab + tab_slots<R>();
            const int n_slots = (int)h0.y;
            if (h0.z & kFlagMonotone) {
                // ---- sweep the slots in source order; emit the rows whose lower tap is the slot ---
                const int win0 = arena + (U ? (int)ld32(tab + 32) : 0);
                const uint32_t rp_s = smem_s + (uint32_t)(tab + kTabRows);
                const uint32_t ocol_s = smem_s + (uint32_t)ocol;
                if (C == 3) {
                    if (U) sweep_c3<true>(n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h1.w, rp_s, ocol_s, wA, wB);
                    else sweep_c3<false>(n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA, wB);
                } else if (C == 1) {
                    if (U) sweep_c1<true>(n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h1.w, rp_s, ocol_s, wA);
                    else sweep_c1<false>(n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA);
                } else {
                    uint32_t PQ[C];
#pragma unroll
                    for (int k = 0; k < C; ++k) PQ[k] = 0u;
                    int rp = tab + kTabRows;
                    uint4 re = ld128(rp);
                    for (int s = 0; s < n_slots; ++s) {
                        uint32_t h[C];
                        const int win = U ? win0 + s * (int)h1.w : arena + (int)ld32(slot_tab + 4 * s);
                        hblend_row<C>(win & ~3, (uint32_t)win << 3, wA, wB, h);
#pragma unroll
                        for (int k = 0; k < C; ++k) PQ[k] = __byte_perm(PQ[k], h[k], 0x5432);
                        while (re.z == (uint32_t)s) {
                            vblend_store<C>(PQ, re.x, ocol + (int)re.y);
                            rp += 16;
                            re = ld128(rp);
                        }
                    }
                }
            } else {
                // ---- rows in arbitrary source order: two fresh horizontal blends per row ----------
#pragma unroll 1
                for (int i = 0; i < n_rows; ++i) {
                    const uint4 re = ld128(tab + kTabRows + 16 * i);
                    uint32_t hA[C], hB[C], PQ[C];
                    const int a0 = arena + (int)ld32(slot_tab + 4 * (int)re.z - 4);
                    const int a1 = arena + (int)ld32(slot_tab + 4 * (int)re.z);
                    hblend_row<C>(a0 & ~3, (uint32_t)a0 << 3, wA, wB, hA);
                    hblend_row<C>(a1 & ~3, (uint32_t)a1 << 3, wA, wB, hB);
#pragma unroll
                    for (int k = 0; k < C; ++k) PQ[k] = hA[k] | (hB[k] << 16);
                    vblend_store<C>(PQ, re.x, ocol + (int)re.y);
                }
            }
        }
        // Stores of earlier passes issued by this thread have finished reading shared memory (their
        // tile is the one the NEXT pass writes); then publish this pass's tile to the async proxy.
        if (tid < R) bulk_wait_read0();
        fence_proxy_async();
        consumer_sync(Wt);
        if (tid == 0) mbar_arrive(bars_s + 8u * (kStages + st));   // stage + table free for the producer

        if (n_rows > 0) {
            // ---- ship the output rows: bulk store for the 16-byte aligned interior, bytes for the ends
            const int len = ncols * C;
            uint8_t* g0 = dimg + ((int64_t)y0 * Wo + x_first) * C;
            const bool ragged = !rows_aligned || ((reinterpret_cast<uintptr_t>(g0) | (uintptr_t)len) & 15) != 0;
            if (tid < n_rows) {
                uint8_t* g = g0 + (int64_t)tid * Wo * C;
                const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                const int head = (16 - off) & 15;
                const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                if (body > 0)
                    bulk_s2g(g + head, smem_s + (uint32_t)(obuf + tid * a.out_pitch + off + head), (uint32_t)body);
                bulk_commit();
            }
            if (ragged) {
                // <= 15 head bytes and <= 15 tail bytes per row, one thread per byte
                for (int t = tid; t < n_rows * 32; t += Wt) {
                    const int i = t >> 5, j = t & 31;
                    uint8_t* g = g0 + (int64_t)i * Wo * C;
                    const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                    const int head = min((16 - off) & 15, len);
                    const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                    const int s = obuf + i * a.out_pitch + off;
                    if (j < 16) {
                        if (j < head) g[j] = smem[s + j];
                    } else {
                        const int qq = head + body + (j - 16);
                        if (qq < len) g[qq] = smem[s + qq];
                    }
                }
            }
        }
    }
    if (tid < R) bulk_wait_read0();    // the CTA's shared memory must outlive its bulk stores
}

template <int C, int R>
int launch_tiled(const uint8_t* src, uint8_t* dst, int n_img, int H, int W, int Ho, int Wo,
                 const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    // one consumer thread per output column; strips as wide as possible up to kMaxThreads
    const int n_strips = (Wo + kMaxThreads - 1) / kMaxThreads;
    int Wt = ((Wo + n_strips - 1) / n_strips + 31) & ~31;
    if (Wt < 32) Wt = 32;
    RemapArgs a;
    a.src = src; a.dst = dst; a.map_x = map_x; a.map_y = map_y;
    a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.map_div = map_div;
    a.n_strips = (Wo + Wt - 1) / Wt;
    a.n_rowtiles = (Ho + R - 1) / R;
    const int64_t total = (int64_t)n_img * a.n_strips * a.n_rowtiles;
    if (total > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "tiled remap: too many tiles");
    a.total_tiles = (int)total;
    a.out_pitch = (Wt * C + 15 + 15) & ~15;
    // a stage holds R + R/4 + 2 source rows at unit scale (slack for local minification)
    const int unit_pitch = (((Wt + 1) * C + 30) & ~15) + 16;
    a.stage_bytes = ((R + R / 4 + 2) * unit_pitch + 127) & ~127;
    const size_t smem_bytes = (size_t)kStages * a.stage_bytes + 2 * (size_t)R * a.out_pitch +
                              kStages * tab_bytes<R>() + 2 * kStages * sizeof(uint64_t);
    // uniform phase: every source row of every image starts at the same 16-byte phase
    const bool uniform = ((W * C) & 15) == 0 && (((int64_t)H * W * C) & 15) == 0;
    auto kern = uniform ? remap_u8_tiled_kernel<C, R, true> : remap_u8_tiled_kernel<C, R, false>;
    // the opt-in and the occupancy query cost microseconds of host time: once per configuration
    struct Cfg { size_t smem; int threads, dev, occ; };
    static thread_local Cfg cache[2] = {{0, 0, -1, 0}, {0, 0, -1, 0}};
    Cfg& c = cache[uniform ? 1 : 0];
    int dev = 0;
    AW_CUDA(cudaGetDevice(&dev));
    if (c.smem != smem_bytes || c.threads != Wt + 32 || c.dev != dev) {
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int o = 0;
        AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, Wt + 32, smem_bytes));
        c = Cfg{smem_bytes, Wt + 32, dev, o};
    }
    const int occ = c.occ;
    if (occ < 1) return fail(ATTWARP_ERR_CUDA, "tiled remap: kernel does not fit an SM (%zu B shared)", smem_bytes);
    const int64_t cap = (int64_t)sm_count() * occ;
    const int grid = (int)(total < cap ? total : cap);
    kern<<<grid, Wt + 32, smem_bytes, st>>>(a);
    return check_launch("remap_u8_tiled_kernel");
}

int rows_per_pass() {
    static const int v = [] {
        const char* e = getenv("ATTWARP_REMAP_ROWS");
        const int r = e ? atoi(e) : 12;
        return (r == 8 || r == 16) ? r : 12;
    }();
    return v;
}

template <int C>
int launch_tiled_c(const uint8_t* s, uint8_t* d, int n_img, int H, int W, int Ho, int Wo, const float* map_x,
                   const float* map_y, int map_div, cudaStream_t st) {
    switch (rows_per_pass()) {
        case 8: return launch_tiled<C, 8>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 16: return launch_tiled<C, 16>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        default: return launch_tiled<C, 12>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
    }
}

}  // namespace

// uint8 images with H, W >= 2: HWC with C in {1,3,4} or planar (n_img = B*C single-channel planes,
// map_div = C).
int launch_remap_u8_tiled(const void* src, void* dst, int n_img, int C, int H, int W, int Ho, int Wo,
                          const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    switch (C) {
        case 1: return launch_tiled_c<1>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 3: return launch_tiled_c<3>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 4: return launch_tiled_c<4>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        default: return fail(ATTWARP_ERR_UNSUPPORTED, "tiled remap supports C in {1,3,4} (got %d)", C);
    }
}

}  // namespace aw
