// remap_stream.cu -- stage 5 for uint8 images: persistent, warp-specialised, streaming resample.
//
// Same arithmetic as remap_direct_kernel (cv2.remap INTER_LINEAR + BORDER_REPLICATE, see
// warp_math.h; reference call sites "Attention Guided Warping/new_method.py:268-271",
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198").  A gather-resample of 3-byte
// pixels runs out of issue slots long before it runs out of HBM bandwidth, so the kernel is
// organised around instructions per output pixel and around never making a warp wait for another:
//
//   * work = column strips (<= 352 output columns) of images, split over a persistent grid (three CTAs
//     per SM) in units of kTileRows output rows; every CTA walks a contiguous range of units.  The units
//     it owns in one strip form a SEGMENT that is streamed top to bottom in CHUNKS of <= R (16) output
//     rows through a ring of shared-memory source stages and a ring of output tiles;
//   * the PRODUCER warp plans a chunk -- the contiguous range of source rows its output rows tap,
//     minus the (at most two) rows whose horizontal blends the consumers still hold in registers
//     from the previous chunk -- writes the chunk's row table and fetches the rows with
//     cp.async.bulk (TMA bulk copy, global -> shared): ONE copy for the whole range when the strip
//     spans full image rows that are multiples of 16 bytes, else one copy per row (the 16-byte
//     aligned span around the strip's source columns); completion lands on the stage's `full`
//     mbarrier.  Planning costs ~60 dependent instructions per chunk and reads map_y through a
//     register window refilled 32 rows ahead: the single producer warp must stay well ahead of the
//     consumer warps;
//   * CONSUMER threads own one output column each (two, half a strip apart, for 3-channel images: the
//     table loads and loop control are shared by both).  The warp map is separable, so the
//     horizontal blend of a source row is computed ONCE per (row, output column) -- a funnel-shifted
//     8-byte window and dp4a with byte-positioned weights -- and kept packed with the previous
//     row's blend as the two 16-bit halves of one register; each output row is then ONE dp2a
//     (vertical blend + rounding constant) and a shift per channel.  The packed pair is carried
//     across chunks, so no source row is fetched or blended twice.  Consumer warps never
//     synchronise with each other: a warp that finishes a chunk arrives on the chunk's mbarriers
//     and moves on to the next one;
//   * the STORE warp waits until every consumer warp has written a chunk's output tile, ships it
//     with cp.async.bulk (shared -> global; byte stores for ragged row ends) and frees the tile
//     once it has been read.
//
// Maps need not be monotone: a chunk ends before the first output row that taps an earlier
// source row than its predecessor, so arbitrary maps degrade to one-row chunks (same arithmetic).
// A strip whose source span does not fit two rows of a stage is gathered straight from global
// memory by the consumers, so the kernel is total.
#include <stdlib.h>

#include "bulk_ptx.cuh"
#include "common.cuh"

namespace aw {
namespace {

using namespace ptx;

// Tuned on B200 at 336^2 and 1344^2 (profiles/run_tune.sh, run_tune2.sh): 2 source stages x 16 rows, 2 output
// tiles, 3 CTAs per SM (72 KB each).  3 x 12 rows is 1-2 % slower (a quarter more chunk prologues), 4 CTAs per
// SM with 2 x 12 / 2 x 10 / 3 x 8 rows 2-8 % slower, a third output tile (3 x 10 + 3) 3 % slower.
#ifndef AW_SRC_STAGES
#define AW_SRC_STAGES 2
#endif
#ifndef AW_OUT_STAGES
#define AW_OUT_STAGES 2
#endif
#ifndef AW_ROWS
#define AW_ROWS 16
#endif
#ifndef AW_MIN_CTAS
#define AW_MIN_CTAS 3
#endif
#ifndef AW_SKIP_SCAN
#define AW_SKIP_SCAN 1
#endif
constexpr int kSrcStages = AW_SRC_STAGES;          // source-row stages (chunks whose loads are in flight) per CTA
constexpr int kOutStages = AW_OUT_STAGES;          // output tiles per CTA
// output columns per strip: 352, the widest strip whose stages let three CTAs share an SM.  One column per
// thread: 11 + 2 warps (x 3 CTAs keeps 48 registers); two columns per thread: 6 + 2 warps.
constexpr int max_cols(int) { return 352; }
constexpr int max_threads(int cpt) { return (max_cols(cpt) + 32 * cpt - 1) / (32 * cpt) * 32 + 64; }
constexpr int kRoleThreads = 64;       // producer warp + store warp
// Work is split over the CTAs in units of kTileRows output rows (finer than a chunk: with ~16 chunks per CTA
// a split in whole chunks leaves the slowest CTA 5 % more work than the average)
#ifndef AW_TILE_ROWS
#define AW_TILE_ROWS 1
#endif
constexpr int kTileRows = AW_TILE_ROWS;

// Shared memory is addressed as byte offsets from the one dynamic array below.
extern __shared__ __align__(128) uint8_t smem[];
__device__ __forceinline__ uint32_t ld32(int off) { return *reinterpret_cast<const uint32_t*>(smem + off); }
__device__ __forceinline__ uint4 ld128(int off) { return *reinterpret_cast<const uint4*>(smem + off); }
__device__ __forceinline__ void st32(int off, uint32_t v) { *reinterpret_cast<uint32_t*>(smem + off) = v; }
__device__ __forceinline__ void st128(int off, uint4 v) { *reinterpret_cast<uint4*>(smem + off) = v; }
__device__ __forceinline__ void st8(int off, uint32_t v) { smem[off] = (uint8_t)v; }

// ---- rows of the per-stage chunk table (the header is described next to the kernel) -----------------
//   uint4 row[R + 1]:  x = dp2a weight word  wa | (32 - wa) << 8  (upper row, lower row)
//                      y = byte offset of the row inside the output tile
//                          (i * out_pitch + (dst address & 15))
//                      z = slot after which the row is emitted (= slot of its LOWER tap);
//                          0xffffffff: both taps are the carried pair, emit before slot 0;
//                          entry n_rows is a sentinel (z = kRowSentinel)
//                      w = 512, the rounding constant of the vertical blend (as a literal it would be
//                          rematerialised inside the row loop; here it arrives with the row entry)
//   uint32 slot_base[2 R]:  slot * slot_pitch + (global address of its first byte & 15)
constexpr int kTabRows = 64;
constexpr uint32_t kRowSentinel = 0x7fffffffu;
template <int R> constexpr int tab_slots() { return kTabRows + 16 * (R + 1); }
template <int R> constexpr int tab_bytes() { return tab_slots<R>() + 4 * 2 * R + 16; }   // + the entry after the last slot
constexpr uint32_t kFlagNewStrip = 1u, kFlagUniform = 2u;
constexpr int kNoCarry = -(1 << 29);

// base source column and tap weights of one output column (border replicate folded into weights)
__device__ __forceinline__ void column_taps(float m, int W, int& xb, int& w0, int& w1) {
    const int sx = quantise_coord(m);
    const int ix = sx >> 5, ax = sx & 31;
    if (ix < 0) { xb = 0; w0 = 32; w1 = 0; }
    else if (ix >= W - 1) { xb = W - 2; w0 = 0; w1 = 32; }
    else { xb = ix; w0 = 32 - ax; w1 = ax; }
}

// Horizontal blend of the C channels of one output column on one staged source row (generic C).
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
            h[0] = __dp4a(A, wA[0], 0u);
            h[1] = __dp4a(Bv, wB[1], __dp4a(A, wA[1], 0u));
            h[2] = __dp4a(Bv, wB[2], __dp4a(A, wA[2], 0u));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __dp4a(Bv, wB[k], __dp4a(A, wA[k], 0u));
        }
    }
}
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

// ---- the sweep over one chunk, hand-scheduled in PTX ---------------------------------------------
// The table-driven loops branch on values loaded from shared memory.  They are uniform over the
// warp, but the compiler cannot prove it and would wrap every branch in reconvergence bookkeeping;
// PTX with `bra.uni` and explicit shared-space addresses avoids that.
//   P*        : packed (previous row blend | this row blend << 16) per channel, carried across chunks
//   cur/arena : shared-space address of the window (U: word holding its first byte in slot 0;
//               !U: start of the staged span + this column's window offset)
//   sh        : U only, 8 * (window byte offset inside its word)
//   sp        : !U only, shared-space address of slot_base[0]
//   rp        : shared-space address of row[0];  ocol: of this column inside the output tile
// Order: rows with z == -1 (no new slot), then per slot: blend, shift into P, emit its rows.
#define AW_SW_ROWCTL_BEGIN                                      \
    "setp.ne.u32 q, ez, s;\n"                                   \
    "@q bra.uni NEXT;\n"                                        \
    "ROW:\n"
#define AW_SW_ROWCTL_END(NSLOTS)                                \
    "add.u32 rp, rp, 16;\n"                                     \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"                \
    "setp.eq.u32 q, ez, s;\n"                                   \
    "@q bra.uni ROW;\n"                                         \
    "NEXT:\n"                                                   \
    "add.s32 s, s, 1;\n"                                        \
    "setp.lt.s32 p, s, " NSLOTS ";\n"
// rows whose two taps are the carried pair: emitted before the first slot of the chunk
#define AW_SW_PRE_BEGIN                                         \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"                \
    "setp.ne.u32 q, ez, 0xffffffff;\n"                          \
    "@q bra.uni PRE_DONE;\n"                                    \
    "PRE_ROW:\n"
#define AW_SW_PRE_END(NSLOTS)                                   \
    "add.u32 rp, rp, 16;\n"                                     \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"                \
    "setp.eq.u32 q, ez, 0xffffffff;\n"                          \
    "@q bra.uni PRE_ROW;\n"                                     \
    "PRE_DONE:\n"                                               \
    "mov.b32 s, 0;\n"                                           \
    "setp.lt.s32 p, s, " NSLOTS ";\n"                           \
    "@!p bra.uni DONE;\n"
#define AW_SW_WINDOW_U                                          \
    "ld.shared.b32 lo, [cur];\n"                                \
    "ld.shared.b32 mid, [cur+4];\n"
#define AW_SW_WINDOW_T(ARENA)                                   \
    "ld.shared.b32 t, [sp];\n"                                  \
    "add.u32 t, t, " ARENA ";\n"                                \
    "and.b32 cur, t, 0xfffffffc;\n"                             \
    "shl.b32 sh, t, 3;\n"                                       \
    "add.u32 sp, sp, 4;\n"                                      \
    "ld.shared.b32 lo, [cur];\n"                                \
    "ld.shared.b32 mid, [cur+4];\n"
#define AW_SW_EMIT3                                             \
    "dp2a.lo.u32.u32 r0, %0, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r1, %1, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r2, %2, ex, ew;\n"                         \
    "add.u32 o, ey, %8;\n"                                      \
    "shr.u32 r0, r0, 10;\n shr.u32 r1, r1, 10;\n shr.u32 r2, r2, 10;\n" \
    "st.shared.u8 [o], r0;\n st.shared.u8 [o+1], r1;\n st.shared.u8 [o+2], r2;\n"
#define AW_SW_BLEND3(SH)                                        \
    "ld.shared.b32 hi, [cur+8];\n"                              \
    "shf.r.wrap.b32 A, lo, mid, " SH ";\n"                      \
    "shf.r.wrap.b32 B, mid, hi, " SH ";\n"                      \
    "dp4a.u32.u32 h0, A, %9, 0;\n"                              \
    "dp4a.u32.u32 t, A, %10, 0;\n"                              \
    "dp4a.u32.u32 h1, B, %12, t;\n"                             \
    "dp4a.u32.u32 t, A, %11, 0;\n"                              \
    "dp4a.u32.u32 h2, B, %13, t;\n"                             \
    "prmt.b32 %0, %0, h0, 0x5432;\n"                            \
    "prmt.b32 %1, %1, h1, 0x5432;\n"                            \
    "prmt.b32 %2, %2, h2, 0x5432;\n"
#define AW_SW_EMIT1                                             \
    "dp2a.lo.u32.u32 r0, %0, ex, ew;\n"                         \
    "add.u32 o, ey, %6;\n"                                      \
    "shr.u32 r0, r0, 10;\n"                                     \
    "st.shared.u8 [o], r0;\n"
#define AW_SW_BLEND1(SH)                                        \
    "shf.r.wrap.b32 A, lo, mid, " SH ";\n"                      \
    "dp4a.u32.u32 h0, A, %7, 0;\n"                              \
    "prmt.b32 %0, %0, h0, 0x5432;\n"

template <bool U>
__device__ __forceinline__ void sweep_c3(uint32_t* P, int n_slots, uint32_t cur_or_arena, uint32_t sh_or_sp,
                                         uint32_t pitch, uint32_t rp, uint32_t ocol, const uint32_t* wA,
                                         const uint32_t* wB, uint32_t rnd) {
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, hi, A, B, h0, h1, h2, t, r0, r1, r2, o, cur, rp, ex, ey, ez, ew;\n"
            "mov.b32 cur, %4;\n mov.b32 rp, %7;\n" AW_SW_PRE_BEGIN AW_SW_EMIT3 AW_SW_PRE_END("%3")
            "SLOT:\n" AW_SW_WINDOW_U AW_SW_BLEND3("%5")
            "add.u32 cur, cur, %6;\n" AW_SW_ROWCTL_BEGIN AW_SW_EMIT3 AW_SW_ROWCTL_END("%3")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1]), "+r"(P[2])
            : "r"(n_slots), "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]),
              "r"(wA[1]), "r"(wA[2]), "r"(wB[1]), "r"(wB[2]), "r"(rnd)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, hi, A, B, h0, h1, h2, t, r0, r1, r2, o, cur, sh, sp, rp, ex, ey, ez, ew;\n"
            "mov.b32 sp, %5;\n mov.b32 rp, %7;\n" AW_SW_PRE_BEGIN AW_SW_EMIT3 AW_SW_PRE_END("%3")
            "SLOT:\n" AW_SW_WINDOW_T("%4") AW_SW_BLEND3("sh") AW_SW_ROWCTL_BEGIN AW_SW_EMIT3 AW_SW_ROWCTL_END("%3")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1]), "+r"(P[2])
            : "r"(n_slots), "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]),
              "r"(wA[1]), "r"(wA[2]), "r"(wB[1]), "r"(wB[2]), "r"(rnd)
            : "memory");
    }
}

template <bool U>
__device__ __forceinline__ void sweep_c1(uint32_t* P, int n_slots, uint32_t cur_or_arena, uint32_t sh_or_sp,
                                         uint32_t pitch, uint32_t rp, uint32_t ocol, const uint32_t* wA,
                                         uint32_t rnd) {
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, A, h0, t, r0, o, cur, rp, ex, ey, ez, ew;\n"
            "mov.b32 cur, %2;\n mov.b32 rp, %5;\n" AW_SW_PRE_BEGIN AW_SW_EMIT1 AW_SW_PRE_END("%1")
            "SLOT:\n" AW_SW_WINDOW_U AW_SW_BLEND1("%3")
            "add.u32 cur, cur, %4;\n" AW_SW_ROWCTL_BEGIN AW_SW_EMIT1 AW_SW_ROWCTL_END("%1")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0])
            : "r"(n_slots), "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]), "r"(rnd)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            ".reg .b32 s, lo, mid, A, h0, t, r0, o, cur, sh, sp, rp, ex, ey, ez, ew;\n"
            "mov.b32 sp, %3;\n mov.b32 rp, %5;\n" AW_SW_PRE_BEGIN AW_SW_EMIT1 AW_SW_PRE_END("%1")
            "SLOT:\n" AW_SW_WINDOW_T("%2") AW_SW_BLEND1("sh") AW_SW_ROWCTL_BEGIN AW_SW_EMIT1 AW_SW_ROWCTL_END("%1")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0])
            : "r"(n_slots), "r"(cur_or_arena), "r"(sh_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(wA[0]), "r"(rnd)
            : "memory");
    }
}

// Two output columns per thread (C = 3), half a strip apart so that a warp's accesses to shared
// memory keep the conflict-free 3-byte lane stride: the table loads and the loop control (a third of
// the one-column sweep) are shared by both columns and each thread carries two independent chains.
//   bdelta : byte distance of column b from column a inside an output row;  bvalid: column b exists
//   ca/cb : window addresses of column a / b (as cur/arena above);  sha/shb: their shift amounts (U)
#define AW_SW2_LOAD                                             \
    "ld.shared.b32 loa, [ca];\n"                                \
    "ld.shared.b32 mida, [ca+4];\n"                             \
    "ld.shared.b32 hia, [ca+8];\n"                              \
    "ld.shared.b32 lob, [cb];\n"                                \
    "ld.shared.b32 midb, [cb+4];\n"                             \
    "ld.shared.b32 hib, [cb+8];\n"
#define AW_SW2_ALIGN(SHA, SHB)                                  \
    "shf.r.wrap.b32 Aa, loa, mida, " SHA ";\n"                  \
    "shf.r.wrap.b32 Ba, mida, hia, " SHA ";\n"                  \
    "shf.r.wrap.b32 Ab, lob, midb, " SHB ";\n"                  \
    "shf.r.wrap.b32 Bb, midb, hib, " SHB ";\n"
#define AW_SW2_DOT                                              \
    "dp4a.u32.u32 h0, Aa, %14, 0;\n"                            \
    "dp4a.u32.u32 t, Aa, %15, 0;\n"                             \
    "dp4a.u32.u32 h1, Ba, %17, t;\n"                            \
    "dp4a.u32.u32 t, Aa, %16, 0;\n"                             \
    "dp4a.u32.u32 h2, Ba, %18, t;\n"                            \
    "dp4a.u32.u32 h3, Ab, %19, 0;\n"                            \
    "dp4a.u32.u32 t, Ab, %20, 0;\n"                             \
    "dp4a.u32.u32 h4, Bb, %22, t;\n"                            \
    "dp4a.u32.u32 t, Ab, %21, 0;\n"                             \
    "dp4a.u32.u32 h5, Bb, %23, t;\n"                            \
    "prmt.b32 %0, %0, h0, 0x5432;\n"                            \
    "prmt.b32 %1, %1, h1, 0x5432;\n"                            \
    "prmt.b32 %2, %2, h2, 0x5432;\n"                            \
    "prmt.b32 %3, %3, h3, 0x5432;\n"                            \
    "prmt.b32 %4, %4, h4, 0x5432;\n"                            \
    "prmt.b32 %5, %5, h5, 0x5432;\n"
#define AW_SW2_BLEND(SHA, SHB) AW_SW2_LOAD AW_SW2_ALIGN(SHA, SHB) AW_SW2_DOT
// non-uniform phase: window addresses and shifts of the next slot from slot_base[] (sp walks the table)
#define AW_SW2_ADDR_T                                           \
    "ld.shared.b32 t, [sp];\n"                                  \
    "add.u32 sp, sp, 4;\n"                                      \
    "add.u32 ta, t, %7;\n"                                      \
    "add.u32 tb, t, %8;\n"                                      \
    "and.b32 ca, ta, 0xfffffffc;\n"                             \
    "and.b32 cb, tb, 0xfffffffc;\n"                             \
    "shl.b32 sha, ta, 3;\n"                                     \
    "shl.b32 shb, tb, 3;\n"
#define AW_SW2_EMIT                                             \
    "dp2a.lo.u32.u32 r0, %0, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r1, %1, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r2, %2, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r3, %3, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r4, %4, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r5, %5, ex, ew;\n"                         \
    "add.u32 o, ey, %12;\n"                                     \
    "add.u32 o2, o, %25;\n"                                     \
    "shr.u32 r0, r0, 10;\n shr.u32 r1, r1, 10;\n shr.u32 r2, r2, 10;\n" \
    "shr.u32 r3, r3, 10;\n shr.u32 r4, r4, 10;\n shr.u32 r5, r5, 10;\n" \
    "st.shared.u8 [o], r0;\n st.shared.u8 [o+1], r1;\n st.shared.u8 [o+2], r2;\n" \
    "@pb st.shared.u8 [o2], r3;\n @pb st.shared.u8 [o2+1], r4;\n @pb st.shared.u8 [o2+2], r5;\n"

template <bool U>
__device__ __forceinline__ void sweep_c3x2(uint32_t* P, int n_slots, uint32_t ca, uint32_t cb, uint32_t sha_or_sp,
                                           uint32_t shb, uint32_t pitch, uint32_t rp, uint32_t ocol,
                                           const uint32_t* wA, const uint32_t* wB, uint32_t rnd,
                                           uint32_t bdelta, uint32_t bvalid) {
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q, pb;\n"
            ".reg .b32 s, loa, mida, hia, lob, midb, hib, Aa, Ba, Ab, Bb, h0, h1, h2, h3, h4, h5, t;\n"
            ".reg .b32 r0, r1, r2, r3, r4, r5, o, o2, ca, cb, rp, ex, ey, ez, ew;\n"
            "setp.ne.u32 pb, %26, 0;\n"
            "mov.b32 ca, %7;\n mov.b32 cb, %8;\n mov.b32 rp, %11;\n" AW_SW_PRE_BEGIN AW_SW2_EMIT AW_SW_PRE_END("%6")
            // software pipeline: the window of slot s + 1 is requested before slot s is blended (the slot
            // after the last one is read too -- still inside the CTA's shared memory -- and never used)
            AW_SW2_LOAD
            "SLOT:\n" AW_SW2_ALIGN("%9", "%24")
            "add.u32 ca, ca, %10;\n add.u32 cb, cb, %10;\n" AW_SW2_LOAD AW_SW2_DOT
            AW_SW_ROWCTL_BEGIN AW_SW2_EMIT AW_SW_ROWCTL_END("%6")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1]), "+r"(P[2]), "+r"(P[3]), "+r"(P[4]), "+r"(P[5])
            : "r"(n_slots), "r"(ca), "r"(cb), "r"(sha_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(rnd),
              "r"(wA[0]), "r"(wA[1]), "r"(wA[2]), "r"(wB[1]), "r"(wB[2]),
              "r"(wA[3]), "r"(wA[4]), "r"(wA[5]), "r"(wB[4]), "r"(wB[5]), "r"(shb), "r"(bdelta), "r"(bvalid)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q, pb;\n"
            ".reg .b32 s, loa, mida, hia, lob, midb, hib, Aa, Ba, Ab, Bb, h0, h1, h2, h3, h4, h5, t, ta, tb, sha, shb;\n"
            ".reg .b32 r0, r1, r2, r3, r4, r5, o, o2, ca, cb, sp, rp, ex, ey, ez, ew;\n"
            "setp.ne.u32 pb, %26, 0;\n"
            "mov.b32 sp, %9;\n mov.b32 rp, %11;\n" AW_SW_PRE_BEGIN AW_SW2_EMIT AW_SW_PRE_END("%6")
            // software pipeline as in the uniform variant: the window of slot s + 1 is addressed and requested
            // before slot s is blended (slot_base has a harmless entry after the last slot)
            AW_SW2_ADDR_T AW_SW2_LOAD
            "SLOT:\n" AW_SW2_ALIGN("sha", "shb") AW_SW2_ADDR_T AW_SW2_LOAD AW_SW2_DOT
            AW_SW_ROWCTL_BEGIN AW_SW2_EMIT AW_SW_ROWCTL_END("%6")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1]), "+r"(P[2]), "+r"(P[3]), "+r"(P[4]), "+r"(P[5])
            : "r"(n_slots), "r"(ca), "r"(cb), "r"(sha_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(rnd),
              "r"(wA[0]), "r"(wA[1]), "r"(wA[2]), "r"(wB[1]), "r"(wB[2]),
              "r"(wA[3]), "r"(wA[4]), "r"(wA[5]), "r"(wB[4]), "r"(wB[5]), "r"(shb), "r"(bdelta), "r"(bvalid)
            : "memory");
    }
}

// Two output columns per thread for single-channel rows (grey images and the planes of CHW images), half a
// strip apart like sweep_c3x2: the row-table loads and the loop control are most of the one-column sweep
// when a column is one byte, so sharing them nearly halves the instructions per output byte.
#define AW_SW1X2_LOAD                                           \
    "ld.shared.b32 loa, [ca];\n"                                \
    "ld.shared.b32 mida, [ca+4];\n"                             \
    "ld.shared.b32 lob, [cb];\n"                                \
    "ld.shared.b32 midb, [cb+4];\n"
#define AW_SW1X2_ALIGN(SHA, SHB)                                \
    "shf.r.wrap.b32 Aa, loa, mida, " SHA ";\n"                  \
    "shf.r.wrap.b32 Ab, lob, midb, " SHB ";\n"
#define AW_SW1X2_DOT                                            \
    "dp4a.u32.u32 h0, Aa, %10, 0;\n"                            \
    "dp4a.u32.u32 h1, Ab, %11, 0;\n"                            \
    "prmt.b32 %0, %0, h0, 0x5432;\n"                            \
    "prmt.b32 %1, %1, h1, 0x5432;\n"
#define AW_SW1X2_EMIT                                           \
    "dp2a.lo.u32.u32 r0, %0, ex, ew;\n"                         \
    "dp2a.lo.u32.u32 r1, %1, ex, ew;\n"                         \
    "add.u32 o, ey, %8;\n"                                      \
    "add.u32 o2, o, %13;\n"                                     \
    "shr.u32 r0, r0, 10;\n shr.u32 r1, r1, 10;\n"               \
    "st.shared.u8 [o], r0;\n"                                   \
    "@pb st.shared.u8 [o2], r1;\n"
#define AW_SW1X2_ADDR_T                                         \
    "ld.shared.b32 t, [sp];\n"                                  \
    "add.u32 sp, sp, 4;\n"                                      \
    "add.u32 ta, t, %3;\n"                                      \
    "add.u32 tb, t, %4;\n"                                      \
    "and.b32 ca, ta, 0xfffffffc;\n"                             \
    "and.b32 cb, tb, 0xfffffffc;\n"                             \
    "shl.b32 sha, ta, 3;\n"                                     \
    "shl.b32 shb, tb, 3;\n"

template <bool U>
__device__ __forceinline__ void sweep_c1x2(uint32_t* P, int n_slots, uint32_t ca, uint32_t cb, uint32_t sha_or_sp,
                                           uint32_t shb, uint32_t pitch, uint32_t rp, uint32_t ocol,
                                           const uint32_t* wA, uint32_t rnd, uint32_t bdelta, uint32_t bvalid) {
    // operands: %0 %1 P | %2 n_slots | %3 ca | %4 cb | %5 sha / sp | %6 pitch | %7 rp | %8 ocol | %9 rnd |
    //           %10 %11 weights | %12 shb | %13 bdelta | %14 bvalid
    if (U) {
        asm volatile(
            "{\n"
            ".reg .pred p, q, pb;\n"
            ".reg .b32 s, loa, mida, lob, midb, Aa, Ab, h0, h1, t, r0, r1, o, o2, ca, cb, rp, ex, ey, ez, ew;\n"
            "setp.ne.u32 pb, %14, 0;\n"
            "mov.b32 ca, %3;\n mov.b32 cb, %4;\n mov.b32 rp, %7;\n" AW_SW_PRE_BEGIN AW_SW1X2_EMIT AW_SW_PRE_END("%2")
            AW_SW1X2_LOAD
            "SLOT:\n" AW_SW1X2_ALIGN("%5", "%12")
            "add.u32 ca, ca, %6;\n add.u32 cb, cb, %6;\n" AW_SW1X2_LOAD AW_SW1X2_DOT
            AW_SW_ROWCTL_BEGIN AW_SW1X2_EMIT AW_SW_ROWCTL_END("%2")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1])
            : "r"(n_slots), "r"(ca), "r"(cb), "r"(sha_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(rnd),
              "r"(wA[0]), "r"(wA[1]), "r"(shb), "r"(bdelta), "r"(bvalid)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p, q, pb;\n"
            ".reg .b32 s, loa, mida, lob, midb, Aa, Ab, h0, h1, t, ta, tb, sha, shb, r0, r1, o, o2, ca, cb, sp, rp;\n"
            ".reg .b32 ex, ey, ez, ew;\n"
            "setp.ne.u32 pb, %14, 0;\n"
            "mov.b32 sp, %5;\n mov.b32 rp, %7;\n" AW_SW_PRE_BEGIN AW_SW1X2_EMIT AW_SW_PRE_END("%2")
            AW_SW1X2_ADDR_T AW_SW1X2_LOAD
            "SLOT:\n" AW_SW1X2_ALIGN("sha", "shb") AW_SW1X2_ADDR_T AW_SW1X2_LOAD AW_SW1X2_DOT
            AW_SW_ROWCTL_BEGIN AW_SW1X2_EMIT AW_SW_ROWCTL_END("%2")
            "@p bra.uni SLOT;\n"
            "DONE:\n"
            "}\n"
            : "+r"(P[0]), "+r"(P[1])
            : "r"(n_slots), "r"(ca), "r"(cb), "r"(sha_or_sp), "r"(pitch), "r"(rp), "r"(ocol), "r"(rnd),
              "r"(wA[0]), "r"(wA[1]), "r"(shb), "r"(bdelta), "r"(bvalid)
            : "memory");
    }
}

struct StreamArgs {
    // uniform batch (imgs == nullptr): n_img dense images of one shape, maps [n_img / map_div][..]
    const uint8_t* src;
    uint8_t* dst;
    const float* map_x;
    const float* map_y;
    int H, W, Ho, Wo;
    int map_div;             // CHW planes share their image's maps
    int n_strips, n_rowtiles;
    int strip_cols;          // output columns per strip (<= consumer threads x CPT)
    // ragged batch: per-image descriptors, n_img + 1 entries (the last one only carries unit_begin)
    const RaggedImage* imgs;
    int n_img;
    int total_units;         // length of the cost axis (uniform batch: one unit per tile)
    int stage_bytes;         // bytes of the source-row arena of one stage (multiple of 128)
    int out_pitch;           // bytes per row of an output tile
    int rnd;                 // 512, the rounding constant of the vertical blend (kept out of the
                             // instruction stream: as a literal it is rematerialised inside the row loop)
    int dbg;                 // ATTWARP_REMAP_DBG experiments: 1 skip the sweep, 2 skip the tile stores
};

// What the kernel needs to know about one image.
struct View {
    const uint8_t* src;
    uint8_t* dst;
    const float* mx;
    const float* my;
    int H, W, Ho, Wo, n_strips, strip_cols, n_rowtiles, unit_begin, tile_units;
};
template <int C>
__device__ __forceinline__ View get_view(const StreamArgs& a, int img) {
    View v;
    if (a.imgs != nullptr) {
        const uint4* q = reinterpret_cast<const uint4*>(a.imgs + img);
        const uint4 p0 = __ldg(q), p1 = __ldg(q + 1), p2 = __ldg(q + 2), p3 = __ldg(q + 3);
        v.src = reinterpret_cast<const uint8_t*>(((uint64_t)p0.y << 32) | p0.x);
        v.dst = reinterpret_cast<uint8_t*>(((uint64_t)p0.w << 32) | p0.z);
        v.mx = reinterpret_cast<const float*>(((uint64_t)p1.y << 32) | p1.x);
        v.my = reinterpret_cast<const float*>(((uint64_t)p1.w << 32) | p1.z);
        v.H = (int)p2.x; v.W = (int)p2.y; v.Ho = (int)p2.z; v.Wo = (int)p2.w;
        v.n_strips = (int)(p3.x & 0xffffu); v.tile_units = (int)(p3.x >> 16);
        v.strip_cols = (int)p3.y; v.n_rowtiles = (int)p3.z; v.unit_begin = (int)p3.w;
    } else {
        const int mrow = img / a.map_div;
        v.src = a.src + (int64_t)img * a.H * a.W * C;
        v.dst = a.dst + (int64_t)img * a.Ho * a.Wo * C;
        v.mx = a.map_x + (int64_t)mrow * a.Wo;
        v.my = a.map_y + (int64_t)mrow * a.Ho;
        v.H = a.H; v.W = a.W; v.Ho = a.Ho; v.Wo = a.Wo;
        v.n_strips = a.n_strips; v.strip_cols = a.strip_cols; v.n_rowtiles = a.n_rowtiles;
        v.unit_begin = img * a.n_strips * a.n_rowtiles;
        v.tile_units = 1;
    }
    return v;
}

// ---- per-stage chunk table (byte offsets), written by the producer, read by the consumers ----------
//   +0   uint4 {n_rows, n_slots | flags << 16, slot_pitch, phase of slot 0 (global address & 15)}
//              n_rows 0: output row y0 takes the direct path; -1: stop
//   +16  uint4 {img, x_first, y0, c_lo}
//   +32  uint4 {address of the chunk's first output byte (lo, hi), bytes per tile row, Wo * C}
//              (copied to the tile header for the store warp)
//   +48  uint4 {address of map_x[x_first] (lo, hi), W, columns in the strip}      (new strip only)
//   +64  uint4 row[R + 1]: see kTabRows
//   then uint32 slot_base[2 R]  (non-uniform phase only)
constexpr int kTabStore = 32, kTabStrip = 48;

// Requires H >= 2 and W >= 2 for every image (the launchers route degenerate images to the direct kernel).
// blockDim.x = Wt / CPT consumer threads (CPT output columns each) + 32 producer threads +
// 32 store threads.
// Uniform phase (per image): W*C is a multiple of 16, so every staged row starts at the same 16-byte
// phase and slot k sits at k * slot_pitch + phase -- the sweep advances by one add per slot.
//
// Shared memory: [kSrcStages source arenas][kOutStages output tiles][kSrcStages chunk tables]
//                [kOutStages tile headers][mbarriers].
// Chunk c lives in source stage c % kSrcStages and output tile c % kOutStages.  mbarriers:
//   full[s]  producer -> consumers   table written, source rows landed (transaction bytes)
//   sfree[s] consumers -> producer   every consumer warp is done with the stage's rows and table
//   odone[o] consumers -> store warp every consumer warp has written its columns of the tile
//   ofree[o] store warp -> consumers the tile has been read out of shared memory
// The two rings are decoupled so that the loads of chunk c + kSrcStages start as soon as the
// consumers leave chunk c, without waiting for its tile to be shipped.
template <int C, int R, int CPT>
__global__ void __launch_bounds__(max_threads(CPT), CPT == 2 ? AW_MIN_CTAS : 3)
remap_u8_stream_kernel(const StreamArgs a) {
    static_assert(CPT == 1 || (CPT == 2 && (C == 3 || C == 1)), "two columns per thread are implemented for C = 1 and 3");
    const int Wt = ((int)blockDim.x - kRoleThreads) * CPT;
    const int tid = threadIdx.x;
    const int out_bytes = R * a.out_pitch;
    const int out_off0 = kSrcStages * a.stage_bytes;
    const int tab_off0 = out_off0 + kOutStages * out_bytes;
    const int ohdr_off0 = tab_off0 + kSrcStages * tab_bytes<R>();
    const int bar_off0 = ohdr_off0 + kOutStages * 32;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t full_s = smem_s + (uint32_t)bar_off0;
    const uint32_t sfree_s = full_s + 8u * kSrcStages;
    const uint32_t odone_s = sfree_s + 8u * kSrcStages;
    const uint32_t ofree_s = odone_s + 8u * kOutStages;
    const int n_cons_warps = ((int)blockDim.x - kRoleThreads) >> 5;

    if (tid == 0) {
        for (int s = 0; s < kSrcStages; ++s) {
            mbar_init(full_s + 8u * s, 1);
            mbar_init(sfree_s + 8u * s, n_cons_warps);
        }
        for (int s = 0; s < kOutStages; ++s) {
            mbar_init(odone_s + 8u * s, n_cons_warps);
            mbar_init(ofree_s + 8u * s, 1);
        }
        mbar_init_fence();
    }
    __syncthreads();

    // contiguous, balanced range of the cost axis for this CTA: it owns the tiles that START inside it
    const int u0 = (int)(((int64_t)a.total_units * blockIdx.x) / gridDim.x);
    const int u1 = (int)(((int64_t)a.total_units * (blockIdx.x + 1)) / gridDim.x);

    // warp-uniform role split (the shuffle tells the compiler the branch does not diverge)
    const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;

    // Role split nested under ONE warp-uniform branch: with two sibling branches ptxas stops treating
    // the consumers' table-driven control flow as uniform and wraps every sweep branch in BSSY/BSYNC.
    if (warp_idx >= n_cons_warps) {
        if (warp_idx == n_cons_warps) {
            // =========================== producer warp =========================================
            int st = 0;
            uint32_t ph = 0;                 // parity of the stage's current use
            // first image: the last one whose first unit is <= u0 (32-ary search over the table)
            int img;
            if (a.imgs == nullptr) {
                img = u0 / (a.n_strips * a.n_rowtiles);
            } else {
                int lo = 0, hi = a.n_img;                              // answer in [lo, hi)
                while (hi - lo > 1) {
                    const int step = (hi - lo + 31) / 32;
                    const int probe = min(lo + (lane + 1) * step, hi);
                    const bool le = probe < hi && __ldg(&a.imgs[probe].unit_begin) <= u0;
                    const int k = __popc(__ballot_sync(0xffffffffu, le));       // probes are monotone
                    const int nlo = lo + k * step;
                    hi = min(lo + (k + 1) * step, hi);
                    lo = nlo;
                }
                img = lo;
            }
            View v = get_view<C>(a, img);
            int local = (u0 - v.unit_begin + v.tile_units - 1) / v.tile_units;   // first tile starting at >= u0
            for (;;) {
                if (local >= v.n_strips * v.n_rowtiles) {              // next image
                    if (++img >= a.n_img) break;
                    v = get_view<C>(a, img);
                    local = 0;
                }
                if (v.unit_begin + local * v.tile_units >= u1) break;
                // ---- segment: the tiles [local, local_end) of one (image, strip) ----------------
                const int H = v.H, W = v.W, Ho = v.Ho, Wo = v.Wo;
                const int strip = local / v.n_rowtiles, rt = local % v.n_rowtiles;
                const int mine_end = (u1 - v.unit_begin + v.tile_units - 1) / v.tile_units;   // tiles starting before u1
                const int local_end = min((strip + 1) * v.n_rowtiles, mine_end);
                const int y_end = min(Ho, (rt + (local_end - local)) * kTileRows);
                const int x_first = strip * v.strip_cols;
                const int ncols = min(v.strip_cols, Wo - x_first);
                const uint8_t* simg = v.src;
                const uintptr_t dimg = reinterpret_cast<uintptr_t>(v.dst);
                const float* my = v.my;
                const float* mx = v.mx + x_first;
                // source column span of the strip
                int c_lo, row_bytes, slot_pitch, max_slots;
                // A single strip over an image whose whole rows (multiples of 16 bytes) fit a stage R + 2 at a
                // time: stage whole rows without looking for the span first (the maps of this library cover
                // the image, so the span is the whole row anyway; the scan is 11 dependent-latency global
                // loads on the critical path of the CTA's first chunk).
                if (AW_SKIP_SCAN && v.n_strips == 1 && ((W * C) & 15) == 0 && (a.stage_bytes - 32) / (W * C) >= R + 2) {
                    c_lo = 0;
                    row_bytes = W * C;
                    slot_pitch = row_bytes;
                    max_slots = 0;      // both set below (one_copy)
                } else {
                    int lo = 0x7fffffff, hi = -1;
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
                const uint8_t* scol = simg + (int64_t)c_lo * C;
                const int64_t row_pitch = (int64_t)W * C;
                // phase of the staged rows (uniform: identical for every row of the image)
                const int phase = (int)(reinterpret_cast<uintptr_t>(scol) & 15);
                const bool uni = ((W * C) & 15) == 0;
                // Full-width strip of an image whose rows are multiples of 16 bytes: consecutive source
                // rows are contiguous in global memory, so a whole chunk is ONE bulk copy (slot pitch =
                // row pitch).  Otherwise one copy per source row, issued by the lane that owns the slot.
                const bool one_copy = uni && row_bytes == W * C;
                if (one_copy) {
                    slot_pitch = W * C;
                    max_slots = min((a.stage_bytes - 32) / slot_pitch, 2 * R);
                }
                uint32_t seg_flags = kFlagNewStrip | (uni ? kFlagUniform : 0u);
                int carry_row = kNoCarry;    // source row whose blend sits in the upper half of P
                int y_cur = rt * kTileRows;
                // map_y is read through a register window of 2 x 32 rows (lane i holds rows y_win + i
                // and y_win + 32 + i) refilled 32 rows ahead of use: a global-load latency per chunk
                // on the planning path would cap the whole CTA at one chunk per microsecond.
                int y_win = y_cur;
                int sy_cur = quantise_coord(__ldg(my + min(y_win + lane, Ho - 1)));
                int sy_nxt = quantise_coord(__ldg(my + min(y_win + 32 + lane, Ho - 1)));
                while (y_cur < y_end) {
                    const int tab = tab_off0 + st * tab_bytes<R>();
                    const uint32_t stage_s = smem_s + (uint32_t)(st * a.stage_bytes);
                    if (y_cur - y_win >= 32) {
                        y_win += 32;
                        sy_cur = sy_nxt;
                        sy_nxt = quantise_coord(__ldg(my + min(y_win + 32 + lane, Ho - 1)));
                    }
                    // ---- plan: lane i <-> output row y_cur + i --------------------------------
                    const int y = y_cur + lane;
                    const bool live = y < y_end && lane < R;
                    const int wsel = y - y_win;                       // 0 .. 31 + R - 1
                    const int sy_a = __shfl_sync(0xffffffffu, sy_cur, wsel & 31);
                    const int sy_b = __shfl_sync(0xffffffffu, sy_nxt, wsel & 31);
                    int ra = 0x3fffffff, wa = 32;       // upper source row (lower = ra + 1), its weight
                    if (live) {
                        const int sy = wsel < 32 ? sy_a : sy_b;
                        const int iy = sy >> 5, ay = sy & 31;
                        if (iy < 0) { ra = 0; wa = 32; }
                        else if (iy >= H - 1) { ra = H - 2; wa = 0; }
                        else { ra = iy; wa = 32 - ay; }
                    }
                    // The chunk stages the CONTIGUOUS source rows r_lo .. ra(last) + 1 and takes output
                    // rows while they run in non-decreasing source order and the range fits the stage.
                    // When its first row starts inside the pair (carry_row - 1, carry_row) whose blends
                    // the consumers still hold, staging continues after that pair.
                    const int prev_ra = __shfl_up_sync(0xffffffffu, ra, 1);
                    const int r0 = __shfl_sync(0xffffffffu, ra, 0);
                    const int r_lo = (r0 == carry_row - 1 || r0 == carry_row) ? carry_row + 1 : r0;
                    const int need = ra + 2 - r_lo;                  // slots up to and including this row's taps
                    const unsigned bad = __ballot_sync(0xffffffffu, !live || (lane > 0 && ra < prev_ra) || need > max_slots);
                    const int n_rows = max_slots >= 2 ? (bad ? (__ffs(bad) - 1) : 32) : 0;
                    const int ra_last = __shfl_sync(0xffffffffu, ra, max(n_rows - 1, 0));
                    const int n_slots = n_rows > 0 ? ra_last + 2 - r_lo : 0;
                    // lane j stages source row r_lo + j into slot j
                    const uint8_t* p = scol + (int64_t)(r_lo + lane) * row_pitch;
                    const int off = (int)(reinterpret_cast<uintptr_t>(p) & 15);
                    const uint32_t bytes = lane < n_slots ? (uint32_t)((off + row_bytes + 15) & ~15) : 0u;
                    const uint32_t tx = one_copy ? (uint32_t)((phase + n_slots * slot_pitch + 15) & ~15)
                                                 : __reduce_add_sync(0xffffffffu, bytes);
                    const uintptr_t gd = dimg + (uintptr_t)(((int64_t)y * Wo + x_first) * C);   // this row's first byte

                    mbar_wait(sfree_s + 8u * st, ph ^ 1u);             // stage free again
                    if (lane < n_rows) {
                        st128(tab + kTabRows + 16 * lane,
                              make_uint4((uint32_t)wa | ((uint32_t)(32 - wa) << 8),
                                         (uint32_t)(lane * a.out_pitch) + (uint32_t)(gd & 15),
                                         (uint32_t)(ra + 1 - r_lo), 512u));
                    } else if (lane == n_rows) {
                        st128(tab + kTabRows + 16 * lane, make_uint4(0u, 0u, kRowSentinel, 0u));
                    }
                    if (!uni) {
                        if (lane < n_slots) st32(tab + tab_slots<R>() + 4 * lane, (uint32_t)(lane * slot_pitch + off));
                        // the two-column sweep requests one slot ahead: a valid address after the last slot
                        if (lane == 0) st32(tab + tab_slots<R>() + 4 * n_slots, 0u);
                    }
                    if (lane == 0) {
                        st128(tab, make_uint4((uint32_t)n_rows, (uint32_t)n_slots | (seg_flags << 16),
                                              (uint32_t)slot_pitch, (uint32_t)phase));
                        st128(tab + 16, make_uint4((uint32_t)img, (uint32_t)x_first, (uint32_t)y_cur, (uint32_t)c_lo));
                        st128(tab + kTabStore, make_uint4((uint32_t)gd, (uint32_t)((uint64_t)gd >> 32),
                                                          (uint32_t)(ncols * C), (uint32_t)(Wo * C)));
                        if (seg_flags & kFlagNewStrip) {
                            const uintptr_t mxa = reinterpret_cast<uintptr_t>(mx);
                            st128(tab + kTabStrip, make_uint4((uint32_t)mxa, (uint32_t)((uint64_t)mxa >> 32),
                                                              (uint32_t)W, (uint32_t)ncols));
                        }
                    }
                    seg_flags &= ~kFlagNewStrip;
                    __syncwarp();
                    if (lane == 0) {
                        if (n_slots > 0) mbar_arrive_expect_tx(full_s + 8u * st, tx);
                        else mbar_arrive(full_s + 8u * st);
                    }
                    __syncwarp();
                    if (one_copy) {
                        if (lane == 0 && n_slots > 0) bulk_g2s(stage_s, p - off, tx, full_s + 8u * st);
                    } else if (lane < n_slots) {
                        bulk_g2s(stage_s + (uint32_t)(lane * slot_pitch), p - off, bytes, full_s + 8u * st);
                    }
                    // the pair the consumers hold after this chunk: rows (carry_row - 1, carry_row)
                    carry_row = n_rows > 0 ? ra_last + 1 : kNoCarry;
                    y_cur += max(n_rows, 1);
                    if (++st == kSrcStages) { st = 0; ph ^= 1u; }
                }
                local = local_end;
            }
            // terminator
            mbar_wait(sfree_s + 8u * st, ph ^ 1u);
            if (lane == 0) {
                st128(tab_off0 + st * tab_bytes<R>(), make_uint4(0xffffffffu, 0u, 0u, 0u));
                mbar_arrive(full_s + 8u * st);
            }
        } else {
            // =============================== store warp ======================================
            int ot = 0;
            uint32_t ph = 0;
            for (;;) {
                mbar_wait(odone_s + 8u * ot, ph);                      // every consumer warp is through
                const uint4 hd = ld128(ohdr_off0 + 32 * ot);           // {n_rows, -, -, -}
                const int n_rows = (int)hd.x;
                if (n_rows < 0) break;
                if (n_rows > 0 && !(a.dbg & 2)) {
                    // ---- ship the rows: bulk store for the 16-byte aligned interior, bytes for the ends
                    const uint4 hs = ld128(ohdr_off0 + 32 * ot + 16);  // {dst lo, dst hi, row bytes, dst pitch}
                    const int len = (int)hs.z;
                    const int64_t dpitch = (int64_t)hs.w;
                    const int obuf = out_off0 + ot * out_bytes;
                    uint8_t* g0 = reinterpret_cast<uint8_t*>(((uint64_t)hs.y << 32) | hs.x);
                    const bool ragged = ((reinterpret_cast<uintptr_t>(g0) | (uintptr_t)len | (uintptr_t)dpitch) & 15) != 0;
                    if (lane < n_rows) {
                        uint8_t* g = g0 + (int64_t)lane * dpitch;
                        const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                        const int head = (16 - off) & 15;
                        const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                        if (body > 0)
                            bulk_s2g(g + head, smem_s + (uint32_t)(obuf + lane * a.out_pitch + off + head), (uint32_t)body);
                    }
                    bulk_commit();
                    if (ragged) {
                        // <= 15 head bytes and <= 15 tail bytes per row, one lane per byte
                        for (int i = 0; i < n_rows; ++i) {
                            uint8_t* g = g0 + (int64_t)i * dpitch;
                            const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                            const int head = min((16 - off) & 15, len);
                            const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                            const int s = obuf + i * a.out_pitch + off;
                            if (lane < 16) {
                                if (lane < head) g[lane] = smem[s + lane];
                            } else {
                                const int qq = head + body + (lane - 16);
                                if (qq < len) g[qq] = smem[s + qq];
                            }
                        }
                    }
                    bulk_wait_read0();                                 // the tile has left shared memory
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(ofree_s + 8u * ot);        // tile free for the consumers
                if (++ot == kOutStages) { ot = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // =============================== consumer warps ==============================================
    // this thread's output columns inside the strip: xl, and for CPT = 2 also xl + Wt / 2
    const int xl = tid;
    const int xstep = Wt / CPT;
    int wo[CPT];                      // byte offset of each column's window inside a staged row span
    bool xvalid = false;
    uint32_t bvalid = 0u;             // CPT = 2: the second column exists in this strip
    uint32_t wA[C * CPT], wB[C * CPT], P[C * CPT];
#pragma unroll
    for (int k = 0; k < C * CPT; ++k) wA[k] = wB[k] = P[k] = 0u;
#pragma unroll
    for (int j = 0; j < CPT; ++j) wo[j] = 0;
    const int out_col = xl * C;
    const uint32_t rnd = (uint32_t)a.rnd;

    int st = 0, ot = 0;               // source stage / output tile of the current chunk
    uint32_t sph = 0u, oph = 1u;      // parities to wait for: stage filled / tile shipped and free
    bool warp_b = false;              // CPT = 2: some lane of this warp owns a second column
    for (;; st = st + 1 == kSrcStages ? 0 : st + 1, sph ^= st == 0 ? 1u : 0u,
            ot = ot + 1 == kOutStages ? 0 : ot + 1, oph ^= ot == 0 ? 1u : 0u) {
        const int tab = tab_off0 + st * tab_bytes<R>();
        mbar_wait(full_s + 8u * st, sph);
        const uint4 h0 = ld128(tab);
        const int n_rows = (int)h0.x;
        mbar_wait(ofree_s + 8u * ot, oph);                                       // tile shipped and free
        if (n_rows < 0) {                                                        // pass the stop on
            if (tid == 0) st128(ohdr_off0 + 32 * ot, make_uint4(0xffffffffu, 0u, 0u, 0u));
            __syncwarp();
            if (lane == 0) mbar_arrive(odone_s + 8u * ot);
            break;
        }
        if (h0.y & (kFlagNewStrip << 16)) {                  // new strip: per-column taps and weights
            const uint4 h1 = ld128(tab + 16);
            const uint4 hx = ld128(tab + kTabStrip);
            const float* mx = reinterpret_cast<const float*>(((uint64_t)hx.y << 32) | hx.x);
            const int W = (int)hx.z, ncols = (int)hx.w;
            xvalid = xl < ncols;
            bvalid = (CPT == 2 && xl + xstep < ncols) ? 1u : 0u;
            warp_b = __any_sync(0xffffffffu, bvalid != 0u);
            int xb = (int)h1.w;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                // a column past the strip's end keeps the first column's window and gets zero weights
                int w0 = j == 0 ? 32 : 0, w1 = 0;
                if (xl + j * xstep < ncols) column_taps(__ldg(mx + xl + j * xstep), W, xb, w0, w1);
                wo[j] = (xb - (int)h1.w) * C;
                // dp4a weight words: tap0 of channel k at byte k of the 8-byte window, tap1 at byte k+C
#pragma unroll
                for (int k = 0; k < C; ++k) {
                    wA[j * C + k] = (uint32_t)w0 << (8 * k);
                    wB[j * C + k] = 0u;
                    if (k + C < 4) wA[j * C + k] |= (uint32_t)w1 << (8 * (k + C));
                    else wB[j * C + k] = (uint32_t)w1 << (8 * (k + C - 4));
                }
            }
        }
        if (tid == 0) {                                      // what the store warp needs to ship the tile
            st128(ohdr_off0 + 32 * ot, make_uint4(h0.x, 0u, 0u, 0u));
            st128(ohdr_off0 + 32 * ot + 16, ld128(tab + kTabStore));
        }
        if (n_rows == 0) {
            // ---- direct path for one output row whose source span does not fit a stage ----------
            if (xvalid) {
                const uint4 h1 = ld128(tab + 16);
                const int img = (int)h1.x, x_first = (int)h1.y, y0 = (int)h1.z;
                const View v = get_view<C>(a, img);
                const int H = v.H, W = v.W, Wo = v.Wo;
                const int ncols = min(v.strip_cols, Wo - x_first);
                const int sy = quantise_coord(__ldg(v.my + y0));
                const int ay = sy & 31;
                const int ya = clampi(sy >> 5, 0, H - 1), yb = clampi((sy >> 5) + 1, 0, H - 1);
                for (int j = 0; j < CPT && xl + j * xstep < ncols; ++j) {
                    const int sx = quantise_coord(__ldg(v.mx + x_first + xl + j * xstep));
                    const int ax = sx & 31;
                    const int x0 = clampi(sx >> 5, 0, W - 1), x1 = clampi((sx >> 5) + 1, 0, W - 1);
                    uint8_t* o = v.dst + ((int64_t)y0 * Wo + x_first + xl + j * xstep) * C;
#pragma unroll
                    for (int k = 0; k < C; ++k)
                        o[k] = bilinear_u8(__ldg(v.src + ((int64_t)ya * W + x0) * C + k),
                                           __ldg(v.src + ((int64_t)ya * W + x1) * C + k),
                                           __ldg(v.src + ((int64_t)yb * W + x0) * C + k),
                                           __ldg(v.src + ((int64_t)yb * W + x1) * C + k), ax, ay);
                }
            }
        } else if (xvalid && !(a.dbg & 1)) {
            const int n_slots = (int)(h0.y & 0xffffu);
            const bool uni = (h0.y & (kFlagUniform << 16)) != 0u;
            const int arena = st * a.stage_bytes + wo[0];
            const int ocol = out_off0 + ot * out_bytes + out_col;
            const uint32_t ocol_s = smem_s + (uint32_t)ocol;
            const uint32_t rp_s = smem_s + (uint32_t)(tab + kTabRows);
            const int slot_tab = tab + tab_slots<R>();
            const int win0 = arena + (uni ? (int)h0.w : 0);
            if (CPT == 2 && C == 1) {
                const int win1 = st * a.stage_bytes + wo[CPT - 1] + (uni ? (int)h0.w : 0);
                if (!warp_b) {
                    if (uni) sweep_c1<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h0.z, rp_s, ocol_s, wA, rnd);
                    else sweep_c1<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA, rnd);
                } else if (uni) {
                    sweep_c1x2<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), smem_s + (uint32_t)(win1 & ~3), (uint32_t)win0 << 3, (uint32_t)win1 << 3, h0.z, rp_s, ocol_s, wA, rnd, (uint32_t)(xstep * C), bvalid);
                } else {
                    sweep_c1x2<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)win1, smem_s + (uint32_t)slot_tab, 0u, 0u, rp_s, ocol_s, wA, rnd, (uint32_t)(xstep * C), bvalid);
                }
            } else if (CPT == 2 && !warp_b) {
                // no lane of this warp has a second column in this strip (the strip is narrower than the
                // consumer threads x 2): the one-column sweep does the same work in 3/5 of the instructions
                if (uni) sweep_c3<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h0.z, rp_s, ocol_s, wA, wB, rnd);
                else sweep_c3<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA, wB, rnd);
            } else if (CPT == 2) {
                const int win1 = st * a.stage_bytes + wo[CPT - 1] + (uni ? (int)h0.w : 0);
                if (uni) sweep_c3x2<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), smem_s + (uint32_t)(win1 & ~3), (uint32_t)win0 << 3, (uint32_t)win1 << 3, h0.z, rp_s, ocol_s, wA, wB, rnd, (uint32_t)(xstep * C), bvalid);
                else sweep_c3x2<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)win1, smem_s + (uint32_t)slot_tab, 0u, 0u, rp_s, ocol_s, wA, wB, rnd, (uint32_t)(xstep * C), bvalid);
            } else if (C == 3) {
                if (uni) sweep_c3<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h0.z, rp_s, ocol_s, wA, wB, rnd);
                else sweep_c3<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA, wB, rnd);
            } else if (C == 1) {
                if (uni) sweep_c1<true>(P, n_slots, smem_s + (uint32_t)(win0 & ~3), (uint32_t)win0 << 3, h0.z, rp_s, ocol_s, wA, rnd);
                else sweep_c1<false>(P, n_slots, smem_s + (uint32_t)win0, smem_s + (uint32_t)slot_tab, 0u, rp_s, ocol_s, wA, rnd);
            } else {
                int rp = tab + kTabRows;
                uint4 re = ld128(rp);
                for (int s = -1; s < n_slots; ++s) {
                    if (s >= 0) {
                        uint32_t h[C];
                        const int win = uni ? win0 + s * (int)h0.z : arena + (int)ld32(slot_tab + 4 * s);
                        hblend_row<C>(win & ~3, (uint32_t)win << 3, wA, wB, h);
#pragma unroll
                        for (int k = 0; k < C; ++k) P[k] = __byte_perm(P[k], h[k], 0x5432);
                    }
                    while (re.z == (uint32_t)s) {
                        vblend_store<C>(P, re.x, ocol + (int)re.y);
                        rp += 16;
                        re = ld128(rp);
                    }
                }
            }
        }
        // publish this warp's part of the tile to the async proxy, then count the warp in
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(sfree_s + 8u * st);
            mbar_arrive(odone_s + 8u * ot);
        }
    }
}

// Strip geometry of one image: as few, equal strips as possible, a multiple of 16 columns so that
// strip boundaries keep the 16-byte phase of a row.
struct StripPlan { int n_strips, strip_cols; };
inline StripPlan plan_strips(int Wo, int max_cols_) {
    StripPlan p;
    p.n_strips = (Wo + max_cols_ - 1) / max_cols_;
    p.strip_cols = (((Wo + p.n_strips - 1) / p.n_strips) + 15) & ~15;
    p.n_strips = (Wo + p.strip_cols - 1) / p.strip_cols;
    return p;
}

// Shared-memory geometry for strips of at most `cols` columns, and the launch itself.
template <int C, int R, int CPT>
int launch_kernel(StreamArgs& a, int cols, cudaStream_t st) {
    const int Wt = (cols + 32 * CPT - 1) / (32 * CPT) * (32 * CPT);   // columns covered by the consumer threads
    a.rnd = 512;
    {
        const char* e = getenv("ATTWARP_REMAP_DBG");
        a.dbg = e ? atoi(e) : 0;
    }
    a.out_pitch = (cols * C + 15 + 15) & ~15;              // + the 16-byte phase of the destination
    // an arena holds R + 2 source rows at unit scale (first chunk of a segment: R + 1)
    const int unit_pitch = (((cols + 1) * C + 30) & ~15) + 16;
    a.stage_bytes = ((R + 2) * unit_pitch + 127) & ~127;
    const size_t smem_bytes = (size_t)kSrcStages * (a.stage_bytes + tab_bytes<R>()) +
                              (size_t)kOutStages * ((size_t)R * a.out_pitch + 32) +
                              2 * (kSrcStages + kOutStages) * sizeof(uint64_t);
    auto kern = remap_u8_stream_kernel<C, R, CPT>;
    const int threads = Wt / CPT + kRoleThreads;
    // the opt-in and the occupancy query cost microseconds of host time: once per configuration
    struct Cfg { size_t smem; int threads, dev, occ; };
    static thread_local Cfg c = {0, 0, -1, 0};
    int dev = 0;
    AW_CUDA(cudaGetDevice(&dev));
    if (c.smem != smem_bytes || c.threads != threads || c.dev != dev) {
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int o = 0;
        AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem_bytes));
        // ATTWARP_REMAP_CTAS_PER_SM caps the persistent grid (leaves shared memory for a kernel of another
        // stream to co-run on the same SMs)
        if (const char* e = getenv("ATTWARP_REMAP_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v < o) o = v; }
        c = Cfg{smem_bytes, threads, dev, o};
    }
    if (c.occ < 1) return fail(ATTWARP_ERR_CUDA, "remap: kernel does not fit an SM (%zu B shared)", smem_bytes);
    const int64_t cap = (int64_t)sm_count() * c.occ;
    const int grid = (int)(a.total_units < cap ? a.total_units : cap);
    kern<<<grid, threads, smem_bytes, st>>>(a);
    return check_launch("remap_u8_stream_kernel");
}

template <int C, int R, int CPT>
int launch_stream(const uint8_t* src, uint8_t* dst, int n_img, int H, int W, int Ho, int Wo,
                  const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    StreamArgs a;
    const StripPlan sp = plan_strips(Wo, max_cols(CPT));
    a.n_strips = sp.n_strips;
    a.strip_cols = sp.strip_cols;
    a.src = src; a.dst = dst; a.map_x = map_x; a.map_y = map_y;
    a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.map_div = map_div;
    a.n_rowtiles = (Ho + kTileRows - 1) / kTileRows;
    a.imgs = nullptr;
    a.n_img = n_img;
    const int64_t total = (int64_t)n_img * a.n_strips * a.n_rowtiles;
    if (total > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: too many tiles");
    a.total_units = (int)total;
    return launch_kernel<C, R, CPT>(a, Wo < a.strip_cols ? Wo : a.strip_cols, st);
}

// Ragged batch, step 1: strip plan and cost prefix of host[0..n) (host[n] only carries the total), upload.
// A tile weighs as many units as it keeps consumer warps busy, so that CTAs splitting the unit axis evenly
// get even work whatever the mix of strip widths.
template <int R, int CPT>
int ragged_prepare(RaggedImage* host, int n, RaggedImage* dev_table, cudaStream_t st) {
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        const StripPlan sp = plan_strips(host[i].Wo, max_cols(CPT));
        const int cols = host[i].Wo < sp.strip_cols ? host[i].Wo : sp.strip_cols;
        const int units = (cols + 32 * CPT - 1) / (32 * CPT);
        if (sp.n_strips > 0xffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: image %d is too wide", i);
        host[i].strips_units = sp.n_strips | (units << 16);
        host[i].strip_cols = sp.strip_cols;
        host[i].n_rowtiles = (host[i].Ho + kTileRows - 1) / kTileRows;
        host[i].unit_begin = (int)total;
        total += (int64_t)sp.n_strips * host[i].n_rowtiles * units;
        if (total > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: too many tiles");
    }
    host[n] = RaggedImage{};
    host[n].unit_begin = (int)total;
    AW_CUDA(cudaMemcpyAsync(dev_table, host, sizeof(RaggedImage) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    return ATTWARP_OK;
}
// Step 2: the launch.
template <int C, int R, int CPT>
int ragged_run(const RaggedImage* host, int n, const RaggedImage* dev_table, cudaStream_t st) {
    int max_strip = 0;
    for (int i = 0; i < n; ++i) {
        const int cols = host[i].Wo < host[i].strip_cols ? host[i].Wo : host[i].strip_cols;
        if (cols > max_strip) max_strip = cols;
    }
    StreamArgs a{};
    a.map_div = 1;
    a.imgs = dev_table;
    a.n_img = n;
    a.total_units = host[n].unit_begin;
    if (a.total_units == 0) return ATTWARP_OK;
    return launch_kernel<C, R, CPT>(a, max_strip, st);
}

// ATTWARP_REMAP_CPT=1 forces one column per thread for C = 1 and 3 (A/B comparisons).
int columns_per_thread_c3() {
    static const int v = [] {
        const char* e = getenv("ATTWARP_REMAP_CPT");
        return (e && atoi(e) == 1) ? 1 : 2;
    }();
    return v;
}

}  // namespace

// uint8 images with H, W >= 2: HWC with C in {1,3,4} or planar (n_img = B*C single-channel planes,
// map_div = C).
int launch_remap_u8_stream(const void* src, void* dst, int n_img, int C, int H, int W, int Ho, int Wo,
                           const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    switch (C) {
        case 1:
            if (columns_per_thread_c3() == 2) return launch_stream<1, AW_ROWS, 2>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
            return launch_stream<1, AW_ROWS, 1>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 3:
            if (columns_per_thread_c3() == 2) return launch_stream<3, AW_ROWS, 2>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
            return launch_stream<3, AW_ROWS, 1>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        case 4: return launch_stream<4, AW_ROWS, 1>(s, d, n_img, H, W, Ho, Wo, map_x, map_y, map_div, st);
        default: return fail(ATTWARP_ERR_UNSUPPORTED, "remap supports C in {1,3,4} (got %d)", C);
    }
}

// Ragged batch of HWC uint8 images (every H, W >= 2), see common.cuh.
int launch_remap_u8_stream_ragged_prepare(RaggedImage* h, int n, int C, RaggedImage* d, cudaStream_t st) {
    if ((C == 3 || C == 1) && columns_per_thread_c3() == 2) return ragged_prepare<AW_ROWS, 2>(h, n, d, st);
    return ragged_prepare<AW_ROWS, 1>(h, n, d, st);
}
int launch_remap_u8_stream_ragged_run(const RaggedImage* h, int n, int C, const RaggedImage* d, cudaStream_t st) {
    switch (C) {
        case 1:
            if (columns_per_thread_c3() == 2) return ragged_run<1, AW_ROWS, 2>(h, n, d, st);
            return ragged_run<1, AW_ROWS, 1>(h, n, d, st);
        case 3:
            if (columns_per_thread_c3() == 2) return ragged_run<3, AW_ROWS, 2>(h, n, d, st);
            return ragged_run<3, AW_ROWS, 1>(h, n, d, st);
        case 4: return ragged_run<4, AW_ROWS, 1>(h, n, d, st);
        default: return fail(ATTWARP_ERR_UNSUPPORTED, "remap supports C in {1,3,4} (got %d)", C);
    }
}

}  // namespace aw
