// remap_quad.cu -- stage 5 for 3-channel interleaved uint8 images (the benchmark format): persistent,
// warp-specialised streaming resample, FOUR ADJACENT output pixels per consumer thread.
//
// Same arithmetic as remap_direct_kernel (cv2.remap INTER_LINEAR + BORDER_REPLICATE, see warp_math.h;
// reference call sites "Attention Guided Warping/new_method.py:268-271",
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198").  Same skeleton as remap_stream.cu (producer
// warp plans chunks of output rows and fetches the source rows they tap with cp.async.bulk, consumer warps
// sweep them once), with the consumer side rebuilt around the two resources the round-1 kernel ran out of --
// shared-memory wavefronts and issue slots (profiles/README.md) -- and WITHOUT output tiles or store warps: the
// consumer warps write their rows to global memory themselves:
//
//   * a thread owns output columns 4t .. 4t+3 of its strip (QUAD mapping: its four source windows sit 3 words
//     apart at unit scale, 12-byte lane stride: every 32-bit load of a warp is conflict-free) or columns t, t + 32,
//     t + 64, t + 96 of its warp's 128-column block (LANE mapping, chosen per warp and strip where the map's local
//     scale would make the QUAD windows collide);
//   * each lane stores the 12 bytes of its four adjacent pixels itself: three 32-bit stores at a 12-byte lane
//     stride (a warp's three stores cover 384 contiguous bytes; L2 merges the partial sectors).  Destination rows
//     at any byte alignment are handled by a second build (MODE 2: one halo pixel per warp, lane-local funnel
//     shifts + one shuffle, byte stores only at the two ends of a row), so no destination needs an output tile;
//   * the horizontal blends of the two most recent source rows are held per channel in an EVEN-row and an
//     ODD-row register (source row r goes to E when r is even), so a new source row overwrites one of them
//     without re-packing; the vertical blend of an output row is  t = wO*O + (wE*E + 512 * 2^14)  with the row's
//     weights pre-shifted by 14 bits -- two 32-bit multiply-adds (the sum stays below 2^32) whose TOP BYTE is the
//     output byte ((v + 512) >> 10 with nothing to shift or mask: prmt picks the top bytes when packing);
//   * the horizontal blend of a pixel is three dp4a: channel 0 on the aligned window word itself (bytes p0c0 p0c1
//     p0c2 p1c0, weights w0 . . w1), channels 1 and 2 on one prmt of the window (p0c1 p1c1 p0c2 p1c2);
//   * source rows are staged at a UNIFORM shared-memory pitch whatever the alignment of the image: a strip that
//     spans whole rows is fetched with ONE bulk copy per chunk (global rows are contiguous; the shared-memory
//     image is the global one shifted by a multiple of 16 bytes), narrower strips with one copy per row into
//     slots whose pitch is congruent to the row pitch modulo 16.  When the pitch is a multiple of 4 the
//     per-pixel window addresses advance by a constant and their byte shifts never change (fixed-shift sweep);
//     otherwise address and shift are re-derived per slot (three more ALU operations per pixel and slot);
//   * a CTA has exactly the consumer warps its strips need (3 .. 16: strips of up to 2048 columns), so a
//     1344-wide image is processed in whole rows fetched as contiguous 4 KB rows, and the images of a ragged batch
//     are grouped by that number.
//
// Maps need not be monotone (a chunk ends before the first output row that taps an earlier source row than its
// predecessor), a strip whose source span does not fit a stage is gathered from global memory (the kernel is
// total), ragged batches run in one launch over a descriptor table.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <utility>
#include <vector>

#include "bulk_ptx.cuh"
#include "common.cuh"

namespace aw {
namespace {

using namespace ptx;

constexpr int kMaxRing = 4;            // source-row stages per CTA are a launch parameter (2 .. 4)
constexpr int kMaxRows = 30;          // capacity of a chunk table (a lane plans a row, one carries the sentinel); the rows per chunk are a launch parameter
constexpr int kC = 3;

extern __shared__ __align__(128) uint8_t smem[];
__device__ __forceinline__ uint4 ld128(int off) { return *reinterpret_cast<const uint4*>(smem + off); }
__device__ __forceinline__ void st128(int off, uint4 v) { *reinterpret_cast<uint4*>(smem + off) = v; }

// ---- per-stage chunk table (byte offsets), written by the producer, read by the consumers ----------
//   +0   uint4 {n_rows, n_slots | flags << 16, slot_pitch, byte offset of slot 0's first byte in the arena}
//              n_rows 0: output row y0 takes the direct path; -1: stop
//   +16  uint4 {img, x_first, y0, c_lo}
//   +32  uint4 {the 4-byte aligned address at or below the image's first destination byte (lo, hi), bytes per strip
//               row, Wo * 3}
//   +48  uint4 {address of map_x[x_first] (lo, hi), W, columns in the strip}      (new strip only)
//   +64  uint4 row[kMaxRows + 1]:  x = wE << 14, y = wO << 14  (weights of the even / odd source row)
//                                  z = offset of the row's first destination byte (of this strip) from the address
//                                      in +32: its low two bits are the row's misalignment
//                                  w = slot after which the row is emitted (= slot of its LOWER tap);
//                                      0xffffffff: both taps are the carried pair, emit before slot 0;
//                                      the entry after the last row is a sentinel
constexpr int kTabStore = 32, kTabStrip = 48, kTabRows = 64;
constexpr int kTabBytes = kTabRows + 16 * (kMaxRows + 1);
constexpr uint32_t kRowSentinel = 0x7fffffffu;
constexpr uint32_t kFlagNewStrip = 1u, kFlagFixedShift = 2u, kFlagOddFirst = 8u;
constexpr int kNoCarry = -(1 << 29);

// base source column and tap weights of one output column (border replicate folded into weights)
__device__ __forceinline__ void column_taps(float m, int W, int& xb, int& w0, int& w1) {
    const int sx = quantise_coord(m);
    const int ix = sx >> 5, ax = sx & 31;
    if (ix < 0) { xb = 0; w0 = 32; w1 = 0; }
    else if (ix >= W - 1) { xb = W - 2; w0 = 0; w1 = 32; }
    else { xb = ix; w0 = 32 - ax; w1 = ax; }
}

// ---- the sweep over one chunk, hand-scheduled in PTX ---------------------------------------------
// Table-driven loops branch on values loaded from shared memory; they are uniform over the warp, but the compiler
// cannot prove it: PTX with bra.uni avoids the reconvergence bookkeeping.  Per pixel j in {a,b,c,d}:
//   E?0..2 / O?0..2 : horizontal blends of the latest even / odd source row (carried across chunks)
//   w?l / w?h       : dp4a weight words  w0 | w1 << 8  and the same shifted by 16
//   FIXED:  k? = shared address of the word holding the window's first byte in slot 0, s? = 8 * byte offset
//   !FIXED: u? = shared BYTE address of the window in slot 0 (word address and shift derived per slot)
// Common: n_slots, pitch (bytes between slots), rp (shared address of row[0]), first slot parity.
#define AWQ_LOAD1(J)                                            \
    "ld.shared.b32 lo" #J ", [k" #J "];\n"                      \
    "ld.shared.b32 mi" #J ", [k" #J "+4];\n"                    \
    "ld.shared.b32 hi" #J ", [k" #J "+8];\n"
#define AWQ_LOAD AWQ_LOAD1(a) AWQ_LOAD1(b) AWQ_LOAD1(c) AWQ_LOAD1(d)
// address of the next slot's window (fixed shift: one add; else byte address -> word address + shift)
#define AWQ_ADDR_F1(J) "add.u32 k" #J ", k" #J ", %36;\n"
#define AWQ_ADDR_V1(J)                                          \
    "add.u32 u" #J ", u" #J ", %36;\n"                          \
    "and.b32 k" #J ", u" #J ", 0xfffffffc;\n"                   \
    "shl.b32 n" #J ", u" #J ", 3;\n"
#define AWQ_ADDR_F AWQ_ADDR_F1(a) AWQ_ADDR_F1(b) AWQ_ADDR_F1(c) AWQ_ADDR_F1(d)
#define AWQ_ADDR_V AWQ_ADDR_V1(a) AWQ_ADDR_V1(b) AWQ_ADDR_V1(c) AWQ_ADDR_V1(d)
// align the 8-byte window: A = p0c0 p0c1 p0c2 p1c0 (channel 0 is a dp4a of A itself with the weights w0 . . w1),
// B = p1c1 p1c2 . . ; pair the taps of the other two channels: X = p0c1 p1c1 p0c2 p1c2
#define AWQ_ALIGN1(J, SH)                                       \
    "shf.r.wrap.b32 A" #J ", lo" #J ", mi" #J ", " SH ";\n"     \
    "shf.r.wrap.b32 B" #J ", mi" #J ", hi" #J ", " SH ";\n"     \
    "prmt.b32 X" #J ", A" #J ", B" #J ", 0x5241;\n"
#define AWQ_ALIGN_F AWQ_ALIGN1(a, "sa") AWQ_ALIGN1(b, "sb") AWQ_ALIGN1(c, "sc") AWQ_ALIGN1(d, "sd")
#define AWQ_ALIGN_V AWQ_ALIGN1(a, "ma") AWQ_ALIGN1(b, "mb") AWQ_ALIGN1(c, "mc") AWQ_ALIGN1(d, "md")
// the shift of the CURRENT slot must survive the address update of the next one
#define AWQ_KEEP_V "mov.b32 ma, na;\n mov.b32 mb, nb;\n mov.b32 mc, nc;\n mov.b32 md, nd;\n"
#define AWQ_DOT1(P, J)                                          \
    "dp4a.u32.u32 " #P #J "0, A" #J ", wa" #J ", 0;\n"          \
    "dp4a.u32.u32 " #P #J "1, X" #J ", wl" #J ", 0;\n"          \
    "dp4a.u32.u32 " #P #J "2, X" #J ", wh" #J ", 0;\n"
#define AWQ_DOT(P) AWQ_DOT1(P, a) AWQ_DOT1(P, b) AWQ_DOT1(P, c) AWQ_DOT1(P, d)
// vertical blend of one byte: top byte of the 32-bit  E * (wE << 14) + O * (wO << 14) + (512 << 14)
#define AWQ_V1(J, K)                                            \
    "mad.lo.u32 v" #J #K ", E" #J #K ", ex, 0x800000;\n"        \
    "mad.lo.u32 v" #J #K ", O" #J #K ", ey, v" #J #K ";\n"
#define AWQ_VBLEND                                              \
    AWQ_V1(a, 0) AWQ_V1(a, 1) AWQ_V1(a, 2) AWQ_V1(b, 0) AWQ_V1(b, 1) AWQ_V1(b, 2)  \
    AWQ_V1(c, 0) AWQ_V1(c, 1) AWQ_V1(c, 2) AWQ_V1(d, 0) AWQ_V1(d, 1) AWQ_V1(d, 2)
// Emit of one output row.  Either mapping ends with the lane holding the 12 output bytes of four ADJACENT pixels (its
// QUAD position: bytes 12 t .. 12 t + 11 of the warp's 384-byte block) in q0, q2, q4:
//   QUAD mapping: the 12 top bytes of the vertical blends packed into 3 words (prmt: x.b3 | y.b3 << 8, then the low
//                 halves of two pairs);
//   LANE mapping: the thread's four pixels are 32 columns apart (pixel j of lane t is column 32 j + t of the warp's
//                 128-column block), so that the window loads of a warp stay inside ~128 bytes and never conflict
//                 whatever the local scale of the map; they change hands through the warp's scratch as RGBX pixels
//                 (X + 4 * lane + 128 j in, 16 bytes at X + 16 * lane out, squeezed to 12 by three prmt; a
//                 bar.warp.sync before the scratch is written and one before it is read).
// Destination rows 4-byte aligned (MODE 1): the lane stores its three words itself (12-byte lane stride: the three
// stores of a warp cover 384 contiguous bytes; L2 merges the partial sectors -- measured 3 % faster at configs[1] than
// exchanging the words through shared memory for 128-byte-contiguous stores, profiles/r02w_*).  %49 = address of the
// warp's first destination byte at row offset 0 (64-bit), %35 = 12 * lane, pv = the lane owns a column of the strip.
// The row's table entry (ex, ey: weights, ez: offset) is dead once the blends and the address are computed: the NEXT
// row's entry is loaded straight over it, early enough for the loop-carried compare of `ew` (no register rotation).
#define AWQ_OWN12_BEGIN                                         \
    AWQ_VBLEND                                                  \
    "cvt.u64.u32 ro64, ez;\n"                                   \
    "add.u64 oa, %49, ro64;\n"                                  \
    "cvt.u64.u32 ro64, %35;\n"                                  \
    "add.u64 oa, oa, ro64;\n"                                   \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp+16];\n"
#define AWQ_STORE_OWN12                                         \
    "@pv st.global.b32 [oa], q0;\n"                             \
    "@pv st.global.b32 [oa+4], q2;\n"                           \
    "@pv st.global.b32 [oa+8], q4;\n"
#define AWQ_EMIT_WD                                             \
    AWQ_OWN12_BEGIN                                             \
    "prmt.b32 q0, va0, va1, 0x0073;\n"                          \
    "prmt.b32 q1, va2, vb0, 0x0073;\n"                          \
    "prmt.b32 q2, vb1, vb2, 0x0073;\n"                          \
    "prmt.b32 q3, vc0, vc1, 0x0073;\n"                          \
    "prmt.b32 q4, vc2, vd0, 0x0073;\n"                          \
    "prmt.b32 q5, vd1, vd2, 0x0073;\n"                          \
    "prmt.b32 q0, q0, q1, 0x5410;\n"                            \
    "prmt.b32 q2, q2, q3, 0x5410;\n"                            \
    "prmt.b32 q4, q4, q5, 0x5410;\n"                            \
    AWQ_STORE_OWN12
#define AWQ_EMIT_LD                                             \
    AWQ_OWN12_BEGIN                                             \
    "prmt.b32 q0, va0, va1, 0x0073;\n prmt.b32 q0, q0, va2, 0x0710;\n"  \
    "prmt.b32 q1, vb0, vb1, 0x0073;\n prmt.b32 q1, q1, vb2, 0x0710;\n"  \
    "prmt.b32 q2, vc0, vc1, 0x0073;\n prmt.b32 q2, q2, vc2, 0x0710;\n"  \
    "prmt.b32 q3, vd0, vd1, 0x0073;\n prmt.b32 q3, q3, vd2, 0x0710;\n"  \
    "bar.warp.sync 0xffffffff;\n"                               \
    "st.shared.b32 [sx], q0;\n st.shared.b32 [sx+128], q1;\n st.shared.b32 [sx+256], q2;\n st.shared.b32 [sx+384], q3;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "ld.shared.v4.b32 {q1, q3, q5, tm}, [sq];\n"                \
    "prmt.b32 q0, q1, q3, 0x4210;\n"                            \
    "prmt.b32 q2, q3, q5, 0x5421;\n"                            \
    "prmt.b32 q4, q5, tm, 0x6542;\n"                            \
    AWQ_STORE_OWN12
// Destination rows at ANY byte alignment (MODE 2).  A warp's 32 lanes x 4 pixels start ONE PIXEL BEFORE its block of
// 127 pixels: lane 0's first pixel is the last pixel of the previous warp's block (computed twice, stored once).
// With P0 = row offset of the strip (from the 4-byte aligned address below the image: `ez`) + 3 * (first column of
// the warp's lanes) (%50) and a = P0 & 3, lane t holds stream bytes P0 + 12 t .. + 11; the aligned word that starts
// a bytes before them is the funnel shift of the previous lane's last word (one shuffle) and its own first word,
// the next two of its own words.  A warp stores every aligned word whose LAST byte lies in its 127 pixels: the word
// that straddles two blocks belongs to the second one, which has all its bytes thanks to the halo pixel -- no byte
// stores between blocks.  Only the first word of a row (first warp of the first strip) and the last one (last warp
// of the last strip) can be partial: those two warps (%52 != 0) put their bytes into P and lanes 4..6 / 0..2 store
// the <= 3 + 3 bytes one by one (%60: bit a = this lane stores its edge byte at alignment a, %61 = the byte's
// offset from P0, %62 = its address in P).  %53: validity bits [4 a + m] of the lane's three words.
#define AWQ_HALO_BEGIN                                          \
    AWQ_VBLEND                                                  \
    "add.s32 up0, ez, %50;\n"                                   \
    "and.b32 uk, up0, 3;\n"                                     \
    "add.s32 tm, up0, %35;\n"                                   \
    "and.b32 tm, tm, 0xfffffffc;\n"                             \
    "cvt.s64.s32 ro64, tm;\n add.s64 oa, %49, ro64;\n"          \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp+16];\n"             \
    "shl.b32 us, uk, 3;\n sub.u32 us, 32, us;\n"                \
    "shl.b32 tm, uk, 2;\n shr.u32 tm, %53, tm;\n"               \
    "and.b32 o, tm, 1;\n setp.ne.u32 pw0, o, 0;\n"              \
    "and.b32 o, tm, 2;\n setp.ne.u32 pw1, o, 0;\n"              \
    "and.b32 o, tm, 4;\n setp.ne.u32 pw2, o, 0;\n"
#define AWQ_HALO_STORE                                          \
    "shfl.sync.up.b32 q1, q4, 1, 0, 0xffffffff;\n"              \
    "shf.r.clamp.b32 q1, q1, q0, us;\n"                         \
    "shf.r.clamp.b32 q3, q0, q2, us;\n"                         \
    "shf.r.clamp.b32 q5, q2, q4, us;\n"                         \
    "@pw0 st.global.b32 [oa], q1;\n"                            \
    "@pw1 st.global.b32 [oa+4], q3;\n"                          \
    "@pw2 st.global.b32 [oa+8], q5;\n"                          \
    "{\n"                                                       \
    "@!pew bra.uni HALO_DONE;\n"                                \
    "bar.warp.sync 0xffffffff;\n"                               \
    "st.shared.b32 [pw], q0;\n st.shared.b32 [pw+4], q2;\n st.shared.b32 [pw+8], q4;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "shr.u32 tm, %60, uk;\n and.b32 tm, tm, 1;\n setp.ne.u32 pe, tm, 0;\n" \
    "@pe ld.shared.u8 o, [pedge];\n"                            \
    "add.s32 tm, up0, %61;\n cvt.s64.s32 ro64, tm;\n add.s64 oae, %49, ro64;\n" \
    "@pe st.global.u8 [oae], o;\n"                              \
    "HALO_DONE:\n"                                              \
    "}\n"
#define AWQ_EMIT_WU                                             \
    AWQ_HALO_BEGIN                                              \
    "prmt.b32 q0, va0, va1, 0x0073;\n"                          \
    "prmt.b32 q1, va2, vb0, 0x0073;\n"                          \
    "prmt.b32 q2, vb1, vb2, 0x0073;\n"                          \
    "prmt.b32 q3, vc0, vc1, 0x0073;\n"                          \
    "prmt.b32 q4, vc2, vd0, 0x0073;\n"                          \
    "prmt.b32 q5, vd1, vd2, 0x0073;\n"                          \
    "prmt.b32 q0, q0, q1, 0x5410;\n"                            \
    "prmt.b32 q2, q2, q3, 0x5410;\n"                            \
    "prmt.b32 q4, q4, q5, 0x5410;\n"                            \
    AWQ_HALO_STORE
#define AWQ_EMIT_LU                                             \
    AWQ_HALO_BEGIN                                              \
    "prmt.b32 q0, va0, va1, 0x0073;\n prmt.b32 q0, q0, va2, 0x0710;\n"  \
    "prmt.b32 q1, vb0, vb1, 0x0073;\n prmt.b32 q1, q1, vb2, 0x0710;\n"  \
    "prmt.b32 q2, vc0, vc1, 0x0073;\n prmt.b32 q2, q2, vc2, 0x0710;\n"  \
    "prmt.b32 q3, vd0, vd1, 0x0073;\n prmt.b32 q3, q3, vd2, 0x0710;\n"  \
    "bar.warp.sync 0xffffffff;\n"                               \
    "st.shared.b32 [sx], q0;\n st.shared.b32 [sx+128], q1;\n st.shared.b32 [sx+256], q2;\n st.shared.b32 [sx+384], q3;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "ld.shared.v4.b32 {q1, q3, q5, tm}, [sq];\n"                \
    "prmt.b32 q0, q1, q3, 0x4210;\n"                            \
    "prmt.b32 q2, q3, q5, 0x5421;\n"                            \
    "prmt.b32 q4, q5, tm, 0x6542;\n"                            \
    AWQ_HALO_STORE
// rows emitted after slot s (label prefix L keeps the two unrolled halves apart).  Every emit variant also fetches the
// table entry of the row AFTER the one being emitted, early enough for the loop-carried compare (the entry after the
// sentinel is read too: still inside the CTA's shared memory, never used)
#define AWQ_ROW_STEP(EMIT)                                      \
    EMIT                                                        \
    "add.u32 rp, rp, 16;\n"
#define AWQ_ROWS(L, EMIT)                                       \
    "setp.ne.u32 q, ew, s;\n"                                   \
    "@q bra.uni " L "_NEXT;\n"                                  \
    L "_ROW:\n" AWQ_ROW_STEP(EMIT)                              \
    "setp.eq.u32 q, ew, s;\n"                                   \
    "@q bra.uni " L "_ROW;\n"                                   \
    L "_NEXT:\n"                                                \
    "add.s32 s, s, 1;\n"
#define AWQ_DECL                                                \
    ".reg .pred p, q, pv, podd;\n"                              \
    ".reg .b32 s, rp, ex, ey, ez, ew, o, sx, sq, pw, tm, uk, us, up0, pedge;\n" \
    ".reg .pred pw0, pw1, pw2, pe, pew;\n"                               \
    ".reg .b64 ro64, oa, oae;\n"                 \
    ".reg .b32 loa, mia, hia, lob, mib, hib, loc, mic, hic, lod, mid, hid;\n"   \
    ".reg .b32 Aa, Ba, Ab, Bb, Ac, Bc, Ad, Bd, Xa, Xb, Xc, Xd;\n" \
    ".reg .b32 ka, kb, kc, kd, ua, ub, uc, ud, sa, sb, sc, sd, ma, mb, mc, md, na, nb, nc, nd;\n" \
    ".reg .b32 wla, wha, wlb, whb, wlc, whc, wld, whd, waa, wab, wac, wad;\n" \
    ".reg .b32 Ea0, Ea1, Ea2, Eb0, Eb1, Eb2, Ec0, Ec1, Ec2, Ed0, Ed1, Ed2;\n"   \
    ".reg .b32 Oa0, Oa1, Oa2, Ob0, Ob1, Ob2, Oc0, Oc1, Oc2, Od0, Od1, Od2;\n"   \
    ".reg .b32 va0, va1, va2, vb0, vb1, vb2, vc0, vc1, vc2, vd0, vd1, vd2;\n"   \
    ".reg .b32 q0, q1, q2, q3, q4, q5;\n"
// operands: %0-%11 E, %12-%23 O (read/write) | %24-%27 window address (a..d) | %28-%31 shift (a..d) |
//           %32 n_slots | %33 LANE mapping: scratch address of my RGBX pixels | %34, %35 unused | %36 pitch | %37 rp |
//           %38 unused | %39 first slot odd | %40 unused | %41-%48 weight words (la, ha, lb, hb, lc, hc, ld, hd) |
//           %49-%62 see the emit variants
#define AWQ_PROLOGUE                                            \
    "mov.b32 Ea0, %0;\n mov.b32 Ea1, %1;\n mov.b32 Ea2, %2;\n mov.b32 Eb0, %3;\n mov.b32 Eb1, %4;\n mov.b32 Eb2, %5;\n" \
    "mov.b32 Ec0, %6;\n mov.b32 Ec1, %7;\n mov.b32 Ec2, %8;\n mov.b32 Ed0, %9;\n mov.b32 Ed1, %10;\n mov.b32 Ed2, %11;\n" \
    "mov.b32 Oa0, %12;\n mov.b32 Oa1, %13;\n mov.b32 Oa2, %14;\n mov.b32 Ob0, %15;\n mov.b32 Ob1, %16;\n mov.b32 Ob2, %17;\n" \
    "mov.b32 Oc0, %18;\n mov.b32 Oc1, %19;\n mov.b32 Oc2, %20;\n mov.b32 Od0, %21;\n mov.b32 Od1, %22;\n mov.b32 Od2, %23;\n" \
    "mov.b32 wla, %41;\n mov.b32 wha, %42;\n mov.b32 wlb, %43;\n mov.b32 whb, %44;\n"  \
    "mov.b32 wlc, %45;\n mov.b32 whc, %46;\n mov.b32 wld, %47;\n mov.b32 whd, %48;\n"  \
    "mov.b32 waa, %54;\n mov.b32 wab, %55;\n mov.b32 wac, %56;\n mov.b32 wad, %57;\n"  \
    "mov.b32 sx, %33;\n mov.b32 sq, %34;\n"                     \
    "mov.b32 pw, %51;\n mov.b32 pedge, %62;\n setp.ne.u32 pew, %52, 0;\n" \
    "and.b32 tm, %53, 1;\n setp.ne.u32 pw0, tm, 0;\n"           \
    "and.b32 tm, %53, 2;\n setp.ne.u32 pw1, tm, 0;\n"           \
    "and.b32 tm, %53, 4;\n setp.ne.u32 pw2, tm, 0;\n"           \
    "setp.ne.u32 pv, %40, 0;\n"                                 \
    "setp.ne.u32 podd, %39, 0;\n"                               \
    "mov.b32 rp, %37;\n"                                        \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"
#define AWQ_EPILOGUE                                            \
    "mov.b32 %0, Ea0;\n mov.b32 %1, Ea1;\n mov.b32 %2, Ea2;\n mov.b32 %3, Eb0;\n mov.b32 %4, Eb1;\n mov.b32 %5, Eb2;\n" \
    "mov.b32 %6, Ec0;\n mov.b32 %7, Ec1;\n mov.b32 %8, Ec2;\n mov.b32 %9, Ed0;\n mov.b32 %10, Ed1;\n mov.b32 %11, Ed2;\n" \
    "mov.b32 %12, Oa0;\n mov.b32 %13, Oa1;\n mov.b32 %14, Oa2;\n mov.b32 %15, Ob0;\n mov.b32 %16, Ob1;\n mov.b32 %17, Ob2;\n" \
    "mov.b32 %18, Oc0;\n mov.b32 %19, Oc1;\n mov.b32 %20, Oc2;\n mov.b32 %21, Od0;\n mov.b32 %22, Od1;\n mov.b32 %23, Od2;\n"
// rows whose two taps are the carried pair: emitted before the first slot of the chunk
#define AWQ_PRE(EMIT)                                           \
    "setp.ne.u32 q, ew, 0xffffffff;\n"                          \
    "@q bra.uni PRE_DONE;\n"                                    \
    "PRE_ROW:\n" AWQ_ROW_STEP(EMIT)                             \
    "setp.eq.u32 q, ew, 0xffffffff;\n"                          \
    "@q bra.uni PRE_ROW;\n"                                     \
    "PRE_DONE:\n"                                               \
    "mov.b32 s, 0;\n"                                           \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@!p bra.uni DONE;\n"
// The slot loop, unrolled over the parity of the source row.  The window of slot s + 1 is requested before slot s
// is blended (the slot after the last one is read too -- still inside the CTA's shared memory -- never used).
#define AWQ_BODY_F(EMIT)                                        \
    "mov.b32 ka, %24;\n mov.b32 kb, %25;\n mov.b32 kc, %26;\n mov.b32 kd, %27;\n"      \
    "mov.b32 sa, %28;\n mov.b32 sb, %29;\n mov.b32 sc, %30;\n mov.b32 sd, %31;\n"      \
    AWQ_PRE(EMIT) AWQ_LOAD                                      \
    "@podd bra.uni ODD;\n"                                      \
    "EVEN:\n" AWQ_ALIGN_F AWQ_ADDR_F AWQ_LOAD AWQ_DOT(E) AWQ_ROWS("EV", EMIT)          \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@!p bra.uni DONE;\n"                                       \
    "ODD:\n" AWQ_ALIGN_F AWQ_ADDR_F AWQ_LOAD AWQ_DOT(O) AWQ_ROWS("OD", EMIT)           \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@p bra.uni EVEN;\n"                                        \
    "DONE:\n"
#define AWQ_BODY_V(EMIT)                                        \
    "mov.b32 ua, %24;\n mov.b32 ub, %25;\n mov.b32 uc, %26;\n mov.b32 ud, %27;\n"      \
    "and.b32 ka, ua, 0xfffffffc;\n and.b32 kb, ub, 0xfffffffc;\n and.b32 kc, uc, 0xfffffffc;\n and.b32 kd, ud, 0xfffffffc;\n" \
    "shl.b32 na, ua, 3;\n shl.b32 nb, ub, 3;\n shl.b32 nc, uc, 3;\n shl.b32 nd, ud, 3;\n" \
    AWQ_PRE(EMIT) AWQ_LOAD                                      \
    "@podd bra.uni ODD;\n"                                      \
    "EVEN:\n" AWQ_KEEP_V AWQ_ADDR_V AWQ_ALIGN_V AWQ_LOAD AWQ_DOT(E) AWQ_ROWS("EV", EMIT) \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@!p bra.uni DONE;\n"                                       \
    "ODD:\n" AWQ_KEEP_V AWQ_ADDR_V AWQ_ALIGN_V AWQ_LOAD AWQ_DOT(O) AWQ_ROWS("OD", EMIT)  \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@p bra.uni EVEN;\n"                                        \
    "DONE:\n"

#define AWQ_OPERANDS                                                                                        \
    : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),       \
      "+r"(E[8]), "+r"(E[9]), "+r"(E[10]), "+r"(E[11]), "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]),     \
      "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7]), "+r"(O[8]), "+r"(O[9]), "+r"(O[10]), "+r"(O[11])      \
    : "r"(win[0]), "r"(win[1]), "r"(win[2]), "r"(win[3]), "r"(sh[0]), "r"(sh[1]), "r"(sh[2]), "r"(sh[3]),   \
      "r"(n_slots), "r"(sx), "r"(d.sq), "r"(d.lane12), "r"(pitch), "r"(rp), "r"(0), "r"(odd_first), "r"(store_ok), \
      "r"(wl[0]), "r"(wh[0]), "r"(wl[1]), "r"(wh[1]), "r"(wl[2]), "r"(wh[2]), "r"(wl[3]), "r"(wh[3]),       \
      "l"(d.obase), "r"(d.cb), "r"(d.pw), "r"(d.edge_warp), "r"(d.wmask), "r"(wa[0]), "r"(wa[1]), "r"(wa[2]),     \
      "r"(wa[3]), "r"(0), "r"(0), "r"(d.emask), "r"(d.eoff), "r"(d.pedge)                                   \
    : "memory"

// what the emit variants need (see AWQ_OWN12_BEGIN / AWQ_HALO_BEGIN / AWQ_HALO_STORE)
struct DirectOps {
    uint64_t obase;                     // MODE 1: the warp's first destination byte at row offset 0; MODE 2: the
                                        // 4-byte aligned address at or below the image's first byte
    uint32_t sq, lane12;                // scratch address of my four adjacent RGBX pixels (LANE mapping), 12 * lane
    // MODE 2 only
    uint32_t cb;                        // 3 * (strip-relative column of the warp's first lane pixel; -1 for warp 0)
    uint32_t pw;                        // scratch address of my 12 bytes in P (edge warps)
    uint32_t edge_warp;                 // this warp holds the first or the last word of the row
    uint32_t wmask;                     // validity bits [4 a + m] of the lane's three words
    uint32_t emask, eoff, pedge;        // edge byte: alignments at which this lane stores it, offset from P0, address in P
};

// FIXED: win = word addresses, sh = shifts.  !FIXED: win = byte addresses (sh unused).
// LANE: false = QUAD mapping, true = LANE mapping.  MODE 1: destination rows 4-byte aligned (three coalesced word
// stores per lane); 2: rows at any alignment (+ byte stores at the two ends of the warp's block).
template <bool FIXED, bool LANE, int MODE>
__device__ __forceinline__ void sweep_quad(uint32_t* E, uint32_t* O, const uint32_t* win, const uint32_t* sh,
                                           const uint32_t* wl, const uint32_t* wh, const uint32_t* wa, int n_slots,
                                           uint32_t pitch,
                                           uint32_t rp, uint32_t odd_first, uint32_t sx, uint32_t store_ok,
                                           const DirectOps& d) {
#define AWQ_RUN(BODY, EMIT) asm volatile("{\n" AWQ_DECL AWQ_PROLOGUE BODY(EMIT) AWQ_EPILOGUE "}\n" AWQ_OPERANDS)
    if (FIXED) {
        if (MODE == 1) { if (!LANE) AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_WD); else AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_LD); }
        else { if (!LANE) AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_WU); else AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_LU); }
    } else {
        if (MODE == 1) { if (!LANE) AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_WD); else AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_LD); }
        else { if (!LANE) AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_WU); else AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_LU); }
    }
#undef AWQ_RUN
}

struct QuadArgs {
    // uniform batch (imgs == nullptr): n_img dense images of one shape, maps [n_img][..]
    const uint8_t* src;
    uint8_t* dst;
    const float* map_x;
    const float* map_y;
    int H, W, Ho, Wo;
    int n_strips, n_rowtiles;
    int strip_cols;          // output columns per strip (<= consumer threads x 4)
    // ragged batch: per-image descriptors, n_img + 1 entries (the last one only carries unit_begin)
    const RaggedImage* imgs;
    int n_img;
    int total_units;         // length of the cost axis (uniform batch: one unit per output row of a strip)
    int stage_bytes;         // bytes of the source-row arena of one stage (multiple of 128)
    int smem_total;          // dynamic shared memory of the CTA
    int first_rows;          // rows of a CTA's first chunk (the second has twice as many, then `rows`)
    int rows;                // output rows per chunk (<= kMaxRows)
    int stages;              // ring depth: source-row stages (chunks whose loads are in flight)
    int map_policy;          // 0: per warp and strip (LANE when the map's local scale would make QUAD loads conflict),
                             // 1: always QUAD, 2: LANE wherever word stores apply
    int drift;               // map_policy 0: pixels of drift from "4 source columns per lane" that switch a warp to LANE
    int dbg;                 // ATTWARP_REMAP_DBG experiments: 1 skip the sweep (loads and pipeline only)
    unsigned long long* trace;   // ATTWARP_REMAP_TRACE: 8 global-timer stamps per CTA (nullptr: off)
};

__device__ __forceinline__ void trace_stamp(const QuadArgs& a, int slot) {
    if (a.trace != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

struct View {
    const uint8_t* src;
    uint8_t* dst;
    const float* mx;
    const float* my;
    int H, W, Ho, Wo, n_strips, strip_cols, n_rowtiles, unit_begin, tile_units;
};
__device__ __forceinline__ View get_view(const QuadArgs& a, int img) {
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
        v.src = a.src + (int64_t)img * a.H * a.W * kC;
        v.dst = a.dst + (int64_t)img * a.Ho * a.Wo * kC;
        v.mx = a.map_x + (int64_t)img * a.Wo;
        v.my = a.map_y + (int64_t)img * a.Ho;
        v.H = a.H; v.W = a.W; v.Ho = a.Ho; v.Wo = a.Wo;
        v.n_strips = a.n_strips; v.strip_cols = a.strip_cols; v.n_rowtiles = a.n_rowtiles;
        v.unit_begin = img * a.n_strips * a.n_rowtiles;
        v.tile_units = 1;
    }
    return v;
}

// the image whose tiles contain unit u0: the last one whose first unit is <= u0 (32-ary search over the table;
// called by whole warps)
__device__ __forceinline__ int first_image(const QuadArgs& a, int u0, int lane) {
    if (a.imgs == nullptr) return u0 / (a.n_strips * a.n_rowtiles);
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
    return lo;
}

// Requires H >= 2 and W >= 2 for every image (the launchers route degenerate images to the direct kernel).
// blockDim.x = consumer threads (a multiple of 32; 4 output columns each) + 32 producer threads.
// Shared memory: [stages source arenas][stages chunk tables][mbarriers][1 KB of scratch per consumer warp].
// Chunk c lives in source stage c % stages.  mbarriers:
//   full[s]  producer -> consumers   table written, source rows landed (transaction bytes)
//   sfree[s] consumers -> producer   every consumer warp is done with the stage's rows and table
// MODE 1: every destination row of the launch is 4-byte aligned; MODE 2: any alignment (see AWQ_UNALIGNED_TAIL).
template <int MAXT, int MINB, int MODE>
__global__ void __launch_bounds__(MAXT, MINB) remap_u8_quad_kernel(const QuadArgs a) {
    static_assert(MODE == 1 || MODE == 2, "MODE: 1 = aligned rows, 2 = any alignment");
    const int R = a.rows;
    const int kStages = a.stages;
    const int tab_off0 = kStages * a.stage_bytes;
    const int bar_off0 = tab_off0 + kStages * kTabBytes;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t full_s = smem_s + (uint32_t)bar_off0;
    const uint32_t sfree_s = full_s + 8u * kStages;
    const int n_cons_warps = ((int)blockDim.x >> 5) - 1;
    // per consumer warp 1 KB of scratch: 512 bytes of RGBX pixels (LANE mapping) and 512 bytes for the packed row
    const uint32_t scratch_s = (sfree_s + 8u * kStages + 15u) & ~15u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_s + 8u * s, 1);
            mbar_init(sfree_s + 8u * s, n_cons_warps);
        }
        mbar_init_fence();
        trace_stamp(a, 0);                                   // CTA started
    }
    __syncthreads();

    // contiguous, balanced range of the cost axis for this CTA: it owns the tiles that START inside it
    // (a share skewed towards the CTAs that reach an SM first -- they finish ~15 % ahead of the last ones, see
    // profiles/r02g_cta_timeline_c2.txt -- was measured and is slower: 47.4 -> 49 us at configs[1])
    const int u0 = (int)(((int64_t)a.total_units * blockIdx.x) / gridDim.x);
    const int u1 = (int)(((int64_t)a.total_units * (blockIdx.x + 1)) / gridDim.x);

    // warp roles: consumers 0 .. n_cons_warps-1, then the producer.  The consumers' waits are plain try_wait loops;
    // the producer, which runs ahead, sleeps between polls (suspend-time hints, sleep lengths and the producer as the
    // CTA's first warp were measured: no change in run time, profiles/r02p_*, r03b_*)
    const int warp_idx = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    const int lane = (int)threadIdx.x & 31;
    const int tid = (int)threadIdx.x;
    auto wait = [&](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };

    if (warp_idx >= n_cons_warps) {
        {
            // =========================== producer warp =========================================
            int st = 0;
            uint32_t ph = 0;                 // parity of the stage's current use
            bool first_copy = true;
            int n_chunks_done = 0;
            int img = first_image(a, u0, lane);
            View v = get_view(a, img);
            int local = (u0 - v.unit_begin + v.tile_units - 1) / v.tile_units;   // first tile starting at >= u0
            for (;;) {
                if (local >= v.n_strips * v.n_rowtiles) {              // next image
                    if (++img >= a.n_img) break;
                    v = get_view(a, img);
                    local = 0;
                }
                if (v.unit_begin + local * v.tile_units >= u1) break;
                // ---- segment: the output rows [local, local_end) of one (image, strip) ----------
                const int H = v.H, W = v.W, Ho = v.Ho, Wo = v.Wo;
                const int strip = local / v.n_rowtiles, rt = local % v.n_rowtiles;
                const int mine_end = (u1 - v.unit_begin + v.tile_units - 1) / v.tile_units;   // tiles starting before u1
                const int local_end = min((strip + 1) * v.n_rowtiles, mine_end);
                const int y_end = min(Ho, rt + (local_end - local));
                const int x_first = strip * v.strip_cols;
                const int ncols = min(v.strip_cols, Wo - x_first);
                const uint8_t* simg = v.src;
                const uintptr_t dimg = reinterpret_cast<uintptr_t>(v.dst);
                const float* my = v.my;
                const float* mx = v.mx + x_first;
                const int64_t row_pitch = (int64_t)W * kC;
                // source column span of the strip.  A single strip over an image whose whole rows fit a stage a
                // few at a time: stage whole rows without looking for the span first (the maps of this library
                // cover the image, so the span is the whole row anyway; the scan is a chain of dependent global
                // loads on the critical path of the CTA's first chunk) -- ONE bulk copy per chunk, because
                // consecutive rows are contiguous in global memory.
                int c_lo, row_bytes, slot_pitch;
                const bool one_copy = v.n_strips == 1 && (a.stage_bytes - 64) / (W * kC) >= 4;
                if (one_copy) {
                    c_lo = 0;
                    row_bytes = W * kC;
                    slot_pitch = row_bytes;
                } else {
                    int lo = 0x7fffffff, hi = -1;
                    // (MODE 2: the first warp of a strip that is not the image's first also computes the column before it)
                    for (int x = lane - ((MODE == 2 && x_first > 0) ? 1 : 0); x < ncols; x += 32) {
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
                    row_bytes = (c_hi - c_lo + 1) * kC;
                    // one copy per row: [16-byte alignment head][span][alignment tail + window over-read], and a
                    // pitch congruent to the row pitch modulo 16 so that slot k keeps the phase of source row k
                    slot_pitch = ((row_bytes + 45 + 15) & ~15) + (int)(row_pitch & 15);
                }
                // (one bulk copy per slot is issued by one lane each: at most 32 slots without the single-copy path)
                const int arena_slots = min(min((a.stage_bytes - 64) / slot_pitch, 2 * R), one_copy ? 2 * kMaxRows : 32);
                const uint8_t* scol = simg + (int64_t)c_lo * kC;
                const bool fixed = (slot_pitch & 3) == 0;
                uint32_t seg_flags = kFlagNewStrip | (fixed ? kFlagFixedShift : 0u);
                int carry_row = kNoCarry;    // the consumers hold the blends of rows carry_row - 1 and carry_row
                int y_cur = rt;
                // map_y is read through a register window of 2 x 32 rows (lane i holds rows y_win + i
                // and y_win + 32 + i) refilled 32 rows ahead of use
                int y_win = y_cur;
                int sy_cur = quantise_coord(__ldg(my + min(y_win + lane, Ho - 1)));
                int sy_nxt = quantise_coord(__ldg(my + min(y_win + 32 + lane, Ho - 1)));
                while (y_cur < y_end) {
                    const int tab = tab_off0 + st * kTabBytes;
                    const uint32_t stage_s = smem_s + (uint32_t)(st * a.stage_bytes);
                    // The sweep requests the window of slot s + 1 before it blends slot s, also after the last slot: that
                    // read (never used) must stay inside the CTA's shared memory.  With slots much wider than the
                    // strip (a narrow strip that taps whole source rows) the last stage has less room behind it.
                    const int max_slots = min(arena_slots, (a.smem_total - st * a.stage_bytes - 32) / slot_pitch - 1);
                    if (y_cur - y_win >= 32) {
                        y_win += 32;
                        sy_cur = sy_nxt;
                        sy_nxt = quantise_coord(__ldg(my + min(y_win + 32 + lane, Ho - 1)));
                    }
                    // ---- plan: lane i <-> output row y_cur + i --------------------------------
                    const int y = y_cur + lane;
                    // (the CTA's first chunks are short: their rows land sooner and the sweep starts earlier)
                    const bool live = y < y_end && lane < (n_chunks_done < 2 ? min(R, a.first_rows << n_chunks_done) : R);
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
                    // lane j stages source row r_lo + j into slot j; slot j's first byte sits at
                    // phase0 + j * slot_pitch, which has the 16-byte phase of the row's global address
                    const uint8_t* p0 = scol + (int64_t)r_lo * row_pitch;
                    const int phase0 = (int)(reinterpret_cast<uintptr_t>(p0) & 15);
                    const uint8_t* p = p0 + (int64_t)lane * row_pitch;
                    const int off = (int)(reinterpret_cast<uintptr_t>(p) & 15);
                    const uint32_t bytes = lane < n_slots ? (uint32_t)((off + row_bytes + 15) & ~15) : 0u;
                    const uint32_t tx = one_copy ? (uint32_t)((phase0 + n_slots * slot_pitch + 15) & ~15)
                                                 : __reduce_add_sync(0xffffffffu, bytes);
                    const uintptr_t gd = dimg + (uintptr_t)(((int64_t)y * Wo + x_first) * kC);   // this row's first byte

                    mbar_wait_relaxed(sfree_s + 8u * st, ph ^ 1u);    // stage free again (the producer runs ahead)
                    if (lane < n_rows) {
                        const uint32_t wu = (uint32_t)wa << 14, wl_ = (uint32_t)(32 - wa) << 14;   // upper / lower tap
                        const bool up_even = (ra & 1) == 0;
                        // z: offset of the row's first byte (of this strip) from the 4-byte aligned address at or
                        // below the image's first byte
                        st128(tab + kTabRows + 16 * lane,
                              make_uint4(up_even ? wu : wl_, up_even ? wl_ : wu,
                                         (uint32_t)(gd - (dimg & ~(uintptr_t)3)),
                                         (uint32_t)(ra + 1 - r_lo)));
                    } else if (lane == n_rows) {
                        st128(tab + kTabRows + 16 * lane, make_uint4(0u, 0u, 0u, kRowSentinel));
                    }
                    if (lane == 0) {
                        const uint32_t fl = seg_flags | ((r_lo & 1) ? kFlagOddFirst : 0u);
                        st128(tab, make_uint4((uint32_t)n_rows, (uint32_t)n_slots | (fl << 16),
                                              (uint32_t)slot_pitch, (uint32_t)phase0));
                        st128(tab + 16, make_uint4((uint32_t)img, (uint32_t)x_first, (uint32_t)y_cur, (uint32_t)c_lo));
                        const uintptr_t gbase = dimg & ~(uintptr_t)3;
                        st128(tab + kTabStore, make_uint4((uint32_t)gbase, (uint32_t)((uint64_t)gbase >> 32),
                                                          (uint32_t)(ncols * kC), (uint32_t)(Wo * kC)));
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
                        if (lane == 0 && n_slots > 0) bulk_g2s(stage_s, p0 - phase0, tx, full_s + 8u * st);
                    } else if (lane < n_slots) {
                        bulk_g2s(stage_s + (uint32_t)((phase0 + lane * slot_pitch) & ~15), p - off, bytes, full_s + 8u * st);
                    }
                    if (first_copy && lane == 0) trace_stamp(a, 1);    // first chunk planned, its copy issued
                    first_copy = false;
                    ++n_chunks_done;
                    carry_row = n_rows > 0 ? ra_last + 1 : kNoCarry;
                    y_cur += max(n_rows, 1);
                    if (++st == kStages) { st = 0; ph ^= 1u; }
                }
                local = local_end;
            }
            // terminator
            mbar_wait_relaxed(sfree_s + 8u * st, ph ^ 1u);
            if (lane == 0) {
                st128(tab_off0 + st * kTabBytes, make_uint4(0xffffffffu, 0u, 0u, 0u));
                mbar_arrive(full_s + 8u * st);
            }
        }
        return;
    }

    // =============================== consumer warps ==============================================
    // this thread's output columns inside the strip.  A warp's lanes cover 128 consecutive columns from xw: its block
    // of kBlockPx columns and, in MODE 2, the column before it (the halo: see AWQ_HALO_BEGIN).  QUAD mapping: columns
    // xw + 4 lane .. + 3.  LANE mapping: columns xw + 32 j + lane.
    constexpr int kBlockPx = MODE == 2 ? 127 : 128, kHalo = MODE == 2 ? 1 : 0;
    const int xw = warp_idx * kBlockPx - kHalo;
    const int x0 = xw + 4 * lane;
    int wo[4];                        // byte offset of each column's window inside a staged row span
    uint32_t wl[4], wh[4], wa[4], E[12], O[12];   // weight words per pixel: w0 w1 . . | . . w0 w1 | w0 . . w1
#pragma unroll
    for (int k = 0; k < 12; ++k) E[k] = O[k] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) { wo[j] = 0; wl[j] = wh[j] = wa[j] = 0u; }
    bool warp_live = false;           // some lane of this warp owns a column of the strip
    bool lane_map = false;            // this warp runs the LANE mapping in the current strip
    uint32_t store_ok = 0u;           // this thread owns at least one column of the strip (QUAD position)
    const uint32_t sx_s = scratch_s + (uint32_t)(warp_idx * 1024 + lane * 4);     // scratch: my RGBX pixels in
    DirectOps dops{};
    dops.lane12 = 12u * (uint32_t)lane;
    dops.sq = scratch_s + (uint32_t)(warp_idx * 1024 + lane * 16);
    dops.cb = (uint32_t)(3 * xw);
    const uint32_t p_s = scratch_s + (uint32_t)(warp_idx * 1024 + 512);       // P: the warp's 384 bytes (edge warps)
    dops.pw = p_s + (uint32_t)(lane * 12);
    int xba[4];                       // source column of each of my pixels' left tap (-1: none yet)

    // per-strip setup: taps and weights of this thread's columns, choice of the mapping.  A column outside the strip
    // gets zero weights and the window of a neighbour; the halo column (-1) exists when the strip is not the first.
    auto setup_strip = [&](const float* mx, int W, int ncols, int x_first) {
        const int x_min = (kHalo && x_first > 0) ? -1 : 0;
        auto in_strip = [&](int x) { return x >= x_min && x < ncols; };
        store_ok = (x0 + 3 >= x_min && x0 < ncols) ? 1u : 0u;          // one of my four columns is in the strip
        // the warp works if it has a column of its own block (the halo alone does not count)
        warp_live = warp_idx * kBlockPx < ncols;
        int xb = -1, xb_first = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int w0 = 0, w1 = 0;
            if (in_strip(x0 + j)) column_taps(__ldg(mx + x0 + j), W, xb, w0, w1);
            xba[j] = xb;
            wl[j] = (uint32_t)w0 | ((uint32_t)w1 << 8);
            wh[j] = wl[j] << 16;
            wa[j] = (uint32_t)w0 | ((uint32_t)w1 << 24);
        }
        // where the lane's first pixel taps (extrapolated from the second one when the first is outside the strip)
        xb_first = xba[0] >= 0 ? xba[0] : xba[1] - 1;
        // QUAD loads are conflict-free only while the lanes' windows stay 3 words apart: when the source column
        // of some lane's first pixel has drifted two or more pixels from "4 per lane", switch the warp to the
        // LANE mapping
        lane_map = false;
        if (warp_live && a.map_policy != 1) {
            const int xb_lane0 = __shfl_sync(0xffffffffu, xb_first, 0);
            const int dev = store_ok ? abs(xb_first - xb_lane0 - 4 * lane) : 0;
            lane_map = a.map_policy == 2 || __any_sync(0xffffffffu, dev >= a.drift);
        }
        if (lane_map) {
            xb = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = xw + 32 * j + lane;
                int w0 = 0, w1 = 0;
                if (in_strip(x)) column_taps(__ldg(mx + x), W, xb, w0, w1);
                xba[j] = xb;
                wl[j] = (uint32_t)w0 | ((uint32_t)w1 << 8);
                wh[j] = wl[j] << 16;
                wa[j] = (uint32_t)w0 | ((uint32_t)w1 << 24);
            }
        }
    };
    // The first strip of this CTA is known before the producer says so (same arithmetic on u0): set it up while
    // the first rows are still in flight -- the map_x loads and the tap arithmetic leave the CTA's start-up path.
    int pre_img = -1, pre_x_first = -1;
    {
        const int img = first_image(a, u0, lane);
        if (img < a.n_img) {
            const View v = get_view(a, img);
            int local = (u0 - v.unit_begin + v.tile_units - 1) / v.tile_units;
            if (local < v.n_strips * v.n_rowtiles && v.unit_begin + local * v.tile_units < u1) {
                const int strip = local / v.n_rowtiles;
                const int x_first = strip * v.strip_cols;
                setup_strip(v.mx + x_first, v.W, min(v.strip_cols, v.Wo - x_first), x_first);
                pre_img = img;
                pre_x_first = x_first;
            }
        }
    }

    int st = 0;                       // source stage of the current chunk
    uint32_t sph = 0u;                // parity to wait for: stage filled
    for (;; st = st + 1 == kStages ? 0 : st + 1, sph ^= st == 0 ? 1u : 0u) {
        const int tab = tab_off0 + st * kTabBytes;
        wait(full_s + 8u * st, sph);
        if (tid == 0 && a.trace != nullptr && a.trace[(size_t)blockIdx.x * 8 + 2] == 0ull) trace_stamp(a, 2);   // first rows landed
        const uint4 h0 = ld128(tab);
        const int n_rows = (int)h0.x;
        if (n_rows < 0) break;                               // the producer's stop
        const uint32_t flags = h0.y >> 16;
        if (flags & kFlagNewStrip) {                         // new strip: per-column taps and weights
            const uint4 h1 = ld128(tab + 16);
            if ((int)h1.x != pre_img || (int)h1.y != pre_x_first) {
                const uint4 hx = ld128(tab + kTabStrip);
                const float* mx = reinterpret_cast<const float*>(((uint64_t)hx.y << 32) | hx.x);
                setup_strip(mx, (int)hx.z, (int)hx.w, (int)h1.y);
            }
            pre_img = -1;                                    // the early setup serves the first segment only
#pragma unroll
            for (int j = 0; j < 4; ++j) wo[j] = xba[j] < 0 ? 0 : (xba[j] - (int)h1.w) * kC;
        }
        if (n_rows == 0) {
            // ---- direct path for one output row whose source span does not fit a stage ----------
            if (store_ok) {
                const uint4 h1 = ld128(tab + 16);
                const int img = (int)h1.x, x_first = (int)h1.y, y0 = (int)h1.z;
                const View v = get_view(a, img);
                const int H = v.H, W = v.W, Wo = v.Wo;
                const int ncols = min(v.strip_cols, Wo - x_first);
                const int sy = quantise_coord(__ldg(v.my + y0));
                const int ay = sy & 31;
                const int ya = clampi(sy >> 5, 0, H - 1), yb = clampi((sy >> 5) + 1, 0, H - 1);
                for (int j = 0; j < 4 && x0 + j < ncols; ++j) {
                    if (x0 + j < warp_idx * kBlockPx) continue;          // the halo column belongs to the previous block
                    const int sx = quantise_coord(__ldg(v.mx + x_first + x0 + j));
                    const int ax = sx & 31;
                    const int xa = clampi(sx >> 5, 0, W - 1), xc = clampi((sx >> 5) + 1, 0, W - 1);
                    uint8_t* o = v.dst + ((int64_t)y0 * Wo + x_first + x0 + j) * kC;
#pragma unroll
                    for (int k = 0; k < kC; ++k)
                        o[k] = bilinear_u8(__ldg(v.src + ((int64_t)ya * W + xa) * kC + k),
                                           __ldg(v.src + ((int64_t)ya * W + xc) * kC + k),
                                           __ldg(v.src + ((int64_t)yb * W + xa) * kC + k),
                                           __ldg(v.src + ((int64_t)yb * W + xc) * kC + k), ax, ay);
                }
            }
        } else if (warp_live && !(a.dbg & 1)) {
            const int n_slots = (int)(h0.y & 0xffffu);
            const uint32_t rp_s = smem_s + (uint32_t)(tab + kTabRows);
            const uint32_t base = smem_s + (uint32_t)(st * a.stage_bytes) + h0.w;    // first byte of slot 0
            const uint32_t odd = (flags & kFlagOddFirst) ? 1u : 0u;
            uint32_t win[4], sh[4];
            {
                const uint4 hs = ld128(tab + kTabStore);                         // {image lo, hi, strip row bytes, Wo * 3}
                const uint64_t img_al = ((uint64_t)hs.y << 32) | hs.x;
                if (MODE == 1) {
                    // the warp's 384 bytes of a row start 384 * warp bytes after the strip's first byte
                    dops.obase = img_al + (uint64_t)(warp_idx * 384);
                } else {
                    // stream positions relative to P0 (lane 0's first byte): the block's own bytes are [3, 3 + 3 n)
                    const uint4 h1 = ld128(tab + 16);
                    const int ncols = (int)hs.z / kC, x_first = (int)h1.y;
                    const int n = min(kBlockPx, ncols - warp_idx * kBlockPx);
                    const bool head_warp = warp_idx == 0 && x_first == 0;
                    const bool tail_warp = warp_idx * kBlockPx + n == ncols && (x_first + ncols) * kC == (int)hs.w;
                    uint32_t vm = 0u, em = 0u;
                    const int rel = lane < 3 ? 3 * n + lane : 3 + (lane - 4);       // tail byte i | head byte i
#pragma unroll
                    for (int al = 0; al < 4; ++al) {
#pragma unroll
                        for (int m = 0; m < 3; ++m) {
                            const int r = 12 * lane + 4 * m - al;                   // first byte of the aligned word
                            if (r >= (head_warp ? 3 : 0) && r < 3 * n) vm |= 1u << (4 * al + m);
                        }
                        const int tc = (al + 3 + 3 * n) & 3;                        // bytes of the row's last, partial word
                        const int h0 = (al + 3) & 3;                                // misalignment of the row's first byte
                        const bool tail = tail_warp && lane < 3 && lane >= 3 - tc;
                        const bool head = head_warp && lane >= 4 && lane < 7 && h0 != 0 && (lane - 4) < 4 - h0 &&
                                          (lane - 4) < 3 * n;
                        if (tail || head) em |= 1u << al;
                    }
                    dops.obase = img_al;
                    dops.wmask = vm;
                    dops.emask = em;
                    dops.eoff = (uint32_t)rel;
                    dops.pedge = p_s + (uint32_t)rel;
                    dops.edge_warp = (head_warp || tail_warp) ? 1u : 0u;
                }
            }
            if (flags & kFlagFixedShift) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t b = base + (uint32_t)wo[j];
                    win[j] = b & ~3u;
                    sh[j] = b << 3;
                }
                if (!lane_map) sweep_quad<true, false, MODE>(E, O, win, sh, wl, wh, wa, n_slots, h0.z, rp_s, odd, sx_s, store_ok, dops);
                else sweep_quad<true, true, MODE>(E, O, win, sh, wl, wh, wa, n_slots, h0.z, rp_s, odd, sx_s, store_ok, dops);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { win[j] = base + (uint32_t)wo[j]; sh[j] = 0u; }
                if (!lane_map) sweep_quad<false, false, MODE>(E, O, win, sh, wl, wh, wa, n_slots, h0.z, rp_s, odd, sx_s, store_ok, dops);
                else sweep_quad<false, true, MODE>(E, O, win, sh, wl, wh, wa, n_slots, h0.z, rp_s, odd, sx_s, store_ok, dops);
            }
        }
        __syncwarp();
        if (tid == 0 && a.trace != nullptr) {
            if (a.trace[(size_t)blockIdx.x * 8 + 3] == 0ull) trace_stamp(a, 3);   // first chunk swept (warp 0)
            trace_stamp(a, 6);                                                     // latest chunk swept
        }
        if (lane == 0) mbar_arrive(sfree_s + 8u * st);       // this warp is done with the stage
    }
}

// ---- launch geometry ------------------------------------------------------------------------------
struct StripPlan { int n_strips, strip_cols; };
inline StripPlan plan_strips(int Wo, int max_cols_) {
    StripPlan p;
    p.n_strips = (Wo + max_cols_ - 1) / max_cols_;
    p.strip_cols = (((Wo + p.n_strips - 1) / p.n_strips) + 15) & ~15;
    p.n_strips = (Wo + p.strip_cols - 1) / p.strip_cols;
    return p;
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// One launch: `kern` (a build of the kernel whose launch bounds admit it) run with `warps` consumer warps + the
// producer warp per CTA, `ctas` CTAs per SM wanted.
using QuadKernel = void (*)(const QuadArgs);
int launch_core(QuadKernel kern, int warps, int ctas, QuadArgs& a, int cols, cudaStream_t st) {
    const int threads = (warps + 1) * 32;
    a.dbg = env_int("ATTWARP_REMAP_DBG", 0);
    const int unit_pitch = (((cols + 1) * kC + 45 + 15) & ~15) + 16;      // a slot at unit scale
    // ring depth (ATTWARP_QUAD_RING = 2 .. 4 stages, tuning experiments) and rows per chunk: as many as the shared
    // memory of 1 / ctas of an SM holds (stages of R + 2 slots), at most kMaxRows
    int stages = 2;
    {
        const int ring = env_int("ATTWARP_QUAD_RING", 0);
        if (ring >= 2 && ring <= kMaxRing) stages = ring;
    }
    a.stages = stages;
    const int scratch = 1024 * warps + 16;
    const int budget = (227 * 1024) / ctas - 1024 - stages * kTabBytes - 16 * kMaxRing - 256 - scratch;
    int R = (budget - stages * (2 * unit_pitch + 64 + 128)) / (stages * unit_pitch);
    R = R > kMaxRows ? kMaxRows : R;
    const int forced = env_int("ATTWARP_QUAD_ROWS", 0);
    if (forced >= 2 && forced <= R) R = forced;
    if (R < 2) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: strips of %d columns do not fit shared memory", cols);
    a.rows = R;
    a.stage_bytes = ((R + 2) * unit_pitch + 64 + 127) & ~127;
    const size_t smem_bytes = (size_t)stages * (a.stage_bytes + kTabBytes) + 2 * (size_t)stages * sizeof(uint64_t) + 16 +
                              (size_t)scratch;
    a.smem_total = (int)smem_bytes;
    a.first_rows = env_int("ATTWARP_QUAD_FIRST_ROWS", 8);    // 8, 16, then full chunks: -1..2 % (profiles/r04e_first_rows.txt)
    a.map_policy = env_int("ATTWARP_QUAD_MAP", 0);          // 0 auto, 1 QUAD only, 2 LANE wherever possible
    a.drift = env_int("ATTWARP_QUAD_DRIFT", 2);
    // per (kernel, device): the largest shared-memory size configured so far; per (kernel, device, threads, smem): occupancy
    struct Key {
        QuadKernel k; int dev, threads; size_t smem;
        bool operator<(const Key& o) const { return std::tie(k, dev, threads, smem) < std::tie(o.k, o.dev, o.threads, o.smem); }
    };
    static std::mutex mu;
    static std::map<std::pair<QuadKernel, int>, size_t> configured;
    static std::map<Key, int> occupancy;
    int dev = 0, occ_q = 0;
    AW_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        size_t& conf = configured[{kern, dev}];
        if (smem_bytes > conf) {
            AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            conf = smem_bytes;
        }
        const Key key{kern, dev, threads, smem_bytes};
        auto it = occupancy.find(key);
        if (it == occupancy.end()) {
            int o = 0;
            AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem_bytes));
            it = occupancy.emplace(key, o).first;
        }
        occ_q = it->second;
    }
    {
        const int cap = env_int("ATTWARP_REMAP_CTAS_PER_SM", 0);
        if (cap >= 1 && cap < occ_q) occ_q = cap;
    }
    if (occ_q < 1) return fail(ATTWARP_ERR_CUDA, "remap: kernel does not fit an SM (%zu B shared)", smem_bytes);
    // attwarp_set_sm_share(2): leave half of every SM to the kernels of another stream
    const int share = sm_share();
    const int occ = occ_q >= 2 * share ? occ_q / share : (occ_q >= 2 && share > 1 ? occ_q / 2 : occ_q);
    const int64_t cap = (int64_t)sm_count() * occ;
    const int grid = (int)(a.total_units < cap ? a.total_units : cap);
    // ATTWARP_REMAP_TRACE=<file>: per-CTA global-timer stamps of every launch, appended to the file (debugging
    // only: synchronises the stream)
    const char* trace_path = getenv("ATTWARP_REMAP_TRACE");
    a.trace = nullptr;
    if (trace_path != nullptr && trace_path[0] != 0) {
        static unsigned long long* dbuf = nullptr;
        static int dcap = 0;
        if (grid > dcap) {
            if (dbuf != nullptr) cudaFree(dbuf);
            AW_CUDA(cudaMalloc(&dbuf, sizeof(unsigned long long) * 8 * (size_t)grid));
            dcap = grid;
        }
        AW_CUDA(cudaMemsetAsync(dbuf, 0, sizeof(unsigned long long) * 8 * (size_t)grid, st));
        a.trace = dbuf;
    }
    kern<<<grid, threads, smem_bytes, st>>>(a);
    if (a.trace != nullptr) {
        std::vector<unsigned long long> h((size_t)grid * 8);
        AW_CUDA(cudaStreamSynchronize(st));
        AW_CUDA(cudaMemcpy(h.data(), a.trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "a")) {
            fprintf(f, "launch grid=%d threads=%d rows=%d smem=%zu\n", grid, threads, a.rows, smem_bytes);
            for (int i = 0; i < grid; ++i) {
                fprintf(f, "%d", i);
                for (int k = 0; k < 8; ++k) fprintf(f, " %llu", h[(size_t)i * 8 + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return check_launch("remap_u8_quad_kernel");
}

// The CTA gets exactly the consumer warps the strips need -- ceil(columns / 128), at least 3 -- and the SM as many
// CTAs as its registers hold.  Builds of the kernel by register budget (launch bounds): <= 3 consumer warps (5 CTAs
// per SM by registers, run with 4), <= 6 (2-3 CTAs), <= 11 and 12 (one CTA, ~120 registers) and a 96-register build
// that runs 7..9 warps with two CTAs per SM and 13..16 (strips of <= 2048 columns) with one.
constexpr int kDirectMaxWarps = 16;
// columns a consumer warp owns: 128, or 127 + the halo column when rows may be unaligned (MODE 2)
inline int block_px(bool aligned) { return aligned ? 128 : 127; }
inline int direct_warps(int cols, bool aligned) {
    const int w = (cols + block_px(aligned) - 1) / block_px(aligned);
    return w < 3 ? 3 : w;
}
// ATTWARP_QUAD_MAXW = 3 .. 16 narrows the widest strip to that many warps (tuning experiments); strips are cut at
// multiples of 16 columns
inline int direct_max_cols(bool aligned) {
    const int w = env_int("ATTWARP_QUAD_MAXW", kDirectMaxWarps);
    return ((w >= 3 && w <= kDirectMaxWarps ? w : kDirectMaxWarps) * block_px(aligned)) & ~15;
}
template <int MODE>
int launch_direct_mode(QuadArgs& a, int cols, cudaStream_t st) {
    const int w = direct_warps(cols, MODE == 1);
    // (<= 3 warps: four CTAs per SM with ~22-row chunks beat five with 16-row chunks -- fewer chunk turn-arounds:
    // 43.9 -> 42.6 us at configs[1], profiles/r04b_rows.txt)
    if (w <= 3) return launch_core(remap_u8_quad_kernel<128, 5, MODE>, w, 4, a, cols, st);
    if (w <= 6) return launch_core(remap_u8_quad_kernel<224, 2, MODE>, w, w == 4 ? 3 : 2, a, cols, st);
    if (w <= 9) return launch_core(remap_u8_quad_kernel<(kDirectMaxWarps + 1) * 32, 1, MODE>, w, 2, a, cols, st);
    if (w <= 11) return launch_core(remap_u8_quad_kernel<384, 1, MODE>, w, 1, a, cols, st);
    if (w <= 12) return launch_core(remap_u8_quad_kernel<416, 1, MODE>, w, 1, a, cols, st);
    return launch_core(remap_u8_quad_kernel<(kDirectMaxWarps + 1) * 32, 1, MODE>, w, 1, a, cols, st);
}
// aligned: every destination row of the launch starts at a multiple of 4 bytes
int launch_direct(QuadArgs& a, int cols, bool aligned, cudaStream_t st) {
    return aligned ? launch_direct_mode<1>(a, cols, st) : launch_direct_mode<2>(a, cols, st);
}

inline bool rows_word_aligned(const void* dst, int Wo) {
    return ((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)(Wo * kC)) & 3) == 0;
}

}  // namespace

// ATTWARP_REMAP_QUAD=0 keeps the round-1 kernel (A/B comparisons).
bool remap_quad_enabled() {
    static const bool v = [] {
        const char* e = getenv("ATTWARP_REMAP_QUAD");
        return !(e != nullptr && atoi(e) == 0);
    }();
    return v;
}

// Uniform batch of HWC uint8 images with 3 channels, H, W >= 2.
int launch_remap_u8_quad(const void* src, void* dst, int n_img, int H, int W, int Ho, int Wo, const float* map_x,
                         const float* map_y, cudaStream_t st) {
    QuadArgs a{};
    // row offsets inside an image travel as 32-bit values (signed in MODE 2)
    if ((int64_t)Ho * Wo * kC > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: output image of %d x %d exceeds 2 GB", Ho, Wo);
    const bool aligned = rows_word_aligned(dst, Wo);
    const StripPlan sp = plan_strips(Wo, direct_max_cols(aligned));
    a.n_strips = sp.n_strips;
    a.strip_cols = sp.strip_cols;
    a.src = static_cast<const uint8_t*>(src);
    a.dst = static_cast<uint8_t*>(dst);
    a.map_x = map_x;
    a.map_y = map_y;
    a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo;
    a.n_rowtiles = Ho;
    a.imgs = nullptr;
    a.n_img = n_img;
    const int64_t total = (int64_t)n_img * a.n_strips * a.n_rowtiles;
    if (total > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: too many tiles");
    a.total_units = (int)total;
    const int cols = Wo < a.strip_cols ? Wo : a.strip_cols;
    return launch_direct(a, cols, aligned, st);
}

// Ragged batch.  Images are grouped into width classes, one launch per class with the geometry that fits it (a
// 300-wide image in a CTA built for 1408 columns would leave eight of its eleven consumer warps idle):
//   classes 0 .. 13   destination rows 4-byte aligned (MODE 1): strips of <= 384 columns (3 consumer warps), then one
//                     class per extra 128 columns up to 2048 (16 warps); wider images are cut into strips.  A class
//                     with less than ~24 M output pixels (the launch would cost more than the warps it saves) joins
//                     the next wider non-empty one.
//   classes 14 .. 27  the same widths for rows at any alignment (MODE 2)
// Step 1: `host` (n + 1 entries, batch order) gets each image's strip plan and is uploaded to dev_main (the maps
// kernel reads shapes and map pointers from it); a copy grouped by class, every group followed by an entry that
// carries its unit total, is uploaded to dev_sorted (n + kRaggedClasses entries).
constexpr int kDirectClasses = kDirectMaxWarps - 2;
static_assert(2 * kDirectClasses == kRaggedClasses, "class table");

int launch_remap_u8_quad_ragged_prepare(RaggedImage* host, int n, RaggedImage* dev_main, RaggedImage* dev_sorted,
                                        RaggedQuadPlan* plan, cudaStream_t st) {
    static thread_local std::vector<RaggedImage> sorted;
    static thread_local std::vector<int> cls;
    sorted.assign((size_t)n + kRaggedClasses, RaggedImage{});
    cls.assign((size_t)n, 0);
    *plan = RaggedQuadPlan{};
    int64_t px[kRaggedClasses] = {};
    for (int i = 0; i < n; ++i) {
        const RaggedImage& im = host[i];
        if ((int64_t)im.Ho * im.Wo * kC > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: output image %d exceeds 2 GB", i);
        const bool aligned = rows_word_aligned(im.dst, im.Wo);
        const StripPlan sp = plan_strips(im.Wo, direct_max_cols(aligned));
        const int c = direct_warps(im.Wo < sp.strip_cols ? im.Wo : sp.strip_cols, aligned) - 3 +
                      (aligned ? 0 : kDirectClasses);
        cls[(size_t)i] = c;
        px[c] += (int64_t)im.Wo * im.Ho;
    }
    // An aligned class next to a non-empty any-alignment class of the same width joins it unless it is big enough
    // to pay for a launch of its own (MODE 2 handles aligned rows too, ~20 % slower; a launch's ramp and tail cost
    // ~10 us, a third of it exposed with three streams).  Then small classes join the next wider non-empty one: a
    // class wastes ~1 / warps of its work there, so below ~24 M output pixels the launch costs more than the waste
    // (profiles/r03c_c4_class_min_px.txt: 0.377 -> 0.35 ms for a 128-image shard, nothing lost at 1024 images).
    int join[kRaggedClasses];
    for (int c = 0; c < kRaggedClasses; ++c) join[c] = c;
    const int64_t min_px = (int64_t)env_int("ATTWARP_QUAD_CLASS_MIN_PX", 24000000);
    const int64_t min_aligned_px = (int64_t)env_int("ATTWARP_QUAD_ALIGNED_MIN_PX", 32000000);
    for (int c = 0; c < kDirectClasses; ++c) {
        if (px[c] == 0 || px[c] >= min_aligned_px || px[c + kDirectClasses] == 0) continue;
        px[c + kDirectClasses] += px[c];
        px[c] = 0;
        join[c] = c + kDirectClasses;
    }
    for (int half = 0; half < 2; ++half) {
        const int c0 = half * kDirectClasses, c1 = c0 + kDirectClasses;
        for (int c = c0; c + 1 < c1; ++c) {
            if (px[c] == 0 || px[c] >= min_px) continue;
            int up = c + 1;
            while (up < c1 && px[up] == 0) ++up;
            if (up == c1) continue;
            px[up] += px[c];
            px[c] = 0;
            join[c] = up;
        }
    }
    for (int i = 0; i < n; ++i) {
        int c = cls[(size_t)i];
        while (join[c] != c) c = join[c];
        cls[(size_t)i] = c;
        plan->count[c]++;
    }
    int pos = 0;
    for (int c = 0; c < kRaggedClasses; ++c) {
        plan->offset[c] = pos;
        plan->mode[c] = c < kDirectClasses ? 1 : 2;          // the kernel's MODE
        pos += plan->count[c] + 1;
    }
    int fill[kRaggedClasses] = {};
    int64_t total[kRaggedClasses] = {};
    for (int i = 0; i < n; ++i) {
        const int c = cls[(size_t)i];
        // (an aligned image that joined an any-alignment class is planned like the rest of that class)
        const bool aligned = plan->mode[c] == 1;
        const StripPlan sp = plan_strips(host[i].Wo, direct_max_cols(aligned));
        const int cols = host[i].Wo < sp.strip_cols ? host[i].Wo : sp.strip_cols;
        const int units = (cols + block_px(aligned) - 1) / block_px(aligned);
        if (sp.n_strips > 0xffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: image %d is too wide", i);
        host[i].strips_units = sp.n_strips | (units << 16);
        host[i].strip_cols = sp.strip_cols;
        host[i].n_rowtiles = host[i].Ho;
        host[i].unit_begin = (int)total[c];
        total[c] += (int64_t)sp.n_strips * host[i].n_rowtiles * units;
        if (total[c] > 0x7fffffff) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: too many tiles");
        if (cols > plan->max_strip[c]) plan->max_strip[c] = cols;
        sorted[(size_t)(plan->offset[c] + fill[c]++)] = host[i];
    }
    for (int c = 0; c < kRaggedClasses; ++c) {
        plan->total_units[c] = (int)total[c];
        sorted[(size_t)(plan->offset[c] + plan->count[c])].unit_begin = (int)total[c];
    }
    host[n] = RaggedImage{};
    AW_CUDA(cudaMemcpyAsync(dev_main, host, sizeof(RaggedImage) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    AW_CUDA(cudaMemcpyAsync(dev_sorted, sorted.data(), sizeof(RaggedImage) * sorted.size(), cudaMemcpyHostToDevice, st));
    return ATTWARP_OK;
}
// Step 2: one launch per non-empty class.  The classes write disjoint images, so their launches fan out over up to
// three streams (the caller's + two of the library's, forked and joined with events): every launch is a
// persistent grid with a ramp (first chunk ~8 us) and a ragged tail, which the next class's CTAs fill as SMs come free.  ATTWARP_RAGGED_STREAMS=1 keeps everything on the caller's stream.
namespace {
struct FanOut {
    static constexpr int kAux = 2;
    int dev = -1;
    cudaStream_t aux[kAux] = {};
    cudaEvent_t fork = nullptr, join[kAux] = {};
    void release() {
        if (dev < 0) return;
        for (int i = 0; i < kAux; ++i) {
            if (aux[i]) cudaStreamDestroy(aux[i]);
            if (join[i]) cudaEventDestroy(join[i]);
            aux[i] = nullptr;
            join[i] = nullptr;
        }
        if (fork) cudaEventDestroy(fork);
        fork = nullptr;
        dev = -1;
    }
    int ensure(int d) {
        if (dev == d) return ATTWARP_OK;
        release();
        for (int i = 0; i < kAux; ++i) {
            AW_CUDA(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
            AW_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
        AW_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        dev = d;
        return ATTWARP_OK;
    }
    ~FanOut() { release(); }
};
}  // namespace

int ragged_quad_launches(const RaggedQuadPlan& plan) {
    int n_launch = 0;
    for (int c = 0; c < kRaggedClasses; ++c) n_launch += plan.count[c] != 0 && plan.total_units[c] != 0;
    return n_launch;
}

int launch_remap_u8_quad_ragged_run(const RaggedQuadPlan& plan, const RaggedImage* dev_sorted, cudaStream_t st) {
    const int n_launch = ragged_quad_launches(plan);
    int lanes = env_int("ATTWARP_RAGGED_STREAMS", 1 + FanOut::kAux);
    lanes = lanes < 1 ? 1 : (lanes > 1 + FanOut::kAux ? 1 + FanOut::kAux : lanes);
    if (lanes > n_launch) lanes = n_launch;
    static thread_local FanOut fo;
    if (lanes > 1) {
        int dev = 0;
        AW_CUDA(cudaGetDevice(&dev));
        const int rc = fo.ensure(dev);
        if (rc != ATTWARP_OK) return rc;
        AW_CUDA(cudaEventRecord(fo.fork, st));
        for (int i = 0; i + 1 < lanes; ++i) AW_CUDA(cudaStreamWaitEvent(fo.aux[i], fo.fork, 0));
    }
    // biggest classes first (longest-processing-time order over the lanes): the small ones fill the gaps at the end
    int order[kRaggedClasses], n_order = 0;
    for (int c = 0; c < kRaggedClasses; ++c)
        if (plan.count[c] != 0 && plan.total_units[c] != 0) order[n_order++] = c;
    std::sort(order, order + n_order, [&](int x, int y) {
        return plan.total_units[x] != plan.total_units[y] ? plan.total_units[x] > plan.total_units[y] : x < y;
    });
    int rc = ATTWARP_OK;
    for (int k = 0; k < n_order && rc == ATTWARP_OK; ++k) {
        const int c = order[k];
        QuadArgs a{};
        a.imgs = dev_sorted + plan.offset[c];
        a.n_img = plan.count[c];
        a.total_units = plan.total_units[c];
        const int lane = k % lanes;
        rc = launch_direct(a, plan.max_strip[c], plan.mode[c] == 1, lane == 0 ? st : fo.aux[lane - 1]);
    }
    // join even after a failed launch: the caller's stream must not run ahead of work already enqueued
    for (int i = 0; i + 1 < lanes; ++i) {
        AW_CUDA(cudaEventRecord(fo.join[i], fo.aux[i]));
        AW_CUDA(cudaStreamWaitEvent(st, fo.join[i], 0));
    }
    return rc;
}

}  // namespace aw
