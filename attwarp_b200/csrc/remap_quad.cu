// remap_quad.cu -- stage 5 for 3-channel interleaved uint8 images (the benchmark format): persistent,
// warp-specialised streaming resample, FOUR ADJACENT output pixels per consumer thread.
//
// Same arithmetic as remap_direct_kernel (cv2.remap INTER_LINEAR + BORDER_REPLICATE, see warp_math.h;
// reference call sites "Attention Guided Warping/new_method.py:268-271",
// "model/marginalnet_full_dataset/checkpoint_utils.py:195-198").  Same skeleton as remap_stream.cu (producer
// warp plans chunks of output rows and fetches the source rows they tap with cp.async.bulk, consumer warps
// sweep them once, store warp ships output tiles with cp.async.bulk), with the consumer side rebuilt around the
// two resources the round-1 kernel ran out of -- shared-memory wavefronts and issue slots (profiles/README.md):
//
//   * a thread owns output columns 4t .. 4t+3 of its strip: 12 contiguous output bytes = three aligned 32-bit
//     stores per output row instead of twelve byte stores, and its four source windows sit 3 words apart at
//     unit scale (12-byte lane stride: every 32-bit load of a warp is conflict-free);
//   * the horizontal blends of the two most recent source rows are held per channel in an EVEN-row and an
//     ODD-row register (source row r goes to E when r is even), so a new source row overwrites one of them
//     without re-packing; the vertical blend of an output row is  t = wO*O + (wE*E + 512 * 2^14)  with the row's
//     weights pre-shifted by 14 bits -- two 32-bit multiply-adds (the sum stays below 2^32) whose TOP BYTE is the
//     output byte ((v + 512) >> 10 with nothing to shift or mask: prmt picks the top bytes when packing);
//   * the horizontal blend pairs the taps of a channel with one prmt per two channels (bytes p0c0 p1c0 p0c1
//     p1c1) so that a pixel needs two weight words instead of five;
//   * source rows are staged at a UNIFORM shared-memory pitch whatever the alignment of the image: a strip that
//     spans whole rows is fetched with ONE bulk copy per chunk (global rows are contiguous; the shared-memory
//     image is the global one shifted by a multiple of 16 bytes), narrower strips with one copy per row into
//     slots whose pitch is congruent to the row pitch modulo 16.  When the pitch is a multiple of 4 the
//     per-pixel window addresses advance by a constant and their byte shifts never change (fixed-shift sweep);
//     otherwise address and shift are re-derived per slot (three more ALU operations per pixel and slot);
//   * strips are as wide as the consumer threads allow (up to 4 columns x 352 threads = 1408): a 1344-wide
//     image is processed in whole rows, fetched and shipped as contiguous 4 KB rows.
//
// Maps need not be monotone (a chunk ends before the first output row that taps an earlier source row than its
// predecessor), a strip whose source span does not fit a stage is gathered from global memory (the kernel is
// total), ragged batches run in one launch over a descriptor table.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "bulk_ptx.cuh"
#include "common.cuh"

namespace aw {
namespace {

using namespace ptx;

constexpr int kMaxRing = 4;            // source-row stages / output tiles per CTA are launch parameters (2 .. 4)
constexpr int kMaxRows = 16;          // capacity of a chunk table; the rows per chunk are a launch parameter
// role warps: one producer warp + `store_warps` store warps (a launch parameter: wide strips with odd row pitches keep
// several warps busy shifting rows on their way out)
constexpr int kC = 3;

extern __shared__ __align__(128) uint8_t smem[];
__device__ __forceinline__ uint32_t ld32(int off) { return *reinterpret_cast<const uint32_t*>(smem + off); }
__device__ __forceinline__ uint4 ld128(int off) { return *reinterpret_cast<const uint4*>(smem + off); }
__device__ __forceinline__ void st32(int off, uint32_t v) { *reinterpret_cast<uint32_t*>(smem + off) = v; }
__device__ __forceinline__ void st128(int off, uint4 v) { *reinterpret_cast<uint4*>(smem + off) = v; }

// ---- per-stage chunk table (byte offsets), written by the producer, read by the consumers ----------
//   +0   uint4 {n_rows, n_slots | flags << 16, slot_pitch, byte offset of slot 0's first byte in the arena}
//              n_rows 0: output row y0 takes the direct path; -1: stop
//   +16  uint4 {img, x_first, y0, c_lo}
//   +32  uint4 {address of the chunk's first output byte -- DIRECT: of the image's -- (lo, hi), bytes per strip row,
//               Wo * 3}
//   +48  uint4 {address of map_x[x_first] (lo, hi), W, columns in the strip}      (new strip only)
//   +64  uint4 row[kMaxRows + 1]:  x = wE << 14, y = wO << 14  (weights of the even / odd source row)
//                                  z = byte offset of the row inside the output tile: row * pitch + (address of the
//                                      row's first destination byte & 12) -- always 4-byte aligned for the
//                                      consumers' word stores; the store warp bridges the remaining 0..3 bytes
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
// Common: n_slots, pitch (bytes between slots), rp (shared address of row[0]), ocol (shared address of this
// thread's 12 bytes in a tile row at offset 0), first slot parity, store predicate.
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
// align the 8-byte window, pair the taps: X = p0c0 p1c0 p0c1 p1c1,  Y = p0c2 p1c2 . .
#define AWQ_ALIGN1(J, SH)                                       \
    "shf.r.wrap.b32 A" #J ", lo" #J ", mi" #J ", " SH ";\n"     \
    "shf.r.wrap.b32 B" #J ", mi" #J ", hi" #J ", " SH ";\n"     \
    "prmt.b32 X" #J ", A" #J ", B" #J ", 0x4130;\n"             \
    "prmt.b32 Y" #J ", A" #J ", B" #J ", 0x0052;\n"
#define AWQ_ALIGN_F AWQ_ALIGN1(a, "sa") AWQ_ALIGN1(b, "sb") AWQ_ALIGN1(c, "sc") AWQ_ALIGN1(d, "sd")
#define AWQ_ALIGN_V AWQ_ALIGN1(a, "ma") AWQ_ALIGN1(b, "mb") AWQ_ALIGN1(c, "mc") AWQ_ALIGN1(d, "md")
// the shift of the CURRENT slot must survive the address update of the next one
#define AWQ_KEEP_V "mov.b32 ma, na;\n mov.b32 mb, nb;\n mov.b32 mc, nc;\n mov.b32 md, nd;\n"
#define AWQ_DOT1(P, J)                                          \
    "dp4a.u32.u32 " #P #J "0, X" #J ", wl" #J ", 0;\n"          \
    "dp4a.u32.u32 " #P #J "1, X" #J ", wh" #J ", 0;\n"          \
    "dp4a.u32.u32 " #P #J "2, Y" #J ", wl" #J ", 0;\n"
#define AWQ_DOT(P) AWQ_DOT1(P, a) AWQ_DOT1(P, b) AWQ_DOT1(P, c) AWQ_DOT1(P, d)
// vertical blend of one byte: top byte of the 32-bit  E * (wE << 14) + O * (wO << 14) + (512 << 14)
#define AWQ_V1(J, K)                                            \
    "mad.lo.u32 v" #J #K ", E" #J #K ", ex, 0x800000;\n"        \
    "mad.lo.u32 v" #J #K ", O" #J #K ", ey, v" #J #K ";\n"
#define AWQ_VBLEND                                              \
    AWQ_V1(a, 0) AWQ_V1(a, 1) AWQ_V1(a, 2) AWQ_V1(b, 0) AWQ_V1(b, 1) AWQ_V1(b, 2)  \
    AWQ_V1(c, 0) AWQ_V1(c, 1) AWQ_V1(c, 2) AWQ_V1(d, 0) AWQ_V1(d, 1) AWQ_V1(d, 2)
// 12 top bytes -> 3 words (prmt: x.b3 | y.b3 << 8, then the low halves of two pairs), three aligned stores
#define AWQ_EMIT_W                                              \
    AWQ_VBLEND                                                  \
    "prmt.b32 q0, va0, va1, 0x0073;\n"                          \
    "prmt.b32 q1, va2, vb0, 0x0073;\n"                          \
    "prmt.b32 q2, vb1, vb2, 0x0073;\n"                          \
    "prmt.b32 q3, vc0, vc1, 0x0073;\n"                          \
    "prmt.b32 q4, vc2, vd0, 0x0073;\n"                          \
    "prmt.b32 q5, vd1, vd2, 0x0073;\n"                          \
    "prmt.b32 q0, q0, q1, 0x5410;\n"                            \
    "prmt.b32 q2, q2, q3, 0x5410;\n"                            \
    "prmt.b32 q4, q4, q5, 0x5410;\n"                            \
    "add.u32 o, ez, %38;\n"                                     \
    "@pv st.shared.b32 [o], q0;\n"                              \
    "@pv st.shared.b32 [o+4], q2;\n"                            \
    "@pv st.shared.b32 [o+8], q4;\n"
// LANE mapping: the thread's four pixels are 32 columns apart (pixel j of lane t is column 32 j + t of the warp's
// 128-column block), so that the window loads of a warp stay inside ~128 bytes and never conflict whatever the
// local scale of the map.  The output bytes change hands through a per-warp scratch (two 512-byte buffers used
// alternately: one bar.warp.sync per row): 4-byte RGBX per pixel in, the 16 bytes of four adjacent pixels out,
// squeezed to 12 bytes -- the tile stores are the same three aligned words as in the QUAD mapping.
#define AWQ_EMIT_L                                              \
    AWQ_VBLEND                                                  \
    "prmt.b32 q0, va0, va1, 0x0073;\n prmt.b32 q0, q0, va2, 0x0710;\n"  \
    "prmt.b32 q1, vb0, vb1, 0x0073;\n prmt.b32 q1, q1, vb2, 0x0710;\n"  \
    "prmt.b32 q2, vc0, vc1, 0x0073;\n prmt.b32 q2, q2, vc2, 0x0710;\n"  \
    "prmt.b32 q3, vd0, vd1, 0x0073;\n prmt.b32 q3, q3, vd2, 0x0710;\n"  \
    "st.shared.b32 [sx], q0;\n st.shared.b32 [sx+128], q1;\n st.shared.b32 [sx+256], q2;\n st.shared.b32 [sx+384], q3;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "ld.shared.v4.b32 {q0, q1, q2, q3}, [sq];\n"                \
    "xor.b32 sx, sx, 512;\n xor.b32 sq, sq, 512;\n"             \
    "prmt.b32 q4, q0, q1, 0x4210;\n"                            \
    "prmt.b32 q5, q1, q2, 0x5421;\n"                            \
    "prmt.b32 q1, q2, q3, 0x6542;\n"                            \
    "add.u32 o, ez, %38;\n"                                     \
    "@pv st.shared.b32 [o], q4;\n"                              \
    "@pv st.shared.b32 [o+4], q5;\n"                            \
    "@pv st.shared.b32 [o+8], q1;\n"
// DIRECT stores (every destination row 4-byte aligned, no output tile, no store warp): the warp's 384 output bytes of
// a row change hands through the scratch so that lane t ends up with WORDS t, t + 32, t + 64 of them -- three
// fully coalesced 128-byte global stores per row.  %49 = address of the warp's first destination byte at row offset
// 0 (64-bit), %50 = 4 * lane, bits 0..2 of %53 = word k lies inside the strip.
//   QUAD mapping: packed words to P + 12 * lane (%51), back from P + 4 * lane (%52)
#define AWQ_DIRECT_ADDR                                         \
    "cvt.u64.u32 ro64, ez;\n"                                   \
    "add.u64 oa, %49, ro64;\n"                                  \
    "cvt.u64.u32 ro64, %50;\n"                                  \
    "add.u64 oa, oa, ro64;\n"
#define AWQ_EMIT_WD                                             \
    AWQ_VBLEND                                                  \
    "prmt.b32 q0, va0, va1, 0x0073;\n"                          \
    "prmt.b32 q1, va2, vb0, 0x0073;\n"                          \
    "prmt.b32 q2, vb1, vb2, 0x0073;\n"                          \
    "prmt.b32 q3, vc0, vc1, 0x0073;\n"                          \
    "prmt.b32 q4, vc2, vd0, 0x0073;\n"                          \
    "prmt.b32 q5, vd1, vd2, 0x0073;\n"                          \
    "prmt.b32 q0, q0, q1, 0x5410;\n"                            \
    "prmt.b32 q2, q2, q3, 0x5410;\n"                            \
    "prmt.b32 q4, q4, q5, 0x5410;\n"                            \
    "st.shared.b32 [pw], q0;\n st.shared.b32 [pw+4], q2;\n st.shared.b32 [pw+8], q4;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "ld.shared.b32 q0, [pr];\n ld.shared.b32 q1, [pr+128];\n ld.shared.b32 q2, [pr+256];\n" \
    AWQ_DIRECT_ADDR                                             \
    "xor.b32 pw, pw, 512;\n xor.b32 pr, pr, 512;\n"             \
    "@pw0 st.global.b32 [oa], q0;\n"                            \
    "@pw1 st.global.b32 [oa+128], q1;\n"                        \
    "@pw2 st.global.b32 [oa+256], q2;\n"
//   LANE mapping: RGBX pixels to X + 4 * lane (+ 128 j), word k of the lane = bytes of two adjacent RGBX pixels
//   (addresses xa0..2, byte selectors xs0..2: per-lane constants)
#define AWQ_EMIT_LD                                             \
    AWQ_VBLEND                                                  \
    "prmt.b32 q0, va0, va1, 0x0073;\n prmt.b32 q0, q0, va2, 0x0710;\n"  \
    "prmt.b32 q1, vb0, vb1, 0x0073;\n prmt.b32 q1, q1, vb2, 0x0710;\n"  \
    "prmt.b32 q2, vc0, vc1, 0x0073;\n prmt.b32 q2, q2, vc2, 0x0710;\n"  \
    "prmt.b32 q3, vd0, vd1, 0x0073;\n prmt.b32 q3, q3, vd2, 0x0710;\n"  \
    "st.shared.b32 [sx], q0;\n st.shared.b32 [sx+128], q1;\n st.shared.b32 [sx+256], q2;\n st.shared.b32 [sx+384], q3;\n" \
    "bar.warp.sync 0xffffffff;\n"                               \
    "ld.shared.b32 q0, [xa0];\n ld.shared.b32 q1, [xa0+4];\n"   \
    "ld.shared.b32 q2, [xa1];\n ld.shared.b32 q3, [xa1+4];\n"   \
    "ld.shared.b32 q4, [xa2];\n ld.shared.b32 q5, [xa2+4];\n"   \
    AWQ_DIRECT_ADDR                                             \
    "xor.b32 sx, sx, 512;\n xor.b32 xa0, xa0, 512;\n xor.b32 xa1, xa1, 512;\n xor.b32 xa2, xa2, 512;\n" \
    "prmt.b32 q0, q0, q1, xs0;\n"                               \
    "prmt.b32 q2, q2, q3, xs1;\n"                               \
    "prmt.b32 q4, q4, q5, xs2;\n"                               \
    "@pw0 st.global.b32 [oa], q0;\n"                            \
    "@pw1 st.global.b32 [oa+128], q2;\n"                        \
    "@pw2 st.global.b32 [oa+256], q4;\n"
// rows emitted after slot s (label prefix L keeps the two unrolled halves apart).  The entry of the row AFTER the
// one being emitted is requested before the emit, so the loop-carried compare never waits for a shared-memory load
// (the entry after the sentinel is read too: still inside the CTA's shared memory, never used)
#define AWQ_ROW_STEP(EMIT)                                      \
    "ld.shared.v4.b32 {fx, fy, fz, fw}, [rp+16];\n"             \
    EMIT                                                        \
    "add.u32 rp, rp, 16;\n"                                     \
    "mov.b32 ex, fx;\n mov.b32 ey, fy;\n mov.b32 ez, fz;\n mov.b32 ew, fw;\n"
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
    ".reg .b32 s, rp, ex, ey, ez, ew, fx, fy, fz, fw, o, sx, sq, pw, pr, xa0, xa1, xa2, xs0, xs1, xs2, tm;\n" \
    ".reg .pred pw0, pw1, pw2;\n"                               \
    ".reg .b64 ro64, oa;\n"                 \
    ".reg .b32 loa, mia, hia, lob, mib, hib, loc, mic, hic, lod, mid, hid;\n"   \
    ".reg .b32 Aa, Ba, Ab, Bb, Ac, Bc, Ad, Bd, Xa, Ya, Xb, Yb, Xc, Yc, Xd, Yd;\n" \
    ".reg .b32 ka, kb, kc, kd, ua, ub, uc, ud, sa, sb, sc, sd, ma, mb, mc, md, na, nb, nc, nd;\n" \
    ".reg .b32 wla, wha, wlb, whb, wlc, whc, wld, whd;\n"       \
    ".reg .b32 Ea0, Ea1, Ea2, Eb0, Eb1, Eb2, Ec0, Ec1, Ec2, Ed0, Ed1, Ed2;\n"   \
    ".reg .b32 Oa0, Oa1, Oa2, Ob0, Ob1, Ob2, Oc0, Oc1, Oc2, Od0, Od1, Od2;\n"   \
    ".reg .b32 va0, va1, va2, vb0, vb1, vb2, vc0, vc1, vc2, vd0, vd1, vd2;\n"   \
    ".reg .b32 q0, q1, q2, q3, q4, q5;\n"
// operands: %0-%11 E, %12-%23 O (read/write) | %24-%27 window address (a..d) | %28-%31 shift (a..d) |
//           %32 n_slots | %33, %34 LANE mapping: scratch addresses (write, read) | %35 unused | %36 pitch | %37 rp | %38 ocol | %39 first slot odd |
//           %40 store predicate | %41-%48 weight words (la, ha, lb, hb, lc, hc, ld, hd)
#define AWQ_PROLOGUE                                            \
    "mov.b32 Ea0, %0;\n mov.b32 Ea1, %1;\n mov.b32 Ea2, %2;\n mov.b32 Eb0, %3;\n mov.b32 Eb1, %4;\n mov.b32 Eb2, %5;\n" \
    "mov.b32 Ec0, %6;\n mov.b32 Ec1, %7;\n mov.b32 Ec2, %8;\n mov.b32 Ed0, %9;\n mov.b32 Ed1, %10;\n mov.b32 Ed2, %11;\n" \
    "mov.b32 Oa0, %12;\n mov.b32 Oa1, %13;\n mov.b32 Oa2, %14;\n mov.b32 Ob0, %15;\n mov.b32 Ob1, %16;\n mov.b32 Ob2, %17;\n" \
    "mov.b32 Oc0, %18;\n mov.b32 Oc1, %19;\n mov.b32 Oc2, %20;\n mov.b32 Od0, %21;\n mov.b32 Od1, %22;\n mov.b32 Od2, %23;\n" \
    "mov.b32 wla, %41;\n mov.b32 wha, %42;\n mov.b32 wlb, %43;\n mov.b32 whb, %44;\n"  \
    "mov.b32 wlc, %45;\n mov.b32 whc, %46;\n mov.b32 wld, %47;\n mov.b32 whd, %48;\n"  \
    "mov.b32 sx, %33;\n mov.b32 sq, %34;\n"                     \
    "mov.b32 pw, %51;\n mov.b32 pr, %52;\n"                     \
    "mov.b32 xa0, %54;\n mov.b32 xa1, %55;\n mov.b32 xa2, %56;\n"      \
    "mov.b32 xs0, %57;\n mov.b32 xs1, %58;\n mov.b32 xs2, %59;\n"      \
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
      "r"(n_slots), "r"(sx), "r"(sq), "r"(0), "r"(pitch), "r"(rp), "r"(ocol), "r"(odd_first), "r"(store_ok),  \
      "r"(wl[0]), "r"(wh[0]), "r"(wl[1]), "r"(wh[1]), "r"(wl[2]), "r"(wh[2]), "r"(wl[3]), "r"(wh[3]),       \
      "l"(d.obase), "r"(d.lane4), "r"(d.pw), "r"(d.pr), "r"(d.wmask), "r"(d.xa[0]), "r"(d.xa[1]), "r"(d.xa[2]),    \
      "r"(d.xs[0]), "r"(d.xs[1]), "r"(d.xs[2])                                                              \
    : "memory"

// what the DIRECT emit variants need besides the tile variants' operands (see AWQ_EMIT_WD / AWQ_EMIT_LD)
struct DirectOps {
    uint64_t obase;
    uint32_t lane4, pw, pr, wmask, xa[3], xs[3];
};

// FIXED: win = word addresses, sh = shifts.  !FIXED: win = byte addresses (sh unused).
// LANE: false = QUAD mapping, true = LANE mapping.  DIRECT: rows go straight to global memory (three coalesced word
// stores per lane), else into the output tile (three aligned word stores at the QUAD position).
template <bool FIXED, bool LANE, bool DIRECT>
__device__ __forceinline__ void sweep_quad(uint32_t* E, uint32_t* O, const uint32_t* win, const uint32_t* sh,
                                           const uint32_t* wl, const uint32_t* wh, int n_slots, uint32_t pitch,
                                           uint32_t rp, uint32_t ocol, uint32_t odd_first, uint32_t store_ok,
                                           uint32_t sx, uint32_t sq, const DirectOps& d) {
#define AWQ_RUN(BODY, EMIT) asm volatile("{\n" AWQ_DECL AWQ_PROLOGUE BODY(EMIT) AWQ_EPILOGUE "}\n" AWQ_OPERANDS)
    if (FIXED) {
        if (!DIRECT) { if (!LANE) AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_W); else AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_L); }
        else { if (!LANE) AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_WD); else AWQ_RUN(AWQ_BODY_F, AWQ_EMIT_LD); }
    } else {
        if (!DIRECT) { if (!LANE) AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_W); else AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_L); }
        else { if (!LANE) AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_WD); else AWQ_RUN(AWQ_BODY_V, AWQ_EMIT_LD); }
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
    int out_pitch;           // bytes per row of an output tile
    int rows;                // output rows per chunk (<= kMaxRows)
    int stages, tiles;       // ring depths: source-row stages (chunks whose loads are in flight), output tiles
    int store_warps;         // store warps per CTA (rows of a tile are dealt round-robin to them)
    int wait_hint_ns;        // suspend-time hint of the mbarrier waits (0: plain try_wait polling)
    int roles_first;         // 1: the producer / store warps are the CTA's first warps (the consumers get the higher
                             // warp ids, which the issue arbiter favours), 0: they are its last warps
    int map_policy;          // 0: per warp and strip (LANE when the map's local scale would make QUAD loads conflict),
                             // 1: always QUAD, 2: LANE wherever word stores apply
    int dbg;                 // ATTWARP_REMAP_DBG experiments: 1 skip the sweep, 2 skip the tile stores
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
// blockDim.x = consumer threads (a multiple of 32; 4 output columns each) + 32 producer threads + 32 x store_warps
// store threads.
// Shared memory: [stages source arenas][tiles output tiles][stages chunk tables][tiles tile headers][mbarriers].
// Chunk c lives in source stage c % stages and output tile c % tiles.  mbarriers:
//   full[s]  producer -> consumers   table written, source rows landed (transaction bytes)
//   sfree[s] consumers -> producer   every consumer warp is done with the stage's rows and table
//   odone[o] consumers -> store warp every consumer warp has written its columns of the tile
//   ofree[o] store warp -> consumers the tile has been read out of shared memory
// DIRECT (every destination row of the launch is 4-byte aligned): no output tiles, no store warps -- the consumers
// write their rows to global memory themselves (AWQ_EMIT_WD / AWQ_EMIT_LD).
template <int MAXT, int MINB, bool DIRECT>
__global__ void __launch_bounds__(MAXT, MINB) remap_u8_quad_kernel(const QuadArgs a) {
    const int R = a.rows;
    const int kStages = a.stages, kTiles = DIRECT ? 0 : a.tiles;
    const int out_bytes = R * a.out_pitch;
    const int out_off0 = kStages * a.stage_bytes;
    const int tab_off0 = out_off0 + kTiles * out_bytes;
    const int ohdr_off0 = tab_off0 + kStages * kTabBytes;
    const int bar_off0 = ohdr_off0 + kTiles * 32;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t full_s = smem_s + (uint32_t)bar_off0;
    const uint32_t sfree_s = full_s + 8u * kStages;
    const uint32_t odone_s = sfree_s + 8u * kStages;
    const uint32_t ofree_s = odone_s + 8u * kTiles;
    const int n_store_warps = DIRECT ? 0 : a.store_warps;
    const int n_cons_warps = ((int)blockDim.x >> 5) - 1 - n_store_warps;
    // per consumer warp 2 KB of scratch at a 1 KB aligned shared address: two 512-byte RGBX buffers (LANE mapping)
    // and two 512-byte packed-row buffers (DIRECT stores), each pair toggled with xor 512
    const uint32_t scratch_s = (ofree_s + 8u * kTiles + 1023u) & ~1023u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_s + 8u * s, 1);
            mbar_init(sfree_s + 8u * s, n_cons_warps);
        }
        for (int s = 0; s < kTiles; ++s) {
            mbar_init(odone_s + 8u * s, n_cons_warps);
            mbar_init(ofree_s + 8u * s, n_store_warps);
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

    // warp roles: consumers 0 .. n_cons_warps-1, then the producer, then the store warps (logical indices)
    const int n_roles = 1 + n_store_warps;
    const int hw_warp = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    const int warp_idx = a.roles_first ? (hw_warp < n_roles ? n_cons_warps + hw_warp : hw_warp - n_roles) : hw_warp;
    const int lane = (int)threadIdx.x & 31;
    const int tid = warp_idx * 32 + lane;                  // logical thread index: consumers first
    const uint32_t hint = (uint32_t)a.wait_hint_ns;
    auto wait = [&](uint32_t bar, uint32_t parity) {
        if (hint != 0u) mbar_wait_hint(bar, parity, hint);
        else mbar_wait(bar, parity);
    };

    if (warp_idx >= n_cons_warps) {
        if (warp_idx == n_cons_warps) {
            // =========================== producer warp =========================================
            int st = 0;
            uint32_t ph = 0;                 // parity of the stage's current use
            bool first_copy = true;
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
                    row_bytes = (c_hi - c_lo + 1) * kC;
                    // one copy per row: [16-byte alignment head][span][alignment tail + window over-read], and a
                    // pitch congruent to the row pitch modulo 16 so that slot k keeps the phase of source row k
                    slot_pitch = ((row_bytes + 45 + 15) & ~15) + (int)(row_pitch & 15);
                }
                const int max_slots = min((a.stage_bytes - 64) / slot_pitch, 2 * R);
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

                    wait(sfree_s + 8u * st, ph ^ 1u);             // stage free again
                    if (lane < n_rows) {
                        const uint32_t wu = (uint32_t)wa << 14, wl_ = (uint32_t)(32 - wa) << 14;   // upper / lower tap
                        const bool up_even = (ra & 1) == 0;
                        // z: tile mode -> offset of the row in the output tile; DIRECT -> offset of the row's first
                        // byte (of this strip) from the image's first destination byte
                        st128(tab + kTabRows + 16 * lane,
                              make_uint4(up_even ? wu : wl_, up_even ? wl_ : wu,
                                         DIRECT ? (uint32_t)(gd - dimg)
                                                : (uint32_t)(lane * a.out_pitch) + (uint32_t)(gd & 12),
                                         (uint32_t)(ra + 1 - r_lo)));
                    } else if (lane == n_rows) {
                        st128(tab + kTabRows + 16 * lane, make_uint4(0u, 0u, 0u, kRowSentinel));
                    }
                    if (lane == 0) {
                        const uint32_t fl = seg_flags | ((r_lo & 1) ? kFlagOddFirst : 0u);
                        st128(tab, make_uint4((uint32_t)n_rows, (uint32_t)n_slots | (fl << 16),
                                              (uint32_t)slot_pitch, (uint32_t)phase0));
                        st128(tab + 16, make_uint4((uint32_t)img, (uint32_t)x_first, (uint32_t)y_cur, (uint32_t)c_lo));
                        const uintptr_t gbase = DIRECT ? dimg : gd;         // DIRECT: the image, else the chunk's first row
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
                    carry_row = n_rows > 0 ? ra_last + 1 : kNoCarry;
                    y_cur += max(n_rows, 1);
                    if (++st == kStages) { st = 0; ph ^= 1u; }
                }
                local = local_end;
            }
            // terminator
            wait(sfree_s + 8u * st, ph ^ 1u);
            if (lane == 0) {
                st128(tab_off0 + st * kTabBytes, make_uint4(0xffffffffu, 0u, 0u, 0u));
                mbar_arrive(full_s + 8u * st);
            }
        } else {
            // =============================== store warps =====================================
            // store warp k ships rows k, k + n_store_warps, ... of every tile
            const int sw = warp_idx - n_cons_warps - 1;
            int ot = 0;
            uint32_t ph = 0;
            for (;;) {
                wait(odone_s + 8u * ot, ph);                      // every consumer warp is through
                const uint4 hd = ld128(ohdr_off0 + 32 * ot);           // {n_rows, -, -, -}
                const int n_rows = (int)hd.x;
                if (n_rows < 0) break;
                if (n_rows > 0 && !(a.dbg & 2)) {
                    // Tile row i sits at  i * out_pitch + (address of its first destination byte & 12).
                    //   rows whose destination is 4-byte aligned: bulk store of the 16-byte aligned interior (shared
                    //     and global addresses have the same 16-byte phase), <= 15 head and tail bytes by byte stores;
                    //   other rows (odd widths): the interior is shifted by 1..3 bytes on its way out -- per 16-byte
                    //     destination chunk one aligned 128-bit load + the word before it, four funnel shifts, one
                    //     aligned 128-bit store (lanes take consecutive chunks: full-sector writes).
                    const uint4 hs = ld128(ohdr_off0 + 32 * ot + 16);  // {dst lo, dst hi, row bytes, dst pitch}
                    const int len = (int)hs.z;
                    const int64_t dpitch = (int64_t)hs.w;
                    const int obuf = out_off0 + ot * out_bytes;
                    uint8_t* g0 = reinterpret_cast<uint8_t*>(((uint64_t)hs.y << 32) | hs.x);
                    const bool plain = ((reinterpret_cast<uintptr_t>(g0) | (uintptr_t)len | (uintptr_t)dpitch) & 15) == 0;
                    const int my_row = sw + lane * n_store_warps;
                    if (my_row < n_rows) {
                        uint8_t* g = g0 + (int64_t)my_row * dpitch;
                        const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                        const int head = (16 - off) & 15;
                        const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                        if ((off & 3) == 0 && body > 0)
                            bulk_s2g(g + head, smem_s + (uint32_t)(obuf + my_row * a.out_pitch + off + head), (uint32_t)body);
                    }
                    bulk_commit();
                    if (!plain) {
                        for (int i = sw; i < n_rows; i += n_store_warps) {
                            uint8_t* g = g0 + (int64_t)i * dpitch;
                            const int off = (int)(reinterpret_cast<uintptr_t>(g) & 15);
                            const int head = min((16 - off) & 15, len);
                            const int body = (len - head) > 0 ? ((len - head) & ~15) : 0;
                            const int s = obuf + i * a.out_pitch + (off & 12);       // the row's first byte
                            // <= 15 head bytes and <= 15 tail bytes, one lane per byte
                            if (lane < 16) {
                                if (lane < head) g[lane] = smem[s + lane];
                            } else {
                                const int qq = head + body + (lane - 16);
                                if (qq < len) g[qq] = smem[s + qq];
                            }
                            const int r = off & 3;
                            if (r != 0) {
                                // destination chunk c = bytes [head + 16 c, + 16) of the row = shared bytes
                                // [A - r, A - r + 16) with A = s + head + 16 c + r a multiple of 16.  Four chunks per
                                // lane and pass, loads first; the word before a chunk is the last word of the chunk
                                // of the lane below (lane 0 reads its own).
                                const uint32_t shift = 8u * (uint32_t)(4 - r);
                                const int nch = body >> 4;
                                const int A0 = s + head + r;
                                uint8_t* gp = g + head;
                                for (int c0 = 0; c0 < nch; c0 += 128) {
                                    uint4 w[4];
                                    uint32_t wm[4];
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int c = c0 + 32 * k + lane;
                                        w[k] = c < nch ? ld128(A0 + 16 * c) : make_uint4(0u, 0u, 0u, 0u);
                                    }
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        wm[k] = __shfl_up_sync(0xffffffffu, w[k].w, 1);
                                        if (lane == 0) wm[k] = ld32(A0 + 16 * (c0 + 32 * k) - 4);
                                    }
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int c = c0 + 32 * k + lane;
                                        uint4 o;
                                        o.x = __funnelshift_r(wm[k], w[k].x, shift);
                                        o.y = __funnelshift_r(w[k].x, w[k].y, shift);
                                        o.z = __funnelshift_r(w[k].y, w[k].z, shift);
                                        o.w = __funnelshift_r(w[k].z, w[k].w, shift);
                                        if (c < nch) *reinterpret_cast<uint4*>(gp + 16 * c) = o;
                                    }
                                }
                            }
                        }
                    }
                    bulk_wait_read0();                                 // the tile has left shared memory
                }
                __syncwarp();
                if (lane == 0 && sw == 0 && a.trace != nullptr) {
                    if (a.trace[(size_t)blockIdx.x * 8 + 4] == 0ull) trace_stamp(a, 4);   // first tile shipped
                    trace_stamp(a, 5);                                                     // latest tile shipped
                }
                if (lane == 0) mbar_arrive(ofree_s + 8u * ot);        // tile free for the consumers
                if (++ot == kTiles) { ot = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // =============================== consumer warps ==============================================
    // this thread's output columns inside the strip.  QUAD mapping: 4 tid .. 4 tid + 3.  LANE mapping: columns
    // 32 j + lane of the warp's 128-column block (the STORES are those of the QUAD mapping in both cases).
    const int x0 = tid * 4;
    const int xw = (tid & ~31) * 4;   // first column of the warp's block
    int wo[4];                        // byte offset of each column's window inside a staged row span
    uint32_t wl[4], wh[4], E[12], O[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) E[k] = O[k] = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) { wo[j] = 0; wl[j] = wh[j] = 0u; }
    bool warp_live = false;           // some lane of this warp owns a column of the strip
    bool lane_map = false;            // this warp runs the LANE mapping in the current strip
    uint32_t store_ok = 0u;           // this thread stores at least one column (QUAD position)
    const int out_col = x0 * kC;
    const uint32_t sx_s = scratch_s + (uint32_t)(warp_idx * 2048 + lane * 4);     // scratch: my RGBX pixels in
    const uint32_t sq_s = scratch_s + (uint32_t)(warp_idx * 2048 + lane * 16);    //          my four adjacent pixels out
    DirectOps dops{};
    if (DIRECT) {
        dops.lane4 = 4u * (uint32_t)lane;
        dops.pw = scratch_s + (uint32_t)(warp_idx * 2048 + 1024 + lane * 12);     // packed row: my 12 bytes in
        dops.pr = scratch_s + (uint32_t)(warp_idx * 2048 + 1024 + lane * 4);      //             my three words out
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // output word w = lane + 32 k of the warp's row = bytes 4 w .. 4 w + 3 of the RGB stream: pixel
            // p0 = floor(4 w / 3) (and the next one), starting at channel 4 w - 3 p0
            const int w = lane + 32 * k, p0 = (4 * w) / 3, c0 = 4 * w - 3 * p0;
            dops.xa[k] = scratch_s + (uint32_t)(warp_idx * 2048 + 4 * p0);
            dops.xs[k] = c0 == 0 ? 0x4210u : (c0 == 1 ? 0x5421u : 0x6542u);
        }
    }
    int xba[4];                       // source column of each of my pixels' left tap (-1: none yet)

    // per-strip setup: taps and weights of this thread's columns, choice of the mapping
    auto setup_strip = [&](const float* mx, int W, int ncols) {
        store_ok = x0 < ncols ? 1u : 0u;
        warp_live = __any_sync(0xffffffffu, store_ok != 0u);
        int xb = -1, xb_first = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // a column past the strip's end keeps the previous column's window and gets zero weights
            int w0 = 0, w1 = 0;
            if (x0 + j < ncols) column_taps(__ldg(mx + x0 + j), W, xb, w0, w1);
            if (j == 0) xb_first = xb;
            xba[j] = xb;
            wl[j] = (uint32_t)w0 | ((uint32_t)w1 << 8);
            wh[j] = wl[j] << 16;
        }
        // QUAD loads are conflict-free only while the lanes' windows stay 3 words apart: when the source column
        // of some lane's first pixel has drifted two or more pixels from "4 per lane", switch the warp to the
        // LANE mapping
        lane_map = false;
        if (warp_live && a.map_policy != 1) {
            const int xb_lane0 = __shfl_sync(0xffffffffu, xb_first, 0);
            const int dev = store_ok ? abs(xb_first - xb_lane0 - 4 * lane) : 0;
            lane_map = a.map_policy == 2 || __any_sync(0xffffffffu, dev >= 2);
        }
        if (lane_map) {
            xb = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = xw + 32 * j + lane;
                int w0 = 0, w1 = 0;
                if (x < ncols) column_taps(__ldg(mx + x), W, xb, w0, w1);
                xba[j] = xb;
                wl[j] = (uint32_t)w0 | ((uint32_t)w1 << 8);
                wh[j] = wl[j] << 16;
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
                setup_strip(v.mx + x_first, v.W, min(v.strip_cols, v.Wo - x_first));
                pre_img = img;
                pre_x_first = x_first;
            }
        }
    }

    int st = 0, ot = 0;               // source stage / output tile of the current chunk
    uint32_t sph = 0u, oph = 1u;      // parities to wait for: stage filled / tile shipped and free
    for (;; st = st + 1 == kStages ? 0 : st + 1, sph ^= st == 0 ? 1u : 0u,
            ot = ot + 1 == kTiles ? 0 : ot + 1, oph ^= ot == 0 ? 1u : 0u) {
        const int tab = tab_off0 + st * kTabBytes;
        wait(full_s + 8u * st, sph);
        if (tid == 0 && a.trace != nullptr && a.trace[(size_t)blockIdx.x * 8 + 2] == 0ull) trace_stamp(a, 2);   // first rows landed
        const uint4 h0 = ld128(tab);
        const int n_rows = (int)h0.x;
        if (!DIRECT) wait(ofree_s + 8u * ot, oph);                          // tile shipped and free
        if (n_rows < 0) {                                                        // pass the stop on
            if (!DIRECT) {
                if (tid == 0) st128(ohdr_off0 + 32 * ot, make_uint4(0xffffffffu, 0u, 0u, 0u));
                __syncwarp();
                if (lane == 0) mbar_arrive(odone_s + 8u * ot);
            }
            break;
        }
        const uint32_t flags = h0.y >> 16;
        if (flags & kFlagNewStrip) {                         // new strip: per-column taps and weights
            const uint4 h1 = ld128(tab + 16);
            if ((int)h1.x != pre_img || (int)h1.y != pre_x_first) {
                const uint4 hx = ld128(tab + kTabStrip);
                const float* mx = reinterpret_cast<const float*>(((uint64_t)hx.y << 32) | hx.x);
                setup_strip(mx, (int)hx.z, (int)hx.w);
            }
            pre_img = -1;                                    // the early setup serves the first segment only
#pragma unroll
            for (int j = 0; j < 4; ++j) wo[j] = xba[j] < 0 ? 0 : (xba[j] - (int)h1.w) * kC;
        }
        if (!DIRECT && tid == 0) {                           // what the store warp needs to ship the tile
            st128(ohdr_off0 + 32 * ot, make_uint4(h0.x, 0u, 0u, 0u));
            st128(ohdr_off0 + 32 * ot + 16, ld128(tab + kTabStore));
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
            const uint32_t ocol_s = smem_s + (uint32_t)(out_off0 + ot * out_bytes + out_col);
            const uint32_t base = smem_s + (uint32_t)(st * a.stage_bytes) + h0.w;    // first byte of slot 0
            const uint32_t odd = (flags & kFlagOddFirst) ? 1u : 0u;
            uint32_t win[4], sh[4];
            if (DIRECT) {
                // the warp's 384 bytes of a row start 384 * warp bytes after the strip's first byte
                const uint4 hs = ld128(tab + kTabStore);                         // {image lo, hi, strip row bytes, -}
                dops.obase = (((uint64_t)hs.y << 32) | hs.x) + (uint64_t)(warp_idx * 384);
                const int nb = (int)hs.z - warp_idx * 384;                       // bytes of the warp's block in the strip
                dops.wmask = (4 * lane + 4 <= nb ? 1u : 0u) | (4 * lane + 132 <= nb ? 2u : 0u) | (4 * lane + 260 <= nb ? 4u : 0u);
            }
            if (flags & kFlagFixedShift) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t b = base + (uint32_t)wo[j];
                    win[j] = b & ~3u;
                    sh[j] = b << 3;
                }
                if (!lane_map) sweep_quad<true, false, DIRECT>(E, O, win, sh, wl, wh, n_slots, h0.z, rp_s, ocol_s, odd, store_ok, sx_s, sq_s, dops);
                else sweep_quad<true, true, DIRECT>(E, O, win, sh, wl, wh, n_slots, h0.z, rp_s, ocol_s, odd, store_ok, sx_s, sq_s, dops);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { win[j] = base + (uint32_t)wo[j]; sh[j] = 0u; }
                if (!lane_map) sweep_quad<false, false, DIRECT>(E, O, win, sh, wl, wh, n_slots, h0.z, rp_s, ocol_s, odd, store_ok, sx_s, sq_s, dops);
                else sweep_quad<false, true, DIRECT>(E, O, win, sh, wl, wh, n_slots, h0.z, rp_s, ocol_s, odd, store_ok, sx_s, sq_s, dops);
            }
        }
        // publish this warp's part of the tile to the async proxy, then count the warp in
        if (!DIRECT) fence_proxy_async();
        __syncwarp();
        if (tid == 0 && a.trace != nullptr) {
            if (a.trace[(size_t)blockIdx.x * 8 + 3] == 0ull) trace_stamp(a, 3);   // first chunk swept (warp 0)
            trace_stamp(a, 6);                                                     // latest chunk swept
        }
        if (lane == 0) {
            mbar_arrive(sfree_s + 8u * st);
            if (!DIRECT) mbar_arrive(odone_s + 8u * ot);
        }
    }
}

// ---- launch geometry ------------------------------------------------------------------------------
// A configuration = consumer warps per CTA (each covers 128 output columns) and CTAs per SM; the rows per chunk
// follow from the shared memory that leaves.  Registers: the quad sweep needs ~120, so about 16 warps fit an SM.
struct Geometry {
    int warps;          // consumer warps per CTA
    int ctas;           // CTAs per SM the kernel is built for
    int max_cols;       // widest strip (multiple of 16)
    int store_warps;    // store warps per CTA
};
constexpr Geometry kGeo[3] = {{3, 4, 352, 1}, {6, 2, 704, 2}, {11, 1, 1408, 4}};      // with output tiles
constexpr Geometry kGeoD[3] = {{3, 5, 352, 0}, {6, 2, 704, 0}, {11, 1, 1408, 0}};     // DIRECT stores

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

// geometry index for strips of `Wo`-wide outputs: the narrowest configuration that takes the image in one strip,
// else the widest (ATTWARP_QUAD_GEO = 0 .. 2 forces one: tuning experiments)
int pick_geometry(int Wo) {
    const int forced = env_int("ATTWARP_QUAD_GEO", -1);
    if (forced >= 0 && forced <= 2) return forced;
    for (int g = 0; g < 3; ++g)
        if (Wo <= kGeo[g].max_cols) return g;
    return 2;
}

template <int G, bool DIRECT>
int launch_geo(QuadArgs& a, int cols, cudaStream_t st) {
    constexpr Geometry geo = DIRECT ? kGeoD[G] : kGeo[G];
    constexpr int kThreads = (geo.warps + 1 + geo.store_warps) * 32;
    a.store_warps = geo.store_warps;
    auto kern = remap_u8_quad_kernel<kThreads, geo.ctas, DIRECT>;
    a.dbg = env_int("ATTWARP_REMAP_DBG", 0);
    a.out_pitch = (cols * kC + 15 + 15) & ~15;              // + the 16-byte phase of the destination
    const int unit_pitch = (((cols + 1) * kC + 45 + 15) & ~15) + 16;      // a slot at unit scale
    // ring depths (ATTWARP_QUAD_RING = stages * 10 + tiles, tuning experiments) and rows per chunk: as many as the
    // shared memory of 1 / ctas of an SM holds (stages of R + 2 slots, tiles of R rows), at most kMaxRows
    int stages = 2, tiles = DIRECT ? 0 : 2;
    {
        const int ring = env_int("ATTWARP_QUAD_RING", 0);
        if (ring / 10 >= 2 && ring / 10 <= kMaxRing && ring % 10 >= 2 && ring % 10 <= kMaxRing) { stages = ring / 10; tiles = DIRECT ? 0 : ring % 10; }
    }
    a.stages = stages;
    a.tiles = tiles;
    const int scratch = 2048 * geo.warps + 1024;
    const int budget = (227 * 1024) / geo.ctas - 1024 - stages * kTabBytes - tiles * 32 - 16 * kMaxRing - 256 - scratch;
    int R = (budget - stages * (2 * unit_pitch + 64 + 128)) / (stages * unit_pitch + tiles * a.out_pitch);
    R = R > kMaxRows ? kMaxRows : R;
    const int forced = env_int("ATTWARP_QUAD_ROWS", 0);
    if (forced >= 2 && forced <= R) R = forced;
    if (R < 2) return fail(ATTWARP_ERR_UNSUPPORTED, "remap: strips of %d columns do not fit shared memory", cols);
    a.rows = R;
    a.stage_bytes = ((R + 2) * unit_pitch + 64 + 127) & ~127;
    const size_t smem_bytes = (size_t)stages * (a.stage_bytes + kTabBytes) + (size_t)tiles * ((size_t)R * a.out_pitch + 32) +
                              2 * (size_t)(stages + tiles) * sizeof(uint64_t) + 16 + (size_t)scratch;
    a.map_policy = env_int("ATTWARP_QUAD_MAP", 0);          // 0 auto, 1 QUAD only, 2 LANE wherever possible
    a.wait_hint_ns = env_int("ATTWARP_QUAD_WAIT_HINT", 0);
    a.roles_first = env_int("ATTWARP_QUAD_ROLES_FIRST", 0);
    struct Cfg { size_t smem; int dev, occ; };
    static thread_local Cfg c = {0, -1, 0};
    int dev = 0;
    AW_CUDA(cudaGetDevice(&dev));
    if (c.smem != smem_bytes || c.dev != dev) {
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int o = 0;
        AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem_bytes));
        const int cap = env_int("ATTWARP_REMAP_CTAS_PER_SM", 0);
        if (cap >= 1 && cap < o) o = cap;
        c = Cfg{smem_bytes, dev, o};
    }
    if (c.occ < 1) return fail(ATTWARP_ERR_CUDA, "remap: kernel does not fit an SM (%zu B shared)", smem_bytes);
    // attwarp_set_sm_share(2): leave half of every SM to the kernels of another stream
    const int share = sm_share();
    const int occ = c.occ >= 2 * share ? c.occ / share : (c.occ >= 2 && share > 1 ? c.occ / 2 : c.occ);
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
    kern<<<grid, kThreads, smem_bytes, st>>>(a);
    if (a.trace != nullptr) {
        std::vector<unsigned long long> h((size_t)grid * 8);
        AW_CUDA(cudaStreamSynchronize(st));
        AW_CUDA(cudaMemcpy(h.data(), a.trace, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(trace_path, "a")) {
            fprintf(f, "launch grid=%d threads=%d rows=%d smem=%zu\n", grid, kThreads, a.rows, smem_bytes);
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

// direct: every destination row of the launch is 4-byte aligned (no output tile, consumers store to global memory)
int launch_by_geometry(int g, bool direct, QuadArgs& a, int cols, cudaStream_t st) {
    if (direct) {
        switch (g) {
            case 0: return launch_geo<0, true>(a, cols, st);
            case 1: return launch_geo<1, true>(a, cols, st);
            default: return launch_geo<2, true>(a, cols, st);
        }
    }
    switch (g) {
        case 0: return launch_geo<0, false>(a, cols, st);
        case 1: return launch_geo<1, false>(a, cols, st);
        default: return launch_geo<2, false>(a, cols, st);
    }
}

// ATTWARP_QUAD_DIRECT=0 keeps the output tiles + store warps for aligned images too (A/B comparisons)
bool direct_allowed() { return env_int("ATTWARP_QUAD_DIRECT", 1) != 0; }
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
    const int g = pick_geometry(Wo);
    const StripPlan sp = plan_strips(Wo, kGeo[g].max_cols);
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
    return launch_by_geometry(g, direct_allowed() && rows_word_aligned(dst, Wo), a, Wo < a.strip_cols ? Wo : a.strip_cols, st);
}

// Ragged batch.  Images are grouped into width classes, one launch per class with the geometry that fits it
// (a 300-wide image in a CTA built for 1408 columns would leave eight of its eleven consumer warps idle):
//   class 0: Wo <= 352 (3 consumer warps per CTA), class 1: Wo <= 704 (6), class 2: wider (11; strips of <= 1408);
//   classes 3..5: the same widths for images whose destination rows are not 4-byte aligned (output tiles + store
//   warps instead of direct stores).
// Step 1: `host` (n + 1 entries, batch order) gets each image's strip plan and is uploaded to dev_main (the maps
// kernel reads shapes and map pointers from it); a copy grouped by class, every group followed by an entry that
// carries its unit total, is uploaded to dev_sorted (n + 3 entries).
constexpr int kClassGeo[3] = {0, 1, 2};
inline int width_class(int Wo) { return Wo <= kGeo[kClassGeo[0]].max_cols ? 0 : (Wo <= kGeo[kClassGeo[1]].max_cols ? 1 : 2); }

int launch_remap_u8_quad_ragged_prepare(RaggedImage* host, int n, RaggedImage* dev_main, RaggedImage* dev_sorted,
                                        RaggedQuadPlan* plan, cudaStream_t st) {
    static thread_local std::vector<RaggedImage> sorted;
    sorted.assign((size_t)n + kRaggedClasses, RaggedImage{});
    *plan = RaggedQuadPlan{};
    const int forced = env_int("ATTWARP_QUAD_GEO", -1);
    const bool direct_ok = direct_allowed();
    auto cls = [&](const RaggedImage& im) {
        const int wc = forced >= 0 ? 2 : width_class(im.Wo);
        return wc + ((direct_ok && rows_word_aligned(im.dst, im.Wo)) ? 0 : 3);
    };
    for (int i = 0; i < n; ++i) plan->count[cls(host[i])]++;
    int pos = 0;
    for (int c = 0; c < kRaggedClasses; ++c) {
        plan->offset[c] = pos;
        plan->geo[c] = forced >= 0 && forced <= 2 ? forced : kClassGeo[c % 3];
        plan->direct[c] = c < 3 ? 1 : 0;
        pos += plan->count[c] + 1;
    }
    int fill[kRaggedClasses] = {};
    int64_t total[kRaggedClasses] = {};
    for (int i = 0; i < n; ++i) {
        const int c = cls(host[i]);
        const StripPlan sp = plan_strips(host[i].Wo, kGeo[plan->geo[c]].max_cols);
        const int cols = host[i].Wo < sp.strip_cols ? host[i].Wo : sp.strip_cols;
        const int units = (cols + 127) / 128;
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
// Step 2: one launch per non-empty class.
int launch_remap_u8_quad_ragged_run(const RaggedQuadPlan& plan, const RaggedImage* dev_sorted, cudaStream_t st) {
    for (int c = 0; c < kRaggedClasses; ++c) {
        if (plan.count[c] == 0 || plan.total_units[c] == 0) continue;
        QuadArgs a{};
        a.imgs = dev_sorted + plan.offset[c];
        a.n_img = plan.count[c];
        a.total_units = plan.total_units[c];
        const int rc = launch_by_geometry(plan.geo[c], plan.direct[c] != 0, a, plan.max_strip[c], st);
        if (rc != ATTWARP_OK) return rc;
    }
    return ATTWARP_OK;
}

}  // namespace aw
