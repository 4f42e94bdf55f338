// remap_f32_stream.cu -- stage 5 for float32 images (the torch path: warp_from_cdf_torch on float images,
// "model/marginalnet_full_dataset/checkpoint_utils.py":195-198; BASELINE configs[4]): persistent, warp-specialised
// streaming resample.
//
// Same arithmetic as remap_f32_rows_kernel / remap_direct_kernel (OpenCV's float path, see warp_math.h): weights
// w = fl32(wy * wx), pixel = ((p00 w00 + p01 w01) + p10 w10) + p11 w11 in float32 with no FMA contraction, taps
// clamped to the image (BORDER_REPLICATE) -- bit-equal to cv2.remap.
//
// The round-1 kernel gathered its taps from global memory through L1 with ~50 compiler-generated instructions per
// output element; it was bound by issue slots, not by HBM.  This one has the skeleton of remap_quad.cu:
//   * a producer warp plans chunks of output rows and fetches the contiguous range of source rows they tap with
//     cp.async.bulk into a ring of shared-memory stages (one copy per chunk for strips that span whole rows: c5's
//     512-float rows are contiguous), uniform slot pitch for any alignment;
//   * consumer threads own four output ELEMENTS (floats) 32 apart -- lane t of warp w: elements 128 w + 32 j + t of
//     the strip, so that the shared-memory loads of a warp are consecutive words and the global stores of a warp are
//     128 contiguous bytes (no output tile, no store warp: float stores coalesce by themselves);
//   * the two taps of an element on the latest EVEN and the latest ODD source row stay in registers (a new source row
//     overwrites one pair: two shared loads per element and NEW source row), an output row is 4 weight products + 4
//     products + 3 sums per element, in cv2's association order -- which of the even / odd pair is the upper row
//     comes with the row's table entry (four straight-line variants, uniform branch);
//   * the sweep is hand-written PTX with bra.uni like the uint8 kernels: ~80 instructions per 128 elements and row.
// Interleaved float images (HWC, E floats per pixel) are the same kernel: an element's taps are E floats apart.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bulk_ptx.cuh"
#include "common.cuh"

namespace aw {
namespace {

using namespace ptx;

constexpr int kStages = 2;
constexpr int kMaxRows = 16;

extern __shared__ __align__(128) uint8_t smem[];
__device__ __forceinline__ uint4 ld128(int off) { return *reinterpret_cast<const uint4*>(smem + off); }
__device__ __forceinline__ void st128(int off, uint4 v) { *reinterpret_cast<uint4*>(smem + off) = v; }

// ---- per-stage chunk table (byte offsets) -----------------------------------------------------------
//   +0   uint4 {n_rows, n_slots | flags << 16, slot_pitch (bytes), byte offset of slot 0's first byte in the arena}
//              n_rows 0: output row y0 takes the direct path; -1: stop
//   +16  uint4 {plane, first element of the strip, y0, first staged source ELEMENT of a row (c_lo * E)}
//   +32  uint4 {address of the plane's first output float (lo, hi), elements in the strip, -}
//   +48  uint4 {address of map_x (row of this plane's image) (lo, hi), W, -}                  (new strip only)
//   +64  uint4 row[kMaxRows + 1]:  x = bits of wy of the UPPER tap row (1 - fy), y = of the LOWER one (fy)
//                                  z = byte offset of the output row inside the plane | variant
//                                      (variant = parity of the upper source row | parity of the lower one << 1:
//                                       which of the even / odd register pairs holds each)
//                                  w = slot after which the row is emitted (= slot of its LOWER tap row);
//                                      0xffffffff: both rows are held already; sentinel after the last row
constexpr int kTabOut = 32, kTabStrip = 48, kTabRows = 64;
constexpr int kTabBytes = kTabRows + 16 * (kMaxRows + 1);
constexpr uint32_t kRowSentinel = 0x7fffffffu;
constexpr uint32_t kFlagNewStrip = 1u, kFlagOddFirst = 8u;
constexpr int kNoCarry = -(1 << 29);

// ---- the sweep over one chunk (PTX) -----------------------------------------------------------------
// columns a..d; P?0 / P?1: taps (left, right) of the latest row of parity P in {E, O}; wx?0 / wx?1: 1 - fx, fx
#define AWF_LOAD1(P, J)                                         \
    "ld.shared.f32 " #P #J "0, [k" #J "0];\n"                   \
    "ld.shared.f32 " #P #J "1, [k" #J "1];\n"                   \
    "add.u32 k" #J "0, k" #J "0, %33;\n"                        \
    "add.u32 k" #J "1, k" #J "1, %33;\n"
#define AWF_LOAD(P) AWF_LOAD1(P, a) AWF_LOAD1(P, b) AWF_LOAD1(P, c) AWF_LOAD1(P, d)
#define AWF_W1(J)                                               \
    "mul.rn.f32 w0" #J ", ex, wx" #J "0;\n"                     \
    "mul.rn.f32 w1" #J ", ex, wx" #J "1;\n"                     \
    "mul.rn.f32 w2" #J ", ey, wx" #J "0;\n"                     \
    "mul.rn.f32 w3" #J ", ey, wx" #J "1;\n"
// ((U0 w00 + U1 w01) + L0 w10) + L1 w11, stored 128 * IDX bytes after the thread's first element
#define AWF_ACC1(U, L, J, IDX)                                  \
    "mul.rn.f32 t0, " #U #J "0, w0" #J ";\n"                    \
    "mul.rn.f32 t1, " #U #J "1, w1" #J ";\n"                    \
    "add.rn.f32 t0, t0, t1;\n"                                  \
    "mul.rn.f32 t1, " #L #J "0, w2" #J ";\n"                    \
    "add.rn.f32 t0, t0, t1;\n"                                  \
    "mul.rn.f32 t1, " #L #J "1, w3" #J ";\n"                    \
    "add.rn.f32 t0, t0, t1;\n"                                  \
    "@pv" #J " st.global.f32 [oa+" #IDX "], t0;\n"
#define AWF_ACC(U, L) AWF_ACC1(U, L, a, 0) AWF_ACC1(U, L, b, 128) AWF_ACC1(U, L, c, 256) AWF_ACC1(U, L, d, 384)
// one output row: address, weights, then the variant (which pair is the upper row)
#define AWF_EMIT(T)                                             \
    "and.b32 var, ez, 3;\n"                                     \
    "and.b32 ro, ez, 0xfffffffc;\n"                             \
    "cvt.u64.u32 ro64, ro;\n"                                   \
    "add.u64 oa, %36, ro64;\n"                                  \
    AWF_W1(a) AWF_W1(b) AWF_W1(c) AWF_W1(d)                     \
    "setp.eq.u32 q, var, 0;\n @q bra.uni " T "_EE;\n"           \
    "setp.eq.u32 q, var, 1;\n @q bra.uni " T "_OE;\n"           \
    "setp.eq.u32 q, var, 2;\n @q bra.uni " T "_EO;\n"           \
    AWF_ACC(O, O) "bra.uni " T "_END;\n"                        \
    T "_EE:\n" AWF_ACC(E, E) "bra.uni " T "_END;\n"             \
    T "_OE:\n" AWF_ACC(O, E) "bra.uni " T "_END;\n"             \
    T "_EO:\n" AWF_ACC(E, O)                                    \
    T "_END:\n"
// the entry of the row after the one being emitted is requested before the emit
#define AWF_ROW_STEP(T)                                         \
    "ld.shared.v4.b32 {fx, fy, fz, fw}, [rp+16];\n"             \
    AWF_EMIT(T)                                                 \
    "add.u32 rp, rp, 16;\n"                                     \
    "mov.b32 ex, fx;\n mov.b32 ey, fy;\n mov.b32 ez, fz;\n mov.b32 ew, fw;\n"
#define AWF_ROWS(L)                                             \
    "setp.ne.u32 q, ew, s;\n"                                   \
    "@q bra.uni " L "_NEXT;\n"                                  \
    L "_ROW:\n" AWF_ROW_STEP(L "V")                             \
    "setp.eq.u32 q, ew, s;\n"                                   \
    "@q bra.uni " L "_ROW;\n"                                   \
    L "_NEXT:\n"                                                \
    "add.s32 s, s, 1;\n"
#define AWF_BODY                                                \
    "{\n"                                                       \
    ".reg .pred p, q, podd, pva, pvb, pvc, pvd;\n"              \
    ".reg .b32 s, rp, ex, ey, ez, ew, fx, fy, fz, fw, var, ro, t;\n"   \
    ".reg .b32 ka0, ka1, kb0, kb1, kc0, kc1, kd0, kd1;\n"       \
    ".reg .f32 Ea0, Ea1, Eb0, Eb1, Ec0, Ec1, Ed0, Ed1, Oa0, Oa1, Ob0, Ob1, Oc0, Oc1, Od0, Od1;\n" \
    ".reg .f32 wxa0, wxa1, wxb0, wxb1, wxc0, wxc1, wxd0, wxd1;\n"      \
    ".reg .f32 w0a, w1a, w2a, w3a, w0b, w1b, w2b, w3b, w0c, w1c, w2c, w3c, w0d, w1d, w2d, w3d, t0, t1;\n" \
    ".reg .b64 ro64, oa;\n"                                     \
    "mov.f32 Ea0, %0;\n mov.f32 Ea1, %1;\n mov.f32 Eb0, %2;\n mov.f32 Eb1, %3;\n"     \
    "mov.f32 Ec0, %4;\n mov.f32 Ec1, %5;\n mov.f32 Ed0, %6;\n mov.f32 Ed1, %7;\n"     \
    "mov.f32 Oa0, %8;\n mov.f32 Oa1, %9;\n mov.f32 Ob0, %10;\n mov.f32 Ob1, %11;\n"   \
    "mov.f32 Oc0, %12;\n mov.f32 Oc1, %13;\n mov.f32 Od0, %14;\n mov.f32 Od1, %15;\n" \
    "mov.f32 wxa0, %16;\n mov.f32 wxa1, %17;\n mov.f32 wxb0, %18;\n mov.f32 wxb1, %19;\n" \
    "mov.f32 wxc0, %20;\n mov.f32 wxc1, %21;\n mov.f32 wxd0, %22;\n mov.f32 wxd1, %23;\n" \
    "mov.b32 ka0, %24;\n mov.b32 ka1, %25;\n mov.b32 kb0, %26;\n mov.b32 kb1, %27;\n" \
    "mov.b32 kc0, %28;\n mov.b32 kc1, %29;\n mov.b32 kd0, %30;\n mov.b32 kd1, %31;\n" \
    "and.b32 t, %37, 1;\n setp.ne.u32 pva, t, 0;\n"             \
    "and.b32 t, %37, 2;\n setp.ne.u32 pvb, t, 0;\n"             \
    "and.b32 t, %37, 4;\n setp.ne.u32 pvc, t, 0;\n"             \
    "and.b32 t, %37, 8;\n setp.ne.u32 pvd, t, 0;\n"             \
    "setp.ne.u32 podd, %35, 0;\n"                               \
    "mov.b32 rp, %34;\n"                                        \
    "ld.shared.v4.b32 {ex, ey, ez, ew}, [rp];\n"                \
    "setp.ne.u32 q, ew, 0xffffffff;\n"                          \
    "@q bra.uni PRE_DONE;\n"                                    \
    "PRE_ROW:\n" AWF_ROW_STEP("PRV")                            \
    "setp.eq.u32 q, ew, 0xffffffff;\n"                          \
    "@q bra.uni PRE_ROW;\n"                                     \
    "PRE_DONE:\n"                                               \
    "mov.b32 s, 0;\n"                                           \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@!p bra.uni DONE;\n"                                       \
    "@podd bra.uni ODD;\n"                                      \
    "EVEN:\n" AWF_LOAD(E) AWF_ROWS("EV")                        \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@!p bra.uni DONE;\n"                                       \
    "ODD:\n" AWF_LOAD(O) AWF_ROWS("OD")                         \
    "setp.lt.s32 p, s, %32;\n"                                  \
    "@p bra.uni EVEN;\n"                                        \
    "DONE:\n"                                                   \
    "mov.f32 %0, Ea0;\n mov.f32 %1, Ea1;\n mov.f32 %2, Eb0;\n mov.f32 %3, Eb1;\n"     \
    "mov.f32 %4, Ec0;\n mov.f32 %5, Ec1;\n mov.f32 %6, Ed0;\n mov.f32 %7, Ed1;\n"     \
    "mov.f32 %8, Oa0;\n mov.f32 %9, Oa1;\n mov.f32 %10, Ob0;\n mov.f32 %11, Ob1;\n"   \
    "mov.f32 %12, Oc0;\n mov.f32 %13, Oc1;\n mov.f32 %14, Od0;\n mov.f32 %15, Od1;\n" \
    "}\n"

// P[0..7] = even-row taps (a0 a1 b0 b1 c0 c1 d0 d1), P[8..15] = odd-row taps; wx[2 j], wx[2 j + 1] = 1 - fx, fx;
// k[2 j], k[2 j + 1] = shared addresses of the two taps of column j in slot 0
__device__ __forceinline__ void sweep_f32(float* P, const float* wx, const uint32_t* k, int n_slots, uint32_t pitch,
                                          uint32_t rp, uint32_t odd_first, uint64_t obase, uint32_t store_mask) {
    asm volatile(AWF_BODY
                 : "+f"(P[0]), "+f"(P[1]), "+f"(P[2]), "+f"(P[3]), "+f"(P[4]), "+f"(P[5]), "+f"(P[6]), "+f"(P[7]),
                   "+f"(P[8]), "+f"(P[9]), "+f"(P[10]), "+f"(P[11]), "+f"(P[12]), "+f"(P[13]), "+f"(P[14]), "+f"(P[15])
                 : "f"(wx[0]), "f"(wx[1]), "f"(wx[2]), "f"(wx[3]), "f"(wx[4]), "f"(wx[5]), "f"(wx[6]), "f"(wx[7]),
                   "r"(k[0]), "r"(k[1]), "r"(k[2]), "r"(k[3]), "r"(k[4]), "r"(k[5]), "r"(k[6]), "r"(k[7]),
                   "r"(n_slots), "r"(pitch), "r"(rp), "r"(odd_first), "l"(obase), "r"(store_mask)
                 : "memory");
}

struct F32Args {
    const float* src;        // [n_planes][H][W * E]
    float* dst;              // [n_planes][Ho][Wo * E]
    const float* map_x;      // [n_planes / map_div][Wo]
    const float* map_y;      // [n_planes / map_div][Ho]
    int H, W, Ho, Wo, E;
    int map_div;             // planes of a CHW image share its maps
    int n_planes, n_strips;
    int strip_el;            // output elements per strip (a multiple of E; <= consumer threads x 4)
    int total_units;         // n_planes * n_strips * Ho
    int stage_bytes, rows;
};

// blockDim.x = consumer threads (a multiple of 32) + 32 producer threads.
// Shared memory: [kStages source arenas][kStages chunk tables][mbarriers full[s], sfree[s]].
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) remap_f32_stream_kernel(const F32Args a) {
    const int tid = threadIdx.x;
    const int R = a.rows;
    const int E = a.E;
    const int tab_off0 = kStages * a.stage_bytes;
    const int bar_off0 = tab_off0 + kStages * kTabBytes;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t full_s = smem_s + (uint32_t)bar_off0;
    const uint32_t sfree_s = full_s + 8u * kStages;
    const int n_cons_warps = ((int)blockDim.x >> 5) - 1;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_s + 8u * s, 1);
            mbar_init(sfree_s + 8u * s, n_cons_warps);
        }
        mbar_init_fence();
    }
    __syncthreads();

    const int u0 = (int)(((int64_t)a.total_units * blockIdx.x) / gridDim.x);
    const int u1 = (int)(((int64_t)a.total_units * (blockIdx.x + 1)) / gridDim.x);
    const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    const int H = a.H, W = a.W, Ho = a.Ho, Wo = a.Wo;
    const int64_t row_pitch = (int64_t)W * E * 4;             // bytes
    const int64_t out_pitch = (int64_t)Wo * E * 4;
    const int units_per_plane = a.n_strips * Ho;

    if (warp_idx == n_cons_warps) {
        // =========================== producer warp =========================================
        int st = 0;
        uint32_t ph = 0;
        int u = u0;
        while (u < u1) {
            // ---- segment: output rows [rt, y_end) of one (plane, strip) ----------------
            const int plane = u / units_per_plane;
            const int local = u - plane * units_per_plane;
            const int strip = local / Ho, rt = local - strip * Ho;
            const int seg_rows = min(Ho - rt, u1 - u);
            const int y_end = rt + seg_rows;
            const int mrow = plane / a.map_div;
            const int el_first = strip * a.strip_el;
            const int n_el = min(a.strip_el, Wo * E - el_first);
            const int x_first = el_first / E, ncols = (n_el + E - 1) / E;
            const uint8_t* simg = reinterpret_cast<const uint8_t*>(a.src + (int64_t)plane * H * W * E);
            const uintptr_t dplane = reinterpret_cast<uintptr_t>(a.dst + (int64_t)plane * Ho * Wo * E);
            const float* my = a.map_y + (int64_t)mrow * Ho;
            const float* mx = a.map_x + (int64_t)mrow * Wo;
            int c_lo, row_bytes, slot_pitch;
            const bool one_copy = a.n_strips == 1 && (int64_t)(a.stage_bytes - 64) / row_pitch >= 3;
            if (one_copy) {
                c_lo = 0;
                row_bytes = (int)row_pitch;
                slot_pitch = row_bytes;
            } else {
                int lo = 0x7fffffff, hi = -1;
                for (int x = lane; x < ncols; x += 32) {
                    const int sx = quantise_coord(__ldg(mx + x_first + x));
                    lo = min(lo, clampi(sx >> 5, 0, W - 1));
                    hi = max(hi, clampi((sx >> 5) + 1, 0, W - 1));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                c_lo = lo;
                row_bytes = (hi - lo + 1) * E * 4;
                slot_pitch = ((row_bytes + 31 + 15) & ~15) + (int)(row_pitch & 15);
            }
            const int max_slots = min((a.stage_bytes - 64) / slot_pitch, 2 * R);
            const uint8_t* scol = simg + (int64_t)c_lo * E * 4;
            uint32_t seg_flags = kFlagNewStrip;
            int carry_row = kNoCarry;    // the consumers hold source rows carry_row - 1 and carry_row
            int y_cur = rt;
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
                const int y = y_cur + lane;
                const bool live = y < y_end && lane < R;
                const int wsel = y - y_win;
                const int sy_a = __shfl_sync(0xffffffffu, sy_cur, wsel & 31);
                const int sy_b = __shfl_sync(0xffffffffu, sy_nxt, wsel & 31);
                int ya = 0x3fffffff, yb = 0x3fffffff, ay = 0;
                if (live) {
                    const int sy = wsel < 32 ? sy_a : sy_b;
                    ay = sy & 31;
                    ya = clampi(sy >> 5, 0, H - 1);
                    yb = clampi((sy >> 5) + 1, 0, H - 1);
                }
                // contiguous staging of source rows r_lo .. yb(last); rows held by the consumers are skipped
                const int prev_ya = __shfl_up_sync(0xffffffffu, ya, 1);
                const int prev_yb = __shfl_up_sync(0xffffffffu, yb, 1);
                const int r0 = __shfl_sync(0xffffffffu, ya, 0);
                const int yb0 = __shfl_sync(0xffffffffu, yb, 0);
                // A chunk either extends the pair the consumers hold or starts afresh.  A fresh start stages at least
                // TWO contiguous rows, so that "the consumers hold carry_row - 1 and carry_row" stays true: a first row
                // on the bottom border taps row H - 1 twice (ya == yb), and staging that row alone would leave an
                // older row in the other register for a later chunk that starts at H - 2 to pick up.
                const int r_lo = (r0 == carry_row - 1 || r0 == carry_row) ? carry_row + 1
                                                                          : ((r0 == yb0 && r0 > 0) ? r0 - 1 : r0);
                const int need = yb + 1 - r_lo;
                // (rows are emitted in table order as the slots advance: both taps must be non-decreasing)
                const unsigned bad = __ballot_sync(0xffffffffu, !live || (lane > 0 && (ya < prev_ya || yb < prev_yb)) ||
                                                                    need > max_slots);
                const int n_rows = max_slots >= 2 ? (bad ? (__ffs(bad) - 1) : 32) : 0;
                const int yb_last = __shfl_sync(0xffffffffu, yb, max(n_rows - 1, 0));
                const int n_slots = n_rows > 0 ? max(yb_last + 1 - r_lo, 0) : 0;
                const uint8_t* p0 = scol + (int64_t)r_lo * row_pitch;
                const int phase0 = (int)(reinterpret_cast<uintptr_t>(p0) & 15);
                const uint8_t* p = p0 + (int64_t)lane * row_pitch;
                const int off = (int)(reinterpret_cast<uintptr_t>(p) & 15);
                const uint32_t bytes = lane < n_slots ? (uint32_t)((off + row_bytes + 15) & ~15) : 0u;
                const uint32_t tx = one_copy ? (uint32_t)((phase0 + n_slots * slot_pitch + 15) & ~15)
                                             : __reduce_add_sync(0xffffffffu, bytes);
                mbar_wait_relaxed(sfree_s + 8u * st, ph ^ 1u);       // (the producer runs ahead: it sleeps between polls)
                if (lane < n_rows) {
                    const float fy = fmul_nofma((float)ay, 1.0f / 32.0f);
                    st128(tab + kTabRows + 16 * lane,
                          make_uint4(__float_as_uint(fadd_nofma(1.0f, -fy)), __float_as_uint(fy),
                                     (uint32_t)((int64_t)y * out_pitch) | (uint32_t)((ya & 1) | ((yb & 1) << 1)),
                                     yb < r_lo ? 0xffffffffu : (uint32_t)(yb - r_lo)));
                } else if (lane == n_rows) {
                    st128(tab + kTabRows + 16 * lane, make_uint4(0u, 0u, 0u, kRowSentinel));
                }
                if (lane == 0) {
                    const uint32_t fl = seg_flags | ((r_lo & 1) ? kFlagOddFirst : 0u);
                    st128(tab, make_uint4((uint32_t)n_rows, (uint32_t)n_slots | (fl << 16), (uint32_t)slot_pitch,
                                          (uint32_t)phase0));
                    st128(tab + 16, make_uint4((uint32_t)plane, (uint32_t)el_first, (uint32_t)y_cur, (uint32_t)(c_lo * E)));
                    st128(tab + kTabOut, make_uint4((uint32_t)dplane, (uint32_t)((uint64_t)dplane >> 32), (uint32_t)n_el, 0u));
                    if (seg_flags & kFlagNewStrip) {
                        const uintptr_t mxa = reinterpret_cast<uintptr_t>(mx);
                        st128(tab + kTabStrip, make_uint4((uint32_t)mxa, (uint32_t)((uint64_t)mxa >> 32), (uint32_t)W, 0u));
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
                // after this chunk the consumers hold rows (latest - 1, latest)
                carry_row = n_rows > 0 ? (n_slots > 0 ? yb_last : carry_row) : kNoCarry;
                y_cur += max(n_rows, 1);
                if (++st == kStages) { st = 0; ph ^= 1u; }
            }
            u += seg_rows;
        }
        mbar_wait_relaxed(sfree_s + 8u * st, ph ^ 1u);
        if (lane == 0) {
            st128(tab_off0 + st * kTabBytes, make_uint4(0xffffffffu, 0u, 0u, 0u));
            mbar_arrive(full_s + 8u * st);
        }
        return;
    }

    // =============================== consumer warps ==============================================
    float P[16], wx[8];
    int so[8];                        // byte offsets of the taps inside a staged row span
#pragma unroll
    for (int k = 0; k < 16; ++k) P[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { wx[k] = 0.f; so[k] = 0; }
    uint32_t store_mask = 0u;
    bool warp_live = false;
    int st = 0;
    uint32_t sph = 0u;
    for (;; st = st + 1 == kStages ? 0 : st + 1, sph ^= st == 0 ? 1u : 0u) {
        const int tab = tab_off0 + st * kTabBytes;
        mbar_wait(full_s + 8u * st, sph);
        const uint4 h0 = ld128(tab);
        const int n_rows = (int)h0.x;
        if (n_rows < 0) break;
        const uint32_t flags = h0.y >> 16;
        const uint4 h1 = ld128(tab + 16);
        const uint4 ho = ld128(tab + kTabOut);
        const int el_first = (int)h1.y, n_el = (int)ho.z;
        if (flags & kFlagNewStrip) {
            const uint4 hx = ld128(tab + kTabStrip);
            const float* mx = reinterpret_cast<const float*>(((uint64_t)hx.y << 32) | hx.x);
            const int Wsrc = (int)hx.z;
            store_mask = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = warp_idx * 128 + 32 * j + lane;
                so[2 * j] = so[2 * j + 1] = 0;
                if (e < n_el) {
                    store_mask |= 1u << j;
                    const int ge = el_first + e;
                    const int x = ge / E, c = ge - x * E;
                    const int sx = quantise_coord(__ldg(mx + x));
                    const float fx = fmul_nofma((float)(sx & 31), 1.0f / 32.0f);
                    wx[2 * j] = fadd_nofma(1.0f, -fx);
                    wx[2 * j + 1] = fx;
                    so[2 * j] = (clampi(sx >> 5, 0, Wsrc - 1) * E + c - (int)h1.w) * 4;
                    so[2 * j + 1] = (clampi((sx >> 5) + 1, 0, Wsrc - 1) * E + c - (int)h1.w) * 4;
                }
            }
            warp_live = __any_sync(0xffffffffu, store_mask != 0u);
        }
        if (n_rows == 0) {
            // ---- direct path: one output row gathered from global memory ----------------------
            const int plane = (int)h1.x, y0 = (int)h1.z;
            const int mrow = plane / a.map_div;
            const int sy = quantise_coord(__ldg(a.map_y + (int64_t)mrow * Ho + y0));
            const int ya = clampi(sy >> 5, 0, H - 1), yb = clampi((sy >> 5) + 1, 0, H - 1);
            const float* sp = a.src + (int64_t)plane * H * W * E;
            float* dp = a.dst + (int64_t)plane * Ho * Wo * E + (int64_t)y0 * Wo * E;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = warp_idx * 128 + 32 * j + lane;
                if (e < n_el) {
                    const int ge = el_first + e;
                    const int x = ge / E, c = ge - x * E;
                    const int sx = quantise_coord(__ldg(a.map_x + (int64_t)mrow * Wo + x));
                    const int xa = clampi(sx >> 5, 0, W - 1), xc = clampi((sx >> 5) + 1, 0, W - 1);
                    const BilinearWeightsF32 w = bilinear_weights_f32(sx & 31, sy & 31);
                    dp[ge] = bilinear_f32(__ldg(sp + ((int64_t)ya * W + xa) * E + c), __ldg(sp + ((int64_t)ya * W + xc) * E + c),
                                          __ldg(sp + ((int64_t)yb * W + xa) * E + c), __ldg(sp + ((int64_t)yb * W + xc) * E + c), w);
                }
            }
        } else if (warp_live) {
            const int n_slots = (int)(h0.y & 0xffffu);
            const uint32_t base = smem_s + (uint32_t)(st * a.stage_bytes) + h0.w;
            uint32_t k[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) k[i] = base + (uint32_t)so[i];
            const uint64_t obase = (((uint64_t)ho.y << 32) | ho.x) + (uint64_t)(el_first + warp_idx * 128 + lane) * 4u;
            sweep_f32(P, wx, k, n_slots, h0.z, smem_s + (uint32_t)(tab + kTabRows), (flags & kFlagOddFirst) ? 1u : 0u,
                      obase, store_mask);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sfree_s + 8u * st);
    }
}

struct GeometryF { int warps, ctas, max_el; };
constexpr GeometryF kGeoF[3] = {{4, 4, 512}, {8, 2, 1024}, {12, 1, 1536}};

int env_int_f(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int G>
int launch_geo_f(F32Args& a, cudaStream_t st) {
    constexpr int kThreads = (kGeoF[G].warps + 1) * 32;
    auto kern = remap_f32_stream_kernel<kThreads, kGeoF[G].ctas>;
    // a slot at unit scale: the strip's elements + one pixel, + alignment head and tail
    const int unit_pitch = (((a.strip_el + a.E) * 4 + 31 + 15) & ~15) + 16;
    const int budget = (227 * 1024) / kGeoF[G].ctas - 1024 - kStages * kTabBytes - 64 - 256;
    int R = (budget / kStages - 2 * unit_pitch - 64 - 128) / unit_pitch;
    R = R > kMaxRows ? kMaxRows : R;
    const int forced = env_int_f("ATTWARP_F32_ROWS", 0);
    if (forced >= 2 && forced <= R) R = forced;
    if (R < 2) return ATTWARP_ERR_UNSUPPORTED;
    a.rows = R;
    a.stage_bytes = ((R + 2) * unit_pitch + 64 + 127) & ~127;
    const size_t smem_bytes = (size_t)kStages * (a.stage_bytes + kTabBytes) + 4 * kStages * sizeof(uint64_t) + 16;
    struct Cfg { size_t smem; int dev, occ; };
    static thread_local Cfg c = {0, -1, 0};
    int dev = 0;
    AW_CUDA(cudaGetDevice(&dev));
    if (c.smem != smem_bytes || c.dev != dev) {
        AW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        int o = 0;
        AW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kThreads, smem_bytes));
        c = Cfg{smem_bytes, dev, o};
    }
    if (c.occ < 1) return ATTWARP_ERR_UNSUPPORTED;
    const int64_t cap = (int64_t)sm_count() * c.occ;
    const int grid = (int)(a.total_units < cap ? a.total_units : cap);
    kern<<<grid, kThreads, smem_bytes, st>>>(a);
    return check_launch("remap_f32_stream_kernel");
}

}  // namespace

// ATTWARP_REMAP_F32=rows keeps the round-1 kernel (A/B comparisons).
bool remap_f32_stream_enabled() {
    static const bool v = [] {
        const char* e = getenv("ATTWARP_REMAP_F32");
        return !(e != nullptr && strcmp(e, "rows") == 0);
    }();
    return v;
}

// float32 images: n_planes dense planes [H][W * E]; returns ATTWARP_ERR_UNSUPPORTED for shapes it does not take
// (the caller falls back to remap_f32_rows_kernel)
int launch_remap_f32_stream(const float* src, float* dst, int n_planes, int E, int H, int W, int Ho, int Wo,
                            const float* map_x, const float* map_y, int map_div, cudaStream_t st) {
    if ((int64_t)Ho * Wo * E * 4 >= (1ll << 30) || (int64_t)H * W * E * 4 >= (1ll << 31)) return ATTWARP_ERR_UNSUPPORTED;
    const int64_t n_e = (int64_t)Wo * E;
    F32Args a{};
    int g = env_int_f("ATTWARP_F32_GEO", -1);
    if (g < 0 || g > 2) g = n_e <= kGeoF[0].max_el ? 0 : (n_e <= kGeoF[1].max_el ? 1 : 2);
    // strips: as few, equal strips as possible, multiples of E elements (and of 4 elements when that is free)
    const int max_px = kGeoF[g].max_el / E;
    if (max_px < 1) return ATTWARP_ERR_UNSUPPORTED;
    a.n_strips = (Wo + max_px - 1) / max_px;
    const int strip_px = (Wo + a.n_strips - 1) / a.n_strips;
    a.strip_el = strip_px * E;
    a.n_strips = (Wo + strip_px - 1) / strip_px;
    a.src = src; a.dst = dst; a.map_x = map_x; a.map_y = map_y;
    a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.E = E; a.map_div = map_div; a.n_planes = n_planes;
    const int64_t total = (int64_t)n_planes * a.n_strips * Ho;
    if (total > 0x7fffffff) return ATTWARP_ERR_UNSUPPORTED;
    a.total_units = (int)total;
    switch (g) {
        case 0: return launch_geo_f<0>(a, st);
        case 1: return launch_geo_f<1>(a, st);
        default: return launch_geo_f<2>(a, st);
    }
}

}  // namespace aw
