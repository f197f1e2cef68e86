// video_fused_tma.cu -- the B200 fast path of the fused kernel: motion compensation + 8x8 IDCT +
// residual add / intra store (predictMacroblock/copyMacroblock video.go:608-637, video_noasm.go:28-80;
// idct video.go:801-928; copy/add*ToDest video.go:943-1002), with every tile moved by TMA.
//
// v1 (video_kernels.cu) spent 2/3 of its instructions on address arithmetic for staging and on
// re-aligning unaligned windows (profiles/r1_v1_fused_summary.md).  Here:
//   * coefficients: one cp.async.bulk.tensor per 32 blocks, 128-byte swizzle, so that one thread per
//     block reads its eight 16-byte rows bank-conflict free;
//   * reference windows: two cp.async.bulk.tensor per predicted macroblock (32x20 luma: 17 rows of window + a row
//     phase that spreads macroblocks over the shared-memory banks; one rank-4 box
//     for 32x9 Cb + 32x9 Cr -- the TMA unit serves about one box per 46 cycles per SM however small, and that
//     rate, not bytes or instructions, bounded the three-box version).  The TMA unit wants the innermost start coordinate on a 16-byte boundary (measured:
//     tools/tma_probe.cu, an unaligned x raises "illegal instruction"), so the box starts at x & ~15
//     and pixel (0,0) sits at byte x & 15 of every staged row -- one offset per macroblock instead of
//     per-row address arithmetic.  Rows past a plane continue into the next plane and columns past the
//     right edge continue on the next row, exactly like the reference's linear indexing
//     (video_noasm.go:49-50), because the tensor map describes the whole frame buffer as overlapping
//     rows (common.cuh, SlabMaps);
//   * one mbarrier per CTA collects all of it (expect_tx = sum of box bytes).
// The arithmetic is the reference's, bit for bit.
#include <cuda.h>

#include <climits>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>


#include "common.cuh"

// One interpolation path for the four half-pel modes (block_load): 1 = the generic path, 0 = one path per mode.
#ifndef MPEGB200_GENERIC_INTERP
#define MPEGB200_GENERIC_INTERP 1
#endif
#ifndef MPEGB200_PAIRED
#define MPEGB200_PAIRED 1
#endif

namespace mpegb200 {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#ifdef MPEGB200_EXPERIMENTS
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// 16 bytes of a plan, straight from global memory.  The plans are written by the pre-pass grid that this kernel follows under
// programmatic dependent launch: they may be read only after griddepcontrol.wait.  A volatile asm load keeps its place behind that
// (volatile asm) instruction; an ordinary load of `const __restrict__` memory, or __ldg, is an invariant load (LDG.E.CONSTANT) that
// the compiler is free to hoist to the top of the kernel -- above the wait -- where it reads whatever the previous launch left in the
// plan buffer (a wrong byte count for the mbarrier: the CTA never wakes up).  That happened once the measurement branches that used
// to sit between the wait and these loads were compiled out.
__device__ __forceinline__ uint4 ld_plan16(const void* p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// the same boxes, fetched into L2 only (for a group that a later CTA will decode)
__device__ __forceinline__ void tma_prefetch_2d(const void* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const void* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const void* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

#if !MPEGB200_PAIRED
// videoPremultiplierMatrix, video.go:1077-1086
__device__ __forceinline__ constexpr int premult(int i) {
    constexpr int t[64] = {32, 44, 42, 38, 32, 25, 17, 9,  44, 62, 58, 52, 44, 35, 24, 12, 42, 58, 55, 49, 42, 33,
                           23, 12, 38, 52, 49, 44, 38, 30, 20, 10, 32, 44, 42, 38, 32, 25, 17, 9,  25, 35, 33, 30,
                           25, 20, 14, 7,  17, 24, 23, 20, 17, 14, 9,  5,  9,  12, 12, 10, 9,  7,  5,  2};
    return t[i];
}

// one 8-point pass, video.go:870-895 (int32 suffices: SURVEY Q10, tests/test_host_logic.py)
__device__ __forceinline__ void idct_pass8(int& s0, int& s1, int& s2, int& s3, int& s4, int& s5, int& s6, int& s7) {
    const int b1 = s4;
    const int b3 = s2 + s6;
    const int b4 = s5 - s3;
    const int tmp1 = s1 + s7;
    const int tmp2 = s3 + s5;
    const int b6 = s1 - s7;
    const int b7 = tmp1 + tmp2;
    const int m0 = s0;
    const int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    // -(u >> 8) == (-u + 255) >> 8 for the arithmetic shift, so the subtraction of video.go:879 becomes shift-and-add (one LEA.HI.SX32)
    const int x0 = x4 + (((tmp2 - tmp1) * 362 + 127) >> 8);
    const int x1 = m0 - b1;
    const int x2 = (((s2 - s6) * 362 + 128) >> 8) - b3;
    const int x3 = m0 + b1;
    const int y3 = x1 + x2;
    const int y4 = x3 + b3;
    const int y5 = x1 - x2;
    const int y6 = x3 - b3;
    const int ny7 = x0 + ((b4 * 473 + b6 * 196 + 128) >> 8);   // -y7 of video.go:887
    s0 = b7 + y4;
    s1 = x4 + y3;
    s2 = y5 - x0;
    s3 = y6 + ny7;
    s4 = y6 - ny7;
    s5 = x0 + y5;
    s6 = y3 - x4;
    s7 = y4 - b7;
}

#endif  // !MPEGB200_PAIRED

// premultiplier pairs for dp2a, in constant memory so that they are instruction operands (c[bank][offset]) instead of
// one uniform-register move each: [pair (0,4) (1,7) (3,5) (2,6)][column]; sum = pa | pb << 8, difference = pa | -pb << 8
// (pair (3,5): -pa | pb << 8, video.go:872 b4 = s5 - s3)
__constant__ int kPairSum[32] = {0x2020, 0x2c2c, 0x2a2a, 0x2626, 0x2020, 0x1919, 0x1111, 0x0909, 0x092c, 0x0c3e, 0x0c3a, 0x0a34, 0x092c, 0x0723, 0x0518, 0x020c, 0x1926, 0x2334, 0x2131, 0x1e2c, 0x1926, 0x141e, 0x0e14, 0x070a, 0x112a, 0x183a, 0x1737, 0x1431, 0x112a, 0x0e21, 0x0917, 0x050c};
__constant__ int kPairDif[32] = {0xe020, 0xd42c, 0xd62a, 0xda26, 0xe020, 0xe719, 0xef11, 0xf709, 0xf72c, 0xf43e, 0xf43a, 0xf634, 0xf72c, 0xf923, 0xfb18, 0xfe0c, 0x19da, 0x23cc, 0x21cf, 0x1ed4, 0x19da, 0x14e2, 0x0eec, 0x07f6, 0xef2a, 0xe83a, 0xe937, 0xec31, 0xef2a, 0xf221, 0xf717, 0xfb0c};

// column pass whose first additions were already made by the coefficient load (block_load, MPEGB200_PAIRED):
// in: s0 = x3 (s0+s4), s1 = tmp1, s2 = b3, s3 = tmp2, s4 = x1 (s0-s4), s5 = b4, s6 = s2-s6, s7 = b6
__device__ __forceinline__ void idct_pass8_pre(int& s0, int& s1, int& s2, int& s3, int& s4, int& s5, int& s6, int& s7) {
    const int x3 = s0, tmp1 = s1, b3 = s2, tmp2 = s3, x1 = s4, b4 = s5, d26 = s6, b6 = s7;
    const int b7 = tmp1 + tmp2;
    const int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    const int x0 = x4 + (((tmp2 - tmp1) * 362 + 127) >> 8);
    const int x2 = ((d26 * 362 + 128) >> 8) - b3;
    const int y3 = x1 + x2;
    const int y4 = x3 + b3;
    const int y5 = x1 - x2;
    const int y6 = x3 - b3;
    const int ny7 = x0 + ((b4 * 473 + b6 * 196 + 128) >> 8);
    s0 = b7 + y4;
    s1 = x4 + y3;
    s2 = y5 - x0;
    s3 = y6 + ny7;
    s4 = y6 - ny7;
    s5 = x0 + y5;
    s6 = y3 - x4;
    s7 = y4 - b7;
}

// row pass, video.go:899-926; the caller has put the +128 on the DC term and passes k[i] = prediction << 8, so
// o[i] = 256 * (result before the final >> 8) + low bits; the shift happens in the saturating pack
__device__ __forceinline__ void idct_row8(const int* s, const int* k, int* o) {
    const int b1 = s[4];
    const int b3 = s[2] + s[6];
    const int b4 = s[5] - s[3];
    const int tmp1 = s[1] + s[7];
    const int tmp2 = s[3] + s[5];
    const int b6 = s[1] - s[7];
    const int b7 = tmp1 + tmp2;
    const int m0 = s[0];
    const int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    const int x0 = x4 + (((tmp2 - tmp1) * 362 + 127) >> 8);
    const int x1 = m0 - b1;
    const int x2 = (((s[2] - s[6]) * 362 + 128) >> 8) - b3;
    const int x3 = m0 + b1;
    const int y3 = x1 + x2;
    const int y4 = x3 + b3;
    const int y5 = x1 - x2;
    const int y6 = x3 - b3;
    const int ny7 = x0 + ((b4 * 473 + b6 * 196 + 128) >> 8);
    o[0] = b7 + y4 + k[0];
    o[1] = x4 + y3 + k[1];
    o[2] = y5 - x0 + k[2];
    o[3] = y6 + ny7 + k[3];
    o[4] = y6 - ny7 + k[4];
    o[5] = x0 + y5 + k[5];
    o[6] = y3 - x4 + k[6];
    o[7] = y4 - b7 + k[7];
}

#if !MPEGB200_GENERIC_INTERP
__device__ __forceinline__ uint32_t avg2(uint32_t a, uint32_t b) { return __vavgu4(a, b); }  // (a+b+1)>>1 per byte
#endif

constexpr int kG = 16;                 // macroblock records per CTA
constexpr int kNT = 6 * kG;            // one thread per output block (8x8)
constexpr int kCoefBox = 32;           // blocks per coefficient TMA box (32 x 128 B = 4 KiB)

// Two ways to stage the reference windows of a group (the plan pre-pass picks one per group):
//  * strip mode: when every predicted macroblock of the group reads the same reference buffer and all windows fit a
//    304x48 luma / 160x24 chroma rectangle (16 neighbouring macroblocks with vectors within +-16 pixels do), the
//    whole rectangle comes in with ONE luma box and ONE rank-4 chroma box.  The TMA unit serves about one box per
//    46 cycles per SM however small (profiles/r1_final_video_summary.md); two boxes per macroblock made the box
//    rate, not bytes or instructions, the limit of the predicted steps.  Overlapping parts of neighbouring strips
//    are L2 hits, DRAM traffic stays the algorithmic one.
//  * box mode (any vectors, any record order): two boxes per predicted macroblock (32x20 luma, 32x9 Cb + 32x9 Cr),
//    staged with a per-macroblock row phase (macroblock j's luma box starts (j & 3) rows above its window) so that
//    the same window row of different macroblocks lands in different shared-memory banks.
// Either way a block thread sees: byte offset of its pixel (0,0) in the window area + a row pitch.
constexpr int kWinY = (32 * kLumaBoxRows + 127) / 128 * 128;          // luma box 32x20 = 640 B
constexpr int kWinC = 32 * kChromaBoxRows;                            // chroma box: Cb 32x9 then Cr 32x9 = 576 B in 640
constexpr int kWinBytes = kWinY + (2 * kWinC + 127) / 128 * 128;      // 1280 per macroblock, both boxes 128-byte aligned
constexpr int kWinTx = 32 * kLumaBoxRows + 2 * kWinC;                 // bytes the two boxes deliver
constexpr uint32_t kPhaseY = (uint32_t)(kLumaBoxRows - 17), kPhaseC = (uint32_t)(kChromaBoxRows - 9);  // masks: 3 / 0
constexpr int kStripLBytes = kStripLW * kStripLH;                     // 14592
constexpr int kStripCBytes = kStripCW * kStripCH;                     // 3840 per plane
constexpr int kStripTx = kStripLBytes + 2 * kStripCBytes;             // 22272
constexpr int kWinArea = (kG * kWinBytes > kStripTx ? kG * kWinBytes : kStripTx + 127) / 128 * 128;
constexpr int kDefaultPrefetchDist = 444; // groups ahead to prefetch into L2: three per SM (MPEGB200_PREFETCH_DIST overrides; measured 0: 0.333 ms, 444: 0.327, 888: 0.327, 1776: 0.332, 3552: 0.402)
constexpr int kStripMinPred = 3;       // fewer predicted macroblocks than this: their own boxes are cheaper

// ------------------------------------------------------------------------------------------------
// Group plan: everything a CTA needs to know about its kG records, computed once by a pre-pass
// (plan_kernel) so that the decode, the record -> picture -> stream pointer chase and the sorting of
// the output blocks are neither repeated nor serialised inside the arithmetic kernel.
// ------------------------------------------------------------------------------------------------
struct PlanMb {               // 32 bytes; the pre-pass builds one per record, the plan stores it field by field
    uint8_t* dst_y;           // destination of the macroblock's luma (row 0, col 0 of the MB)
    uint32_t dst_c_off;       // Cb destination = dst_y + dst_c_off
    uint32_t dst_cr_off;      // Cr destination = Cb destination + dst_cr_off
    uint16_t luma_w;
    uint16_t mcw;             // byte 0 luma, byte 1 chroma: mode << 4 | 0x80 if predicted;
                              // mode bit 0 = horizontal half-pel, bit 1 = vertical half-pel
    uint16_t woff_y, woff_c;  // byte offset of luma / Cb pixel (0,0) of the staged window in the window area
    uint32_t pad[2];
};
static_assert(sizeof(PlanMb) == 32, "PlanMb size");

struct PlanBox {              // 16 bytes, box mode only: where the two boxes of a macroblock come from
    int16_t lx, ly, cx, cy;   // box origins (x folded into [0, pitch) and aligned down to 16, y minus the row phase)
    uint16_t z;               // 3 * slot + reference buffer
    uint16_t slab;            // tensor-map set of the stream's slab
    uint16_t pred;            // 1 if the macroblock has windows
    uint16_t pad;
};
static_assert(sizeof(PlanBox) == 16, "PlanBox size");

// A group plan in global memory: head (what the fetch needs), body (what the block threads need), boxes (box mode only).
struct PlanHead {             // 48 bytes
    uint32_t n_box, n_pred, block0, tx_bytes;   // coefficient boxes, predicted macroblocks, first block, bytes the boxes deliver
    uint16_t strip, slab, z;                    // strip mode: one luma + one chroma box at (sx, sy) / (scx, scy)
    uint16_t pitch_y, pitch_c, cr_win;          // window row pitches; Cb -> Cr window distance
    int16_t sx, sy, scx, scy;                   // strip origins in bytes / rows (x multiples of 16)
    uint32_t pad0[3];
};
static_assert(sizeof(PlanHead) == 48, "PlanHead size");

struct PlanBody {             // 576 bytes
    uint16_t map[kNT];        // output-block list, sorted by (coded?, interpolation mode)
    // The per-macroblock fields as arrays (PlanMb, field by field): the 32 block threads of a warp read the fields of up to
    // 16 different macroblocks, and an array of 16 entries of 8 bytes or less never puts two of them on one bank.
    uint8_t* dst_y[kG];
    uint32_t dst_c_off[kG], dst_cr_off[kG];
    uint16_t luma_w[kG], mcw[kG], woff_y[kG], woff_c[kG];
};
static_assert(sizeof(PlanBody) == 576, "PlanBody size");

struct GroupPlan {            // 1024 bytes = 64 x 16
    PlanHead h;
    PlanBody b;
    uint8_t pad1[768 - 48 - 576];
    PlanBox box[kG];
};
static_assert(sizeof(GroupPlan) == 1024, "GroupPlan size");
static_assert(offsetof(GroupPlan, b) == 48 && offsetof(GroupPlan, box) == 768, "GroupPlan layout");

// Output-block list entry: [3:0] macroblock in group, [6:4] block 0..5, [7] coded, [14:8] coefficient slot.
constexpr uint32_t kNoBlock = 0xFFFFu;

constexpr int kPlanGroupsPerCta = 8;   // 16 lanes per group

// max per 16-bit lane over the 16 lanes of a group (xor butterflies of 8, 4, 2, 1 never leave the half-warp, so the
// whole warp runs them together: no divergence between the two groups of a warp)
__device__ __forceinline__ uint32_t group_max_s16x2(uint32_t v) {
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) v = __vmaxs2(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
// (hi, lo) = (v, -v) as 16-bit lanes: one max reduction yields max(v) and -min(v).  Lanes that do not take part pass the identity.
__device__ __forceinline__ uint32_t span_s16x2(int v, bool take) {
    return take ? (((uint32_t)v << 16) | ((uint32_t)(-v) & 0xffffu)) : 0x80008000u;
}
__device__ __forceinline__ int span_max(uint32_t r) { return (int)r >> 16; }
__device__ __forceinline__ int span_min(uint32_t r) { return -(int)(int16_t)(r & 0xffffu); }

// The plan pre-pass: 16 lanes = one group, two groups per warp, no CTA-wide barrier (a group never leaves its half-warp):
// bounding boxes by 16-bit-pair max butterflies over the half-warp, the block order from 64 shared 16-bit counters per group
// (two per word, so that a rank is four packed min + add steps instead of eight scalar ones).
__global__ void __launch_bounds__(16 * kPlanGroupsPerCta) plan_kernel(GroupPlan* __restrict__ plans,
                                                                     const StreamInfo* __restrict__ streams,
                                                                     int max_streams,
                                                                     const mpegb200_picture* __restrict__ pics,
                                                                     int n_pics, const mpegb200_mb* __restrict__ mbs,
                                                                     uint32_t n_mb, uint32_t n_blocks, int allow_strip) {
    // Programmatic dependent launch: once every CTA of this grid has started, the CTAs of the arithmetic kernel may take
    // the SM slots that free up and run their prologue; they wait for this grid's completion (griddepcontrol.wait)
    // before they touch a plan.
    asm volatile("griddepcontrol.launch_dependents;");
    __shared__ __align__(16) GroupPlan s_plan[kPlanGroupsPerCta];
    __shared__ __align__(16) uint32_t s_cnt2[kPlanGroupsPerCta][8][4];   // [bin][(coefficient slot & 7) >> 1], counter of an odd slot in the high half
    __shared__ __align__(16) uint4 s_lt[16 * kPlanGroupsPerCta / 32][8]; // per warp: [r] = 16-bit lanes b < r set to 1
    const int tid = threadIdx.x, gl = tid >> 4, lane = tid & 15;
    const int seg0 = tid & 16;                             // first lane of this group inside the warp
    const unsigned seg = 0xffffu << seg0;                  // and all its lanes
    const uint32_t n_groups = (n_mb + kG - 1) / kG;
    const uint32_t group = blockIdx.x * kPlanGroupsPerCta + gl;
    GroupPlan& P = s_plan[gl];
    if (lane < 8) reinterpret_cast<uint4*>(&s_cnt2[gl][0][0])[lane] = make_uint4(0, 0, 0, 0);
    if (lane < 12) reinterpret_cast<uint4*>(P.b.map)[lane] = make_uint4(~0u, ~0u, ~0u, ~0u);   // kNoBlock everywhere
    if ((tid & 31) < 8) {
        const int r = tid & 31;
        s_lt[tid >> 5][r] = make_uint4((r > 0 ? 1u : 0u) | (r > 1 ? 0x10000u : 0u), (r > 2 ? 1u : 0u) | (r > 3 ? 0x10000u : 0u),
                                       (r > 4 ? 1u : 0u) | (r > 5 ? 0x10000u : 0u), (r > 6 ? 1u : 0u) | (r > 7 ? 0x10000u : 0u));
    }
    __syncwarp();

    const uint32_t m0 = group * (uint32_t)kG;
    const bool have = group < n_groups && m0 + lane < n_mb;
    uint32_t out_mask = 0, cbp_r = 0, rel_r = 0, bin_y = 0, bin_c = 0, pos = 0, pos_hi = 0, top = 0;
    uint32_t block0 = 0;
    bool predicted = false;
    int lx = 0, ly = 0, cx = 0, cy = 0, lw = 0, zslab = 0;
    PlanMb c;
    memset(&c, 0, sizeof(c));
    if (group < n_groups) block0 = mbs[m0].coeff_block;
    if (have) {
        const uint4 raw = reinterpret_cast<const uint4*>(mbs)[m0 + lane];
        const uint32_t row = raw.x & 0xffffu, col = raw.x >> 16;
        const int mv_h = (int16_t)(raw.y & 0xffffu), mv_v = (int16_t)(raw.y >> 16);
        const uint32_t flags = raw.z & 0xffu, cbp = (raw.z >> 8) & 0x3fu, pic_i = raw.z >> 16;
        const uint32_t cblock = raw.w;
        const int ncoded = __popc(cbp);
        bool ok = pic_i < (uint32_t)n_pics;
        if (ok) {
            const uint4 praw = reinterpret_cast<const uint4*>(pics)[pic_i];
            const int stream = (int)praw.x;
            const uint32_t dst_b = (praw.y >> 8) & 0xffu, fwd_b = (praw.y >> 16) & 0xffu, bwd_b = praw.y >> 24;
            ok = stream >= 0 && stream < max_streams && dst_b < 3 && fwd_b < 3 && bwd_b < 3;
            if (ok) {
                const StreamInfo si = streams[stream];
                ok = si.open && si.tma_ok && row < si.mb_h && col < si.mb_w;
                const uint32_t rel = cblock - block0;
                if (ncoded) ok = ok && rel <= (uint32_t)kNT && rel + ncoded <= (uint32_t)kNT && cblock + ncoded <= n_blocks;
                if (ok) {
                    lw = si.luma_w;
                    const uint32_t cw = (uint32_t)lw >> 1;
                    uint8_t* dst = si.base + (size_t)dst_b * si.buf_stride;
                    c.dst_y = dst + (size_t)(row << 4) * lw + (col << 4);
                    c.dst_c_off = (uint32_t)((size_t)lw * si.luma_h + (size_t)(row << 3) * cw + (col << 3) -
                                             ((size_t)(row << 4) * lw + (col << 4)));
                    c.dst_cr_off = cw * (uint32_t)(si.luma_h >> 1);
                    c.luma_w = (uint16_t)lw;
                    predicted = (flags & MPEGB200_MB_PREDICT) != 0 && !(flags & MPEGB200_MB_INTRA);
                    uint32_t mode_y = 0, mode_c = 0;
                    if (predicted) {  // window origins, video_noasm.go:29-42 (not yet folded into the row)
                        lx = (int)(col << 4) + (mv_h >> 1);
                        ly = (int)(row << 4) + (mv_v >> 1);
                        const int cmh = mv_h / 2, cmv = mv_v / 2;  // toward zero
                        cx = (int)(col << 3) + (cmh >> 1);
                        cy = (int)(row << 3) + (cmv >> 1);
                        mode_y = (uint32_t)((mv_h & 1) | ((mv_v & 1) << 1));
                        mode_c = (uint32_t)((cmh & 1) | ((cmv & 1) << 1));
                        c.mcw = (uint16_t)(((mode_y << 4) | 0x80) | (((mode_c << 4) | 0x80) << 8));
                        zslab = (int)(((uint32_t)si.slab << 16) | (uint32_t)(si.slot * 3 + ((flags & MPEGB200_MB_REF_BWD) ? bwd_b : fwd_b)));
                    }
                    if (ncoded) top = rel + ncoded;
                    // every 8x8 block whose pixels this record defines goes on the output list, binned by
                    // (coded?, interpolation mode) so that the threads of a warp take the same code path
                    out_mask = predicted ? 0x3fu : cbp;
                    cbp_r = cbp;
                    rel_r = rel;
                    // Inside a bin the blocks are ordered round-robin over (coefficient slot & 7): the 128-byte swizzle of the
                    // coefficient tile makes the eight LDS.128 of a quarter-warp conflict free exactly when its eight
                    // slots differ mod 8 (profiles/r1_final_video_summary.md: 9.1 wavefronts per instruction, ideal 4).
                    // Bin order H, copy, V, HV: a warp that straddles two bins runs both interpolation paths, and
                    // with 96 blocks in four bins the two middle bins are the ones that get split over warps --
                    // so the two cheapest paths (copy, V) sit in the middle.  Intra blocks (no window) count as copy.
#if MPEGB200_GENERIC_INTERP
                    // With ONE interpolation path for all modes only coded / not coded is left to sort by: two bins, and the coded
                    // blocks of the whole group go round-robin over slot & 7 -- the eight LDS.128 of every quarter-warp hit eight
                    // different columns of the coefficient tile across the whole block list, not just inside a mode's bin
                    // (dense step 0.3117 -> 0.2959 ms: worth more than the shorter interpolation itself).
                    bin_y = bin_c = 0;
#else
                    bin_y = (0x3201u >> (4 * mode_y)) & 3u;
                    bin_c = (0x3201u >> (4 * mode_c)) & 3u;
#endif
                }
            }
        }
    }
    // (A shortcut for groups whose 16 records code all six blocks -- block (j, k) in thread 6 j + k, no sorting -- was measured
    // and dropped: 0.2974 ms against 0.2959 ms on the dense step, the sorted order stages the windows with fewer bank conflicts.)
    {
        uint32_t slot_run = rel_r;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (out_mask & (0x20u >> k)) {
                const bool coded = (cbp_r & (0x20u >> k)) != 0;
                const uint32_t bin = (k < 4 ? bin_y : bin_c) + (coded ? 0u : 4u);
                const uint32_t res = coded ? (slot_run & 7u) : ((uint32_t)(lane + k) & 7u);
                const uint32_t sh = (res & 1u) << 4;
                const uint32_t pp = (atomicAdd(&s_cnt2[gl][bin][res >> 1], 1u << sh) >> sh) & 0xffffu;  // < 96
                if (k < 4) pos |= pp << (8 * k); else pos_hi |= pp << (8 * (k - 4));
                if (coded) slot_run++;
            }
        }
    }
    __syncwarp();   // the counters are complete
    // strip or boxes?  Every lane of the group takes the same decision from reductions over the half-warp.  All window
    // origins fit 16 bits (pictures are at most 4095 wide and high, vectors are int16 half-pels).
    const unsigned pmask = __ballot_sync(0xffffffffu, predicted) & seg;
    const uint32_t n_pred = (uint32_t)__popc(pmask);
    const int first = pmask ? __ffs((int)pmask) - 1 : seg0;       // the group's first predicted lane
    const int z0 = __shfl_sync(0xffffffffu, zslab, first);
    const int slw = __shfl_sync(0xffffffffu, lw, first), scw = slw >> 1;   // one reference buffer => one stream => one pitch
    const bool one_ref = (__ballot_sync(0xffffffffu, predicted && zslab != z0) & seg) == 0;
    const uint32_t r_lx = group_max_s16x2(span_s16x2(lx, predicted)), r_ly = group_max_s16x2(span_s16x2(ly, predicted));
    const uint32_t r_cx = group_max_s16x2(span_s16x2(cx, predicted)), r_cy = group_max_s16x2(span_s16x2(cy, predicted));
    const uint32_t n_top = group_max_s16x2(top) & 0xffffu;        // top <= 96
    const int min_lx = span_min(r_lx), max_lx = span_max(r_lx), min_ly = span_min(r_ly), max_ly = span_max(r_ly);
    const int min_cx = span_min(r_cx), max_cx = span_max(r_cx), min_cy = span_min(r_cy), max_cy = span_max(r_cy);
    bool strip = false;
    int X0 = 0, Y0 = 0, CX0 = 0, CY0 = 0;
    if (allow_strip && n_pred >= (uint32_t)kStripMinPred && one_ref) {
        X0 = min_lx & ~15;   // two's complement: rounds toward minus infinity
        Y0 = min_ly;
        CX0 = min_cx & ~15;
        CY0 = min_cy;
        // The rectangle must hold every window, start inside the row (x >= 0: a window left of column 0 belongs to the
        // end of the row above in the reference's linear addressing -- box mode folds it, a strip cannot) and every
        // needed byte must lie inside the tensor's rows (luma_w + 32 wide; what the box reads beyond is zero fill
        // that no window uses).
        strip = X0 >= 0 && max_lx + 17 - X0 <= kStripLW && max_lx + 17 <= slw + 32 &&
                max_ly + 17 - Y0 <= kStripLH &&
                CX0 >= 0 && max_cx + 9 - CX0 <= kStripCW && max_cx + 9 <= scw + 32 &&
                max_cy + 9 - CY0 <= kStripCH;
    }
    {
        PlanBox b;
        memset(&b, 0, sizeof(b));
        if (predicted) {
            if (strip) {
                c.woff_y = (uint16_t)((ly - Y0) * kStripLW + (lx - X0));
                c.woff_c = (uint16_t)(kStripLBytes + (cy - CY0) * kStripCW + (cx - CX0));
            } else {
                // The reference indexes linearly (si = y*stride + x, video_noasm.go:31,39), so a window
                // starting left of column 0 really starts near the end of the row above.  Fold x into
                // [0, pitch): same bytes, and inside the tensor's (overlapping) rows.
                const int cw = lw >> 1;
                int q = lx / lw, r = lx - q * lw;
                if (r < 0) { r += lw; q--; }
                lx = r;
                ly += q;
                q = cx / cw;
                r = cx - q * cw;
                if (r < 0) { r += cw; q--; }
                cx = r;
                cy += q;
                const int phy = lane & (int)kPhaseY, phc = lane & (int)kPhaseC;
                c.woff_y = (uint16_t)(lane * kWinBytes + phy * 32 + (lx & 15));
                c.woff_c = (uint16_t)(lane * kWinBytes + kWinY + phc * 32 + (cx & 15));
                b.lx = (int16_t)(lx & ~15);  // the TMA unit needs x on a 16-byte boundary
                b.ly = (int16_t)(ly - phy);
                b.cx = (int16_t)(cx & ~15);
                b.cy = (int16_t)(cy - phc);
                b.z = (uint16_t)(zslab & 0xffff);
                b.slab = (uint16_t)((uint32_t)zslab >> 16);
                b.pred = 1;
            }
        }
        P.b.dst_y[lane] = c.dst_y;
        P.b.dst_c_off[lane] = c.dst_c_off;
        P.b.dst_cr_off[lane] = c.dst_cr_off;
        P.b.luma_w[lane] = c.luma_w;
        P.b.mcw[lane] = c.mcw;
        P.b.woff_y[lane] = c.woff_y;
        P.b.woff_c[lane] = c.woff_c;
        P.box[lane] = b;
    }
    {
        // bin totals: lane b (and b + 8) sums bin b, an exclusive scan over each 8 lanes turns them into the bins' first
        // positions; a block fetches its bin's by shuffle
        const uint4 ta = *reinterpret_cast<const uint4*>(&s_cnt2[gl][lane & 7][0]);
        const uint32_t t2 = ta.x + ta.y + ta.z + ta.w;            // two 16-bit sums, each <= 96
        const uint32_t tot = (t2 & 0xffffu) + (t2 >> 16);
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d, 8);
            if ((lane & 7) >= d) incl += up;
        }
        const uint32_t base = incl - tot;
        uint32_t slot = rel_r;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const bool coded = cbp_r & (0x20u >> k);
            const uint32_t bin = (k < 4 ? bin_y : bin_c) + (coded ? 0u : 4u);   // as in the counting loop above
            const uint32_t res = coded ? (slot & 7u) : ((uint32_t)(lane + k) & 7u);
            const uint32_t bsel = __shfl_sync(0xffffffffu, base, seg0 + (int)bin);   // every lane, whether it has block k or not
            if (out_mask & (0x20u >> k)) {
                const uint32_t idx = (k < 4 ? pos >> (8 * k) : pos_hi >> (8 * (k - 4))) & 0xffu;
                // rank of (idx, res) among the bin's entries in (idx, res) order: sum over r of min(count[r], idx + (r < res))
                const uint4 cn = *reinterpret_cast<const uint4*>(&s_cnt2[gl][bin][0]);
                const uint4 lt = s_lt[tid >> 5][res];
                const uint32_t iv = idx * 0x00010001u;
                const uint32_t s2 = __vminu2(cn.x, iv + lt.x) + __vminu2(cn.y, iv + lt.y) + __vminu2(cn.z, iv + lt.z) + __vminu2(cn.w, iv + lt.w);
                const uint32_t pp = (s2 & 0xffffu) + (s2 >> 16);
                P.b.map[bsel + pp] = (uint16_t)((uint32_t)(lane + 16 * k) | (coded ? 0x80u : 0u) | (slot << 8));
                if (coded) slot++;
            }
        }
    }
    if (lane == 0) {
        const uint32_t n_box = (n_top + kCoefBox - 1) / kCoefBox;
        PlanHead h;
        memset(&h, 0, sizeof(h));
        h.n_box = n_box;
        h.n_pred = n_pred;
        h.block0 = block0;
        h.strip = strip ? 1 : 0;
        if (strip) {
            h.tx_bytes = n_box * (kCoefBox * 128) + kStripTx;
            h.slab = (uint16_t)((uint32_t)z0 >> 16);
            h.z = (uint16_t)(z0 & 0xffff);
            h.pitch_y = kStripLW;
            h.pitch_c = kStripCW;
            h.cr_win = kStripCBytes;
            h.sx = (int16_t)X0;
            h.sy = (int16_t)Y0;
            h.scx = (int16_t)CX0;
            h.scy = (int16_t)CY0;
        } else {
            h.tx_bytes = n_box * (kCoefBox * 128) + n_pred * kWinTx;
            h.pitch_y = 32;
            h.pitch_c = 32;
            h.cr_win = kWinC;
        }
        P.h = h;
        *reinterpret_cast<uint4*>(P.pad1) = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (group < n_groups) {  // head + body (640 bytes with padding) out, 16 bytes per lane per step; the boxes only in box mode
        const uint4* src = reinterpret_cast<const uint4*>(&P);
        uint4* dstp = reinterpret_cast<uint4*>(plans + group);
        dstp[lane] = src[lane];
        dstp[lane + 16] = src[lane + 16];
        if (lane < 8) dstp[lane + 32] = src[lane + 32];
        if (!strip) dstp[lane + 48] = src[lane + 48];
    }
}

// One output block in three steps, so that a kernel can put the next group's fetch between the part that reads
// shared memory and the part that only works on registers:
//   block_setup : the block-list entry -> destination, window address, mode (reads the plan)
//   block_load  : interpolate the prediction from the staged window, fetch + premultiply the coefficients
//   block_finish: IDCT, add, saturate, store
struct BlockCtx {
    uint8_t* dst;
    uint32_t pitch;
    uint32_t win;      // byte offset of the block's pixel (0,0) in the window area
    uint32_t wpitch;   // row pitch of the staged window
    uint32_t slot;     // coefficient slot in the tile
    uint32_t mode;     // bit 0 horizontal, bit 1 vertical half-pel
    bool live, pred, coded;
};

__device__ __forceinline__ void block_setup(const PlanHead& H, const PlanBody& P, int t, BlockCtx& B) {
    const uint32_t e = P.map[t];
    B.live = e != kNoBlock;
    const int j = e & 15, k = (e >> 4) & 7;
    B.coded = B.live && (e & 0x80u);
    B.slot = e >> 8;
    const uint32_t mcb = (uint32_t)(P.mcw[j] >> (k < 4 ? 0 : 8)) & 0xffu;
    B.pred = B.live && (mcb & 0x80u);
    B.mode = (mcb >> 4) & 3u;
    const uint32_t lw = P.luma_w[j];
    if (k < 4) {
        B.pitch = lw;
        B.dst = P.dst_y[j] + (size_t)((k >> 1) * 8) * B.pitch + (k & 1) * 8;
        B.wpitch = H.pitch_y;
        B.win = P.woff_y[j] + (uint32_t)((k >> 1) * 8) * B.wpitch + (k & 1) * 8;
    } else {
        B.pitch = lw >> 1;
        B.dst = P.dst_y[j] + P.dst_c_off[j] + (k == 5 ? P.dst_cr_off[j] : 0u);
        B.wpitch = H.pitch_c;
        B.win = P.woff_c[j] + (k == 5 ? (uint32_t)H.cr_win : 0u);
    }
}

// (a + b + c + d + 2) >> 2 per byte without cascaded averages (video_noasm.go:72-77), on sums kept as two 16-bit lanes:
// HSum holds, for one source row, a[x] + a[x+1] for x = 0,2 / 1,3 / 4,6 / 5,7.
struct HSum { uint32_t p02, p13, p46, p57; };
#if !MPEGB200_GENERIC_INTERP
__device__ __forceinline__ HSum hsum_row(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t sh) {
    const uint32_t u0 = __funnelshift_rc(w0, w1, sh), u1 = __funnelshift_rc(w1, w2, sh);   // bytes 0..3, 4..7
    const uint32_t s1 = __funnelshift_rc(w1, w2, sh + 8);                                   // bytes 5..8
    const uint32_t e0 = u0 & 0x00ff00ffu, o0 = __byte_perm(u0, 0, 0x4341);                  // [b0,b2] [b1,b3]
    const uint32_t e1 = u1 & 0x00ff00ffu, o1 = __byte_perm(u1, 0, 0x4341);                  // [b4,b6] [b5,b7]
    const uint32_t m0 = __funnelshift_r(e0, e1, 16);                                        // [b2,b4]
    const uint32_t m1 = __byte_perm(s1, 0, 0x4341);                                         // [b6,b8]
    HSum h;
    h.p02 = e0 + o0;
    h.p13 = o0 + m0;
    h.p46 = e1 + o1;
    h.p57 = o1 + m1;
    return h;
}
#endif
// The same with the right-hand neighbour replaced by the pixel itself where the vector has no horizontal half (mh = 0):
// a[x] + a[x] instead of a[x] + a[x+1].  With the lower row replaced by the row itself where it has no vertical half, the
// four-sample rounding (s + 2) >> 2 gives all four modes of video_noasm.go:44-80 exactly: (4a + 2) >> 2 = a,
// (2a + 2b + 2) >> 2 = (a + b + 1) >> 1, likewise vertically, and (a + b + c + d + 2) >> 2.  One path for every lane of a
// warp instead of up to four one after the other (natural pictures: 18.9 of 32 lanes active per instruction before).
__device__ __forceinline__ uint32_t sel32(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }   // one LOP3
__device__ __forceinline__ HSum hsum_row_sel(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t sh, uint32_t mh) {
    const uint32_t u0 = __funnelshift_rc(w0, w1, sh), u1 = __funnelshift_rc(w1, w2, sh);   // bytes 0..3, 4..7
    const uint32_t s1 = __funnelshift_rc(w1, w2, sh + 8);                                   // bytes 5..8
    const uint32_t e0 = u0 & 0x00ff00ffu, o0 = __byte_perm(u0, 0, 0x4341);                  // [b0,b2] [b1,b3]
    const uint32_t e1 = u1 & 0x00ff00ffu, o1 = __byte_perm(u1, 0, 0x4341);                  // [b4,b6] [b5,b7]
    const uint32_t m0 = __funnelshift_r(e0, e1, 16);                                        // [b2,b4]
    const uint32_t m1 = __byte_perm(s1, 0, 0x4341);                                         // [b6,b8]
    HSum h;
    h.p02 = e0 + sel32(mh, o0, e0);
    h.p13 = o0 + sel32(mh, m0, o0);
    h.p46 = e1 + sel32(mh, o1, e1);
    h.p57 = o1 + sel32(mh, m1, o1);
    return h;
}
__device__ __forceinline__ uint32_t vsum_pack(uint32_t a_even, uint32_t b_even, uint32_t a_odd, uint32_t b_odd) {
    const uint32_t ve = (a_even + b_even + 0x00020002u) >> 2, vo = (a_odd + b_odd + 0x00020002u) >> 2;
    return __byte_perm(ve, vo, 0x6240);   // bytes: even lane 0, odd lane 0, even lane 1, odd lane 1
}

__device__ __forceinline__ void block_load(const BlockCtx& B, const uint8_t* s_coef, const uint8_t* s_win,
                                           uint32_t (&p0)[8], uint32_t (&p1)[8], int (&c)[64]) {
    // prediction: eight rows of eight bytes, straight from the staged window (video_noasm.go:44-80)
    if (B.pred) {
        // s_win is 128-byte aligned and every pitch a multiple of 4, so the byte offset decides the word alignment
        const uint32_t a = B.win;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_win + (a & ~3u));
        const uint32_t wq = B.wpitch >> 2;
        const uint32_t sh = (a & 3u) * 8;
        const uint32_t mode = B.mode;
#define LOAD_ROW(R, W0, W1, W2)  \
    do {                         \
        W0 = wp[(R) * wq];       \
        W1 = wp[(R) * wq + 1];   \
        W2 = wp[(R) * wq + 2];   \
    } while (0)
#if MPEGB200_GENERIC_INTERP
        {
            const uint32_t mh = (mode & 1u) ? 0xffffffffu : 0u, mv = (mode & 2u) ? 0xffffffffu : 0u;
            uint32_t w0, w1, w2;
            LOAD_ROW(0, w0, w1, w2);
            HSum up = hsum_row_sel(w0, w1, w2, sh, mh);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                LOAD_ROW(r + 1, w0, w1, w2);   // the row below is staged for every mode (17 / 9 window rows)
                const HSum nx = hsum_row_sel(w0, w1, w2, sh, mh);
                p0[r] = vsum_pack(up.p02, sel32(mv, nx.p02, up.p02), up.p13, sel32(mv, nx.p13, up.p13));
                p1[r] = vsum_pack(up.p46, sel32(mv, nx.p46, up.p46), up.p57, sel32(mv, nx.p57, up.p57));
                up = nx;
            }
        }
#else
        if (mode == 0) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                uint32_t w0, w1, w2;
                LOAD_ROW(r, w0, w1, w2);
                p0[r] = __funnelshift_rc(w0, w1, sh);
                p1[r] = __funnelshift_rc(w1, w2, sh);
            }
        } else if (mode == 1) {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                uint32_t w0, w1, w2;
                LOAD_ROW(r, w0, w1, w2);
                p0[r] = avg2(__funnelshift_rc(w0, w1, sh), __funnelshift_rc(w0, w1, sh + 8));
                p1[r] = avg2(__funnelshift_rc(w1, w2, sh), __funnelshift_rc(w1, w2, sh + 8));
            }
        } else if (mode == 2) {
            uint32_t w0, w1, w2;
            LOAD_ROW(0, w0, w1, w2);
            uint32_t u0 = __funnelshift_rc(w0, w1, sh), u1 = __funnelshift_rc(w1, w2, sh);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                LOAD_ROW(r + 1, w0, w1, w2);
                const uint32_t n0 = __funnelshift_rc(w0, w1, sh), n1 = __funnelshift_rc(w1, w2, sh);
                p0[r] = avg2(u0, n0);
                p1[r] = avg2(u1, n1);
                u0 = n0;
                u1 = n1;
            }
        } else {
            uint32_t w0, w1, w2;
            LOAD_ROW(0, w0, w1, w2);
            HSum up = hsum_row(w0, w1, w2, sh);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                LOAD_ROW(r + 1, w0, w1, w2);
                const HSum dn = hsum_row(w0, w1, w2, sh);
                p0[r] = vsum_pack(up.p02, dn.p02, up.p13, dn.p13);
                p1[r] = vsum_pack(up.p46, dn.p46, up.p57, dn.p57);
                up = dn;
            }
        }
#endif
#undef LOAD_ROW
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) p0[r] = p1[r] = 0;   // intra: the residual is added to nothing
    }
    if (B.coded) {
        const int slot = (int)B.slot;
        const uint8_t* src = s_coef + slot * 128;
        const int sw = (slot & 7) << 4;  // 128-byte swizzle: 16-byte chunk index ^= row index mod 8
#if MPEGB200_PAIRED
        // The first step of the column pass only ever needs s1 +- s7, s3 +- s5, s2 +- s6 and s0 +- s4 (video.go:870-878).
        // With the two levels of such a pair side by side in one register, dp2a (int16 pair x int8 pair, summed)
        // premultiplies (video.go:744) AND forms the sum or the difference: 1 byte permute + 2 dp2a per pair
        // instead of 2 dp2a + 2 additions.  c[] then holds, per column i:
        //   c[i] = s0 + s4, c[32+i] = s0 - s4, c[8+i] = s1 + s7, c[56+i] = s1 - s7,
        //   c[24+i] = s3 + s5, c[40+i] = s5 - s3, c[16+i] = s2 + s6, c[48+i] = s2 - s6.
        uint32_t w[8][4];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + ((r << 4) ^ sw));
            w[r][0] = q.x; w[r][1] = q.y; w[r][2] = q.z; w[r][3] = q.w;
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = 2 * p + h;
                const uint32_t sel = h ? 0x7632u : 0x5410u;
#define PAIR(K, RA, RB)                                                                              \
    do {                                                                                            \
        const int ab = (int)__byte_perm(w[RA][p], w[RB][p], sel);                                   \
        c[RA * 8 + i] = __dp2a_lo(ab, kPairSum[K * 8 + i], K == 0 && i == 0 ? 128 : 0);             \
        c[RB * 8 + i] = __dp2a_lo(ab, kPairDif[K * 8 + i], K == 0 && i == 0 ? 128 : 0);             \
    } while (0)
                PAIR(0, 0, 4);
                PAIR(1, 1, 7);
                PAIR(2, 3, 5);
                PAIR(3, 2, 6);
#undef PAIR
            }
        }
#else
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 w = *reinterpret_cast<const uint4*>(src + ((r << 4) ^ sw));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int p = 0; p < 4; p++) {  // level * premultiplier (video.go:744) straight from the int16 pairs
                c[r * 8 + 2 * p] = __dp2a_lo((int)ww[p], premult(r * 8 + 2 * p), 0);
                c[r * 8 + 2 * p + 1] = __dp2a_lo((int)ww[p], premult(r * 8 + 2 * p + 1) << 8, 0);
            }
        }
        // The +128 of the final (x + 128) >> 8 (video.go:918-925) rides on the DC term: s0 enters every output of both
        // passes exactly once with weight +1 and never passes through a rounding shift (x1 = m0 - b1, x3 = m0 + b1).
        c[0] += 128;
#endif
    }
}

// d = (sat_u16(a) << 16) | sat_u16(b)
__device__ __forceinline__ uint32_t pack_sat_u16(int a, int b) {
    uint32_t d;
    asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// ((v >> 8) + pred) clamped to 0..255 for four pixels (addBlockToDest video.go:957-971; copyBlockToDest with pred = 0):
// t = v + (pred << 8) is clamped to 0..65535 first, its high byte is floor(t / 256) clamped to 0..255.
__device__ __forceinline__ uint32_t finish4(int t0, int t1, int t2, int t3) {
    return __byte_perm(pack_sat_u16(t1, t0), pack_sat_u16(t3, t2), 0x7531);
}

__device__ __forceinline__ void block_finish(const BlockCtx& B, const uint32_t (&p0)[8], const uint32_t (&p1)[8], int (&c)[64]) {
    if (!B.live) return;
    uint8_t* dst = B.dst;
    const uint32_t pitch = B.pitch;
    if (!B.coded) {  // predicted block without residual: the prediction is the result (skipped / cbp bit clear)
#pragma unroll
        for (int r = 0; r < 8; r++) *reinterpret_cast<uint2*>(dst + (size_t)r * pitch) = make_uint2(p0[r], p1[r]);
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; i++)  // columns, video.go:869-896
#if MPEGB200_PAIRED
        idct_pass8_pre(c[i], c[8 + i], c[16 + i], c[24 + i], c[32 + i], c[40 + i], c[48 + i], c[56 + i]);
#else
        idct_pass8(c[i], c[8 + i], c[16 + i], c[24 + i], c[32 + i], c[40 + i], c[48 + i], c[56 + i]);
#endif
#pragma unroll
    for (int r = 0; r < 8; r++) {  // rows, video.go:899-926, then copy/addBlockToDest (:943-971)
        int k[8], v[8];
#pragma unroll
        for (int x = 0; x < 4; x++) {   // prediction pixel x, shifted left by 8
            k[x] = (int)__byte_perm(p0[r], 0, 0x4404 | (x << 4));
            k[4 + x] = (int)__byte_perm(p1[r], 0, 0x4404 | (x << 4));
        }
        idct_row8(&c[r * 8], k, v);
        *reinterpret_cast<uint2*>(dst + (size_t)r * pitch) =
            make_uint2(finish4(v[0], v[1], v[2], v[3]), finish4(v[4], v[5], v[6], v[7]));
    }
}

#if defined(MPEGB200_EXPERIMENTS) && defined(MPEGB200_EXP_COEF_ALIAS)
// Occupancy probe (WRONG PIXELS): the third coefficient box lands on the first one, 8 KiB instead of 12 -> seven CTAs per SM
// with the same bytes fetched and the same instructions executed.
constexpr int kCoefSmem = 2 * kCoefBox * 128, kMinCtas = 7;
#define COEF_BOX_OFFSET(lane) (((lane) & 1) * (kCoefBox * 128))
#else
constexpr int kCoefSmem = kNT * 128, kMinCtas = 6;
#define COEF_BOX_OFFSET(lane) ((lane) * (kCoefBox * 128))
#endif

struct Smem {
    static constexpr int coef = 0;                               // kNT x 128, 1024-aligned, swizzled by TMA
    static constexpr int win = coef + kCoefSmem;                 // the window area: one strip or kG x 1280
    static constexpr int plan = win + kWinArea;                  // 640: the group's plan (head + body), then the mbarrier
    static constexpr int total_oneshot = plan + 640 + 16;
};

// The TMA boxes of one group.  Strip mode: lane 0 of warp 0 the luma strip, lane 0 of warp 1 the chroma strip.  Box mode:
// warp 0 the luma boxes, warp 1 the chroma boxes, lane = macroblock.  Warp 2 the coefficient boxes and the expect_tx;
// complete_tx of boxes issued before that arrival is fine: the phase cannot complete before the (single) arrival, and
// the transaction count is allowed to run negative meanwhile.
__device__ __forceinline__ void issue_group(const PlanHead& P, const PlanBox* boxes, const CUtensorMap* coef_map,
                                            const SlabMaps* slab_maps, uint8_t* s_coef, uint8_t* s_win, uint64_t* bar,
                                            int warp, int lane, uint32_t extra_tx) {
    if (warp == 2) {
        if (lane == 0) mbar_arrive_expect_tx(bar, P.tx_bytes + extra_tx);
        if (lane < (int)P.n_box)  // rows past n_blocks are zero-filled by the TMA unit
            tma_load_2d(s_coef + COEF_BOX_OFFSET(lane), coef_map, bar, 0, (int)(P.block0 + lane * kCoefBox));
    } else if (P.strip) {
        if (lane == 0) {
            const SlabMaps* maps = slab_maps + P.slab;
            if (warp == 0)
                tma_load_3d(s_win, maps->luma_strip, bar, P.sx >> 3, P.sy, P.z);
            else
                tma_load_4d(s_win + kStripLBytes, maps->chroma_strip, bar, P.scx >> 3, P.scy, 0, P.z);
        }
    } else if (lane < kG) {
        PlanBox t;                       // from the plan in global memory: like the head, not before griddepcontrol.wait
        *reinterpret_cast<uint4*>(&t) = ld_plan16(boxes + lane);
        if (t.pred) {
            const SlabMaps* maps = slab_maps + t.slab;
            uint8_t* w = s_win + lane * kWinBytes;
            if (warp == 0)
                tma_load_3d(w, maps->luma, bar, t.lx, t.ly, t.z);
            else  // one rank-4 box fetches the Cb and the Cr window (plane is the third dimension)
                tma_load_4d(w + kWinY, maps->chroma, bar, t.cx, t.cy, 0, t.z);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One-shot kernel: one CTA = one group plan = kG records.  The lanes that issue the boxes read the plan head straight
// from global memory (an L2 hit: the pre-pass has just written it) while a bulk copy brings head + body into shared
// memory on the same mbarrier as the tiles; then one thread per output block: interpolate, IDCT, add, saturate, store.
// While it waits, the CTA also asks for the tiles of group blockIdx.x + prefetch_dist to be fetched into L2
// (cp.async.bulk.prefetch.tensor): shared memory allows only six groups in flight per SM, which leaves the CTAs
// waiting on DRAM for a quarter of their life (profiles/); the CTA that decodes that group later finds its tiles in L2.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ PlanHead load_head(const GroupPlan* gp) {
    PlanHead h;
    const uint4* src = reinterpret_cast<const uint4*>(&gp->h);
    uint4* dst = reinterpret_cast<uint4*>(&h);
    dst[0] = ld_plan16(src);
    dst[1] = ld_plan16(src + 1);
    dst[2] = ld_plan16(src + 2);
    return h;
}

constexpr int kPlanSmemBytes = 640;   // head + body (624) rounded up to 16
static_assert(sizeof(PlanHead) + sizeof(PlanBody) <= kPlanSmemBytes, "plan copy size");

__global__ void __launch_bounds__(kNT, kMinCtas) fused_tma_kernel(const __grid_constant__ CUtensorMap coef_map,
                                                          const SlabMaps* __restrict__ slab_maps,
                                                          const GroupPlan* plans, uint32_t n_groups,
                                                          uint32_t prefetch_dist, int dbg) {
#ifndef MPEGB200_EXPERIMENTS
    (void)dbg;   // the fetch-only / arithmetic-only timing modes exist in experiment builds only
#endif
    extern __shared__ __align__(1024) uint8_t smem[];  // the swizzled TMA tile needs 1024-byte alignment
    uint8_t* s_coef = smem + Smem::coef;
    uint8_t* s_win = smem + Smem::win;
    const GroupPlan& P = *reinterpret_cast<const GroupPlan*>(smem + Smem::plan);   // head and body only
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + Smem::plan + kPlanSmemBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GroupPlan* gp = plans + blockIdx.x;

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        mbar_init(s_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the plan pre-pass has completed and its plans are visible
#ifdef MPEGB200_EXPERIMENTS
    if (dbg == 2) {   // measurement only: the arithmetic and the stores without the tile fetch (works on whatever shared memory holds)
        if (tid == 64) {
            mbar_arrive_expect_tx(s_bar, kPlanSmemBytes);
            bulk_load(smem + Smem::plan, gp, kPlanSmemBytes, s_bar);
        }
    } else
#endif
    {
        const PlanHead h = load_head(gp);
        issue_group(h, gp->box, &coef_map, slab_maps, s_coef, s_win, s_bar, warp, lane, kPlanSmemBytes);
        if (tid == 64) bulk_load(smem + Smem::plan, gp, kPlanSmemBytes, s_bar);
    }
    if (prefetch_dist && blockIdx.x + prefetch_dist < n_groups) {
        const GroupPlan* fp = gp + prefetch_dist;
        const PlanHead f = load_head(fp);
        if (warp == 2) {
            if (lane < (int)f.n_box) tma_prefetch_2d(&coef_map, 0, (int)(f.block0 + lane * kCoefBox));
        } else if (f.strip && lane == 0) {
            const SlabMaps* maps = slab_maps + f.slab;
            if (warp == 0)
                tma_prefetch_3d(maps->luma_strip, f.sx >> 3, f.sy, f.z);
            else
                tma_prefetch_4d(maps->chroma_strip, f.scx >> 3, f.scy, 0, f.z);
        }
    }

    mbar_wait(s_bar, 0);   // all tiles of the group and its plan have landed
#ifdef MPEGB200_EXPERIMENTS
    if (dbg == 1) return;  // measurement only: the fetch without the arithmetic
#endif
    BlockCtx B;
    block_setup(P.h, P.b, tid, B);
    if (!B.live) return;
    uint32_t p0[8], p1[8];
    int c[64];
    block_load(B, s_coef, s_win, p0, p1, c);
    block_finish(B, p0, p1, c);
}

#ifdef MPEGB200_EXPERIMENTS
// ------------------------------------------------------------------------------------------------
// (experiment builds only: measured slower than the one-shot kernel, profiles/r1_s3_video_summary.md)
// Streaming kernel: the registers are the second pipeline stage.  A thread needs shared memory only until its
// prediction (16 registers) and its premultiplied coefficients (64 registers) are loaded; the IDCT, the add and the
// stores run on registers.  Each CTA walks over groups b, b + grid, ...; the warp that is LAST to leave the load step
// of group i (a shared-memory counter tells it) fetches the tiles of group i + 1 into the same shared memory, so
// that they land while the IDCT of group i runs, and nobody waits at a CTA-wide barrier.  Plans are double-buffered
// and fetched two groups ahead with a bulk copy.
// ------------------------------------------------------------------------------------------------

struct SmemS {
    static constexpr int coef = 0;                                   // kNT x 128, 1024-aligned, swizzled by TMA
    static constexpr int win = coef + kNT * 128;
    static constexpr int body = win + kWinArea;                      // the current group's PlanBody
    static constexpr int head = body + (int)sizeof(PlanBody);        // 2 x PlanHead: current and next group
    static constexpr int bar = head + 2 * (int)sizeof(PlanHead);     // full, head[0], head[1], counter
    static constexpr int total = bar + 32;
};
// 6 CTAs per SM need at most 35,600 bytes each (measured: the one-shot kernel's 35,600 fit six times, 36,672 did not)
// and at most 96 registers (16,384 per scheduler / (96 x 32) = 5 warps per scheduler = 20 per SM; 104 gives 16).
static_assert(SmemS::total <= 35600, "streaming kernel: shared memory budget for 6 CTAs per SM");

__global__ void __maxnreg__(96) fused_stream_kernel(const __grid_constant__ CUtensorMap coef_map,
                                                   const SlabMaps* __restrict__ slab_maps,
                                                   const GroupPlan* __restrict__ plans, uint32_t n_groups) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* s_coef = smem + SmemS::coef;
    uint8_t* s_win = smem + SmemS::win;
    PlanBody& s_body = *reinterpret_cast<PlanBody*>(smem + SmemS::body);
    PlanHead* s_head = reinterpret_cast<PlanHead*>(smem + SmemS::head);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + SmemS::bar);
    uint64_t* bar_head = bar_full + 1;   // [2]
    uint32_t* s_left = reinterpret_cast<uint32_t*>(bar_full + 3);   // warps that have left the load step
    const uint32_t stride = gridDim.x;
    uint32_t g = blockIdx.x;

    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        mbar_init(bar_full, 1);
        mbar_init(bar_head, 1);
        mbar_init(bar_head + 1, 1);
        *s_left = 0;
        fence_barrier_init();
        // the heads of the first two groups
        mbar_arrive_expect_tx(bar_head, (uint32_t)sizeof(PlanHead));
        bulk_load(&s_head[0], &plans[g].h, (uint32_t)sizeof(PlanHead), bar_head);
        if (g + stride < n_groups) {
            mbar_arrive_expect_tx(bar_head + 1, (uint32_t)sizeof(PlanHead));
            bulk_load(&s_head[1], &plans[g + stride].h, (uint32_t)sizeof(PlanHead), bar_head + 1);
        }
    }
    __syncthreads();
    mbar_wait(bar_head, 0);
    // first group: the three warps share the fetch (the body rides on the same mbarrier as the tiles)
    issue_group(s_head[0], plans[g].box, &coef_map, slab_maps, s_coef, s_win, bar_full, warp, lane, (uint32_t)sizeof(PlanBody));
    if (tid == 64) bulk_load(&s_body, &plans[g].b, (uint32_t)sizeof(PlanBody), bar_full);

    for (uint32_t i = 0; g < n_groups; i++, g += stride) {
        const uint32_t buf = i & 1;
        mbar_wait(bar_head + buf, (i >> 1) & 1);   // long since complete (the fetch below needed it); makes the head visible here
        mbar_wait(bar_full, i & 1);           // the group's tiles and plan body have landed
        BlockCtx B;
        block_setup(s_head[buf], s_body, tid, B);
        uint32_t p0[8], p1[8];
        int c[64];
        if (B.live) block_load(B, s_coef, s_win, p0, p1, c);
        // This warp has left shared memory (windows, coefficients, plan).  The last warp to get here fetches the next
        // group's tiles; the other two go straight on to their arithmetic.  (In the first round the three warps
        // issued their own share above; nobody can pass the wait on bar_full before all of that has landed.)
        __syncwarp();
        uint32_t left = 0;
        if (lane == 0) {
            __threadfence_block();
            left = atomicAdd(s_left, 1u);
        }
        left = __shfl_sync(0xffffffffu, left, 0);
        if (left == 2) {
            if (lane == 0) {
                *s_left = 0;
                __threadfence_block();
                fence_proxy_async();          // generic-proxy reads of the tiles before the async-proxy writes
            }
            __syncwarp();
            const uint32_t gn = g + stride;
            if (gn < n_groups) {
                mbar_wait(bar_head + (buf ^ 1), ((i + 1) >> 1) & 1);
                const PlanHead& N = s_head[buf ^ 1];
                // one warp plays all three roles of issue_group
                issue_group(N, plans[gn].box, &coef_map, slab_maps, s_coef, s_win, bar_full, 2, lane, (uint32_t)sizeof(PlanBody));
                if (lane == 0) bulk_load(&s_body, &plans[gn].b, (uint32_t)sizeof(PlanBody), bar_full);
                issue_group(N, plans[gn].box, &coef_map, slab_maps, s_coef, s_win, bar_full, 0, lane, 0);
                issue_group(N, plans[gn].box, &coef_map, slab_maps, s_coef, s_win, bar_full, 1, lane, 0);
                if (lane == 0 && gn + stride < n_groups) {   // and the head after it, into the slot this group has left
                    mbar_arrive_expect_tx(bar_head + buf, (uint32_t)sizeof(PlanHead));
                    bulk_load(&s_head[buf], &plans[gn + stride].h, (uint32_t)sizeof(PlanHead), bar_head + buf);
                }
            }
        }
        block_finish(B, p0, p1, c);
    }
}
#endif  // MPEGB200_EXPERIMENTS

}  // namespace

size_t fused_plan_bytes(uint32_t n_mb) { return (size_t)((n_mb + kG - 1) / kG) * sizeof(GroupPlan); }

cudaError_t launch_fused_tma(const void* coef_map, const SlabMaps* d_maps, void* d_plans, const StreamInfo* d_streams,
                             int max_streams, const mpegb200_picture* d_pics, int n_pics, const mpegb200_mb* d_mbs,
                             uint32_t n_mb, uint32_t n_blocks, cudaStream_t stream, const cudaEvent_t* timing) {
    if (n_mb == 0) return cudaSuccess;
    static bool configured = false;
    int allow_strip = 1, prefetch_dist = kDefaultPrefetchDist, dbg = 0, pdl = 1;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total_oneshot);
        if (e != cudaSuccess) return e;
        configured = true;
    }
#ifdef MPEGB200_EXPERIMENTS
    // Experiment builds (make EXTRA=-DMPEGB200_EXPERIMENTS): A/B switches read once from the environment.
    // MPEGB200_FUSED=oneshot|stream picks the arithmetic kernel, MPEGB200_STRIP=0 makes the plan pre-pass stage every window
    // with its own boxes, MPEGB200_STREAM_CTAS sets the streaming grid, MPEGB200_PREFETCH_DIST the L2 prefetch distance,
    // MPEGB200_MEASURE=fetch|math times the tile fetch / the arithmetic alone (WRONG PIXELS), MPEGB200_PDL=0 plain stream order.
    static int x_variant = -1, x_strip = 1, x_ctas = 0, x_pd = kDefaultPrefetchDist, x_dbg = 0, x_pdl = 1;
    if (x_variant < 0) {
        const char* v = getenv("MPEGB200_FUSED");
        const int want = (v && strcmp(v, "stream") == 0) ? 1 : 0;
        cudaError_t e = cudaFuncSetAttribute(fused_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemS::total);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, per_sm = 0, per_sm1 = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_stream_kernel, kNT, SmemS::total);
        if (e != cudaSuccess) return e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, fused_tma_kernel, kNT, Smem::total_oneshot);
        x_ctas = sms * (per_sm > 0 ? per_sm : 1);
        const char* c = getenv("MPEGB200_STREAM_CTAS");
        if (c && atoi(c) > 0) x_ctas = atoi(c);
        const char* pd = getenv("MPEGB200_PREFETCH_DIST");
        if (pd) x_pd = atoi(pd) > 0 ? atoi(pd) : 0;
        const char* dm = getenv("MPEGB200_MEASURE");
        x_dbg = dm && strcmp(dm, "fetch") == 0 ? 1 : dm && strcmp(dm, "math") == 0 ? 2 : 0;
        const char* pl = getenv("MPEGB200_PDL");
        x_pdl = !(pl && pl[0] == '0');
        const char* st = getenv("MPEGB200_STRIP");
        x_strip = !(st && st[0] == '0');
        fprintf(stderr, "[mpegb200 EXPERIMENT BUILD] fused variant %d; CTAs per SM: streaming %d, one-shot %d; smem %d\n", want, per_sm,
                per_sm1, Smem::total_oneshot);
        x_variant = want;
    }
    allow_strip = x_strip, prefetch_dist = x_pd, dbg = x_dbg, pdl = x_pdl;
#endif
    const uint32_t n_groups = (n_mb + kG - 1) / kG;
    GroupPlan* plans = reinterpret_cast<GroupPlan*>(d_plans);
    if (timing) cudaEventRecord(timing[0], stream);
    plan_kernel<<<(n_groups + kPlanGroupsPerCta - 1) / kPlanGroupsPerCta, 16 * kPlanGroupsPerCta, 0, stream>>>(
        plans, d_streams, max_streams, d_pics, n_pics, d_mbs, n_mb, n_blocks, allow_strip);
    if (timing) cudaEventRecord(timing[1], stream);
    const CUtensorMap& cm = *reinterpret_cast<const CUtensorMap*>(coef_map);
#ifdef MPEGB200_EXPERIMENTS
    if (x_variant == 1) {
        const uint32_t grid = n_groups < (uint32_t)x_ctas ? n_groups : (uint32_t)x_ctas;
        fused_stream_kernel<<<grid, kNT, SmemS::total, stream>>>(cm, d_maps, plans, n_groups);
        if (timing) cudaEventRecord(timing[2], stream);
        return cudaGetLastError();
    }
#endif
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_groups);
    cfg.blockDim = dim3(kNT);
    cfg.dynamicSmemBytes = Smem::total_oneshot;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && !timing) ? 1 : 0;   // with timing events between the two kernels there is nothing to overlap
    const SlabMaps* maps_arg = d_maps;
    const GroupPlan* plans_arg = plans;
    const uint32_t pf = (uint32_t)prefetch_dist;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fused_tma_kernel, cm, maps_arg, plans_arg, n_groups, pf, dbg);
    if (e != cudaSuccess) return e;
    if (timing) cudaEventRecord(timing[2], stream);
    return cudaGetLastError();
}

}  // namespace mpegb200
