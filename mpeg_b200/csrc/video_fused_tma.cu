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


#include "common.cuh"

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

// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack4_sat_u8(int v0, int v1, int v2, int v3) {
    return pack_sat_u8(v1, v0, pack_sat_u8(v3, v2, 0u));
}

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
    const int x0 = x4 - (((tmp1 - tmp2) * 362 + 128) >> 8);
    const int x1 = m0 - b1;
    const int x2 = (((s2 - s6) * 362 + 128) >> 8) - b3;
    const int x3 = m0 + b1;
    const int y3 = x1 + x2;
    const int y4 = x3 + b3;
    const int y5 = x1 - x2;
    const int y6 = x3 - b3;
    const int y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
    s0 = b7 + y4;
    s1 = x4 + y3;
    s2 = y5 - x0;
    s3 = y6 - y7;
    s4 = y6 + y7;
    s5 = x0 + y5;
    s6 = y3 - x4;
    s7 = y4 - b7;
}

// row pass with the final (x + 128) >> 8 of video.go:918-925 folded into the last additions
__device__ __forceinline__ void idct_row8(const int* s, int* o) {
    const int b1 = s[4];
    const int b3 = s[2] + s[6];
    const int b4 = s[5] - s[3];
    const int tmp1 = s[1] + s[7];
    const int tmp2 = s[3] + s[5];
    const int b6 = s[1] - s[7];
    const int b7 = tmp1 + tmp2;
    const int m0 = s[0];
    const int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    const int x0 = x4 - (((tmp1 - tmp2) * 362 + 128) >> 8);
    const int x1 = m0 - b1;
    const int x2 = (((s[2] - s[6]) * 362 + 128) >> 8) - b3;
    const int x3 = m0 + b1;
    const int y3 = x1 + x2;
    const int y4 = x3 + b3;
    const int y5 = x1 - x2;
    const int y6 = x3 - b3;
    const int y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
    o[0] = (b7 + y4 + 128) >> 8;
    o[1] = (x4 + y3 + 128) >> 8;
    o[2] = (y5 - x0 + 128) >> 8;
    o[3] = (y6 - y7 + 128) >> 8;
    o[4] = (y6 + y7 + 128) >> 8;
    o[5] = (x0 + y5 + 128) >> 8;
    o[6] = (y3 - x4 + 128) >> 8;
    o[7] = (y4 - b7 + 128) >> 8;
}

__device__ __forceinline__ uint32_t avg2(uint32_t a, uint32_t b) { return __vavgu4(a, b); }  // (a+b+1)>>1 per byte
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // (a+b+c+d+2)>>2 per byte
    const uint32_t m = 0x00ff00ffu;
    const uint32_t lo = (((a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u) >> 2) & m;
    const uint32_t hi = ((((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u) >> 2) & m;
    return lo | (hi << 8);
}

constexpr int kG = 16;                 // macroblock records per CTA
constexpr int kNT = 6 * kG;            // one thread per output block (8x8)
constexpr int kCoefBox = 32;           // blocks per coefficient TMA box (32 x 128 B = 4 KiB)
// Windows are staged with a per-macroblock row phase: macroblock j's luma box starts (j & 3) rows above its window
// (the phase masks follow from the box heights in common.cuh), so that the same window row of different macroblocks lands in
// different shared-memory banks (a 32-byte row is 8 banks; without the phase the 32 block threads of a warp
// all pulled their row from the same 8 banks: profiles/r1_final_video_summary.md).
constexpr int kWinY = (32 * kLumaBoxRows + 127) / 128 * 128;          // luma box 32x20 = 640 B
constexpr int kWinC = 32 * kChromaBoxRows;                            // chroma box: Cb 32x9 then Cr 32x9 = 576 B in 640
constexpr int kWinBytes = kWinY + (2 * kWinC + 127) / 128 * 128;      // 1280 per macroblock, both boxes 128-byte aligned
constexpr int kWinTx = 32 * kLumaBoxRows + 2 * kWinC;                 // bytes the two boxes deliver

// ------------------------------------------------------------------------------------------------
// Group plan: everything a CTA needs to know about its kG records, computed once by a pre-pass
// (plan_kernel) so that the decode, the record -> picture -> stream pointer chase and the sorting of
// the output blocks are neither repeated nor serialised inside the arithmetic kernel.
// ------------------------------------------------------------------------------------------------
struct PlanMb {               // 32 bytes
    uint8_t* dst_y;           // destination of the macroblock's luma (row 0, col 0 of the MB)
    uint32_t dst_c_off;       // Cb destination = dst_y + dst_c_off; Cr at + (luma_w/2) * chroma_h
    uint16_t luma_w, chroma_h;
    uint16_t mcw;             // byte 0 luma, byte 1 chroma: (x & 15) | mode << 4 | 0x80 if predicted;
                              // mode bit 0 = horizontal half-pel, bit 1 = vertical half-pel
    uint16_t slab;            // tensor-map pair of the stream's slab
    int16_t lx, ly, cx, cy;   // box origins (x folded into [0, pitch) and aligned down to 16)
    uint16_t z;               // 3 * slot + reference buffer
    uint16_t pad;
};
static_assert(sizeof(PlanMb) == 32, "PlanMb size");

struct GroupPlan {            // 768 bytes = 48 x 16
    uint32_t n_box, n_pred, block0, pad0;
    uint16_t map[kNT];        // output-block list, sorted by (coded?, interpolation mode)
    PlanMb mb[kG];
    uint8_t pad1[768 - 16 - 2 * kNT - 32 * kG];
};
static_assert(sizeof(GroupPlan) == 768, "GroupPlan size");

// Output-block list entry: [3:0] macroblock in group, [6:4] block 0..5, [7] coded, [14:8] coefficient slot.
constexpr uint32_t kNoBlock = 0xFFFFu;

constexpr int kPlanGroupsPerCta = 8;   // 16 lanes per group

__global__ void __launch_bounds__(16 * kPlanGroupsPerCta) plan_kernel(GroupPlan* __restrict__ plans,
                                                                     const StreamInfo* __restrict__ streams,
                                                                     int max_streams,
                                                                     const mpegb200_picture* __restrict__ pics,
                                                                     int n_pics, const mpegb200_mb* __restrict__ mbs,
                                                                     uint32_t n_mb, uint32_t n_blocks) {
    __shared__ __align__(16) GroupPlan s_plan[kPlanGroupsPerCta];
    __shared__ uint32_t s_cnt[kPlanGroupsPerCta][10];  // [0..7] bins, [8] coded blocks to fetch, [9] predicted MBs
    const int tid = threadIdx.x, gl = tid >> 4, lane = tid & 15;
    const uint32_t n_groups = (n_mb + kG - 1) / kG;
    const uint32_t group = blockIdx.x * kPlanGroupsPerCta + gl;
    GroupPlan& P = s_plan[gl];
    for (int i = lane; i < kNT; i += 16) P.map[i] = (uint16_t)kNoBlock;
    if (lane < 10) s_cnt[gl][lane] = 0;
    {
        PlanMb z;
        memset(&z, 0, sizeof(z));
        P.mb[lane] = z;
    }
    __syncthreads();

    const uint32_t m0 = group * (uint32_t)kG;
    const bool have = group < n_groups && m0 + lane < n_mb;
    uint32_t out_mask = 0, cbp_r = 0, rel_r = 0, bins = 0, pos = 0, pos_hi = 0;
    uint32_t block0 = 0;
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
                    const uint32_t lw = si.luma_w, cw = lw >> 1;
                    uint8_t* dst = si.base + (size_t)dst_b * si.buf_stride;
                    PlanMb c;
                    memset(&c, 0, sizeof(c));
                    c.dst_y = dst + (size_t)(row << 4) * lw + (col << 4);
                    c.dst_c_off = (uint32_t)((size_t)lw * si.luma_h + (size_t)(row << 3) * cw + (col << 3) -
                                             ((size_t)(row << 4) * lw + (col << 4)));
                    c.luma_w = (uint16_t)lw;
                    c.chroma_h = (uint16_t)(si.luma_h >> 1);
                    c.slab = si.slab;
                    const bool predicted = (flags & MPEGB200_MB_PREDICT) != 0 && !(flags & MPEGB200_MB_INTRA);
                    uint32_t mode_y = 0, mode_c = 0;
                    if (predicted) {  // window origins, video_noasm.go:29-42
                        int lx = (int)(col << 4) + (mv_h >> 1);
                        int ly = (int)(row << 4) + (mv_v >> 1);
                        const int cmh = mv_h / 2, cmv = mv_v / 2;  // toward zero
                        int cx = (int)(col << 3) + (cmh >> 1);
                        int cy = (int)(row << 3) + (cmv >> 1);
                        // The reference indexes linearly (si = y*stride + x, video_noasm.go:31,39), so a window
                        // starting left of column 0 really starts near the end of the row above.  Fold x into
                        // [0, pitch): same bytes, and inside the tensor's (overlapping) rows.
                        {
                            int q = lx / (int)lw, r = lx - q * (int)lw;
                            if (r < 0) { r += (int)lw; q--; }
                            lx = r;
                            ly += q;
                            q = cx / (int)cw;
                            r = cx - q * (int)cw;
                            if (r < 0) { r += (int)cw; q--; }
                            cx = r;
                            cy += q;
                        }
                        mode_y = (uint32_t)((mv_h & 1) | ((mv_v & 1) << 1));
                        mode_c = (uint32_t)((cmh & 1) | ((cmv & 1) << 1));
                        c.mcw = (uint16_t)(((lx & 15) | (mode_y << 4) | 0x80) | (((cx & 15) | (mode_c << 4) | 0x80) << 8));
                        c.lx = (int16_t)(lx & ~15);  // the TMA unit needs x on a 16-byte boundary
                        c.ly = (int16_t)ly;
                        c.cx = (int16_t)(cx & ~15);
                        c.cy = (int16_t)cy;
                        c.z = (uint16_t)(si.slot * 3 + ((flags & MPEGB200_MB_REF_BWD) ? bwd_b : fwd_b));
                        atomicAdd(&s_cnt[gl][9], 1u);
                    }
                    P.mb[lane] = c;
                    if (ncoded) atomicMax(&s_cnt[gl][8], rel + ncoded);
                    // every 8x8 block whose pixels this record defines goes on the output list, binned by
                    // (coded?, interpolation mode) so that the threads of a warp take the same code path
                    out_mask = predicted ? 0x3fu : cbp;
                    cbp_r = cbp;
                    rel_r = rel;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        if (out_mask & (0x20u >> k)) {
                            const uint32_t bin = ((cbp & (0x20u >> k)) ? 0u : 4u) + (predicted ? (k < 4 ? mode_y : mode_c) : 0u);
                            const uint32_t pp = atomicAdd(&s_cnt[gl][bin], 1u);  // < 96
                            bins |= bin << (4 * k);
                            if (k < 4) pos |= pp << (8 * k); else pos_hi |= pp << (8 * (k - 4));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (out_mask) {
        uint32_t run = 0, base_of[8];
#pragma unroll
        for (int b = 0; b < 8; b++) {
            base_of[b] = run;
            run += s_cnt[gl][b];
        }
        uint32_t slot = rel_r;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (out_mask & (0x20u >> k)) {
                const uint32_t bin = (bins >> (4 * k)) & 15u;
                const uint32_t pp = (k < 4 ? pos >> (8 * k) : pos_hi >> (8 * (k - 4))) & 0xffu;
                const bool coded = cbp_r & (0x20u >> k);
                uint32_t bsel = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) bsel = bin == (uint32_t)b ? base_of[b] : bsel;
                P.map[bsel + pp] = (uint16_t)((uint32_t)lane | ((uint32_t)k << 4) | (coded ? 0x80u : 0u) | (slot << 8));
                if (coded) slot++;
            }
        }
    }
    if (lane == 0) {
        P.n_box = (s_cnt[gl][8] + kCoefBox - 1) / kCoefBox;
        P.n_pred = s_cnt[gl][9];
        P.block0 = block0;
        P.pad0 = 0;
    }
    __syncthreads();
    if (group < n_groups) {  // 768 bytes out, 16 bytes per lane per step
        const uint4* src = reinterpret_cast<const uint4*>(&P);
        uint4* dstp = reinterpret_cast<uint4*>(plans + group);
#pragma unroll
        for (int i = 0; i < 3; i++) dstp[lane + 16 * i] = src[lane + 16 * i];
    }
}

struct Smem {
    static constexpr int coef = 0;                               // kNT x 128, 1024-aligned, swizzled by TMA
    static constexpr int win = coef + kNT * 128;                 // kG x 1408
    static constexpr int plan = win + kG * kWinBytes;            // 768: the group's plan
    static constexpr int bar = plan + (int)sizeof(GroupPlan);    // 8 bytes
    static constexpr int total = bar + 16;
};

// One output block in three steps, so that a kernel can put a barrier between the part that reads shared memory
// and the part that only works on registers:
//   block_setup : the block-list entry -> destination, window address, mode (reads the plan)
//   block_load  : interpolate the prediction from the staged window, fetch + premultiply the coefficients
//   block_finish: IDCT, add, saturate, store
struct BlockCtx {
    uint8_t* dst;
    uint32_t pitch;
    uint32_t win;      // byte offset of the block's pixel (0,0) in the window area
    uint32_t slot;     // coefficient slot in the tile
    uint32_t mode;     // bit 0 horizontal, bit 1 vertical half-pel
    bool live, pred, coded;
};

__device__ __forceinline__ void block_setup(const GroupPlan& P, int t, uint32_t phases, BlockCtx& B) {
    const uint32_t e = P.map[t];
    B.live = e != kNoBlock;
    const int j = e & 15, k = (e >> 4) & 7;
    B.coded = B.live && (e & 0x80u);
    B.slot = e >> 8;
    const PlanMb& cx = P.mb[j];
    const uint32_t mcb = (uint32_t)(cx.mcw >> (k < 4 ? 0 : 8)) & 0xffu;
    B.pred = B.live && (mcb & 0x80u);
    B.mode = (mcb >> 4) & 3u;
    if (k < 4) {
        B.pitch = cx.luma_w;
        B.dst = cx.dst_y + (size_t)((k >> 1) * 8) * B.pitch + (k & 1) * 8;
    } else {
        B.pitch = cx.luma_w >> 1;
        B.dst = cx.dst_y + cx.dst_c_off + (k == 5 ? B.pitch * (uint32_t)cx.chroma_h : 0u);
    }
    const uint32_t ph = (uint32_t)j & (k < 4 ? phases & 0xffu : phases >> 8);
    B.win = j * kWinBytes + (k < 4 ? (k >> 1) * 256 + (k & 1) * 8 : kWinY + (k - 4) * kWinC) + ph * 32 + (mcb & 15u);
}

__device__ __forceinline__ void block_load(const BlockCtx& B, const uint8_t* s_coef, const uint8_t* s_win,
                                           uint32_t (&p0)[8], uint32_t (&p1)[8], int (&c)[64]) {
    // prediction: eight rows of eight bytes, straight from the staged window (video_noasm.go:44-80)
    if (B.pred) {
        // s_win is 128-byte aligned, so the byte offset decides the word alignment
        const uint32_t a = B.win;
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_win + (a & ~3u));
        const uint32_t sh = (a & 3u) * 8;
        const uint32_t mode = B.mode;
#define LOAD_ROW(R, W0, W1, W2) \
    do {                        \
        W0 = wp[(R) * 8];       \
        W1 = wp[(R) * 8 + 1];   \
        W2 = wp[(R) * 8 + 2];   \
    } while (0)
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
            uint32_t u0 = __funnelshift_rc(w0, w1, sh), u1 = __funnelshift_rc(w1, w2, sh);
            uint32_t s0 = __funnelshift_rc(w0, w1, sh + 8), s1 = __funnelshift_rc(w1, w2, sh + 8);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                LOAD_ROW(r + 1, w0, w1, w2);
                const uint32_t n0 = __funnelshift_rc(w0, w1, sh), n1 = __funnelshift_rc(w1, w2, sh);
                const uint32_t t0 = __funnelshift_rc(w0, w1, sh + 8), t1 = __funnelshift_rc(w1, w2, sh + 8);
                p0[r] = avg4(u0, s0, n0, t0);
                p1[r] = avg4(u1, s1, n1, t1);
                u0 = n0; u1 = n1; s0 = t0; s1 = t1;
            }
        }
#undef LOAD_ROW
    }
    if (B.coded) {
        const int slot = (int)B.slot;
        const uint8_t* src = s_coef + slot * 128;
        const int sw = (slot & 7) << 4;  // 128-byte swizzle: 16-byte chunk index ^= row index mod 8
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
    }
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
        idct_pass8(c[i], c[8 + i], c[16 + i], c[24 + i], c[32 + i], c[40 + i], c[48 + i], c[56 + i]);
#pragma unroll
    for (int r = 0; r < 8; r++) {  // rows, video.go:899-926, then copy/addBlockToDest (:943-971)
        int v[8];
        idct_row8(&c[r * 8], v);
        if (B.pred) {
#pragma unroll
            for (int x = 0; x < 4; x++) {
                v[x] = (int)__dp4a(p0[r], 1u << (8 * x), (uint32_t)v[x]);
                v[4 + x] = (int)__dp4a(p1[r], 1u << (8 * x), (uint32_t)v[4 + x]);
            }
        }
        *reinterpret_cast<uint2*>(dst + (size_t)r * pitch) =
            make_uint2(pack4_sat_u8(v[0], v[1], v[2], v[3]), pack4_sat_u8(v[4], v[5], v[6], v[7]));
    }
}

__device__ __forceinline__ void process_block(const GroupPlan& P, const uint8_t* s_coef, const uint8_t* s_win, int t,
                                              uint64_t* bar, uint32_t parity, uint32_t phases) {
    BlockCtx B;
    block_setup(P, t, phases, B);
    if (!B.live) return;
    mbar_wait(bar, parity);   // all tiles of the group have landed
    uint32_t p0[8], p1[8];
    int c[64];
    block_load(B, s_coef, s_win, p0, p1, c);
    block_finish(B, p0, p1, c);
}

// ------------------------------------------------------------------------------------------------
// The arithmetic kernel: one CTA = one group plan = kG records.  Load the plan (48 x 16 bytes), issue
// every TMA box (warp w: plane w of each predicted macroblock; warp 0 also the coefficient boxes), wait on
// the one mbarrier, then one thread per output block: interpolate, IDCT, add, saturate, store.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT) fused_tma_kernel(const __grid_constant__ CUtensorMap coef_map,
                                                       const SlabMaps* __restrict__ slab_maps,
                                                       const GroupPlan* __restrict__ plans, uint32_t phases) {
    extern __shared__ __align__(1024) uint8_t smem[];  // the swizzled TMA tile needs 1024-byte alignment
    uint8_t* s_coef = smem + Smem::coef;
    uint8_t* s_win = smem + Smem::win;
    GroupPlan& P = *reinterpret_cast<GroupPlan*>(smem + Smem::plan);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + Smem::bar);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid < 48) reinterpret_cast<uint4*>(&P)[tid] = reinterpret_cast<const uint4*>(plans + blockIdx.x)[tid];
    if (tid == 64) {
        if (smem_u32(smem) & 1023u) __trap();
        mbar_init(s_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();

    if (lane < kG && (P.mb[lane].mcw & 0x80u)) {
        const PlanMb& t = P.mb[lane];
        const SlabMaps* maps = slab_maps + t.slab;
        uint8_t* w = s_win + lane * kWinBytes;
        if (warp == 0)
            tma_load_3d(w, maps->luma, s_bar, t.lx, t.ly - (int)(lane & (phases & 0xffu)), t.z);
        else if (warp == 1)  // one rank-4 box fetches the Cb and the Cr window (plane is the third dimension)
            tma_load_4d(w + kWinY, maps->chroma, s_bar, t.cx, t.cy - (int)(lane & (phases >> 8)), 0, t.z);
    }
    if (warp == 0) {
        // complete_tx of boxes issued before this arrival is fine: the phase cannot complete before the
        // (single) arrival, and the transaction count is allowed to run negative meanwhile
        if (lane == 0) mbar_arrive_expect_tx(s_bar, P.n_box * (kCoefBox * 128) + P.n_pred * kWinTx);
        if (lane < (int)P.n_box)  // rows past n_blocks are zero-filled by the TMA unit
            tma_load_2d(s_coef + lane * (kCoefBox * 128), &coef_map, s_bar, 0, (int)(P.block0 + lane * kCoefBox));
    }

    process_block(P, s_coef, s_win, tid, s_bar, 0, phases);
}


}  // namespace

size_t fused_plan_bytes(uint32_t n_mb) { return (size_t)((n_mb + kG - 1) / kG) * sizeof(GroupPlan); }

cudaError_t launch_fused_tma(const void* coef_map, const SlabMaps* d_maps, void* d_plans, const StreamInfo* d_streams,
                             int max_streams, const mpegb200_picture* d_pics, int n_pics, const mpegb200_mb* d_mbs,
                             uint32_t n_mb, uint32_t n_blocks, cudaStream_t stream) {
    if (n_mb == 0) return cudaSuccess;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const uint32_t n_groups = (n_mb + kG - 1) / kG;
    GroupPlan* plans = reinterpret_cast<GroupPlan*>(d_plans);
    plan_kernel<<<(n_groups + kPlanGroupsPerCta - 1) / kPlanGroupsPerCta, 16 * kPlanGroupsPerCta, 0, stream>>>(
        plans, d_streams, max_streams, d_pics, n_pics, d_mbs, n_mb, n_blocks);
    constexpr uint32_t phases = (uint32_t)(kLumaBoxRows - 17) | ((uint32_t)(kChromaBoxRows - 9) << 8);
    fused_tma_kernel<<<n_groups, kNT, Smem::total, stream>>>(*reinterpret_cast<const CUtensorMap*>(coef_map), d_maps, plans, phases);
    return cudaGetLastError();
}

}  // namespace mpegb200
