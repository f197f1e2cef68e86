// video_fused_tma.cu -- the B200 fast path of the fused kernel: motion compensation + 8x8 IDCT +
// residual add / intra store (predictMacroblock/copyMacroblock video.go:608-637, video_noasm.go:28-80;
// idct video.go:801-928; copy/add*ToDest video.go:943-1002), with every tile moved by TMA.
//
// v1 (video_kernels.cu) spent 2/3 of its instructions on address arithmetic for staging and on
// re-aligning unaligned windows (profiles/r1_v1_fused_summary.md).  Here:
//   * coefficients: one cp.async.bulk.tensor per 32 blocks, 128-byte swizzle, so that one thread per
//     block reads its eight 16-byte rows bank-conflict free;
//   * reference windows: three cp.async.bulk.tensor per predicted macroblock (32x17 luma, 32x9 Cb,
//     32x9 Cr).  The TMA unit wants the innermost start coordinate on a 16-byte boundary (measured:
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
constexpr int kNT = 6 * kG;            // one thread per potential coded block
constexpr int kCoefBox = 32;           // blocks per coefficient TMA box (32 x 128 B = 4 KiB)
constexpr int kWinY = 640, kWinC = 384;            // bytes reserved per window (544 / 288 used), 128-aligned
constexpr int kWinBytes = kWinY + 2 * kWinC;       // 1408 per macroblock
constexpr int kWinTx = 32 * 17 + 2 * 32 * 9;       // bytes the three boxes deliver: 1120
constexpr int kBlkPitch = 72;          // prediction/pixel tile: 64 B per 8x8 block + 8: conflict-free for block threads

struct MbCtx {                // 32 bytes
    uint8_t* dst_y;           // destination of the macroblock's luma (row 0, col 0 of the MB)
    uint8_t* dst_c;           // destination of its Cb; Cr at + chroma_bytes
    uint32_t chroma_bytes;
    uint16_t luma_w;
    uint8_t flags, cbp, valid, mask;  // mask: blocks whose pixels are defined (all six if predicted, else cbp)
    uint8_t ox_y, ox_c;       // byte offset of pixel (0,0) inside the staged 32-byte rows (x & 15)
    uint8_t mode_y, mode_c;   // bit 0: horizontal half-pel, bit 1: vertical half-pel (luma / chroma vector)
    uint16_t rel_block;
};
static_assert(sizeof(MbCtx) == 32, "MbCtx size");

struct Smem {
    static constexpr int coef = 0;                               // kNT x 128, 1024-aligned, swizzled by TMA
    static constexpr int win = coef + kNT * 128;                 // kG x 1152
    static constexpr int pix = win + kG * kWinBytes;             // kNT x 72
    static constexpr int ctx = pix + kNT * kBlkPitch;            // kG x 40
    static constexpr int map = ctx + kG * (int)sizeof(MbCtx);    // kNT bytes
    static constexpr int bar = (map + kNT + 7) & ~7;             // 8 bytes
    static constexpr int mc = bar + 8;                           // kG x 4: interpolation word per macroblock
    static constexpr int total = mc + kG * 4 + 8;                // + the two CTA counters
};

__global__ void __launch_bounds__(kNT) fused_tma_kernel(const __grid_constant__ CUtensorMap coef_map,
                                                       const SlabMaps* __restrict__ slab_maps,
                                                       const StreamInfo* __restrict__ streams, int max_streams,
                                                       const mpegb200_picture* __restrict__ pics, int n_pics,
                                                       const mpegb200_mb* __restrict__ mbs, uint32_t n_mb,
                                                       uint32_t n_blocks) {
    extern __shared__ __align__(1024) uint8_t smem[];  // the swizzled TMA tile needs 1024-byte alignment
    uint8_t* s_coef = smem + Smem::coef;
    uint8_t* s_win = smem + Smem::win;
    uint8_t* s_pix = smem + Smem::pix;
    MbCtx* s_ctx = reinterpret_cast<MbCtx*>(smem + Smem::ctx);
    uint8_t* s_map = smem + Smem::map;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + Smem::bar);
    uint32_t* s_mc = reinterpret_cast<uint32_t*>(smem + Smem::mc);
    uint32_t& s_nb = s_mc[kG];
    uint32_t& s_npred = s_mc[kG + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t m0 = blockIdx.x * (uint32_t)kG;
    const int n_here = (int)min((uint32_t)kG, n_mb - m0);

    s_map[tid] = 0xFF;
    if (tid == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        mbar_init(s_bar, 1);
        fence_barrier_init();
        s_nb = 0;
        s_npred = 0;
    }
    __syncthreads();

    // ---------------- producer: record j is owned by lane j/3 of warp j%3, so that the (per lane
    // serialised) TMA issue is spread over all three warps ----------------
    {
        const int j = lane * 3 + warp;
        if (lane < (kG + 2) / 3 && j < kG) {
            const uint32_t block0 = mbs[m0].coeff_block;
            MbCtx c;
            c.valid = 0;
            c.flags = 0;
            c.cbp = 0;
            c.mask = 0;
            c.ox_y = c.ox_c = c.mode_y = c.mode_c = 0;
            uint32_t mcw = 0;
            if (j < n_here) {
                const uint4 raw = reinterpret_cast<const uint4*>(mbs)[m0 + j];
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
                            const uint32_t luma_bytes = lw * si.luma_h;
                            uint8_t* dst = si.base + (size_t)dst_b * si.buf_stride;
                            c.dst_y = dst + (size_t)(row << 4) * lw + (col << 4);
                            c.dst_c = dst + luma_bytes + (size_t)(row << 3) * cw + (col << 3);
                            c.chroma_bytes = cw * (si.luma_h >> 1);
                            c.luma_w = (uint16_t)lw;
                            c.flags = (uint8_t)flags;
                            c.cbp = (uint8_t)cbp;
                            c.rel_block = (uint16_t)rel;
                            c.valid = 1;
                            const bool predicted = (flags & MPEGB200_MB_PREDICT) != 0;
                            c.mask = predicted ? 0x3f : (uint8_t)cbp;
                            int k = 0;
                            for (int b = 0; b < 6; b++)
                                if (cbp & (0x20u >> b)) s_map[rel + k++] = (uint8_t)((j << 3) | b);
                            if (ncoded) atomicMax(&s_nb, rel + ncoded);
                            if (predicted) {  // window origins, video_noasm.go:29-42
                                const SlabMaps* maps = slab_maps + si.slab;
                                const int z = si.slot * 3 + (int)((flags & MPEGB200_MB_REF_BWD) ? bwd_b : fwd_b);
                                int lx = (int)(col << 4) + (mv_h >> 1);
                                int ly = (int)(row << 4) + (mv_v >> 1);
                                const int cmh = mv_h / 2, cmv = mv_v / 2;  // toward zero
                                int cx = (int)(col << 3) + (cmh >> 1);
                                int cy = (int)(row << 3) + (cmv >> 1);
                                // The reference indexes linearly (si = y*stride + x, video_noasm.go:31,39), so a
                                // window starting left of column 0 really starts near the end of the row above.
                                // Fold x into [0, pitch): same bytes, and inside the tensor's (overlapping) rows.
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
                                c.ox_y = (uint8_t)(lx & 15);
                                c.ox_c = (uint8_t)(cx & 15);
                                c.mode_y = (uint8_t)((mv_h & 1) | ((mv_v & 1) << 1));
                                c.mode_c = (uint8_t)((cmh & 1) | ((cmv & 1) << 1));
                                // interpolation word: byte 0 luma, byte 1 chroma: offset | mode << 4 | 0x80 (predicted)
                                mcw = (uint32_t)(c.ox_y | (c.mode_y << 4) | 0x80) | ((uint32_t)(c.ox_c | (c.mode_c << 4) | 0x80) << 8);
                                atomicAdd(&s_npred, 1u);
                                uint8_t* w = s_win + j * kWinBytes;
                                const int chroma_h = si.luma_h >> 1;
                                // the TMA unit needs a 16-byte aligned innermost coordinate
                                tma_load_3d(w, maps->luma, s_bar, lx & ~15, ly, z);
                                tma_load_3d(w + kWinY, maps->chroma, s_bar, cx & ~15, cy, z);
                                tma_load_3d(w + kWinY + kWinC, maps->chroma, s_bar, cx & ~15, cy + chroma_h, z);
                            }
                        }
                    }
                }
            }
            s_ctx[j] = c;
            s_mc[j] = mcw;
        }
    }
    __syncthreads();       // contexts, block map and totals visible
    if (warp == 0) {
        const uint32_t n_box = (s_nb + kCoefBox - 1) / kCoefBox;
        // complete_tx of the window boxes may already have been counted: the phase cannot complete
        // before this (single) arrival, and the transaction count is allowed to run negative meanwhile
        if (lane == 0) mbar_arrive_expect_tx(s_bar, n_box * (kCoefBox * 128) + s_npred * kWinTx);
        __syncwarp();
        if (lane < (int)n_box)  // rows past n_blocks are zero-filled by the TMA unit
            tma_load_2d(s_coef + lane * (kCoefBox * 128), &coef_map, s_bar, 0, (int)(mbs[m0].coeff_block + lane * kCoefBox));
    }
    mbar_wait(s_bar, 0);   // all tiles have landed

    // ---------------- interpolation: 8-pixel row pieces; per pass warp 0 / 1 take the luma of two
    // macroblocks, warp 2 their chroma; a thread's piece geometry is fixed, only the macroblock moves ----
    {
        int sub, wrow, pix_off;  // which of the pass's two macroblocks, byte offset of the row in its window, tile offset
        if (warp < 2) {
            const int y = lane >> 1, x0 = (lane & 1) * 8;
            sub = warp;
            wrow = y * 32 + x0;
            pix_off = ((y >> 3) * 2 + (x0 >> 3)) * kBlkPitch + (y & 7) * 8;
        } else {
            const int p = (lane >> 3) & 1, y = lane & 7;
            sub = lane >> 4;
            wrow = kWinY + p * kWinC + y * 32;
            pix_off = (4 + p) * kBlkPitch + y * 8;
        }
        const bool chroma = warp >= 2;
        const uint32_t win0 = smem_u32(s_win) + wrow;
        const int mc_shift = chroma ? 8 : 0;
#pragma unroll 2
        for (int j = sub; j < kG; j += 2) {
            const uint32_t mcw = s_mc[j] >> mc_shift;
            if (!(mcw & 0x80u)) continue;
            const uint32_t a = win0 + j * kWinBytes + (mcw & 15u);
            const uint32_t mode = mcw >> 4;
            const uint32_t aw = a & ~3u, sh = (a & 3u) * 8;
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                         : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(aw));
            uint32_t lo = __funnelshift_rc(w0, w1, sh), hi = __funnelshift_rc(w1, w2, sh);
            if (mode & 1) {
                const uint32_t lo1 = __funnelshift_rc(w0, w1, sh + 8), hi1 = __funnelshift_rc(w1, w2, sh + 8);
                if (mode & 2) {
                    uint32_t v0, v1, v2;
                    asm volatile("ld.shared.u32 %0, [%3+32];\n\tld.shared.u32 %1, [%3+36];\n\tld.shared.u32 %2, [%3+40];"
                                 : "=r"(v0), "=r"(v1), "=r"(v2) : "r"(aw));
                    lo = avg4(lo, lo1, __funnelshift_rc(v0, v1, sh), __funnelshift_rc(v0, v1, sh + 8));
                    hi = avg4(hi, hi1, __funnelshift_rc(v1, v2, sh), __funnelshift_rc(v1, v2, sh + 8));
                } else {
                    lo = avg2(lo, lo1);
                    hi = avg2(hi, hi1);
                }
            } else if (mode & 2) {
                uint32_t v0, v1, v2;
                asm volatile("ld.shared.u32 %0, [%3+32];\n\tld.shared.u32 %1, [%3+36];\n\tld.shared.u32 %2, [%3+40];"
                             : "=r"(v0), "=r"(v1), "=r"(v2) : "r"(aw));
                lo = avg2(lo, __funnelshift_rc(v0, v1, sh));
                hi = avg2(hi, __funnelshift_rc(v1, v2, sh));
            }
            *reinterpret_cast<uint2*>(s_pix + j * 6 * kBlkPitch + pix_off) = make_uint2(lo, hi);
        }
    }
    __syncthreads();

    // ---------------- one thread per coded block: premultiply, IDCT, add, saturate ----------------
    {
        const uint32_t bm = s_map[tid];
        if (bm != 0xFF) {
            const int j = bm >> 3, k = bm & 7;
            const uint32_t fl = s_ctx[j].flags;
            const bool add = (fl & MPEGB200_MB_PREDICT) && !(fl & MPEGB200_MB_INTRA);
            int c[64];
            const uint8_t* src = s_coef + tid * 128;
            const int sw = (tid & 7) << 4;  // 128-byte swizzle: 16-byte chunk index ^= row index mod 8
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
#pragma unroll
            for (int i = 0; i < 8; i++)  // columns, video.go:869-896
                idct_pass8(c[i], c[8 + i], c[16 + i], c[24 + i], c[32 + i], c[40 + i], c[48 + i], c[56 + i]);
            uint8_t* tile = s_pix + (j * 6 + k) * kBlkPitch;
#pragma unroll
            for (int r = 0; r < 8; r++) {  // rows, video.go:899-926, then copy/addBlockToDest (:943-971)
                int v[8];
                idct_row8(&c[r * 8], v);
                uint2* tp = reinterpret_cast<uint2*>(tile + r * 8);
                if (add) {
                    const uint2 pr = *tp;
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        v[x] = (int)__dp4a(pr.x, 1u << (8 * x), (uint32_t)v[x]);
                        v[4 + x] = (int)__dp4a(pr.y, 1u << (8 * x), (uint32_t)v[4 + x]);
                    }
                }
                *tp = make_uint2(pack4_sat_u8(v[0], v[1], v[2], v[3]), pack4_sat_u8(v[4], v[5], v[6], v[7]));
            }
        }
    }
    __syncthreads();

    // ---------------- tile -> frame: lanes across neighbouring macroblocks ----------------
    for (int it = tid; it < 16 * kG; it += kNT) {  // luma rows
        const int r = it / kG, j = it - r * kG;
        const MbCtx& c = s_ctx[j];
        if (!c.valid) continue;
        const int kl = (r >> 3) * 2;
        const uint8_t* t = s_pix + (j * 6 + kl) * kBlkPitch + (r & 7) * 8;
        const uint2 l = *reinterpret_cast<const uint2*>(t), rr = *reinterpret_cast<const uint2*>(t + kBlkPitch);
        uint8_t* d = c.dst_y + (size_t)r * c.luma_w;
        const bool left = c.mask & (0x20u >> kl), right = c.mask & (0x10u >> kl);
        if (left && right) {
            *reinterpret_cast<uint4*>(d) = make_uint4(l.x, l.y, rr.x, rr.y);
        } else if (left) {
            *reinterpret_cast<uint2*>(d) = l;
        } else if (right) {
            *reinterpret_cast<uint2*>(d + 8) = rr;
        }
    }
    for (int it = tid; it < 16 * kG; it += kNT) {  // chroma rows: 2 planes x 8
        const int pr = it / kG, j = it - pr * kG;
        const int p = pr >> 3, r = pr & 7;
        const MbCtx& c = s_ctx[j];
        if (!c.valid || !(c.mask & (0x02u >> p))) continue;
        uint8_t* d = c.dst_c + (p ? c.chroma_bytes : 0) + (size_t)r * (c.luma_w >> 1);
        *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(s_pix + (j * 6 + 4 + p) * kBlkPitch + r * 8);
    }
}

}  // namespace

cudaError_t launch_fused_tma(const void* coef_map, const SlabMaps* d_maps, const StreamInfo* d_streams, int max_streams,
                             const mpegb200_picture* d_pics, int n_pics, const mpegb200_mb* d_mbs, uint32_t n_mb,
                             uint32_t n_blocks, cudaStream_t stream) {
    if (n_mb == 0) return cudaSuccess;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const uint32_t grid = (n_mb + kG - 1) / kG;
    fused_tma_kernel<<<grid, kNT, Smem::total, stream>>>(*reinterpret_cast<const CUtensorMap*>(coef_map), d_maps,
                                                         d_streams, max_streams, d_pics, n_pics, d_mbs, n_mb, n_blocks);
    return cudaGetLastError();
}

}  // namespace mpegb200
