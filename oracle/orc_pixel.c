/*
 * orc_pixel.c -- CPU restatement of the reference's pixel kernels (TEST INFRASTRUCTURE).
 * See mpeg_oracle.h for the rules: nothing in the product may call this file.
 */
#include "mpeg_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* hash/fnv New64a, as mpeg_test.go:174,216 uses it */
uint64_t orc_fnv1a64(uint64_t h, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ULL;
    }
    return h;
}

/* ------------------------------------------------------------------------------------------
 * idct, video.go:801-928.
 *
 * One 8-point pass (video.go:870-895 for columns, :900-925 for rows, the two differ only
 * in the final (x+128)>>8).  Inputs s0..s7 are the eight samples of a column or row.
 * All arithmetic is Go `int` = int64; >> is arithmetic.
 * ---------------------------------------------------------------------------------------- */
static inline void pass8(int64_t s0, int64_t s1, int64_t s2, int64_t s3, int64_t s4, int64_t s5,
                         int64_t s6, int64_t s7, int64_t o[8]) {
    int64_t b1 = s4;
    int64_t b3 = s2 + s6;
    int64_t b4 = s5 - s3;
    int64_t tmp1 = s1 + s7;
    int64_t tmp2 = s3 + s5;
    int64_t b6 = s1 - s7;
    int64_t b7 = tmp1 + tmp2;
    int64_t m0 = s0;
    int64_t x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
    int64_t x0 = x4 - (((tmp1 - tmp2) * 362 + 128) >> 8);
    int64_t x1 = m0 - b1;
    int64_t x2 = (((s2 - s6) * 362 + 128) >> 8) - b3;
    int64_t x3 = m0 + b1;
    int64_t y3 = x1 + x2;
    int64_t y4 = x3 + b3;
    int64_t y5 = x1 - x2;
    int64_t y6 = x3 - b3;
    int64_t y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
    o[0] = b7 + y4;
    o[1] = x4 + y3;
    o[2] = y5 - x0;
    o[3] = y6 - y7;
    o[4] = y6 + y7;
    o[5] = x0 + y5;
    o[6] = y3 - x4;
    o[7] = y4 - b7;
}

void orc_idct_full(int64_t block[64]) {
    int64_t o[8];
    for (int i = 0; i < 8; i++) { /* columns, video.go:869-896 */
        pass8(block[0 * 8 + i], block[1 * 8 + i], block[2 * 8 + i], block[3 * 8 + i], block[4 * 8 + i],
              block[5 * 8 + i], block[6 * 8 + i], block[7 * 8 + i], o);
        for (int k = 0; k < 8; k++) block[k * 8 + i] = o[k];
    }
    for (int i = 0; i < 64; i += 8) { /* rows, video.go:899-926 */
        pass8(block[i + 0], block[i + 1], block[i + 2], block[i + 3], block[i + 4], block[i + 5], block[i + 6],
              block[i + 7], o);
        for (int k = 0; k < 8; k++) block[i + k] = (o[k] + 128) >> 8;
    }
}

void orc_idct(int64_t block[64], int max_index) {
    if (max_index >= 10) {
        orc_idct_full(block);
        return;
    }
    /* video.go:807-866: coefficients with zig-zag index < 10 live in rows 0..3 x columns 0..3,
     * so rows 4..7 are taken as zero and only columns 0..3 are transformed; whatever else is
     * in the array is ignored (and overwritten by the outputs). */
    int64_t o[8];
    for (int i = 0; i < 4; i++) {
        pass8(block[0 * 8 + i], block[1 * 8 + i], block[2 * 8 + i], block[3 * 8 + i], 0, 0, 0, 0, o);
        for (int k = 0; k < 8; k++) block[k * 8 + i] = o[k];
    }
    for (int i = 0; i < 64; i += 8) {
        pass8(block[i + 0], block[i + 1], block[i + 2], block[i + 3], 0, 0, 0, 0, o);
        for (int k = 0; k < 8; k++) block[i + k] = (o[k] + 128) >> 8;
    }
}

/* clamp, video.go:1014-1016 */
static inline uint8_t clamp_u8(int64_t n) { return (uint8_t)(n < 0 ? 0 : (n > 255 ? 255 : n)); }

/* video.go:943-956 */
void orc_copy_block_to_dest(const int64_t block[64], uint8_t* dest, int64_t index, int64_t scan) {
    for (int n = 0; n < 64; n += 8) {
        for (int k = 0; k < 8; k++) dest[index + k] = clamp_u8(block[n + k]);
        index += scan + 8;
    }
}

/* video.go:958-971 */
void orc_add_block_to_dest(const int64_t block[64], uint8_t* dest, int64_t index, int64_t scan) {
    for (int n = 0; n < 64; n += 8) {
        for (int k = 0; k < 8; k++) dest[index + k] = clamp_u8((int64_t)dest[index + k] + block[n + k]);
        index += scan + 8;
    }
}

/* video.go:973-987 */
void orc_copy_value_to_dest(int64_t value, uint8_t* dest, int64_t index, int64_t scan) {
    uint8_t val = clamp_u8(value);
    for (int n = 0; n < 64; n += 8) {
        for (int k = 0; k < 8; k++) dest[index + k] = val;
        index += scan + 8;
    }
}

/* video.go:989-1002 */
void orc_add_value_to_dest(int64_t value, uint8_t* dest, int64_t index, int64_t scan) {
    for (int n = 0; n < 64; n += 8) {
        for (int k = 0; k < 8; k++) dest[index + k] = clamp_u8((int64_t)dest[index + k] + value);
        index += scan + 8;
    }
}

/* ------------------------------------------------------------------------------------------
 * Frame buffers, video.go:333-355.
 * ---------------------------------------------------------------------------------------- */
int orc_frame_init(orc_frame* f, int width, int height) {
    memset(f, 0, sizeof(*f));
    if (width <= 0 || height <= 0) return -1;
    int mb_w = (width + 15) >> 4, mb_h = (height + 15) >> 4; /* video.go:314-315 */
    f->width = width;
    f->height = height;
    f->luma_w = mb_w << 4;
    f->luma_h = mb_h << 4;
    f->chroma_w = mb_w << 3;
    f->chroma_h = mb_h << 3;
    size_t luma = (size_t)f->luma_w * f->luma_h, chroma = (size_t)f->chroma_w * f->chroma_h;
    f->buf_bytes = luma + 2 * chroma + (size_t)f->luma_w * 16; /* video.go:340 */
    f->base = (uint8_t*)calloc(f->buf_bytes, 1);
    if (!f->base) return -1;
    f->y = f->base;
    f->cb = f->base + luma;
    f->cr = f->cb + chroma;
    return 0;
}

void orc_frame_free(orc_frame* f) {
    free(f->base);
    memset(f, 0, sizeof(*f));
}

/* ------------------------------------------------------------------------------------------
 * copyMacroblock.  Scalar form follows the reference's own test oracle (video_test.go:10-43),
 * which the portable path (video_noasm.go:28-80) and both assembler back-ends are verified
 * against.  A plane's readable extent is the rest of the shared allocation
 * (video_noasm.go:49-50, video.go:338-340).
 * ---------------------------------------------------------------------------------------- */
static int window_ok(const orc_frame* s, const uint8_t* plane, int64_t si, int stride, int size, int odd_h,
                     int odd_v) {
    int64_t lo = si;
    int64_t hi = si + (int64_t)(size - 1 + odd_v) * stride + (size - 1 + odd_h);
    int64_t avail = (int64_t)((s->base + s->buf_bytes) - plane);
    return lo >= 0 && hi < avail;
}

static void mc_plane_scalar(const uint8_t* src, uint8_t* dst, int stride, int size, int motion_h, int motion_v,
                            int mb_row, int mb_col) {
    int hp = motion_h >> 1, vp = motion_v >> 1; /* arithmetic shift: floor (video_test.go:12-13) */
    int odd_h = (motion_h & 1) == 1, odd_v = (motion_v & 1) == 1;
    for (int y = 0; y < size; y++) {
        for (int x = 0; x < size; x++) {
            int64_t si = ((int64_t)(mb_row * size) + vp + y) * stride + (mb_col * size) + hp + x;
            int64_t di = ((int64_t)(mb_row * size) + y) * stride + (mb_col * size) + x;
            int v;
            if (!odd_h && !odd_v)
                v = src[si];
            else if (odd_h && !odd_v)
                v = (src[si] + src[si + 1] + 1) >> 1;
            else if (!odd_h && odd_v)
                v = (src[si] + src[si + stride] + 1) >> 1;
            else
                v = (src[si] + src[si + 1] + src[si + stride] + src[si + stride + 1] + 2) >> 2;
            dst[di] = (uint8_t)v;
        }
    }
}

static int mc_check(int motion_h, int motion_v, int mb_row, int mb_col, const orc_frame* s) {
    int hp = motion_h >> 1, vp = motion_v >> 1;
    int64_t lsi = ((int64_t)(mb_row << 4) + vp) * s->luma_w + (mb_col << 4) + hp; /* video_noasm.go:31 */
    if (!window_ok(s, s->y, lsi, s->luma_w, 16, motion_h & 1, motion_v & 1)) return -1;
    int cm_h = motion_h / 2, cm_v = motion_v / 2; /* truncation toward zero, video_noasm.go:35-36 */
    hp = cm_h >> 1;
    vp = cm_v >> 1;
    int64_t csi = ((int64_t)(mb_row << 3) + vp) * s->chroma_w + (mb_col << 3) + hp;
    if (!window_ok(s, s->cb, csi, s->chroma_w, 8, cm_h & 1, cm_v & 1)) return -1;
    if (!window_ok(s, s->cr, csi, s->chroma_w, 8, cm_h & 1, cm_v & 1)) return -1;
    return 0;
}

int orc_copy_macroblock(int motion_h, int motion_v, int mb_row, int mb_col, const orc_frame* s, orc_frame* d) {
    if (mc_check(motion_h, motion_v, mb_row, mb_col, s) != 0) return -1;
    mc_plane_scalar(s->y, d->y, s->luma_w, 16, motion_h, motion_v, mb_row, mb_col);
    int cm_h = motion_h / 2, cm_v = motion_v / 2;
    mc_plane_scalar(s->cb, d->cb, s->chroma_w, 8, cm_h, cm_v, mb_row, mb_col);
    mc_plane_scalar(s->cr, d->cr, s->chroma_w, 8, cm_h, cm_v, mb_row, mb_col);
    return 0;
}

/* SWAR form, video_noasm.go:7-80: eight packed bytes per step. */
static inline uint64_t ld64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;
}
static inline void st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
/* video_noasm.go:15-17 */
static inline uint64_t round_avg(uint64_t a, uint64_t b) { return (a | b) - (((a ^ b) >> 1) & 0x7f7f7f7f7f7f7f7fULL); }
/* video_noasm.go:22-26 */
static inline uint64_t bilin_avg(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    const uint64_t lo_mask = 0x00ff00ff00ff00ffULL, two = 0x0002000200020002ULL;
    uint64_t lo = (((a & lo_mask) + (b & lo_mask) + (c & lo_mask) + (d & lo_mask) + two) >> 2) & lo_mask;
    uint64_t hi =
        ((((a >> 8) & lo_mask) + ((b >> 8) & lo_mask) + ((c >> 8) & lo_mask) + ((d >> 8) & lo_mask) + two) >> 2) &
        lo_mask;
    return lo | (hi << 8);
}

static void mc_block_swar(const uint8_t* src, uint8_t* dst, int stride, int64_t si, int64_t di, int size, int odd_h,
                          int odd_v) {
    for (int r = 0; r < size; r++) { /* video_noasm.go:52-79 */
        if (!odd_h && !odd_v) {
            memcpy(dst + di, src + si, (size_t)size);
        } else if (odd_h && !odd_v) {
            for (int x = 0; x < size; x += 8) st64(dst + di + x, round_avg(ld64(src + si + x), ld64(src + si + x + 1)));
        } else if (!odd_h && odd_v) {
            for (int x = 0; x < size; x += 8)
                st64(dst + di + x, round_avg(ld64(src + si + x), ld64(src + si + x + stride)));
        } else {
            for (int x = 0; x < size; x += 8)
                st64(dst + di + x, bilin_avg(ld64(src + si + x), ld64(src + si + x + 1), ld64(src + si + x + stride),
                                             ld64(src + si + x + stride + 1)));
        }
        si += stride;
        di += stride;
    }
}

int orc_copy_macroblock_swar(int motion_h, int motion_v, int mb_row, int mb_col, const orc_frame* s, orc_frame* d) {
    if (mc_check(motion_h, motion_v, mb_row, mb_col, s) != 0) return -1;
    /* the 8-byte loads of the SWAR form may touch up to 7 bytes beyond the scalar window; the
     * reference relies on the trailing pad (video.go:340) for that.  Guard the same way. */
    int hp = motion_h >> 1, vp = motion_v >> 1;
    int64_t lsi = ((int64_t)(mb_row << 4) + vp) * s->luma_w + (mb_col << 4) + hp;
    int64_t ldi = (int64_t)(mb_row << 4) * s->luma_w + (mb_col << 4);
    mc_block_swar(s->y, d->y, s->luma_w, lsi, ldi, 16, (motion_h & 1) == 1, (motion_v & 1) == 1);
    int cm_h = motion_h / 2, cm_v = motion_v / 2;
    hp = cm_h >> 1;
    vp = cm_v >> 1;
    int64_t csi = ((int64_t)(mb_row << 3) + vp) * s->chroma_w + (mb_col << 3) + hp;
    int64_t cdi = (int64_t)(mb_row << 3) * s->chroma_w + (mb_col << 3);
    mc_block_swar(s->cb, d->cb, s->chroma_w, csi, cdi, 8, (cm_h & 1) == 1, (cm_v & 1) == 1);
    mc_block_swar(s->cr, d->cr, s->chroma_w, csi, cdi, 8, (cm_h & 1) == 1, (cm_v & 1) == 1);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Frame.RGBA(), video.go:31-36 -> Go 1.23 image/draw.Draw(dst *image.RGBA, src *image.YCbCr,
 * op Src) -> image/internal/imageutil.DrawYCbCr, case YCbCrSubsampleRatio420.  That source is
 * not part of /root/reference; restated from the published standard library (the inlined
 * color.YCbCrToRGB): yy1 = Y*0x10101, cb1 = Cb-128, cr1 = Cr-128,
 *   r = (yy1 + 91881*cr1) >> 16, g = (yy1 - 22554*cb1 - 46802*cr1) >> 16, b = (yy1 + 116130*cb1) >> 16
 * each saturated to 0..255 (values with bits 24..31 set become ^(v>>31)), alpha 255, chroma
 * sample (x/2, y/2).  PARITY UNPINNED: no reference test asserts RGBA pixel values.
 * ---------------------------------------------------------------------------------------- */
static inline uint8_t sat_shift16(int32_t v) {
    if (((uint32_t)v & 0xff000000u) == 0) return (uint8_t)(v >> 16);
    return (uint8_t)(~(v >> 31));
}

void orc_rgba(const orc_frame* f, uint8_t* rgba) {
    for (int y = 0; y < f->height; y++) {
        uint8_t* dp = rgba + (size_t)y * 4 * f->width; /* Stride 4*width, video.go:369 */
        const uint8_t* yrow = f->y + (size_t)y * f->luma_w;
        const uint8_t* cbrow = f->cb + (size_t)(y / 2) * f->chroma_w;
        const uint8_t* crrow = f->cr + (size_t)(y / 2) * f->chroma_w;
        for (int x = 0; x < f->width; x++) {
            int32_t yy1 = (int32_t)yrow[x] * 0x10101;
            int32_t cb1 = (int32_t)cbrow[x / 2] - 128;
            int32_t cr1 = (int32_t)crrow[x / 2] - 128;
            dp[4 * x + 0] = sat_shift16(yy1 + 91881 * cr1);
            dp[4 * x + 1] = sat_shift16(yy1 - 22554 * cb1 - 46802 * cr1);
            dp[4 * x + 2] = sat_shift16(yy1 + 116130 * cb1);
            dp[4 * x + 3] = 255;
        }
    }
}

/* Frame.RGBA() of many frames: frame i = frames[3 * streams[i] + bufs[i]] -> out + i * stride (one frame per task) */
void orc_rgba_batch(const orc_frame* frames, int n, const int32_t* streams, const uint8_t* bufs, uint8_t* out, size_t stride,
                    int threads) {
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < n; i++) orc_rgba(&frames[(size_t)streams[i] * 3 + bufs[i]], out + (size_t)i * stride);
    (void)threads;
}

/* ------------------------------------------------------------------------------------------
 * Record-level executor: what decodeMacroblock + decodeBlock do once the bitstream has been
 * parsed (video.go:544-561, 747-798), driven by the packed records of include/mpegb200.h.
 * ---------------------------------------------------------------------------------------- */
/* videoPremultiplierMatrix, video.go:1077-1086.  It is the outer product of
 * {32,44,42,38,32,25,17,9}/32-scaled cosine factors; kept as data. */
static const uint8_t k_premultiplier[64] = {
    32, 44, 42, 38, 32, 25, 17, 9,  44, 62, 58, 52, 44, 35, 24, 12, 42, 58, 55, 49, 42, 33,
    23, 12, 38, 52, 49, 44, 38, 30, 20, 10, 32, 44, 42, 38, 32, 25, 17, 9,  25, 35, 33, 30,
    25, 20, 14, 7,  17, 24, 23, 20, 17, 14, 9,  5,  9,  12, 12, 10, 9,  7,  5,  2,
};
const uint8_t* orc_premultiplier(void) { return k_premultiplier; }

/* Which restatement of copyMacroblock the record executor uses: 0 = per pixel (default), 1 = 8 bytes per step like the
 * reference's portable Go path (video_noasm.go:44-80).  Same bytes (tests/test_oracle_golden.py sweeps both); bench.py's
 * CPU arm takes the faster one. */
static int g_swar_mc = 0;
void orc_use_swar_mc(int on) { g_swar_mc = on != 0; }

int orc_exec_macroblock(const mpegb200_mb* mb, const int16_t* coeffs, orc_frame* dst, const orc_frame* fwd,
                        const orc_frame* bwd) {
    int rc = 0;
    if (mb->flags & MPEGB200_MB_PREDICT) { /* video.go:544 -> predictMacroblock -> copyMacroblock */
        const orc_frame* ref = (mb->flags & MPEGB200_MB_REF_BWD) ? bwd : fwd;
        if ((g_swar_mc ? orc_copy_macroblock_swar : orc_copy_macroblock)(mb->mv_h, mb->mv_v, mb->mb_row, mb->mb_col, ref, dst) != 0) rc = -1;
    }
    const int16_t* blk = coeffs + (size_t)mb->coeff_block * 64;
    const int intra = (mb->flags & MPEGB200_MB_INTRA) != 0;
    for (int block = 0; block < 6; block++) { /* video.go:555-561 */
        if (!(mb->cbp & (0x20 >> block))) continue;
        int64_t data[64];
        int last = 0; /* highest natural index holding a non-zero coefficient */
        for (int i = 0; i < 64; i++) {
            data[i] = (int64_t)blk[i] * k_premultiplier[i]; /* video.go:744 (and :672 for intra DC) */
            if (blk[i] != 0) last = i;
        }
        blk += 64;
        /* destination, video.go:752-770 */
        uint8_t* d;
        int64_t di, scan;
        if (block < 4) {
            d = dst->y;
            di = ((int64_t)mb->mb_row * dst->luma_w + mb->mb_col) << 4;
            scan = dst->luma_w - 8;
            if (block & 1) di += 8;
            if (block & 2) di += (int64_t)dst->luma_w << 3;
        } else {
            d = block == 4 ? dst->cb : dst->cr;
            di = (((int64_t)mb->mb_row * dst->luma_w) << 2) + (mb->mb_col << 3);
            scan = (dst->luma_w >> 1) - 8;
        }
        /* video.go:772-798.  The DC shortcut (n == 1) and the full transform agree on a block
         * whose only non-zero coefficient is DC, so either may be used; take the shortcut when
         * it applies, like the reference does for n == 1. */
        if (last == 0) {
            int64_t value = (data[0] + 128) >> 8;
            if (intra)
                orc_copy_value_to_dest(value, d, di, scan);
            else
                orc_add_value_to_dest(value, d, di, scan);
        } else {
            orc_idct_full(data);
            if (intra)
                orc_copy_block_to_dest(data, d, di, scan);
            else
                orc_add_block_to_dest(data, d, di, scan);
        }
    }
    return rc;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_exec_pictures(orc_frame* frames, int n_pictures, const mpegb200_picture* pics, size_t n_mb,
                      const mpegb200_mb* mbs, const int16_t* coeffs, int threads) {
    (void)n_mb;
    int rc = 0;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(| : rc)
#endif
    for (int p = 0; p < n_pictures; p++) {
        const mpegb200_picture* pic = &pics[p];
        orc_frame* base = frames + (size_t)pic->stream * 3;
        orc_frame* dst = base + pic->dst_buf;
        const orc_frame* fwd = base + pic->fwd_buf;
        const orc_frame* bwd = base + pic->bwd_buf;
        for (uint32_t i = 0; i < pic->n_mb; i++) {
            const mpegb200_mb* mb = &mbs[pic->first_mb + i];
            if (orc_exec_macroblock(mb, coeffs, dst, fwd, bwd) != 0) rc |= 1;
        }
    }
    return rc ? -1 : 0;
}

void orc_free(void* p) { free(p); }
