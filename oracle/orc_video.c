/*
 * orc_video.c -- CPU restatement of the reference's MPEG-1 video decoder (TEST INFRASTRUCTURE).
 * Follows video.go (parse, dequantisation, frame rotation, display order) on top of the
 * pixel kernels in orc_pixel.c.  See mpeg_oracle.h for the rules and the pinning vectors.
 */
#include "mpeg_oracle.h"
#include "orc_bits.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Variable-length codes.  The rows in vlc_codes.inc are ISO/IEC 11172-2 Annex B as printed
 * (codeword, value); a binary decoding tree is grown from them once.  Decoding walks one bit
 * at a time exactly like Buffer.readVlc (buffer.go:352-376): a leaf returns its value, an
 * unassigned prefix returns 0 after consuming the bit that identified it (the reference's
 * {-1, 0} nodes, e.g. video.go:1103).
 * ---------------------------------------------------------------------------------------- */
#define VLC_INVALID (-32768)
typedef struct vlc_code {
    const char* bits;
    int value;
} vlc_code;
#include "vlc_codes.inc"

typedef struct vlc_node {
    int16_t child[2]; /* 0 = not grown */
    int32_t value;
    uint8_t leaf; /* 1 = symbol, 2 = unassigned prefix */
} vlc_node;

typedef struct vlc_tree {
    vlc_node* nodes;
    int count;
} vlc_tree;

static void vlc_grow(vlc_tree* t, const vlc_code* rows, size_t n_rows) {
    size_t cap = 2;
    for (size_t i = 0; i < n_rows; i++) cap += strlen(rows[i].bits);
    t->nodes = (vlc_node*)calloc(cap, sizeof(vlc_node));
    t->count = 1;
    for (size_t i = 0; i < n_rows; i++) {
        int at = 0;
        for (const char* c = rows[i].bits; *c; c++) {
            int bit = *c == '1';
            if (!t->nodes[at].child[bit]) t->nodes[at].child[bit] = (int16_t)t->count++;
            at = t->nodes[at].child[bit];
        }
        if (rows[i].value == VLC_INVALID) {
            t->nodes[at].leaf = 2;
            t->nodes[at].value = 0;
        } else {
            t->nodes[at].leaf = 1;
            t->nodes[at].value = rows[i].value;
        }
    }
}

static inline int vlc_read(orc_bits* b, const vlc_tree* t) {
    int at = 0;
    for (;;) {
        at = t->nodes[at].child[bits_read1(b)];
        if (at == 0) return 0; /* cannot happen with complete tables */
        if (t->nodes[at].leaf) return t->nodes[at].value;
    }
}

static vlc_tree T_ADDR_INC, T_TYPE_I, T_TYPE_P, T_TYPE_B, T_CBP, T_MOTION, T_DC_LUMA, T_DC_CHROMA, T_COEFF;
static int g_tables_ready = 0;
#define GROW(t, rows) vlc_grow(&(t), rows, sizeof(rows) / sizeof(rows[0]))
static void tables_init(void) {
    if (g_tables_ready) return;
    GROW(T_ADDR_INC, VLC_MB_ADDR_INC);
    GROW(T_TYPE_I, VLC_MB_TYPE_I);
    GROW(T_TYPE_P, VLC_MB_TYPE_P);
    GROW(T_TYPE_B, VLC_MB_TYPE_B);
    GROW(T_CBP, VLC_CBP);
    GROW(T_MOTION, VLC_MOTION);
    GROW(T_DC_LUMA, VLC_DC_SIZE_LUMA);
    GROW(T_DC_CHROMA, VLC_DC_SIZE_CHROMA);
    GROW(T_COEFF, VLC_DCT_COEFF);
    g_tables_ready = 1;
}

/* video.go:1034-1042 */
static const double k_picture_rate[16] = {0.000,  23.976, 24.000, 25.000, 29.970, 30.000, 50.000, 59.940,
                                          60.000, 0.000,  0.000,  0.000,  0.000,  0.000,  0.000,  0.000};

/* Zig-zag scan (video.go:1044-1053): generated, position n -> natural index. */
static uint8_t g_zigzag[64];
static void zigzag_init(void) {
    int r = 0, c = 0, up = 1;
    for (int n = 0; n < 64; n++) {
        g_zigzag[n] = (uint8_t)(r * 8 + c);
        if (up) {
            if (c == 7) { r++; up = 0; }
            else if (r == 0) { c++; up = 0; }
            else { r--; c++; }
        } else {
            if (r == 7) { c++; up = 1; }
            else if (c == 0) { r++; up = 1; }
            else { r++; c--; }
        }
    }
}

/* Default quantiser matrices (ISO 11172-2 2.4.3.2; video.go:1055-1075). */
static const uint8_t k_intra_quant[64] = {
    8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34,
    34, 38, 22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32,
    35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83,
};
const uint8_t* orc_premultiplier(void); /* orc_pixel.c */

enum { PIC_I = 1, PIC_P = 2, PIC_B = 3 };                                          /* video.go:931-933 */
enum { START_PICTURE = 0x00, START_SLICE_FIRST = 0x01, START_SLICE_LAST = 0xAF,    /* video.go:935-940 */
       START_USER_DATA = 0xB2, START_SEQUENCE = 0xB3, START_EXTENSION = 0xB5 };

typedef struct motion {
    int full_px, r_size, h, v, is_set; /* video.go:1026-1032 */
} motion;

struct orc_video {
    uint8_t* data;
    orc_bits buf;

    double frame_rate, time;
    int frames_decoded;
    int width, height, mb_width, mb_height, mb_size;
    int luma_width, luma_height, chroma_width, chroma_height;
    int start_code, picture_type;
    motion motion_forward, motion_backward;
    int has_sequence_header;
    int quantizer_scale, slice_begin, macroblock_address;
    int mb_row, mb_col, macroblock_type, macroblock_intra;
    int64_t dc_predictor[3];
    orc_frame frame_current, frame_forward, frame_backward; /* struct copies rotate, video.go:406-433 */
    orc_frame storage[3];
    int64_t block_data[64]; /* video.go:101 */
    int32_t level_data[64]; /* the same coefficients before the premultiply (tap output) */
    uint8_t intra_quant[64], non_intra_quant[64];
    int has_reference_frame, assume_no_b_frames;
    const orc_frame* last;
    int oob_count;

    /* record tap */
    int tap_on;
    mpegb200_picture* tap_pics;
    int tap_n_pics, tap_cap_pics;
    mpegb200_mb* tap_mbs;
    size_t tap_n_mbs, tap_cap_mbs;
    int16_t* tap_coeffs;
    size_t tap_n_blocks, tap_cap_blocks;
    int tap_cur_mb; /* index of the record being filled, -1 if none */
};

static int frame_id(const orc_video* v, const orc_frame* f) {
    for (int i = 0; i < 3; i++)
        if (f->base == v->storage[i].base) return i;
    return -1;
}

/* ---- tap helpers -------------------------------------------------------------------------- */
static void tap_begin_picture(orc_video* v) {
    if (!v->tap_on) return;
    if (v->tap_n_pics == v->tap_cap_pics) {
        v->tap_cap_pics = v->tap_cap_pics ? v->tap_cap_pics * 2 : 4;
        v->tap_pics = (mpegb200_picture*)realloc(v->tap_pics, sizeof(mpegb200_picture) * (size_t)v->tap_cap_pics);
    }
    mpegb200_picture* p = &v->tap_pics[v->tap_n_pics++];
    memset(p, 0, sizeof(*p));
    p->stream = 0;
    p->type = (uint8_t)v->picture_type;
    p->dst_buf = (uint8_t)frame_id(v, &v->frame_current);
    p->fwd_buf = (uint8_t)frame_id(v, &v->frame_forward);
    p->bwd_buf = (uint8_t)frame_id(v, &v->frame_backward);
    p->first_mb = (uint32_t)v->tap_n_mbs;
    p->n_mb = 0;
}

static mpegb200_mb* tap_new_mb(orc_video* v) {
    if (!v->tap_on || v->tap_n_pics == 0) {
        v->tap_cur_mb = -1;
        return NULL;
    }
    if (v->tap_n_mbs == v->tap_cap_mbs) {
        v->tap_cap_mbs = v->tap_cap_mbs ? v->tap_cap_mbs * 2 : 1024;
        v->tap_mbs = (mpegb200_mb*)realloc(v->tap_mbs, sizeof(mpegb200_mb) * v->tap_cap_mbs);
    }
    mpegb200_mb* m = &v->tap_mbs[v->tap_n_mbs];
    memset(m, 0, sizeof(*m));
    m->mb_row = (uint16_t)v->mb_row;
    m->mb_col = (uint16_t)v->mb_col;
    m->pic = (uint16_t)(v->tap_n_pics - 1);
    m->coeff_block = (uint32_t)v->tap_n_blocks;
    v->tap_cur_mb = (int)v->tap_n_mbs;
    v->tap_n_mbs++;
    v->tap_pics[v->tap_n_pics - 1].n_mb++;
    return m;
}

static void tap_block(orc_video* v, int block, const int32_t* levels) {
    if (!v->tap_on || v->tap_cur_mb < 0) return;
    if (v->tap_n_blocks == v->tap_cap_blocks) {
        v->tap_cap_blocks = v->tap_cap_blocks ? v->tap_cap_blocks * 2 : 4096;
        v->tap_coeffs = (int16_t*)realloc(v->tap_coeffs, sizeof(int16_t) * 64 * v->tap_cap_blocks);
    }
    int16_t* dst = v->tap_coeffs + 64 * v->tap_n_blocks++;
    for (int i = 0; i < 64; i++) {
        int32_t l = levels[i];
        if (l > 32767) l = 32767; /* only reachable through a drifting DC predictor in corrupt streams */
        if (l < -32768) l = -32768;
        dst[i] = (int16_t)l;
    }
    v->tap_mbs[v->tap_cur_mb].cbp |= (uint8_t)(0x20 >> block);
}

/* ---- sequence header, video.go:270-331 ------------------------------------------------------ */
static int decode_sequence_header(orc_video* v) {
    int max_header_size = 64 + 2 * 64 * 8;
    if (!bits_has(&v->buf, max_header_size)) return 0;
    v->width = (int)bits_read(&v->buf, 12);
    v->height = (int)bits_read(&v->buf, 12);
    if (v->width <= 0 || v->height <= 0) return 0;
    bits_read(&v->buf, 4); /* aspect ratio index: not needed by the hot path */
    v->frame_rate = k_picture_rate[bits_read(&v->buf, 4)];
    bits_read(&v->buf, 18); /* bit rate */
    bits_skip(&v->buf, 1 + 10 + 1);
    if (bits_read1(&v->buf) != 0) {
        for (int i = 0; i < 64; i++) v->intra_quant[g_zigzag[i]] = (uint8_t)bits_read(&v->buf, 8);
    } else {
        memcpy(v->intra_quant, k_intra_quant, 64);
    }
    if (bits_read1(&v->buf) != 0) {
        for (int i = 0; i < 64; i++) v->non_intra_quant[g_zigzag[i]] = (uint8_t)bits_read(&v->buf, 8);
    } else {
        memset(v->non_intra_quant, 16, 64); /* video.go:1066-1075: flat 16 */
    }
    v->mb_width = (v->width + 15) >> 4;
    v->mb_height = (v->height + 15) >> 4;
    v->mb_size = v->mb_width * v->mb_height;
    v->luma_width = v->mb_width << 4;
    v->luma_height = v->mb_height << 4;
    v->chroma_width = v->mb_width << 3;
    v->chroma_height = v->mb_height << 3;
    for (int i = 0; i < 3; i++) { /* initFrame x3, video.go:324-326 */
        orc_frame_free(&v->storage[i]);
        if (orc_frame_init(&v->storage[i], v->width, v->height) != 0) return 0;
    }
    v->frame_current = v->storage[0];
    v->frame_forward = v->storage[1];
    v->frame_backward = v->storage[2];
    v->has_sequence_header = 1;
    return 1;
}

/* video.go:130-147 */
int orc_video_has_header(orc_video* v) {
    if (v->has_sequence_header) return 1;
    if (v->start_code != START_SEQUENCE) v->start_code = bits_find_start_code(&v->buf, START_SEQUENCE);
    if (v->start_code == -1) return 0;
    return decode_sequence_header(v);
}

/* ---- motion vectors, video.go:564-606 ---------------------------------------------------------- */
static int decode_motion_vector(orc_video* v, int r_size, int mot) {
    int fscale = 1 << r_size;
    int m_code = vlc_read(&v->buf, &T_MOTION);
    int r, d;
    if (m_code != 0 && fscale != 1) {
        r = (int)bits_read(&v->buf, r_size);
        d = (((m_code < 0 ? -m_code : m_code) - 1) << r_size) + r + 1;
        if (m_code < 0) d = -d;
    } else {
        d = m_code;
    }
    mot += d;
    if (mot > (fscale << 4) - 1)
        mot -= fscale << 5;
    else if (mot < ((-fscale) * 16))
        mot += fscale << 5;
    return mot;
}

static void decode_motion_vectors(orc_video* v) {
    if (v->motion_forward.is_set) {
        int r = v->motion_forward.r_size;
        v->motion_forward.h = decode_motion_vector(v, r, v->motion_forward.h);
        v->motion_forward.v = decode_motion_vector(v, r, v->motion_forward.v);
    } else if (v->picture_type == PIC_P) {
        v->motion_forward.h = 0;
        v->motion_forward.v = 0;
    }
    if (v->motion_backward.is_set) {
        int r = v->motion_backward.r_size;
        v->motion_backward.h = decode_motion_vector(v, r, v->motion_backward.h);
        v->motion_backward.v = decode_motion_vector(v, r, v->motion_backward.v);
    }
}

/* ---- predictMacroblock, video.go:608-637 -------------------------------------------------------- */
static void predict_macroblock(orc_video* v, mpegb200_mb* rec) {
    int fw_h = v->motion_forward.h, fw_v = v->motion_forward.v;
    if (v->motion_forward.full_px) {
        fw_h *= 2;
        fw_v *= 2;
    }
    int use_bwd = 0, mv_h = fw_h, mv_v = fw_v;
    if (v->picture_type == PIC_B) {
        int bw_h = v->motion_backward.h, bw_v = v->motion_backward.v;
        if (v->motion_backward.full_px) {
            bw_h *= 2;
            bw_v *= 2;
        }
        if (v->motion_forward.is_set) {
            if (orc_copy_macroblock(fw_h, fw_v, v->mb_row, v->mb_col, &v->frame_forward, &v->frame_current) != 0)
                v->oob_count++;
            if (v->motion_backward.is_set) {
                /* the backward copy simply overwrites the forward one: no averaging (video.go:626-630) */
                if (orc_copy_macroblock(bw_h, bw_v, v->mb_row, v->mb_col, &v->frame_backward, &v->frame_current) != 0)
                    v->oob_count++;
                use_bwd = 1;
            }
        } else {
            if (orc_copy_macroblock(bw_h, bw_v, v->mb_row, v->mb_col, &v->frame_backward, &v->frame_current) != 0)
                v->oob_count++;
            use_bwd = 1;
        }
        if (use_bwd) {
            mv_h = bw_h;
            mv_v = bw_v;
        }
    } else {
        if (orc_copy_macroblock(fw_h, fw_v, v->mb_row, v->mb_col, &v->frame_forward, &v->frame_current) != 0)
            v->oob_count++;
    }
    if (rec) {
        rec->flags |= MPEGB200_MB_PREDICT | (use_bwd ? MPEGB200_MB_REF_BWD : 0);
        rec->mv_h = (int16_t)mv_h;
        rec->mv_v = (int16_t)mv_v;
    }
}

/* ---- decodeBlock, video.go:639-799 ---------------------------------------------------------------- */
static void decode_block(orc_video* v, int block) {
    int n = 0;
    const uint8_t* quant_matrix;
    const uint8_t* premult = orc_premultiplier();

    if (v->macroblock_intra) {
        int plane_index = block > 3 ? block - 3 : 0;
        int64_t predictor = v->dc_predictor[plane_index];
        int dct_size = vlc_read(&v->buf, plane_index == 0 ? &T_DC_LUMA : &T_DC_CHROMA);
        if (dct_size > 0) {
            int64_t differential = bits_read(&v->buf, dct_size);
            if ((differential & ((int64_t)1 << (dct_size - 1))) != 0)
                v->block_data[0] = predictor + differential;
            else
                v->block_data[0] = predictor + ((-((int64_t)1 << dct_size)) | (differential + 1));
        } else {
            v->block_data[0] = predictor;
        }
        v->dc_predictor[plane_index] = v->block_data[0];
        v->level_data[0] = (int32_t)(v->block_data[0] * 8); /* level form: dc*8, times premult 32 == dc<<8 */
        v->block_data[0] *= 256;                            /* <<= 3+5, video.go:672 (multiply: also right for negatives) */
        quant_matrix = v->intra_quant;
        n = 1;
    } else {
        quant_matrix = v->non_intra_quant;
    }

    int64_t level = 0;
    for (;;) {
        int run = 0;
        int coeff = vlc_read(&v->buf, &T_COEFF);
        if (coeff == 0x0001 && n > 0 && bits_read1(&v->buf) == 0) break; /* end_of_block */
        if (coeff == 0xffff) {                                           /* escape */
            run = (int)bits_read(&v->buf, 6);
            level = bits_read(&v->buf, 8);
            if (level == 0)
                level = bits_read(&v->buf, 8);
            else if (level == 128)
                level = bits_read(&v->buf, 8) - 256;
            else if (level > 128)
                level -= 256;
        } else {
            run = coeff >> 8;
            level = coeff & 0xff;
            if (bits_read1(&v->buf) != 0) level = -level;
        }
        n += run;
        if (n < 0 || n >= 64) return; /* invalid: block_data is left dirty, video.go:712-714 */
        int dz = g_zigzag[n];
        n++;
        /* dequantise, oddify, clip: video.go:719-741 */
        level *= 2;
        if (!v->macroblock_intra) level += level < 0 ? -1 : 1;
        level = (level * v->quantizer_scale * (int64_t)quant_matrix[dz]) >> 4;
        if ((level & 1) == 0) level -= level > 0 ? 1 : -1;
        if (level > 2047)
            level = 2047;
        else if (level < -2048)
            level = -2048;
        v->level_data[dz] = (int32_t)level;
        v->block_data[dz] = level * premult[dz]; /* video.go:744 */
    }

    /* destination, video.go:747-770 */
    uint8_t* d;
    int64_t di, scan;
    if (block < 4) {
        d = v->frame_current.y;
        di = ((int64_t)v->mb_row * v->luma_width + v->mb_col) << 4;
        scan = v->luma_width - 8;
        if (block & 1) di += 8;
        if (block & 2) di += (int64_t)v->luma_width << 3;
    } else {
        d = block == 4 ? v->frame_current.cb : v->frame_current.cr;
        di = (((int64_t)v->mb_row * v->luma_width) << 2) + (v->mb_col << 3);
        scan = (v->luma_width >> 1) - 8;
    }

    /* tap: emit the coefficients the reference is about to transform.  n == 1 uses DC only and
     * leaves the rest of block_data alone; n < 10 ignores everything outside rows/cols 0..3. */
    if (v->tap_on) {
        int32_t out[64];
        if (n == 1) {
            memset(out, 0, sizeof(out));
            out[0] = v->level_data[0];
        } else if (n < 10) {
            for (int i = 0; i < 64; i++) out[i] = ((i >> 3) < 4 && (i & 7) < 4) ? v->level_data[i] : 0;
        } else {
            memcpy(out, v->level_data, sizeof(out));
        }
        tap_block(v, block, out);
    }

    if (n == 1) { /* video.go:774-777, 787-790 */
        int64_t value = (v->block_data[0] + 128) >> 8;
        if (v->macroblock_intra)
            orc_copy_value_to_dest(value, d, di, scan);
        else
            orc_add_value_to_dest(value, d, di, scan);
        v->block_data[0] = 0;
        v->level_data[0] = 0;
    } else {
        orc_idct(v->block_data, n);
        if (v->macroblock_intra)
            orc_copy_block_to_dest(v->block_data, d, di, scan);
        else
            orc_add_block_to_dest(v->block_data, d, di, scan);
        memset(v->block_data, 0, sizeof(v->block_data));
        memset(v->level_data, 0, sizeof(v->level_data));
    }
}

/* ---- decodeMacroblock, video.go:462-562 ------------------------------------------------------------ */
static void decode_macroblock(orc_video* v) {
    int increment = 0;
    int t = vlc_read(&v->buf, &T_ADDR_INC);
    while (t == 34) t = vlc_read(&v->buf, &T_ADDR_INC); /* macroblock_stuffing */
    while (t == 35) {                                    /* macroblock_escape */
        increment += 33;
        t = vlc_read(&v->buf, &T_ADDR_INC);
    }
    increment += t;

    if (v->slice_begin) {
        v->slice_begin = 0;
        v->macroblock_address += increment;
    } else {
        if (v->macroblock_address + increment >= v->mb_size) return; /* invalid */
        if (increment > 1) {
            v->dc_predictor[0] = v->dc_predictor[1] = v->dc_predictor[2] = 128;
            if (v->picture_type == PIC_P) {
                v->motion_forward.h = 0;
                v->motion_forward.v = 0;
            }
        }
        while (increment > 1) { /* skipped macroblocks are predicted, video.go:503-510 */
            v->macroblock_address++;
            v->mb_row = v->macroblock_address / v->mb_width;
            v->mb_col = v->macroblock_address % v->mb_width;
            predict_macroblock(v, tap_new_mb(v));
            increment--;
        }
        v->macroblock_address++;
    }

    v->mb_row = v->macroblock_address / v->mb_width;
    v->mb_col = v->macroblock_address % v->mb_width;
    if (v->macroblock_address < 0 || v->mb_col >= v->mb_width || v->mb_row >= v->mb_height) return; /* corrupt stream (address -1: slice 1 + increment code of value 0; the Go code panics on the index) */

    const vlc_tree* type_tree = v->picture_type == PIC_I ? &T_TYPE_I : (v->picture_type == PIC_P ? &T_TYPE_P : &T_TYPE_B);
    v->macroblock_type = vlc_read(&v->buf, type_tree);
    v->macroblock_intra = (v->macroblock_type & 0x01) != 0;
    v->motion_forward.is_set = (v->macroblock_type & 0x08) != 0;
    v->motion_backward.is_set = (v->macroblock_type & 0x04) != 0;
    if ((v->macroblock_type & 0x10) != 0) v->quantizer_scale = (int)bits_read(&v->buf, 5);

    mpegb200_mb* rec = tap_new_mb(v);
    if (v->macroblock_intra) {
        v->motion_backward.h = v->motion_forward.h = 0;
        v->motion_backward.v = v->motion_forward.v = 0;
        if (rec) rec->flags |= MPEGB200_MB_INTRA;
    } else {
        v->dc_predictor[0] = v->dc_predictor[1] = v->dc_predictor[2] = 128;
        decode_motion_vectors(v);
        predict_macroblock(v, rec);
    }

    int cbp = 0;
    if ((v->macroblock_type & 0x02) != 0)
        cbp = vlc_read(&v->buf, &T_CBP);
    else if (v->macroblock_intra)
        cbp = 0x3f;
    for (int block = 0, mask = 0x20; block < 6; block++, mask >>= 1)
        if (cbp & mask) decode_block(v, block);
}

/* ---- decodeSlice, video.go:436-460 ------------------------------------------------------------------ */
static void decode_slice(orc_video* v, int slice) {
    v->slice_begin = 1;
    v->macroblock_address = (slice - 1) * v->mb_width - 1;
    v->motion_backward.h = v->motion_forward.h = 0;
    v->motion_backward.v = v->motion_forward.v = 0;
    v->dc_predictor[0] = v->dc_predictor[1] = v->dc_predictor[2] = 128;
    v->quantizer_scale = (int)bits_read(&v->buf, 5);
    while (bits_read1(&v->buf) != 0) bits_skip(&v->buf, 8); /* extra information */
    do {
        decode_macroblock(v);
    } while (v->macroblock_address < v->mb_size - 1 && bits_peek_non_zero(&v->buf, 23));
}

/* ---- decodePicture, video.go:374-434 ------------------------------------------------------------------ */
static void decode_picture(orc_video* v) {
    bits_skip(&v->buf, 10); /* temporal reference */
    v->picture_type = (int)bits_read(&v->buf, 3);
    bits_skip(&v->buf, 16); /* vbv_delay */
    if (v->picture_type <= 0 || v->picture_type > PIC_B) return; /* D pictures / unknown */
    if (v->picture_type == PIC_P || v->picture_type == PIC_B) {
        v->motion_forward.full_px = bits_read1(&v->buf);
        int f_code = (int)bits_read(&v->buf, 3);
        if (f_code == 0) return;
        v->motion_forward.r_size = f_code - 1;
    }
    if (v->picture_type == PIC_B) {
        v->motion_backward.full_px = bits_read1(&v->buf);
        int f_code = (int)bits_read(&v->buf, 3);
        if (f_code == 0) return;
        v->motion_backward.r_size = f_code - 1;
    }

    orc_frame frame_temp = v->frame_forward;
    if (v->picture_type == PIC_I || v->picture_type == PIC_P) v->frame_forward = v->frame_backward;

    tap_begin_picture(v);

    do { /* first slice start code; skip extension and user data */
        v->start_code = bits_next_start_code(&v->buf);
    } while (v->start_code == START_EXTENSION || v->start_code == START_USER_DATA);

    while (v->start_code >= START_SLICE_FIRST && v->start_code <= START_SLICE_LAST) {
        decode_slice(v, v->start_code & 0xFF);
        if (v->macroblock_address >= v->mb_size - 2) break;
        v->start_code = bits_next_start_code(&v->buf);
    }

    if (v->picture_type == PIC_I || v->picture_type == PIC_P) { /* rotate, video.go:430-433 */
        v->frame_backward = v->frame_current;
        v->frame_current = frame_temp;
    }
}

/* ---- Decode, video.go:209-268 ---------------------------------------------------------------------------- */
const orc_frame* orc_video_decode(orc_video* v) {
    v->tap_n_pics = 0;
    v->tap_n_mbs = 0;
    v->tap_n_blocks = 0;
    v->tap_cur_mb = -1;
    if (!orc_video_has_header(v)) return NULL;
    orc_frame* frame = NULL;
    for (;;) {
        if (v->start_code != START_PICTURE) {
            v->start_code = bits_find_start_code(&v->buf, START_PICTURE);
            if (v->start_code == -1) {
                if (v->has_reference_frame && !v->assume_no_b_frames && v->buf.has_ended &&
                    (v->picture_type == PIC_I || v->picture_type == PIC_P)) {
                    v->has_reference_frame = 0;
                    frame = &v->frame_backward;
                    break;
                }
                return NULL;
            }
        }
        if (bits_has_start_code(&v->buf, START_PICTURE) == -1 && !v->buf.has_ended) return NULL;
        decode_picture(v);
        if (v->assume_no_b_frames)
            frame = &v->frame_backward;
        else if (v->picture_type == PIC_B)
            frame = &v->frame_current;
        else if (v->has_reference_frame)
            frame = &v->frame_forward;
        else
            v->has_reference_frame = 1;
        if (frame) break;
    }
    frame->time = v->time;
    v->frames_decoded++;
    v->time = (double)v->frames_decoded / v->frame_rate;
    v->last = frame;
    return frame;
}

int orc_video_last_buf(orc_video* v) { return v->last ? frame_id(v, v->last) : -1; }

/* NewVideo, video.go:110-121 */
orc_video* orc_video_open(const uint8_t* data, size_t len) {
    if (!g_tables_ready) {
        zigzag_init();
        tables_init();
    }
    orc_video* v = (orc_video*)calloc(1, sizeof(orc_video));
    if (!v) return NULL;
    v->data = (uint8_t*)malloc(len + 8);
    memcpy(v->data, data, len);
    memset(v->data + len, 0, 8);
    v->buf.bytes = v->data;
    v->buf.len = len;
    v->tap_cur_mb = -1;
    v->start_code = bits_find_start_code(&v->buf, START_SEQUENCE);
    if (v->start_code != -1) decode_sequence_header(v);
    return v;
}

void orc_video_close(orc_video* v) {
    if (!v) return;
    for (int i = 0; i < 3; i++) orc_frame_free(&v->storage[i]);
    free(v->tap_pics);
    free(v->tap_mbs);
    free(v->tap_coeffs);
    free(v->data);
    free(v);
}

int orc_video_width(orc_video* v) { return orc_video_has_header(v) ? v->width : 0; }
int orc_video_height(orc_video* v) { return orc_video_has_header(v) ? v->height : 0; }
double orc_video_framerate(orc_video* v) { return orc_video_has_header(v) ? v->frame_rate : 0; }
void orc_video_set_no_delay(orc_video* v, int no_delay) { v->assume_no_b_frames = no_delay; }

/* video.go:195-201 (+ Buffer.seek(0) for a reader-less buffer, buffer.go:158-176) */
void orc_video_rewind(orc_video* v) {
    v->buf.bit_index = 0;
    v->buf.has_ended = 0;
    v->time = 0;
    v->frames_decoded = 0;
    v->has_reference_frame = 0;
    v->start_code = -1;
}

void orc_video_tap_enable(orc_video* v, int on) { v->tap_on = on; }
int orc_video_tap_pictures(orc_video* v, const mpegb200_picture** pics) {
    *pics = v->tap_pics;
    return v->tap_n_pics;
}
size_t orc_video_tap_mbs(orc_video* v, const mpegb200_mb** mbs) {
    *mbs = v->tap_mbs;
    return v->tap_n_mbs;
}
size_t orc_video_tap_blocks(orc_video* v, const int16_t** coeffs) {
    *coeffs = v->tap_coeffs;
    return v->tap_n_blocks;
}
int orc_video_oob_count(orc_video* v) { return v->oob_count; }
