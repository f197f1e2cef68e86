/*
 * orc_audio.c -- CPU restatement of the reference's MP2 decoder (TEST INFRASTRUCTURE).
 * Follows audio.go (header, allocation, scale factors, requantisation, synthesis loop),
 * audio_noasm.go (window).  See mpeg_oracle.h for the rules and the pinning vectors.
 * Build with -ffp-contract=off: every float32 operation is rounded on its own, as on the
 * reference's amd64 non-FMA path (audio_test.go:7-8, audio_amd64.s:5-8).
 */
#include "mpeg_oracle.h"
#include "orc_bits.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * idct36, audio.go:492-772 ("matrixing": a 32-point DCT written to the V buffer).
 *
 * The reference spells out B.G. Lee's recursive fast DCT as 273 straight-line statements.
 * Restated here as the recursion itself; each level does
 *     e[i] = x[i] + x[n-1-i],   o[i] = (x[i] - x[n-1-i]) * C_n[i]           (audio.go:497-555 ...)
 *     E = DCT(e), O = DCT(o),  O[k] += O[k+1] (k ascending)                  (audio.go:562, 571-574, 600-610, 692-706)
 *     X[2k] = E[k], X[2k+1] = O[k]
 * with the same operand order, so every float32 rounding is the same.  The first level adds
 * and subtracts the integer samples before the conversion to float32 (audio.go:497-528).
 * C_n[i] = 1 / (2 cos((2i+1) pi / 2n)), given with the reference's 12 significant digits so that
 * the float32 roundings of the literals agree.
 * ---------------------------------------------------------------------------------------- */
static const float C32[16] = {0.500602998235f, 0.505470959898f, 0.515447309923f, 0.53104259109f,
                              0.553103896034f, 0.582934968206f, 0.622504123036f, 0.674808341455f,
                              0.744536271002f, 0.839349645416f, 0.972568237862f, 1.16943993343f,
                              1.48416461631f,  2.05778100995f,  3.40760841847f,  10.1900081235f};
static const float C16[8] = {0.502419286188f, 0.52249861494f,  0.566944034816f, 0.64682178336f,
                             0.788154623451f, 1.06067768599f, 1.72244709824f,  5.10114861869f};
static const float C8[4] = {0.509795579104f, 0.601344886935f, 0.899976223136f, 2.56291544774f};
static const float C4[2] = {0.541196100146f, 1.30656296488f};
static const float C2[1] = {0.707106781187f};

static const float* lee_coeffs(int n) {
    switch (n) {
        case 16: return C16;
        case 8: return C8;
        case 4: return C4;
        default: return C2;
    }
}

/* in place, natural-order output; n in {16, 8, 4, 2, 1} */
static void lee_dct(float* x, int n) {
    if (n == 1) return;
    int h = n / 2;
    const float* c = lee_coeffs(n);
    float e[8], o[8];
    for (int i = 0; i < h; i++) {
        float a = x[i], b = x[n - 1 - i];
        e[i] = a + b;
        o[i] = (a - b) * c[i];
    }
    lee_dct(e, h);
    lee_dct(o, h);
    for (int k = 0; k + 1 < h; k++) o[k] += o[k + 1];
    for (int k = 0; k < h; k++) {
        x[2 * k] = e[k];
        x[2 * k + 1] = o[k];
    }
}

/* X[0..31] of one time slot; samples s[sb] are the requantised integers */
static void dct32_from_ints(const int64_t* s, int stride, float X[32]) {
    float e[16], o[16];
    for (int i = 0; i < 16; i++) {
        int64_t a = s[i * stride], b = s[(31 - i) * stride];
        e[i] = (float)(a + b);               /* integer add first, audio.go:497 */
        o[i] = (float)(a - b) * C32[i];      /* audio.go:498 */
    }
    lee_dct(e, 16);
    lee_dct(o, 16);
    for (int k = 0; k < 15; k++) o[k] += o[k + 1]; /* audio.go:692-706 */
    for (int k = 0; k < 16; k++) {
        X[2 * k] = e[k];
        X[2 * k + 1] = o[k];
    }
}

/* Placement into V, audio.go:708-771: with X the DCT above,
 *   d[dp+j]      =  X[16+j]          j = 0..15
 *   d[dp+16]     =  0
 *   d[dp+32-j]   = -X[16+j]          j = 0..15  (=> d[dp+17..32])
 *   d[dp+48+-k]  = -X[k]             k = 0..15  (=> d[dp+33..63]) */
static void place_v(const float X[32], float* d, int dp) {
    for (int j = 0; j < 16; j++) {
        d[dp + j] = X[16 + j];
        d[dp + 32 - j] = -X[16 + j];
    }
    d[dp + 16] = 0.0f;
    for (int k = 0; k < 16; k++) {
        d[dp + 48 + k] = -X[k];
        d[dp + 48 - k] = -X[k];
    }
}

void orc_idct36(const int64_t s[32][3], int ss, float* d, int dp) {
    float X[32];
    dct32_from_ints(&s[0][ss], 3, X);
    place_v(X, d, dp);
}

/* ------------------------------------------------------------------------------------------
 * synthesisWindow, audio.go:812-899: the ISO 11172-3 Table 3-B.3 window D[i] scaled by 2^16 and
 * negated in alternate 64-blocks, as kjmp2/pl_mpeg store it.  Kept as data, run-length free.
 * ---------------------------------------------------------------------------------------- */
#include "synth_window.inc"

static float g_window1024[1024];
static int g_window_ready = 0;
const float* orc_synthesis_window_1024(void) {
    if (!g_window_ready) {
        for (int i = 0; i < 512; i++) { /* audio.go:95-98 */
            g_window1024[i] = (float)k_synthesis_window_x2[i] * 0.5f;
            g_window1024[i + 512] = g_window1024[i];
        }
        g_window_ready = 1;
    }
    return g_window1024;
}

/* synthWindow, audio_noasm.go:8-38: 8 + 8 taps of 32 lanes, accumulated in tap order */
void orc_synth_window(float u[32], const float d[1024], const float v[1024], int v_pos) {
    for (int i = 0; i < 32; i++) u[i] = 0;
    int d_index = 512 - (v_pos >> 1);
    int v_index = (v_pos % 128) >> 1;
    while (v_index < 1024) {
        for (int i = 0; i < 32; i++) u[i] += d[d_index + i] * v[v_index + i];
        v_index += 128;
        d_index += 64;
    }
    d_index -= 512 - 32;
    v_index = (128 - 32 + 1024) - v_index;
    while (v_index < 1024) {
        for (int i = 0; i < 32; i++) u[i] += d[d_index + i] * v[v_index + i];
        v_index += 128;
        d_index += 64;
    }
}

/* the fused variant of audio_amd64.s:107-156 (VFMADD231PS) / audio_arm64.s:36-85 (FMLA) */
void orc_synth_window_fma(float u[32], const float d[1024], const float v[1024], int v_pos) {
    for (int i = 0; i < 32; i++) u[i] = 0;
    int d_index = 512 - (v_pos >> 1);
    int v_index = (v_pos % 128) >> 1;
    while (v_index < 1024) {
        for (int i = 0; i < 32; i++) u[i] = __builtin_fmaf(d[d_index + i], v[v_index + i], u[i]);
        v_index += 128;
        d_index += 64;
    }
    d_index -= 512 - 32;
    v_index = (128 - 32 + 1024) - v_index;
    while (v_index < 1024) {
        for (int i = 0; i < 32; i++) u[i] = __builtin_fmaf(d[d_index + i], v[v_index + i], u[i]);
        v_index += 128;
        d_index += 64;
    }
}

/* Output formatting, audio.go:386-418.  float32 -> int16 conversions truncate toward zero; for
 * values outside int16 Go's result is implementation-specific, amd64 keeps the low 16 bits of
 * the int32 truncation, restated here. */
static inline void emit_sample(float uj, int format, void* out, int pos, int ch) {
    float s = uj / -1090519040.0f; /* a true float32 division */
    switch (format) {
        case MPEGB200_AUDIO_F32N:
            ((float*)out)[(pos << 1) + ch] = s;
            break;
        case MPEGB200_AUDIO_F32NLR:
            ((float*)out)[ch * MPEGB200_SAMPLES_PER_FRAME + pos] = s;
            break;
        case MPEGB200_AUDIO_S16:
            ((int16_t*)out)[(pos << 1) + ch] =
                (int16_t)(int32_t)(s < 0 ? s * 32768.0f /*0x8000*/ : s * 32767.0f /*0x7FFF*/);
            break;
        default: /* MPEGB200_AUDIO_F32: 0x80000000 and 0x7FFFFFFF both round to 2^31 as float32 */
            ((float*)out)[(pos << 1) + ch] = s * 2147483648.0f;
            break;
    }
}

/* Synthesis section of decodeFrame, audio.go:377-422, for a whole frame (36 sub-steps). */
void orc_synth_frame(orc_synth_state* st, const int32_t* samples, int format, void* out, int fma) {
    const float* dwin = orc_synthesis_window_1024();
    float u[32];
    int out_pos = 0;
    for (int step = 0; step < 36; step++) {
        st->v_pos = (st->v_pos - 64) & 1023; /* audio.go:380 */
        for (int ch = 0; ch < 2; ch++) {
            const int32_t* s = samples + (ch * 36 + step) * 32;
            int64_t s64[32];
            for (int i = 0; i < 32; i++) s64[i] = s[i];
            float X[32];
            dct32_from_ints(s64, 1, X);
            place_v(X, st->v[ch], st->v_pos);
            if (fma)
                orc_synth_window_fma(u, dwin, st->v[ch], st->v_pos);
            else
                orc_synth_window(u, dwin, st->v[ch], st->v_pos);
            for (int j = 0; j < 32; j++) emit_sample(u[j], format, out, out_pos + j, ch);
        }
        out_pos += 32;
    }
}

void orc_synth_batch(orc_synth_state* states, int n_streams, int frames_per_stream, const int32_t* samples,
                     int format, void* out, int fma, int threads) {
    size_t out_bytes = (format == MPEGB200_AUDIO_S16 ? 2 : 4) * (size_t)2 * MPEGB200_SAMPLES_PER_FRAME;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int s = 0; s < n_streams; s++) {
        for (int f = 0; f < frames_per_stream; f++) {
            size_t idx = (size_t)s * frames_per_stream + f;
            orc_synth_frame(&states[s], samples + idx * 2 * 36 * 32, format, (uint8_t*)out + idx * out_bytes, fma);
        }
    }
    (void)threads;
}

/* ------------------------------------------------------------------------------------------
 * Frame parse (audio.go:184-490).  Tables are ISO 11172-3 Annex B (3-B.2a..d) in the compact
 * form used by kjmp2 and the reference (audio.go:798-973).
 * ---------------------------------------------------------------------------------------- */
enum { FRAME_SYNC = 0x7ff, MPEG_1 = 3, LAYER_II = 2 };
enum { MODE_STEREO = 0, MODE_JOINT_STEREO = 1, MODE_DUAL = 2, MODE_MONO = 3 };

static const uint16_t k_samplerate[4] = {44100, 48000, 32000, 0};
static const int16_t k_bitrate[14] = {32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384};
static const int64_t k_scalefactor_base[3] = {0x02000000, 0x01965FEA, 0x01428A30};

/* bitrate class by (mono?0:1, bitrate index), audio.go:902-907 */
static const uint8_t k_rate_class[2][14] = {
    {0, 0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2},
    {0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2},
};
/* (table id << 6) | sblimit by (class, samplerate index), audio.go:910-920 */
static const uint8_t k_table_pick[3][3] = {
    {8, 8, 12},
    {27 | 64, 27 | 64, 27 | 64},
    {30 | 64, 27 | 64, 30 | 64},
};
/* per subband: (nbal << 4) | row of k_alloc_rows, audio.go:923-943 (the two MPEG-1 tables) */
static const uint8_t k_sb_alloc[2][30] = {
    {0x44, 0x44, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34},
    {0x43, 0x43, 0x43, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x31, 0x31, 0x31, 0x31,
     0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20},
};
/* allocation code -> quantiser number (0 = no bits), audio.go:946-953 */
static const uint8_t k_alloc_rows[6][16] = {
    {0, 1, 2, 17},
    {0, 1, 2, 3, 4, 5, 6, 17},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 17},
    {0, 1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17},
    {0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
};
/* quantiser classes, audio.go:955-973: {levels, grouped, bits} */
typedef struct quantizer {
    uint16_t levels;
    uint8_t group, bits;
} quantizer;
static const quantizer k_quant[17] = {
    {3, 1, 5},     {5, 1, 7},     {7, 0, 3},      {9, 1, 10},     {15, 0, 4},     {31, 0, 5},
    {63, 0, 6},    {127, 0, 7},   {255, 0, 8},    {511, 0, 9},    {1023, 0, 10},  {2047, 0, 11},
    {4095, 0, 12}, {8191, 0, 13}, {16383, 0, 14}, {32767, 0, 15}, {65535, 0, 16},
};

struct orc_audio {
    uint8_t* data;
    orc_bits buf;
    double time;
    int samples_decoded, samplerate_index, bitrate_index, version, layer, mode, channels, bound;
    int next_frame_data_size, has_header;
    const quantizer* allocation[2][32];
    uint8_t scale_factor_info[2][32];
    int64_t scale_factor[2][32][3];
    int64_t sample[2][32][3];
    int format, fma;
    orc_synth_state st;
    float u[32];
    uint8_t out[2 * MPEGB200_SAMPLES_PER_FRAME * 4];
    int32_t last_samples[2 * 36 * 32];
};

/* decodeHeader, audio.go:184-272 */
static int decode_header(orc_audio* a) {
    if (!bits_has(&a->buf, 48)) return 0;
    bits_skip_bytes(&a->buf, 0x00);
    int sync = (int)bits_read(&a->buf, 11);
    if (sync != FRAME_SYNC && !bits_find_frame_sync(&a->buf)) return 0;
    a->version = (int)bits_read(&a->buf, 2);
    a->layer = (int)bits_read(&a->buf, 2);
    int has_crc = bits_read1(&a->buf) == 0;
    if (a->version != MPEG_1 || a->layer != LAYER_II) return 0;
    int bitrate_index = (int)bits_read(&a->buf, 4) - 1;
    if (bitrate_index > 13) return 0;
    int samplerate_index = (int)bits_read(&a->buf, 2);
    if (samplerate_index == 3) return 0;
    int padding = bits_read1(&a->buf);
    bits_skip(&a->buf, 1);
    int mode = (int)bits_read(&a->buf, 2);
    if (a->has_header &&
        (a->bitrate_index != bitrate_index || a->samplerate_index != samplerate_index || a->mode != mode))
        return 0;
    a->bitrate_index = bitrate_index;
    a->samplerate_index = samplerate_index;
    a->mode = mode;
    a->has_header = 1;
    if (mode == MODE_STEREO || mode == MODE_JOINT_STEREO)
        a->channels = 2;
    else if (mode == MODE_MONO)
        a->channels = 1;
    if (mode == MODE_JOINT_STEREO) {
        a->bound = ((int)bits_read(&a->buf, 2) + 1) << 2;
    } else {
        bits_skip(&a->buf, 2);
        a->bound = mode == MODE_MONO ? 0 : 32;
    }
    bits_skip(&a->buf, 4);
    if (has_crc) bits_skip(&a->buf, 16);
    /* a bitrate index of -1 ("free format") indexes before the table in Go and panics; guard */
    if (bitrate_index < 0) return 0;
    int frame_size = (144000 * (int)k_bitrate[a->bitrate_index] / (int)k_samplerate[a->samplerate_index]) + padding;
    return frame_size - (has_crc ? 6 : 4);
}

/* readAllocation, audio.go:429-438 */
static const quantizer* read_allocation(orc_audio* a, int sb, int tab3) {
    int tab4 = k_sb_alloc[tab3][sb];
    int qtab = k_alloc_rows[tab4 & 15][bits_read(&a->buf, tab4 >> 4)];
    return qtab ? &k_quant[qtab - 1] : NULL;
}

/* readSamples, audio.go:440-490 */
static void read_samples(orc_audio* a, int ch, int sb, int part) {
    const quantizer* q = a->allocation[ch][sb];
    int64_t sf = a->scale_factor[ch][sb][part];
    int64_t* smp = a->sample[ch][sb];
    if (!q) {
        smp[0] = smp[1] = smp[2] = 0;
        return;
    }
    if (sf == 63) {
        sf = 0;
    } else {
        int shift = (int)(sf / 3);
        sf = (k_scalefactor_base[sf % 3] + (((int64_t)1 << shift) >> 1)) >> shift;
    }
    int64_t adj = q->levels;
    if (q->group) {
        int64_t val = bits_read(&a->buf, q->bits);
        smp[0] = val % adj;
        val /= adj;
        smp[1] = val % adj;
        smp[2] = val / adj;
    } else {
        smp[0] = bits_read(&a->buf, q->bits);
        smp[1] = bits_read(&a->buf, q->bits);
        smp[2] = bits_read(&a->buf, q->bits);
    }
    int64_t scale = 65536 / (adj + 1);
    adj = ((adj + 1) >> 1) - 1;
    for (int i = 0; i < 3; i++) {
        int64_t val = (adj - smp[i]) * scale;
        smp[i] = (val * (sf >> 12) + ((val * (sf & 4095) + 2048) >> 12)) >> 12;
    }
}

/* decodeFrame, audio.go:274-427 */
static void decode_frame(orc_audio* a) {
    int tab1 = a->mode == MODE_MONO ? 0 : 1;
    int tab2 = k_rate_class[tab1][a->bitrate_index];
    int tab3 = k_table_pick[tab2][a->samplerate_index];
    int sblimit = tab3 & 63;
    tab3 >>= 6;
    if (a->bound > sblimit) a->bound = sblimit;

    for (int sb = 0; sb < a->bound; sb++) {
        a->allocation[0][sb] = read_allocation(a, sb, tab3);
        a->allocation[1][sb] = read_allocation(a, sb, tab3);
    }
    for (int sb = a->bound; sb < sblimit; sb++) {
        a->allocation[0][sb] = read_allocation(a, sb, tab3);
        a->allocation[1][sb] = a->allocation[0][sb];
    }
    int channels = a->mode == MODE_MONO ? 1 : 2;
    for (int sb = 0; sb < sblimit; sb++) {
        for (int ch = 0; ch < channels; ch++)
            if (a->allocation[ch][sb]) a->scale_factor_info[ch][sb] = (uint8_t)bits_read(&a->buf, 2);
        if (a->mode == MODE_MONO) a->scale_factor_info[1][sb] = a->scale_factor_info[0][sb];
    }
    for (int sb = 0; sb < sblimit; sb++) {
        for (int ch = 0; ch < channels; ch++) {
            if (!a->allocation[ch][sb]) continue;
            int64_t* sf = a->scale_factor[ch][sb];
            switch (a->scale_factor_info[ch][sb]) { /* audio.go:322-342 */
                case 0:
                    sf[0] = bits_read(&a->buf, 6);
                    sf[1] = bits_read(&a->buf, 6);
                    sf[2] = bits_read(&a->buf, 6);
                    break;
                case 1:
                    sf[0] = sf[1] = bits_read(&a->buf, 6);
                    sf[2] = bits_read(&a->buf, 6);
                    break;
                case 2:
                    sf[0] = sf[1] = sf[2] = bits_read(&a->buf, 6);
                    break;
                case 3:
                    sf[0] = bits_read(&a->buf, 6);
                    sf[1] = sf[2] = bits_read(&a->buf, 6);
                    break;
            }
        }
        if (a->mode == MODE_MONO)
            for (int i = 0; i < 3; i++) a->scale_factor[1][sb][i] = a->scale_factor[0][sb][i];
    }

    const float* dwin = orc_synthesis_window_1024();
    int out_pos = 0, step = 0;
    for (int part = 0; part < 3; part++) {
        for (int granule = 0; granule < 4; granule++) {
            for (int sb = 0; sb < a->bound; sb++) {
                read_samples(a, 0, sb, part);
                read_samples(a, 1, sb, part);
            }
            for (int sb = a->bound; sb < sblimit; sb++) {
                read_samples(a, 0, sb, part);
                for (int i = 0; i < 3; i++) a->sample[1][sb][i] = a->sample[0][sb][i];
            }
            for (int sb = sblimit; sb < 32; sb++)
                for (int i = 0; i < 3; i++) a->sample[0][sb][i] = a->sample[1][sb][i] = 0;

            for (int p = 0; p < 3; p++, step++) { /* synthesis loop, audio.go:377-422 */
                a->st.v_pos = (a->st.v_pos - 64) & 1023;
                for (int ch = 0; ch < 2; ch++) {
                    for (int sb = 0; sb < 32; sb++)
                        a->last_samples[(ch * 36 + step) * 32 + sb] = (int32_t)a->sample[ch][sb][p];
                    orc_idct36((const int64_t(*)[3])a->sample[ch], p, a->st.v[ch], a->st.v_pos);
                    if (a->fma)
                        orc_synth_window_fma(a->u, dwin, a->st.v[ch], a->st.v_pos);
                    else
                        orc_synth_window(a->u, dwin, a->st.v[ch], a->st.v_pos);
                    for (int j = 0; j < 32; j++) emit_sample(a->u[j], a->format, a->out, out_pos + j, ch);
                }
                out_pos += 32;
            }
        }
    }
    bits_align(&a->buf);
}

/* Decode, audio.go:163-182 */
const void* orc_audio_decode(orc_audio* a, double* time) {
    if (a->next_frame_data_size == 0) a->next_frame_data_size = decode_header(a);
    if (a->next_frame_data_size == 0 || !bits_has(&a->buf, (int64_t)a->next_frame_data_size << 3)) return NULL;
    decode_frame(a);
    a->next_frame_data_size = 0;
    if (time) *time = a->time;
    a->samples_decoded += MPEGB200_SAMPLES_PER_FRAME;
    a->time = (double)a->samples_decoded / (double)k_samplerate[a->samplerate_index];
    return a->out;
}

/* NewAudio, audio.go:83-104 */
orc_audio* orc_audio_open(const uint8_t* data, size_t len) {
    orc_audio* a = (orc_audio*)calloc(1, sizeof(orc_audio));
    if (!a) return NULL;
    a->data = (uint8_t*)malloc(len + 8);
    memcpy(a->data, data, len);
    memset(a->data + len, 0, 8);
    a->buf.bytes = a->data;
    a->buf.len = len;
    a->samplerate_index = 3;
    a->next_frame_data_size = decode_header(a);
    return a;
}

void orc_audio_close(orc_audio* a) {
    if (!a) return;
    free(a->data);
    free(a);
}

int orc_audio_has_header(orc_audio* a) { /* audio.go:112-120 */
    if (a->has_header) return 1;
    a->next_frame_data_size = decode_header(a);
    return a->has_header;
}
int orc_audio_samplerate(orc_audio* a) { return orc_audio_has_header(a) ? k_samplerate[a->samplerate_index] : 0; }
int orc_audio_channels(orc_audio* a) { return a->channels; }
void orc_audio_set_format(orc_audio* a, int format) { a->format = format; }
void orc_audio_set_fma(orc_audio* a, int fma) { a->fma = fma; }
/* audio.go:149-154: v and vPos deliberately survive */
void orc_audio_rewind(orc_audio* a) {
    a->buf.bit_index = 0;
    a->buf.has_ended = 0;
    a->time = 0;
    a->samples_decoded = 0;
    a->next_frame_data_size = 0;
}
const int32_t* orc_audio_last_samples(orc_audio* a) { return a->last_samples; }
const orc_synth_state* orc_audio_state(orc_audio* a) { return &a->st; }
