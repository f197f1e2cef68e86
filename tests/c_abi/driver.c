/*
 * driver.c -- plain-C (C11, -Wall -Werror) consumer of include/mpegb200.h + include/mpegb200_host.h.
 *
 * 1. compile time: the header is valid C (no C++-isms), struct layouts are what the kernels and every binding assume;
 * 2. `driver abi`   (no GPU needed): version, exported host-side helpers, creation fails cleanly without a device;
 * 3. `driver gpu`   (on the B200 box): one tiny decode through the C-ABI with values a human can check --
 *    an intra picture whose blocks are DC only (every pixel = dc), a predicted picture with zero vectors and no
 *    residual (a copy), a predicted block plus a DC residual (pixel + value), RGBA of a grey frame, one MP2 frame of
 *    silence (-0.0f everywhere, audio.go:390), the variable-width transfer form, and the validation switch.
 * Built and run by tests/test_c_abi.py.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpegb200.h"
#include "mpegb200_host.h"

_Static_assert(sizeof(mpegb200_mb) == 16, "mpegb200_mb is 16 bytes");
_Static_assert(offsetof(mpegb200_mb, mb_row) == 0 && offsetof(mpegb200_mb, mb_col) == 2, "mb position");
_Static_assert(offsetof(mpegb200_mb, mv_h) == 4 && offsetof(mpegb200_mb, mv_v) == 6, "mb vector");
_Static_assert(offsetof(mpegb200_mb, flags) == 8 && offsetof(mpegb200_mb, cbp) == 9 && offsetof(mpegb200_mb, pic) == 10, "mb flags");
_Static_assert(offsetof(mpegb200_mb, coeff_block) == 12, "mb coefficient index");
_Static_assert(sizeof(mpegb200_picture) == 16, "mpegb200_picture is 16 bytes");
_Static_assert(offsetof(mpegb200_picture, stream) == 0 && offsetof(mpegb200_picture, type) == 4, "picture head");
_Static_assert(offsetof(mpegb200_picture, dst_buf) == 5 && offsetof(mpegb200_picture, fwd_buf) == 6 && offsetof(mpegb200_picture, bwd_buf) == 7, "picture buffers");
_Static_assert(offsetof(mpegb200_picture, first_mb) == 8 && offsetof(mpegb200_picture, n_mb) == 12, "picture range");
_Static_assert(sizeof(mpegb200_launch) == 32, "mpegb200_launch is 32 bytes");
_Static_assert(sizeof(mpegb200_launch_vlen) == 24, "mpegb200_launch_vlen is 24 bytes");
_Static_assert(MPEGB200_SAMPLES_PER_FRAME == 1152, "audio.go:9");

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            fprintf(stderr, "driver.c:%d: check failed: %s\n", __LINE__, #cond);     \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static int run_abi(void) {
    CHECK(mpegb200_abi_version() == MPEGB200_ABI_VERSION);
    CHECK(mpegb200_vlen_payload_bound(10) >= 10 * 128 + 16);
    /* the converter and its checker are pure host code */
    int16_t blocks[2][64];
    memset(blocks, 0, sizeof blocks);
    blocks[0][0] = 8 * 100;  /* an intra DC (even): raw 12-bit group */
    blocks[0][1] = -3;
    blocks[1][63] = 2047;
    uint32_t headers[2];
    uint64_t chunks[1];
    uint8_t payload[2 * 128 + 16];
    size_t used = 0;
    CHECK(mpegb200_pack_coeffs_vlen(&blocks[0][0], 2, headers, chunks, payload, sizeof payload, &used) == 0);
    CHECK((headers[0] & 15u) == 13u && chunks[0] == 0 && used >= 16);
    CHECK(mpegb200_vlen_validate(headers, chunks, 2, used) == 0);
    CHECK(mpegb200_vlen_validate(headers, chunks, 2, used - 1) != 0);
    /* the host parser works without a GPU: garbage has no sequence header, Decode() == nil */
    uint8_t junk[256];
    for (int i = 0; i < 256; i++) junk[i] = (uint8_t)i;
    mpegb200_video_parser* vp = mpegb200_video_parser_new(junk, sizeof junk);
    CHECK(vp != NULL && !mpegb200_video_parser_has_header(vp));
    mpegb200_video_step st;
    CHECK(mpegb200_video_parser_next(vp, &st) == 0 && !st.has_frame);
    mpegb200_video_parser_free(vp);
    CHECK(mpegb200_video_parser_next(NULL, &st) == MPEGB200_EINVAL);
    /* null handling of the context entry points */
    CHECK(mpegb200_sync(NULL) == MPEGB200_EINVAL && mpegb200_launch_count(NULL) == 0);
    puts("abi ok");
    return 0;
}

static int run_gpu(void) {
    int err = 0;
    mpegb200_ctx* ctx = mpegb200_create(0, 4, &err);
    CHECK(ctx != NULL && err == MPEGB200_OK);
    const int W = 32, H = 32;                      /* 2 x 2 macroblocks */
    CHECK(mpegb200_video_open(ctx, 1, W, H) == 0);
    int lw, lh, cw, ch;
    size_t fb;
    CHECK(mpegb200_video_geometry(ctx, 1, &lw, &lh, &cw, &ch, &fb) == 0);
    CHECK(lw == 32 && lh == 32 && cw == 16 && ch == 16 && fb == 32 * 32 + 2 * 16 * 16 + 32 * 16);
    CHECK(mpegb200_set_validate(ctx, 1) == 0);

    /* picture 1, intra: four macroblocks, six DC-only blocks each; luma = 50 + 10 * mb, Cb = 90, Cr = 200 */
    mpegb200_mb mbs[4];
    static int16_t coeffs[24][64];
    memset(mbs, 0, sizeof mbs);
    memset(coeffs, 0, sizeof coeffs);
    for (int m = 0; m < 4; m++) {
        mbs[m].mb_row = (uint16_t)(m / 2);
        mbs[m].mb_col = (uint16_t)(m % 2);
        mbs[m].flags = MPEGB200_MB_INTRA;
        mbs[m].cbp = 0x3f;
        mbs[m].coeff_block = (uint32_t)(6 * m);
        for (int b = 0; b < 6; b++) coeffs[6 * m + b][0] = (int16_t)(8 * (b < 4 ? 50 + 10 * m : (b == 4 ? 90 : 200)));
    }
    mpegb200_picture pic = {.stream = 1, .type = MPEGB200_PIC_I, .dst_buf = 0, .fwd_buf = 1, .bwd_buf = 2, .first_mb = 0, .n_mb = 4};
    CHECK(mpegb200_video_validate(ctx, 1, &pic, 4, mbs, 24) == 0);
    CHECK(mpegb200_video_decode_pictures(ctx, 1, &pic, 4, mbs, 24, &coeffs[0][0]) == 0);
    static uint8_t y[32 * 32], cb[16 * 16], cr[16 * 16];
    CHECK(mpegb200_video_read_planes(ctx, 1, 0, y, cb, cr) == 0);
    for (int r = 0; r < 32; r++)
        for (int c = 0; c < 32; c++) CHECK(y[r * 32 + c] == 50 + 10 * ((r / 16) * 2 + c / 16));
    for (int i = 0; i < 256; i++) CHECK(cb[i] == 90 && cr[i] == 200);

    /* picture 2, predicted from buffer 0 into buffer 2: zero vectors; macroblock 3 adds a DC residual of +7 to its luma */
    mpegb200_mb pm[4];
    static int16_t pco[4][64];
    memset(pm, 0, sizeof pm);
    memset(pco, 0, sizeof pco);
    for (int m = 0; m < 4; m++) {
        pm[m].mb_row = (uint16_t)(m / 2);
        pm[m].mb_col = (uint16_t)(m % 2);
        pm[m].flags = MPEGB200_MB_PREDICT;
        pm[m].coeff_block = 0;
    }
    pm[3].cbp = 0x3c;                                /* the four luma blocks */
    for (int b = 0; b < 4; b++) pco[b][0] = 8 * 7 + 1;   /* odd level (video.go:732-736): (57 * 32 + 128) >> 8 = 7 */
    mpegb200_picture pp = {.stream = 1, .type = MPEGB200_PIC_P, .dst_buf = 2, .fwd_buf = 0, .bwd_buf = 0, .first_mb = 0, .n_mb = 4};
    /* through the variable-width transfer form */
    uint32_t hd[4];
    uint64_t ck[1];
    uint8_t pl[4 * 128 + 16];
    size_t used = 0;
    CHECK(mpegb200_pack_coeffs_vlen(&pco[0][0], 4, hd, ck, pl, sizeof pl, &used) == 0);
    CHECK(mpegb200_video_decode_pictures_vlen(ctx, 1, &pp, 4, pm, 4, hd, ck, pl, used) == 0);
    CHECK(mpegb200_video_read_planes(ctx, 1, 2, y, cb, cr) == 0);
    for (int r = 0; r < 32; r++)
        for (int c = 0; c < 32; c++) {
            const int m = (r / 16) * 2 + c / 16;
            CHECK(y[r * 32 + c] == 50 + 10 * m + (m == 3 ? 7 : 0));
        }
    for (int i = 0; i < 256; i++) CHECK(cb[i] == 90 && cr[i] == 200);

    /* validation: a vector that reads in front of the frame buffer is refused, nothing is decoded */
    pm[0].mv_v = -200;
    CHECK(mpegb200_video_decode_pictures_vlen(ctx, 1, &pp, 4, pm, 4, hd, ck, pl, used) == MPEGB200_ERECORD);
    CHECK(strlen(mpegb200_last_error(ctx)) > 0);

    /* Frame.RGBA of buffer 0: Cb = 90, Cr = 200 over luma 50: R = (50*0x10101 + 91881*72) >> 16 = 151, G = 11, B = sat(-18) = 0 */
    static uint8_t rgba[32 * 32 * 4];
    CHECK(mpegb200_video_rgba(ctx, 1, 0, rgba) == 0);
    CHECK(rgba[0] == 151 && rgba[1] == 11 && rgba[2] == 0 && rgba[3] == 255);

    /* one MP2 frame of silence: u = +0, and +0 / -1090519040 = -0 (audio.go:390) */
    CHECK(mpegb200_audio_open(ctx, 2) == 0);
    static int32_t samples[2 * 36 * 32];
    static float out[2 * MPEGB200_SAMPLES_PER_FRAME];
    const int32_t id = 2;
    CHECK(mpegb200_audio_synth(ctx, 1, &id, 1, samples, MPEGB200_AUDIO_F32N, out) == 0);
    for (int i = 0; i < 2 * MPEGB200_SAMPLES_PER_FRAME; i++) {
        uint32_t bits;
        memcpy(&bits, &out[i], 4);
        CHECK(bits == 0x80000000u);
    }
    CHECK(mpegb200_audio_synth(ctx, 1, &id, 1, samples, MPEGB200_AUDIO_F32N | MPEGB200_AUDIO_WINDOW_FMA, out) == 0);
    CHECK(mpegb200_audio_synth(ctx, 1, &id, 1, samples, 7, out) == MPEGB200_EINVAL);
    CHECK(mpegb200_launch_count(ctx) >= 5);
    CHECK(mpegb200_audio_close(ctx, 2) == 0 && mpegb200_video_close(ctx, 1) == 0);
    mpegb200_destroy(ctx);
    puts("gpu ok");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) return run_gpu();
    return run_abi();
}
