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
/* the tables of the slice-parallel VLC stage (mpegb200_video_decode_bitstream): 32-byte entries, read by the device as they are */
_Static_assert(sizeof(mpegb200_vlc_picture) == 32 && offsetof(mpegb200_vlc_picture, type) == 4 && offsetof(mpegb200_vlc_picture, fwd_full_px) == 8, "vlc picture head");
_Static_assert(offsetof(mpegb200_vlc_picture, first_slice) == 12 && offsetof(mpegb200_vlc_picture, mb_slot) == 20 && offsetof(mpegb200_vlc_picture, quant) == 28, "vlc picture tail");
_Static_assert(sizeof(mpegb200_vlc_slice) == 32 && offsetof(mpegb200_vlc_slice, next_code) == 8 && offsetof(mpegb200_vlc_slice, stream_left) == 12, "vlc slice head");
_Static_assert(offsetof(mpegb200_vlc_slice, pic) == 16 && offsetof(mpegb200_vlc_slice, vpos) == 20 && offsetof(mpegb200_vlc_slice, mb_slot) == 24 && offsetof(mpegb200_vlc_slice, mb_cap) == 28, "vlc slice tail");
_Static_assert(sizeof(mpegb200_scan_slice) == 24 && sizeof(mpegb200_scan_picture) == 32, "scan step entries");

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            fprintf(stderr, "driver.c:%d: check failed: %s\n", __LINE__, #cond);     \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static const uint8_t kTinyStream[178];

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
    /* scan mode (the host half of the slice-parallel VLC stage) is host code too */
    vp = mpegb200_video_parser_new(kTinyStream, sizeof kTinyStream);
    mpegb200_video_scan_step ss;
    CHECK(vp != NULL && mpegb200_video_parser_next_scan(vp, &ss) == 0 && ss.has_frame && ss.n_pictures == 1 && ss.pictures[0].n_slices == 2);
    CHECK(mpegb200_video_parser_unscan(vp) == 0 && mpegb200_video_parser_unscan(vp) == MPEGB200_EINVAL);     /* one step back, not two */
    CHECK(mpegb200_video_parser_next_scan(vp, &ss) == 0 && ss.has_frame && ss.slices[1].vpos == 2);
    CHECK(mpegb200_video_parser_redo(vp, 0, &st) == 0 && st.has_frame && st.n_launches == 1 && st.launches[0].n_mb == 4);
    const uint64_t wrong[2] = {1, 2};
    CHECK(mpegb200_video_parser_set_start_codes(vp, wrong, 2) == MPEGB200_EINVAL);
    mpegb200_video_parser_free(vp);
    CHECK(mpegb200_video_decode_bitstream(NULL, 0, NULL, 0, NULL, NULL, 0, NULL, 0, 0) == MPEGB200_EINVAL);
    /* null handling of the context entry points */
    CHECK(mpegb200_sync(NULL) == MPEGB200_EINVAL && mpegb200_launch_count(NULL) == 0);
    puts("abi ok");
    return 0;
}

/* A 32 x 32 I picture written by hand: sequence header (default matrices), picture header, two slices (one per macroblock row) of
 * two intra macroblocks each -- address increment "1", type "1", six blocks of dct_dc_size 0 ("100" luma, "00" chroma) and end of
 * block "10" -- so every DC equals its predictor 128 and every pixel of the frame is 128; sequence end code; zero padding (the
 * reference wants 136 bytes behind a sequence start code before it reads the header, video.go:271). */
static const uint8_t kTinyStream[178] = {
    0x00, 0x00, 0x01, 0xb3, 0x02, 0x00, 0x20, 0x15, 0xff, 0xff, 0xe0, 0xa0, 0x00, 0x00, 0x01, 0x00, 0x00, 0x0f, 0xff, 0xf8,
    0x00, 0x00, 0x01, 0x01, 0x0b, 0x94, 0xa5, 0x22, 0x2e, 0x52, 0x94, 0x88, 0x80, 0x00, 0x00, 0x01, 0x02, 0x0b, 0x94, 0xa5,
    0x22, 0x2e, 0x52, 0x94, 0x88, 0x80, 0x00, 0x00, 0x01, 0xb7};

/* Bitstream in: the host scans headers and start codes, the device parses the slices (mpegb200_video_decode_bitstream) --
 * first with the picture's bytes travelling with the wave, then with the stream resident in device memory and its start codes
 * indexed there.  The call sequence a cgo binding would make. */
static int run_bitstream(mpegb200_ctx* ctx, int stream, int resident) {
    mpegb200_video_parser* vp = mpegb200_video_parser_new(kTinyStream, sizeof kTinyStream);
    CHECK(vp != NULL && mpegb200_video_parser_has_header(vp) && mpegb200_video_parser_width(vp) == 32);
    if (resident) {
        uint64_t at[8];
        size_t n = 0;
        CHECK(mpegb200_video_stream_upload(ctx, stream, kTinyStream, sizeof kTinyStream) == 0);
        CHECK(mpegb200_video_stream_index(ctx, stream, at, 8, &n) == 0);
        CHECK(n == 5 && at[0] == 0 && at[1] == 12 && at[2] == 20 && at[3] == 33 && at[4] == 46);   /* b3, picture, slice 1, slice 2, b7 */
        CHECK(mpegb200_video_parser_set_start_codes(vp, at, n) == 0);
    }
    mpegb200_video_scan_step ss;
    CHECK(mpegb200_video_parser_next_scan(vp, &ss) == 0 && ss.has_frame && ss.n_pictures == 1 && ss.host_step == NULL);
    CHECK(ss.pictures[0].type == MPEGB200_PIC_I && ss.pictures[0].n_slices == 2 && ss.mb_w == 2 && ss.mb_h == 2);
    CHECK(ss.slices[0].offset == 24 && ss.slices[0].next_code == 33 && ss.slices[0].vpos == 1 && ss.slices[1].offset == 37 && ss.slices[1].vpos == 2);
    mpegb200_vlc_picture pic;
    memset(&pic, 0, sizeof pic);
    pic.stream = stream;
    pic.type = ss.pictures[0].type;
    pic.dst_buf = ss.pictures[0].dst_buf;
    pic.fwd_buf = ss.pictures[0].fwd_buf;
    pic.bwd_buf = ss.pictures[0].bwd_buf;
    pic.n_slices = 2;
    pic.n_mb_slots = 32;               /* two slices of two macroblocks: 16 record slots each */
    mpegb200_vlc_slice sl[2];
    memset(sl, 0, sizeof sl);
    const uint64_t first = ss.slices[0].offset;
    for (int k = 0; k < 2; k++) {
        sl[k].data_offset = resident ? ss.slices[k].offset : ss.slices[k].offset - first;
        sl[k].next_code = (uint32_t)(ss.slices[k].next_code - ss.slices[k].offset);
        sl[k].stream_left = (uint32_t)(ss.stream_len - ss.slices[k].offset);
        sl[k].vpos = ss.slices[k].vpos;
        sl[k].mb_slot = 16u * (uint32_t)k;
        sl[k].mb_cap = 16;
    }
    const size_t n_bytes = (size_t)(ss.slices[1].next_code + 8 - first);
    CHECK(mpegb200_video_decode_bitstream(ctx, 1, &pic, 2, sl, resident ? NULL : ss.stream + first, resident ? 0 : n_bytes, ss.quant, 1, 32) == 0);
    int flag = -1;
    CHECK(mpegb200_video_bitstream_flags(ctx, &flag, 1) == 0 && flag == 0);
    static mpegb200_mb recs[32];
    CHECK(mpegb200_video_bitstream_records(ctx, recs, NULL) == 0);
    CHECK(recs[0].flags == MPEGB200_MB_INTRA && recs[0].cbp == 0x3f && recs[1].mb_col == 1 && recs[2].pic == 0xffff);
    CHECK(recs[16].mb_row == 1 && recs[17].mb_col == 1 && recs[17].coeff_block == 6 * 16 + 6 && recs[18].pic == 0xffff);
    static uint8_t y[32 * 32], cb[16 * 16], cr[16 * 16];
    memset(y, 0, sizeof y);
    CHECK(mpegb200_video_read_planes(ctx, stream, ss.frame_buf, y, cb, cr) == 0);
    for (int i = 0; i < 32 * 32; i++) CHECK(y[i] == 128);
    for (int i = 0; i < 16 * 16; i++) CHECK(cb[i] == 128 && cr[i] == 128);
    /* a slot table with a hole is refused before anything is enqueued */
    sl[1].mb_slot = 32;
    CHECK(mpegb200_video_decode_bitstream(ctx, 1, &pic, 2, sl, resident ? NULL : ss.stream + first, resident ? 0 : n_bytes, ss.quant, 1, 48) == MPEGB200_EINVAL);
    CHECK(mpegb200_video_parser_next_scan(vp, &ss) == 0 && !ss.has_frame);   /* one picture, handed out by the flush above */
    if (resident) CHECK(mpegb200_video_stream_upload(ctx, stream, NULL, 0) == 0);
    mpegb200_video_parser_free(vp);
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
    /* bitstream in: slices parsed on the device */
    CHECK(mpegb200_video_open(ctx, 3, 32, 32) == 0);
    if (run_bitstream(ctx, 3, 0) != 0) return 1;
    if (run_bitstream(ctx, 3, 1) != 0) return 1;
    CHECK(mpegb200_video_close(ctx, 3) == 0);
    /* the same for a lock-step batch of two streams as ONE call per step (mpegb200_device_stepper_*) */
    {
        mpegb200_video_batch* vb = mpegb200_video_batch_new(2, 1, NULL, NULL);
        CHECK(vb != NULL && mpegb200_video_batch_set_stream(vb, 0, kTinyStream, sizeof kTinyStream) == 0);
        CHECK(mpegb200_video_batch_set_stream(vb, 1, kTinyStream, sizeof kTinyStream) == 0);
        CHECK(mpegb200_video_open(ctx, 2, 32, 32) == 0 && mpegb200_video_open(ctx, 3, 32, 32) == 0);
        mpegb200_device_stepper* ds = mpegb200_device_stepper_new(ctx, vb, 2, 1);
        CHECK(ds != NULL);
        int has[2], buf[2];
        double tm[2];
        CHECK(mpegb200_device_stepper_step(ds, has, buf, tm) == 0 && has[0] == 1 && has[1] == 1 && buf[0] == buf[1] && tm[0] == 0.0);
        static uint8_t y[32 * 32];
        memset(y, 0, sizeof y);
        CHECK(mpegb200_video_read_planes(ctx, 3, buf[1], y, NULL, NULL) == 0);
        for (int i = 0; i < 32 * 32; i++) CHECK(y[i] == 128);
        CHECK(mpegb200_device_stepper_step(ds, has, buf, tm) == 0 && has[0] == 0 && has[1] == 0);
        mpegb200_device_stepper_stats stats;
        CHECK(mpegb200_device_stepper_get_stats(ds, &stats) == 0 && stats.steps == 2 && stats.waves == 1 && stats.flagged_pictures == 0);
        mpegb200_device_stepper_free(ds);
        mpegb200_video_batch_free(vb);
        CHECK(mpegb200_video_close(ctx, 2) == 0 && mpegb200_video_close(ctx, 3) == 0);
    }
    mpegb200_destroy(ctx);
    puts("gpu ok");
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) return run_gpu();
    return run_abi();
}
