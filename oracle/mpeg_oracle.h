/*
 * mpeg_oracle.h -- CPU restatement of gen2brain/mpeg's decode path (TEST INFRASTRUCTURE).
 *
 * This directory is the parity oracle.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it, and only as the checker or as the reported CPU baseline.  The product library
 * (mpeg_b200/csrc -> libmpegb200.so) never links or calls anything in here.
 *
 * Provenance: the Go reference cannot be built in this image (no Go toolchain), so
 * every function below restates the reference's arithmetic in plain C, citing the
 * reference file:line it follows.  The restatement is pinned by the reference's own
 * golden vectors (tests/test_oracle_golden.py):
 *   video  FNV-1a-64 0xea6d7fcb1340ba3f over testdata/test.mpeg1video (mpeg_test.go:203-231)
 *   audio  FNV-1a-64 0xf1b76cdf8e6cdea5 over testdata/test.mp2, the non-FMA amd64 value
 *          (mpeg_test.go:164-201)
 *   motion-compensation sweep (video_test.go:63-103) and window sweep (audio_test.go:36-64)
 * and by vectors produced by mechanically evaluating the reference's own Go source text
 * for idct / idct36 (tests/golden/make_golden_from_go.py).
 * Frame.RGBA() is Go standard-library arithmetic that no reference test pins:
 * "parity unpinned" for orc_rgba.
 *
 * Go semantics carried over: `int` is int64_t, >> on negatives is arithmetic, / and %
 * truncate toward zero, float32 operations are individually rounded with no FMA
 * contraction (build with -ffp-contract=off).
 */
#ifndef MPEG_ORACLE_H
#define MPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/mpegb200.h" /* packed record types shared with the product ABI */

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- hashing (hash/fnv New64a as used by mpeg_test.go:174,216) ---------- */
uint64_t orc_fnv1a64(uint64_t h, const void* data, size_t n);
#define ORC_FNV_OFFSET 0xcbf29ce484222325ULL

/* ---------- pixel kernels (the hot path) ---------- */
/* idct: video.go:801-928 (both the maxIndex<10 branch and the full transform). */
void orc_idct(int64_t block[64], int max_index);
/* full transform only (video.go:867-927); used to check the two branches agree. */
void orc_idct_full(int64_t block[64]);
/* video.go:943-1002 */
void orc_copy_block_to_dest(const int64_t block[64], uint8_t* dest, int64_t index, int64_t scan);
void orc_add_block_to_dest(const int64_t block[64], uint8_t* dest, int64_t index, int64_t scan);
void orc_copy_value_to_dest(int64_t value, uint8_t* dest, int64_t index, int64_t scan);
void orc_add_value_to_dest(int64_t value, uint8_t* dest, int64_t index, int64_t scan);

/* A frame buffer: one allocation Y|Cb|Cr|pad (video.go:333-355). */
typedef struct orc_frame {
    int width, height;           /* display size */
    int luma_w, luma_h, chroma_w, chroma_h;
    size_t buf_bytes;            /* luma + 2*chroma + luma_w*16 */
    uint8_t* base;               /* y = base, cb = base+luma, cr = cb+chroma */
    uint8_t *y, *cb, *cr;
    double time;
} orc_frame;
int  orc_frame_init(orc_frame* f, int width, int height);
void orc_frame_free(orc_frame* f);

/* copyMacroblock: scalar form of video_test.go:10-43 == video_noasm.go:28-80.
 * Returns 0, or -1 when a source window leaves the frame buffer (where the Go code
 * would panic, video_noasm.go:49-50); nothing is written in that case. */
int orc_copy_macroblock(int motion_h, int motion_v, int mb_row, int mb_col, const orc_frame* s, orc_frame* d);
/* SWAR form of video_noasm.go:14-80 (8 packed bytes per step); same results. */
int orc_copy_macroblock_swar(int motion_h, int motion_v, int mb_row, int mb_col, const orc_frame* s, orc_frame* d);

/* Frame.RGBA(): Go 1.23 image/draw -> image/internal/imageutil.DrawYCbCr, 4:2:0 case
 * (restated from the published standard-library source; parity unpinned). */
void orc_rgba(const orc_frame* f, uint8_t* rgba /* width*height*4, stride 4*width */);
void orc_rgba_batch(const orc_frame* frames, int n, const int32_t* streams, const uint8_t* bufs, uint8_t* out, size_t stride, int threads);
void orc_use_swar_mc(int on);

/* ---------- record-level executor (what the CUDA kernel is compared with) ---------- */
/* Apply one packed macroblock record exactly as decodeMacroblock/decodeBlock would
 * (video.go:544-561, 747-798): prediction first, then per coded block premultiply,
 * idct (or the DC shortcut when only coefficient 0 is non-zero), add or copy. */
int orc_exec_macroblock(const mpegb200_mb* mb, const int16_t* coeffs /* base of coefficient array */,
                        orc_frame* dst, const orc_frame* fwd, const orc_frame* bwd);
/* Whole batch; frames[stream*3 + buf].  threads<=1: serial; otherwise pictures are
 * distributed over OpenMP threads (independent streams).  Returns 0 or -1. */
int orc_exec_pictures(orc_frame* frames, int n_pictures, const mpegb200_picture* pics,
                      size_t n_mb, const mpegb200_mb* mbs, const int16_t* coeffs, int threads);
int orc_max_threads(void);

/* ---------- MPEG-1 video elementary-stream decoder (video.go) ---------- */
typedef struct orc_video orc_video;
orc_video* orc_video_open(const uint8_t* data, size_t len);       /* NewVideo, video.go:110 */
void   orc_video_close(orc_video* v);
int    orc_video_has_header(orc_video* v);                        /* video.go:130 */
int    orc_video_width(orc_video* v);
int    orc_video_height(orc_video* v);
double orc_video_framerate(orc_video* v);
void   orc_video_set_no_delay(orc_video* v, int no_delay);        /* video.go:178 */
void   orc_video_rewind(orc_video* v);                            /* video.go:195 */
/* Decode(): returns NULL at end (video.go:209-268).  The frame stays valid until the next call. */
const orc_frame* orc_video_decode(orc_video* v);
/* physical buffer index (0..2) of the frame last returned */
int    orc_video_last_buf(orc_video* v);

/* Record tap: when enabled, every picture decoded appends its packed records so the
 * same work can be replayed through orc_exec_pictures or the CUDA path. */
void   orc_video_tap_enable(orc_video* v, int on);
/* Pictures decoded during the last orc_video_decode call (may be 0, 1 or more). */
int    orc_video_tap_pictures(orc_video* v, const mpegb200_picture** pics);
size_t orc_video_tap_mbs(orc_video* v, const mpegb200_mb** mbs);
size_t orc_video_tap_blocks(orc_video* v, const int16_t** coeffs);
/* number of macroblock windows the reference would have panicked on (corrupt streams) */
int    orc_video_oob_count(orc_video* v);

/* ---------- MP2 audio ---------- */
/* idct36 (audio.go:492-772): one time slot of 32 subband samples -> 64 floats at d[dp..dp+63]. */
void orc_idct36(const int64_t s[32][3], int ss, float* d, int dp);
/* synthWindow (audio_noasm.go:8-38), unfused multiply-add. */
void orc_synth_window(float u[32], const float d[1024], const float v[1024], int v_pos);
/* the AVX2 / NEON variant: fused multiply-add per lane in tap order (audio_amd64.s:107-156) */
void orc_synth_window_fma(float u[32], const float d[1024], const float v[1024], int v_pos);
const float* orc_synthesis_window_1024(void);   /* table duplicated to 1024, audio.go:95-98 */

typedef struct orc_synth_state { float v[2][1024]; int v_pos; } orc_synth_state;
/* Synthesis section of decodeFrame (audio.go:377-422) for one frame given
 * samples[2][36][32] (layout of mpegb200_audio_synth). out: 2304 values. fma: 0/1 window policy. */
void orc_synth_frame(orc_synth_state* st, const int32_t* samples, int format, void* out, int fma);
/* batch: [n_streams][frames_per_stream] with states[n_streams]; OpenMP over streams */
void orc_synth_batch(orc_synth_state* states, int n_streams, int frames_per_stream,
                     const int32_t* samples, int format, void* out, int fma, int threads);

typedef struct orc_audio orc_audio;
orc_audio* orc_audio_open(const uint8_t* data, size_t len);       /* NewAudio, audio.go:83 */
void  orc_audio_close(orc_audio* a);
int   orc_audio_has_header(orc_audio* a);
int   orc_audio_samplerate(orc_audio* a);
int   orc_audio_channels(orc_audio* a);
void  orc_audio_set_format(orc_audio* a, int format);
void  orc_audio_set_fma(orc_audio* a, int fma);
void  orc_audio_rewind(orc_audio* a);
/* Decode(): returns pointer to 2304 output values (float or int16 per format) or NULL. */
const void* orc_audio_decode(orc_audio* a, double* time);
/* requantised samples of the frame last decoded, [2][36][32] int32 (the kernel's input) */
const int32_t* orc_audio_last_samples(orc_audio* a);
const orc_synth_state* orc_audio_state(orc_audio* a);

/* ---------- MPEG-PS demux (demux.go), enough for testdata/test.mpg ---------- */
/* Splits a program stream into its video (0xE0) and first audio (0xC0) elementary
 * streams; buffers are malloc'ed, caller frees with orc_free. */
int  orc_demux_split(const uint8_t* data, size_t len, uint8_t** video, size_t* video_len,
                     uint8_t** audio, size_t* audio_len, int* n_video_packets, int* n_audio_packets);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
