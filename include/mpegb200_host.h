/*
 * mpegb200_host.h -- host half of the drop-in: the serial work gen2brain/mpeg keeps on the CPU
 * (bit reader, VLC parse, dequantisation, PS demux), emitting the packed records of mpegb200.h.
 *
 * In the reference these are the Go functions around the kernels (file:line @ 27c6f084):
 *   Video:  NewVideo video.go:110, HasHeader :130, Decode :209-268 (display order, end flush),
 *           decodeSequenceHeader :270-331, decodePicture :374-434, decodeSlice :436-460,
 *           decodeMacroblock :462-562, decodeMotionVector(s) :564-606, predictMacroblock :608-637
 *           (decision only), decodeBlock :639-746 (parse + dequantise + oddify + clip)
 *   Audio:  NewAudio audio.go:83, Decode :163-182, decodeHeader :184-272, decodeFrame :274-375
 *           (allocation, scale factors, readSamples :440-490 requantisation)
 *   Demux:  demux.go:85-138 (headers), :473-568 (packet walk)
 * A Go build keeps using its own Go code for this half; this C++ restatement exists
 * because no Go toolchain is available here, and it is what the Python mirror (mpeg_b200.Video /
 * Audio / MPEG) drives.  It is product code: written independently of oracle/ and cross-checked
 * against it record for record in tests/test_host_parser.py.
 *
 * All functions are plain C; parsers are CPU-only objects and need no GPU.
 */
#ifndef MPEGB200_HOST_H
#define MPEGB200_HOST_H

#include "mpegb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- video elementary stream ------------------------------------------------------------------ */

typedef struct mpegb200_video_parser mpegb200_video_parser;

/* One kernel launch worth of records: a picture, or one wave of a picture that rewrites macroblocks
 * (a later record that defines all six blocks replaces the earlier one; one that defines only some
 * blocks goes into the next wave -- the reference decodes serially, video.go:454-459). */
typedef struct mpegb200_launch {
    mpegb200_picture picture;   /* stream = 0; the caller patches its own stream id in */
    uint32_t first_mb, n_mb;    /* slice of the step's mb array; coeff_block is relative to first_block */
    uint32_t first_block, n_blocks;
} mpegb200_launch;

/* Where a launch's coefficients sit when the parser emits the variable-width transfer form (mpegb200.h, "vlen")
 * instead of int16 blocks: headers at vlen_headers + first_block (n_blocks of them, as for the int16 form), its chunk
 * offsets at vlen_chunk_offsets + first_chunk ((n_blocks + 31) / 32 of them, relative to the launch's payload), its
 * payload at vlen_payload + payload_offset, payload_bytes long (16 bytes of padding included). */
typedef struct mpegb200_launch_vlen {
    uint32_t first_chunk;
    uint32_t reserved;
    uint64_t payload_offset;
    uint64_t payload_bytes;
} mpegb200_launch_vlen;

/* What one Video.Decode() call amounts to (video.go:209-268). */
typedef struct mpegb200_video_step {
    int has_frame;              /* 0: end of stream (Decode() == nil) */
    int frame_buf;              /* physical buffer 0..2 holding the frame to return */
    double time;                /* Frame.Time (video.go:263) */
    int n_launches;             /* pictures decoded on the way (0 for the end-of-stream flush, video.go:223-229) */
    const mpegb200_launch* launches;
    const mpegb200_mb* mbs;
    const int16_t* coeffs;      /* 64 per block; NULL in vlen mode */
    /* vlen mode (mpegb200_video_parser_set_vlen), else NULL: */
    const mpegb200_launch_vlen* vlen_launches;   /* one per launch */
    const uint32_t* vlen_headers;
    const uint64_t* vlen_chunk_offsets;
    const uint8_t* vlen_payload;
} mpegb200_video_step;

mpegb200_video_parser* mpegb200_video_parser_new(const uint8_t* data, size_t len);   /* NewVideo; copies data */
void   mpegb200_video_parser_free(mpegb200_video_parser* v);
int    mpegb200_video_parser_has_header(mpegb200_video_parser* v);
int    mpegb200_video_parser_width(mpegb200_video_parser* v);
int    mpegb200_video_parser_height(mpegb200_video_parser* v);
double mpegb200_video_parser_framerate(mpegb200_video_parser* v);
void   mpegb200_video_parser_set_no_delay(mpegb200_video_parser* v, int no_delay);   /* video.go:178 */
/* Emit the coefficients in the variable-width transfer form: decodeBlock (video.go:639-746) walks a block's
 * coefficients in zig-zag order, which is the group order of that form, so the parser writes headers and payload as it
 * goes and no int16[64] array (128 bytes per block) is ever materialised or converted.  Byte for byte what
 * mpegb200_pack_coeffs_vlen makes of the int16 blocks (tests/test_host_parser.py).  Takes effect at the next picture. */
void   mpegb200_video_parser_set_vlen(mpegb200_video_parser* v, int on);
void   mpegb200_video_parser_rewind(mpegb200_video_parser* v);                       /* video.go:195 */
int    mpegb200_video_parser_has_ended(mpegb200_video_parser* v);
/* Parse up to and including the picture that makes a frame due.  Pointers in *out stay valid until
 * the next call on the same parser.  Returns 0, or MPEGB200_EINVAL. */
int    mpegb200_video_parser_next(mpegb200_video_parser* v, mpegb200_video_step* out);

/* ---- scan mode: headers and start codes on the host, slices to the device (mpegb200_video_decode_bitstream) ---- */

typedef struct mpegb200_scan_slice {
    uint64_t offset;            /* byte offset, in the elementary stream, of the first byte behind the slice start code */
    uint64_t next_code;         /* byte offset of the next start code (its 00 00 01), or the stream's length if there is none */
    uint32_t vpos;              /* vertical position 1..175 */
    uint32_t reserved;
} mpegb200_scan_slice;

typedef struct mpegb200_scan_picture {
    uint8_t  type, dst_buf, fwd_buf, bwd_buf;
    uint8_t  fwd_full_px, fwd_r_size, bwd_full_px, bwd_r_size;
    uint32_t first_slice, n_slices;    /* into the step's slice array */
    uint64_t begin, end;               /* byte range of the picture in the elementary stream: its start code .. behind the start code that ended its slices */
} mpegb200_scan_picture;

/* What one Video.Decode() call amounts to when the slices are left unparsed. */
typedef struct mpegb200_video_scan_step {
    int has_frame;
    int frame_buf;
    double time;
    int n_pictures;
    const mpegb200_scan_picture* pictures;
    const mpegb200_scan_slice* slices;
    const uint8_t* stream;             /* the parser's copy of the elementary stream, stream_len bytes (+ 16 readable zero bytes) */
    size_t stream_len;
    const uint8_t* quant;              /* 128 bytes: intra then non-intra quantiser matrix, natural order (video.go:299-309) */
    int mb_w, mb_h;
    /* Not NULL: the host parsed this step itself (n_pictures == 0) and these are its launches.  Happens while coefficients of a
     * dropped block (video.go:712-714) are pending: the serial reference leaks them into the next block it decodes, and only
     * the host parser carries that state. */
    const mpegb200_video_step* host_step;
} mpegb200_video_scan_step;

/* Like mpegb200_video_parser_next, stopping at the slice start codes.  The two calls may alternate between steps. */
int mpegb200_video_parser_next_scan(mpegb200_video_parser* v, mpegb200_video_scan_step* out);
/* Every start code prefix (00 00 01) of the stream, byte offsets ascending -- e.g. from mpegb200_video_stream_index, which finds
 * them on the device: the parser then never searches a byte of the stream (it checks that the positions are start codes of its
 * stream; MPEGB200_EINVAL otherwise). */
int mpegb200_video_parser_set_start_codes(mpegb200_video_parser* v, const uint64_t* positions, size_t n);
/* Withdraw the last scan step: the parser stands where it stood before that mpegb200_video_parser_next_scan call and the
 * step before it is the "last scan step" again (for mpegb200_video_parser_redo).  One step back only.  This is what lets a
 * caller scan step k + 1 while the device still works on step k: if step k then flags a picture, step k + 1 is withdrawn,
 * step k's tail re-parsed, and step k + 1 scanned again from where the serial reference really stands. */
int mpegb200_video_parser_unscan(mpegb200_video_parser* v);
/* The device stage flagged picture k of the last scan step: the parser goes back in front of that picture and parses the
 * REST OF THE STEP (picture k and whatever Video.Decode() would decode behind it) in full, with the serial semantics of
 * the reference; it stands where the reference would stand afterwards, which may differ from where the scan stood (a
 * damaged slice may swallow the start codes behind it).  *out is the step's tail: its launches replace picture k and every
 * later picture of the scan step, its has_frame / frame_buf / time replace the scan step's. */
int mpegb200_video_parser_redo(mpegb200_video_parser* v, int k, mpegb200_video_step* out);

/* ---- many video streams in lock-step (the batched deployment of INTEGRATION.md section 6) ------- */

typedef struct mpegb200_video_batch mpegb200_video_batch;

/* One kernel launch for many streams: the w-th launch of every stream's current Decode() step.
 * At most one picture per stream; pics[i].stream is the stream's index in the batch. */
typedef struct mpegb200_wave {
    int n_pictures;
    const mpegb200_picture* pics;
    size_t n_mb;
    const mpegb200_mb* mbs;        /* pic = index into pics, coeff_block = index into coeffs */
    size_t n_blocks;
    const int16_t* coeffs;         /* NULL in vlen mode */
    /* vlen mode (mpegb200_video_batch_set_vlen): arguments of mpegb200_video_decode_pictures_vlen */
    const uint32_t* vlen_headers;
    const uint64_t* vlen_chunk_offsets;
    const uint8_t* vlen_payload;
    size_t vlen_payload_bytes;
} mpegb200_wave;

typedef struct mpegb200_batch_step {
    int n_streams;
    const int* has_frame;          /* per stream: 1 if this step returns a frame (0 = that stream has ended) */
    const int* frame_buf;          /* per stream: physical buffer 0..2 of the frame */
    const double* time;            /* per stream: Frame.Time */
    int n_waves;
    const mpegb200_wave* waves;    /* launch in order; each wave is duplicate-free (see mpegb200_launch) */
} mpegb200_batch_step;

/* n_streams parsers driven by `threads` host threads.  The wave arrays are allocated with alloc/free when
 * given (pass mpegb200_host_alloc / mpegb200_host_free to get pinned staging), else with malloc. */
mpegb200_video_batch* mpegb200_video_batch_new(int n_streams, int threads, void* (*alloc)(size_t), void (*free_fn)(void*));
void mpegb200_video_batch_free(mpegb200_video_batch* b);
int  mpegb200_video_batch_set_stream(mpegb200_video_batch* b, int index, const uint8_t* data, size_t len);
int  mpegb200_video_batch_stream_size(mpegb200_video_batch* b, int index, int* width, int* height);
/* Every parser of the batch emits the variable-width form; the waves carry headers / chunk offsets / payload. */
int  mpegb200_video_batch_set_vlen(mpegb200_video_batch* b, int on);
/* One Video.Decode() step of every stream, parsed in parallel and merged into waves. */
int  mpegb200_video_batch_next(mpegb200_video_batch* b, mpegb200_batch_step* out);

/* The same in scan mode: the waves carry slice tables and the pictures' compressed bytes for mpegb200_video_decode_bitstream. */
typedef struct mpegb200_vlc_wave {
    int n_pictures;
    const mpegb200_vlc_picture* pics;   /* pics[i].stream is the stream's index in the batch */
    const int32_t* step_picture;        /* per picture: its index in its stream's scan step (argument of mpegb200_video_batch_redo) */
    size_t n_slices;
    const mpegb200_vlc_slice* slices;
    const uint8_t* bitstream;
    size_t bitstream_bytes;
    const uint8_t* quant;               /* n_pictures x 128 bytes */
    size_t n_quant;                     /* = n_pictures */
    size_t n_mb_slots;
} mpegb200_vlc_wave;

typedef struct mpegb200_batch_scan_step {
    int n_streams;
    const int* has_frame;
    const int* frame_buf;
    const double* time;
    int n_waves;
    const mpegb200_vlc_wave* waves;
    /* streams whose step the host parsed itself (mpegb200_video_scan_step.host_step): run their launches like those of
     * mpegb200_video_parser_next; they take no part in the waves */
    int n_host;
    const int* host_index;
    const mpegb200_video_step* host_steps;
} mpegb200_batch_scan_step;

int  mpegb200_video_batch_next_scan(mpegb200_video_batch* b, mpegb200_batch_scan_step* out);
/* The parser of stream `index` (owned by the batch), for the per-stream controls: mpegb200_video_parser_set_no_delay, _rewind,
 * _has_ended, _framerate.  Do not call its _next functions. */
mpegb200_video_parser* mpegb200_video_batch_parser(mpegb200_video_batch* b, int index);
/* Resident mode: the streams were uploaded with mpegb200_video_stream_upload (stream id = index + the caller's offset); the waves
 * then carry no bytes (bitstream = NULL) and the slices' data_offset counts from the first byte of their stream. */
int  mpegb200_video_batch_set_resident(mpegb200_video_batch* b, int on);
int  mpegb200_video_batch_set_start_codes(mpegb200_video_batch* b, int index, const uint64_t* positions, size_t n);
/* mpegb200_video_parser_unscan for every stream of the batch (the wave arrays of the step before stay valid). */
int  mpegb200_video_batch_unscan(mpegb200_video_batch* b);
/* mpegb200_video_parser_redo for stream `index` of the batch: the rest of its step from picture `step_picture` on, parsed on
 * the host.  The stream's pictures in the later waves of this step are void (set their type to 0 before the wave is decoded:
 * the device then skips them); has_frame / frame_buf / time of the stream are those of *out. */
int  mpegb200_video_batch_redo(mpegb200_video_batch* b, int index, int step_picture, mpegb200_video_step* out);

/* ---- the whole lock-step step with the slices parsed on the device, as one call -------------------------------------------
 * What a binding would otherwise write itself around mpegb200_video_batch_next_scan / mpegb200_video_decode_bitstream /
 * mpegb200_video_bitstream_flags / mpegb200_video_batch_redo / _unscan: host-parsed steps, waves,
 * scan-ahead, flags, re-parse of a flagged picture's step tail.  Stream i of the batch is stream id first_stream + i of the
 * context (opened by the caller with mpegb200_video_open; resident mode: uploaded and indexed by the caller, see
 * mpegb200_video_batch_set_resident).  One stepper per batch; do not call the batch's _next functions beside it. */
typedef struct mpegb200_device_stepper mpegb200_device_stepper;
typedef struct mpegb200_device_stepper_stats {
    uint64_t steps, waves;
    uint64_t flagged_pictures;      /* pictures the device flagged: their step's tail took the host parser */
    uint64_t host_steps;            /* steps the host parsed itself (stale coefficients pending) */
    uint64_t withdrawn_scans;       /* scans made ahead of time and withdrawn because the step before flagged */
    double seconds_host_scan, seconds_submit, seconds_waiting;   /* host time in the scan, in the submission, waiting for the flags */
} mpegb200_device_stepper_stats;

mpegb200_device_stepper* mpegb200_device_stepper_new(mpegb200_ctx* ctx, mpegb200_video_batch* batch, int first_stream, int scan_ahead);
void mpegb200_device_stepper_free(mpegb200_device_stepper* s);
/* One Video.Decode() of every stream: has_frame / frame_buf / time receive n_streams entries each (has_frame 0: that stream has
 * ended).  The kernels of the step's last wave may still run when it returns (mpegb200_sync, or a read-back, waits for them). */
int  mpegb200_device_stepper_step(mpegb200_device_stepper* s, int* has_frame, int* frame_buf, double* time);
/* Withdraw a scan made ahead of time (before a control call that moves a parser: rewind, no-delay). */
int  mpegb200_device_stepper_drop_scan_ahead(mpegb200_device_stepper* s);
int  mpegb200_device_stepper_get_stats(mpegb200_device_stepper* s, mpegb200_device_stepper_stats* out);

/* ---- MP2 elementary stream --------------------------------------------------------------------- */

typedef struct mpegb200_audio_parser mpegb200_audio_parser;

mpegb200_audio_parser* mpegb200_audio_parser_new(const uint8_t* data, size_t len);   /* NewAudio */
void   mpegb200_audio_parser_free(mpegb200_audio_parser* a);
int    mpegb200_audio_parser_has_header(mpegb200_audio_parser* a);
int    mpegb200_audio_parser_samplerate(mpegb200_audio_parser* a);
int    mpegb200_audio_parser_channels(mpegb200_audio_parser* a);
void   mpegb200_audio_parser_rewind(mpegb200_audio_parser* a);                       /* audio.go:149 */
/* Parse one frame: fills samples[2][36][32] (layout of mpegb200_audio_synth) and *time.
 * Returns 1 if a frame was parsed, 0 at the end (Decode() == nil). */
int    mpegb200_audio_parser_next(mpegb200_audio_parser* a, int32_t* samples, double* time);
/* The same frame without the requantisation of audio.go:476-489: per (channel, subband) the quantiser and the three
 * scale-factor indices, and the sample codes as the bitstream has them (degrouped) -- the input of
 * mpegb200_audio_synth_coded, which requantises on the device (4.9 KB per frame over PCIe instead of 9.2 KB). */
int    mpegb200_audio_parser_next_coded(mpegb200_audio_parser* a, mpegb200_audio_frame_info* info, uint16_t* codes, double* time);

/* ---- many MP2 streams in lock-step -------------------------------------------------------------- */

typedef struct mpegb200_audio_batch mpegb200_audio_batch;

/* One step: every stream that still has data contributes up to frames_per_stream frames.  Streams that delivered all
 * frames_per_stream frames form the rectangular batch of mpegb200_audio_synth (n_full of them, batch indices in
 * full_index, samples [n_full][frames_per_stream][2][36][32]); a stream that ran dry on the way delivers its last
 * k < frames_per_stream frames separately (tail_*; one mpegb200_audio_synth call each with frames_per_stream = k). */
typedef struct mpegb200_audio_batch_step {
    int n_streams;
    const int* n_frames;           /* per stream: frames parsed in this step (0 = ended) */
    const double* time;            /* per stream: Samples.Time of its first frame of the step (audio.go:176) */
    int frames_per_stream;
    int n_full;
    const int32_t* full_index;     /* batch indices of the streams in the rectangular part */
    const int32_t* full_samples;
    int n_tail;
    const int32_t* tail_index;     /* batch index of each tail stream */
    const int32_t* tail_frames;    /* its number of frames */
    const int32_t* tail_samples;   /* the tails' frames back to back: [sum of tail_frames][2][36][32] */
} mpegb200_audio_batch_step;

mpegb200_audio_batch* mpegb200_audio_batch_new(int n_streams, int threads, void* (*alloc)(size_t), void (*free_fn)(void*));
void mpegb200_audio_batch_free(mpegb200_audio_batch* b);
int  mpegb200_audio_batch_set_stream(mpegb200_audio_batch* b, int index, const uint8_t* data, size_t len);
int  mpegb200_audio_batch_stream_info(mpegb200_audio_batch* b, int index, int* samplerate, int* channels);
/* Parse up to frames_per_stream frames of every stream on the batch's threads (Audio.Decode's parse half, audio.go:163-375). */
int  mpegb200_audio_batch_next(mpegb200_audio_batch* b, int frames_per_stream, mpegb200_audio_batch_step* out);

/* ---- MPEG program stream ----------------------------------------------------------------------- */

/* Split a program stream into its video (0xE0) and first audio (0xC0) elementary streams.  The
 * buffers are malloc'ed; release with mpegb200_buffer_free.  Returns 0 or MPEGB200_EINVAL
 * (no pack / system header: ErrInvalidHeader, demux.go:32). */
int  mpegb200_demux_split(const uint8_t* data, size_t len, uint8_t** video, size_t* video_len, uint8_t** audio,
                          size_t* audio_len, int* n_video_packets, int* n_audio_packets);
void mpegb200_buffer_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MPEGB200_HOST_H */
