/*
 * mpegb200.h -- C-ABI of the B200-native MPEG-1 video / MP2 audio decode hot path.
 *
 * This is the drop-in boundary for gen2brain/mpeg's data-parallel kernels.  The host
 * (the Go package through cgo, or the C++/Python host layer in this repo) keeps the
 * serial work -- PS demux, bit reader, VLC parse, dequantisation -- and hands packed
 * arrays to the entry points below; everything behind them is hand-written sm_100a CUDA.
 *
 * Reference interfaces replaced (file:line in gen2brain/mpeg @ 27c6f084):
 *   idct                         video.go:801-928
 *   copyBlockToDest/addBlockToDest/copyValueToDest/addValueToDest
 *                                video.go:943-1002 (dispatch video.go:747-798)
 *   predictMacroblock            video.go:608-637   (decision resolved by the packer)
 *   copyMacroblock               video_noasm.go:28-80, video_amd64.s:22-595, video_arm64.s:24-319
 *   Frame.RGBA / Pixels          video.go:31-43     (Go stdlib image/draw YCbCr->RGBA)
 *   idct36                       audio.go:492-772
 *   synthWindow                  audio_noasm.go:8-38, audio_amd64.s:33-156, audio_arm64.s:36-85
 *   synthesis loop + scaling     audio.go:377-422
 *
 * Conventions: every function is extern "C", takes plain pointers and sizes, returns
 * 0 on success or a negative MPEGB200_E* code (never throws, never aborts).  One
 * context per GPU and per host thread; distinct contexts are fully independent, which
 * mirrors the reference ("distinct decoders are independent", no shared state).  The
 * caller owns all host memory; the library owns all device memory.  There is no CPU
 * fallback: without a CUDA device every entry point fails with MPEGB200_ECUDA.
 */
#ifndef MPEGB200_H
#define MPEGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPEGB200_ABI_VERSION 1

/* error codes */
#define MPEGB200_OK        0
#define MPEGB200_EINVAL   -1   /* bad argument (null pointer, stream id out of range, bad geometry ...) */
#define MPEGB200_ECUDA    -2   /* CUDA runtime error; text via mpegb200_last_error() */
#define MPEGB200_ENOMEM   -3   /* device or pinned-host allocation failed */
#define MPEGB200_ESTATE   -4   /* stream not opened / already opened / geometry mismatch */
#define MPEGB200_ERECORD  -5   /* a packed record failed validation (see mpegb200_video_validate) */

/* picture coding types (video.go:930-933) */
#define MPEGB200_PIC_I 1
#define MPEGB200_PIC_P 2
#define MPEGB200_PIC_B 3

/* macroblock flags */
#define MPEGB200_MB_INTRA    0x01  /* coded blocks overwrite the destination (video.go:772-784) */
#define MPEGB200_MB_PREDICT  0x02  /* motion-compensated prediction is written first (video.go:544) */
#define MPEGB200_MB_REF_BWD  0x04  /* prediction source is the backward reference, else the forward one.
                                      The packer resolves predictMacroblock (video.go:608-637): for a
                                      B macroblock with both vectors set the backward copy overwrites the
                                      forward one, so only the backward prediction is observable. */

/*
 * One macroblock of work, 16 bytes.  Produced by the host where the reference calls
 * predictMacroblock (video.go:508,544) and decodeBlock (video.go:556-561).
 */
typedef struct mpegb200_mb {
    uint16_t mb_row;       /* macroblock row    (video.go:514) */
    uint16_t mb_col;       /* macroblock column (video.go:515) */
    int16_t  mv_h;         /* half-pel motion vector, FullPx doubling already applied (video.go:612-624) */
    int16_t  mv_v;
    uint8_t  flags;        /* MPEGB200_MB_* */
    uint8_t  cbp;          /* coded block pattern: bit 5 = block 0 (Y top-left) ... bit 0 = block 5 (Cr) */
    uint16_t pic;          /* index of the owning picture in this call's picture array */
    uint32_t coeff_block;  /* index, in 64-coefficient blocks, of this macroblock's first coded block.
                              Blocks of one macroblock are consecutive, in block order 0..5; macroblocks
                              must be packed in array order: coeff_block[i+1] = coeff_block[i]+popcount(cbp[i]) */
} mpegb200_mb;

/*
 * One picture of one stream, 16 bytes.  Buffer indices are the three physical frame
 * buffers of the stream (frameCurrent / frameForward / frameBackward, video.go:97-99);
 * the host mirrors the reference's rotation (video.go:406-409, 430-433) and passes
 * which physical buffer plays which role for this picture.
 */
typedef struct mpegb200_picture {
    int32_t  stream;       /* stream id given to mpegb200_video_open */
    uint8_t  type;         /* MPEGB200_PIC_* (informational; the kernels only look at macroblock flags) */
    uint8_t  dst_buf;      /* 0..2: buffer written by this picture (frameCurrent) */
    uint8_t  fwd_buf;      /* 0..2: forward reference  (frameForward)  */
    uint8_t  bwd_buf;      /* 0..2: backward reference (frameBackward) */
    uint32_t first_mb;     /* informational: index of the picture's first record in the mb array */
    uint32_t n_mb;         /* informational: number of records */
} mpegb200_picture;

/*
 * Coefficients: int16_t[64] per coded block, natural (de-zigzagged, row-major) order,
 * holding the dequantised, oddified and clipped level of video.go:729-741, i.e. the
 * value just before the premultiply of video.go:744 (the kernel multiplies by
 * videoPremultiplierMatrix, video.go:1077-1086, on chip).  Intra DC is carried as
 * dc*8 (so that dc*8*32 == dc<<8, video.go:672); the host must keep |dc| <= 4095.
 */

/* audio output formats (audio.go:12-23) */
#define MPEGB200_AUDIO_F32N    0   /* interleaved normalised float32 (Samples.Interleaved) */
#define MPEGB200_AUDIO_F32NLR  1   /* planar: 1152 left then 1152 right (Samples.Left/Right) */
#define MPEGB200_AUDIO_F32     2   /* interleaved float32 scaled to int32 range (Samples.F32) */
#define MPEGB200_AUDIO_S16     3   /* interleaved int16 (Samples.S16) */
#define MPEGB200_AUDIO_FORMAT_MASK 0xff
/* Window arithmetic, OR-ed into `format`.  The reference has two back-ends for synthWindow that differ in the last
 * bits and holds a golden hash for each (mpeg_test.go:192-196): the pure-Go / SSE one rounds the product and the sum
 * of every tap separately (audio_noasm.go:8-38, audio_amd64.s:33-105; hash 0xf1b76cdf8e6cdea5) -- the default here --
 * and the AVX2 / NEON one uses one fused multiply-add per tap (audio_amd64.s:107-156, audio_arm64.s:36-85; hash
 * 0x50f3ab75f5fb0fb5).  With this flag the kernel reproduces the fused back-end bit for bit (and runs faster). */
#define MPEGB200_AUDIO_WINDOW_FMA  0x100

#define MPEGB200_SAMPLES_PER_FRAME 1152   /* audio.go:9 */

typedef struct mpegb200_ctx mpegb200_ctx;

/* ---- context ---------------------------------------------------------------------- */

int  mpegb200_abi_version(void);
/* Create a context on CUDA device `device` able to hold `max_streams` video and
 * `max_streams` audio streams.  Fails (returns NULL, *err set) when no sm_100 device. */
mpegb200_ctx* mpegb200_create(int device, int max_streams, int* err);
void mpegb200_destroy(mpegb200_ctx* ctx);
const char* mpegb200_last_error(mpegb200_ctx* ctx);
/* Work is enqueued on the context's stream; by default a private non-blocking stream.
 * A caller that wants to time with its own events passes its cudaStream_t here. */
int  mpegb200_set_stream(mpegb200_ctx* ctx, void* cuda_stream);
void* mpegb200_get_stream(mpegb200_ctx* ctx);
int  mpegb200_sync(mpegb200_ctx* ctx);
/* Wait only for the host-to-device copies enqueued so far: after it the host arrays passed to the
 * host-pointer entry points may be overwritten while the kernels are still running. */
int  mpegb200_sync_uploads(mpegb200_ctx* ctx);
/* Make the context's compute stream wait for every asynchronous read-back enqueued so far (mpegb200_video_read_pictures_host
 * runs on a private copy stream): afterwards an event recorded on the compute stream also covers those copies. */
int  mpegb200_join_readbacks(mpegb200_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t mpegb200_launch_count(mpegb200_ctx* ctx);
/* Self-validation.  The decode entry points trust their records (a trusted packer pays nothing); the kernels are
 * memory safe on malformed records but silently drop or clamp them.  With validation on, every host-pointer decode
 * entry point (mpegb200_video_decode_pictures, _packed, _vlen) first runs mpegb200_video_validate (and
 * mpegb200_vlen_validate) on its arrays and returns MPEGB200_ERECORD without enqueuing anything -- the loud failure
 * the reference has as a Go panic (index out of range in copyMacroblock, video_noasm.go:49-50).  Use it for records
 * parsed from untrusted bitstreams; the Python mirror (mpeg_b200.Video / VideoBatch) turns it on.  Setting the
 * environment variable MPEGB200_VALIDATE=1 turns it on for every context (debug aid). */
int  mpegb200_set_validate(mpegb200_ctx* ctx, int on);
/* Measurement aid (bench.py's roofline): while on, every mpegb200_video_decode_pictures* call brackets its plan
 * pre-pass and its arithmetic kernel with CUDA events on the context's stream.  mpegb200_kernel_times synchronises
 * the stream, writes the per-call durations in milliseconds (oldest first, at most `cap` calls; either array may be
 * NULL), forgets them and returns the number of calls written, or a negative error code. */
int  mpegb200_set_kernel_timing(mpegb200_ctx* ctx, int on);
int  mpegb200_kernel_times(mpegb200_ctx* ctx, float* plan_ms, float* fused_ms, int cap);

/* ---- video ------------------------------------------------------------------------ */

/* Allocate and zero the three frame buffers of a stream: each is one allocation
 * Y | Cb | Cr | pad(lumaWidth*16) exactly as initFrame lays it out (video.go:333-355),
 * so half-pel reads past a plane's end see the same bytes the reference sees. */
int mpegb200_video_open(mpegb200_ctx* ctx, int stream, int width, int height);
int mpegb200_video_close(mpegb200_ctx* ctx, int stream);
/* Geometry queries: luma_width = ((width+15)>>4)<<4 etc. (video.go:314-322). */
int mpegb200_video_geometry(mpegb200_ctx* ctx, int stream, int* luma_w, int* luma_h,
                            int* chroma_w, int* chroma_h, size_t* frame_bytes);

/* Check a batch the way the kernels assume it: indices in range, coeff_block packing,
 * motion windows inside the stream's frame buffer (the reference would panic or read
 * foreign memory there, video_noasm.go:49-50), no macroblock written twice in a picture,
 * at most one picture per stream.  Returns 0 or MPEGB200_ERECORD (text in last_error). */
int mpegb200_video_validate(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics,
                            size_t n_mb, const mpegb200_mb* mbs, size_t n_blocks);

/* Decode a batch of mutually independent pictures (at most one per stream): fused
 * motion compensation + 8x8 IDCT + residual add / intra store, one kernel launch.
 * Host-pointer form: copies the three arrays host->device on the context stream first
 * (pinned memory makes that asynchronous), then launches. */
int mpegb200_video_decode_pictures(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics,
                                   size_t n_mb, const mpegb200_mb* mbs,
                                   size_t n_blocks, const int16_t* coeffs);
/* Device-pointer form: the three arrays are already resident in device memory. */
int mpegb200_video_decode_pictures_dev(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* d_pics,
                                       size_t n_mb, const mpegb200_mb* d_mbs,
                                       size_t n_blocks, const int16_t* d_coeffs);

/* 12-bit transfer form of the coefficients: every level of video.go:737-741 lies in [-2048, 2047] and so does
 * a valid intra DC in its dc*8 form (dc <= 255), hence 64 x 12 bits = 96 bytes per block instead of 128 -- a quarter
 * less host-to-device traffic, which is what bounds the end-to-end rate.  Value i of a block occupies bits
 * [12i, 12i+12) of its 96-byte little-endian bit string (two's complement).  mpegb200_pack_coeffs12 converts the
 * int16 form (returns MPEGB200_ERECORD if a value does not fit: use the 16-bit entry point then); the _packed entry
 * point copies the packed blocks to the device, expands them there and continues like
 * mpegb200_video_decode_pictures. */
int mpegb200_pack_coeffs12(const int16_t* coeffs, size_t n_blocks, uint8_t* packed /* 96 * n_blocks */);
int mpegb200_video_decode_pictures_packed(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics,
                                          size_t n_mb, const mpegb200_mb* mbs, size_t n_blocks,
                                          const uint8_t* coeffs12);

/* Variable-width transfer form ("vlen") of the coefficients.  Every level the bitstream carried is odd after the
 * oddification of video.go:732-736, every level it did not carry is 0, only an intra DC (dc*8, video.go:672) is even;
 * neighbours in zig-zag order (video.go:1044-1053) have similar magnitude.  A block travels as
 *   header  : uint32, eight 4-bit codes, code g for zig-zag positions 8g..8g+7:
 *             0      all eight values are zero, no payload
 *             1..12  eight w-bit two's-complement fields (w = code, value i in bits [i*w, i*w+w) of the group's w-byte
 *                    little-endian string) holding c = (x + sign(x)) / 2, i.e. x = 2c - sign(c)
 *             13     eight raw 12-bit two's-complement values (a group with an even non-zero value)
 *             14     eight raw 16-bit values (a group with a value outside [-2048, 2047]: only an intra DC of a damaged
 *                    stream gets there, video.go:666-672 does not bound the predictor)
 *   payload : the groups' bytes back to back, blocks back to back; 32 blocks form a chunk,
 *             chunk_offsets[k] = byte offset of block 32k in the payload.
 * The dense blocks of BASELINE config 3 take about 49 bytes instead of 128 (96 in the 12-bit form), sparse blocks of a
 * real stream 4 bytes plus a few.  mpegb200_pack_coeffs_vlen converts the int16 form (multi-threaded; returns
 * MPEGB200_EINVAL if payload_cap is too small;
 * mpegb200_vlen_payload_bound(n_blocks) always suffices; *payload_bytes includes 16 bytes of padding that must be
 * transferred with it).  The _vlen entry point copies headers, chunk offsets and payload to the device, expands them
 * there (expand_vlen_kernel) and continues like mpegb200_video_decode_pictures.
 * mpegb200_vlen_validate (pure host, one pass over the headers) checks what a foreign packer produced: codes 0..14, chunk offsets
 * back to back and in order, payload_bytes = sum of the group sizes + 16; the device expansion clamps its reads to the payload
 * either way, so a malformed stream yields wrong coefficients, never a fault. */
size_t mpegb200_vlen_payload_bound(size_t n_blocks);
int mpegb200_vlen_validate(const uint32_t* headers, const uint64_t* chunk_offsets, size_t n_blocks, size_t payload_bytes);
int mpegb200_pack_coeffs_vlen(const int16_t* coeffs, size_t n_blocks, uint32_t* headers /* n_blocks */,
                              uint64_t* chunk_offsets /* (n_blocks + 31) / 32 */, uint8_t* payload, size_t payload_cap,
                              size_t* payload_bytes);
int mpegb200_video_decode_pictures_vlen(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics,
                                        size_t n_mb, const mpegb200_mb* mbs, size_t n_blocks, const uint32_t* headers,
                                        const uint64_t* chunk_offsets, const uint8_t* payload, size_t payload_bytes);

/* ---- slice-parallel VLC stage on the device -------------------------------------------------------------
 * The serial walk of decodeSlice / decodeMacroblock / decodeBlock (video.go:436-746) is what bounds bitstream -> frames
 * (tens of 720p pictures per second and host thread against 8 x 10^5 in the decode kernel).  MPEG-1 slices start
 * byte-aligned at start codes and reset every predictor (video.go:436-446), so they parse independently: here the host
 * keeps only the headers and the start-code scan (mpegb200_video_parser_next_scan / mpegb200_video_batch_next_scan in
 * mpegb200_host.h), the compressed bytes go to the device as they are, and one GPU thread per slice walks the macroblocks
 * (vlc_parse_kernel: bit reader, table-driven codes, motion vectors, dequantisation -- video.go:462-746) and writes the SAME
 * packed records the host parser would write (mpegb200_mb + int16[64] blocks) into device memory, which the decode
 * kernels then consume -- no coefficient ever crosses PCIe.
 *
 * A wave = at most one picture per stream.  Record slots: slice s owns mb_cap(s) record slots (a multiple of 16, so a
 * group of 16 records never straddles two slices) and 6 * mb_cap(s) block slots starting at block 6 * mb_slot(s); unused
 * slots hold null records (pic == 0xffff) that the kernels skip.
 *
 * What the serial reference resolves by order of arrival is not reproduced on the device but DETECTED there: a picture one
 * of whose slices runs past the next start code (or into a start code the scan took for a slice), overflows its slots,
 * hits an invalid run (video.go:712-714), drops a macroblock (address outside the picture), reads outside its frame buffer
 * (the reference panics, video_noasm.go:49-50), whose slices overlap, come out of order or end the picture early
 * (video.go:424-426) is FLAGGED: none of its records is executed (its destination buffer stays untouched) and the caller
 * decodes it through the host parser (mpegb200_video_parser_redo / mpegb200_video_batch_redo), which has the reference's
 * serial semantics.  Streams a conforming encoder wrote never flag. */
typedef struct mpegb200_vlc_picture {
    int32_t  stream;                 /* stream id given to mpegb200_video_open */
    uint8_t  type, dst_buf, fwd_buf, bwd_buf;
    uint8_t  fwd_full_px, fwd_r_size, bwd_full_px, bwd_r_size;   /* picture header, video.go:393-404 */
    uint32_t first_slice, n_slices;  /* its slices in the wave's slice array, in bitstream order */
    uint32_t mb_slot, n_mb_slots;    /* record slots of the picture: [mb_slot, mb_slot + n_mb_slots) */
    uint32_t quant;                  /* index of its 128-byte quantiser pair (intra then non-intra, natural order) */
} mpegb200_vlc_picture;              /* 32 bytes */

typedef struct mpegb200_vlc_slice {
    uint64_t data_offset;            /* byte offset, in the wave's bitstream buffer, of the first byte behind the slice start code */
    uint32_t next_code;              /* bytes from there to the next start code of the stream (to its end if there is none):
                                        a slice that consumes more is flagged */
    uint32_t stream_left;            /* bytes from there to the END of the elementary stream, saturated at 2^32 - 1 (the
                                        reader delivers zero bits beyond, buffer.go:203-255) */
    uint32_t pic;                    /* index into the wave's pictures */
    uint32_t vpos;                   /* slice vertical position 1..175 (the start code's last byte) */
    uint32_t mb_slot, mb_cap;        /* record slots of the slice: [mb_slot, mb_slot + mb_cap), both multiples of 16 */
} mpegb200_vlc_slice;                /* 32 bytes */

/* per-picture flags (0 = decoded on the device) */
#define MPEGB200_VLC_INVALID_RUN  0x01   /* a run left the block (video.go:712-714) */
#define MPEGB200_VLC_OVERFLOW     0x02   /* more macroblocks than the slice's record slots */
#define MPEGB200_VLC_OVERRUN      0x04   /* the slice consumed bits beyond next_code */
#define MPEGB200_VLC_DROPPED      0x08   /* a macroblock address outside the picture */
#define MPEGB200_VLC_WINDOW       0x10   /* a motion vector reads outside the frame buffer / reference == destination */
#define MPEGB200_VLC_ORDER        0x20   /* slices overlap or are out of order */
#define MPEGB200_VLC_EARLY_END    0x40   /* a slice other than the last reaches the end of the picture */
#define MPEGB200_VLC_BAD_ARG      0x80   /* the slice or picture table itself is inconsistent */

/* Elementary streams resident in device memory.  mpegb200_video_stream_upload copies a stream's bytes to the device once (any
 * stream id below max_streams, opened or not; a second upload replaces the first; len 0 frees it); a wave whose `bitstream`
 * argument is NULL then reads its slices from there: data_offset counts from the first byte of the stream of the slice's
 * picture.  No compressed byte crosses PCIe per step, only the tables.  mpegb200_video_stream_index finds every start code
 * prefix 00 00 01 xx of the resident stream on the device (the search of buffer.go:279-302, all at once) and writes the byte
 * offsets of the prefixes in ascending order: *n = how many there are; at most `cap` are written (MPEGB200_EINVAL if cap is too
 * small, *n still valid).  mpegb200_video_parser_set_start_codes (mpegb200_host.h) hands them to the host parser, which then
 * never searches a byte. */
int mpegb200_video_stream_upload(mpegb200_ctx* ctx, int stream, const uint8_t* data, size_t len);
int mpegb200_video_stream_index(mpegb200_ctx* ctx, int stream, uint64_t* positions, size_t cap, size_t* n);

/* Parse + decode one wave: uploads the tables and the bytes, runs vlc_parse_kernel and vlc_check_kernel, then the decode
 * kernels on the records they left in device memory.  All arrays are host memory (pinned for asynchronous uploads);
 * `bitstream` holds the pictures' bytes, bitstream_bytes < 2^32 (NULL: the streams are resident, see above).  quant: n_quant x 128 bytes.  n_mb_slots: total record
 * slots of the wave (a multiple of 16, < 2^32 / 6).  Asynchronous like the other decode entry points; returns 0 or a negative code. */
int mpegb200_video_decode_bitstream(mpegb200_ctx* ctx, int n_pictures, const mpegb200_vlc_picture* pics,
                                    size_t n_slices, const mpegb200_vlc_slice* slices,
                                    const uint8_t* bitstream, size_t bitstream_bytes,
                                    const uint8_t* quant, size_t n_quant, size_t n_mb_slots);
/* Waits for the last mpegb200_video_decode_bitstream call and writes its per-picture flags (n ints, n = that call's
 * n_pictures).  Returns the number of flagged pictures or a negative code. */
int mpegb200_video_bitstream_flags(mpegb200_ctx* ctx, int* flags_out, int n);
/* Test / debug aid: the records the last mpegb200_video_decode_bitstream call produced, copied to host memory
 * (n_mb_slots records, 6 * n_mb_slots blocks of 64 int16; null records have pic == 0xffff; the blocks behind a slice's
 * last coded block are undefined). */
int mpegb200_video_bitstream_records(mpegb200_ctx* ctx, mpegb200_mb* mbs, int16_t* coeffs);
/* Measurement aid: duration of the parse + check kernels of the last call in milliseconds (needs mpegb200_set_kernel_timing). */
int mpegb200_video_bitstream_parse_ms(mpegb200_ctx* ctx, float* ms);

/* Plane read-back for *Frame (Plane.Data, video.go:50-54): copies the macroblock-padded
 * planes of physical buffer `buf` to host memory.  Any of y/cb/cr may be NULL. */
int mpegb200_video_read_planes(mpegb200_ctx* ctx, int stream, int buf,
                               uint8_t* y, uint8_t* cb, uint8_t* cr);
/* Write planes (test set-up / resume from a checkpointed reference frame). */
int mpegb200_video_write_planes(mpegb200_ctx* ctx, int stream, int buf,
                                const uint8_t* y, const uint8_t* cb, const uint8_t* cr);
/* Whole frame buffer (Y|Cb|Cr|pad) to/from host. */
int mpegb200_video_read_frame(mpegb200_ctx* ctx, int stream, int buf, uint8_t* dst, size_t dst_bytes);
int mpegb200_video_write_frame(mpegb200_ctx* ctx, int stream, int buf, const uint8_t* src, size_t src_bytes);
/* Batched, asynchronous picture read-back: frame i = (streams[i], bufs[i]); its Y|Cb|Cr bytes
 * (luma + 2*chroma, without the pad) go to dst + i*dst_stride.  `dst` is host memory (pinned for
 * true asynchrony) for _host, device memory for _dev.  Enqueued on the context stream: call
 * mpegb200_sync before touching the bytes.  This is what feeds Plane.Data of many *Frames at once
 * and, in the multi-GPU layout, the send buffer of the NCCL gather. */
int mpegb200_video_read_pictures_host(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                      uint8_t* dst, size_t dst_stride);
int mpegb200_video_read_pictures_dev(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                     uint8_t* d_dst, size_t dst_stride);
/* Device address of a frame buffer (Y at +0, Cb at +luma_w*luma_h, Cr after it). */
void* mpegb200_video_frame_dev(mpegb200_ctx* ctx, int stream, int buf);

/* Display ring: the frames Video.Decode() returns, kept in device memory in display order for `depth` steps, so that
 * consumers (RGBA conversion of an older frame, read-back, the NCCL gather) may lag behind the decoder instead of
 * having to finish before the next Decode() re-uses the frame buffer ("valid until the next call", mpeg.go:413-415).
 * A ring covers n streams; a slot holds their n pictures (Y|Cb|Cr without the pad) `stride` bytes apart.
 * mpegb200_video_ring_push copies, on the context's stream, the display frame of every stream (bufs[i] = the physical
 * buffer the host parser named as frame_buf; 255 = this stream returned no frame in this step, its place in the slot is
 * left as it was) into the next slot and returns that slot's index (or a negative error code). */
typedef struct mpegb200_ring mpegb200_ring;
mpegb200_ring* mpegb200_video_ring_new(mpegb200_ctx* ctx, int n, const int32_t* streams, int depth);
void  mpegb200_video_ring_free(mpegb200_ring* ring);
int   mpegb200_video_ring_push(mpegb200_ring* ring, const uint8_t* bufs);
/* Device address of slot `slot` (picture i at + i * *stride); valid until the slot is overwritten depth pushes later. */
void* mpegb200_video_ring_slot_dev(mpegb200_ring* ring, int slot, size_t* stride);
/* Asynchronous copy of a slot to (pinned) host memory on the read-back stream, ordered behind the pushes so far. */
int   mpegb200_video_ring_read_host(mpegb200_ring* ring, int slot, uint8_t* dst, size_t dst_stride);

/* Frame.RGBA(): YCbCr 4:2:0 -> RGBA8 over the display rectangle width x height,
 * destination stride 4*width (video.go:31-36, 367-371).  Host form copies the result
 * into `rgba` (width*height*4 bytes). */
int mpegb200_video_rgba(mpegb200_ctx* ctx, int stream, int buf, uint8_t* rgba);
/* Batched device form: frame i = (streams[i], bufs[i]) -> d_rgba + i*rgba_stride_bytes. */
int mpegb200_video_rgba_batch_dev(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                  uint8_t* d_rgba, size_t rgba_stride_bytes);

/* ---- audio ------------------------------------------------------------------------ */

/* Allocate the synthesis state of a stream: V[2][1024] = 0, vPos = 0 (audio.go:77-79). */
int mpegb200_audio_open(mpegb200_ctx* ctx, int stream);
int mpegb200_audio_close(mpegb200_ctx* ctx, int stream);
/* Synthesis of a rectangular batch: each of the n_streams listed streams (distinct ids)
 * advances by frames_per_stream MP2 frames, in decode order (the V history and vPos carry
 * from one frame to the next, audio.go:380, and survive across calls like they survive
 * Rewind, audio.go:149).
 * samples: int32[n_streams][frames_per_stream][2][36][32] = requantised subband samples
 * sample[ch][sb][p] of audio.go:72 laid out [ch][3*(4*part+granule)+p][sb] (both channels
 * always present: mono streams mirror channel 0, audio.go:362-367).
 * out: per frame 2304 values in `format` (float32, or int16 for MPEGB200_AUDIO_S16),
 * frames in the same [stream][frame] order.  stream_ids is always a HOST array. */
int mpegb200_audio_synth(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids,
                         int frames_per_stream, const int32_t* samples, int format, void* out);
int mpegb200_audio_synth_dev(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids,
                             int frames_per_stream, const int32_t* d_samples, int format, void* d_out);
/* Coded form: requantisation (audio.go:440-490, second half) on the device.  Per frame the host passes
 *   info : for every (channel, subband) the quantiser number 1..17 of audio.go:955-973 (0 = no bits allocated) and the
 *          scale-factor index 0..63 of each of the three parts (audio.go:322-342); 256 bytes;
 *   codes: uint16[2][36][32], the sample codes as read from the bitstream (degrouped, audio.go:462-475), laid out like
 *          `samples` above; 4608 bytes.
 * The device computes sample = ((adj - code) * scale) x scalefactor exactly as audio.go:476-489 (int32 suffices: |val| <=
 * 2^15, scalefactor <= 2^25 split at 12 bits) and continues like mpegb200_audio_synth. */
typedef struct mpegb200_audio_frame_info {
    uint8_t quant[2][32];
    uint8_t scf[2][32][3];
} mpegb200_audio_frame_info;
int mpegb200_audio_synth_coded(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids, int frames_per_stream,
                               const mpegb200_audio_frame_info* info, const uint16_t* codes, int format, void* out);
/* State read-back in the reference's own form: v = float[2][1024], *v_pos (audio.go:63,78). */
int mpegb200_audio_read_state(mpegb200_ctx* ctx, int stream, float* v, int* v_pos);
int mpegb200_audio_write_state(mpegb200_ctx* ctx, int stream, const float* v, int v_pos);

/* ---- pinned host memory helpers (for the packer's staging arrays) ------------------- */
void* mpegb200_host_alloc(size_t bytes);
void  mpegb200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MPEGB200_H */
