// common.cuh -- device-side structures shared by the sm_100a kernels and the C-ABI glue.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mpegb200.h"

namespace mpegb200 {

// Per-stream video geometry and storage, mirrored in device memory (one entry per stream id).
// The three physical frame buffers (frameCurrent/Forward/Backward, video.go:97-99) of a stream
// live in one allocation, `buf_stride` bytes apart; each is Y | Cb | Cr | pad(luma_w*16) exactly
// as initFrame lays it out (video.go:333-355) followed by >= 64 bytes of slack so that 16-byte
// aligned super-set loads of a motion window never leave the allocation.
struct StreamInfo {
    uint8_t* base;        // buffer b at base + b * buf_stride
    uint32_t buf_stride;  // multiple of luma_w*16 (and of 256)
    uint32_t buf_bytes;   // luma + 2*chroma + luma_w*16 (video.go:340)
    uint16_t luma_w, luma_h;
    uint16_t mb_w, mb_h;
    uint16_t width, height;  // display size (video.go:342-343)
    uint16_t slab;           // index into the tensor-map table (TMA kernel)
    uint16_t slot;           // stream slot inside the slab: buffer b is tensor z = slot*3 + b
    uint8_t open;
    uint8_t tma_ok;          // chroma pitch is a multiple of 16 bytes (even mb_w): TMA path usable
    uint8_t pad[10];
};
static_assert(sizeof(StreamInfo) == 48, "StreamInfo layout");

// Tensor maps of one slab (streams of one geometry sharing one allocation), kept in global memory.
//   luma  : 3-D u8 tensor {luma_w + 32, buf_stride / luma_w, 3 * capacity}, strides {luma_w, buf_stride}:
//           the whole frame buffer seen as rows of luma_w bytes, rows overlapping by 32 bytes so that a
//           window crossing the right edge continues on the next row exactly like linear addressing does.
//   chroma: 4-D {chroma_w + 32, rows to the end of the buffer, plane (Cb, Cr: chroma_bytes apart), 3 * capacity}:
//           one box fetches the Cb and the Cr window of a macroblock together.
// Box heights: 17 rows of luma window plus 3 rows of head-room for the per-macroblock row phase of the staging
// (video_fused_tma.cu); 9 rows of chroma, no phase (measured: 20/9 0.426 ms, 20/10 0.430, 18/9 0.432, 17/9 0.449).
#ifndef MPEGB200_LUMA_BOX_ROWS
#define MPEGB200_LUMA_BOX_ROWS 20
#endif
#ifndef MPEGB200_CHROMA_BOX_ROWS
#define MPEGB200_CHROMA_BOX_ROWS 9
#endif
constexpr int kLumaBoxRows = MPEGB200_LUMA_BOX_ROWS, kChromaBoxRows = MPEGB200_CHROMA_BOX_ROWS;

// Strip maps: the same two tensors seen as 8-byte elements (box dimensions are limited to 256 ELEMENTS, so the
// wider element buys a 304-byte wide box): one box holds every window of a group of 16 neighbouring macroblocks
// whose vectors stay within +-16 pixels (video_fused_tma.cu, strip mode).
constexpr int kStripLW = 304, kStripLH = 48;   // luma strip: 16*16 + 32 (vector range) + 1 + 15 (alignment) x 16 + 32
constexpr int kStripCW = 160, kStripCH = 24;   // chroma strip (per plane): 8*16 + 16 + 1 + 15 x 8 + 16

struct alignas(128) SlabMaps {
    unsigned char luma[128];
    unsigned char chroma[128];
    unsigned char luma_strip[128];
    unsigned char chroma_strip[128];
};

// Per-stream MP2 synthesis state in the reference's own form (audio.go:63,78).
struct AudioState {
    float v[2][1024];
    int32_t v_pos;
    int32_t open;
    int32_t pad[2];
};

// launch wrappers (video_kernels.cu / audio_kernels.cu); all return cudaGetLastError()
cudaError_t launch_fused_mc_idct(const StreamInfo* d_streams, int max_streams, const mpegb200_picture* d_pics,
                                 int n_pics, const mpegb200_mb* d_mbs, uint32_t n_mb, const int16_t* d_coeffs,
                                 uint32_t n_blocks, cudaStream_t stream, bool skip_tma_streams = false);
cudaError_t launch_rgba(const StreamInfo* d_streams, int max_streams, const int32_t* d_stream_ids,
                        const uint8_t* d_bufs, int n, int max_w, int max_h, uint8_t* d_rgba,
                        size_t rgba_stride_bytes, cudaStream_t stream);
cudaError_t launch_audio_synth(AudioState* d_states, int max_streams, const int32_t* d_stream_ids, int n_streams,
                               int frames_per_stream, const int32_t* d_samples, int format, void* d_out,
                               const float* d_window, cudaStream_t stream);
// TMA kernel (video_fused_tma.cu).  coef_map: 128-byte CUtensorMap over the coefficient array (host copy,
// passed by value as a __grid_constant__ parameter); d_maps: per-slab window tensor maps in global memory.
// d_plans: scratch of fused_plan_bytes(n_mb) bytes, 16-byte aligned (one 1024-byte plan per 16 records).
// timing: optional three events recorded before the pre-pass, between the two kernels and after the arithmetic kernel.
cudaError_t launch_fused_tma(const void* coef_map, const SlabMaps* d_maps, void* d_plans, const StreamInfo* d_streams,
                             int max_streams, const mpegb200_picture* d_pics, int n_pics, const mpegb200_mb* d_mbs,
                             uint32_t n_mb, uint32_t n_blocks, cudaStream_t stream, const cudaEvent_t* timing = nullptr);
size_t fused_plan_bytes(uint32_t n_mb);
// 12-bit packed coefficient blocks (96 B) -> int16 blocks (128 B)
cudaError_t launch_unpack12(const uint8_t* d_packed, int16_t* d_coeffs, size_t n_blocks, cudaStream_t stream);
// variable-width transfer form (coeff_vlen.cu) -> int16 blocks
cudaError_t launch_expand_vlen(const uint32_t* d_headers, const uint64_t* d_chunk_offsets, const uint8_t* d_payload,
                               int16_t* d_coeffs, size_t n_blocks, size_t payload_bytes, cudaStream_t stream);
// MP2 requantisation on the device (audio_requant.cu): info + codes -> int32 samples in the layout of mpegb200_audio_synth
cudaError_t launch_audio_requant(const mpegb200_audio_frame_info* d_info, const uint16_t* d_codes, int32_t* d_samples,
                                 size_t n_frames, cudaStream_t stream);
// slice-parallel VLC stage (vlc_slices.cu): bitstream -> records + int16 blocks in device memory, flagged pictures disabled
struct VlcDeviceTables;
struct ResidentStream {      // an elementary stream kept in device memory (mpegb200_video_stream_upload)
    const uint8_t* bytes;    // 16-byte aligned, 32 zero bytes behind the last
    uint32_t n_words;        // readable 32-bit words
    uint32_t pad;
};
size_t vlc_summary_bytes(size_t n_slices);
cudaError_t launch_startcode_index(const uint8_t* d_bytes, uint64_t len, uint64_t* d_out, uint32_t cap, uint32_t* d_count,
                                   cudaStream_t stream);
cudaError_t launch_vlc_parse(const VlcDeviceTables* d_tables, const mpegb200_vlc_picture* d_vpics, mpegb200_picture* d_pics,
                             int n_pics, const mpegb200_vlc_slice* d_slices, uint32_t n_slices, const uint8_t* d_bitstream,
                             uint32_t n_words, const uint8_t* d_quant, uint32_t n_quant, const StreamInfo* d_streams,
                             int max_streams, mpegb200_mb* d_mbs, uint32_t n_mb_slots, int16_t* d_coeffs, void* d_summary,
                             int32_t* d_flags, int sm_count, cudaStream_t stream, const ResidentStream* d_resident = nullptr);
cudaError_t configure_vlc_kernel();
cudaError_t configure_kernels();  // opt-in to large dynamic shared memory; call once per device

}  // namespace mpegb200
