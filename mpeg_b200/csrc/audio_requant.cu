// audio_requant.cu -- MP2 requantisation on the device (readSamples, audio.go:440-490, second half): sample codes as the
// bitstream has them + per (channel, subband) quantiser and scale-factor indices -> the int32 subband samples that
// audio_synth_kernel consumes.  One thread per sample; int32 arithmetic is exact here: |val| <= 2^15 and the scale
// factor (<= 2^25) enters split at 12 bits, so no product exceeds 2^28 (the reference computes in Go int = int64).
#include "common.cuh"

namespace mpegb200 {

namespace {

// quantiser levels, audio.go:955-973 (ISO 11172-3 table 3-B.4): index = quantiser number - 1
__constant__ int kLevels[17] = {3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535};
// scale factor of index i < 63: (base[i % 3] + ((1 << (i / 3)) >> 1)) >> (i / 3), audio.go:476-481 (base: audio.go:976); 63 -> 0
__constant__ int kScaleBaseDev[3] = {0x02000000, 0x01965FEA, 0x01428A30};

__global__ void __launch_bounds__(256) audio_requant_kernel(const mpegb200_audio_frame_info* __restrict__ info,
                                                           const uint16_t* __restrict__ codes, int32_t* __restrict__ samples,
                                                           size_t n_frames) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // [frame][ch][slot][sb]
    if (i >= n_frames * 2304) return;
    const size_t frame = i / 2304;
    const int r = (int)(i - frame * 2304), ch = r / 1152, slot = (r % 1152) >> 5, sb = r & 31;
    const mpegb200_audio_frame_info& fi = info[frame];
    const int qn = fi.quant[ch][sb];
    int out = 0;
    if (qn >= 1 && qn <= 17) {
        const int sfi = fi.scf[ch][sb][slot / 12];
        int sf = 0;
        if (sfi < 63) {
            const int shift = sfi / 3;
            sf = (kScaleBaseDev[sfi % 3] + ((1 << shift) >> 1)) >> shift;
        }
        const int levels = kLevels[qn - 1];
        const int scale = 65536 / (levels + 1);
        const int adj = ((levels + 1) >> 1) - 1;
        const int val = (adj - (int)codes[i]) * scale;
        out = (val * (sf >> 12) + ((val * (sf & 4095) + 2048) >> 12)) >> 12;
    }
    samples[i] = out;
}

}  // namespace

cudaError_t launch_audio_requant(const mpegb200_audio_frame_info* d_info, const uint16_t* d_codes, int32_t* d_samples,
                                 size_t n_frames, cudaStream_t stream) {
    if (n_frames == 0) return cudaSuccess;
    const size_t n = n_frames * 2304;
    audio_requant_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_info, d_codes, d_samples, n_frames);
    return cudaGetLastError();
}

}  // namespace mpegb200
