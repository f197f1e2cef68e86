// vlc_emu.cpp -- TEST INFRASTRUCTURE: runs the device-side slice walker (mpeg_b200/csrc/vlc_slice_walk.h, the very code of
// vlc_parse_kernel / vlc_check_kernel) on the CPU, one slice after the other, so that the CPU test suite can compare it with
// the host parser record for record on clean and damaged streams.  Never linked into the product library.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mpeg_b200/csrc/vlc_slice_walk.h"

using namespace mpegb200;

extern "C" int vlc_emu_wave(const void* tables, int n_pictures, const mpegb200_vlc_picture* pics, size_t n_slices,
                            const mpegb200_vlc_slice* slices, const uint8_t* bitstream, size_t bitstream_bytes,
                            const uint8_t* quant, size_t n_quant, size_t n_mb_slots, const int* mb_w, const int* mb_h,
                            mpegb200_mb* mbs, int16_t* coeffs, int* flags_out) {
    const VlcDeviceTables* T = static_cast<const VlcDeviceTables*>(tables);
    const uint32_t n_words = (uint32_t)(bitstream_bytes / 4 + 2);
    std::vector<uint32_t> words(n_words + 4, 0);                      // as ctx.cu stages it: zero behind the last byte
    memcpy(words.data(), bitstream, bitstream_bytes);
    std::vector<SliceSummary> summary(n_slices ? n_slices : 1);
    alignas(16) uint8_t scratch[128];
    memset(scratch, 0, sizeof(scratch));
    for (size_t s = 0; s < n_slices; s++) {
        const mpegb200_vlc_slice& sl = slices[s];
        if (sl.pic >= (uint32_t)n_pictures || sl.mb_slot + (size_t)sl.mb_cap > n_mb_slots || pics[sl.pic].quant >= n_quant) return -1;
        const mpegb200_vlc_picture& P = pics[sl.pic];
        VlcGeometry g;
        g.mb_w = mb_w[sl.pic];
        g.mb_h = mb_h[sl.pic];
        g.luma_w = g.mb_w * 16;
        g.luma_h = g.mb_h * 16;
        g.buf_bytes = (uint32_t)(g.luma_w * g.luma_h * 3 / 2 + g.luma_w * 16);   // video.go:340
        summary[s] = walk_slice(T, T->coef_fast, T->zigzag, scratch, (uint32_t)(s & 7u), sl, P, g, words.data(), n_words, quant, mbs, coeffs);
        for (int i = 0; i < 128; i++)
            if (scratch[i]) return -2;                                 // the walker must leave its scratch zeroed
    }
    for (int p = 0; p < n_pictures; p++) flags_out[p] = (int)vlc_check_picture(pics[p], summary.data(), mb_w[p] * mb_h[p]);
    return 0;
}
