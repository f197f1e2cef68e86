// AddressSanitizer / UBSan harness for the host parser: mutates an elementary stream and walks it with
// mpegb200_video_parser_next / mpegb200_audio_parser_next.  Build + run: tools/asan_parser.sh
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/mpegb200_host.h"
#include "../mpeg_b200/csrc/vlc_slice_walk.h"   // the device-side slice walker, compiled for the host: its table, record and block indices under ASan

extern "C" size_t mpegb200_internal_vlc_tables(void* out, size_t cap);

// One stream through the scan-mode batch, every wave walked by the device walker into buffers of exactly the size the
// entry point would allocate, flagged pictures re-parsed by the host, the scan withdrawn and repeated now and then.
static long scan_walk(const std::vector<uint8_t>& d, const mpegb200::VlcDeviceTables* T, std::mt19937_64& rng, int resident) {
    using namespace mpegb200;
    mpegb200_video_batch* b = mpegb200_video_batch_new(1, 1, nullptr, nullptr);
    mpegb200_video_batch_set_stream(b, 0, d.data(), d.size());
    if (resident) {
        std::vector<uint64_t> at;
        for (size_t i = 0; i + 2 < d.size(); i++)
            if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) at.push_back(i);
        mpegb200_video_batch_set_resident(b, 1);
        mpegb200_video_batch_set_start_codes(b, 0, at.data(), at.size());
    }
    int w = 0, h = 0;
    mpegb200_video_batch_stream_size(b, 0, &w, &h);
    long pictures = 0;
    if (w > 0 && h > 0 && w <= 4096 && h <= 4096) {
        VlcGeometry g;
        g.mb_w = (w + 15) >> 4;
        g.mb_h = (h + 15) >> 4;
        g.luma_w = g.mb_w * 16;
        g.luma_h = g.mb_h * 16;
        g.buf_bytes = (uint32_t)(g.luma_w * g.luma_h * 3 / 2 + g.luma_w * 16);
        mpegb200_batch_scan_step st;
        int steps = 0;
        while (mpegb200_video_batch_next_scan(b, &st) == 0 && st.has_frame[0] && steps++ < 400) {
            if (rng() % 7 == 0) {   // withdraw the scan and make it again
                mpegb200_video_batch_unscan(b);
                if (mpegb200_video_batch_next_scan(b, &st) != 0 || !st.has_frame[0]) break;
            }
            for (int wv = 0; wv < st.n_waves; wv++) {
                const mpegb200_vlc_wave& W = st.waves[wv];
                if (W.n_pictures != 1) continue;
                const uint8_t* bytes = resident ? d.data() : W.bitstream;
                const size_t n_bytes = resident ? d.size() : W.bitstream_bytes;
                const uint32_t n_words = (uint32_t)(n_bytes / 4 + 2);
                std::vector<uint32_t> words(n_words, 0);
                memcpy(words.data(), bytes, n_bytes);
                std::vector<mpegb200_mb> mbs(W.n_mb_slots ? W.n_mb_slots : 1);
                std::vector<int16_t> coeffs((W.n_mb_slots ? W.n_mb_slots : 1) * 6 * 64);
                std::vector<SliceSummary> sum(W.n_slices ? W.n_slices : 1);
                alignas(16) uint8_t scratch[128] = {0};
                for (size_t s = 0; s < W.n_slices; s++)
                    sum[s] = walk_slice(T, T->coef_fast, T->zigzag, scratch, (uint32_t)(s & 7u), W.slices[s], W.pics[W.slices[s].pic], g, words.data(),
                                        n_words, W.quant, mbs.data(), coeffs.data());
                pictures++;
                if (vlc_check_picture(W.pics[0], sum.data(), g.mb_w * g.mb_h)) {
                    mpegb200_video_step tail;
                    mpegb200_video_batch_redo(b, 0, W.step_picture[0], &tail);
                    break;   // the host finished the step
                }
            }
        }
    }
    mpegb200_video_batch_free(b);
    return pictures;
}

static std::vector<uint8_t> slurp(const char* path) {
    std::vector<uint8_t> d;
    FILE* f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s video.es audio.mp2 trials [crafted.es]\n", argv[0]); return 2; }
    const std::vector<uint8_t> video = slurp(argv[1]), audio = slurp(argv[2]);
    const int trials = atoi(argv[3]);
    std::mt19937_64 rng(12345);
    long pictures = 0, frames = 0, walked = 0;
    std::vector<uint8_t> table_bytes(sizeof(mpegb200::VlcDeviceTables));
    if (mpegb200_internal_vlc_tables(table_bytes.data(), table_bytes.size()) == 0) { fprintf(stderr, "device tables have an unexpected shape\n"); return 2; }
    const mpegb200::VlcDeviceTables* T = reinterpret_cast<const mpegb200::VlcDeviceTables*>(table_bytes.data());
    if (argc > 4) {  // the crafted stream first, unmutated
        const std::vector<uint8_t> c = slurp(argv[4]);
        mpegb200_video_parser* v = mpegb200_video_parser_new(c.data(), c.size());
        mpegb200_video_step st;
        while (mpegb200_video_parser_next(v, &st) == 0 && st.has_frame) pictures++;
        mpegb200_video_parser_free(v);
    }
    for (int t = 0; t < trials; t++) {
        std::vector<uint8_t> d(video.begin(), video.begin() + std::min<size_t>(video.size(), 12000 + rng() % 30000));
        const int flips = 1 + (int)(rng() % 80);
        for (int k = 0; k < flips; k++) d[12 + rng() % (d.size() - 12)] ^= (uint8_t)(1u << (rng() % 8));
        mpegb200_video_parser* v = mpegb200_video_parser_new(d.data(), d.size());
        mpegb200_video_step st;
        int steps = 0;
        while (mpegb200_video_parser_next(v, &st) == 0 && st.has_frame && steps++ < 400) pictures++;
        mpegb200_video_parser_free(v);
        if (t % 2 == 0) walked += scan_walk(d, T, rng, (t >> 1) & 1);
        if (t % 4 == 0) {
            std::vector<uint8_t> a(audio.begin(), audio.begin() + std::min<size_t>(audio.size(), 4000 + rng() % 20000));
            for (int k = 0; k < flips; k++) a[rng() % a.size()] ^= (uint8_t)(1u << (rng() % 8));
            mpegb200_audio_parser* p = mpegb200_audio_parser_new(a.data(), a.size());
            std::vector<int32_t> s(2 * 36 * 32);
            double tm;
            int n = 0;
            while (mpegb200_audio_parser_next(p, s.data(), &tm) && n++ < 400) frames++;
            mpegb200_audio_parser_free(p);
        }
    }
    printf("asan harness: %d trials, %ld video steps, %ld pictures through the device walker (scan mode, redo, unscan), %ld audio frames, no report\n",
           trials, pictures, walked, frames);
    return 0;
}
