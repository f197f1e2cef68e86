// AddressSanitizer / UBSan harness for the host parser: mutates an elementary stream and walks it with
// mpegb200_video_parser_next / mpegb200_audio_parser_next.  Build + run: tools/asan_parser.sh
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../include/mpegb200_host.h"

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
    long pictures = 0, frames = 0;
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
    printf("asan harness: %d trials, %ld video steps, %ld audio frames, no report\n", trials, pictures, frames);
    return 0;
}
