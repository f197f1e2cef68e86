// TEST INFRASTRUCTURE: the memoised start-code search of the host parser's bit reader (host_parser.cpp, BitReader::next_start_code)
// against the reference's byte-by-byte walk (buffer.go:279-302) on adversarial buffers: dense and overlapping 00 00 01 patterns,
// forward and backward jumps (hasStartCode rewinds), gaps larger than the memo tolerates, the forget threshold, the complete-index
// mode.  Includes the product source to reach its internals; built and run by tests/test_host_parser.py.
#include "../../mpeg_b200/csrc/host_parser.cpp"

#include <cstdio>
#include <random>

namespace {
struct Naive {
    int code;
    uint64_t pos;
    bool ended;
};
Naive naive_next(const uint8_t* p, size_t len, uint64_t pos_bits) {   // buffer.go:279-302, literally
    size_t i = (size_t)(((pos_bits + 7) & ~(uint64_t)7) >> 3);
    while (i + 5 <= len) {
        if (p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 1) return {p[i + 3], (uint64_t)(i + 4) << 3, false};
        i++;
    }
    return {-1, (uint64_t)i << 3, true};
}
}  // namespace

int main() {
    std::mt19937_64 rng(20260925);
    long checks = 0;
    for (int trial = 0; trial < 400; trial++) {
        const size_t len = trial < 20 ? (size_t)trial : 16 + rng() % (trial % 7 == 0 ? 300000 : 5000);
        std::vector<uint8_t> d(len + 16, 0);
        const int style = trial % 5;   // 0 random, 1 mostly zeros, 2 dense start codes, 3 sparse start codes, 4 zeros and ones only
        for (size_t i = 0; i < len; i++) {
            const uint64_t r = rng();
            d[i] = style == 0 ? (uint8_t)r : style == 1 ? (r % 11 ? 0 : (uint8_t)(r >> 8)) : style == 2 ? (uint8_t)((r % 3) == 0 ? 1 : 0)
                 : style == 3 ? (r % 97 ? (uint8_t)(1 + (r >> 8) % 255) : 0) : (uint8_t)((r >> 3) & 1);
        }
        if (style == 3)
            for (size_t i = 0; i + 4 < len; i += 1 + rng() % 4000) { d[i] = 0; d[i + 1] = 0; d[i + 2] = 1; }
        BitReader br;
        br.p = d.data();
        br.len = len;
        if (trial % 3 == 0) {   // complete index handed in (what mpegb200_video_parser_set_start_codes does)
            for (size_t i = 0; i + 5 <= len; i++)
                if (d[i] == 0 && d[i + 1] == 0 && d[i + 2] == 1) br.sc_at.push_back(i);
            br.sc_from = 0;
            br.sc_to = len;
            br.sc_complete = true;
        }
        uint64_t pos = 0;
        for (int op = 0; op < 600; op++) {
            const uint64_t r = rng();
            switch (r % 8) {
                case 0: pos = (r >> 8) % (len * 8 + 9); break;                                     // anywhere, bit granular
                case 1: pos = pos > 4000 ? pos - (r >> 8) % 4000 : 0; break;                        // a little back (hasStartCode)
                case 2: pos += (r >> 8) % 64; break;                                                // a few bits on
                case 3: pos += ((r >> 8) % 200000) * 8; if (pos > len * 8 + 8) pos = len * 8; break;  // a long jump forward
                default: break;                                                                      // go on where the last search ended
            }
            br.pos = pos;
            br.ended = false;
            const Naive want = naive_next(d.data(), len, pos);
            const int code = br.next_start_code();
            if (code != want.code || br.pos != want.pos || br.ended != want.ended) {
                fprintf(stderr, "trial %d op %d: from bit %llu got (%d, %llu, %d), want (%d, %llu, %d)\n", trial, op, (unsigned long long)pos, code,
                        (unsigned long long)br.pos, (int)br.ended, want.code, (unsigned long long)want.pos, (int)want.ended);
                return 1;
            }
            pos = br.pos;
            checks++;
        }
    }
    printf("start-code memo ok: %ld searches equal the byte-by-byte walk\n", checks);
    return 0;
}
