// coeff_pack.cpp -- host side of the variable-width transfer form ("vlen", include/mpegb200.h; the device side is
// coeff_vlen.cu): int16 blocks -> headers + payload, and the checker for streams a foreign packer produced.
//
// One pass per block: the 64 levels are gathered in zig-zag order (video.go:1044-1053), turned into codes
// c = (x + sign(x)) / 2, and every group of eight gets its width from the OR of c ^ (c >> 15) (the widest code decides).
// Eight w-bit fields are w bytes; they are assembled in a 128-bit integer and stored with one 16-byte write where the
// thread's own output range allows it.  The arithmetic over the 64 values has an AVX2 form (chosen at run time); blocks
// are spread over the host threads in whole chunks of 32; each thread packs its range into a scratch buffer, a prefix sum
// over the chunk sizes gives the offsets, and the ranges' bytes are copied to their places.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/mpegb200.h"
#include "vlen_encode.h"

namespace {

struct ZigZag {
    uint8_t nat[64];
    constexpr ZigZag() : nat() {
        int r = 0, c = 0;
        bool up = true;
        for (int p = 0; p < 64; p++) {
            nat[p] = (uint8_t)(r * 8 + c);
            if (up) {
                if (c == 7) { r++; up = false; }
                else if (r == 0) { c++; up = false; }
                else { r--; c++; }
            } else {
                if (r == 7) { c++; up = true; }
                else if (c == 0) { r++; up = true; }
                else { r++; c--; }
            }
        }
    }
};
constexpr ZigZag kZigZag;

// What one block turns into: per group the values to store (codes, or the raw levels of a group with an even value) and
// its 4-bit code.
struct Packed {
    alignas(32) int16_t v[64];   // zig-zag order
    uint32_t header;
    uint32_t bytes;
    bool ok;
};

inline uint32_t bit_length(uint32_t a) { return a ? 32u - (uint32_t)__builtin_clz(a) : 0u; }

// scalar form of the per-block arithmetic
inline void encode_scalar(const int16_t* blk, Packed& p) {
    uint32_t h = 0, total = 0;
    for (int g = 0; g < 8; g++) {
        uint32_t mag = 0, any = 0, even = 0, wide = 0;
        int16_t x[8], c[8];
        for (int i = 0; i < 8; i++) {
            x[i] = blk[kZigZag.nat[8 * g + i]];
            const int s = (x[i] > 0) - (x[i] < 0);
            c[i] = (int16_t)((x[i] + s) >> 1);           // x + sign(x) is even (or 0 / an even raw value, see below)
            mag |= (uint32_t)(uint16_t)(c[i] ^ (c[i] >> 15));
            any |= (uint32_t)(uint16_t)x[i];
            even |= (uint32_t)((x[i] & 1) == 0 && x[i] != 0);
            wide |= (uint32_t)(x[i] < -2048 || x[i] > 2047);
        }
        const uint32_t w = any ? 1u + bit_length(mag) : 0u;
        const uint32_t code = wide ? 14u : even ? 13u : w;
        h |= code << (4 * g);
        total += wide ? 16u : even ? 12u : w;
        memcpy(&p.v[8 * g], (even | wide) ? x : c, 16);
    }
    p.header = h;
    p.bytes = total;
    p.ok = true;   // every int16 value has a representation (code 14)
}

#if defined(__x86_64__)
// AVX2 form: two groups per 256-bit register
__attribute__((target("avx2"))) inline void encode_avx2(const int16_t* blk, Packed& p) {
    alignas(32) int16_t z[64];
    for (int i = 0; i < 64; i++) z[i] = blk[kZigZag.nat[i]];
    const __m256i one = _mm256_set1_epi16(1), zero = _mm256_setzero_si256();
    const __m256i lo = _mm256_set1_epi16(-2048), hi = _mm256_set1_epi16(2047);
    uint32_t h = 0, total = 0;
    for (int q = 0; q < 4; q++) {
        const __m256i x = _mm256_load_si256(reinterpret_cast<const __m256i*>(z + 16 * q));
        const __m256i sgn = _mm256_sign_epi16(one, x);                       // -1, 0, +1
        const __m256i c = _mm256_srai_epi16(_mm256_add_epi16(x, sgn), 1);
        const __m256i mag = _mm256_xor_si256(c, _mm256_srai_epi16(c, 15));
        const __m256i nz = _mm256_xor_si256(_mm256_cmpeq_epi16(x, zero), _mm256_set1_epi16(-1));
        const __m256i ev = _mm256_and_si256(_mm256_cmpeq_epi16(_mm256_and_si256(x, one), zero), nz);
        const uint32_t widem = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpgt_epi16(lo, x), _mm256_cmpgt_epi16(x, hi)));
        // horizontal OR inside each 128-bit half (one group each)
        __m256i m = mag;
        m = _mm256_or_si256(m, _mm256_srli_si256(m, 8));
        m = _mm256_or_si256(m, _mm256_srli_si256(m, 4));
        m = _mm256_or_si256(m, _mm256_srli_si256(m, 2));
        const uint32_t mag0 = (uint16_t)_mm256_extract_epi16(m, 0), mag1 = (uint16_t)_mm256_extract_epi16(m, 8);
        const uint32_t nzm = (uint32_t)_mm256_movemask_epi8(nz), evm = (uint32_t)_mm256_movemask_epi8(ev);
        const uint32_t rawm = evm | widem;
        const __m256i take_raw = _mm256_set_m128i(_mm_set1_epi16((rawm >> 16) ? -1 : 0), _mm_set1_epi16((rawm & 0xffffu) ? -1 : 0));
        _mm256_store_si256(reinterpret_cast<__m256i*>(p.v + 16 * q), _mm256_blendv_epi8(c, x, take_raw));
        for (int half = 0; half < 2; half++) {
            const uint32_t any = (nzm >> (16 * half)) & 0xffffu, even = (evm >> (16 * half)) & 0xffffu;
            const uint32_t wide = (widem >> (16 * half)) & 0xffffu;
            const uint32_t w = any ? 1u + bit_length(half ? mag1 : mag0) : 0u;
            const uint32_t code = wide ? 14u : even ? 13u : w;
            h |= code << (4 * (2 * q + half));
            total += wide ? 16u : even ? 12u : w;
        }
    }
    p.header = h;
    p.bytes = total;
    p.ok = true;
}
#endif

using EncodeFn = void (*)(const int16_t*, Packed&);
bool portable_only() {   // MPEGB200_PACK_PORTABLE=1: the plain C++ forms (tests run both)
    const char* e = getenv("MPEGB200_PACK_PORTABLE");
    return e && e[0] == '1';
}
EncodeFn pick_encoder() {
#if defined(__x86_64__)
    if (!portable_only() && __builtin_cpu_supports("avx2")) return encode_avx2;
#endif
    return encode_scalar;
}

// the bytes of one block; `room` = bytes that may be written from out on (the thread's own range)
inline uint8_t* emit(const Packed& p, uint8_t* out, size_t room) {
    for (int g = 0; g < 8; g++) {
        const uint32_t code = (p.header >> (4 * g)) & 15u;
        if (code == 0) continue;
        const uint32_t w = code == 13u ? 12u : code == 14u ? 16u : code;
        const uint32_t mask = (1u << w) - 1u;
        unsigned __int128 acc = 0;
        for (int i = 7; i >= 0; i--) acc = (acc << w) | ((uint32_t)(uint16_t)p.v[8 * g + i] & mask);
        if (room >= 16) {
            memcpy(out, &acc, 16);            // the bytes past w belong to later groups of this range and get rewritten
        } else {
            memcpy(out, &acc, w);
        }
        out += w;
        room -= w;
    }
    return out;
}

#if defined(__x86_64__)
// the same with BMI2: pext squeezes the low w bits of four 16-bit lanes into 4w contiguous bits
__attribute__((target("bmi2"))) inline uint8_t* emit_bmi2(const Packed& p, uint8_t* out, size_t room) {
    for (int g = 0; g < 8; g++) {
        const uint32_t code = (p.header >> (4 * g)) & 15u;
        if (code == 0) continue;
        const uint32_t w = code == 13u ? 12u : code == 14u ? 16u : code;
        const uint64_t lanes = 0x0001000100010001ull * ((1u << w) - 1u);
        uint64_t a, b;
        memcpy(&a, &p.v[8 * g], 8);
        memcpy(&b, &p.v[8 * g + 4], 8);
        const uint64_t lo4 = _pext_u64(a, lanes), hi4 = _pext_u64(b, lanes);   // 4w <= 48 bits each
        const unsigned __int128 acc = (unsigned __int128)lo4 | ((unsigned __int128)hi4 << (4 * w));
        if (room >= 16) {
            memcpy(out, &acc, 16);
        } else {
            memcpy(out, &acc, w);
        }
        out += w;
        room -= w;
    }
    return out;
}
#endif

using EmitFn = uint8_t* (*)(const Packed&, uint8_t*, size_t);
EmitFn pick_emitter() {
#if defined(__x86_64__)
    // pext is microcoded (hundreds of cycles) on AMD Zen 1 / Zen 2: the shift-and-or form is faster there
    if (!portable_only() && __builtin_cpu_supports("bmi2") && !__builtin_cpu_is("znver1") && !__builtin_cpu_is("znver2"))
        return emit_bmi2;
#endif
    return emit;
}

// ranges of whole chunks, one per host thread; f(range index, first block, end block)
struct Ranges {
    size_t n = 0, per = 0;
    unsigned count = 0;
    explicit Ranges(size_t n_blocks) : n(n_blocks) {
        unsigned t = std::thread::hardware_concurrency();
        if (t == 0) t = 1;
        if (t > 64) t = 64;
        if (n < 4096) t = 1;
        per = ((n + t - 1) / t + 31) / 32 * 32;
        count = per ? (unsigned)((n + per - 1) / per) : 0;
    }
    template <class F>
    void run(F f) const {
        if (count <= 1) {
            if (n) f(0u, (size_t)0, n);
            return;
        }
        std::vector<std::thread> th;
        for (unsigned i = 0; i < count; i++) {
            const size_t lo = (size_t)i * per, hi = lo + per < n ? lo + per : n;
            th.emplace_back([=] { f(i, lo, hi); });
        }
        for (auto& x : th) x.join();
    }
};

}  // namespace

namespace mpegb200 {
VlenBlock vlen_encode_block(const int16_t* blk, uint8_t* out) {
    static const EncodeFn encode = pick_encoder();
    static const EmitFn emit_block = pick_emitter();
    Packed p;
    encode(blk, p);
    emit_block(p, out, (size_t)1 << 30);   // the caller guarantees 16 bytes of slack behind the block: every group takes the 16-byte store
    return VlenBlock{p.header, p.bytes, p.ok};
}
}  // namespace mpegb200

extern "C" {

size_t mpegb200_vlen_payload_bound(size_t n_blocks) { return n_blocks * 128 + 16; }

int mpegb200_vlen_validate(const uint32_t* headers, const uint64_t* chunk_offsets, size_t n_blocks, size_t payload_bytes) {
    if (n_blocks == 0) return 0;
    if (!headers || !chunk_offsets) return MPEGB200_EINVAL;
    if (payload_bytes < 16) return MPEGB200_ERECORD;
    uint64_t run = 0;
    for (size_t b = 0; b < n_blocks; b++) {
        if (b % 32 == 0) {
            if (chunk_offsets[b / 32] != run) return MPEGB200_ERECORD;   // chunks back to back, in order
        }
        for (int g = 0; g < 8; g++) {
            const uint32_t code = (headers[b] >> (4 * g)) & 15u;
            if (code > 14u) return MPEGB200_ERECORD;
            run += code == 13u ? 12u : code == 14u ? 16u : code;
        }
    }
    return run + 16 == payload_bytes ? 0 : MPEGB200_ERECORD;
}

int mpegb200_pack_coeffs_vlen(const int16_t* coeffs, size_t n_blocks, uint32_t* headers, uint64_t* chunk_offsets,
                              uint8_t* payload, size_t payload_cap, size_t* payload_bytes) try {
    if (!payload_bytes || (n_blocks && (!coeffs || !headers || !chunk_offsets || !payload))) return MPEGB200_EINVAL;
    static const EncodeFn encode = pick_encoder();
    static const EmitFn emit_block = pick_emitter();
    const size_t chunks = (n_blocks + 31) / 32;
    const Ranges ranges(n_blocks);
    std::vector<uint32_t> chunk_bytes(chunks, 0);
    std::vector<std::vector<uint8_t>> scratch(ranges.count);   // each range's payload, until the chunk offsets are known
    std::atomic<bool> ok{true};
    // pass 1: every block once -- header, size, and its bytes into the range's own scratch
    ranges.run([&](unsigned r, size_t lo, size_t hi) {
        std::vector<uint8_t>& mine = scratch[r];
        mine.resize((hi - lo) * 128 + 16);
        uint8_t* out = mine.data();
        Packed p;
        bool good = true;
        for (size_t b = lo; b < hi; b++) {
            encode(coeffs + b * 64, p);
            good &= p.ok;
            headers[b] = p.header;
            chunk_bytes[b / 32] += p.bytes;
            out = emit_block(p, out, (size_t)(mine.data() + mine.size() - out));   // >= 16 everywhere: every group takes the single 16-byte store
        }
        if (!good) ok = false;
    });
    if (!ok) return MPEGB200_ERECORD;
    uint64_t run = 0;
    for (size_t c = 0; c < chunks; c++) {
        chunk_offsets[c] = run;
        run += chunk_bytes[c];
    }
    *payload_bytes = (size_t)run + 16;   // 16 bytes of padding: the kernel reads whole words around a group
    if (*payload_bytes > payload_cap) return MPEGB200_EINVAL;
    // pass 2: the ranges' bytes to their places (a range is a run of whole chunks, so its payload is contiguous)
    ranges.run([&](unsigned r, size_t lo, size_t hi) {
        const size_t last_chunk = (hi + 31) / 32;
        const uint64_t begin = chunk_offsets[lo / 32], end = last_chunk < chunks ? chunk_offsets[last_chunk] : run;
        memcpy(payload + begin, scratch[r].data(), (size_t)(end - begin));
    });
    memset(payload + run, 0, 16);
    return 0;
} catch (...) {  // bad_alloc / system_error (scratch vectors, threads) must not cross the C boundary
    return MPEGB200_ENOMEM;
}

}  // extern "C"
