// host_parser.cpp -- the host half of the drop-in (include/mpegb200_host.h): MPEG-1 video and MP2
// bitstream parsing, dequantisation / requantisation and PS demux on the CPU, producing the packed
// records the sm_100a kernels consume.  CPU only; no CUDA in this file.
//
// Behaviour follows the reference's serial half (citations in mpegb200_host.h and below); the
// machinery is its own: a 64-bit window bit reader and table-driven multi-bit VLC decoding instead of
// the reference's one-bit-per-step tree walk (buffer.go:352-376), records instead of pixels.
#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/mpegb200_host.h"
#include "vlc_device_tables.h"
#include "vlen_encode.h"

namespace {

// ------------------------------------------------------------------------------------------------
// bit reader (most significant bit first, buffer.go:223-255), zero bits past the end
// ------------------------------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p = nullptr;  // has >= 8 readable zero bytes after len
    size_t len = 0;
    uint64_t pos = 0;  // in bits
    bool ended = false;

    int64_t left() const { return (int64_t)(len * 8) - (int64_t)pos; }
    bool has(int64_t n) {  // buffer.go:203-221 for a fully resident source
        if (left() >= n) return true;
        ended = true;
        return false;
    }
    uint32_t peek(int n) const {  // n <= 32
        const size_t b = (size_t)(pos >> 3);
        if (b >= len) return 0;
        uint64_t w;
        memcpy(&w, p + b, 8);           // >= 8 readable bytes follow every position below len
        w = __builtin_bswap64(w);       // most significant bit first
        w <<= (pos & 7);
        return n ? (uint32_t)(w >> (64 - n)) : 0;
    }
    // the next (at least) 57 bits, left-aligned: one load serves a whole code + sign / escape sequence
    uint64_t window() const {
        const size_t b = (size_t)(pos >> 3);
        if (b >= len) return 0;
        uint64_t w;
        memcpy(&w, p + b, 8);
        return __builtin_bswap64(w) << (pos & 7);
    }
    void skip_bits(int n) { pos += (uint64_t)n; }
    uint32_t read(int n) {
        const uint32_t v = peek(n);
        pos += (uint64_t)n;
        return v;
    }
    int read1() { return (int)read(1); }
    void align() { pos = (pos + 7) & ~(uint64_t)7; }                // buffer.go:257
    void skip(int n) { if (has(n)) pos += (uint64_t)n; }            // buffer.go:261
    int skip_bytes(uint8_t v) {                                     // buffer.go:267
        align();
        int n = 0;
        while (has(8) && p[pos >> 3] == v) {
            pos += 8;
            n++;
        }
        return n;
    }
    // Start codes found so far: every 00 00 01 xx whose first byte lies in [sc_from, sc_to) is in sc_at, in order.  Video.Decode
    // looks for the next picture start code before it decodes a picture (hasStartCode, video.go:239) and the scan mode walks the
    // slice start codes of the same bytes again: with the memo every byte of the stream is searched once.
    std::vector<uint64_t> sc_at;
    uint64_t sc_from = 0, sc_to = 0;
    bool sc_complete = false;   // sc_at holds every start code of the stream (indexed elsewhere, e.g. on the device): never searched, never forgotten

    // first index f >= i with p[f..f+2] == 00 00 01 and f + 5 <= len, or len if there is none (16 bytes at a time: SSE2 is baseline x86-64)
    size_t raw_find(size_t i) const {
        if (len < 5) return len;
        const size_t last = len - 5;
#if defined(__SSE2__)
        const __m128i zero = _mm_setzero_si128(), one = _mm_set1_epi8(1);
        while (i + 18 <= len) {   // three loads of 16 bytes at i, i + 1, i + 2
            const __m128i a = _mm_loadu_si128((const __m128i*)(p + i)), b = _mm_loadu_si128((const __m128i*)(p + i + 1)),
                          c = _mm_loadu_si128((const __m128i*)(p + i + 2));
            const unsigned m = (unsigned)_mm_movemask_epi8(_mm_and_si128(_mm_and_si128(_mm_cmpeq_epi8(a, zero), _mm_cmpeq_epi8(b, zero)), _mm_cmpeq_epi8(c, one)));
            if (m) {
                const size_t f = i + (size_t)__builtin_ctz(m);
                return f <= last ? f : len;
            }
            i += 16;
        }
#endif
        for (; i <= last; i++)
            if (p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 1) return i;
        return len;
    }
    int next_start_code() {  // buffer.go:279-302
        align();
        const uint64_t i = pos >> 3;
        if (!sc_complete && (i < sc_from || i > sc_to + 65536 || sc_at.size() > 4096)) {   // outside what the memo covers (or time to forget): start over here
            sc_at.clear();
            sc_from = sc_to = i;
        }
        const auto it = std::lower_bound(sc_at.begin(), sc_at.end(), i);
        uint64_t f = len;
        if (it != sc_at.end()) {
            f = *it;
        } else {
            for (;;) {   // extend the memo until it holds a start code at or behind i (the bytes between sc_to and i are searched too)
                f = raw_find((size_t)sc_to);
                if (f >= len) {
                    sc_to = len;
                    break;
                }
                sc_at.push_back(f);
                sc_to = f + 3;   // no 00 00 01 can begin at f + 1 or f + 2
                if (f >= i) break;
            }
        }
        if (f < len) {
            pos = (f + 4) << 3;
            return p[f + 3];
        }
        // where the reference's byte-by-byte walk stops: the first position that has no room for a start code
        const uint64_t stop = len >= 4 ? (uint64_t)len - 4 : 0;
        pos = (i > stop ? i : stop) << 3;
        ended = true;
        return -1;
    }
    int find_start_code(int code) {  // buffer.go:304-311
        for (;;) {
            const int c = next_start_code();
            if (c == code || c == -1) return c;
        }
    }
    int has_start_code(int code) {  // buffer.go:313-324
        const uint64_t save = pos;
        const int c = find_start_code(code);
        pos = save;
        return c;
    }
    bool peek_non_zero(int n) { return has(n) && peek(n) != 0; }   // buffer.go:341-350
};

// ------------------------------------------------------------------------------------------------
// variable-length codes: prefix rows -> direct lookup tables
// ------------------------------------------------------------------------------------------------
#define VLC_INVALID (-32768)
struct vlc_code {
    const char* bits;
    int value;
};
#include "mpeg1_vlc_codes.inc"

struct VlcTable {
    struct Entry {
        int32_t value;
        uint8_t len;  // 0 = continue in a second-level table (value = its index)
    };
    int first_bits = 0, max_bits = 0;
    std::vector<Entry> first;
    std::vector<std::vector<Entry>> second;

    void build(const vlc_code* rows, size_t n, int first_level_bits) {
        max_bits = 0;
        for (size_t i = 0; i < n; i++) max_bits = std::max(max_bits, (int)strlen(rows[i].bits));
        first_bits = std::min(first_level_bits, max_bits);
        first.assign((size_t)1 << first_bits, Entry{0, 0xff});
        for (size_t i = 0; i < n; i++) {
            const int len = (int)strlen(rows[i].bits);
            uint32_t code = 0;
            for (int b = 0; b < len; b++) code = (code << 1) | (uint32_t)(rows[i].bits[b] == '1');
            const int32_t value = rows[i].value == VLC_INVALID ? 0 : rows[i].value;  // unassigned prefix -> 0
            if (len <= first_bits) {
                const uint32_t lo = code << (first_bits - len);
                for (uint32_t f = 0; f < (1u << (first_bits - len)); f++) first[lo + f] = Entry{value, (uint8_t)len};
            } else {
                const uint32_t head = code >> (len - first_bits);
                if (first[head].len != 0) {
                    first[head] = Entry{(int32_t)second.size(), 0};
                    second.emplace_back((size_t)1 << (max_bits - first_bits), Entry{0, 0xff});
                }
                auto& sub = second[(size_t)first[head].value];
                const int rest = len - first_bits, sub_bits = max_bits - first_bits;
                const uint32_t lo = (code & ((1u << rest) - 1)) << (sub_bits - rest);
                for (uint32_t f = 0; f < (1u << (sub_bits - rest)); f++) sub[lo + f] = Entry{value, (uint8_t)len};
            }
        }
    }
    // the code at the top of a left-aligned window (>= max_bits valid bits): value, and its length through *len
    // (max_bits for a prefix the tables of the standard do not assign, like read() below)
    int lookup(uint64_t w, int* len) const {
        const uint32_t bits = (uint32_t)(w >> (64 - max_bits));
        Entry e = first[bits >> (max_bits - first_bits)];
        if (e.len == 0) e = second[(size_t)e.value][bits & ((1u << (max_bits - first_bits)) - 1)];
        if (e.len == 0xff) {
            *len = max_bits;
            return 0;
        }
        *len = e.len;
        return e.value;
    }
    int read(BitReader& br) const {
        const uint32_t bits = br.peek(max_bits);
        Entry e = first[bits >> (max_bits - first_bits)];
        if (e.len == 0) e = second[(size_t)e.value][bits & ((1u << (max_bits - first_bits)) - 1)];
        if (e.len == 0xff) {  // not reachable with the complete tables of the standard
            br.skip_bits(max_bits);
            return 0;
        }
        br.skip_bits(e.len);
        return e.value;
    }
};

// Direct table for the DCT coefficient codes AFTER the first coefficient of a block (video.go:686-708), indexed by the next
// kCoefBits bits: code and sign bit resolved together, "10" = end of block.  len = bits consumed, 0 = not in this table
// (escape, codes longer than kCoefBits - 1, unassigned prefixes: the general two-level table takes those).
constexpr int kCoefBits = 12;
struct CoefEntry {
    int16_t level;   // signed
    uint8_t run;     // kCoefEob: end of block
    uint8_t len;
};
constexpr uint8_t kCoefEob = 0xff;

struct Tables {
    VlcTable addr_inc, type_i, type_p, type_b, cbp, motion, dc_luma, dc_chroma, coeff;
    uint8_t zigzag[64];
    CoefEntry coef_fast[1 << kCoefBits];
    void build_coef_fast() {
        for (auto& e : coef_fast) e = CoefEntry{0, 0, 0};
        auto fill = [&](uint32_t code, int len, CoefEntry e) {   // every kCoefBits-bit word that starts with the code
            e.len = (uint8_t)len;
            const uint32_t lo = code << (kCoefBits - len);
            for (uint32_t f = 0; f < (1u << (kCoefBits - len)); f++) coef_fast[lo + f] = e;
        };
        for (const vlc_code& r : VLC_DCT_COEFF) {
            if (r.value == VLC_INVALID || r.value == 0xffff) continue;
            const int len = (int)strlen(r.bits);
            uint32_t code = 0;
            for (int b = 0; b < len; b++) code = (code << 1) | (uint32_t)(r.bits[b] == '1');
            const int run = r.value >> 8, level = r.value & 0xff;
            if (len == 1) {                         // "1": "10" ends the block, "11s" is (0, +-1)
                fill(0b10, 2, CoefEntry{0, kCoefEob, 0});
                fill(0b110, 3, CoefEntry{(int16_t)level, (uint8_t)run, 0});
                fill(0b111, 3, CoefEntry{(int16_t)-level, (uint8_t)run, 0});
            } else if (len + 1 <= kCoefBits) {
                fill(code << 1, len + 1, CoefEntry{(int16_t)level, (uint8_t)run, 0});
                fill((code << 1) | 1, len + 1, CoefEntry{(int16_t)-level, (uint8_t)run, 0});
            }
        }
    }
    Tables() {
        build_coef_fast();
#define BUILD(t, rows, fb) t.build(rows, sizeof(rows) / sizeof(rows[0]), fb)
        BUILD(addr_inc, VLC_MB_ADDR_INC, 11);
        BUILD(type_i, VLC_MB_TYPE_I, 8);
        BUILD(type_p, VLC_MB_TYPE_P, 8);
        BUILD(type_b, VLC_MB_TYPE_B, 8);
        BUILD(cbp, VLC_CBP, 9);
        BUILD(motion, VLC_MOTION, 11);
        BUILD(dc_luma, VLC_DC_SIZE_LUMA, 8);
        BUILD(dc_chroma, VLC_DC_SIZE_CHROMA, 8);
        BUILD(coeff, VLC_DCT_COEFF, 9);
#undef BUILD
        // zig-zag scan (video.go:1044-1053): walk the anti-diagonals
        int i = 0;
        for (int d = 0; d < 15; d++) {
            const int lo = d < 8 ? 0 : d - 7, hi = d < 8 ? d : 7;
            for (int t = lo; t <= hi; t++) {
                const int r = (d & 1) ? t : d - t, c = d - r;  // odd diagonals run downwards
                zigzag[i++] = (uint8_t)(r * 8 + c);
            }
        }
    }
};
const Tables& tables() {
    static const Tables t;
    return t;
}

// A single-level table as the device reads it: value | length << 16 per index.
bool flatten(const VlcTable& t, int bits, uint32_t* out) {
    if (t.max_bits != bits || t.first_bits != bits || !t.second.empty()) return false;
    for (size_t i = 0; i < ((size_t)1 << bits); i++) {
        const VlcTable::Entry e = t.first[i];
        if (e.len == 0 || e.len == 0xff || e.value < -32768 || e.value > 32767) return false;
        out[i] = ((uint32_t)e.value & 0xffffu) | ((uint32_t)e.len << 16);
    }
    return true;
}

const uint8_t kIntraQuant[64] = {  // ISO 11172-2 default intra matrix (video.go:1055-1064)
    8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83};
const double kPictureRate[16] = {0.000, 23.976, 24.000, 25.000, 29.970, 30.000, 50.000, 59.940,
                                 60.000, 0.000,  0.000,  0.000,  0.000,  0.000,  0.000,  0.000};

enum { kPicI = 1, kPicP = 2, kPicB = 3 };
enum { kStartPicture = 0x00, kSliceFirst = 0x01, kSliceLast = 0xAF, kUserData = 0xB2, kSequence = 0xB3, kExtension = 0xB5 };

struct Motion {
    int full_px = 0, r_size = 0, h = 0, v = 0;
    bool is_set = false;
};

}  // namespace

// The tables above in the flat form of the device-side slice parser (vlc_device_tables.h).
bool mpegb200::fill_vlc_device_tables(mpegb200::VlcDeviceTables* out) {
    using namespace mpegb200;
    const Tables& t = tables();
    memset(out, 0, sizeof(*out));
    static_assert(kVlcCoefFastBits == kCoefBits && sizeof(CoefEntry) == 4, "fast coefficient table layout");
    for (size_t i = 0; i < ((size_t)1 << kCoefBits); i++) {
        const CoefEntry e = t.coef_fast[i];
        out->coef_fast[i] = ((uint32_t)(uint16_t)e.level) | ((uint32_t)e.run << 16) | ((uint32_t)e.len << 24);
    }
    const VlcTable& c = t.coeff;
    if (c.first_bits != kVlcCoefFirstBits || c.max_bits != kVlcCoefFirstBits + kVlcCoefSecondBits ||
        c.second.size() > (size_t)kVlcCoefSecondTables)
        return false;
    auto entry = [](const VlcTable::Entry& e, uint32_t* o) {
        if (e.len == 0xff || e.value < 0 || e.value > 0xffff) return false;
        *o = e.len == 0 ? ((uint32_t)e.value | kVlcLink) : ((uint32_t)e.value | ((uint32_t)e.len << 16));
        return true;
    };
    for (size_t i = 0; i < c.first.size(); i++)
        if (!entry(c.first[i], &out->coeff_first[i])) return false;
    for (size_t k = 0; k < c.second.size(); k++)
        for (size_t i = 0; i < c.second[k].size(); i++)
            if (c.second[k][i].len == 0 || !entry(c.second[k][i], &out->coeff_second[k][i])) return false;
    memcpy(out->zigzag, t.zigzag, 64);
    return flatten(t.addr_inc, kVlcAddrIncBits, out->addr_inc) && flatten(t.motion, kVlcMotionBits, out->motion) &&
           flatten(t.cbp, kVlcCbpBits, out->cbp) && flatten(t.type_i, kVlcTypeIBits, out->type_i) &&
           flatten(t.type_p, kVlcTypePBits, out->type_p) && flatten(t.type_b, kVlcTypeBBits, out->type_b) &&
           flatten(t.dc_luma, kVlcDcLumaBits, out->dc_luma) && flatten(t.dc_chroma, kVlcDcChromaBits, out->dc_chroma);
}

// ------------------------------------------------------------------------------------------------
// video
// ------------------------------------------------------------------------------------------------
struct mpegb200_video_parser {
    std::vector<uint8_t> data;
    BitReader br;
    double frame_rate = 0, time = 0;
    int frames_decoded = 0;
    int width = 0, height = 0, mb_w = 0, mb_h = 0, mb_size = 0;
    int start_code = -1, picture_type = 0;
    Motion fwd, bwd;
    bool has_header = false, has_reference = false, no_delay = false;
    int quantizer_scale = 0, mb_addr = 0, mb_row = 0, mb_col = 0;
    bool slice_begin = false, intra = false;
    int dc_pred[3] = {128, 128, 128};
    int cur = 0, fwd_buf = 1, bwd_buf = 2;  // physical buffers playing frameCurrent / Forward / Backward
    uint8_t intra_q[64], non_intra_q[64];
    int32_t level[64];  // the reference's blockData before the premultiply; persists like it (video.go:101)

    // records of the step being built
    std::vector<mpegb200_launch> launches;
    std::vector<mpegb200_mb> mbs;
    std::vector<int16_t> coeffs;
    // the same in the variable-width transfer form (vlen mode: `coeffs` stays empty)
    bool vlen = false, vlen_next = false;        // vlen_next: what mpegb200_video_parser_set_vlen asked for (applies per picture)
    std::vector<mpegb200_launch_vlen> vl_launches;
    std::vector<uint32_t> vl_headers;
    std::vector<uint64_t> vl_chunks;
    std::vector<uint8_t> vl_payload;
    // picture under construction
    std::vector<mpegb200_mb> pic_mbs;
    std::vector<int16_t> pic_coeffs;
    std::vector<uint32_t> pic_headers;           // vlen mode: one header per coded block of the picture ...
    std::vector<uint64_t> pic_block_at;          // ... where its bytes start in pic_payload (n + 1 entries) ...
    std::vector<uint8_t> pic_payload;            // ... and the bytes (16 bytes of slack behind the last block)
    size_t pic_blocks = 0;                       // coded blocks of the picture so far (both modes)
    // scan mode (mpegb200_video_parser_next_scan): pictures and slice start codes of the step, and where to resume a host re-parse
    // everything the serial walk carries from picture to picture, as it is when decode_picture is entered
    struct Saved {
        uint64_t br_pos;       // the reader's position and end flag (its start-code memo stays: it describes the bytes, not the walk)
        bool br_ended;
        int start_code, picture_type, cur, fwd_buf, bwd_buf;
        bool has_reference;
        Motion fwd, bwd;
    };
    int step_frames_decoded = 0;       // the step counters in front of the current scan step (a re-parse restarts the step's tail)
    double step_time = 0;
    // One step back (mpegb200_video_parser_unscan): the caller scans step k + 1 while the device still works on step k; if step k
    // then flags a picture, step k + 1 is withdrawn and step k's saved states are current again.
    struct StepBegin {
        bool valid = false;
        Saved at;
        int frames_decoded = 0;
        double time = 0;
        int32_t level[64];
        std::vector<Saved> prev_saved;
        int prev_step_frames_decoded = 0;
        double prev_step_time = 0;
    } step_begin;
    mpegb200_video_step host_step;     // scan mode: a step the host had to parse itself (stale coefficients pending, see next_scan)
    std::vector<mpegb200_scan_picture> scan_pics;
    std::vector<mpegb200_scan_slice> scan_slices;
    std::vector<Saved> scan_saved;
    uint8_t quant128[128];
    std::vector<int32_t> last_writer;  // per macroblock address: index into pic_mbs, -1 = none
    bool pic_has_rewrites = false;
    int rec = -1;

    void reset_levels() { memset(level, 0, sizeof(level)); }
};

namespace {

using VP = mpegb200_video_parser;

bool decode_sequence_header(VP* v) {  // video.go:270-331
    BitReader& br = v->br;
    if (!br.has(64 + 2 * 64 * 8)) return false;
    v->width = (int)br.read(12);
    v->height = (int)br.read(12);
    if (v->width <= 0 || v->height <= 0) return false;
    br.read(4);  // aspect ratio
    v->frame_rate = kPictureRate[br.read(4)];
    br.read(18);  // bit rate
    br.skip(1 + 10 + 1);
    const Tables& t = tables();
    if (br.read1()) {
        for (int i = 0; i < 64; i++) v->intra_q[t.zigzag[i]] = (uint8_t)br.read(8);
    } else {
        memcpy(v->intra_q, kIntraQuant, 64);
    }
    if (br.read1()) {
        for (int i = 0; i < 64; i++) v->non_intra_q[t.zigzag[i]] = (uint8_t)br.read(8);
    } else {
        memset(v->non_intra_q, 16, 64);
    }
    v->mb_w = (v->width + 15) >> 4;
    v->mb_h = (v->height + 15) >> 4;
    v->mb_size = v->mb_w * v->mb_h;
    v->cur = 0;      // initFrame x3 (video.go:324-326): current, forward, backward = buffers 0, 1, 2
    v->fwd_buf = 1;
    v->bwd_buf = 2;
    v->has_header = true;
    return true;
}

bool ensure_header(VP* v) {  // video.go:130-147
    if (v->has_header) return true;
    if (v->start_code != kSequence) v->start_code = v->br.find_start_code(kSequence);
    if (v->start_code == -1) return false;
    return decode_sequence_header(v);
}

int decode_motion_vector(VP* v, int r_size, int motion) {  // video.go:583-606
    const int fscale = 1 << r_size;
    const int m_code = tables().motion.read(v->br);
    int d;
    if (m_code != 0 && fscale != 1) {
        const int r = (int)v->br.read(r_size);
        d = ((abs(m_code) - 1) << r_size) + r + 1;
        if (m_code < 0) d = -d;
    } else {
        d = m_code;
    }
    motion += d;
    if (motion > (fscale << 4) - 1)
        motion -= fscale << 5;
    else if (motion < -(fscale << 4))
        motion += fscale << 5;
    return motion;
}

// A new record for the macroblock at (mb_row, mb_col); remembers rewrites (SURVEY: serial semantics).
mpegb200_mb& new_record(VP* v) {
    mpegb200_mb m;
    memset(&m, 0, sizeof(m));
    m.mb_row = (uint16_t)v->mb_row;
    m.mb_col = (uint16_t)v->mb_col;
    m.coeff_block = (uint32_t)v->pic_blocks;
    v->rec = (int)v->pic_mbs.size();
    v->pic_mbs.push_back(m);
    const int addr = v->mb_row * v->mb_w + v->mb_col;
    if (v->last_writer[addr] >= 0) v->pic_has_rewrites = true;
    v->last_writer[addr] = v->rec;
    return v->pic_mbs.back();
}

// predictMacroblock (video.go:608-637), decision only: which reference, which vector.
void pack_prediction(VP* v, mpegb200_mb& m) {
    int h = v->fwd.h, w = v->fwd.v;
    if (v->fwd.full_px) {
        h *= 2;
        w *= 2;
    }
    bool use_bwd = false;
    if (v->picture_type == kPicB) {
        // forward copy, then -- if set -- the backward copy on top of it: only the latter is observable
        if (!v->fwd.is_set || v->bwd.is_set) {
            use_bwd = true;
            h = v->bwd.h;
            w = v->bwd.v;
            if (v->bwd.full_px) {
                h *= 2;
                w *= 2;
            }
        }
    }
    m.flags |= MPEGB200_MB_PREDICT | (use_bwd ? MPEGB200_MB_REF_BWD : 0);
    m.mv_h = (int16_t)h;
    m.mv_v = (int16_t)w;
}

void decode_block(VP* v, int block) {  // video.go:639-799, up to the hand-over to the kernels
    BitReader& br = v->br;
    const Tables& t = tables();
    int n = 0;
    const uint8_t* q;
    if (v->intra) {
        const int plane = block > 3 ? block - 3 : 0;
        const int size = (plane == 0 ? t.dc_luma : t.dc_chroma).read(br);
        int dc = v->dc_pred[plane];
        if (size > 0) {
            const int diff = (int)br.read(size);
            dc += (diff & (1 << (size - 1))) ? diff : (-(1 << size) | (diff + 1));
        }
        v->dc_pred[plane] = dc;
        v->level[0] = dc * 8;  // dc << 8 == (dc * 8) * premultiplier[0] (32), video.go:672
        q = v->intra_q;
        n = 1;
    } else {
        q = v->non_intra_q;
    }
    for (;;) {
        // one window per coefficient: the code (at most 17 bits), then its sign bit or the 6 + 8 (+ 8) bits of an escape
        uint64_t w = br.window();
        int used, run, lv;
        const CoefEntry fe = t.coef_fast[w >> (64 - kCoefBits)];
        if (fe.len != 0 && n > 0) {   // the common case: a short code behind the first coefficient
            if (fe.run == kCoefEob) {
                br.skip_bits(fe.len);
                break;
            }
            run = fe.run;
            lv = fe.level;
            used = fe.len;
        } else {
            const int c = t.coeff.lookup(w, &used);
            w <<= used;
            if (c == 0x0001 && n > 0) {   // "1" after the first coefficient: a 0 bit behind it ends the block (video.go:686) ...
                if ((w >> 63) == 0) {
                    br.skip_bits(used + 1);
                    break;
                }
                w <<= 1;                  // ... a 1 bit is consumed, and the sign follows like after any other code
                used += 1;
            }
            if (c == 0xffff) {  // escape
                run = (int)(w >> 58);
                lv = (int)((w >> 50) & 0xff);
                used += 14;
                if (lv == 0) {
                    lv = (int)((w >> 42) & 0xff);
                    used += 8;
                } else if (lv == 128) {
                    lv = (int)((w >> 42) & 0xff) - 256;
                    used += 8;
                } else if (lv > 128) {
                    lv -= 256;
                }
            } else {
                run = c >> 8;
                lv = c & 0xff;
                if (w >> 63) lv = -lv;
                used += 1;
            }
        }
        br.skip_bits(used);
        n += run;
        if (n < 0 || n >= 64) return;  // invalid run: the block is dropped, its coefficients stay (video.go:712-714)
        const int dz = t.zigzag[n++];
        lv *= 2;  // dequantise, oddify, clip: video.go:719-741
        if (!v->intra) lv += lv < 0 ? -1 : 1;
        lv = (lv * v->quantizer_scale * (int)q[dz]) >> 4;
        if ((lv & 1) == 0) lv -= lv > 0 ? 1 : -1;
        if (lv > 2047) lv = 2047;
        if (lv < -2048) lv = -2048;
        v->level[dz] = lv;
    }
    // hand-over: what idct + copy/add*ToDest would consume (video.go:772-798)
    alignas(32) int16_t scratch[64];
    int16_t* out = scratch;
    if (!v->vlen) {
        const size_t at = v->pic_coeffs.size();
        v->pic_coeffs.resize(at + 64, 0);
        out = &v->pic_coeffs[at];
    } else {
        memset(scratch, 0, sizeof(scratch));
    }
    auto put = [&](int i) {
        const int32_t l = v->level[i];
        out[i] = (int16_t)(l > 32767 ? 32767 : (l < -32768 ? -32768 : l));
    };
    if (n == 1) {  // DC only: the rest of the array is ignored and survives (video.go:774-777)
        put(0);
        v->level[0] = 0;
    } else {
        if (n < 10) {  // sparse transform: rows/cols 0..3 only (video.go:807-866)
            for (int r = 0; r < 4; r++)
                for (int c2 = 0; c2 < 4; c2++) put(r * 8 + c2);
        } else {
            for (int i = 0; i < 64; i++) put(i);
        }
        v->reset_levels();
    }
    if (v->vlen) {  // the block goes out as variable-width groups: header + a few bytes instead of 128
        const size_t at = (size_t)v->pic_block_at.back();
        if (v->pic_payload.size() < at + 128 + 16) v->pic_payload.resize(std::max(v->pic_payload.size() * 2, at + 4096));
        const mpegb200::VlenBlock b = mpegb200::vlen_encode_block(scratch, v->pic_payload.data() + at);
        v->pic_headers.push_back(b.header);
        v->pic_block_at.push_back(at + b.bytes);
    }
    v->pic_blocks++;
    v->pic_mbs[(size_t)v->rec].cbp |= (uint8_t)(0x20 >> block);
}

void decode_macroblock(VP* v) {  // video.go:462-562
    BitReader& br = v->br;
    const Tables& t = tables();
    int inc = 0, code = t.addr_inc.read(br);
    while (code == 34) code = t.addr_inc.read(br);  // stuffing
    while (code == 35) {                            // escape
        inc += 33;
        code = t.addr_inc.read(br);
    }
    inc += code;
    if (v->slice_begin) {
        v->slice_begin = false;
        v->mb_addr += inc;
    } else {
        if (v->mb_addr + inc >= v->mb_size) return;
        if (inc > 1) {
            v->dc_pred[0] = v->dc_pred[1] = v->dc_pred[2] = 128;
            if (v->picture_type == kPicP) v->fwd.h = v->fwd.v = 0;
        }
        while (inc > 1) {  // skipped macroblocks are pure predictions
            v->mb_addr++;
            v->mb_row = v->mb_addr / v->mb_w;
            v->mb_col = v->mb_addr % v->mb_w;
            pack_prediction(v, new_record(v));
            inc--;
        }
        v->mb_addr++;
    }
    v->mb_row = v->mb_addr / v->mb_w;
    v->mb_col = v->mb_addr % v->mb_w;
    // mb_addr < 0: slice 1 whose first address increment is an unassigned code (value 0) leaves the address at -1;
    // the reference indexes out of range there and panics, here the macroblock is dropped
    if (v->mb_addr < 0 || v->mb_col >= v->mb_w || v->mb_row >= v->mb_h) return;

    const VlcTable& tt = v->picture_type == kPicI ? t.type_i : (v->picture_type == kPicP ? t.type_p : t.type_b);
    const int type = tt.read(br);
    v->intra = type & 0x01;
    v->fwd.is_set = type & 0x08;
    v->bwd.is_set = type & 0x04;
    if (type & 0x10) v->quantizer_scale = (int)br.read(5);

    mpegb200_mb& m = new_record(v);
    if (v->intra) {
        v->fwd.h = v->fwd.v = v->bwd.h = v->bwd.v = 0;
        m.flags |= MPEGB200_MB_INTRA;
    } else {
        v->dc_pred[0] = v->dc_pred[1] = v->dc_pred[2] = 128;
        if (v->fwd.is_set) {  // decodeMotionVectors, video.go:564-581
            v->fwd.h = decode_motion_vector(v, v->fwd.r_size, v->fwd.h);
            v->fwd.v = decode_motion_vector(v, v->fwd.r_size, v->fwd.v);
        } else if (v->picture_type == kPicP) {
            v->fwd.h = v->fwd.v = 0;
        }
        if (v->bwd.is_set) {
            v->bwd.h = decode_motion_vector(v, v->bwd.r_size, v->bwd.h);
            v->bwd.v = decode_motion_vector(v, v->bwd.r_size, v->bwd.v);
        }
        pack_prediction(v, m);
    }
    int cbp = 0;
    if (type & 0x02)
        cbp = t.cbp.read(br);
    else if (v->intra)
        cbp = 0x3f;
    for (int block = 0; block < 6; block++)
        if (cbp & (0x20 >> block)) decode_block(v, block);
}

void decode_slice(VP* v, int slice) {  // video.go:436-460
    BitReader& br = v->br;
    v->slice_begin = true;
    v->mb_addr = (slice - 1) * v->mb_w - 1;
    v->fwd.h = v->fwd.v = v->bwd.h = v->bwd.v = 0;
    v->dc_pred[0] = v->dc_pred[1] = v->dc_pred[2] = 128;
    v->quantizer_scale = (int)br.read(5);
    while (br.read1()) br.skip(8);
    do {
        decode_macroblock(v);
    } while (v->mb_addr < v->mb_size - 1 && br.peek_non_zero(23));
}

// Close the picture: split it into launches without double writes and append them to the step.
void emit_picture(VP* v, int type, int dst, int fwd, int bwd) {
    const size_t n = v->pic_mbs.size();
    std::vector<uint8_t> dead(n, 0);
    std::vector<size_t> cuts;  // record indices where a new wave starts
    if (v->pic_has_rewrites) {
        std::vector<int32_t> seen((size_t)v->mb_size, -1);
        for (size_t i = 0; i < n; i++) {
            const mpegb200_mb& m = v->pic_mbs[i];
            const size_t addr = (size_t)m.mb_row * v->mb_w + m.mb_col;
            if (seen[addr] >= 0) {
                const bool complete = (m.flags & MPEGB200_MB_PREDICT) || ((m.flags & MPEGB200_MB_INTRA) && m.cbp == 0x3f);
                if (complete) {
                    dead[(size_t)seen[addr]] = 1;  // the later record defines every pixel: it simply wins
                } else {
                    cuts.push_back(i);  // partial rewrite must see the earlier result: next launch
                    std::fill(seen.begin(), seen.end(), -1);
                }
            }
            seen[addr] = (int32_t)i;
        }
    }
    cuts.push_back(n);
    size_t begin = 0;
    for (size_t cut : cuts) {
        mpegb200_launch L;
        memset(&L, 0, sizeof(L));
        L.picture.stream = 0;
        L.picture.type = (uint8_t)type;
        L.picture.dst_buf = (uint8_t)dst;
        L.picture.fwd_buf = (uint8_t)fwd;
        L.picture.bwd_buf = (uint8_t)bwd;
        L.first_mb = (uint32_t)v->mbs.size();
        L.first_block = v->vlen ? (uint32_t)v->vl_headers.size() : (uint32_t)(v->coeffs.size() / 64);
        mpegb200_launch_vlen LV;
        memset(&LV, 0, sizeof(LV));
        LV.first_chunk = (uint32_t)v->vl_chunks.size();
        LV.payload_offset = v->vl_payload.size();
        uint32_t blocks = 0;
        for (size_t i = begin; i < cut; i++) {
            if (dead[i]) continue;
            mpegb200_mb m = v->pic_mbs[i];
            const int nc = __builtin_popcount(m.cbp);
            if (!v->vlen) {
                const int16_t* src = &v->pic_coeffs[(size_t)m.coeff_block * 64];
                v->coeffs.insert(v->coeffs.end(), src, src + (size_t)nc * 64);
            } else {
                for (int k = 0; k < nc; k++) {  // a chunk offset at every 32nd block of the launch
                    const size_t pb = (size_t)m.coeff_block + (size_t)k;
                    if (((blocks + (uint32_t)k) & 31u) == 0) v->vl_chunks.push_back(v->vl_payload.size() - LV.payload_offset);
                    v->vl_headers.push_back(v->pic_headers[pb]);
                    v->vl_payload.insert(v->vl_payload.end(), v->pic_payload.begin() + (ptrdiff_t)v->pic_block_at[pb],
                                         v->pic_payload.begin() + (ptrdiff_t)v->pic_block_at[pb + 1]);
                }
            }
            m.coeff_block = blocks;
            m.pic = 0;
            v->mbs.push_back(m);
            blocks += (uint32_t)nc;
        }
        L.n_mb = (uint32_t)v->mbs.size() - L.first_mb;
        L.n_blocks = blocks;
        L.picture.first_mb = 0;
        L.picture.n_mb = L.n_mb;
        if (L.n_mb || cut == n) {
            v->launches.push_back(L);
            if (v->vlen) {
                v->vl_payload.insert(v->vl_payload.end(), 16, 0);   // the padding the device expansion may read
                LV.payload_bytes = v->vl_payload.size() - LV.payload_offset;
                v->vl_launches.push_back(LV);
            }
        }
        begin = cut;
    }
}

void decode_picture(VP* v, bool scan = false) {  // video.go:374-434
    BitReader& br = v->br;
    const VP::Saved at_entry{br.pos, br.ended, v->start_code, v->picture_type, v->cur, v->fwd_buf, v->bwd_buf, v->has_reference, v->fwd, v->bwd};
    const uint64_t begin_byte = (br.pos >> 3) - 4;   // the picture start code
    br.skip(10);
    v->picture_type = (int)br.read(3);
    br.skip(16);
    if (v->picture_type <= 0 || v->picture_type > kPicB) return;
    if (v->picture_type == kPicP || v->picture_type == kPicB) {
        v->fwd.full_px = br.read1();
        const int f = (int)br.read(3);
        if (f == 0) return;
        v->fwd.r_size = f - 1;
    }
    if (v->picture_type == kPicB) {
        v->bwd.full_px = br.read1();
        const int f = (int)br.read(3);
        if (f == 0) return;
        v->bwd.r_size = f - 1;
    }
    const int temp = v->fwd_buf;  // rotation by index instead of by struct copy (video.go:406-409)
    if (v->picture_type == kPicI || v->picture_type == kPicP) v->fwd_buf = v->bwd_buf;

    if (scan) {
        // Headers and start codes only: the slices stay unparsed (the device walks them, one thread per slice).  For a stream
        // whose slices are what the reference expects this consumes exactly the bytes the full parse consumes.
        mpegb200_scan_picture P;
        memset(&P, 0, sizeof(P));
        P.type = (uint8_t)v->picture_type;
        P.dst_buf = (uint8_t)v->cur;
        P.fwd_buf = (uint8_t)v->fwd_buf;
        P.bwd_buf = (uint8_t)v->bwd_buf;
        P.fwd_full_px = (uint8_t)v->fwd.full_px;
        P.fwd_r_size = (uint8_t)v->fwd.r_size;
        P.bwd_full_px = (uint8_t)v->bwd.full_px;
        P.bwd_r_size = (uint8_t)v->bwd.r_size;
        P.first_slice = (uint32_t)v->scan_slices.size();
        P.begin = begin_byte;
        do {
            v->start_code = br.next_start_code();
        } while (v->start_code == kExtension || v->start_code == kUserData);
        while (v->start_code >= kSliceFirst && v->start_code <= kSliceLast) {
            mpegb200_scan_slice S;
            memset(&S, 0, sizeof(S));
            S.offset = br.pos >> 3;
            S.vpos = (uint32_t)(v->start_code & 0xff);
            v->start_code = br.next_start_code();
            S.next_code = v->start_code == -1 ? (uint64_t)br.len : (br.pos >> 3) - 4;
            v->scan_slices.push_back(S);
        }
        P.n_slices = (uint32_t)v->scan_slices.size() - P.first_slice;
        const uint64_t end = br.pos >> 3;            // behind the start code that ended the slices (or the end of the stream)
        P.end = end < br.len ? end : br.len;
        v->scan_pics.push_back(P);
        v->scan_saved.push_back(at_entry);
        if (v->picture_type == kPicI || v->picture_type == kPicP) {  // video.go:430-433
            v->bwd_buf = v->cur;
            v->cur = temp;
        }
        return;
    }

    v->pic_mbs.clear();
    v->pic_coeffs.clear();
    v->vlen = v->vlen_next;
    v->pic_headers.clear();
    v->pic_block_at.assign(1, 0);
    v->pic_blocks = 0;
    v->last_writer.assign((size_t)v->mb_size, -1);
    v->pic_has_rewrites = false;
    v->rec = -1;

    do {
        v->start_code = br.next_start_code();
    } while (v->start_code == kExtension || v->start_code == kUserData);
    while (v->start_code >= kSliceFirst && v->start_code <= kSliceLast) {
        decode_slice(v, v->start_code & 0xff);
        if (v->mb_addr >= v->mb_size - 2) break;
        v->start_code = br.next_start_code();
    }
    emit_picture(v, v->picture_type, v->cur, v->fwd_buf, v->bwd_buf);

    if (v->picture_type == kPicI || v->picture_type == kPicP) {  // video.go:430-433
        v->bwd_buf = v->cur;
        v->cur = temp;
    }
}

}  // namespace

extern "C" {

mpegb200_video_parser* mpegb200_video_parser_new(const uint8_t* data, size_t len) {
    if (!data && len) return nullptr;
    auto* v = new (std::nothrow) mpegb200_video_parser();
    if (!v) return nullptr;
    try {
        v->data.assign(data, data + len);
        v->data.resize(len + 16, 0);
    } catch (...) {  // bad_alloc must not cross the C boundary
        delete v;
        return nullptr;
    }
    v->br.p = v->data.data();
    v->br.len = len;
    v->reset_levels();
    v->start_code = v->br.find_start_code(kSequence);  // NewVideo, video.go:114-118
    if (v->start_code != -1) decode_sequence_header(v);
    return v;
}

void mpegb200_video_parser_free(mpegb200_video_parser* v) { delete v; }
int mpegb200_video_parser_has_header(mpegb200_video_parser* v) { return v && ensure_header(v); }
int mpegb200_video_parser_width(mpegb200_video_parser* v) { return v && ensure_header(v) ? v->width : 0; }
int mpegb200_video_parser_height(mpegb200_video_parser* v) { return v && ensure_header(v) ? v->height : 0; }
double mpegb200_video_parser_framerate(mpegb200_video_parser* v) { return v && ensure_header(v) ? v->frame_rate : 0; }
void mpegb200_video_parser_set_no_delay(mpegb200_video_parser* v, int no_delay) {
    if (v) v->no_delay = no_delay != 0;
}
int mpegb200_video_parser_has_ended(mpegb200_video_parser* v) { return v ? v->br.ended : 1; }
void mpegb200_video_parser_set_vlen(mpegb200_video_parser* v, int on) {
    if (v) v->vlen_next = on != 0;
}

void mpegb200_video_parser_rewind(mpegb200_video_parser* v) {  // video.go:195-201
    if (!v) return;
    v->br.pos = 0;
    v->br.ended = false;
    v->time = 0;
    v->frames_decoded = 0;
    v->has_reference = false;
    v->start_code = -1;
}

// resume >= 0: the tail of the last scan step, from its picture `resume` on, parsed in full (mpegb200_video_parser_redo)
static int video_parser_next_impl(mpegb200_video_parser* v, mpegb200_video_step* out, bool scan = false, int resume = -1) {  // Video.Decode, video.go:209-268
    memset(out, 0, sizeof(*out));
    bool resuming = resume >= 0;
    if (resuming) {
        const VP::Saved sv = v->scan_saved[(size_t)resume];
        v->br.pos = sv.br_pos;
        v->br.ended = sv.br_ended;
        v->start_code = sv.start_code;
        v->picture_type = sv.picture_type;
        v->cur = sv.cur;
        v->fwd_buf = sv.fwd_buf;
        v->bwd_buf = sv.bwd_buf;
        v->has_reference = sv.has_reference;
        v->fwd = sv.fwd;
        v->bwd = sv.bwd;
        v->frames_decoded = v->step_frames_decoded;
        v->time = v->step_time;
    } else {
        v->step_frames_decoded = v->frames_decoded;
        v->step_time = v->time;
    }
    v->scan_pics.clear();
    v->scan_slices.clear();
    v->scan_saved.clear();
    v->launches.clear();
    v->mbs.clear();
    v->coeffs.clear();
    v->vl_launches.clear();
    v->vl_headers.clear();
    v->vl_chunks.clear();
    v->vl_payload.clear();
    if (!ensure_header(v)) return 0;
    int frame = -1;
    for (;;) {
        if (!resuming) {
            if (v->start_code != kStartPicture) {
                v->start_code = v->br.find_start_code(kStartPicture);
                if (v->start_code == -1) {
                    if (v->has_reference && !v->no_delay && v->br.ended && (v->picture_type == kPicI || v->picture_type == kPicP)) {
                        v->has_reference = false;  // flush the last reference frame (video.go:223-229)
                        frame = v->bwd_buf;
                        break;
                    }
                    return 0;
                }
            }
            if (v->br.has_start_code(kStartPicture) == -1 && !v->br.ended) return 0;
        }
        resuming = false;
        decode_picture(v, scan);
        if (v->no_delay)
            frame = v->bwd_buf;
        else if (v->picture_type == kPicB)
            frame = v->cur;
        else if (v->has_reference)
            frame = v->fwd_buf;
        else
            v->has_reference = true;
        if (frame >= 0) break;
    }
    out->has_frame = 1;
    out->frame_buf = frame;
    out->time = v->time;
    v->frames_decoded++;
    v->time = (double)v->frames_decoded / v->frame_rate;
    out->n_launches = (int)v->launches.size();
    out->launches = v->launches.data();
    out->mbs = v->mbs.data();
    // a step is in one form throughout: the mode switches between steps only (set_vlen is applied at a picture start,
    // and the caller flips it between two parser_next calls)
    if (!v->vl_launches.empty() && v->vl_launches.size() == v->launches.size()) {
        out->coeffs = nullptr;
        out->vlen_launches = v->vl_launches.data();
        out->vlen_headers = v->vl_headers.data();
        out->vlen_chunk_offsets = v->vl_chunks.data();
        out->vlen_payload = v->vl_payload.data();
    } else {
        out->coeffs = v->coeffs.data();
    }
    return 0;
}

// The record vectors grow while parsing; an allocation failure must not unwind through the C boundary (cgo, ctypes)
// nor reach std::terminate on a pool thread: it is reported as MPEGB200_ENOMEM and the step is empty.
int mpegb200_video_parser_next(mpegb200_video_parser* v, mpegb200_video_step* out) {
    if (!v || !out) return MPEGB200_EINVAL;
    try {
        return video_parser_next_impl(v, out);
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

int mpegb200_video_parser_next_scan(mpegb200_video_parser* v, mpegb200_video_scan_step* out) {
    if (!v || !out) return MPEGB200_EINVAL;
    memset(out, 0, sizeof(*out));
    try {
        // Coefficients a dropped block left behind (video.go:712-714) leak into the next block the serial reference decodes,
        // possibly a picture later.  Only the host parser carries that state: while it is pending the host parses the step itself.
        bool stale = false;
        for (int i = 0; i < 64 && !stale; i++) stale = v->level[i] != 0;
        {   // what mpegb200_video_parser_unscan goes back to
            auto& sb = v->step_begin;
            sb.valid = true;
            sb.at = VP::Saved{v->br.pos, v->br.ended, v->start_code, v->picture_type, v->cur, v->fwd_buf, v->bwd_buf, v->has_reference, v->fwd, v->bwd};
            sb.frames_decoded = v->frames_decoded;
            sb.time = v->time;
            memcpy(sb.level, v->level, sizeof(sb.level));
            sb.prev_saved.swap(v->scan_saved);          // next_impl clears scan_saved: the previous step's states move here first
            sb.prev_step_frames_decoded = v->step_frames_decoded;
            sb.prev_step_time = v->step_time;
        }
        mpegb200_video_step st;
        const int rc = video_parser_next_impl(v, stale ? &v->host_step : &st, !stale);
        if (stale) st = v->host_step;
        if (rc != 0 || !st.has_frame) return rc;
        memcpy(v->quant128, v->intra_q, 64);
        memcpy(v->quant128 + 64, v->non_intra_q, 64);
        out->has_frame = 1;
        out->frame_buf = st.frame_buf;
        out->time = st.time;
        out->n_pictures = (int)v->scan_pics.size();
        out->pictures = v->scan_pics.data();
        out->slices = v->scan_slices.data();
        out->stream = v->data.data();
        out->stream_len = v->br.len;
        out->quant = v->quant128;
        out->mb_w = v->mb_w;
        out->mb_h = v->mb_h;
        out->host_step = stale ? &v->host_step : nullptr;
        return 0;
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

int mpegb200_video_parser_set_start_codes(mpegb200_video_parser* v, const uint64_t* positions, size_t n) {
    if (!v || (n && !positions)) return MPEGB200_EINVAL;
    try {
        BitReader& br = v->br;
        br.sc_at.clear();
        br.sc_at.reserve(n);
        uint64_t prev = 0;
        for (size_t k = 0; k < n; k++) {
            const uint64_t f = positions[k];
            if (k && f <= prev) return MPEGB200_EINVAL;                      // ascending, please
            prev = f;
            if (br.len < 5 || f > (uint64_t)br.len - 5) continue;             // no room for the code byte and one more (buffer.go:284)
            if (br.p[f] != 0 || br.p[f + 1] != 0 || br.p[f + 2] != 1) return MPEGB200_EINVAL;   // not a start code of THIS stream
            br.sc_at.push_back(f);
        }
        br.sc_from = 0;
        br.sc_to = br.len;
        br.sc_complete = true;
        return 0;
    } catch (...) {
        return MPEGB200_ENOMEM;
    }
}

int mpegb200_video_parser_unscan(mpegb200_video_parser* v) {
    if (!v || !v->step_begin.valid) return MPEGB200_EINVAL;
    auto& sb = v->step_begin;
    v->br.pos = sb.at.br_pos;
    v->br.ended = sb.at.br_ended;
    v->start_code = sb.at.start_code;
    v->picture_type = sb.at.picture_type;
    v->cur = sb.at.cur;
    v->fwd_buf = sb.at.fwd_buf;
    v->bwd_buf = sb.at.bwd_buf;
    v->has_reference = sb.at.has_reference;
    v->fwd = sb.at.fwd;
    v->bwd = sb.at.bwd;
    v->frames_decoded = sb.frames_decoded;
    v->time = sb.time;
    memcpy(v->level, sb.level, sizeof(sb.level));
    v->scan_saved.swap(sb.prev_saved);
    v->step_frames_decoded = sb.prev_step_frames_decoded;
    v->step_time = sb.prev_step_time;
    v->scan_pics.clear();
    v->scan_slices.clear();
    sb.valid = false;   // one step back, not two
    return 0;
}

int mpegb200_video_parser_redo(mpegb200_video_parser* v, int k, mpegb200_video_step* out) {
    if (!v || !out || k < 0 || k >= (int)v->scan_saved.size()) return MPEGB200_EINVAL;
    try {
        return video_parser_next_impl(v, out, false, k);
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// MP2 (audio.go:184-490): header, allocation, scale factors, requantised samples
// ------------------------------------------------------------------------------------------------
namespace {

struct Quantizer {
    uint16_t levels;
    uint8_t group, bits;
};
// ISO 11172-3 Annex B tables 3-B.2a-d in kjmp2's compact form (audio.go:901-973)
const Quantizer kQuant[17] = {{3, 1, 5},     {5, 1, 7},     {7, 0, 3},      {9, 1, 10},     {15, 0, 4},     {31, 0, 5},
                              {63, 0, 6},    {127, 0, 7},   {255, 0, 8},    {511, 0, 9},    {1023, 0, 10},  {2047, 0, 11},
                              {4095, 0, 12}, {8191, 0, 13}, {16383, 0, 14}, {32767, 0, 15}, {65535, 0, 16}};
const uint8_t kRateClass[2][14] = {{0, 0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2}, {0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2}};
const uint8_t kTablePick[3][3] = {{8, 8, 12}, {27 | 64, 27 | 64, 27 | 64}, {30 | 64, 27 | 64, 30 | 64}};
const uint8_t kSbAlloc[2][30] = {
    {0x44, 0x44, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34, 0x34},
    {0x43, 0x43, 0x43, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x42, 0x31, 0x31, 0x31, 0x31,
     0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x31, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20, 0x20}};
const uint8_t kAllocRows[6][16] = {{0, 1, 2, 17},
                                   {0, 1, 2, 3, 4, 5, 6, 17},
                                   {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 17},
                                   {0, 1, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17},
                                   {0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16},
                                   {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}};
const int kSampleRate[4] = {44100, 48000, 32000, 0};
const int kBitRate[14] = {32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384};
const int64_t kScaleBase[3] = {0x02000000, 0x01965FEA, 0x01428A30};
enum { kModeStereo = 0, kModeJoint = 1, kModeDual = 2, kModeMono = 3 };

}  // namespace

struct mpegb200_audio_parser {
    std::vector<uint8_t> data;
    BitReader br;
    double time = 0;
    int samples_decoded = 0, samplerate_index = 3, bitrate_index = 0, mode = 0, channels = 0, bound = 0;
    int next_frame_data_size = 0;
    bool has_header = false;
    const Quantizer* alloc[2][32] = {};
    uint8_t scfsi[2][32] = {};
    int64_t scale[2][32][3] = {};
};

namespace {

using AP = mpegb200_audio_parser;

bool find_frame_sync(BitReader& br) {  // buffer.go:326-339
    size_t i;
    for (i = (size_t)(br.pos >> 3); i + 1 < br.len; i++)
        if (br.p[i] == 0xFF && (br.p[i + 1] & 0xFE) == 0xFC) {
            br.pos = ((uint64_t)(i + 1) << 3) + 3;
            return true;
        }
    br.pos = (uint64_t)(i + 1) << 3;
    return false;
}

int decode_header(AP* a) {  // audio.go:184-272
    BitReader& br = a->br;
    if (!br.has(48)) return 0;
    br.skip_bytes(0x00);
    const int sync = (int)br.read(11);
    if (sync != 0x7ff && !find_frame_sync(br)) return 0;
    const int version = (int)br.read(2), layer = (int)br.read(2);
    const bool has_crc = br.read1() == 0;
    if (version != 3 || layer != 2) return 0;  // MPEG-1, layer II
    const int bitrate_index = (int)br.read(4) - 1;
    if (bitrate_index > 13) return 0;
    const int samplerate_index = (int)br.read(2);
    if (samplerate_index == 3) return 0;
    const int padding = br.read1();
    br.skip(1);
    const int mode = (int)br.read(2);
    if (a->has_header && (a->bitrate_index != bitrate_index || a->samplerate_index != samplerate_index || a->mode != mode))
        return 0;
    a->bitrate_index = bitrate_index;
    a->samplerate_index = samplerate_index;
    a->mode = mode;
    a->has_header = true;
    if (mode == kModeStereo || mode == kModeJoint)
        a->channels = 2;
    else if (mode == kModeMono)
        a->channels = 1;
    if (mode == kModeJoint) {
        a->bound = ((int)br.read(2) + 1) << 2;
    } else {
        br.skip(2);
        a->bound = mode == kModeMono ? 0 : 32;
    }
    br.skip(4);
    if (has_crc) br.skip(16);
    if (bitrate_index < 0) return 0;  // "free format": the Go code would index out of range
    const int frame_size = 144000 * kBitRate[bitrate_index] / kSampleRate[samplerate_index] + padding;
    return frame_size - (has_crc ? 6 : 4);
}

const Quantizer* read_allocation(AP* a, int sb, int tab3) {  // audio.go:429-438
    const int tab4 = kSbAlloc[tab3][sb];
    const int qtab = kAllocRows[tab4 & 15][a->br.read(tab4 >> 4)];
    return qtab ? &kQuant[qtab - 1] : nullptr;
}

// readSamples (audio.go:440-490), first half: the three sample codes of one subband as the bitstream has them (degrouped)
bool read_codes(AP* a, int ch, int sb, int64_t out[3]) {
    const Quantizer* q = a->alloc[ch][sb];
    if (!q) {
        out[0] = out[1] = out[2] = 0;
        return false;
    }
    const int64_t adj = q->levels;
    if (q->group) {
        int64_t val = a->br.read(q->bits);
        out[0] = val % adj;
        val /= adj;
        out[1] = val % adj;
        out[2] = val / adj;
    } else {
        out[0] = a->br.read(q->bits);
        out[1] = a->br.read(q->bits);
        out[2] = a->br.read(q->bits);
    }
    return true;
}

// ... second half: requantisation (audio.go:476-489)
void requantise(const AP* a, int ch, int sb, int part, int64_t out[3]) {
    const Quantizer* q = a->alloc[ch][sb];
    int64_t sf = a->scale[ch][sb][part];
    if (sf == 63) {
        sf = 0;
    } else {
        const int shift = (int)(sf / 3);
        sf = (kScaleBase[sf % 3] + (((int64_t)1 << shift) >> 1)) >> shift;
    }
    int64_t adj = q->levels;
    const int64_t scale = 65536 / (adj + 1);
    adj = ((adj + 1) >> 1) - 1;
    for (int i = 0; i < 3; i++) {
        const int64_t val = (adj - out[i]) * scale;
        out[i] = (val * (sf >> 12) + ((val * (sf & 4095) + 2048) >> 12)) >> 12;
    }
}

void read_samples(AP* a, int ch, int sb, int part, int64_t out[3]) {
    if (read_codes(a, ch, sb, out)) requantise(a, ch, sb, part, out);
}

// samples != nullptr: requantised samples (the layout of mpegb200_audio_synth); else info + codes (mpegb200_audio_synth_coded)
void decode_frame(AP* a, int32_t* samples, mpegb200_audio_frame_info* info = nullptr, uint16_t* codes = nullptr) {  // audio.go:274-375
    BitReader& br = a->br;
    const int tab2 = kRateClass[a->mode == kModeMono ? 0 : 1][a->bitrate_index];
    int tab3 = kTablePick[tab2][a->samplerate_index];
    const int sblimit = tab3 & 63;
    tab3 >>= 6;
    if (a->bound > sblimit) a->bound = sblimit;
    for (int sb = 0; sb < a->bound; sb++) {
        a->alloc[0][sb] = read_allocation(a, sb, tab3);
        a->alloc[1][sb] = read_allocation(a, sb, tab3);
    }
    for (int sb = a->bound; sb < sblimit; sb++) a->alloc[1][sb] = a->alloc[0][sb] = read_allocation(a, sb, tab3);
    const int channels = a->mode == kModeMono ? 1 : 2;
    for (int sb = 0; sb < sblimit; sb++) {
        for (int ch = 0; ch < channels; ch++)
            if (a->alloc[ch][sb]) a->scfsi[ch][sb] = (uint8_t)br.read(2);
        if (a->mode == kModeMono) a->scfsi[1][sb] = a->scfsi[0][sb];
    }
    for (int sb = 0; sb < sblimit; sb++) {
        for (int ch = 0; ch < channels; ch++) {
            if (!a->alloc[ch][sb]) continue;
            int64_t* s = a->scale[ch][sb];
            switch (a->scfsi[ch][sb]) {  // audio.go:322-342
                case 0: s[0] = br.read(6); s[1] = br.read(6); s[2] = br.read(6); break;
                case 1: s[0] = s[1] = br.read(6); s[2] = br.read(6); break;
                case 2: s[0] = s[1] = s[2] = br.read(6); break;
                default: s[0] = br.read(6); s[1] = s[2] = br.read(6); break;
            }
        }
        if (a->mode == kModeMono)
            for (int i = 0; i < 3; i++) a->scale[1][sb][i] = a->scale[0][sb][i];
    }
    if (samples) {
        memset(samples, 0, sizeof(int32_t) * 2 * 36 * 32);  // subbands >= sblimit stay zero (audio.go:368-375)
    } else {
        memset(codes, 0, sizeof(uint16_t) * 2 * 36 * 32);
        memset(info, 0, sizeof(*info));
        for (int sb = 0; sb < sblimit; sb++)
            for (int ch = 0; ch < 2; ch++) {
                // above the joint-stereo bound (and for mono) channel 1 mirrors channel 0's requantised samples
                // (audio.go:362-367): it gets channel 0's quantiser AND scale factors, so the device computes the same values
                const int src = sb < a->bound ? ch : 0;
                const Quantizer* q = a->alloc[src][sb];
                info->quant[ch][sb] = q ? (uint8_t)(q - kQuant + 1) : 0;
                for (int part = 0; part < 3; part++) info->scf[ch][sb][part] = (uint8_t)a->scale[src][sb][part];
            }
    }
    for (int part = 0; part < 3; part++)
        for (int granule = 0; granule < 4; granule++) {
            const int step0 = 3 * (4 * part + granule);
            int64_t s0[3], s1[3];
            for (int sb = 0; sb < sblimit; sb++) {
                if (samples) {
                    read_samples(a, 0, sb, part, s0);
                    if (sb < a->bound)
                        read_samples(a, 1, sb, part, s1);
                    else
                        memcpy(s1, s0, sizeof(s0));  // joint/mono: channel 1 mirrors channel 0 (audio.go:362-367)
                    for (int p = 0; p < 3; p++) {
                        samples[(0 * 36 + step0 + p) * 32 + sb] = (int32_t)s0[p];
                        samples[(1 * 36 + step0 + p) * 32 + sb] = (int32_t)s1[p];
                    }
                } else {
                    read_codes(a, 0, sb, s0);
                    if (sb < a->bound)
                        read_codes(a, 1, sb, s1);
                    else
                        memcpy(s1, s0, sizeof(s0));
                    for (int p = 0; p < 3; p++) {
                        codes[(0 * 36 + step0 + p) * 32 + sb] = (uint16_t)s0[p];
                        codes[(1 * 36 + step0 + p) * 32 + sb] = (uint16_t)s1[p];
                    }
                }
            }
        }
    br.align();
}

}  // namespace

extern "C" {

mpegb200_audio_parser* mpegb200_audio_parser_new(const uint8_t* data, size_t len) {
    if (!data && len) return nullptr;
    auto* a = new (std::nothrow) mpegb200_audio_parser();
    if (!a) return nullptr;
    try {
        a->data.assign(data, data + len);
        a->data.resize(len + 16, 0);
    } catch (...) {
        delete a;
        return nullptr;
    }
    a->br.p = a->data.data();
    a->br.len = len;
    a->next_frame_data_size = decode_header(a);
    return a;
}
void mpegb200_audio_parser_free(mpegb200_audio_parser* a) { delete a; }
int mpegb200_audio_parser_has_header(mpegb200_audio_parser* a) {
    if (!a) return 0;
    if (a->has_header) return 1;
    a->next_frame_data_size = decode_header(a);
    return a->has_header;
}
int mpegb200_audio_parser_samplerate(mpegb200_audio_parser* a) {
    return mpegb200_audio_parser_has_header(a) ? kSampleRate[a->samplerate_index] : 0;
}
int mpegb200_audio_parser_channels(mpegb200_audio_parser* a) { return a ? a->channels : 0; }
void mpegb200_audio_parser_rewind(mpegb200_audio_parser* a) {  // audio.go:149-154
    if (!a) return;
    a->br.pos = 0;
    a->br.ended = false;
    a->time = 0;
    a->samples_decoded = 0;
    a->next_frame_data_size = 0;
}
static int audio_next(mpegb200_audio_parser* a, int32_t* samples, mpegb200_audio_frame_info* info, uint16_t* codes, double* time) {
    if (a->next_frame_data_size == 0) a->next_frame_data_size = decode_header(a);
    if (a->next_frame_data_size == 0 || !a->br.has((int64_t)a->next_frame_data_size << 3)) return 0;
    decode_frame(a, samples, info, codes);
    a->next_frame_data_size = 0;
    if (time) *time = a->time;
    a->samples_decoded += MPEGB200_SAMPLES_PER_FRAME;
    a->time = (double)a->samples_decoded / (double)kSampleRate[a->samplerate_index];
    return 1;
}
int mpegb200_audio_parser_next(mpegb200_audio_parser* a, int32_t* samples, double* time) {  // audio.go:163-182
    if (!a || !samples) return 0;
    return audio_next(a, samples, nullptr, nullptr, time);
}
int mpegb200_audio_parser_next_coded(mpegb200_audio_parser* a, mpegb200_audio_frame_info* info, uint16_t* codes, double* time) {
    if (!a || !info || !codes) return 0;
    return audio_next(a, nullptr, info, codes, time);
}

// ------------------------------------------------------------------------------------------------
// program stream (demux.go): pack + system header, then PES packets of the wanted ids
// ------------------------------------------------------------------------------------------------
static int demux_split_impl(const uint8_t* data, size_t len, uint8_t** video, size_t* video_len, uint8_t** audio,
                            size_t* audio_len, int* n_video_packets, int* n_audio_packets) {
    std::vector<uint8_t> padded(data, data + len);
    padded.resize(len + 16, 0);
    BitReader br;
    br.p = padded.data();
    br.len = len;
    auto skip_clock = [&]() {  // decodeTime, demux.go:520-529
        br.read(3); br.skip(1); br.read(15); br.skip(1); br.read(15); br.skip(1);
    };
    if (br.find_start_code(0xBA) == -1 || !br.has(64) || br.read(4) != 0x02) return MPEGB200_EINVAL;  // demux.go:91-113
    skip_clock();
    br.skip(1); br.skip(22); br.skip(1);
    if (br.find_start_code(0xBB) == -1 || !br.has(56)) return MPEGB200_EINVAL;                        // demux.go:116-133
    br.skip(16); br.skip(24); br.read(6); br.skip(5); br.read(5);
    std::vector<uint8_t> v, a;
    int nv = 0, na = 0;
    for (;;) {  // demux.go:500-510
        const int code = br.next_start_code();
        if (code == -1) break;
        if (!(code == 0xE0 || code == 0xBD || (code >= 0xC0 && code <= 0xC3))) continue;
        if (!br.has(16 << 3)) break;  // decodePacket, demux.go:531-568
        int64_t length = br.read(16);
        length -= br.skip_bytes(0xff);
        if (br.read(2) == 0x01) {
            br.skip(16);
            length -= 2;
        }
        const int marker = (int)br.read(2);
        if (marker == 0x03) {
            skip_clock();
            br.skip(40);
            length -= 10;
        } else if (marker == 0x02) {
            skip_clock();
            length -= 5;
        } else if (marker == 0x00) {
            br.skip(4);
            length -= 1;
        } else {
            continue;
        }
        if (length < 0 || !br.has(length << 3)) break;
        const uint8_t* payload = padded.data() + (br.pos >> 3);
        if (code == 0xE0) {
            v.insert(v.end(), payload, payload + length);
            nv++;
        } else if (code == 0xC0) {
            a.insert(a.end(), payload, payload + length);
            na++;
        }
        br.pos += (uint64_t)length << 3;
    }
    *video = (uint8_t*)malloc(v.size() ? v.size() : 1);
    *audio = (uint8_t*)malloc(a.size() ? a.size() : 1);
    if (!*video || !*audio) {
        free(*video);
        free(*audio);
        *video = *audio = nullptr;
        return MPEGB200_ENOMEM;
    }
    memcpy(*video, v.data(), v.size());
    memcpy(*audio, a.data(), a.size());
    *video_len = v.size();
    *audio_len = a.size();
    if (n_video_packets) *n_video_packets = nv;
    if (n_audio_packets) *n_audio_packets = na;
    return 0;
}

int mpegb200_demux_split(const uint8_t* data, size_t len, uint8_t** video, size_t* video_len, uint8_t** audio,
                         size_t* audio_len, int* n_video_packets, int* n_audio_packets) {
    if (!data || !video || !video_len || !audio || !audio_len) return MPEGB200_EINVAL;
    *video = *audio = nullptr;
    try {
        return demux_split_impl(data, len, video, video_len, audio, audio_len, n_video_packets, n_audio_packets);
    } catch (...) {
        return MPEGB200_ENOMEM;
    }
}

void mpegb200_buffer_free(void* p) { free(p); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Many streams in lock-step: parse one Decode() step of every stream on a pool of host threads and
// merge the per-stream launches into waves (one kernel launch each, at most one picture per stream).
// ------------------------------------------------------------------------------------------------
#include <atomic>
#include <memory>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {

// Worker threads that live as long as the batch: a step hands them its two loops (parse every stream, merge every
// stream's records into the waves) instead of creating and joining a set of threads twice per step -- with 64 threads
// that was a couple of milliseconds per step, as much as the parsing of 4096 small pictures itself.
class WorkerPool {
public:
    explicit WorkerPool(int threads) {
        for (int k = 1; k < threads; k++) workers_.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_job_.notify_all();
        for (auto& t : workers_) t.join();
    }
    bool failed() { return failed_.exchange(false); }
    // f(i) for i in [0, n), the calling thread takes part; returns when all are done
    void run(int n, const std::function<void(int)>& f) {
        if (workers_.empty() || n <= 1) {
            next_.store(0);
            work(f, n);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            n_ = n;
            next_.store(0);
            active_ = (int)workers_.size();
            gen_++;
        }
        cv_job_.notify_all();
        work(f, n);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return active_ == 0; });
        job_ = nullptr;
    }

private:
    void work(const std::function<void(int)>& f, int n) {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n) break;
            try {
                f(i);
            } catch (...) {  // an exception on a pool thread would be std::terminate; the caller asks failed()
                failed_.store(true);
            }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* f;
            int n;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_job_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                f = job_;
                n = n_;
            }
            work(*f, n);
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) cv_done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_job_, cv_done_;
    const std::function<void(int)>* job_ = nullptr;
    int n_ = 0, active_ = 0;
    std::atomic<int> next_{0};
    std::atomic<bool> failed_{false};
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace

struct mpegb200_video_batch {
    std::unique_ptr<WorkerPool> pool;
    int n = 0, threads = 1;
    void* (*alloc)(size_t) = nullptr;
    void (*free_fn)(void*) = nullptr;
    std::vector<mpegb200_video_parser*> parsers;
    std::vector<mpegb200_video_step> steps;
    std::vector<int> has_frame, frame_buf;
    std::vector<double> time;
    struct WaveBuf {
        mpegb200_picture* pics = nullptr;
        mpegb200_mb* mbs = nullptr;
        int16_t* coeffs = nullptr;
        uint32_t* headers = nullptr;       // vlen mode
        uint64_t* chunks = nullptr;
        uint8_t* payload = nullptr;
        size_t cap_pics = 0, cap_mbs = 0, cap_blocks = 0, cap_headers = 0, cap_chunks = 0, cap_payload = 0;
    };
    bool vlen = false;
    std::vector<WaveBuf> bufs[2];   // double-buffered: the previous step's arrays stay intact while the next is parsed
    int flip = 0;
    std::vector<mpegb200_wave> waves;
    // scan mode (mpegb200_video_batch_next_scan): slice tables and compressed bytes instead of records
    struct ScanBuf {
        mpegb200_vlc_picture* pics = nullptr;
        mpegb200_vlc_slice* slices = nullptr;
        uint8_t* bits = nullptr;
        uint8_t* quant = nullptr;
        int32_t* step_picture = nullptr;
        size_t cap_pics = 0, cap_slices = 0, cap_bits = 0, cap_quant = 0, cap_step = 0;
    };
    std::vector<mpegb200_video_scan_step> scan_steps;
    std::vector<ScanBuf> scan_bufs[2];
    std::vector<mpegb200_vlc_wave> scan_waves[2];   // the descriptors are double-buffered like the arrays they point to
    std::vector<int> host_index;
    std::vector<mpegb200_video_step> host_steps;
    bool resident = false;          // the streams live in device memory: the waves carry tables only

    void* get(size_t bytes) { return alloc ? alloc(bytes ? bytes : 1) : malloc(bytes ? bytes : 1); }
    void put(void* p) {
        if (!p) return;
        if (free_fn) free_fn(p); else free(p);
    }
    template <typename T>
    bool reserve(T*& p, size_t& cap, size_t need) {
        if (need <= cap) return true;
        put(p);
        cap = need + need / 4 + 64;
        p = (T*)get(cap * sizeof(T));
        return p != nullptr;
    }
};

extern "C" {

mpegb200_video_batch* mpegb200_video_batch_new(int n_streams, int threads, void* (*alloc)(size_t), void (*free_fn)(void*)) {
    if (n_streams <= 0 || n_streams > 65536) return nullptr;  // 16-bit picture index per launch
    auto* b = new (std::nothrow) mpegb200_video_batch();
    if (!b) return nullptr;
    try {
        b->n = n_streams;
        b->threads = threads > 0 ? threads : 1;
        b->pool.reset(new WorkerPool(b->threads));   // std::thread may throw system_error
        b->alloc = alloc;
        b->free_fn = free_fn;
        b->parsers.assign((size_t)n_streams, nullptr);
        b->steps.resize((size_t)n_streams);
        b->has_frame.assign((size_t)n_streams, 0);
        b->frame_buf.assign((size_t)n_streams, 0);
        b->time.assign((size_t)n_streams, 0.0);
    } catch (...) {
        delete b;
        return nullptr;
    }
    return b;
}

void mpegb200_video_batch_free(mpegb200_video_batch* b) {
    if (!b) return;
    for (auto* p : b->parsers) delete p;
    for (auto& set : b->bufs)
        for (auto& w : set) {
            b->put(w.pics);
            b->put(w.mbs);
            b->put(w.coeffs);
            b->put(w.headers);
            b->put(w.chunks);
            b->put(w.payload);
        }
    for (auto& set : b->scan_bufs)
        for (auto& w : set) {
            b->put(w.pics);
            b->put(w.slices);
            b->put(w.bits);
            b->put(w.quant);
            b->put(w.step_picture);
        }
    delete b;
}

int mpegb200_video_batch_set_stream(mpegb200_video_batch* b, int index, const uint8_t* data, size_t len) {
    if (!b || index < 0 || index >= b->n) return MPEGB200_EINVAL;
    delete b->parsers[(size_t)index];
    b->parsers[(size_t)index] = mpegb200_video_parser_new(data, len);
    mpegb200_video_parser_set_vlen(b->parsers[(size_t)index], b->vlen);
    return b->parsers[(size_t)index] ? 0 : MPEGB200_ENOMEM;
}

int mpegb200_video_batch_set_vlen(mpegb200_video_batch* b, int on) {
    if (!b) return MPEGB200_EINVAL;
    b->vlen = on != 0;
    for (auto* p : b->parsers) mpegb200_video_parser_set_vlen(p, on);
    return 0;
}

int mpegb200_video_batch_stream_size(mpegb200_video_batch* b, int index, int* width, int* height) {
    if (!b || index < 0 || index >= b->n || !b->parsers[(size_t)index]) return MPEGB200_EINVAL;
    if (width) *width = mpegb200_video_parser_width(b->parsers[(size_t)index]);
    if (height) *height = mpegb200_video_parser_height(b->parsers[(size_t)index]);
    return 0;
}

static int video_batch_next_impl(mpegb200_video_batch* b, mpegb200_batch_step* out) {
    memset(out, 0, sizeof(*out));
    const int n = b->n;
    b->pool->run(n, [&](int i) {
        mpegb200_video_step& st = b->steps[(size_t)i];
        memset(&st, 0, sizeof(st));
        if (b->parsers[(size_t)i] && mpegb200_video_parser_next(b->parsers[(size_t)i], &st) != 0) throw std::bad_alloc();
        b->has_frame[(size_t)i] = st.has_frame;
        b->frame_buf[(size_t)i] = st.frame_buf;
        b->time[(size_t)i] = st.time;
    });
    if (b->pool->failed()) return MPEGB200_ENOMEM;
    int n_waves = 0;
    for (int i = 0; i < n; i++) n_waves = std::max(n_waves, b->steps[(size_t)i].has_frame ? b->steps[(size_t)i].n_launches : 0);
    b->flip ^= 1;
    auto& bufs = b->bufs[b->flip];
    if ((int)bufs.size() < n_waves) bufs.resize((size_t)n_waves);
    b->waves.assign((size_t)n_waves, mpegb200_wave{});
    std::vector<uint32_t> pic_of((size_t)n), mb_off((size_t)n), blk_off((size_t)n);
    std::vector<uint64_t> byte_off((size_t)n);
    for (int w = 0; w < n_waves; w++) {
        // sizes and offsets of wave w (streams that have a w-th launch with work in it)
        uint32_t np = 0, nm = 0, nb = 0;
        uint64_t nbytes = 0;
        bool wave_vlen = false;
        for (int i = 0; i < n; i++) {
            const mpegb200_video_step& st = b->steps[(size_t)i];
            pic_of[(size_t)i] = 0xffffffffu;
            if (!st.has_frame || w >= st.n_launches || st.launches[w].n_mb == 0) continue;
            pic_of[(size_t)i] = np++;
            mb_off[(size_t)i] = nm;
            blk_off[(size_t)i] = nb;
            byte_off[(size_t)i] = nbytes;
            nm += st.launches[w].n_mb;
            nb += st.launches[w].n_blocks;
            if (st.vlen_launches) {
                wave_vlen = true;
                nbytes += st.vlen_launches[w].payload_bytes - 16;   // the per-launch padding is dropped, the wave gets its own
            }
        }
        auto& buf = bufs[(size_t)w];
        if (!b->reserve(buf.pics, buf.cap_pics, np) || !b->reserve(buf.mbs, buf.cap_mbs, nm)) return MPEGB200_ENOMEM;
        if (wave_vlen) {
            if (!b->reserve(buf.headers, buf.cap_headers, (size_t)nb) || !b->reserve(buf.chunks, buf.cap_chunks, ((size_t)nb + 31) / 32) ||
                !b->reserve(buf.payload, buf.cap_payload, (size_t)nbytes + 16))
                return MPEGB200_ENOMEM;
        } else if (!b->reserve(buf.coeffs, buf.cap_blocks, (size_t)nb * 64)) {
            return MPEGB200_ENOMEM;
        }
        b->pool->run(n, [&](int i) {
            const uint32_t p = pic_of[(size_t)i];
            if (p == 0xffffffffu) return;
            const mpegb200_video_step& st = b->steps[(size_t)i];
            const mpegb200_launch& L = st.launches[w];
            mpegb200_picture pic = L.picture;
            pic.stream = i;
            pic.first_mb = mb_off[(size_t)i];
            pic.n_mb = L.n_mb;
            buf.pics[p] = pic;
            const mpegb200_mb* src = st.mbs + L.first_mb;
            mpegb200_mb* dst = buf.mbs + mb_off[(size_t)i];
            for (uint32_t k = 0; k < L.n_mb; k++) {
                dst[k] = src[k];
                dst[k].pic = (uint16_t)p;
                dst[k].coeff_block += blk_off[(size_t)i];
            }
            if (!wave_vlen) {
                memcpy(buf.coeffs + (size_t)blk_off[(size_t)i] * 64, st.coeffs + (size_t)L.first_block * 64, (size_t)L.n_blocks * 128);
                return;
            }
            // variable-width form: headers and payload move as they are; the wave's chunk offsets (one per 32 blocks of the
            // merged numbering) come from a running sum over this stream's headers
            const mpegb200_launch_vlen& LV = st.vlen_launches[w];
            const uint32_t* hs = st.vlen_headers + L.first_block;
            memcpy(buf.headers + blk_off[(size_t)i], hs, (size_t)L.n_blocks * 4);
            memcpy(buf.payload + byte_off[(size_t)i], st.vlen_payload + LV.payload_offset, (size_t)(LV.payload_bytes - 16));
            uint64_t at = byte_off[(size_t)i];
            for (uint32_t k = 0; k < L.n_blocks; k++) {
                const uint32_t gbl = blk_off[(size_t)i] + k;
                if ((gbl & 31u) == 0) buf.chunks[gbl >> 5] = at;
                at += mpegb200::vlen_block_bytes(hs[k]);
            }
        });
        mpegb200_wave& W = b->waves[(size_t)w];
        W.n_pictures = (int)np;
        W.pics = buf.pics;
        W.n_mb = nm;
        W.mbs = buf.mbs;
        W.n_blocks = nb;
        if (wave_vlen) {
            memset(buf.payload + nbytes, 0, 16);
            W.coeffs = nullptr;
            W.vlen_headers = buf.headers;
            W.vlen_chunk_offsets = buf.chunks;
            W.vlen_payload = buf.payload;
            W.vlen_payload_bytes = (size_t)nbytes + 16;
        } else {
            W.coeffs = buf.coeffs;
        }
    }
    out->n_streams = n;
    out->has_frame = b->has_frame.data();
    out->frame_buf = b->frame_buf.data();
    out->time = b->time.data();
    out->n_waves = n_waves;
    out->waves = b->waves.data();
    return 0;
}

int mpegb200_video_batch_next(mpegb200_video_batch* b, mpegb200_batch_step* out) {
    if (!b || !out) return MPEGB200_EINVAL;
    try {
        return video_batch_next_impl(b, out);
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

// Scan mode: every stream's parser stops at the slice start codes; wave w holds the w-th picture of every stream's step as
// slice tables over one buffer of compressed bytes (the argument list of mpegb200_video_decode_bitstream).
static int video_batch_next_scan_impl(mpegb200_video_batch* b, mpegb200_batch_scan_step* out) {
    memset(out, 0, sizeof(*out));
    const int n = b->n;
    b->scan_steps.resize((size_t)n);
    b->pool->run(n, [&](int i) {
        mpegb200_video_scan_step& st = b->scan_steps[(size_t)i];
        memset(&st, 0, sizeof(st));
        if (b->parsers[(size_t)i] && mpegb200_video_parser_next_scan(b->parsers[(size_t)i], &st) != 0) throw std::bad_alloc();
        b->has_frame[(size_t)i] = st.has_frame;
        b->frame_buf[(size_t)i] = st.frame_buf;
        b->time[(size_t)i] = st.time;
    });
    if (b->pool->failed()) return MPEGB200_ENOMEM;
    int n_waves = 0;
    b->host_index.clear();
    b->host_steps.clear();
    for (int i = 0; i < n; i++) {
        const mpegb200_video_scan_step& st = b->scan_steps[(size_t)i];
        n_waves = std::max(n_waves, st.has_frame ? st.n_pictures : 0);
        if (st.has_frame && st.host_step) {
            b->host_index.push_back(i);
            b->host_steps.push_back(*st.host_step);
        }
    }
    b->flip ^= 1;
    auto& bufs = b->scan_bufs[b->flip];
    if ((int)bufs.size() < n_waves) bufs.resize((size_t)n_waves);
    auto& scan_waves = b->scan_waves[b->flip];
    scan_waves.assign((size_t)n_waves, mpegb200_vlc_wave{});
    std::vector<uint32_t> pic_of((size_t)n), slice_off((size_t)n), slot_off((size_t)n);
    std::vector<uint64_t> byte_off((size_t)n), first_byte((size_t)n), n_bytes((size_t)n);
    for (int w = 0; w < n_waves; w++) {
        uint32_t np = 0, ns = 0;
        uint64_t nbytes = 0, nslots = 0;
        for (int i = 0; i < n; i++) {
            const mpegb200_video_scan_step& st = b->scan_steps[(size_t)i];
            pic_of[(size_t)i] = 0xffffffffu;
            if (!st.has_frame || w >= st.n_pictures) continue;
            const mpegb200_scan_picture& P = st.pictures[w];
            pic_of[(size_t)i] = np++;
            slice_off[(size_t)i] = ns;
            slot_off[(size_t)i] = (uint32_t)nslots;
            byte_off[(size_t)i] = nbytes;
            ns += P.n_slices;
            // record slots: a slice may write the addresses from its own row to the next slice's (or to the picture's end)
            const mpegb200_scan_slice* S = st.slices + P.first_slice;
            const int64_t mb_size = (int64_t)st.mb_w * st.mb_h;
            for (uint32_t k = 0; k < P.n_slices; k++) {
                const int64_t from = ((int64_t)S[k].vpos - 1) * st.mb_w;
                const int64_t to = k + 1 < P.n_slices ? ((int64_t)S[k + 1].vpos - 1) * st.mb_w : mb_size;
                const int64_t cap = std::max<int64_t>(0, std::min(to, mb_size) - from);
                nslots += (uint64_t)((cap + 15) & ~(int64_t)15);
            }
            // bytes: from the first slice to eight bytes behind the start code that ends the last one (what a reader that
            // stops in front of that code may still look at)
            if (P.n_slices && !b->resident) {
                first_byte[(size_t)i] = S[0].offset;
                const uint64_t last = std::min<uint64_t>(S[P.n_slices - 1].next_code + 8, st.stream_len);
                n_bytes[(size_t)i] = last - S[0].offset;
            } else {
                first_byte[(size_t)i] = 0;
                n_bytes[(size_t)i] = 0;
            }
            nbytes += (n_bytes[(size_t)i] + 15) & ~(uint64_t)15;
        }
        if (nslots > 0xffffffffull / 6 || nbytes >= 0xffffff00ull) return MPEGB200_EINVAL;
        auto& buf = bufs[(size_t)w];
        if (!b->reserve(buf.pics, buf.cap_pics, np) || !b->reserve(buf.slices, buf.cap_slices, ns) ||
            !b->reserve(buf.bits, buf.cap_bits, (size_t)nbytes + 16) || !b->reserve(buf.quant, buf.cap_quant, (size_t)np * 128) ||
            !b->reserve(buf.step_picture, buf.cap_step, np))
            return MPEGB200_ENOMEM;
        b->pool->run(n, [&](int i) {
            const uint32_t p = pic_of[(size_t)i];
            if (p == 0xffffffffu) return;
            const mpegb200_video_scan_step& st = b->scan_steps[(size_t)i];
            const mpegb200_scan_picture& P = st.pictures[w];
            const mpegb200_scan_slice* S = st.slices + P.first_slice;
            mpegb200_vlc_picture V;
            memset(&V, 0, sizeof(V));
            V.stream = i;
            V.type = P.type;
            V.dst_buf = P.dst_buf;
            V.fwd_buf = P.fwd_buf;
            V.bwd_buf = P.bwd_buf;
            V.fwd_full_px = P.fwd_full_px;
            V.fwd_r_size = P.fwd_r_size;
            V.bwd_full_px = P.bwd_full_px;
            V.bwd_r_size = P.bwd_r_size;
            V.first_slice = slice_off[(size_t)i];
            V.n_slices = P.n_slices;
            V.mb_slot = slot_off[(size_t)i];
            V.quant = p;
            memcpy(buf.quant + (size_t)p * 128, st.quant, 128);
            buf.step_picture[p] = w;
            const int64_t mb_size = (int64_t)st.mb_w * st.mb_h;
            uint32_t slot = slot_off[(size_t)i];
            for (uint32_t k = 0; k < P.n_slices; k++) {
                const int64_t from = ((int64_t)S[k].vpos - 1) * st.mb_w;
                const int64_t to = k + 1 < P.n_slices ? ((int64_t)S[k + 1].vpos - 1) * st.mb_w : mb_size;
                const int64_t cap = std::max<int64_t>(0, std::min(to, mb_size) - from);
                mpegb200_vlc_slice L;
                memset(&L, 0, sizeof(L));
                L.data_offset = b->resident ? S[k].offset : byte_off[(size_t)i] + (S[k].offset - first_byte[(size_t)i]);
                L.next_code = (uint32_t)std::min<uint64_t>(S[k].next_code - S[k].offset, 0xffffffffull);
                L.stream_left = (uint32_t)std::min<uint64_t>(st.stream_len - S[k].offset, 0xffffffffull);
                L.pic = p;
                L.vpos = S[k].vpos;
                L.mb_slot = slot;
                L.mb_cap = (uint32_t)((cap + 15) & ~(int64_t)15);
                slot += L.mb_cap;
                buf.slices[slice_off[(size_t)i] + k] = L;
            }
            V.n_mb_slots = slot - slot_off[(size_t)i];
            buf.pics[p] = V;
            if (!b->resident) {
                uint8_t* dst = buf.bits + byte_off[(size_t)i];
                memcpy(dst, st.stream + first_byte[(size_t)i], (size_t)n_bytes[(size_t)i]);
                memset(dst + n_bytes[(size_t)i], 0, (size_t)(((n_bytes[(size_t)i] + 15) & ~(uint64_t)15) - n_bytes[(size_t)i]));
            }
        });
        if (b->pool->failed()) return MPEGB200_ENOMEM;
        memset(buf.bits + nbytes, 0, 16);
        mpegb200_vlc_wave& W = scan_waves[(size_t)w];
        W.n_pictures = (int)np;
        W.pics = buf.pics;
        W.step_picture = buf.step_picture;
        W.n_slices = ns;
        W.slices = buf.slices;
        W.bitstream = b->resident ? nullptr : buf.bits;
        W.bitstream_bytes = (size_t)nbytes;
        W.quant = buf.quant;
        W.n_quant = np;
        W.n_mb_slots = (size_t)nslots;
    }
    out->n_streams = n;
    out->has_frame = b->has_frame.data();
    out->frame_buf = b->frame_buf.data();
    out->time = b->time.data();
    out->n_waves = n_waves;
    out->waves = scan_waves.data();
    out->n_host = (int)b->host_index.size();
    out->host_index = b->host_index.data();
    out->host_steps = b->host_steps.data();
    return 0;
}

int mpegb200_video_batch_next_scan(mpegb200_video_batch* b, mpegb200_batch_scan_step* out) {
    if (!b || !out) return MPEGB200_EINVAL;
    try {
        return video_batch_next_scan_impl(b, out);
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

// Not part of the C-ABI (no declaration in include/): the device tables as bytes, for the test that runs the device-side
// slice walker on the CPU (tests/vlc_emu).  Returns their size, or 0 if `cap` is too small or the tables have an unexpected shape.
size_t mpegb200_internal_vlc_tables(void* out, size_t cap) {
    if (!out || cap < sizeof(mpegb200::VlcDeviceTables)) return 0;
    return mpegb200::fill_vlc_device_tables(static_cast<mpegb200::VlcDeviceTables*>(out)) ? sizeof(mpegb200::VlcDeviceTables) : 0;
}

mpegb200_video_parser* mpegb200_video_batch_parser(mpegb200_video_batch* b, int index) {
    return b && index >= 0 && index < b->n ? b->parsers[(size_t)index] : nullptr;
}

int mpegb200_video_batch_set_resident(mpegb200_video_batch* b, int on) {
    if (!b) return MPEGB200_EINVAL;
    b->resident = on != 0;
    return 0;
}

int mpegb200_video_batch_set_start_codes(mpegb200_video_batch* b, int index, const uint64_t* positions, size_t n) {
    if (!b || index < 0 || index >= b->n || !b->parsers[(size_t)index]) return MPEGB200_EINVAL;
    return mpegb200_video_parser_set_start_codes(b->parsers[(size_t)index], positions, n);
}

int mpegb200_video_batch_unscan(mpegb200_video_batch* b) {
    if (!b) return MPEGB200_EINVAL;
    int rc = 0;
    for (auto* p : b->parsers)
        if (p && mpegb200_video_parser_unscan(p) != 0) rc = MPEGB200_ESTATE;
    return rc;
}

int mpegb200_video_batch_redo(mpegb200_video_batch* b, int index, int step_picture, mpegb200_video_step* out) {
    if (!b || !out || index < 0 || index >= b->n || !b->parsers[(size_t)index]) return MPEGB200_EINVAL;
    const int rc = mpegb200_video_parser_redo(b->parsers[(size_t)index], step_picture, out);
    if (rc == 0) {   // the step's result for this stream is the re-parsed tail's
        b->has_frame[(size_t)index] = out->has_frame;
        b->frame_buf[(size_t)index] = out->frame_buf;
        b->time[(size_t)index] = out->time;
    }
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Many MP2 streams in lock-step: parse the next frames of every stream on the pool, hand them over as one rectangular
// batch (one mpegb200_audio_synth launch for all streams) plus the tails of the streams that end inside the step.
// ------------------------------------------------------------------------------------------------
struct mpegb200_audio_batch {
    std::unique_ptr<WorkerPool> pool;
    int n = 0;
    void* (*alloc)(size_t) = nullptr;
    void (*free_fn)(void*) = nullptr;
    std::vector<mpegb200_audio_parser*> parsers;
    std::vector<int> n_frames;
    std::vector<double> time;
    std::vector<int32_t> full_index, tail_index, tail_frames;
    int32_t* stage = nullptr;        // [n][frames][2][36][32], every stream parses into its own rows
    int32_t* full = nullptr;         // compacted rectangular part (pinned when alloc is given)
    int32_t* tail = nullptr;
    size_t cap_stage = 0, cap_full = 0, cap_tail = 0;

    void* get(size_t bytes) { return alloc ? alloc(bytes ? bytes : 1) : malloc(bytes ? bytes : 1); }
    void put(void* p) {
        if (!p) return;
        if (free_fn) free_fn(p); else free(p);
    }
    bool reserve(int32_t*& p, size_t& cap, size_t need) {
        if (need <= cap) return true;
        put(p);
        cap = need + need / 4 + 64;
        p = (int32_t*)get(cap * sizeof(int32_t));
        return p != nullptr;
    }
};

extern "C" {

mpegb200_audio_batch* mpegb200_audio_batch_new(int n_streams, int threads, void* (*alloc)(size_t), void (*free_fn)(void*)) {
    if (n_streams <= 0 || n_streams > (1 << 20)) return nullptr;
    auto* b = new (std::nothrow) mpegb200_audio_batch();
    if (!b) return nullptr;
    try {
        b->n = n_streams;
        b->pool.reset(new WorkerPool(threads > 0 ? threads : 1));
        b->alloc = alloc;
        b->free_fn = free_fn;
        b->parsers.assign((size_t)n_streams, nullptr);
        b->n_frames.assign((size_t)n_streams, 0);
        b->time.assign((size_t)n_streams, 0.0);
    } catch (...) {
        delete b;
        return nullptr;
    }
    return b;
}

void mpegb200_audio_batch_free(mpegb200_audio_batch* b) {
    if (!b) return;
    for (auto* p : b->parsers) mpegb200_audio_parser_free(p);
    b->put(b->stage);
    b->put(b->full);
    b->put(b->tail);
    delete b;
}

int mpegb200_audio_batch_set_stream(mpegb200_audio_batch* b, int index, const uint8_t* data, size_t len) {
    if (!b || index < 0 || index >= b->n) return MPEGB200_EINVAL;
    mpegb200_audio_parser_free(b->parsers[(size_t)index]);
    b->parsers[(size_t)index] = mpegb200_audio_parser_new(data, len);
    return b->parsers[(size_t)index] ? 0 : MPEGB200_ENOMEM;
}

int mpegb200_audio_batch_stream_info(mpegb200_audio_batch* b, int index, int* samplerate, int* channels) {
    if (!b || index < 0 || index >= b->n || !b->parsers[(size_t)index]) return MPEGB200_EINVAL;
    if (samplerate) *samplerate = mpegb200_audio_parser_samplerate(b->parsers[(size_t)index]);
    if (channels) *channels = mpegb200_audio_parser_channels(b->parsers[(size_t)index]);
    return 0;
}

static int audio_batch_next_impl(mpegb200_audio_batch* b, int F, mpegb200_audio_batch_step* out) {
    const size_t frame = 2 * 36 * 32;
    const int n = b->n;
    if (!b->reserve(b->stage, b->cap_stage, (size_t)n * F * frame)) return MPEGB200_ENOMEM;
    b->pool->run(n, [&](int i) {
        int k = 0;
        double t0 = 0, t = 0;
        mpegb200_audio_parser* p = b->parsers[(size_t)i];
        while (p && k < F && mpegb200_audio_parser_next(p, b->stage + ((size_t)i * F + k) * frame, &t)) {
            if (k == 0) t0 = t;
            k++;
        }
        b->n_frames[(size_t)i] = k;
        b->time[(size_t)i] = t0;
    });
    b->full_index.clear();
    b->tail_index.clear();
    b->tail_frames.clear();
    size_t tail_total = 0;
    for (int i = 0; i < n; i++) {
        const int k = b->n_frames[(size_t)i];
        if (k == F) {
            b->full_index.push_back(i);
        } else if (k > 0) {
            b->tail_index.push_back(i);
            b->tail_frames.push_back(k);
            tail_total += (size_t)k;
        }
    }
    const size_t n_full = b->full_index.size();
    if (!b->reserve(b->full, b->cap_full, n_full * F * frame) || !b->reserve(b->tail, b->cap_tail, tail_total * frame))
        return MPEGB200_ENOMEM;
    b->pool->run((int)n_full, [&](int j) {
        memcpy(b->full + (size_t)j * F * frame, b->stage + (size_t)b->full_index[(size_t)j] * F * frame, (size_t)F * frame * sizeof(int32_t));
    });
    size_t at = 0;
    for (size_t j = 0; j < b->tail_index.size(); j++) {
        memcpy(b->tail + at * frame, b->stage + (size_t)b->tail_index[j] * F * frame, (size_t)b->tail_frames[j] * frame * sizeof(int32_t));
        at += (size_t)b->tail_frames[j];
    }
    out->n_streams = n;
    out->n_frames = b->n_frames.data();
    out->time = b->time.data();
    out->frames_per_stream = F;
    out->n_full = (int)n_full;
    out->full_index = b->full_index.data();
    out->full_samples = b->full;
    out->n_tail = (int)b->tail_index.size();
    out->tail_index = b->tail_index.data();
    out->tail_frames = b->tail_frames.data();
    out->tail_samples = b->tail;
    return 0;
}

int mpegb200_audio_batch_next(mpegb200_audio_batch* b, int frames_per_stream, mpegb200_audio_batch_step* out) {
    if (!b || !out || frames_per_stream <= 0 || frames_per_stream > 4096) return MPEGB200_EINVAL;
    memset(out, 0, sizeof(*out));
    try {
        return audio_batch_next_impl(b, frames_per_stream, out);
    } catch (...) {
        memset(out, 0, sizeof(*out));
        return MPEGB200_ENOMEM;
    }
}

}  // extern "C"

