// vlc_slice_walk.h -- one MPEG-1 slice walked by one thread: the core of the device-side VLC stage (vlc_slices.cu).
//
// decodeSlice / decodeMacroblock / decodeMotionVectors / decodeBlock of the reference (video.go:436-746) with the same
// table-driven reading as the host parser (host_parser.cpp: one bit window per coefficient, a 12-bit direct table for the
// codes behind the first coefficient, a two-level table for the rest), writing packed records -- mpegb200_mb + int16[64] per
// coded block -- into the record slots the slice owns.  What the serial reference resolves by order of arrival is detected
// and flagged here, never emulated (include/mpegb200.h, "slice-parallel VLC stage").
//
// The functions are __host__ __device__ so that tests/vlc_emu.cpp can run this very code on the CPU against the host parser
// (no GPU in the development container); the product calls it from vlc_parse_kernel only.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/mpegb200.h"
#include "vlc_device_tables.h"

#if defined(__CUDACC__)
#define VLC_HD __host__ __device__ __forceinline__
#else
#define VLC_HD inline
#endif

static_assert(sizeof(mpegb200_vlc_picture) == 32 && sizeof(mpegb200_vlc_slice) == 32, "slice tables: 32-byte entries");

namespace mpegb200 {

struct VlcGeometry {         // what the walker needs of a stream (StreamInfo, common.cuh)
    int mb_w, mb_h, luma_w, luma_h;
    uint32_t buf_bytes;
};

struct SliceSummary {        // 16 bytes per slice, read by vlc_check_picture
    uint32_t flags;
    int32_t first_addr;      // address of the first record, -1 if none
    int32_t last_addr;       // address of the last record
    int32_t end_addr;        // the reference's mbAddress when the slice ended
};

// bitstream words and quantiser bytes come from global memory through the read-only path; the code tables are read with plain
// loads (they sit in shared memory on the device: the walker is inlined into the kernel and the compiler sees the address space)
VLC_HD uint32_t vlc_ldg32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
VLC_HD uint32_t vlc_ld32(const uint32_t* p) { return *p; }
VLC_HD uint32_t vlc_ld8(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
VLC_HD int vlc_clz(uint32_t v) {   // v != 0
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}
VLC_HD uint32_t vlc_bswap(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}
VLC_HD void vlc_store16(mpegb200_mb* dst, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(dst) = make_uint4(x, y, z, w);
#else
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    d[0] = x;
    d[1] = y;
    d[2] = z;
    d[3] = w;
#endif
}
// one 16-byte chunk of the block scratch leaves for the block slot (if there is room) and is zero again
VLC_HD void vlc_flush16(uint8_t* scratch, int16_t* out, bool room) {
#if defined(__CUDA_ARCH__)
    uint4* const s = reinterpret_cast<uint4*>(scratch);
    if (room) *reinterpret_cast<uint4*>(out) = *s;
    *s = make_uint4(0, 0, 0, 0);
#else
    uint64_t* const s = reinterpret_cast<uint64_t*>(scratch);
    if (room) {
        reinterpret_cast<uint64_t*>(out)[0] = s[0];
        reinterpret_cast<uint64_t*>(out)[1] = s[1];
    }
    s[0] = s[1] = 0;
#endif
}

struct Bits {                // most significant bit first (buffer.go:223-255)
    const uint32_t* w;
    uint32_t n_words, next, end_byte;   // end_byte: where the elementary stream ends in the buffer (zero bits beyond)
    uint32_t origin;                    // word index and byte phase of the slice's first bit: pos() counts from there
    uint32_t ahead;                     // word `next`, already loaded: its latency hides behind the bits in front of it
    uint64_t buf;
    int cnt;                            // valid bits at the top of buf; >= 32 between two operations

    // Word i of the buffer, most significant bit first, zero behind the end of the stream and of the buffer -- without a branch:
    // the index is clamped, the bytes that count are kept by a mask (1.154 against 1.195 ms per wave for the branching form; a
    // fast path for words that need no masking with the rest in a function of its own: 1.37 ms).
    uint32_t end_eff, last_word;        // min(end_byte, 4 * n_words); n_words - 1
    VLC_HD uint32_t load(uint32_t i) const {
        const uint32_t v = vlc_bswap(vlc_ldg32(w + (i < last_word ? i : last_word)));
        long long vb = (long long)end_eff - 4ll * (long long)i;          // bytes of word i in front of the end
        vb = vb < 0 ? 0 : (vb > 4 ? 4 : vb);
        return v & (uint32_t)(0xffffffff00000000ull >> (8 * (int)vb));
    }
    VLC_HD void refill() {
        if (cnt < 32) {
            buf |= (uint64_t)ahead << (32 - cnt);
            cnt += 32;
            ahead = load(++next);
        }
    }
    VLC_HD void init(uint32_t byte_offset) {
        end_eff = end_byte < n_words * 4u ? end_byte : n_words * 4u;   // n_words < 2^30 (the entry point bounds the buffer)
        last_word = n_words - 1u;
        next = byte_offset >> 2;
        const int mis = (int)(byte_offset & 3u);
        origin = next * 32u + 8u * (uint32_t)mis;   // modulo 2^32, like pos(): a slice is shorter than 2^32 bits
        buf = (uint64_t)load(next) << (32 + 8 * mis);
        cnt = 32 - 8 * mis;
        ahead = load(++next);
        refill();
    }
    VLC_HD uint32_t pos() const { return next * 32u - (uint32_t)cnt - origin; }   // bits consumed since init
    VLC_HD uint32_t peek(int n) const { return n ? (uint32_t)(buf >> (64 - n)) : 0u; }   // n <= 32
    VLC_HD void skip(int n) {                                                            // n <= 32
        buf <<= n;
        cnt -= n;
        refill();
    }
    VLC_HD uint32_t read(int n) {
        const uint32_t v = peek(n);
        skip(n);
        return v;
    }
};

VLC_HD int vlc_entry_value(uint32_t e) { return (int)(int16_t)(e & 0xffffu); }
VLC_HD int vlc_entry_len(uint32_t e) { return (int)((e >> 16) & 0xffu); }

// one code from a single-level table
VLC_HD int read_vlc(Bits& br, const uint32_t* table, int bits) {
    const uint32_t e = vlc_ld32(table + br.peek(bits));
    br.skip(vlc_entry_len(e));
    return vlc_entry_value(e);
}

struct VlcMotion {
    int h, v;
    bool is_set;
};

// decodeMotionVector, video.go:583-606
VLC_HD int read_motion(Bits& br, const VlcDeviceTables* T, int r_size, int motion) {
    const int fscale = 1 << r_size;
    const int m_code = read_vlc(br, T->motion, kVlcMotionBits);
    int d;
    if (m_code != 0 && fscale != 1) {
        const int r = (int)br.read(r_size);
        d = (((m_code < 0 ? -m_code : m_code) - 1) << r_size) + r + 1;
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

// 32-bit arithmetic suffices: a frame buffer of 4096 x 4096 luma is 2^24.6 bytes, a vector at most 2^11 pixels
VLC_HD bool vlc_window_inside(int off, int stride, int size, int odd_h, int odd_v, int avail) {
    const int hi = off + (size - 1 + odd_v) * stride + (size - 1 + odd_h);
    return off >= 0 && hi < avail;
}

// Walk slice `sl` of picture P.  fast / zigzag: the hot tables (shared memory on the device); blk_base: this thread's 128 bytes of
// zeroed block scratch, int16 position p in 16-byte chunk (p >> 3) ^ swz; mbs / coeffs: the wave's record and block arrays.
// Fills every record slot of the slice (null records behind the last macroblock).
VLC_HD SliceSummary walk_slice(const VlcDeviceTables* T, const uint32_t* fast, const uint8_t* zigzag, uint8_t* blk_base, uint32_t swz,
                               const mpegb200_vlc_slice& sl, const mpegb200_vlc_picture& P, const VlcGeometry& g, const uint32_t* words,
                               uint32_t n_words, const uint8_t* quant, mpegb200_mb* mbs, int16_t* coeffs) {
    SliceSummary sum;
    sum.flags = 0;
    sum.first_addr = -1;
    sum.last_addr = -1;
    sum.end_addr = 0;
    uint32_t slot = sl.mb_slot, block = 6u * sl.mb_slot;
    const uint32_t slot_end = sl.mb_slot + sl.mb_cap;
    const uint32_t swz16 = swz << 4;
    auto put = [&](int p, int value) {   // byte offset 2p with the 16-byte chunk index (bits 4..6) swizzled
        *reinterpret_cast<int16_t*>(blk_base + (((uint32_t)p << 1) ^ swz16)) = (int16_t)value;
    };
    {
        const int mb_w = g.mb_w, mb_h = g.mb_h, mb_size = mb_w * mb_h;
        (void)mb_h;
        const uint8_t* const q_intra = quant + (size_t)P.quant * 128, * const q_inter = q_intra + 64;
        const int type = P.type;
        const uint32_t* const type_table = type == MPEGB200_PIC_I ? T->type_i : (type == MPEGB200_PIC_P ? T->type_p : T->type_b);
        const int type_bits = type == MPEGB200_PIC_I ? kVlcTypeIBits : (type == MPEGB200_PIC_P ? kVlcTypePBits : kVlcTypeBBits);
        const int total = (int)g.buf_bytes, luma_bytes = g.luma_w * g.luma_h;
        const int chroma_w = g.luma_w >> 1;

        Bits br;
        br.w = words;
        br.n_words = n_words;
        const uint32_t start = (uint32_t)sl.data_offset;
        const uint64_t end64 = sl.data_offset + sl.stream_left;
        br.end_byte = end64 > 0xffffffffull ? 0xffffffffu : (uint32_t)end64;
        br.init(start);
        const long long stream_bits = (long long)sl.stream_left * 8;
        auto has = [&](int n) { return stream_bits - (long long)br.pos() >= n; };   // buffer.go:203-221
        const uint32_t limit_bits = sl.next_code > 0x0fffffffu ? 0x7fffffffu : sl.next_code * 8u;

        // decodeSlice, video.go:436-460
        bool slice_begin = true;
        int mb_addr = ((int)sl.vpos - 1) * mb_w - 1;
        VlcMotion fwd{0, 0, false}, bwd{0, 0, false};
        int dc_pred0 = 128, dc_pred1 = 128, dc_pred2 = 128;
        int qscale = (int)br.read(5);
        while (br.read(1)) {
            if (has(8)) br.skip(8);
        }
        uint32_t flags = 0;

        // one record; vectors resolved like predictMacroblock (video.go:608-637): in a B picture the backward copy
        // overwrites the forward one, so a record carries exactly one reference
        auto emit = [&](int addr, bool intra, uint32_t cbp, uint32_t first_block) {
            const int row = addr / mb_w, col = addr - row * mb_w;
            uint32_t f = MPEGB200_MB_INTRA;
            int h = 0, v = 0;
            if (!intra) {
                h = fwd.h;
                v = fwd.v;
                if (P.fwd_full_px) {
                    h *= 2;
                    v *= 2;
                }
                bool use_bwd = false;
                if (type == MPEGB200_PIC_B && (!fwd.is_set || bwd.is_set)) {
                    use_bwd = true;
                    h = bwd.h;
                    v = bwd.v;
                    if (P.bwd_full_px) {
                        h *= 2;
                        v *= 2;
                    }
                }
                f = MPEGB200_MB_PREDICT | (use_bwd ? MPEGB200_MB_REF_BWD : 0u);
                h = (int)(int16_t)h;   // the record holds 16 bits, like the host parser's
                v = (int)(int16_t)v;
                // the windows copyMacroblock reads must lie inside the frame buffer (mpegb200_video_validate)
                const int hp = h >> 1, vp = v >> 1;
                const int lsi = ((row << 4) + vp) * g.luma_w + (col << 4) + hp;
                const int cmh = h / 2, cmv = v / 2;
                const int csi = ((row << 3) + (cmv >> 1)) * chroma_w + (col << 3) + (cmh >> 1);
                const int cb0 = luma_bytes, cr0 = cb0 + luma_bytes / 4;
                if ((use_bwd ? P.bwd_buf : P.fwd_buf) == P.dst_buf || !vlc_window_inside(lsi, g.luma_w, 16, h & 1, v & 1, total) ||
                    !vlc_window_inside(csi, chroma_w, 8, cmh & 1, cmv & 1, total - cb0) ||
                    !vlc_window_inside(csi, chroma_w, 8, cmh & 1, cmv & 1, total - cr0))
                    flags |= MPEGB200_VLC_WINDOW;
            }
            if (slot >= slot_end) {
                flags |= MPEGB200_VLC_OVERFLOW;
                return;
            }
            struct { uint32_t x, y, z, w; } r;
            r.x = (uint32_t)row | ((uint32_t)col << 16);
            r.y = ((uint32_t)h & 0xffffu) | ((uint32_t)v << 16);
            r.z = f | (cbp << 8) | (sl.pic << 16);
            r.w = first_block;
            vlc_store16(mbs + slot++, r.x, r.y, r.z, r.w);
            if (sum.first_addr < 0) sum.first_addr = addr;
            sum.last_addr = addr;
        };

        do {
            // ---- decodeMacroblock, video.go:462-562 ----
            int inc = 0, code = read_vlc(br, T->addr_inc, kVlcAddrIncBits);
            while (code == 34) code = read_vlc(br, T->addr_inc, kVlcAddrIncBits);   // stuffing
            while (code == 35) {                                                    // escape
                inc += 33;
                code = read_vlc(br, T->addr_inc, kVlcAddrIncBits);
                if (inc > 65536) break;                                             // damaged stream: flagged below (address leaves the picture)
            }
            inc += code;
            bool parse_mb = true;
            if (slice_begin) {
                slice_begin = false;
                mb_addr += inc;
            } else {
                if (mb_addr + inc >= mb_size) {
                    flags |= MPEGB200_VLC_DROPPED;
                    parse_mb = false;
                } else {
                    if (inc > 1) {
                        dc_pred0 = dc_pred1 = dc_pred2 = 128;
                        if (type == MPEGB200_PIC_P) fwd.h = fwd.v = 0;
                    }
                    while (inc > 1) {   // skipped macroblocks are pure predictions
                        mb_addr++;
                        emit(mb_addr, false, 0, block);
                        inc--;
                    }
                    mb_addr++;
                }
            }
            if (parse_mb && (mb_addr < 0 || mb_addr >= mb_size)) {
                flags |= MPEGB200_VLC_DROPPED;
                parse_mb = false;
            }
            if (parse_mb) {
                const int mtype = read_vlc(br, type_table, type_bits);
                const bool intra = mtype & 0x01;
                fwd.is_set = mtype & 0x08;
                bwd.is_set = mtype & 0x04;
                if (mtype & 0x10) qscale = (int)br.read(5);
                if (intra) {
                    fwd.h = fwd.v = bwd.h = bwd.v = 0;
                } else {
                    dc_pred0 = dc_pred1 = dc_pred2 = 128;
                    if (fwd.is_set) {   // decodeMotionVectors, video.go:564-581
                        fwd.h = read_motion(br, T, P.fwd_r_size, fwd.h);
                        fwd.v = read_motion(br, T, P.fwd_r_size, fwd.v);
                    } else if (type == MPEGB200_PIC_P) {
                        fwd.h = fwd.v = 0;
                    }
                    if (bwd.is_set) {
                        bwd.h = read_motion(br, T, P.bwd_r_size, bwd.h);
                        bwd.v = read_motion(br, T, P.bwd_r_size, bwd.v);
                    }
                }
                uint32_t cbp = 0;
                if (mtype & 0x02)
                    cbp = (uint32_t)read_vlc(br, T->cbp, kVlcCbpBits);
                else if (intra)
                    cbp = 0x3f;
                const uint32_t first_block = block;
                const bool room = slot < slot_end;
                const uint8_t* const q = intra ? q_intra : q_inter;
                // the k-th CODED block of every slice of the warp in the same iteration (not block k: the slices' patterns differ).
                // (One loop over all coefficients of the macroblock, a slice entering its next block inside it, raises the active
                // lanes from 1.97 to 2.28 but measured slower, 1.30 against 1.19 ms: profiles/r2_vlc_summary.md.)
                for (uint32_t left = cbp & 0x3fu; left && !(flags & MPEGB200_VLC_INVALID_RUN);) {
                    const int b = vlc_clz(left) - 26;
                    left &= ~(0x20u >> b);
                    // ---- decodeBlock, video.go:639-746 ----
                    int n = 0;
                    if (intra) {
                        const int size = b < 4 ? read_vlc(br, T->dc_luma, kVlcDcLumaBits) : read_vlc(br, T->dc_chroma, kVlcDcChromaBits);
                        int dc = b < 4 ? dc_pred0 : (b == 4 ? dc_pred1 : dc_pred2);
                        if (size > 0) {
                            const int diff = (int)br.read(size);
                            dc += (diff & (1 << (size - 1))) ? diff : (-(1 << size) | (diff + 1));
                        }
                        if (b < 4) dc_pred0 = dc; else if (b == 4) dc_pred1 = dc; else dc_pred2 = dc;
                        const int l = dc * 8;   // dc << 8 == (dc * 8) * premultiplier[0], video.go:672
                        put(0, l > 32767 ? 32767 : (l < -32768 ? -32768 : l));
                        n = 1;
                    }
                    for (;;) {
                        uint64_t w = br.buf;
                        int used, run, lv;
                        const uint32_t fe = fast[(uint32_t)(w >> (64 - kVlcCoefFastBits))];
                        if ((fe >> 24) != 0 && n > 0) {   // a short code behind the first coefficient: code, sign and end of block in one look-up
                            used = (int)(fe >> 24);
                            run = (int)((fe >> 16) & 0xffu);
                            if (run == 0xff) {
                                br.skip(used);
                                break;
                            }
                            lv = (int)(int16_t)(fe & 0xffffu);
                        } else {
                            const uint32_t bits16 = (uint32_t)(w >> 48);
                            uint32_t c;
                            if ((bits16 >> 10) == 1u) {   // "000001": escape, no look-up needed
                                c = 0xffffu;
                                used = 6;
                            } else {
                                uint32_t e = vlc_ld32(&T->coeff_first[bits16 >> kVlcCoefSecondBits]);
                                if (e & kVlcLink) e = vlc_ld32(&T->coeff_second[e & 0xffu][bits16 & ((1u << kVlcCoefSecondBits) - 1u)]);
                                c = e & 0xffffu;
                                used = vlc_entry_len(e);
                            }
                            w <<= used;
                            if (c == 0x0001u && n > 0) {   // "1" behind the first coefficient: a 0 bit ends the block (video.go:686) ...
                                if ((w >> 63) == 0) {
                                    br.skip(used + 1);
                                    break;
                                }
                                w <<= 1;                   // ... a 1 bit is consumed and the sign follows
                                used += 1;
                            }
                            if (c == 0xffffu) {            // escape: 6 bits of run, 8 (or 16) bits of level
                                run = (int)(w >> 58);
                                lv = (int)((w >> 50) & 0xffu);
                                used += 14;
                                if (lv == 0) {
                                    lv = (int)((w >> 42) & 0xffu);
                                    used += 8;
                                } else if (lv == 128) {
                                    lv = (int)((w >> 42) & 0xffu) - 256;
                                    used += 8;
                                } else if (lv > 128) {
                                    lv -= 256;
                                }
                            } else {
                                run = (int)(c >> 8);
                                lv = (int)(c & 0xffu);
                                if (w >> 63) lv = -lv;
                                used += 1;
                            }
                        }
                        br.skip(used);
                        n += run;
                        if (n >= 64) {   // invalid run: the reference drops the block and keeps its coefficients for the next one -- serial state, host's job
                            flags |= MPEGB200_VLC_INVALID_RUN;
                            break;
                        }
                        const int dz = zigzag[n++];
                        lv *= 2;         // dequantise, oddify, clip: video.go:719-741
                        if (!intra) lv += (lv >> 31) | 1;
                        lv = (lv * (qscale * (int)vlc_ld8(q + dz))) >> 4;
                        if ((lv & 1) == 0) lv -= lv > 0 ? 1 : -1;
                        lv = lv > 2047 ? 2047 : (lv < -2048 ? -2048 : lv);
                        put(dz, lv);
                    }
                    // hand-over: eight 16-byte chunks leave for the block slot and the scratch is zero again
                    int16_t* const out = coeffs + (size_t)block * 64;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                    for (uint32_t c = 0; c < 8; c++) vlc_flush16(blk_base + ((c ^ swz) << 4), out + c * 8, room);
                    if (room) block++;
                }
                emit(mb_addr, intra, cbp, first_block);
            }
            if (br.pos() > limit_bits) flags |= MPEGB200_VLC_OVERRUN;
        } while (!(flags & (MPEGB200_VLC_OVERRUN | MPEGB200_VLC_INVALID_RUN | MPEGB200_VLC_OVERFLOW)) && mb_addr < mb_size - 1 &&
                 has(23) && br.peek(23) != 0);
        // the next start code is searched from the next byte boundary (buffer.go:279-302): the slice must end in front of the scan's
        if (((br.pos() + 7u) >> 3) > sl.next_code) flags |= MPEGB200_VLC_OVERRUN;
        sum.flags = flags;
        sum.end_addr = mb_addr;
    }
    // unused slots: null records whose block index continues the packing (a group's first record names the group's first block)
    for (; slot < slot_end; slot++) vlc_store16(mbs + slot, 0u, 0u, 0xffffu << 16, block);
    return sum;
}

// What the slices of a picture owe each other (one thread per picture): the OR of their flags, plus order and early end.
VLC_HD uint32_t vlc_check_picture(const mpegb200_vlc_picture& P, const SliceSummary* summary, int mb_size) {
    uint32_t flags = 0;
    int prev_last = -1;
    for (uint32_t k = 0; k < P.n_slices; k++) {
        const SliceSummary sm = summary[P.first_slice + k];
        flags |= sm.flags;
        if (sm.first_addr >= 0) {
            if (sm.first_addr <= prev_last) flags |= MPEGB200_VLC_ORDER;
            prev_last = sm.last_addr;
        }
        if (k + 1 < P.n_slices && sm.end_addr >= mb_size - 2) flags |= MPEGB200_VLC_EARLY_END;   // video.go:424-426
    }
    return flags;
}

}  // namespace mpegb200
