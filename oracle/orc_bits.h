/*
 * orc_bits.h -- in-memory bit reader (TEST INFRASTRUCTURE, see mpeg_oracle.h).
 *
 * Restates buffer.go for a source that is completely in memory.  The reference pulls
 * 128 KiB chunks through a load callback (buffer.go:131-156, 203-221); with the whole
 * stream resident the only observable part of that machinery is "a has() that cannot be
 * satisfied marks the buffer as ended" (buffer.go:142-153, 216-218), which is what
 * bits_has does.  Reads past the end return zero bits where the Go code would panic.
 */
#ifndef ORC_BITS_H
#define ORC_BITS_H

#include <stdint.h>
#include <stddef.h>

typedef struct orc_bits {
    const uint8_t* bytes;
    size_t len;        /* bytes */
    int64_t bit_index; /* buffer.go:21 */
    int has_ended;     /* buffer.go:24 */
} orc_bits;

static inline int64_t bits_left(const orc_bits* b) { return ((int64_t)b->len << 3) - b->bit_index; }

/* buffer.go:203-221 */
static inline int bits_has(orc_bits* b, int64_t count) {
    if (bits_left(b) >= count) return 1;
    b->has_ended = 1;
    return 0;
}

/* buffer.go:246-255 */
static inline int bits_read1(orc_bits* b) {
    size_t byte = (size_t)(b->bit_index >> 3);
    int v = 0;
    if (byte < b->len) v = (b->bytes[byte] >> (7 - (b->bit_index & 7))) & 1;
    b->bit_index += 1;
    return v;
}

/* buffer.go:223-244: most significant bit first, count <= 32 in all call sites */
static inline int64_t bits_read(orc_bits* b, int count) {
    int64_t value = 0;
    while (count != 0) {
        size_t byte = (size_t)(b->bit_index >> 3);
        int current = byte < b->len ? b->bytes[byte] : 0;
        int remaining = 8 - (int)(b->bit_index & 7);
        int take = remaining < count ? remaining : count;
        int shift = remaining - take;
        int mask = 0xff >> (8 - take);
        value = (value << take) | ((current & (mask << shift)) >> shift);
        b->bit_index += take;
        count -= take;
    }
    return value;
}

/* buffer.go:257-259 */
static inline void bits_align(orc_bits* b) { b->bit_index = ((b->bit_index + 7) >> 3) << 3; }

/* buffer.go:261-265 */
static inline void bits_skip(orc_bits* b, int count) {
    if (bits_has(b, count)) b->bit_index += count;
}

/* buffer.go:267-277 */
static inline int bits_skip_bytes(orc_bits* b, uint8_t v) {
    bits_align(b);
    int skipped = 0;
    while (bits_has(b, 8) && b->bytes[b->bit_index >> 3] == v) {
        b->bit_index += 8;
        skipped++;
    }
    return skipped;
}

/* buffer.go:279-302 */
static inline int bits_next_start_code(orc_bits* b) {
    bits_align(b);
    for (;;) {
        while (bits_left(b) >= (5 << 3)) {
            size_t i = (size_t)(b->bit_index >> 3);
            if (b->bytes[i] == 0x00 && b->bytes[i + 1] == 0x00 && b->bytes[i + 2] == 0x01) {
                b->bit_index = (int64_t)(i + 4) << 3;
                return b->bytes[i + 3];
            }
            b->bit_index += 8;
        }
        if (!bits_has(b, 5 << 3)) return -1;
    }
}

/* buffer.go:304-311 */
static inline int bits_find_start_code(orc_bits* b, int code) {
    for (;;) {
        int current = bits_next_start_code(b);
        if (current == code || current == -1) return current;
    }
}

/* buffer.go:313-324 */
static inline int bits_has_start_code(orc_bits* b, int code) {
    int64_t prev = b->bit_index;
    int current = bits_find_start_code(b, code);
    b->bit_index = prev;
    return current;
}

/* buffer.go:326-339 */
static inline int bits_find_frame_sync(orc_bits* b) {
    size_t i;
    for (i = (size_t)(b->bit_index >> 3); i + 1 < b->len; i++) {
        if (b->bytes[i] == 0xFF && (b->bytes[i + 1] & 0xFE) == 0xFC) {
            b->bit_index = ((int64_t)(i + 1) << 3) + 3;
            return 1;
        }
    }
    b->bit_index = (int64_t)(i + 1) << 3;
    return 0;
}

/* buffer.go:341-350 */
static inline int bits_peek_non_zero(orc_bits* b, int count) {
    if (!bits_has(b, count)) return 0;
    int64_t v = bits_read(b, count);
    b->bit_index -= count;
    return v != 0;
}

#endif
