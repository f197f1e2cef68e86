// vlen_encode.h -- internal: one block of the variable-width transfer form (include/mpegb200.h), shared by the
// converter (coeff_pack.cpp, mpegb200_pack_coeffs_vlen) and the host parser, which emits the form directly.
#pragma once
#include <cstddef>
#include <cstdint>

namespace mpegb200 {

struct VlenBlock {
    uint32_t header;   // eight 4-bit group codes
    uint32_t bytes;    // payload bytes of the block
    bool ok;           // always true since groups with 16-bit values have their own code (14)
};

// Encodes one block of 64 levels in natural order and writes its payload bytes at `out`; up to 16 bytes past the block's
// own bytes may be written (the caller provides the slack and overwrites it with the next block).
VlenBlock vlen_encode_block(const int16_t* blk, uint8_t* out);

// Payload bytes of a block, from its header alone.
inline uint32_t vlen_block_bytes(uint32_t header) {
    uint32_t n = 0;
    for (int g = 0; g < 8; g++) {
        const uint32_t code = (header >> (4 * g)) & 15u;
        n += code == 13u ? 12u : code == 14u ? 16u : code;
    }
    return n;
}

}  // namespace mpegb200
