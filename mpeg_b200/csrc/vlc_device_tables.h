// vlc_device_tables.h -- the variable-length-code tables of ISO 11172-2 in the flat form the device-side slice parser
// (vlc_slices.cu) reads.  Filled on the host by fill_vlc_device_tables (host_parser.cpp) from the very tables the host parser
// uses, uploaded once per context (ctx.cu).  Internal: not part of the C-ABI.
#pragma once

#include <stdint.h>

namespace mpegb200 {

// Entry of a direct table, indexed by the next `bits` bits of the stream:
//   bits  0..15  value (int16; 0xffff = escape in the coefficient table; for a link: index of the second-level table)
//   bits 16..23  code length in bits (what the reference's tree walk of buffer.go:352-376 consumes, unassigned prefixes included)
//   bit  24      link: continue in coeff_second[value] with the next kVlcCoefSecondBits bits
constexpr uint32_t kVlcLink = 1u << 24;
constexpr int kVlcCoefFastBits = 12;     // host_parser.cpp: kCoefBits
constexpr int kVlcCoefFirstBits = 9, kVlcCoefSecondBits = 7, kVlcCoefSecondTables = 16;
constexpr int kVlcAddrIncBits = 11, kVlcMotionBits = 11, kVlcCbpBits = 9, kVlcTypeIBits = 2, kVlcTypePBits = 6, kVlcTypeBBits = 6,
              kVlcDcLumaBits = 7, kVlcDcChromaBits = 8;

struct VlcDeviceTables {
    // coefficient codes behind the first coefficient of a block, code and sign (or end of block) in one look-up:
    // level (int16) | run << 16 (0xff = end of block) | length << 24 (0 = not in this table: escape, long codes)
    uint32_t coef_fast[1 << kVlcCoefFastBits];
    uint32_t coeff_first[1 << kVlcCoefFirstBits];
    uint32_t coeff_second[kVlcCoefSecondTables][1 << kVlcCoefSecondBits];
    uint32_t addr_inc[1 << kVlcAddrIncBits];
    uint32_t motion[1 << kVlcMotionBits];
    uint32_t cbp[1 << kVlcCbpBits];
    uint32_t type_i[1 << kVlcTypeIBits], type_p[1 << kVlcTypePBits], type_b[1 << kVlcTypeBBits];
    uint32_t dc_luma[1 << kVlcDcLumaBits], dc_chroma[1 << kVlcDcChromaBits];
    uint8_t zigzag[64];
};

// false if the host tables do not have the shape assumed above (they do: checked once at context creation)
bool fill_vlc_device_tables(VlcDeviceTables* out);

}  // namespace mpegb200
