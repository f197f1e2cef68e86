// coeff_vlen.cu -- variable-width transfer form of the coefficient blocks (include/mpegb200.h, "vlen"): the kernel
// that expands it on the device to the int16[64] working form the decode kernel reads.
//
// Why: the end-to-end rate of the host-pointer entry points is bound by PCIe, and the coefficients are nine tenths of
// the upload.  What the reference leaves in blockData before the premultiply (video.go:729-741) is, for every
// coefficient the bitstream carried, an ODD number (the "oddification" of :732-736), and 0 for the ones it did not
// carry; only an intra DC (dc*8, :672) is even.  So a block is sent as eight groups of eight values in zig-zag order
// (the order the bitstream has them in: neighbours have similar magnitude), each group with its own width w:
//     code c = (x + sign(x)) / 2      (0 -> 0, +-1 -> +-1, +-3 -> +-2, ...: a bijection from {0, odd} to the integers)
//     w      = bits of the widest c of the group in two's complement (0 when all eight are zero)
// eight values of w bits are exactly w bytes, so groups never straddle bytes.  A group holding an even non-zero value
// (intra DC) travels raw, 12 bits per value.  A block is a 32-bit header (eight 4-bit group codes: 0..12 = w, 13 = raw 12-bit, 14 = raw 16-bit)
// plus sum(w) payload bytes; 32 blocks form a chunk with one 64-bit payload offset, offsets inside a chunk come from a
// warp scan over the headers.  The dense blocks of the benchmark shrink from 128 (96 in the 12-bit form) to about 49
// bytes, typical sparse blocks of a real stream to 4 + a few bytes.
#include <cstdint>

#include "common.cuh"

namespace mpegb200 {

namespace {

// zig-zag scan, video.go:1044-1053: position p of the scan -> index in the natural 8x8 order
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
#define MPEGB200_ZIGZAG_LIST                                                                                         \
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, \
        42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63
__constant__ uint8_t kZigZagDev[64] = {MPEGB200_ZIGZAG_LIST};
constexpr bool zigzag_list_matches() {
    constexpr uint8_t lit[64] = {MPEGB200_ZIGZAG_LIST};
    for (int i = 0; i < 64; i++)
        if (lit[i] != kZigZag.nat[i]) return false;
    return true;
}
static_assert(zigzag_list_matches(), "the device table is the generated zig-zag scan");

// ------------------------------------------------------------------------------------------------
// Expansion kernel.  One warp = one chunk of 32 blocks: lane l reads the header of block l, a warp scan gives the
// blocks' payload offsets; then eight rounds of four blocks, lane = (block in round, group): the lane shifts its
// group's w bytes through a 128-bit window, undoes the code and puts the eight values at their natural positions in a
// 512-byte staging tile, which the warp writes out as 32 x 16 bytes.
// ------------------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 8;

__global__ void __launch_bounds__(32 * kWarpsPerCta) expand_vlen_kernel(const uint32_t* __restrict__ headers,
                                                                        const uint64_t* __restrict__ chunk_offsets,
                                                                        const uint8_t* __restrict__ payload,
                                                                        uint4* __restrict__ out, size_t n_blocks,
                                                                        size_t payload_bytes) {
    __shared__ __align__(16) int16_t s_tile[kWarpsPerCta][4 * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t chunk = (size_t)blockIdx.x * kWarpsPerCta + warp;
    const size_t b0 = chunk * 32;
    if (b0 >= n_blocks) return;
    const int g = lane & 7, q = lane >> 3;
    uint32_t nat[8];
#pragma unroll
    for (int i = 0; i < 8; i++) nat[i] = kZigZagDev[8 * g + i];

    const uint32_t my_header = b0 + lane < n_blocks ? headers[b0 + lane] : 0u;
    uint32_t my_bytes = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t code = (my_header >> (4 * j)) & 15u;
        my_bytes += code == 13u ? 12u : code == 14u ? 16u : code;
    }
    uint32_t incl = my_bytes;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const uint32_t my_off = incl - my_bytes;
    // memory safety only: offsets that mpegb200_pack_coeffs_vlen produced never need the clamps (mpegb200_vlen_validate)
    const size_t chunk_off = min((size_t)chunk_offsets[chunk], payload_bytes - 16);
    const uint8_t* base = payload + chunk_off;
    const uint32_t off_max = (uint32_t)min(payload_bytes - 16 - chunk_off, (size_t)0xffffffffu);
    int16_t* tile = s_tile[warp];

#pragma unroll 1
    for (int round = 0; round < 8; round++) {
        const int src = 4 * round + q;
        const uint32_t h = __shfl_sync(0xffffffffu, my_header, src);
        uint32_t off = __shfl_sync(0xffffffffu, my_off, src);
        if (b0 + 4 * round >= n_blocks) break;   // warp-uniform
        const uint32_t code = (h >> (4 * g)) & 15u;
        const uint32_t w = code == 14u ? 16u : code - (code == 13u ? 1u : 0u);
        {   // the group's offset inside its block: exclusive scan of the widths over the block's eight lanes
            uint32_t incl_w = w;
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const uint32_t up_w = __shfl_up_sync(0xffffffffu, incl_w, d, 8);
                if (g >= d) incl_w += up_w;
            }
            off += incl_w - w;
        }
        int v[8];
        if (w == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0;
        } else if (code == 14u) {   // eight raw 16-bit values (a level outside 12 bits: an intra DC of a damaged stream)
            const uint8_t* p = base + min(off, off_max);
            const uint32_t* p4 = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
            uint32_t r[5];
#pragma unroll
            for (int k = 0; k < 5; k++) r[k] = __ldg(p4 + k);   // 16 bytes of padding follow the payload, the device buffer has slack behind
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t word = __funnelshift_r(r[k], r[k + 1], sh);
                v[2 * k] = (int)(int16_t)(word & 0xffffu);
                v[2 * k + 1] = (int)(int16_t)(word >> 16);
            }
        } else {
            const uint8_t* p = base + min(off, off_max);
            const uint32_t* p4 = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
            // the payload buffer is padded, so the fourth word is always readable
            uint32_t r0 = __ldg(p4), r1 = __ldg(p4 + 1), r2 = __ldg(p4 + 2);
            const uint32_t r3 = __ldg(p4 + 3);
            r0 = __funnelshift_r(r0, r1, sh);   // the group's (at most 96) bits, aligned
            r1 = __funnelshift_r(r1, r2, sh);
            r2 = __funnelshift_r(r2, r3, sh);
            const uint32_t up = 32u - w;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v[i] = (int)(r0 << up) >> up;   // sign-extended w-bit field
                r0 = __funnelshift_r(r0, r1, w);
                r1 = __funnelshift_r(r1, r2, w);
                r2 >>= w;
            }
            if (code != 13u) {   // undo c = (x + sign(x)) / 2: x = 2c - sign(c)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int c = v[i];
                    const int x = 2 * c - ((c >> 31) | 1);   // right for c != 0
                    v[i] = c == 0 ? 0 : x;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) tile[q * 64 + nat[i]] = (int16_t)v[i];
        __syncwarp();
        const size_t blk = b0 + 4 * round + (lane >> 3);   // 16 bytes per lane: block lane/8, part lane%8
        if (blk < n_blocks) out[blk * 8 + (lane & 7)] = reinterpret_cast<const uint4*>(tile)[lane];
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_expand_vlen(const uint32_t* d_headers, const uint64_t* d_chunk_offsets, const uint8_t* d_payload,
                               int16_t* d_coeffs, size_t n_blocks, size_t payload_bytes, cudaStream_t stream) {
    if (n_blocks == 0) return cudaSuccess;
    if (payload_bytes < 16) return cudaErrorInvalidValue;
    const size_t chunks = (n_blocks + 31) / 32;
    expand_vlen_kernel<<<(unsigned)((chunks + kWarpsPerCta - 1) / kWarpsPerCta), 32 * kWarpsPerCta, 0, stream>>>(
        d_headers, d_chunk_offsets, d_payload, reinterpret_cast<uint4*>(d_coeffs), n_blocks, payload_bytes);
    return cudaGetLastError();
}

}  // namespace mpegb200
