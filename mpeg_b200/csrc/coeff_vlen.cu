// coeff_vlen.cu -- variable-width transfer form of the coefficient blocks (include/mpegb200.h, "vlen"): the host
// packer and the kernel that expands it on the device to the int16[64] working form the decode kernel reads.
//
// Why: the end-to-end rate of the host-pointer entry points is bound by PCIe, and the coefficients are nine tenths of
// the upload.  What the reference leaves in blockData before the premultiply (video.go:729-741) is, for every
// coefficient the bitstream carried, an ODD number (the "oddification" of :732-736), and 0 for the ones it did not
// carry; only an intra DC (dc*8, :672) is even.  So a block is sent as eight groups of eight values in zig-zag order
// (the order the bitstream has them in: neighbours have similar magnitude), each group with its own width w:
//     code c = (x + sign(x)) / 2      (0 -> 0, +-1 -> +-1, +-3 -> +-2, ...: a bijection from {0, odd} to the integers)
//     w      = bits of the widest c of the group in two's complement (0 when all eight are zero)
// eight values of w bits are exactly w bytes, so groups never straddle bytes.  A group holding an even non-zero value
// (intra DC) travels raw, 12 bits per value.  A block is a 32-bit header (eight 4-bit group codes: 0..12 = w, 13 = raw)
// plus sum(w) payload bytes; 32 blocks form a chunk with one 64-bit payload offset, offsets inside a chunk come from a
// warp scan over the headers.  The dense blocks of the benchmark shrink from 128 (96 in the 12-bit form) to about 49
// bytes, typical sparse blocks of a real stream to 4 + a few bytes.
#include <cstdint>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

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

inline int width_of(int c) {  // bits of c in two's complement
    const unsigned a = (unsigned)(c < 0 ? ~c : c);
    int b = 1;
    while ((a >> (b - 1)) != 0) b++;
    return b;
}

// header and payload size of one block; returns false if a value does not fit 12 bits
inline bool plan_block(const int16_t* blk, uint32_t* header, uint32_t* bytes) {
    uint32_t h = 0, total = 0;
    for (int g = 0; g < 8; g++) {
        int w = 0;
        bool raw = false, any = false;
        for (int i = 0; i < 8; i++) {
            const int x = blk[kZigZag.nat[8 * g + i]];
            if (x < -2048 || x > 2047) return false;
            if (x == 0) continue;
            any = true;
            if ((x & 1) == 0) { raw = true; continue; }
            const int c = (x + (x > 0 ? 1 : -1)) / 2;
            const int b = width_of(c);
            if (b > w) w = b;
        }
        const uint32_t code = raw ? 13u : (any ? (uint32_t)w : 0u);
        h |= code << (4 * g);
        total += raw ? 12u : (any ? (uint32_t)w : 0u);
    }
    *header = h;
    *bytes = total;
    return true;
}

inline void write_block(const int16_t* blk, uint32_t header, uint8_t* out) {
    for (int g = 0; g < 8; g++) {
        const uint32_t code = (header >> (4 * g)) & 15u;
        if (code == 0) continue;
        const int w = code == 13 ? 12 : (int)code;
        uint64_t lo = 0, hi = 0;  // up to 96 bits
        for (int i = 0; i < 8; i++) {
            const int x = blk[kZigZag.nat[8 * g + i]];
            const int v = code == 13 ? x : (x == 0 ? 0 : (x + (x > 0 ? 1 : -1)) / 2);
            const uint64_t u = (uint64_t)((uint32_t)v & ((1u << w) - 1u));
            const int bit = i * w;
            if (bit < 64) {
                lo |= u << bit;
                if (bit + w > 64) hi |= u >> (64 - bit);
            } else {
                hi |= u << (bit - 64);
            }
        }
        for (int k = 0; k < w; k++) out[k] = (uint8_t)(k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8)));
        out += w;
    }
}

template <class F>
void parallel_for(size_t n, F f) {
    unsigned t = std::thread::hardware_concurrency();
    if (t == 0) t = 1;
    if (t > 32) t = 32;
    if (n < 4096) t = 1;
    if (t == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const size_t per = ((n + t - 1) / t + 31) / 32 * 32;  // whole chunks
    for (unsigned i = 0; i < t; i++) {
        const size_t lo = (size_t)i * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

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
        my_bytes += code == 13u ? 12u : code;
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
        const uint32_t w = code - (code == 13u ? 1u : 0u);
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

extern "C" {

size_t mpegb200_vlen_payload_bound(size_t n_blocks) { return n_blocks * 96 + 16; }

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
            if (code > 13u) return MPEGB200_ERECORD;
            run += code == 13u ? 12u : code;
        }
    }
    return run + 16 == payload_bytes ? 0 : MPEGB200_ERECORD;
}

int mpegb200_pack_coeffs_vlen(const int16_t* coeffs, size_t n_blocks, uint32_t* headers, uint64_t* chunk_offsets,
                              uint8_t* payload, size_t payload_cap, size_t* payload_bytes) {
    using namespace mpegb200;
    if (!payload_bytes || (n_blocks && (!coeffs || !headers || !chunk_offsets || !payload))) return MPEGB200_EINVAL;
    const size_t chunks = (n_blocks + 31) / 32;
    std::vector<uint32_t> chunk_bytes(chunks, 0);
    std::atomic<bool> ok{true};
    parallel_for(n_blocks, [&](size_t lo, size_t hi) {  // ranges are whole chunks
        for (size_t b = lo; b < hi; b++) {
            uint32_t bytes = 0;
            if (!plan_block(coeffs + b * 64, &headers[b], &bytes)) {
                ok = false;
                return;
            }
            chunk_bytes[b / 32] += bytes;
        }
    });
    if (!ok) return MPEGB200_ERECORD;
    uint64_t run = 0;
    for (size_t c = 0; c < chunks; c++) {
        chunk_offsets[c] = run;
        run += chunk_bytes[c];
    }
    *payload_bytes = (size_t)run + 16;   // 16 bytes of padding: the kernel reads whole words around a group
    if (*payload_bytes > payload_cap) return MPEGB200_EINVAL;
    parallel_for(n_blocks, [&](size_t lo, size_t hi) {
        for (size_t c = lo / 32; c * 32 < hi; c++) {
            uint8_t* out = payload + chunk_offsets[c];
            for (size_t b = c * 32; b < hi && b < c * 32 + 32; b++) {
                write_block(coeffs + b * 64, headers[b], out);
                uint32_t bytes = 0;
                for (int g = 0; g < 8; g++) {
                    const uint32_t code = (headers[b] >> (4 * g)) & 15u;
                    bytes += code == 13u ? 12u : code;
                }
                out += bytes;
            }
        }
    });
    memset(payload + run, 0, 16);
    return 0;
}

}  // extern "C"
