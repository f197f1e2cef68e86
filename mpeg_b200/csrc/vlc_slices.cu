// vlc_slices.cu -- slice-parallel MPEG-1 VLC stage on the device (include/mpegb200.h "slice-parallel VLC stage").
//
// One thread walks one slice (vlc_slice_walk.h: the reference's decodeSlice .. decodeBlock, video.go:436-746, table-driven
// like the host parser) and writes packed records -- mpegb200_mb + int16[64] per coded block -- into the slots the slice
// owns.  Slices are independent in MPEG-1 (byte-aligned start code, every predictor reset, video.go:436-446); what is NOT
// independent (overlapping or over-running slices, a picture ended early) is detected, never emulated: vlc_check_kernel
// flags such pictures, disables their records and the caller re-parses them on the host.
//
// Shape of the kernel.  The walk is a chain of dependent look-ups (window -> table -> shift -> next window): latency bound,
// and the slices of a warp diverge.  So a warp carries only a FEW slices (`lanes` of its 32 lanes, chosen by the launcher so
// that the wave fills the SMs with sixteen warps each) and the idle lanes only help to load the tables: more warps hide
// the latency, fewer slices per warp serialise less.  All code tables (47 KB) live in shared memory; the bitstream is read
// through L1 with one aligned 32-bit load per 32 bits consumed, issued one word ahead; a block is assembled in shared memory
// (128 bytes per slice, 16-byte chunks swizzled by lane) and leaves as eight 16-byte stores.
#include "common.cuh"
#include "vlc_slice_walk.h"

namespace mpegb200 {

namespace {
#ifndef MPEGB200_VLC_WARPS
#define MPEGB200_VLC_WARPS 4
#endif
#ifndef MPEGB200_VLC_CTAS
#define MPEGB200_VLC_CTAS 4
#endif
constexpr int kVlcWarps = MPEGB200_VLC_WARPS, kVlcCtasPerSm = MPEGB200_VLC_CTAS;   // warps per CTA; CTAs per SM (51 KB of shared memory each, <= 128 registers)
constexpr int kVlcMaxLanes = 8;                               // slices per warp, at most
constexpr size_t kVlcSmem = sizeof(VlcDeviceTables) + (size_t)kVlcWarps * kVlcMaxLanes * 128;
static_assert(sizeof(VlcDeviceTables) % 16 == 0, "the block scratch behind the tables must be 16-byte aligned");
}

__global__ void __launch_bounds__(kVlcWarps * 32, kVlcCtasPerSm) vlc_parse_kernel(
    const VlcDeviceTables* __restrict__ tables, int lanes, const mpegb200_vlc_picture* __restrict__ pics, int n_pics,
    const mpegb200_vlc_slice* __restrict__ slices, uint32_t n_slices, const uint32_t* __restrict__ words, uint32_t n_words,
    const uint8_t* __restrict__ quant, uint32_t n_quant, const StreamInfo* __restrict__ streams, int max_streams,
    mpegb200_mb* __restrict__ mbs, uint32_t n_mb_slots, int16_t* __restrict__ coeffs, SliceSummary* __restrict__ summary,
    const ResidentStream* __restrict__ resident) {
    extern __shared__ __align__(16) uint8_t smem[];
    VlcDeviceTables* const T = reinterpret_cast<VlcDeviceTables*>(smem);
    uint8_t* const s_block = smem + sizeof(VlcDeviceTables);
    {
        const uint4* src = reinterpret_cast<const uint4*>(tables);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(VlcDeviceTables) / 16); i += kVlcWarps * 32) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < kVlcWarps * kVlcMaxLanes * 8; i += kVlcWarps * 32) reinterpret_cast<uint4*>(s_block)[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= lanes) return;
    const uint32_t s = (blockIdx.x * kVlcWarps + warp) * (uint32_t)lanes + (uint32_t)lane;
    if (s >= n_slices) return;

    const mpegb200_vlc_slice sl = slices[s];
    // the entry point has checked the slot tables on the host; a slice whose slots are off anyway writes nothing
    const bool slots_ok = sl.pic < (uint32_t)n_pics && (sl.mb_slot & 15u) == 0 && (sl.mb_cap & 15u) == 0 && sl.mb_slot <= n_mb_slots &&
                          sl.mb_cap <= n_mb_slots - sl.mb_slot;
    bool ok = slots_ok && (resident != nullptr || sl.data_offset < (uint64_t)n_words * 4u);
    mpegb200_vlc_picture P;
    VlcGeometry g;
    if (ok) {
        P = pics[sl.pic];
        ok = P.stream >= 0 && P.stream < max_streams && P.quant < n_quant && P.dst_buf < 3 && P.fwd_buf < 3 && P.bwd_buf < 3 &&
             P.type >= MPEGB200_PIC_I && P.type <= MPEGB200_PIC_B && P.fwd_r_size < 7 && P.bwd_r_size < 7;
        if (ok) {
            const StreamInfo si = streams[P.stream];
            ok = si.open != 0;
            g.mb_w = si.mb_w;
            g.mb_h = si.mb_h;
            g.luma_w = si.luma_w;
            g.luma_h = si.luma_h;
            g.buf_bytes = si.buf_bytes;
            if (resident) {   // the slice's bytes come from its stream's own copy in device memory
                const ResidentStream rs = resident[P.stream];
                words = reinterpret_cast<const uint32_t*>(rs.bytes);
                n_words = rs.n_words;
                ok = ok && rs.bytes != nullptr && sl.data_offset < (uint64_t)n_words * 4u;
            }
        }
    }
    SliceSummary sum;
    if (ok) {
        sum = walk_slice(T, T->coef_fast, T->zigzag, s_block + (warp * kVlcMaxLanes + lane) * 128, (uint32_t)lane & 7u, sl, P, g, words,
                         n_words, quant, mbs, coeffs);
    } else {
        // a picture that is void (type 0: the caller withdrew it) or inconsistent: its slots hold null records, its picture is flagged
        sum.flags = MPEGB200_VLC_BAD_ARG;
        sum.first_addr = sum.last_addr = -1;
        sum.end_addr = 0;
        if (slots_ok)
            for (uint32_t slot = sl.mb_slot; slot < sl.mb_slot + sl.mb_cap; slot++) vlc_store16(mbs + slot, 0u, 0u, 0xffffu << 16, 6u * sl.mb_slot);
    }
    reinterpret_cast<uint4*>(summary)[s] = make_uint4(sum.flags, (uint32_t)sum.first_addr, (uint32_t)sum.last_addr, (uint32_t)sum.end_addr);
}

// One thread per picture.  A flagged picture loses its stream id, which makes the decode kernels drop every record of it
// (their own validity check), and reports its flags.
__global__ void vlc_check_kernel(const mpegb200_vlc_picture* __restrict__ vpics, mpegb200_picture* __restrict__ pics, int n_pics,
                                 const SliceSummary* __restrict__ summary, uint32_t n_slices,
                                 const StreamInfo* __restrict__ streams, int max_streams, int32_t* __restrict__ flags_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pics) return;
    const mpegb200_vlc_picture P = vpics[p];
    uint32_t flags;
    if (P.stream < 0 || P.stream >= max_streams || P.first_slice > n_slices || P.n_slices > n_slices - P.first_slice) {
        flags = MPEGB200_VLC_BAD_ARG;
    } else {
        const StreamInfo si = streams[P.stream];
        flags = vlc_check_picture(P, summary, (int)si.mb_w * (int)si.mb_h);
    }
    mpegb200_picture out;
    out.stream = flags ? -1 : P.stream;
    out.type = P.type;
    out.dst_buf = P.dst_buf;
    out.fwd_buf = P.fwd_buf;
    out.bwd_buf = P.bwd_buf;
    out.first_mb = P.mb_slot;
    out.n_mb = P.n_mb_slots;
    pics[p] = out;
    flags_out[p] = (int32_t)flags;
}

// Every start code prefix 00 00 01 xx of a resident stream (buffer.go:279-302 for all positions at once): one thread per 16
// bytes, positions appended through an atomic counter (the host sorts them: a few thousand per stream).
__global__ void startcode_index_kernel(const uint8_t* __restrict__ p, uint64_t len, uint64_t* __restrict__ out, uint32_t cap,
                                       uint32_t* __restrict__ count) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16u;
    if (i0 >= len) return;
    // bytes i0 .. i0 + 17 (the buffer is padded with zeros: 32 readable bytes behind len)
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p + i0));
    const uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(p + i0 + 16));
    const uint32_t w[5] = {a.x, a.y, a.z, a.w, b};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // three consecutive bytes, little-endian words
        const uint32_t lo = w[k >> 2], hi = w[(k >> 2) + 1];
        const uint32_t three = (uint32_t)((((uint64_t)hi << 32) | lo) >> (8 * (k & 3))) & 0xffffffu;
        if (three == 0x010000u && i0 + (uint64_t)k + 5u <= len) {
            const uint32_t at = atomicAdd(count, 1u);
            if (at < cap) out[at] = i0 + (uint64_t)k;
        }
    }
}

cudaError_t launch_startcode_index(const uint8_t* d_bytes, uint64_t len, uint64_t* d_out, uint32_t cap, uint32_t* d_count,
                                   cudaStream_t stream) {
    if (len) {
        const uint64_t threads = (len + 15) / 16;
        startcode_index_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_bytes, len, d_out, cap, d_count);
    }
    return cudaGetLastError();
}

size_t vlc_summary_bytes(size_t n_slices) { return sizeof(SliceSummary) * (n_slices ? n_slices : 1); }

cudaError_t configure_vlc_kernel() {
    return cudaFuncSetAttribute(vlc_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVlcSmem);
}

cudaError_t launch_vlc_parse(const VlcDeviceTables* d_tables, const mpegb200_vlc_picture* d_vpics, mpegb200_picture* d_pics,
                             int n_pics, const mpegb200_vlc_slice* d_slices, uint32_t n_slices, const uint8_t* d_bitstream,
                             uint32_t n_words, const uint8_t* d_quant, uint32_t n_quant, const StreamInfo* d_streams,
                             int max_streams, mpegb200_mb* d_mbs, uint32_t n_mb_slots, int16_t* d_coeffs, void* d_summary,
                             int32_t* d_flags, int sm_count, cudaStream_t stream, const ResidentStream* d_resident) {
    if (n_slices) {
        // slices per warp: as few as put the whole wave on the SMs at once (sixteen resident warps each)
        const uint32_t resident = (uint32_t)sm_count * kVlcCtasPerSm * kVlcWarps;
        int lanes = (int)((n_slices + resident - 1) / resident);
        lanes = lanes < 1 ? 1 : (lanes > kVlcMaxLanes ? kVlcMaxLanes : lanes);
#ifdef MPEGB200_EXPERIMENTS
        if (const char* e = getenv("MPEGB200_VLC_LANES")) lanes = atoi(e) < 1 ? 1 : (atoi(e) > kVlcMaxLanes ? kVlcMaxLanes : atoi(e));
#endif
        const uint32_t per_cta = (uint32_t)(kVlcWarps * lanes);
        vlc_parse_kernel<<<(n_slices + per_cta - 1) / per_cta, kVlcWarps * 32, kVlcSmem, stream>>>(
            d_tables, lanes, d_vpics, n_pics, d_slices, n_slices, reinterpret_cast<const uint32_t*>(d_bitstream), n_words, d_quant,
            n_quant, d_streams, max_streams, d_mbs, n_mb_slots, d_coeffs, reinterpret_cast<SliceSummary*>(d_summary), d_resident);
    }
    if (n_pics)
        vlc_check_kernel<<<(n_pics + 127) / 128, 128, 0, stream>>>(d_vpics, d_pics, n_pics,
                                                                  reinterpret_cast<const SliceSummary*>(d_summary), n_slices,
                                                                  d_streams, max_streams, d_flags);
    return cudaGetLastError();
}

}  // namespace mpegb200
