// ctx.cu -- the C-ABI of include/mpegb200.h: context, device memory, transfers, launches.
// No CPU fallback anywhere: if CUDA is unavailable every entry point reports MPEGB200_ECUDA.
#include <cuda.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "mp2_window.inc"
#include "vlc_device_tables.h"

namespace mpegb200 {
cudaError_t configure_audio_kernel();
}
using namespace mpegb200;

struct HostStream {
    bool open = false;
    int width = 0, height = 0;
    int luma_w = 0, luma_h = 0, chroma_w = 0, chroma_h = 0;
    size_t luma_bytes = 0, chroma_bytes = 0, buf_bytes = 0, buf_stride = 0;
    uint8_t* dev = nullptr;  // first of the three buffers (inside the slab)
    int slab = -1, slot = -1;
};

// Streams of one geometry share one allocation ("slab") so that a single pair of tensor maps
// (SlabMaps) addresses every frame buffer of the slab: TMA coordinate z = slot*3 + buffer.
struct Slab {
    bool alive = false;
    int width = 0, height = 0;
    size_t buf_stride = 0;
    int capacity = 0, used = 0;
    std::vector<uint8_t> slot_used;
    uint8_t* dev = nullptr;
    bool tma_ok = false;
};
static const int kMaxSlabs = 1024;

struct DevBuf {  // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
};

struct mpegb200_ctx {
    int device = 0;
    int max_streams = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::vector<HostStream> vs;
    std::vector<StreamInfo> h_info;
    StreamInfo* d_info = nullptr;
    bool info_dirty = false;
    std::vector<uint8_t> audio_open;
    AudioState* d_audio = nullptr;
    float* d_window = nullptr;
    DevBuf s_pics[2], s_mbs[2], s_coeffs[2], s_packed[2], s_headers[2], s_chunks[2], s_ids, s_bufs, s_rgba, s_samples, s_out, s_plans, s_ainfo, s_acodes;
    // host-pointer pipeline: uploads and read-backs run on their own streams so that the H2D copy of the
    // next step overlaps the kernels and the D2H copy of the current one (double-buffered staging)
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_kernel = nullptr;
    // one read-back event per physical buffer index: a decode only waits for the read-backs of the buffers it writes
    cudaEvent_t ev_down[3] = {nullptr, nullptr, nullptr};
    uint64_t upload_seq = 0;
    bool down_pending[3] = {false, false, false};
    std::vector<Slab> slabs;
    SlabMaps* d_maps = nullptr;   // kMaxSlabs entries
    int n_generic_streams = 0;    // open streams that cannot use the TMA kernel (odd mb_w or encode failure)
    int n_tma_streams = 0;        // open streams that can
    bool force_generic = false;   // MPEGB200_FUSED=generic (A/B measurements)
    void* encode_fn = nullptr;    // cuTensorMapEncodeTiled
    uint64_t launches = 0;
    bool validate = false;                   // mpegb200_set_validate / MPEGB200_VALIDATE=1
    bool kernel_timing = false;              // mpegb200_set_kernel_timing
    std::vector<cudaEvent_t> timing_events;  // three per timed decode call
    // slice-parallel VLC stage (mpegb200_video_decode_bitstream): tables, double-buffered uploads, the records it leaves on the device
    VlcDeviceTables* d_vlc_tables = nullptr;
    DevBuf s_vpics[2], s_slices[2], s_bits[2], s_quant[2], s_vlc_pics, s_vlc_mbs, s_vlc_coeffs, s_vlc_summary, s_vlc_flags;
    int32_t* h_vlc_flags = nullptr;          // pinned
    size_t h_vlc_flags_cap = 0;
    int vlc_n_pictures = 0, sm_count = 148;
    size_t vlc_n_mb_slots = 0;
    cudaEvent_t ev_vlc_flags = nullptr, ev_vlc_t0 = nullptr, ev_vlc_t1 = nullptr;
    bool vlc_timed = false;
    // elementary streams resident in device memory (mpegb200_video_stream_upload): per stream id its bytes and length
    std::vector<void*> resident_dev;
    std::vector<size_t> resident_len;
    std::vector<ResidentStream> h_resident;
    ResidentStream* d_resident = nullptr;
    bool resident_dirty = false;
    DevBuf s_index;
    int max_w = 0, max_h = 0;
    char err[512] = {0};
};

static int fail(mpegb200_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, MPEGB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                            \
    } while (0)

static int ensure(mpegb200_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->up_stream) CU(cudaStreamSynchronize(ctx->up_stream));
        if (ctx->down_stream) CU(cudaStreamSynchronize(ctx->down_stream));
        CU(cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = bytes + bytes / 4 + 256;
    if (cudaMalloc(&b.p, cap) != cudaSuccess) {
        b.p = nullptr;
        cudaGetLastError();
        return fail(ctx, MPEGB200_ENOMEM, "device allocation of %zu bytes failed", cap);
    }
    b.cap = cap;
    return 0;
}

// the compute stream must not overwrite a frame buffer that an asynchronous read-back is still copying
// (buf_mask: bit b = the work about to be enqueued writes physical buffer b of some stream; with the reference's
// rotation (video.go:406-433) the buffer a picture is decoded into is never the one just handed out for reading, so
// the read-back of step k overlaps the decode of step k+1)
static int join_readback(mpegb200_ctx* ctx, unsigned buf_mask = 7u) {
    for (int b = 0; b < 3; b++)
        if ((buf_mask >> b & 1u) && ctx->down_pending[b]) {
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_down[b], 0));
            ctx->down_pending[b] = false;
        }
    return 0;
}

// which physical buffers a batch of pictures (host copy) writes
static unsigned dst_buffers(const mpegb200_picture* pics, int n_pictures) {
    unsigned m = 0;
    for (int i = 0; i < n_pictures; i++) m |= pics[i].dst_buf < 3 ? 1u << pics[i].dst_buf : 7u;
    return m;
}

static void drop_timing_events(mpegb200_ctx* ctx) {
    for (cudaEvent_t e : ctx->timing_events) cudaEventDestroy(e);
    ctx->timing_events.clear();
}

static int flush_info(mpegb200_ctx* ctx) {
    if (!ctx->info_dirty) return 0;
    // stream-ordered so that it lands before any kernel enqueued afterwards; h_info outlives the copy
    CU(cudaMemcpyAsync(ctx->d_info, ctx->h_info.data(), sizeof(StreamInfo) * ctx->max_streams, cudaMemcpyHostToDevice,
                       ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->info_dirty = false;
    return 0;
}

extern "C" {

int mpegb200_abi_version(void) { return MPEGB200_ABI_VERSION; }

mpegb200_ctx* mpegb200_create(int device, int max_streams, int* err) try {
    auto set = [&](int e) {
        if (err) *err = e;
    };
    if (max_streams <= 0 || max_streams > (1 << 20)) {
        set(MPEGB200_EINVAL);
        return nullptr;
    }
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        set(MPEGB200_ECUDA);
        return nullptr;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        prop.major != 10) {  // the kernels are sm_100a only: no other path exists
        cudaGetLastError();
        set(MPEGB200_ECUDA);
        return nullptr;
    }
    mpegb200_ctx* ctx = new (std::nothrow) mpegb200_ctx();
    if (!ctx) {
        set(MPEGB200_ENOMEM);
        return nullptr;
    }
    ctx->device = device;
    ctx->max_streams = max_streams;
    ctx->vs.resize(max_streams);
    ctx->h_info.assign(max_streams, StreamInfo{});
    ctx->audio_open.assign(max_streams, 0);
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->own_stream = ok;
    ok = ok && cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; i++) {
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_kernel, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; b < 3; b++) ok = ok && cudaEventCreateWithFlags(&ctx->ev_down[b], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_info, sizeof(StreamInfo) * max_streams) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_info, 0, sizeof(StreamInfo) * max_streams) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_window, sizeof(float) * 1024) == cudaSuccess;
    if (ok) {
        float w[1024];
        for (int i = 0; i < 512; i++) w[i] = w[i + 512] = (float)kSynthesisWindowX2[i] * 0.5f;  // audio.go:95-98
        ok = cudaMemcpy(ctx->d_window, w, sizeof(w), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&ctx->d_maps, sizeof(SlabMaps) * kMaxSlabs) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_maps, 0, sizeof(SlabMaps) * kMaxSlabs) == cudaSuccess;
    if (ok) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ctx->encode_fn = nullptr;
        cudaGetLastError();
#ifdef MPEGB200_EXPERIMENTS
        const char* sel = getenv("MPEGB200_FUSED");   // experiment builds: "generic" sends every stream to the cp.async kernel
        ctx->force_generic = sel && strcmp(sel, "generic") == 0;
#endif
    }
    {
        const char* val = getenv("MPEGB200_VALIDATE");   // debug aid: same as mpegb200_set_validate(ctx, 1)
        ctx->validate = val && val[0] == '1';
    }
    ok = ok && configure_kernels() == cudaSuccess && configure_audio_kernel() == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        mpegb200_destroy(ctx);
        set(MPEGB200_ECUDA);
        return nullptr;
    }
    set(MPEGB200_OK);
    return ctx;
} catch (...) {  // bad_alloc from the host-side bookkeeping must not cross the C boundary
    if (err) *err = MPEGB200_ENOMEM;
    return nullptr;
}

void mpegb200_destroy(mpegb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& sl : ctx->slabs)
        if (sl.dev) cudaFree(sl.dev);
    if (ctx->d_maps) cudaFree(ctx->d_maps);
    if (ctx->up_stream) cudaStreamSynchronize(ctx->up_stream);
    if (ctx->down_stream) cudaStreamSynchronize(ctx->down_stream);
    for (DevBuf* b : {&ctx->s_pics[0], &ctx->s_pics[1], &ctx->s_mbs[0], &ctx->s_mbs[1], &ctx->s_coeffs[0],
                      &ctx->s_coeffs[1], &ctx->s_ids, &ctx->s_bufs, &ctx->s_rgba, &ctx->s_samples, &ctx->s_out,
                      &ctx->s_plans, &ctx->s_ainfo, &ctx->s_acodes, &ctx->s_packed[0], &ctx->s_packed[1], &ctx->s_headers[0], &ctx->s_headers[1],
                      &ctx->s_chunks[0], &ctx->s_chunks[1], &ctx->s_vpics[0], &ctx->s_vpics[1], &ctx->s_slices[0], &ctx->s_slices[1],
                      &ctx->s_bits[0], &ctx->s_bits[1], &ctx->s_quant[0], &ctx->s_quant[1], &ctx->s_vlc_pics, &ctx->s_vlc_mbs,
                      &ctx->s_vlc_coeffs, &ctx->s_vlc_summary, &ctx->s_vlc_flags})
        if (b->p) cudaFree(b->p);
    if (ctx->d_vlc_tables) cudaFree(ctx->d_vlc_tables);
    for (void* r : ctx->resident_dev)
        if (r) cudaFree(r);
    if (ctx->d_resident) cudaFree(ctx->d_resident);
    if (ctx->s_index.p) cudaFree(ctx->s_index.p);
    if (ctx->h_vlc_flags) cudaFreeHost(ctx->h_vlc_flags);
    for (cudaEvent_t e : {ctx->ev_vlc_flags, ctx->ev_vlc_t0, ctx->ev_vlc_t1})
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
        if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
    }
    drop_timing_events(ctx);
    if (ctx->ev_kernel) cudaEventDestroy(ctx->ev_kernel);
    for (int b = 0; b < 3; b++)
        if (ctx->ev_down[b]) cudaEventDestroy(ctx->ev_down[b]);
    if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
    if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
    if (ctx->d_info) cudaFree(ctx->d_info);
    if (ctx->d_audio) cudaFree(ctx->d_audio);
    if (ctx->d_window) cudaFree(ctx->d_window);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mpegb200_last_error(mpegb200_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int mpegb200_set_stream(mpegb200_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MPEGB200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return 0;
}

void* mpegb200_get_stream(mpegb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int mpegb200_sync(mpegb200_ctx* ctx) {
    if (!ctx) return MPEGB200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->up_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->down_stream));
    for (int b = 0; b < 3; b++) ctx->down_pending[b] = false;
    return 0;
}

int mpegb200_sync_uploads(mpegb200_ctx* ctx) {
    if (!ctx) return MPEGB200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->up_stream));
    return 0;
}

int mpegb200_join_readbacks(mpegb200_ctx* ctx) {
    if (!ctx) return MPEGB200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    return join_readback(ctx, 7u);
}

uint64_t mpegb200_launch_count(mpegb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mpegb200_set_validate(mpegb200_ctx* ctx, int on) {
    if (!ctx) return MPEGB200_EINVAL;
    ctx->validate = on != 0;
    return 0;
}

int mpegb200_set_kernel_timing(mpegb200_ctx* ctx, int on) {
    if (!ctx) return MPEGB200_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (!on) {
        CU(cudaStreamSynchronize(ctx->stream));
        drop_timing_events(ctx);
    }
    ctx->kernel_timing = on != 0;
    return 0;
}

int mpegb200_kernel_times(mpegb200_ctx* ctx, float* plan_ms, float* fused_ms, int cap) {
    if (!ctx || cap < 0) return fail(ctx, MPEGB200_EINVAL, "bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const int calls = (int)(ctx->timing_events.size() / 3), n = calls < cap ? calls : cap;
    for (int i = 0; i < n; i++) {
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, ctx->timing_events[3 * i], ctx->timing_events[3 * i + 1]));
        CU(cudaEventElapsedTime(&b, ctx->timing_events[3 * i + 1], ctx->timing_events[3 * i + 2]));
        if (plan_ms) plan_ms[i] = a;
        if (fused_ms) fused_ms[i] = b;
    }
    drop_timing_events(ctx);
    return n;
}

/* ---------------------------------------------------------------------------------------- video */

static HostStream* vstream(mpegb200_ctx* ctx, int stream) {
    if (!ctx || stream < 0 || stream >= ctx->max_streams) return nullptr;
    return &ctx->vs[stream];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Window tensor maps of a slab (see SlabMaps in common.cuh).  Returns false if the driver refuses them.
static bool encode_slab_maps(mpegb200_ctx* ctx, const Slab& sl, const HostStream& g, SlabMaps* out) {
    if (!ctx->encode_fn) return false;
    EncodeTiledFn enc = (EncodeTiledFn)ctx->encode_fn;
    const cuuint32_t ones[3] = {1, 1, 1};
    {
        const cuuint64_t dims[3] = {(cuuint64_t)g.luma_w + 32, sl.buf_stride / g.luma_w, (cuuint64_t)3 * sl.capacity};
        const cuuint64_t strides[2] = {(cuuint64_t)g.luma_w, sl.buf_stride};
        const cuuint32_t box[3] = {32, kLumaBoxRows, 1};
        if (enc((CUtensorMap*)out->luma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, sl.dev, dims, strides, box, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        // rank 4: {x (overlapping rows), y (to the end of the buffer), plane (Cb, Cr), 3*slot+buffer}
        const cuuint64_t dims[4] = {(cuuint64_t)g.chroma_w + 32, (sl.buf_stride - g.luma_bytes - g.chroma_bytes) / g.chroma_w, 2,
                                    (cuuint64_t)3 * sl.capacity};
        const cuuint64_t strides[3] = {(cuuint64_t)g.chroma_w, g.chroma_bytes, sl.buf_stride};
        const cuuint32_t box[4] = {32, kChromaBoxRows, 2, 1}, ones4[4] = {1, 1, 1, 1};
        if (enc((CUtensorMap*)out->chroma, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, sl.dev + g.luma_bytes, dims, strides, box,
                ones4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    // the strip maps: same bytes, 8-byte elements (luma_w + 32 and chroma_w + 32 are multiples of 8)
    {
        const cuuint64_t dims[3] = {((cuuint64_t)g.luma_w + 32) / 8, sl.buf_stride / g.luma_w, (cuuint64_t)3 * sl.capacity};
        const cuuint64_t strides[2] = {(cuuint64_t)g.luma_w, sl.buf_stride};
        const cuuint32_t box[3] = {kStripLW / 8, kStripLH, 1};
        if (enc((CUtensorMap*)out->luma_strip, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, sl.dev, dims, strides, box, ones,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    {
        const cuuint64_t dims[4] = {((cuuint64_t)g.chroma_w + 32) / 8, (sl.buf_stride - g.luma_bytes - g.chroma_bytes) / g.chroma_w, 2,
                                    (cuuint64_t)3 * sl.capacity};
        const cuuint64_t strides[3] = {(cuuint64_t)g.chroma_w, g.chroma_bytes, sl.buf_stride};
        const cuuint32_t box[4] = {kStripCW / 8, kStripCH, 2, 1}, ones4[4] = {1, 1, 1, 1};
        if (enc((CUtensorMap*)out->chroma_strip, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, sl.dev + g.luma_bytes, dims, strides, box,
                ones4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

int mpegb200_video_open(mpegb200_ctx* ctx, int stream, int width, int height) try {
    HostStream* s = vstream(ctx, stream);
    if (!s) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (s->open) return fail(ctx, MPEGB200_ESTATE, "video stream %d already open", stream);
    if (width <= 0 || height <= 0 || width > 4095 || height > 4095)  // 12-bit fields, video.go:276-277
        return fail(ctx, MPEGB200_EINVAL, "bad picture size %dx%d", width, height);
    CU(cudaSetDevice(ctx->device));
    const int mb_w = (width + 15) >> 4, mb_h = (height + 15) >> 4;  // video.go:314-315
    HostStream g;
    g.width = width;
    g.height = height;
    g.luma_w = mb_w << 4;
    g.luma_h = mb_h << 4;
    g.chroma_w = mb_w << 3;
    g.chroma_h = mb_h << 3;
    g.luma_bytes = (size_t)g.luma_w * g.luma_h;
    g.chroma_bytes = (size_t)g.chroma_w * g.chroma_h;
    g.buf_bytes = g.luma_bytes + 2 * g.chroma_bytes + (size_t)g.luma_w * 16;  // video.go:340
    const size_t unit = (size_t)g.luma_w * 16;                                  // multiple of 256
    g.buf_stride = (g.buf_bytes + 64 + unit - 1) / unit * unit;

    // find a slab of this geometry with a free slot, or make one
    int si = -1;
    for (size_t i = 0; i < ctx->slabs.size(); i++) {
        const Slab& sl = ctx->slabs[i];
        if (sl.alive && sl.width == width && sl.height == height && sl.used < sl.capacity) {
            si = (int)i;
            break;
        }
    }
    if (si < 0) {
        for (size_t i = 0; i < ctx->slabs.size(); i++)
            if (!ctx->slabs[i].alive) {
                si = (int)i;
                break;
            }
        if (si < 0) {
            if ((int)ctx->slabs.size() >= kMaxSlabs) return fail(ctx, MPEGB200_ENOMEM, "too many geometry slabs");
            ctx->slabs.emplace_back();
            si = (int)ctx->slabs.size() - 1;
        }
        Slab& sl = ctx->slabs[si];
        sl = Slab{};
        const size_t per_stream = 3 * g.buf_stride;
        size_t cap = ((size_t)1 << 30) / per_stream;
        if (cap < 1) cap = 1;
        if (cap > (size_t)ctx->max_streams) cap = (size_t)ctx->max_streams;
        if (cap > 21000) cap = 21000;  // 3 * slot + buffer travels as 16 bits in the group plans
        sl.capacity = (int)cap;
        sl.width = width;
        sl.height = height;
        sl.buf_stride = g.buf_stride;
        sl.slot_used.assign(cap, 0);
        if (cudaMalloc(&sl.dev, cap * per_stream + 256) != cudaSuccess) {
            cudaGetLastError();
            sl.dev = nullptr;
            return fail(ctx, MPEGB200_ENOMEM, "frame-buffer slab for %dx%d (%zu bytes)", width, height, cap * per_stream);
        }
        CU(cudaMemsetAsync(sl.dev, 0, cap * per_stream + 256, ctx->stream));  // make([]byte, ...) zeroes, video.go:340
        sl.alive = true;
        SlabMaps maps;
        memset(&maps, 0, sizeof(maps));
        sl.tma_ok = (mb_w % 2 == 0) && encode_slab_maps(ctx, sl, g, &maps);  // chroma pitch must be a multiple of 16 B
        if (sl.tma_ok) CU(cudaMemcpy(&ctx->d_maps[si], &maps, sizeof(maps), cudaMemcpyHostToDevice));
    }
    Slab& sl = ctx->slabs[si];
    int slot = 0;
    while (sl.slot_used[slot]) slot++;
    sl.slot_used[slot] = 1;
    sl.used++;
    *s = g;
    s->slab = si;
    s->slot = slot;
    s->dev = sl.dev + (size_t)slot * 3 * g.buf_stride;
    CU(cudaMemsetAsync(s->dev, 0, 3 * g.buf_stride, ctx->stream));  // a re-used slot starts zeroed like a new one
    StreamInfo& info = ctx->h_info[stream];
    memset(&info, 0, sizeof(info));
    info.base = s->dev;
    info.buf_stride = (uint32_t)g.buf_stride;
    info.buf_bytes = (uint32_t)g.buf_bytes;
    info.luma_w = (uint16_t)g.luma_w;
    info.luma_h = (uint16_t)g.luma_h;
    info.mb_w = (uint16_t)mb_w;
    info.mb_h = (uint16_t)mb_h;
    info.width = (uint16_t)width;
    info.height = (uint16_t)height;
    info.slab = (uint16_t)si;
    info.slot = (uint16_t)slot;
    info.open = 1;
    info.tma_ok = sl.tma_ok ? 1 : 0;
    s->open = true;
    if (!sl.tma_ok) ctx->n_generic_streams++; else ctx->n_tma_streams++;
    ctx->info_dirty = true;
    if (width > ctx->max_w) ctx->max_w = width;
    if (height > ctx->max_h) ctx->max_h = height;
    return 0;
} catch (...) {  // bad_alloc from the host-side bookkeeping must not cross the C boundary
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

int mpegb200_video_close(mpegb200_ctx* ctx, int stream) {
    HostStream* s = vstream(ctx, stream);
    if (!s) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (!s->open) return fail(ctx, MPEGB200_ESTATE, "video stream %d not open", stream);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    Slab& sl = ctx->slabs[s->slab];
    sl.slot_used[s->slot] = 0;
    sl.used--;
    if (!sl.tma_ok) ctx->n_generic_streams--; else ctx->n_tma_streams--;
    if (sl.used == 0) {
        cudaFree(sl.dev);
        sl = Slab{};
    }
    *s = HostStream{};
    ctx->h_info[stream] = StreamInfo{};
    ctx->info_dirty = true;
    return 0;
}

int mpegb200_video_geometry(mpegb200_ctx* ctx, int stream, int* luma_w, int* luma_h, int* chroma_w, int* chroma_h,
                            size_t* frame_bytes) {
    HostStream* s = vstream(ctx, stream);
    if (!s) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (!s->open) return fail(ctx, MPEGB200_ESTATE, "video stream %d not open", stream);
    if (luma_w) *luma_w = s->luma_w;
    if (luma_h) *luma_h = s->luma_h;
    if (chroma_w) *chroma_w = s->chroma_w;
    if (chroma_h) *chroma_h = s->chroma_h;
    if (frame_bytes) *frame_bytes = s->buf_bytes;
    return 0;
}

static bool window_inside(long off, int stride, int size, int odd_h, int odd_v, long avail) {
    const long hi = off + (long)(size - 1 + odd_v) * stride + (size - 1 + odd_h);
    return off >= 0 && hi < avail;
}

int mpegb200_video_validate(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics, size_t n_mb,
                            const mpegb200_mb* mbs, size_t n_blocks) try {
    if (!ctx || n_pictures < 0 || (n_pictures && !pics) || (n_mb && !mbs))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (n_pictures > 65536) return fail(ctx, MPEGB200_ERECORD, "at most 65536 pictures per call (16-bit pic index)");
    std::vector<uint8_t> seen_stream(ctx->max_streams, 0);
    for (int p = 0; p < n_pictures; p++) {
        const mpegb200_picture& pic = pics[p];
        HostStream* s = vstream(ctx, pic.stream);
        if (!s || !s->open) return fail(ctx, MPEGB200_ERECORD, "picture %d: stream %d not open", p, pic.stream);
        if (seen_stream[pic.stream]++)
            return fail(ctx, MPEGB200_ERECORD, "picture %d: second picture of stream %d in one call", p, pic.stream);
        if (pic.dst_buf > 2 || pic.fwd_buf > 2 || pic.bwd_buf > 2)
            return fail(ctx, MPEGB200_ERECORD, "picture %d: buffer index out of range", p);
    }
    std::vector<std::vector<uint8_t>> written(n_pictures);
    uint64_t expect_block = n_mb ? mbs[0].coeff_block : 0;
    for (size_t i = 0; i < n_mb; i++) {
        const mpegb200_mb& m = mbs[i];
        if (m.pic >= n_pictures) return fail(ctx, MPEGB200_ERECORD, "mb %zu: picture index %u out of range", i, m.pic);
        const mpegb200_picture& pic = pics[m.pic];
        const HostStream& s = ctx->vs[pic.stream];
        const int mb_w = s.luma_w >> 4, mb_h = s.luma_h >> 4;
        if (m.mb_row >= mb_h || m.mb_col >= mb_w)
            return fail(ctx, MPEGB200_ERECORD, "mb %zu: position (%u,%u) outside %dx%d macroblocks", i, m.mb_row,
                        m.mb_col, mb_h, mb_w);
        if (m.cbp & ~0x3f) return fail(ctx, MPEGB200_ERECORD, "mb %zu: cbp has more than 6 bits", i);
        const bool intra = m.flags & MPEGB200_MB_INTRA, pred = m.flags & MPEGB200_MB_PREDICT;
        if (intra == pred) return fail(ctx, MPEGB200_ERECORD, "mb %zu: exactly one of INTRA / PREDICT must be set", i);
        if ((m.flags & MPEGB200_MB_REF_BWD) && !pred) return fail(ctx, MPEGB200_ERECORD, "mb %zu: REF_BWD without PREDICT", i);
        const int ncoded = __builtin_popcount(m.cbp);
        if (m.coeff_block != expect_block)
            return fail(ctx, MPEGB200_ERECORD, "mb %zu: coeff_block %u breaks the packing rule (expected %llu)", i,
                        m.coeff_block, (unsigned long long)expect_block);
        expect_block += ncoded;
        if (expect_block > n_blocks) return fail(ctx, MPEGB200_ERECORD, "mb %zu: coefficient blocks run past n_blocks", i);
        auto& w = written[m.pic];
        if (w.empty()) w.assign((size_t)mb_w * mb_h, 0);
        if (w[(size_t)m.mb_row * mb_w + m.mb_col]++)
            return fail(ctx, MPEGB200_ERECORD, "mb %zu: macroblock (%u,%u) written twice in picture %u", i, m.mb_row,
                        m.mb_col, m.pic);
        if (pred) {  // the windows copyMacroblock reads must lie inside the frame buffer (video_noasm.go:49-50)
            if (((m.flags & MPEGB200_MB_REF_BWD) ? pic.bwd_buf : pic.fwd_buf) == pic.dst_buf)
                return fail(ctx, MPEGB200_ERECORD, "mb %zu: reference buffer equals destination buffer", i);
            const long total = (long)s.buf_bytes;
            const int hp = m.mv_h >> 1, vp = m.mv_v >> 1;
            const long lsi = (long)((m.mb_row << 4) + vp) * s.luma_w + (m.mb_col << 4) + hp;
            const int cmh = m.mv_h / 2, cmv = m.mv_v / 2;
            const long csi = (long)((m.mb_row << 3) + (cmv >> 1)) * s.chroma_w + (m.mb_col << 3) + (cmh >> 1);
            const long cb0 = (long)s.luma_bytes, cr0 = cb0 + (long)s.chroma_bytes;
            if (!window_inside(lsi, s.luma_w, 16, m.mv_h & 1, m.mv_v & 1, total) ||
                !window_inside(csi, s.chroma_w, 8, cmh & 1, cmv & 1, total - cb0) ||
                !window_inside(csi, s.chroma_w, 8, cmh & 1, cmv & 1, total - cr0))
                return fail(ctx, MPEGB200_ERECORD, "mb %zu: motion vector (%d,%d) reads outside the frame buffer", i,
                            m.mv_h, m.mv_v);
        }
    }
    return 0;
} catch (...) {  // bad_alloc from the host-side bookkeeping must not cross the C boundary
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

static int decode_pictures_dev(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* d_pics, size_t n_mb,
                               const mpegb200_mb* d_mbs, size_t n_blocks, const int16_t* d_coeffs, unsigned dst_mask) try {
    if (!ctx || n_pictures < 0 || (n_mb && (!d_pics || !d_mbs)) || (n_blocks && !d_coeffs))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (n_mb > 0xffffffffull || n_blocks > 0xffffffffull || n_pictures > 65536)
        return fail(ctx, MPEGB200_EINVAL, "batch too large");
    if ((reinterpret_cast<uintptr_t>(d_mbs) & 15) || (reinterpret_cast<uintptr_t>(d_pics) & 15) ||
        (reinterpret_cast<uintptr_t>(d_coeffs) & 15))
        return fail(ctx, MPEGB200_EINVAL, "device arrays must be 16-byte aligned");
    if (n_mb == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    if (int rc = flush_info(ctx)) return rc;
    if (int rc = join_readback(ctx, dst_mask)) return rc;
    // The TMA kernel serves every stream whose chroma pitch is a multiple of 16 bytes; streams of odd macroblock width
    // go to the generic kernel.  A batch may mix both: each kernel skips the other's records.
    bool use_tma = !ctx->force_generic && ctx->encode_fn && ctx->n_tma_streams > 0;
    CUtensorMap coef_map;
    if (use_tma) {
        // 2-D view of the coefficient array: n_blocks rows of 64 int16, boxes of 32 rows, 128-byte swizzle
        const cuuint64_t dims[2] = {64, n_blocks ? (cuuint64_t)n_blocks : 1};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, 32}, ones[2] = {1, 1};
        void* base = n_blocks ? (void*)d_coeffs : (void*)ctx->d_window;  // any valid 16-byte aligned address when empty
        use_tma = ((EncodeTiledFn)ctx->encode_fn)(&coef_map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box,
                                                  ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (use_tma) {
        if (int rc = ensure(ctx, ctx->s_plans, fused_plan_bytes((uint32_t)n_mb))) return rc;
        const cudaEvent_t* timing = nullptr;
        if (ctx->kernel_timing) {
            cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
            for (int i = 0; i < 3; i++) CU(cudaEventCreate(&ev[i]));
            ctx->timing_events.insert(ctx->timing_events.end(), ev, ev + 3);
            timing = ctx->timing_events.data() + ctx->timing_events.size() - 3;
        }
        CU(launch_fused_tma(&coef_map, ctx->d_maps, ctx->s_plans.p, ctx->d_info, ctx->max_streams, d_pics, n_pictures,
                            d_mbs, (uint32_t)n_mb, (uint32_t)n_blocks, ctx->stream, timing));
        ctx->launches += 2;  // the plan pre-pass and the arithmetic kernel
    }
    if (!use_tma || ctx->n_generic_streams > 0) {
        CU(launch_fused_mc_idct(ctx->d_info, ctx->max_streams, d_pics, n_pictures, d_mbs, (uint32_t)n_mb, d_coeffs,
                                (uint32_t)n_blocks, ctx->stream, use_tma));
        ctx->launches++;
    }
    return 0;
} catch (...) {  // bad_alloc from the host-side bookkeeping must not cross the C boundary
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

// device-resident records: which buffers the pictures write is not known on the host, so every pending read-back is joined
int mpegb200_video_decode_pictures_dev(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* d_pics, size_t n_mb,
                                       const mpegb200_mb* d_mbs, size_t n_blocks, const int16_t* d_coeffs) {
    return decode_pictures_dev(ctx, n_pictures, d_pics, n_mb, d_mbs, n_blocks, d_coeffs, 7u);
}

int mpegb200_video_decode_pictures(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics, size_t n_mb,
                                   const mpegb200_mb* mbs, size_t n_blocks, const int16_t* coeffs) {
    if (!ctx || n_pictures < 0 || (n_mb && (!pics || !mbs)) || (n_blocks && !coeffs))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (ctx->validate) {  // records from an untrusted bitstream: fail loudly like the reference panics, before anything is enqueued
        if (int rc = mpegb200_video_validate(ctx, n_pictures, pics, n_mb, mbs, n_blocks)) return rc;
    }
    if (n_mb == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    const int slot = (int)(ctx->upload_seq++ & 1);
    if (int rc = ensure(ctx, ctx->s_pics[slot], sizeof(mpegb200_picture) * (size_t)n_pictures)) return rc;
    if (int rc = ensure(ctx, ctx->s_mbs[slot], sizeof(mpegb200_mb) * n_mb)) return rc;
    if (int rc = ensure(ctx, ctx->s_coeffs[slot], 128 * (n_blocks ? n_blocks : 1))) return rc;
    // the staging slot is free once the kernel that consumed it two calls ago has finished
    CU(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_free[slot], 0));
    CU(cudaMemcpyAsync(ctx->s_pics[slot].p, pics, sizeof(mpegb200_picture) * (size_t)n_pictures, cudaMemcpyHostToDevice,
                       ctx->up_stream));
    CU(cudaMemcpyAsync(ctx->s_mbs[slot].p, mbs, sizeof(mpegb200_mb) * n_mb, cudaMemcpyHostToDevice, ctx->up_stream));
    if (n_blocks)
        CU(cudaMemcpyAsync(ctx->s_coeffs[slot].p, coeffs, 128 * n_blocks, cudaMemcpyHostToDevice, ctx->up_stream));
    CU(cudaEventRecord(ctx->ev_up[slot], ctx->up_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[slot], 0));
    int rc = decode_pictures_dev(ctx, n_pictures, (const mpegb200_picture*)ctx->s_pics[slot].p, n_mb,
                                 (const mpegb200_mb*)ctx->s_mbs[slot].p, n_blocks, (const int16_t*)ctx->s_coeffs[slot].p,
                                 dst_buffers(pics, n_pictures));
    CU(cudaEventRecord(ctx->ev_free[slot], ctx->stream));
    return rc;
}

int mpegb200_pack_coeffs12(const int16_t* coeffs, size_t n_blocks, uint8_t* packed) {
    if ((!coeffs || !packed) && n_blocks) return MPEGB200_EINVAL;
    for (size_t b = 0; b < n_blocks; b++) {
        const int16_t* in = coeffs + b * 64;
        uint8_t* out = packed + b * 96;
        for (int i = 0; i < 64; i += 2) {  // two values -> three bytes
            const int a = in[i], c = in[i + 1];
            if (a < -2048 || a > 2047 || c < -2048 || c > 2047) return MPEGB200_ERECORD;
            const uint32_t u = ((uint32_t)a & 0xfffu) | (((uint32_t)c & 0xfffu) << 12);
            out[0] = (uint8_t)u;
            out[1] = (uint8_t)(u >> 8);
            out[2] = (uint8_t)(u >> 16);
            out += 3;
        }
    }
    return 0;
}

int mpegb200_video_decode_pictures_packed(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics, size_t n_mb,
                                          const mpegb200_mb* mbs, size_t n_blocks, const uint8_t* coeffs12) {
    if (!ctx || n_pictures < 0 || (n_mb && (!pics || !mbs)) || (n_blocks && !coeffs12))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (ctx->validate) {  // records from an untrusted bitstream: fail loudly like the reference panics, before anything is enqueued
        if (int rc = mpegb200_video_validate(ctx, n_pictures, pics, n_mb, mbs, n_blocks)) return rc;
    }
    if (n_mb == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    const int slot = (int)(ctx->upload_seq++ & 1);
    if (int rc = ensure(ctx, ctx->s_pics[slot], sizeof(mpegb200_picture) * (size_t)n_pictures)) return rc;
    if (int rc = ensure(ctx, ctx->s_mbs[slot], sizeof(mpegb200_mb) * n_mb)) return rc;
    if (int rc = ensure(ctx, ctx->s_packed[slot], 96 * (n_blocks ? n_blocks : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_coeffs[slot], 128 * (n_blocks ? n_blocks : 1))) return rc;
    CU(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_free[slot], 0));
    CU(cudaMemcpyAsync(ctx->s_pics[slot].p, pics, sizeof(mpegb200_picture) * (size_t)n_pictures, cudaMemcpyHostToDevice,
                       ctx->up_stream));
    CU(cudaMemcpyAsync(ctx->s_mbs[slot].p, mbs, sizeof(mpegb200_mb) * n_mb, cudaMemcpyHostToDevice, ctx->up_stream));
    if (n_blocks)
        CU(cudaMemcpyAsync(ctx->s_packed[slot].p, coeffs12, 96 * n_blocks, cudaMemcpyHostToDevice, ctx->up_stream));
    CU(cudaEventRecord(ctx->ev_up[slot], ctx->up_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[slot], 0));
    CU(launch_unpack12((const uint8_t*)ctx->s_packed[slot].p, (int16_t*)ctx->s_coeffs[slot].p, n_blocks, ctx->stream));
    ctx->launches++;
    int rc = decode_pictures_dev(ctx, n_pictures, (const mpegb200_picture*)ctx->s_pics[slot].p, n_mb,
                                 (const mpegb200_mb*)ctx->s_mbs[slot].p, n_blocks, (const int16_t*)ctx->s_coeffs[slot].p,
                                 dst_buffers(pics, n_pictures));
    CU(cudaEventRecord(ctx->ev_free[slot], ctx->stream));
    return rc;
}

int mpegb200_video_decode_pictures_vlen(mpegb200_ctx* ctx, int n_pictures, const mpegb200_picture* pics, size_t n_mb,
                                        const mpegb200_mb* mbs, size_t n_blocks, const uint32_t* headers,
                                        const uint64_t* chunk_offsets, const uint8_t* payload, size_t payload_bytes) {
    if (!ctx || n_pictures < 0 || (n_mb && (!pics || !mbs)) || (n_blocks && (!headers || !chunk_offsets || !payload)))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (n_blocks && payload_bytes < 16) return fail(ctx, MPEGB200_EINVAL, "payload without its 16 bytes of padding");
    if (ctx->validate) {  // records from an untrusted bitstream: fail loudly like the reference panics, before anything is enqueued
        if (int rc = mpegb200_video_validate(ctx, n_pictures, pics, n_mb, mbs, n_blocks)) return rc;
        if (int rc = mpegb200_vlen_validate(headers, chunk_offsets, n_blocks, payload_bytes)) return fail(ctx, rc, "malformed variable-width coefficient stream");
    }
    if (n_mb == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    const size_t chunks = (n_blocks + 31) / 32;
    const int slot = (int)(ctx->upload_seq++ & 1);
    if (int rc = ensure(ctx, ctx->s_pics[slot], sizeof(mpegb200_picture) * (size_t)n_pictures)) return rc;
    if (int rc = ensure(ctx, ctx->s_mbs[slot], sizeof(mpegb200_mb) * n_mb)) return rc;
    if (int rc = ensure(ctx, ctx->s_headers[slot], 4 * (n_blocks ? n_blocks : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_chunks[slot], 8 * (chunks ? chunks : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_packed[slot], payload_bytes ? payload_bytes : 16)) return rc;
    if (int rc = ensure(ctx, ctx->s_coeffs[slot], 128 * (n_blocks ? n_blocks : 1))) return rc;
    CU(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_free[slot], 0));
    CU(cudaMemcpyAsync(ctx->s_pics[slot].p, pics, sizeof(mpegb200_picture) * (size_t)n_pictures, cudaMemcpyHostToDevice,
                       ctx->up_stream));
    CU(cudaMemcpyAsync(ctx->s_mbs[slot].p, mbs, sizeof(mpegb200_mb) * n_mb, cudaMemcpyHostToDevice, ctx->up_stream));
    if (n_blocks) {
        CU(cudaMemcpyAsync(ctx->s_headers[slot].p, headers, 4 * n_blocks, cudaMemcpyHostToDevice, ctx->up_stream));
        CU(cudaMemcpyAsync(ctx->s_chunks[slot].p, chunk_offsets, 8 * chunks, cudaMemcpyHostToDevice, ctx->up_stream));
        CU(cudaMemcpyAsync(ctx->s_packed[slot].p, payload, payload_bytes, cudaMemcpyHostToDevice, ctx->up_stream));
    }
    CU(cudaEventRecord(ctx->ev_up[slot], ctx->up_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[slot], 0));
    CU(launch_expand_vlen((const uint32_t*)ctx->s_headers[slot].p, (const uint64_t*)ctx->s_chunks[slot].p,
                          (const uint8_t*)ctx->s_packed[slot].p, (int16_t*)ctx->s_coeffs[slot].p, n_blocks, payload_bytes,
                          ctx->stream));
    if (n_blocks) ctx->launches++;
    int rc = decode_pictures_dev(ctx, n_pictures, (const mpegb200_picture*)ctx->s_pics[slot].p, n_mb,
                                 (const mpegb200_mb*)ctx->s_mbs[slot].p, n_blocks, (const int16_t*)ctx->s_coeffs[slot].p,
                                 dst_buffers(pics, n_pictures));
    CU(cudaEventRecord(ctx->ev_free[slot], ctx->stream));
    return rc;
}

// ---- slice-parallel VLC stage: compressed slices in, records parsed and executed on the device ----

int mpegb200_video_stream_upload(mpegb200_ctx* ctx, int stream, const uint8_t* data, size_t len) try {
    if (!ctx || stream < 0 || stream >= ctx->max_streams || (len && !data)) return fail(ctx, MPEGB200_EINVAL, "bad argument");
    if (len >= 0xffffff00ull) return fail(ctx, MPEGB200_EINVAL, "stream too long for one resident buffer (2^32 bytes)");
    CU(cudaSetDevice(ctx->device));
    if (ctx->resident_dev.empty()) {
        ctx->resident_dev.assign((size_t)ctx->max_streams, nullptr);
        ctx->resident_len.assign((size_t)ctx->max_streams, 0);
        ctx->h_resident.assign((size_t)ctx->max_streams, ResidentStream{nullptr, 0, 0});
        if (cudaMalloc(&ctx->d_resident, sizeof(ResidentStream) * (size_t)ctx->max_streams) != cudaSuccess) {
            ctx->d_resident = nullptr;
            cudaGetLastError();
            ctx->resident_dev.clear();
            return fail(ctx, MPEGB200_ENOMEM, "device allocation failed");
        }
        ctx->resident_dirty = true;
    }
    if (ctx->resident_dev[(size_t)stream]) {   // a wave in flight may still read the old bytes
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(ctx->resident_dev[(size_t)stream]));
        ctx->resident_dev[(size_t)stream] = nullptr;
        ctx->resident_len[(size_t)stream] = 0;
        ctx->h_resident[(size_t)stream] = ResidentStream{nullptr, 0, 0};
        ctx->resident_dirty = true;
    }
    if (len == 0) return 0;
    void* d = nullptr;
    const size_t padded = ((len + 3) & ~(size_t)3) + 64;
    if (cudaMalloc(&d, padded) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MPEGB200_ENOMEM, "device allocation of %zu bytes failed", padded);
    }
    cudaError_t e = cudaMemsetAsync((uint8_t*)d + (len & ~(size_t)3), 0, padded - (len & ~(size_t)3), ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, data, len, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // `data` may go away when this returns
    if (e != cudaSuccess) {
        cudaFree(d);
        return fail(ctx, MPEGB200_ECUDA, "upload of the stream failed: %s", cudaGetErrorString(e));
    }
    ctx->resident_dev[(size_t)stream] = d;
    ctx->resident_len[(size_t)stream] = len;
    ctx->h_resident[(size_t)stream] = ResidentStream{(const uint8_t*)d, (uint32_t)(len / 4 + 2), 0};
    ctx->resident_dirty = true;
    return 0;
} catch (...) {
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

int mpegb200_video_stream_index(mpegb200_ctx* ctx, int stream, uint64_t* positions, size_t cap, size_t* n) try {
    if (!ctx || !n || stream < 0 || stream >= ctx->max_streams || (cap && !positions)) return fail(ctx, MPEGB200_EINVAL, "bad argument");
    *n = 0;
    if (ctx->resident_dev.empty() || !ctx->resident_dev[(size_t)stream]) return fail(ctx, MPEGB200_ESTATE, "stream %d is not resident", stream);
    if (cap > 0x7fffffffull) cap = 0x7fffffffull;
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure(ctx, ctx->s_index, 8 * cap + 16)) return rc;
    uint32_t* d_count = (uint32_t*)ctx->s_index.p;
    uint64_t* d_out = (uint64_t*)((uint8_t*)ctx->s_index.p + 16);
    CU(cudaMemsetAsync(d_count, 0, 16, ctx->stream));
    CU(launch_startcode_index((const uint8_t*)ctx->resident_dev[(size_t)stream], ctx->resident_len[(size_t)stream], d_out, (uint32_t)cap,
                              d_count, ctx->stream));
    ctx->launches++;
    uint32_t count = 0;
    CU(cudaMemcpyAsync(&count, d_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n = count;
    const size_t got = count < cap ? count : cap;
    if (got) CU(cudaMemcpy(positions, d_out, 8 * got, cudaMemcpyDeviceToHost));
    std::sort(positions, positions + got);
    if (count > cap) return fail(ctx, MPEGB200_EINVAL, "%u start codes, room for %zu", count, cap);
    return 0;
} catch (...) {
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

int mpegb200_video_decode_bitstream(mpegb200_ctx* ctx, int n_pictures, const mpegb200_vlc_picture* pics, size_t n_slices,
                                    const mpegb200_vlc_slice* slices, const uint8_t* bitstream, size_t bitstream_bytes,
                                    const uint8_t* quant, size_t n_quant, size_t n_mb_slots) try {
    if (!ctx || n_pictures < 0 || (n_pictures && !pics) || (n_slices && (!slices || !quant)))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    const bool resident = bitstream == nullptr;   // the slices read their streams' resident copies
    if (resident) {
        bitstream_bytes = 0;
        if (n_slices && ctx->resident_dev.empty()) return fail(ctx, MPEGB200_ESTATE, "no resident stream (mpegb200_video_stream_upload) and no bitstream given");
    }
    if (n_pictures >= 65535 || n_slices > 0x7fffffffull || bitstream_bytes >= 0xffffff00ull || n_mb_slots > 0xffffffffull / 6 ||
        (n_mb_slots & 15) || n_quant > 0xffffffffull)
        return fail(ctx, MPEGB200_EINVAL, "wave too large (or record slots not a multiple of 16)");
    // the slot tables must tile the wave's record slots in order: a hole would be executed as records
    size_t expect = 0;
    for (size_t i = 0; i < n_slices; i++) {
        const mpegb200_vlc_slice& sl = slices[i];
        if (sl.mb_slot != expect || (sl.mb_cap & 15u) || sl.pic >= (uint32_t)n_pictures || (!resident && sl.data_offset >= bitstream_bytes))
            return fail(ctx, MPEGB200_EINVAL, "slice %zu: record slots must follow each other in multiples of 16, picture and offset must be in range", i);
        expect += sl.mb_cap;
    }
    if (expect != n_mb_slots) return fail(ctx, MPEGB200_EINVAL, "the slices' record slots add up to %zu, not %zu", expect, n_mb_slots);
    unsigned dst_mask = 0;
    for (int i = 0; i < n_pictures; i++) {
        const mpegb200_vlc_picture& P = pics[i];
        if (P.quant >= n_quant || P.first_slice > n_slices || P.n_slices > n_slices - P.first_slice)
            return fail(ctx, MPEGB200_EINVAL, "picture %d: quantiser or slice range out of range", i);
        dst_mask |= P.dst_buf < 3 ? 1u << P.dst_buf : 7u;
        if (resident && P.n_slices && (P.stream < 0 || P.stream >= ctx->max_streams || !ctx->resident_dev[(size_t)P.stream]))
            return fail(ctx, MPEGB200_ESTATE, "picture %d: stream %d is not resident", i, P.stream);
    }
    ctx->vlc_n_pictures = 0;   // until this wave is enqueued: a failed call leaves no flags to read
    ctx->vlc_n_mb_slots = 0;
    if (n_pictures == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_vlc_tables) {   // once per context
        VlcDeviceTables* h = new VlcDeviceTables;
        const bool ok = fill_vlc_device_tables(h);
        cudaError_t e = ok ? cudaMalloc(&ctx->d_vlc_tables, sizeof(VlcDeviceTables)) : cudaErrorUnknown;
        if (e == cudaSuccess) e = cudaMemcpy(ctx->d_vlc_tables, h, sizeof(VlcDeviceTables), cudaMemcpyHostToDevice);
        delete h;
        if (e != cudaSuccess) {
            if (ctx->d_vlc_tables) cudaFree(ctx->d_vlc_tables);
    for (void* r : ctx->resident_dev)
        if (r) cudaFree(r);
    if (ctx->d_resident) cudaFree(ctx->d_resident);
    if (ctx->s_index.p) cudaFree(ctx->s_index.p);
            ctx->d_vlc_tables = nullptr;
            return fail(ctx, ok ? MPEGB200_ECUDA : MPEGB200_ESTATE, "variable-length-code tables: %s", ok ? cudaGetErrorString(e) : "host tables have an unexpected shape");
        }
        CU(configure_vlc_kernel());
        CU(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
        CU(cudaEventCreateWithFlags(&ctx->ev_vlc_flags, cudaEventDisableTiming));
        CU(cudaEventCreate(&ctx->ev_vlc_t0));
        CU(cudaEventCreate(&ctx->ev_vlc_t1));
    }
    if ((size_t)n_pictures > ctx->h_vlc_flags_cap) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_vlc_flags) cudaFreeHost(ctx->h_vlc_flags);
        ctx->h_vlc_flags = nullptr;
        ctx->h_vlc_flags_cap = 0;
        const size_t cap = (size_t)n_pictures + (size_t)n_pictures / 4 + 64;
        if (cudaMallocHost(&ctx->h_vlc_flags, cap * sizeof(int32_t)) != cudaSuccess) {
            ctx->h_vlc_flags = nullptr;
            cudaGetLastError();
            return fail(ctx, MPEGB200_ENOMEM, "pinned allocation failed");
        }
        ctx->h_vlc_flags_cap = cap;
    }
    const int slot = (int)(ctx->upload_seq++ & 1);
    const size_t n_blocks = 6 * n_mb_slots;
    if (int rc = ensure(ctx, ctx->s_vpics[slot], sizeof(mpegb200_vlc_picture) * (size_t)n_pictures)) return rc;
    if (int rc = ensure(ctx, ctx->s_slices[slot], sizeof(mpegb200_vlc_slice) * (n_slices ? n_slices : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_bits[slot], bitstream_bytes + 32)) return rc;
    if (int rc = ensure(ctx, ctx->s_quant[slot], 128 * (n_quant ? n_quant : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_vlc_pics, sizeof(mpegb200_picture) * (size_t)n_pictures)) return rc;
    if (int rc = ensure(ctx, ctx->s_vlc_mbs, sizeof(mpegb200_mb) * (n_mb_slots ? n_mb_slots : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_vlc_coeffs, 128 * (n_blocks ? n_blocks : 1))) return rc;
    if (int rc = ensure(ctx, ctx->s_vlc_summary, vlc_summary_bytes(n_slices))) return rc;
    if (int rc = ensure(ctx, ctx->s_vlc_flags, sizeof(int32_t) * (size_t)n_pictures)) return rc;
    CU(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_free[slot], 0));
    CU(cudaMemcpyAsync(ctx->s_vpics[slot].p, pics, sizeof(mpegb200_vlc_picture) * (size_t)n_pictures, cudaMemcpyHostToDevice, ctx->up_stream));
    if (n_slices) {
        CU(cudaMemcpyAsync(ctx->s_slices[slot].p, slices, sizeof(mpegb200_vlc_slice) * n_slices, cudaMemcpyHostToDevice, ctx->up_stream));
        CU(cudaMemcpyAsync(ctx->s_quant[slot].p, quant, 128 * n_quant, cudaMemcpyHostToDevice, ctx->up_stream));
        if (!resident) {
            // the bit reader loads whole 32-bit words: zero what lies behind the last byte, then the bytes
            CU(cudaMemsetAsync((uint8_t*)ctx->s_bits[slot].p + (bitstream_bytes & ~(size_t)3), 0, 16, ctx->up_stream));
            CU(cudaMemcpyAsync(ctx->s_bits[slot].p, bitstream, bitstream_bytes, cudaMemcpyHostToDevice, ctx->up_stream));
        }
    }
    if (resident && ctx->resident_dirty) {
        CU(cudaMemcpyAsync(ctx->d_resident, ctx->h_resident.data(), sizeof(ResidentStream) * (size_t)ctx->max_streams, cudaMemcpyHostToDevice,
                           ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->resident_dirty = false;
    }
    CU(cudaEventRecord(ctx->ev_up[slot], ctx->up_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[slot], 0));
    if (int rc = flush_info(ctx)) return rc;
    ctx->vlc_timed = ctx->kernel_timing;
    if (ctx->vlc_timed) CU(cudaEventRecord(ctx->ev_vlc_t0, ctx->stream));
    CU(launch_vlc_parse(ctx->d_vlc_tables, (const mpegb200_vlc_picture*)ctx->s_vpics[slot].p, (mpegb200_picture*)ctx->s_vlc_pics.p,
                        n_pictures, (const mpegb200_vlc_slice*)ctx->s_slices[slot].p, (uint32_t)n_slices,
                        (const uint8_t*)ctx->s_bits[slot].p, (uint32_t)(bitstream_bytes / 4 + 2), (const uint8_t*)ctx->s_quant[slot].p,
                        (uint32_t)n_quant, ctx->d_info, ctx->max_streams, (mpegb200_mb*)ctx->s_vlc_mbs.p, (uint32_t)n_mb_slots,
                        (int16_t*)ctx->s_vlc_coeffs.p, ctx->s_vlc_summary.p, (int32_t*)ctx->s_vlc_flags.p, ctx->sm_count, ctx->stream,
                        resident ? ctx->d_resident : nullptr));
    ctx->launches += n_slices ? 2 : 1;
    if (ctx->vlc_timed) CU(cudaEventRecord(ctx->ev_vlc_t1, ctx->stream));
    CU(cudaEventRecord(ctx->ev_free[slot], ctx->stream));   // the uploads are consumed: the staging slot may be refilled
    CU(cudaMemcpyAsync(ctx->h_vlc_flags, ctx->s_vlc_flags.p, sizeof(int32_t) * (size_t)n_pictures, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev_vlc_flags, ctx->stream));
    ctx->vlc_n_pictures = n_pictures;
    ctx->vlc_n_mb_slots = n_mb_slots;
    return decode_pictures_dev(ctx, n_pictures, (const mpegb200_picture*)ctx->s_vlc_pics.p, n_mb_slots,
                               (const mpegb200_mb*)ctx->s_vlc_mbs.p, n_blocks, (const int16_t*)ctx->s_vlc_coeffs.p, dst_mask);
} catch (...) {
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

int mpegb200_video_bitstream_flags(mpegb200_ctx* ctx, int* flags_out, int n) {
    if (!ctx || n < 0 || (n && !flags_out)) return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (n != ctx->vlc_n_pictures) return fail(ctx, MPEGB200_ESTATE, "the last wave had %d pictures, not %d", ctx->vlc_n_pictures, n);
    if (n == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_vlc_flags));
    int flagged = 0;
    for (int i = 0; i < n; i++) {
        flags_out[i] = ctx->h_vlc_flags[i];
        flagged += flags_out[i] != 0;
    }
    return flagged;
}

int mpegb200_video_bitstream_records(mpegb200_ctx* ctx, mpegb200_mb* mbs, int16_t* coeffs) {
    if (!ctx) return MPEGB200_EINVAL;
    if (ctx->vlc_n_pictures == 0 || ctx->vlc_n_mb_slots == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (mbs) CU(cudaMemcpy(mbs, ctx->s_vlc_mbs.p, sizeof(mpegb200_mb) * ctx->vlc_n_mb_slots, cudaMemcpyDeviceToHost));
    if (coeffs) CU(cudaMemcpy(coeffs, ctx->s_vlc_coeffs.p, 128 * 6 * ctx->vlc_n_mb_slots, cudaMemcpyDeviceToHost));
    return 0;
}

int mpegb200_video_bitstream_parse_ms(mpegb200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return MPEGB200_EINVAL;
    if (!ctx->vlc_timed || ctx->vlc_n_pictures == 0) return fail(ctx, MPEGB200_ESTATE, "the last wave was not timed (mpegb200_set_kernel_timing)");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->ev_vlc_t1));
    CU(cudaEventElapsedTime(ms, ctx->ev_vlc_t0, ctx->ev_vlc_t1));
    return 0;
}

static int check_buf(mpegb200_ctx* ctx, int stream, int buf, HostStream** out) {
    HostStream* s = vstream(ctx, stream);
    if (!s) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (!s->open) return fail(ctx, MPEGB200_ESTATE, "video stream %d not open", stream);
    if (buf < 0 || buf > 2) return fail(ctx, MPEGB200_EINVAL, "buffer index %d out of range", buf);
    *out = s;
    return 0;
}

int mpegb200_video_read_planes(mpegb200_ctx* ctx, int stream, int buf, uint8_t* y, uint8_t* cb, uint8_t* cr) {
    HostStream* s = nullptr;
    if (int rc = check_buf(ctx, stream, buf, &s)) return rc;
    CU(cudaSetDevice(ctx->device));
    const uint8_t* base = s->dev + (size_t)buf * s->buf_stride;
    if (y) CU(cudaMemcpyAsync(y, base, s->luma_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (cb) CU(cudaMemcpyAsync(cb, base + s->luma_bytes, s->chroma_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (cr)
        CU(cudaMemcpyAsync(cr, base + s->luma_bytes + s->chroma_bytes, s->chroma_bytes, cudaMemcpyDeviceToHost,
                           ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mpegb200_video_write_planes(mpegb200_ctx* ctx, int stream, int buf, const uint8_t* y, const uint8_t* cb,
                                const uint8_t* cr) {
    HostStream* s = nullptr;
    if (int rc = check_buf(ctx, stream, buf, &s)) return rc;
    CU(cudaSetDevice(ctx->device));
    if (int rc = join_readback(ctx, 1u << buf)) return rc;
    uint8_t* base = s->dev + (size_t)buf * s->buf_stride;
    if (y) CU(cudaMemcpyAsync(base, y, s->luma_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (cb) CU(cudaMemcpyAsync(base + s->luma_bytes, cb, s->chroma_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (cr)
        CU(cudaMemcpyAsync(base + s->luma_bytes + s->chroma_bytes, cr, s->chroma_bytes, cudaMemcpyHostToDevice,
                           ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mpegb200_video_read_frame(mpegb200_ctx* ctx, int stream, int buf, uint8_t* dst, size_t dst_bytes) {
    HostStream* s = nullptr;
    if (int rc = check_buf(ctx, stream, buf, &s)) return rc;
    if (!dst || dst_bytes < s->buf_bytes) return fail(ctx, MPEGB200_EINVAL, "destination smaller than %zu bytes", s->buf_bytes);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, s->dev + (size_t)buf * s->buf_stride, s->buf_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mpegb200_video_write_frame(mpegb200_ctx* ctx, int stream, int buf, const uint8_t* src, size_t src_bytes) {
    HostStream* s = nullptr;
    if (int rc = check_buf(ctx, stream, buf, &s)) return rc;
    if (!src || src_bytes != s->buf_bytes) return fail(ctx, MPEGB200_EINVAL, "source must be %zu bytes", s->buf_bytes);
    CU(cudaSetDevice(ctx->device));
    if (int rc = join_readback(ctx, 1u << buf)) return rc;
    CU(cudaMemcpyAsync(s->dev + (size_t)buf * s->buf_stride, src, s->buf_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int read_pictures(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs, uint8_t* dst,
                         size_t dst_stride, cudaMemcpyKind kind) {
    if (!ctx || n < 0 || (n && (!streams || !bufs || !dst))) return fail(ctx, MPEGB200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    // host read-backs run on the download stream, behind everything enqueued on the compute stream so far
    cudaStream_t q = ctx->stream;
    if (kind == cudaMemcpyDeviceToHost) {
        CU(cudaEventRecord(ctx->ev_kernel, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->down_stream, ctx->ev_kernel, 0));
        q = ctx->down_stream;
    }
    // Streams that sit at equal distances in one slab (consecutive slots: the normal case of a batch) and hand out the
    // same buffer index go as ONE strided copy: a copy-engine operation costs a few microseconds however small, and a
    // batch of 256 pictures should not pay that 256 times.
    for (int i = 0; i < n;) {
        HostStream* s = nullptr;
        if (int rc = check_buf(ctx, streams[i], bufs[i], &s)) return rc;
        const size_t bytes = s->luma_bytes + 2 * s->chroma_bytes;
        if (bytes > dst_stride) return fail(ctx, MPEGB200_EINVAL, "stride %zu smaller than a picture (%zu)", dst_stride, bytes);
        const uint8_t* src = s->dev + (size_t)bufs[i] * s->buf_stride;
        int j = i + 1;
        ptrdiff_t pitch = 0;
        const HostStream* prev = s;
        for (; j < n; j++) {
            HostStream* t = nullptr;
            if (int rc = check_buf(ctx, streams[j], bufs[j], &t)) return rc;
            const ptrdiff_t d = t->dev - prev->dev;
            if (bufs[j] != bufs[i] || t->luma_bytes != s->luma_bytes || t->chroma_bytes != s->chroma_bytes ||
                t->buf_stride != s->buf_stride || d < (ptrdiff_t)bytes || d > (ptrdiff_t)0x7fffffff || (pitch && d != pitch))
                break;
            pitch = d;
            prev = t;
        }
        if (j - i >= 2)
            CU(cudaMemcpy2DAsync(dst + (size_t)i * dst_stride, dst_stride, src, (size_t)pitch, bytes, (size_t)(j - i), kind, q));
        else
            CU(cudaMemcpyAsync(dst + (size_t)i * dst_stride, src, bytes, kind, q));
        i = j;
    }
    if (kind == cudaMemcpyDeviceToHost) {
        unsigned m = 0;
        for (int i = 0; i < n; i++) m |= 1u << bufs[i];
        for (int b = 0; b < 3; b++)
            if (m >> b & 1u) {
                CU(cudaEventRecord(ctx->ev_down[b], ctx->down_stream));
                ctx->down_pending[b] = true;
            }
    }
    return 0;
}

int mpegb200_video_read_pictures_host(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                      uint8_t* dst, size_t dst_stride) {
    return read_pictures(ctx, n, streams, bufs, dst, dst_stride, cudaMemcpyDeviceToHost);
}

int mpegb200_video_read_pictures_dev(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                     uint8_t* d_dst, size_t dst_stride) {
    return read_pictures(ctx, n, streams, bufs, d_dst, dst_stride, cudaMemcpyDeviceToDevice);
}

void* mpegb200_video_frame_dev(mpegb200_ctx* ctx, int stream, int buf) {
    HostStream* s = nullptr;
    if (check_buf(ctx, stream, buf, &s)) return nullptr;
    return s->dev + (size_t)buf * s->buf_stride;
}

int mpegb200_video_rgba_batch_dev(mpegb200_ctx* ctx, int n, const int32_t* streams, const uint8_t* bufs,
                                  uint8_t* d_rgba, size_t rgba_stride_bytes) {
    if (!ctx || n < 0 || (n && (!streams || !bufs || !d_rgba))) return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (n == 0) return 0;
    int max_w = 0, max_h = 0;
    for (int i = 0; i < n; i++) {
        HostStream* s = nullptr;
        if (int rc = check_buf(ctx, streams[i], bufs[i], &s)) return rc;
        if ((size_t)s->width * s->height * 4 > rgba_stride_bytes)
            return fail(ctx, MPEGB200_EINVAL, "rgba stride %zu too small for stream %d", rgba_stride_bytes, streams[i]);
        if (s->width > max_w) max_w = s->width;
        if (s->height > max_h) max_h = s->height;
    }
    CU(cudaSetDevice(ctx->device));
    if (int rc = flush_info(ctx)) return rc;
    // small pageable copies are staged by the driver at call time: asynchronous and safe to reuse
    const size_t id_bytes = sizeof(int32_t) * (size_t)n;
    if (int rc = ensure(ctx, ctx->s_ids, id_bytes)) return rc;
    if (int rc = ensure(ctx, ctx->s_bufs, (size_t)n)) return rc;
    CU(cudaMemcpyAsync(ctx->s_ids.p, streams, id_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_bufs.p, bufs, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_rgba(ctx->d_info, ctx->max_streams, (const int32_t*)ctx->s_ids.p, (const uint8_t*)ctx->s_bufs.p, n, max_w,
                   max_h, d_rgba, rgba_stride_bytes, ctx->stream));
    ctx->launches += 1;
    return 0;
}

int mpegb200_video_rgba(mpegb200_ctx* ctx, int stream, int buf, uint8_t* rgba) {
    HostStream* s = nullptr;
    if (int rc = check_buf(ctx, stream, buf, &s)) return rc;
    if (!rgba) return fail(ctx, MPEGB200_EINVAL, "null destination");
    const size_t bytes = (size_t)s->width * s->height * 4;
    CU(cudaSetDevice(ctx->device));
    if (int rc = ensure(ctx, ctx->s_rgba, bytes)) return rc;
    const int32_t id = stream;
    const uint8_t b = (uint8_t)buf;
    if (int rc = mpegb200_video_rgba_batch_dev(ctx, 1, &id, &b, (uint8_t*)ctx->s_rgba.p, bytes)) return rc;
    CU(cudaMemcpyAsync(rgba, ctx->s_rgba.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* ---------------------------------------------------------------------------------------- display ring */

}  // extern "C"

struct mpegb200_ring {
    mpegb200_ctx* ctx = nullptr;
    int n = 0, depth = 0;
    uint64_t head = 0;
    size_t stride = 0;
    std::vector<int32_t> streams;
    uint8_t* dev = nullptr;
    cudaEvent_t pushed = nullptr;
};

extern "C" {

mpegb200_ring* mpegb200_video_ring_new(mpegb200_ctx* ctx, int n, const int32_t* streams, int depth) try {
    if (!ctx || n <= 0 || !streams || depth <= 0 || depth > 4096) return nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    size_t stride = 0;
    for (int i = 0; i < n; i++) {
        HostStream* s = vstream(ctx, streams[i]);
        if (!s || !s->open) {
            fail(ctx, MPEGB200_ESTATE, "ring: video stream %d not open", streams[i]);
            return nullptr;
        }
        const size_t bytes = s->luma_bytes + 2 * s->chroma_bytes;
        if (bytes > stride) stride = bytes;
    }
    stride = (stride + 255) / 256 * 256;
    mpegb200_ring* r = new mpegb200_ring();
    r->ctx = ctx;
    r->n = n;
    r->depth = depth;
    r->stride = stride;
    r->streams.assign(streams, streams + n);
    if (cudaMalloc(&r->dev, stride * (size_t)n * (size_t)depth) != cudaSuccess ||
        cudaEventCreateWithFlags(&r->pushed, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (r->dev) cudaFree(r->dev);
        delete r;
        fail(ctx, MPEGB200_ENOMEM, "ring of %d x %d pictures (%zu bytes each)", depth, n, stride);
        return nullptr;
    }
    cudaMemsetAsync(r->dev, 0, stride * (size_t)n * (size_t)depth, ctx->stream);
    return r;
} catch (...) {
    return nullptr;
}

void mpegb200_video_ring_free(mpegb200_ring* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    if (r->ctx->down_stream) cudaStreamSynchronize(r->ctx->down_stream);
    if (r->pushed) cudaEventDestroy(r->pushed);
    if (r->dev) cudaFree(r->dev);
    delete r;
}

int mpegb200_video_ring_push(mpegb200_ring* r, const uint8_t* bufs) {
    if (!r || !bufs) return MPEGB200_EINVAL;
    mpegb200_ctx* ctx = r->ctx;
    const int slot = (int)(r->head % (uint64_t)r->depth);
    uint8_t* base = r->dev + (size_t)slot * r->stride * (size_t)r->n;
    // a read-back of this slot that is still in flight must finish before the slot is overwritten
    CU(cudaSetDevice(ctx->device));
    if (ctx->down_stream) {
        CU(cudaEventRecord(ctx->ev_kernel, ctx->down_stream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_kernel, 0));
    }
    for (int i = 0; i < r->n;) {   // runs of streams that return a frame: one batched device-to-device copy each
        if (bufs[i] == 255) {
            i++;
            continue;
        }
        int j = i;
        while (j < r->n && bufs[j] != 255) j++;
        if (int rc = mpegb200_video_read_pictures_dev(ctx, j - i, r->streams.data() + i, bufs + i, base + (size_t)i * r->stride, r->stride))
            return rc;
        i = j;
    }
    CU(cudaEventRecord(r->pushed, ctx->stream));
    r->head++;
    return slot;
}

void* mpegb200_video_ring_slot_dev(mpegb200_ring* r, int slot, size_t* stride) {
    if (!r || slot < 0 || slot >= r->depth) return nullptr;
    if (stride) *stride = r->stride;
    return r->dev + (size_t)slot * r->stride * (size_t)r->n;
}

int mpegb200_video_ring_read_host(mpegb200_ring* r, int slot, uint8_t* dst, size_t dst_stride) {
    if (!r || !dst || slot < 0 || slot >= r->depth || dst_stride < 1) return MPEGB200_EINVAL;
    mpegb200_ctx* ctx = r->ctx;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamWaitEvent(ctx->down_stream, r->pushed, 0));
    const uint8_t* src = r->dev + (size_t)slot * r->stride * (size_t)r->n;
    const size_t width = dst_stride < r->stride ? dst_stride : r->stride;
    CU(cudaMemcpy2DAsync(dst, dst_stride, src, r->stride, width, (size_t)r->n, cudaMemcpyDeviceToHost, ctx->down_stream));
    return 0;
}

/* ---------------------------------------------------------------------------------------- audio */

static int audio_alloc(mpegb200_ctx* ctx) {
    if (ctx->d_audio) return 0;
    if (cudaMalloc(&ctx->d_audio, sizeof(AudioState) * (size_t)ctx->max_streams) != cudaSuccess) {
        cudaGetLastError();
        ctx->d_audio = nullptr;
        return fail(ctx, MPEGB200_ENOMEM, "audio state allocation failed");
    }
    CU(cudaMemsetAsync(ctx->d_audio, 0, sizeof(AudioState) * (size_t)ctx->max_streams, ctx->stream));
    return 0;
}

int mpegb200_audio_open(mpegb200_ctx* ctx, int stream) {
    if (!ctx || stream < 0 || stream >= ctx->max_streams) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (ctx->audio_open[stream]) return fail(ctx, MPEGB200_ESTATE, "audio stream %d already open", stream);
    CU(cudaSetDevice(ctx->device));
    if (int rc = audio_alloc(ctx)) return rc;
    // V = 0, vPos = 0 (a fresh Audio struct, audio.go:83-84), open = 1
    CU(cudaMemsetAsync(&ctx->d_audio[stream], 0, sizeof(AudioState), ctx->stream));
    static const int32_t one = 1;
    CU(cudaMemcpyAsync(&ctx->d_audio[stream].open, &one, sizeof(one), cudaMemcpyHostToDevice, ctx->stream));
    ctx->audio_open[stream] = 1;
    return 0;
}

int mpegb200_audio_close(mpegb200_ctx* ctx, int stream) {
    if (!ctx || stream < 0 || stream >= ctx->max_streams) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (!ctx->audio_open[stream]) return fail(ctx, MPEGB200_ESTATE, "audio stream %d not open", stream);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(&ctx->d_audio[stream], 0, sizeof(AudioState), ctx->stream));
    ctx->audio_open[stream] = 0;
    return 0;
}

static bool audio_format_ok(int format) {
    return format >= 0 && (format & ~(MPEGB200_AUDIO_FORMAT_MASK | MPEGB200_AUDIO_WINDOW_FMA)) == 0 && (format & MPEGB200_AUDIO_FORMAT_MASK) <= 3;
}
static size_t audio_out_bytes(int format) {
    return ((format & MPEGB200_AUDIO_FORMAT_MASK) == MPEGB200_AUDIO_S16 ? 2 : 4) * (size_t)2 * MPEGB200_SAMPLES_PER_FRAME;
}

int mpegb200_audio_synth_dev(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids, int frames_per_stream,
                             const int32_t* d_samples, int format, void* d_out) try {
    if (!ctx || n_streams < 0 || frames_per_stream < 0 || (n_streams && (!stream_ids || !d_samples || !d_out)))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (!audio_format_ok(format)) return fail(ctx, MPEGB200_EINVAL, "unknown audio format %d", format);
    if (n_streams == 0 || frames_per_stream == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(d_samples) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail(ctx, MPEGB200_EINVAL, "device arrays must be 16-byte aligned");
    {
        std::vector<uint8_t> seen(ctx->max_streams, 0);
        for (int i = 0; i < n_streams; i++) {
            const int s = stream_ids[i];
            if (s < 0 || s >= ctx->max_streams || !ctx->audio_open[s])
                return fail(ctx, MPEGB200_ESTATE, "audio stream %d not open", s);
            if (seen[s]++) return fail(ctx, MPEGB200_EINVAL, "audio stream %d listed twice", s);
        }
    }
    CU(cudaSetDevice(ctx->device));
    const size_t id_bytes = sizeof(int32_t) * (size_t)n_streams;
    if (int rc = ensure(ctx, ctx->s_ids, id_bytes)) return rc;
    CU(cudaMemcpyAsync(ctx->s_ids.p, stream_ids, id_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_audio_synth(ctx->d_audio, ctx->max_streams, (const int32_t*)ctx->s_ids.p, n_streams, frames_per_stream,
                          d_samples, format, d_out, ctx->d_window, ctx->stream));
    ctx->launches++;
    return 0;
} catch (...) {  // bad_alloc from the host-side bookkeeping must not cross the C boundary
    return fail(ctx, MPEGB200_ENOMEM, "host allocation failed");
}

int mpegb200_audio_synth(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids, int frames_per_stream,
                         const int32_t* samples, int format, void* out) {
    if (!ctx || n_streams < 0 || frames_per_stream < 0 || (n_streams && (!stream_ids || !samples || !out)))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (!audio_format_ok(format)) return fail(ctx, MPEGB200_EINVAL, "unknown audio format %d", format);
    if (n_streams == 0 || frames_per_stream == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    const size_t n_frames = (size_t)n_streams * frames_per_stream;
    const size_t in_bytes = n_frames * 2 * 36 * 32 * sizeof(int32_t), out_bytes = n_frames * audio_out_bytes(format);
    if (int rc = ensure(ctx, ctx->s_samples, in_bytes)) return rc;
    if (int rc = ensure(ctx, ctx->s_out, out_bytes)) return rc;
    CU(cudaMemcpyAsync(ctx->s_samples.p, samples, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = mpegb200_audio_synth_dev(ctx, n_streams, stream_ids, frames_per_stream, (const int32_t*)ctx->s_samples.p,
                                          format, ctx->s_out.p))
        return rc;
    CU(cudaMemcpyAsync(out, ctx->s_out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mpegb200_audio_synth_coded(mpegb200_ctx* ctx, int n_streams, const int32_t* stream_ids, int frames_per_stream,
                               const mpegb200_audio_frame_info* info, const uint16_t* codes, int format, void* out) {
    if (!ctx || n_streams < 0 || frames_per_stream < 0 || (n_streams && (!stream_ids || !info || !codes || !out)))
        return fail(ctx, MPEGB200_EINVAL, "null argument");
    if (!audio_format_ok(format)) return fail(ctx, MPEGB200_EINVAL, "unknown audio format %d", format);
    if (n_streams == 0 || frames_per_stream == 0) return 0;
    CU(cudaSetDevice(ctx->device));
    const size_t n_frames = (size_t)n_streams * frames_per_stream;
    const size_t out_bytes = n_frames * audio_out_bytes(format);
    if (int rc = ensure(ctx, ctx->s_ainfo, n_frames * sizeof(mpegb200_audio_frame_info))) return rc;
    if (int rc = ensure(ctx, ctx->s_acodes, n_frames * 2304 * sizeof(uint16_t))) return rc;
    if (int rc = ensure(ctx, ctx->s_samples, n_frames * 2304 * sizeof(int32_t))) return rc;
    if (int rc = ensure(ctx, ctx->s_out, out_bytes)) return rc;
    CU(cudaMemcpyAsync(ctx->s_ainfo.p, info, n_frames * sizeof(mpegb200_audio_frame_info), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->s_acodes.p, codes, n_frames * 2304 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_audio_requant((const mpegb200_audio_frame_info*)ctx->s_ainfo.p, (const uint16_t*)ctx->s_acodes.p, (int32_t*)ctx->s_samples.p,
                            n_frames, ctx->stream));
    ctx->launches++;
    if (int rc = mpegb200_audio_synth_dev(ctx, n_streams, stream_ids, frames_per_stream, (const int32_t*)ctx->s_samples.p, format, ctx->s_out.p))
        return rc;
    CU(cudaMemcpyAsync(out, ctx->s_out.p, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int mpegb200_audio_read_state(mpegb200_ctx* ctx, int stream, float* v, int* v_pos) {
    if (!ctx || stream < 0 || stream >= ctx->max_streams) return fail(ctx, MPEGB200_EINVAL, "stream id %d out of range", stream);
    if (!ctx->audio_open[stream]) return fail(ctx, MPEGB200_ESTATE, "audio stream %d not open", stream);
    CU(cudaSetDevice(ctx->device));
    AudioState* st = &ctx->d_audio[stream];
    int32_t vp = 0;
    if (v) CU(cudaMemcpyAsync(v, st->v, sizeof(float) * 2048, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&vp, &st->v_pos, sizeof(vp), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (v_pos) *v_pos = vp;
    return 0;
}

int mpegb200_audio_write_state(mpegb200_ctx* ctx, int stream, const float* v, int v_pos) {
    if (!ctx || stream < 0 || stream >= ctx->max_streams || !v) return fail(ctx, MPEGB200_EINVAL, "bad argument");
    if (!ctx->audio_open[stream]) return fail(ctx, MPEGB200_ESTATE, "audio stream %d not open", stream);
    if (v_pos < 0 || v_pos > 1023 || (v_pos & 63)) return fail(ctx, MPEGB200_EINVAL, "vPos must be a multiple of 64 in 0..960");
    CU(cudaSetDevice(ctx->device));
    AudioState* st = &ctx->d_audio[stream];
    const int32_t vp = v_pos;
    CU(cudaMemcpyAsync(st->v, v, sizeof(float) * 2048, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(&st->v_pos, &vp, sizeof(vp), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* ---------------------------------------------------------------------------------------- host memory */

void* mpegb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void mpegb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
