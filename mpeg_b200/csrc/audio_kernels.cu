// audio_kernels.cu -- sm_100a kernel for the MP2 synthesis filterbank:
//   idct36 ("matrixing", audio.go:492-772) + synthWindow (audio_noasm.go:8-38) + the synthesis
//   loop with its output scaling (audio.go:377-422), for a rectangular batch of streams x frames.
//
// Numerics.  Two window modes, both bit-exact to a back-end of the reference:
//   * default: every multiply and every add of the window rounds on its own (mul.rn.f32 / add.rn), in the reference's
//     operand and accumulation order -- the amd64 SSE / pure-Go path, golden hash 0xf1b76cdf8e6cdea5 (mpeg_test.go:194);
//   * MPEGB200_AUDIO_WINDOW_FMA: each tap is one fused multiply-add, u = fma(d, v, u), like the reference's AVX2 / NEON
//     back-ends (audio_amd64.s:107-156 VFMADD231PS, audio_arm64.s:36-85 FMLA), golden hash 0x50f3ab75f5fb0fb5
//     (mpeg_test.go:195).
// The matrixing and the output scaling are the same in both (explicit .rn operations, never contracted).
//
// Where the time goes (profiles/r2_audio_summary.md): the first versions of this kernel read every window tap from shared
// memory -- 16 loads per output sample -- and ran into the LSU pipe (74 % of its wavefront peak) long before the FP32
// pipes.  Now a warp owns one channel and a run of consecutive time slots and keeps the 16-slice V ring in registers:
// consecutive slots share 15 of their 16 slices, so a slot costs two shared-memory loads (its own slice) and 16 register
// taps.  (Packed FP32 was tried for the two channels: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever
// --fmad says, so the default mode cannot use it; scalar mul.rn.f32 / add.rn.f32 carry PTX's no-contraction guarantee.)
#include "common.cuh"

namespace mpegb200 {

// C_n[i] = 1 / (2 cos((2i+1) pi / 2n)) with the reference's decimal literals (audio.go:498-555 ...)
__device__ __forceinline__ constexpr float lee_c(int n, int i) {
    constexpr float c32[16] = {0.500602998235f, 0.505470959898f, 0.515447309923f, 0.53104259109f,
                               0.553103896034f, 0.582934968206f, 0.622504123036f, 0.674808341455f,
                               0.744536271002f, 0.839349645416f, 0.972568237862f, 1.16943993343f,
                               1.48416461631f,  2.05778100995f,  3.40760841847f,  10.1900081235f};
    constexpr float c16[8] = {0.502419286188f, 0.52249861494f,  0.566944034816f, 0.64682178336f,
                              0.788154623451f, 1.06067768599f, 1.72244709824f,  5.10114861869f};
    constexpr float c8[4] = {0.509795579104f, 0.601344886935f, 0.899976223136f, 2.56291544774f};
    constexpr float c4[2] = {0.541196100146f, 1.30656296488f};
    return n == 32 ? c32[i] : n == 16 ? c16[i] : n == 8 ? c8[i] : n == 4 ? c4[i] : 0.707106781187f;
}

// B.G. Lee's recursive DCT, natural-order output, fully unrolled into registers.  Same data flow as
// the reference's 273 straight-line statements: e[i] = x[i] + x[n-1-i], o[i] = (x[i] - x[n-1-i]) * C,
// E = DCT(e), O = DCT(o), O[k] += O[k+1] ascending, X[2k] = E[k], X[2k+1] = O[k].
template <int N>
__device__ __forceinline__ void lee_dct(float (&x)[N]) {
    if constexpr (N > 1) {
        constexpr int H = N / 2;
        float e[H], o[H];
#pragma unroll
        for (int i = 0; i < H; i++) {
            e[i] = __fadd_rn(x[i], x[N - 1 - i]);
            o[i] = __fmul_rn(__fsub_rn(x[i], x[N - 1 - i]), lee_c(N, i));
        }
        lee_dct<H>(e);
        lee_dct<H>(o);
#pragma unroll
        for (int k = 0; k + 1 < H; k++) o[k] = __fadd_rn(o[k], o[k + 1]);
#pragma unroll
        for (int k = 0; k < H; k++) {
            x[2 * k] = e[k];
            x[2 * k + 1] = o[k];
        }
    }
}

constexpr int kSlicePitch = 65;  // floats per V slice in shared memory (64 + 1: conflict-free column writes)
constexpr int kHist = 15;        // slices of history a window reaches back over (taps of age 1..15)
constexpr int kLin = kHist + 36; // a frame's slices lie linearly behind their history: slice of age a = slot's slice - a
constexpr int kAudioThreads = 96;  // three warps: 72 of 96 threads hold a DCT in phase (a); 80 registers each with eight CTAs per SM
constexpr int kRunLen = 24;        // window phase: the 72 (channel, time slot) windows of a frame in three runs of 24

struct AudioSmem {
    float v[2][kLin * kSlicePitch];       // 26,520 B: eight CTAs per SM, so 1024 streams are resident in one wave
};

// u / -1090519040.0 (audio.go:390), correctly rounded.  Fast path: q = u * y, r = fma(-q, c, u) (exact), q' = fma(r, y, q)
// with y = RN(1 / c) equals the IEEE quotient for every float32 u with 2^-102 <= |u| < inf (checked over all 2^32 bit
// patterns on the host, tools/check_fast_div.c); anything else -- tiny values, inf, nan and, to keep the test at one
// compare, zero -- takes the generic division.  The test is made once per warp (vote): all lanes take the same path.
__device__ __forceinline__ float scale_out(float u) {
    constexpr float c = -1090519040.0f;
    constexpr float y = -0x1.f81f82p-31f;   // RN(1 / c)
    const uint32_t mag = __float_as_uint(u) & 0x7fffffffu;
    const bool odd = (mag - 0x0d800000u) >= (0x7f800000u - 0x0d800000u);   // |u| < 2^-100 (zero included), inf, nan
    if (__any_sync(0xffffffffu, odd)) return __fdiv_rn(u, c);
    const float q = __fmul_rn(u, y);
    const float r = __fmaf_rn(-q, c, u);
    return __fmaf_rn(r, y, q);
}

// Output formats (audio.go:386-418): element type, distance between a channel's consecutive samples, and the store.
template <int FORMAT> struct OutFmt { using T = float; static constexpr int kStride = 2; };
template <> struct OutFmt<MPEGB200_AUDIO_F32NLR> { using T = float; static constexpr int kStride = 1; };
template <> struct OutFmt<MPEGB200_AUDIO_S16> { using T = int16_t; static constexpr int kStride = 2; };

template <int FORMAT>
__device__ __forceinline__ typename OutFmt<FORMAT>::T* out_pointer(void* out, size_t fidx, int ch, int first, int lane) {
    using T = typename OutFmt<FORMAT>::T;
    T* base = reinterpret_cast<T*>(out) + fidx * (2 * MPEGB200_SAMPLES_PER_FRAME);
    if constexpr (FORMAT == MPEGB200_AUDIO_F32NLR) return base + ch * MPEGB200_SAMPLES_PER_FRAME + first * 32 + lane;
    return base + 2 * (first * 32 + lane) + ch;
}

template <int FORMAT>
__device__ __forceinline__ void emit_sample(typename OutFmt<FORMAT>::T* o, float s) {
    if constexpr (FORMAT == MPEGB200_AUDIO_S16) {  // audio.go:400-408
        *o = (int16_t)__float2int_rz(s < 0 ? __fmul_rn(s, 32768.0f) : __fmul_rn(s, 32767.0f));
    } else if constexpr (FORMAT == MPEGB200_AUDIO_F32) {  // audio.go:409-417 (both constants are 2^31 as float32)
        *o = __fmul_rn(s, 2147483648.0f);
    } else {
        *o = s;
    }
}

// synthWindow (audio_noasm.go:8-38) for one time slot whose vPos/64 is the compile-time P, one channel, lane = output
// sample.  The V ring lives in REGISTERS: R[q][h] = this lane's element of half h (32 floats) of the slice at ring
// position q (v[64 q ...]).  The slot's own slice (ring position P) comes in from shared memory -- two loads -- and
// every tap then is register arithmetic with compile-time indices: window index t and half follow from P like in the
// reference's two loops (first: V positions 128 m + 32 (P & 1), second: 64 + 128 m + 32 (1 - (P & 1))).
template <int P, bool FMA>
__device__ __forceinline__ float window_slot(const float* __restrict__ own, float (&R)[16][2], const float (&D)[32]) {
    R[P][0] = own[0];
    R[P][1] = own[32];
    float u = 0.0f;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            constexpr int kOdd = P & 1;
            const int q = 2 * m + pass;                                      // ring position = V position / 64
            const int half = pass == 0 ? kOdd : 1 - kOdd;
            const int t = ((pass == 0 ? 512 : 544) - 32 * P + 64 * m) / 32;  // window index / 32 (audio_noasm.go:9,24)
            if constexpr (FMA)
                u = __fmaf_rn(D[t], R[q][half], u);                          // audio_amd64.s:123-131
            else
                u = __fadd_rn(u, __fmul_rn(D[t], R[q][half]));
        }
    }
    return u;
}

// The 15 slices of history before a run's first slot (vPos/64 = P0) into the register ring: the slice written `age`
// slots earlier sits at ring position (P0 + age) & 15 -- a compile-time index once P0 is a template parameter.
template <int P0>
__device__ __forceinline__ void preload_ring(const float* __restrict__ own, float (&R)[16][2]) {
#pragma unroll
    for (int age = 1; age < 16; age++) {
        R[(P0 + age) & 15][0] = own[-age * kSlicePitch];
        R[(P0 + age) & 15][1] = own[-age * kSlicePitch + 32];
    }
}

// A run of `count` consecutive time slots of one channel, starting at slot `first` whose vPos/64 is p0: the history comes
// into registers once, then the slots go through an unrolled ring of the 16 values of vPos/64, entered at p0 (Duff's
// device; vPos decreases by 64 per slot, audio.go:380).
template <bool FMA, int FORMAT>
__device__ __forceinline__ void window_run(const float* __restrict__ vch, const float (&D)[32], int ch, int first, int count, int p0,
                                           size_t fidx, void* __restrict__ out, int lane) {
    const float* own = vch + (kHist + first) * kSlicePitch + lane;
    float R[16][2];
    switch (p0) {
#define MPEGB200_PRE(PV) case PV: preload_ring<PV>(own, R); break;
        MPEGB200_PRE(0) MPEGB200_PRE(1) MPEGB200_PRE(2) MPEGB200_PRE(3) MPEGB200_PRE(4) MPEGB200_PRE(5) MPEGB200_PRE(6) MPEGB200_PRE(7)
        MPEGB200_PRE(8) MPEGB200_PRE(9) MPEGB200_PRE(10) MPEGB200_PRE(11) MPEGB200_PRE(12) MPEGB200_PRE(13) MPEGB200_PRE(14)
        default: preload_ring<15>(own, R); break;
#undef MPEGB200_PRE
    }
    typename OutFmt<FORMAT>::T* o = out_pointer<FORMAT>(out, fidx, ch, first, lane);
    int left = count;
#define MPEGB200_SLOT(PV)                                                           \
    case PV: {                                                                      \
        emit_sample<FORMAT>(o, scale_out(window_slot<PV, FMA>(own, R, D)));         \
        if (--left == 0) break;                                                     \
        own += kSlicePitch;                                                         \
        o += 32 * OutFmt<FORMAT>::kStride;                                          \
    }
    switch (p0) {
        do {
            MPEGB200_SLOT(15) MPEGB200_SLOT(14) MPEGB200_SLOT(13) MPEGB200_SLOT(12) MPEGB200_SLOT(11) MPEGB200_SLOT(10)
            MPEGB200_SLOT(9) MPEGB200_SLOT(8) MPEGB200_SLOT(7) MPEGB200_SLOT(6) MPEGB200_SLOT(5) MPEGB200_SLOT(4)
            MPEGB200_SLOT(3) MPEGB200_SLOT(2) MPEGB200_SLOT(1) MPEGB200_SLOT(0)
        } while (true);
    }
#undef MPEGB200_SLOT
}

// One CTA per stream; frames are processed in order, each in three barriers:
//   (a) 72 threads: one 32-point DCT each (channel, time slot) on samples read straight from global memory
//       -> V slice in shared memory (audio.go:708-771 placement)
//   (b) 3 warps: the frame's 72 (channel, time slot) windows in runs of consecutive slots of one channel (window_run):
//       two shared-memory loads per slot instead of 16
//   (c) the last 15 slices move to the front as the next frame's history
template <bool FMA, int FORMAT>
__global__ void __launch_bounds__(kAudioThreads, 8) audio_synth_kernel(AudioState* __restrict__ states, int max_streams,
                                                                       const int32_t* __restrict__ stream_ids,
                                                                       int frames_per_stream,
                                                                       const int32_t* __restrict__ samples,
                                                                       void* __restrict__ out,
                                                                       const float* __restrict__ window) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    AudioSmem& sm = *reinterpret_cast<AudioSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sidx = blockIdx.x;
    const int stream = stream_ids[sidx];
    if (stream < 0 || stream >= max_streams) return;
    AudioState& st = states[stream];
    if (!st.open) return;

    // History: the slice at V position q (64 floats at v[64q]) was written `age` steps ago with
    // age = (q - p) mod 16, p = vPos/64 (age 0 = newest).  A new time slot reaches back 15 slices, so ages
    // 0..14 go to linear slices 14..0 and the frame's 36 slices follow at 15..50.
    const int p_init = (st.v_pos >> 6) & 15;
    for (int i = tid; i < 2 * 16 * 64; i += kAudioThreads) {
        const int ch = i >> 10, q = (i >> 6) & 15, e = i & 63;
        const int age = (q - p_init) & 15;
        if (age < kHist) sm.v[ch][(kHist - 1 - age) * kSlicePitch + e] = st.v[ch][q * 64 + e];
    }
    // (the slice of age 15 is overwritten by the first new slot in the reference's ring and never read)

    for (int f = 0; f < frames_per_stream; f++) {
        const size_t fidx = (size_t)sidx * frames_per_stream + f;
        // (a) matrixing
        if (tid < 72) {
            const int ch = tid / 36, step = tid - ch * 36;
            // the time slot's 32 subband samples: 128 contiguous bytes, read straight from global memory
            const int4* sp = reinterpret_cast<const int4*>(samples + fidx * (2 * 36 * 32) + tid * 32);
            int s[32];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int4 w = __ldg(sp + i);
                s[4 * i] = w.x; s[4 * i + 1] = w.y; s[4 * i + 2] = w.z; s[4 * i + 3] = w.w;
            }
            if (f + 1 < frames_per_stream)  // the next frame's samples on their way to L2 while this one is computed
                asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + (2 * 36 * 32) / 4));
            float e[16], o[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int a = s[i], b = s[31 - i];
                e[i] = (float)(a + b);                                   // integer add, then convert: audio.go:497
                o[i] = __fmul_rn((float)(a - b), lee_c(32, i));          // audio.go:498
            }
            lee_dct<16>(e);
            lee_dct<16>(o);
#pragma unroll
            for (int k = 0; k < 15; k++) o[k] = __fadd_rn(o[k], o[k + 1]);  // audio.go:692-706
            // X[2k] = e[k], X[2k+1] = o[k]; placement audio.go:708-771
            float* d = &sm.v[ch][(kHist + step) * kSlicePitch];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float xe = e[k], xo = o[k];  // X[2k], X[2k+1]
                // X[m], m < 16:  d[48+m] = d[48-m] = -X[m];   m >= 16: d[m-16] = X[m], d[48-m] = -X[m]
                if (2 * k < 16) {
                    d[48 + 2 * k] = -xe;
                    d[48 - 2 * k] = -xe;
                    d[48 + 2 * k + 1] = -xo;
                    d[48 - 2 * k - 1] = -xo;
                } else {
                    d[2 * k - 16] = xe;
                    d[48 - 2 * k] = -xe;
                    d[2 * k + 1 - 16] = xo;
                    d[48 - 2 * k - 1] = -xo;
                }
            }
            d[16] = 0.0f;
        }
        __syncthreads();

        // (b) windows: warp w takes the windows 24 w .. 24 w + 23 of the frame's 72 (channel-major), i.e. one or two runs of
        // consecutive time slots of one channel
        {
            float D[32];  // the whole window table across the warp: D[t] = window[32 t + lane] (4 KiB, L1 resident; loaded per
                          // frame so that it does not occupy 32 registers during the matrixing)
#pragma unroll
            for (int t = 0; t < 32; t++) D[t] = __ldg(&window[32 * t + lane]);
            int wdx = warp * kRunLen;
            const int wend = wdx + kRunLen;
            while (wdx < wend) {
                const int ch = wdx >= 36 ? 1 : 0, first = wdx - 36 * ch;
                const int count = min(wend - wdx, 36 - first);
                window_run<FMA, FORMAT>(sm.v[ch], D, ch, first, count, (p_init - f * 36 - first - 1) & 15, fidx, out, lane);
                wdx += count;
            }
        }
        __syncthreads();

        // (c) the frame's last 15 slices (36..50) become the next frame's history (0..14)
        if (f + 1 < frames_per_stream) {
            constexpr int kMove = 2 * kHist * 64;             // floats to move (the padding elements stay)
            constexpr int kPer = (kMove + kAudioThreads - 1) / kAudioThreads;
            float keep[kPer];
#pragma unroll
            for (int i = 0; i < kPer; i++) {
                const int x = tid + i * kAudioThreads;        // [channel][slice][element]
                if (x < kMove) keep[i] = sm.v[x / (kHist * 64)][(36 + (x % (kHist * 64)) / 64) * kSlicePitch + (x & 63)];
            }
            __syncthreads();   // every read of 36..50 is done before the stores below and the next frame's DCTs overwrite 15..50
#pragma unroll
            for (int i = 0; i < kPer; i++) {
                const int x = tid + i * kAudioThreads;
                if (x < kMove) sm.v[x / (kHist * 64)][((x % (kHist * 64)) / 64) * kSlicePitch + (x & 63)] = keep[i];
            }
            // the barrier after (a) orders these stores before the windows read them
        }
    }

    // write the state back in the reference's form: position q holds the slice of age (q - p_final) mod 16
    const int total = frames_per_stream * 36;
    const int p_final = (p_init - total) & 15;
    for (int i = tid; i < 2 * 16 * 64; i += kAudioThreads) {
        const int ch = i >> 10, q = (i >> 6) & 15, e = i & 63;
        const int age = (q - p_final) & 15;
        st.v[ch][q * 64 + e] = sm.v[ch][(kLin - 1 - age) * kSlicePitch + e];
    }
    if (tid == 0) st.v_pos = p_final * 64;
}

template <bool FMA, int FORMAT>
static cudaError_t launch_one(AudioState* d_states, int max_streams, const int32_t* d_stream_ids, int n_streams, int frames_per_stream,
                              const int32_t* d_samples, void* d_out, const float* d_window, cudaStream_t stream) {
    audio_synth_kernel<FMA, FORMAT><<<n_streams, kAudioThreads, sizeof(AudioSmem), stream>>>(
        d_states, max_streams, d_stream_ids, frames_per_stream, d_samples, d_out, d_window);
    return cudaGetLastError();
}

cudaError_t launch_audio_synth(AudioState* d_states, int max_streams, const int32_t* d_stream_ids, int n_streams,
                               int frames_per_stream, const int32_t* d_samples, int format, void* d_out,
                               const float* d_window, cudaStream_t stream) {
    if (n_streams <= 0 || frames_per_stream <= 0) return cudaSuccess;
    const bool fma = (format & MPEGB200_AUDIO_WINDOW_FMA) != 0;
#define MPEGB200_GO(F, FMT) return launch_one<F, FMT>(d_states, max_streams, d_stream_ids, n_streams, frames_per_stream, d_samples, d_out, d_window, stream)
    switch (format & MPEGB200_AUDIO_FORMAT_MASK) {
        case MPEGB200_AUDIO_F32N: if (fma) MPEGB200_GO(true, MPEGB200_AUDIO_F32N); else MPEGB200_GO(false, MPEGB200_AUDIO_F32N);
        case MPEGB200_AUDIO_F32NLR: if (fma) MPEGB200_GO(true, MPEGB200_AUDIO_F32NLR); else MPEGB200_GO(false, MPEGB200_AUDIO_F32NLR);
        case MPEGB200_AUDIO_F32: if (fma) MPEGB200_GO(true, MPEGB200_AUDIO_F32); else MPEGB200_GO(false, MPEGB200_AUDIO_F32);
        case MPEGB200_AUDIO_S16: if (fma) MPEGB200_GO(true, MPEGB200_AUDIO_S16); else MPEGB200_GO(false, MPEGB200_AUDIO_S16);
        default: return cudaErrorInvalidValue;
    }
#undef MPEGB200_GO
}

cudaError_t configure_audio_kernel() {
    cudaError_t e = cudaSuccess;
#define MPEGB200_CFG(F, FMT)                                                                                                         \
    if (e == cudaSuccess)                                                                                                            \
        e = cudaFuncSetAttribute(audio_synth_kernel<F, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AudioSmem))
    MPEGB200_CFG(false, MPEGB200_AUDIO_F32N); MPEGB200_CFG(true, MPEGB200_AUDIO_F32N);
    MPEGB200_CFG(false, MPEGB200_AUDIO_F32NLR); MPEGB200_CFG(true, MPEGB200_AUDIO_F32NLR);
    MPEGB200_CFG(false, MPEGB200_AUDIO_F32); MPEGB200_CFG(true, MPEGB200_AUDIO_F32);
    MPEGB200_CFG(false, MPEGB200_AUDIO_S16); MPEGB200_CFG(true, MPEGB200_AUDIO_S16);
#undef MPEGB200_CFG
    return e;
}

}  // namespace mpegb200
