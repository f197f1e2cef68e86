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
// Layout.  The two channels of a stream sit side by side: a V slice is 64 float2 (channel 0, channel 1), so a window tap
// is one LDS.64 and -- with the packed FP32 pipe of sm_100 -- one FFMA2 (fused mode: the window coefficient is a scalar
// broadcast operand) or two FMUL + one FADD2 (default mode; ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// whatever --fmad says, so the products stay scalar mul.rn.f32, whose no-contraction guarantee PTX documents).
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

constexpr int kSlicePitch = 65;  // float2 per V slice in shared memory (64 + 1): 130 words = 2 mod 32, so the 32 (slot, channel)
                                 // threads of a warp write one element of their slices to 32 different banks
constexpr int kHist = 15;        // slices of history a window reaches back over (taps of age 1..15)
constexpr int kLin = kHist + 36; // a frame's slices lie linearly behind their history: tap address = slice - age * pitch
constexpr int kAudioThreads = 128;

struct AudioSmem {
    float2 v[kLin * kSlicePitch];         // 26,520 B: 8 CTAs per SM, so 1024 streams are resident in one wave
};

__device__ __forceinline__ float2 fma2_bcast(float d, float2 v, float2 acc) {   // (d * v.x + acc.x, d * v.y + acc.y), fused
    uint64_t r;
    asm("{\n\t.reg .b64 dd;\n\tmov.b64 dd, {%1, %1};\n\tfma.rn.f32x2 %0, dd, %2, %3;\n\t}"
        : "=l"(r)
        : "r"(__float_as_uint(d)), "l"(*reinterpret_cast<const uint64_t*>(&v)), "l"(*reinterpret_cast<const uint64_t*>(&acc)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}

// synthWindow (audio_noasm.go:8-38) for one time slot whose vPos/64 is the compile-time P.  `s` points at
// lane's element of the slot's own slice (both channels); D[t] = window[32 t + lane].
// With P fixed every tap's age (= which slice), half (= which 32 elements of it) and window index are constants:
// a tap is one LDS.64 with an immediate offset plus its arithmetic, accumulated in the reference's order
// (first loop: V positions 128 m + base, second loop: 64 + 128 m + (32 - base)).
template <int P, bool FMA>
__device__ __forceinline__ float2 window_taps(const float2* __restrict__ s, const float (&D)[32]) {
    float2 u = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int q = 2 * m + pass;                                  // V position / 64
            const int age = (q - P) & 15;
            const int half = pass == 0 ? (P & 1) : 1 - (P & 1);
            const int t = ((pass == 0 ? 512 : 544) - 32 * P + 64 * m) / 32;  // window index / 32 (audio_noasm.go:9,24)
            const float2 vv = s[half * 32 - age * kSlicePitch];
            if constexpr (FMA) {
                u = fma2_bcast(D[t], vv, u);                             // audio_amd64.s:123-131
            } else {
                const float2 p = make_float2(__fmul_rn(D[t], vv.x), __fmul_rn(D[t], vv.y));
                u = add2_rn(u, p);
            }
        }
    }
    return u;
}

// u / -1090519040.0 (audio.go:390), correctly rounded.  Fast path: q = u * y, r = fma(-q, c, u) (exact), q' = fma(r, y, q)
// with y = RN(1 / c) equals the IEEE quotient for every float32 u with 2^-102 <= |u| < inf and for +-0 (checked over all
// 2^32 bit patterns on the host, tools/check_fast_div.c); anything else takes the generic division.
__device__ __forceinline__ float scale_out(float u) {
    constexpr float c = -1090519040.0f;
    constexpr float y = -0x1.f81f82p-31f;   // RN(1 / c)
    const uint32_t mag = __float_as_uint(u) & 0x7fffffffu;
    if (((mag - 1u) < 0x0d7fffffu) | (mag >= 0x7f800000u)) return __fdiv_rn(u, c);   // 0 < |u| < 2^-100, inf, nan
    const float q = __fmul_rn(u, y);
    const float r = __fmaf_rn(-q, c, u);
    return __fmaf_rn(r, y, q);
}

// One CTA per stream; frames are processed in order, each in three barriers:
//   (a) 72 threads: one 32-point DCT each (time slot, channel) on samples read straight from global memory
//       -> V slice (audio.go:708-771 placement)
//   (b) 4 warps: the 36 time slots' windows, lane = output sample, both channels at once -> coalesced store
//   (c) the last 15 slices move to the front as the next frame's history
template <bool FMA>
__global__ void __launch_bounds__(kAudioThreads, 7) audio_synth_kernel(AudioState* __restrict__ states, int max_streams,
                                                                       const int32_t* __restrict__ stream_ids,
                                                                       int frames_per_stream,
                                                                       const int32_t* __restrict__ samples, int format,
                                                                       void* __restrict__ out,
                                                                       const float* __restrict__ window) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    AudioSmem& sm = *reinterpret_cast<AudioSmem*>(smem_raw);
    float* vf = reinterpret_cast<float*>(sm.v);      // word view: element e of channel ch of slice i at (i * pitch + e) * 2 + ch
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
        if (age < kHist) vf[((kHist - 1 - age) * kSlicePitch + e) * 2 + ch] = st.v[ch][q * 64 + e];
    }
    // (the slice of age 15 is overwritten by the first new slot in the reference's ring and never read)

    float D[32];  // the whole window table across the warp: D[t] = window[32 t + lane]
#pragma unroll
    for (int t = 0; t < 32; t++) D[t] = __ldg(&window[32 * t + lane]);

    const size_t frame_vals = 2 * MPEGB200_SAMPLES_PER_FRAME;
    for (int f = 0; f < frames_per_stream; f++) {
        const size_t fidx = (size_t)sidx * frames_per_stream + f;
        // (a) matrixing
        if (tid < 72) {
            const int step = tid >> 1, ch = tid & 1;
            // the time slot's 32 subband samples: 128 contiguous bytes, read straight from global memory
            const int4* sp = reinterpret_cast<const int4*>(samples + fidx * (2 * 36 * 32) + (ch * 36 + step) * 32);
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
            float* d = vf + (kHist + step) * kSlicePitch * 2 + ch;       // element i at d[2 i]
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float xe = e[k], xo = o[k];  // X[2k], X[2k+1]
                // X[m], m < 16:  d[48+m] = d[48-m] = -X[m];   m >= 16: d[m-16] = X[m], d[48-m] = -X[m]
                if (2 * k < 16) {
                    d[2 * (48 + 2 * k)] = -xe;
                    d[2 * (48 - 2 * k)] = -xe;
                    d[2 * (48 + 2 * k + 1)] = -xo;
                    d[2 * (48 - 2 * k - 1)] = -xo;
                } else {
                    d[2 * (2 * k - 16)] = xe;
                    d[2 * (48 - 2 * k)] = -xe;
                    d[2 * (2 * k + 1 - 16)] = xo;
                    d[2 * (48 - 2 * k - 1)] = -xo;
                }
            }
            d[2 * 16] = 0.0f;
        }
        __syncthreads();

        // (b) windows: warp w takes time slots w, w+4, ...
        {
            const int p_frame = p_init - f * 36;
            for (int step = warp; step < 36; step += 4) {
                const int p = (p_frame - step - 1) & 15;            // vPos / 64 for this slot (audio.go:380)
                const float2* s = &sm.v[(kHist + step) * kSlicePitch + lane];
                float2 u;
                switch (p) {
                    case 0: u = window_taps<0, FMA>(s, D); break;
                    case 1: u = window_taps<1, FMA>(s, D); break;
                    case 2: u = window_taps<2, FMA>(s, D); break;
                    case 3: u = window_taps<3, FMA>(s, D); break;
                    case 4: u = window_taps<4, FMA>(s, D); break;
                    case 5: u = window_taps<5, FMA>(s, D); break;
                    case 6: u = window_taps<6, FMA>(s, D); break;
                    case 7: u = window_taps<7, FMA>(s, D); break;
                    case 8: u = window_taps<8, FMA>(s, D); break;
                    case 9: u = window_taps<9, FMA>(s, D); break;
                    case 10: u = window_taps<10, FMA>(s, D); break;
                    case 11: u = window_taps<11, FMA>(s, D); break;
                    case 12: u = window_taps<12, FMA>(s, D); break;
                    case 13: u = window_taps<13, FMA>(s, D); break;
                    case 14: u = window_taps<14, FMA>(s, D); break;
                    default: u = window_taps<15, FMA>(s, D); break;
                }
                const float s0 = scale_out(u.x), s1 = scale_out(u.y);  // audio.go:390
                const int pos = step * 32 + lane;
                if (format == MPEGB200_AUDIO_F32N) {
                    reinterpret_cast<float2*>(out)[fidx * MPEGB200_SAMPLES_PER_FRAME + pos] = make_float2(s0, s1);
                } else if (format == MPEGB200_AUDIO_F32NLR) {
                    float* o = reinterpret_cast<float*>(out) + fidx * frame_vals;
                    o[pos] = s0;
                    o[MPEGB200_SAMPLES_PER_FRAME + pos] = s1;
                } else if (format == MPEGB200_AUDIO_S16) {  // audio.go:400-408
                    const int a = __float2int_rz(s0 < 0 ? __fmul_rn(s0, 32768.0f) : __fmul_rn(s0, 32767.0f));
                    const int b = __float2int_rz(s1 < 0 ? __fmul_rn(s1, 32768.0f) : __fmul_rn(s1, 32767.0f));
                    reinterpret_cast<uint32_t*>(out)[fidx * MPEGB200_SAMPLES_PER_FRAME + pos] =
                        ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16);
                } else {  // MPEGB200_AUDIO_F32, audio.go:409-417 (both constants are 2^31 as float32)
                    reinterpret_cast<float2*>(out)[fidx * MPEGB200_SAMPLES_PER_FRAME + pos] =
                        make_float2(__fmul_rn(s0, 2147483648.0f), __fmul_rn(s1, 2147483648.0f));
                }
            }
        }
        __syncthreads();

        // (c) the frame's last 15 slices (36..50) become the next frame's history (0..14)
        if (f + 1 < frames_per_stream) {
            constexpr int kMove = kHist * 64;             // float2 elements to move (the padding element stays)
            constexpr int kPer = (kMove + kAudioThreads - 1) / kAudioThreads;
            float2 keep[kPer];
#pragma unroll
            for (int i = 0; i < kPer; i++) {
                const int x = tid + i * kAudioThreads;
                if (x < kMove) keep[i] = sm.v[(36 + (x >> 6)) * kSlicePitch + (x & 63)];
            }
            __syncthreads();   // every read of 36..50 is done before the stores below and the next frame's DCTs overwrite 15..50
#pragma unroll
            for (int i = 0; i < kPer; i++) {
                const int x = tid + i * kAudioThreads;
                if (x < kMove) sm.v[(x >> 6) * kSlicePitch + (x & 63)] = keep[i];
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
        st.v[ch][q * 64 + e] = vf[((kLin - 1 - age) * kSlicePitch + e) * 2 + ch];
    }
    if (tid == 0) st.v_pos = p_final * 64;
}

cudaError_t launch_audio_synth(AudioState* d_states, int max_streams, const int32_t* d_stream_ids, int n_streams,
                               int frames_per_stream, const int32_t* d_samples, int format, void* d_out,
                               const float* d_window, cudaStream_t stream) {
    if (n_streams <= 0 || frames_per_stream <= 0) return cudaSuccess;
    const int fmt = format & MPEGB200_AUDIO_FORMAT_MASK;
    if (format & MPEGB200_AUDIO_WINDOW_FMA)
        audio_synth_kernel<true><<<n_streams, kAudioThreads, sizeof(AudioSmem), stream>>>(
            d_states, max_streams, d_stream_ids, frames_per_stream, d_samples, fmt, d_out, d_window);
    else
        audio_synth_kernel<false><<<n_streams, kAudioThreads, sizeof(AudioSmem), stream>>>(
            d_states, max_streams, d_stream_ids, frames_per_stream, d_samples, fmt, d_out, d_window);
    return cudaGetLastError();
}

cudaError_t configure_audio_kernel() {
    cudaError_t e = cudaFuncSetAttribute(audio_synth_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AudioSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(audio_synth_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AudioSmem));
}

}  // namespace mpegb200
