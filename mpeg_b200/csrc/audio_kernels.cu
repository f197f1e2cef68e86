// audio_kernels.cu -- sm_100a kernel for the MP2 synthesis filterbank:
//   idct36 ("matrixing", audio.go:492-772) + synthWindow (audio_noasm.go:8-38) + the synthesis
//   loop with its output scaling (audio.go:377-422), for a rectangular batch of streams x frames.
//
// Numerics: every float32 operation is an explicitly rounded __fadd_rn/__fsub_rn/__fmul_rn/__fdiv_rn
// (never contracted into FMA), in the reference's operand and accumulation order, so the result
// is bit-identical to the reference's amd64 non-FMA path (golden hash 0xf1b76cdf8e6cdea5).
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
constexpr int kLin = kHist + 36; // a frame's slices lie linearly behind their history: tap address = slice - age * pitch
constexpr int kAudioThreads = 128;

struct AudioSmem {
    float v[2][kLin * kSlicePitch];       // 26,520 B: 8 CTAs per SM, so 1024 streams are resident in one wave
};

// synthWindow (audio_noasm.go:8-38) for one time slot whose vPos/64 is the compile-time P.  `s` points at
// lane's element of the slot's own slice (channel 0; channel 1 lies kLin slices further); D[t] = window[32 t + lane].
// With P fixed every tap's age (= which slice), half (= which 32 floats of it) and window index are constants:
// the 16 taps are two loads, two multiplies and two adds each, accumulated in the reference's order
// (first loop: V positions 128 m + base, second loop: 64 + 128 m + (32 - base)).
template <int P>
__device__ __forceinline__ void window_taps(const float* __restrict__ s, const float (&D)[32], float& u0, float& u1) {
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int q = 2 * m + pass;                                  // V position / 64
            const int age = (q - P) & 15;
            const int half = pass == 0 ? (P & 1) : 1 - (P & 1);
            const int t = ((pass == 0 ? 512 : 544) - 32 * P + 64 * m) / 32;  // window index / 32 (audio_noasm.go:9,24)
            const int off = half * 32 - age * kSlicePitch;
            u0 = __fadd_rn(u0, __fmul_rn(D[t], s[off]));
            u1 = __fadd_rn(u1, __fmul_rn(D[t], s[off + kLin * kSlicePitch]));
        }
    }
}

// One CTA per stream; frames are processed in order, each in three barriers:
//   (a) 72 threads: one 32-point DCT each (channel, time slot) on samples read straight from global memory
//       -> V slice (audio.go:708-771 placement)
//   (b) 4 warps: the 36 time slots' windows, lane = output sample, both channels -> coalesced store
//   (c) the last 15 slices move to the front as the next frame's history
__global__ void __launch_bounds__(kAudioThreads) audio_synth_kernel(AudioState* __restrict__ states, int max_streams,
                                                                    const int32_t* __restrict__ stream_ids,
                                                                    int frames_per_stream,
                                                                    const int32_t* __restrict__ samples, int format,
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

    const size_t frame_vals = 2 * MPEGB200_SAMPLES_PER_FRAME;
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

        // (b) windows: warp w takes time slots w, w+4, ...
        {
            float D[32];  // the whole window table across the warp: D[t] = window[32 t + lane]
#pragma unroll
            for (int t = 0; t < 32; t++) D[t] = __ldg(&window[32 * t + lane]);
            const int p_frame = p_init - f * 36;
            for (int step = warp; step < 36; step += 4) {
                const int p = (p_frame - step - 1) & 15;            // vPos / 64 for this slot (audio.go:380)
                const float* s = &sm.v[0][(kHist + step) * kSlicePitch + lane];
                float u0 = 0.0f, u1 = 0.0f;
                switch (p) {
                    case 0: window_taps<0>(s, D, u0, u1); break;
                    case 1: window_taps<1>(s, D, u0, u1); break;
                    case 2: window_taps<2>(s, D, u0, u1); break;
                    case 3: window_taps<3>(s, D, u0, u1); break;
                    case 4: window_taps<4>(s, D, u0, u1); break;
                    case 5: window_taps<5>(s, D, u0, u1); break;
                    case 6: window_taps<6>(s, D, u0, u1); break;
                    case 7: window_taps<7>(s, D, u0, u1); break;
                    case 8: window_taps<8>(s, D, u0, u1); break;
                    case 9: window_taps<9>(s, D, u0, u1); break;
                    case 10: window_taps<10>(s, D, u0, u1); break;
                    case 11: window_taps<11>(s, D, u0, u1); break;
                    case 12: window_taps<12>(s, D, u0, u1); break;
                    case 13: window_taps<13>(s, D, u0, u1); break;
                    case 14: window_taps<14>(s, D, u0, u1); break;
                    default: window_taps<15>(s, D, u0, u1); break;
                }
                const float s0 = __fdiv_rn(u0, -1090519040.0f), s1 = __fdiv_rn(u1, -1090519040.0f);  // audio.go:390
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
            float keep[kHist];
            const int ch = tid >> 6, e = tid & 63;
#pragma unroll
            for (int i = 0; i < kHist; i++) keep[i] = sm.v[ch][(36 + i) * kSlicePitch + e];
            __syncthreads();   // every read of 36..50 is done before the next frame's DCTs overwrite 15..50
#pragma unroll
            for (int i = 0; i < kHist; i++) sm.v[ch][i * kSlicePitch + e] = keep[i];
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

cudaError_t launch_audio_synth(AudioState* d_states, int max_streams, const int32_t* d_stream_ids, int n_streams,
                               int frames_per_stream, const int32_t* d_samples, int format, void* d_out,
                               const float* d_window, cudaStream_t stream) {
    if (n_streams <= 0 || frames_per_stream <= 0) return cudaSuccess;
    audio_synth_kernel<<<n_streams, kAudioThreads, sizeof(AudioSmem), stream>>>(
        d_states, max_streams, d_stream_ids, frames_per_stream, d_samples, format, d_out, d_window);
    return cudaGetLastError();
}

cudaError_t configure_audio_kernel() {
    return cudaFuncSetAttribute(audio_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AudioSmem));
}

}  // namespace mpegb200
