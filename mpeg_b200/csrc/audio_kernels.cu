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
constexpr int kSlots = 52;       // slices kept per channel (>= 15 history + 36 of the frame)
constexpr int kAudioThreads = 128;

struct AudioSmem {
    float v[2][kSlots * kSlicePitch];       // 27,040 B: 8 CTAs per SM, so 1024 streams are resident in one wave
};

// One CTA per stream; frames are processed in order, each in three barriers:
//   (b) 72 threads: one 32-point DCT each (channel, time slot) on samples read straight from global memory
//       -> V slice (audio.go:708-771 placement)
//   (c) 4 warps: the 36 time slots' windows, lane = output sample, both channels -> coalesced store
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
    // age = (q - p) mod 16, p = vPos/64.  Step numbering: the first new time slot is step 16, so
    // history occupies steps 0..15 (step 15 = newest) and step s lives in slot s % kSlots.
    const int p_init = (st.v_pos >> 6) & 15;
    for (int i = tid; i < 2 * 16 * 64; i += kAudioThreads) {
        const int ch = i >> 10, q = (i >> 6) & 15, e = i & 63;
        const int age = (q - p_init) & 15;
        sm.v[ch][(15 - age) * kSlicePitch + e] = st.v[ch][q * 64 + e];
    }

    const size_t frame_vals = 2 * MPEGB200_SAMPLES_PER_FRAME;
    for (int f = 0; f < frames_per_stream; f++) {
        const size_t fidx = (size_t)sidx * frames_per_stream + f;
        // (b) matrixing
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
            const int gstep = 16 + f * 36 + step;
            float* d = &sm.v[ch][(gstep % kSlots) * kSlicePitch];
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

        // (c) windows: warp w takes time slots w, w+4, ...
        for (int step = warp; step < 36; step += 4) {
            const int gstep = 16 + f * 36 + step;
            const int t_total = f * 36 + step + 1;              // vPos has been decremented this many times
            const int p = (p_init - t_total) & 15;              // vPos / 64 for this step (audio.go:380)
            float u[2] = {0.0f, 0.0f};
            // first loop of synthWindow: positions 128m + base, then second loop: 64 + 128m + (32 - base)
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const int dbase = (pass == 0 ? 512 : 544) - 32 * p;
                const int half = pass == 0 ? (p & 1) : 1 - (p & 1);
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int q = 2 * m + pass;              // V position / 64
                    const int age = (q - p) & 15;
                    const int slot = (gstep - age) % kSlots;
                    const float dw = __ldg(&window[dbase + 64 * m + lane]);
                    const int idx = slot * kSlicePitch + half * 32 + lane;
                    u[0] = __fadd_rn(u[0], __fmul_rn(dw, sm.v[0][idx]));
                    u[1] = __fadd_rn(u[1], __fmul_rn(dw, sm.v[1][idx]));
                }
            }
            const float s0 = __fdiv_rn(u[0], -1090519040.0f), s1 = __fdiv_rn(u[1], -1090519040.0f);  // audio.go:390
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
        __syncthreads();
    }

    // write the state back in the reference's form: position q holds the slice of age (q - p_final) mod 16
    const int total = frames_per_stream * 36;
    const int p_final = (p_init - total) & 15;
    const int last = 16 + total - 1;
    for (int i = tid; i < 2 * 16 * 64; i += kAudioThreads) {
        const int ch = i >> 10, q = (i >> 6) & 15, e = i & 63;
        const int age = (q - p_final) & 15;
        st.v[ch][q * 64 + e] = sm.v[ch][((last - age) % kSlots) * kSlicePitch + e];
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
