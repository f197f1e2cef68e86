// tma_rate.cu -- how fast does one SM's TMA unit serve small 3-D u8 boxes from an L2-resident tensor, as a function of
// the box shape?  (Is box mode of the fused kernel bound per box or per box ROW?)  One CTA per SM, one thread issues
// `n` boxes of W x R bytes round-robin into kSlots shared-memory slots, each with its own mbarrier.
// Build: nvcc -arch=sm_100a -O3 -o tma_rate tma_rate.cu ; run: ./tma_rate
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kSlots = 16;

__global__ void __launch_bounds__(32) rate_kernel(const __grid_constant__ CUtensorMap map, int n, int box_bytes, int slot_bytes,
                                                  int xmax, int ymax, int zmax, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[kSlots];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t rng = 12345u + blockIdx.x * 7919u;
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        const int s = i % kSlots;
        if (i >= kSlots) {  // wait for the previous use of the slot
            const uint32_t parity = ((i / kSlots) - 1) & 1;
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
            } while (!done);
        }
        rng = rng * 1664525u + 1013904223u;
        const int x = (int)((rng >> 8) % (uint32_t)xmax) & ~15, y = (int)((rng >> 4) % (uint32_t)ymax), z = (int)(rng % (uint32_t)zmax);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(sm + s * slot_bytes)), "l"(&map), "r"(smem_u32(&bar[s])), "r"(x), "r"(y), "r"(z) : "memory");
    }
    for (int s = 0; s < kSlots && s < n; s++) {  // drain
        const int uses = (n - s + kSlots - 1) / kSlots;
        const uint32_t parity = (uses - 1) & 1;
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(parity) : "memory");
        } while (!done);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // an L2-resident "slab": 16 frames of 1280 x 760
    const int W = 1280, H = 760, Z = 16;
    uint8_t* buf;
    cudaMalloc(&buf, (size_t)W * H * Z + 4096);
    cudaMemset(buf, 1, (size_t)W * H * Z + 4096);
    long long* d_cyc;
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    const int shapes[][2] = {{32, 1}, {32, 4}, {32, 9}, {32, 17}, {32, 20}, {32, 32}, {64, 17}, {128, 17}, {16, 17}, {48, 20}, {256, 8}, {256, 48}};
    printf("box WxR   bytes   cycles/box/SM   GB/s (all SMs, at 1.965 GHz)\n");
    for (auto& sh : shapes) {
        const int bw = sh[0], br = sh[1];
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)W + 32, (cuuint64_t)H - 1, (cuuint64_t)Z};
        const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)br, 1}, ones[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("%dx%d: encode failed\n", bw, br);
            continue;
        }
        const int box_bytes = bw * br, slot_bytes = (box_bytes + 127) / 128 * 128;
        const int n = 20000;
        cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * slot_bytes);
        for (int rep = 0; rep < 2; rep++)
            rate_kernel<<<sms, 32, kSlots * slot_bytes>>>(map, n, box_bytes, slot_bytes, W - bw, H - br - 1, Z, d_cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%dx%d: %s\n", bw, br, cudaGetErrorString(cudaGetLastError())); return 1; }
        long long* h = (long long*)malloc(sizeof(long long) * sms);
        cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; i++) mean += (double)h[i];
        mean /= sms;
        const double cpb = mean / n;
        printf("%3dx%-3d  %6d   %8.1f        %8.1f\n", bw, br, box_bytes, cpb, (double)box_bytes * sms / cpb * 1.965);
        free(h);
    }
    return 0;
}
