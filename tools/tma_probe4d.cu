// tma_probe4d.cu -- can one TMA box fetch the Cb and the Cr window together?  Rank-4 u8 tensor over a
// buffer holding two planes `plane_stride` apart inside frames `frame_stride` apart, where neither stride
// divides the other.  Tries both dimension orders.  Build: nvcc -arch=sm_100a.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 9 * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(sm)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __syncthreads();
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < 576; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
    const int order = argc > 1 ? atoi(argv[1]) : 0;  // 0: {x,y,plane,z}, 1: {x,y,z,plane}
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const int W = 48, ROWS = 40, Z = 5;                 // plane = 48*40 = 1920 B; frame stride 4864 B (= 48*16*... not a multiple of 1920)
    const size_t plane = (size_t)W * ROWS, frame = 4864;
    std::vector<uint8_t> h(frame * Z + 4096);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 13 + (i >> 7));
    uint8_t *d, *out;
    cudaMalloc(&d, h.size()); cudaMalloc(&out, 1024);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    CUtensorMap map; const cuuint32_t ones[4] = {1, 1, 1, 1};
    CUresult r;
    if (order == 0) {
        const cuuint64_t dims[4] = {(cuuint64_t)W + 32, (cuuint64_t)ROWS * 2 + 8, 2, Z}, strides[3] = {(cuuint64_t)W, plane, frame};
        const cuuint32_t box[4] = {32, 9, 2, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[4] = {(cuuint64_t)W + 32, (cuuint64_t)ROWS * 2 + 8, Z, 2}, strides[3] = {(cuuint64_t)W, frame, plane};
        const cuuint32_t box[4] = {32, 9, 1, 2};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    printf("order %d: encode %d; ", order, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
    const int x = 16, y = 7, z = 3;
    if (order == 0) probe<<<1, 128, 4096>>>(map, out, x, y, 0, z); else probe<<<1, 128, 4096>>>(map, out, x, y, z, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> o(576);
    cudaMemcpy(o.data(), out, 576, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int p = 0; p < 2; p++) for (int rr = 0; rr < 9; rr++) for (int c = 0; c < 32; c++)
        bad += o[p * 288 + rr * 32 + c] != h[z * frame + p * plane + (size_t)(y + rr) * W + x + c];
    printf("%s (%d mismatches)\n", bad ? "MISMATCH" : "ok", bad);
    return 0;
}
