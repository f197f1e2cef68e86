// tma_probe.cu -- isolates which TMA feature a B200 accepts: (1) 2-D swizzled load through a
// __grid_constant__ map, (2) 3-D u8 window load, (3) the same with overlapping rows (x extent >
// row pitch), (4) the same with the map read from global memory.  Build: nvcc -arch=sm_100a.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

__global__ void probe2d(const __grid_constant__ CUtensorMap map, uint8_t* out, int row0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* tile = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(row0) : "memory");
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = tile[i];
}

__global__ void probe3d(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, uint8_t* out,
                        int x, int y, int z) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* tile = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        const void* m = use_global ? (const void*)gmap : (const void*)&pmap;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 17) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(m), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 32 * 17; i += blockDim.x) out[i] = tile[i];
}

__global__ void probe2d8(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int use_global, uint8_t* out,
                         int x, int y) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* tile = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        const void* m = use_global ? (const void*)gmap : (const void*)&pmap;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(32 * 17) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(m), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 32 * 17; i += blockDim.x) out[i] = tile[i];
}

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e = (x);                                                         \
        if (e != cudaSuccess) {                                                      \
            printf("  CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static int main_2d() {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const cuuint32_t ones[3] = {1, 1, 1};

    // ---- (1) 2-D swizzled int16 rows
    {
        const int rows = 100;
        std::vector<uint16_t> h(rows * 64);
        for (size_t i = 0; i < h.size(); i++) h[i] = (uint16_t)i;
        uint16_t* d;
        uint8_t* out;
        CK(cudaMalloc(&d, h.size() * 2));
        CK(cudaMalloc(&out, 4096));
        CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        CUtensorMap map;
        const cuuint64_t dims[2] = {64, rows}, strides[1] = {128};
        const cuuint32_t box[2] = {64, 32};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("(1) encode 2d swizzle128: %d\n", (int)r);
        probe2d<<<1, 128, 8192>>>(map, out, 80);  // rows 80..111: 20 valid + 12 out of bounds
        CK(cudaDeviceSynchronize());
        std::vector<uint16_t> o(2048);
        CK(cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int t = 0; t < 32; t++)
            for (int c = 0; c < 8; c++)
                for (int e = 0; e < 8; e++) {
                    uint16_t want = (80 + t < rows) ? (uint16_t)((80 + t) * 64 + c * 8 + e) : 0;
                    uint16_t got = o[t * 64 + ((c ^ (t & 7)) * 8) + e];
                    bad += got != want;
                }
        printf("(1) swizzled 2d load: %s (%d mismatches)\n", bad ? "MISMATCH" : "ok", bad);
    }

    return 0;
}

// usage: tma_probe 3d <x> <y> <z> <overlap> <global_map> <boxw> | tma_probe 2d
int main(int argc, char** argv) {
    if (argc < 2 || !strcmp(argv[1], "2d")) return main_2d();
    const int x = atoi(argv[2]), y = atoi(argv[3]), z = atoi(argv[4]), overlap = atoi(argv[5]), use_global = atoi(argv[6]);
    const int rank = argc > 7 ? atoi(argv[7]) : 3;
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    const cuuint32_t ones[3] = {1, 1, 1};
    const int W = 64, ROWS = 40, Z = 6;
    const size_t stride_z = (size_t)W * ROWS;
    std::vector<uint8_t> h(stride_z * Z + 256);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *out;
    CUtensorMap* gmap;
    CK(cudaMalloc(&d, h.size()));
    CK(cudaMalloc(&out, 1024));
    CK(cudaMalloc(&gmap, sizeof(CUtensorMap) * 2));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)W + (overlap ? 32 : 0), rank == 3 ? (cuuint64_t)ROWS : (cuuint64_t)ROWS * Z, Z};
    const cuuint64_t strides[2] = {W, stride_z};
    const cuuint32_t box[3] = {32, 17, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d x=%d y=%d z=%d overlap=%d global=%d: encode %d; ", rank, x, y, z, overlap, use_global, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 0; }
    CK(cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice));
    if (rank == 3) probe3d<<<1, 128, 8192>>>(map, gmap, use_global, out, x, y, z);
    else probe2d8<<<1, 128, 8192>>>(map, gmap, use_global, out, x, y + z * ROWS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> o(32 * 17);
    CK(cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int r2 = 0; r2 < 17; r2++)
        for (int c = 0; c < 32; c++) bad += o[r2 * 32 + c] != h[z * stride_z + (size_t)(y + r2) * W + x + c];  // linear addressing
    printf("%s (%d mismatches)\n", bad ? "MISMATCH" : "ok", bad);
    return 0;
}
