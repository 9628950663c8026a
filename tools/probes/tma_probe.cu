// development probe: which tensor-map / box configurations does UTMALDG accept on this part?
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int z, int bytes, uint8_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = (uint64_t*)(smem + 65536);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(smem)), "l"(&tm), "r"(s32(bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(smem)), "l"(&tm), "r"(s32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled)fp;
    const int W = 4096, H = 300, N = 2;
    int esz = 4, rank = 2, bw = 128, bh = 32, x = 0, y = 0;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    switch (variant) {
        case 0: break;                                         // f32 2D 128x32 at (0,0)
        case 1: bw = 136; bh = 34; x = -4; y = -1; break;      // f32 2D 136x34 at (-4,-1)
        case 2: rank = 3; break;                               // f32 3D 128x32x1
        case 3: esz = 2; dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;            // u16 2D 128x32
        case 4: esz = 2; dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; rank = 3; break;  // u16 3D 128x32x1
        case 5: esz = 2; dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; rank = 3; bw = 136; bh = 34; x = -4; y = -1; break;
        case 6: bw = 136; bh = 34; x = -4; y = -1; l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
        case 7: bw = 136; bh = 32; break;                      // only the width is odd
        case 8: bw = 128; bh = 34; break;                      // only the height is odd
        case 9: bw = 144; bh = 34; break;
        case 10: bw = 160; bh = 34; break;
        case 11: bw = 192; bh = 34; break;
        case 12: bw = 256; bh = 34; break;
    }
    std::vector<uint8_t> h((size_t)W * H * N * esz);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    const int bytes = bw * bh * esz;
    cudaMalloc(&o, bytes);
    cuuint64_t dims[3] = {W, H, N};
    cuuint64_t strides[2] = {(cuuint64_t)W * esz, (cuuint64_t)W * H * esz};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, dt, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: rank %d esz %d box %dx%d at (%d,%d): encode=%d ", variant, rank, esz, bw, bh, x, y, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    auto k = rank == 2 ? probe<2> : probe<3>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    k<<<1, 128, 65536 + 64>>>(tm, x, y, 1, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorName(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> got(bytes);
        cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        const int zf = rank == 3 ? 1 : 0;
        for (int r2 = 0; r2 < bh; ++r2)
            for (int c = 0; c < bw * esz; ++c) {
                int gy = y + r2, gxb = x * esz + c;
                uint8_t want = 0;
                if (gy >= 0 && gy < H && gxb >= 0 && gxb < W * esz) want = h[((size_t)zf * H + gy) * W * esz + gxb];
                if (got[(size_t)r2 * bw * esz + c] != want) ++bad;
            }
        printf("mismatch=%zu", bad);
    }
    printf("\n");
    return 0;
}
