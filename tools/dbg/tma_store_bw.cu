// TMA tensor-map store bandwidth probes (diagnostic; GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tsb tools/dbg/tma_store_bw.cu -lcuda && /tmp/tsb
// The output is [M][64] floats (the fused conv's C tile = 128 rows x 256 B = 32 KB contiguous).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: two 2-D stores per tile, coords {0|32, row}; mode 1: one 2-D store {0, row}; mode 2: one 2-D store on the
// [2M][32] view {0, 2*row}; mode 3: one 3-D store {0, row, 0}
__global__ void tstore_kernel(const __grid_constant__ CUtensorMap map, int ntiles, int mode, int depth)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* al = sm + ((1024 - ((uint32_t)__cvta_generic_to_shared(sm) & 1023)) & 1023);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(al)[i] = hash32(i * 2654435761u + blockIdx.x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(al);
        int inflight = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int row = t * 128;
            if (mode == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(&map), "r"(0), "r"(row), "r"(src) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(&map), "r"(32), "r"(row), "r"(src + 16384) : "memory");
            } else if (mode == 1) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(&map), "r"(0), "r"(row), "r"(src) : "memory");
            } else if (mode == 2) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(&map), "r"(0), "r"(2 * row), "r"(src) : "memory");
            } else {
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" :: "l"(&map), "r"(0), "r"(row), "r"(0), "r"(src) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= depth) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main()
{
    const size_t M = 3154176, N = 64, bytes = M * N * 4;
    float* buf[3];
    for (auto& b : buf) cudaMalloc(&b, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(tstore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    auto mk = [&](int mode, float* base, bool sw) {
        CUtensorMap m;
        cuuint64_t dims[3]; cuuint64_t str[2]; cuuint32_t box[3]; cuuint32_t es[3] = {1, 1, 1};
        int rank = 2;
        if (mode == 0) { dims[0] = 64; dims[1] = M; str[0] = 256; box[0] = 32; box[1] = 128; }
        if (mode == 1) { dims[0] = 64; dims[1] = M; str[0] = 256; box[0] = 64; box[1] = 128; }
        if (mode == 2) { dims[0] = 32; dims[1] = 2 * M; str[0] = 128; box[0] = 32; box[1] = 256; }
        if (mode == 3) { rank = 3; dims[0] = 32; dims[1] = M; dims[2] = 2; str[0] = 256; str[1] = 128; box[0] = 32; box[1] = 128; box[2] = 2; }
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) printf("encode mode %d failed: %d\n", mode, (int)r);
        return m;
    };
    cudaFree(0);
    const char* names[4] = {"2 x box{32,128} sw128 (current)", "1 x box{64,128} no swizzle", "1 x box{32,256} on [2M][32] sw128", "1 x box{32,128,2} 3-D sw128"};
    for (int mode = 0; mode < 4; mode++) {
        CUtensorMap maps[3];
        for (int i = 0; i < 3; i++) maps[i] = mk(mode, buf[i], mode != 1);
        for (int depth : {1, 2, 4}) {
            for (int i = 0; i < 3; i++) tstore_kernel<<<148, 128, 40960>>>(maps[i], (int)(M / 128), mode, depth);
            cudaDeviceSynchronize();
            float best = 1e9f;
            for (int r = 0; r < 3; r++) {
                cudaEventRecord(e0);
                for (int i = 0; i < 12; i++) tstore_kernel<<<148, 128, 40960>>>(maps[i % 3], (int)(M / 128), mode, depth);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 12; if (ms < best) best = ms;
            }
            printf("%-36s depth %d  %.4f ms  %7.0f GB/s  (%s)\n", names[mode], depth, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
