// Pure-write HBM bandwidth probes with incompressible data (diagnostic; build + run on the GPU box):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/write_bw tools/dbg/write_bw.cu && /tmp/write_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void stg_kernel(uint4* out, size_t n16, uint32_t seed, int streaming)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) {
        uint32_t h = hash32((uint32_t)i ^ seed);
        uint4 v = make_uint4(h, h * 3u + 1u, h ^ 0x9e3779b9u, h + (uint32_t)i);
        if (streaming) __stcs(out + i, v); else out[i] = v;
    }
}

// each CTA: fill `chunk` bytes of smem with hash data once, then bulk-store it to successive chunks (TMA 1-D stores),
// keeping `depth` stores in flight
__global__ void bulk_kernel(uint8_t* out, size_t nchunks, int chunk, int depth)
{
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = hash32(i * 2654435761u + blockIdx.x);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
        int inflight = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + c * (size_t)chunk), "r"(src), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= depth) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main()
{
    const size_t bytes = 807469056ull / 32768 * 32768;      // the conv output of BASELINE config C4
    uint8_t* buf[3];
    for (auto& b : buf) cudaMalloc(&b, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch(i % 3);
        cudaDeviceSynchronize();
        float best = 1e9f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            for (int i = 0; i < 12; i++) launch(i % 3);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 12; if (ms < best) best = ms;
        }
        printf("%-34s %.4f ms  %7.0f GB/s   (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int ctas : {148 * 4, 148 * 8, 148 * 16}) {
        char nm[64];
        snprintf(nm, sizeof nm, "STG.128 plain, %d CTAs x 256", ctas);
        time(nm, [&](int i) { stg_kernel<<<ctas, 256>>>((uint4*)buf[i], bytes / 16, 17u * i, 0); });
        snprintf(nm, sizeof nm, "STG.128 .cs,   %d CTAs x 256", ctas);
        time(nm, [&](int i) { stg_kernel<<<ctas, 256>>>((uint4*)buf[i], bytes / 16, 17u * i, 1); });
    }
    for (int chunk : {16384, 32768}) {
        for (int depth : {1, 2, 4}) {
            for (int per_sm : {1, 2}) {
                char nm[64];
                snprintf(nm, sizeof nm, "bulk store %d B, depth %d, %d CTA/SM", chunk, depth, per_sm);
                cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
                time(nm, [&](int i) { bulk_kernel<<<148 * per_sm, 128, chunk>>>(buf[i], bytes / chunk, chunk, depth); });
            }
        }
    }
    return 0;
}
