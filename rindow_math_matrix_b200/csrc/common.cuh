// common.cuh -- shared state and helpers of librindow_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rindow_b200.h"

namespace rb200 {

struct Context {
    bool        initialised = false;
    int         device = -1;
    int         sm_count = 0;
    int         cc_major = 0, cc_minor = 0;
    size_t      total_mem = 0;
    cudaStream_t own_stream = nullptr;   // created by rb200_init
    cudaStream_t stream = nullptr;       // the stream kernels are launched on
    cudaEvent_t  switch_event = nullptr; // orders a newly adopted stream after the previous one (rb200_set_stream)
    int         default_gemm_mode = RB200_GEMM_TF32X3;
    int64_t     launches = 0;
    const char* last_gemm_path = "none";
    // scratch for two-pass reductions / split-K (grown on demand, stream ordered)
    void*       workspace = nullptr;
    size_t      workspace_bytes = 0;
    // self-resetting arrival tickets of the split reductions (zero between launches)
    int*        counters = nullptr;
    size_t      counters_len = 0;
    // pinned 64-byte slot the scalar-result calls (sum / imax / imin / onehot's error flag) copy into
    void*       pinned_scalar = nullptr;
    // peer copies of C the next tensor-core GEMM also writes from its epilogue (rb200_sgemm_allgather)
    float*      gemm_peers[RB200_MAX_PEERS] = {};
    int         n_gemm_peers = 0;
    bool        gemm_peers_written = false;
    // multicast alias of C (+ offC) the next tensor-core GEMM stores its final values to (rb200_sgemm_allgather_mc)
    float*      gemm_mc = nullptr;
    // arrival / exit counters of the CTA-pair GEMM's meeting points (gemm_tcgen05.cu, TcParams::sync_cnt)
    unsigned int* gemm_sync = nullptr;
};

Context& ctx();
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int  ensure_workspace(size_t bytes);
int  ensure_counters(int64_t n, int** out);

inline void count_launch(int n = 1) { ctx().launches += n; }

}  // namespace rb200

#define RB_CUDA(call)                                                            \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return rb200::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define RB_REQUIRE_INIT()                                                        \
    do {                                                                         \
        if (!rb200::ctx().initialised) {                                         \
            rb200::set_error("rb200_init() has not been called");                \
            return RB200_E_NOTINIT;                                              \
        }                                                                        \
    } while (0)

#define RB_INVALID(...)                                                          \
    do {                                                                         \
        rb200::set_error(__VA_ARGS__);                                           \
        return RB200_E_INVALID;                                                  \
    } while (0)

#define RB_UNSUPPORTED(...)                                                      \
    do {                                                                         \
        rb200::set_error(__VA_ARGS__);                                           \
        return RB200_E_UNSUPPORTED;                                              \
    } while (0)

// post-launch check: catches bad launch configurations immediately (async faults surface
// at the next sync, as with any CUDA library).
#define RB_LAUNCH_CHECK(name)                                                    \
    do {                                                                         \
        cudaError_t _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess) return rb200::cuda_fail(_e, name, __FILE__, __LINE__); \
        rb200::count_launch();                                                   \
    } while (0)

static inline int rb_dtype_size(int dtype)
{
    switch (dtype) {
        case RB200_BOOL: case RB200_INT8: case RB200_UINT8: return 1;
        case RB200_INT16: case RB200_UINT16: return 2;
        case RB200_INT32: case RB200_UINT32: case RB200_FLOAT32: return 4;
        case RB200_INT64: case RB200_UINT64: case RB200_FLOAT64: case RB200_COMPLEX64: return 8;
        case RB200_COMPLEX128: return 16;
        default: return 0;
    }
}

static inline int64_t rb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// 128-bit streaming load / store (read-once data: keep it out of L1).
// ld_stream_f4: read-only inputs (.nc).  ld_inplace_f4: data the same thread overwrites.
__device__ __forceinline__ float4 ld_inplace_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
