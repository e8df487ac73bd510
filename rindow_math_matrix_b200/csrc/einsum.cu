// einsum.cu -- PhpMath::einsum (PhpMath.php:3321-3357) and einsum4p1 (:3360-3440): C = sum over the trailing index
// axes of A[...] * B[...], both operands addressed through per-axis strides (0 for an axis an operand lacks).
//
// One thread per element of C walks the summed axes in the reference's order (last axis fastest) and accumulates in
// double with separate multiply and add, exactly as PHP evaluates `$sumC += $A[..] * $B[..]`: float32 products are
// exact in double, so the result is the reference's bit for bit (one rounding at the store); float64 matches too
// (no FMA contraction).  Correctness first: contractions that are plain (batched) matrix products are far faster
// through rb200_sgemm / rb200_gemm_strided_batched, which LinearAlgebra::matmul already uses.
#include "common.cuh"

namespace {

constexpr int ES_THREADS = 128;
constexpr int ES_MAX_DEPTH = 16;

struct EinsumShape {
    int depth, ndimC;
    int64_t sizeC, sizeSum;
    int64_t size[ES_MAX_DEPTH], ldA[ES_MAX_DEPTH], ldB[ES_MAX_DEPTH];
};

template <typename T>
__global__ void __launch_bounds__(ES_THREADS)
einsum_kernel(EinsumShape s, const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C)
{
    for (int64_t ic = (int64_t)blockIdx.x * ES_THREADS + threadIdx.x; ic < s.sizeC; ic += (int64_t)gridDim.x * ES_THREADS) {
        int64_t baseA = 0, baseB = 0, r = ic;
        for (int axis = s.ndimC - 1; axis >= 0; axis--) {
            const int64_t idx = r % s.size[axis];
            r /= s.size[axis];
            baseA += idx * s.ldA[axis];
            baseB += idx * s.ldB[axis];
        }
        double sum = 0.0;
        for (int64_t t = 0; t < s.sizeSum; t++) {
            int64_t ia = baseA, ib = baseB, q = t;
            for (int axis = s.depth - 1; axis >= s.ndimC; axis--) {
                const int64_t idx = q % s.size[axis];
                q /= s.size[axis];
                ia += idx * s.ldA[axis];
                ib += idx * s.ldB[axis];
            }
            sum = __dadd_rn(sum, __dmul_rn((double)A[ia], (double)B[ib]));
        }
        C[ic] = (T)sum;
    }
}

template <typename T>
int einsum_launch(const EinsumShape& s, const void* A, const void* B, void* C)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(s.sizeC, ES_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 64;
    if (grid > cap) grid = cap;
    einsum_kernel<T><<<(unsigned)grid, ES_THREADS, 0, c.stream>>>(s, reinterpret_cast<const T*>(A), reinterpret_cast<const T*>(B),
                                                                 reinterpret_cast<T*>(C));
    RB_LAUNCH_CHECK("einsum_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_einsum(int dtype, int depth, int ndimC, const int64_t* sizeOfIndices, const void* A, int64_t offA,
                            const int64_t* ldA, const void* B, int64_t offB, const int64_t* ldB, void* C, int64_t offC)
{
    RB_REQUIRE_INIT();
    if (!A || !B || !C || !sizeOfIndices || !ldA || !ldB) RB_INVALID("NULL buffer");
    if (depth < 1 || depth > ES_MAX_DEPTH || ndimC < 0 || ndimC > depth || offA < 0 || offB < 0 || offC < 0) {
        RB_INVALID("einsum: bad depth / ndimC / offset");
    }
    EinsumShape s;
    s.depth = depth; s.ndimC = ndimC; s.sizeC = 1; s.sizeSum = 1;
    for (int h = 0; h < ES_MAX_DEPTH; h++) { s.size[h] = 1; s.ldA[h] = 0; s.ldB[h] = 0; }
    for (int h = 0; h < depth; h++) {
        if (sizeOfIndices[h] < 1 || ldA[h] < 0 || ldB[h] < 0) RB_INVALID("einsum: bad size / stride at axis %d", h);
        s.size[h] = sizeOfIndices[h]; s.ldA[h] = ldA[h]; s.ldB[h] = ldB[h];
        if (h < ndimC) s.sizeC *= sizeOfIndices[h]; else s.sizeSum *= sizeOfIndices[h];
    }
    const int es = rb_dtype_size(dtype);
    const char* a = reinterpret_cast<const char*>(A) + offA * es;
    const char* b = reinterpret_cast<const char*>(B) + offB * es;
    char* c = reinterpret_cast<char*>(C) + offC * es;
    switch (dtype) {
        case RB200_FLOAT32: return einsum_launch<float>(s, a, b, c);
        case RB200_FLOAT64: return einsum_launch<double>(s, a, b, c);
        default: RB_UNSUPPORTED("einsum: unsupported dtype %d", dtype);
    }
}
