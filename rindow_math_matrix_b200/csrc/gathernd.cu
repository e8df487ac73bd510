// gathernd.cu -- gatherND / scatterND / scatterNDAdd: PhpMath::gathernd (PhpMath.php:1271-1325), what
// LinearAlgebra::gatherND / scatterND / scatterNDAdd (LinearAlgebra.php:2925-3090) call.
//
//   A[m, paramShape..., k]   X[m, n, indexDepth] (integer)   B[m, n, k]
//   forward:  B[i,j,:] (=|+=) A[i, X[i,j,0], ..., X[i,j,depth-1], :]      reverse: the other way, j ascending
//   a tuple with any index >= its paramShape entry (or negative) skips the row, as the reference does.
//   The reference reads X without its offset (PhpMath.php:1292: $X[$i*$n*$indexDepth + ...]); so does this.
//
// Same design as gatherb.cu: one thread per element of B; the reverse modes find the last / first j of a destination
// by walking the n index tuples of the batch row (deterministic, the reference's order, no atomics).
#include "common.cuh"

namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_DEPTH = 8;

struct GnShape {
    int64_t m, n, k, total, paramSize;
    int depth;
    int64_t dims[GN_MAX_DEPTH];
};

template <typename I>
__device__ __forceinline__ int64_t gn_offset(const GnShape& s, const I* __restrict__ x)
{
    int64_t off = 0;
    for (int h = 0; h < s.depth; h++) {
        const int64_t index = (int64_t)x[h];
        if (index < 0 || index >= s.dims[h]) return -1;
        off = off * s.dims[h] + index;
    }
    return off;
}

template <typename T, typename I, int MODE>    // MODE 0 gather, 1 gather-add, 2 scatter (last wins), 3 scatter-add
__global__ void __launch_bounds__(GN_THREADS)
gathernd_kernel(GnShape s, const I* __restrict__ X, T* A, T* B)
{
    for (int64_t t = (int64_t)blockIdx.x * GN_THREADS + threadIdx.x; t < s.total; t += (int64_t)gridDim.x * GN_THREADS) {
        const int64_t e = t % s.k;
        const int64_t r = t / s.k;
        const int64_t j = r % s.n, i = r / s.n;
        const I* xs = X + i * s.n * s.depth;
        const int64_t off = gn_offset<I>(s, xs + j * s.depth);
        if (off < 0) continue;
        const int64_t ia = (i * s.paramSize + off) * s.k + e;
        if (MODE == 0) { B[t] = A[ia]; continue; }
        if (MODE == 1) { B[t] = B[t] + A[ia]; continue; }
        if (MODE == 2) {
            bool last = true;
            for (int64_t jj = j + 1; jj < s.n; jj++) { if (gn_offset<I>(s, xs + jj * s.depth) == off) { last = false; break; } }
            if (last) A[ia] = B[t];
            continue;
        }
        bool first = true;
        for (int64_t jj = 0; jj < j; jj++) { if (gn_offset<I>(s, xs + jj * s.depth) == off) { first = false; break; } }
        if (!first) continue;
        T acc = A[ia] + B[t];
        for (int64_t jj = j + 1; jj < s.n; jj++) {
            if (gn_offset<I>(s, xs + jj * s.depth) == off) acc = acc + B[t + (jj - j) * s.k];
        }
        A[ia] = acc;
    }
}

template <typename T, typename I>
int gathernd_launch(int mode, const GnShape& s, const void* X, void* A, void* B)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(s.total, GN_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    const I* x = reinterpret_cast<const I*>(X);
    T* a = reinterpret_cast<T*>(A);
    T* b = reinterpret_cast<T*>(B);
    switch (mode) {
        case 0: gathernd_kernel<T, I, 0><<<(unsigned)grid, GN_THREADS, 0, c.stream>>>(s, x, a, b); break;
        case 1: gathernd_kernel<T, I, 1><<<(unsigned)grid, GN_THREADS, 0, c.stream>>>(s, x, a, b); break;
        case 2: gathernd_kernel<T, I, 2><<<(unsigned)grid, GN_THREADS, 0, c.stream>>>(s, x, a, b); break;
        default: gathernd_kernel<T, I, 3><<<(unsigned)grid, GN_THREADS, 0, c.stream>>>(s, x, a, b); break;
    }
    RB_LAUNCH_CHECK("gathernd_kernel");
    return RB200_OK;
}

template <typename I>
int gathernd_index(int dtype, int mode, const GnShape& s, const void* X, void* A, void* B)
{
    if (mode & 1) {
        switch (dtype) {
            case RB200_FLOAT32: return gathernd_launch<float, I>(mode, s, X, A, B);
            case RB200_FLOAT64: return gathernd_launch<double, I>(mode, s, X, A, B);
            case RB200_INT32:   return gathernd_launch<int32_t, I>(mode, s, X, A, B);
            case RB200_INT64:   return gathernd_launch<long long, I>(mode, s, X, A, B);
            default: RB_UNSUPPORTED("gathernd: add mode supports float32 / float64 / int32 / int64, not dtype %d", dtype);
        }
    }
    switch (rb_dtype_size(dtype)) {
        case 1: return gathernd_launch<uint8_t, I>(mode, s, X, A, B);
        case 2: return gathernd_launch<uint16_t, I>(mode, s, X, A, B);
        case 4: return gathernd_launch<uint32_t, I>(mode, s, X, A, B);
        case 8: return gathernd_launch<unsigned long long, I>(mode, s, X, A, B);
        default: RB_UNSUPPORTED("gathernd: unsupported dtype %d", dtype);
    }
}

}  // namespace

extern "C" int rb200_gathernd(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k,
                              int index_depth, const int64_t* param_shape, void* A, int64_t offA, const void* X,
                              void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !X || !B || !param_shape) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || index_depth < 1 || offA < 0 || offB < 0) RB_INVALID("gathernd: bad shape / offset");
    if (index_depth > GN_MAX_DEPTH) RB_UNSUPPORTED("gathernd: index depth %d exceeds %d", index_depth, GN_MAX_DEPTH);
    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("gathernd: unsupported dtype %d", dtype);
    GnShape s;
    s.m = m; s.n = n; s.k = k; s.depth = index_depth; s.total = m * n * k; s.paramSize = 1;
    for (int h = 0; h < GN_MAX_DEPTH; h++) s.dims[h] = 1;
    for (int h = 0; h < index_depth; h++) {
        if (param_shape[h] < 1) RB_INVALID("gathernd: paramShape[%d] must be positive", h);
        s.dims[h] = param_shape[h];
        s.paramSize *= param_shape[h];
    }
    const int mode = (reverse ? 2 : 0) | (add_mode ? 1 : 0);
    char* a = reinterpret_cast<char*>(A) + offA * es;
    char* b = reinterpret_cast<char*>(B) + offB * es;
    switch (index_dtype) {
        case RB200_INT32:  return gathernd_index<int32_t>(dtype, mode, s, X, a, b);
        case RB200_UINT32: return gathernd_index<uint32_t>(dtype, mode, s, X, a, b);
        case RB200_INT64:  return gathernd_index<long long>(dtype, mode, s, X, a, b);
        case RB200_UINT64: return gathernd_index<long long>(dtype, mode, s, X, a, b);
        default: RB_UNSUPPORTED("gathernd: index dtype %d is not a 32- or 64-bit integer", index_dtype);
    }
}
