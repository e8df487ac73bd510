// gatherb.cu -- batched gather / scatter / scatterAdd: PhpMath::gatherb (PhpMath.php:1209-1260), what
// LinearAlgebra::gatherb / scatterb / scatterbAdd (LinearAlgebra.php:2690-2920) call.
//
//   A[batches, m, numClass, k, len]   X[batches, n, k] (integer)   B[batches, m, n, k, len]
//   forward:  B[b,i,j,h,:] (=|+=) A[b,i,X[b,j,h],h,:]
//   reverse:  A[b,i,X[b,j,h],h,:] (=|+=) B[b,i,j,h,:]      j ascending, so the last j wins a copy and sums add in j order
//   an index >= numClass (or negative) skips the element, as the reference does.
//
// HBM-bound element moves, one thread per element of B.  The reverse modes stay deterministic and identical to the
// reference's order without atomics or a sort: for a copy only the last j holding an index writes; for an add the
// first j holding an index sums every later j of that index in ascending order on top of A.  Finding them is a walk
// over the n indices of (b, h) -- O(n) small cached reads per element, meant for the index counts of the layers that
// call it (argmax / top-k gathers, attention position lookups), not for vocabulary-sized scatters (rb200_gather has
// the sorted scatterAdd for those).
#include "common.cuh"

namespace {

constexpr int GB_THREADS = 256;

struct GbShape { int64_t batches, m, n, k, len, numClass, total; };

template <typename T, typename I, int MODE>    // MODE 0 gather, 1 gather-add, 2 scatter (last wins), 3 scatter-add
__global__ void __launch_bounds__(GB_THREADS)
gatherb_kernel(GbShape s, const I* __restrict__ X, T* A, T* B)
{
    for (int64_t t = (int64_t)blockIdx.x * GB_THREADS + threadIdx.x; t < s.total; t += (int64_t)gridDim.x * GB_THREADS) {
        const int64_t l = t % s.len;
        int64_t r = t / s.len;
        const int64_t h = r % s.k; r /= s.k;
        const int64_t j = r % s.n; r /= s.n;
        const int64_t i = r % s.m;
        const int64_t b = r / s.m;
        const I* xs = X + b * s.n * s.k + h;                 // X[b, :, h], stride k
        const int64_t index = (int64_t)xs[j * s.k];
        if (index < 0 || index >= s.numClass) continue;
        const int64_t ia = ((((b * s.m + i) * s.numClass + index) * s.k + h) * s.len) + l;
        if (MODE == 0) { B[t] = A[ia]; continue; }
        if (MODE == 1) { B[t] = B[t] + A[ia]; continue; }
        if (MODE == 2) {
            bool last = true;
            for (int64_t jj = j + 1; jj < s.n; jj++) { if ((int64_t)xs[jj * s.k] == index) { last = false; break; } }
            if (last) A[ia] = B[t];
            continue;
        }
        bool first = true;
        for (int64_t jj = 0; jj < j; jj++) { if ((int64_t)xs[jj * s.k] == index) { first = false; break; } }
        if (!first) continue;
        T acc = A[ia] + B[t];
        const int64_t jstride = s.k * s.len;
        for (int64_t jj = j + 1; jj < s.n; jj++) {
            if ((int64_t)xs[jj * s.k] == index) acc = acc + B[t + (jj - j) * jstride];
        }
        A[ia] = acc;
    }
}

template <typename T, typename I>
int gatherb_launch(int mode, const GbShape& s, const void* X, void* A, void* B)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(s.total, GB_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    const I* x = reinterpret_cast<const I*>(X);
    T* a = reinterpret_cast<T*>(A);
    T* b = reinterpret_cast<T*>(B);
    switch (mode) {
        case 0: gatherb_kernel<T, I, 0><<<(unsigned)grid, GB_THREADS, 0, c.stream>>>(s, x, a, b); break;
        case 1: gatherb_kernel<T, I, 1><<<(unsigned)grid, GB_THREADS, 0, c.stream>>>(s, x, a, b); break;
        case 2: gatherb_kernel<T, I, 2><<<(unsigned)grid, GB_THREADS, 0, c.stream>>>(s, x, a, b); break;
        default: gatherb_kernel<T, I, 3><<<(unsigned)grid, GB_THREADS, 0, c.stream>>>(s, x, a, b); break;
    }
    RB_LAUNCH_CHECK("gatherb_kernel");
    return RB200_OK;
}

template <typename I>
int gatherb_index(int dtype, int mode, const GbShape& s, const void* X, void* A, void* B)
{
    const bool add = (mode & 1) != 0;
    if (add) {
        switch (dtype) {
            case RB200_FLOAT32: return gatherb_launch<float, I>(mode, s, X, A, B);
            case RB200_FLOAT64: return gatherb_launch<double, I>(mode, s, X, A, B);
            case RB200_INT32:   return gatherb_launch<int32_t, I>(mode, s, X, A, B);
            case RB200_INT64:   return gatherb_launch<long long, I>(mode, s, X, A, B);
            default: RB_UNSUPPORTED("gatherb: add mode supports float32 / float64 / int32 / int64, not dtype %d", dtype);
        }
    }
    switch (rb_dtype_size(dtype)) {                          // copies move words: any dtype, bit-exact
        case 1: return gatherb_launch<uint8_t, I>(mode, s, X, A, B);
        case 2: return gatherb_launch<uint16_t, I>(mode, s, X, A, B);
        case 4: return gatherb_launch<uint32_t, I>(mode, s, X, A, B);
        case 8: return gatherb_launch<unsigned long long, I>(mode, s, X, A, B);
        default: RB_UNSUPPORTED("gatherb: unsupported dtype %d", dtype);
    }
}

}  // namespace

extern "C" int rb200_gatherb(int dtype, int index_dtype, int reverse, int add_mode, int64_t batches, int64_t m, int64_t n,
                             int64_t k, int64_t len, int64_t numClass, void* A, int64_t offA, const void* X, int64_t offX,
                             void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !X || !B) RB_INVALID("NULL buffer");
    if (batches < 1 || m < 1 || n < 1 || k < 1 || len < 1 || numClass < 1 || offA < 0 || offX < 0 || offB < 0) {
        RB_INVALID("gatherb: bad shape / offset");
    }
    const int es = rb_dtype_size(dtype), is = rb_dtype_size(index_dtype);
    if (es == 0) RB_UNSUPPORTED("gatherb: unsupported dtype %d", dtype);
    GbShape s;
    s.batches = batches; s.m = m; s.n = n; s.k = k; s.len = len; s.numClass = numClass;
    s.total = batches * m * n * k * len;
    const int mode = (reverse ? 2 : 0) | (add_mode ? 1 : 0);
    char* a = reinterpret_cast<char*>(A) + offA * es;
    char* b = reinterpret_cast<char*>(B) + offB * es;
    const char* x = reinterpret_cast<const char*>(X) + offX * is;
    switch (index_dtype) {
        case RB200_INT32:  return gatherb_index<int32_t>(dtype, mode, s, x, a, b);
        case RB200_UINT32: return gatherb_index<uint32_t>(dtype, mode, s, x, a, b);
        case RB200_INT64:  return gatherb_index<long long>(dtype, mode, s, x, a, b);
        case RB200_UINT64: return gatherb_index<long long>(dtype, mode, s, x, a, b);
        case RB200_INT16:  return gatherb_index<int16_t>(dtype, mode, s, x, a, b);
        case RB200_UINT16: return gatherb_index<uint16_t>(dtype, mode, s, x, a, b);
        case RB200_INT8:   return gatherb_index<int8_t>(dtype, mode, s, x, a, b);
        case RB200_UINT8:  return gatherb_index<uint8_t>(dtype, mode, s, x, a, b);
        default: RB_UNSUPPORTED("gatherb: index dtype %d is not an integer type", index_dtype);
    }
}
