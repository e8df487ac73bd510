// masking.cu -- masking: A[m,n,k,len] gets `fill` written (mode 0) or added (mode 1) wherever the boolean mask
// X[m,k] is false; X broadcasts over n and len (attention / sequence masks).  HBM-bound: only the masked-out
// elements are touched (mode 0 never reads A).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:3229-3270):
//   for i<m, j<n, h<k: if (!X[i*k+h]) for l<len: a = ((i*n+j)*k+h)*len+l;
//        mode 0: A[a] = fill        mode 1: A[a] = A[a] + fill   (bool A: A[a] || fill)
// `fill` is a PHP float stored into a buffer of A's dtype: truncated toward zero for integer buffers, (fill != 0) for bool.
#include "common.cuh"

namespace {

constexpr int MK_THREADS = 256;

template <typename T> __device__ __forceinline__ T mk_add(T a, T f) { return a + f; }

// one thread per element of A; (i, h) recovered with 64-bit divisions only when the mask is not a plain broadcast
template <typename T, bool ADD, bool IS_BOOL>
__global__ void __launch_bounds__(MK_THREADS)
masking_kernel(int64_t total, int64_t n, int64_t k, int64_t len, T fill, const uint8_t* __restrict__ X, T* __restrict__ A)
{
    const int64_t kl = k * len, nkl = n * kl;
    for (int64_t t = (int64_t)blockIdx.x * MK_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * MK_THREADS) {
        const int64_t i = t / nkl;
        const int64_t h = ((t - i * nkl) % kl) / len;
        if (!X[i * k + h]) {
            if (!ADD) A[t] = fill;
            else if (IS_BOOL) A[t] = (T)((A[t] != 0) || (fill != 0));
            else A[t] = mk_add<T>(A[t], fill);
        }
    }
}

template <typename T, bool IS_BOOL>
int masking_launch(int64_t m, int64_t n, int64_t k, int64_t len, double fill, int mode, const void* X, void* A)
{
    rb200::Context& c = rb200::ctx();
    const int64_t total = m * n * k * len;
    int64_t grid = rb_ceil_div(total, MK_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const T f = IS_BOOL ? (T)(fill != 0.0) : (T)fill;
    const uint8_t* x = reinterpret_cast<const uint8_t*>(X);
    T* a = reinterpret_cast<T*>(A);
    if (mode == 0) masking_kernel<T, false, IS_BOOL><<<(unsigned)grid, MK_THREADS, 0, c.stream>>>(total, n, k, len, f, x, a);
    else           masking_kernel<T, true, IS_BOOL><<<(unsigned)grid, MK_THREADS, 0, c.stream>>>(total, n, k, len, f, x, a);
    RB_LAUNCH_CHECK("masking_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_masking(int dtype, int64_t m, int64_t n, int64_t k, int64_t len, double fill, int mode,
                             const void* X, int64_t offX, void* A, int64_t offA)
{
    RB_REQUIRE_INIT();
    if (!X || !A) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || len < 1 || offX < 0 || offA < 0) RB_INVALID("masking: bad shape / offset");
    if (mode != 0 && mode != 1) RB_INVALID("masking: mode must be 0 (set) or 1 (add)");
    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("masking: unsupported dtype %d", dtype);
    const char* x = reinterpret_cast<const char*>(X) + offX;          // bool buffer: one byte per element
    char* a = reinterpret_cast<char*>(A) + offA * es;
    switch (dtype) {
        case RB200_FLOAT32: return masking_launch<float, false>(m, n, k, len, fill, mode, x, a);
        case RB200_FLOAT64: return masking_launch<double, false>(m, n, k, len, fill, mode, x, a);
        case RB200_INT32:   return masking_launch<int32_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_INT64:   return masking_launch<int64_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_INT8:    return masking_launch<int8_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_UINT8:   return masking_launch<uint8_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_INT16:   return masking_launch<int16_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_UINT16:  return masking_launch<uint16_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_UINT32:  return masking_launch<uint32_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_UINT64:  return masking_launch<uint64_t, false>(m, n, k, len, fill, mode, x, a);
        case RB200_BOOL:    return masking_launch<uint8_t, true>(m, n, k, len, fill, mode, x, a);
        default: RB_UNSUPPORTED("masking: unsupported dtype %d", dtype);
    }
}
