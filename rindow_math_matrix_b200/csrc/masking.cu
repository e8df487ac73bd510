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

// rows of A are (i, j) pairs, rowlen = k*len elements each, masked by the k bytes of X[i,:].  One thread per 16-byte
// vector of A: its V = 16/sizeof(T) consecutive elements either share one mask byte (len % V == 0) or map to V
// consecutive mask bytes (len == 1).  Row / position by multiply-high divisions when they are exact.  Mode 0 never
// reads A.
struct MkVec {
    int64_t total;          // vectors in A
    int64_t vpr, n, k, len; // vectors per row, rows per mask row
    unsigned m_vpr, m_n, m_len;
    int small;              // 32-bit multiply-high divisions are exact
    int x_word;             // k % 4 == 0 and X 4-byte aligned: mask bytes can be read as 32-bit words
};

template <typename T, bool ADD, bool IS_BOOL, bool LEN1>
__global__ void __launch_bounds__(MK_THREADS)
masking_vec_kernel(MkVec p, T fill, const uint8_t* __restrict__ X, T* __restrict__ A)
{
    constexpr int V = 16 / (int)sizeof(T);
    for (int64_t g = (int64_t)blockIdx.x * MK_THREADS + threadIdx.x; g < p.total; g += (int64_t)gridDim.x * MK_THREADS) {
        int64_t r, ev, i, h = 0;
        if (p.small) {
            const unsigned gg = (unsigned)g;
            const unsigned rr = p.vpr == 1 ? gg : __umulhi(gg, p.m_vpr);
            ev = gg - rr * (unsigned)p.vpr;
            i = p.n == 1 ? rr : __umulhi(rr, p.m_n);
            r = rr;
            if (!LEN1) h = p.len == V ? ev : (int64_t)__umulhi((unsigned)(ev * V), p.m_len);
        } else {
            r = g / p.vpr; ev = g - r * p.vpr;
            i = r / p.n;
            if (!LEN1) h = (ev * V) / p.len;
        }
        const uint8_t* xrow = X + i * p.k;
        T* a = A + g * V;
        bool off[V];
        bool any = false, all = true;
        if (LEN1) {
            if (V == 4 && p.x_word) {                       // the four mask bytes of this vector in one load
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(xrow + ev * 4));
#pragma unroll
                for (int e = 0; e < V; e++) { off[e] = ((w >> (8 * e)) & 0xffu) == 0; any |= off[e]; all &= off[e]; }
            } else {
#pragma unroll
                for (int e = 0; e < V; e++) { off[e] = !xrow[ev * V + e]; any |= off[e]; all &= off[e]; }
            }
        } else {
            any = all = !xrow[h];
#pragma unroll
            for (int e = 0; e < V; e++) off[e] = any;
        }
        if (!any) continue;
        union { int4 raw; T v[V]; } pk;
        if (!ADD && all) {
#pragma unroll
            for (int e = 0; e < V; e++) pk.v[e] = fill;
            __stcs(reinterpret_cast<int4*>(a), pk.raw);
        } else if (!ADD) {
#pragma unroll
            for (int e = 0; e < V; e++) if (off[e]) a[e] = fill;
        } else {
            pk.raw = *reinterpret_cast<const int4*>(a);
#pragma unroll
            for (int e = 0; e < V; e++) {
                if (off[e]) pk.v[e] = IS_BOOL ? (T)((pk.v[e] != 0) || (fill != 0)) : mk_add<T>(pk.v[e], fill);
            }
            *reinterpret_cast<int4*>(a) = pk.raw;
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
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t rowlen = k * len;
    if ((rowlen % V) == 0 && (len == 1 || (len % V) == 0) && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
        MkVec q;
        q.vpr = rowlen / V; q.n = n; q.k = k; q.len = len;
        q.total = m * n * q.vpr;
        const int64_t dmax = q.vpr > n ? (q.vpr > len ? q.vpr : len) : (n > len ? n : len);
        q.small = dmax < (1LL << 16) && q.total < (1LL << 32) / dmax && rowlen < (1LL << 32) / (len > 1 ? len : 1);
        auto magic = [](int64_t d) { return (unsigned)((1ULL << 32) / (unsigned long long)d) + 1u; };
        q.m_vpr = magic(q.vpr); q.m_n = magic(n); q.m_len = magic(len);
        q.x_word = (k % 4) == 0 && (reinterpret_cast<uintptr_t>(X) & 3) == 0;
        int64_t g2 = rb_ceil_div(q.total, MK_THREADS);
        if (g2 > cap) g2 = cap;
        if (len == 1) {
            if (mode == 0) masking_vec_kernel<T, false, IS_BOOL, true><<<(unsigned)g2, MK_THREADS, 0, c.stream>>>(q, f, x, a);
            else           masking_vec_kernel<T, true, IS_BOOL, true><<<(unsigned)g2, MK_THREADS, 0, c.stream>>>(q, f, x, a);
        } else {
            if (mode == 0) masking_vec_kernel<T, false, IS_BOOL, false><<<(unsigned)g2, MK_THREADS, 0, c.stream>>>(q, f, x, a);
            else           masking_vec_kernel<T, true, IS_BOOL, false><<<(unsigned)g2, MK_THREADS, 0, c.stream>>>(q, f, x, a);
        }
        RB_LAUNCH_CHECK("masking_vec_kernel");
        return RB200_OK;
    }
    if (mode == 0) masking_kernel<T, false, IS_BOOL><<<(unsigned)grid, MK_THREADS, 0, c.stream>>>(total, n, k, len, f, x, a);
    else           masking_kernel<T, true, IS_BOOL><<<(unsigned)grid, MK_THREADS, 0, c.stream>>>(total, n, k, len, f, x, a);
    RB_LAUNCH_CHECK("masking_kernel");
    return RB200_OK;
}

// bandpart (PhpMath.php bandpart): A[b,i,j] = 0 where (lower >= 0 && i-j > lower) || (upper >= 0 && j-i > upper).
// One thread per element word; zero stores only (A is never read).
template <typename W>
__global__ void __launch_bounds__(MK_THREADS)
bandpart_kernel(int64_t total, int64_t n, int64_t k, int64_t lower, int64_t upper, W* __restrict__ A)
{
    const int64_t nk = n * k;
    for (int64_t t = (int64_t)blockIdx.x * MK_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * MK_THREADS) {
        const int64_t r = t % nk;
        const int64_t i = r / k, j = r - i * k;
        if ((lower >= 0 && (i - j) > lower) || (upper >= 0 && (j - i) > upper)) A[t] = (W)0;
    }
}

template <typename W>
int bandpart_launch(int64_t m, int64_t n, int64_t k, int64_t lower, int64_t upper, void* A)
{
    rb200::Context& c = rb200::ctx();
    const int64_t total = m * n * k;
    int64_t grid = rb_ceil_div(total, MK_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    bandpart_kernel<W><<<(unsigned)grid, MK_THREADS, 0, c.stream>>>(total, n, k, lower, upper, reinterpret_cast<W*>(A));
    RB_LAUNCH_CHECK("bandpart_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_bandpart(int dtype, int64_t m, int64_t n, int64_t k, void* A, int64_t offA, int64_t lower, int64_t upper)
{
    RB_REQUIRE_INIT();
    if (!A) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || offA < 0) RB_INVALID("bandpart: bad shape / offset");
    const int es = rb_dtype_size(dtype);
    char* a = reinterpret_cast<char*>(A) + offA * es;
    switch (es) {                                   // zero is all-bits-zero for every supported dtype
        case 1: return bandpart_launch<uint8_t>(m, n, k, lower, upper, a);
        case 2: return bandpart_launch<uint16_t>(m, n, k, lower, upper, a);
        case 4: return bandpart_launch<uint32_t>(m, n, k, lower, upper, a);
        case 8: return bandpart_launch<uint64_t>(m, n, k, lower, upper, a);
        default: RB_UNSUPPORTED("bandpart: unsupported dtype %d", dtype);
    }
}

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
