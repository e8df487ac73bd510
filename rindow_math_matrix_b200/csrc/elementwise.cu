// elementwise.cu -- broadcast add / multiply, the rest of the broadcast
// family (section 8f rank 2: maximum, minimum, greater, greaterEqual, less, lessEqual, pow, duplicate --
// PhpMath.php:274-528, 687-722, 860-891) and fill.  All HBM-bound.
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:530-606):
//   trans == false:  A[i*ldA + j] = alpha*X[j*incX] + A[i*ldA + j]     i<m, j<n
//   trans == true :  A[i*ldA + j] = alpha*X[i*incX] + A[i*ldA + j]     i<m, j<n
// (multiply: A = X * A).  In place, one read and one write of A per element:
// algorithmic bytes = 2*m*n*sizeof(T) + len(X)*sizeof(T).
//
// Fast path: 128-bit accesses (4 x f32 / 2 x f64 / ...) when rows are 16-byte aligned and
// n is a multiple of the vector width.  Each thread owns one 16-byte column chunk, keeps
// the X chunk in registers and streams ROWS rows through it with all loads issued before
// the first store (8 x 16 B in flight per thread), so HBM latency is covered by ILP and
// by 8 resident CTAs per SM.  No shared memory: there is no reuse beyond X.
#include "common.cuh"

namespace {

// OP_ADD / OP_MUL are rb200_add / rb200_multiply; the others follow the RB200_BC_* codes + 2
enum { OP_ADD = 0, OP_MUL = 1, OP_MAXIMUM = 2, OP_MINIMUM = 3, OP_GREATER = 4, OP_GREATER_EQUAL = 5, OP_LESS = 6,
       OP_LESS_EQUAL = 7, OP_POW = 8, OP_DUPLICATE = 9 };

template <typename T> __device__ __forceinline__ bool ew_isnan(T) { return false; }
template <> __device__ __forceinline__ bool ew_isnan<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool ew_isnan<double>(double v) { return v != v; }
template <typename T> __device__ __forceinline__ T ew_pow(T a, T x)
{
    // integer dtypes: PHP pow() on ints -- computed in double, stored back truncated
    return (T)pow((double)a, (double)x);
}
template <> __device__ __forceinline__ float ew_pow<float>(float a, float x) { return powf(a, x); }
template <> __device__ __forceinline__ double ew_pow<double>(double a, double x) { return pow(a, x); }

template <typename T> struct VecOf { static constexpr int N = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ T ew_apply_add(T alpha, T x, T a) { return alpha * x + a; }
template <> __device__ __forceinline__ float  ew_apply_add<float>(float alpha, float x, float a)   { return fmaf(alpha, x, a); }
template <> __device__ __forceinline__ double ew_apply_add<double>(double alpha, double x, double a) { return fma(alpha, x, a); }

template <typename T, int OP>
__device__ __forceinline__ T ew_apply(T alpha, T x, T a)
{
    if (OP == OP_ADD) return ew_apply_add<T>(alpha, x, a);
    if (OP == OP_MUL) return x * a;
    // maximum / minimum (PhpMath.php:274-357): a NaN in X is written; a NaN already in A stays
    if (OP == OP_MAXIMUM) return ew_isnan(x) ? x : ((a < x) ? x : a);
    if (OP == OP_MINIMUM) return ew_isnan(x) ? x : ((a > x) ? x : a);
    // comparisons (:362-528): 1 / 0 in the buffer's dtype, false for every NaN operand
    if (OP == OP_GREATER) return (a > x) ? (T)1 : (T)0;
    if (OP == OP_GREATER_EQUAL) return (a >= x) ? (T)1 : (T)0;
    if (OP == OP_LESS) return (a < x) ? (T)1 : (T)0;
    if (OP == OP_LESS_EQUAL) return (a <= x) ? (T)1 : (T)0;
    if (OP == OP_POW) return ew_pow<T>(a, x);          // :687-722
    return x;                                          // OP_DUPLICATE (:860-891)
}

template <typename T> union Pack16 {
    int4 raw;
    T v[16 / sizeof(T)];
};

constexpr int EW_ROWS = 8;      // rows per thread, all loads in flight before the stores
constexpr int EW_THREADS = 256;

// blockDim = (TX, TY), TX*TY == 256.  Thread (tx,ty) owns vector column cv and rows
// row0 + ty + i*TY (i < EW_ROWS), striding over blockIdx.y for matrices taller than the grid.
template <typename T, int OP, bool TRANS>
__global__ void __launch_bounds__(EW_THREADS)
ew_vec_kernel(int64_t m, int64_t nvec, T alpha,
              const T* __restrict__ X, int64_t incX,
              T* __restrict__ A, int64_t ldA)
{
    constexpr int V = VecOf<T>::N;
    const int64_t cv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cv >= nvec) return;
    const int TY = blockDim.y;

    Pack16<T> xv;
    if (!TRANS) {
        xv.raw = __ldg(reinterpret_cast<const int4*>(X) + cv);      // X chunk: reused for every row
    }
    for (int64_t row0 = (int64_t)blockIdx.y * (TY * EW_ROWS) + threadIdx.y; row0 < m;
         row0 += (int64_t)gridDim.y * (TY * EW_ROWS)) {
        Pack16<T> a[EW_ROWS];
        T xs[EW_ROWS];
#pragma unroll
        for (int i = 0; i < EW_ROWS; i++) {
            const int64_t r = row0 + (int64_t)i * TY;
            if (r < m) {
                if (OP != OP_DUPLICATE) a[i].raw = __ldcs(reinterpret_cast<const int4*>(A + r * ldA) + cv);
                if (TRANS) xs[i] = __ldg(X + r * incX);
            }
        }
#pragma unroll
        for (int i = 0; i < EW_ROWS; i++) {
            const int64_t r = row0 + (int64_t)i * TY;
            if (r < m) {
#pragma unroll
                for (int e = 0; e < V; e++) {
                    a[i].v[e] = ew_apply<T, OP>(alpha, TRANS ? xs[i] : xv.v[e], a[i].v[e]);
                }
                __stcs(reinterpret_cast<int4*>(A + r * ldA) + cv, a[i].raw);
            }
        }
    }
}

// General path: any increment / leading dimension / alignment.  One element per thread,
// threadIdx.x along the contiguous j axis.
template <typename T, int OP, bool TRANS>
__global__ void __launch_bounds__(EW_THREADS)
ew_scalar_kernel(int64_t m, int64_t n, T alpha,
                 const T* __restrict__ X, int64_t incX,
                 T* __restrict__ A, int64_t ldA)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int64_t i = blockIdx.y; i < m; i += gridDim.y) {
        const T x = TRANS ? X[i * incX] : X[j * incX];
        T* p = A + i * ldA + j;
        *p = ew_apply<T, OP>(alpha, x, (OP == OP_DUPLICATE) ? x : *p);
    }
}

template <typename T, int OP>
int ew_launch(int trans, int64_t m, int64_t n, double alpha_d,
              const void* Xv, int64_t offX, int64_t incX, void* Av, int64_t offA, int64_t ldA)
{
    rb200::Context& c = rb200::ctx();
    const T* X = reinterpret_cast<const T*>(Xv) + offX;
    T* A = reinterpret_cast<T*>(Av) + offA;
    const T alpha = (T)alpha_d;
    constexpr int V = VecOf<T>::N;

    // a single X element broadcast over a dense column (LinearAlgebra::calcBroadcastFormat's scalar case:
    // m = size, n = 1, ldA = 1; or trans with one row): view A as [m/w, w] rows of X[0]
    if ((!trans && n == 1 && ldA == 1) || (trans && m == 1)) {
        const int64_t total = trans ? n : m;
        int64_t w = 1;
        while (w < 2048 && total % (w * 2) == 0) w *= 2;
        if (w < V) w = total;                 // odd lengths: one long row (scalar kernel when unaligned)
        trans = 1; m = total / w; n = w; ldA = w; incX = 0;
    }

    const bool a_vec_ok = (n % V == 0) && (ldA % V == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool x_vec_ok = trans ? true : (incX == 1 && (reinterpret_cast<uintptr_t>(X) & 15) == 0);
    if (a_vec_ok && x_vec_ok) {
        const int64_t nvec = n / V;
        int tx = 1;
        while (tx < EW_THREADS && tx < nvec) tx <<= 1;
        const int ty = EW_THREADS / tx;
        dim3 block(tx, ty);
        int64_t gy = rb_ceil_div(m, (int64_t)ty * EW_ROWS);
        if (gy > 65535) gy = 65535;          // the kernel strides over the remaining rows
        dim3 grid((unsigned)rb_ceil_div(nvec, tx), (unsigned)gy);
        if (trans) ew_vec_kernel<T, OP, true><<<grid, block, 0, c.stream>>>(m, nvec, alpha, X, incX, A, ldA);
        else       ew_vec_kernel<T, OP, false><<<grid, block, 0, c.stream>>>(m, nvec, alpha, X, incX, A, ldA);
        RB_LAUNCH_CHECK("ew_vec_kernel");
        return RB200_OK;
    }
    dim3 block(EW_THREADS);
    int64_t gy = m < 65535 ? m : 65535;
    dim3 grid((unsigned)rb_ceil_div(n, EW_THREADS), (unsigned)gy);
    if (trans) ew_scalar_kernel<T, OP, true><<<grid, block, 0, c.stream>>>(m, n, alpha, X, incX, A, ldA);
    else       ew_scalar_kernel<T, OP, false><<<grid, block, 0, c.stream>>>(m, n, alpha, X, incX, A, ldA);
    RB_LAUNCH_CHECK("ew_scalar_kernel");
    return RB200_OK;
}

template <int OP>
int ew_dispatch(int dtype, int trans, int64_t m, int64_t n, double alpha,
                const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA)
{
    RB_REQUIRE_INIT();
    if (!X || !A) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1) RB_INVALID("m and n must be greater than 0");
    if (incX < 1 || ldA < 1 || offX < 0 || offA < 0) RB_INVALID("bad increment / leading dimension / offset");
    if (ldA < n) RB_INVALID("ldA (%lld) is smaller than n (%lld)", (long long)ldA, (long long)n);
    switch (dtype) {
        case RB200_FLOAT32: return ew_launch<float, OP>(trans, m, n, alpha, X, offX, incX, A, offA, ldA);
        case RB200_FLOAT64: return ew_launch<double, OP>(trans, m, n, alpha, X, offX, incX, A, offA, ldA);
        case RB200_INT32:   return ew_launch<int32_t, OP>(trans, m, n, alpha, X, offX, incX, A, offA, ldA);
        case RB200_INT64:   return ew_launch<int64_t, OP>(trans, m, n, alpha, X, offX, incX, A, offA, ldA);
        default: RB_UNSUPPORTED("add/multiply: unsupported dtype %d", dtype);
    }
}

// ---- fill ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* __restrict__ p, int64_t n, T value)
{
    constexpr int V = 16 / sizeof(T);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // head: unaligned prefix, body: 16-byte stores, tail: remainder
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    if (i < head) p[i] = value;
    T* body = p + head;
    const int64_t nvec = (n - head) / V;
    Pack16<T> pk;
#pragma unroll
    for (int e = 0; e < V; e++) pk.v[e] = value;
    for (int64_t v = i; v < nvec; v += stride) {
        __stcs(reinterpret_cast<int4*>(body) + v, pk.raw);
    }
    const int64_t tail0 = head + nvec * V;
    if (tail0 + i < n) p[tail0 + i] = value;
}

template <typename T>
int fill_launch(void* dptr, int64_t off, int64_t n, const void* value)
{
    rb200::Context& c = rb200::ctx();
    T v;
    memcpy(&v, value, sizeof(T));
    T* p = reinterpret_cast<T*>(dptr) + off;
    int64_t blocks = rb_ceil_div(n, 256 * 8);
    const int64_t max_blocks = (int64_t)c.sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    fill_kernel<T><<<(unsigned)blocks, 256, 0, c.stream>>>(p, n, v);
    RB_LAUNCH_CHECK("fill_kernel");
    return RB200_OK;
}


__global__ void __launch_bounds__(256) fill_pairs_kernel(uint64_t* __restrict__ p, int64_t n, uint64_t a, uint64_t b)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        *reinterpret_cast<ulonglong2*>(p + 2 * i) = make_ulonglong2(a, b);
    }
}

int fill_pairs_launch(void* dptr, int64_t off, int64_t n, uint64_t a, uint64_t b)
{
    rb200::Context& c = rb200::ctx();
    uint64_t* p = reinterpret_cast<uint64_t*>(dptr) + 2 * off;
    int64_t blocks = rb_ceil_div(n, 256 * 4);
    const int64_t max_blocks = (int64_t)c.sm_count * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    fill_pairs_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(p, n, a, b);
    RB_LAUNCH_CHECK("fill_pairs_kernel");
    return RB200_OK;
}

template <typename T>
int bc_launch(int op, int trans, int64_t m, int64_t n, const void* X, int64_t offX, int64_t incX,
              void* A, int64_t offA, int64_t ldA)
{
    switch (op) {
        case RB200_BC_MAXIMUM:       return ew_launch<T, OP_MAXIMUM>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_MINIMUM:       return ew_launch<T, OP_MINIMUM>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_GREATER:       return ew_launch<T, OP_GREATER>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_GREATER_EQUAL: return ew_launch<T, OP_GREATER_EQUAL>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_LESS:          return ew_launch<T, OP_LESS>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_LESS_EQUAL:    return ew_launch<T, OP_LESS_EQUAL>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_POW:           return ew_launch<T, OP_POW>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        case RB200_BC_DUPLICATE:     return ew_launch<T, OP_DUPLICATE>(trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
        default: RB_INVALID("unknown broadcast op %d", op);
    }
}

}  // namespace

extern "C" {

int rb200_broadcast(int dtype, int op, int trans, int64_t m, int64_t n,
                    void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX)
{
    RB_REQUIRE_INIT();
    if (!X || !A) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1) RB_INVALID("m and n must be greater than 0");
    if (incX < 1 || ldA < 1 || offX < 0 || offA < 0) RB_INVALID("bad increment / leading dimension / offset");
    if (ldA < n) RB_INVALID("ldA (%lld) is smaller than n (%lld)", (long long)ldA, (long long)n);
    switch (dtype) {
        case RB200_FLOAT32: return bc_launch<float>(op, trans, m, n, X, offX, incX, A, offA, ldA);
        case RB200_FLOAT64: return bc_launch<double>(op, trans, m, n, X, offX, incX, A, offA, ldA);
        case RB200_INT32:   return bc_launch<int32_t>(op, trans, m, n, X, offX, incX, A, offA, ldA);
        case RB200_INT64:   return bc_launch<int64_t>(op, trans, m, n, X, offX, incX, A, offA, ldA);
        default: RB_UNSUPPORTED("broadcast op: unsupported dtype %d", dtype);
    }
}

int rb200_add(int dtype, int trans, int64_t m, int64_t n, double alpha,
              const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA)
{
    return ew_dispatch<OP_ADD>(dtype, trans, m, n, alpha, X, offX, incX, A, offA, ldA);
}

int rb200_multiply(int dtype, int trans, int64_t m, int64_t n,
                   const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA)
{
    return ew_dispatch<OP_MUL>(dtype, trans, m, n, 1.0, X, offX, incX, A, offA, ldA);
}

int rb200_buffer_fill(int dtype, void* dptr, int64_t off, int64_t n, const void* value)
{
    RB_REQUIRE_INIT();
    if (!dptr || !value) RB_INVALID("NULL argument");
    if (n < 0 || off < 0) RB_INVALID("negative size/offset");
    if (n == 0) return RB200_OK;
    switch (rb_dtype_size(dtype)) {
        case 1: return fill_launch<uint8_t>(dptr, off, n, value);
        case 2: return fill_launch<uint16_t>(dptr, off, n, value);
        case 4: return fill_launch<uint32_t>(dptr, off, n, value);
        case 8: return fill_launch<uint64_t>(dptr, off, n, value);
        case 16: {
            // complex128: pairs of 8-byte words; equal halves fill as 2n words, anything else word by word would need
            // a 16-byte element type -- two strided fills of n words each do it with the kernel at hand
            uint64_t w[2];
            memcpy(w, value, 16);
            if (w[0] == w[1]) return fill_launch<uint64_t>(dptr, 2 * off, 2 * n, &w[0]);
            int rc = fill_pairs_launch(dptr, off, n, w[0], w[1]);
            return rc;
        }
        default: RB_UNSUPPORTED("fill: unsupported dtype %d", dtype);
    }
}

}  // extern "C"
