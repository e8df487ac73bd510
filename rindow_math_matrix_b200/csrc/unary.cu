// unary.cu -- the in-place unary family, equal / notEqual, astype, the scalar reductions sum / imax / imin and
// updateAddOnehot (the elementwise callers either side of the hot path).
// All HBM-bound streaming kernels: 128-bit accesses with four packs in flight per thread when the increment is 1
// and the operand 16-byte aligned, strided scalar otherwise.
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php): increment :221-239, reciprocal :244-268, square
// :611-629, sqrt :634-652, rsqrt :657-682, exp :727-744, log :749-767, tanh :772-789, sin :794-811, cos :816-833,
// tan :838-855, equal :1470-1493, notEqual :1499-1522, not :1528-1547, astype :1710-1778, nan2num :2830-2851,
// isnan :2853-2875, sum :150-166, imax :171-191, imin :196-216, updateAddOnehot :1438-1464.
#include "common.cuh"

#include <math.h>

namespace {

constexpr int UN_THREADS = 256;
constexpr int UN_UNROLL = 8;

template <typename T> union UnPack { int4 raw; T v[16 / sizeof(T)]; };

__device__ __forceinline__ float  un_mul(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double un_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  un_add(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double un_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  un_inf(float)  { return INFINITY; }
__device__ __forceinline__ double un_inf(double) { return (double)INFINITY; }

// float / double ops.  alpha*x + beta is one multiply and one add (PHP evaluates them separately), never an FMA.
template <typename T, int OP>
__device__ __forceinline__ T un_apply(T x, T alpha, T beta)
{
    if (OP == RB200_UN_INCREMENT) return un_add(un_mul(alpha, x), beta);
    if (OP == RB200_UN_RECIPROCAL) {
        const T t = un_add(un_mul(alpha, x), beta);
        return (t == (T)0) ? un_inf(t) : (T)1 / t;
    }
    if (OP == RB200_UN_SQUARE) return un_mul(x, x);
    if (OP == RB200_UN_SQRT) return sqrt(x);
    if (OP == RB200_UN_RSQRT) {
        const T t = un_add(un_mul(alpha, sqrt(x)), beta);
        return (t != (T)0) ? (T)1 / t : un_inf(t);
    }
    if (OP == RB200_UN_EXP) return exp(x);
    if (OP == RB200_UN_LOG) return log(x);
    if (OP == RB200_UN_TANH) return tanh(x);
    if (OP == RB200_UN_SIN) return sin(x);
    if (OP == RB200_UN_COS) return cos(x);
    if (OP == RB200_UN_TAN) return tan(x);
    if (OP == RB200_UN_NOT) return (x == (T)0) ? (T)1 : (T)0;
    if (OP == RB200_UN_NAN2NUM) return (x != x) ? alpha : x;
    return (x != x) ? (T)1 : (T)0;     // RB200_UN_ISNAN
}

template <typename T, int OP>
__global__ void __launch_bounds__(UN_THREADS)
unary_vec_kernel(int64_t nvec, T alpha, T beta, T* __restrict__ X)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t chunk = (int64_t)UN_THREADS * UN_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
        UnPack<T> x[UN_UNROLL];
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < nvec) x[u].raw = __ldcs(reinterpret_cast<const int4*>(X) + i);
        }
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < nvec) {
#pragma unroll
                for (int e = 0; e < V; e++) x[u].v[e] = un_apply<T, OP>(x[u].v[e], alpha, beta);
                __stcs(reinterpret_cast<int4*>(X) + i, x[u].raw);
            }
        }
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(UN_THREADS)
unary_strided_kernel(int64_t n, T alpha, T beta, T* __restrict__ X, int64_t incX)
{
    const int64_t stride = (int64_t)gridDim.x * UN_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < n; i += stride) {
        T* p = X + i * incX;
        *p = un_apply<T, OP>(*p, alpha, beta);
    }
}

// 16-byte aligned body as packs, unaligned head / tail through the strided kernel
template <typename T, int OP>
int unary_launch(int64_t n, double alpha_d, double beta_d, void* Xv, int64_t offX, int64_t incX)
{
    rb200::Context& c = rb200::ctx();
    T* X = reinterpret_cast<T*>(Xv) + offX;
    const T alpha = (T)alpha_d, beta = (T)beta_d;
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (incX == 1 && n >= 4 * V) {
        int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(X) & 15)) & 15) / sizeof(T));
        if (head) {
            unary_strided_kernel<T, OP><<<1, 32, 0, c.stream>>>(head, alpha, beta, X, 1);
            RB_LAUNCH_CHECK("unary_strided_kernel");
        }
        const int64_t nvec = (n - head) / V;
        int64_t grid = rb_ceil_div(nvec, (int64_t)UN_THREADS * UN_UNROLL);
        if (grid > cap) grid = cap;
        unary_vec_kernel<T, OP><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(nvec, alpha, beta, X + head);
        RB_LAUNCH_CHECK("unary_vec_kernel");
        const int64_t done = head + nvec * V;
        if (done < n) {
            unary_strided_kernel<T, OP><<<1, 32, 0, c.stream>>>(n - done, alpha, beta, X + done, 1);
            RB_LAUNCH_CHECK("unary_strided_kernel");
        }
        return RB200_OK;
    }
    int64_t grid = rb_ceil_div(n, UN_THREADS);
    if (grid > cap) grid = cap;
    unary_strided_kernel<T, OP><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(n, alpha, beta, X, incX);
    RB_LAUNCH_CHECK("unary_strided_kernel");
    return RB200_OK;
}

template <typename T>
int unary_float_dispatch(int op, int64_t n, double alpha, double beta, void* X, int64_t offX, int64_t incX)
{
    switch (op) {
#define RB_UN_CASE(code) case code: return unary_launch<T, code>(n, alpha, beta, X, offX, incX);
        RB_UN_CASE(RB200_UN_INCREMENT) RB_UN_CASE(RB200_UN_RECIPROCAL) RB_UN_CASE(RB200_UN_SQUARE)
        RB_UN_CASE(RB200_UN_SQRT) RB_UN_CASE(RB200_UN_RSQRT) RB_UN_CASE(RB200_UN_EXP) RB_UN_CASE(RB200_UN_LOG)
        RB_UN_CASE(RB200_UN_TANH) RB_UN_CASE(RB200_UN_SIN) RB_UN_CASE(RB200_UN_COS) RB_UN_CASE(RB200_UN_TAN)
        RB_UN_CASE(RB200_UN_NOT) RB_UN_CASE(RB200_UN_NAN2NUM) RB_UN_CASE(RB200_UN_ISNAN)
#undef RB_UN_CASE
        default: RB_INVALID("unknown unary op %d", op);
    }
}

// `not` on integer / bool buffers: X := (X == 0) ? 1 : 0
template <typename T>
__global__ void __launch_bounds__(UN_THREADS) not_int_kernel(int64_t n, T* __restrict__ X, int64_t incX)
{
    const int64_t stride = (int64_t)gridDim.x * UN_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < n; i += stride) {
        T* p = X + i * incX;
        *p = (*p == (T)0) ? (T)1 : (T)0;
    }
}

template <typename T>
int not_int_launch(int64_t n, void* Xv, int64_t offX, int64_t incX)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(n, UN_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (grid > cap) grid = cap;
    not_int_kernel<T><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(n, reinterpret_cast<T*>(Xv) + offX, incX);
    RB_LAUNCH_CHECK("not_int_kernel");
    return RB200_OK;
}

// ---- equal / notEqual ---------------------------------------------------------------------------------------
// EQ semantics per element type: IEEE for floats (NaN != NaN, -0 == 0), value equality for integers (U = the
// unsigned integer of the element's size: equality and the stored 1 / 0 do not depend on signedness).
template <typename T, bool NE>
__global__ void __launch_bounds__(UN_THREADS)
compare_kernel(int64_t n, const T* __restrict__ X, int64_t incX, T* __restrict__ Y, int64_t incY)
{
    const int64_t stride = (int64_t)gridDim.x * UN_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < n; i += stride) {
        T* y = Y + i * incY;
        const bool eq = (*y == X[i * incX]);
        *y = (eq != NE) ? (T)1 : (T)0;
    }
}

template <typename T, bool NE>
__global__ void __launch_bounds__(UN_THREADS)
compare_vec_kernel(int64_t nvec, const T* __restrict__ X, T* __restrict__ Y)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t chunk = (int64_t)UN_THREADS * UN_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
        UnPack<T> x[UN_UNROLL], y[UN_UNROLL];
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < nvec) {
                x[u].raw = __ldcs(reinterpret_cast<const int4*>(X) + i);
                y[u].raw = __ldcs(reinterpret_cast<const int4*>(Y) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < nvec) {
#pragma unroll
                for (int e = 0; e < V; e++) y[u].v[e] = ((y[u].v[e] == x[u].v[e]) != NE) ? (T)1 : (T)0;
                __stcs(reinterpret_cast<int4*>(Y) + i, y[u].raw);
            }
        }
    }
}

template <typename T, bool NE>
int compare_launch(int64_t n, const void* Xv, int64_t offX, int64_t incX, void* Yv, int64_t offY, int64_t incY)
{
    rb200::Context& c = rb200::ctx();
    const T* X = reinterpret_cast<const T*>(Xv) + offX;
    T* Y = reinterpret_cast<T*>(Yv) + offY;
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t cap = (int64_t)c.sm_count * 16;
    const bool aligned = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0;
    int64_t done = 0;
    if (incX == 1 && incY == 1 && aligned && n >= V) {
        const int64_t nvec = n / V;
        int64_t grid = rb_ceil_div(nvec, (int64_t)UN_THREADS * UN_UNROLL);
        if (grid > cap) grid = cap;
        compare_vec_kernel<T, NE><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(nvec, X, Y);
        RB_LAUNCH_CHECK("compare_vec_kernel");
        done = nvec * V;
        if (done == n) return RB200_OK;
    }
    int64_t grid = rb_ceil_div(n - done, UN_THREADS);
    if (grid > cap) grid = cap;
    compare_kernel<T, NE><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(n - done, X + done * incX, incX,
                                                                        Y + done * incY, incY);
    RB_LAUNCH_CHECK("compare_kernel");
    return RB200_OK;
}

template <typename T>
int compare_op(int op, int64_t n, const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY)
{
    if (op == RB200_CMP_EQUAL) return compare_launch<T, false>(n, X, offX, incX, Y, offY, incY);
    if (op == RB200_CMP_NOT_EQUAL) return compare_launch<T, true>(n, X, offX, incX, Y, offY, incY);
    RB_INVALID("unknown compare op %d", op);
}

// ---- astype -------------------------------------------------------------------------------------------------
// PHP rules (:1739-1777): bool source -> 1 / 0; bool target -> (x != 0) (NaN is truthy); float target ->
// (float)x; integer target -> (int)x = truncation toward zero, unsigned targets keep the low bits ($mask & x).
struct BoolTag { uint8_t b; };

template <typename S> __device__ __forceinline__ bool as_truthy(S x) { return x != (S)0; }

template <typename D, typename S> struct AsCast {
    __device__ static __forceinline__ D run(S x) { return (D)x; }
};
// float -> integer: through int64 (PHP's (int) on a float), then the target keeps its low bits
template <typename D> struct AsCast<D, float>  { __device__ static __forceinline__ D run(float x)  { return (D)(int64_t)x; } };
template <typename D> struct AsCast<D, double> { __device__ static __forceinline__ D run(double x) { return (D)(int64_t)x; } };
template <> struct AsCast<float, float>   { __device__ static __forceinline__ float  run(float x)  { return x; } };
template <> struct AsCast<double, float>  { __device__ static __forceinline__ double run(float x)  { return (double)x; } };
template <> struct AsCast<float, double>  { __device__ static __forceinline__ float  run(double x) { return (float)x; } };
template <> struct AsCast<double, double> { __device__ static __forceinline__ double run(double x) { return x; } };
template <> struct AsCast<uint64_t, float>  { __device__ static __forceinline__ uint64_t run(float x)  { return (uint64_t)(int64_t)x; } };
template <> struct AsCast<uint64_t, double> { __device__ static __forceinline__ uint64_t run(double x) { return (uint64_t)(int64_t)x; } };

template <typename D, typename S, bool TO_BOOL>
__global__ void __launch_bounds__(UN_THREADS)
astype_kernel(int64_t n, const S* __restrict__ X, int64_t incX, D* __restrict__ Y, int64_t incY)
{
    const int64_t chunk = (int64_t)UN_THREADS * UN_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < n; base += (int64_t)gridDim.x * chunk) {
        S x[UN_UNROLL];
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < n) x[u] = X[i * incX];
        }
#pragma unroll
        for (int u = 0; u < UN_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < n) Y[i * incY] = TO_BOOL ? (D)(as_truthy<S>(x[u]) ? 1 : 0) : AsCast<D, S>::run(x[u]);
        }
    }
}

// unit-stride form: a thread moves groups of four consecutive elements (4..32-byte accesses on either side),
// four groups in flight
template <typename T> struct alignas(sizeof(T) * 4) AsVec4 { T v[4]; };
constexpr int AS_UNROLL = 4;

template <typename D, typename S, bool TO_BOOL>
__global__ void __launch_bounds__(UN_THREADS)
astype_vec_kernel(int64_t ngroups, const S* __restrict__ X, D* __restrict__ Y)
{
    const AsVec4<S>* xv = reinterpret_cast<const AsVec4<S>*>(X);
    AsVec4<D>* yv = reinterpret_cast<AsVec4<D>*>(Y);
    const int64_t chunk = (int64_t)UN_THREADS * AS_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < ngroups; base += (int64_t)gridDim.x * chunk) {
        AsVec4<S> x[AS_UNROLL];
#pragma unroll
        for (int u = 0; u < AS_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < ngroups) x[u] = xv[i];
        }
#pragma unroll
        for (int u = 0; u < AS_UNROLL; u++) {
            const int64_t i = base + u * UN_THREADS + threadIdx.x;
            if (i < ngroups) {
                AsVec4<D> y;
#pragma unroll
                for (int e = 0; e < 4; e++) y.v[e] = TO_BOOL ? (D)(as_truthy<S>(x[u].v[e]) ? 1 : 0) : AsCast<D, S>::run(x[u].v[e]);
                yv[i] = y;
            }
        }
    }
}

template <typename D, typename S>
int astype_launch(bool to_bool, int64_t n, const void* Xv, int64_t offX, int64_t incX, void* Yv, int64_t offY, int64_t incY)
{
    rb200::Context& c = rb200::ctx();
    const S* X = reinterpret_cast<const S*>(Xv) + offX;
    D* Y = reinterpret_cast<D*>(Yv) + offY;
    const int64_t cap = (int64_t)c.sm_count * 16;
    int64_t done = 0;
    const bool aligned = (reinterpret_cast<uintptr_t>(X) % (sizeof(S) * 4) == 0) && (reinterpret_cast<uintptr_t>(Y) % (sizeof(D) * 4) == 0);
    if (incX == 1 && incY == 1 && aligned && n >= 4) {
        const int64_t ngroups = n / 4;
        int64_t grid = rb_ceil_div(ngroups, (int64_t)UN_THREADS * AS_UNROLL);
        if (grid > cap) grid = cap;
        if (to_bool) astype_vec_kernel<D, S, true><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(ngroups, X, Y);
        else         astype_vec_kernel<D, S, false><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(ngroups, X, Y);
        RB_LAUNCH_CHECK("astype_vec_kernel");
        done = ngroups * 4;
        if (done == n) return RB200_OK;
    }
    const int64_t rest = n - done;
    int64_t grid = rb_ceil_div(rest, (int64_t)UN_THREADS * UN_UNROLL);
    if (grid > cap) grid = cap;
    if (to_bool) astype_kernel<D, S, true><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(rest, X + done * incX, incX, Y + done * incY, incY);
    else         astype_kernel<D, S, false><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(rest, X + done * incX, incX, Y + done * incY, incY);
    RB_LAUNCH_CHECK("astype_kernel");
    return RB200_OK;
}

template <typename D>
int astype_from(int from, bool to_bool, int64_t n, const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY)
{
    switch (from) {
        case RB200_BOOL: case RB200_UINT8: return astype_launch<D, uint8_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_INT8:    return astype_launch<D, int8_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_INT16:   return astype_launch<D, int16_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT16:  return astype_launch<D, uint16_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_INT32:   return astype_launch<D, int32_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT32:  return astype_launch<D, uint32_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_INT64:   return astype_launch<D, int64_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT64:  return astype_launch<D, uint64_t>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_FLOAT32: return astype_launch<D, float>(to_bool, n, X, offX, incX, Y, offY, incY);
        case RB200_FLOAT64: return astype_launch<D, double>(to_bool, n, X, offX, incX, Y, offY, incY);
        default: RB_UNSUPPORTED("astype: unsupported source dtype %d", from);
    }
}

// ---- sum / imax / imin: one launch, CTA partials + ticket, the last CTA merges them in index order -------------
constexpr int SR_THREADS = 256;
enum { SR_SUM = 0, SR_IMAX = 1, SR_IMIN = 2, SR_ASUM = 3, SR_NRM2 = 4, SR_DOT = 5, SR_IAMAX = 6, SR_IAMIN = 7 };

template <typename T> __device__ __forceinline__ bool sr_isnan(T) { return false; }
template <> __device__ __forceinline__ bool sr_isnan<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool sr_isnan<double>(double v) { return v != v; }

template <typename V> __device__ __forceinline__ V sr_shfl(V v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
template <> __device__ __forceinline__ int8_t sr_shfl<int8_t>(int8_t v, int d) { return (int8_t)__shfl_down_sync(0xffffffffu, (int)v, d); }
template <> __device__ __forceinline__ uint8_t sr_shfl<uint8_t>(uint8_t v, int d) { return (uint8_t)__shfl_down_sync(0xffffffffu, (int)v, d); }
template <> __device__ __forceinline__ int16_t sr_shfl<int16_t>(int16_t v, int d) { return (int16_t)__shfl_down_sync(0xffffffffu, (int)v, d); }
template <> __device__ __forceinline__ uint16_t sr_shfl<uint16_t>(uint16_t v, int d) { return (uint16_t)__shfl_down_sync(0xffffffffu, (int)v, d); }

// SUM family: double accumulator for every dtype (PhpMath::sum starts from 0.0 and adds PHP numbers, :161-164).
// F selects the term: the element (sum), |x| (PhpBlas::asum :266-286), x*x (nrm2 :366-392, sqrt at the end) or
// x*y (dot :196-222) -- each a double operation on the widened operands, as PHP evaluates it.
enum { SRF_ID = 0, SRF_ABS = 1, SRF_SQR = 2, SRF_DOT = 3 };
template <typename T, int F> struct SrSumF {
    struct Acc { double val; };
    __device__ static __forceinline__ double term(T x, T y)
    {
        const double dx = (double)x;
        if (F == SRF_ABS) return fabs(dx);
        if (F == SRF_SQR) return dx * dx;
        if (F == SRF_DOT) return dx * (double)y;
        return dx;
    }
    __device__ static __forceinline__ Acc init() { Acc a; a.val = 0.0; return a; }
    __device__ static __forceinline__ void take(Acc& a, T x, T y, int64_t) { a.val += term(x, y); }
    // a 16-byte pack is summed as a tree before it joins the running sum: the dependent DADD chain per thread is one
    // add per pack instead of one per element (the kernel is latency-bound otherwise: 40 % of DRAM throughput)
    template <int V> __device__ static __forceinline__ void take_pack(Acc& a, const T* x, const T* y, int64_t)
    {
        double s[V];
#pragma unroll
        for (int e = 0; e < V; e++) s[e] = term(x[e], y[e]);
#pragma unroll
        for (int w = V / 2; w >= 1; w >>= 1) {
#pragma unroll
            for (int e = 0; e < w; e++) s[e] += s[e + w];
        }
        a.val += s[0];
    }
    __device__ static __forceinline__ Acc merge(const Acc& a, const Acc& b) { Acc r; r.val = a.val + b.val; return r; }
    __device__ static __forceinline__ Acc shfl(const Acc& a, int d) { Acc r; r.val = sr_shfl(a.val, d); return r; }
    __device__ static __forceinline__ int64_t finish(const Acc& a, int64_t, const T*)
    {
        return __double_as_longlong(F == SRF_SQR ? sqrt(a.val) : a.val);
    }
};

template <typename T> __device__ __forceinline__ T sr_abs(T v) { return v < (T)0 ? (T)(-v) : v; }
template <> __device__ __forceinline__ uint8_t sr_abs<uint8_t>(uint8_t v) { return v; }
template <> __device__ __forceinline__ uint16_t sr_abs<uint16_t>(uint16_t v) { return v; }
template <> __device__ __forceinline__ uint32_t sr_abs<uint32_t>(uint32_t v) { return v; }
template <> __device__ __forceinline__ float sr_abs<float>(float v) { return fabsf(v); }
template <> __device__ __forceinline__ double sr_abs<double>(double v) { return fabs(v); }

// IMAX / IMIN: the candidate is the first strict extremum among the non-NaN elements.  A thread visits its elements
// in increasing index order, so the strict comparison alone keeps the lowest index; merges break ties by index.
// ABS: the BLAS forms iamax / iamin (PhpBlas.php:288-348) compare |x| and start from element 0, so a NaN there sticks.
template <typename T, bool MAX, bool ABS> struct SrArg {
    struct Acc { T val; int64_t idx; };      // idx < 0: no candidate yet
    __device__ static __forceinline__ Acc init() { Acc a; a.val = (T)0; a.idx = -1; return a; }
    __device__ static __forceinline__ void take(Acc& a, T x, T, int64_t i)
    {
        if (sr_isnan<T>(x)) return;                       // NaN elements never win (imax :185, imin :210)
        const T v = ABS ? sr_abs<T>(x) : x;
        if (a.idx < 0 || (MAX ? (v > a.val) : (v < a.val))) { a.val = v; a.idx = i; }
    }
    template <int V> __device__ static __forceinline__ void take_pack(Acc& a, const T* x, const T*, int64_t i0)
    {
#pragma unroll
        for (int e = 0; e < V; e++) take(a, x[e], x[e], i0 + e);
    }
    __device__ static __forceinline__ Acc merge(const Acc& a, const Acc& b)
    {
        if (b.idx < 0) return a;
        if (a.idx < 0) return b;
        const bool better = (MAX ? (b.val > a.val) : (b.val < a.val)) || (b.val == a.val && b.idx < a.idx);
        return better ? b : a;
    }
    __device__ static __forceinline__ Acc shfl(const Acc& a, int d) { Acc r; r.val = sr_shfl(a.val, d); r.idx = sr_shfl(a.idx, d); return r; }
    __device__ static __forceinline__ int64_t finish(const Acc& a, int64_t n, const T* X)
    {
        if (MAX && !ABS) return (a.idx < 0) ? n - 1 : a.idx;      // all NaN: the scan ends on n-1 (:185)
        return (a.idx < 0 || sr_isnan<T>(X[0])) ? 0 : a.idx;      // a NaN at 0 is never displaced (:210, blas :296)
    }
};

// out[0] = the sum (as double bits) or the index
template <typename T, typename Pol, bool VEC, bool TWO>
__global__ void __launch_bounds__(SR_THREADS)
scalar_reduce_kernel(int64_t n, const T* __restrict__ X, int64_t incX, const T* __restrict__ Y, int64_t incY, int64_t seg,
                     typename Pol::Acc* __restrict__ partials, int* __restrict__ counter, int64_t* __restrict__ out)
{
    using Acc = typename Pol::Acc;
    constexpr int V = 16 / (int)sizeof(T);
    __shared__ Acc warp_acc[SR_THREADS / 32];
    __shared__ bool is_last;
    const int64_t begin = (int64_t)blockIdx.x * seg;
    const int64_t end = (begin + seg < n) ? begin + seg : n;
    Acc acc = Pol::init();
    if (VEC) {                                      // unit increments, operands 16-byte aligned, seg % V == 0
        const int64_t v0 = begin / V, v1 = end / V;
        const int4* xv = reinterpret_cast<const int4*>(X);
        const int4* yv = reinterpret_cast<const int4*>(Y);
        for (int64_t base = v0; base < v1; base += SR_THREADS * UN_UNROLL) {
            UnPack<T> p[UN_UNROLL], q[TWO ? UN_UNROLL : 1];
#pragma unroll
            for (int u = 0; u < UN_UNROLL; u++) {
                const int64_t i = base + u * SR_THREADS + threadIdx.x;
                if (i < v1) { p[u].raw = __ldcs(xv + i); if (TWO) q[u].raw = __ldcs(yv + i); }
            }
#pragma unroll
            for (int u = 0; u < UN_UNROLL; u++) {
                const int64_t i = base + u * SR_THREADS + threadIdx.x;
                if (i < v1) Pol::template take_pack<V>(acc, p[u].v, TWO ? q[u].v : p[u].v, i * V);
            }
        }
        for (int64_t i = v1 * V + threadIdx.x; i < end; i += SR_THREADS) Pol::take(acc, X[i], TWO ? Y[i] : X[i], i);   // last CTA's tail
    } else {
        for (int64_t i = begin + threadIdx.x; i < end; i += SR_THREADS) {
            const T x = X[i * incX];
            Pol::take(acc, x, TWO ? Y[i * incY] : x, i);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = Pol::merge(acc, Pol::shfl(acc, d));
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Acc a = warp_acc[0];
        for (int w = 1; w < SR_THREADS / 32; w++) a = Pol::merge(a, warp_acc[w]);
        partials[blockIdx.x] = a;
        __threadfence();
        const int ticket = atomicAdd(counter, 1);
        is_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // the last CTA merges the partials: thread-strided loads (L2, bypassing L1), shuffle tree, one writer
    static_assert(sizeof(Acc) % 8 == 0, "partials are moved as 8-byte words");
    acc = Pol::init();
    bool have = false;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += SR_THREADS) {
        Acc pb;
        const long long* src = reinterpret_cast<const long long*>(partials + b);
        long long* dst = reinterpret_cast<long long*>(&pb);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(Acc) / 8); k++) dst[k] = __ldcg(src + k);
        acc = have ? Pol::merge(acc, pb) : pb;
        have = true;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = Pol::merge(acc, Pol::shfl(acc, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Acc a = warp_acc[0];
        for (int w = 1; w < SR_THREADS / 32; w++) a = Pol::merge(a, warp_acc[w]);
        out[0] = Pol::finish(a, n, X);
        *counter = 0;                                                          // self-resetting ticket
    }
}

template <typename T, typename Pol, bool TWO>
int scalar_reduce_launch(int64_t n, const void* Xv, int64_t offX, int64_t incX, const void* Yv, int64_t offY, int64_t incY,
                         int64_t* host_out)
{
    using Acc = typename Pol::Acc;
    rb200::Context& c = rb200::ctx();
    const T* X = reinterpret_cast<const T*>(Xv) + offX;
    const T* Y = TWO ? reinterpret_cast<const T*>(Yv) + offY : X;
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = (incX == 1) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && n >= V &&
                     (!TWO || ((incY == 1) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0)));
    // contiguous segments, a multiple of one CTA pass, sized for one wave of 8 CTAs per SM
    const int64_t quantum = (int64_t)SR_THREADS * UN_UNROLL * V;
    int64_t ctas = rb_ceil_div(n, quantum);
    static int per_sm = 0;                                       // resident CTAs per SM of this instantiation
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scalar_reduce_kernel<T, Pol, true, TWO>, SR_THREADS, 0) != cudaSuccess ||
            per_sm < 1) per_sm = 4;
    }
    const int64_t cap = (int64_t)c.sm_count * per_sm;            // exactly one wave: no tail
    if (ctas > cap) ctas = cap;
    int64_t seg = rb_ceil_div(rb_ceil_div(n, ctas), quantum) * quantum;
    ctas = rb_ceil_div(n, seg);
    const size_t bytes = (size_t)ctas * sizeof(Acc) + 16;
    int rc = rb200::ensure_workspace(bytes);
    if (rc != RB200_OK) return rc;
    int* counter = nullptr;
    rc = rb200::ensure_counters(1, &counter);
    if (rc != RB200_OK) return rc;
    int64_t* out = reinterpret_cast<int64_t*>(c.workspace);
    Acc* partials = reinterpret_cast<Acc*>(reinterpret_cast<char*>(c.workspace) + 16);
    if (vec) scalar_reduce_kernel<T, Pol, true, TWO><<<(unsigned)ctas, SR_THREADS, 0, c.stream>>>(n, X, 1, Y, 1, seg, partials, counter, out);
    else     scalar_reduce_kernel<T, Pol, false, TWO><<<(unsigned)ctas, SR_THREADS, 0, c.stream>>>(n, X, incX, Y, incY, seg, partials, counter, out);
    RB_LAUNCH_CHECK("scalar_reduce_kernel");
    RB_CUDA(cudaMemcpyAsync(c.pinned_scalar, out, sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
    RB_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(host_out, c.pinned_scalar, sizeof(int64_t));
    return RB200_OK;
}

template <typename T, int OP>
int scalar_reduce_op(int64_t n, const void* X, int64_t offX, int64_t incX, const void* Y, int64_t offY, int64_t incY, int64_t* host_out)
{
    if (OP == SR_SUM)   return scalar_reduce_launch<T, SrSumF<T, SRF_ID>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    if (OP == SR_ASUM)  return scalar_reduce_launch<T, SrSumF<T, SRF_ABS>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    if (OP == SR_NRM2)  return scalar_reduce_launch<T, SrSumF<T, SRF_SQR>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    if (OP == SR_DOT)   return scalar_reduce_launch<T, SrSumF<T, SRF_DOT>, true>(n, X, offX, incX, Y, offY, incY, host_out);
    if (OP == SR_IMAX)  return scalar_reduce_launch<T, SrArg<T, true, false>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    if (OP == SR_IMIN)  return scalar_reduce_launch<T, SrArg<T, false, false>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    if (OP == SR_IAMAX) return scalar_reduce_launch<T, SrArg<T, true, true>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
    return scalar_reduce_launch<T, SrArg<T, false, true>, false>(n, X, offX, incX, nullptr, 0, 1, host_out);
}

template <int OP>
int scalar_reduce_dispatch(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* host_out,
                           const void* Y = nullptr, int64_t offY = 0, int64_t incY = 1)
{
    RB_REQUIRE_INIT();
    if (!X || !host_out || (OP == SR_DOT && !Y)) RB_INVALID("NULL buffer");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (incX < 1 || offX < 0 || incY < 1 || offY < 0) RB_INVALID("bad increment / offset");
    switch (dtype) {
        case RB200_FLOAT32: return scalar_reduce_op<float, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_FLOAT64: return scalar_reduce_op<double, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_INT32:   return scalar_reduce_op<int32_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_INT64:   return scalar_reduce_op<int64_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_BOOL: case RB200_UINT8: return scalar_reduce_op<uint8_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_INT8:    return scalar_reduce_op<int8_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_INT16:   return scalar_reduce_op<int16_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_UINT16:  return scalar_reduce_op<uint16_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        case RB200_UINT32:  return scalar_reduce_op<uint32_t, OP>(n, X, offX, incX, Y, offY, incY, host_out);
        default: RB_UNSUPPORTED("scalar reduction: unsupported dtype %d", dtype);
    }
}

// ---- updateAddOnehot ------------------------------------------------------------------------------------------
template <typename TX, typename TY>
__global__ void __launch_bounds__(UN_THREADS)
onehot_kernel(int64_t m, int64_t n, TY a, const TX* __restrict__ X, int64_t incX, TY* __restrict__ Y, int64_t ldY,
              int* __restrict__ bad)
{
    const int64_t stride = (int64_t)gridDim.x * UN_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * UN_THREADS + threadIdx.x; i < m; i += stride) {
        const int64_t label = (int64_t)X[i * incX];            // (int)$X[$idx], :1459
        if (label >= n || label < 0) { *bad = 1; continue; }
        TY* p = Y + i * ldY + label;
        *p = *p + a;
    }
}

template <typename TX, typename TY>
int onehot_launch(int64_t m, int64_t n, double a, const void* Xv, int64_t offX, int64_t incX, void* Yv, int64_t offY, int64_t ldY)
{
    rb200::Context& c = rb200::ctx();
    int rc = rb200::ensure_workspace(16);
    if (rc != RB200_OK) return rc;
    int* bad = reinterpret_cast<int*>(c.workspace);
    RB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c.stream));
    int64_t grid = rb_ceil_div(m, UN_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (grid > cap) grid = cap;
    onehot_kernel<TX, TY><<<(unsigned)grid, UN_THREADS, 0, c.stream>>>(
        m, n, (TY)a, reinterpret_cast<const TX*>(Xv) + offX, incX, reinterpret_cast<TY*>(Yv) + offY, ldY, bad);
    RB_LAUNCH_CHECK("onehot_kernel");
    RB_CUDA(cudaMemcpyAsync(c.pinned_scalar, bad, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    RB_CUDA(cudaStreamSynchronize(c.stream));
    int host_bad = 0;
    memcpy(&host_bad, c.pinned_scalar, sizeof(int));
    if (host_bad) {
        rb200::set_error("Label number is out of bounds.");
        return RB200_E_RANGE;
    }
    return RB200_OK;
}

template <typename TY>
int onehot_x(int dtype_x, int64_t m, int64_t n, double a, const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t ldY)
{
    switch (dtype_x) {
        case RB200_INT32:   return onehot_launch<int32_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_INT64:   return onehot_launch<int64_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_UINT8:   return onehot_launch<uint8_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_INT8:    return onehot_launch<int8_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_INT16:   return onehot_launch<int16_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_UINT16:  return onehot_launch<uint16_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_UINT32:  return onehot_launch<uint32_t, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_FLOAT32: return onehot_launch<float, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_FLOAT64: return onehot_launch<double, TY>(m, n, a, X, offX, incX, Y, offY, ldY);
        default: RB_UNSUPPORTED("onehot: unsupported label dtype %d", dtype_x);
    }
}

}  // namespace

extern "C" {

int rb200_unary(int dtype, int op, int64_t n, double alpha, void* X, int64_t offX, int64_t incX, double beta)
{
    RB_REQUIRE_INIT();
    if (!X) RB_INVALID("NULL buffer");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (incX < 1 || offX < 0) RB_INVALID("bad increment / offset");
    switch (dtype) {
        case RB200_FLOAT32: return unary_float_dispatch<float>(op, n, alpha, beta, X, offX, incX);
        case RB200_FLOAT64: return unary_float_dispatch<double>(op, n, alpha, beta, X, offX, incX);
        default: break;
    }
    if (op != RB200_UN_NOT) RB_UNSUPPORTED("unary op %d: unsupported dtype %d", op, dtype);
    switch (rb_dtype_size(dtype)) {
        case 1: return not_int_launch<uint8_t>(n, X, offX, incX);
        case 2: return not_int_launch<uint16_t>(n, X, offX, incX);
        case 4: return not_int_launch<uint32_t>(n, X, offX, incX);
        case 8: return not_int_launch<uint64_t>(n, X, offX, incX);
        default: RB_UNSUPPORTED("not: unsupported dtype %d", dtype);
    }
}

int rb200_compare(int dtype, int op, int64_t n, const void* X, int64_t offX, int64_t incX,
                  void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!X || !Y) RB_INVALID("NULL buffer");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (incX < 1 || incY < 1 || offX < 0 || offY < 0) RB_INVALID("bad increment / offset");
    if (dtype == RB200_FLOAT32) return compare_op<float>(op, n, X, offX, incX, Y, offY, incY);
    if (dtype == RB200_FLOAT64) return compare_op<double>(op, n, X, offX, incX, Y, offY, incY);
    switch (rb_dtype_size(dtype)) {
        case 1: return compare_op<uint8_t>(op, n, X, offX, incX, Y, offY, incY);
        case 2: return compare_op<uint16_t>(op, n, X, offX, incX, Y, offY, incY);
        case 4: return compare_op<uint32_t>(op, n, X, offX, incX, Y, offY, incY);
        case 8: return compare_op<uint64_t>(op, n, X, offX, incX, Y, offY, incY);
        default: RB_UNSUPPORTED("equal / notEqual: unsupported dtype %d", dtype);
    }
}

int rb200_astype(int64_t n, int from_dtype, const void* X, int64_t offX, int64_t incX,
                 int to_dtype, void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!X || !Y) RB_INVALID("NULL buffer");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (incX < 1 || incY < 1 || offX < 0 || offY < 0) RB_INVALID("bad increment / offset");
    switch (to_dtype) {
        case RB200_BOOL:    return astype_from<uint8_t>(from_dtype, true, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT8:   return astype_from<uint8_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_INT8:    return astype_from<int8_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_INT16:   return astype_from<int16_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT16:  return astype_from<uint16_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_INT32:   return astype_from<int32_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT32:  return astype_from<uint32_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_INT64:   return astype_from<int64_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_UINT64:  return astype_from<uint64_t>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_FLOAT32: return astype_from<float>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        case RB200_FLOAT64: return astype_from<double>(from_dtype, false, n, X, offX, incX, Y, offY, incY);
        default: RB_UNSUPPORTED("astype: dtype must be type of integer or float: %d", to_dtype);
    }
}

int rb200_sum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result)
{
    int64_t bits = 0;
    if (!result) RB_INVALID("result is NULL");
    const int rc = scalar_reduce_dispatch<SR_SUM>(dtype, n, X, offX, incX, &bits);
    if (rc == RB200_OK) memcpy(result, &bits, sizeof(double));
    return rc;
}

int rb200_imax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result)
{
    return scalar_reduce_dispatch<SR_IMAX>(dtype, n, X, offX, incX, result);
}

int rb200_imin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result)
{
    return scalar_reduce_dispatch<SR_IMIN>(dtype, n, X, offX, incX, result);
}


int rb200_asum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result)
{
    int64_t bits = 0;
    if (!result) RB_INVALID("result is NULL");
    const int rc = scalar_reduce_dispatch<SR_ASUM>(dtype, n, X, offX, incX, &bits);
    if (rc == RB200_OK) memcpy(result, &bits, sizeof(double));
    return rc;
}

int rb200_nrm2(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result)
{
    int64_t bits = 0;
    if (!result) RB_INVALID("result is NULL");
    const int rc = scalar_reduce_dispatch<SR_NRM2>(dtype, n, X, offX, incX, &bits);
    if (rc == RB200_OK) memcpy(result, &bits, sizeof(double));
    return rc;
}

int rb200_dot(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX,
              const void* Y, int64_t offY, int64_t incY, double* result)
{
    int64_t bits = 0;
    if (!result) RB_INVALID("result is NULL");
    const int rc = scalar_reduce_dispatch<SR_DOT>(dtype, n, X, offX, incX, &bits, Y, offY, incY);
    if (rc == RB200_OK) memcpy(result, &bits, sizeof(double));
    return rc;
}

int rb200_iamax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result)
{
    return scalar_reduce_dispatch<SR_IAMAX>(dtype, n, X, offX, incX, result);
}

int rb200_iamin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result)
{
    return scalar_reduce_dispatch<SR_IAMIN>(dtype, n, X, offX, incX, result);
}

int rb200_onehot(int dtype_x, int dtype_y, int64_t m, int64_t n, double a,
                 const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t ldY)
{
    RB_REQUIRE_INIT();
    if (!X || !Y) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1) RB_INVALID("m and n must be greater than 0");
    if (incX < 1 || ldY < 1 || offX < 0 || offY < 0) RB_INVALID("bad increment / leading dimension / offset");
    switch (dtype_y) {
        case RB200_FLOAT32: return onehot_x<float>(dtype_x, m, n, a, X, offX, incX, Y, offY, ldY);
        case RB200_FLOAT64: return onehot_x<double>(dtype_x, m, n, a, X, offX, incX, Y, offY, ldY);
        default: RB_UNSUPPORTED("onehot: unsupported output dtype %d", dtype_y);
    }
}

}  // extern "C"
