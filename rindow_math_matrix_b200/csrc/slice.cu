// slice.cu -- slice / stick (sub-block copy or add between A[m,n,k,size] and Y[s0,s1,s2,size]) and repeat
// (B[m,repeats,k] = A[m,k]): the data movement behind LinearAlgebra::slice / stick / stack / concat / split / repeat
// (HBM-bound; SURVEY.md section 8b lists them among the math methods LinearAlgebra calls).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php):
//   slice :2691-2776   for (i0,i1,i2) in the sub-block:  pa = ((i0+start0)*n*k + (i1+start1)*k + (i2+start2))*size + offA,
//                      py = ((i0*size1 + i1)*size2 + i2)*size + offY;  forward Y[py..] (=|+=) A[pa..], reverse A[pa..]
//                      (=|+=) Y[py..]; `size` items with increments incA / incY (math_copy / math_add :3197-3226).
//   repeat :1400-1422  B[(i*repeats + j)*k + e] = A[i*k + e].
// Copies are bit-exact for every dtype; add is one rounded add per element.
//
// The host folds axis 2 and every trailing axis that is taken whole into contiguous runs; long runs are copied by one
// CTA row per run in the widest word their alignment allows (16 bytes when everything is 16-byte aligned), short runs
// by one thread per word with multiply-high divisions.  Non-unit increments take a plain per-element kernel.
#include "common.cuh"

namespace {

constexpr int SL_THREADS = 256;

// After the host has folded axis 2 (always a contiguous run) and every trailing axis that is taken whole into the run:
// `runs` runs of `run` words; run r = (i0, i1) starts at start + i0*stride0 + i1*stride1 in A and at r*run in Y.
struct SliceParams {
    int64_t s1;                 // runs per i0
    int64_t stride0, stride1;   // source strides of i0 / i1, in words
    int64_t start;              // first source word
    int64_t run, runs;
    unsigned m_run, m_s1;       // floor(2^32/d)+1 for the 32-bit path of the flat kernel
    int small;                  // flat kernel: every dividend * divisor < 2^32
};

__device__ __forceinline__ unsigned sl_div(unsigned t, unsigned magic, unsigned d) { return d == 1 ? t : __umulhi(t, magic); }

// MODE 0: copy A -> Y, 1: copy Y -> A, 2: Y += A, 3: A += Y
template <typename W, int MODE>
__device__ __forceinline__ void sl_move(W* A, int64_t a, W* Y, int64_t y)
{
    if constexpr (MODE == 0) Y[y] = A[a];
    else if constexpr (MODE == 1) A[a] = Y[y];
    else if constexpr (MODE == 2) Y[y] = Y[y] + A[a];
    else A[a] = A[a] + Y[y];
}

// flat: one thread per word, run and position from the word index (short runs)
template <typename W, int MODE>
__global__ void __launch_bounds__(SL_THREADS)
slice_flat_kernel(SliceParams p, W* __restrict__ A, W* __restrict__ Y)
{
    const int64_t total = p.run * p.runs;
    for (int64_t t = (int64_t)blockIdx.x * SL_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * SL_THREADS) {
        int64_t e, i1, i0;
        if (p.small) {
            const unsigned tt = (unsigned)t;
            const unsigned r = sl_div(tt, p.m_run, (unsigned)p.run);
            e = tt - r * (unsigned)p.run;
            const unsigned q = sl_div(r, p.m_s1, (unsigned)p.s1);
            i1 = r - q * (unsigned)p.s1;
            i0 = q;
        } else {
            const int64_t r = t / p.run; e = t - r * p.run;
            i0 = r / p.s1; i1 = r - i0 * p.s1;
        }
        sl_move<W, MODE>(A, p.start + i0 * p.stride0 + i1 * p.stride1 + e, Y, t);
    }
}

// runs: a CTA takes chunks of SL_CHUNK consecutive words of one run (one division per chunk, not per word) and keeps
// SL_UNROLL independent loads in flight per thread before the first store
constexpr int SL_UNROLL = 4;
constexpr int SL_CHUNK = SL_THREADS * SL_UNROLL;

template <typename W, int MODE>
__global__ void __launch_bounds__(SL_THREADS)
slice_runs_kernel(SliceParams p, int64_t cpr, W* __restrict__ A, W* __restrict__ Y)
{
    const int64_t chunks = p.runs * cpr;
    for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const int64_t r = c / cpr, e0 = (c - r * cpr) * SL_CHUNK + threadIdx.x;
        const int64_t i0 = r / p.s1, i1 = r - i0 * p.s1;
        W* a = A + p.start + i0 * p.stride0 + i1 * p.stride1;
        W* y = Y + r * p.run;
        W va[SL_UNROLL], vy[SL_UNROLL];
#pragma unroll
        for (int u = 0; u < SL_UNROLL; u++) {
            const int64_t e = e0 + u * SL_THREADS;
            if (e < p.run) {
                if (MODE != 1) va[u] = a[e];
                if (MODE != 0) vy[u] = y[e];
            }
        }
#pragma unroll
        for (int u = 0; u < SL_UNROLL; u++) {
            const int64_t e = e0 + u * SL_THREADS;
            if (e < p.run) {
                if constexpr (MODE == 0) y[e] = va[u];
                else if constexpr (MODE == 1) a[e] = vy[u];
                else if constexpr (MODE == 2) y[e] = vy[u] + va[u];
                else a[e] = va[u] + vy[u];
            }
        }
    }
}

// non-unit increments (the reference applies them inside an item only: pa + e*incA, py + e*incY, :2748-2770)
struct SliceStrided {
    int64_t n, k, size, s1, s2, st0, st1, st2, incA, incY, total;
};
template <typename T, int MODE>
__global__ void __launch_bounds__(SL_THREADS)
slice_strided_kernel(SliceStrided p, T* __restrict__ A, T* __restrict__ Y)
{
    for (int64_t t = (int64_t)blockIdx.x * SL_THREADS + threadIdx.x; t < p.total; t += (int64_t)gridDim.x * SL_THREADS) {
        const int64_t q = t / p.size, e = t - q * p.size;
        const int64_t q2 = q / p.s2, i2 = q - q2 * p.s2;
        const int64_t i0 = q2 / p.s1, i1 = q2 - i0 * p.s1;
        const int64_t pa = (((i0 + p.st0) * p.n + (i1 + p.st1)) * p.k + (i2 + p.st2)) * p.size + e * p.incA;
        const int64_t py = q * p.size + e * p.incY;
        sl_move<T, MODE>(A, pa, Y, py);
    }
}

static unsigned sl_magic(int64_t d) { return (unsigned)((1ULL << 32) / (unsigned long long)d) + 1u; }

template <typename W, int MODE>
int slice_launch(const SliceParams& p, void* A, void* Y)
{
    rb200::Context& c = rb200::ctx();
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (p.run >= 128) {
        const int64_t cpr = rb_ceil_div(p.run, SL_CHUNK);
        int64_t grid = p.runs * cpr;
        if (grid > cap) grid = cap;
        slice_runs_kernel<W, MODE><<<(unsigned)grid, SL_THREADS, 0, c.stream>>>(p, cpr, reinterpret_cast<W*>(A), reinterpret_cast<W*>(Y));
        RB_LAUNCH_CHECK("slice_runs_kernel");
        return RB200_OK;
    }
    int64_t grid = rb_ceil_div(p.run * p.runs, SL_THREADS);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    slice_flat_kernel<W, MODE><<<(unsigned)grid, SL_THREADS, 0, c.stream>>>(p, reinterpret_cast<W*>(A), reinterpret_cast<W*>(Y));
    RB_LAUNCH_CHECK("slice_flat_kernel");
    return RB200_OK;
}

template <typename T, int MODE>
int slice_strided_launch(const SliceStrided& p, void* A, void* Y)
{
    rb200::Context& c = rb200::ctx();
    int64_t grid = rb_ceil_div(p.total, SL_THREADS);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    slice_strided_kernel<T, MODE><<<(unsigned)grid, SL_THREADS, 0, c.stream>>>(p, reinterpret_cast<T*>(A), reinterpret_cast<T*>(Y));
    RB_LAUNCH_CHECK("slice_strided_kernel");
    return RB200_OK;
}

template <typename T>
int slice_strided(bool reverse, bool add, const SliceStrided& p, void* A, void* Y)
{
    if (!add) return reverse ? slice_strided_launch<T, 1>(p, A, Y) : slice_strided_launch<T, 0>(p, A, Y);
    return reverse ? slice_strided_launch<T, 3>(p, A, Y) : slice_strided_launch<T, 2>(p, A, Y);
}

template <typename W>
int slice_copy(bool reverse, const SliceParams& p, void* A, void* Y)
{
    return reverse ? slice_launch<W, 1>(p, A, Y) : slice_launch<W, 0>(p, A, Y);
}

template <typename T>
int slice_add(bool reverse, const SliceParams& p, void* A, void* Y)
{
    return reverse ? slice_launch<T, 3>(p, A, Y) : slice_launch<T, 2>(p, A, Y);
}

// repeat: a CTA takes a chunk of one row of A, loads it once and stores it `repeats` times
template <typename W>
__global__ void __launch_bounds__(SL_THREADS)
repeat_kernel(int64_t m, int64_t k, int64_t repeats, int64_t cpr, const W* __restrict__ A, W* __restrict__ B)
{
    const int64_t chunks = m * cpr;
    for (int64_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const int64_t i = c / cpr, e0 = (c - i * cpr) * SL_CHUNK + threadIdx.x;
        const W* a = A + i * k;
        W v[SL_UNROLL];
#pragma unroll
        for (int u = 0; u < SL_UNROLL; u++) {
            const int64_t e = e0 + u * SL_THREADS;
            if (e < k) v[u] = __ldg(a + e);
        }
        for (int64_t j = 0; j < repeats; j++) {
            W* b = B + (i * repeats + j) * k;
#pragma unroll
            for (int u = 0; u < SL_UNROLL; u++) {
                const int64_t e = e0 + u * SL_THREADS;
                if (e < k) __stcs(b + e, v[u]);
            }
        }
    }
}

template <typename W>
int repeat_launch(int64_t m, int64_t k, int64_t repeats, const void* A, void* B)
{
    rb200::Context& c = rb200::ctx();
    const int64_t cpr = rb_ceil_div(k, SL_CHUNK);
    int64_t grid = m * cpr;
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    repeat_kernel<W><<<(unsigned)grid, SL_THREADS, 0, c.stream>>>(m, k, repeats, cpr, reinterpret_cast<const W*>(A),
                                                                 reinterpret_cast<W*>(B));
    RB_LAUNCH_CHECK("repeat_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_slice(int dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int64_t size,
                           void* A, int64_t offA, int64_t incA, void* Y, int64_t offY, int64_t incY,
                           int64_t start0, int64_t size0, int64_t start1, int64_t size1, int64_t start2, int64_t size2)
{
    RB_REQUIRE_INIT();
    if (!A || !Y) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || size < 1 || incA < 1 || incY < 1 || offA < 0 || offY < 0) RB_INVALID("slice: bad shape / increment / offset");
    if (start0 < 0 || start0 >= m || size0 < 0 || size0 + start0 > m) RB_INVALID("Axis0 range is too large for source array.");
    if (start1 < 0 || start1 >= n || size1 < 0 || size1 + start1 > n) RB_INVALID("Axis1 range is too large for source array.");
    if (start2 < 0 || start2 >= k || size2 < 0 || size2 + start2 > k) RB_INVALID("Axis2 range is too large for source array.");
    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("slice: unsupported dtype %d", dtype);
    if (size0 == 0 || size1 == 0 || size2 == 0) return RB200_OK;
    char* a = reinterpret_cast<char*>(A) + offA * es;
    char* y = reinterpret_cast<char*>(Y) + offY * es;
    if (incA != 1 || incY != 1) {
        SliceStrided q;
        q.n = n; q.k = k; q.size = size; q.s1 = size1; q.s2 = size2; q.st0 = start0; q.st1 = start1; q.st2 = start2;
        q.incA = incA; q.incY = incY; q.total = size0 * size1 * size2 * size;
        const bool rev = reverse != 0, add = add_mode != 0;
        switch (dtype) {
            case RB200_FLOAT32: return slice_strided<float>(rev, add, q, a, y);
            case RB200_FLOAT64: return slice_strided<double>(rev, add, q, a, y);
            case RB200_INT32:   return slice_strided<int32_t>(rev, add, q, a, y);
            case RB200_INT64:   return slice_strided<int64_t>(rev, add, q, a, y);
            case RB200_INT8:    return slice_strided<int8_t>(rev, add, q, a, y);
            case RB200_BOOL: case RB200_UINT8: return slice_strided<uint8_t>(rev, add, q, a, y);
            case RB200_INT16:   return slice_strided<int16_t>(rev, add, q, a, y);
            case RB200_UINT16:  return slice_strided<uint16_t>(rev, add, q, a, y);
            case RB200_UINT32:  return slice_strided<uint32_t>(rev, add, q, a, y);
            case RB200_UINT64:  return slice_strided<uint64_t>(rev, add, q, a, y);
            default: RB_UNSUPPORTED("slice: unsupported dtype %d", dtype);
        }
    }
    // contiguous runs: axis 2 always folds into the run; axis 1 too when axis 2 is whole, then axis 0 when axis 1 is
    int64_t run = size2 * size, s0 = size0, s1 = size1;
    int64_t stride1 = k * size, stride0 = n * k * size;
    int64_t first = ((start0 * n + start1) * k + start2) * size;
    if (size2 == k) {
        run *= s1; s1 = 1;
        if (size1 == n) { run *= s0; s0 = 1; }
    }
    // widest word the runs allow (copy only): run bytes, strides, first word and both bases 16 / 8 / 4 ... aligned
    int wb = es;
    if (!add_mode) {
        for (int cand = 16; cand > es; cand >>= 1) {
            if ((run * es) % cand == 0 && (stride0 * es) % cand == 0 && (stride1 * es) % cand == 0 && (first * es) % cand == 0 &&
                (reinterpret_cast<uintptr_t>(a) % cand) == 0 && (reinterpret_cast<uintptr_t>(y) % cand) == 0) {
                wb = cand;
                break;
            }
        }
    }
    const int64_t f = wb / es;
    SliceParams p;
    p.s1 = s1; p.run = run / f; p.runs = s0 * s1; p.stride0 = stride0 / f; p.stride1 = stride1 / f; p.start = first / f;
    const int64_t dmax = p.run > s1 ? p.run : s1;
    p.small = dmax < (1LL << 16) && p.run * p.runs < (1LL << 32) / dmax;
    p.m_run = sl_magic(p.run); p.m_s1 = sl_magic(s1);
    if (!add_mode) {
        switch (wb) {
            case 1:  return slice_copy<uint8_t>(reverse != 0, p, a, y);
            case 2:  return slice_copy<uint16_t>(reverse != 0, p, a, y);
            case 4:  return slice_copy<uint32_t>(reverse != 0, p, a, y);
            case 8:  return slice_copy<uint64_t>(reverse != 0, p, a, y);
            default: return slice_copy<int4>(reverse != 0, p, a, y);
        }
    }
    switch (dtype) {
        case RB200_FLOAT32: return slice_add<float>(reverse != 0, p, a, y);
        case RB200_FLOAT64: return slice_add<double>(reverse != 0, p, a, y);
        case RB200_INT32:   return slice_add<int32_t>(reverse != 0, p, a, y);
        case RB200_INT64:   return slice_add<int64_t>(reverse != 0, p, a, y);
        case RB200_INT8:    return slice_add<int8_t>(reverse != 0, p, a, y);
        case RB200_BOOL: case RB200_UINT8: return slice_add<uint8_t>(reverse != 0, p, a, y);
        case RB200_INT16:   return slice_add<int16_t>(reverse != 0, p, a, y);
        case RB200_UINT16:  return slice_add<uint16_t>(reverse != 0, p, a, y);
        case RB200_UINT32:  return slice_add<uint32_t>(reverse != 0, p, a, y);
        case RB200_UINT64:  return slice_add<uint64_t>(reverse != 0, p, a, y);
        default: RB_UNSUPPORTED("slice (add mode): unsupported dtype %d", dtype);
    }
}

extern "C" int rb200_repeat(int dtype, int64_t m, int64_t k, int64_t repeats, const void* A, int64_t offA, void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !B) RB_INVALID("NULL buffer");
    if (m < 1 || k < 1 || repeats < 1 || offA < 0 || offB < 0) RB_INVALID("repeat: bad shape / offset");
    const int es = rb_dtype_size(dtype);
    if (es == 0) RB_UNSUPPORTED("repeat: unsupported dtype %d", dtype);
    const char* a = reinterpret_cast<const char*>(A) + offA * es;
    char* b = reinterpret_cast<char*>(B) + offB * es;
    int wb = es;
    for (int cand = 16; cand > es; cand >>= 1) {
        if ((k * es) % cand == 0 && (reinterpret_cast<uintptr_t>(a) % cand) == 0 && (reinterpret_cast<uintptr_t>(b) % cand) == 0) { wb = cand; break; }
    }
    const int64_t kw = k * es / wb;
    switch (wb) {
        case 1:  return repeat_launch<uint8_t>(m, kw, repeats, a, b);
        case 2:  return repeat_launch<uint16_t>(m, kw, repeats, a, b);
        case 4:  return repeat_launch<uint32_t>(m, kw, repeats, a, b);
        case 8:  return repeat_launch<uint64_t>(m, kw, repeats, a, b);
        default: return repeat_launch<int4>(m, kw, repeats, a, b);
    }
}
