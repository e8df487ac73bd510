// gather.cu -- gather / scatter / scatterAdd along the first axis and reduceGather (SURVEY.md section 8f rank 4:
// embedding lookup and its gradient, sparse cross-entropy pick).  HBM-bound row movement.
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php):
//   gather :1063-1130   for j < n: index = X[j] (index >= numClass is clamped to numClass-1, :1107-1111)
//        forward copy   B[j,:] = A[index,:]          forward add   B[j,:] += A[index,:]
//        reverse copy   A[index,:] = B[j,:]  -- sequential in j: for duplicate indices the LAST j wins
//        reverse add    scatterAdd2 :1362-1395: A[index,t] += B[j,t], j ascending for every column t
//   reduceGather :1135-1208   B[i,j] = A[i, X[i,j], j]  (and the reverse copy / add forms; no two (i,j) share a target)
// Determinism: the reverse forms reproduce the sequential results exactly -- "last j wins" through a winner table
// (atomicMax of j per class), scatterAdd by grouping the (class, j) pairs with a stable radix sort (cub::DeviceRadixSort
// -- index plumbing only; the arithmetic below is this file's) and summing every class's rows in ascending j on top
// of the existing A, the same association as the reference loop.  A negative index (which the reference would turn
// into an out-of-range SplFixedArray access) is clamped to 0.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace {

constexpr int GA_THREADS = 256;

template <typename I>
__device__ __forceinline__ int64_t ga_index(const I* X, int64_t j, int64_t numClass)
{
    int64_t idx = (int64_t)X[j];
    if (idx >= numClass) idx = numClass - 1;
    if (idx < 0) idx = 0;
    return idx;
}

template <typename T, int V> struct alignas(sizeof(T) * V) GaPack { T v[V]; };

// MODE 0: B[j] = A[idx]   1: B[j] += A[idx]   2: A[idx] = B[j] iff winner[idx] == j
template <typename T, typename I, int V, int MODE>
__global__ void __launch_bounds__(GA_THREADS)
gather_rows_kernel(int64_t n, int64_t kv, int64_t numClass, const I* __restrict__ X, T* __restrict__ A, T* __restrict__ B,
                   const int* __restrict__ winner)
{
    using Pack = GaPack<T, V>;
    const int64_t total = n * kv;
    const int64_t stride = (int64_t)gridDim.x * GA_THREADS;
    for (int64_t g = (int64_t)blockIdx.x * GA_THREADS + threadIdx.x; g < total; g += stride) {
        const int64_t j = g / kv, piece = g - j * kv;
        const int64_t idx = ga_index<I>(X, j, numClass);
        Pack* a = reinterpret_cast<Pack*>(A + idx * kv * V) + piece;
        Pack* b = reinterpret_cast<Pack*>(B + j * kv * V) + piece;
        if (MODE == 0) {
            *b = *a;
        } else if (MODE == 1) {
            Pack pa = *a, pb = *b;
#pragma unroll
            for (int e = 0; e < V; e++) pb.v[e] = pb.v[e] + pa.v[e];
            *b = pb;
        } else {
            if (winner[idx] == (int)j) *a = *b;
        }
    }
}

template <typename I>
__global__ void __launch_bounds__(GA_THREADS)
scatter_winner_kernel(int64_t n, int64_t numClass, const I* __restrict__ X, int* __restrict__ winner)
{
    const int64_t stride = (int64_t)gridDim.x * GA_THREADS;
    for (int64_t j = (int64_t)blockIdx.x * GA_THREADS + threadIdx.x; j < n; j += stride) {
        atomicMax(&winner[ga_index<I>(X, j, numClass)], (int)j);
    }
}

template <typename I>
__global__ void __launch_bounds__(GA_THREADS)
scatter_keys_kernel(int64_t n, int64_t numClass, const I* __restrict__ X, int* __restrict__ keys, int* __restrict__ vals)
{
    const int64_t stride = (int64_t)gridDim.x * GA_THREADS;
    for (int64_t j = (int64_t)blockIdx.x * GA_THREADS + threadIdx.x; j < n; j += stride) {
        keys[j] = (int)ga_index<I>(X, j, numClass);
        vals[j] = (int)j;
    }
}

// one warp per sorted position; the warp that sits on the first entry of a class sums the class's rows in
// ascending j (the stable sort kept them in order) on top of A, TPL columns per lane at a time
template <typename T>
__global__ void __launch_bounds__(GA_THREADS)
scatter_add_sorted_kernel(int64_t n, int64_t k, const int* __restrict__ keys, const int* __restrict__ js,
                          T* __restrict__ A, const T* __restrict__ B)
{
    constexpr int TPL = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (GA_THREADS / 32);
    for (int64_t q = (int64_t)blockIdx.x * (GA_THREADS / 32) + (threadIdx.x >> 5); q < n; q += warps) {
        const int cls = keys[q];
        if (q > 0 && keys[q - 1] == cls) continue;                 // not the head of its class
        int64_t end = q + 1;
        while (end < n && keys[end] == cls) end++;
        T* arow = A + (int64_t)cls * k;
        for (int64_t t0 = 0; t0 < k; t0 += 32 * TPL) {
            T acc[TPL];
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const int64_t t = t0 + u * 32 + lane;
                acc[u] = t < k ? arow[t] : (T)0;
            }
            for (int64_t e = q; e < end; e++) {
                const T* brow = B + (int64_t)js[e] * k;
#pragma unroll
                for (int u = 0; u < TPL; u++) {
                    const int64_t t = t0 + u * 32 + lane;
                    if (t < k) acc[u] = acc[u] + brow[t];
                }
            }
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const int64_t t = t0 + u * 32 + lane;
                if (t < k) arow[t] = acc[u];
            }
        }
    }
}

int64_t ga_grid(int64_t work)
{
    int64_t grid = rb_ceil_div(work, GA_THREADS);
    const int64_t cap = (int64_t)rb200::ctx().sm_count * 32;
    if (grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}

template <typename T, typename I, int V>
int gather_launch_v(int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass, const I* X, T* A, T* B)
{
    rb200::Context& c = rb200::ctx();
    const int64_t kv = k / V;
    const unsigned grid = (unsigned)ga_grid(n * kv);
    if (!reverse) {
        if (!add_mode) gather_rows_kernel<T, I, V, 0><<<grid, GA_THREADS, 0, c.stream>>>(n, kv, numClass, X, A, B, nullptr);
        else           gather_rows_kernel<T, I, V, 1><<<grid, GA_THREADS, 0, c.stream>>>(n, kv, numClass, X, A, B, nullptr);
        RB_LAUNCH_CHECK("gather_rows_kernel");
        return RB200_OK;
    }
    // reverse copy: winner table in the workspace
    int rc = rb200::ensure_workspace((size_t)numClass * sizeof(int));
    if (rc != RB200_OK) return rc;
    int* winner = reinterpret_cast<int*>(c.workspace);
    RB_CUDA(cudaMemsetAsync(winner, 0xff, (size_t)numClass * sizeof(int), c.stream));      // -1
    scatter_winner_kernel<I><<<(unsigned)ga_grid(n), GA_THREADS, 0, c.stream>>>(n, numClass, X, winner);
    RB_LAUNCH_CHECK("scatter_winner_kernel");
    gather_rows_kernel<T, I, V, 2><<<grid, GA_THREADS, 0, c.stream>>>(n, kv, numClass, X, A, B, winner);
    RB_LAUNCH_CHECK("gather_rows_kernel");
    return RB200_OK;
}

template <typename T, typename I>
int scatter_add_launch(int64_t n, int64_t k, int64_t numClass, const I* X, T* A, const T* B)
{
    rb200::Context& c = rb200::ctx();
    int bits = 1;
    while (bits < 31 && (1LL << bits) < numClass) bits++;
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                                    (int*)nullptr, (int)n, 0, bits, c.stream);
    if (e != cudaSuccess) return rb200::cuda_fail(e, "cub::DeviceRadixSort::SortPairs (size)", __FILE__, __LINE__);
    const size_t arr = (((size_t)n * sizeof(int)) + 255) & ~(size_t)255;
    int rc = rb200::ensure_workspace(4 * arr + cub_bytes + 256);
    if (rc != RB200_OK) return rc;
    char* w = reinterpret_cast<char*>(c.workspace);
    int* keys_in = reinterpret_cast<int*>(w), *vals_in = reinterpret_cast<int*>(w + arr);
    int* keys_out = reinterpret_cast<int*>(w + 2 * arr), *vals_out = reinterpret_cast<int*>(w + 3 * arr);
    void* cub_tmp = w + 4 * arr;
    scatter_keys_kernel<I><<<(unsigned)ga_grid(n), GA_THREADS, 0, c.stream>>>(n, numClass, X, keys_in, vals_in);
    RB_LAUNCH_CHECK("scatter_keys_kernel");
    e = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, bits, c.stream);
    if (e != cudaSuccess) return rb200::cuda_fail(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
    rb200::count_launch(1);
    scatter_add_sorted_kernel<T><<<(unsigned)ga_grid(n * 32), GA_THREADS, 0, c.stream>>>(n, k, keys_out, vals_out, A, B);
    RB_LAUNCH_CHECK("scatter_add_sorted_kernel");
    return RB200_OK;
}

template <typename T, typename I>
int gather_launch(int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass,
                  const void* Xv, int64_t offX, void* Av, int64_t offA, void* Bv, int64_t offB)
{
    const I* X = reinterpret_cast<const I*>(Xv) + offX;
    T* A = reinterpret_cast<T*>(Av) + offA;
    T* B = reinterpret_cast<T*>(Bv) + offB;
    if (reverse && add_mode) {
        if (n >= (1LL << 31) || numClass >= (1LL << 31)) RB_UNSUPPORTED("scatterAdd: n and numClass must be below 2^31");
        return scatter_add_launch<T, I>(n, k, numClass, X, A, B);
    }
    if (reverse && n >= (1LL << 31)) RB_UNSUPPORTED("scatter: n must be below 2^31");
    constexpr int VMAX = 16 / (int)sizeof(T);
    const bool vec = VMAX > 1 && (k % VMAX == 0) &&
                     ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
    if (vec) return gather_launch_v<T, I, VMAX>(reverse, add_mode, n, k, numClass, X, A, B);
    return gather_launch_v<T, I, 1>(reverse, add_mode, n, k, numClass, X, A, B);
}

// ---- reduceGather: B[i,j] (op)= A[i, X[i,j], j] and the reverse forms; one thread per (i,j) ---------------------------
template <typename T, typename I, int MODE>       // 0 fwd copy, 1 fwd add, 2 rev copy, 3 rev add
__global__ void __launch_bounds__(GA_THREADS)
reduce_gather_kernel(int64_t m, int64_t n, int64_t numClass, const I* __restrict__ X, T* __restrict__ A, T* __restrict__ B)
{
    const int64_t total = m * n;
    const int64_t stride = (int64_t)gridDim.x * GA_THREADS;
    for (int64_t g = (int64_t)blockIdx.x * GA_THREADS + threadIdx.x; g < total; g += stride) {
        const int64_t i = g / n, j = g - i * n;
        const int64_t idx = ga_index<I>(X, g, numClass);
        T* a = A + (i * numClass + idx) * n + j;
        T* b = B + g;
        if (MODE == 0) *b = *a;
        else if (MODE == 1) *b = *b + *a;
        else if (MODE == 2) *a = *b;
        else *a = *a + *b;
    }
}

template <typename T, typename I>
int reduce_gather_launch(int reverse, int add_mode, int64_t m, int64_t n, int64_t numClass,
                         const void* Xv, int64_t offX, void* Av, int64_t offA, void* Bv, int64_t offB)
{
    rb200::Context& c = rb200::ctx();
    const I* X = reinterpret_cast<const I*>(Xv) + offX;
    T* A = reinterpret_cast<T*>(Av) + offA;
    T* B = reinterpret_cast<T*>(Bv) + offB;
    const unsigned grid = (unsigned)ga_grid(m * n);
    const int mode = (reverse ? 2 : 0) + (add_mode ? 1 : 0);
    switch (mode) {
        case 0: reduce_gather_kernel<T, I, 0><<<grid, GA_THREADS, 0, c.stream>>>(m, n, numClass, X, A, B); break;
        case 1: reduce_gather_kernel<T, I, 1><<<grid, GA_THREADS, 0, c.stream>>>(m, n, numClass, X, A, B); break;
        case 2: reduce_gather_kernel<T, I, 2><<<grid, GA_THREADS, 0, c.stream>>>(m, n, numClass, X, A, B); break;
        default: reduce_gather_kernel<T, I, 3><<<grid, GA_THREADS, 0, c.stream>>>(m, n, numClass, X, A, B); break;
    }
    RB_LAUNCH_CHECK("reduce_gather_kernel");
    return RB200_OK;
}

template <typename T, bool REDUCE>
int ga_by_index(int index_dtype, int reverse, int add_mode, int64_t p0, int64_t p1, int64_t numClass,
                const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB)
{
#define GA_CASE(code, I) case code: return REDUCE ? reduce_gather_launch<T, I>(reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB) \
                                                   : gather_launch<T, I>(reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
    switch (index_dtype) {
        GA_CASE(RB200_INT32, int32_t) GA_CASE(RB200_INT64, int64_t) GA_CASE(RB200_UINT8, uint8_t) GA_CASE(RB200_INT8, int8_t)
        GA_CASE(RB200_INT16, int16_t) GA_CASE(RB200_UINT16, uint16_t) GA_CASE(RB200_UINT32, uint32_t)
        GA_CASE(RB200_FLOAT32, float) GA_CASE(RB200_FLOAT64, double)
        default: RB_UNSUPPORTED("gather: unsupported index dtype %d", index_dtype);
    }
#undef GA_CASE
}

template <bool REDUCE>
int ga_dispatch(int dtype, int index_dtype, int reverse, int add_mode, int64_t p0, int64_t p1, int64_t numClass,
                const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!X || !A || !B) RB_INVALID("NULL buffer");
    if (p0 < 1 || p1 < 1) RB_INVALID("sizes must be greater than 0");
    if (numClass <= 0) RB_INVALID("numClass must be grator than zero.");
    if (offX < 0 || offA < 0 || offB < 0) RB_INVALID("bad offset");
    switch (dtype) {
        case RB200_FLOAT32: return ga_by_index<float, REDUCE>(index_dtype, reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
        case RB200_FLOAT64: return ga_by_index<double, REDUCE>(index_dtype, reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
        case RB200_INT32:   return ga_by_index<int32_t, REDUCE>(index_dtype, reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
        case RB200_INT64:   return ga_by_index<int64_t, REDUCE>(index_dtype, reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
        case RB200_BOOL: case RB200_UINT8:
            return ga_by_index<uint8_t, REDUCE>(index_dtype, reverse, add_mode, p0, p1, numClass, X, offX, A, offA, B, offB);
        default: RB_UNSUPPORTED("gather: unsupported dtype %d", dtype);
    }
}

}  // namespace

extern "C" {

int rb200_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass,
                 const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB)
{
    return ga_dispatch<false>(dtype, index_dtype, reverse, add_mode, n, k, numClass, X, offX, A, offA, B, offB);
}

int rb200_reduce_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t numClass,
                        const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB)
{
    return ga_dispatch<true>(dtype, index_dtype, reverse, add_mode, m, n, numClass, X, offX, A, offA, B, offB);
}

}  // extern "C"
