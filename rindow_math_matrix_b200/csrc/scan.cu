// scan.cu -- prefix sums and sorted search.
//
//   rb200_cumsumb      B[m,n,k] = running sum of A[m,n,k] along n           (PhpMath.php:1861-1904, LinearAlgebra::cumsum)
//   rb200_cumsum       Y[n]     = running sum of X[n] with increments       (PhpMath.php:1822-1859)
//   rb200_searchsorted Y[i]     = number of leading A[row(i), j] below X[i] (PhpMath.php:1780-1820)
//
// HBM-bound.  A scan along n is split into chunks only when (m, k) alone cannot fill the machine: pass 1 writes one
// total per chunk, the totals are scanned, pass 2 re-reads the chunk behind its carry (2 reads + 1 write, every
// sum in a fixed order -- no look-back whose association would depend on timing); otherwise one pass (1 read +
// 1 write).  k == 1 scans run a warp per (row, chunk) over 256-element tiles (coalesced striped
// loads, transposed through shared memory to 8 consecutive elements per lane); k >= 2 scans run a thread per (i, h)
// column, coalesced across h.  The reference accumulates serially; these add in tree order -- the tests hold them
// to the rounding bound of the dtype, and integer dtypes are exact.
#include "common.cuh"

namespace {

constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr int SC_VPT = 8;
constexpr int SC_TILE = 32 * SC_VPT;            // elements a warp scans per step
constexpr int SC_TILE_PAD = SC_TILE + SC_TILE / 32;

struct ScanRows {
    int64_t m, n;            // rows, elements per row
    int64_t ldx, ldy;        // row strides (elements)
    int64_t incx, incy;
    int64_t chunk;           // logical steps per chunk (multiple of SC_TILE)
    int64_t chunks;          // chunks per row
    int revx, revy, exclusive;
};

template <typename T, int V> struct alignas(sizeof(T) * V) ScVec { T e[V]; };

template <typename T> __device__ __forceinline__ T sc_shfl_up(T v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
template <typename T> __device__ __forceinline__ T sc_shfl(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

// PHASE 0: chunk totals only.  PHASE 1: scan behind the totals of the earlier chunks (totals == nullptr: one chunk).
// PHASE 1 reads the carry of a chunk from `totals`, which the host has turned into exclusive prefixes by then
// (scan_totals_kernel).  The next tile's loads are issued before the current tile is scanned.
template <typename T, int PHASE>
__global__ void __launch_bounds__(SC_THREADS)
scan_rows_kernel(ScanRows p, const T* __restrict__ X, T* __restrict__ Y, T* __restrict__ totals)
{
    __shared__ T tile_s[SC_WARPS][SC_TILE_PAD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t units = p.m * p.chunks;
    T* ts = tile_s[warp];
    for (int64_t uu = (int64_t)blockIdx.x * SC_WARPS + warp; uu < units; uu += (int64_t)gridDim.x * SC_WARPS) {
        // pass 2 starts with the chunks pass 1 read last: they are still in L2
        const int64_t u = (PHASE == 1 && totals != nullptr) ? units - 1 - uu : uu;
        const int64_t i = u / p.chunks, c = u - i * p.chunks;
        const int64_t s0 = c * p.chunk;
        const int64_t s1 = s0 + p.chunk < p.n ? s0 + p.chunk : p.n;
        const T* x = X + i * p.ldx;
        T* y = Y + i * p.ldy;
        if (PHASE == 0) {
            // totals: any order of addition, no transposition
            T part = (T)0;
            for (int64_t t0 = s0; t0 < s1; t0 += 2 * SC_TILE) {
                T v[2 * SC_VPT];
#pragma unroll
                for (int r = 0; r < 2 * SC_VPT; r++) {
                    const int64_t s = t0 + r * 32 + lane;
                    v[r] = s < s1 ? x[(p.revx ? p.n - 1 - s : s) * p.incx] : (T)0;
                }
#pragma unroll
                for (int r = 0; r < 2 * SC_VPT; r++) part += v[r];
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            if (lane == 0) totals[i * p.chunks + c] = part;
            continue;
        }
        T carry = totals != nullptr ? totals[i * p.chunks + c] : (T)0;
        T v[SC_VPT];
#pragma unroll
        for (int r = 0; r < SC_VPT; r++) {
            const int64_t s = s0 + r * 32 + lane;
            v[r] = s < s1 ? x[(p.revx ? p.n - 1 - s : s) * p.incx] : (T)0;
        }
        for (int64_t t0 = s0; t0 < s1; t0 += SC_TILE) {
#pragma unroll
            for (int r = 0; r < SC_VPT; r++) { const int e = r * 32 + lane; ts[e + (e >> 5)] = v[r]; }
            __syncwarp();
            if (t0 + SC_TILE < s1) {
#pragma unroll
                for (int r = 0; r < SC_VPT; r++) {
                    const int64_t s = t0 + SC_TILE + r * 32 + lane;
                    v[r] = s < s1 ? x[(p.revx ? p.n - 1 - s : s) * p.incx] : (T)0;
                }
            }
            T a[SC_VPT], tot = (T)0;
#pragma unroll
            for (int q = 0; q < SC_VPT; q++) { const int e = lane * SC_VPT + q; a[q] = ts[e + (e >> 5)]; tot += a[q]; }
            // inclusive scan of the lane totals
            T inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const T o = sc_shfl_up(inc, d); if (lane >= d) inc += o; }
            // a lane starts behind the carry and the inclusive total of the lane before it
            const T prev = sc_shfl_up(inc, 1);
            T run = lane == 0 ? carry : carry + prev;
            const T tile_total = sc_shfl(inc, 31);
#pragma unroll
            for (int q = 0; q < SC_VPT; q++) {
                T o;
                if (p.exclusive) { o = run; run += a[q]; } else { run += a[q]; o = run; }
                const int e = lane * SC_VPT + q;
                ts[e + (e >> 5)] = o;
            }
            carry += tile_total;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < SC_VPT; r++) {
                const int64_t s = t0 + r * 32 + lane;
                const int e = r * 32 + lane;
                if (s < s1) y[(p.revy ? p.n - 1 - s : s) * p.incy] = ts[e + (e >> 5)];
            }
            __syncwarp();
        }
    }
}

// Forward scans of contiguous, 16-byte aligned rows (the common case) skip shared memory: a lane loads SC_NV 16-byte
// vectors, vector j of lane l covering elements (j * 32 + l) * V ... + V of the tile -- fully coalesced -- and the
// tile is scanned as SC_NV sub-tiles of 32 * V elements, one shuffle scan each (independent, so they pipeline).
constexpr int SC_NV = 4;

template <typename T, int PHASE>
__global__ void __launch_bounds__(SC_THREADS)
scan_rows_vec_kernel(ScanRows p, const T* __restrict__ X, T* __restrict__ Y, T* __restrict__ totals)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int TILE = 32 * V * SC_NV;
    typedef ScVec<T, V> Vec;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t units = p.m * p.chunks;
    for (int64_t uu = (int64_t)blockIdx.x * SC_WARPS + warp; uu < units; uu += (int64_t)gridDim.x * SC_WARPS) {
        // pass 2 starts with the chunks pass 1 read last: they are still in L2
        const int64_t u = (PHASE == 1 && totals != nullptr) ? units - 1 - uu : uu;
        const int64_t i = u / p.chunks, c = u - i * p.chunks;
        const int64_t s0 = c * p.chunk;
        const int64_t s1 = s0 + p.chunk < p.n ? s0 + p.chunk : p.n;
        const T* x = X + i * p.ldx;
        T* y = Y + i * p.ldy;
        Vec zero;
#pragma unroll
        for (int w = 0; w < V; w++) zero.e[w] = (T)0;
        if (PHASE == 0) {
            T part = (T)0;
            for (int64_t t0 = s0; t0 < s1; t0 += TILE) {
                Vec v[SC_NV];
#pragma unroll
                for (int j = 0; j < SC_NV; j++) {
                    const int64_t s = t0 + (int64_t)(j * 32 + lane) * V;
                    v[j] = s < s1 ? *reinterpret_cast<const Vec*>(x + s) : zero;
                }
#pragma unroll
                for (int j = 0; j < SC_NV; j++) {
#pragma unroll
                    for (int w = 0; w < V; w++) part += v[j].e[w];
                }
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
            if (lane == 0) totals[u] = part;
            continue;
        }
        T carry = totals != nullptr ? totals[u] : (T)0;
        Vec v[SC_NV];
#pragma unroll
        for (int j = 0; j < SC_NV; j++) {
            const int64_t s = s0 + (int64_t)(j * 32 + lane) * V;
            v[j] = s < s1 ? *reinterpret_cast<const Vec*>(x + s) : zero;
        }
        for (int64_t t0 = s0; t0 < s1; t0 += TILE) {
            Vec a[SC_NV];
            T inc[SC_NV];
#pragma unroll
            for (int j = 0; j < SC_NV; j++) {
                a[j] = v[j];
                T t = (T)0;
#pragma unroll
                for (int w = 0; w < V; w++) t += a[j].e[w];
                inc[j] = t;
            }
            if (t0 + TILE < s1) {
#pragma unroll
                for (int j = 0; j < SC_NV; j++) {
                    const int64_t s = t0 + TILE + (int64_t)(j * 32 + lane) * V;
                    v[j] = s < s1 ? *reinterpret_cast<const Vec*>(x + s) : zero;
                }
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int j = 0; j < SC_NV; j++) { const T o = sc_shfl_up(inc[j], d); if (lane >= d) inc[j] += o; }
            }
#pragma unroll
            for (int j = 0; j < SC_NV; j++) {
                const T prev = sc_shfl_up(inc[j], 1);
                const T sub_total = sc_shfl(inc[j], 31);
                T run = lane == 0 ? carry : carry + prev;
                Vec o;
#pragma unroll
                for (int w = 0; w < V; w++) {
                    if (p.exclusive) { o.e[w] = run; run += a[j].e[w]; } else { run += a[j].e[w]; o.e[w] = run; }
                }
                const int64_t s = t0 + (int64_t)(j * 32 + lane) * V;
                if (s < s1) *reinterpret_cast<Vec*>(y + s) = o;
                carry += sub_total;
            }
        }
    }
}

// exclusive scan of the chunk totals of each row, in place: one block per row, SC_THREADS * 4 totals per step
template <typename T>
__global__ void __launch_bounds__(SC_THREADS)
scan_totals_kernel(int64_t chunks, T* __restrict__ totals)
{
    __shared__ T warp_s[SC_WARPS];
    __shared__ T carry_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* t = totals + (int64_t)blockIdx.x * chunks;
    if (threadIdx.x == 0) carry_s = (T)0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < chunks; t0 += SC_THREADS * 4) {
        T a[4], tot = (T)0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int64_t e = t0 + (int64_t)threadIdx.x * 4 + q;
            a[q] = e < chunks ? t[e] : (T)0;
            tot += a[q];
        }
        T inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const T o = sc_shfl_up(inc, d); if (lane >= d) inc += o; }
        if (lane == 31) warp_s[warp] = inc;
        __syncthreads();
        T run = carry_s;
        for (int w = 0; w < warp; w++) run += warp_s[w];
        const T prev = sc_shfl_up(inc, 1);
        if (lane != 0) run += prev;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int64_t e = t0 + (int64_t)threadIdx.x * 4 + q;
            if (e < chunks) t[e] = run;
            run += a[q];
        }
        __syncthreads();
        if (threadIdx.x == SC_THREADS - 1) carry_s = run;
        __syncthreads();
    }
}

struct ScanCols {
    int64_t cols;            // m * k columns
    int64_t n, k;
    int64_t chunk, chunks;
    int reverse, exclusive;
};


// one thread per V adjacent (i, h) columns and chunk (V > 1: k % V == 0 and 16-byte aligned buffers); adjacent threads
// read adjacent columns, SC_UNROLL steps of the scan are in flight.
constexpr int SC_UNROLL = 8;

template <typename T, int V, int PHASE>
__global__ void __launch_bounds__(SC_THREADS)
scan_cols_kernel(ScanCols p, const T* __restrict__ A, T* __restrict__ B, T* __restrict__ totals)
{
    typedef ScVec<T, V> Vec;
    const int64_t vcols = p.cols / V;
    const int64_t col_blocks = (vcols + SC_THREADS - 1) / SC_THREADS;
    for (int64_t bb = blockIdx.x; bb < col_blocks * p.chunks; bb += gridDim.x) {
        const int64_t blk = (PHASE == 1 && totals != nullptr) ? col_blocks * p.chunks - 1 - bb : bb;
        const int64_t c = blk / col_blocks;
        const int64_t g = ((blk - c * col_blocks) * SC_THREADS + threadIdx.x) * V;
        if (g >= p.cols) continue;
        const int64_t i = g / p.k, h = g - i * p.k;
        const int64_t s0 = c * p.chunk;
        const int64_t s1 = s0 + p.chunk < p.n ? s0 + p.chunk : p.n;
        const int64_t base = i * p.n * p.k + h;
        Vec run;
#pragma unroll
        for (int w = 0; w < V; w++) run.e[w] = (T)0;
        if (PHASE == 1 && totals != nullptr) {
            for (int64_t cc = 0; cc < c; cc++) {
                const Vec t = *reinterpret_cast<const Vec*>(totals + cc * p.cols + g);
#pragma unroll
                for (int w = 0; w < V; w++) run.e[w] += t.e[w];
            }
        }
        int64_t s = s0;
        for (; s + SC_UNROLL <= s1; s += SC_UNROLL) {
            Vec a[SC_UNROLL];
#pragma unroll
            for (int q = 0; q < SC_UNROLL; q++) {
                a[q] = *reinterpret_cast<const Vec*>(A + base + (p.reverse ? p.n - 1 - (s + q) : s + q) * p.k);
            }
#pragma unroll
            for (int q = 0; q < SC_UNROLL; q++) {
                Vec o;
#pragma unroll
                for (int w = 0; w < V; w++) {
                    if (p.exclusive) { o.e[w] = run.e[w]; run.e[w] += a[q].e[w]; } else { run.e[w] += a[q].e[w]; o.e[w] = run.e[w]; }
                }
                if (PHASE == 1) *reinterpret_cast<Vec*>(B + base + (p.reverse ? p.n - 1 - (s + q) : s + q) * p.k) = o;
            }
        }
        for (; s < s1; s++) {
            const int64_t at = base + (p.reverse ? p.n - 1 - s : s) * p.k;
            const Vec a = *reinterpret_cast<const Vec*>(A + at);
            Vec o;
#pragma unroll
            for (int w = 0; w < V; w++) {
                if (p.exclusive) { o.e[w] = run.e[w]; run.e[w] += a.e[w]; } else { run.e[w] += a.e[w]; o.e[w] = run.e[w]; }
            }
            if (PHASE == 1) *reinterpret_cast<Vec*>(B + at) = o;
        }
        if (PHASE == 0) *reinterpret_cast<Vec*>(totals + c * p.cols + g) = run;
    }
}

template <typename T>
int scan_rows_launch(ScanRows p, const T* X, T* Y)
{
    rb200::Context& c = rb200::ctx();
    // a warp per (row, chunk): split rows only while the machine has idle warps, never below 4 tiles per chunk
    const int64_t want = (int64_t)c.sm_count * SC_WARPS * 6;
    int64_t chunks = 1;
    if (p.m < want) {
        chunks = want / p.m;
        const int64_t maxc = rb_ceil_div(p.n, 4 * SC_TILE);
        if (chunks > maxc) chunks = maxc;
        if (chunks < 1) chunks = 1;
    }
    constexpr int V = 16 / (int)sizeof(T);
    const bool vec = p.incx == 1 && p.incy == 1 && !p.revx && !p.revy && (p.n % V) == 0 && (p.ldx % V) == 0 && (p.ldy % V) == 0 &&
                     ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0;
    const int64_t tile = vec ? 32 * V * SC_NV : SC_TILE;
    p.chunk = rb_ceil_div(rb_ceil_div(p.n, chunks), tile) * tile;
    p.chunks = rb_ceil_div(p.n, p.chunk);
    int64_t grid = rb_ceil_div(p.m * p.chunks, SC_WARPS);
    const int64_t cap = (int64_t)c.sm_count * 8;
    if (grid > cap) grid = cap;
    T* totals = nullptr;
    if (p.chunks > 1) {
        int rc = rb200::ensure_workspace((size_t)(p.m * p.chunks) * sizeof(T));
        if (rc != RB200_OK) return rc;
        totals = reinterpret_cast<T*>(c.workspace);
        if (vec) scan_rows_vec_kernel<T, 0><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, X, Y, totals);
        else     scan_rows_kernel<T, 0><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, X, Y, totals);
        RB_LAUNCH_CHECK("scan_rows_kernel<totals>");
        scan_totals_kernel<T><<<(unsigned)p.m, SC_THREADS, 0, c.stream>>>(p.chunks, totals);
        RB_LAUNCH_CHECK("scan_totals_kernel");
    }
    if (vec) scan_rows_vec_kernel<T, 1><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, X, Y, totals);
    else     scan_rows_kernel<T, 1><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, X, Y, totals);
    RB_LAUNCH_CHECK("scan_rows_kernel");
    return RB200_OK;
}

template <typename T, int V>
int scan_cols_launch_v(ScanCols p, const T* A, T* B)
{
    rb200::Context& c = rb200::ctx();
    const int64_t col_blocks = rb_ceil_div(p.cols / V, SC_THREADS);
    // two blocks per SM keep SC_UNROLL 16-byte loads per thread in flight; the carry of a chunk is the sum of the totals
    // of the chunks before it, so their number stays small
    const int64_t want = (int64_t)c.sm_count * 2;
    int64_t chunks = 1;
    if (col_blocks < want) {
        chunks = want / col_blocks;
        const int64_t maxc = rb_ceil_div(p.n, 4 * SC_UNROLL);
        if (chunks > maxc) chunks = maxc;
        if (chunks < 1) chunks = 1;
    }
    p.chunk = rb_ceil_div(p.n, chunks);
    p.chunks = rb_ceil_div(p.n, p.chunk);
    int64_t grid = col_blocks * p.chunks;
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (grid > cap) grid = cap;
    T* totals = nullptr;
    if (p.chunks > 1) {
        int rc = rb200::ensure_workspace((size_t)(p.chunks * p.cols) * sizeof(T));
        if (rc != RB200_OK) return rc;
        totals = reinterpret_cast<T*>(c.workspace);
        scan_cols_kernel<T, V, 0><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, A, B, totals);
        RB_LAUNCH_CHECK("scan_cols_kernel<totals>");
    }
    scan_cols_kernel<T, V, 1><<<(unsigned)grid, SC_THREADS, 0, c.stream>>>(p, A, B, totals);
    RB_LAUNCH_CHECK("scan_cols_kernel");
    return RB200_OK;
}

template <typename T>
int scan_cols_launch(int64_t m, int64_t n, int64_t k, bool exclusive, bool reverse, const T* A, T* B)
{
    ScanCols p;
    p.cols = m * k; p.n = n; p.k = k; p.reverse = reverse; p.exclusive = exclusive;
    constexpr int V = 16 / (int)sizeof(T);
    const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
    // the totals sit at the 256-byte aligned workspace base and c * cols + g stays a multiple of V when k is
    if (aligned && (k % V) == 0) return scan_cols_launch_v<T, V>(p, A, B);
    return scan_cols_launch_v<T, 1>(p, A, B);
}

template <typename T>
int cumsumb_typed(int64_t m, int64_t n, int64_t k, bool exclusive, bool reverse, const void* A, void* B)
{
    const T* a = reinterpret_cast<const T*>(A);
    T* b = reinterpret_cast<T*>(B);
    if (k == 1) {
        ScanRows p;
        p.m = m; p.n = n; p.ldx = n; p.ldy = n; p.incx = 1; p.incy = 1;
        p.revx = reverse; p.revy = reverse; p.exclusive = exclusive;
        return scan_rows_launch<T>(p, a, b);
    }
    return scan_cols_launch<T>(m, n, k, exclusive, reverse, a, b);
}

template <typename T>
int cumsum_typed(int64_t n, int64_t incX, int64_t incY, bool exclusive, bool reverse, const void* X, void* Y)
{
    ScanRows p;
    p.m = 1; p.n = n; p.ldx = 0; p.ldy = 0; p.incx = incX; p.incy = incY;
    // the reference walks X forwards in both directions and only reverses where Y is written (PhpMath.php:1835-1842)
    p.revx = 0; p.revy = reverse; p.exclusive = exclusive;
    return scan_rows_launch<T>(p, reinterpret_cast<const T*>(X), reinterpret_cast<T*>(Y));
}

// ---- searchsorted -----------------------------------------------------------------------------------------------

// flags[row] != 0 when row `row` of A is not non-decreasing (NaN included): the linear walk of the reference and a
// bisection then disagree, and the search kernel walks that row instead
template <typename T>
__global__ void __launch_bounds__(SC_THREADS)
ss_unsorted_kernel(int64_t rows, int64_t n, int64_t ldA, const T* __restrict__ A, int* __restrict__ flags)
{
    const int64_t total = rows * n;
    for (int64_t t = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x; t < total; t += (int64_t)gridDim.x * SC_THREADS) {
        const int64_t r = t / n, j = t - r * n;
        const T a = A[r * ldA + j];
        bool bad = !(a == a);
        if (j + 1 < n) bad = bad || !(a <= A[r * ldA + j + 1]);
        if (bad) flags[r] = 1;
    }
}

template <typename T, typename I, bool RIGHT>
__global__ void __launch_bounds__(SC_THREADS)
ss_search_kernel(int64_t m, int64_t n, int64_t ldA, const T* __restrict__ A, const T* __restrict__ X, int64_t incX,
                 I* __restrict__ Y, int64_t incY, const int* __restrict__ flags)
{
    for (int64_t i = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x; i < m; i += (int64_t)gridDim.x * SC_THREADS) {
        const T v = X[i * incX];
        const T* a = A + i * ldA;
        int64_t res;
        if (flags[ldA ? i : 0]) {
            // PhpMath.php:1799-1814: stop at the first element the value is not above
            int64_t j = 0;
            for (; j < n; j++) { if (RIGHT ? !(v >= a[j]) : !(v > a[j])) break; }
            res = j;
        } else {
            int64_t lo = 0, hi = n;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                const bool above = RIGHT ? (v >= a[mid]) : (v > a[mid]);
                if (above) lo = mid + 1; else hi = mid;
            }
            res = lo;
        }
        Y[i * incY] = (I)res;
    }
}

template <typename T, typename I>
int searchsorted_typed(int64_t m, int64_t n, const void* A, int64_t ldA, const void* X, int64_t incX, bool right,
                       void* Y, int64_t incY)
{
    rb200::Context& c = rb200::ctx();
    const int64_t rows = ldA ? m : 1;
    int rc = rb200::ensure_workspace((size_t)rows * sizeof(int));
    if (rc != RB200_OK) return rc;
    int* flags = reinterpret_cast<int*>(c.workspace);
    RB_CUDA(cudaMemsetAsync(flags, 0, (size_t)rows * sizeof(int), c.stream));
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (n > 0) {
        int64_t g = rb_ceil_div(rows * n, SC_THREADS);
        if (g > cap) g = cap;
        ss_unsorted_kernel<T><<<(unsigned)g, SC_THREADS, 0, c.stream>>>(rows, n, ldA, reinterpret_cast<const T*>(A), flags);
        RB_LAUNCH_CHECK("ss_unsorted_kernel");
    }
    int64_t g = rb_ceil_div(m, SC_THREADS);
    if (g > cap) g = cap;
    if (right) {
        ss_search_kernel<T, I, true><<<(unsigned)g, SC_THREADS, 0, c.stream>>>(
            m, n, ldA, reinterpret_cast<const T*>(A), reinterpret_cast<const T*>(X), incX, reinterpret_cast<I*>(Y), incY, flags);
    } else {
        ss_search_kernel<T, I, false><<<(unsigned)g, SC_THREADS, 0, c.stream>>>(
            m, n, ldA, reinterpret_cast<const T*>(A), reinterpret_cast<const T*>(X), incX, reinterpret_cast<I*>(Y), incY, flags);
    }
    RB_LAUNCH_CHECK("ss_search_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" int rb200_cumsumb(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, int exclusive,
                             int reverse, void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !B) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1 || offA < 0 || offB < 0) RB_INVALID("cumsumb: bad shape / offset");
    const int es = rb_dtype_size(dtype);
    const char* a = reinterpret_cast<const char*>(A) + offA * es;
    char* b = reinterpret_cast<char*>(B) + offB * es;
    switch (dtype) {
        case RB200_FLOAT32: return cumsumb_typed<float>(m, n, k, exclusive != 0, reverse != 0, a, b);
        case RB200_FLOAT64: return cumsumb_typed<double>(m, n, k, exclusive != 0, reverse != 0, a, b);
        case RB200_INT32:   return cumsumb_typed<int32_t>(m, n, k, exclusive != 0, reverse != 0, a, b);
        case RB200_INT64:   return cumsumb_typed<long long>(m, n, k, exclusive != 0, reverse != 0, a, b);
        default: RB_UNSUPPORTED("cumsumb: unsupported dtype %d", dtype);
    }
}

extern "C" int rb200_cumsum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int exclusive, int reverse,
                            void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!X || !Y) RB_INVALID("NULL buffer");
    if (n < 1 || incX < 1 || incY < 1 || offX < 0 || offY < 0) RB_INVALID("cumsum: bad shape / increment / offset");
    const int es = rb_dtype_size(dtype);
    const char* x = reinterpret_cast<const char*>(X) + offX * es;
    char* y = reinterpret_cast<char*>(Y) + offY * es;
    switch (dtype) {
        case RB200_FLOAT32: return cumsum_typed<float>(n, incX, incY, exclusive != 0, reverse != 0, x, y);
        case RB200_FLOAT64: return cumsum_typed<double>(n, incX, incY, exclusive != 0, reverse != 0, x, y);
        case RB200_INT32:   return cumsum_typed<int32_t>(n, incX, incY, exclusive != 0, reverse != 0, x, y);
        case RB200_INT64:   return cumsum_typed<long long>(n, incX, incY, exclusive != 0, reverse != 0, x, y);
        default: RB_UNSUPPORTED("cumsum: unsupported dtype %d", dtype);
    }
}

extern "C" int rb200_searchsorted(int dtype, int index_dtype, int64_t m, int64_t n, const void* A, int64_t offA, int64_t ldA,
                                  const void* X, int64_t offX, int64_t incX, int right, void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!A || !X || !Y) RB_INVALID("NULL buffer");
    if (m < 1 || n < 0 || ldA < 0 || incX < 1 || incY < 1 || offA < 0 || offX < 0 || offY < 0) {
        RB_INVALID("searchsorted: bad shape / increment / offset");
    }
    const int es = rb_dtype_size(dtype), is = rb_dtype_size(index_dtype);
    const bool i32 = index_dtype == RB200_INT32 || index_dtype == RB200_UINT32;
    const bool i64 = index_dtype == RB200_INT64 || index_dtype == RB200_UINT64;
    if (!i32 && !i64) RB_UNSUPPORTED("searchsorted: index dtype %d is not a 32- or 64-bit integer", index_dtype);
    const char* a = reinterpret_cast<const char*>(A) + offA * es;
    const char* x = reinterpret_cast<const char*>(X) + offX * es;
    char* y = reinterpret_cast<char*>(Y) + offY * is;
    if (dtype == RB200_FLOAT32) {
        return i32 ? searchsorted_typed<float, int32_t>(m, n, a, ldA, x, incX, right != 0, y, incY)
                   : searchsorted_typed<float, int64_t>(m, n, a, ldA, x, incX, right != 0, y, incY);
    }
    if (dtype == RB200_FLOAT64) {
        return i32 ? searchsorted_typed<double, int32_t>(m, n, a, ldA, x, incX, right != 0, y, incY)
                   : searchsorted_typed<double, int64_t>(m, n, a, ldA, x, incX, right != 0, y, incY);
    }
    RB_UNSUPPORTED("searchsorted: unsupported dtype %d", dtype);
}
