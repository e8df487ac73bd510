// reduce.cu -- reduceSum / reduceMax / reduceArgMax over the middle axis of A[m,n,k] -> B[m,k]
// (HBM-bound).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:83-130, 1552-1643):
//   sum    : acc = 0.0; acc += A[i,t,j] for t = 0..n-1
//   max    : NaN-sticky maximum (the rule the reference pins for accelerated drivers,
//            LinearAlgebraTest.php:7617-7643; equals math_max on NaN-free / NaN-last rows)
//   argmax : index of the first strict maximum; ties -> lowest index; a NaN never wins,
//            except that a NaN in position 0 is never displaced (math_imax :114-130)
// Algorithmic bytes: m*n*k*sizeof(T) read + m*k*sizeof(out) written.
//
// Kernel families (all warp-shuffle + shared-memory tree; 128-bit loads when aligned):
//   k == 1, n <= 4096 : one warp per row            (reduce_rows_warp)
//   k == 1, long rows : one CTA per (row, segment)  (reduce_rows_block)
//   k  > 1            : threads along k, CTA tile TX x TY over (k, n), optional split of n
//                       (reduce_cols)
// Splitting the reduced axis keeps every SM busy when m*k alone is too small to fill 148 SMs
// (e.g. axis=0 of [65536,1024], or 64 rows of 1e6); the last CTA of an output to finish (atomic
// ticket) merges the partials in a fixed order inside the same launch.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace {

constexpr int RD_THREADS = 256;

template <typename T> __device__ __forceinline__ T rd_neg_inf();
template <> __device__ __forceinline__ float  rd_neg_inf<float>()  { return -INFINITY; }
template <> __device__ __forceinline__ double rd_neg_inf<double>() { return -(double)INFINITY; }
template <typename T> __device__ __forceinline__ T rd_nan();
template <> __device__ __forceinline__ float  rd_nan<float>()  { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double rd_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

template <typename T> __device__ __forceinline__ T rd_shfl_xor(T v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }

// ---- reduction policies ------------------------------------------------------------------
template <typename T> struct SumOp {
    struct Acc { T s; };
    static __device__ __forceinline__ Acc identity() { return Acc{(T)0}; }
    static __device__ __forceinline__ void accum(Acc& a, T v, int) { a.s += v; }
    static __device__ __forceinline__ Acc combine(const Acc& a, const Acc& b) { return Acc{a.s + b.s}; }
    static __device__ __forceinline__ Acc shfl(const Acc& a, int mask) { return Acc{rd_shfl_xor(a.s, mask)}; }
    static __device__ __forceinline__ void store(void* B, int64_t pos, const Acc& a, int) { reinterpret_cast<T*>(B)[pos] = a.s; }
};

template <typename T> struct MaxOp {
    struct Acc { T m; int nan; };
    static __device__ __forceinline__ Acc identity() { return Acc{rd_neg_inf<T>(), 0}; }
    static __device__ __forceinline__ void accum(Acc& a, T v, int) { a.nan |= (v != v); a.m = fmax(a.m, v); }   // fmax drops NaN
    static __device__ __forceinline__ Acc combine(const Acc& a, const Acc& b) { return Acc{(T)fmax(a.m, b.m), a.nan | b.nan}; }
    static __device__ __forceinline__ Acc shfl(const Acc& a, int mask) { return Acc{rd_shfl_xor(a.m, mask), rd_shfl_xor(a.nan, mask)}; }
    static __device__ __forceinline__ void store(void* B, int64_t pos, const Acc& a, int) { reinterpret_cast<T*>(B)[pos] = a.nan ? rd_nan<T>() : a.m; }
};

template <typename T> struct ArgMaxOp {
    struct Acc { T v; int idx; };
    static __device__ __forceinline__ Acc identity() { return Acc{rd_neg_inf<T>(), INT_MAX}; }
    static __device__ __forceinline__ void accum(Acc& a, T v, int idx)
    {
        if (v != v) {
            if (idx != 0) return;             // a NaN never wins ...
            v = -rd_neg_inf<T>();             // ... but a NaN at position 0 is never displaced
        }
        if (v > a.v || (v == a.v && idx < a.idx)) { a.v = v; a.idx = idx; }
    }
    static __device__ __forceinline__ Acc combine(const Acc& a, const Acc& b)
    {
        return (b.v > a.v || (b.v == a.v && b.idx < a.idx)) ? b : a;
    }
    static __device__ __forceinline__ Acc shfl(const Acc& a, int mask) { return Acc{rd_shfl_xor(a.v, mask), rd_shfl_xor(a.idx, mask)}; }
    static __device__ __forceinline__ void store(void* B, int64_t pos, const Acc& a, int idx64)
    {
        if (idx64) reinterpret_cast<int64_t*>(B)[pos] = (int64_t)a.idx;
        else       reinterpret_cast<int32_t*>(B)[pos] = (int32_t)a.idx;
    }
};

// Arg-max of ONE STRIPE of a row-sharded axis (rb200_reduce_argmax_partial): the record (value, index) a rank
// contributes to the cross-rank merge.  GLOBAL0 = the stripe starts at global position 0: math_imax's rule as above,
// with the never-displaced NaN at position 0 reported as +INF (nothing is greater than either, and ties keep the lower
// index, so the merge cannot tell them apart).  Otherwise the stripe continues a running maximum that is never NaN
// there (PhpMath.php:121-127: `if($value>$max)` is false for a NaN on either side), so NaNs are skipped wherever they
// sit; an all-NaN stripe reports (-INF, INT_MAX), which never wins.
template <typename T, bool GLOBAL0> struct ArgMaxPairOp {
    using Acc = typename ArgMaxOp<T>::Acc;
    static __device__ __forceinline__ Acc identity() { return ArgMaxOp<T>::identity(); }
    static __device__ __forceinline__ void accum(Acc& a, T v, int idx)
    {
        if (GLOBAL0) { ArgMaxOp<T>::accum(a, v, idx); return; }
        if (v != v) return;
        if (v > a.v || (v == a.v && idx < a.idx)) { a.v = v; a.idx = idx; }
    }
    static __device__ __forceinline__ Acc combine(const Acc& a, const Acc& b) { return ArgMaxOp<T>::combine(a, b); }
    static __device__ __forceinline__ Acc shfl(const Acc& a, int mask) { return ArgMaxOp<T>::shfl(a, mask); }
    static __device__ __forceinline__ void store(void* B, int64_t pos, const Acc& a, int)
    {
        // 16-byte records {double value; int64 index}
        int4 rec;
        const double v = (double)a.v;
        const long long i = (long long)a.idx;
        memcpy(&rec.x, &v, 8);
        memcpy(&rec.z, &i, 8);
        reinterpret_cast<int4*>(B)[pos] = rec;
    }
};

template <typename Op> __device__ __forceinline__ typename Op::Acc warp_reduce(typename Op::Acc a)
{
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) a = Op::combine(a, Op::shfl(a, mask));
    return a;
}

// whole-CTA reduce; result valid in thread 0
template <typename Op> __device__ __forceinline__ typename Op::Acc block_reduce(typename Op::Acc a)
{
    __shared__ typename Op::Acc warp_part[RD_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warp_reduce<Op>(a);
    if (lane == 0) warp_part[warp] = a;
    __syncthreads();
    if (warp == 0) {
        a = lane < (int)(blockDim.x >> 5) ? warp_part[lane] : Op::identity();
        a = warp_reduce<Op>(a);
    }
    return a;
}

// partial accumulators written by other CTAs: read through L2 (volatile), never from a stale L1 line
template <typename Op> __device__ __forceinline__ typename Op::Acc ld_acc(const typename Op::Acc* p)
{
    typename Op::Acc a;
    if (sizeof(a) == 4)      { const int v  = __ldcg(reinterpret_cast<const int*>(p));  memcpy(&a, &v, 4); }
    else if (sizeof(a) == 8) { const int2 v = __ldcg(reinterpret_cast<const int2*>(p)); memcpy(&a, &v, 8); }
    else                     { const int4 v = __ldcg(reinterpret_cast<const int4*>(p)); memcpy(&a, &v, 16); }
    return a;
}

template <typename T> struct RdVec { static constexpr int N = 16 / sizeof(T); };
template <typename T> union RdPack { int4 raw; T v[16 / sizeof(T)]; };

// accumulate elements [begin,end) of one contiguous run, `nthreads` cooperating threads,
// this thread's rank `rank`.  VEC: 128-bit loads (run base 16-byte aligned, begin % V == 0).
template <typename T, typename Op, bool VEC>
__device__ __forceinline__ void accum_run(typename Op::Acc& acc, const T* __restrict__ run,
                                          int64_t begin, int64_t end, int rank, int nthreads)
{
    if (VEC) {
        constexpr int V = RdVec<T>::N;
        const int4* run4 = reinterpret_cast<const int4*>(run);
        const int64_t vbegin = begin / V, vend = end / V;        // full vectors
        int64_t c = vbegin + rank;
        // 4 independent 128-bit loads in flight per thread
        for (; c + 3LL * nthreads < vend; c += 4LL * nthreads) {
            RdPack<T> p0, p1, p2, p3;
            p0.raw = __ldcs(run4 + c);
            p1.raw = __ldcs(run4 + c + nthreads);
            p2.raw = __ldcs(run4 + c + 2LL * nthreads);
            p3.raw = __ldcs(run4 + c + 3LL * nthreads);
#pragma unroll
            for (int e = 0; e < V; e++) Op::accum(acc, p0.v[e], (int)(c * V + e));
#pragma unroll
            for (int e = 0; e < V; e++) Op::accum(acc, p1.v[e], (int)((c + nthreads) * V + e));
#pragma unroll
            for (int e = 0; e < V; e++) Op::accum(acc, p2.v[e], (int)((c + 2LL * nthreads) * V + e));
#pragma unroll
            for (int e = 0; e < V; e++) Op::accum(acc, p3.v[e], (int)((c + 3LL * nthreads) * V + e));
        }
        for (; c < vend; c += nthreads) {
            RdPack<T> p;
            p.raw = __ldcs(run4 + c);
#pragma unroll
            for (int e = 0; e < V; e++) Op::accum(acc, p.v[e], (int)(c * V + e));
        }
        for (int64_t t = vend * V + rank; t < end; t += nthreads) Op::accum(acc, __ldcs(run + t), (int)t);
    } else {
        for (int64_t t = begin + rank; t < end; t += nthreads) Op::accum(acc, __ldcs(run + t), (int)t);
    }
}

// ---- k == 1: one warp per row ----------------------------------------------------------------
template <typename T, typename Op, bool VEC>
__global__ void __launch_bounds__(RD_THREADS)
reduce_rows_warp(int64_t m, int64_t n, const T* __restrict__ A, void* __restrict__ B, int64_t offB, int idx64)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps_per_grid = (int64_t)gridDim.x * (RD_THREADS / 32);
    for (int64_t row = (int64_t)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5); row < m; row += warps_per_grid) {
        typename Op::Acc acc = Op::identity();
        accum_run<T, Op, VEC>(acc, A + row * n, 0, n, lane, 32);
        acc = warp_reduce<Op>(acc);
        if (lane == 0) Op::store(B, offB + row, acc, idx64);
    }
}

// ---- k == 1: one CTA per (row, segment) --------------------------------------------------------
// With S > 1 segments per row each CTA publishes its partial, and the LAST CTA of a row to finish
// (atomic ticket; the counter resets itself) merges the S partials in a fixed order -- one launch,
// deterministic result.
template <typename T, typename Op, bool VEC>
__global__ void __launch_bounds__(RD_THREADS)
reduce_rows_block(int64_t m, int64_t n, int S, int64_t seglen, const T* __restrict__ A,
                  void* __restrict__ B, int64_t offB, int idx64, typename Op::Acc* partial, int* counters)
{
    __shared__ int is_last;
    const int seg = blockIdx.x;
    for (int64_t row = blockIdx.y; row < m; row += gridDim.y) {
        const int64_t begin = (int64_t)seg * seglen;
        int64_t end = begin + seglen;
        if (end > n) end = n;
        typename Op::Acc acc = Op::identity();
        if (begin < end) accum_run<T, Op, VEC>(acc, A + row * n, begin, end, threadIdx.x, RD_THREADS);
        acc = block_reduce<Op>(acc);
        if (S == 1) {
            if (threadIdx.x == 0) Op::store(B, offB + row, acc, idx64);
        } else {
            if (threadIdx.x == 0) {
                partial[row * S + seg] = acc;
                __threadfence();
                const int ticket = atomicAdd(&counters[row], 1);
                is_last = (ticket == S - 1);
                if (is_last) counters[row] = 0;
            }
            __syncthreads();
            if (is_last) {
                __threadfence();
                typename Op::Acc a2 = Op::identity();
                for (int s2 = threadIdx.x; s2 < S; s2 += RD_THREADS) a2 = Op::combine(a2, ld_acc<Op>(partial + row * S + s2));
                __syncthreads();
                a2 = block_reduce<Op>(a2);
                if (threadIdx.x == 0) Op::store(B, offB + row, a2, idx64);
            }
        }
        __syncthreads();      // block_reduce's shared scratch and is_last are reused by the next row
    }
}

// ---- k > 1: threads along k -----------------------------------------------------------------------
// blockDim = (TX, TY), TX*TY == 256.  Thread (tx,ty) owns V consecutive columns and the
// rows t = seg*tlen + ty, +TY, ... of its segment; a shared-memory tree merges the TY partials.
template <typename T, typename Op, bool VEC>
__global__ void __launch_bounds__(RD_THREADS)
reduce_cols(int64_t m, int64_t n, int64_t k, int S, int64_t tlen, const T* __restrict__ A,
            void* __restrict__ B, int64_t offB, int idx64, typename Op::Acc* partial, int* counters)
{
    constexpr int V = VEC ? RdVec<T>::N : 1;
    __shared__ typename Op::Acc sm[RD_THREADS * V];
    __shared__ int is_last;
    const int TX = blockDim.x, TY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t j0 = ((int64_t)blockIdx.x * TX + tx) * V;
    const int64_t units = m * S;
    for (int64_t u = blockIdx.y; u < units; u += gridDim.y) {
        const int64_t i = u / S;
        const int seg = (int)(u - i * S);
        const int64_t tbegin = (int64_t)seg * tlen;
        int64_t tend = tbegin + tlen;
        if (tend > n) tend = n;
        typename Op::Acc acc[V];
#pragma unroll
        for (int e = 0; e < V; e++) acc[e] = Op::identity();
        if (j0 < k) {
            const T* base = A + i * n * k + j0;
            int64_t t = tbegin + ty;
            if (VEC) {
                for (; t + 3LL * TY < tend; t += 4LL * TY) {
                    RdPack<T> p0, p1, p2, p3;
                    p0.raw = __ldcs(reinterpret_cast<const int4*>(base + t * k));
                    p1.raw = __ldcs(reinterpret_cast<const int4*>(base + (t + TY) * k));
                    p2.raw = __ldcs(reinterpret_cast<const int4*>(base + (t + 2LL * TY) * k));
                    p3.raw = __ldcs(reinterpret_cast<const int4*>(base + (t + 3LL * TY) * k));
#pragma unroll
                    for (int e = 0; e < V; e++) {
                        Op::accum(acc[e], p0.v[e], (int)t);
                        Op::accum(acc[e], p1.v[e], (int)(t + TY));
                        Op::accum(acc[e], p2.v[e], (int)(t + 2LL * TY));
                        Op::accum(acc[e], p3.v[e], (int)(t + 3LL * TY));
                    }
                }
                for (; t < tend; t += TY) {
                    RdPack<T> p;
                    p.raw = __ldcs(reinterpret_cast<const int4*>(base + t * k));
#pragma unroll
                    for (int e = 0; e < V; e++) Op::accum(acc[e], p.v[e], (int)t);
                }
            } else {
                for (; t + 3LL * TY < tend; t += 4LL * TY) {
                    const T v0 = __ldcs(base + t * k), v1 = __ldcs(base + (t + TY) * k);
                    const T v2 = __ldcs(base + (t + 2LL * TY) * k), v3 = __ldcs(base + (t + 3LL * TY) * k);
                    Op::accum(acc[0], v0, (int)t);
                    Op::accum(acc[0], v1, (int)(t + TY));
                    Op::accum(acc[0], v2, (int)(t + 2LL * TY));
                    Op::accum(acc[0], v3, (int)(t + 3LL * TY));
                }
                for (; t < tend; t += TY) Op::accum(acc[0], __ldcs(base + t * k), (int)t);
            }
        }
        // merge the TY partials of each column: sm[(ty*TX + tx)*V + e]
#pragma unroll
        for (int e = 0; e < V; e++) sm[(ty * TX + tx) * V + e] = acc[e];
        __syncthreads();
        for (int half = TY >> 1; half > 0; half >>= 1) {
            if (ty < half) {
#pragma unroll
                for (int e = 0; e < V; e++) {
                    sm[(ty * TX + tx) * V + e] = Op::combine(sm[(ty * TX + tx) * V + e], sm[((ty + half) * TX + tx) * V + e]);
                }
            }
            __syncthreads();
        }
        if (ty == 0 && j0 < k) {
#pragma unroll
            for (int e = 0; e < V; e++) {
                if (j0 + e < k) {
                    if (S == 1) Op::store(B, offB + i * k + j0 + e, sm[tx * V + e], idx64);
                    else        partial[(i * S + seg) * k + j0 + e] = sm[tx * V + e];
                }
            }
        }
        if (S > 1) {
            // the last CTA of this (i, column block) to finish merges the S partials (fixed order:
            // ty-strided over s, then the same shared-memory tree) -- one launch, deterministic
            __threadfence();
            __syncthreads();
            if (tx == 0 && ty == 0) {
                int* ctr = counters + i * gridDim.x + blockIdx.x;
                const int ticket = atomicAdd(ctr, 1);
                is_last = (ticket == S - 1);
                if (is_last) *ctr = 0;
            }
            __syncthreads();
            if (is_last) {
                __threadfence();
#pragma unroll
                for (int e = 0; e < V; e++) {
                    typename Op::Acc a2 = Op::identity();
                    if (j0 + e < k) {
                        const typename Op::Acc* pp = partial + (i * S) * k + j0 + e;
                        int s2 = ty;
                        for (; s2 + 3 * TY < S; s2 += 4 * TY) {       // 4 independent L2 loads in flight
                            const typename Op::Acc p0 = ld_acc<Op>(pp + (int64_t)s2 * k);
                            const typename Op::Acc p1 = ld_acc<Op>(pp + (int64_t)(s2 + TY) * k);
                            const typename Op::Acc p2 = ld_acc<Op>(pp + (int64_t)(s2 + 2 * TY) * k);
                            const typename Op::Acc p3 = ld_acc<Op>(pp + (int64_t)(s2 + 3 * TY) * k);
                            a2 = Op::combine(Op::combine(Op::combine(Op::combine(a2, p0), p1), p2), p3);
                        }
                        for (; s2 < S; s2 += TY) a2 = Op::combine(a2, ld_acc<Op>(pp + (int64_t)s2 * k));
                    }
                    sm[(ty * TX + tx) * V + e] = a2;
                }
                __syncthreads();
                for (int half = TY >> 1; half > 0; half >>= 1) {
                    if (ty < half) {
#pragma unroll
                        for (int e = 0; e < V; e++) {
                            sm[(ty * TX + tx) * V + e] = Op::combine(sm[(ty * TX + tx) * V + e], sm[((ty + half) * TX + tx) * V + e]);
                        }
                    }
                    __syncthreads();
                }
                if (ty == 0 && j0 < k) {
#pragma unroll
                    for (int e = 0; e < V; e++) {
                        if (j0 + e < k) Op::store(B, offB + i * k + j0 + e, sm[tx * V + e], idx64);
                    }
                }
            }
        }
        __syncthreads();
    }
}

template <typename T, typename Op>
int reduce_launch(int64_t m, int64_t n, int64_t k, const void* Av, int64_t offA, void* B, int64_t offB, int idx64)
{
    rb200::Context& c = rb200::ctx();
    const T* A = reinterpret_cast<const T*>(Av) + offA;
    constexpr int V = RdVec<T>::N;
    using Acc = typename Op::Acc;
    const int64_t target_ctas = (int64_t)c.sm_count * 8;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    // CTAs per SM actually resident for the split kernels: splitting to exactly one full wave avoids
    // a partially filled second wave (1184 CTAs on 888 slots cost 25% at axis-0 of [65536,1024])
    static int occ_rows = 0, occ_cols = 0;
    if (!occ_rows) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_rows, reduce_rows_block<T, Op, true>, RD_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_cols, reduce_cols<T, Op, true>, RD_THREADS, 0);
        if (occ_rows < 1) occ_rows = 4;
        if (occ_cols < 1) occ_cols = 4;
    }
    const int64_t wave_rows = (int64_t)c.sm_count * occ_rows, wave_cols = (int64_t)c.sm_count * occ_cols;
    if (n >= (int64_t)INT_MAX) RB_INVALID("reduced axis too long (n=%lld)", (long long)n);

    if (k == 1) {
        const bool vec = base_aligned && (n % V == 0);
        const int rows_per_cta = RD_THREADS / 32;
        if (n <= 4096 && m * 32 >= target_ctas * RD_THREADS / 4) {
            int64_t grid = rb_ceil_div(m, rows_per_cta);
            const int64_t cap = (int64_t)c.sm_count * 64;
            if (grid > cap) grid = cap;
            if (vec) reduce_rows_warp<T, Op, true><<<(unsigned)grid, RD_THREADS, 0, c.stream>>>(m, n, A, B, offB, idx64);
            else     reduce_rows_warp<T, Op, false><<<(unsigned)grid, RD_THREADS, 0, c.stream>>>(m, n, A, B, offB, idx64);
            RB_LAUNCH_CHECK("reduce_rows_warp");
            return RB200_OK;
        }
        // CTA per (row, segment)
        int64_t S = 1;
        const int64_t min_seg = (int64_t)RD_THREADS * V * 4;          // >= 4 vectors per thread
        if (m < target_ctas) {
            S = wave_rows / m;                       // floor: never more than one full wave
            if (S < 1) S = 1;
            const int64_t smax = rb_ceil_div(n, min_seg);
            if (S > smax) S = smax;
            if (S < 1) S = 1;
            if (S > 4096) S = 4096;
        }
        int64_t seglen = rb_ceil_div(n, S);
        seglen = rb_ceil_div(seglen, (int64_t)V) * V;                 // keep segments 16-byte aligned
        S = rb_ceil_div(n, seglen);
        Acc* partial = nullptr;
        int* counters = nullptr;
        if (S > 1) {
            int rc = rb200::ensure_workspace((size_t)(m * S) * sizeof(Acc));
            if (rc != RB200_OK) return rc;
            partial = reinterpret_cast<Acc*>(c.workspace);
            rc = rb200::ensure_counters(m, &counters);
            if (rc != RB200_OK) return rc;
        }
        dim3 grid((unsigned)S, (unsigned)(m < 65535 ? m : 65535));
        if (vec) reduce_rows_block<T, Op, true><<<grid, RD_THREADS, 0, c.stream>>>(m, n, (int)S, seglen, A, B, offB, idx64, partial, counters);
        else     reduce_rows_block<T, Op, false><<<grid, RD_THREADS, 0, c.stream>>>(m, n, (int)S, seglen, A, B, offB, idx64, partial, counters);
        RB_LAUNCH_CHECK("reduce_rows_block");
        return RB200_OK;
    }

    // k > 1
    const bool vec = base_aligned && (k % V == 0);
    const int64_t kv = vec ? k / V : k;
    // threads along k.  Unsplit (enough (row, column block) units to fill the SMs): 32, i.e. 512-byte row pieces per
    // warp.  When the reduced axis has to be split into row segments, the last CTA of each column block merges the
    // segments' partials, and that tail is on the critical path (ncu at axis 0 of [65536,1024]: SMs active 80 k of
    // 101 k elapsed cycles, the merging SMs 94 k): 8 threads along k give four times as many, four times smaller
    // merges -- 0.82 -> 0.87 of HBM peak there; 16: 0.83; 64 / 128 / 256: 0.67 / 0.41 / 0.16
    // (profiles/hbm_straggler_experiments_r02.txt).  RB200_REDUCE_COLS_TX forces a width.
    static const int tx_env = [] { const char* e = getenv("RB200_REDUCE_COLS_TX"); const int v = e ? atoi(e) : 0;
                                   return (v == 8 || v == 16 || v == 32 || v == 64 || v == 128 || v == 256) ? v : 0; }();
    auto threads_along_k = [&](int cap) { int t = 1; while (t < cap && t < kv) t <<= 1; return t; };
    int tx = threads_along_k(32);
    if (tx_env) tx = threads_along_k(tx_env);
    else if (m * rb_ceil_div(kv, (int64_t)tx) < target_ctas) tx = threads_along_k(8);
    const int ty = RD_THREADS / tx;
    const int64_t col_blocks = rb_ceil_div(kv, tx);
    int64_t S = 1;
    if (m * col_blocks < target_ctas) {
        S = wave_cols / (m * col_blocks);            // floor: one full wave
        if (S < 1) S = 1;
        const int64_t smax = rb_ceil_div(n, (int64_t)ty * 4);
        if (S > smax) S = smax;
        if (S < 1) S = 1;
        if (S > 1024) S = 1024;
    }
    int64_t tlen = rb_ceil_div(n, S);
    S = rb_ceil_div(n, tlen);
    Acc* partial = nullptr;
    int* counters = nullptr;
    if (S > 1) {
        int rc = rb200::ensure_workspace((size_t)(m * S * k) * sizeof(Acc));
        if (rc != RB200_OK) return rc;
        partial = reinterpret_cast<Acc*>(c.workspace);
        rc = rb200::ensure_counters(m * col_blocks, &counters);
        if (rc != RB200_OK) return rc;
    }
    const int64_t units = m * S;
    dim3 block(tx, ty);
    dim3 grid((unsigned)col_blocks, (unsigned)(units < 65535 ? units : 65535));
    if (col_blocks > 0x7fffffffLL) RB_INVALID("k too large");
    if (vec) reduce_cols<T, Op, true><<<grid, block, 0, c.stream>>>(m, n, k, (int)S, tlen, A, B, offB, idx64, partial, counters);
    else     reduce_cols<T, Op, false><<<grid, block, 0, c.stream>>>(m, n, k, (int)S, tlen, A, B, offB, idx64, partial, counters);
    RB_LAUNCH_CHECK("reduce_cols");
    return RB200_OK;
}

int reduce_check(int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, const void* B, int64_t offB)
{
    RB_REQUIRE_INIT();
    if (!A || !B) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1 || k < 1) RB_INVALID("m, n and k must be greater than 0");
    if (offA < 0 || offB < 0) RB_INVALID("negative offset");
    return RB200_OK;
}

}  // namespace

extern "C" {

int rb200_reduce_sum(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, void* B, int64_t offB)
{
    int rc = reduce_check(m, n, k, A, offA, B, offB);
    if (rc != RB200_OK) return rc;
    switch (dtype) {
        case RB200_FLOAT32: return reduce_launch<float, SumOp<float>>(m, n, k, A, offA, B, offB, 0);
        case RB200_FLOAT64: return reduce_launch<double, SumOp<double>>(m, n, k, A, offA, B, offB, 0);
        default: RB_UNSUPPORTED("reduceSum: unsupported dtype %d", dtype);
    }
}

int rb200_reduce_max(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, void* B, int64_t offB)
{
    int rc = reduce_check(m, n, k, A, offA, B, offB);
    if (rc != RB200_OK) return rc;
    switch (dtype) {
        case RB200_FLOAT32: return reduce_launch<float, MaxOp<float>>(m, n, k, A, offA, B, offB, 0);
        case RB200_FLOAT64: return reduce_launch<double, MaxOp<double>>(m, n, k, A, offA, B, offB, 0);
        default: RB_UNSUPPORTED("reduceMax: unsupported dtype %d", dtype);
    }
}

int rb200_reduce_argmax(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA,
                        int index_dtype, void* B, int64_t offB)
{
    int rc = reduce_check(m, n, k, A, offA, B, offB);
    if (rc != RB200_OK) return rc;
    int idx64;
    switch (index_dtype) {
        case RB200_INT32: case RB200_UINT32: idx64 = 0; break;
        case RB200_INT64: case RB200_UINT64: idx64 = 1; break;
        default: RB_UNSUPPORTED("reduceArgMax: index dtype must be int32/uint32/int64/uint64 (got %d)", index_dtype);
    }
    switch (dtype) {
        case RB200_FLOAT32: return reduce_launch<float, ArgMaxOp<float>>(m, n, k, A, offA, B, offB, idx64);
        case RB200_FLOAT64: return reduce_launch<double, ArgMaxOp<double>>(m, n, k, A, offA, B, offB, idx64);
        default: RB_UNSUPPORTED("reduceArgMax: unsupported dtype %d", dtype);
    }
}

int rb200_reduce_argmax_partial(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA,
                                int stripe_starts_at_global_0, void* records, int64_t off_records)
{
    int rc = reduce_check(m, n, k, A, offA, records, off_records);
    if (rc != RB200_OK) return rc;
    if ((reinterpret_cast<uintptr_t>(records) & 15) != 0) RB_INVALID("records must be 16-byte aligned");
    // B is indexed in records by the kernels: offB counts records
    if (stripe_starts_at_global_0) {
        switch (dtype) {
            case RB200_FLOAT32: return reduce_launch<float, ArgMaxPairOp<float, true>>(m, n, k, A, offA, records, off_records, 0);
            case RB200_FLOAT64: return reduce_launch<double, ArgMaxPairOp<double, true>>(m, n, k, A, offA, records, off_records, 0);
            default: RB_UNSUPPORTED("reduceArgMax: unsupported dtype %d", dtype);
        }
    }
    switch (dtype) {
        case RB200_FLOAT32: return reduce_launch<float, ArgMaxPairOp<float, false>>(m, n, k, A, offA, records, off_records, 0);
        case RB200_FLOAT64: return reduce_launch<double, ArgMaxPairOp<double, false>>(m, n, k, A, offA, records, off_records, 0);
        default: RB_UNSUPPORTED("reduceArgMax: unsupported dtype %d", dtype);
    }
}

}  // extern "C"
