// blas1.cu -- scal / axpy / copy (BLAS level 1) and gemv on the device (SURVEY.md section 8f, ranks 2-3:
// the callers either side of the hot path -- MatrixOperator::op / the tests' equalTest use scal/axpy/copy,
// MatrixOperator::cross with a rank-1 right operand uses gemv).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpBlas.php): scal :141-161  X[i] = X[i]*alpha;
// axpy :165-194  Y[i] = alpha*X[i] + Y[i];  copy :350-364  Y[i] = X[i];
// gemv :762-844  Y[i] = sum_j (alpha*opA[i,j])*X[j] (+ beta*Y[i] iff beta != 0).  All HBM-bound:
// 128-bit accesses when the increments are 1 and the operands 16-byte aligned, strided scalar otherwise.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int B1_THREADS = 256;
enum { B1_SCAL = 0, B1_AXPY = 1, B1_COPY = 2, B1_SWAP = 3 };     // swap (PhpBlas.php:744-760) also writes X

template <typename T> union B1Pack { int4 raw; T v[16 / sizeof(T)]; };

constexpr int B1_UNROLL = 4;     // 16-byte packs in flight per thread and operand

template <typename T, int OP>
__global__ void __launch_bounds__(B1_THREADS)
blas1_vec_kernel(int64_t nvec, T alpha, const T* X, T* Y)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t chunk = (int64_t)B1_THREADS * B1_UNROLL;
    for (int64_t base = (int64_t)blockIdx.x * chunk; base < nvec; base += (int64_t)gridDim.x * chunk) {
        B1Pack<T> x[B1_UNROLL], y[B1_UNROLL];
#pragma unroll
        for (int u = 0; u < B1_UNROLL; u++) {
            const int64_t i = base + u * B1_THREADS + threadIdx.x;
            if (i < nvec) {
                if (OP != B1_SCAL) x[u].raw = __ldcs(reinterpret_cast<const int4*>(X) + i);
                if (OP != B1_COPY) y[u].raw = __ldcs(reinterpret_cast<const int4*>(Y) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < B1_UNROLL; u++) {
            const int64_t i = base + u * B1_THREADS + threadIdx.x;
            if (i < nvec) {
                if (OP == B1_SWAP) {
                    __stcs(reinterpret_cast<int4*>(const_cast<T*>(X)) + i, y[u].raw);
                    __stcs(reinterpret_cast<int4*>(Y) + i, x[u].raw);
                } else {
#pragma unroll
                    for (int e = 0; e < V; e++) {
                        if (OP == B1_SCAL) y[u].v[e] = y[u].v[e] * alpha;
                        else if (OP == B1_AXPY) y[u].v[e] = fma(alpha, x[u].v[e], y[u].v[e]);
                        else y[u].v[e] = x[u].v[e];
                    }
                    __stcs(reinterpret_cast<int4*>(Y) + i, y[u].raw);
                }
            }
        }
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(B1_THREADS)
blas1_strided_kernel(int64_t n, T alpha, const T* X, int64_t incX, T* Y, int64_t incY)
{
    const int64_t stride = (int64_t)gridDim.x * B1_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * B1_THREADS + threadIdx.x; i < n; i += stride) {
        T* y = Y + i * incY;
        if (OP == B1_SCAL) *y = *y * alpha;
        else if (OP == B1_AXPY) *y = fma(alpha, X[i * incX], *y);
        else if (OP == B1_SWAP) { T* x = const_cast<T*>(X) + i * incX; const T t = *y; *y = *x; *x = t; }
        else *y = X[i * incX];
    }
}

template <typename T, int OP>
int blas1_launch(int64_t n, double alpha_d, const void* Xv, int64_t offX, int64_t incX, void* Yv, int64_t offY, int64_t incY)
{
    rb200::Context& c = rb200::ctx();
    const T* X = Xv ? reinterpret_cast<const T*>(Xv) + offX : nullptr;
    T* Y = reinterpret_cast<T*>(Yv) + offY;
    const T alpha = (T)alpha_d;
    constexpr int V = 16 / (int)sizeof(T);
    const bool aligned = ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) && (OP == B1_SCAL || (reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const int64_t cap = (int64_t)c.sm_count * 16;
    if (incY == 1 && (OP == B1_SCAL || incX == 1) && aligned && n >= V) {
        const int64_t nvec = n / V;
        int64_t grid = rb_ceil_div(nvec, (int64_t)B1_THREADS * B1_UNROLL);
        if (grid > cap) grid = cap;
        blas1_vec_kernel<T, OP><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>(nvec, alpha, X, Y);
        RB_LAUNCH_CHECK("blas1_vec_kernel");
        const int64_t done = nvec * V;
        if (done < n) {
            blas1_strided_kernel<T, OP><<<1, B1_THREADS, 0, c.stream>>>(n - done, alpha, X ? X + done : nullptr, 1, Y + done, 1);
            RB_LAUNCH_CHECK("blas1_strided_kernel");
        }
        return RB200_OK;
    }
    int64_t grid = rb_ceil_div(n, B1_THREADS);
    if (grid > cap) grid = cap;
    blas1_strided_kernel<T, OP><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>(n, alpha, X, incX, Y, incY);
    RB_LAUNCH_CHECK("blas1_strided_kernel");
    return RB200_OK;
}

template <int OP>
int blas1_dispatch(int dtype, int64_t n, double alpha, const void* X, int64_t offX, int64_t incX,
                   void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!Y || (OP != B1_SCAL && !X)) RB_INVALID("NULL buffer");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (incY < 1 || (OP != B1_SCAL && incX < 1) || offX < 0 || offY < 0) RB_INVALID("bad increment / offset");
    switch (dtype) {
        case RB200_FLOAT32: return blas1_launch<float, OP>(n, alpha, X, offX, incX, Y, offY, incY);
        case RB200_FLOAT64: return blas1_launch<double, OP>(n, alpha, X, offX, incX, Y, offY, incY);
        // complex buffers: pure data movement only (copy / swap), as 8-byte words
        case RB200_COMPLEX64:
            if (OP == B1_COPY || OP == B1_SWAP) return blas1_launch<double, OP>(n, alpha, X, offX, incX, Y, offY, incY);
            break;
        case RB200_COMPLEX128:
            if ((OP == B1_COPY || OP == B1_SWAP) && incX == 1 && incY == 1)
                return blas1_launch<double, OP>(2 * n, alpha, X, 2 * offX, 1, Y, 2 * offY, 1);
            break;
        default: break;
    }
    if (dtype == RB200_COMPLEX64 || dtype == RB200_COMPLEX128) {
        rb200::set_error("blas level 1: complex buffers support (contiguous) copy and swap only");
        return RB200_E_UNSUPPORTED;
    }
    rb200::set_error("blas level 1: unsupported dtype %d", dtype);
    return RB200_E_UNSUPPORTED;
}

// ---- gemv -------------------------------------------------------------------------------------------
// NoTrans: SPLIT warps per row of A (coalesced along the row, shuffle reduction, shared-memory add of the SPLIT
// parts in index order; SPLIT > 1 only for matrices with too few rows to occupy the SMs).  Compiled for 8
// CTAs per SM (<= 32 registers) so that the 8192 rows of the bench shape are all resident in ONE wave: with 4
// CTAs per SM the same launch is 1.73 waves and pays the tail.  VEC: 16-byte loads of A and X (ldA, the
// offsets and the base pointers 16-byte aligned, incX == 1), four packs in flight per lane.
constexpr int GEMV_PACKS = 2;     // 16-byte packs of A in flight per lane (3 also fits 32 registers; measured the same: 24.8 vs 25.0 us)

template <typename T, bool VEC, int SPLIT, int PACKS>
__global__ void __launch_bounds__(B1_THREADS, SPLIT == 1 ? 8 : 6)
gemv_n_kernel(int64_t m, int64_t n, int64_t part_cols, T alpha, const T* __restrict__ A, int64_t ldA,
              const T* __restrict__ X, int64_t incX, T beta, T* __restrict__ Y, int64_t incY)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int WARPS = B1_THREADS / 32;
    constexpr int RPC = WARPS / SPLIT;              // rows per CTA-unit
    __shared__ T parts[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_local = warp / SPLIT, part = warp % SPLIT;
    const int64_t c_begin = (int64_t)part * part_cols;
    const int64_t c_end = (c_begin + part_cols < n) ? c_begin + part_cols : n;
    const int64_t units = (m + RPC - 1) / RPC;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t i = u * RPC + row_local;
        T acc = (T)0;
        if (i < m && c_begin < c_end) {
            const T* row = A + i * ldA;
            int64_t j0 = c_begin;
            if (VEC) {                               // part_cols is a multiple of V on this path
                const int64_t v0 = c_begin / V, v1 = c_end / V;
                const int4* rv = reinterpret_cast<const int4*>(row);
                const int4* xv = reinterpret_cast<const int4*>(X);
                for (int64_t jb = v0; jb < v1; jb += 32 * PACKS) {
                    B1Pack<T> a[PACKS];
#pragma unroll
                    for (int q = 0; q < PACKS; q++) {
                        const int64_t jv = jb + q * 32 + lane;
                        if (jv < v1) a[q].raw = __ldcs(rv + jv);
                    }
#pragma unroll
                    for (int q = 0; q < PACKS; q++) {
                        const int64_t jv = jb + q * 32 + lane;
                        if (jv < v1) {
                            B1Pack<T> x;
                            x.raw = __ldg(xv + jv);
#pragma unroll
                            for (int e = 0; e < V; e++) acc = fma(alpha * a[q].v[e], x.v[e], acc);
                        }
                    }
                }
                j0 = v1 * V;
            }
#pragma unroll 4
            for (int64_t j = j0 + lane; j < c_end; j += 32) acc = fma(alpha * __ldcs(row + j), __ldg(X + j * incX), acc);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
        if (SPLIT == 1) {
            if (lane == 0 && i < m) {
                T* y = Y + i * incY;
                *y = (beta == (T)0) ? acc : fma(beta, *y, acc);
            }
        } else {
            if (lane == 0) parts[warp] = acc;
            __syncthreads();
            if (part == 0 && lane == 0 && i < m) {
                T sum = parts[warp];
#pragma unroll
                for (int q = 1; q < SPLIT; q++) sum += parts[warp + q];
                T* y = Y + i * incY;
                *y = (beta == (T)0) ? sum : fma(beta, *y, sum);
            }
            __syncthreads();
        }
    }
}

// Trans: Y[j] = sum_i (alpha*A[i,j])*X[i]; thread (tx,ty) owns CPT adjacent columns (CPT = 16 bytes worth when
// A's rows are 16-byte aligned, else 1) of column block blockIdx.x and rows ty, ty+8, ... of the row segment
// blockIdx.y (coalesced along j); shared-memory tree over ty.  With one segment the CTA finishes Y itself,
// otherwise it writes a partial row into the workspace and the last CTA of the column block to finish (ticket)
// adds the segments in a fixed order -- one launch, deterministic.
constexpr int GEMV_T_PACKS = 4;

template <typename T, int CPT, int TPACKS>
__global__ void __launch_bounds__(B1_THREADS)
gemv_t_kernel(int64_t m, int64_t n, int64_t seg_rows, T alpha, const T* __restrict__ A, int64_t ldA,
              const T* __restrict__ X, int64_t incX, T beta, T* __restrict__ Y, int64_t incY, T* __restrict__ partial,
              int* __restrict__ counters)
{
    __shared__ T sm[8][32 * CPT + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t j = ((int64_t)blockIdx.x * 32 + tx) * CPT;
    const int64_t i0 = (int64_t)blockIdx.y * seg_rows;
    const int64_t i1 = (i0 + seg_rows < m) ? i0 + seg_rows : m;
    T acc[CPT];
#pragma unroll
    for (int e = 0; e < CPT; e++) acc[e] = (T)0;
    if (CPT == 1) {
        if (j < n) {
            const T* a = A + j;
#pragma unroll 4
            for (int64_t i = i0 + ty; i < i1; i += 8) acc[0] = fma(alpha * __ldcs(a + i * ldA), __ldg(X + i * incX), acc[0]);
        }
    } else if (j < n) {                    // n % CPT == 0 on this path: a pack never straddles the edge
        const T* a = A + j;
        for (int64_t ib = i0 + ty; ib < i1; ib += 8 * TPACKS) {
            B1Pack<T> v[TPACKS];
#pragma unroll
            for (int u = 0; u < TPACKS; u++) {
                const int64_t i = ib + u * 8;
                if (i < i1) v[u].raw = __ldcs(reinterpret_cast<const int4*>(a + i * ldA));
            }
#pragma unroll
            for (int u = 0; u < TPACKS; u++) {
                const int64_t i = ib + u * 8;
                if (i < i1) {
                    const T x = __ldg(X + i * incX);
#pragma unroll
                    for (int e = 0; e < CPT; e++) acc[e] = fma(alpha * v[u].v[e], x, acc[e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < CPT; e++) sm[ty][tx * CPT + e] = acc[e];
    __syncthreads();
    // 32*CPT columns finished by the first 32*CPT threads
    for (int c = threadIdx.x; c < 32 * CPT; c += B1_THREADS) {
        const int64_t jc = (int64_t)blockIdx.x * 32 * CPT + c;
        if (jc < n) {
            T s = sm[0][c];
#pragma unroll
            for (int r = 1; r < 8; r++) s += sm[r][c];
            if (partial) partial[(int64_t)blockIdx.y * n + jc] = s;
            else {
                T* y = Y + jc * incY;
                *y = (beta == (T)0) ? s : fma(beta, *y, s);
            }
        }
    }
    if (!partial) return;
    // the last CTA of this column block to finish adds the segment partials in a fixed order (one launch,
    // deterministic): thread (c, r) sums segments r, r+8, ...; shared-memory tree over r
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* ctr = counters + blockIdx.x;
        const int ticket = atomicAdd(ctr, 1);
        is_last = (ticket == (int)gridDim.y - 1);
        if (is_last) *ctr = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int segs = (int)gridDim.y;
    for (int c0 = 0; c0 < 32 * CPT; c0 += 32) {
        const int c = c0 + tx;
        const int64_t jc = (int64_t)blockIdx.x * 32 * CPT + c;
        T a = (T)0;
        if (jc < n) {
            for (int g = ty; g < segs; g += 8) a += __ldcg(partial + (int64_t)g * n + jc);
        }
        sm[ty][c] = a;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * CPT; c += B1_THREADS) {
        const int64_t jc = (int64_t)blockIdx.x * 32 * CPT + c;
        if (jc < n) {
            T s2 = sm[0][c];
#pragma unroll
            for (int r = 1; r < 8; r++) s2 += sm[r][c];
            T* y = Y + jc * incY;
            *y = (beta == (T)0) ? s2 : fma(beta, *y, s2);
        }
    }
}

template <typename T>
int gemv_launch(bool trans, int64_t m, int64_t n, double alpha, const void* Av, int64_t offA, int64_t ldA,
                const void* Xv, int64_t offX, int64_t incX, double beta, void* Yv, int64_t offY, int64_t incY)
{
    rb200::Context& c = rb200::ctx();
    constexpr int V = 16 / (int)sizeof(T);
    const T* A = reinterpret_cast<const T*>(Av) + offA;
    const T* X = reinterpret_cast<const T*>(Xv) + offX;
    T* Y = reinterpret_cast<T*>(Yv) + offY;
    const bool a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (ldA % V == 0);
    if (!trans) {
        const bool vec = a_vec && incX == 1 && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && n >= V;
        static int resident = 0;
        if (!resident) {
            int per_sm = 0;
            RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemv_n_kernel<T, true, 1, GEMV_PACKS>, B1_THREADS, 0));
            resident = (per_sm > 0 ? per_sm : 1) * c.sm_count;
        }
        // one warp per row; rows are split over 2/4/8 warps only when there are too few of them to occupy the SMs
        int split = 1;
        while (split < 8 && n / (split * 2) >= 512 && rb_ceil_div(m * split, B1_THREADS / 32) * 2 <= resident) split *= 2;
        const int64_t units = rb_ceil_div(m * split, B1_THREADS / 32);
        int64_t grid = units < resident ? units : resident;
        int64_t part_cols = rb_ceil_div(rb_ceil_div(n, split), 128) * 128;
        static const int n_packs = getenv("RB200_GEMV_PACKS") ? atoi(getenv("RB200_GEMV_PACKS")) : GEMV_PACKS;
#define RB_GEMV_N(VEC_, SP_) do { if (n_packs == 2) gemv_n_kernel<T, VEC_, SP_, 2><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>( \
            m, n, part_cols, (T)alpha, A, ldA, X, incX, (T)beta, Y, incY); \
        else gemv_n_kernel<T, VEC_, SP_, 3><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>( \
            m, n, part_cols, (T)alpha, A, ldA, X, incX, (T)beta, Y, incY); } while (0)
        if (vec) { if (split == 1) RB_GEMV_N(true, 1); else if (split == 2) RB_GEMV_N(true, 2);
                   else if (split == 4) RB_GEMV_N(true, 4); else RB_GEMV_N(true, 8); }
        else     { if (split == 1) RB_GEMV_N(false, 1); else if (split == 2) RB_GEMV_N(false, 2);
                   else if (split == 4) RB_GEMV_N(false, 4); else RB_GEMV_N(false, 8); }
#undef RB_GEMV_N
        RB_LAUNCH_CHECK("gemv_n_kernel");
    } else {
        const bool vec = a_vec && (n % V == 0);
        const int64_t gx = rb_ceil_div(n, vec ? 32 * V : 32);
        if (gx > 0x7fffffffLL) RB_INVALID("gemv: n too large");
        // row segments: one full wave of resident CTAs, at least 64 rows each
        static int t_resident = 0;
        if (!t_resident) {
            int per_sm = 0;
            RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemv_t_kernel<T, V, GEMV_T_PACKS>, B1_THREADS, 0));
            t_resident = (per_sm > 0 ? per_sm : 1) * c.sm_count;
        }
        int64_t segs = t_resident / gx;
        if (segs > m / 64) segs = m / 64;
        if (segs < 1) segs = 1;
        if (segs > 1024) segs = 1024;
        const int64_t seg_rows = rb_ceil_div(m, segs);
        segs = rb_ceil_div(m, seg_rows);
        T* partial = nullptr;
        int* counters = nullptr;
        if (segs > 1) {
            int rc = rb200::ensure_workspace((size_t)segs * (size_t)n * sizeof(T));
            if (rc != RB200_OK) return rc;
            partial = reinterpret_cast<T*>(c.workspace);
            rc = rb200::ensure_counters(gx, &counters);
            if (rc != RB200_OK) return rc;
        }
        const dim3 grid((unsigned)gx, (unsigned)segs);
        static const int t_packs = getenv("RB200_GEMV_T_PACKS") ? atoi(getenv("RB200_GEMV_T_PACKS")) : GEMV_T_PACKS;
        if (vec && t_packs >= 8) gemv_t_kernel<T, V, 8><<<grid, B1_THREADS, 0, c.stream>>>(m, n, seg_rows, (T)alpha, A, ldA, X, incX, (T)beta, Y, incY, partial, counters);
        else if (vec) gemv_t_kernel<T, V, 4><<<grid, B1_THREADS, 0, c.stream>>>(m, n, seg_rows, (T)alpha, A, ldA, X, incX, (T)beta, Y, incY, partial, counters);
        else gemv_t_kernel<T, 1, 4><<<grid, B1_THREADS, 0, c.stream>>>(m, n, seg_rows, (T)alpha, A, ldA, X, incX, (T)beta, Y, incY, partial, counters);
        RB_LAUNCH_CHECK("gemv_t_kernel");
    }
    return RB200_OK;
}

// ---- omatcopy ---------------------------------------------------------------------------------------
// B[rows,cols] = alpha * op(A); Trans goes through a 32x33 shared tile so both sides stay coalesced; each CTA
// (32x8 threads) walks tiles in a grid-stride loop.
template <typename T, bool TRANS>
__global__ void __launch_bounds__(B1_THREADS)
omatcopy_kernel(int64_t rows, int64_t cols, T alpha, const T* __restrict__ A, int64_t ldA, T* __restrict__ B, int64_t ldB)
{
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
    for (int64_t t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
        const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
        if (TRANS) {
            // A is cols x rows (row-major, ldA): read A[c0+.., r0+tx] coalesced along r
#pragma unroll
            for (int y = ty; y < 32; y += 8) {
                const int64_t ar = c0 + y, ac = r0 + tx;
                if (ar < cols && ac < rows) tile[y][tx] = __ldcs(A + ar * ldA + ac);
            }
            __syncthreads();
#pragma unroll
            for (int y = ty; y < 32; y += 8) {
                const int64_t br = r0 + y, bc = c0 + tx;
                if (br < rows && bc < cols) __stcs(B + br * ldB + bc, alpha * tile[tx][y]);
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int y = ty; y < 32; y += 8) {
                const int64_t br = r0 + y, bc = c0 + tx;
                if (br < rows && bc < cols) __stcs(B + br * ldB + bc, alpha * __ldcs(A + br * ldA + bc));
            }
        }
    }
}

// Transposing copy with 16-byte global accesses on both sides: 64x64 tile, 256 threads, all of a thread's
// loads issued before its first shared store.  Needs 16-byte aligned A/B rows and rows, cols multiples of the
// pack width; everything else takes omatcopy_kernel.
template <typename T>
__global__ void __launch_bounds__(B1_THREADS)
omatcopy_t_vec_kernel(int64_t rows, int64_t cols, T alpha, const T* __restrict__ A, int64_t ldA,
                      T* __restrict__ B, int64_t ldB)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int TS = 64;                       // tile edge in elements
    constexpr int VPR = TS / V;                  // packs per tile row
    constexpr int RPP = B1_THREADS / VPR;        // tile rows per pass
    constexpr int PASSES = TS / RPP;
    __shared__ T tile[TS][TS + 1];
    const int xv = threadIdx.x % VPR, y0 = threadIdx.x / VPR;
    const int64_t tiles_c = (cols + TS - 1) / TS, tiles_r = (rows + TS - 1) / TS;
    for (int64_t t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
        const int64_t r0 = (t / tiles_c) * TS, c0 = (t % tiles_c) * TS;
        B1Pack<T> v[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
            const int64_t ar = c0 + y0 + p * RPP, ac = r0 + xv * V;      // A is cols x rows
            if (ar < cols && ac < rows) v[p].raw = __ldcs(reinterpret_cast<const int4*>(A + ar * ldA + ac));
        }
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
#pragma unroll
            for (int e = 0; e < V; e++) tile[y0 + p * RPP][xv * V + e] = v[p].v[e];
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < PASSES; p++) {
            const int y = y0 + p * RPP;
            const int64_t br = r0 + y, bc = c0 + xv * V;
            if (br < rows && bc < cols) {
                B1Pack<T> o;
#pragma unroll
                for (int e = 0; e < V; e++) o.v[e] = alpha * tile[xv * V + e][y];
                __stcs(reinterpret_cast<int4*>(B + br * ldB + bc), o.raw);
            }
        }
        __syncthreads();
    }
}

template <typename T>
int omatcopy_launch(bool trans, int64_t rows, int64_t cols, double alpha, const void* Av, int64_t offA, int64_t ldA,
                    void* Bv, int64_t offB, int64_t ldB)
{
    rb200::Context& c = rb200::ctx();
    const T* A = reinterpret_cast<const T*>(Av) + offA;
    T* B = reinterpret_cast<T*>(Bv) + offB;
    constexpr int V = 16 / (int)sizeof(T);
    const int64_t cap = (int64_t)c.sm_count * 32;
    if (trans && rows % V == 0 && cols % V == 0 && ldA % V == 0 && ldB % V == 0 &&
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0) {
        int64_t grid = rb_ceil_div(rows, 64) * rb_ceil_div(cols, 64);
        if (grid > cap) grid = cap;
        omatcopy_t_vec_kernel<T><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>(rows, cols, (T)alpha, A, ldA, B, ldB);
        RB_LAUNCH_CHECK("omatcopy_t_vec_kernel");
        return RB200_OK;
    }
    int64_t grid = rb_ceil_div(rows, 32) * rb_ceil_div(cols, 32);
    if (grid > cap) grid = cap;
    if (trans) omatcopy_kernel<T, true><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>(rows, cols, (T)alpha, A, ldA, B, ldB);
    else omatcopy_kernel<T, false><<<(unsigned)grid, B1_THREADS, 0, c.stream>>>(rows, cols, (T)alpha, A, ldA, B, ldB);
    RB_LAUNCH_CHECK("omatcopy_kernel");
    return RB200_OK;
}

}  // namespace

extern "C" {

int rb200_scal(int dtype, int64_t n, double alpha, void* X, int64_t offX, int64_t incX)
{
    return blas1_dispatch<B1_SCAL>(dtype, n, alpha, nullptr, 0, 1, X, offX, incX);
}

int rb200_axpy(int dtype, int64_t n, double alpha, const void* X, int64_t offX, int64_t incX,
               void* Y, int64_t offY, int64_t incY)
{
    return blas1_dispatch<B1_AXPY>(dtype, n, alpha, X, offX, incX, Y, offY, incY);
}

int rb200_copy(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY)
{
    return blas1_dispatch<B1_COPY>(dtype, n, 1.0, X, offX, incX, Y, offY, incY);
}

int rb200_swap(int dtype, int64_t n, void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY)
{
    return blas1_dispatch<B1_SWAP>(dtype, n, 1.0, X, offX, incX, Y, offY, incY);
}

int rb200_gemv(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
               const void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX,
               double beta, void* Y, int64_t offY, int64_t incY)
{
    RB_REQUIRE_INIT();
    if (!A || !X || !Y) RB_INVALID("NULL buffer");
    if (order == RB200_COL_MAJOR) { int64_t t = m; m = n; n = t; }
    else if (order != RB200_ROW_MAJOR) RB_INVALID("Invalid Order type");
    bool t;
    switch (trans) {
        case RB200_NO_TRANS: case RB200_CONJ_NO_TRANS: t = false; break;
        case RB200_TRANS: case RB200_CONJ_TRANS: t = true; break;
        default: RB_INVALID("Unknown Tranpose Code: %d", trans);
    }
    if (m < 1) RB_INVALID("Argument m must be greater than 0.");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (ldA < 1 || incX < 1 || incY < 1 || offA < 0 || offX < 0 || offY < 0) RB_INVALID("bad ld / increment / offset");
    switch (dtype) {
        case RB200_FLOAT32: return gemv_launch<float>(t, m, n, alpha, A, offA, ldA, X, offX, incX, beta, Y, offY, incY);
        case RB200_FLOAT64: return gemv_launch<double>(t, m, n, alpha, A, offA, ldA, X, offX, incX, beta, Y, offY, incY);
        default: RB_UNSUPPORTED("gemv: unsupported dtype %d", dtype);
    }
}

int rb200_omatcopy(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
                   const void* A, int64_t offA, int64_t ldA, void* B, int64_t offB, int64_t ldB)
{
    RB_REQUIRE_INIT();
    if (!A || !B) RB_INVALID("NULL buffer");
    if (order == RB200_COL_MAJOR) { int64_t t = m; m = n; n = t; }
    else if (order != RB200_ROW_MAJOR) RB_INVALID("Invalid Order type");
    bool t;
    switch (trans) {
        case RB200_NO_TRANS: case RB200_CONJ_NO_TRANS: t = false; break;
        case RB200_TRANS: case RB200_CONJ_TRANS: t = true; break;
        default: RB_INVALID("Unknown Tranpose Code: %d", trans);
    }
    if (m < 1) RB_INVALID("Argument m must be greater than 0.");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (ldA < 1 || ldB < 1 || offA < 0 || offB < 0) RB_INVALID("bad ld / offset");
    const int64_t rows = t ? n : m, cols = t ? m : n;
    switch (dtype) {
        case RB200_FLOAT32: return omatcopy_launch<float>(t, rows, cols, alpha, A, offA, ldA, B, offB, ldB);
        case RB200_FLOAT64: return omatcopy_launch<double>(t, rows, cols, alpha, A, offA, ldA, B, offB, ldB);
        default: RB_UNSUPPORTED("omatcopy: unsupported dtype %d", dtype);
    }
}

}  // extern "C"
