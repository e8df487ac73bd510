// softmax.cu -- in-place row softmax (HBM-bound).
//
// Reference semantics (src/Drivers/MatlibPHP/PhpMath.php:1645-1670), per row of A[m,n] (row
// stride ldA):  mx = max(x);  t_j = exp(x_j - mx);  s = sum_j t_j;  x_j = t_j / s.
// The reference's "Zero divide in softmax." branch (s == 0) is unreachable for n >= 1: the
// maximum element contributes exp(0) = 1, and a NaN/Inf row gives s = NaN, which compares
// unequal to 0 -- so the device path needs no error flag.  Any NaN or +Inf in a row makes
// the whole row NaN in both implementations (x - mx is NaN for at least the max element).
// Algorithmic bytes: 2*m*n*sizeof(T) (single pass: the row stays on chip).
//
//   softmax_warp<ITEMS> : one warp per row, the row lives in registers (n <= 128*ITEMS,
//                         128-bit loads/stores), shuffle reductions only -- the C3 config
//                         [65536,1024] runs here with ITEMS = 8.
//   softmax_block       : one CTA per row, any n / ldA / alignment; three passes over the
//                         row (max, exp-sum, normalise), the 2nd and 3rd served by L2.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int SM_THREADS = 256;

template <typename T> __device__ __forceinline__ T sm_exp(T x);
template <> __device__ __forceinline__ float  sm_exp<float>(float x)   { return expf(x); }
template <> __device__ __forceinline__ double sm_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T sm_neg_inf();
template <> __device__ __forceinline__ float  sm_neg_inf<float>()  { return -INFINITY; }
template <> __device__ __forceinline__ double sm_neg_inf<double>() { return -(double)INFINITY; }

template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, mask));
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int mask = 16; mask > 0; mask >>= 1) v += __shfl_xor_sync(0xffffffffu, v, mask);
    return v;
}

// float32 only: each lane holds ITEMS float4 chunks of the row: chunk index lane + 32*i.
template <int ITEMS>
__global__ void __launch_bounds__(SM_THREADS)
softmax_warp_f32(int64_t m, int n, float* __restrict__ A, int64_t ldA)
{
    const int lane = threadIdx.x & 31;
    const int nvec = n >> 2;
    const int64_t warps_per_grid = (int64_t)gridDim.x * (SM_THREADS / 32);
    for (int64_t row = (int64_t)blockIdx.x * (SM_THREADS / 32) + (threadIdx.x >> 5); row < m; row += warps_per_grid) {
        float4* row4 = reinterpret_cast<float4*>(A + row * ldA);
        float4 v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int c = lane + 32 * i;
            if (c < nvec) v[i] = ld_inplace_f4(row4 + c);
            else          v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
        // NaN handling: fmaxf drops NaN from the max, but exp(NaN - mx) = NaN poisons the sum,
        // so the row still comes out all-NaN like the reference.
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
        mx = warp_max(mx);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx);
            v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        s = warp_sum(s);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const int c = lane + 32 * i;
            if (c < nvec) {
                v[i].x = v[i].x / s; v[i].y = v[i].y / s; v[i].z = v[i].z / s; v[i].w = v[i].w / s;
                st_stream_f4(row4 + c, v[i]);
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS)
softmax_block(int64_t m, int64_t n, T* __restrict__ A, int64_t ldA)
{
    __shared__ T red[SM_THREADS / 32];
    __shared__ T bcast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
        T* x = A + row * ldA;
        T mx = sm_neg_inf<T>();
        for (int64_t j = threadIdx.x; j < n; j += SM_THREADS) mx = fmax(mx, x[j]);
        mx = warp_max(mx);
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (warp == 0) {
            T t = lane < SM_THREADS / 32 ? red[lane] : sm_neg_inf<T>();
            t = warp_max(t);
            if (lane == 0) bcast = t;
        }
        __syncthreads();
        mx = bcast;
        T s = 0;
        for (int64_t j = threadIdx.x; j < n; j += SM_THREADS) s += sm_exp<T>(x[j] - mx);
        s = warp_sum(s);
        __syncthreads();
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (warp == 0) {
            T t = lane < SM_THREADS / 32 ? red[lane] : (T)0;
            t = warp_sum(t);
            if (lane == 0) bcast = t;
        }
        __syncthreads();
        s = bcast;
        for (int64_t j = threadIdx.x; j < n; j += SM_THREADS) x[j] = sm_exp<T>(x[j] - mx) / s;
        __syncthreads();
    }
}

}  // namespace

extern "C" int rb200_softmax(int dtype, int64_t m, int64_t n, void* Av, int64_t offA, int64_t ldA)
{
    RB_REQUIRE_INIT();
    rb200::Context& c = rb200::ctx();
    if (!Av) RB_INVALID("NULL buffer");
    if (m < 1 || n < 1) RB_INVALID("m and n must be greater than 0");
    if (offA < 0 || ldA < n) RB_INVALID("bad offset / leading dimension");
    if (dtype == RB200_FLOAT32) {
        float* A = reinterpret_cast<float*>(Av) + offA;
        const bool vec = (n % 4 == 0) && (ldA % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
        if (vec && n <= 1024) {
            int64_t grid = rb_ceil_div(m, SM_THREADS / 32);
            const int64_t cap = (int64_t)c.sm_count * 64;
            if (grid > cap) grid = cap;
            const int nn = (int)n;
            if (n <= 128)      softmax_warp_f32<1><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, nn, A, ldA);
            else if (n <= 256) softmax_warp_f32<2><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, nn, A, ldA);
            else if (n <= 512) softmax_warp_f32<4><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, nn, A, ldA);
            else               softmax_warp_f32<8><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, nn, A, ldA);
            RB_LAUNCH_CHECK("softmax_warp_f32");
            return RB200_OK;
        }
        int64_t grid = m < (int64_t)c.sm_count * 16 ? m : (int64_t)c.sm_count * 16;
        softmax_block<float><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, n, A, ldA);
        RB_LAUNCH_CHECK("softmax_block");
        return RB200_OK;
    }
    if (dtype == RB200_FLOAT64) {
        double* A = reinterpret_cast<double*>(Av) + offA;
        int64_t grid = m < (int64_t)c.sm_count * 16 ? m : (int64_t)c.sm_count * 16;
        softmax_block<double><<<(unsigned)grid, SM_THREADS, 0, c.stream>>>(m, n, A, ldA);
        RB_LAUNCH_CHECK("softmax_block");
        return RB200_OK;
    }
    RB_UNSUPPORTED("softmax: unsupported dtype %d", dtype);
}
