// gemm_simt.cu -- FFMA (float32) / DFMA (float64) GEMM for every shape the tensor-core path
// cannot take: arbitrary m/n/k, leading dimensions, element offsets and transposes (the
// reference tests use ld = 3 and offset = 9, MatrixOperatorTest.php:308-318), and float64.
//
// Reference semantics (src/Drivers/MatlibPHP/PhpBlas.php:849-948):
//   C[i,j] = alpha * sum_k opA[i,k]*opB[k,j]  (+ beta*C[i,j] iff beta != 0; C not read otherwise)
// Each thread accumulates its 4x4 micro-tile with k ascending (fma chain), i.e. the same
// summation order as the reference's inner loop, in the operand precision.
//
// 64x64 CTA tile, BK = 16, 256 threads, shared-memory staging with the global loads mapped
// along whichever operand dimension is contiguous (coalesced for both NoTrans and Trans).
#include "common.cuh"

namespace {

constexpr int GS_BM = 64, GS_BN = 64, GS_BK = 16, GS_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, T alpha,
                 const T* __restrict__ A, int64_t sAm, int64_t sAk,
                 const T* __restrict__ B, int64_t sBk, int64_t sBn,
                 T beta, T* __restrict__ C, int64_t ldC)
{
    __shared__ T As[GS_BK][GS_BM + 1];
    __shared__ T Bs[GS_BK][GS_BN + 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4x4 outputs each
    const int64_t m0 = (int64_t)blockIdx.y * GS_BM, n0 = (int64_t)blockIdx.x * GS_BN;

    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = (T)0;

    for (int64_t k0 = 0; k0 < K; k0 += GS_BK) {
        // A tile: GS_BM x GS_BK.  If k is the contiguous dim (sAk == 1) let consecutive threads
        // walk k, otherwise walk m.
#pragma unroll
        for (int it = 0; it < (GS_BM * GS_BK) / GS_THREADS; it++) {
            const int e = tid + it * GS_THREADS;
            int mi, ki;
            if (sAk == 1) { ki = e % GS_BK; mi = e / GS_BK; } else { mi = e % GS_BM; ki = e / GS_BM; }
            const int64_t gm = m0 + mi, gk = k0 + ki;
            As[ki][mi] = (gm < M && gk < K) ? A[gm * sAm + gk * sAk] : (T)0;
        }
#pragma unroll
        for (int it = 0; it < (GS_BN * GS_BK) / GS_THREADS; it++) {
            const int e = tid + it * GS_THREADS;
            int ni, ki;
            if (sBn == 1) { ni = e % GS_BN; ki = e / GS_BN; } else { ki = e % GS_BK; ni = e / GS_BK; }
            const int64_t gn = n0 + ni, gk = k0 + ki;
            Bs[ki][ni] = (gn < N && gk < K) ? B[gk * sBk + gn * sBn] : (T)0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GS_BK; kk++) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            T* c = C + gm * ldC + gn;
            *c = (beta == (T)0) ? alpha * acc[i][j] : fma(beta, *c, alpha * acc[i][j]);
        }
    }
}

}  // namespace

namespace rb200 {

template <typename T>
int gemm_simt(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk,
              const T* B, int64_t sBk, int64_t sBn, T beta, T* C, int64_t ldC)
{
    Context& c = ctx();
    const int64_t gx = rb_ceil_div(N, GS_BN), gy = rb_ceil_div(M, GS_BM);
    if (gy > 65535) {
        // very tall: run stripes of 65535*64 rows
        const int64_t stripe = 65535LL * GS_BM;
        for (int64_t r0 = 0; r0 < M; r0 += stripe) {
            const int64_t rows = (M - r0) < stripe ? (M - r0) : stripe;
            int rc = gemm_simt<T>(rows, N, K, alpha, A + r0 * sAm, sAm, sAk, B, sBk, sBn, beta, C + r0 * ldC, ldC);
            if (rc != RB200_OK) return rc;
        }
        return RB200_OK;
    }
    dim3 grid((unsigned)gx, (unsigned)gy);
    gemm_simt_kernel<T><<<grid, GS_THREADS, 0, c.stream>>>(M, N, K, alpha, A, sAm, sAk, B, sBk, sBn, beta, C, ldC);
    RB_LAUNCH_CHECK("gemm_simt_kernel");
    return RB200_OK;
}

template int gemm_simt<float>(int64_t, int64_t, int64_t, float, const float*, int64_t, int64_t,
                              const float*, int64_t, int64_t, float, float*, int64_t);
template int gemm_simt<double>(int64_t, int64_t, int64_t, double, const double*, int64_t, int64_t,
                               const double*, int64_t, int64_t, double, double*, int64_t);

}  // namespace rb200
