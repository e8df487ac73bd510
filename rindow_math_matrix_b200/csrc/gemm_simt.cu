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
                 T beta, T* __restrict__ C, int64_t ldC,
                 int64_t strideA, int64_t strideB, int64_t strideC)
{
    // blockIdx.z = matrix of a strided batch (strides in elements; 0 broadcasts an operand)
    A += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
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

// ---- short-K kernel (K <= 64, NoTrans/NoTrans, float32): the conv-as-GEMM shape of BASELINE config C4
// ([B*oh*ow, kh*kw*C] x [kh*kw*C, filters], K = 27, N = 64).  HBM-bound: the whole K extent of a
// 128-row A tile and of a 64-column B tile sits in shared memory, each thread owns a 4 x 16 register
// tile (64 FFMA per 4 scalar + 4 vector shared loads), k ascending as in the reference loop.
constexpr int SK_BM = 128, SK_BN = 64, SK_THREADS = 128, SK_MAXK = 64, SK_MLP = 9;

__global__ void __launch_bounds__(SK_THREADS)
gemm_shortk_kernel(int64_t M, int N, int K, float alpha,
                   const float* __restrict__ A, int64_t ldA,
                   const float* __restrict__ B, int64_t ldB,
                   float beta, float* __restrict__ C, int64_t ldC, unsigned k_magic)
{
    extern __shared__ __align__(16) float sk_smem[];
    const int KS = K | 1;                                  // odd row pitch: conflict-free column reads
    float* As = sk_smem;                                   // [SK_BM][KS]
    float* Bs = sk_smem + ((SK_BM * KS + 3) & ~3);         // [K][SK_BN], 16-byte aligned
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * SK_BM;
    const int n0 = blockIdx.y * SK_BN;
    const int rows = (int)min((int64_t)SK_BM, M - m0);

    for (int e0 = 0; e0 < K * SK_BN; e0 += SK_MLP * SK_THREADS) {
        float v[SK_MLP];
#pragma unroll
        for (int u = 0; u < SK_MLP; u++) {
            const int e = e0 + u * SK_THREADS + tid;
            const int kk = e / SK_BN, j = e - kk * SK_BN;
            v[u] = (e < K * SK_BN && n0 + j < N) ? __ldg(B + (int64_t)kk * ldB + n0 + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < SK_MLP; u++) {
            const int e = e0 + u * SK_THREADS + tid;
            if (e < K * SK_BN) Bs[e] = v[u];
        }
    }
    // A tile: SK_MLP independent loads in flight per thread before the first shared-memory store
    const float* Ablk = A + m0 * ldA;
    const int total = SK_BM * K;
    for (int e0 = 0; e0 < total; e0 += SK_MLP * SK_THREADS) {
        float v[SK_MLP];
#pragma unroll
        for (int u = 0; u < SK_MLP; u++) {
            const int e = e0 + u * SK_THREADS + tid;
            v[u] = 0.f;
            if (e < total) {
                const int r = K == 1 ? e : (int)__umulhi((unsigned)e, k_magic);
                const int kk = e - r * K;
                if (r < rows) v[u] = __ldcs(Ablk + (int64_t)r * ldA + kk);
            }
        }
#pragma unroll
        for (int u = 0; u < SK_MLP; u++) {
            const int e = e0 + u * SK_THREADS + tid;
            if (e < total) {
                const int r = K == 1 ? e : (int)__umulhi((unsigned)e, k_magic);
                As[r * KS + (e - r * K)] = v[u];
            }
        }
    }
    __syncthreads();

    const int rg = tid >> 2, cg = tid & 3;                 // 32 row groups x 4 column groups
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) acc[i][j] = 0.f;
    const float* arow = As + (rg * 4) * KS;
    const float4* brow = reinterpret_cast<const float4*>(Bs) + cg * 4;
    for (int kk = 0; kk < K; kk++) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = arow[i * KS + kk];
        float bv[16];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const float4 t = brow[kk * (SK_BN / 4) + v];
            bv[4 * v] = t.x; bv[4 * v + 1] = t.y; bv[4 * v + 2] = t.z; bv[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 16; j++) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    const bool vec_ok = ((ldC & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (n0 + SK_BN <= N);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = rg * 4 + i;
        if (r >= rows) continue;
        float* crow = C + (m0 + r) * ldC + n0 + cg * 16;
        if (vec_ok) {
#pragma unroll
            for (int v = 0; v < 4; v++) {
                float4 o;
                o.x = alpha * acc[i][4 * v]; o.y = alpha * acc[i][4 * v + 1];
                o.z = alpha * acc[i][4 * v + 2]; o.w = alpha * acc[i][4 * v + 3];
                float4* dst = reinterpret_cast<float4*>(crow) + v;
                if (beta != 0.f) {
                    const float4 old = *dst;
                    o.x = fmaf(beta, old.x, o.x); o.y = fmaf(beta, old.y, o.y);
                    o.z = fmaf(beta, old.z, o.z); o.w = fmaf(beta, old.w, o.w);
                }
                __stcs(dst, o);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (n0 + cg * 16 + j < N) {
                    float o = alpha * acc[i][j];
                    if (beta != 0.f) o = fmaf(beta, crow[j], o);
                    crow[j] = o;
                }
            }
        }
    }
}

}  // namespace

namespace rb200 {

bool gemm_shortk_eligible(int64_t M, int64_t N, int64_t K, int64_t sAk, int64_t sBn)
{
    return K >= 1 && K <= SK_MAXK && sAk == 1 && sBn == 1 && M >= 4 * SK_BM && N <= 0x7fffffffLL &&
           rb_ceil_div(M, SK_BM) <= 0x7fffffffLL && rb_ceil_div(N, SK_BN) <= 65535;
}

int gemm_shortk(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                const float* B, int64_t ldB, float beta, float* C, int64_t ldC)
{
    Context& c = ctx();
    const int KS = (int)K | 1;
    const size_t smem = (size_t)(((SK_BM * KS + 3) & ~3) + (int)K * SK_BN) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(gemm_shortk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    dim3 grid((unsigned)rb_ceil_div(M, SK_BM), (unsigned)rb_ceil_div(N, SK_BN));
    const unsigned k_magic = (unsigned)((1ULL << 32) / (unsigned long long)K) + 1u;
    gemm_shortk_kernel<<<grid, SK_THREADS, smem, c.stream>>>(M, (int)N, (int)K, alpha, A, ldA, B, ldB, beta, C, ldC,
                                                            k_magic);
    RB_LAUNCH_CHECK("gemm_shortk_kernel");
    return RB200_OK;
}


template <typename T>
int gemm_simt_batched(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk, int64_t strideA,
                      const T* B, int64_t sBk, int64_t sBn, int64_t strideB, T beta, T* C, int64_t ldC, int64_t strideC,
                      int64_t batch);

template <typename T>
int gemm_simt(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk,
              const T* B, int64_t sBk, int64_t sBn, T beta, T* C, int64_t ldC)
{
    return gemm_simt_batched<T>(M, N, K, alpha, A, sAm, sAk, 0, B, sBk, sBn, 0, beta, C, ldC, 0, 1);
}

// `batch` matrices at element strides strideA / strideB / strideC (0 = the operand is shared) in one launch
template <typename T>
int gemm_simt_batched(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk, int64_t strideA,
                      const T* B, int64_t sBk, int64_t sBn, int64_t strideB, T beta, T* C, int64_t ldC, int64_t strideC,
                      int64_t batch)
{
    Context& c = ctx();
    const int64_t gx = rb_ceil_div(N, GS_BN), gy = rb_ceil_div(M, GS_BM);
    if (gy > 65535) {
        // very tall: run stripes of 65535*64 rows
        const int64_t stripe = 65535LL * GS_BM;
        for (int64_t r0 = 0; r0 < M; r0 += stripe) {
            const int64_t rows = (M - r0) < stripe ? (M - r0) : stripe;
            int rc = gemm_simt_batched<T>(rows, N, K, alpha, A + r0 * sAm, sAm, sAk, strideA, B, sBk, sBn, strideB, beta,
                                          C + r0 * ldC, ldC, strideC, batch);
            if (rc != RB200_OK) return rc;
        }
        return RB200_OK;
    }
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t nb = (batch - b0) < 65535 ? (batch - b0) : 65535;
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nb);
        gemm_simt_kernel<T><<<grid, GS_THREADS, 0, c.stream>>>(M, N, K, alpha, A + b0 * strideA, sAm, sAk, B + b0 * strideB,
                                                                sBk, sBn, beta, C + b0 * strideC, ldC, strideA, strideB, strideC);
        RB_LAUNCH_CHECK("gemm_simt_kernel");
    }
    return RB200_OK;
}

template int gemm_simt_batched<float>(int64_t, int64_t, int64_t, float, const float*, int64_t, int64_t, int64_t,
                                      const float*, int64_t, int64_t, int64_t, float, float*, int64_t, int64_t, int64_t);
template int gemm_simt_batched<double>(int64_t, int64_t, int64_t, double, const double*, int64_t, int64_t, int64_t,
                                       const double*, int64_t, int64_t, int64_t, double, double*, int64_t, int64_t, int64_t);
template int gemm_simt<float>(int64_t, int64_t, int64_t, float, const float*, int64_t, int64_t,
                              const float*, int64_t, int64_t, float, float*, int64_t);
template int gemm_simt<double>(int64_t, int64_t, int64_t, double, const double*, int64_t, int64_t,
                               const double*, int64_t, int64_t, double, double*, int64_t);

}  // namespace rb200
