// gemm_simt.cu -- FFMA (float32) / DFMA (float64) GEMM for every shape the tensor-core path
// cannot take: arbitrary m/n/k, leading dimensions, element offsets and transposes (the
// reference tests use ld = 3 and offset = 9, MatrixOperatorTest.php:308-318), and float64.
//
// Reference semantics (src/Drivers/MatlibPHP/PhpBlas.php:849-948):
//   C[i,j] = alpha * sum_k opA[i,k]*opB[k,j]  (+ beta*C[i,j] iff beta != 0; C not read otherwise)
// Each thread accumulates its 4x4 micro-tile with k ascending (fma chain), i.e. the same
// summation order as the reference's inner loop, in the operand precision.
//
// Small / odd problems: 64x64 CTA tile, BK = 16, 256 threads, shared-memory staging with the global loads mapped
// along whichever operand dimension is contiguous (coalesced for both NoTrans and Trans).  Problems whose tiles
// fill the machine take the large-tile double-buffered kernel below.
#include <stdlib.h>

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


// ---- large-tile kernel -------------------------------------------------------------------------------------------
// The kernel every big FFMA / DFMA problem takes (float64 always, float32 in FP32_SIMT mode or when TMA cannot address
// the operands): BM x BN x 16 CTA tile, 256 threads, two shared-memory stages.  Global -> registers -> shared memory,
// with the NEXT k-tile's loads in flight while the current one is multiplied; 128-bit global loads whenever the operand
// is 16-byte addressable along its contiguous dimension and scalar loads otherwise (any ld / offset / transpose).
// k ascends per output element: the reference's summation order (PhpBlas.php:927-947), in the operand precision.
//   float : 128 x 128 x 16, 8 x 8 per thread  (64 FFMA per 4 LDS.128), fragments double-buffered in registers
//   double: 128 x  64 x 16, 8 x 4 per thread  (32 DFMA per 6 LDS.128)
// What ncu changed (profiles/ncu_simt_big_r02.txt): (1) a thread's rows / columns are 16-byte groups interleaved across
// the lanes -- 4 consecutive doubles per thread made every B fragment load 2-way bank conflicted; (2) operands whose
// contiguous dimension is k (NoTrans A, Trans B) are read as whole row segments by 4 neighbouring threads (one cache
// line per row instead of one per thread: those shapes ran 10-25 % behind) and transposed into shared memory with a
// rotated store order that is conflict-free; (3) FFMA needs every issue slot (one warp instruction per clock per
// scheduler), so everything that is not an FFMA in the main loop costs throughput one for one: pointers, bounds
// predicates and the interior / edge decision are hoisted out of the k loop.
template <typename T> struct BigCfg;
template <> struct BigCfg<float>  { static constexpr int BM = 128, BN = 128, BK = 16; static constexpr bool FRAG_DB = true; };
template <> struct BigCfg<double> { static constexpr int BM = 128, BN = 64,  BK = 16; static constexpr bool FRAG_DB = false; };

template <typename T> struct Vec16;
template <> struct Vec16<float>  { using type = float4;  static constexpr int N = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int N = 2; };

// One operand tile [BK][BX] (x = the m or n dimension), EPT elements = NV 16-byte words per thread.
//   x-contiguous operand: element (x, k) at base[x + k * ld]  -> word v of a thread: k row tid / RUNS,
//             x = v * (RUNS * V) + (tid % RUNS) * V: consecutive threads touch consecutive words both in global
//             memory (coalesced) and in shared memory (conflict-free vector stores)
//   k-contiguous operand: element (x, k) at base[x * ld + k]  -> BK / V threads share a row (its k-run is one
//             contiguous segment), 256 / (BK / V) rows per pass, word v = pass v
template <typename T, int BX, int BK, int LD>
struct TileMover {
    static constexpr int V = Vec16<T>::N;
    using VT = typename Vec16<T>::type;
    static constexpr int EPT = BX * BK / 256, NV = EPT / V;
    static constexpr int RUNS = BX / EPT;                 // x-contiguous: threads per k row
    static constexpr int TPR = BK / V, RPP = 256 / TPR;   // k-contiguous: threads per row, rows per pass

    const T* p[NV];         // this thread's word v of k-tile 0
    int64_t step;           // pointer advance per k-tile
    int xleft[NV];          // valid elements of word v along x (x-contiguous) / 1 or 0 (k-contiguous: row inside)
    int kofs;               // k offset of the thread's words inside a tile
    int soff[NV];           // shared-memory offset of word v (x-contiguous) / of its first element (k-contiguous)
    bool xc, vec;
    int rot;

    __device__ __forceinline__ void init(const T* base, int64_t ld, int64_t x0, int64_t X, bool x_contig, bool vec_ok, int tid)
    {
        xc = x_contig; vec = vec_ok;
        if (xc) {
            const int kk = tid / RUNS, lr = tid % RUNS;
            kofs = kk; step = (int64_t)BK * ld; rot = 0;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int xo = v * (RUNS * V) + lr * V;
                p[v] = base + (int64_t)kk * ld + x0 + xo;
                const int64_t left = X - (x0 + xo);
                xleft[v] = left >= V ? V : (left > 0 ? (int)left : 0);
                soff[v] = kk * LD + xo;
            }
        } else {
            const int kq = tid % TPR, rr = tid / TPR;
            kofs = kq * V; step = BK;
            rot = (sizeof(T) == 4) ? (kq & 2) : 0;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                const int xo = v * RPP + rr;
                p[v] = base + (x0 + xo) * ld + kq * V;
                xleft[v] = (x0 + xo < X) ? 1 : 0;
                soff[v] = (kq * V) * LD + xo;
            }
        }
    }

    // k-tile starting at k0 (p already advanced): `interior` = the tile lies inside K and the operand is vector-addressable
    __device__ __forceinline__ void load(T (&regs)[EPT], int64_t k0, int64_t K, bool interior)
    {
        if (xc) {
            if (interior) {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    if (xleft[v] == V) {
                        const VT t = __ldg(reinterpret_cast<const VT*>(p[v]));
                        memcpy(&regs[v * V], &t, sizeof(VT));
                    } else {
#pragma unroll
                        for (int e = 0; e < V; e++) regs[v * V + e] = e < xleft[v] ? __ldg(p[v] + e) : (T)0;
                    }
                }
            } else {
                const bool kin = k0 + kofs < K;
#pragma unroll
                for (int v = 0; v < NV; v++)
#pragma unroll
                    for (int e = 0; e < V; e++) regs[v * V + e] = (kin && e < xleft[v]) ? __ldg(p[v] + e) : (T)0;
            }
        } else {
            if (interior) {
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    if (xleft[v]) {
                        const VT t = __ldg(reinterpret_cast<const VT*>(p[v]));
                        memcpy(&regs[v * V], &t, sizeof(VT));
                    } else {
#pragma unroll
                        for (int e = 0; e < V; e++) regs[v * V + e] = (T)0;
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < NV; v++)
#pragma unroll
                    for (int e = 0; e < V; e++) regs[v * V + e] = (xleft[v] && k0 + kofs + e < K) ? __ldg(p[v] + e) : (T)0;
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++) p[v] += step;
    }

    __device__ __forceinline__ void store(T* __restrict__ smem, const T (&regs)[EPT]) const
    {
        if (xc) {
#pragma unroll
            for (int v = 0; v < NV; v++) {
                VT t;
                memcpy(&t, &regs[v * V], sizeof(VT));
                *reinterpret_cast<VT*>(smem + soff[v]) = t;
            }
        } else {
            // transposing store: element e of the word goes to row (kofs + e) of the tile.  With 4-byte elements the
            // lanes kq and kq + 2 of a store instruction would hit the same banks (4 * LD = 16 mod 32): those lanes walk
            // their word rotated by two, which makes every store instruction conflict-free
#pragma unroll
            for (int v = 0; v < NV; v++) {
                T w[V];
#pragma unroll
                for (int e = 0; e < V; e++) w[e] = regs[v * V + e];
                if (sizeof(T) == 4 && rot) {
                    T t0 = w[0]; w[0] = w[2 % V]; w[2 % V] = t0;
                    T t1 = w[1]; w[1] = w[3 % V]; w[3 % V] = t1;
                }
#pragma unroll
                for (int e = 0; e < V; e++) smem[soff[v] + (e ^ rot) * LD] = w[e];
            }
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(256, 2)
gemm_simt_big_kernel(int64_t M, int64_t N, int64_t K, T alpha,
                     const T* __restrict__ A, int64_t sAm, int64_t sAk,
                     const T* __restrict__ B, int64_t sBk, int64_t sBn,
                     T beta, T* __restrict__ C, int64_t ldC,
                     int64_t strideA, int64_t strideB, int64_t strideC, int vecA, int vecB, int vecC, int group_m)
{
    using Cfg = BigCfg<T>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
    constexpr int V = Vec16<T>::N;
    using VT = typename Vec16<T>::type;
    // 8 warps as 4 (m) x 2 (n) warp tiles of WM x WN; inside a warp the lanes form a 4 (ly) x 8 (lx) grid and a
    // thread's rows / columns come in groups of V (one 16-byte shared-memory word): row group g at g * 4 * V + ly * V,
    // column group h at h * 8 * V + lx * V (conflict-free fragment loads, 128-byte row segments in the epilogue)
    constexpr int WM = BM / 4, WN = BN / 2;
    constexpr int GA = WM / (4 * V), GB = WN / (8 * V);                      // float: 2, 2   double: 4, 2
    constexpr int LDA = BM + 16 / sizeof(T), LDB = BN + 16 / sizeof(T);      // 16-byte pad: rows stay vector aligned
    using MoveA = TileMover<T, BM, BK, LDA>;
    using MoveB = TileMover<T, BN, BK, LDB>;
    extern __shared__ __align__(16) unsigned char big_smem[];
    T* As0 = reinterpret_cast<T*>(big_smem);                                 // [2][BK * LDA]
    T* Bs0 = As0 + 2 * BK * LDA;                                             // [2][BK * LDB]

    A += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
    // grouped rasterisation: `group_m` row blocks share each column block before the next column block starts
    const int num_m = (int)((M + BM - 1) / BM), num_n = (int)((N + BN - 1) / BN);
    int bm, bn;
    {
        const int tile = blockIdx.x;
        const int per_group = group_m * num_n;
        const int g = tile / per_group;
        const int first = g * group_m;
        const int gsz = min(num_m - first, group_m);
        const int in_g = tile - g * per_group;
        bm = first + in_g % gsz;
        bn = in_g / gsz;
    }
    const int64_t m0 = (int64_t)bm * BM, n0 = (int64_t)bn * BN;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int row_base = (warp >> 1) * WM + (lane >> 3) * V;       // + g * 4 * V + i
    const int col_base = (warp & 1) * WN + (lane & 7) * V;         // + h * 8 * V + j
    // contiguous dimension of each operand (ties -- a single row or column -- take the k-contiguous path)
    const bool A_x = (sAm == 1 && sAk != 1), B_x = (sBn == 1);
    MoveA ma;
    MoveB mb;
    ma.init(A, A_x ? sAk : sAm, m0, M, A_x, vecA != 0, tid);
    mb.init(B, B_x ? sBk : sBn, n0, N, B_x, vecB != 0, tid);

    T acc[GA][V][GB][V];
#pragma unroll
    for (int g = 0; g < GA; g++)
#pragma unroll
        for (int i = 0; i < V; i++)
#pragma unroll
            for (int h = 0; h < GB; h++)
#pragma unroll
                for (int j = 0; j < V; j++) acc[g][i][h][j] = (T)0;

    T ra[MoveA::EPT], rb[MoveB::EPT];
    const int64_t num_kt = (K + BK - 1) / BK;
    ma.load(ra, 0, K, vecA && BK <= K);
    mb.load(rb, 0, K, vecB && BK <= K);
    ma.store(As0, ra);
    mb.store(Bs0, rb);
    __syncthreads();

    for (int64_t kt = 0; kt < num_kt; kt++) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < num_kt) {
            const bool full = (kt + 2) * BK <= K;
            ma.load(ra, (kt + 1) * BK, K, vecA && full);
            mb.load(rb, (kt + 1) * BK, K, vecB && full);
        }
        const T* as = As0 + cur * (BK * LDA) + row_base;
        const T* bs = Bs0 + cur * (BK * LDB) + col_base;
        if (Cfg::FRAG_DB) {
            // register fragments double-buffered over kk: the shared-memory loads of step kk + 1 are in flight while
            // step kk multiplies
            VT fa[2][GA], fb[2][GB];
#pragma unroll
            for (int g = 0; g < GA; g++) fa[0][g] = *reinterpret_cast<const VT*>(as + g * 4 * V);
#pragma unroll
            for (int h = 0; h < GB; h++) fb[0][h] = *reinterpret_cast<const VT*>(bs + h * 8 * V);
#pragma unroll
            for (int kk = 0; kk < BK; kk++) {
                const int c = kk & 1, nx = c ^ 1;
                if (kk + 1 < BK) {
#pragma unroll
                    for (int g = 0; g < GA; g++) fa[nx][g] = *reinterpret_cast<const VT*>(as + (kk + 1) * LDA + g * 4 * V);
#pragma unroll
                    for (int h = 0; h < GB; h++) fb[nx][h] = *reinterpret_cast<const VT*>(bs + (kk + 1) * LDB + h * 8 * V);
                }
#pragma unroll
                for (int g = 0; g < GA; g++) {
                    T a[V];
                    memcpy(a, &fa[c][g], sizeof(VT));
#pragma unroll
                    for (int h = 0; h < GB; h++) {
                        T b[V];
                        memcpy(b, &fb[c][h], sizeof(VT));
#pragma unroll
                        for (int i = 0; i < V; i++)
#pragma unroll
                            for (int j = 0; j < V; j++) acc[g][i][h][j] = fma(a[i], b[j], acc[g][i][h][j]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < BK; kk++) {
                VT fa[GA], fb[GB];
#pragma unroll
                for (int g = 0; g < GA; g++) fa[g] = *reinterpret_cast<const VT*>(as + kk * LDA + g * 4 * V);
#pragma unroll
                for (int h = 0; h < GB; h++) fb[h] = *reinterpret_cast<const VT*>(bs + kk * LDB + h * 8 * V);
#pragma unroll
                for (int g = 0; g < GA; g++) {
                    T a[V];
                    memcpy(a, &fa[g], sizeof(VT));
#pragma unroll
                    for (int h = 0; h < GB; h++) {
                        T b[V];
                        memcpy(b, &fb[h], sizeof(VT));
#pragma unroll
                        for (int i = 0; i < V; i++)
#pragma unroll
                            for (int j = 0; j < V; j++) acc[g][i][h][j] = fma(a[i], b[j], acc[g][i][h][j]);
                    }
                }
            }
        }
        if (kt + 1 < num_kt) {
            ma.store(As0 + (cur ^ 1) * (BK * LDA), ra);
            mb.store(Bs0 + (cur ^ 1) * (BK * LDB), rb);
        }
        __syncthreads();
    }

#pragma unroll
    for (int g = 0; g < GA; g++)
#pragma unroll
        for (int i = 0; i < V; i++) {
            const int64_t gm = m0 + row_base + g * 4 * V + i;
            if (gm >= M) continue;
#pragma unroll
            for (int h = 0; h < GB; h++) {
                const int64_t gn = n0 + col_base + h * 8 * V;
                if (gn >= N) continue;
                T* c = C + gm * ldC + gn;
                if (vecC && gn + V <= N) {
                    T o[V];
#pragma unroll
                    for (int e = 0; e < V; e++) o[e] = alpha * acc[g][i][h][e];
                    if (beta != (T)0) {
                        const VT old = *reinterpret_cast<const VT*>(c);
                        T oe[V];
                        memcpy(oe, &old, sizeof(VT));
#pragma unroll
                        for (int e = 0; e < V; e++) o[e] = fma(beta, oe[e], o[e]);
                    }
                    VT t;
                    memcpy(&t, o, sizeof(VT));
                    *reinterpret_cast<VT*>(c) = t;
                } else {
#pragma unroll
                    for (int j = 0; j < V; j++) {
                        if (gn + j < N) {
                            const T o = alpha * acc[g][i][h][j];
                            c[j] = (beta == (T)0) ? o : fma(beta, c[j], o);
                        }
                    }
                }
            }
        }
}

template <typename T> constexpr size_t big_smem_bytes()
{
    using Cfg = BigCfg<T>;
    return (size_t)2 * Cfg::BK * ((Cfg::BM + 16 / sizeof(T)) + (Cfg::BN + 16 / sizeof(T))) * sizeof(T);
}

// ---- float64 on the FP64 tensor pipe (DMMA) ---------------------------------------------------------------------------
// mma.sync.aligned.m16n8k8.row.col.f64: one warp instruction = 1024 FMAs, against 32 for a DFMA.  The DFMA kernel above
// tops out at ~58 % of the FP64 pipe (ncu: every DFMA needs an issue slot pair and 6 LDS.128 feed 32 of them); here a
// warp's 32 x 32 tile takes 8 DMMAs and 16 LDS.64 per 8 k -- a sixth of the shared-memory traffic and an eighth of the
// math instructions.  128 x 64 x 16 CTA tile, 8 warps as 4 (m) x 2 (n), same global -> register -> shared-memory
// double buffering and the same TileMover (any ld / offset / transpose) as the FFMA / DFMA kernel.
// Fragment layout (PTX ISA; cute/atom/mma_traits_sm90.hpp SM90_16x8x8_F64F64F64F64_TN), g = lane / 4, t = lane % 4:
//   a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)      b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//   c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
// Shared-memory rows of 132 / 68 doubles (LD = 4 mod 16): the four k rows a half-warp touches land 8 banks apart and
// its four consecutive m (or n) fill those 8 banks -- every fragment load is conflict-free.
// Summation: the tensor pipe accumulates the 8 products of a step inside the instruction and k steps ascend; results
// agree with the k-ascending FMA chain of the reference loop to ~1e-15 relative (tests: < 1e-14).
__device__ __forceinline__ void dmma_16x8x8(double (&d)[4], const double (&a)[4], const double (&b)[2])
{
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

constexpr int DM_BM = 128, DM_BN = 64, DM_BK = 16, DM_LDA = DM_BM + 4, DM_LDB = DM_BN + 4;
constexpr size_t DM_SMEM = (size_t)2 * DM_BK * (DM_LDA + DM_LDB) * sizeof(double);

__global__ void __launch_bounds__(256, 2)
gemm_dmma_kernel(int64_t M, int64_t N, int64_t K, double alpha,
                 const double* __restrict__ A, int64_t sAm, int64_t sAk,
                 const double* __restrict__ B, int64_t sBk, int64_t sBn,
                 double beta, double* __restrict__ C, int64_t ldC,
                 int64_t strideA, int64_t strideB, int64_t strideC, int vecA, int vecB, int group_m)
{
    constexpr int BM = DM_BM, BN = DM_BN, BK = DM_BK, LDA = DM_LDA, LDB = DM_LDB;
    using MoveA = TileMover<double, BM, BK, LDA>;
    using MoveB = TileMover<double, BN, BK, LDB>;
    extern __shared__ __align__(16) unsigned char big_smem[];
    double* As0 = reinterpret_cast<double*>(big_smem);
    double* Bs0 = As0 + 2 * BK * LDA;

    A += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
    const int num_m = (int)((M + BM - 1) / BM), num_n = (int)((N + BN - 1) / BN);
    int bm, bn;
    {
        const int tile = blockIdx.x;
        const int per_group = group_m * num_n;
        const int g = tile / per_group;
        const int first = g * group_m;
        const int gsz = min(num_m - first, group_m);
        const int in_g = tile - g * per_group;
        bm = first + in_g % gsz;
        bn = in_g / gsz;
    }
    const int64_t m0 = (int64_t)bm * BM, n0 = (int64_t)bn * BN;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;          // this warp's 32 x 32 tile
    const bool A_x = (sAm == 1 && sAk != 1), B_x = (sBn == 1);
    MoveA ma;
    MoveB mb;
    ma.init(A, A_x ? sAk : sAm, m0, M, A_x, vecA != 0, tid);
    mb.init(B, B_x ? sBk : sBn, n0, N, B_x, vecB != 0, tid);

    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0.0;

    double ra[MoveA::EPT], rb[MoveB::EPT];
    const int64_t num_kt = (K + BK - 1) / BK;
    ma.load(ra, 0, K, vecA && BK <= K);
    mb.load(rb, 0, K, vecB && BK <= K);
    ma.store(As0, ra);
    mb.store(Bs0, rb);
    __syncthreads();

    for (int64_t kt = 0; kt < num_kt; kt++) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < num_kt) {
            const bool full = (kt + 2) * BK <= K;
            ma.load(ra, (kt + 1) * BK, K, vecA && full);
            mb.load(rb, (kt + 1) * BK, K, vecB && full);
        }
        const double* as = As0 + cur * (BK * LDA) + t * LDA + wm + g;
        const double* bs = Bs0 + cur * (BK * LDB) + t * LDB + wn + g;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ks++) {
            double a[2][4], b[4][2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                a[i][0] = as[(ks * 8) * LDA + i * 16];
                a[i][1] = as[(ks * 8) * LDA + i * 16 + 8];
                a[i][2] = as[(ks * 8 + 4) * LDA + i * 16];
                a[i][3] = as[(ks * 8 + 4) * LDA + i * 16 + 8];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                b[j][0] = bs[(ks * 8) * LDB + j * 8];
                b[j][1] = bs[(ks * 8 + 4) * LDB + j * 8];
            }
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_16x8x8(acc[i][j], a[i], b[j]);
        }
        if (kt + 1 < num_kt) {
            ma.store(As0 + (cur ^ 1) * (BK * LDA), ra);
            mb.store(Bs0 + (cur ^ 1) * (BK * LDB), rb);
        }
        __syncthreads();
    }

    const bool pair_ok = (ldC % 2 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int64_t gm = m0 + wm + i * 16 + h * 8 + g;
            if (gm >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t gn = n0 + wn + j * 8 + 2 * t;
                if (gn >= N) continue;
                double* c = C + gm * ldC + gn;
                double o0 = alpha * acc[i][j][2 * h], o1 = alpha * acc[i][j][2 * h + 1];
                if (pair_ok && gn + 2 <= N) {
                    if (beta != 0.0) {
                        const double2 old = *reinterpret_cast<const double2*>(c);
                        o0 = fma(beta, old.x, o0);
                        o1 = fma(beta, old.y, o1);
                    }
                    *reinterpret_cast<double2*>(c) = make_double2(o0, o1);
                } else {
                    c[0] = (beta == 0.0) ? o0 : fma(beta, c[0], o0);
                    if (gn + 1 < N) c[1] = (beta == 0.0) ? o1 : fma(beta, c[1], o1);
                }
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

static int launch_dmma(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t sAm, int64_t sAk,
                       int64_t strideA, const double* B, int64_t sBk, int64_t sBn, int64_t strideB, double beta, double* C,
                       int64_t ldC, int64_t strideC, int64_t batch, int vecA, int vecB)
{
    Context& c = ctx();
    static bool configured = false;
    if (!configured) {
        RB_CUDA(cudaFuncSetAttribute(gemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DM_SMEM));
        configured = true;
    }
    const int64_t tiles = rb_ceil_div(M, (int64_t)DM_BM) * rb_ceil_div(N, (int64_t)DM_BN);
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const int64_t nb = (batch - b0) < 65535 ? (batch - b0) : 65535;
        dim3 grid((unsigned)tiles, 1, (unsigned)nb);
        gemm_dmma_kernel<<<grid, 256, DM_SMEM, c.stream>>>(M, N, K, alpha, A + b0 * strideA, sAm, sAk, B + b0 * strideB, sBk,
                                                           sBn, beta, C + b0 * strideC, ldC, strideA, strideB, strideC,
                                                           vecA, vecB, 8);
        RB_LAUNCH_CHECK("gemm_dmma_kernel");
    }
    return RB200_OK;
}
static int launch_dmma(int64_t, int64_t, int64_t, float, const float*, int64_t, int64_t, int64_t, const float*, int64_t,
                       int64_t, int64_t, float, float*, int64_t, int64_t, int64_t, int, int)
{
    return RB200_E_UNSUPPORTED;
}

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
    // big problems: the large-tile double-buffered kernel, as soon as its tiles fill the machine
    {
        using Cfg = BigCfg<T>;
        const int64_t tiles = rb_ceil_div(M, (int64_t)Cfg::BM) * rb_ceil_div(N, (int64_t)Cfg::BN);
        static const int use_big = getenv("RB200_SIMT_BIG") ? atoi(getenv("RB200_SIMT_BIG")) : 1;
        if (use_big && K >= 2 * Cfg::BK && tiles * batch >= c.sm_count / 2 && tiles <= 0x7fffffffLL) {
            constexpr int V = 16 / (int)sizeof(T);
            auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
            // 128-bit loads along the contiguous dimension: base, leading dimension and batch stride 16-byte aligned
            const bool A_x = (sAm == 1 && sAk != 1);
            const int64_t ldA = A_x ? sAk : sAm, ldB = (sBn == 1) ? sBk : sBn;
            const int vecA = (al16(A) && ldA % V == 0 && strideA % V == 0 && (A_x ? true : sAk == 1)) ? 1 : 0;
            const int vecB = (al16(B) && ldB % V == 0 && strideB % V == 0 && ((sBn == 1) ? true : sBk == 1)) ? 1 : 0;
            const int vecC = (al16(C) && ldC % V == 0 && strideC % V == 0) ? 1 : 0;
            static const int use_dmma = getenv("RB200_DGEMM_DMMA") ? atoi(getenv("RB200_DGEMM_DMMA")) : 1;
            if (sizeof(T) == 8 && use_dmma) {
                int rc = launch_dmma(M, N, K, alpha, A, sAm, sAk, strideA, B, sBk, sBn, strideB, beta, C, ldC, strideC,
                                     batch, vecA, vecB);
                if (rc != RB200_OK) return rc;
                c.last_gemm_path = batch > 1 ? "dmma_f64_batched" : "dmma_f64";
                return RB200_OK;
            }
            for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
                const int64_t nb = (batch - b0) < 65535 ? (batch - b0) : 65535;
                dim3 grid((unsigned)tiles, 1, (unsigned)nb);
                static bool configured = false;
                if (!configured) {
                    RB_CUDA(cudaFuncSetAttribute(gemm_simt_big_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)big_smem_bytes<T>()));
                    configured = true;
                }
                gemm_simt_big_kernel<T><<<grid, 256, big_smem_bytes<T>(), c.stream>>>(M, N, K, alpha, A + b0 * strideA, sAm, sAk,
                                                                    B + b0 * strideB, sBk, sBn, beta, C + b0 * strideC, ldC,
                                                                    strideA, strideB, strideC, vecA, vecB, vecC, 8);
                RB_LAUNCH_CHECK("gemm_simt_big_kernel");
            }
            c.last_gemm_path = sizeof(T) == 4 ? (batch > 1 ? "simt_f32_big_batched" : "simt_f32_big")
                                              : (batch > 1 ? "simt_f64_big_batched" : "simt_f64_big");
            return RB200_OK;
        }
    }
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
