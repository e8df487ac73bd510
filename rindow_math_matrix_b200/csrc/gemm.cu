// gemm.cu -- rb200_sgemm / rb200_dgemm: argument decoding and dispatch between the tensor-core
// path (gemm_tcgen05.cu) and the FFMA/DFMA path (gemm_simt.cu).
//
// Mirrors the argument handling of PhpBlas::gemm (src/Drivers/MatlibPHP/PhpBlas.php:849-886):
// ColMajor merely swaps m and n (:862-863); Trans/ConjTrans and NoTrans/ConjNoTrans are the
// same for real dtypes (:104-123); trans only swaps the two strides of an operand (:883-886).
#include <stdlib.h>

#include "common.cuh"

namespace rb200 {
template <typename T>
int gemm_simt(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk,
              const T* B, int64_t sBk, int64_t sBn, T beta, T* C, int64_t ldC);
template <typename T>
int gemm_simt_batched(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk, int64_t strideA,
                      const T* B, int64_t sBk, int64_t sBn, int64_t strideB, T beta, T* C, int64_t ldC, int64_t strideC,
                      int64_t batch);
bool gemm_shortk_eligible(int64_t M, int64_t N, int64_t K, int64_t sAk, int64_t sBn);
int gemm_shortk(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
bool gemm_tcgen05_skinny_eligible(int mode, int64_t M, int64_t N, int64_t K);
int gemm_tcgen05_skinny(int mode, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                        const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
bool conv2d_tcgen05_eligible(int mode, int64_t M, int64_t N, int64_t K);
int conv2d_tcgen05(int mode, const float* images, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                   int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                   int64_t dilation_h, int64_t dilation_w, int64_t pad_h, int64_t pad_w, int64_t out_h, int64_t out_w,
                   const float* W, int64_t filters, const float* bias, float* out);
bool gemm_tcgen05_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t ldA,
                           const float* B, int64_t ldB, const float* C, int64_t ldC);
int gemm_tcgen05(int mode, bool transA, bool transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t ldA, const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
bool gemm_tcgen05_batched_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t ldA, int64_t strideA,
                                   const float* B, int64_t ldB, int64_t strideB, const float* C, int64_t ldC,
                                   int64_t strideC, int64_t batch);
int mc_push_rows(const float* src, float* mc, int64_t rows, int64_t cols, int64_t ld);
int gemm_tcgen05_batched(int mode, bool transA, bool transB, int64_t M, int64_t N, int64_t K, float alpha,
                         const float* A, int64_t ldA, int64_t strideA, const float* B, int64_t ldB, int64_t strideB,
                         float beta, float* C, int64_t ldC, int64_t strideC, int64_t batch);
}  // namespace rb200

namespace {

int decode(int order, int transA, int transB, int64_t& m, int64_t& n, int64_t k, bool& tA, bool& tB,
           const void* A, const void* B, const void* C, int64_t offA, int64_t ldA, int64_t offB, int64_t ldB,
           int64_t offC, int64_t ldC)
{
    RB_REQUIRE_INIT();
    if (order == RB200_COL_MAJOR) { int64_t t = m; m = n; n = t; }
    else if (order != RB200_ROW_MAJOR) RB_INVALID("Invalid Order type");
    switch (transA) {
        case RB200_NO_TRANS: case RB200_CONJ_NO_TRANS: tA = false; break;
        case RB200_TRANS: case RB200_CONJ_TRANS: tA = true; break;
        default: RB_INVALID("Unknown Tranpose Code: %d", transA);
    }
    switch (transB) {
        case RB200_NO_TRANS: case RB200_CONJ_NO_TRANS: tB = false; break;
        case RB200_TRANS: case RB200_CONJ_TRANS: tB = true; break;
        default: RB_INVALID("Unknown Tranpose Code: %d", transB);
    }
    if (m < 1) RB_INVALID("Argument m must be greater than 0.");
    if (n < 1) RB_INVALID("Argument n must be greater than 0.");
    if (k < 1) RB_INVALID("Argument k must be greater than 0.");
    if (!A || !B || !C) RB_INVALID("NULL buffer");
    if (offA < 0 || offB < 0 || offC < 0) RB_INVALID("Argument offset must be greater than equals 0.");
    if (ldA < 1 || ldB < 1 || ldC < 1) RB_INVALID("Argument ld must be greater than 0.");
    return RB200_OK;
}

}  // namespace

extern "C" {

int rb200_sgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                float alpha, const float* A, int64_t offA, int64_t ldA,
                const float* B, int64_t offB, int64_t ldB,
                float beta, float* C, int64_t offC, int64_t ldC, int mode)
{
    bool tA, tB;
    int rc = decode(order, transA, transB, m, n, k, tA, tB, A, B, C, offA, ldA, offB, ldB, offC, ldC);
    if (rc != RB200_OK) return rc;
    rb200::Context& c = rb200::ctx();
    if (mode == RB200_GEMM_AUTO) mode = c.default_gemm_mode;
    if (mode < RB200_GEMM_FP32_SIMT || mode > RB200_GEMM_TF32X3) RB_INVALID("unknown gemm mode %d", mode);
    const float* a = A + offA;
    const float* b = B + offB;
    float* cc = C + offC;
    const int64_t sAm = tA ? 1 : ldA, sAk = tA ? ldA : 1;
    const int64_t sBk = tB ? 1 : ldB, sBn = tB ? ldB : 1;
    if (mode != RB200_GEMM_FP32_SIMT) {
        // The TMA-fed path pays off from one full 128 x 128 tile upwards; the known-answer sized
        // problems (3x3 ... 64x64) and anything TMA cannot address run elsewhere.
        const bool big = (m >= 128 && n >= 128 && k >= 32);
        if (big && rb200::gemm_tcgen05_eligible(m, n, k, a, ldA, b, ldB, cc, ldC)) {
            return rb200::gemm_tcgen05(mode, tA, tB, m, n, k, alpha, a, ldA, b, ldB, beta, cc, ldC);
        }
        // tall-and-skinny NoTrans products (conv-as-GEMM: [B*oh*ow, kh*kw*C] x [kh*kw*C, filters]) of any
        // alignment: tensor cores fed by software-swizzled staging
        if (!tA && !tB && rb200::gemm_tcgen05_skinny_eligible(mode, m, n, k)) {
            return rb200::gemm_tcgen05_skinny(mode, m, n, k, alpha, a, ldA, b, ldB, beta, cc, ldC);
        }
    }
    // FFMA paths (exact fp32): short-K kernel for tall NoTrans shapes, generic tiled kernel otherwise
    if (rb200::gemm_shortk_eligible(m, n, k, sAk, sBn)) {
        c.last_gemm_path = "simt_f32_shortk";
        return rb200::gemm_shortk(m, n, k, alpha, a, ldA, b, ldB, beta, cc, ldC);
    }
    c.last_gemm_path = "simt_f32";
    return rb200::gemm_simt<float>(m, n, k, alpha, a, sAm, sAk, b, sBk, sBn, beta, cc, ldC);
}

// `batch` products at fixed element strides (0 shares an operand): what LinearAlgebra::matmul issues as a loop of
// gemm calls (LinearAlgebra.php:1091-1106).  Matrices of at least one tensor-core tile each go through rb200_sgemm
// one after the other (every launch already fills the machine); smaller / unaligned / float64 ones run as ONE
// launch of the FFMA / DFMA kernel with the batch on blockIdx.z.
int rb200_gemm_strided_batched(int dtype, int transA, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                               const void* A, int64_t offA, int64_t ldA, int64_t strideA,
                               const void* B, int64_t offB, int64_t ldB, int64_t strideB, double beta,
                               void* C, int64_t offC, int64_t ldC, int64_t strideC, int64_t batch, int mode)
{
    bool tA, tB;
    int rc = decode(RB200_ROW_MAJOR, transA, transB, m, n, k, tA, tB, A, B, C, offA, ldA, offB, ldB, offC, ldC);
    if (rc != RB200_OK) return rc;
    if (batch < 1) RB_INVALID("Argument batch must be greater than 0.");
    if (strideA < 0 || strideB < 0 || strideC < 0) RB_INVALID("Argument stride must be greater than equals 0.");
    if (strideC == 0 && batch > 1) RB_INVALID("strideC must not be 0: the products would overwrite each other");
    rb200::Context& c = rb200::ctx();
    const int64_t sAm = tA ? 1 : ldA, sAk = tA ? ldA : 1;
    const int64_t sBk = tB ? 1 : ldB, sBn = tB ? ldB : 1;
    if (dtype == RB200_FLOAT64) {
        c.last_gemm_path = "simt_f64_batched";
        return rb200::gemm_simt_batched<double>(m, n, k, alpha, reinterpret_cast<const double*>(A) + offA, sAm, sAk, strideA,
                                                reinterpret_cast<const double*>(B) + offB, sBk, sBn, strideB, beta,
                                                reinterpret_cast<double*>(C) + offC, ldC, strideC, batch);
    }
    if (dtype != RB200_FLOAT32) RB_UNSUPPORTED("gemm_strided_batched: unsupported dtype %d", dtype);
    if (mode == RB200_GEMM_AUTO) mode = c.default_gemm_mode;
    if (mode < RB200_GEMM_FP32_SIMT || mode > RB200_GEMM_TF32X3) RB_INVALID("unknown gemm mode %d", mode);
    const float* a = reinterpret_cast<const float*>(A) + offA;
    const float* b = reinterpret_cast<const float*>(B) + offB;
    float* cc = reinterpret_cast<float*>(C) + offC;
    // big matrices: one launch each already fills the machine with the CTA-pair kernel
    const bool big = mode != RB200_GEMM_FP32_SIMT && m >= 1024 && n >= 1024 && k >= 32 &&
                     (double)m * (double)n * (double)k >= 2048.0 * 2048.0 * 2048.0 &&
                     strideA % 4 == 0 && strideB % 4 == 0 && strideC % 4 == 0 &&
                     rb200::gemm_tcgen05_eligible(m, n, k, a, ldA, b, ldB, cc, ldC);
    if (big) {
        for (int64_t i = 0; i < batch; i++) {
            rc = rb200::gemm_tcgen05(mode, tA, tB, m, n, k, (float)alpha, a + i * strideA, ldA, b + i * strideB, ldB,
                                     (float)beta, cc + i * strideC, ldC);
            if (rc != RB200_OK) return rc;
        }
        return RB200_OK;
    }
    // tensor-core sized matrices: ONE persistent launch, the batch as a tile-index digit (chunks of 32768 matrices)
    if (mode != RB200_GEMM_FP32_SIMT &&
        rb200::gemm_tcgen05_batched_eligible(m, n, k, a, ldA, strideA, b, ldB, strideB, cc, ldC, strideC, batch < 32768 ? batch : 32768)) {
        for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
            const int64_t nb = (batch - b0) < 32768 ? (batch - b0) : 32768;
            rc = rb200::gemm_tcgen05_batched(mode, tA, tB, m, n, k, (float)alpha, a + b0 * strideA, ldA, strideA,
                                             b + b0 * strideB, ldB, strideB, (float)beta, cc + b0 * strideC, ldC, strideC, nb);
            if (rc != RB200_OK) return rc;
        }
        return RB200_OK;
    }
    c.last_gemm_path = "simt_f32_batched";
    return rb200::gemm_simt_batched<float>(m, n, k, (float)alpha, a, sAm, sAk, strideA, b, sBk, sBn, strideB, (float)beta,
                                           cc, ldC, strideC, batch);
}

// Row-sharded sgemm + all-gather of C through peer stores (include/rindow_b200.h).  The CTA-pair tensor-core kernel
// writes every finished value to the peers from its epilogue; any other kernel family multiplies first and the
// stripe is pushed with peer-to-peer copies on the same stream.
int rb200_sgemm_allgather(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                          float alpha, const float* A, int64_t offA, int64_t ldA,
                          const float* B, int64_t offB, int64_t ldB,
                          float beta, float* C, int64_t offC, int64_t ldC, int mode,
                          int n_peers, void* const* peer_C)
{
    RB_REQUIRE_INIT();
    if (n_peers < 0 || n_peers > RB200_MAX_PEERS || (n_peers > 0 && !peer_C)) RB_INVALID("bad peer list");
    if (order != RB200_ROW_MAJOR) RB_UNSUPPORTED("rb200_sgemm_allgather: row-major only (the stripes are rows)");
    rb200::Context& c = rb200::ctx();
    for (int j = 0; j < n_peers; j++) {
        if (!peer_C[j]) RB_INVALID("peer_C[%d] is NULL", j);
        c.gemm_peers[j] = reinterpret_cast<float*>(peer_C[j]) + offC;
    }
    c.n_gemm_peers = n_peers;
    c.gemm_peers_written = false;
    const int rc = rb200_sgemm(order, transA, transB, m, n, k, alpha, A, offA, ldA, B, offB, ldB, beta, C, offC, ldC, mode);
    c.n_gemm_peers = 0;
    if (rc != RB200_OK || n_peers == 0 || c.gemm_peers_written) return rc;
    for (int j = 0; j < n_peers; j++) {
        float* dst = reinterpret_cast<float*>(peer_C[j]) + offC;
        if (ldC == n) {
            RB_CUDA(cudaMemcpyAsync(dst, C + offC, (size_t)(m * n) * sizeof(float), cudaMemcpyDefault, c.stream));
        } else {
            RB_CUDA(cudaMemcpy2DAsync(dst, (size_t)ldC * sizeof(float), C + offC, (size_t)ldC * sizeof(float),
                                      (size_t)n * sizeof(float), (size_t)m, cudaMemcpyDefault, c.stream));
        }
    }
    return RB200_OK;
}

// The NVLS form: `mc_C` is the MULTICAST alias of the symmetric allocation C lives in (same element indexing).  The
// CTA-pair kernel's epilogue writes every finished 128-byte row segment once with multimem.st; the NVSwitch replicates
// it into all ranks' copies, so a stripe leaves this GPU once instead of once per peer.
int rb200_sgemm_allgather_mc(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                             float alpha, const float* A, int64_t offA, int64_t ldA,
                             const float* B, int64_t offB, int64_t ldB,
                             float beta, float* C, int64_t offC, int64_t ldC, int mode, void* mc_C)
{
    RB_REQUIRE_INIT();
    if (!mc_C) RB_INVALID("mc_C is NULL");
    if (order != RB200_ROW_MAJOR) RB_UNSUPPORTED("rb200_sgemm_allgather_mc: row-major only (the stripes are rows)");
    rb200::Context& c = rb200::ctx();
    // RB200_MC_EPILOGUE=0 (diagnostic): multiply first, then push the stripe with the contiguous multimem copy kernel
    static const int mc_epilogue = getenv("RB200_MC_EPILOGUE") ? atoi(getenv("RB200_MC_EPILOGUE")) : 1;
    c.n_gemm_peers = 0;
    c.gemm_mc = mc_epilogue ? reinterpret_cast<float*>(mc_C) + offC : nullptr;
    c.gemm_peers_written = false;
    const int rc = rb200_sgemm(order, transA, transB, m, n, k, alpha, A, offA, ldA, B, offB, ldB, beta, C, offC, ldC, mode);
    c.gemm_mc = nullptr;
    if (rc != RB200_OK || c.gemm_peers_written) return rc;
    return rb200::mc_push_rows(C + offC, reinterpret_cast<float*>(mc_C) + offC, m, n, ldC);
}

}  // extern "C"

// ---- rb200_sgemm_streamed: the host-operand form of sgemm -------------------------------------------
// For callers whose operands live in (pinned) host memory -- the PHP side's host-indexable buffers.  The
// copies are part of the call, so they are pipelined instead of serialised: B goes up first, then A (and C
// when beta != 0) in row stripes on a copy-in stream while the compute stream multiplies the stripes that
// have landed and a copy-out stream brings finished C stripes back over the other PCIe direction.
namespace {

struct StreamedState {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t ev_b = nullptr, ev_start = nullptr;
    cudaEvent_t ev_in[16] = {}, ev_done[16] = {};
    bool ready = false;
};
StreamedState g_streamed;

int streamed_init()
{
    StreamedState& st = g_streamed;
    if (st.ready) return RB200_OK;
    RB_CUDA(cudaStreamCreateWithFlags(&st.in, cudaStreamNonBlocking));
    RB_CUDA(cudaStreamCreateWithFlags(&st.out, cudaStreamNonBlocking));
    RB_CUDA(cudaEventCreateWithFlags(&st.ev_b, cudaEventDisableTiming));
    RB_CUDA(cudaEventCreateWithFlags(&st.ev_start, cudaEventDisableTiming));
    for (int i = 0; i < 16; i++) {
        RB_CUDA(cudaEventCreateWithFlags(&st.ev_in[i], cudaEventDisableTiming));
        RB_CUDA(cudaEventCreateWithFlags(&st.ev_done[i], cudaEventDisableTiming));
    }
    st.ready = true;
    return RB200_OK;
}

// copy the [rows x cols] view at element offset `off` with leading dimension ld (host <-> device mirrors share
// the element indexing): one contiguous copy when the rows are dense, a pitched copy otherwise
int copy_view(float* dst, const float* src, int64_t off, int64_t rows, int64_t cols, int64_t ld,
              cudaMemcpyKind kind, cudaStream_t stream)
{
    if (rows < 1 || cols < 1) return RB200_OK;
    if (ld == cols || rows == 1) {
        RB_CUDA(cudaMemcpyAsync(dst + off, src + off, (size_t)((rows - 1) * ld + cols) * sizeof(float), kind, stream));
    } else {
        RB_CUDA(cudaMemcpy2DAsync(dst + off, (size_t)ld * sizeof(float), src + off, (size_t)ld * sizeof(float),
                                  (size_t)cols * sizeof(float), (size_t)rows, kind, stream));
    }
    return RB200_OK;
}

}  // namespace

extern "C" {

int rb200_sgemm_streamed(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                         float alpha, float* dA, const float* hA, int64_t offA, int64_t ldA,
                         float* dB, const float* hB, int64_t offB, int64_t ldB,
                         float beta, float* dC, float* hC, int hC_is_current, int64_t offC, int64_t ldC, int mode)
{
    bool tA, tB;
    int rc = decode(order, transA, transB, m, n, k, tA, tB, dA, dB, dC, offA, ldA, offB, ldB, offC, ldC);
    if (rc != RB200_OK) return rc;
    rb200::Context& c = rb200::ctx();
    rc = streamed_init();
    if (rc != RB200_OK) return rc;
    StreamedState& st = g_streamed;
    // decode() swapped m,n for ColMajor: from here on the problem is row-major m x n
    // everything queued so far on the compute stream (earlier kernels that wrote dA/dB/dC, or read the regions
    // about to be overwritten) happens-before the first copy
    RB_CUDA(cudaEventRecord(st.ev_start, c.stream));
    RB_CUDA(cudaStreamWaitEvent(st.in, st.ev_start, 0));
    RB_CUDA(cudaStreamWaitEvent(st.out, st.ev_start, 0));

    const int64_t rowsB = tB ? n : k, colsB = tB ? k : n;
    if (hB) {
        rc = copy_view(dB, hB, offB, rowsB, colsB, ldB, cudaMemcpyHostToDevice, st.in);
        if (rc != RB200_OK) return rc;
    }
    RB_CUDA(cudaEventRecord(st.ev_b, st.in));
    RB_CUDA(cudaStreamWaitEvent(c.stream, st.ev_b, 0));

    // row stripes of C (and of op(A)): multiples of 256 rows (the CTA-pair tile), at most 16, at least ~1024 rows
    int64_t stripes = m / 1024;
    if (stripes > 16) stripes = 16;
    if (stripes < 1 || (!hA && !hC)) stripes = 1;
    int64_t stripe_rows = rb_ceil_div(rb_ceil_div(m, stripes), 256) * 256;
    stripes = rb_ceil_div(m, stripe_rows);
    const bool upload_c = (beta != 0.0f) && hC && hC_is_current;
    for (int64_t s = 0; s < stripes; s++) {
        const int64_t r0 = s * stripe_rows;
        const int64_t mr = (m - r0 < stripe_rows) ? m - r0 : stripe_rows;
        if (hA) {
            // NoTrans: rows r0.. of the m x k matrix; Trans: columns r0.. of the stored k x m matrix
            rc = tA ? copy_view(dA, hA, offA + r0, k, mr, ldA, cudaMemcpyHostToDevice, st.in)
                    : copy_view(dA, hA, offA + r0 * ldA, mr, k, ldA, cudaMemcpyHostToDevice, st.in);
            if (rc != RB200_OK) return rc;
        }
        if (upload_c) {
            rc = copy_view(dC, hC, offC + r0 * ldC, mr, n, ldC, cudaMemcpyHostToDevice, st.in);
            if (rc != RB200_OK) return rc;
        }
        RB_CUDA(cudaEventRecord(st.ev_in[s], st.in));
        RB_CUDA(cudaStreamWaitEvent(c.stream, st.ev_in[s], 0));
        rc = rb200_sgemm(RB200_ROW_MAJOR, transA, transB, mr, n, k, alpha, dA, tA ? offA + r0 : offA + r0 * ldA, ldA,
                         dB, offB, ldB, beta, dC, offC + r0 * ldC, ldC, mode);
        if (rc != RB200_OK) return rc;
        if (hC) {
            RB_CUDA(cudaEventRecord(st.ev_done[s], c.stream));
            RB_CUDA(cudaStreamWaitEvent(st.out, st.ev_done[s], 0));
            rc = copy_view(hC, dC, offC + r0 * ldC, mr, n, ldC, cudaMemcpyDeviceToHost, st.out);
            if (rc != RB200_OK) return rc;
        }
    }
    if (hC) RB_CUDA(cudaStreamSynchronize(st.out));      // the host may read hC as soon as the call returns
    // the caller may overwrite hA / hB (its pinned mirrors) as soon as the call returns: the uploads must have left
    // host memory by then.  Only the copies are waited for -- the multiplies of the last stripes stay asynchronous.
    else if (hA || hB || upload_c) RB_CUDA(cudaStreamSynchronize(st.in));
    return RB200_OK;
}

int rb200_conv2d_forward(int dtype, const void* images, int64_t images_offset,
                         int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                         int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                         int padding, int64_t dilation_h, int64_t dilation_w,
                         const void* kernel, int64_t kernel_offset, int64_t filters,
                         const void* bias, int64_t bias_offset, void* out, int64_t out_offset, int mode)
{
    RB_REQUIRE_INIT();
    rb200::Context& c = rb200::ctx();
    if (dtype != RB200_FLOAT32) RB_UNSUPPORTED("conv2d_forward: float32 only (use im2col2d + gemm)");
    if (!images || !kernel || !out) RB_INVALID("NULL buffer");
    if (batches < 1 || im_h < 1 || im_w < 1 || channels < 1 || filter_h < 1 || filter_w < 1 || stride_h < 1 ||
        stride_w < 1 || dilation_h < 1 || dilation_w < 1 || filters < 1) {
        RB_INVALID("conv2d_forward: shape parameters must be greater than 0");
    }
    if (images_offset < 0 || kernel_offset < 0 || out_offset < 0 || bias_offset < 0) RB_INVALID("negative offset");
    // same output-shape rule as PhpMath::im2col2d (PhpMath.php:2213-2241)
    int64_t out_h = (im_h - (filter_h - 1) * dilation_h - 1) / stride_h + 1;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_h <= 0 || out_w <= 0) RB_INVALID("Invalid shape or parameters.");
    int64_t pad_h = 0, pad_w = 0;
    if (padding) {
        pad_h = ((im_h - 1) * stride_h - im_h + (filter_h - 1) * dilation_h + 1) / 2;
        pad_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_h = im_h;
        out_w = im_w;
    }
    if (mode == RB200_GEMM_AUTO) mode = c.default_gemm_mode;
    const int64_t M = batches * out_h * out_w, K = filter_h * filter_w * channels;
    if (batches * im_h * im_w * channels >= (1LL << 31)) RB_UNSUPPORTED("conv2d_forward: images exceed 2^31 elements");
    // long reductions over channels % 32 == 0 take the TMA-fed implicit GEMM (conv_tcgen05_tma.cu), any filter count;
    // everything else needs filters <= 128 and K <= 128 (TMEM-operand kernel for K <= 32, skinny kernel above)
    const bool tma_shape = (mode == RB200_GEMM_TF32 || mode == RB200_GEMM_TF32X3) && K > 32 && channels % 32 == 0 &&
                           filters % 4 == 0 && filters >= 8;
    if (!tma_shape && !rb200::conv2d_tcgen05_eligible(mode, M, filters, K)) {
        RB_UNSUPPORTED("conv2d_forward: needs a TF32 mode and either channels %% 32 == 0 with filters %% 4 == 0, or "
                       "filters <= 128 and filter_h*filter_w*channels <= 128 (use im2col2d + gemm)");
    }
    return rb200::conv2d_tcgen05(mode, reinterpret_cast<const float*>(images) + images_offset, batches, im_h, im_w,
                                 channels, filter_h, filter_w, stride_h, stride_w, dilation_h, dilation_w, pad_h, pad_w,
                                 out_h, out_w, reinterpret_cast<const float*>(kernel) + kernel_offset, filters,
                                 bias ? reinterpret_cast<const float*>(bias) + bias_offset : nullptr,
                                 reinterpret_cast<float*>(out) + out_offset);
}

int rb200_dgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                double alpha, const double* A, int64_t offA, int64_t ldA,
                const double* B, int64_t offB, int64_t ldB,
                double beta, double* C, int64_t offC, int64_t ldC)
{
    bool tA, tB;
    int rc = decode(order, transA, transB, m, n, k, tA, tB, A, B, C, offA, ldA, offB, ldB, offC, ldC);
    if (rc != RB200_OK) return rc;
    rb200::ctx().last_gemm_path = "simt_f64";
    const int64_t sAm = tA ? 1 : ldA, sAk = tA ? ldA : 1;
    const int64_t sBk = tB ? 1 : ldB, sBn = tB ? ldB : 1;
    return rb200::gemm_simt<double>(m, n, k, alpha, A + offA, sAm, sAk, B + offB, sBk, sBn, beta, C + offC, ldC);
}

}  // extern "C"
