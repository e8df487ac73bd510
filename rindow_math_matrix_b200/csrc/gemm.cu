// gemm.cu -- rb200_sgemm / rb200_dgemm: argument decoding and dispatch between the tensor-core
// path (gemm_tcgen05.cu) and the FFMA/DFMA path (gemm_simt.cu).
//
// Mirrors the argument handling of PhpBlas::gemm (src/Drivers/MatlibPHP/PhpBlas.php:849-886):
// ColMajor merely swaps m and n (:862-863); Trans/ConjTrans and NoTrans/ConjNoTrans are the
// same for real dtypes (:104-123); trans only swaps the two strides of an operand (:883-886).
#include "common.cuh"

namespace rb200 {
template <typename T>
int gemm_simt(int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t sAm, int64_t sAk,
              const T* B, int64_t sBk, int64_t sBn, T beta, T* C, int64_t ldC);
bool gemm_shortk_eligible(int64_t M, int64_t N, int64_t K, int64_t sAk, int64_t sBn);
int gemm_shortk(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
bool gemm_tcgen05_skinny_eligible(int mode, int64_t M, int64_t N, int64_t K);
int gemm_tcgen05_skinny(int mode, int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t ldA,
                        const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
bool gemm_tcgen05_eligible(int64_t M, int64_t N, int64_t K, const float* A, int64_t ldA,
                           const float* B, int64_t ldB, const float* C, int64_t ldC);
int gemm_tcgen05(int mode, bool transA, bool transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float* A, int64_t ldA, const float* B, int64_t ldB, float beta, float* C, int64_t ldC);
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
