/*
 * rindow_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference's pure-PHP compute driver (PhpBlas / PhpMath on
 * PhpBuffer) for the hot path of SURVEY.md section 8.  It exists so that tests/, the
 * smoke() check and bench.py's cpu_baseline / --impl reference legs have something to
 * compare the CUDA path with.  Nothing under rindow_math_matrix_b200/ may include,
 * link or call this file: the product path has no CPU fallback.
 *
 * Storage model (reference: src/Drivers/MatlibPHP/PhpBuffer.php:19,86-93,124-133):
 * a PhpBuffer is an SplFixedArray of PHP scalars, so a "float32" buffer holds PHP
 * doubles that are never rounded to binary32, and every PhpBlas/PhpMath loop runs in
 * double precision, strictly sequentially in index order.  Hence every function here
 * works on `double` arrays; callers widen float32 inputs to double first (exactly the
 * values the PHP buffer would hold) and compare device float32 results against the
 * double result with the tolerance the test states.
 *
 * Parity pinning: the PHP interpreter is absent from this image and from the GPU
 * boxes, so the restatement cannot be diffed against a live run of the reference.  It
 * is pinned instead against every known-answer vector the reference's own PHPUnit
 * suite holds for this path (the JSON files in tests/golden, transcribed/extracted from
 * the PHP files of tests/RindowTest/Math/Matrix by tools/extract_golden.py) -- see
 * tests/test_oracle_golden.py.  Results on random data at BASELINE.json sizes are
 * pinned by this restatement only ("parity pinned by restatement").
 *
 * Compile:  see oracle/Makefile  (gcc -O2 -fno-fast-math -ffp-contract=off: no
 * reassociation, no FMA contraction -- PHP evaluates `$a * $b` and `+` separately).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define RO_OK            0
#define RO_E_ARG        -1   /* InvalidArgumentException in the reference */
#define RO_E_ZERODIV    -2   /* RuntimeException("Zero divide in softmax.") */

/* ---- helpers: PhpMath.php:83-130 ------------------------------------------------ */

/* math_sum, PhpMath.php:83-93: acc starts at 0.0, sequential += */
static double ro_math_sum(int64_t n, const double *X, int64_t offX, int64_t incX)
{
    int64_t idx = offX;
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++, idx += incX) {
        acc += X[idx];
    }
    return acc;
}

/* math_max, PhpMath.php:95-112: `if(!($value<=$max)) $max=$value;` -- a NaN replaces
 * the running max, and any later ordinary value replaces a NaN max again (quirk). */
static double ro_math_max(int64_t n, const double *X, int64_t offX, int64_t incX)
{
    int64_t idx = offX;
    double mx = X[idx];
    idx += incX;
    for (int64_t i = 1; i < n; i++, idx += incX) {
        double v = X[idx];
        if (!(v <= mx)) {
            mx = v;
        }
    }
    return mx;
}

/* NaN-sticky variant: the rule of the reference's accelerated path
 * (src/Drivers/MatlibCL/OpenCLMath.php:142-157 'lrmax' / :88-98 'qmax') which
 * LinearAlgebraTest.php:7617-7643 pins for NaN in the middle of a row.  Identical to
 * ro_math_max on NaN-free rows and on rows whose only NaN is the last element. */
static double ro_math_max_sticky(int64_t n, const double *X, int64_t offX, int64_t incX)
{
    int64_t idx = offX;
    double mx = X[idx];
    idx += incX;
    for (int64_t i = 1; i < n; i++, idx += incX) {
        double v = X[idx];
        if (isnan(mx)) {
            continue;
        }
        if (isnan(v) || v > mx) {
            mx = v;
        }
    }
    return mx;
}

/* math_imax, PhpMath.php:114-130: strict `>`; ties keep the lowest index; NaN never
 * wins; NaN at position 0 is never displaced. */
static int64_t ro_math_imax(int64_t n, const double *X, int64_t offX, int64_t incX)
{
    int64_t idx = offX;
    double mx = X[idx];
    int64_t imax = 0;
    idx += incX;
    for (int64_t i = 1; i < n; i++, idx += incX) {
        double v = X[idx];
        if (v > mx) {
            mx = v;
            imax = i;
        }
    }
    return imax;
}

/* ---- BLAS: PhpBlas.php:849-948 (real path :927-947) ------------------------------- */

/* order: 101 RowMajor / 102 ColMajor; trans: 111 NoTrans, 112 Trans, 113 ConjTrans,
 * 114 ConjNoTrans (CBLAS values; polite-math BLAS constants).  countA/B/C are the
 * buffer lengths (count($A)), used for the overflow checks of Utils.php:52-65. */
int ro_gemm(int order, int transA, int transB,
            int64_t m, int64_t n, int64_t k,
            double alpha,
            const double *A, int64_t countA, int64_t offA, int64_t ldA,
            const double *B, int64_t countB, int64_t offB, int64_t ldB,
            double beta,
            double *C, int64_t countC, int64_t offC, int64_t ldC)
{
    if (order == 102) {                       /* PhpBlas.php:862-863 */
        int64_t t = m; m = n; n = t;
    } else if (order != 101) {
        return RO_E_ARG;
    }
    int tA, tB;
    switch (transA) { case 111: case 114: tA = 0; break; case 112: case 113: tA = 1; break; default: return RO_E_ARG; }
    switch (transB) { case 111: case 114: tB = 0; break; case 112: case 113: tB = 1; break; default: return RO_E_ARG; }
    if (m < 1 || n < 1 || k < 1) return RO_E_ARG;            /* Utils.php:30-36 */
    int64_t rowsA = !tA ? m : k, colsA = !tA ? k : m;
    int64_t rowsB = !tB ? k : n, colsB = !tB ? n : k;
    if (offA < 0 || ldA < 1 || offA + (rowsA - 1) * ldA + (colsA - 1) >= countA) return RO_E_ARG;
    if (offB < 0 || ldB < 1 || offB + (rowsB - 1) * ldB + (colsB - 1) >= countB) return RO_E_ARG;
    if (offC < 0 || ldC < 1 || offC + (m - 1) * ldC + (n - 1) >= countC) return RO_E_ARG;

    int64_t ldA_m = !tA ? ldA : 1, ldA_k = !tA ? 1 : ldA;    /* PhpBlas.php:883-886 */
    int64_t ldB_k = !tB ? ldB : 1, ldB_n = !tB ? 1 : ldB;

    int64_t idA_m = offA, idC_m = offC;
    for (int64_t im = 0; im < m; im++, idA_m += ldA_m, idC_m += ldC) {
        int64_t idB_n = offB, idC = idC_m;
        for (int64_t in = 0; in < n; in++, idB_n += ldB_n, idC++) {
            int64_t idA = idA_m, idB = idB_n;
            double acc = 0.0;
            for (int64_t ik = 0; ik < k; ik++, idA += ldA_k, idB += ldB_k) {
                acc += A[idA] * B[idB];
            }
            if (beta == 0.0) {
                C[idC] = alpha * acc;                        /* C is not read */
            } else {
                C[idC] = alpha * acc + beta * C[idC];
            }
        }
    }
    return RO_OK;
}

/* Bounded-sample variant for bench.py's CPU baseline: identical arithmetic, but only
 * rows [row0, row0+rows) of C are produced.  Not a reference function. */
int ro_gemm_rows(int64_t row0, int64_t rows,
                 int64_t n, int64_t k, double alpha,
                 const double *A, int64_t ldA, const double *B, int64_t ldB,
                 double beta, double *C, int64_t ldC)
{
    for (int64_t im = row0; im < row0 + rows; im++) {
        for (int64_t in = 0; in < n; in++) {
            double acc = 0.0;
            const double *a = A + im * ldA;
            const double *b = B + in;
            for (int64_t ik = 0; ik < k; ik++) {
                acc += a[ik] * b[ik * ldB];
            }
            double *c = C + im * ldC + in;
            if (beta == 0.0) *c = alpha * acc; else *c = alpha * acc + beta * *c;
        }
    }
    return RO_OK;
}

/* Same results as ro_gemm_rows, bit for bit, with the n and k loops interchanged: every
 * C[im,in] still receives its products in k = 0..K-1 order (no reassociation inside an
 * element, no FMA), only independent elements are interleaved.  Used by the tests to check
 * sampled rows of large GEMMs in seconds; never timed as the baseline. */
int ro_gemm_rows_interleaved(int64_t row0, int64_t rows,
                             int64_t n, int64_t k, double alpha,
                             const double *A, int64_t ldA, const double *B, int64_t ldB,
                             double beta, double *C, int64_t ldC, double *acc /* n scratch */)
{
    for (int64_t im = row0; im < row0 + rows; im++) {
        const double *a = A + im * ldA;
        for (int64_t in = 0; in < n; in++) acc[in] = 0.0;
        for (int64_t ik = 0; ik < k; ik++) {
            const double av = a[ik];
            const double *b = B + ik * ldB;
            for (int64_t in = 0; in < n; in++) {
                acc[in] += av * b[in];
            }
        }
        double *c = C + im * ldC;
        for (int64_t in = 0; in < n; in++) {
            if (beta == 0.0) c[in] = alpha * acc[in]; else c[in] = alpha * acc[in] + beta * c[in];
        }
    }
    return RO_OK;
}

/* MatrixOperator::gemm3, src/MatrixOperator.php:510-529 (cross() with rank(B) > 2). */
int ro_gemm3(int64_t M, int64_t N, int64_t K, int64_t L, double alpha,
             const double *A, int64_t offA, const double *B, int64_t offB,
             double beta, double *C, int64_t offC)
{
    for (int64_t m = 0; m < M; m++) {
        for (int64_t l = 0; l < L; l++) {
            for (int64_t n = 0; n < N; n++) {
                double acc = 0.0;
                for (int64_t k = 0; k < K; k++) {
                    acc += A[offA + m * K + k] * B[offB + N * K * l + k * N + n];
                }
                int64_t ic = offC + m * L * N + N * l + n;
                C[ic] = alpha * acc + beta * C[ic];
            }
        }
    }
    return RO_OK;
}

/* ---- elementwise: PhpMath.php:530-606 --------------------------------------------- */

int ro_add(int trans, int64_t m, int64_t n, double alpha,
           const double *X, int64_t countX, int64_t offX, int64_t incX,
           double *A, int64_t countA, int64_t offA, int64_t ldA)
{
    int64_t rows = !trans ? m : n, cols = !trans ? n : m;
    if (offX + (cols - 1) * incX >= countX) return RO_E_ARG;      /* :590-591 */
    if (offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG; /* :592-593 */
    int64_t incAj = !trans ? ldA : 1, incAi = !trans ? 1 : ldA;
    int64_t idAj = offA;
    for (int64_t j = 0; j < rows; j++, idAj += incAj) {
        int64_t idA = idAj, idX = offX;
        for (int64_t i = 0; i < cols; i++, idA += incAi, idX += incX) {
            A[idA] = alpha * X[idX] + A[idA];                     /* one mul, one add */
        }
    }
    return RO_OK;
}

int ro_multiply(int trans, int64_t m, int64_t n,
                const double *X, int64_t countX, int64_t offX, int64_t incX,
                double *A, int64_t countA, int64_t offA, int64_t ldA)
{
    int64_t rows = !trans ? m : n, cols = !trans ? n : m;
    if (offX + (cols - 1) * incX >= countX) return RO_E_ARG;
    if (offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG;
    int64_t incAj = !trans ? ldA : 1, incAi = !trans ? 1 : ldA;
    int64_t idAj = offA;
    for (int64_t j = 0; j < rows; j++, idAj += incAj) {
        int64_t idA = idAj, idX = offX;
        for (int64_t i = 0; i < cols; i++, idA += incAi, idX += incX) {
            A[idA] = X[idX] * A[idA];
        }
    }
    return RO_OK;
}

/* ---- reductions over the middle axis of A[m,n,k]: PhpMath.php:1552-1643 ----------- */

int ro_reduce_sum(int64_t m, int64_t n, int64_t k,
                  const double *A, int64_t countA, int64_t offA,
                  double *B, int64_t countB, int64_t offB)
{
    if (offA + m * n * k > countA) return RO_E_ARG;
    if (offB + m * k > countB) return RO_E_ARG;
    int64_t idxA = offA, idxB = offB, ldA = n * k, ldB = k;
    for (int64_t i = 0; i < m; i++, idxA += ldA, idxB += ldB) {
        for (int64_t j = 0; j < k; j++) {
            B[idxB + j] = ro_math_sum(n, A, idxA + j, k);
        }
    }
    return RO_OK;
}

/* sticky=0: the PHP rule (math_max); sticky=1: the accelerated-path NaN-sticky rule. */
int ro_reduce_max(int sticky, int64_t m, int64_t n, int64_t k,
                  const double *A, int64_t countA, int64_t offA,
                  double *B, int64_t countB, int64_t offB)
{
    if (offA + m * n * k > countA) return RO_E_ARG;
    if (offB + m * k > countB) return RO_E_ARG;
    int64_t idxA = offA, idxB = offB, ldA = n * k, ldB = k;
    for (int64_t i = 0; i < m; i++, idxA += ldA, idxB += ldB) {
        for (int64_t j = 0; j < k; j++) {
            B[idxB + j] = sticky ? ro_math_max_sticky(n, A, idxA + j, k)
                                 : ro_math_max(n, A, idxA + j, k);
        }
    }
    return RO_OK;
}

int ro_reduce_argmax(int64_t m, int64_t n, int64_t k,
                     const double *A, int64_t countA, int64_t offA,
                     int64_t *B, int64_t countB, int64_t offB)
{
    if (offA + m * n * k > countA) return RO_E_ARG;
    if (offB + m * k > countB) return RO_E_ARG;
    int64_t idxA = offA, idxB = offB, ldA = n * k, ldB = k;
    for (int64_t i = 0; i < m; i++, idxA += ldA, idxB += ldB) {
        for (int64_t j = 0; j < k; j++) {
            B[idxB + j] = ro_math_imax(n, A, idxA + j, k);
        }
    }
    return RO_OK;
}

/* ---- softmax: PhpMath.php:1645-1670 ------------------------------------------------ */

int ro_softmax(int64_t m, int64_t n, double *A, int64_t countA, int64_t offA, int64_t ldA)
{
    if (offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG;
    int64_t idA = offA;
    for (int64_t i = 0; i < m; i++, idA += ldA) {
        double max_a = ro_math_max(n, A, idA, 1);
        double sum_exp = 0;
        for (int64_t j = 0; j < n; j++) {
            double t = exp(A[idA + j] - max_a);
            sum_exp += t;
            A[idA + j] = t;
        }
        if (sum_exp == 0.0) {
            return RO_E_ZERODIV;
        }
        for (int64_t j = 0; j < n; j++) {
            A[idA + j] = A[idA + j] / sum_exp;
        }
    }
    return RO_OK;
}

/* ---- im2col2d / col2im: PhpMath.php:2091-2148 (copyCell2d), :2158-2314 ------------- */

static void ro_copy_cell2d(int reverse, double *images, int64_t images_pos,
                           int64_t im_h, int64_t im_w, int64_t channels, int64_t channel_step,
                           int64_t filter_h_step, int64_t filter_w_step,
                           int64_t vim_y, int64_t vim_x, int64_t vfilter_h, int64_t vfilter_w,
                           int64_t dilation_h, int64_t dilation_w,
                           double *out, int64_t out_pos, int64_t out_filter_step,
                           int64_t out_channel_step)
{
    int64_t filter_h_pos = images_pos;
    int64_t out_filter_pos = out_pos;
    for (int64_t vfy = 0; vfy < vfilter_h; vfy += dilation_h) {
        int64_t filter_w_pos = filter_h_pos;
        for (int64_t vfx = 0; vfx < vfilter_w; vfx += dilation_w) {
            int64_t channel_pos = filter_w_pos;
            int64_t out_channel_pos = out_filter_pos;
            int64_t input_y = vim_y + vfy;
            int64_t input_x = vim_x + vfx;
            for (int64_t c = 0; c < channels; c++) {
                if (input_y < 0 || input_y >= im_h || input_x < 0 || input_x >= im_w) {
                    if (!reverse) {
                        out[out_channel_pos] = 0;
                    }
                } else {
                    if (!reverse) {
                        out[out_channel_pos] = images[channel_pos];
                    } else {
                        images[channel_pos] += out[out_channel_pos];
                    }
                }
                out_channel_pos += out_channel_step;
                channel_pos += channel_step;
            }
            out_filter_pos += out_filter_step;
            filter_w_pos += filter_w_step;
        }
        filter_h_pos += filter_h_step;
    }
}

int ro_im2col2d(int reverse,
                double *images, int64_t count_images, int64_t images_offset, int64_t images_size,
                int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                int padding, int channels_first, int64_t dilation_h, int64_t dilation_w,
                int cols_channels_first,
                double *cols, int64_t count_cols, int64_t cols_offset, int64_t cols_size)
{
    int64_t images_buf_size = batches * im_h * im_w * channels;
    if (images_size != images_buf_size || count_images - images_offset < images_buf_size) {
        return RO_E_ARG;                                  /* 'images buffer size is invalid' */
    }
    /* PHP intdiv truncates toward zero, like C integer division. */
    int64_t out_h = (im_h - (filter_h - 1) * dilation_h - 1) / stride_h + 1;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_h <= 0 || out_w <= 0) {
        return RO_E_ARG;                                  /* 'Invalid shape or parameters.' */
    }
    int64_t out_buf_size, padding_h, padding_w;
    if (padding) {
        out_buf_size = batches * im_h * filter_h * im_w * filter_w * channels;
        padding_h = ((im_h - 1) * stride_h - im_h + (filter_h - 1) * dilation_h + 1) / 2;
        padding_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_h = im_h;
        out_w = im_w;
    } else {
        out_buf_size = batches * out_h * filter_h * out_w * filter_w * channels;
        padding_h = 0;
        padding_w = 0;
    }
    /* PhpMath.php:2243-2246: note the reference's `>` -- a cols buffer LARGER than the
     * view is rejected; a smaller one would fault in SplFixedArray, so reject it too. */
    if (cols_size != out_buf_size || count_cols - cols_offset > out_buf_size
        || count_cols - cols_offset < out_buf_size) {
        return RO_E_ARG;                                  /* 'output buffer size is invalid' */
    }
    int64_t im_w_step, im_h_step, channel_step, batch_step;
    if (channels_first) {
        im_w_step = 1; im_h_step = im_w; channel_step = im_w * im_h; batch_step = im_w * im_h * channels;
    } else {
        channel_step = 1; im_w_step = channels; im_h_step = channels * im_w; batch_step = channels * im_w * im_h;
    }
    int64_t stride_w_step = im_w_step * stride_w, stride_h_step = im_h_step * stride_h;
    int64_t filter_w_step = im_w_step * dilation_w, filter_h_step = im_h_step * dilation_h;
    int64_t out_filter_step, out_channel_step;
    if (cols_channels_first) {
        out_filter_step = 1; out_channel_step = filter_h * filter_w;
    } else {
        out_filter_step = channels; out_channel_step = 1;
    }
    int64_t out_cell_step = filter_h * filter_w * channels;
    int64_t batch_pos = images_offset - im_h_step * padding_h - im_w_step * padding_w;
    int64_t out_pos = cols_offset;
    int64_t vim_h = out_h * stride_h, vim_w = out_w * stride_w;
    int64_t vfilter_h = filter_h * dilation_h, vfilter_w = filter_w * dilation_w;
    for (int64_t batch = 0; batch < batches; batch++) {
        int64_t stride_h_pos = batch_pos;
        for (int64_t vim_y = 0; vim_y < vim_h; vim_y += stride_h) {
            int64_t stride_w_pos = stride_h_pos;
            for (int64_t vim_x = 0; vim_x < vim_w; vim_x += stride_w) {
                ro_copy_cell2d(reverse, images, stride_w_pos, im_h, im_w, channels, channel_step,
                               filter_h_step, filter_w_step, vim_y - padding_h, vim_x - padding_w,
                               vfilter_h, vfilter_w, dilation_h, dilation_w,
                               cols, out_pos, out_filter_step, out_channel_step);
                stride_w_pos += stride_w_step;
                out_pos += out_cell_step;
            }
            stride_h_pos += stride_h_step;
        }
        batch_pos += batch_step;
    }
    return RO_OK;
}

/* PhpMath::gatherb, PhpMath.php:1209-1260.
 * f32 = 0 is the reference's semantics: a PhpBuffer is an SplFixedArray of PHP doubles, so partial sums are never
 * rounded.  f32 != 0 is a float32 STORAGE MODEL (every stored partial sum rounded to float32,
 * what any driver with real float buffers does, the reference's own C/OpenCL drivers included); it is NOT PhpBuffer
 * behaviour.  The GPU tests assert both: bit-exact against the storage model and within terms*eps32*sum|terms| of
 * the unrounded result (tests/test_gpu_gatherb.py, test_gpu_scan.py). */
int ro_gatherb(int reverse, int add_mode, int f32, int64_t batches, int64_t m, int64_t n, int64_t k, int64_t len, int64_t numClass,
               double *A, int64_t countA, int64_t offA, const double *X, int64_t countX, int64_t offX,
               double *B, int64_t countB, int64_t offB)
{
    if (offX + batches * n * k > countX || offA + batches * m * numClass * k * len > countA ||
        offB + batches * m * n * k * len > countB) return RO_E_ARG;
    for (int64_t batch = 0; batch < batches; batch++) {
        for (int64_t i = 0; i < m; i++) {
            for (int64_t j = 0; j < n; j++) {
                for (int64_t h = 0; h < k; h++) {
                    const int64_t index = (int64_t)X[offX + (batch * n + j) * k + h];
                    if (index >= numClass || index < 0) continue;
                    double *a = A + offA + (((batch * m + i) * numClass + index) * k + h) * len;
                    double *b = B + offB + (((batch * m + i) * n + j) * k + h) * len;
                    double *from = reverse ? b : a, *to = reverse ? a : b;
                    for (int64_t t = 0; t < len; t++) {
                        if (!add_mode) to[t] = from[t];
                        else { const double v = to[t] + from[t]; to[t] = f32 ? (double)(float)v : v; }
                    }
                }
            }
        }
    }
    return RO_OK;
}

/* PhpMath::gathernd, PhpMath.php:1271-1325.  X is read from element 0: the reference does not apply offsetX (:1292). */
int ro_gathernd(int reverse, int add_mode, int f32, int64_t m, int64_t n, int64_t k, int64_t depth, const int64_t *paramShape,
                double *A, int64_t countA, int64_t offA, const double *X, int64_t countX,
                double *B, int64_t countB, int64_t offB)
{
    int64_t paramSize = 1;
    for (int64_t h = 0; h < depth; h++) paramSize *= paramShape[h];
    if (m * n * depth > countX || offA + m * paramSize * k > countA || offB + m * n * k > countB) return RO_E_ARG;
    for (int64_t i = 0; i < m; i++) {
        for (int64_t j = 0; j < n; j++) {
            int64_t offset = 0;
            int error = 0;
            for (int64_t h = 0; h < depth; h++) {
                offset *= paramShape[h];
                const int64_t index = (int64_t)X[i * n * depth + j * depth + h];
                if (index >= paramShape[h] || index < 0) { error = 1; break; }
                offset += index;
            }
            if (error) continue;
            double *a = A + offA + i * paramSize * k + offset * k;
            double *b = B + offB + i * n * k + j * k;
            double *from = reverse ? b : a, *to = reverse ? a : b;
            for (int64_t t = 0; t < k; t++) {
                if (!add_mode) to[t] = from[t];
                else { const double v = to[t] + from[t]; to[t] = f32 ? (double)(float)v : v; }
            }
        }
    }
    return RO_OK;
}

/* PhpMath::imagecopy, PhpMath.php:2877-2957 */
int ro_imagecopy(int64_t height, int64_t width, int64_t channels, const double *A, int64_t countA, int64_t offA,
                 double *B, int64_t countB, int64_t offB, int channelsFirst, int64_t heightShift, int64_t widthShift,
                 int verticalFlip, int horizontalFlip, int rgbFlip)
{
    if (width * height * channels + offA > countA || width * height * channels + offB > countB) return RO_E_ARG;
    int64_t ldC, ldY, ldX;
    if (channelsFirst) { ldC = width * height; ldY = width; ldX = 1; }
    else { ldY = width * channels; ldX = channels; ldC = 1; }
    int64_t directionY = 1, directionX = 1, biasY = 0, biasX = 0;
    if (verticalFlip) { directionY = -directionY; biasY = height - 1; }
    if (horizontalFlip) { directionX = -directionX; biasX = width - 1; }
    biasY -= heightShift * directionY;
    biasX -= widthShift * directionX;
    for (int64_t y = 0; y < height; y++) {
        for (int64_t x = 0; x < width; x++) {
            for (int64_t c = 0; c < channels; c++) {
                int64_t sy = y * directionY + biasY, sx = x * directionX + biasX;
                if (sy < 0) sy = 0; else if (sy >= height) sy = height - 1;
                if (sx < 0) sx = 0; else if (sx >= width) sx = width - 1;
                const int64_t srcC = (rgbFlip && c < 3) ? (2 - c) : c;
                if (srcC >= channels) return RO_E_ARG;             /* the PHP buffer would be read outside the image */
                B[y * ldY + x * ldX + c * ldC + offB] = A[sy * ldY + sx * ldX + srcC * ldC + offA];
            }
        }
    }
    return RO_OK;
}

/* PhpMath::einsum, PhpMath.php:3321-3357 (odometer over all `depth` axes, the first ndimC index C; einsum4p1
 * :3360-3440 is the depth 5 / ndimC 4 case).  The sum is a PHP float (double). */
int ro_einsum(int64_t depth, int64_t ndimC, const int64_t *sizeOfIndices, const double *A, int64_t countA, int64_t offA,
              const int64_t *ldA, const double *B, int64_t countB, int64_t offB, const int64_t *ldB,
              double *C, int64_t countC, int64_t offC)
{
    if (depth < 1 || depth > 16) return RO_E_ARG;
    int64_t sizeC = 1, maxA = 0, maxB = 0;
    for (int64_t i = 0; i < depth; i++) {
        if (i < ndimC) sizeC *= sizeOfIndices[i];
        maxA += (sizeOfIndices[i] - 1) * ldA[i];
        maxB += (sizeOfIndices[i] - 1) * ldB[i];
    }
    if (offA + maxA >= countA || offB + maxB >= countB || offC + sizeC > countC) return RO_E_ARG;
    int64_t indices[16] = {0};
    for (int64_t indexC = 0; indexC < sizeC; indexC++) {
        double sumC = 0;
        for (;;) {
            int64_t indexA = offA, indexB = offB;
            for (int64_t axis = 0; axis < depth; axis++) { indexA += indices[axis] * ldA[axis]; indexB += indices[axis] * ldB[axis]; }
            sumC += A[indexA] * B[indexB];
            int64_t i = depth - 1;                                   /* einsum_next_indices, :3297-3318 */
            while (i >= 0) {
                if (indices[i] < sizeOfIndices[i] - 1) break;
                indices[i] = 0;
                i--;
            }
            if (i >= 0) indices[i]++;
            if (i < ndimC) break;
        }
        C[offC + indexC] = sumC;
    }
    return RO_OK;
}

/* PhpMath::topk, PhpMath.php:933-1058: min-heap of the first k, the rest offered in order, optional heap sort */
static void ro_topk_heapify(int64_t size, double *heap, double *idx, int64_t parent)
{
    int64_t left = 2 * parent + 1, right = 2 * parent + 2;
    while (left < size) {
        int64_t smallest = (right < size && heap[right] < heap[left]) ? right : left;
        if (heap[parent] <= heap[smallest]) break;
        double t = heap[parent]; heap[parent] = heap[smallest]; heap[smallest] = t;
        t = idx[parent]; idx[parent] = idx[smallest]; idx[smallest] = t;
        parent = smallest;
        left = 2 * parent + 1;
        right = 2 * parent + 2;
    }
}

int ro_topk(int64_t m, int64_t n, const double *input, int64_t countIn, int64_t offIn, int64_t k, int sorted,
            double *values, int64_t countV, int64_t offV, double *indices, int64_t countI, int64_t offI)
{
    if (m < 1 || n < 1 || k < 1) return RO_E_ARG;
    if (offIn + m * n > countIn || offV + m * k > countV || offI + m * k > countI) return RO_E_ARG;
    if (k > n) return RO_OK;
    for (int64_t r = 0; r < m; r++) {
        const double *arr = input + offIn + r * n;
        double *top = values + offV + r * k, *idx = indices + offI + r * k;
        for (int64_t i = 0; i < k; i++) { top[i] = arr[i]; idx[i] = (double)i; }
        for (int64_t i = k / 2 - 1; i >= 0; --i) ro_topk_heapify(k, top, idx, i);
        for (int64_t i = k; i < n; i++) {
            if (arr[i] > top[0]) { top[0] = arr[i]; idx[0] = (double)i; ro_topk_heapify(k, top, idx, 0); }
        }
        if (sorted) {
            for (int64_t i = k - 1; i > 0; --i) {
                double t = top[0]; top[0] = top[i]; top[i] = t;
                t = idx[0]; idx[0] = idx[i]; idx[i] = t;
                ro_topk_heapify(i, top, idx, 0);
            }
        }
    }
    return RO_OK;
}

/* PhpMath::cumsumb, PhpMath.php:1861-1904.  f32 = 0: the reference's semantics (PhpBuffer holds doubles, nothing is
 * rounded between steps).  f32 != 0: float32 storage model -- every stored partial sum is read back by the next step,
 * so a real float32 buffer rounds at every step (see ro_gatherb). */
int ro_cumsumb(int64_t m, int64_t n, int64_t k, const double *A, int64_t countA, int64_t offA, int exclusive, int reverse,
               int f32, double *B, int64_t countB, int64_t offB)
{
    if (offA + m * n * k > countA || offB + m * n * k > countB) return RO_E_ARG;
    for (int64_t i = 0; i < m; i++) {
        int64_t ida = offA + i * n * k + (reverse ? (n - 1) * k : 0);
        int64_t idb = offB + i * n * k + (reverse ? (n - 1) * k : 0);
        const int64_t ld = reverse ? -k : k;
        for (int64_t h = 0; h < k; h++) B[idb + h] = exclusive ? 0.0 : A[ida + h];
        for (int64_t j = 0; j < n - 1; j++) {
            idb += ld;
            if (!exclusive) ida += ld;
            for (int64_t h = 0; h < k; h++) {
                const double v = B[idb - ld + h] + A[ida + h];
                B[idb + h] = f32 ? (double)(float)v : v;
            }
            if (exclusive) ida += ld;
        }
    }
    return RO_OK;
}

/* PhpMath::cumsum, PhpMath.php:1822-1859: the running value is a PHP float (double); reverse only turns Y around. */
int ro_cumsum(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, int exclusive, int reverse,
              double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (n < 1 || offX + (n - 1) * incX >= countX || offY + (n - 1) * incY >= countY) return RO_E_ARG;
    int64_t idx = offX, idy = offY;
    if (reverse) { idy = offY + incY * (n - 1); incY = -incY; }
    double value = 0.0;
    for (int64_t i = 0; i < n; i++) {
        if (exclusive) { Y[idy] = value; value += X[idx]; }
        else { value += X[idx]; Y[idy] = value; }
        idx += incX; idy += incY;
    }
    return RO_OK;
}

/* PhpMath::searchsorted, PhpMath.php:1780-1820 */
int ro_searchsorted(int64_t m, int64_t n, const double *A, int64_t countA, int64_t offA, int64_t ldA,
                    const double *X, int64_t countX, int64_t offX, int64_t incX, int right,
                    double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (m < 1 || offA + (m - 1) * ldA + n > countA || offX + (m - 1) * incX >= countX || offY + (m - 1) * incY >= countY) {
        return RO_E_ARG;
    }
    for (int64_t i = 0; i < m; i++) {
        const double v = X[offX + i * incX];
        const double *a = A + offA + i * ldA;
        int64_t j = 0;
        for (; j < n; j++) {
            if (right ? !(v >= a[j]) : !(v > a[j])) break;
        }
        Y[offY + i * incY] = (double)j;
    }
    return RO_OK;
}

/* PhpMath::bandpart */
int ro_bandpart(int64_t m, int64_t n, int64_t k, double *A, int64_t countA, int64_t offA, int64_t lower, int64_t upper)
{
    if (offA + m * n * k > countA) return RO_E_ARG;
    for (int64_t b = 0; b < m; b++) {
        for (int64_t i = 0; i < n; i++) {
            for (int64_t j = 0; j < k; j++) {
                if ((lower >= 0 && (i - j) > lower) || (upper >= 0 && (j - i) > upper)) A[offA + b * n * k + i * k + j] = 0;
            }
        }
    }
    return RO_OK;
}

/* PhpMath::masking, PhpMath.php:3229-3270.  X holds the bool mask as 0 / 1 doubles; kind: 0 float, 1 integer (fill
 * truncated toward zero as the C buffer stores it), 2 bool (mode 1: A || fill). */
int ro_masking(int64_t m, int64_t n, int64_t k, int64_t len, double fill, int mode, int kind,
               const double *X, int64_t countX, int64_t offX, double *A, int64_t countA, int64_t offA)
{
    if (offX + m * k > countX) return RO_E_ARG;
    if (offA + m * n * k * len > countA) return RO_E_ARG;
    if (kind == 1) fill = (double)(int64_t)fill;
    if (kind == 2) fill = fill != 0.0;
    for (int64_t i = 0; i < m; i++) {
        for (int64_t j = 0; j < n; j++) {
            for (int64_t h = 0; h < k; h++) {
                if (X[offX + i * k + h] != 0.0) continue;
                for (int64_t l = 0; l < len; l++) {
                    int64_t a = offA + ((i * n + j) * k + h) * len + l;
                    if (mode == 0) A[a] = fill;
                    else if (kind == 2) A[a] = (A[a] != 0.0 || fill != 0.0);
                    else A[a] = A[a] + fill;
                }
            }
        }
    }
    return RO_OK;
}

/* PhpMath::slice, PhpMath.php:2691-2776 (math_copy / math_add :3197-3226) */
int ro_slice(int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int64_t size,
             double *A, int64_t countA, int64_t offA, int64_t incA,
             double *Y, int64_t countY, int64_t offY, int64_t incY,
             int64_t start0, int64_t size0, int64_t start1, int64_t size1, int64_t start2, int64_t size2)
{
    if (m * n * k * size * incA + offA > countA) return RO_E_ARG;          /* 'unmatch BufferA size and m,n,k' */
    if (start0 < 0 || start0 >= m || size0 < 0 || size0 + start0 > m) return RO_E_ARG;
    if (start1 < 0 || start1 >= n || size1 < 0 || size1 + start1 > n) return RO_E_ARG;
    if (start2 < 0 || start2 >= k || size2 < 0 || size2 + start2 > k) return RO_E_ARG;
    if (size0 * size1 * size2 * size * incY > countY - offY) return RO_E_ARG;   /* 'BufferY size is too small' */
    for (int64_t i0 = 0; i0 < size0; i0++) {
        for (int64_t i1 = 0; i1 < size1; i1++) {
            for (int64_t i2 = 0; i2 < size2; i2++) {
                int64_t pa = (i0 + start0) * n * k * size + (i1 + start1) * k * size + (i2 + start2) * size + offA;
                int64_t py = i0 * size1 * size2 * size + i1 * size2 * size + i2 * size + offY;
                for (int64_t e = 0; e < size; e++, pa += incA, py += incY) {
                    if (!reverse) { if (add_mode) Y[py] = Y[py] + A[pa]; else Y[py] = A[pa]; }
                    else          { if (add_mode) A[pa] = A[pa] + Y[py]; else A[pa] = Y[py]; }
                }
            }
        }
    }
    return RO_OK;
}

/* PhpMath::repeat, PhpMath.php:1400-1422 */
int ro_repeat(int64_t m, int64_t k, int64_t repeats, const double *A, int64_t countA, int64_t offA,
              double *B, int64_t countB, int64_t offB)
{
    if (offA + m * k > countA) return RO_E_ARG;
    if (offB + m * repeats * k > countB) return RO_E_ARG;
    int64_t idA = offA, idB = offB;
    for (int64_t i = 0; i < m; i++, idA += k) {
        for (int64_t j = 0; j < repeats; j++, idB += k) {
            for (int64_t e = 0; e < k; e++) B[idB + e] = A[idA + e];
        }
    }
    return RO_OK;
}

/* PhpMath::im2col1d, PhpMath.php:1967-2088 (copyCell1d :1912-1958): the loops of im2col2d without the y axis. */
int ro_im2col1d(int reverse,
                double *images, int64_t count_images, int64_t images_offset, int64_t images_size,
                int64_t batches, int64_t im_w, int64_t channels, int64_t filter_w, int64_t stride_w,
                int padding, int channels_first, int64_t dilation_w, int cols_channels_first,
                double *cols, int64_t count_cols, int64_t cols_offset, int64_t cols_size)
{
    int64_t images_buf_size = batches * im_w * channels;
    if (images_size != images_buf_size || count_images - images_offset < images_buf_size) return RO_E_ARG;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_w <= 0) return RO_E_ARG;
    int64_t out_buf_size, padding_w;
    if (padding) {
        out_buf_size = batches * im_w * filter_w * channels;
        padding_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_w = im_w;
    } else {
        out_buf_size = batches * out_w * filter_w * channels;
        padding_w = 0;
    }
    if (cols_size != out_buf_size || count_cols - cols_offset != out_buf_size) return RO_E_ARG;
    int64_t im_w_step, channel_step, batch_step;
    if (channels_first) { im_w_step = 1; channel_step = im_w; batch_step = im_w * channels; }
    else { channel_step = 1; im_w_step = channels; batch_step = channels * im_w; }
    int64_t stride_w_step = im_w_step * stride_w, filter_w_step = im_w_step * dilation_w;
    int64_t out_filter_step, out_channel_step;
    if (cols_channels_first) { out_filter_step = 1; out_channel_step = filter_w; }
    else { out_filter_step = channels; out_channel_step = 1; }
    int64_t out_cell_step = filter_w * channels;
    int64_t batch_pos = images_offset - im_w_step * padding_w;
    int64_t out_pos = cols_offset;
    int64_t vim_w = out_w * stride_w, vfilter_w = filter_w * dilation_w;
    for (int64_t batch = 0; batch < batches; batch++) {
        int64_t stride_w_pos = batch_pos;
        for (int64_t vim_x = 0; vim_x < vim_w; vim_x += stride_w) {
            /* copyCell1d */
            int64_t filter_w_pos = stride_w_pos, out_filter_pos = out_pos;
            for (int64_t vfx = 0; vfx < vfilter_w; vfx += dilation_w) {
                int64_t channel_pos = filter_w_pos, out_channel_pos = out_filter_pos;
                int64_t input_x = vim_x - padding_w + vfx;
                for (int64_t c = 0; c < channels; c++) {
                    if (input_x < 0 || input_x >= im_w) {
                        if (!reverse) cols[out_channel_pos] = 0;
                    } else if (!reverse) {
                        cols[out_channel_pos] = images[channel_pos];
                    } else {
                        images[channel_pos] += cols[out_channel_pos];
                    }
                    out_channel_pos += out_channel_step;
                    channel_pos += channel_step;
                }
                out_filter_pos += out_filter_step;
                filter_w_pos += filter_w_step;
            }
            stride_w_pos += stride_w_step;
            out_pos += out_cell_step;
        }
        batch_pos += batch_step;
    }
    return RO_OK;
}

/* PhpMath::im2col3d, PhpMath.php:2396-2600 (copyCell3d :2319-2390).  padding_d uses im_h as the reference does (:2459),
 * and only out_h / out_w are tested (:2453). */
int ro_im2col3d(int reverse,
                double *images, int64_t count_images, int64_t images_offset, int64_t images_size,
                int64_t batches, int64_t im_d, int64_t im_h, int64_t im_w, int64_t channels,
                int64_t filter_d, int64_t filter_h, int64_t filter_w,
                int64_t stride_d, int64_t stride_h, int64_t stride_w,
                int padding, int channels_first, int64_t dilation_d, int64_t dilation_h, int64_t dilation_w,
                int cols_channels_first,
                double *cols, int64_t count_cols, int64_t cols_offset, int64_t cols_size)
{
    int64_t images_buf_size = batches * im_d * im_h * im_w * channels;
    if (images_size != images_buf_size || count_images - images_offset < images_buf_size) return RO_E_ARG;
    int64_t out_d = (im_d - (filter_d - 1) * dilation_d - 1) / stride_d + 1;
    int64_t out_h = (im_h - (filter_h - 1) * dilation_h - 1) / stride_h + 1;
    int64_t out_w = (im_w - (filter_w - 1) * dilation_w - 1) / stride_w + 1;
    if (out_h <= 0 || out_w <= 0) return RO_E_ARG;
    int64_t out_buf_size, padding_d, padding_h, padding_w;
    if (padding) {
        out_buf_size = batches * im_d * filter_d * im_h * filter_h * im_w * filter_w * channels;
        padding_d = ((im_d - 1) * stride_d - im_h + (filter_d - 1) * dilation_d + 1) / 2;
        padding_h = ((im_h - 1) * stride_h - im_h + (filter_h - 1) * dilation_h + 1) / 2;
        padding_w = ((im_w - 1) * stride_w - im_w + (filter_w - 1) * dilation_w + 1) / 2;
        out_d = im_d; out_h = im_h; out_w = im_w;
    } else {
        out_buf_size = batches * out_d * filter_d * out_h * filter_h * out_w * filter_w * channels;
        padding_d = padding_h = padding_w = 0;
    }
    if (cols_size != out_buf_size || count_cols - cols_offset != out_buf_size) return RO_E_ARG;
    int64_t im_w_step, im_h_step, im_d_step, channel_step, batch_step;
    if (channels_first) {
        im_w_step = 1; im_h_step = im_w; im_d_step = im_w * im_h; channel_step = im_w * im_h * im_d;
        batch_step = im_w * im_h * im_d * channels;
    } else {
        channel_step = 1; im_w_step = channels; im_h_step = channels * im_w; im_d_step = channels * im_w * im_h;
        batch_step = channels * im_w * im_h * im_d;
    }
    int64_t stride_w_step = im_w_step * stride_w, stride_h_step = im_h_step * stride_h, stride_d_step = im_d_step * stride_d;
    int64_t filter_w_step = im_w_step * dilation_w, filter_h_step = im_h_step * dilation_h, filter_d_step = im_d_step * dilation_d;
    int64_t out_filter_step, out_channel_step;
    if (cols_channels_first) { out_filter_step = 1; out_channel_step = filter_d * filter_h * filter_w; }
    else { out_filter_step = channels; out_channel_step = 1; }
    int64_t out_cell_step = filter_d * filter_h * filter_w * channels;
    int64_t batch_pos = images_offset - im_d_step * padding_d - im_h_step * padding_h - im_w_step * padding_w;
    int64_t out_pos = cols_offset;
    int64_t vim_d = out_d * stride_d, vim_h = out_h * stride_h, vim_w = out_w * stride_w;
    int64_t vfilter_d = filter_d * dilation_d, vfilter_h = filter_h * dilation_h, vfilter_w = filter_w * dilation_w;
    for (int64_t batch = 0; batch < batches; batch++) {
        int64_t stride_d_pos = batch_pos;
        for (int64_t vim_z = 0; vim_z < vim_d; vim_z += stride_d) {
            int64_t stride_h_pos = stride_d_pos;
            for (int64_t vim_y = 0; vim_y < vim_h; vim_y += stride_h) {
                int64_t stride_w_pos = stride_h_pos;
                for (int64_t vim_x = 0; vim_x < vim_w; vim_x += stride_w) {
                    /* copyCell3d */
                    int64_t filter_d_pos = stride_w_pos, out_filter_pos = out_pos;
                    for (int64_t vfz = 0; vfz < vfilter_d; vfz += dilation_d) {
                        int64_t filter_h_pos = filter_d_pos;
                        for (int64_t vfy = 0; vfy < vfilter_h; vfy += dilation_h) {
                            int64_t filter_w_pos = filter_h_pos;
                            for (int64_t vfx = 0; vfx < vfilter_w; vfx += dilation_w) {
                                int64_t channel_pos = filter_w_pos, out_channel_pos = out_filter_pos;
                                int64_t iz = vim_z - padding_d + vfz, iy = vim_y - padding_h + vfy, ix = vim_x - padding_w + vfx;
                                for (int64_t c = 0; c < channels; c++) {
                                    if (iz < 0 || iz >= im_d || iy < 0 || iy >= im_h || ix < 0 || ix >= im_w) {
                                        if (!reverse) cols[out_channel_pos] = 0;
                                    } else if (!reverse) {
                                        cols[out_channel_pos] = images[channel_pos];
                                    } else {
                                        images[channel_pos] += cols[out_channel_pos];
                                    }
                                    out_channel_pos += out_channel_step;
                                    channel_pos += channel_step;
                                }
                                out_filter_pos += out_filter_step;
                                filter_w_pos += filter_w_step;
                            }
                            filter_h_pos += filter_h_step;
                        }
                        filter_d_pos += filter_d_step;
                    }
                    stride_w_pos += stride_w_step;
                    out_pos += out_cell_step;
                }
                stride_h_pos += stride_h_step;
            }
            stride_d_pos += stride_d_step;
        }
        batch_pos += batch_step;
    }
    return RO_OK;
}

/* ---- BLAS level 1 / 2 and omatcopy (SURVEY.md section 8f widening) ------------------------------- */

/* PhpBlas::scal, PhpBlas.php:141-161 (real path :156-159): X[i] = X[i] * alpha */
int ro_scal(int64_t n, double alpha, double *X, int64_t countX, int64_t offX, int64_t incX)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    for (int64_t i = 0, id = offX; i < n; i++, id += incX) X[id] = X[id] * alpha;
    return RO_OK;
}

/* PhpBlas::axpy, PhpBlas.php:165-194: alpha == 1 adds, otherwise one mul + one add */
int ro_axpy(int64_t n, double alpha, const double *X, int64_t countX, int64_t offX, int64_t incX,
            double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    if (offY < 0 || incY < 1 || offY + (n - 1) * incY >= countY) return RO_E_ARG;
    int64_t ix = offX, iy = offY;
    if (alpha == 1.0) { for (int64_t i = 0; i < n; i++, ix += incX, iy += incY) Y[iy] = X[ix] + Y[iy]; }
    else { for (int64_t i = 0; i < n; i++, ix += incX, iy += incY) Y[iy] = alpha * X[ix] + Y[iy]; }
    return RO_OK;
}

/* PhpBlas::copy, PhpBlas.php:350-364 */
int ro_copy(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX,
            double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    if (offY < 0 || incY < 1 || offY + (n - 1) * incY >= countY) return RO_E_ARG;
    for (int64_t i = 0, ix = offX, iy = offY; i < n; i++, ix += incX, iy += incY) Y[iy] = X[ix];
    return RO_OK;
}

/* PhpBlas::dot, PhpBlas.php:196-222 (real path :216-220): acc += X*Y from 0.0 in index order */
int ro_dot(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX,
           const double *Y, int64_t countY, int64_t offY, int64_t incY, double *result)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    if (offY < 0 || incY < 1 || offY + (n - 1) * incY >= countY) return RO_E_ARG;
    double acc = 0.0;
    for (int64_t i = 0, ix = offX, iy = offY; i < n; i++, ix += incX, iy += incY) acc += X[ix] * Y[iy];
    *result = acc;
    return RO_OK;
}

/* PhpBlas::asum, PhpBlas.php:266-286 (real path :279-283): acc += abs(X) */
int ro_asum(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, double *result)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    double acc = 0.0;
    for (int64_t i = 0, ix = offX; i < n; i++, ix += incX) acc += fabs(X[ix]);
    *result = acc;
    return RO_OK;
}

/* PhpBlas::iamax / iamin, PhpBlas.php:288-317 / :319-348 (real paths :306-314 / :337-345): acc = abs(X[0]), idx = 0,
 * then a strict comparison per element -- a NaN at 0 is never displaced, a later NaN never wins */
int ro_iamax(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, int64_t *result)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    double acc = fabs(X[offX]);
    int64_t best = 0;
    for (int64_t i = 1, ix = offX + incX; i < n; i++, ix += incX) {
        const double a = fabs(X[ix]);
        if (acc < a) { acc = a; best = i; }
    }
    *result = best;
    return RO_OK;
}

int ro_iamin(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, int64_t *result)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    double acc = fabs(X[offX]);
    int64_t best = 0;
    for (int64_t i = 1, ix = offX + incX; i < n; i++, ix += incX) {
        const double a = fabs(X[ix]);
        if (acc > a) { acc = a; best = i; }
    }
    *result = best;
    return RO_OK;
}

/* PhpBlas::nrm2, PhpBlas.php:366-392 (real path :384-389): sqrt(sum v*v), no scaling */
int ro_nrm2(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, double *result)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    double sum = 0.0;
    for (int64_t i = 0, ix = offX; i < n; i++, ix += incX) { const double v = X[ix]; sum += v * v; }
    *result = sqrt(sum);
    return RO_OK;
}

/* PhpBlas::swap, PhpBlas.php:744-760 */
int ro_swap(int64_t n, double *X, int64_t countX, int64_t offX, int64_t incX,
            double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (n < 1 || offX < 0 || incX < 1 || offX + (n - 1) * incX >= countX) return RO_E_ARG;
    if (offY < 0 || incY < 1 || offY + (n - 1) * incY >= countY) return RO_E_ARG;
    for (int64_t i = 0, ix = offX, iy = offY; i < n; i++, ix += incX, iy += incY) {
        const double tmp = Y[iy];
        Y[iy] = X[ix];
        X[ix] = tmp;
    }
    return RO_OK;
}

/* PhpBlas::gemv, PhpBlas.php:762-844 (real path :826-842): acc += (alpha*A)*X in index order, then
 * Y = acc + beta*Y iff beta != 0.  order/trans are the CBLAS codes. */
int ro_gemv(int order, int trans, int64_t m, int64_t n, double alpha,
            const double *A, int64_t countA, int64_t offA, int64_t ldA,
            const double *X, int64_t countX, int64_t offX, int64_t incX,
            double beta, double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (order == 102) { int64_t t = m; m = n; n = t; } else if (order != 101) return RO_E_ARG;
    int t;
    if (trans == 111 || trans == 114) t = 0; else if (trans == 112 || trans == 113) t = 1; else return RO_E_ARG;
    int64_t rows = !t ? m : n, cols = !t ? n : m;
    if (m < 1 || n < 1) return RO_E_ARG;
    if (offA < 0 || ldA < 1 || offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG;
    if (offX < 0 || incX < 1 || offX + (cols - 1) * incX >= countX) return RO_E_ARG;
    if (offY < 0 || incY < 1 || offY + (rows - 1) * incY >= countY) return RO_E_ARG;
    int64_t ldA_i = !t ? ldA : 1, ldA_j = !t ? 1 : ldA;
    int64_t idA_i = offA, idY = offY;
    for (int64_t i = 0; i < rows; i++, idA_i += ldA_i, idY += incY) {
        int64_t idA = idA_i, idX = offX;
        double acc = 0.0;
        for (int64_t j = 0; j < cols; j++, idA += ldA_j, idX += incX) acc += alpha * A[idA] * X[idX];
        Y[idY] = (beta != 0.0) ? acc + beta * Y[idY] : acc;
    }
    return RO_OK;
}

/* PhpBlas::omatcopy, PhpBlas.php:1550-1612 (real path :1603-1610): B[i,j] = alpha * op(A)[i,j] */
int ro_omatcopy(int order, int trans, int64_t m, int64_t n, double alpha,
                const double *A, int64_t countA, int64_t offA, int64_t ldA,
                double *B, int64_t countB, int64_t offB, int64_t ldB)
{
    if (order == 102) { int64_t t = m; m = n; n = t; } else if (order != 101) return RO_E_ARG;
    int t;
    if (trans == 111 || trans == 114) t = 0; else if (trans == 112 || trans == 113) t = 1; else return RO_E_ARG;
    if (m < 1 || n < 1) return RO_E_ARG;
    int64_t rows = !t ? m : n, cols = !t ? n : m;
    if (offA < 0 || ldA < 1 || offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG;
    if (offB < 0 || ldB < 1 || offB + (rows - 1) * ldB + (cols - 1) >= countB) return RO_E_ARG;
    int64_t ldA_i = !t ? ldA : 1, ldA_j = !t ? 1 : ldA;
    int64_t idA_i = offA, idB_i = offB;
    for (int64_t i = 0; i < rows; i++, idA_i += ldA_i, idB_i += ldB) {
        int64_t idA = idA_i, idB = idB_i;
        for (int64_t j = 0; j < cols; j++, idA += ldA_j, idB++) B[idB] = alpha * A[idA];
    }
    return RO_OK;
}

/* ---- the rest of the broadcast family: PhpMath.php:274-528 (maximum, minimum,
 *      greater, greaterEqual, less, lessEqual: A[m,n] against X[n], no $trans), :687-722 (pow) and :860-891
 *      (duplicate), which take $trans like add.  op codes = RB200_BC_* of include/rindow_b200.h. ------------ */
int ro_broadcast(int op, int trans, int64_t m, int64_t n,
                 double *A, int64_t countA, int64_t offA, int64_t ldA,
                 const double *X, int64_t countX, int64_t offX, int64_t incX)
{
    if (op <= 5) {                                               /* maximum .. lessEqual */
        if (trans) return RO_E_ARG;                              /* these methods have no $trans */
        if (m <= 0 || n <= 0) return RO_E_ARG;                   /* :285-287 */
        if ((m - 1) * ldA + (n - 1) + offA >= countA) return RO_E_ARG;   /* "Matrix A is too small" */
        if ((n - 1) * incX + offX >= countX) return RO_E_ARG;            /* "Buffer X is too small" */
        int64_t lna = offA;
        for (int64_t i = 0; i < m; i++, lna += ldA) {
            int64_t idx = offX, ida = lna;
            for (int64_t j = 0; j < n; j++, idx += incX, ida++) {
                double v = X[idx];
                switch (op) {
                    case 0: if (isnan(v)) A[ida] = v; else if (A[ida] < v) A[ida] = v; break;   /* :300-309 */
                    case 1: if (isnan(v)) A[ida] = v; else if (A[ida] > v) A[ida] = v; break;   /* :344-353 */
                    case 2: A[ida] = (A[ida] >  v) ? 1.0 : 0.0; break;                          /* :390-395 */
                    case 3: A[ida] = (A[ida] >= v) ? 1.0 : 0.0; break;
                    case 4: A[ida] = (A[ida] <  v) ? 1.0 : 0.0; break;
                    default: A[ida] = (A[ida] <= v) ? 1.0 : 0.0; break;
                }
            }
        }
        return RO_OK;
    }
    if (op != 6 && op != 7) return RO_E_ARG;
    int64_t rows = !trans ? m : n, cols = !trans ? n : m;       /* :700-712 / :873-885 */
    if (offX + (cols - 1) * incX >= countX) return RO_E_ARG;
    if (offA + (m - 1) * ldA + (n - 1) >= countA) return RO_E_ARG;
    int64_t incAj = !trans ? ldA : 1, incAi = !trans ? 1 : ldA;
    int64_t idAj = offA;
    for (int64_t j = 0; j < rows; j++, idAj += incAj) {
        int64_t idA = idAj, idX = offX;
        for (int64_t i = 0; i < cols; i++, idA += incAi, idX += incX) {
            A[idA] = (op == 6) ? pow(A[idA], X[idX]) : X[idX];
        }
    }
    return RO_OK;
}

/* ---- unary family, in place: PhpMath.php:221-268 (increment, reciprocal), :611-682 (square, sqrt, rsqrt),
 *      :727-855 (exp, log, tanh, sin, cos, tan), :1528-1547 (not), :2830-2875 (nan2num, isnan).
 *      op codes = RB200_UN_*. ------------------------------------------------------------------------------- */
int ro_unary(int op, int64_t n, double alpha, double *X, int64_t countX, int64_t offX, int64_t incX, double beta)
{
    if (offX + (n - 1) * incX >= countX) return RO_E_ARG;
    int64_t idx = offX;
    for (int64_t i = 0; i < n; i++, idx += incX) {
        double t = X[idx], r;
        switch (op) {
            case 0: r = alpha * t + beta; break;
            case 1: t = alpha * t + beta; r = (t == 0.0) ? INFINITY : 1 / t; break;
            case 2: r = t * t; break;
            case 3: r = sqrt(t); break;
            case 4: t = alpha * sqrt(t) + beta; r = (t != 0.0) ? 1 / t : INFINITY; break;
            case 5: r = exp(t); break;
            case 6: r = log(t); break;
            case 7: r = tanh(t); break;
            case 8: r = sin(t); break;
            case 9: r = cos(t); break;
            case 10: r = tan(t); break;
            case 11: r = (t == 0.0) ? 1.0 : 0.0; break;
            case 12: r = isnan(t) ? alpha : t; break;
            case 13: r = isnan(t) ? 1.0 : 0.0; break;
            default: return RO_E_ARG;
        }
        X[idx] = r;
    }
    return RO_OK;
}

/* equal / notEqual: PhpMath.php:1470-1522.  ne = 0: Y := (Y == X) ? 1 : 0;  ne = 1: Y := (Y != X) ? 1 : 0 */
int ro_compare(int ne, int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX,
               double *Y, int64_t countY, int64_t offY, int64_t incY)
{
    if (offX + (n - 1) * incX >= countX) return RO_E_ARG;
    if (offY + (n - 1) * incY >= countY) return RO_E_ARG;
    int64_t idX = offX, idY = offY;
    for (int64_t i = 0; i < n; i++, idX += incX, idY += incY) {
        if (!ne) Y[idY] = (Y[idY] == X[idX]) ? 1.0 : 0.0;
        else     Y[idY] = (Y[idY] != X[idX]) ? 1.0 : 0.0;
    }
    return RO_OK;
}

/* sum / imax / imin: PhpMath.php:150-216 */
int ro_sum(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, double *result)
{
    if (offX + (n - 1) * incX >= countX) return RO_E_ARG;
    *result = ro_math_sum(n, X, offX, incX);
    return RO_OK;
}

int ro_imax(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, int64_t *result)
{
    if (offX + (n - 1) * incX >= countX) return RO_E_ARG;
    int64_t idx = offX + incX, best = 0;
    double acc = X[offX];
    for (int64_t i = 1; i < n; i++, idx += incX) {
        if (acc < X[idx] || isnan(acc)) {                        /* :185 */
            acc = X[idx];
            best = i;
        }
    }
    *result = best;
    return RO_OK;
}

int ro_imin(int64_t n, const double *X, int64_t countX, int64_t offX, int64_t incX, int64_t *result)
{
    if (offX + (n - 1) * incX >= countX) return RO_E_ARG;
    int64_t idx = offX + incX, best = 0;
    double acc = X[offX];
    for (int64_t i = 1; i < n; i++, idx += incX) {
        if (acc > X[idx]) {                                      /* :210 */
            acc = X[idx];
            best = i;
        }
    }
    *result = best;
    return RO_OK;
}

#define RO_E_RANGE -3   /* RuntimeException('Label number is out of bounds.') */
/* updateAddOnehot: PhpMath.php:1438-1464 (rows before a bad label are already updated, as in the PHP loop) */
int ro_onehot(int64_t m, int64_t n, double a, const double *X, int64_t countX, int64_t offX, int64_t incX,
              double *Y, int64_t countY, int64_t offY, int64_t ldY)
{
    if (offX + (m - 1) * incX >= countX) return RO_E_ARG;
    if (offY + (m - 1) * ldY + (n - 1) >= countY) return RO_E_ARG;
    int64_t idx = offX, idy = offY;
    for (int64_t i = 0; i < m; i++, idy += ldY, idx += incX) {
        int64_t label = (int64_t)X[idx];
        if (label >= n || label < 0) return RO_E_RANGE;
        Y[idy + label] = Y[idy + label] + a;
    }
    return RO_OK;
}

/* ---- N-D transpose: PhpMath.php:2959-3025 (strides) + transCopy :3043-3067 (the recursive copy) --------------- */
static void ro_trans_copy(const double *A, int64_t offA, double *B, int64_t offB, int depth, int ndim,
                          const int64_t *shape, const int64_t *strides, const int64_t *tstrides)
{
    const int64_t repeat = shape[depth], stride = strides[depth], tstride = tstrides[depth];
    if (depth == ndim - 1) {                                     /* math_copy(repeat, A, offA, stride, B, offB, tstride) */
        for (int64_t i = 0; i < repeat; i++) B[offB + i * tstride] = A[offA + i * stride];
        return;
    }
    for (int64_t pos = 0; pos < repeat; pos++) {
        ro_trans_copy(A, offA + stride * pos, B, offB + tstride * pos, depth + 1, ndim, shape, strides, tstrides);
    }
}

int ro_transpose(int ndim, const int64_t *shape, const int64_t *perm,
                 const double *A, int64_t countA, int64_t offA, double *B, int64_t countB, int64_t offB)
{
    int64_t strides[16], tstrides[16];
    if (ndim <= 0 || ndim > 16) return RO_E_ARG;                 /* 'Matrix must not be a scalar' */
    int64_t stride = 1, tstride = 1;
    for (int d = 0; d < ndim; d++) tstrides[d] = -1;
    for (int d = ndim - 1; d >= 0; d--) {                        /* :2991-3001 */
        strides[d] = stride;
        stride *= shape[d];
        int64_t targ = perm[d];
        if (targ >= ndim || targ < 0) return RO_E_ARG;           /* 'dim value in the perm is out of range.' */
        tstrides[targ] = tstride;
        tstride *= shape[targ];
    }
    if (stride != tstride) return RO_E_ARG;                      /* 'duplicate axis in perm option' */
    for (int d = 0; d < ndim; d++) if (tstrides[d] < 0) return RO_E_ARG;
    if (offA + stride > countA) return RO_E_ARG;
    if (offB + tstride > countB) return RO_E_ARG;
    ro_trans_copy(A, offA, B, offB, 0, ndim, shape, strides, tstrides);
    return RO_OK;
}

/* ---- gather / scatterAdd2 / reduceGather: PhpMath.php:1063-1130, :1362-1395, :1135-1208.  X holds PHP numbers
 *      (doubles here); the index >= numClass clamp is the reference's (:1107-1111). ------------------------------ */
int ro_gather(int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass,
              const double *X, int64_t countX, int64_t offX, double *A, int64_t countA, int64_t offA,
              double *B, int64_t countB, int64_t offB)
{
    if (offX + n > countX) return RO_E_ARG;
    if (offA + numClass * k > countA) return RO_E_ARG;
    if (offB + n * k > countB) return RO_E_ARG;
    if (reverse && add_mode) {                                   /* scatterAdd2: column by column, j ascending */
        for (int64_t i = 0; i < k; i++) {
            for (int64_t j = 0; j < n; j++) {
                int64_t index = (int64_t)X[offX + j];
                if (index >= numClass) index = numClass - 1;
                A[offA + index * k + i] = A[offA + index * k + i] + B[offB + j * k + i];
            }
        }
        return RO_OK;
    }
    if (numClass <= 0) return RO_E_ARG;
    for (int64_t j = 0; j < n; j++) {
        int64_t index = (int64_t)X[offX + j];
        if (index >= numClass) index = numClass - 1;
        double *a = A + offA + index * k, *b = B + offB + j * k;
        for (int64_t t = 0; t < k; t++) {
            if (reverse) { if (add_mode) a[t] = a[t] + b[t]; else a[t] = b[t]; }          /* math_add / math_copy */
            else         { if (add_mode) b[t] = b[t] + a[t]; else b[t] = a[t]; }
        }
    }
    return RO_OK;
}

int ro_reduce_gather(int reverse, int add_mode, int64_t m, int64_t n, int64_t numClass,
                     const double *X, int64_t countX, int64_t offX, double *A, int64_t countA, int64_t offA,
                     double *B, int64_t countB, int64_t offB)
{
    if (offX + n > countX) return RO_E_ARG;                      /* :1164 (sic: n, not m*n) */
    if (offX + m * n > countX) return RO_E_ARG;                  /* what the loop actually touches */
    if (offA + m * numClass > countA) return RO_E_ARG;           /* :1166 */
    if (offA + m * numClass * n > countA) return RO_E_ARG;
    if (offB + m * n > countB) return RO_E_ARG;
    if (numClass <= 0) return RO_E_ARG;
    int64_t idxX = offX, idxA = offA, idxB = offB;
    for (int64_t i = 0; i < m; i++, idxX += n, idxA += n * numClass, idxB += n) {
        for (int64_t j = 0; j < n; j++) {
            int64_t index = (int64_t)X[idxX + j];
            if (index >= numClass) index = numClass - 1;
            int64_t iA = idxA + j + index * n, iB = idxB + j;
            if (reverse) { if (add_mode) A[iA] = A[iA] + B[iB]; else A[iA] = B[iB]; }
            else         { if (add_mode) B[iB] = B[iB] + A[iA]; else B[iB] = A[iA]; }
        }
    }
    return RO_OK;
}

int ro_abi_version(void) { return 1; }
