/*
 * rindow_b200.h -- C ABI of librindow_b200.so, the B200 (sm_100a) compute driver for
 * rindow-math-matrix.
 *
 * This is the drop-in boundary: the entry points below are what a
 * PHP-FFI (`FFI::cdef`) shim class -- MatlibCuda\{CudaBlas,CudaMath,CudaBuffer} plugged
 * into Rindow\Math\Matrix\Drivers\Service (reference src/Drivers/Service.php:4-26) via
 * AbstractMatlibService's three factories (src/Drivers/AbstractMatlibService.php:43-75)
 * -- binds, one per reference driver method on the hot path.  In this repository the
 * same ABI is bound by ctypes (rindow_math_matrix_b200/_capi.py), because the image has
 * no PHP; php/MatlibCuda holds the PHP side a maintainer would add.
 *
 * Conventions
 *  - extern "C", plain pointers and 64-bit sizes; no C++/torch types; no callbacks.
 *  - every function returns 0 on success and a negative rb200 status on failure;
 *    rb200_last_error() gives the message (thread-local).  No C++ exception crosses.
 *  - device pointers are plain `void*` from rb200_buffer_alloc (or any cudaMalloc'd /
 *    torch-owned allocation on the current device).  Offsets `off*`, leading dimensions
 *    `ld*` and increments `inc*` are in ELEMENTS, exactly as in the reference driver
 *    signatures; rb200_buffer_* offsets are in BYTES like the reference's DeviceBuffer
 *    (src/NDArrayCL.php:261-270).
 *  - all compute calls are asynchronous on the library stream (rb200_set_stream to
 *    adopt the caller's); rb200_sync() before the host reads results.  Buffer h2d/d2h
 *    are synchronous on return.
 *  - argument VALIDATION with the reference's exact exception strings lives in the host
 *    shim (Utils.php:30-65 semantics); the ABI re-checks only what would make the
 *    device code fault (null pointers, non-positive sizes, unsupported dtype).
 *  - not re-entrant: one host thread drives one device (PHP CLI is single-threaded,
 *    PhpBlas.php:49-54 reports 1 thread).  One process per GPU for multi-GPU.
 */
#ifndef RINDOW_B200_H_
#define RINDOW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_ABI_VERSION 1

/* status codes */
#define RB200_OK              0
#define RB200_E_INVALID      -1   /* bad argument (host shim should have thrown InvalidArgumentException) */
#define RB200_E_CUDA         -2   /* CUDA runtime/driver error; message in rb200_last_error() */
#define RB200_E_UNSUPPORTED  -3   /* dtype / mode not supported by this entry point */
#define RB200_E_NOTINIT      -4   /* rb200_init not called or no device */
#define RB200_E_NOMEM        -5

/* dtype codes.  The ABI owns this enum; the host shim maps Interop\Polite\Math\Matrix\
 * NDArray::float32 etc. to it symbolically in one table (the polite-math constant values
 * are not vendored under the reference). */
#define RB200_BOOL      0
#define RB200_INT8      1
#define RB200_INT16     2
#define RB200_INT32     3
#define RB200_INT64     4
#define RB200_UINT8     5
#define RB200_UINT16    6
#define RB200_UINT32    7
#define RB200_UINT64    8
#define RB200_FLOAT32   12
#define RB200_FLOAT64   13
/* complex storage (PhpBuffer.php:38-41, 60-61: complex_float 8 bytes, complex_double 16 bytes, interleaved real / imag
 * as dump() packs them :153-158).  Buffers of these dtypes can be allocated, filled, copied, swapped and moved between
 * host and device; the arithmetic entry points take real dtypes only (the reference's complex gemm branch,
 * PhpBlas.php:890-926, is outside the hot path). */
#define RB200_COMPLEX64  16
#define RB200_COMPLEX128 17

/* CBLAS-valued enums (reference: Interop\Polite\Math\Matrix\BLAS; LinearAlgebra.php:16-17) */
#define RB200_ROW_MAJOR      101
#define RB200_COL_MAJOR      102
#define RB200_NO_TRANS       111
#define RB200_TRANS          112
#define RB200_CONJ_TRANS     113
#define RB200_CONJ_NO_TRANS  114

/* sgemm arithmetic modes (BASELINE.json north_star: "TF32 or 3xTF32 mode ... plus FFMA") */
#define RB200_GEMM_AUTO      0    /* library default: see rb200_set_default_gemm_mode */
#define RB200_GEMM_FP32_SIMT 1    /* FFMA, fp32 accumulate, any shape/stride */
#define RB200_GEMM_TF32      2    /* tcgen05 kind::tf32, 1 pass  (max rel err ~1e-3 of max|C|) */
#define RB200_GEMM_TF32X3    3    /* tcgen05 kind::tf32, 3 passes hi/lo split (~fp32 accuracy) */

/* ---- library / device ------------------------------------------------------------- */
int         rb200_abi_version(void);
const char* rb200_last_error(void);
int         rb200_device_count(int* count);
/* Select `device`, create the library stream, query SM count.  Idempotent per device. */
int         rb200_init(int device);
int         rb200_shutdown(void);
int         rb200_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor,
                              int64_t* total_mem_bytes);
/* Adopt a caller-owned cudaStream_t (e.g. torch's current stream; 0 = the legacy default stream);
 * (void*)-1 restores the library-owned stream. */
int         rb200_set_stream(void* cuda_stream);
int         rb200_get_stream(void** cuda_stream);
int         rb200_sync(void);
int         rb200_set_default_gemm_mode(int mode);
int         rb200_get_default_gemm_mode(int* mode);
/* How many kernels this library has launched since init / since the last reset
 * (bench.py's "gpu_launches"). */
int         rb200_launch_count(int64_t* count, int reset);
/* Name of the kernel family the last rb200_sgemm/dgemm/conv2d_forward dispatched to: "tcgen05_tf32_2cta",
 * "tcgen05_tf32", "tcgen05_tf32_128", "tcgen05_tf32_64", "tcgen05_tf32x3", "tcgen05_tf32x3_64", "tcgen05_tf32x3_2cta", "tcgen05_tf32[x3]_skinny", "tcgen05_tf32[x3]_conv", "simt_f32_shortk",
 * "simt_f32", "simt_f64" (64 x 64 tiles: small / odd problems), "simt_f32_big" (128 x 128 x 16 double-buffered FFMA
 * tiles: every float32 FFMA problem that fills the machine), "dmma_f64" (128 x 64 x 16 tiles on the FP64 tensor pipe,
 * mma.sync m16n8k8: every float64 problem that fills the machine; "simt_f64_big" is its DFMA form, RB200_DGEMM_DMMA=0). */
const char* rb200_last_gemm_path(void);

/* ---- buffers: replaces PhpBLASFactory::Buffer (PhpBLASFactory.php:35) storage and the
 *      DeviceBuffer read/write/copy/fill calls of NDArrayCL (NDArrayCL.php:147-168,
 *      252-272, 390-442) ------------------------------------------------------------- */
/* Buffers come from the device's stream-ordered memory pool (cudaMallocAsync, freed memory stays cached) and are
 * released without a synchronisation: LinearAlgebra allocates an output per call (src/LinearAlgebra.php:255-260).
 * rb200_buffer_alloc_shareable uses plain cudaMalloc: the allocations rb200_ipc_export can export. */
int rb200_buffer_alloc(int64_t bytes, void** dptr);
int rb200_buffer_alloc_shareable(int64_t bytes, void** dptr);
int rb200_buffer_free(void* dptr);
int rb200_buffer_h2d(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes);
int rb200_buffer_d2h(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes);
int rb200_buffer_h2d_async(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes);
int rb200_buffer_d2h_async(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes);
int rb200_buffer_d2d(void* dst, int64_t dst_offset_bytes, const void* src, int64_t src_offset_bytes,
                     int64_t bytes);
/* fill n elements of `dtype` starting at element offset `off` with *value (value points to
 * one element of that dtype on the host).  Replaces PhpMath::fill (PhpMath.php:2811) /
 * zeros (:908) for the alloc+zeros every LinearAlgebra::gemm/im2col does. */
int rb200_buffer_fill(int dtype, void* dptr, int64_t off, int64_t n, const void* value);
/* pinned host staging memory for the host-indexable mirror of an LV_ADVANCED buffer */
int rb200_host_alloc(int64_t bytes, void** hptr);
int rb200_host_free(void* hptr);
/* Stream-ordered completion markers: the counterpart of the OpenCL event lists the reference's device
 * variants take as trailing ($events, $waitEvents) arguments (OpenCLMath.php:1851-1859,
 * LinearAlgebraCL.php:1867-1875) and of DeviceBuffer::read/write's blocking flag (NDArrayCL.php:261-270).
 * record = "everything queued on the library stream so far"; synchronize blocks the host until then;
 * query sets *done to 1/0 without blocking; wait makes the library stream wait for an event recorded
 * earlier (possibly while another stream was adopted with rb200_set_stream).  The host mirror of a
 * CudaBuffer records one after its asynchronous upload and waits on it before the host writes the
 * pinned memory again. */
int rb200_event_create(void** event);
int rb200_event_destroy(void* event);
int rb200_event_record(void* event);
int rb200_event_synchronize(void* event);
int rb200_event_query(void* event, int* done);
int rb200_event_wait(void* event);
/* CUDA-graph capture of a launch-bound sequence of calls (many small gemm / elementwise calls whose kernels run for a
 * few microseconds: the host-side cost of issuing them dominates).  Between begin and end every compute call on the
 * library stream is RECORDED instead of executed; rb200_graph_launch replays the recorded kernels with one driver
 * call.  Pointers, sizes and scalars are baked in.  Calls that allocate or synchronise are not capturable: issue the
 * sequence once before capturing (workspace and kernel attributes are then in place) and do not call the
 * scalar-returning entry points (asum, sum, imax ...) or buffer h2d / d2h inside a capture.  The reference has no
 * counterpart (its OpenCL queue is issued call by call, LinearAlgebraCL.php:145-152); a PHP caller would wrap a
 * training step's driver calls the same way. */
int rb200_graph_begin(void);
int rb200_graph_end(void** graph_exec);
int rb200_graph_launch(void* graph_exec);
int rb200_graph_destroy(void* graph_exec);

/* ---- BLAS: replaces PhpBlas::gemm (src/Drivers/MatlibPHP/PhpBlas.php:849-948) ------- */
/* C[m,n] = alpha * op(A)[m,k] * op(B)[k,n] (+ beta * C iff beta != 0; C is not read when
 * beta == 0, PhpBlas.php:940-944).  Row-major fast path; RB200_COL_MAJOR swaps m,n like
 * PhpBlas.php:862-863.  ConjTrans == Trans, ConjNoTrans == NoTrans for real dtypes
 * (PhpBlas.php:104-123). */
int rb200_sgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                float alpha, const float* A, int64_t offA, int64_t ldA,
                const float* B, int64_t offB, int64_t ldB,
                float beta, float* C, int64_t offC, int64_t ldC, int mode);
int rb200_dgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                double alpha, const double* A, int64_t offA, int64_t ldA,
                const double* B, int64_t offB, int64_t ldB,
                double beta, double* C, int64_t offC, int64_t ldC);

/* `batch` row-major products C_i = alpha * op(A_i) * op(B_i) (+ beta * C_i) with A_i = A + offA + i*strideA etc.
 * (strides in elements; strideA or strideB == 0 shares that operand across the batch) -- the inner loop of
 * LinearAlgebra::matmul (src/LinearAlgebra.php:1091-1106; the reference's accelerated driver has the same entry:
 * LinearAlgebraCL::matmul -> gemmStridedBatched, src/LinearAlgebraCL.php:2001).  dtype float32 (mode as in
 * rb200_sgemm) or float64. */
int rb200_gemm_strided_batched(int dtype, int transA, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                               const void* A, int64_t offA, int64_t ldA, int64_t strideA,
                               const void* B, int64_t offB, int64_t ldB, int64_t strideB, double beta,
                               void* C, int64_t offC, int64_t ldC, int64_t strideC, int64_t batch, int mode);

/* sgemm for operands whose current copy is in (pinned) HOST memory -- the e2e form of PhpBlas::gemm for a
 * caller that fills buffers from PHP and reads the product back (MatrixOperator::cross followed by
 * toArray(), MatrixOperator.php:429-508).  dX is the device allocation, hX the host mirror with the same
 * element indexing; hA / hB NULL = the device copy is current, hC NULL = leave C on the device.  The call
 * pipelines the transfers with the multiply: B first, then A (and C when beta != 0 and hC_is_current) in
 * row stripes on a copy-in stream, one sgemm per landed stripe on the library stream, finished C stripes
 * back on a copy-out stream (both PCIe directions busy).  Returns after hC is complete.  Same arithmetic,
 * modes and argument rules as rb200_sgemm. */
int rb200_sgemm_streamed(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                         float alpha, float* dA, const float* hA, int64_t offA, int64_t ldA,
                         float* dB, const float* hB, int64_t offB, int64_t ldB,
                         float beta, float* dC, float* hC, int hC_is_current, int64_t offC, int64_t ldC, int mode);

/* ---- multi-GPU: row-sharded sgemm whose C stripe is ALSO written into the peers' copies of
 *      the full C from the GEMM epilogue (peer stores over NVLink / NVSwitch), i.e. the all-gather of C overlaps
 *      the multiply tile by tile instead of following it as an NCCL call.  One process per GPU: a rank exports its
 *      full-C allocation with rb200_ipc_export (64-byte CUDA IPC handle, exchanged by the host -- e.g.
 *      torch.distributed all_gather_object), opens the other ranks' handles with rb200_ipc_open and passes the
 *      mapped base pointers as peer_C.  C / offC / ldC address this rank's stripe inside its own full C; every
 *      peer_C[j] is indexed with the same offC / ldC.  The caller orders the ranks afterwards (a barrier) before
 *      any rank reads rows it did not compute.  Same arithmetic, modes and argument rules as rb200_sgemm; shapes
 *      that do not take the CTA-pair tensor-core kernel are multiplied first and pushed with peer copies. ------ */
#define RB200_MAX_PEERS        15
#define RB200_IPC_HANDLE_BYTES 64
int rb200_ipc_export(void* dptr, void* handle64);
int rb200_ipc_open(const void* handle64, void** dptr);
int rb200_ipc_close(void* dptr);
int rb200_sgemm_allgather(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                          float alpha, const float* A, int64_t offA, int64_t ldA,
                          const float* B, int64_t offB, int64_t ldB,
                          float beta, float* C, int64_t offC, int64_t ldC, int mode,
                          int n_peers, void* const* peer_C);
/* The same call over an NVLS multicast mapping (NVSwitch): `mc_C` is the multicast alias of the symmetric allocation
 * that holds C on every rank (cuMulticastCreate / cuMulticastBindMem; torch.distributed._symmetric_memory provides it
 * as handle.multicast_ptr), indexed like C.  In the CTA-pair kernel a pusher warp per CTA reads every finished tile back
 * through L2 and stores it ONCE with multimem.st (512 contiguous bytes per instruction, decoupled from the TMEM drain by
 * a 4-tile mbarrier ring); the switch replicates it into every rank's copy (this rank's included) -- a stripe leaves the
 * GPU once instead of once per peer.  beta != 0 stores from the epilogue instead; other kernel families multiply first
 * and push the stripe with a multimem copy kernel on the same stream.  Same ordering contract as rb200_sgemm_allgather. */
int rb200_sgemm_allgather_mc(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                             float alpha, const float* A, int64_t offA, int64_t ldA,
                             const float* B, int64_t offB, int64_t ldB,
                             float beta, float* C, int64_t offC, int64_t ldC, int mode, void* mc_C);

/* ---- BLAS level 1 / 2: replace PhpBlas::scal (PhpBlas.php:141-161),
 *      axpy (:165-194), copy (:350-364), gemv (:762-844).  float32 / float64. ------------------------------ */
int rb200_scal(int dtype, int64_t n, double alpha, void* X, int64_t offX, int64_t incX);
int rb200_axpy(int dtype, int64_t n, double alpha, const void* X, int64_t offX, int64_t incX,
               void* Y, int64_t offY, int64_t incY);
int rb200_copy(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY);
/* X <-> Y, PhpBlas::swap (:744-760); float32 / float64; the vectors must not overlap. */
int rb200_swap(int dtype, int64_t n, void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t incY);
/* Y[rows] = alpha * op(A)[rows,cols] * X[cols] (+ beta*Y iff beta != 0); A is m x n row-major with ldA;
 * trans: rows = n, cols = m. */
int rb200_gemv(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
               const void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX,
               double beta, void* Y, int64_t offY, int64_t incY);
/* B[rows,cols] = alpha * op(A), A is m x n; replaces PhpBlas::omatcopy (PhpBlas.php:1550-1612), the transpose
 * LinearAlgebra::transpose2D issues (LinearAlgebra.php:4520-4565). */
int rb200_omatcopy(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
                   const void* A, int64_t offA, int64_t ldA, void* B, int64_t offB, int64_t ldB);

/* N-dimensional transpose B = permute(A, perm): B's axis d is A's axis perm[d] (shape and perm are HOST arrays of
 * ndim entries; A's shape is `shape`).  Replaces PhpMath::transpose (PhpMath.php:2959-3025), what
 * LinearAlgebra::transposeND (LinearAlgebra.php:4471-4519) calls for every rank != 2 and every integer dtype.
 * A pure copy: any dtype, bit-exact.  ndim <= 8. */
int rb200_transpose(int dtype, int ndim, const int64_t* shape, const int64_t* perm,
                    const void* A, int64_t offA, void* B, int64_t offB);

/* gather / scatter / scatterAdd along the first axis: replaces PhpMath::gather (PhpMath.php:1063-1130) incl. its
 * reverse+add form scatterAdd2 (:1362-1395).  X holds n indices (index_dtype: any real dtype; index >= numClass is
 * clamped to numClass-1 like the reference, a negative one to 0), A is [numClass, k], B is [n, k]:
 *   reverse=0 add=0: B[j,:] = A[X[j],:]      reverse=0 add=1: B[j,:] += A[X[j],:]
 *   reverse=1 add=0: A[X[j],:] = B[j,:] (duplicates: the last j wins, as in the sequential loop)
 *   reverse=1 add=1: A[X[j],:] += B[j,:] in ascending j (deterministic).
 * rb200_reduce_gather: B[i,j] = A[i, X[i,j], j] with A [m, numClass, n], X and B [m, n] (PhpMath::reduceGather
 * :1135-1208) and the same reverse / add forms.  dtype float32/float64/int32/int64/uint8/bool. */
int rb200_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass,
                 const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB);
int rb200_reduce_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t numClass,
                        const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB);

/* ---- Math: replaces PhpMath::add / multiply (PhpMath.php:570-606 / 530-565) ---------- */
/* trans==0: A[i,j] = alpha*X[j] + A[i,j]  (i<m, j<n, X has n elements)
 * trans!=0: A[i,j] = alpha*X[i] + A[i,j]  (X has m elements); A is m x n with row stride ldA */
int rb200_add(int dtype, int trans, int64_t m, int64_t n, double alpha,
              const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA);
int rb200_multiply(int dtype, int trans, int64_t m, int64_t n,
                   const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA);

/* ---- the rest of the broadcast family: replaces PhpMath::maximum (:274),
 *      minimum (:318), greater (:362), greaterEqual (:404), less (:446), lessEqual (:488), pow (:687),
 *      duplicate (:860).  Same A[m,n] / X broadcast rule as rb200_add (trans: X has m elements, one per row;
 *      the PHP methods without a $trans argument pass 0).  In place on A; float32/float64/int32/int64.
 *        MAXIMUM/MINIMUM: a NaN in X is written, a NaN already in A stays (:300-309).
 *        GREATER..LESS_EQUAL: A := (A op X) ? 1 : 0 in A's dtype; every NaN operand compares false.
 *        POW: A := A ** X.   DUPLICATE: A := X (A is not read). ------------------------------------------ */
#define RB200_BC_MAXIMUM        0
#define RB200_BC_MINIMUM        1
#define RB200_BC_GREATER        2
#define RB200_BC_GREATER_EQUAL  3
#define RB200_BC_LESS           4
#define RB200_BC_LESS_EQUAL     5
#define RB200_BC_POW            6
#define RB200_BC_DUPLICATE      7
int rb200_broadcast(int dtype, int op, int trans, int64_t m, int64_t n,
                    void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX);

/* ---- unary family, in place on X (n elements, stride incX): replaces PhpMath::increment (:221, X := alpha*X +
 *      beta), reciprocal (:244, 1/(alpha*X+beta), INF when the denominator is 0), square (:611), sqrt (:634),
 *      rsqrt (:657, 1/(alpha*sqrt(X)+beta), INF at 0), exp (:727), log (:749), tanh (:772), sin (:794), cos (:816),
 *      tan (:838), not (:1528, X := (X == 0) ? 1 : 0), nan2num (:2830, NaN -> alpha), isnan (:2853, 1.0 / 0.0).
 *      alpha / beta are ignored by the ops that have none.  float32/float64; NOT also int32/int64/bool/uint8. -- */
#define RB200_UN_INCREMENT   0
#define RB200_UN_RECIPROCAL  1
#define RB200_UN_SQUARE      2
#define RB200_UN_SQRT        3
#define RB200_UN_RSQRT       4
#define RB200_UN_EXP         5
#define RB200_UN_LOG         6
#define RB200_UN_TANH        7
#define RB200_UN_SIN         8
#define RB200_UN_COS         9
#define RB200_UN_TAN         10
#define RB200_UN_NOT         11
#define RB200_UN_NAN2NUM     12
#define RB200_UN_ISNAN       13
int rb200_unary(int dtype, int op, int64_t n, double alpha, void* X, int64_t offX, int64_t incX, double beta);

/* Y := (Y == X) ? 1 : 0  /  (Y != X) ? 1 : 0  -- PhpMath::equal (:1470) / notEqual (:1499); any dtype. */
#define RB200_CMP_EQUAL      0
#define RB200_CMP_NOT_EQUAL  1
int rb200_compare(int dtype, int op, int64_t n, const void* X, int64_t offX, int64_t incX,
                  void* Y, int64_t offY, int64_t incY);

/* Y[to_dtype] := cast(X[from_dtype]) -- PhpMath::astype (:1710-1778): to bool = (x != 0); to float = (float)x;
 * to int = truncation toward zero (PHP (int)), unsigned targets keep the low 8/16/32 bits (the $mask). */
int rb200_astype(int64_t n, int from_dtype, const void* X, int64_t offX, int64_t incX,
                 int to_dtype, void* Y, int64_t offY, int64_t incY);

/* ---- scalar results (the call returns after the value is on the host): PhpMath::sum (:150), imax (:171: first
 *      maximum, a NaN running value is replaced by the next element), imin (:196: first minimum, NaN at 0 stays).
 *      These are what MatrixOperator::sum/max/argMax/min/argMin call (MatrixOperator.php:820-931). --------- */
int rb200_sum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
int rb200_imax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
int rb200_imin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);

/* ---- BLAS level-1 reductions, PhpBlas.php: asum (:266-286) = sum |x|; nrm2 (:366-392) = sqrt(sum x*x); dot
 *      (:196-222) = sum x*y; iamax (:288-317) / iamin (:319-348) = index of the first largest / smallest |x|,
 *      scanning from element 0 with strict comparisons (a NaN at 0 stays, a later NaN never wins).  Every real
 *      dtype; terms and sums in double, as PHP evaluates them.  The call returns after the value is on the host. */
int rb200_asum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
int rb200_nrm2(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
int rb200_dot(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX,
              const void* Y, int64_t offY, int64_t incY, double* result);
int rb200_iamax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
int rb200_iamin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);

/* Y[i, (int)X[i]] += a for i < m -- PhpMath::updateAddOnehot (:1438-1464), what LinearAlgebra::onehot (:3095)
 * calls.  X holds the labels (any real dtype), Y is float32/float64 with row stride ldY.  A label outside
 * [0,n) returns RB200_E_RANGE (the reference throws RuntimeException 'Label number is out of bounds.'); the
 * call synchronises to report it. */
#define RB200_E_RANGE        -6
int rb200_onehot(int dtype_x, int dtype_y, int64_t m, int64_t n, double a,
                 const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t ldY);

/* ---- reductions over the middle axis of A[m,n,k] -> B[m,k]
 *      replaces PhpMath::reduceSum/reduceMax/reduceArgMax (PhpMath.php:1552-1643) ------- */
int rb200_reduce_sum(int dtype, int64_t m, int64_t n, int64_t k,
                     const void* A, int64_t offA, void* B, int64_t offB);
/* NaN-sticky (any NaN in the reduced run gives NaN): the rule the reference pins for its
 * accelerated drivers (LinearAlgebraTest.php:7617-7643, OpenCLMath.php:142-157); equals
 * PhpMath::math_max (:95-112) on NaN-free data and on rows whose only NaN is last. */
int rb200_reduce_max(int dtype, int64_t m, int64_t n, int64_t k,
                     const void* A, int64_t offA, void* B, int64_t offB);
/* index of the first strict maximum (ties -> lowest index, NaN never wins, NaN at 0
 * sticks): PhpMath::math_imax (:114-130).  index_dtype: RB200_INT32 or RB200_INT64. */
int rb200_reduce_argmax(int dtype, int64_t m, int64_t n, int64_t k,
                        const void* A, int64_t offA, int index_dtype, void* B, int64_t offB);
/* Arg-max of one STRIPE of a reduced axis that is sharded over ranks: writes, per
 * output, a 16-byte record {double value; int64 index} (index local to the stripe) into `records` (16-byte aligned,
 * `off_records` counted in records) for the cross-rank merge of math_imax (PhpMath.php:114-130).  The stripe that
 * starts at global position 0 follows math_imax itself, its never-displaced NaN at position 0 reported as +INF (the
 * merge cannot tell the two apart); every other stripe continues a running maximum, where `$value > $max` is false
 * for a NaN wherever it sits, so NaNs are skipped and an all-NaN stripe reports (-INF, INT32_MAX).  Merging the
 * records in rank order with a strict `>` reproduces the single-buffer result bit for bit. */
int rb200_reduce_argmax_partial(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA,
                                int stripe_starts_at_global_0, void* records, int64_t off_records);

/* ---- softmax: replaces PhpMath::softmax (PhpMath.php:1645-1670); in place, row stride ldA */
int rb200_softmax(int dtype, int64_t m, int64_t n, void* A, int64_t offA, int64_t ldA);

/* ---- im2col2d / col2im: replaces PhpMath::im2col2d (PhpMath.php:2158-2314), same 20
 *      arguments in the same order (buffers become pointer + dtype) ---------------------- */
int rb200_im2col2d(int dtype, int reverse,
                   void* images, int64_t images_offset, int64_t images_size,
                   int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                   int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                   int padding, int channels_first, int64_t dilation_h, int64_t dilation_w,
                   int cols_channels_first,
                   void* cols, int64_t cols_offset, int64_t cols_size);

/* ---- slice / stick and repeat: PhpMath::slice (PhpMath.php:2691-2776) moves the sub-block
 *      [start0,+size0) x [start1,+size1) x [start2,+size2) x size of A[m,n,k,size] to / from the dense Y (reverse:
 *      Y -> A, add_mode: += instead of =), `size` items with increments incA / incY; PhpMath::repeat (:1400-1422)
 *      writes B[m,repeats,k] = A[m,k].  Any dtype (copies are bit-exact).  These carry LinearAlgebra::slice / stick /
 *      stack / concat / split / repeat (LinearAlgebra.php:3944-4365). */
int rb200_slice(int dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int64_t size,
                void* A, int64_t offA, int64_t incA, void* Y, int64_t offY, int64_t incY,
                int64_t start0, int64_t size0, int64_t start1, int64_t size1, int64_t start2, int64_t size2);
int rb200_repeat(int dtype, int64_t m, int64_t k, int64_t repeats, const void* A, int64_t offA, void* B, int64_t offB);

/* ---- masking: PhpMath::masking (PhpMath.php:3229-3270).  X is a bool buffer [m,k]; wherever X[i,h] is false the
 *      elements A[i,j,h,l] (A is [m,n,k,len], any real dtype or bool) get fill written (mode 0) or added (mode 1;
 *      bool A: A || fill).  fill is converted to A's dtype as PHP stores it (truncation for integers). */
int rb200_masking(int dtype, int64_t m, int64_t n, int64_t k, int64_t len, double fill, int mode,
                  const void* X, int64_t offX, void* A, int64_t offA);

/* ---- bandpart: PhpMath::bandpart -- zero A[b,i,j] (A is [m,n,k]) outside the band: (lower >= 0 && i-j > lower) ||
 *      (upper >= 0 && j-i > upper); negative bounds keep the whole triangle.  Any dtype (bool included). */
int rb200_bandpart(int dtype, int64_t m, int64_t n, int64_t k, void* A, int64_t offA, int64_t lower, int64_t upper);

/* ---- gatherb: PhpMath::gatherb (PhpMath.php:1209-1260) behind LinearAlgebra::gatherb / scatterb / scatterbAdd
 *      (LinearAlgebra.php:2690-2920).  A[batches,m,numClass,k,len], X[batches,n,k] integer, B[batches,m,n,k,len].
 *      forward: B[b,i,j,h,:] (=|+=) A[b,i,X[b,j,h],h,:]; reverse: A[b,i,X[b,j,h],h,:] (=|+=) B[b,i,j,h,:] in ascending
 *      j (last j wins a copy; sums are added in j order -- deterministic, the reference's association).  Indices
 *      outside [0, numClass) skip the element.  Copies: any dtype; add modes: float32 / float64 / int32 / int64. */
int rb200_gatherb(int dtype, int index_dtype, int reverse, int add_mode, int64_t batches, int64_t m, int64_t n, int64_t k,
                  int64_t len, int64_t numClass, void* A, int64_t offA, const void* X, int64_t offX, void* B, int64_t offB);

/* ---- gathernd: PhpMath::gathernd (PhpMath.php:1271-1325) behind LinearAlgebra::gatherND / scatterND / scatterNDAdd
 *      (LinearAlgebra.php:2925-3090).  A[m, paramShape..., k], X[m,n,index_depth] integer, B[m,n,k]; param_shape is a
 *      HOST array of index_depth entries (<= 8).  forward: B[i,j,:] (=|+=) A[i, X[i,j,:], :]; reverse the other way in
 *      ascending j (deterministic, the reference's order).  A tuple with an index outside its dimension skips the row.
 *      X is read from the pointer given: the reference ignores X's offset (PhpMath.php:1292) and so does the caller. */
int rb200_gathernd(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int index_depth,
                   const int64_t* param_shape, void* A, int64_t offA, const void* X, void* B, int64_t offB);

/* ---- imagecopy: PhpMath::imagecopy (PhpMath.php:2877-2957; LinearAlgebra::imagecopy, LinearAlgebra.php:4590-4660).
 *      B[y,x,c] = A[clamp(y*dy+by), clamp(x*dx+bx), (rgb_flip && c<3) ? 2-c : c] with the reference's shift / flip
 *      biases; [h,w,c] or channels_first [c,h,w]; any dtype, bit-exact.  A and B must not overlap. */
int rb200_imagecopy(int dtype, int64_t height, int64_t width, int64_t channels, const void* A, int64_t offA, void* B,
                    int64_t offB, int channels_first, int64_t height_shift, int64_t width_shift, int vertical_flip,
                    int horizontal_flip, int rgb_flip);

/* ---- einsum: PhpMath::einsum (PhpMath.php:3321-3357) and einsum4p1 (:3360-3440, the same sum with four output
 *      axes and one summed axis).  sizeOfIndices / ldA / ldB are HOST arrays of `depth` entries (<= 16): the first
 *      ndimC axes index C (row-major, dense), the rest are summed, last axis fastest; an operand's stride is 0 on an
 *      axis it lacks.  C[ic] = sum A[offA + idx.ldA] * B[offB + idx.ldB], accumulated in double in the reference's
 *      order with separate multiply and add (bit-identical to PHP's `$sumC += $A[..] * $B[..]`).  float32 / float64.
 *      einsum4p1 callers pass offC = 0: the reference stores from element 0 of C whatever offsetC says (:3432). */
int rb200_einsum(int dtype, int depth, int ndimC, const int64_t* sizeOfIndices, const void* A, int64_t offA,
                 const int64_t* ldA, const void* B, int64_t offB, const int64_t* ldB, void* C, int64_t offC);

/* ---- topk: PhpMath::topk (PhpMath.php:933-1058; LinearAlgebra::topK, LinearAlgebra.php:2368-2440).  Per row of
 *      input[m,n]: the k largest values[m,k] and their indices[m,k], produced by the reference's min-heap so ties, NaN
 *      handling and output order (heap order, or heap-sort order when sorted) are identical.  k > n returns without
 *      writing, as the reference does.  float32 / float64 / int32 / int64 values; 32- or 64-bit integer indices. */
int rb200_topk(int dtype, int index_dtype, int64_t m, int64_t n, const void* input, int64_t offInput, int64_t k, int sorted,
               void* values, int64_t offValues, void* indices, int64_t offIndices);

/* ---- cumsumb: PhpMath::cumsumb (PhpMath.php:1861-1904), what LinearAlgebra::cumsum (LinearAlgebra.php:4764-4812)
 *      calls: B[i,j,h] = sum of A[i,0..j,h] along the middle axis of [m,n,k]; exclusive shifts by one (B[i,0,h] = 0),
 *      reverse runs from j = n-1 down.  float32 / float64 / int32 / int64.  Floating sums are added in tree order
 *      (the reference adds serially). */
int rb200_cumsumb(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, int exclusive, int reverse,
                  void* B, int64_t offB);

/* ---- cumsum: PhpMath::cumsum (PhpMath.php:1822-1859), the 1-D form with increments.  As in the reference, reverse
 *      walks X forwards and writes Y backwards (Y[n-1] = X[0], Y[n-2] = X[0]+X[1] ...). */
int rb200_cumsum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int exclusive, int reverse,
                 void* Y, int64_t offY, int64_t incY);

/* ---- searchsorted: PhpMath::searchsorted (PhpMath.php:1780-1820).  Y[i] = number of leading elements of row i of
 *      A[m,n] (ldA == 0: one shared row) that X[i] is above (right: above or equal), the walk stopping at the first
 *      element it is not -- identical to the reference for unsorted rows and NaN as well (sorted rows are bisected).
 *      dtype float32 / float64; index_dtype int32 / uint32 / int64 / uint64. */
int rb200_searchsorted(int dtype, int index_dtype, int64_t m, int64_t n, const void* A, int64_t offA, int64_t ldA,
                       const void* X, int64_t offX, int64_t incX, int right, void* Y, int64_t offY, int64_t incY);

/* ---- im2col3d / col2im3d: PhpMath::im2col3d (PhpMath.php:2396-2600, copyCell3d :2319-2390), the Conv3D sibling.
 *      images [b,d,h,w,c] (channels_first: [b,c,d,h,w]); cols [b,od,oh,ow,kd,kh,kw,c] (cols_channels_first:
 *      [b,od,oh,ow,c,kd,kh,kw]).  Same out / padding rules per axis as im2col2d, including the reference's
 *      padding_d, which is computed from im_h (:2459).  Forward is a bit-exact copy for every dtype; reverse
 *      accumulates in the reference's loop order (float32/float64/int32/int64).
 *      im2col1d (PhpMath.php:1967-2088) needs no entry of its own: it is rb200_im2col2d with im_h = filter_h = 1. */
int rb200_im2col3d(int dtype, int reverse,
                   void* images, int64_t images_offset, int64_t images_size,
                   int64_t batches, int64_t im_d, int64_t im_h, int64_t im_w, int64_t channels,
                   int64_t filter_d, int64_t filter_h, int64_t filter_w,
                   int64_t stride_d, int64_t stride_h, int64_t stride_w,
                   int padding, int channels_first,
                   int64_t dilation_d, int64_t dilation_h, int64_t dilation_w,
                   int cols_channels_first,
                   void* cols, int64_t cols_offset, int64_t cols_size);

/* ---- fused conv2d forward (precedent for a fused entry in the reference:
 *      LinearAlgebraCL::im2col2dclblast, src/LinearAlgebraCL.php:5001).  Computes exactly what
 *      LinearAlgebra::im2col (src/LinearAlgebra.php:3369) + gemm (:925) (+ add, :1968) compute for the
 *      reference's default layouts -- channels-last images [b,h,w,c], cols [b,oh,ow,kh,kw,c] viewed as
 *      [b*oh*ow, kh*kw*c], kernel [kh*kw*c, filters] -- without materialising cols:
 *        out[b,oy,ox,f] = sum_{ky,kx,c} images[b, oy*sh+ky*dh-ph, ox*sw+kx*dw-pw, c] * kernel[(ky,kx,c), f] (+ bias[f])
 *      float32, TF32 / 3xTF32 modes, filters <= 128, kh*kw*c <= 128; anything else returns
 *      RB200_E_UNSUPPORTED and the caller runs rb200_im2col2d + rb200_sgemm instead. ------------------ */
int rb200_conv2d_forward(int dtype, const void* images, int64_t images_offset,
                         int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                         int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                         int padding, int64_t dilation_h, int64_t dilation_w,
                         const void* kernel, int64_t kernel_offset, int64_t filters,
                         const void* bias, int64_t bias_offset, void* out, int64_t out_offset, int mode);

#ifdef __cplusplus
}
#endif
#endif /* RINDOW_B200_H_ */
