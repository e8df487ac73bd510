<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use FFI;
use RuntimeException;

/**
 * Thin PHP-FFI binding of librindow_b200.so (C ABI: include/rindow_b200.h).
 *
 * NOTE: this file cannot be executed in the build environment of this repository (no PHP
 * interpreter); it is the reference-side binding a maintainer adds, kept mechanical on purpose:
 * every method is validation (the reference's own Utils trait) + one FFI call.  The same ABI is
 * exercised by the ctypes binding rindow_math_matrix_b200/_capi.py, function for function.
 */
class CudaFFI
{
    private static ?FFI $ffi = null;

    public const CDEF = <<<'CDEF'
        int         rb200_abi_version(void);
        const char* rb200_last_error(void);
        int         rb200_device_count(int* count);
        int         rb200_init(int device);
        int         rb200_shutdown(void);
        int         rb200_sync(void);
        int         rb200_set_default_gemm_mode(int mode);
        int         rb200_buffer_alloc(int64_t bytes, void** dptr);
        int         rb200_buffer_free(void* dptr);
        int         rb200_buffer_h2d(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes);
        int         rb200_buffer_d2h(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes);
        int         rb200_buffer_h2d_async(void* dptr, int64_t dst_offset_bytes, const void* host, int64_t bytes);
        int         rb200_buffer_d2d(void* dst, int64_t dst_off, const void* src, int64_t src_off, int64_t bytes);
        int         rb200_buffer_fill(int dtype, void* dptr, int64_t off, int64_t n, const void* value);
        int         rb200_host_alloc(int64_t bytes, void** hptr);
        int         rb200_host_free(void* hptr);
        int rb200_sgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                        float alpha, const float* A, int64_t offA, int64_t ldA,
                        const float* B, int64_t offB, int64_t ldB,
                        float beta, float* C, int64_t offC, int64_t ldC, int mode);
        int rb200_dgemm(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                        double alpha, const double* A, int64_t offA, int64_t ldA,
                        const double* B, int64_t offB, int64_t ldB,
                        double beta, double* C, int64_t offC, int64_t ldC);
        int rb200_sgemm_streamed(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                        float alpha, float* dA, const float* hA, int64_t offA, int64_t ldA,
                        float* dB, const float* hB, int64_t offB, int64_t ldB,
                        float beta, float* dC, float* hC, int hC_is_current, int64_t offC, int64_t ldC, int mode);
        int rb200_scal(int dtype, int64_t n, double alpha, void* X, int64_t offX, int64_t incX);
        int rb200_axpy(int dtype, int64_t n, double alpha, const void* X, int64_t offX, int64_t incX,
                       void* Y, int64_t offY, int64_t incY);
        int rb200_copy(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX,
                       void* Y, int64_t offY, int64_t incY);
        int rb200_swap(int dtype, int64_t n, void* X, int64_t offX, int64_t incX,
                       void* Y, int64_t offY, int64_t incY);
        int rb200_asum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
        int rb200_nrm2(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
        int rb200_dot(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX,
                      const void* Y, int64_t offY, int64_t incY, double* result);
        int rb200_iamax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
        int rb200_iamin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
        int rb200_gemv(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
                       const void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX,
                       double beta, void* Y, int64_t offY, int64_t incY);
        int rb200_omatcopy(int dtype, int order, int trans, int64_t m, int64_t n, double alpha,
                           const void* A, int64_t offA, int64_t ldA, void* B, int64_t offB, int64_t ldB);
        int rb200_add(int dtype, int trans, int64_t m, int64_t n, double alpha,
                      const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA);
        int rb200_multiply(int dtype, int trans, int64_t m, int64_t n,
                           const void* X, int64_t offX, int64_t incX, void* A, int64_t offA, int64_t ldA);
        int rb200_gemm_strided_batched(int dtype, int transA, int transB, int64_t m, int64_t n, int64_t k, double alpha,
                            const void* A, int64_t offA, int64_t ldA, int64_t strideA,
                            const void* B, int64_t offB, int64_t ldB, int64_t strideB, double beta,
                            void* C, int64_t offC, int64_t ldC, int64_t strideC, int64_t batch, int mode);
        int rb200_ipc_export(void* dptr, void* handle64);
        int rb200_ipc_open(const void* handle64, void** dptr);
        int rb200_ipc_close(void* dptr);
        int rb200_sgemm_allgather(int order, int transA, int transB, int64_t m, int64_t n, int64_t k,
                            float alpha, const float* A, int64_t offA, int64_t ldA,
                            const float* B, int64_t offB, int64_t ldB,
                            float beta, float* C, int64_t offC, int64_t ldC, int mode,
                            int n_peers, void* const* peer_C);
        int rb200_transpose(int dtype, int ndim, const int64_t* shape, const int64_t* perm,
                            const void* A, int64_t offA, void* B, int64_t offB);
        int rb200_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t n, int64_t k, int64_t numClass,
                         const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB);
        int rb200_reduce_gather(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t numClass,
                         const void* X, int64_t offX, void* A, int64_t offA, void* B, int64_t offB);
        int rb200_broadcast(int dtype, int op, int trans, int64_t m, int64_t n,
                            void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX);
        int rb200_unary(int dtype, int op, int64_t n, double alpha, void* X, int64_t offX, int64_t incX, double beta);
        int rb200_compare(int dtype, int op, int64_t n, const void* X, int64_t offX, int64_t incX,
                          void* Y, int64_t offY, int64_t incY);
        int rb200_astype(int64_t n, int from_dtype, const void* X, int64_t offX, int64_t incX,
                         int to_dtype, void* Y, int64_t offY, int64_t incY);
        int rb200_sum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, double* result);
        int rb200_imax(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
        int rb200_imin(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int64_t* result);
        int rb200_onehot(int dtype_x, int dtype_y, int64_t m, int64_t n, double a,
                         const void* X, int64_t offX, int64_t incX, void* Y, int64_t offY, int64_t ldY);
        int rb200_reduce_sum(int dtype, int64_t m, int64_t n, int64_t k,
                             const void* A, int64_t offA, void* B, int64_t offB);
        int rb200_reduce_max(int dtype, int64_t m, int64_t n, int64_t k,
                             const void* A, int64_t offA, void* B, int64_t offB);
        int rb200_reduce_argmax(int dtype, int64_t m, int64_t n, int64_t k,
                                const void* A, int64_t offA, int index_dtype, void* B, int64_t offB);
        int rb200_softmax(int dtype, int64_t m, int64_t n, void* A, int64_t offA, int64_t ldA);
        int rb200_im2col2d(int dtype, int reverse,
                           void* images, int64_t images_offset, int64_t images_size,
                           int64_t batches, int64_t im_h, int64_t im_w, int64_t channels,
                           int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w,
                           int padding, int channels_first, int64_t dilation_h, int64_t dilation_w,
                           int cols_channels_first,
                           void* cols, int64_t cols_offset, int64_t cols_size);
        int rb200_bandpart(int dtype, int64_t m, int64_t n, int64_t k, void* A, int64_t offA, int64_t lower, int64_t upper);
        int rb200_gatherb(int dtype, int index_dtype, int reverse, int add_mode, int64_t batches, int64_t m, int64_t n, int64_t k, int64_t len, int64_t numClass, void* A, int64_t offA, const void* X, int64_t offX, void* B, int64_t offB);
        int rb200_gathernd(int dtype, int index_dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int index_depth, const int64_t* param_shape, void* A, int64_t offA, const void* X, void* B, int64_t offB);
        int rb200_imagecopy(int dtype, int64_t height, int64_t width, int64_t channels, const void* A, int64_t offA, void* B, int64_t offB, int channels_first, int64_t height_shift, int64_t width_shift, int vertical_flip, int horizontal_flip, int rgb_flip);
        int rb200_einsum(int dtype, int depth, int ndimC, const int64_t* sizeOfIndices, const void* A, int64_t offA, const int64_t* ldA, const void* B, int64_t offB, const int64_t* ldB, void* C, int64_t offC);
        int rb200_topk(int dtype, int index_dtype, int64_t m, int64_t n, const void* input, int64_t offInput, int64_t k, int sorted, void* values, int64_t offValues, void* indices, int64_t offIndices);
        int rb200_cumsumb(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, int exclusive, int reverse, void* B, int64_t offB);
        int rb200_cumsum(int dtype, int64_t n, const void* X, int64_t offX, int64_t incX, int exclusive, int reverse, void* Y, int64_t offY, int64_t incY);
        int rb200_searchsorted(int dtype, int index_dtype, int64_t m, int64_t n, const void* A, int64_t offA, int64_t ldA, const void* X, int64_t offX, int64_t incX, int right, void* Y, int64_t offY, int64_t incY);
        int rb200_masking(int dtype, int64_t m, int64_t n, int64_t k, int64_t len, double fill, int mode,
                          const void* X, int64_t offX, void* A, int64_t offA);
        int rb200_slice(int dtype, int reverse, int add_mode, int64_t m, int64_t n, int64_t k, int64_t size,
                        void* A, int64_t offA, int64_t incA, void* Y, int64_t offY, int64_t incY,
                        int64_t start0, int64_t size0, int64_t start1, int64_t size1, int64_t start2, int64_t size2);
        int rb200_repeat(int dtype, int64_t m, int64_t k, int64_t repeats, const void* A, int64_t offA, void* B, int64_t offB);
        int rb200_im2col3d(int dtype, int reverse, void* images, int64_t images_offset, int64_t images_size,
                           int64_t batches, int64_t im_d, int64_t im_h, int64_t im_w, int64_t channels,
                           int64_t filter_d, int64_t filter_h, int64_t filter_w,
                           int64_t stride_d, int64_t stride_h, int64_t stride_w, int padding, int channels_first,
                           int64_t dilation_d, int64_t dilation_h, int64_t dilation_w, int cols_channels_first,
                           void* cols, int64_t cols_offset, int64_t cols_size);
        int rb200_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, int64_t* total_mem_bytes);
        int rb200_set_stream(void* cuda_stream);
        int rb200_get_stream(void** cuda_stream);
        int rb200_get_default_gemm_mode(int* mode);
        int rb200_launch_count(int64_t* count, int reset);
        const char* rb200_last_gemm_path(void);
        int rb200_buffer_d2h_async(void* host, const void* dptr, int64_t src_offset_bytes, int64_t bytes);
        int rb200_event_create(void** event);
        int rb200_event_destroy(void* event);
        int rb200_event_record(void* event);
        int rb200_event_synchronize(void* event);
        int rb200_event_query(void* event, int* done);
        int rb200_event_wait(void* event);
        int rb200_sgemm_allgather_mc(int order, int transA, int transB, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t offA, int64_t ldA, const float* B, int64_t offB, int64_t ldB, float beta, float* C, int64_t offC, int64_t ldC, int mode, void* mc_C);
        int rb200_reduce_argmax_partial(int dtype, int64_t m, int64_t n, int64_t k, const void* A, int64_t offA, int stripe_starts_at_global_0, void* records, int64_t off_records);
        int rb200_conv2d_forward(int dtype, const void* images, int64_t images_offset, int64_t batches, int64_t im_h, int64_t im_w, int64_t channels, int64_t filter_h, int64_t filter_w, int64_t stride_h, int64_t stride_w, int padding, int64_t dilation_h, int64_t dilation_w, const void* kernel, int64_t kernel_offset, int64_t filters, const void* bias, int64_t bias_offset, void* out, int64_t out_offset, int mode);
        int rb200_graph_begin(void);
        int rb200_graph_end(void** graph_exec);
        int rb200_graph_launch(void* graph_exec);
        int rb200_graph_destroy(void* graph_exec);
        int rb200_buffer_alloc_shareable(int64_t bytes, void** dptr);
        CDEF;

    public static function lib() : FFI
    {
        if (self::$ffi === null) {
            $path = getenv('RINDOW_B200_LIB') ?: 'librindow_b200.so';
            $ffi = FFI::cdef(self::CDEF, $path);
            $status = $ffi->rb200_init((int)(getenv('RINDOW_B200_DEVICE') ?: 0));
            if ($status != 0) {             // no B200 / wrong architecture: stay unavailable on every later call too
                throw new RuntimeException(
                    'librindow_b200 status '.$status.': '.FFI::string($ffi->rb200_last_error()));
            }
            self::$ffi = $ffi;
        }
        return self::$ffi;
    }

    public static function isAvailable() : bool
    {
        if (!extension_loaded('ffi')) {
            return false;
        }
        try {
            self::lib();
            return true;
        } catch (\Throwable $e) {
            return false;
        }
    }

    public static function check(int $status) : void
    {
        if ($status != 0) {
            throw new RuntimeException(
                'librindow_b200 status '.$status.': '.FFI::string(self::$ffi->rb200_last_error()));
        }
    }
}
