"""CudaMath -- the `math()` object of the B200 service.

Same positional signatures, checks and exception strings as PhpMath
(src/Drivers/MatlibPHP/PhpMath.php); arithmetic in librindow_b200.so.  Hot-path methods:
add :570, multiply :530, reduceSum :1552, reduceMax :1584, reduceArgMax :1616, softmax :1645,
im2col2d :2158, plus zeros :908 / fill :2811 (every alloc+zeros on the path).
"""
from __future__ import annotations

import numpy as np

from . import _capi, dtypes
from .cuda_buffer import CudaBuffer
from .exceptions import InvalidArgumentException, LogicException


def _cuda(name: str, buf) -> CudaBuffer:
    if not isinstance(buf, CudaBuffer):
        raise InvalidArgumentException("buffer %s is not a CudaBuffer" % name)
    return buf


class CudaMath:
    def __init__(self):
        self._lib = _capi.load()

    def getNumThreads(self) -> int:
        return 1

    def getNumProcs(self) -> int:
        return 1

    def getParallel(self) -> int:
        return 0

    # ---- A(m,n) := alpha * X(n) + A(m,n)   (PhpMath.php:570-606) ----------------------------
    def add(self, trans, m, n, alpha, X, offsetX, incX, A, offsetA, ldA) -> None:
        trans = bool(trans)
        rows, cols = (m, n) if not trans else (n, m)
        if offsetX + (cols - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if offsetA + (m - 1) * ldA + (n - 1) >= len(A):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X), _cuda("A", A)
        if X.dtype() != A.dtype():
            raise InvalidArgumentException("Unmatch data type for X and A")
        # The reference walks rows x cols with A strides (ldA,1) or (1,ldA) (:595-605).  Either way A
        # is the m x n row-major matrix with row stride ldA; trans selects X[i] (per row) instead of
        # X[j] (per column) -- which is exactly the ABI's definition.
        _capi.check(self._lib.rb200_add(A.dtype(), 1 if trans else 0, m, n, float(alpha),
                                        X.device_ptr(), offsetX, incX, A.device_ptr_rw(), offsetA, ldA))

    # ---- A(m,n) := X(n) * A(m,n)   (PhpMath.php:530-565) ----------------------------------------
    def multiply(self, trans, m, n, X, offsetX, incX, A, offsetA, ldA) -> None:
        trans = bool(trans)
        rows, cols = (m, n) if not trans else (n, m)
        if offsetX + (cols - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if offsetA + (m - 1) * ldA + (n - 1) >= len(A):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X), _cuda("A", A)
        if X.dtype() != A.dtype():
            raise InvalidArgumentException("Unmatch data type for X and A")
        _capi.check(self._lib.rb200_multiply(A.dtype(), 1 if trans else 0, m, n,
                                             X.device_ptr(), offsetX, incX, A.device_ptr_rw(), offsetA, ldA))

    # ---- reductions over the middle axis of A[m,n,k] (PhpMath.php:1552-1643) ----------------------
    def _reduce_check(self, m, n, k, A, offsetA, B, offsetB) -> None:
        if offsetA + m * n * k > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + m * k > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        _cuda("A", A), _cuda("B", B)

    def reduceSum(self, m, n, k, A, offsetA, B, offsetB) -> None:
        self._reduce_check(m, n, k, A, offsetA, B, offsetB)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        _capi.check(self._lib.rb200_reduce_sum(A.dtype(), m, n, k, A.device_ptr(), offsetA,
                                               B.device_ptr_rw(), offsetB))

    def reduceMax(self, m, n, k, A, offsetA, B, offsetB) -> None:
        self._reduce_check(m, n, k, A, offsetA, B, offsetB)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        _capi.check(self._lib.rb200_reduce_max(A.dtype(), m, n, k, A.device_ptr(), offsetA,
                                               B.device_ptr_rw(), offsetB))

    def reduceArgMax(self, m, n, k, A, offsetA, B, offsetB) -> None:
        self._reduce_check(m, n, k, A, offsetA, B, offsetB)
        _capi.check(self._lib.rb200_reduce_argmax(A.dtype(), m, n, k, A.device_ptr(), offsetA,
                                                  B.dtype(), B.device_ptr_rw(), offsetB))

    # ---- softmax (PhpMath.php:1645-1670) ----------------------------------------------------------------
    def softmax(self, m, n, A, offsetA, ldA) -> None:
        if offsetA + (m - 1) * ldA + (n - 1) >= len(A):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("A", A)
        _capi.check(self._lib.rb200_softmax(A.dtype(), m, n, A.device_ptr_rw(), offsetA, ldA))

    # ---- im2col2d (PhpMath.php:2158-2314) ---------------------------------------------------------------
    def im2col2d(self, reverse, images, images_offset, images_size, batches, im_h, im_w, channels,
                 filter_h, filter_w, stride_h, stride_w, padding, channels_first, dilation_h, dilation_w,
                 cols_channels_first, cols, cols_offset, cols_size) -> None:
        images_buf_size = batches * im_h * im_w * channels
        if images_size != images_buf_size or len(images) - images_offset < images_buf_size:
            raise InvalidArgumentException("images buffer size is invalid")
        out_h = int((im_h - (filter_h - 1) * dilation_h - 1) / stride_h) + 1      # intdiv truncates
        out_w = int((im_w - (filter_w - 1) * dilation_w - 1) / stride_w) + 1
        if out_h <= 0 or out_w <= 0:
            raise InvalidArgumentException("Invalid shape or parameters.")
        if padding:
            out_buf_size = batches * im_h * filter_h * im_w * filter_w * channels
        else:
            out_buf_size = batches * out_h * filter_h * out_w * filter_w * channels
        # PhpMath.php:2243-2246 rejects a cols view larger than needed (`>`); a smaller one would
        # fault inside SplFixedArray -- both are errors here.
        if cols_size != out_buf_size or len(cols) - cols_offset != out_buf_size:
            raise InvalidArgumentException("output buffer size is invalid")
        _cuda("images", images), _cuda("cols", cols)
        if images.dtype() != cols.dtype():
            raise InvalidArgumentException("Unmatch data type for images and cols")
        if reverse:
            im_ptr, co_ptr = images.device_ptr_rw(), cols.device_ptr()
        else:
            im_ptr, co_ptr = images.device_ptr(), cols.device_ptr_rw()
        _capi.check(self._lib.rb200_im2col2d(images.dtype(), 1 if reverse else 0, im_ptr, images_offset,
                                             images_size, batches, im_h, im_w, channels, filter_h, filter_w,
                                             stride_h, stride_w, 1 if padding else 0, 1 if channels_first else 0,
                                             dilation_h, dilation_w, 1 if cols_channels_first else 0,
                                             co_ptr, cols_offset, cols_size))

    # ---- fused conv2d forward (extension) ----------------------------------------
    def conv2d_forward(self, images, images_offset, batches, im_h, im_w, channels, filter_h, filter_w,
                       stride_h, stride_w, padding, dilation_h, dilation_w, kernel, kernel_offset, filters,
                       bias, bias_offset, out, out_offset, gemm_mode=_capi.GEMM_AUTO) -> bool:
        """out[b,oy,ox,f] = im2col2d(images) x kernel (+ bias) without materialising cols.  Returns False
        (nothing launched) when the shape needs the two-call path (rb200 status UNSUPPORTED)."""
        if len(images) - images_offset < batches * im_h * im_w * channels:
            raise InvalidArgumentException("images buffer size is invalid")
        if len(kernel) - kernel_offset < filter_h * filter_w * channels * filters:
            raise InvalidArgumentException("kernel buffer size is invalid")
        _cuda("images", images), _cuda("kernel", kernel), _cuda("out", out)
        if images.dtype() != dtypes.float32:
            return False
        b_ptr = None
        if bias is not None:
            if len(bias) - bias_offset < filters:
                raise InvalidArgumentException("bias buffer size is invalid")
            b_ptr = _cuda("bias", bias).device_ptr()
        rc = self._lib.rb200_conv2d_forward(images.dtype(), images.device_ptr(), images_offset, batches, im_h, im_w,
                                            channels, filter_h, filter_w, stride_h, stride_w, 1 if padding else 0,
                                            dilation_h, dilation_w, kernel.device_ptr(), kernel_offset, filters,
                                            b_ptr, bias_offset, out.device_ptr_rw(), out_offset, gemm_mode)
        if rc == _capi.E_UNSUPPORTED:
            return False
        _capi.check(rc)
        return True

    # ---- zeros / fill (PhpMath.php:908-921, 2811-2827): unit-stride only on the device -------------------
    def zeros(self, n, X, offsetX, incX) -> None:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if incX != 1:
            raise LogicException("CudaMath::zeros supports incX == 1")
        _cuda("X", X).fill(0, n, offsetX)

    def fill(self, n, V, offsetV, X, offsetX, incX) -> None:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if incX != 1:
            raise LogicException("CudaMath::fill supports incX == 1")
        value = V[offsetV]
        _cuda("X", X).fill(value, n, offsetX)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)

        def _not_on_device(*a, **k):
            raise LogicException("CudaMath::%s is not on the B200 hot path (SURVEY.md section 8f); the PHP shim "
                                 "delegates it to PhpMath, this package has no CPU fallback" % name)
        return _not_on_device


class CudaMathFactory:
    """`mathFactory` slot of AbstractMatlibService: isAvailable(), Math()."""

    def isAvailable(self) -> bool:
        return _capi.device_available()

    def Math(self) -> CudaMath:
        return CudaMath()
