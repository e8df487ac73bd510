"""CudaMath -- the `math()` object of the B200 service.

Same positional signatures, checks and exception strings as PhpMath
(src/Drivers/MatlibPHP/PhpMath.php); arithmetic in librindow_b200.so.  Hot-path methods:
add :570, multiply :530, reduceSum :1552, reduceMax :1584, reduceArgMax :1616, softmax :1645,
im2col2d :2158, plus zeros :908 / fill :2811 (every alloc+zeros on the path), and the section 8(f) widening:
the broadcast family (maximum ... lessEqual :274-528, pow :687, duplicate :860), the unary family (:221-268,
:611-682, :727-855, not :1528, nan2num :2830, isnan :2853), equal/notEqual :1470-1522, astype :1710,
sum/imax/imin :150-216 and updateAddOnehot :1438.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi, dtypes
from .cuda_buffer import CudaBuffer
from .exceptions import InvalidArgumentException, LogicException, RuntimeException


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

    # ---- the rest of the broadcast family ------------------------------------
    # maximum :274, minimum :318, greater :362, greaterEqual :404, less :446, lessEqual :488 share one signature,
    # one set of checks and messages; A[m,n] against X[n] in place.
    def _broadcast_mn(self, op, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        if m <= 0 or n <= 0:
            raise InvalidArgumentException("m and n must be greater than 0")
        if (m - 1) * ldA + (n - 1) + offsetA >= len(A):
            raise InvalidArgumentException("Matrix A is too small")
        if (n - 1) * incX + offsetX >= len(X):
            raise InvalidArgumentException("Buffer X is too small")
        _cuda("X", X), _cuda("A", A)
        if X.dtype() != A.dtype():
            raise InvalidArgumentException("Unmatch data type for A and X")
        _capi.check(self._lib.rb200_broadcast(A.dtype(), op, 0, m, n, A.device_ptr_rw(), offsetA, ldA,
                                              X.device_ptr(), offsetX, incX))

    def maximum(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_MAXIMUM, m, n, A, offsetA, ldA, X, offsetX, incX)

    def minimum(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_MINIMUM, m, n, A, offsetA, ldA, X, offsetX, incX)

    def greater(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_GREATER, m, n, A, offsetA, ldA, X, offsetX, incX)

    def greaterEqual(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_GREATER_EQUAL, m, n, A, offsetA, ldA, X, offsetX, incX)

    def less(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_LESS, m, n, A, offsetA, ldA, X, offsetX, incX)

    def lessEqual(self, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_mn(_capi.BC_LESS_EQUAL, m, n, A, offsetA, ldA, X, offsetX, incX)

    # pow :687-722 and duplicate :860-891 take $trans like add
    def _broadcast_trans(self, op, trans, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        trans = bool(trans)
        rows, cols = (m, n) if not trans else (n, m)
        if offsetX + (cols - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if offsetA + (m - 1) * ldA + (n - 1) >= len(A):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X), _cuda("A", A)
        if X.dtype() != A.dtype():
            raise InvalidArgumentException("Unmatch data type for A and X")
        _capi.check(self._lib.rb200_broadcast(A.dtype(), op, 1 if trans else 0, m, n, A.device_ptr_rw(), offsetA, ldA,
                                              X.device_ptr(), offsetX, incX))

    def pow(self, trans, m, n, A, offsetA, ldA, X, offsetX, incX) -> None:
        self._broadcast_trans(_capi.BC_POW, trans, m, n, A, offsetA, ldA, X, offsetX, incX)

    def duplicate(self, trans, m, n, X, offsetX, incX, A, offsetA, ldA) -> None:
        self._broadcast_trans(_capi.BC_DUPLICATE, trans, m, n, A, offsetA, ldA, X, offsetX, incX)

    # ---- unary family, in place (PhpMath.php:221-268, 611-682, 727-855, 1528-1547, 2830-2875) ----------------
    def _unary(self, op, n, X, offsetX, incX, alpha=1.0, beta=0.0) -> None:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X)
        _capi.check(self._lib.rb200_unary(X.dtype(), op, n, float(alpha), X.device_ptr_rw(), offsetX, incX, float(beta)))

    def increment(self, n, alpha, X, offsetX, incX, beta) -> None:
        self._unary(_capi.UN_INCREMENT, n, X, offsetX, incX, alpha, beta)

    def reciprocal(self, n, alpha, X, offsetX, incX, beta) -> None:
        self._unary(_capi.UN_RECIPROCAL, n, X, offsetX, incX, alpha, beta)

    def square(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_SQUARE, n, X, offsetX, incX)

    def sqrt(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_SQRT, n, X, offsetX, incX)

    def rsqrt(self, n, alpha, X, offsetX, incX, beta) -> None:
        self._unary(_capi.UN_RSQRT, n, X, offsetX, incX, alpha, beta)

    def exp(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_EXP, n, X, offsetX, incX)

    def log(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_LOG, n, X, offsetX, incX)

    def tanh(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_TANH, n, X, offsetX, incX)

    def sin(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_SIN, n, X, offsetX, incX)

    def cos(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_COS, n, X, offsetX, incX)

    def tan(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_TAN, n, X, offsetX, incX)

    def nan2num(self, n, X, offsetX, incX, alpha) -> None:
        self._unary(_capi.UN_NAN2NUM, n, X, offsetX, incX, alpha)

    def isnan(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_ISNAN, n, X, offsetX, incX)

    # `not` is a Python keyword: the mirror of PhpMath::not (:1528) is not_, also reachable as getattr(m, "not")
    def not_(self, n, X, offsetX, incX) -> None:
        self._unary(_capi.UN_NOT, n, X, offsetX, incX)

    # ---- equal / notEqual (PhpMath.php:1470-1522): Y := (Y == X) / (Y != X) as 1 / 0 ------------------------
    def _compare(self, op, n, X, offsetX, incX, Y, offsetY, incY) -> None:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        if offsetY + (n - 1) * incY >= len(Y):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X), _cuda("Y", Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        _capi.check(self._lib.rb200_compare(Y.dtype(), op, n, X.device_ptr(), offsetX, incX,
                                            Y.device_ptr_rw(), offsetY, incY))

    def equal(self, n, X, offsetX, incX, Y, offsetY, incY) -> None:
        self._compare(_capi.CMP_EQUAL, n, X, offsetX, incX, Y, offsetY, incY)

    def notEqual(self, n, X, offsetX, incX, Y, offsetY, incY) -> None:
        self._compare(_capi.CMP_NOT_EQUAL, n, X, offsetX, incX, Y, offsetY, incY)

    # ---- astype (PhpMath.php:1710-1778) ----------------------------------------------------------------------
    def astype(self, n, dtype, X, offsetX, incX, Y, offsetY, incY) -> None:
        if dtype not in dtypes.NUMPY:
            raise InvalidArgumentException("dtype must be type of integer or float: %s" % dtype)
        if offsetX + (n - 1) * incX >= len(X) or offsetY + (n - 1) * incY >= len(Y):
            raise InvalidArgumentException("Vector specification too large for buffer.")
        _cuda("X", X), _cuda("Y", Y)
        if Y.dtype() != dtype:
            raise InvalidArgumentException("Unmatch data type for Y")
        _capi.check(self._lib.rb200_astype(n, X.dtype(), X.device_ptr(), offsetX, incX,
                                           dtype, Y.device_ptr_rw(), offsetY, incY))

    # ---- scalar results (PhpMath.php:150-216): the value is on the host when the call returns ---------------
    def sum(self, n, X, offsetX, incX) -> float:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector X specification too large for buffer.")
        out = ctypes.c_double(0.0)
        _capi.check(self._lib.rb200_sum(_cuda("X", X).dtype(), n, X.device_ptr(), offsetX, incX, ctypes.byref(out)))
        return out.value

    def imax(self, n, X, offsetX, incX) -> int:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector X specification too large for buffer.")
        out = ctypes.c_int64(0)
        _capi.check(self._lib.rb200_imax(_cuda("X", X).dtype(), n, X.device_ptr(), offsetX, incX, ctypes.byref(out)))
        return out.value

    def imin(self, n, X, offsetX, incX) -> int:
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector X specification too large for buffer.")
        out = ctypes.c_int64(0)
        _capi.check(self._lib.rb200_imin(_cuda("X", X).dtype(), n, X.device_ptr(), offsetX, incX, ctypes.byref(out)))
        return out.value

    # ---- updateAddOnehot (PhpMath.php:1438-1464) -------------------------------------------------------------
    def updateAddOnehot(self, m, n, a, X, offsetX, incX, Y, offsetY, ldY) -> None:
        if offsetX + (m - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for bufferX.")
        if offsetY + (m - 1) * ldY + (n - 1) >= len(Y):
            raise InvalidArgumentException("Vector specification too large for bufferY.")
        _cuda("X", X), _cuda("Y", Y)
        rc = self._lib.rb200_onehot(X.dtype(), Y.dtype(), m, n, float(a), X.device_ptr(), offsetX, incX,
                                    Y.device_ptr_rw(), offsetY, ldY)
        if rc == _capi.E_RANGE:
            raise RuntimeException("Label number is out of bounds.")
        _capi.check(rc)

    # ---- gather / scatter / scatterAdd (PhpMath.php:1063-1130, 1362-1395) and reduceGather (:1135-1208) ----------
    def gather(self, reverse, addMode, n, k, numClass, X, offsetX, A, offsetA, B, offsetB) -> None:
        if offsetX + n > len(X):
            raise InvalidArgumentException("Matrix X specification too large for buffer.")
        if offsetA + numClass * k > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + n * k > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        if numClass <= 0:
            raise InvalidArgumentException("numClass must be grator than zero.")
        _cuda("X", X), _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        a_ptr = A.device_ptr_rw() if reverse else A.device_ptr()
        b_ptr = B.device_ptr() if reverse else B.device_ptr_rw()
        _capi.check(self._lib.rb200_gather(A.dtype(), X.dtype(), 1 if reverse else 0, 1 if addMode else 0, n, k, numClass,
                                           X.device_ptr(), offsetX, a_ptr, offsetA, b_ptr, offsetB))

    # ---- gatherb (PhpMath.php:1209-1260): batched gather / scatter / scatterAdd ---------------------------------------
    def gatherb(self, reverse, addMode, batches, m, n, k, len_, numClass, A, offsetA, X, offsetX, B, offsetB) -> None:
        if min(batches, m, n, k, len_, numClass) < 1:
            raise InvalidArgumentException("Argument shape must be greater than 0.")
        if offsetX + batches * n * k > len(X):
            raise InvalidArgumentException("Matrix X specification too large for buffer.")
        if offsetA + batches * m * numClass * k * len_ > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + batches * m * n * k * len_ > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        _cuda("X", X), _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        if X.dtype() not in dtypes.INT_TYPES:
            raise InvalidArgumentException("dtype of X must be integer.")
        a_ptr = A.device_ptr_rw() if reverse else A.device_ptr()
        b_ptr = B.device_ptr() if reverse else B.device_ptr_rw()
        _capi.check(self._lib.rb200_gatherb(A.dtype(), X.dtype(), 1 if reverse else 0, 1 if addMode else 0, batches, m, n, k,
                                            len_, numClass, a_ptr, offsetA, X.device_ptr(), offsetX, b_ptr, offsetB))

    # ---- gathernd (PhpMath.php:1271-1325): gatherND / scatterND / scatterNDAdd ---------------------------------------
    def gathernd(self, reverse, addMode, m, n, k, indexDepth, paramShape, A, offsetA, X, offsetX, B, offsetB) -> None:
        """paramShape: the indexDepth leading dimensions of the indexed axes (a host sequence or a buffer).  As in the
        reference, X is read from element 0 whatever offsetX says (PhpMath.php:1292)."""
        dims = [int(paramShape[h]) for h in range(indexDepth)]
        if min([m, n, k, indexDepth] + dims) < 1:
            raise InvalidArgumentException("Argument shape must be greater than 0.")
        paramSize = int(np.prod(dims, dtype=np.int64))
        if m * n * indexDepth > len(X):
            raise InvalidArgumentException("Matrix X specification too large for buffer.")
        if offsetA + m * paramSize * k > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + m * n * k > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        _cuda("X", X), _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        if X.dtype() not in dtypes.INT_TYPES:
            raise InvalidArgumentException("dtype of X must be integer.")
        shape = (ctypes.c_int64 * indexDepth)(*dims)
        a_ptr = A.device_ptr_rw() if reverse else A.device_ptr()
        b_ptr = B.device_ptr() if reverse else B.device_ptr_rw()
        _capi.check(self._lib.rb200_gathernd(A.dtype(), X.dtype(), 1 if reverse else 0, 1 if addMode else 0, m, n, k,
                                             indexDepth, ctypes.cast(shape, ctypes.c_void_p), a_ptr, offsetA, X.device_ptr(),
                                             b_ptr, offsetB))

    def reduceGather(self, reverse, addMode, m, n, numClass, X, offsetX, A, offsetA, B, offsetB) -> None:
        if offsetX + m * n > len(X):          # the reference tests offsetX+n (:1164); the whole [m,n] index must fit
            raise InvalidArgumentException("Matrix X specification too large for buffer.")
        if offsetA + m * numClass * n > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + m * n > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        if numClass <= 0:
            raise InvalidArgumentException("numClass must be grator than zero.")
        _cuda("X", X), _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        a_ptr = A.device_ptr_rw() if reverse else A.device_ptr()
        b_ptr = B.device_ptr() if reverse else B.device_ptr_rw()
        _capi.check(self._lib.rb200_reduce_gather(A.dtype(), X.dtype(), 1 if reverse else 0, 1 if addMode else 0, m, n,
                                                  numClass, X.device_ptr(), offsetX, a_ptr, offsetA, b_ptr, offsetB))

    # ---- N-D transpose (PhpMath.php:2959-3025): sourceShape and perm are int32 Buffers, as in the reference ------
    def transpose(self, sourceShape, perm, A, offsetA, B, offsetB) -> None:
        if len(sourceShape) != len(perm):
            raise InvalidArgumentException("unmatch sourceshape and perm")
        _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("unmatch sourceshape and perm")          # sic (PhpMath.php:2979-2981)
        ndim = len(sourceShape)
        if ndim <= 0:
            raise InvalidArgumentException("Matrix must not be a scalar")
        shape = [int(sourceShape[i]) for i in range(ndim)]
        pm = [int(perm[i]) for i in range(ndim)]
        size = 1
        for s in shape:
            size *= s
        for p in pm:
            if p >= ndim:
                raise InvalidArgumentException("dim value in the perm is out of range.")
        if len(set(pm)) != ndim:
            raise InvalidArgumentException("duplicate axis in perm option")
        if offsetA + size > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA")
        if offsetB + size > len(B):
            raise InvalidArgumentException("Matrix specification too large for bufferB")
        c_shape = (ctypes.c_int64 * ndim)(*shape)
        c_perm = (ctypes.c_int64 * ndim)(*pm)
        _capi.check(self._lib.rb200_transpose(A.dtype(), ndim, c_shape, c_perm, A.device_ptr(), offsetA,
                                              B.device_ptr_rw(), offsetB))

    # ---- strided-batched gemm: the call LinearAlgebraCL::matmul makes on its math object
    #      (src/LinearAlgebraCL.php:2001-2012); row-major, same operand rules as PhpBlas::gemm --------------------
    def gemmStridedBatched(self, order, transA, transB, m, n, k, alpha, A, offsetA, ldA, strideA,
                           B, offsetB, ldB, strideB, beta, C, offsetC, ldC, strideC, batchSize, gemm_mode=None) -> None:
        from .cuda_blas import BLAS, assert_matrix_buffer_spec, assert_shape_parameter, code_to_trans
        if order != BLAS.RowMajor:
            raise InvalidArgumentException("gemmStridedBatched is row-major")
        tA, _ = code_to_trans(transA)
        tB, _ = code_to_trans(transB)
        assert_shape_parameter("m", m)
        assert_shape_parameter("n", n)
        assert_shape_parameter("k", k)
        if batchSize < 1:
            raise InvalidArgumentException("Argument batchSize must be greater than 0.")
        rowsA, colsA = (m, k) if not tA else (k, m)
        rowsB, colsB = (k, n) if not tB else (n, k)
        last = batchSize - 1
        assert_matrix_buffer_spec("A", A, rowsA, colsA, offsetA + last * strideA, ldA)
        assert_matrix_buffer_spec("B", B, rowsB, colsB, offsetB + last * strideB, ldB)
        assert_matrix_buffer_spec("C", C, m, n, offsetC + last * strideC, ldC)
        _cuda("A", A), _cuda("B", B), _cuda("C", C)
        if not (A.dtype() == B.dtype() == C.dtype()) or A.dtype() not in (dtypes.float32, dtypes.float64):
            raise InvalidArgumentException("Unmatch data type for A, B and C")
        mode = _capi.GEMM_AUTO if gemm_mode is None else gemm_mode
        _capi.check(self._lib.rb200_gemm_strided_batched(
            A.dtype(), transA, transB, m, n, k, float(alpha), A.device_ptr(), offsetA, ldA, strideA,
            B.device_ptr(), offsetB, ldB, strideB, float(beta), C.device_ptr_rw(), offsetC, ldC, strideC, batchSize, mode))

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

    def reduceArgMaxPartial(self, m, n, k, A, offsetA, global0: bool, records, offsetRecords) -> None:
        """(value, index) records of one stripe of a row-sharded reduced axis, for sharding.merge_argmax
        (rb200_reduce_argmax_partial).  `records` is an int64 buffer holding two words per output: the bits of the
        double value, then the stripe-local index."""
        _cuda("A", A)
        _cuda("records", records)
        if m < 1 or n < 1 or k < 1:
            raise InvalidArgumentException("m, n and k must be greater than 0.")
        if offsetA < 0 or offsetA + m * n * k > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA.")
        if records.dtype() != dtypes.int64 or offsetRecords < 0 or 2 * (offsetRecords + m * k) > len(records):
            raise InvalidArgumentException("records must be an int64 buffer of two words per output")
        _capi.check(self._lib.rb200_reduce_argmax_partial(A.dtype(), m, n, k, A.device_ptr(), offsetA,
                                                          1 if global0 else 0, records.device_ptr_rw(), offsetRecords))

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

    # ---- topk (PhpMath.php:933-1058): the reference's min-heap per row ------------------------------------------------
    def topk(self, m, n, input, offsetInput, k, sorted, values, offsetValues, indices, offsetIndices) -> None:
        if m < 1 or n < 1 or k < 1:
            raise InvalidArgumentException("Argument m / n / k must be greater than 0.")
        if offsetInput + m * n > len(input):
            raise InvalidArgumentException("Matrix specification too large for bufferinput.")
        if offsetValues + m * k > len(values):
            raise InvalidArgumentException("Matrix specification too large for buffervalues.")
        if offsetIndices + m * k > len(indices):
            raise InvalidArgumentException("Matrix specification too large for bufferindices.")
        _cuda("input", input), _cuda("values", values), _cuda("indices", indices)
        if indices.dtype() not in dtypes.INT_TYPES:
            raise InvalidArgumentException("indices must be integers")
        if input.dtype() != values.dtype():
            raise InvalidArgumentException("Unmatch data type for input and values")
        _capi.check(self._lib.rb200_topk(input.dtype(), indices.dtype(), m, n, input.device_ptr(), offsetInput, k,
                                         1 if sorted else 0, values.device_ptr_rw(), offsetValues, indices.device_ptr_rw(),
                                         offsetIndices))

    topK = topk

    # ---- einsum / einsum4p1 (PhpMath.php:3321-3440) ---------------------------------------------------------------------
    def _einsum(self, sizes, A, offsetA, ldA, B, offsetB, ldB, C, offsetC, ndimC) -> None:
        depth = len(sizes)
        if depth < 1 or depth > 16 or min(sizes) < 1 or min(ldA) < 0 or min(ldB) < 0 or ndimC < 0 or ndimC > depth:
            raise InvalidArgumentException("Invalid einsum specification.")
        if offsetA + sum((d - 1) * l for d, l in zip(sizes, ldA)) >= len(A):
            raise InvalidArgumentException("Matrix specification too large for buffer A.")
        if offsetB + sum((d - 1) * l for d, l in zip(sizes, ldB)) >= len(B):
            raise InvalidArgumentException("Matrix specification too large for buffer B.")
        if offsetC + int(np.prod(sizes[:ndimC], dtype=np.int64)) > len(C):
            raise InvalidArgumentException("Matrix specification too large for buffer C.")
        _cuda("A", A), _cuda("B", B), _cuda("C", C)
        if A.dtype() != B.dtype() or A.dtype() != C.dtype():
            raise InvalidArgumentException("Unmatch data type for A, B and C")
        if A.dtype() not in (dtypes.float32, dtypes.float64):
            raise InvalidArgumentException("Unsupported data type.")
        arr = lambda v: ctypes.cast((ctypes.c_int64 * depth)(*[int(x) for x in v]), ctypes.c_void_p)
        s, la_, lb_ = arr(sizes), arr(ldA), arr(ldB)
        _capi.check(self._lib.rb200_einsum(A.dtype(), depth, ndimC, s, A.device_ptr(), offsetA, la_, B.device_ptr(), offsetB,
                                           lb_, C.device_ptr_rw(), offsetC))

    def einsum(self, sizeOfIndices, A, offsetA, ldA, B, offsetB, ldB, C, offsetC, ndimC) -> None:
        """sizeOfIndices / ldA / ldB: host sequences or buffers of len(sizeOfIndices) entries"""
        depth = len(sizeOfIndices)
        self._einsum([int(sizeOfIndices[h]) for h in range(depth)], A, offsetA, [int(ldA[h]) for h in range(depth)],
                     B, offsetB, [int(ldB[h]) for h in range(depth)], C, offsetC, ndimC)

    def einsum4p1(self, dim0, dim1, dim2, dim3, dim4, A, offsetA, ldA0, ldA1, ldA2, ldA3, ldA4,
                  B, offsetB, ldB0, ldB1, ldB2, ldB3, ldB4, C, offsetC) -> None:
        if len(C) <= offsetC + dim0 * dim1 * dim2 * dim3 - 1:
            raise InvalidArgumentException("Matrix specification too large for buffer C.")
        # the reference stores from element 0 of C whatever offsetC says (PhpMath.php:3432)
        self._einsum([dim0, dim1, dim2, dim3, dim4], A, offsetA, [ldA0, ldA1, ldA2, ldA3, ldA4],
                     B, offsetB, [ldB0, ldB1, ldB2, ldB3, ldB4], C, 0, 4)

    # ---- imagecopy (PhpMath.php:2877-2957): shift / flip / RGB swap of one image, border clamped ---------------------
    def imagecopy(self, height, width, channels, A, offsetA, B, offsetB, channelsFirst, heightShift, widthShift,
                  verticalFlip, horizontalFlip, rgbFlip) -> None:
        if width * height * channels + offsetA > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA")
        if width * height * channels + offsetB > len(B):
            raise InvalidArgumentException("Matrix specification too large for bufferB")
        _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        if rgbFlip and channels < 3:
            raise InvalidArgumentException("rgbFlip needs at least 3 channels.")
        _capi.check(self._lib.rb200_imagecopy(A.dtype(), height, width, channels, A.device_ptr(), offsetA, B.device_ptr_rw(),
                                              offsetB, 1 if channelsFirst else 0, int(heightShift), int(widthShift),
                                              1 if verticalFlip else 0, 1 if horizontalFlip else 0, 1 if rgbFlip else 0))

    # ---- matrixcopy (PhpMath.php:2778-2809): B = alpha * A or its transpose; the kernel of BLAS omatcopy -----------
    def matrixcopy(self, trans, m, n, alpha, A, offsetA, ldA, B, offsetB, ldB) -> None:
        if m < 1 or n < 1:
            raise InvalidArgumentException("Argument m / n must be greater than 0.")
        rows, cols = (n, m) if trans else (m, n)
        if ldA < n or offsetA + (m - 1) * ldA + n > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA.")
        if ldB < cols or offsetB + (rows - 1) * ldB + cols > len(B):
            raise InvalidArgumentException("Matrix specification too large for bufferB.")
        _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        if A.dtype() not in (dtypes.float32, dtypes.float64):
            raise InvalidArgumentException("Unsupported data type.")
        # RowMajor = 101, NoTrans = 111, Trans = 112 (the CBLAS codes rb200_omatcopy takes)
        _capi.check(self._lib.rb200_omatcopy(A.dtype(), 101, 112 if trans else 111, m, n, float(alpha),
                                             A.device_ptr(), offsetA, ldA, B.device_ptr_rw(), offsetB, ldB))

    # ---- cumsum / cumsumb / searchsorted (PhpMath.php:1780-1904) ------------------------------------------------------
    def cumsumb(self, m, n, k, A, offsetA, exclusive, reverse, B, offsetB) -> None:
        if offsetA + m * n * k > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA.")
        if offsetB + m * n * k > len(B):
            raise InvalidArgumentException("Matrix specification too large for bufferB.")
        _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        _capi.check(self._lib.rb200_cumsumb(A.dtype(), m, n, k, A.device_ptr(), offsetA, 1 if exclusive else 0,
                                            1 if reverse else 0, B.device_ptr_rw(), offsetB))

    def cumsum(self, n, X, offsetX, incX, exclusive, reverse, Y, offsetY, incY) -> None:
        if n < 1 or incX < 1 or incY < 1:
            raise InvalidArgumentException("Argument n / inc must be greater than 0.")
        if offsetX + (n - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for bufferX.")
        if offsetY + (n - 1) * incY >= len(Y):
            raise InvalidArgumentException("Vector specification too large for bufferY.")
        _cuda("X", X), _cuda("Y", Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        _capi.check(self._lib.rb200_cumsum(X.dtype(), n, X.device_ptr(), offsetX, incX, 1 if exclusive else 0,
                                           1 if reverse else 0, Y.device_ptr_rw(), offsetY, incY))

    def searchsorted(self, m, n, A, offsetA, ldA, X, offsetX, incX, right, Y, offsetY, incY) -> None:
        if m < 1 or incX < 1 or incY < 1:
            raise InvalidArgumentException("Argument m / inc must be greater than 0.")
        if offsetA + (m - 1) * ldA + n > len(A):
            raise InvalidArgumentException("Matrix specification too large for bufferA.")
        if offsetX + (m - 1) * incX >= len(X):
            raise InvalidArgumentException("Vector specification too large for bufferX.")
        if offsetY + (m - 1) * incY >= len(Y):
            raise InvalidArgumentException("Vector specification too large for bufferY.")
        _cuda("A", A), _cuda("X", X), _cuda("Y", Y)
        if A.dtype() != X.dtype():
            raise InvalidArgumentException("Unmatch data type for A and X")
        _capi.check(self._lib.rb200_searchsorted(A.dtype(), Y.dtype(), m, n, A.device_ptr(), offsetA, ldA, X.device_ptr(),
                                                 offsetX, incX, 1 if right else 0, Y.device_ptr_rw(), offsetY, incY))

    # ---- bandpart (PhpMath::bandpart): zero outside the band, in place ------------------------------------------------
    def bandpart(self, m, n, k, A, offset, lower, upper) -> None:
        if offset + m * n * k > len(A):
            raise InvalidArgumentException("Matrix specification too large for buffer.")
        _cuda("A", A)
        _capi.check(self._lib.rb200_bandpart(A.dtype(), m, n, k, A.device_ptr_rw(), offset, int(lower), int(upper)))

    # ---- masking (PhpMath.php:3229-3270) ---------------------------------------------------------------------------
    def masking(self, m, n, k, len_, fill, mode, X, offsetX, A, offsetA) -> None:
        if X.dtype() != dtypes.bool_:
            raise InvalidArgumentException("dtype of X must be bool.")
        if offsetX + m * k > len(X):
            raise InvalidArgumentException("Matrix specification too large for buffer X.")
        if offsetA + m * n * k * len_ > len(A):
            raise InvalidArgumentException("Matrix specification too large for buffer A.")
        _cuda("X", X), _cuda("A", A)
        _capi.check(self._lib.rb200_masking(A.dtype(), m, n, k, len_, float(fill), int(mode), X.device_ptr(), offsetX,
                                            A.device_ptr_rw(), offsetA))

    # ---- slice / repeat (PhpMath.php:2691-2776, :1400-1422) -------------------------------------------------------
    def slice(self, reverse, addMode, m, n, k, size, A, offsetA, incA, Y, offsetY, incY,
              startAxis0, sizeAxis0, startAxis1, sizeAxis1, startAxis2, sizeAxis2) -> None:
        if m * n * k * size * incA + offsetA > len(A):
            raise InvalidArgumentException("unmatch BufferA size and m,n,k")
        if startAxis0 < 0 or startAxis0 >= m or sizeAxis0 < 0 or sizeAxis0 + startAxis0 > m:
            raise InvalidArgumentException("Axis0 range is too large for source array.")
        if startAxis1 < 0 or startAxis1 >= n or sizeAxis1 < 0 or sizeAxis1 + startAxis1 > n:
            raise InvalidArgumentException("Axis1 range is too large for source array.")
        if startAxis2 < 0 or startAxis2 >= k or sizeAxis2 < 0 or sizeAxis2 + startAxis2 > k:
            raise InvalidArgumentException("Axis2 range is too large for source array.")
        if sizeAxis0 * sizeAxis1 * sizeAxis2 * size * incY > len(Y) - offsetY:
            raise InvalidArgumentException("BufferY size is too small")
        _cuda("A", A), _cuda("Y", Y)
        if A.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for A and Y")
        a_ptr = A.device_ptr_rw() if reverse else A.device_ptr()             # reverse writes A, forward writes Y
        y_ptr = Y.device_ptr() if reverse else Y.device_ptr_rw()
        _capi.check(self._lib.rb200_slice(A.dtype(), 1 if reverse else 0, 1 if addMode else 0, m, n, k, size,
                                          a_ptr, offsetA, incA, y_ptr, offsetY, incY,
                                          startAxis0, sizeAxis0, startAxis1, sizeAxis1, startAxis2, sizeAxis2))

    def repeat(self, m, k, repeats, A, offsetA, B, offsetB) -> None:
        if offsetA + m * k > len(A):
            raise InvalidArgumentException("Matrix A specification too large for buffer.")
        if offsetB + m * repeats * k > len(B):
            raise InvalidArgumentException("Matrix B specification too large for buffer.")
        _cuda("A", A), _cuda("B", B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        _capi.check(self._lib.rb200_repeat(A.dtype(), m, k, repeats, A.device_ptr(), offsetA, B.device_ptr_rw(), offsetB))

    # ---- im2col1d (PhpMath.php:1967-2088): the same copy as im2col2d with one row of height 1 ----------------
    def im2col1d(self, reverse, images, images_offset, images_size, batches, im_w, channels, filter_w, stride_w,
                 padding, channels_first, dilation_w, cols_channels_first, cols, cols_offset, cols_size) -> None:
        images_buf_size = batches * im_w * channels
        if images_size != images_buf_size or len(images) - images_offset < images_buf_size:
            raise InvalidArgumentException("images buffer size is invalid")
        out_w = int((im_w - (filter_w - 1) * dilation_w - 1) / stride_w) + 1
        if out_w <= 0:
            raise InvalidArgumentException("Invalid shape or parameters.")
        out_buf_size = batches * (im_w if padding else out_w) * filter_w * channels
        if cols_size != out_buf_size or len(cols) - cols_offset != out_buf_size:
            raise InvalidArgumentException("output buffer size is invalid")
        self.im2col2d(reverse, images, images_offset, images_size, batches, 1, im_w, channels, 1, filter_w, 1, stride_w,
                      padding, channels_first, 1, dilation_w, cols_channels_first, cols, cols_offset, cols_size)

    # ---- im2col3d (PhpMath.php:2396-2600) ---------------------------------------------------------------------
    def im2col3d(self, reverse, images, images_offset, images_size, batches, im_d, im_h, im_w, channels,
                 filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, padding, channels_first,
                 dilation_d, dilation_h, dilation_w, cols_channels_first, cols, cols_offset, cols_size) -> None:
        images_buf_size = batches * im_d * im_h * im_w * channels
        if images_size != images_buf_size or len(images) - images_offset < images_buf_size:
            raise InvalidArgumentException("images buffer size is invalid")
        out_d = int((im_d - (filter_d - 1) * dilation_d - 1) / stride_d) + 1
        out_h = int((im_h - (filter_h - 1) * dilation_h - 1) / stride_h) + 1
        out_w = int((im_w - (filter_w - 1) * dilation_w - 1) / stride_w) + 1
        if out_h <= 0 or out_w <= 0:                                    # the reference does not test out_d (:2453)
            raise InvalidArgumentException("Invalid shape or parameters.")
        if padding:
            out_buf_size = batches * im_d * filter_d * im_h * filter_h * im_w * filter_w * channels
        else:
            out_buf_size = batches * out_d * filter_d * out_h * filter_h * out_w * filter_w * channels
        if cols_size != out_buf_size or len(cols) - cols_offset != out_buf_size:
            raise InvalidArgumentException("output buffer size is invalid")
        _cuda("images", images), _cuda("cols", cols)
        if images.dtype() != cols.dtype():
            raise InvalidArgumentException("Unmatch data type for images and cols")
        if reverse:
            im_ptr, co_ptr = images.device_ptr_rw(), cols.device_ptr()
        else:
            im_ptr, co_ptr = images.device_ptr(), cols.device_ptr_rw()
        _capi.check(self._lib.rb200_im2col3d(images.dtype(), 1 if reverse else 0, im_ptr, images_offset, images_size,
                                             batches, im_d, im_h, im_w, channels, filter_d, filter_h, filter_w,
                                             stride_d, stride_h, stride_w, 1 if padding else 0,
                                             1 if channels_first else 0, dilation_d, dilation_h, dilation_w,
                                             1 if cols_channels_first else 0, co_ptr, cols_offset, cols_size))

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
        if name == "not":
            return self.not_

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
