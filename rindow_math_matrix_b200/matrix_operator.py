"""MatrixOperator -- the numpy-ish facade, hot-path part.

Mirror of Rindow\\Math\\Matrix\\MatrixOperator::cross (src/MatrixOperator.php:429-440) ->
matrixMultiply (:442-508): fold leading dims into M, allocate a ZERO-filled C and call the
driver's gemm directly with alpha = 1, beta = 1 (:479-503); rank(B) > 2 goes through gemm3
(:510-529), a host loop in the reference -- here it is expressed as L strided driver gemm calls
(one per leading index of B) so it stays on the device.
"""
from __future__ import annotations

import numpy as np

from . import dtypes
from .cuda_blas import BLAS
from .exceptions import InvalidArgumentException, LogicException
from .linear_algebra import LinearAlgebra
from .ndarray import NDArray
from .service import Service


class MatrixOperator:
    def __init__(self, service=None):
        if service is None:
            from .service import MatlibCuda
            service = MatlibCuda()
        self._service = service
        self._la = None
        self.defaultFloatType = dtypes.float32
        self.defaultIntType = dtypes.int32                                # MatrixOperator.php:53

    def service(self):
        return self._service

    def la(self) -> LinearAlgebra:                                       # MatrixOperator.php:1524
        if self._la is None:
            self._la = LinearAlgebra(self._service, self.defaultFloatType)
        return self._la

    def createLinearAlgebraCuda(self, options=None):                      # cf. createLinearAlgebraCL, :1553-1561
        from .linear_algebra_cuda import LinearAlgebraCuda
        queue = self._service.createQueue(options)
        return LinearAlgebraCuda(queue, self._service, self.defaultFloatType)

    def laAccelerated(self, name: str, options=None):                     # MatrixOperator.php:1566-1579
        if name == "cuda":
            if getattr(self, "_cudaLA", None) is None:
                self._cudaLA = self.createLinearAlgebraCuda(options)
            return self._cudaLA
        raise LogicException("Unknown accelerator")

    def isAdvanced(self) -> bool:
        return self._service.serviceLevel() >= Service.LV_ADVANCED

    def isAccelerated(self) -> bool:
        return self._service.serviceLevel() >= Service.LV_ACCELERATED

    def isIntType(self, dtype: int) -> bool:
        return dtype in dtypes.INT_TYPES

    def getBlas(self, dtype: int):                                       # MatrixOperator.php:225-231
        if self.isIntType(dtype):
            return self._basic_or_device("blas")
        return self._service.blas()

    def _basic_or_device(self, kind: str):
        """Integer arrays go to the LV_BASIC (pure-PHP) driver in the reference because OpenBLAS-class drivers have no
        integer kernels.  This package ships no host driver (service.py): where the service has no LV_BASIC slot the
        device driver is used -- its level-1 / scalar entry points take every real dtype."""
        getter = getattr(self._service, kind)
        try:
            return getter(Service.LV_BASIC)
        except LogicException:
            return getter()

    def array(self, array, dtype: int | None = None) -> NDArray:
        if dtype is None:
            a = np.asarray(array)
            dtype = dtypes.bool_ if a.dtype == np.bool_ else self.defaultFloatType
        return NDArray(array, dtype=dtype, service=self._service)

    def zeros(self, shape, dtype: int | None = None) -> NDArray:
        x = NDArray(None, dtype=self.defaultFloatType if dtype is None else dtype, shape=shape,
                    service=self._service)
        return self.la().zeros(x)

    def ones(self, shape, dtype: int | None = None) -> NDArray:
        x = NDArray(None, dtype=self.defaultFloatType if dtype is None else dtype, shape=shape,
                    service=self._service)
        return self.la().ones(x)

    def arange(self, count: int, start=None, step=None, dtype: int | None = None) -> NDArray:
        start = 0 if start is None else start
        step = 1 if step is None else step
        if dtype is None:                                                 # MatrixOperator.php:405-410
            dtype = self.defaultIntType if isinstance(start, (int, np.integer)) else self.defaultFloatType
        return NDArray(np.arange(count) * step + start, dtype=dtype, service=self._service)

    def isFloatType(self, dtype: int) -> bool:                            # :395-398
        return dtype in dtypes.FLOAT_TYPES

    def full(self, shape, value, dtype: int | None = None) -> NDArray:    # :342-366: alloc + la->fill
        if dtype is None:
            dtype = self.defaultFloatType
        if not isinstance(value, (bool, int, float, np.integer, np.floating, np.bool_)):
            raise InvalidArgumentException("value must be scalar")
        x = NDArray(None, dtype=dtype, shape=shape, service=self._service)
        return self.la().fill(value, x)

    def zerosLike(self, array: NDArray) -> NDArray:                       # :368-371
        return self.zeros(array.shape(), array.dtype())

    def fullLike(self, array: NDArray, value: int) -> NDArray:            # :373-376
        return self.full(array.shape(), value, array.dtype())

    def astype(self, array: NDArray, dtype: int) -> NDArray:              # :378-381
        return self.la().astype(array, dtype)

    def copy(self, array: NDArray) -> NDArray:                            # :422-427
        new = NDArray(None, dtype=array.dtype(), shape=array.shape(), service=self._service)
        return self.la().copy(array, new)

    def transpose(self, X: NDArray, perm=None) -> NDArray:                # :587-593
        return self.la().transpose(X, perm=perm)

    def dot(self, A: NDArray, B: NDArray):                                # :622-636
        if A.shape() != B.shape():
            raise InvalidArgumentException("unmatch shape of dimension")
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("unmatch value type")
        return self.getBlas(A.dtype()).dot(A.size(), A.buffer(), A.offset(), 1, B.buffer(), B.offset(), 1)

    def add(self, X: NDArray, Y: NDArray) -> NDArray:                     # :638-647: copy(Y), then axpy
        if X.shape() != Y.shape():
            raise InvalidArgumentException("Unmatch shape of dimension to add: (%s),(%s)" % (
                ",".join(map(str, X.shape())), ",".join(map(str, Y.shape()))))
        return self.la().axpy(X, self.copy(Y))

    def scale(self, a, X: NDArray) -> NDArray:                            # :649-661: copy(X), then blas->scal
        if isinstance(a, bool) or not isinstance(a, (int, float, np.integer, np.floating)):
            raise InvalidArgumentException("the scalar must be a numeric")
        C = self.copy(X)
        self.getBlas(X.dtype()).scal(X.size(), a, C.buffer(), C.offset(), 1)
        return C

    def cross(self, A: NDArray, B: NDArray) -> NDArray:                   # MatrixOperator.php:429-440
        if A.ndim() >= 2 and B.ndim() >= 2:
            return self.matrixMultiply(A, B)
        if A.ndim() >= 2 and B.ndim() == 1:
            return self.vectorTransform(A, B)
        raise InvalidArgumentException("unmatch shape of dimension")

    def vectorTransform(self, A: NDArray, B: NDArray) -> NDArray:         # MatrixOperator.php:531-582
        shapeA = A.shape()
        batchSize = 1
        while len(shapeA) > 2:
            batchSize *= shapeA.pop(0)
        shapeC = A.shape()[:-1]
        C = self.zeros(shapeC, A.dtype())
        sizeA = A.size() // batchSize
        sizeC = C.size() // batchSize
        aM, aN = shapeA
        if aN != B.shape()[0]:
            raise InvalidArgumentException("unmatch shape of dimension")
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("unmatch value type: %s and %s" % (A.dtype(), B.dtype()))
        offA, offC = A.offset(), C.offset()
        blas = self.getBlas(A.dtype())
        for _ in range(batchSize):
            blas.gemv(BLAS.RowMajor, BLAS.NoTrans, aM, aN, 1.0, A.buffer(), offA, aN,
                      B.buffer(), B.offset(), 1, 1.0, C.buffer(), offC, 1)
            offA += sizeA
            offC += sizeC
        return C

    def matrixMultiply(self, A: NDArray, B: NDArray) -> NDArray:          # MatrixOperator.php:442-508
        multiA = multiB = 1
        shapeC = []
        shapeA = A.shape()
        while len(shapeA) > 2:
            n = shapeA.pop(0)
            shapeC.append(n)
            multiA *= n
        aM, aK = shapeA
        shapeC.insert(0, aM)
        shapeB = B.shape()
        while len(shapeB) > 2:
            n = shapeB.pop(0)
            shapeC.append(n)
            multiB *= n
        bK, bN = shapeB
        shapeC.append(bN)
        if aK != bK:
            raise InvalidArgumentException("Unmatch shape of dimension to cross multiple: (%s),(%s)" % (
                ",".join(map(str, A.shape())), ",".join(map(str, B.shape()))))
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("unmatch value type")
        C = self.zeros(shapeC, A.dtype())
        alpha, beta = 1.0, 1.0
        M, N, K, L = aM * multiA, bN, bK, multiB
        blas = self.getBlas(A.dtype())
        if B.ndim() == 2:
            blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, N, K, alpha,
                      A.buffer(), A.offset(), K, B.buffer(), B.offset(), N, beta, C.buffer(), C.offset(), N)
        else:
            # gemm3 (:510-529): C[m, l, n] = sum_k A[m,k] * B[l,k,n]  ->  for each l one gemm writing the
            # [M, N] slab that starts at offC + N*l with row stride L*N.
            for l in range(L):
                blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, N, K, alpha,
                          A.buffer(), A.offset(), K, B.buffer(), B.offset() + N * K * l, N, beta,
                          C.buffer(), C.offset() + N * l, L * N)
        return C

    # ---- reductions along an axis (MatrixOperator.php:824-1023) ----------------------------------------------------
    # The reference walks the remaining axes on the host and asks the driver for ONE scalar per output element
    # (walkAxis :824-857: `$YY[$posY] = $math->sum($N,$XX,$offX+$bufPos,$incX)`), i.e. prod(shape)/shape[axis] driver
    # calls.  Viewed as A[m, n, k] with n = shape[axis] that is exactly the layout of the driver's reduce entry points
    # (PhpMath.php:1552-1643), so for float arrays sum / mean / max / argMax / min / argMin run as one to three kernels:
    #   sum     reduceSum                         mean    reduceSum + scal(1/n)
    #   argMax  reduceArgMax (first strict maximum, the math_imax rule, PhpMath.php:114-130)
    #   max     reduceArgMax, then reduceGather of X at those positions (`$XX[... + $pos*$incX]`, :911-916) -- not
    #           reduceMax, whose NaN rule is the sticky one of the accelerated drivers
    #   argMin / min   the same on -X (math_imin is math_imax mirrored: first strict minimum)
    # Everything else (integer and bool arrays, the |x| family asum / amax / amin / argAmax / argAmin, axis=None) keeps
    # the reference's walk, one driver call per output element.

    def setDefaultIntType(self, dtype: int) -> None:                      # :252-255
        self.defaultIntType = dtype

    def setDefaultFloatType(self, dtype: int) -> None:                    # :257-260
        self.defaultFloatType = dtype

    def getMath(self, dtype: int):                                        # :232-238
        if self.isIntType(dtype):
            return self._basic_or_device("math")
        return self._service.math()

    @staticmethod
    def _axis_layout(X: NDArray, axis: int):
        shape = X.shape()
        if axis < 0 or axis >= len(shape):                                # :838-840
            raise InvalidArgumentException("Invalid axis: axis=%d,shape=[%s]" % (axis, ",".join(map(str, shape))))
        m = int(np.prod(shape[:axis], dtype=np.int64)) if axis > 0 else 1
        k = int(np.prod(shape[axis + 1:], dtype=np.int64)) if axis + 1 < len(shape) else 1
        return m, shape[axis], k, shape[:axis] + shape[axis + 1:]

    def walkAxis(self, func, X: NDArray, axis=None, dtype=None):          # :824-857
        XX, offX = X.buffer(), X.offset()
        if axis is None:
            return func(X.size(), XX, offX, 0, 1)
        m, n, k, outShape = self._axis_layout(X, axis)
        Y = NDArray(None, dtype=X.dtype() if dtype is None else dtype, shape=outShape, service=self._service)
        YY = Y.buffer()
        pos = 0
        for outer in range(m):                                            # createMatrixBufferIterator order (:727-822)
            for inner in range(k):
                YY[pos] = func(n, XX, offX, outer * n * k + inner, k)
                pos += 1
        return Y

    def _on_device_axis(self, X: NDArray, axis) -> bool:
        return axis is not None and X.dtype() in dtypes.FLOAT_TYPES

    def _arg_extreme(self, X: NDArray, axis: int, smallest: bool, dtype=None) -> NDArray:
        la = self.la()
        self._axis_layout(X, axis)                                        # the reference's axis check and message
        src = la.scal(-1.0, la.copy(X)) if smallest else X
        return la.reduceArgMax(src, axis=axis, dtype=dtype)

    def sum(self, X: NDArray, axis=None):                                 # :873-892
        if self._on_device_axis(X, axis):
            self._axis_layout(X, axis)
            return self.la().reduceSum(X, axis=axis)
        if X.dtype() == dtypes.bool_:
            def func(N, XX, offX, bufPos, incX):                          # the reference's own loop bounds (:877-883)
                return sum(1 for i in range(0, N, incX) if XX[offX + bufPos + i])
        else:
            math = self.getMath(X.dtype())

            def func(N, XX, offX, bufPos, incX):
                return math.sum(N, XX, offX + bufPos, incX)
        return self.walkAxis(func, X, axis)

    def mean(self, X: NDArray, axis=None):                                # :1014-1022
        if self._on_device_axis(X, axis):
            _, n, _, _ = self._axis_layout(X, axis)
            la = self.la()
            return la.scal(1.0 / n, la.reduceSum(X, axis=axis))
        math = self.getMath(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: math.sum(N, XX, offX + bufPos, incX) / N, X, axis)

    def asum(self, X: NDArray, axis=None):                                # :894-901
        blas = self.getBlas(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: blas.asum(N, XX, offX + bufPos, incX), X, axis)

    def max(self, X: NDArray, axis=None):                                 # :903-912
        if self._on_device_axis(X, axis):
            return self.la().gather(X, self._arg_extreme(X, axis, False), axis=axis)
        math = self.getMath(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX:
                             XX[offX + bufPos + math.imax(N, XX, offX + bufPos, incX) * incX], X, axis)

    def argMax(self, X: NDArray, axis=None):                              # :914-923
        if self._on_device_axis(X, axis):
            return self._arg_extreme(X, axis, False, self.defaultIntType)
        math = self.getMath(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: math.imax(N, XX, offX + bufPos, incX), X, axis,
                             self.defaultIntType)

    def min(self, X: NDArray, axis=None):                                 # :950-959
        if self._on_device_axis(X, axis):
            return self.la().gather(X, self._arg_extreme(X, axis, True), axis=axis)
        math = self.getMath(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX:
                             XX[offX + bufPos + math.imin(N, XX, offX + bufPos, incX) * incX], X, axis)

    def argMin(self, X: NDArray, axis=None):                              # :961-970
        if self._on_device_axis(X, axis):
            return self._arg_extreme(X, axis, True, self.defaultIntType)
        math = self.getMath(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: math.imin(N, XX, offX + bufPos, incX), X, axis,
                             self.defaultIntType)

    def amax(self, X: NDArray, axis=None):                                # :925-934
        blas = self.getBlas(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX:
                             XX[offX + bufPos + blas.iamax(N, XX, offX + bufPos, incX) * incX], X, axis)

    def argAmax(self, X: NDArray, axis=None):                             # :936-945
        blas = self.getBlas(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: blas.iamax(N, XX, offX + bufPos, incX), X, axis,
                             self.defaultIntType)

    def amin(self, X: NDArray, axis=None):                                # :972-990 (hasIamin() is true on this driver)
        blas = self.getBlas(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX:
                             XX[offX + bufPos + blas.iamin(N, XX, offX + bufPos, incX) * incX], X, axis)

    def argAmin(self, X: NDArray, axis=None):                             # :992-1012
        blas = self.getBlas(X.dtype())
        return self.walkAxis(lambda N, XX, offX, bufPos, incX: blas.iamin(N, XX, offX + bufPos, incX), X, axis,
                             self.defaultIntType)
