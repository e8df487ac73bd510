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
            return self._service.blas(Service.LV_BASIC)
        return self._service.blas()

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
        dtype = dtypes.int32 if dtype is None else dtype
        return NDArray(np.arange(count) * step + start, dtype=dtype, service=self._service)

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
