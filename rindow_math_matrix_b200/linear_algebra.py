"""LinearAlgebra -- host LA API over a driver service.

Mirror of the hot-path callers of Rindow\\Math\\Matrix\\LinearAlgebra (src/LinearAlgebra.php):
shape checks -> (m,n,k,ld,offset) -> one positional driver call, with the reference's exception
strings.  gemm :925-995, add :1968-2016, multiply :1921-1963, softmax :3138-3156,
reduceSum/Max/ArgMax :3161-3336, im2col :3369, col2im :3434, im2col2d :3597-3701, the section 8(f)
widening (copy/scal/axpy :312-371, gemv :861-923, matmul :1000-1110, omatcopy :1443-1500, 2-D transpose
:4438-4565), plus the
array/alloc/zeros/ones/fill plumbing those need (:193-330, :4662).  It is driver-agnostic, exactly
like the PHP class: everything numeric goes through service.blas() / service.math().
"""
from __future__ import annotations

import numpy as np

from . import dtypes
from .cuda_blas import BLAS
from .exceptions import InvalidArgumentException, LogicException
from .ndarray import NDArray, NDArrayInterface
from .service import Service


def _intdiv(a: int, b: int) -> int:
    return int(a / b) if b else 0        # PHP intdiv truncates toward zero


class LinearAlgebra:
    def __init__(self, service, defaultFloatType: int | None = None):
        self.service = service
        self.blas = service.blas()
        self.math = service.math()
        self.defaultFloatType = dtypes.float32 if defaultFloatType is None else defaultFloatType

    # ---- plumbing ---------------------------------------------------------------------------------
    def accelerated(self) -> bool:
        return self.service.serviceLevel() >= Service.LV_ADVANCED and type(self.blas).__name__.startswith("Cuda")

    # ---- accessors (LinearAlgebra.php:61-93, 153-162, 183-189) -------------------------------------------------
    def getBlas(self):
        return self.blas

    def getMath(self):
        return self.math

    def getConfig(self) -> str:
        return self.blas.getConfig()

    def fp64(self) -> bool:
        return True

    def finish(self) -> None:
        """Host drivers have nothing to wait for (:91-93); the CUDA service is stream-ordered and host reads
        synchronise themselves, so the LV_ADVANCED tier keeps the reference's no-op (LinearAlgebraCuda overrides)."""

    def isInt(self, value: NDArray) -> bool:
        return value.dtype() in dtypes.INT_TYPES

    def isFloat(self, value: NDArray) -> bool:
        return value.dtype() in dtypes.FLOAT_TYPES

    def isComplexDtype(self, dtype: int) -> bool:                               # :4955-4958
        return dtype in dtypes.COMPLEX_TYPES

    def scalar(self, array):
        return array.toArray() if isinstance(array, NDArrayInterface) else array

    def abs(self, value) -> float:                                              # :4970-4980 (real scalars)
        if isinstance(value, bool) or not isinstance(value, (int, float, np.integer, np.floating)):
            raise InvalidArgumentException("invalid data type: " + type(value).__name__)
        return abs(value)

    def expandDims(self, x: NDArray, axis: int) -> NDArray:                     # :191-215
        shape = x.shape()
        ndim = len(shape)
        orgAxis = axis
        if axis < 0:
            axis = ndim + axis + 1
        if axis < 0 or axis > ndim:
            raise InvalidArgumentException("axis is out of range: %d" % orgAxis)
        return x.reshape(shape[:axis] + [1] + shape[axis:])

    def squeeze(self, x: NDArray, axis: int | None = None) -> NDArray:          # :217-250
        shape = x.shape()
        if axis is None:
            return x.reshape([n for n in shape if n != 1])
        ndim = len(shape)
        orgAxis = axis
        if axis < 0:
            axis = ndim + axis
        if axis < 0 or axis >= ndim:
            raise InvalidArgumentException("axis is out of range: %d" % orgAxis)
        if shape[axis] != 1:
            raise InvalidArgumentException("Can not squeeze dim[%d]" % axis)
        return x.reshape(shape[:axis] + shape[axis + 1:])

    def isclose(self, a: NDArray, b: NDArray, rtol=None, atol=None, debug=None) -> bool:   # :5439-5464
        """The reference's comparator (its tests' tolerance convention):
        max|b - a| < atol + rtol * max|b|, rtol 1e-4, atol 1e-7 -- copy / axpy / scal / iamax on the driver."""
        rtol = 1e-04 if rtol is None else rtol
        atol = 1e-07 if atol is None else atol
        if a.shape() != b.shape():
            return False
        diff = self.abs(self.scalar(self.amax(self.axpy(a, self.copy(b), -1))))
        close = atol + self.abs(self.scalar(self.amax(self.scal(rtol, self.copy(b)))))
        if debug:
            print("diff=%14.8e, close=%14.8e" % (diff, close))
        return bool(diff < close)

    def array(self, array, dtype: int | None = None) -> NDArray:
        if isinstance(array, NDArrayInterface):
            return array
        if dtype is None:
            a = np.asarray(array)
            if a.dtype == np.bool_:
                dtype = dtypes.bool_
            else:
                dtype = self.defaultFloatType
        return NDArray(array, dtype=dtype, service=self.service)

    def alloc(self, shape, dtype: int | None = None) -> NDArray:
        return NDArray(None, dtype=self.defaultFloatType if dtype is None else dtype, shape=shape,
                       service=self.service)

    def toNDArray(self, ndarray: NDArray) -> NDArray:                           # LinearAlgebra.php:175-181
        return ndarray

    # hooks the device-resident subclass (LinearAlgebraCuda) overrides
    def _host_array(self, array, dtype: int) -> NDArray:
        """A small parameter array the DRIVER reads from the host (shape / perm vectors, fill patterns)."""
        return NDArray(array, dtype=dtype, service=self.service)

    def _element(self, X: NDArray, i: int):
        """One element of X read on the host (`$XX[$offX+$i]`, LinearAlgebra.php:480-492, 1644-1645)."""
        return X.buffer()[X.offset() + i]

    def fill(self, value, X: NDArray) -> NDArray:                               # LinearAlgebra.php:4662
        V = self._host_array([value], X.dtype())
        self.math.fill(X.size(), V.buffer(), V.offset(), X.buffer(), X.offset(), 1)
        return X

    def zeros(self, X: NDArray) -> NDArray:
        self.math.zeros(X.size(), X.buffer(), X.offset(), 1)
        return X

    def ones(self, X: NDArray) -> NDArray:
        return self.fill(1.0, X)

    def zerosLike(self, array: NDArray) -> NDArray:                             # LinearAlgebra.php:298-301
        return self.zeros(self.alloc(array.shape(), array.dtype()))

    # ---- BLAS level 1 mirrors (LinearAlgebra.php:312-371) ----------------------------------------------
    @staticmethod
    def _shape_error(X: NDArray, Y: NDArray) -> str:
        return "Unmatch shape of dimension: (%s),(%s)" % (",".join(map(str, X.shape())), ",".join(map(str, Y.shape())))

    def copy(self, X: NDArray, Y: NDArray | None = None) -> NDArray:            # LinearAlgebra.php:312-331
        if Y is None:
            Y = self.alloc(X.shape(), dtype=X.dtype())
        elif X.shape() != Y.shape():
            raise InvalidArgumentException(self._shape_error(X, Y))
        self.blas.copy(X.size(), X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)
        return Y

    def scal(self, alpha, X: NDArray) -> NDArray:                               # LinearAlgebra.php:336-345
        self.blas.scal(X.size(), alpha, X.buffer(), X.offset(), 1)
        return X

    def axpy(self, X: NDArray, Y: NDArray, alpha=None) -> NDArray:              # LinearAlgebra.php:350-371
        if X.shape() != Y.shape():
            raise InvalidArgumentException(self._shape_error(X, Y))
        if alpha is None:
            alpha = 1.0
        self.blas.axpy(X.size(), alpha, X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)
        return Y

    # ---- BLAS level-1 reductions and swap (LinearAlgebra.php:374-530, :842-858) ---------------------------
    def dot(self, X: NDArray, Y: NDArray, conj=None):                            # :374-398 (real dtypes)
        if X.shape() != Y.shape():
            raise InvalidArgumentException(self._shape_error(X, Y))
        return self.blas.dot(X.size(), X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)

    def asum(self, X: NDArray) -> float:                                         # :403-410
        return self.blas.asum(X.size(), X.buffer(), X.offset(), 1)

    def iamax(self, X: NDArray) -> int:                                          # :415-422
        return self.blas.iamax(X.size(), X.buffer(), X.offset(), 1)

    def iamin(self, X: NDArray) -> int:                                          # :427-438 (hasIamin() is true here)
        return self.blas.iamin(X.size(), X.buffer(), X.offset(), 1)

    def amax(self, X: NDArray) -> float:                                         # :480-492: |X[iamax]|, read from the buffer
        i = self.blas.iamax(X.size(), X.buffer(), X.offset(), 1)
        return abs(self._element(X, i))

    def amin(self, X: NDArray) -> float:                                         # :497-513
        i = self.blas.iamin(X.size(), X.buffer(), X.offset(), 1)
        return abs(self._element(X, i))

    def nrm2(self, X: NDArray) -> float:                                         # :518-527
        return self.blas.nrm2(X.size(), X.buffer(), X.offset(), 1)

    def swap(self, X: NDArray, Y: NDArray) -> None:                              # :842-858
        if X.shape() != Y.shape():
            raise InvalidArgumentException(self._shape_error(X, Y))
        self.blas.swap(X.size(), X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)

    # ---- y := alpha * Ax + beta * y   (LinearAlgebra.php:861-923) -------------------------------------
    def gemv(self, A: NDArray, X: NDArray, alpha=None, beta=None, Y: NDArray | None = None,
             trans=None, conj=None) -> NDArray:
        trans = bool(trans)
        if A.ndim() != 2 or X.ndim() != 1:
            raise InvalidArgumentException('"A" must be 2D-NDArray and "X" must 1D-NDArray.')
        shapeA = A.shape()
        rows, cols = (shapeA[0], shapeA[1]) if not trans else (shapeA[1], shapeA[0])
        if cols != X.shape()[0]:
            raise InvalidArgumentException(
                'The number of columns in "A" and The number of item in "X" must be the same')
        m, n = shapeA
        if alpha is None:
            alpha = 1.0
        if beta is None:
            beta = 0.0
        if Y is not None:
            if Y.ndim() != 1:
                raise InvalidArgumentException('"Y" must 1D-NDArray.')
            if rows != Y.shape()[0]:
                raise InvalidArgumentException(
                    'The number of rows in "A" and The number of item in "Y" must be the same')
        else:
            Y = self.zeros(self.alloc([rows], dtype=X.dtype()))
        self.blas.gemv(BLAS.RowMajor, BLAS.Trans if trans else BLAS.NoTrans, m, n, alpha,
                       A.buffer(), A.offset(), n, X.buffer(), X.offset(), 1, beta, Y.buffer(), Y.offset(), 1)
        return Y

    # ---- batched C := alpha * AB + beta * C with batch broadcast (LinearAlgebra.php:1000-1110) ----------
    def matmul(self, A: NDArray, B: NDArray, transA=None, transB=None, C: NDArray | None = None,
               alpha=None, beta=None, conjA=None, conjB=None) -> NDArray:
        transA, transB = bool(transA), bool(transB)
        shapes = "[%s]<=>[%s]" % (",".join(map(str, A.shape())), ",".join(map(str, B.shape())))
        if A.ndim() < 2 or B.ndim() < 2:
            raise InvalidArgumentException("Dimensions rank must be greater then 2D or equal:" + shapes)
        shapeA, shapeB = A.shape(), B.shape()
        shapeEA, shapeA = shapeA[-2:], shapeA[:-2]
        shapeEB, shapeB = shapeB[-2:], shapeB[:-2]
        batchA = int(np.prod(shapeA, dtype=np.int64)) if shapeA else 1
        batchB = int(np.prod(shapeB, dtype=np.int64)) if shapeB else 1
        if transA:
            shapeEA = shapeEA[::-1]
        if transB:
            shapeEB = shapeEB[::-1]
        if shapeEA[1] != shapeEB[0]:
            raise InvalidArgumentException(
                'The number of columns in "A" and the number of rows in "B" must be the same:' + shapes)
        M, N, K = shapeEA[0], shapeEB[1], shapeEA[1]
        if alpha is None:
            alpha = 1.0
        if beta is None:
            beta = 0.0
        lda = M if transA else K
        ldb = K if transB else N
        ldc = N
        shapeEC = [M, N]
        if batchA > batchB:
            dest, base, orgShapeC = batchA, batchB, shapeA + shapeEC
        else:
            dest, base, orgShapeC = batchB, batchA, shapeB + shapeEC
        if dest % base != 0:
            raise InvalidArgumentException("Matrix size-incompatible for broadcast:" + shapes)
        if C is not None:
            if C.shape() != orgShapeC:
                raise InvalidArgumentException(
                    '"A" and "C" must have the same number of rows."B" and "C" must have the same number of '
                    'columns:[%s] , [%s] => [%s]' % (",".join(map(str, A.shape())), ",".join(map(str, B.shape())),
                                                     ",".join(map(str, C.shape()))))
        else:
            C = self.zeros(self.alloc(orgShapeC, dtype=A.dtype()))
        tA = self._trans_code(transA, conjA)
        tB = self._trans_code(transB, conjB)
        offA, offB, offC = A.offset(), B.offset(), C.offset()
        batched = getattr(self.math, "gemmStridedBatched", None) if type(self.math).__name__ == "CudaMath" else None
        for _ in range(dest // base):
            if batchA > batchB:
                offB = B.offset()
            else:
                offA = A.offset()
            if batched is not None and base > 1:
                # one driver call per broadcast repeat, as LinearAlgebraCL::matmul does (LinearAlgebraCL.php:1995-2016)
                batched(BLAS.RowMajor, tA, tB, M, N, K, alpha, A.buffer(), offA, lda, M * K, B.buffer(), offB, ldb, N * K,
                        beta, C.buffer(), offC, ldc, M * N, base,
                        gemm_mode=getattr(self.blas, "gemm_mode", None))
                offA += M * K * base
                offB += N * K * base
                offC += M * N * base
                continue
            for _ in range(base):
                self.blas.gemm(BLAS.RowMajor, tA, tB, M, N, K, alpha, A.buffer(), offA, lda,
                               B.buffer(), offB, ldb, beta, C.buffer(), offC, ldc)
                offA += M * K
                offB += N * K
                offC += M * N
        return C

    # ---- B := alpha * op(A)   (LinearAlgebra.php:1443-1500) ----------------------------------------------
    def omatcopy(self, A: NDArray, trans=None, conj=None, alpha=None, B: NDArray | None = None) -> NDArray:
        trans = bool(trans)
        if A.ndim() != 2:
            raise InvalidArgumentException("Dimensions must be 2D-NDArray")
        rows, cols = A.shape()
        if trans:
            rows, cols = cols, rows
        if B is None:
            B = self.zeros(self.alloc([rows, cols], dtype=A.dtype()))
        elif B.shape() != [rows, cols]:
            raise InvalidArgumentException(self._shape_error(A, B))
        M, N = A.shape()
        if alpha is None:
            alpha = 1.0
        self.blas.omatcopy(BLAS.RowMajor, BLAS.Trans if trans else BLAS.NoTrans, M, N, alpha,
                           A.buffer(), A.offset(), N, B.buffer(), B.offset(), cols)
        return B

    # ---- transpose (LinearAlgebra.php:4438-4565): 2-D floats through omatcopy, everything else transposeND --------
    def transpose(self, A: NDArray, perm=None, B: NDArray | None = None) -> NDArray:
        if A.ndim() < 1:
            raise InvalidArgumentException("input array must be grator than or equal 1D.")
        has_omatcopy = getattr(self.blas, "hasOmatcopy", lambda: False)()
        if A.ndim() == 2 and A.dtype() in (dtypes.float32, dtypes.float64) and has_omatcopy:
            if perm is not None and len(perm) and len(perm) != 2:
                raise InvalidArgumentException("unmatch sourceshape and perm")
            return self._transpose2D(A, B)
        return self._transposeND(A, perm, B)

    def _transpose2D(self, A: NDArray, B: NDArray | None) -> NDArray:            # LinearAlgebra.php:4520-4565
        shape = A.shape()[::-1]
        if B is None:
            B = self.alloc(shape, dtype=A.dtype())
        else:
            if B.shape() != shape:
                raise InvalidArgumentException("output shape must be transpose matrix of input.")
            if B.dtype() != A.dtype():
                raise InvalidArgumentException("output data type must be same with matrix of input.")
        m, n = shape[1], shape[0]
        self.blas.omatcopy(BLAS.RowMajor, BLAS.Trans, m, n, 1.0, A.buffer(), A.offset(), n,
                           B.buffer(), B.offset(), m)
        return B

    def _transposeND(self, A: NDArray, perm, B: NDArray | None) -> NDArray:      # LinearAlgebra.php:4471-4519
        if perm is None:
            perm = list(range(A.ndim() - 1, -1, -1))
        if isinstance(perm, NDArrayInterface):
            perm = [int(v) for v in np.asarray(perm.toArray()).reshape(-1)]
        perm = [int(v) for v in perm]
        shapeA = A.shape()
        if len(shapeA) != len(perm):
            raise InvalidArgumentException("unmatch sourceshape and perm")
        shapeB, seen = [], set()
        for axis in perm:
            if axis in seen:
                raise InvalidArgumentException("duplicate axis in perm option")
            seen.add(axis)
            if axis >= len(shapeA):
                raise InvalidArgumentException("dim value in the perm is out of range.")
            shapeB.append(shapeA[axis])
        if B is None:
            B = self.alloc(shapeB, dtype=A.dtype())
        else:
            if B.shape() != shapeB:
                raise InvalidArgumentException("output shape must be transpose matrix of input.")
            if B.dtype() != A.dtype():
                raise InvalidArgumentException("output data type must be same with matrix of input.")
        sourceShape = self._host_array(np.asarray(shapeA, np.int32), dtypes.int32).buffer()
        permBuf = self._host_array(np.asarray(perm, np.int32), dtypes.int32).buffer()
        self.math.transpose(sourceShape, permBuf, A.buffer(), A.offset(), B.buffer(), B.offset())
        return B

    # ---- C := alpha * AB + beta * C   (LinearAlgebra.php:925-995) -----------------------------------
    @staticmethod
    def _trans_code(trans, conj) -> int:                                        # complementTrans / transToCode, :125-143
        trans, conj = bool(trans), bool(conj)                                   # real dtypes: conj defaults to false
        if trans:
            return BLAS.ConjTrans if conj else BLAS.Trans
        return BLAS.ConjNoTrans if conj else BLAS.NoTrans

    def gemm(self, A: NDArray, B: NDArray, alpha=None, beta=None, C: NDArray | None = None,
             transA=None, transB=None, conjA=None, conjB=None) -> NDArray:
        transA, transB = bool(transA), bool(transB)
        if A.ndim() != 2 or B.ndim() != 2:
            raise InvalidArgumentException("Dimensions must be 2D-NDArray")
        shapeA = A.shape()
        if transA:
            shapeA = [shapeA[1], shapeA[0]]
        shapeB = B.shape()
        if transB:
            shapeB = [shapeB[1], shapeB[0]]
        if shapeA[1] != shapeB[0]:
            raise InvalidArgumentException(
                'The number of columns in "A" and the number of rows in "B" must be the same')
        M, N, K = shapeA[0], shapeB[1], shapeA[1]
        if alpha is None:
            alpha = 1.0
        if beta is None:
            beta = 0.0
        if C is not None:
            shapeC = C.shape()
            if M != shapeC[0] or N != shapeC[1]:
                raise InvalidArgumentException(
                    '"A" and "C" must have the same number of rows."B" and "C" must have the same number of columns')
        else:
            C = self.zeros(self.alloc([M, N], dtype=A.dtype()))
            beta = 0.0
        lda = M if transA else K
        ldb = K if transB else N
        ldc = N
        self.blas.gemm(BLAS.RowMajor, self._trans_code(transA, conjA), self._trans_code(transB, conjB), M, N, K, alpha,
                       A.buffer(), A.offset(), lda, B.buffer(), B.offset(), ldb, beta,
                       C.buffer(), C.offset(), ldc)
        return C

    # ---- broadcast elementwise (LinearAlgebra.php:1921-2016) ------------------------------------------
    def _broadcast_mn(self, X: NDArray, A: NDArray, trans: bool, reversed_in_message: bool):
        shapeX = X.shape()
        shapeA = A.shape()
        if trans:
            shapeA = shapeA[::-1]
        while True:
            if not shapeX:
                break
            xd = shapeX.pop()
            ad = shapeA.pop() if shapeA else None
            if xd != ad:
                shown = A.shape()[::-1] if (trans and reversed_in_message) else A.shape()
                raise InvalidArgumentException("Unmatch dimension size for broadcast.: [%s] => [%s]" % (
                    ",".join(map(str, X.shape())), ",".join(map(str, shown))))
        n = X.size()
        m = A.size() // n
        if trans:
            m, n = n, m
        return m, n

    def multiply(self, X: NDArray, A: NDArray, trans=None) -> NDArray:
        trans = bool(trans)
        m, n = self._broadcast_mn(X, A, trans, True)
        self.math.multiply(trans, m, n, X.buffer(), X.offset(), 1, A.buffer(), A.offset(), n)
        return A

    def add(self, X: NDArray, A: NDArray, alpha=None, trans=None) -> NDArray:
        trans = bool(trans)
        if alpha is None:
            alpha = 1.0
        m, n = self._broadcast_mn(X, A, trans, False)
        self.math.add(trans, m, n, alpha, X.buffer(), X.offset(), 1, A.buffer(), A.offset(), n)
        return A

    # ---- scalar reductions (LinearAlgebra.php:1602-1659) ---------------------------------------------------
    def sum(self, X: NDArray) -> float:
        return self.math.sum(X.size(), X.buffer(), X.offset(), 1)

    def imax(self, X: NDArray) -> int:
        return self.math.imax(X.size(), X.buffer(), X.offset(), 1)

    def imin(self, X: NDArray) -> int:
        return self.math.imin(X.size(), X.buffer(), X.offset(), 1)

    def max(self, X: NDArray):
        i = self.math.imax(X.size(), X.buffer(), X.offset(), 1)
        return self._element(X, i)

    def min(self, X: NDArray):
        i = self.math.imin(X.size(), X.buffer(), X.offset(), 1)
        return self._element(X, i)

    # ---- unary family, in place (LinearAlgebra.php:1664-1715, 2021-2080, 2143-2246, 2296-2304, 4817-4853) ----
    def increment(self, X: NDArray, beta=None, alpha=None) -> NDArray:
        self.math.increment(X.size(), 1.0 if alpha is None else alpha, X.buffer(), X.offset(), 1,
                            0.0 if beta is None else beta)
        return X

    def reciprocal(self, X: NDArray, beta=None, alpha=None) -> NDArray:
        self.math.reciprocal(X.size(), 1.0 if alpha is None else alpha, X.buffer(), X.offset(), 1,
                             0.0 if beta is None else beta)
        return X

    def rsqrt(self, X: NDArray, beta=None, alpha=None) -> NDArray:
        self.math.rsqrt(X.size(), 1.0 if alpha is None else alpha, X.buffer(), X.offset(), 1,
                        0.0 if beta is None else beta)
        return X

    def _unary(self, name: str, X: NDArray) -> NDArray:
        getattr(self.math, name)(X.size(), X.buffer(), X.offset(), 1)
        return X

    def square(self, X: NDArray) -> NDArray:
        return self._unary("square", X)

    def sqrt(self, X: NDArray) -> NDArray:
        return self._unary("sqrt", X)

    def exp(self, X: NDArray) -> NDArray:
        return self._unary("exp", X)

    def log(self, X: NDArray) -> NDArray:
        return self._unary("log", X)

    def tanh(self, X: NDArray) -> NDArray:
        return self._unary("tanh", X)

    def sin(self, X: NDArray) -> NDArray:
        return self._unary("sin", X)

    def cos(self, X: NDArray) -> NDArray:
        return self._unary("cos", X)

    def tan(self, X: NDArray) -> NDArray:
        return self._unary("tan", X)

    def not_(self, X: NDArray) -> NDArray:                                      # LinearAlgebra::not :2296
        return self._unary("not", X)

    def isnan(self, X: NDArray) -> NDArray:                                     # :4840
        return self._unary("isnan", X)

    def nan2num(self, X: NDArray, alpha=None) -> NDArray:                       # :4817
        self.math.nan2num(X.size(), X.buffer(), X.offset(), 1, 0.0 if alpha is None else alpha)
        return X

    # ---- A[m,n] against X[n] (LinearAlgebra.php:1720-1916) ---------------------------------------------------
    def _calc_broadcast_format(self, A: NDArray, X):
        if isinstance(X, (int, float, np.integer, np.floating)) and not isinstance(X, bool):
            X = self.array(X, dtype=A.dtype())
        if not isinstance(X, NDArrayInterface):
            raise InvalidArgumentException("X must be NDArray or float")
        if X.ndim() == 0:
            X = X.reshape([X.size()])
            m, n = A.size(), 1
        else:
            shapeA, shapeX = A.shape(), X.shape()
            if shapeA == shapeX:
                m, n = 1, A.size()
            else:
                n = 1
                while shapeX:
                    tmpX = shapeX.pop()
                    tmpA = shapeA.pop() if shapeA else None
                    if tmpA != tmpX:
                        raise InvalidArgumentException("A and X is unmatched for broadcast")
                    n *= tmpX
                m = int(np.prod(shapeA, dtype=np.int64)) if shapeA else 1
        if A.dtype() != X.dtype():
            raise InvalidArgumentException("A and X must be same data type")
        return int(m), int(n), X

    def _broadcast_op(self, name: str, A: NDArray, X) -> NDArray:
        m, n, X = self._calc_broadcast_format(A, X)
        getattr(self.math, name)(m, n, A.buffer(), A.offset(), n, X.buffer(), X.offset(), 1)
        return A

    def maximum(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("maximum", A, X)

    def minimum(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("minimum", A, X)

    def greater(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("greater", A, X)

    def greaterEqual(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("greaterEqual", A, X)

    def less(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("less", A, X)

    def lessEqual(self, A: NDArray, X) -> NDArray:
        return self._broadcast_op("lessEqual", A, X)

    # ---- A(m,n) := A(m,n) ** alpha(n)   (LinearAlgebra.php:2085-2138) ----------------------------------------
    def pow(self, A: NDArray, alpha, trans=None) -> NDArray:
        trans = bool(trans)
        shapeA = A.shape()
        if isinstance(alpha, (int, float, np.integer, np.floating)):
            alpha = self.array(alpha, dtype=A.dtype())
        shapeX = alpha.shape()
        if len(shapeX) == 0:
            trans = False
            shapeA = [int(np.prod(shapeA, dtype=np.int64)) if shapeA else 1, 1]
            shapeX = [1]
        if trans:
            shapeA = shapeA[::-1]
        while shapeX:
            xd = shapeX.pop()
            ad = shapeA.pop() if shapeA else None
            if xd != ad:
                shown = A.shape()[::-1] if trans else A.shape()
                raise InvalidArgumentException("Unmatch dimension size for broadcast.: [%s] => [%s]" % (
                    ",".join(map(str, shapeX)), ",".join(map(str, shown))))
        n = alpha.size()
        m = A.size() // n
        if trans:
            m, n = n, m
        self.math.pow(trans, m, n, A.buffer(), A.offset(), n, alpha.buffer(), alpha.offset(), 1)
        return A

    # ---- Y := (X == Y) / (X != Y) as 1 / 0   (LinearAlgebra.php:2252-2290) -----------------------------------
    def _xy_shape_check(self, X: NDArray, Y: NDArray) -> None:
        if X.shape() != Y.shape():
            raise InvalidArgumentException('Unmatch shape of dimension "X" and "Y" and "numClass": (%s),(%s)' % (
                ",".join(map(str, X.shape())), ",".join(map(str, Y.shape()))))

    def equal(self, X: NDArray, Y: NDArray) -> NDArray:
        self._xy_shape_check(X, Y)
        self.math.equal(X.size(), X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)
        return Y

    def notEqual(self, X: NDArray, Y: NDArray) -> NDArray:
        self._xy_shape_check(X, Y)
        self.math.notEqual(X.size(), X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)
        return Y

    # ---- A(m,n) := X(n)   (LinearAlgebra.php:2311-2361) ------------------------------------------------------
    def duplicate(self, input: NDArray, repeats=None, trans=None, output: NDArray | None = None) -> NDArray:
        trans = bool(trans)
        if output is None:
            if repeats is None:
                repeats = 1
            shape = ([repeats] + input.shape()) if not trans else (input.shape() + [repeats])
            output = self.alloc(shape, dtype=input.dtype())
        else:
            shapeX, shapeA = input.shape(), output.shape()
            if trans:
                shapeA = shapeA[::-1]
            while shapeX:
                xd = shapeX.pop()
                ad = shapeA.pop() if shapeA else None
                if xd != ad:
                    raise InvalidArgumentException("Unmatch dimension size for broadcast.: [%s] => [%s]" % (
                        ",".join(map(str, input.shape())), ",".join(map(str, output.shape()))))
        n = input.size()
        m = output.size() // n
        if trans:
            m, n = n, m
        self.math.duplicate(trans, m, n, input.buffer(), input.offset(), 1, output.buffer(), output.offset(), n)
        return output

    # ---- astype (LinearAlgebra.php:291-307) ------------------------------------------------------------------
    def astype(self, X: NDArray, dtype: int) -> NDArray:
        Y = self.alloc(X.shape(), dtype=dtype)
        self.math.astype(X.size(), dtype, X.buffer(), X.offset(), 1, Y.buffer(), Y.offset(), 1)
        return Y

    # ---- gather / scatterAdd / scatter (LinearAlgebra.php:2441-2690) -------------------------------------------
    @staticmethod
    def _printable_shapes(values) -> str:                                        # LinearAlgebra.php:95-115
        parts = []
        for v in values:
            shape = v.shape() if isinstance(v, NDArrayInterface) else list(v)
            parts.append("(" + ",".join(map(str, shape)) + ")")
        return ",".join(parts)

    def _do_gather(self, scatterAdd: bool, A: NDArray, X: NDArray, axis=None, output: NDArray | None = None,
                   dtype=None) -> NDArray:
        if axis is None:
            postfix, prefix = A.shape(), X.shape()
            numClass = postfix.pop(0)
            m, n = 1, int(np.prod(prefix, dtype=np.int64)) if prefix else 1
            k = int(np.prod(postfix, dtype=np.int64)) if postfix else 1
            reduction = False
            outputShape = prefix + postfix
        else:
            ndim = A.ndim()
            if axis < 0:
                axis = ndim + axis
            postfix = A.shape()
            prefix = [postfix.pop(0) for _ in range(axis)]
            m = int(np.prod(prefix, dtype=np.int64)) if prefix else 1
            numClass = postfix.pop(0)
            n = int(np.prod(postfix, dtype=np.int64)) if postfix else 1
            k = 1
            reduction = True
            outputShape = prefix + postfix
            if X.shape() != outputShape:
                raise InvalidArgumentException("Unmatch Shape:" + self._printable_shapes([A, X]))
        if dtype is None:
            dtype = A.dtype()
        if output is None:
            output = self.zeros(self.alloc(outputShape, dtype=dtype))
        elif output.shape() != outputShape:
            raise InvalidArgumentException("Unmatch output shape of dimension: " + self._printable_shapes([outputShape, output]))
        reverse = addMode = bool(scatterAdd)
        if reduction:
            self.math.reduceGather(reverse, addMode, m, n, numClass, X.buffer(), X.offset(), A.buffer(), A.offset(),
                                   output.buffer(), output.offset())
        else:
            self.math.gather(reverse, addMode, n, k, numClass, X.buffer(), X.offset(), A.buffer(), A.offset(),
                             output.buffer(), output.offset())
        return output

    def gather(self, A: NDArray, X: NDArray, axis=None, output: NDArray | None = None, dtype=None) -> NDArray:
        return self._do_gather(False, A, X, axis, output, dtype)

    def scatterAdd(self, X: NDArray, A: NDArray, output: NDArray, axis=None, dtype=None) -> NDArray:
        self._do_gather(True, output, X, axis, A, dtype)
        return output

    def scatter(self, X: NDArray, A: NDArray, numClass: int, axis=None, output: NDArray | None = None, dtype=None) -> NDArray:
        if axis is None:
            postfix, prefix = A.shape(), X.shape()
            tmp = [postfix.pop(0) for _ in range(X.ndim())]
            if tmp != prefix:
                raise InvalidArgumentException("Unmatch Shape:" + self._printable_shapes([X, A]))
            n = int(np.prod(prefix, dtype=np.int64)) if prefix else 1
            k = int(np.prod(postfix, dtype=np.int64)) if postfix else 1
            m, expand = 1, False
            outputShape = [numClass] + postfix
        else:
            ndim = A.ndim()
            if axis < 0:
                axis = ndim + axis
            postfix = A.shape()
            if postfix != X.shape():
                raise InvalidArgumentException("Unmatch Shape X and A:" + self._printable_shapes([X, A]))
            prefix = [postfix.pop(0) for _ in range(axis)]
            m = int(np.prod(prefix, dtype=np.int64)) if prefix else 1
            n = int(np.prod(postfix, dtype=np.int64)) if postfix else 1
            k, expand = 1, True
            outputShape = prefix + [numClass] + postfix
        if dtype is None:
            dtype = A.dtype()
        if output is None:
            output = self.zeros(self.alloc(outputShape, dtype=dtype))
        elif output.shape() != outputShape:
            raise InvalidArgumentException("Unmatch shape of dimension: (%s),(%s)" % (
                ",".join(map(str, A.shape())), ",".join(map(str, output.shape()))))
        if expand:
            self.math.reduceGather(True, False, m, n, numClass, X.buffer(), X.offset(), output.buffer(), output.offset(),
                                   A.buffer(), A.offset())
        else:
            self.math.gather(True, False, n, k, numClass, X.buffer(), X.offset(), output.buffer(), output.offset(),
                             A.buffer(), A.offset())
        return output

    # ---- onehot (LinearAlgebra.php:3095-3133) ----------------------------------------------------------------
    def onehot(self, X: NDArray, numClass: int, alpha=None, output: NDArray | None = None) -> NDArray:
        if X.ndim() != 1:
            raise InvalidArgumentException('"X" must be 1D-NDArray.')
        sizeX = X.size()
        if output is None:
            output = self.zeros(self.alloc([sizeX, numClass], dtype=self.defaultFloatType))
        if output.ndim() != 2:
            raise InvalidArgumentException('"Y" must be 2D-NDArray.')
        m, n = output.shape()
        if m != sizeX or n != numClass:
            raise InvalidArgumentException('Unmatch shape of dimension "X" and "Y" and "numClass": (%s),(%s)' % (
                ",".join(map(str, X.shape())), ",".join(map(str, output.shape()))))
        if alpha is None:
            alpha = 1.0
        self.math.updateAddOnehot(m, n, alpha, X.buffer(), X.offset(), 1, output.buffer(), output.offset(), n)
        return output

    # ---- softmax (LinearAlgebra.php:3138-3156) ---------------------------------------------------------
    def softmax(self, X: NDArray) -> NDArray:
        if X.ndim() != 2:
            raise InvalidArgumentException('"X" must be 2-D dimension')
        m, n = X.shape()
        self.math.softmax(m, n, X.buffer(), X.offset(), n)
        return X

    # ---- reductions (LinearAlgebra.php:3161-3336) --------------------------------------------------------
    def _reduce_args(self, input: NDArray, axis, keepdims, output, dtype, default_dtype):
        ndim = input.ndim()
        origAxis = axis
        if axis is None:
            axis = 0                                      # ** CAUTION ** null is 0
        if axis < 0:
            axis = ndim + axis
        if axis < 0 or axis > ndim - 1:
            raise InvalidArgumentException("Invalid axis: %s" % ("null" if origAxis is None else origAxis))
        shape = input.shape()
        prefix, n, postfix = shape[:axis], shape[axis], shape[axis + 1:]
        m = int(np.prod(prefix, dtype=np.int64)) if prefix else 1
        k = int(np.prod(postfix, dtype=np.int64)) if postfix else 1
        outputShape = prefix + ([1] if keepdims else []) + postfix
        if dtype is None:
            dtype = default_dtype
        if output is None:
            output = self.alloc(outputShape, dtype=dtype)
        elif output.shape() != outputShape:
            raise InvalidArgumentException("Unmatch shape of dimension: (%s),(%s)" % (
                ",".join(map(str, input.shape())), ",".join(map(str, output.shape()))))
        return m, n, k, output

    def reduceSum(self, input: NDArray, axis=None, keepdims=None, output=None, dtype=None) -> NDArray:
        m, n, k, output = self._reduce_args(input, axis, keepdims, output, dtype, input.dtype())
        self.math.reduceSum(m, n, k, input.buffer(), input.offset(), output.buffer(), output.offset())
        return output

    def reduceMax(self, input: NDArray, axis=None, keepdims=None, output=None, dtype=None) -> NDArray:
        m, n, k, output = self._reduce_args(input, axis, keepdims, output, dtype, input.dtype())
        self.math.reduceMax(m, n, k, input.buffer(), input.offset(), output.buffer(), output.offset())
        return output

    def reduceArgMax(self, input: NDArray, axis=None, keepdims=None, output=None, dtype=None) -> NDArray:
        m, n, k, output = self._reduce_args(input, axis, keepdims, output, dtype, dtypes.int32)
        self.math.reduceArgMax(m, n, k, input.buffer(), input.offset(), output.buffer(), output.offset())
        return output

    def reduceMean(self, input: NDArray, axis=None, keepdims=None, output=None, dtype=None) -> NDArray:
        if axis is None:                                                         # LinearAlgebra.php:3338-3362
            axis = 0
        output = self.reduceSum(input, axis=axis, keepdims=keepdims, output=output, dtype=dtype)
        ndim = input.ndim()
        if axis < 0:
            axis = ndim + axis
        if ndim <= axis:
            raise InvalidArgumentException("axis must be less then num of dimension")
        rows = input.shape()[axis]
        self.scal(1 / rows, output)
        return output

    # ---- im2col / col2im (LinearAlgebra.php:3369-3494, 3597-3701) -------------------------------------------
    # ---- gatherb / scatterb / scatterbAdd (LinearAlgebra.php:2690-2920) ------------------------------------------
    def _do_gatherb(self, reverse, addMode, A: NDArray, X: NDArray, axis=None, batchDims=None, detailDepth=None,
                    indexDepth=None, B=None) -> NDArray:
        """LinearAlgebra::doGatherb: params (batches, m, numClass, k, len), indices (batches, n, k),
        outputs (batches, m, n, k, len)"""
        batchDims = 0 if batchDims is None else batchDims
        if batchDims < 0:
            batchDims += A.ndim()
        if batchDims >= A.ndim():
            raise InvalidArgumentException(f"batchDims ({batchDims}) must be less than to ndims of params ({A.ndim()})")
        axis = batchDims if axis is None else axis
        if axis < 0:
            axis += A.ndim()
        if axis >= A.ndim():
            raise InvalidArgumentException(f"axis ({axis}) must be less than to ndims of params ({A.ndim()})")
        detailDepth = axis + 1 if detailDepth is None else detailDepth
        indexDepth = X.ndim() if indexDepth is None else indexDepth
        if batchDims > axis:
            if axis == 0:
                batchDims = 0
            else:
                raise InvalidArgumentException(f"batchDims ({batchDims}) must be less than or equal to axis ({axis})")
        shapeA = A.shape()
        batchShape, outerShape = shapeA[:batchDims], shapeA[batchDims:]
        outerShape, innerShape = outerShape[:axis - batchDims], outerShape[axis - batchDims:]
        numClass = innerShape.pop(0)
        cut = detailDepth - 1 - axis                           # array_splice: a negative offset counts from the end
        if cut < 0:
            cut = max(len(innerShape) + cut, 0)
        innerShape, detailShape = innerShape[:cut], innerShape[cut:]
        shapeX = X.shape()
        batchShapeX, indexShape = shapeX[:batchDims], shapeX[batchDims:]
        cutx = indexDepth - batchDims
        if cutx < 0:
            cutx = max(len(indexShape) + cutx, 0)
        indexShape, innerShapeX = indexShape[:cutx], indexShape[cutx:]
        if batchShape != batchShapeX:
            raise InvalidArgumentException(
                f"Unmatch batch shape of params and indices: {batchShape},{batchShapeX},")
        if innerShape != innerShapeX:
            raise InvalidArgumentException(
                f"Unmatch inner shape of params and indices: param's inner shape is {innerShape},"
                f"index's inner shape is {innerShapeX},")
        prod = lambda s: int(np.prod(s, dtype=np.int64)) if s else 1
        batches, m, n, k, len_ = prod(batchShape), prod(outerShape), prod(indexShape), prod(innerShape), prod(detailShape)
        outputShape = batchShape + outerShape + indexShape + innerShape + detailShape
        if B is None:
            B = self.zeros(self.alloc(outputShape, A.dtype()))
        elif B.shape() != outputShape:
            what = "updates" if reverse else "output"
            raise InvalidArgumentException(f"Unmatch {what} shape: expects {outputShape}, but gives {B.shape()}")
        self.math.gatherb(reverse, addMode, batches, m, n, k, len_, numClass, A.buffer(), A.offset(), X.buffer(), X.offset(),
                          B.buffer(), B.offset())
        return B

    def gatherb(self, params: NDArray, indices: NDArray, axis=None, batchDims=None, detailDepth=None, indexDepth=None,
                outputs=None) -> NDArray:
        return self._do_gatherb(False, False, params, indices, axis, batchDims, detailDepth, indexDepth, outputs)

    def scatterb(self, indices: NDArray, updates: NDArray, shape, axis=None, batchDims=None, detailDepth=None,
                 indexDepth=None, outputs=None) -> NDArray:
        if outputs is None:
            outputs = self.zeros(self.alloc(list(shape), updates.dtype()))
        self._do_gatherb(True, False, outputs, indices, axis, batchDims, detailDepth, indexDepth, updates)
        return outputs

    def scatterbAdd(self, indices: NDArray, updates: NDArray, shape, axis=None, batchDims=None, detailDepth=None,
                    indexDepth=None, outputs=None) -> NDArray:
        if outputs is None:
            outputs = self.zeros(self.alloc(list(shape), updates.dtype()))
        self._do_gatherb(True, True, outputs, indices, axis, batchDims, detailDepth, indexDepth, updates)
        return outputs

    # ---- gatherND / scatterND / scatterNDAdd (LinearAlgebra.php:2925-3090) -----------------------------------------
    def _do_gathernd(self, reverse, addMode, A: NDArray, X: NDArray, batchDims=None, B=None) -> NDArray:
        """LinearAlgebra::doGatherND: params (m, paramShape..., k), indices (m, n, index_depth), outputs (m, n, k)"""
        batchDims = 0 if batchDims is None else batchDims
        batchShape = X.shape()
        indexDepth = batchShape.pop()
        batchShape, outerShape = batchShape[:batchDims], batchShape[batchDims:]
        innerShape = A.shape()
        paramsBatchShape, innerShape = innerShape[:batchDims], innerShape[batchDims:]
        if paramsBatchShape != batchShape:
            raise InvalidArgumentException(f"Unmatch batch shape of params and indices: {paramsBatchShape},{batchShape}")
        if indexDepth > A.ndim() - batchDims:
            raise InvalidArgumentException(
                "ndim of params must greater than index depth or equal: "
                f"ndim of params gives {A.ndim()} , index depth gives {indexDepth}")
        paramShape, innerShape = innerShape[:indexDepth], innerShape[indexDepth:]
        outputShape = batchShape + outerShape + innerShape
        prod = lambda s: int(np.prod(s, dtype=np.int64)) if s else 1
        m, n, k = prod(batchShape), prod(outerShape), prod(innerShape)
        if B is None:
            B = self.zeros(self.alloc(outputShape, A.dtype()))
        elif B.shape() != outputShape:
            what = "updates" if reverse else "output"
            raise InvalidArgumentException(f"Unmatch {what} shape: expects {outputShape}, but gives {B.shape()}")
        self.math.gathernd(reverse, addMode, m, n, k, indexDepth, paramShape, A.buffer(), A.offset(), X.buffer(), X.offset(),
                           B.buffer(), B.offset())
        return B

    def gatherND(self, params: NDArray, indices: NDArray, batchDims=None, outputs=None) -> NDArray:
        return self._do_gathernd(False, False, params, indices, batchDims, outputs)

    def scatterND(self, indices: NDArray, updates: NDArray, shape, batchDims=None, outputs=None) -> NDArray:
        if outputs is None:
            outputs = self.zeros(self.alloc(list(shape), updates.dtype()))
        self._do_gathernd(True, False, outputs, indices, batchDims, updates)
        return outputs

    def scatterNDAdd(self, indices: NDArray, updates: NDArray, shape, batchDims=None, outputs=None) -> NDArray:
        if outputs is None:
            outputs = self.zeros(self.alloc(list(shape), updates.dtype()))
        self._do_gathernd(True, True, outputs, indices, batchDims, updates)
        return outputs

    def topK(self, input: NDArray, k=None, sorted=None, values=None, indices=None):
        """LinearAlgebra::topK (LinearAlgebra.php:2368-2440): [values, indices] of the k largest along the last axis"""
        k = 1 if k is None else k
        sorted = True if sorted is None else sorted
        shape = input.shape()
        n = shape.pop()
        m = int(np.prod(shape, dtype=np.int64)) if shape else 1
        shape.append(k)
        if k > n:
            raise InvalidArgumentException("k must be less than or equal to the entity.")
        if values is None:
            values = self.alloc(shape, input.dtype())
        else:
            if values.shape() != shape:
                raise InvalidArgumentException(f'Unmatch shape of "values" with "input" and "k": {values.shape()}')
            if values.dtype() != input.dtype():
                raise InvalidArgumentException('Unmatch dtype of "values" with input')
        if indices is None:
            indices = self.alloc(shape, dtypes.int32)
        else:
            if indices.ndim() != 2:
                raise InvalidArgumentException('"indices" must be 2-D dimension')
            if indices.shape() != shape:
                raise InvalidArgumentException(f'Unmatch shape of "indices" with "input" and "k": {indices.shape()}')
            if indices.dtype() != dtypes.int32:
                raise InvalidArgumentException('Unmatch dtype of "indices" with input')
        self.math.topk(m, n, input.buffer(), input.offset(), k, sorted, values.buffer(), values.offset(),
                       indices.buffer(), indices.offset())
        return [values, indices]

    def imagecopy(self, A: NDArray, B=None, channels_first=None, heightShift=None, widthShift=None, verticalFlip=None,
                  horizontalFlip=None, rgbFlip=None) -> NDArray:
        """LinearAlgebra::imagecopy (LinearAlgebra.php:4590-4660)"""
        if A.ndim() != 3:
            raise InvalidArgumentException("input array must be 3D.")
        shape = A.shape()
        if B is None:
            B = self.zeros(self.alloc(shape, A.dtype()))
        else:
            if B.shape() != shape:
                raise InvalidArgumentException("output shape must be transpose matrix of input.")
            if B.dtype() != A.dtype():
                raise InvalidArgumentException("output data type must be same with matrix of input.")
        if channels_first is None:                             # the reference tests == null: false selects [h,w,c] too
            channels_first = False
        if not channels_first:
            height, width, channels = shape
        else:
            channels, height, width = shape
        self.math.imagecopy(height, width, channels, A.buffer(), A.offset(), B.buffer(), B.offset(), bool(channels_first),
                            heightShift or 0, widthShift or 0, bool(verticalFlip), bool(horizontalFlip), bool(rgbFlip))
        return B

    def cumsum(self, inputs: NDArray, axis=None, exclusive=None, reverse=None, outputs=None) -> NDArray:
        """LinearAlgebra::cumsum (LinearAlgebra.php:4764-4812): running sum along one axis through cumsumb."""
        ndim = inputs.ndim()
        orig = axis
        axis = 0 if axis is None else axis
        if axis < 0:
            axis += ndim
        if axis < 0 or axis > ndim - 1:
            raise InvalidArgumentException(f"Invalid axis: {'null' if orig is None else orig}")
        shape = inputs.shape()
        m = int(np.prod(shape[:axis])) if axis else 1
        n = shape[axis]
        k = int(np.prod(shape[axis + 1:])) if axis + 1 < ndim else 1
        if outputs is None:
            outputs = self.alloc(shape, inputs.dtype())
        if inputs.shape() != outputs.shape():
            raise InvalidArgumentException(f"Unmatch shape of dimension: {inputs.shape()},{outputs.shape()}")
        self.math.cumsumb(m, n, k, inputs.buffer(), inputs.offset(), bool(exclusive), bool(reverse),
                          outputs.buffer(), outputs.offset())
        return outputs

    def searchsorted(self, A: NDArray, X: NDArray, right=None, dtype=None, Y=None) -> NDArray:
        """LinearAlgebra::searchsorted (LinearAlgebra.php:4700-4762): 1-D A is shared by every X, 2-D A pairs row i
        with X[i]."""
        if A.ndim() == 1:
            individual = False
        elif A.ndim() == 2:
            individual = True
        else:
            raise InvalidArgumentException("A must be 1D or 2D NDArray.")
        right = bool(right)
        if dtype is None:
            dtype = dtypes.int32
        if Y is None:
            Y = self.alloc(X.shape(), dtype)
        if Y.dtype() not in (dtypes.uint32, dtypes.int32, dtypes.uint64, dtypes.int64):
            raise InvalidArgumentException("dtype of Y must be int32 or int64")
        if X.shape() != Y.shape():
            raise InvalidArgumentException(f"Unmatch shape of dimension: {X.shape()},{Y.shape()}")
        if individual:
            m, n = A.shape()
            if m != X.size():
                raise InvalidArgumentException(f"Unmatch shape of dimension A,X: {A.shape()},{X.shape()}")
            ldA = n
        else:
            m, n, ldA = X.size(), A.size(), 0
        self.math.searchsorted(m, n, A.buffer(), A.offset(), ldA, X.buffer(), X.offset(), 1, right,
                               Y.buffer(), Y.offset(), 1)
        return Y

    def bandpart(self, A: NDArray, lower: int, upper: int) -> NDArray:           # LinearAlgebra::bandpart
        if A.ndim() < 2:
            raise InvalidArgumentException("input array must be 2D or upper.")
        shape = A.shape()
        k = shape.pop()
        n = shape.pop()
        m = int(np.prod(shape)) if shape else 1
        self.math.bandpart(m, n, k, A.buffer(), A.offset(), lower, upper)
        return A

    # ---- masking (LinearAlgebra.php:5325-5432) -------------------------------------------------------------------
    def masking(self, mask: NDArray, data: NDArray, batchDims=None, axis=None, fill=None, mode=None) -> NDArray:
        if mask.dtype() != dtypes.bool_:
            raise InvalidArgumentException("dtype of mask must be bool.: " + dtypes.NAMES.get(mask.dtype(), "?"))
        batchDims = 0 if batchDims is None else batchDims
        if batchDims < 0:
            batchDims += data.ndim()
        if batchDims > data.ndim():
            raise InvalidArgumentException("batchDims (%d) must be less than dims of data (%d) or equal." % (batchDims, data.ndim()))
        axis = batchDims if axis is None else axis
        if axis < 0:
            axis += data.ndim()
        if axis > data.ndim():
            raise InvalidArgumentException("axis (%d) must be less than to ndims of data (%d)" % (axis, data.ndim()))
        if batchDims > axis:
            if axis == 0:
                batchDims = 0
            else:
                raise InvalidArgumentException("batchDims (%d) must be less than or equal to axis (%d)" % (batchDims, axis))
        if mask.ndim() < batchDims:
            raise InvalidArgumentException("batchDims (%d) must be less than or equal to ndims of mask (%d)" % (batchDims, mask.ndim()))
        fill = 0 if fill is None else fill
        mode = 0 if mode is None else mode
        shape = data.shape()
        outer, rest = shape[:batchDims], shape[batchDims:]
        broadcast, rest = rest[:axis - batchDims], rest[axis - batchDims:]
        inner, inner_broadcast = rest[:mask.ndim() - batchDims], rest[mask.ndim() - batchDims:]
        mshape = mask.shape()
        outer_x, inner_x = mshape[:batchDims], mshape[batchDims:]
        names = "mask(%s) => data(%s)" % (",".join(map(str, mask.shape())), ",".join(map(str, data.shape())))
        if outer != outer_x:
            raise InvalidArgumentException("Unmatch dimension outer shape.: " + names)
        if inner != inner_x:
            raise InvalidArgumentException("Unmatch dimension inner shape.: " + names)
        prod = lambda v: int(np.prod(v)) if v else 1
        m, n, k, ln = prod(outer), prod(broadcast), prod(inner), prod(inner_broadcast)
        if n == 1:
            k, m = m * k, 1
        if k == 1:
            k, ln, m, n = m, n * ln, 1, 1
        self.math.masking(m, n, k, ln, fill, mode, mask.buffer(), mask.offset(), data.buffer(), data.offset())
        return data

    # ---- slice / stick / stack / concat / split / repeat (LinearAlgebra.php:3944-4365) -----------------------
    def slice(self, input: NDArray, begin, size, output: NDArray | None = None) -> NDArray:
        return self._do_slice(False, input, begin, size, output)

    def stick(self, input: NDArray, output: NDArray, begin, size) -> NDArray:
        return self._do_slice(True, output, begin, size, input)

    def stack(self, values, axis=None) -> NDArray:
        axis = axis or 0
        if axis not in (0, 1, 2):
            raise InvalidArgumentException("unsuppoted axis")
        for value in values:
            if not isinstance(value, NDArrayInterface):
                raise InvalidArgumentException("values must be array of NDArray")
        shape = values[0].shape()
        output = self.alloc(shape[:axis] + [len(values)] + shape[axis:], dtype=values[0].dtype())
        for i, value in enumerate(values):
            vs = value.shape()
            self._do_slice(True, output, [0] * axis + [i], [-1] * axis + [1], value.reshape(vs[:axis] + [1] + vs[axis:]))
        return output

    def concat(self, values, axis=None) -> NDArray:
        if axis is None:
            axis = -1
        if axis < 0:
            axis = values[0].ndim() + axis
        m = base = None
        n = 0
        reshaped = []
        prefix, shape = [], []
        for value in values:
            shape = value.shape()
            prefix, shape = shape[:axis], shape[axis:]
            mm = int(np.prod(prefix)) if prefix else 1
            nn = shape.pop(0)
            if base is None:
                m, base = mm, shape
            elif m != mm or base != shape:
                raise InvalidArgumentException("Unmatch shape: " + ",".join("(" + ",".join(map(str, v.shape())) + ")" for v in values))
            n += nn
            reshaped.append(value.reshape([mm, nn] + shape))
        output = self.alloc([m, n] + shape, dtype=values[0].dtype())
        i = 0
        for value in reshaped:
            nn = value.shape()[1]
            self._do_slice(True, output, [0, i], [-1, nn], value)
            i += nn
        return output.reshape(prefix + [n] + shape)

    def split(self, input: NDArray, sizeSplits, axis=None):
        if axis is None:
            axis = -1
        if axis < 0:
            axis = input.ndim() + axis
        shape = input.shape()
        prefix, shape = shape[:axis], shape[axis:]
        m = int(np.prod(prefix)) if prefix else 1
        n = shape.pop(0)
        input = input.reshape([m, n] + shape)
        i = 0
        outputs = []
        for size in sizeSplits:
            outputs.append(self._do_slice(False, input, [0, i], [-1, size]).reshape(prefix + [size] + shape))
            i += size
        return outputs

    def _do_slice(self, reverse, input: NDArray, begin, size, output: NDArray | None = None) -> NDArray:   # :4163-4303
        message_input = "Output" if reverse else "Input"
        org_begin = list(begin)
        begin, size = list(begin), list(size)
        if not 1 <= len(begin) <= 3:
            raise InvalidArgumentException("begin must has 1 or 2 or 3 integer.")
        if not 1 <= len(size) <= 3:
            raise InvalidArgumentException("Size must has 1 or 2 or 3 integer.")
        if len(begin) != len(size):
            raise InvalidArgumentException("Unmatch shape of begin and size")
        nb = len(begin)
        if input.ndim() < nb:
            raise InvalidArgumentException(message_input + " shape rank is low to slice")
        shape = input.shape()
        dims, starts, sizes = [1, 1, 1], [0, 0, 0], [1, 1, 1]
        for ax in range(nb):
            d = shape.pop(0)
            st = begin[ax]
            if st < 0:
                st = d + st
            if st < 0 or st >= d:
                raise InvalidArgumentException("start of axis %d is invalid value.%s" % (
                    ax, "" if ax == 0 else ":begin=[" + ",".join(map(str, org_begin)) + "]"))
            sz = size[ax]
            if sz < 0:
                sz = d - st + sz + 1
            if sz < 1 or st + sz > d:
                raise InvalidArgumentException("size of axis %d is invalid value." % ax)
            dims[ax], starts[ax], sizes[ax] = d, st, sz
        item_size = int(np.prod(shape)) if shape else 1
        output_shape = sizes[:nb] + shape
        if output is None:
            output = self.alloc(output_shape, dtype=input.dtype())
        elif output_shape != output.shape():
            raise InvalidArgumentException("Unmatch output shape: (%s)<=>(%s)" % (
                ",".join(map(str, output_shape)), ",".join(map(str, output.shape()))))
        self.math.slice(reverse, False, dims[0], dims[1], dims[2], item_size, input.buffer(), input.offset(), 1,
                        output.buffer(), output.offset(), 1, starts[0], sizes[0], starts[1], sizes[1], starts[2], sizes[2])
        return output

    def repeat(self, A: NDArray, repeats: int, axis=None, keepdims=None) -> NDArray:                        # :4308-4365
        if repeats < 1:
            raise InvalidArgumentException("repeats argument must be one or greater.")
        if axis is not None:
            if axis < 0:
                axis = A.ndim() + axis
            if A.ndim() < axis:
                raise InvalidArgumentException("dimension rank must be two or greater.")
        inner = A.shape()
        outer = []
        if axis is not None:
            outer, inner = inner[:axis], inner[axis:]
        base = 1
        if axis is None:
            output_shape = [int(np.prod(outer + [repeats] + inner))]
        elif keepdims:
            if not inner:
                raise InvalidArgumentException("dimension rank must be two or greater on keepdims.")
            base = inner.pop(0)
            output_shape = outer + [repeats * base] + inner
        else:
            output_shape = outer + [repeats] + inner
        B = self.alloc(output_shape, dtype=A.dtype())
        m = int(np.prod(outer)) if outer else 1
        k = (int(np.prod(inner)) if inner else 1) * base
        self.math.repeat(m, k, repeats, A.buffer(), A.offset(), B.buffer(), B.offset())
        return B

    def im2col(self, images: NDArray, filterSize=None, strides=None, padding=None, channels_first=None,
               dilation_rate=None, cols_channels_first=None, cols: NDArray | None = None) -> NDArray:
        fn = {3: self.im2col1d, 4: self.im2col2d, 5: self.im2col3d}.get(images.ndim())      # LinearAlgebra.php:3369-3398
        if fn is None:
            raise InvalidArgumentException("unsuppoted images shape")
        _cols = fn(False, images, filterSize, strides, padding, channels_first, dilation_rate, cols_channels_first, cols)
        return _cols if cols is None else cols

    def col2im(self, cols: NDArray, images: NDArray, filterSize=None, strides=None, padding=None,
               channels_first=None, dilation_rate=None, cols_channels_first=None) -> NDArray:
        fn = {3: self.im2col1d, 4: self.im2col2d, 5: self.im2col3d}.get(images.ndim())      # :3406-3456
        if fn is None:
            raise InvalidArgumentException("unsuppoted images shape")
        fn(True, images, filterSize, strides, padding, channels_first, dilation_rate, cols_channels_first, cols)
        return images

    def im2col1d(self, reverse, images: NDArray, filterSize=None, strides=None, padding=None,
                 channels_first=None, dilation_rate=None, cols_channels_first=None,
                 cols: NDArray | None = None) -> NDArray:                        # LinearAlgebra.php:3496-3590
        if images.ndim() != 3:
            raise InvalidArgumentException("images must be 3D dimension")
        if channels_first:
            batches, channels, in_w = images.shape()
        else:
            batches, in_w, channels = images.shape()
        (filter_w,) = filterSize if filterSize else (3,)
        (stride_w,) = strides if strides else (1,)
        padding = bool(padding)
        channels_first = bool(channels_first)
        (dilation_w,) = dilation_rate if dilation_rate else (1,)
        cols_channels_first = bool(cols_channels_first)
        if cols is None:
            out_w = in_w if padding else _intdiv(in_w - (filter_w - 1) * dilation_w - 1, stride_w) + 1
            if out_w <= 0:
                raise InvalidArgumentException("Invalid shape or paramaters.")
            shape = [batches, out_w, channels, filter_w] if cols_channels_first else [batches, out_w, filter_w, channels]
            cols = self.zeros(self.alloc(shape, dtype=images.dtype()))
        self.math.im2col1d(reverse, images.buffer(), images.offset(), images.size(), batches, in_w, channels,
                           filter_w, stride_w, padding, channels_first, dilation_w, cols_channels_first,
                           cols.buffer(), cols.offset(), cols.size())
        return cols

    def im2col3d(self, reverse, images: NDArray, filterSize=None, strides=None, padding=None,
                 channels_first=None, dilation_rate=None, cols_channels_first=None,
                 cols: NDArray | None = None) -> NDArray:                        # LinearAlgebra.php:3708-3818
        if images.ndim() != 5:
            raise InvalidArgumentException("images must be 5D dimension")
        if channels_first:
            batches, channels, in_d, in_h, in_w = images.shape()
        else:
            batches, in_d, in_h, in_w, channels = images.shape()
        filter_d, filter_h, filter_w = filterSize if filterSize else (3, 3, 3)
        stride_d, stride_h, stride_w = strides if strides else (1, 1, 1)
        padding = bool(padding)
        channels_first = bool(channels_first)
        dilation_d, dilation_h, dilation_w = dilation_rate if dilation_rate else (1, 1, 1)
        cols_channels_first = bool(cols_channels_first)
        if cols is None:
            if padding:
                out_d, out_h, out_w = in_d, in_h, in_w
            else:
                out_d = _intdiv(in_d - (filter_d - 1) * dilation_d - 1, stride_d) + 1
                out_h = _intdiv(in_h - (filter_h - 1) * dilation_h - 1, stride_h) + 1
                out_w = _intdiv(in_w - (filter_w - 1) * dilation_w - 1, stride_w) + 1
            if out_d <= 0 or out_h <= 0 or out_w <= 0:
                raise InvalidArgumentException("Invalid shape or paramaters.")
            if cols_channels_first:
                shape = [batches, out_d, out_h, out_w, channels, filter_d, filter_h, filter_w]
            else:
                shape = [batches, out_d, out_h, out_w, filter_d, filter_h, filter_w, channels]
            cols = self.zeros(self.alloc(shape, dtype=images.dtype()))
        self.math.im2col3d(reverse, images.buffer(), images.offset(), images.size(), batches, in_d, in_h, in_w,
                           channels, filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, padding,
                           channels_first, dilation_d, dilation_h, dilation_w, cols_channels_first,
                           cols.buffer(), cols.offset(), cols.size())
        return cols

    def conv2d(self, images: NDArray, kernel: NDArray, bias: NDArray | None = None, strides=None, padding=None,
               dilation_rate=None, fused: bool | None = None) -> NDArray:
        """Conv2D forward for channels-last images [b,h,w,c] and kernel [kh,kw,c,filters] -- the composition
        the reference's users write by hand (im2col :3369 -> reshape -> gemm :925 -> add :1968; the Conv2D
        layer itself lives in rindow-neuralnetworks).  On any driver: exactly those three calls (on the B200
        driver: tiled im2col2d + software-staged tcgen05 GEMM, 0.25 ms for BASELINE config C4).  On the B200
        driver the single-kernel implicit-GEMM entry (rb200_conv2d_forward: no cols round trip) is taken by
        default for short reductions, kh*kw*c <= 32, where it runs on the TMEM-operand kernel (0.14 ms for C4,
        0.92 of HBM), and for hidden layers with channels % 32 == 0, where the GEMM's A operand is fetched by TMA
        straight from the image tensor (conv_tcgen05_tma.cu); fused=True asks for it at any size it supports,
        fused=False never.  The three calls
        remain the fallback whenever the driver declines the shape."""
        if images.ndim() != 4 or kernel.ndim() != 4:
            raise InvalidArgumentException("images and kernel must be 4D dimension")
        batches, in_h, in_w, channels = images.shape()
        filter_h, filter_w, kc, filters = kernel.shape()
        if kc != channels:
            raise InvalidArgumentException("Unmatch channels of images and kernel")
        stride_h, stride_w = strides if strides else (1, 1)
        dilation_h, dilation_w = dilation_rate if dilation_rate else (1, 1)
        padding = bool(padding)
        if padding:
            out_h, out_w = in_h, in_w
        else:
            out_h = _intdiv(in_h - (filter_h - 1) * dilation_h - 1, stride_h) + 1
            out_w = _intdiv(in_w - (filter_w - 1) * dilation_w - 1, stride_w) + 1
        if out_h <= 0 or out_w <= 0:
            raise InvalidArgumentException("Invalid shape or paramaters.")
        if bias is not None and bias.shape() != [filters]:
            raise InvalidArgumentException("Unmatch shape of bias and kernel")
        if fused is None:
            # the single-kernel forms that beat im2col + gemm: short reductions (TMEM-operand kernel) and hidden layers
            # with channels % 32 == 0 (TMA-fed implicit GEMM, no cols buffer at all)
            fused = filter_h * filter_w * channels <= 32 or (channels % 32 == 0 and filters % 4 == 0 and filters >= 8)
        fused_fn = getattr(self.math, "conv2d_forward", None) if (fused and type(self.math).__name__ == "CudaMath") else None
        if fused_fn is not None:
            fused = fused_fn
            out = self.alloc([batches, out_h, out_w, filters], dtype=images.dtype())
            if fused(images.buffer(), images.offset(), batches, in_h, in_w, channels, filter_h,
                                           filter_w, stride_h, stride_w, padding, dilation_h, dilation_w,
                                           kernel.buffer(), kernel.offset(), filters,
                                           None if bias is None else bias.buffer(), 0 if bias is None else bias.offset(),
                                           out.buffer(), out.offset()):
                return out
        cols = self.im2col(images, filterSize=[filter_h, filter_w], strides=[stride_h, stride_w], padding=padding,
                           dilation_rate=[dilation_h, dilation_w])
        out2d = self.gemm(cols.reshape([batches * out_h * out_w, filter_h * filter_w * channels]),
                          kernel.reshape([filter_h * filter_w * channels, filters]))
        if bias is not None:
            self.add(bias, out2d)
        return out2d.reshape([batches, out_h, out_w, filters])

    def im2col2d(self, reverse, images: NDArray, filterSize=None, strides=None, padding=None,
                 channels_first=None, dilation_rate=None, cols_channels_first=None,
                 cols: NDArray | None = None) -> NDArray:
        if images.ndim() != 4:
            raise InvalidArgumentException("images must be 4D dimension")
        if channels_first:
            batches, channels, in_h, in_w = images.shape()
        else:
            batches, in_h, in_w, channels = images.shape()
        filter_h, filter_w = filterSize if filterSize else (3, 3)
        stride_h, stride_w = strides if strides else (1, 1)
        padding = bool(padding)
        channels_first = bool(channels_first)
        dilation_h, dilation_w = dilation_rate if dilation_rate else (1, 1)
        cols_channels_first = bool(cols_channels_first)
        if cols is None:
            if padding:
                out_h, out_w = in_h, in_w
            else:
                out_h = _intdiv(in_h - (filter_h - 1) * dilation_h - 1, stride_h) + 1
                out_w = _intdiv(in_w - (filter_w - 1) * dilation_w - 1, stride_w) + 1
            if out_h <= 0 or out_w <= 0:
                raise InvalidArgumentException("Invalid shape or paramaters.")
            if cols_channels_first:
                shape = [batches, out_h, out_w, channels, filter_h, filter_w]
            else:
                shape = [batches, out_h, out_w, filter_h, filter_w, channels]
            cols = self.zeros(self.alloc(shape, dtype=images.dtype()))
        self.math.im2col2d(reverse, images.buffer(), images.offset(), images.size(), batches, in_h, in_w,
                           channels, filter_h, filter_w, stride_h, stride_w, padding, channels_first,
                           dilation_h, dilation_w, cols_channels_first, cols.buffer(), cols.offset(), cols.size())
        return cols
