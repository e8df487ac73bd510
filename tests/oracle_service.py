"""A driver service backed by the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Plays the role of the reference's MatlibPhp service (src/Drivers/MatlibPHP/MatlibPhp.php): the
same LinearAlgebra / MatrixOperator / NDArray host classes run over it unchanged, which is how
the reference itself swaps drivers under one test-suite ("driver-swap by
inheritance").  Buffers hold float64 (PHP doubles) or int64 (PHP ints), like PhpBuffer.
"""
import numpy as np

import oracle
from rindow_math_matrix_b200 import dtypes
from rindow_math_matrix_b200.exceptions import InvalidArgumentException, RuntimeException


class OracleBuffer:
    def __init__(self, size, dtype):
        if dtype not in dtypes.NUMPY:
            raise InvalidArgumentException("Invalid data type")
        self._dtype = dtype
        self._is_int = dtype in dtypes.INT_TYPES or dtype == dtypes.bool_
        self.data = np.zeros(size, np.int64 if self._is_int else np.float64)

    def __len__(self):
        return self.data.size

    def dtype(self):
        return self._dtype

    def value_size(self):
        return dtypes.VALUE_SIZE[self._dtype]

    def __getitem__(self, i):
        if i < 0 or i >= self.data.size:
            raise RuntimeException("Index invalid or out of range")
        return self.data[i].item()

    def __setitem__(self, i, v):
        if i < 0 or i >= self.data.size:
            raise RuntimeException("Index invalid or out of range")
        self.data[i] = v

    def from_numpy(self, a, offset=0):
        a = np.asarray(a).reshape(-1)
        self.data[offset:offset + a.size] = a
        return self

    def to_numpy(self):
        return self.data.copy()

    def f64(self):
        if self._is_int:
            raise InvalidArgumentException("integer buffer in a float op")
        return self.data


def _wrap(fn, *args, **kw):
    try:
        return fn(*args, **kw)
    except oracle.OracleArgumentError as e:
        raise InvalidArgumentException(str(e))
    except oracle.OracleZeroDivide as e:
        raise RuntimeException(str(e))


class OracleBlas:
    def getParallel(self):
        return 0

    def hasOmatcopy(self):
        return True

    def gemm(self, order, transA, transB, m, n, k, alpha, A, offA, ldA, B, offB, ldB, beta, C, offC, ldC):
        _wrap(oracle.gemm, order, transA, transB, m, n, k, alpha, A.f64(), offA, ldA, B.f64(), offB, ldB,
              beta, C.f64(), offC, ldC)


    def scal(self, n, alpha, X, offX, incX):
        _wrap(oracle.scal, n, alpha, X.f64(), offX, incX)

    def axpy(self, n, alpha, X, offX, incX, Y, offY, incY):
        _wrap(oracle.axpy, n, alpha, X.f64(), offX, incX, Y.f64(), offY, incY)

    def copy(self, n, X, offX, incX, Y, offY, incY):
        _wrap(oracle.copy, n, X.f64(), offX, incX, Y.f64(), offY, incY)

    def gemv(self, order, trans, m, n, alpha, A, offA, ldA, X, offX, incX, beta, Y, offY, incY):
        _wrap(oracle.gemv, order, trans, m, n, alpha, A.f64(), offA, ldA, X.f64(), offX, incX,
              beta, Y.f64(), offY, incY)

    def omatcopy(self, order, trans, m, n, alpha, A, offA, ldA, B, offB, ldB):
        _wrap(oracle.omatcopy, order, trans, m, n, alpha, A.f64(), offA, ldA, B.f64(), offB, ldB)


class OracleMath:
    sticky_max = False      # False: PhpMath::math_max rule; True: the accelerated-path rule

    def getParallel(self):
        return 0

    def add(self, trans, m, n, alpha, X, offX, incX, A, offA, ldA):
        _wrap(oracle.add, trans, m, n, alpha, X.f64(), offX, incX, A.f64(), offA, ldA)

    def multiply(self, trans, m, n, X, offX, incX, A, offA, ldA):
        _wrap(oracle.multiply, trans, m, n, X.f64(), offX, incX, A.f64(), offA, ldA)

    def reduceSum(self, m, n, k, A, offA, B, offB):
        _wrap(oracle.reduce_sum, m, n, k, A.f64(), offA, B.f64(), offB)

    def reduceMax(self, m, n, k, A, offA, B, offB):
        _wrap(oracle.reduce_max, m, n, k, A.f64(), offA, B.f64(), offB, sticky=self.sticky_max)

    def reduceArgMax(self, m, n, k, A, offA, B, offB):
        _wrap(oracle.reduce_argmax, m, n, k, A.f64(), offA, B.data, offB)

    def softmax(self, m, n, A, offA, ldA):
        _wrap(oracle.softmax, m, n, A.f64(), offA, ldA)

    def im2col2d(self, reverse, images, images_offset, images_size, batches, im_h, im_w, channels,
                 filter_h, filter_w, stride_h, stride_w, padding, channels_first, dilation_h, dilation_w,
                 cols_channels_first, cols, cols_offset, cols_size):
        _wrap(oracle.im2col2d, reverse, images.f64(), images_offset, images_size, batches, im_h, im_w, channels,
              filter_h, filter_w, stride_h, stride_w, padding, channels_first, dilation_h, dilation_w,
              cols_channels_first, cols.f64(), cols_offset, cols_size)

    def zeros(self, n, X, offX, incX):
        X.data[offX:offX + n * incX:incX] = 0

    def fill(self, n, V, offV, X, offX, incX):
        X.data[offX:offX + n * incX:incX] = V[offV]


class OracleBufferFactory:
    def isAvailable(self):
        return True

    def Buffer(self, size, dtype):
        return OracleBuffer(size, dtype)


class OracleService:
    LV_BASIC, LV_ADVANCED = 1, 2

    def __init__(self, sticky_max=False):
        self._blas, self._math, self._buffer = OracleBlas(), OracleMath(), OracleBufferFactory()
        self._math.sticky_max = sticky_max

    def serviceLevel(self):
        return 1

    def name(self):
        return "oracle"

    def blas(self, level=None):
        return self._blas

    def math(self, level=None):
        return self._math

    def buffer(self, level=None):
        return self._buffer
