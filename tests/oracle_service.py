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
    except (oracle.OracleZeroDivide, oracle.OracleLabelRange) as e:
        raise RuntimeException(str(e))


class OracleBlas:
    def getParallel(self):
        return 0

    def hasOmatcopy(self):
        return True

    def hasIamin(self):
        return True

    def dot(self, n, X, offX, incX, Y, offY, incY):
        return _wrap(oracle.dot, n, X.f64(), offX, incX, Y.f64(), offY, incY)

    def asum(self, n, X, offX, incX):
        return _wrap(oracle.asum, n, X.f64(), offX, incX)

    def nrm2(self, n, X, offX, incX):
        return _wrap(oracle.nrm2, n, X.f64(), offX, incX)

    def iamax(self, n, X, offX, incX):
        return _wrap(oracle.iamax, n, X.f64(), offX, incX)

    def iamin(self, n, X, offX, incX):
        return _wrap(oracle.iamin, n, X.f64(), offX, incX)

    def swap(self, n, X, offX, incX, Y, offY, incY):
        _wrap(oracle.swap, n, X.f64(), offX, incX, Y.f64(), offY, incY)

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

    def reduceArgMaxPartial(self, m, n, k, A, offA, global0, records, offRecords):
        """The stripe record of CudaMath.reduceArgMaxPartial, restated from math_imax (PhpMath.php:114-130): the
        stripe at global position 0 IS math_imax (value read back at the winning index, a NaN there as +INF); any
        other stripe continues the loop `if($value > $max)` from a running maximum, i.e. from -INF with NaNs never
        taken.  Pure-Python loop: small cases only."""
        a = A.f64()[offA:offA + m * n * k].reshape(m, n, k)
        vals = np.empty((m, k))
        idxs = np.empty((m, k), np.int64)
        if global0:
            tmp = np.zeros(m * k, np.int64)
            _wrap(oracle.reduce_argmax, m, n, k, A.f64(), offA, tmp, 0)
            idxs[:] = tmp.reshape(m, k)
            for i in range(m):
                for j in range(k):
                    v = a[i, idxs[i, j], j]
                    vals[i, j] = np.inf if np.isnan(v) else v
        else:
            for i in range(m):
                for j in range(k):
                    mx, ix = -np.inf, np.iinfo(np.int32).max
                    for t in range(n):
                        if a[i, t, j] > mx:
                            mx, ix = a[i, t, j], t
                    vals[i, j], idxs[i, j] = mx, ix
        rec = records.data[2 * offRecords:2 * (offRecords + m * k)].reshape(m * k, 2)
        rec[:, 0] = vals.reshape(-1).view(np.int64)
        rec[:, 1] = idxs.reshape(-1)

    def softmax(self, m, n, A, offA, ldA):
        _wrap(oracle.softmax, m, n, A.f64(), offA, ldA)

    def im2col2d(self, reverse, images, images_offset, images_size, batches, im_h, im_w, channels,
                 filter_h, filter_w, stride_h, stride_w, padding, channels_first, dilation_h, dilation_w,
                 cols_channels_first, cols, cols_offset, cols_size):
        _wrap(oracle.im2col2d, reverse, images.f64(), images_offset, images_size, batches, im_h, im_w, channels,
              filter_h, filter_w, stride_h, stride_w, padding, channels_first, dilation_h, dilation_w,
              cols_channels_first, cols.f64(), cols_offset, cols_size)

    def im2col1d(self, reverse, images, images_offset, images_size, batches, im_w, channels, filter_w, stride_w,
                 padding, channels_first, dilation_w, cols_channels_first, cols, cols_offset, cols_size):
        _wrap(oracle.im2col1d, reverse, images.f64(), images_offset, images_size, batches, im_w, channels, filter_w,
              stride_w, padding, channels_first, dilation_w, cols_channels_first, cols.f64(), cols_offset, cols_size)

    def im2col3d(self, reverse, images, images_offset, images_size, batches, im_d, im_h, im_w, channels,
                 filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, padding, channels_first,
                 dilation_d, dilation_h, dilation_w, cols_channels_first, cols, cols_offset, cols_size):
        _wrap(oracle.im2col3d, reverse, images.f64(), images_offset, images_size, batches, im_d, im_h, im_w, channels,
              filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, padding, channels_first,
              dilation_d, dilation_h, dilation_w, cols_channels_first, cols.f64(), cols_offset, cols_size)

    # ---- SURVEY.md section 8(f) widening: broadcast / unary / compare families, astype, scalar reductions, onehot
    def _bc(self, op, trans, m, n, A, offA, ldA, X, offX, incX):
        a, x = self._f(A), self._f(X)
        _wrap(oracle.broadcast, op, trans, m, n, a, offA, ldA, x, offX, incX)
        self._back(A, a)

    @staticmethod
    def _f(buf):
        """float64 working copy of any buffer (PHP ints and floats share one arithmetic)."""
        return buf.data if buf.data.dtype == np.float64 else buf.data.astype(np.float64)

    @staticmethod
    def _back(buf, work):
        if work is not buf.data:
            buf.data[:] = work.astype(np.int64)

    def maximum(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_MAXIMUM, False, m, n, A, offA, ldA, X, offX, incX)

    def minimum(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_MINIMUM, False, m, n, A, offA, ldA, X, offX, incX)

    def greater(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_GREATER, False, m, n, A, offA, ldA, X, offX, incX)

    def greaterEqual(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_GREATER_EQUAL, False, m, n, A, offA, ldA, X, offX, incX)

    def less(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_LESS, False, m, n, A, offA, ldA, X, offX, incX)

    def lessEqual(self, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_LESS_EQUAL, False, m, n, A, offA, ldA, X, offX, incX)

    def pow(self, trans, m, n, A, offA, ldA, X, offX, incX):
        self._bc(oracle.BC_POW, trans, m, n, A, offA, ldA, X, offX, incX)

    def duplicate(self, trans, m, n, X, offX, incX, A, offA, ldA):
        self._bc(oracle.BC_DUPLICATE, trans, m, n, A, offA, ldA, X, offX, incX)

    def _un(self, op, n, X, offX, incX, alpha=1.0, beta=0.0):
        x = self._f(X)
        _wrap(oracle.unary, op, n, x, offX, incX, alpha, beta)
        self._back(X, x)

    def increment(self, n, alpha, X, offX, incX, beta):
        self._un(oracle.UN_INCREMENT, n, X, offX, incX, alpha, beta)

    def reciprocal(self, n, alpha, X, offX, incX, beta):
        self._un(oracle.UN_RECIPROCAL, n, X, offX, incX, alpha, beta)

    def rsqrt(self, n, alpha, X, offX, incX, beta):
        self._un(oracle.UN_RSQRT, n, X, offX, incX, alpha, beta)

    def square(self, n, X, offX, incX):
        self._un(oracle.UN_SQUARE, n, X, offX, incX)

    def sqrt(self, n, X, offX, incX):
        self._un(oracle.UN_SQRT, n, X, offX, incX)

    def exp(self, n, X, offX, incX):
        self._un(oracle.UN_EXP, n, X, offX, incX)

    def log(self, n, X, offX, incX):
        self._un(oracle.UN_LOG, n, X, offX, incX)

    def tanh(self, n, X, offX, incX):
        self._un(oracle.UN_TANH, n, X, offX, incX)

    def sin(self, n, X, offX, incX):
        self._un(oracle.UN_SIN, n, X, offX, incX)

    def cos(self, n, X, offX, incX):
        self._un(oracle.UN_COS, n, X, offX, incX)

    def tan(self, n, X, offX, incX):
        self._un(oracle.UN_TAN, n, X, offX, incX)

    def not_(self, n, X, offX, incX):
        self._un(oracle.UN_NOT, n, X, offX, incX)

    def __getattr__(self, name):
        if name == "not":
            return self.not_
        raise AttributeError(name)

    def nan2num(self, n, X, offX, incX, alpha):
        self._un(oracle.UN_NAN2NUM, n, X, offX, incX, alpha)

    def isnan(self, n, X, offX, incX):
        self._un(oracle.UN_ISNAN, n, X, offX, incX)

    def equal(self, n, X, offX, incX, Y, offY, incY):
        y = self._f(Y)
        _wrap(oracle.compare, False, n, self._f(X), offX, incX, y, offY, incY)
        self._back(Y, y)

    def notEqual(self, n, X, offX, incX, Y, offY, incY):
        y = self._f(Y)
        _wrap(oracle.compare, True, n, self._f(X), offX, incX, y, offY, incY)
        self._back(Y, y)

    def astype(self, n, dtype, X, offX, incX, Y, offY, incY):
        if dtype == dtypes.bool_:
            kind, mask = "bool", 0
        elif dtype in dtypes.FLOAT_TYPES:
            kind, mask = "float", 0
        else:
            kind = "int"
            mask = {dtypes.uint8: 0xff, dtypes.uint16: 0xffff, dtypes.uint32: 0xffffffff}.get(dtype, 0)
        Y.data[offY:offY + (n - 1) * incY + 1:incY] = oracle.astype(n, kind, mask, X.data, offX, incX)

    def sum(self, n, X, offX, incX):
        return _wrap(oracle.sum, n, self._f(X), offX, incX)

    def imax(self, n, X, offX, incX):
        return _wrap(oracle.imax, n, self._f(X), offX, incX)

    def imin(self, n, X, offX, incX):
        return _wrap(oracle.imin, n, self._f(X), offX, incX)

    def updateAddOnehot(self, m, n, a, X, offX, incX, Y, offY, ldY):
        _wrap(oracle.onehot, m, n, a, self._f(X), offX, incX, Y.f64(), offY, ldY)

    def gather(self, reverse, addMode, n, k, numClass, X, offX, A, offA, B, offB):
        a, b = self._f(A), self._f(B)
        _wrap(oracle.gather, reverse, addMode, n, k, numClass, self._f(X), offX, a, offA, b, offB)
        self._back(A, a)
        self._back(B, b)

    def reduceGather(self, reverse, addMode, m, n, numClass, X, offX, A, offA, B, offB):
        a, b = self._f(A), self._f(B)
        _wrap(oracle.reduce_gather, reverse, addMode, m, n, numClass, self._f(X), offX, a, offA, b, offB)
        self._back(A, a)
        self._back(B, b)

    def gatherb(self, reverse, addMode, batches, m, n, k, len_, numClass, A, offA, X, offX, B, offB):
        a, b = self._f(A), self._f(B)
        _wrap(oracle.gatherb, reverse, addMode, A.dtype() == dtypes.float32, batches, m, n, k, len_, numClass, a, offA,
              self._f(X), offX, b, offB)
        self._back(A, a)
        self._back(B, b)

    def gathernd(self, reverse, addMode, m, n, k, indexDepth, paramShape, A, offA, X, offX, B, offB):
        a, b = self._f(A), self._f(B)
        dims = [int(paramShape[h]) for h in range(indexDepth)]
        _wrap(oracle.gathernd, reverse, addMode, A.dtype() == dtypes.float32, m, n, k, indexDepth, dims, a, offA, self._f(X),
              b, offB)
        self._back(A, a)
        self._back(B, b)

    def imagecopy(self, height, width, channels, A, offA, B, offB, channelsFirst, heightShift, widthShift, verticalFlip,
                  horizontalFlip, rgbFlip):
        b = self._f(B)
        _wrap(oracle.imagecopy, height, width, channels, self._f(A), offA, b, offB, channelsFirst, heightShift, widthShift,
              verticalFlip, horizontalFlip, rgbFlip)
        self._back(B, b)

    def einsum(self, sizeOfIndices, A, offA, ldA, B, offB, ldB, C, offC, ndimC):
        depth = len(sizeOfIndices)
        c = self._f(C)
        _wrap(oracle.einsum, [int(sizeOfIndices[h]) for h in range(depth)], self._f(A), offA,
              [int(ldA[h]) for h in range(depth)], self._f(B), offB, [int(ldB[h]) for h in range(depth)], c, offC, ndimC)
        self._back(C, c)

    def topk(self, m, n, input, offInput, k, sorted, values, offValues, indices, offIndices):
        v, ix = self._f(values), self._f(indices)
        _wrap(oracle.topk, m, n, self._f(input), offInput, k, sorted, v, offValues, ix, offIndices)
        self._back(values, v)
        self._back(indices, ix)

    topK = topk

    def cumsumb(self, m, n, k, A, offA, exclusive, reverse, B, offB):
        b = self._f(B)
        _wrap(oracle.cumsumb, m, n, k, self._f(A), offA, exclusive, reverse, A.dtype() == dtypes.float32, b, offB)
        self._back(B, b)

    def cumsum(self, n, X, offX, incX, exclusive, reverse, Y, offY, incY):
        y = self._f(Y)
        _wrap(oracle.cumsum, n, self._f(X), offX, incX, exclusive, reverse, y, offY, incY)
        self._back(Y, y)

    def searchsorted(self, m, n, A, offA, ldA, X, offX, incX, right, Y, offY, incY):
        y = self._f(Y)
        _wrap(oracle.searchsorted, m, n, self._f(A), offA, ldA, self._f(X), offX, incX, right, y, offY, incY)
        self._back(Y, y)

    def bandpart(self, m, n, k, A, offA, lower, upper):
        a = self._f(A)
        _wrap(oracle.bandpart, m, n, k, a, offA, lower, upper)
        self._back(A, a)

    def masking(self, m, n, k, len_, fill, mode, X, offX, A, offA):
        if X.dtype() != dtypes.bool_:
            raise InvalidArgumentException("dtype of X must be bool.")
        a = self._f(A)
        kind = 2 if A.dtype() == dtypes.bool_ else (0 if A.dtype() in (dtypes.float32, dtypes.float64) else 1)
        _wrap(oracle.masking, m, n, k, len_, fill, mode, kind, self._f(X), offX, a, offA)
        self._back(A, a)

    def slice(self, reverse, addMode, m, n, k, size, A, offA, incA, Y, offY, incY, st0, sz0, st1, sz1, st2, sz2):
        a, y = self._f(A), self._f(Y)
        _wrap(oracle.slice, reverse, addMode, m, n, k, size, a, offA, incA, y, offY, incY, st0, sz0, st1, sz1, st2, sz2)
        self._back(A, a)
        self._back(Y, y)

    def repeat(self, m, k, repeats, A, offA, B, offB):
        a, b = self._f(A), self._f(B)
        _wrap(oracle.repeat, m, k, repeats, a, offA, b, offB)
        self._back(B, b)

    def transpose(self, sourceShape, perm, A, offA, B, offB):
        a, b = self._f(A), self._f(B)
        _wrap(oracle.transpose, sourceShape.data, perm.data, a, offA, b, offB)
        self._back(B, b)

    def zeros(self, n, X, offX, incX):
        X.data[offX:offX + n * incX:incX] = 0

    def fill(self, n, V, offV, X, offX, incX):
        X.data[offX:offX + n * incX:incX] = V[offV]


class OracleBufferFactory:
    def isAvailable(self):
        return True

    def Buffer(self, size, dtype):
        return OracleBuffer(size, dtype)


class OracleDeviceBuffer(OracleBuffer):
    """DeviceBuffer stand-in (src/NDArrayCL.php:147-168, 261-270, 416-441): bytes() + queue-ordered read / write / copy
    with BYTE sizes and offsets, backed by the same numpy storage as OracleBuffer -- so NDArrayCuda / LinearAlgebraCuda
    (the host logic of the device tier) run on CPU in the not-gpu suite."""
    is_device_buffer = True

    def __init__(self, context, bytes_, flags=0, hostBuffer=None, hostOffset=0, dtype=None):
        if dtype is None:
            dtype = hostBuffer.dtype()
        vs = dtypes.VALUE_SIZE[dtype]
        super().__init__(bytes_ // vs, dtype)
        self._vs = vs
        self._flags = flags
        if hostBuffer is not None:
            self.write(None, hostBuffer, bytes_, 0, hostOffset * vs)

    def bytes(self):
        return self.data.size * self._vs

    def _partner_vs(self, other):
        return dtypes.VALUE_SIZE[other.dtype()]

    def read(self, queue, hostBuf, size=0, offset=0, hostoffset=0, blocking_read=None, events=None, waitEvents=None):
        n = (size or self.bytes() - offset) // self._vs
        o, h = offset // self._vs, hostoffset // self._partner_vs(hostBuf)
        hostBuf.data[h:h + n] = self.data[o:o + n]
        if events is not None:
            events.record()

    def write(self, queue, hostBuf, size=0, offset=0, hostoffset=0, blocking_write=None, events=None, waitEvents=None):
        pvs = self._partner_vs(hostBuf)
        n = (size // self._vs) if size else min(hostBuf.data.size - hostoffset // pvs, self.data.size - offset // self._vs)
        o, h = offset // self._vs, hostoffset // pvs
        self.data[o:o + n] = hostBuf.data[h:h + n]
        if events is not None:
            events.record()

    def copy(self, queue, src, size=0, src_offset=0, dst_offset=0, events=None, waitEvents=None):
        n = (size or src.bytes() - src_offset) // self._vs
        so, do = src_offset // self._vs, dst_offset // self._vs
        self.data[do:do + n] = src.data[so:so + n]
        if events is not None:
            events.record()


class OracleDeviceBufferFactory:
    def isAvailable(self):
        return True

    def Buffer(self, context, bytes_, flags=0, hostBuffer=None, hostOffset=0, dtype=None):
        return OracleDeviceBuffer(context, bytes_, flags, hostBuffer, hostOffset, dtype)


class OracleEventList:
    def __init__(self):
        self.n = 0

    def record(self):
        self.n += 1

    def wait(self):
        pass

    def __len__(self):
        return self.n


class OracleQueue:
    finished = 0

    def getContext(self):
        return self

    def finish(self):
        self.finished += 1

    def newEventList(self):
        return OracleEventList()


class OracleService:
    LV_BASIC, LV_ADVANCED, LV_ACCELERATED = 1, 2, 3

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
        if level == self.LV_ACCELERATED:
            return OracleDeviceBufferFactory()
        return self._buffer

    def createQueue(self, options=None):
        return OracleQueue()
