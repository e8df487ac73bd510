"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/rindow_oracle.c header).

ctypes binding of oracle/liboracle.so, the C restatement of the reference's pure-PHP
driver (PhpBlas / PhpMath).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product package
`rindow_math_matrix_b200` never does (tests/test_no_oracle_in_product.py enforces it).

All functions follow the reference's *driver-level* signatures (flat buffer + offset +
ld), operate on float64 buffers (a PhpBuffer holds PHP doubles -- PhpBuffer.php:19) and
mutate their output argument in place, exactly like the PHP methods do.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

RowMajor, ColMajor = 101, 102
NoTrans, Trans, ConjTrans, ConjNoTrans = 111, 112, 113, 114


class OracleArgumentError(ValueError):
    """Stands for the reference's InvalidArgumentException."""


class OracleZeroDivide(RuntimeError):
    """Stands for RuntimeException('Zero divide in softmax.') (PhpMath.php:1663)."""


class OracleLabelRange(RuntimeError):
    """Stands for RuntimeException('Label number is out of bounds.') (PhpMath.php:1461)."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rindow_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, dbl, p, i32 = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        L.ro_gemm.argtypes = [i32, i32, i32, i64, i64, i64, dbl, p, i64, i64, i64,
                              p, i64, i64, i64, dbl, p, i64, i64, i64]
        L.ro_gemm_rows.argtypes = [i64, i64, i64, i64, dbl, p, i64, p, i64, dbl, p, i64]
        L.ro_gemm_rows_interleaved.argtypes = [i64, i64, i64, i64, dbl, p, i64, p, i64, dbl, p, i64, p]
        L.ro_gemm3.argtypes = [i64, i64, i64, i64, dbl, p, i64, p, i64, dbl, p, i64]
        L.ro_scal.argtypes = [i64, dbl, p, i64, i64, i64]
        L.ro_axpy.argtypes = [i64, dbl, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_copy.argtypes = [i64, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_dot.argtypes = [i64, p, i64, i64, i64, p, i64, i64, i64, p]
        L.ro_asum.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_nrm2.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_iamax.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_iamin.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_swap.argtypes = [i64, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_gemv.argtypes = [i32, i32, i64, i64, dbl, p, i64, i64, i64, p, i64, i64, i64, dbl, p, i64, i64, i64]
        L.ro_omatcopy.argtypes = [i32, i32, i64, i64, dbl, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_add.argtypes = [i32, i64, i64, dbl, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_multiply.argtypes = [i32, i64, i64, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_reduce_sum.argtypes = [i64, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_reduce_max.argtypes = [i32, i64, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_reduce_argmax.argtypes = [i64, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_softmax.argtypes = [i64, i64, p, i64, i64, i64]
        L.ro_im2col2d.argtypes = [i32, p, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64,
                                  i32, i32, i64, i64, i32, p, i64, i64, i64]
        L.ro_bandpart.argtypes = [i64, i64, i64, p, i64, i64, i64, i64]
        L.ro_gatherb.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, i64, i64, i64, i64, i64, p, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_gathernd.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, i64, i64, i64, i64, p, p, i64, i64, p, i64, p, i64, i64]
        L.ro_imagecopy.argtypes = [i64, i64, i64, p, i64, i64, p, i64, i64, ctypes.c_int, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ro_einsum.argtypes = [i64, i64, p, p, i64, i64, p, p, i64, i64, p, p, i64, i64]
        L.ro_topk.argtypes = [i64, i64, p, i64, i64, i64, ctypes.c_int, p, i64, i64, p, i64, i64]
        L.ro_cumsumb.argtypes = [i64, i64, i64, p, i64, i64, ctypes.c_int, ctypes.c_int, ctypes.c_int, p, i64, i64]
        L.ro_cumsum.argtypes = [i64, p, i64, i64, i64, ctypes.c_int, ctypes.c_int, p, i64, i64, i64]
        L.ro_searchsorted.argtypes = [i64, i64, p, i64, i64, i64, p, i64, i64, i64, ctypes.c_int, p, i64, i64, i64]
        L.ro_masking.argtypes = [i64, i64, i64, i64, dbl, i32, i32, p, i64, i64, p, i64, i64]
        L.ro_slice.argtypes = [i32, i32, i64, i64, i64, i64, p, i64, i64, i64, p, i64, i64, i64, i64, i64, i64, i64, i64, i64]
        L.ro_repeat.argtypes = [i64, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_im2col1d.argtypes = [i32, p, i64, i64, i64, i64, i64, i64, i64, i64, i32, i32, i64, i32, p, i64, i64, i64]
        L.ro_im2col3d.argtypes = [i32, p, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64,
                                  i32, i32, i64, i64, i64, i32, p, i64, i64, i64]
        L.ro_broadcast.argtypes = [i32, i32, i64, i64, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_unary.argtypes = [i32, i64, dbl, p, i64, i64, i64, dbl]
        L.ro_compare.argtypes = [i32, i64, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_sum.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_imax.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_imin.argtypes = [i64, p, i64, i64, i64, p]
        L.ro_onehot.argtypes = [i64, i64, dbl, p, i64, i64, i64, p, i64, i64, i64]
        L.ro_transpose.argtypes = [i32, p, p, p, i64, i64, p, i64, i64]
        L.ro_gather.argtypes = [i32, i32, i64, i64, i64, p, i64, i64, p, i64, i64, p, i64, i64]
        L.ro_reduce_gather.argtypes = [i32, i32, i64, i64, i64, p, i64, i64, p, i64, i64, p, i64, i64]
        for name in ("ro_topk", "ro_einsum", "ro_imagecopy", "ro_gathernd", "ro_gatherb", "ro_cumsumb", "ro_cumsum", "ro_searchsorted", "ro_bandpart", "ro_masking", "ro_slice", "ro_repeat", "ro_im2col1d", "ro_im2col3d", "ro_dot", "ro_asum", "ro_nrm2", "ro_iamax", "ro_iamin", "ro_swap", "ro_gather", "ro_reduce_gather", "ro_transpose", "ro_broadcast", "ro_unary", "ro_compare", "ro_sum", "ro_imax", "ro_imin", "ro_onehot", "ro_gemm", "ro_gemm_rows", "ro_gemm_rows_interleaved", "ro_gemm3", "ro_add", "ro_multiply",
                     "ro_scal", "ro_axpy", "ro_copy", "ro_gemv", "ro_omatcopy",
                     "ro_reduce_sum", "ro_reduce_max", "ro_reduce_argmax", "ro_softmax",
                     "ro_im2col2d", "ro_abi_version"):
            getattr(L, name).restype = ctypes.c_int
        _lib = L
    return _lib


def _buf(a: np.ndarray, dtype=np.float64) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.ndim == 1):
        raise TypeError("oracle buffers are flat C-contiguous %s arrays" % np.dtype(dtype).name)
    return a


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _check(rc: int) -> None:
    if rc == -1:
        raise OracleArgumentError("invalid argument (reference: InvalidArgumentException)")
    if rc == -2:
        raise OracleZeroDivide("Zero divide in softmax.")
    if rc == -3:
        raise OracleLabelRange("Label number is out of bounds.")
    if rc != 0:
        raise RuntimeError("oracle rc=%d" % rc)


def as_buffer(a) -> np.ndarray:
    """Widen any array to the flat float64 buffer the PHP driver would hold."""
    return np.ascontiguousarray(np.asarray(a), dtype=np.float64).reshape(-1).copy()


# PhpBlas::gemm, src/Drivers/MatlibPHP/PhpBlas.php:849-948
def gemm(order, transA, transB, m, n, k, alpha, A, offA, ldA, B, offB, ldB, beta, C, offC, ldC):
    _check(lib().ro_gemm(order, transA, transB, m, n, k, float(alpha),
                         _ptr(_buf(A)), A.size, offA, ldA, _ptr(_buf(B)), B.size, offB, ldB,
                         float(beta), _ptr(_buf(C)), C.size, offC, ldC))


def gemm_rows(row0, rows, n, k, alpha, A, ldA, B, ldB, beta, C, ldC):
    _check(lib().ro_gemm_rows(row0, rows, n, k, float(alpha), _ptr(_buf(A)), ldA,
                              _ptr(_buf(B)), ldB, float(beta), _ptr(_buf(C)), ldC))


def gemm_rows_fast(row0, rows, n, k, alpha, A, ldA, B, ldB, beta, C, ldC):
    """Bit-identical to gemm_rows (per-element k order kept), loops interchanged for speed."""
    acc = np.empty(n)
    _check(lib().ro_gemm_rows_interleaved(row0, rows, n, k, float(alpha), _ptr(_buf(A)), ldA,
                                          _ptr(_buf(B)), ldB, float(beta), _ptr(_buf(C)), ldC, _ptr(acc)))


# MatrixOperator::gemm3, src/MatrixOperator.php:510-529
def gemm3(M, N, K, L, alpha, A, offA, B, offB, beta, C, offC):
    _check(lib().ro_gemm3(M, N, K, L, float(alpha), _ptr(_buf(A)), offA, _ptr(_buf(B)), offB,
                          float(beta), _ptr(_buf(C)), offC))


# PhpBlas::scal / axpy / copy / gemv / omatcopy, PhpBlas.php:141-194, :350-364, :762-844, :1550-1612
def scal(n, alpha, X, offX, incX):
    _check(lib().ro_scal(n, float(alpha), _ptr(_buf(X)), X.size, offX, incX))


def axpy(n, alpha, X, offX, incX, Y, offY, incY):
    _check(lib().ro_axpy(n, float(alpha), _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, incY))


def copy(n, X, offX, incX, Y, offY, incY):
    _check(lib().ro_copy(n, _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, incY))


# PhpBlas::dot / asum / iamax / iamin / nrm2 / swap, PhpBlas.php:196-222, :266-348, :366-392, :744-760
def dot(n, X, offX, incX, Y, offY, incY):
    out = ctypes.c_double(0.0)
    _check(lib().ro_dot(n, _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, incY, ctypes.byref(out)))
    return out.value


def asum(n, X, offX, incX):
    out = ctypes.c_double(0.0)
    _check(lib().ro_asum(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def nrm2(n, X, offX, incX):
    out = ctypes.c_double(0.0)
    _check(lib().ro_nrm2(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def iamax(n, X, offX, incX):
    out = ctypes.c_int64(0)
    _check(lib().ro_iamax(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def iamin(n, X, offX, incX):
    out = ctypes.c_int64(0)
    _check(lib().ro_iamin(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def swap(n, X, offX, incX, Y, offY, incY):
    _check(lib().ro_swap(n, _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, incY))


def gemv(order, trans, m, n, alpha, A, offA, ldA, X, offX, incX, beta, Y, offY, incY):
    _check(lib().ro_gemv(order, trans, m, n, float(alpha), _ptr(_buf(A)), A.size, offA, ldA,
                         _ptr(_buf(X)), X.size, offX, incX, float(beta), _ptr(_buf(Y)), Y.size, offY, incY))


def omatcopy(order, trans, m, n, alpha, A, offA, ldA, B, offB, ldB):
    _check(lib().ro_omatcopy(order, trans, m, n, float(alpha), _ptr(_buf(A)), A.size, offA, ldA,
                             _ptr(_buf(B)), B.size, offB, ldB))


# PhpMath::add, PhpMath.php:570-606
def add(trans, m, n, alpha, X, offX, incX, A, offA, ldA):
    _check(lib().ro_add(int(bool(trans)), m, n, float(alpha), _ptr(_buf(X)), X.size, offX, incX,
                        _ptr(_buf(A)), A.size, offA, ldA))


# PhpMath::multiply, PhpMath.php:530-565
def multiply(trans, m, n, X, offX, incX, A, offA, ldA):
    _check(lib().ro_multiply(int(bool(trans)), m, n, _ptr(_buf(X)), X.size, offX, incX,
                             _ptr(_buf(A)), A.size, offA, ldA))


# broadcast family: PhpMath::maximum :274, minimum :318, greater :362, greaterEqual :404, less :446, lessEqual :488
# (A[m,n] against X[n]; no trans), pow :687 and duplicate :860 (trans like add).  op = BC_* below.
BC_MAXIMUM, BC_MINIMUM, BC_GREATER, BC_GREATER_EQUAL, BC_LESS, BC_LESS_EQUAL, BC_POW, BC_DUPLICATE = range(8)


def broadcast(op, trans, m, n, A, offA, ldA, X, offX, incX):
    _check(lib().ro_broadcast(op, int(bool(trans)), m, n, _ptr(_buf(A)), A.size, offA, ldA,
                              _ptr(_buf(X)), X.size, offX, incX))


# unary family: PhpMath::increment :221, reciprocal :244, square :611, sqrt :634, rsqrt :657, exp :727, log :749,
# tanh :772, sin :794, cos :816, tan :838, not :1528, nan2num :2830, isnan :2853.  op = UN_* below.
(UN_INCREMENT, UN_RECIPROCAL, UN_SQUARE, UN_SQRT, UN_RSQRT, UN_EXP, UN_LOG, UN_TANH, UN_SIN, UN_COS, UN_TAN,
 UN_NOT, UN_NAN2NUM, UN_ISNAN) = range(14)


def unary(op, n, X, offX, incX, alpha=1.0, beta=0.0):
    _check(lib().ro_unary(op, n, float(alpha), _ptr(_buf(X)), X.size, offX, incX, float(beta)))


# PhpMath::equal :1470 / notEqual :1499
def compare(not_equal, n, X, offX, incX, Y, offY, incY):
    _check(lib().ro_compare(int(bool(not_equal)), n, _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, incY))


# PhpMath::sum :150, imax :171, imin :196
def sum(n, X, offX, incX):
    out = ctypes.c_double(0.0)
    _check(lib().ro_sum(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def imax(n, X, offX, incX):
    out = ctypes.c_int64(0)
    _check(lib().ro_imax(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


def imin(n, X, offX, incX):
    out = ctypes.c_int64(0)
    _check(lib().ro_imin(n, _ptr(_buf(X)), X.size, offX, incX, ctypes.byref(out)))
    return out.value


# PhpMath::updateAddOnehot :1438
def onehot(m, n, a, X, offX, incX, Y, offY, ldY):
    _check(lib().ro_onehot(m, n, float(a), _ptr(_buf(X)), X.size, offX, incX, _ptr(_buf(Y)), Y.size, offY, ldY))


# PhpMath::astype :1710-1778 on the PHP value model (numpy restatement: integer / byte work).  X is the float64 or
# int64 flat buffer; returns what the PHP loop stores: bool -> 1/0, float target -> (float)x, integer target ->
# (int)x truncated toward zero, unsigned 8/16/32 targets masked to their low bits.
def astype(n, to_kind, mask, X, offX, incX):
    x = X[offX:offX + (n - 1) * incX + 1:incX]
    if to_kind == "bool":
        return (x != 0).astype(np.int64)
    if to_kind == "float":
        return x.astype(np.float64)
    v = np.trunc(x).astype(np.int64) if x.dtype.kind == "f" else x.astype(np.int64)
    return (v & mask) if mask else v


# PhpMath::gather :1063 (incl. scatterAdd2 :1362) and reduceGather :1135; X is the float64 widening of the indices
def gather(reverse, add_mode, n, k, numClass, X, offX, A, offA, B, offB):
    _check(lib().ro_gather(int(bool(reverse)), int(bool(add_mode)), n, k, numClass, _ptr(_buf(X)), X.size, offX,
                           _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB))


def reduce_gather(reverse, add_mode, m, n, numClass, X, offX, A, offA, B, offB):
    _check(lib().ro_reduce_gather(int(bool(reverse)), int(bool(add_mode)), m, n, numClass, _ptr(_buf(X)), X.size, offX,
                                  _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB))


# PhpMath::transpose, PhpMath.php:2959-3025: B's axis d is A's axis perm[d]
def transpose(shape, perm, A, offA, B, offB):
    sh = np.ascontiguousarray(shape, dtype=np.int64)
    pm = np.ascontiguousarray(perm, dtype=np.int64)
    if sh.size != pm.size:
        raise OracleArgumentError("unmatch sourceshape and perm")
    _check(lib().ro_transpose(int(sh.size), _ptr(sh), _ptr(pm), _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB))


# PhpMath::reduceSum, PhpMath.php:1552-1579
def reduce_sum(m, n, k, A, offA, B, offB):
    _check(lib().ro_reduce_sum(m, n, k, _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB))


# PhpMath::reduceMax, PhpMath.php:1584-1611 (sticky=True: accelerated-path NaN rule)
def reduce_max(m, n, k, A, offA, B, offB, sticky=False):
    _check(lib().ro_reduce_max(int(bool(sticky)), m, n, k, _ptr(_buf(A)), A.size, offA,
                               _ptr(_buf(B)), B.size, offB))


# PhpMath::reduceArgMax, PhpMath.php:1616-1643 (B is an int64 buffer: PHP ints)
def reduce_argmax(m, n, k, A, offA, B, offB):
    _check(lib().ro_reduce_argmax(m, n, k, _ptr(_buf(A)), A.size, offA,
                                  _ptr(_buf(B, np.int64)), B.size, offB))


# PhpMath::softmax, PhpMath.php:1645-1670
def softmax(m, n, A, offA, ldA):
    _check(lib().ro_softmax(m, n, _ptr(_buf(A)), A.size, offA, ldA))


# PhpMath::im2col2d, PhpMath.php:2158-2314
def im2col2d(reverse, images, images_offset, images_size, batches, im_h, im_w, channels,
             filter_h, filter_w, stride_h, stride_w, padding, channels_first,
             dilation_h, dilation_w, cols_channels_first, cols, cols_offset, cols_size):
    _check(lib().ro_im2col2d(int(bool(reverse)), _ptr(_buf(images)), images.size, images_offset,
                             images_size, batches, im_h, im_w, channels, filter_h, filter_w,
                             stride_h, stride_w, int(bool(padding)), int(bool(channels_first)),
                             dilation_h, dilation_w, int(bool(cols_channels_first)),
                             _ptr(_buf(cols)), cols.size, cols_offset, cols_size))


# PhpMath::gatherb :1209-1260
def gatherb(reverse, addMode, f32, batches, m, n, k, len_, numClass, A, offA, X, offX, B, offB):
    _check(lib().ro_gatherb(int(bool(reverse)), int(bool(addMode)), int(bool(f32)), batches, m, n, k, len_, numClass,
                            _ptr(_buf(A)), A.size, offA, _ptr(_buf(X)), X.size, offX, _ptr(_buf(B)), B.size, offB))


# PhpMath::gathernd :1271-1325 (paramShape: host sequence; X read from element 0 as the reference does)
def gathernd(reverse, addMode, f32, m, n, k, depth, paramShape, A, offA, X, B, offB):
    shape = (ctypes.c_int64 * depth)(*[int(v) for v in paramShape[:depth]])
    _check(lib().ro_gathernd(int(bool(reverse)), int(bool(addMode)), int(bool(f32)), m, n, k, depth,
                             ctypes.cast(shape, ctypes.c_void_p), _ptr(_buf(A)), A.size, offA, _ptr(_buf(X)), X.size,
                             _ptr(_buf(B)), B.size, offB))


# PhpMath::imagecopy :2877-2957
def imagecopy(height, width, channels, A, offA, B, offB, channelsFirst, heightShift, widthShift, verticalFlip, horizontalFlip,
              rgbFlip):
    _check(lib().ro_imagecopy(height, width, channels, _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB,
                              int(bool(channelsFirst)), int(heightShift), int(widthShift), int(bool(verticalFlip)),
                              int(bool(horizontalFlip)), int(bool(rgbFlip))))


# PhpMath::einsum :3321-3357 (einsum4p1 :3360-3440 = depth 5, ndimC 4); sizes / strides are host sequences
def einsum(sizeOfIndices, A, offA, ldA, B, offB, ldB, C, offC, ndimC):
    depth = len(sizeOfIndices)
    arr = lambda v: ctypes.cast((ctypes.c_int64 * depth)(*[int(x) for x in v]), ctypes.c_void_p)
    s, la, lb = arr(sizeOfIndices), arr(ldA), arr(ldB)
    _check(lib().ro_einsum(depth, ndimC, s, _ptr(_buf(A)), A.size, offA, la, _ptr(_buf(B)), B.size, offB, lb,
                           _ptr(_buf(C)), C.size, offC))


# PhpMath::topk :933-1058
def topk(m, n, input, offInput, k, sorted, values, offValues, indices, offIndices):
    _check(lib().ro_topk(m, n, _ptr(_buf(input)), input.size, offInput, k, int(bool(sorted)), _ptr(_buf(values)), values.size,
                         offValues, _ptr(_buf(indices)), indices.size, offIndices))


# PhpMath::cumsumb :1861-1904, cumsum :1822-1859, searchsorted :1780-1820
def cumsumb(m, n, k, A, offA, exclusive, reverse, f32, B, offB):
    _check(lib().ro_cumsumb(m, n, k, _ptr(_buf(A)), A.size, offA, int(bool(exclusive)), int(bool(reverse)), int(bool(f32)),
                            _ptr(_buf(B)), B.size, offB))


def cumsum(n, X, offX, incX, exclusive, reverse, Y, offY, incY):
    _check(lib().ro_cumsum(n, _ptr(_buf(X)), X.size, offX, incX, int(bool(exclusive)), int(bool(reverse)),
                           _ptr(_buf(Y)), Y.size, offY, incY))


def searchsorted(m, n, A, offA, ldA, X, offX, incX, right, Y, offY, incY):
    _check(lib().ro_searchsorted(m, n, _ptr(_buf(A)), A.size, offA, ldA, _ptr(_buf(X)), X.size, offX, incX,
                                 int(bool(right)), _ptr(_buf(Y)), Y.size, offY, incY))


def bandpart(m, n, k, A, offA, lower, upper):
    _check(lib().ro_bandpart(m, n, k, _ptr(_buf(A)), A.size, offA, lower, upper))


# PhpMath::masking :3229-3270 (kind: 0 float, 1 integer, 2 bool data)
def masking(m, n, k, len_, fill, mode, kind, X, offX, A, offA):
    _check(lib().ro_masking(m, n, k, len_, float(fill), int(mode), int(kind), _ptr(_buf(X)), X.size, offX,
                            _ptr(_buf(A)), A.size, offA))


# PhpMath::slice :2691-2776, repeat :1400-1422
def slice(reverse, addMode, m, n, k, size, A, offA, incA, Y, offY, incY, start0, size0, start1, size1, start2, size2):
    _check(lib().ro_slice(int(bool(reverse)), int(bool(addMode)), m, n, k, size, _ptr(_buf(A)), A.size, offA, incA,
                          _ptr(_buf(Y)), Y.size, offY, incY, start0, size0, start1, size1, start2, size2))


def repeat(m, k, repeats, A, offA, B, offB):
    _check(lib().ro_repeat(m, k, repeats, _ptr(_buf(A)), A.size, offA, _ptr(_buf(B)), B.size, offB))


# PhpMath::im2col1d :1967-2088, im2col3d :2396-2600
def im2col1d(reverse, images, images_offset, images_size, batches, im_w, channels, filter_w, stride_w, padding,
             channels_first, dilation_w, cols_channels_first, cols, cols_offset, cols_size):
    _check(lib().ro_im2col1d(int(bool(reverse)), _ptr(_buf(images)), images.size, images_offset, images_size,
                             batches, im_w, channels, filter_w, stride_w, int(bool(padding)), int(bool(channels_first)),
                             dilation_w, int(bool(cols_channels_first)), _ptr(_buf(cols)), cols.size, cols_offset, cols_size))


def im2col3d(reverse, images, images_offset, images_size, batches, im_d, im_h, im_w, channels,
             filter_d, filter_h, filter_w, stride_d, stride_h, stride_w, padding, channels_first,
             dilation_d, dilation_h, dilation_w, cols_channels_first, cols, cols_offset, cols_size):
    _check(lib().ro_im2col3d(int(bool(reverse)), _ptr(_buf(images)), images.size, images_offset, images_size,
                             batches, im_d, im_h, im_w, channels, filter_d, filter_h, filter_w,
                             stride_d, stride_h, stride_w, int(bool(padding)), int(bool(channels_first)),
                             dilation_d, dilation_h, dilation_w, int(bool(cols_channels_first)),
                             _ptr(_buf(cols)), cols.size, cols_offset, cols_size))


def im2col2d_out_shape(batches, im_h, im_w, channels, filter_h, filter_w, stride_h, stride_w,
                       padding, dilation_h, dilation_w, cols_channels_first):
    """Shape LinearAlgebra::im2col2d allocates for cols (src/LinearAlgebra.php:3645-3670)."""
    if padding:
        out_h, out_w = im_h, im_w
    else:
        out_h = (im_h - (filter_h - 1) * dilation_h - 1) // stride_h + 1
        out_w = (im_w - (filter_w - 1) * dilation_w - 1) // stride_w + 1
    if cols_channels_first:
        return (batches, out_h, out_w, channels, filter_h, filter_w)
    return (batches, out_h, out_w, filter_h, filter_w, channels)
