"""CudaBlas -- the `blas()` object of the B200 service.

Same positional signatures, validation order and exception strings as PhpBlas
(src/Drivers/MatlibPHP/PhpBlas.php) + the Utils trait (src/Drivers/MatlibPHP/Utils.php:30-65);
the arithmetic runs in librindow_b200.so.  On the device: gemm (the hot path, rb200_sgemm / rb200_dgemm)
and the section 8(f) widening -- scal, axpy, copy, gemv, omatcopy.  The remaining BLAS methods raise: the
PHP shim delegates those to PhpBlas on the host mirror, this package has no CPU path.
"""
from __future__ import annotations

import ctypes

from . import _capi, dtypes
from .cuda_buffer import CudaBuffer
from .exceptions import InvalidArgumentException, LogicException, RuntimeException


class BLAS:
    """Interop\\Polite\\Math\\Matrix\\BLAS constants (CBLAS values; LinearAlgebra.php:16-17)."""
    RowMajor = 101
    ColMajor = 102
    NoTrans = 111
    Trans = 112
    ConjTrans = 113
    ConjNoTrans = 114


def assert_shape_parameter(name: str, n: int) -> None:                     # Utils.php:30-36
    if n < 1:
        raise InvalidArgumentException("Argument %s must be greater than 0." % name)


def assert_vector_buffer_spec(name: str, buffer, n: int, offset: int, inc: int) -> None:   # Utils.php:38-50
    if offset < 0:
        raise InvalidArgumentException("Argument offset%s must be greater than equals 0." % name)
    if inc < 1:
        raise InvalidArgumentException("Argument inc%s must be greater than 0." % name)
    if offset + (n - 1) * inc >= len(buffer):
        raise InvalidArgumentException("Vector specification too large for buffer%s." % name)


def assert_matrix_buffer_spec(name: str, buffer, m: int, n: int, offset: int, ld: int) -> None:  # Utils.php:52-65
    if offset < 0:
        raise InvalidArgumentException("Argument offset%s must be greater than equals 0." % name)
    if ld < 1:
        raise InvalidArgumentException("Argument ld%s must be greater than 0." % name)
    if offset + (m - 1) * ld + (n - 1) >= len(buffer):
        raise InvalidArgumentException("Matrix specification too large for buffer%s." % name)


def code_to_trans(trans: int):                                             # PhpBlas.php:104-123
    if trans == BLAS.NoTrans:
        return False, False
    if trans == BLAS.Trans:
        return True, False
    if trans == BLAS.ConjTrans:
        return True, True
    if trans == BLAS.ConjNoTrans:
        return False, True
    raise InvalidArgumentException("Unknown Tranpose Code: %s" % trans)


class CudaBlas:
    # rows from which the host-operand GEMM pipelines its transfers (below that one copy + one launch wins)
    STREAMED_MIN_ROWS = 2048

    def __init__(self, gemm_mode: int = _capi.GEMM_AUTO):
        self._lib = _capi.load()
        self.gemm_mode = gemm_mode
        # None: bring C back eagerly iff the caller already indexes C from the host (its mirror exists);
        # True / False: always / never.  Only matters for buffers filled or read from the host side.
        self.host_result_prefetch = None

    # ---- meta (PhpBlas.php:49-90) -----------------------------------------------------------
    def getNumThreads(self) -> int:
        return 1

    def getNumProcs(self) -> int:
        return 1

    def getConfig(self) -> str:
        return "CudaBlas sm_100a"

    def getCorename(self) -> str:
        return "B200"

    def getParallel(self) -> int:
        return 0        # 0 = SEQUENTIAL from the host's point of view (AbstractMatlibService.php:225-244)

    def hasIamin(self) -> bool:
        return True

    def hasOmatcopy(self) -> bool:
        return True     # must be true so transposes never go through GEMM

    def setGemmMode(self, mode: int) -> None:
        self.gemm_mode = mode

    def setHostResultPrefetch(self, flag) -> None:
        self.host_result_prefetch = flag

    def _gemm_streamed(self, transA, transB, m, n, k, alpha, A, offsetA, ldA, B, offsetB, ldB, beta,
                       C, offsetC, ldC, rowsA, colsA, rowsB, colsB) -> bool:
        """Host-operand form (rb200_sgemm_streamed): when an operand's current copy is in its pinned mirror
        and/or the caller reads C from the host, the transfers are pipelined with the multiply instead of
        flush -> kernel -> fetch.  Returns False when the plain path should run."""
        if m < self.STREAMED_MIN_ROWS or A is C or B is C:
            return False

        def whole(buf, off, rows, cols, ld):        # the view is the entire buffer, rows dense
            return off == 0 and ld == cols and len(buf) == rows * cols

        upA = A.host_pending() and whole(A, offsetA, rowsA, colsA, ldA)
        upB = B.host_pending() and whole(B, offsetB, rowsB, colsB, ldB) and B is not A
        if B is A and upA:
            upB = False                      # one buffer: uploaded once, whole, as "B"
            upA = False
            A.flush()
        want_c = self.host_result_prefetch
        if want_c is None:
            want_c = C.has_mirror()
        downC = bool(want_c) and (whole(C, offsetC, m, n, ldC) or C.mirror_current())
        if not (upA or upB or downC):
            return False
        if A.host_pending() and not upA:
            A.flush()
        if B.host_pending() and not upB:
            B.flush()
        c_host_current = 0
        if C.host_pending():
            if whole(C, offsetC, m, n, ldC) and downC:
                c_host_current = 1           # uploaded stripe by stripe (only read when beta != 0)
                C.streamed_uploaded()
            else:
                C.flush()
        if downC:
            C._ensure_mirror()               # allocate only: the call itself fills it
        dA, hA = A.raw_pointers()
        dB, hB = B.raw_pointers()
        dC, hC = C.raw_pointers()
        _capi.check(self._lib.rb200_sgemm_streamed(
            BLAS.RowMajor, transA, transB, m, n, k, alpha, dA, hA if upA else None, offsetA, ldA,
            dB, hB if upB else None, offsetB, ldB, beta, dC, hC if downC else None, c_host_current,
            offsetC, ldC, self.gemm_mode))
        if upA:
            A.streamed_uploaded()
        if upB:
            B.streamed_uploaded()
        C.streamed_result(downC)
        return True

    # ---- gemm (PhpBlas.php:849-948) -------------------------------------------------------------
    def gemm(self, order, transA, transB, m, n, k, alpha, A, offsetA, ldA, B, offsetB, ldB,
             beta, C, offsetC, ldC) -> None:
        if order == BLAS.ColMajor:
            m, n = n, m
        elif order != BLAS.RowMajor:
            raise InvalidArgumentException("Invalid Order type")
        tA, _ = code_to_trans(transA)
        tB, _ = code_to_trans(transB)
        assert_shape_parameter("m", m)
        assert_shape_parameter("n", n)
        assert_shape_parameter("k", k)
        rowsA, colsA = (m, k) if not tA else (k, m)
        rowsB, colsB = (k, n) if not tB else (n, k)
        assert_matrix_buffer_spec("A", A, rowsA, colsA, offsetA, ldA)
        assert_matrix_buffer_spec("B", B, rowsB, colsB, offsetB, ldB)
        assert_matrix_buffer_spec("C", C, m, n, offsetC, ldC)
        for name, buf in (("A", A), ("B", B), ("C", C)):
            if not isinstance(buf, CudaBuffer):
                raise InvalidArgumentException("buffer%s is not a CudaBuffer" % name)
        if not (A.dtype() == B.dtype() == C.dtype()):
            raise InvalidArgumentException("Unmatch data type for A, B and C")
        try:
            alpha = float(alpha)
        except (TypeError, ValueError):
            raise RuntimeException("alpha is not float")                   # PhpBlas.php:133-139
        try:
            beta = float(beta)
        except (TypeError, ValueError):
            raise RuntimeException("beta is not float")
        # m,n were already swapped for ColMajor: the ABI is called RowMajor
        if A.dtype() == dtypes.float32:
            if self._gemm_streamed(transA, transB, m, n, k, alpha, A, offsetA, ldA, B, offsetB, ldB, beta,
                                   C, offsetC, ldC, rowsA, colsA, rowsB, colsB):
                return
            a_ptr, b_ptr = A.device_ptr(), B.device_ptr()
            c_ptr = C.device_ptr_rw()
            _capi.check(self._lib.rb200_sgemm(BLAS.RowMajor, transA, transB, m, n, k, alpha, a_ptr, offsetA, ldA,
                                              b_ptr, offsetB, ldB, beta, c_ptr, offsetC, ldC, self.gemm_mode))
        elif A.dtype() == dtypes.float64:
            a_ptr, b_ptr = A.device_ptr(), B.device_ptr()
            c_ptr = C.device_ptr_rw()
            _capi.check(self._lib.rb200_dgemm(BLAS.RowMajor, transA, transB, m, n, k, alpha, a_ptr, offsetA, ldA,
                                              b_ptr, offsetB, ldB, beta, c_ptr, offsetC, ldC))
        else:
            raise LogicException("CudaBlas::gemm supports float32/float64; the reference routes integer "
                                 "dtypes to the LV_BASIC PHP driver (MatrixOperator.php:225-231)")

    # ---- multi-GPU: row-sharded gemm whose C stripe also lands in the peers' full C ----
    def gemmAllgather(self, transA, transB, m, n, k, alpha, A, offsetA, ldA, B, offsetB, ldB,
                      beta, C, offsetC, ldC, peer_ptrs, multicast_ptr: int | None = None) -> None:
        """PhpBlas::gemm semantics (row-major) for this rank's stripe: C[offsetC ...] is the stripe inside this
        rank's full C; peer_ptrs are the other ranks' full-C allocations mapped here (sharding.PeerRows).  The
        tensor-core kernel stores every finished tile to the peers from its epilogue (rb200_sgemm_allgather);
        the caller orders the ranks afterwards (PeerRows.arrive).  With `multicast_ptr` (the NVLS multicast alias of
        the symmetric allocation C lives in, sharding.SymmetricRows) the stripe leaves this GPU once: the epilogue
        stores with multimem.st and the NVSwitch replicates (rb200_sgemm_allgather_mc)."""
        tA, _ = code_to_trans(transA)
        tB, _ = code_to_trans(transB)
        assert_shape_parameter("m", m)
        assert_shape_parameter("n", n)
        assert_shape_parameter("k", k)
        rowsA, colsA = (m, k) if not tA else (k, m)
        rowsB, colsB = (k, n) if not tB else (n, k)
        assert_matrix_buffer_spec("A", A, rowsA, colsA, offsetA, ldA)
        assert_matrix_buffer_spec("B", B, rowsB, colsB, offsetB, ldB)
        assert_matrix_buffer_spec("C", C, m, n, offsetC, ldC)
        if not (A.dtype() == B.dtype() == C.dtype() == dtypes.float32):
            raise LogicException("CudaBlas::gemmAllgather supports float32")
        if multicast_ptr:
            _capi.check(self._lib.rb200_sgemm_allgather_mc(
                BLAS.RowMajor, transA, transB, m, n, k, float(alpha), A.device_ptr(), offsetA, ldA,
                B.device_ptr(), offsetB, ldB, float(beta), C.device_ptr_rw(), offsetC, ldC, self.gemm_mode,
                multicast_ptr))
            return
        peers = (ctypes.c_void_p * max(1, len(peer_ptrs)))(*peer_ptrs)
        _capi.check(self._lib.rb200_sgemm_allgather(
            BLAS.RowMajor, transA, transB, m, n, k, float(alpha), A.device_ptr(), offsetA, ldA,
            B.device_ptr(), offsetB, ldB, float(beta), C.device_ptr_rw(), offsetC, ldC, self.gemm_mode,
            len(peer_ptrs), peers))

    # ---- helpers ----------------------------------------------------------------------------------------
    @staticmethod
    def _clean_float(value, name):                                           # PhpBlas.php:133-139
        try:
            return float(value)
        except (TypeError, ValueError):
            raise RuntimeException("%s is not float" % name)

    @staticmethod
    def _device_buffers(**bufs):
        for name, buf in bufs.items():
            if not isinstance(buf, CudaBuffer):
                raise InvalidArgumentException("buffer%s is not a CudaBuffer" % name)

    @staticmethod
    def _float_dtype(buf, op):
        if buf.dtype() not in (dtypes.float32, dtypes.float64):
            raise LogicException("CudaBlas::%s supports float32/float64; the reference routes integer "
                                 "dtypes to the LV_BASIC PHP driver (MatrixOperator.php:225-231)" % op)
        return buf.dtype()

    # ---- scal (PhpBlas.php:141-161) ---------------------------------------------------------------------
    def scal(self, n, alpha, X, offsetX, incX) -> None:
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        self._device_buffers(X=X)
        alpha = self._clean_float(alpha, "alpha")
        _capi.check(self._lib.rb200_scal(self._float_dtype(X, "scal"), n, alpha, X.device_ptr_rw(), offsetX, incX))

    # ---- axpy (PhpBlas.php:165-194) ---------------------------------------------------------------------
    def axpy(self, n, alpha, X, offsetX, incX, Y, offsetY, incY) -> None:
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        assert_vector_buffer_spec("Y", Y, n, offsetY, incY)
        self._device_buffers(X=X, Y=Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        alpha = self._clean_float(alpha, "alpha")
        _capi.check(self._lib.rb200_axpy(self._float_dtype(X, "axpy"), n, alpha, X.device_ptr(), offsetX, incX,
                                         Y.device_ptr_rw(), offsetY, incY))

    # ---- copy (PhpBlas.php:350-364) ---------------------------------------------------------------------
    def copy(self, n, X, offsetX, incX, Y, offsetY, incY) -> None:
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        assert_vector_buffer_spec("Y", Y, n, offsetY, incY)
        self._device_buffers(X=X, Y=Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        dt = X.dtype() if X.dtype() in dtypes.COMPLEX_TYPES else self._float_dtype(X, "copy")   # complex: data movement only
        _capi.check(self._lib.rb200_copy(dt, n, X.device_ptr(), offsetX, incX,
                                         Y.device_ptr_rw(), offsetY, incY))

    # ---- gemv (PhpBlas.php:762-844) ---------------------------------------------------------------------
    def gemv(self, order, trans, m, n, alpha, A, offsetA, ldA, X, offsetX, incX, beta, Y, offsetY, incY) -> None:
        if order == BLAS.ColMajor:
            m, n = n, m
        elif order != BLAS.RowMajor:
            raise InvalidArgumentException("Invalid Order type")
        t, _ = code_to_trans(trans)
        rows, cols = (m, n) if not t else (n, m)
        assert_shape_parameter("m", m)
        assert_shape_parameter("n", n)
        assert_matrix_buffer_spec("A", A, m, n, offsetA, ldA)
        assert_vector_buffer_spec("X", X, cols, offsetX, incX)
        assert_vector_buffer_spec("Y", Y, rows, offsetY, incY)
        self._device_buffers(A=A, X=X, Y=Y)
        if not (A.dtype() == X.dtype() == Y.dtype()):
            raise InvalidArgumentException("Unmatch data type for A, X and Y")
        alpha = self._clean_float(alpha, "alpha")
        beta = self._clean_float(beta, "beta")
        _capi.check(self._lib.rb200_gemv(self._float_dtype(A, "gemv"), BLAS.RowMajor, trans, m, n, alpha,
                                         A.device_ptr(), offsetA, ldA, X.device_ptr(), offsetX, incX, beta,
                                         Y.device_ptr_rw(), offsetY, incY))

    # ---- omatcopy (PhpBlas.php:1550-1612) ---------------------------------------------------------------
    def omatcopy(self, order, trans, m, n, alpha, A, offsetA, ldA, B, offsetB, ldB) -> None:
        if order == BLAS.ColMajor:
            m, n = n, m
        elif order != BLAS.RowMajor:
            raise InvalidArgumentException("Invalid Order type")
        t, _ = code_to_trans(trans)
        assert_shape_parameter("m", m)
        assert_shape_parameter("n", n)
        rows, cols = (m, n) if not t else (n, m)
        assert_matrix_buffer_spec("A", A, m, n, offsetA, ldA)
        assert_matrix_buffer_spec("B", B, rows, cols, offsetB, ldB)
        self._device_buffers(A=A, B=B)
        if A.dtype() != B.dtype():
            raise InvalidArgumentException("Unmatch data type for A and B")
        alpha = self._clean_float(alpha, "alpha")
        _capi.check(self._lib.rb200_omatcopy(self._float_dtype(A, "omatcopy"), BLAS.RowMajor, trans, m, n, alpha,
                                             A.device_ptr(), offsetA, ldA, B.device_ptr_rw(), offsetB, ldB))

    # ---- level-1 reductions (PhpBlas.php:196-222 dot, :266-286 asum, :288-348 iamax / iamin, :366-392 nrm2) --------
    # every real dtype: PHP's abs / * / + work on ints too; the value is on the host when the call returns
    def _scalar(self, fn, ctype, n, X, offsetX, incX):
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        self._device_buffers(X=X)
        out = ctype(0)
        _capi.check(fn(X.dtype(), n, X.device_ptr(), offsetX, incX, ctypes.byref(out)))
        return out.value

    def asum(self, n, X, offsetX, incX) -> float:
        return self._scalar(self._lib.rb200_asum, ctypes.c_double, n, X, offsetX, incX)

    def nrm2(self, n, X, offsetX, incX) -> float:
        return self._scalar(self._lib.rb200_nrm2, ctypes.c_double, n, X, offsetX, incX)

    def iamax(self, n, X, offsetX, incX) -> int:
        return self._scalar(self._lib.rb200_iamax, ctypes.c_int64, n, X, offsetX, incX)

    def iamin(self, n, X, offsetX, incX) -> int:
        return self._scalar(self._lib.rb200_iamin, ctypes.c_int64, n, X, offsetX, incX)

    def dot(self, n, X, offsetX, incX, Y, offsetY, incY) -> float:
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        assert_vector_buffer_spec("Y", Y, n, offsetY, incY)
        self._device_buffers(X=X, Y=Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        out = ctypes.c_double(0.0)
        _capi.check(self._lib.rb200_dot(X.dtype(), n, X.device_ptr(), offsetX, incX, Y.device_ptr(), offsetY, incY,
                                        ctypes.byref(out)))
        return out.value

    # ---- swap (PhpBlas.php:744-760) ---------------------------------------------------------------------
    def swap(self, n, X, offsetX, incX, Y, offsetY, incY) -> None:
        assert_shape_parameter("n", n)
        assert_vector_buffer_spec("X", X, n, offsetX, incX)
        assert_vector_buffer_spec("Y", Y, n, offsetY, incY)
        self._device_buffers(X=X, Y=Y)
        if X.dtype() != Y.dtype():
            raise InvalidArgumentException("Unmatch data type for X and Y")
        dt = X.dtype() if X.dtype() in dtypes.COMPLEX_TYPES else self._float_dtype(X, "swap")
        _capi.check(self._lib.rb200_swap(dt, n, X.device_ptr_rw(), offsetX, incX,
                                         Y.device_ptr_rw(), offsetY, incY))

    def __getattr__(self, name):
        if name in ("dotu", "dotc", "rotg", "rot", "rotm", "rotmg", "symm", "syrk", "syr2k", "trmm", "trsm"):
            def _not_on_device(*a, **k):
                raise LogicException("CudaBlas::%s is not on the B200 hot path (SURVEY.md section 8f); the PHP "
                                     "shim delegates it to PhpBlas, this package has no CPU fallback" % name)
            return _not_on_device
        raise AttributeError(name)


class CudaBlasFactory:
    """`openblasFactory` slot of AbstractMatlibService (:43-75): isAvailable(), Blas(), Lapack()."""

    def isAvailable(self) -> bool:
        return _capi.device_available()

    def Blas(self) -> CudaBlas:
        return CudaBlas()

    def Lapack(self):
        raise LogicException("LAPACK is out of scope (SURVEY.md section 2); the PHP shim returns PhpLapack")
