"""MatlibCuda -- the B200 driver service.

Implements Rindow\\Math\\Matrix\\Drivers\\Service (src/Drivers/Service.php:4-26) the way
AbstractMatlibService does (src/Drivers/AbstractMatlibService.php:43-75, 155-316): three factories
(buffer / blas / math) whose availability decides the service level, and level-selected
accessors.  With a visible B200 the level is LV_ADVANCED, which is what makes NDArray,
LinearAlgebra and MatrixOperator use the CUDA driver with zero edits.

Differences from the PHP class, by design:
  * the LV_BASIC slot (PhpBlas/PhpMath) lives in the reference; this package has NO CPU
    implementation, so asking for LV_BASIC raises LogicException instead of computing on the host;
  * there is no OpenCL side: openCL()/blasCL()/mathCL()/mathCLBlast()/createQueue() raise.
"""
from __future__ import annotations

from . import _capi
from .cuda_blas import CudaBlasFactory
from .cuda_buffer import CudaBufferFactory
from .cuda_math import CudaMathFactory
from .exceptions import LogicException


class Service:
    LV_UNAVAILABLE = 0
    LV_BASIC = 1        # Pure PHP
    LV_ADVANCED = 2     # Use external library
    LV_ACCELERATED = 3  # Use some accelerater like a GPU


_LEVEL_STRING = {0: "Unavailable", 1: "Basic", 2: "Advanced", 3: "Accelerated"}


class MatlibCuda(Service):
    name_ = "matlib_cuda"

    def __init__(self, bufferFactory=None, blasFactory=None, mathFactory=None, verbose: int | None = None,
                 gemm_mode: int | None = None):
        self._verbose = verbose or 0
        self.bufferFactory = bufferFactory or CudaBufferFactory()
        self.blasFactory = blasFactory or CudaBlasFactory()
        self.mathFactory = mathFactory or CudaMathFactory()
        self._level = None
        self._blas = self._math = self._buffer = None
        self._device_buffer = None
        if self.serviceLevel() >= Service.LV_ADVANCED:
            _capi.init()
            self._blas = self.blasFactory.Blas()
            self._math = self.mathFactory.Math()
            self._buffer = self.bufferFactory
            if gemm_mode is not None:
                self._blas.setGemmMode(gemm_mode)

    def _log(self, message: str) -> None:
        if self._verbose:
            print(message)

    def serviceLevel(self) -> int:                                  # AbstractMatlibService.php:155-223
        if self._level is not None:
            return self._level
        level = Service.LV_BASIC
        self._log("Current level Basic.")
        ok = True
        for f in (self.bufferFactory, self.blasFactory, self.mathFactory):
            if not f.isAvailable():
                self._log("%s is ** not available **." % type(f).__name__)
                ok = False
        if ok:
            level = Service.LV_ADVANCED
            self._log("Update current level to Advanced.")
            # the accelerated tier of AbstractMatlibService.php:193-218 (there: OpenCL + CLBlast factories): device-
            # resident buffers, queue, NDArrayCuda / LinearAlgebraCuda -- same library, so available together
            level = Service.LV_ACCELERATED
            self._log("Update current level to Accelerated.")
        self._level = level
        self._log("The service level was diagnosed as %s." % _LEVEL_STRING[level])
        return level

    def _require(self, level):
        level = Service.LV_ADVANCED if level is None else level
        if level == Service.LV_ACCELERATED:
            level = Service.LV_ADVANCED            # blas()/math() of the accelerated tier are the same driver objects
        if level == Service.LV_BASIC:
            raise LogicException("LV_BASIC is the reference's pure-PHP driver (PhpBlas/PhpMath); the B200 package "
                                 "ships no CPU implementation")
        if level != Service.LV_ADVANCED:
            raise LogicException("Unknown service level.")
        if self._level < Service.LV_ADVANCED:
            raise LogicException("No B200 device / librindow_b200.so available: the CUDA driver has no CPU fallback")

    def info(self) -> str:                                              # AbstractMatlibService.php:225-244
        modes = ["SEQUENTIAL", "THREAD", "OPENMP"]
        s = "Service Level   : %s\n" % _LEVEL_STRING[self.serviceLevel()]
        if self._blas is not None:
            s += "Buffer Factory  : %s\n" % type(self._buffer).__name__
            s += "BLAS Driver     : %s(%s)\n" % (type(self._blas).__name__, modes[self._blas.getParallel()])
            s += "Math Driver     : %s(%s)\n" % (type(self._math).__name__, modes[self._math.getParallel()])
        return s

    def name(self) -> str:
        return self.name_

    def blas(self, level=None):
        self._require(level)
        return self._blas

    def math(self, level=None):
        self._require(level)
        return self._math

    def buffer(self, level=None):
        self._require(level)
        if level == Service.LV_ACCELERATED:                              # AbstractMatlibService.php:316-318
            if self._device_buffer is None:
                from .ndarray_cuda import CudaDeviceBufferFactory
                self._device_buffer = CudaDeviceBufferFactory()
            return self._device_buffer
        return self._buffer

    def lapack(self, level=None):
        raise LogicException("LAPACK is out of scope of the B200 driver (SURVEY.md section 2)")

    def openCL(self):
        raise LogicException("openCL is not supported.")

    def blasCL(self, queue):
        raise LogicException("blasCL is not supported.")

    def mathCL(self, queue):
        raise LogicException("mathCL is not supported.")

    def mathCLBlast(self, queue):
        raise LogicException("mathCLBlast is not supported.")

    def createQueue(self, options=None):
        """The queue NDArrayCuda / LinearAlgebraCuda are constructed with (AbstractMatlibService::createQueue): the
        library stream."""
        self._require(None)
        from .ndarray_cuda import CudaQueue
        return CudaQueue(options)

    def blasCuda(self, queue):
        """Device-tier driver objects (the blasCL($queue) / mathCL($queue) slots of Service.php:16-20): the same
        CudaBlas / CudaMath -- they take device buffers positionally and are ordered on the queue's stream."""
        self._require(None)
        return self._blas

    def mathCuda(self, queue):
        self._require(None)
        return self._math
