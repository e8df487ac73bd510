"""LinearAlgebraCuda -- the accelerated LA object over device-resident arrays (NDArrayCuda).

The B200 counterpart of Rindow\\Math\\Matrix\\LinearAlgebraCL (src/LinearAlgebraCL.php): constructed from a queue and
the service (:75-93), `accelerated()` is true (:234-237), `array()` uploads a host NDArray / nested array into a
device array and passes device arrays through (:321-352), `toNDArray()` downloads (:354-360), `alloc()` creates
device arrays (:441-460), `allocHost()` / `zerosHost()` stay on the host (:793-814), `blocking()` / `finish()` control
whether a call returns before its kernels have run (:145-152, 250-253), scalar-valued calls return rank-0 device
arrays unless `scalarNumeric(true)` (:220-227, 2526-2529), `scalar()` reads one back (:362-368).  (The PHP class,
php/MatlibCuda/LinearAlgebraCuda.php, extends LinearAlgebra whose `sum(): float` etc. cannot be widened by a subclass,
so it always behaves as scalarNumeric(true); `la.scalar(la.sum(x))` is the form that works on both.)

Every operation itself is inherited from the driver-agnostic LinearAlgebra mirror: the shape algebra is the same,
and the driver objects (CudaBlas / CudaMath) take device buffers positionally exactly as the host-facing ones do --
the OpenCL variants' trailing ($events, $waitEvents) have no counterpart because the driver is stream-ordered
(one library stream; rb200_event_* where a completion marker is needed).  MatrixOperator::laAccelerated('cuda')
returns this class (src/MatrixOperator.php:1566-1579 knows only 'clblast'; tests override createLinearAlgebraCL the
same way, tests/.../LinearAlgebraCLTest.php:29-37).
"""
from __future__ import annotations

import numpy as np

from . import dtypes
from .exceptions import InvalidArgumentException
from .linear_algebra import LinearAlgebra
from .ndarray import NDArray
from .ndarray_cuda import CL_MEM_COPY_HOST_PTR, CL_MEM_READ_WRITE, CudaEventList, CudaQueue, NDArrayCuda

_SCALAR_CALLS = ("sum", "asum", "nrm2", "dot", "amax", "amin", "max", "min", "imax", "imin", "iamax", "iamin")
_INDEX_CALLS = ("imax", "imin", "iamax", "iamin")


class LinearAlgebraCuda(LinearAlgebra):
    def __init__(self, queue: CudaQueue, service, defaultFloatType: int | None = None):
        super().__init__(service, defaultFloatType)
        self.queue = queue
        self.context = queue.getContext()
        self._blocking = False
        self._scalar_numeric = False

    # ---- identity / control (LinearAlgebraCL.php:95-253) -------------------------------------------------
    def getConfig(self) -> str:
        return "CudaB200"

    def getContext(self):
        return self.context

    def getQueue(self) -> CudaQueue:
        return self.queue

    def getBlas(self):
        return self.blas

    def getMath(self):
        return self.math

    def accelerated(self) -> bool:
        return True

    def fp64(self) -> bool:
        return True

    def deviceTypes(self):
        return ["GPU"]

    def blocking(self, switch: bool | None = None) -> bool:
        if switch is not None:
            self._blocking = bool(switch)
        return self._blocking

    def scalarNumeric(self, switch: bool | None = None) -> bool:
        if switch is not None:
            self._scalar_numeric = bool(switch)
        return self._scalar_numeric

    def finish(self) -> None:
        self.queue.finish()

    def newEventList(self):
        make = getattr(self.queue, "newEventList", None)         # a non-CUDA queue brings its own event lists
        return make() if make else CudaEventList()

    # ---- arrays (LinearAlgebraCL.php:321-368, 441-460, 793-814) --------------------------------------------------
    def array(self, array, dtype: int | None = None, flags: int | None = None) -> NDArrayCuda:
        if isinstance(array, NDArrayCuda):
            return array
        if not isinstance(array, NDArray):
            if dtype is None:
                a = np.asarray(array)
                dtype = dtypes.bool_ if a.dtype == np.bool_ else self.defaultFloatType
            array = NDArray(array, dtype=dtype, service=self.service)
        flags = (CL_MEM_READ_WRITE if flags is None else flags) | CL_MEM_COPY_HOST_PTR
        return NDArrayCuda(self.queue, array.buffer(), array.dtype(), array.shape(), array.offset(), flags, self.service)

    def toNDArray(self, ndarray):                                               # LinearAlgebraCL.php:354-360
        return ndarray.toNDArray() if isinstance(ndarray, NDArrayCuda) else ndarray

    def scalar(self, array):
        if isinstance(array, (NDArray, NDArrayCuda)):
            return array.toArray()
        return array

    def alloc(self, shape, dtype: int | None = None, flags: int | None = None) -> NDArrayCuda:
        return NDArrayCuda(self.queue, None, self.defaultFloatType if dtype is None else dtype, list(shape), None,
                           flags, self.service)

    def allocHost(self, shape, dtype: int | None = None) -> NDArray:
        return NDArray(None, dtype=self.defaultFloatType if dtype is None else dtype, shape=shape, service=self.service)

    def zerosHost(self, X: NDArray) -> NDArray:
        self.math.zeros(X.size(), X.buffer(), X.offset(), 1)
        return X

    def _element(self, X, i: int):
        return X.buffer()[X.offset() + i]                   # one small d2h (CudaDeviceBuffer.__getitem__)

    # ---- blocking / scalar conventions wrapped around every inherited operation -------------------------------------
    def _as_scalar_result(self, name: str, value, like):
        if self._scalar_numeric or not isinstance(value, (int, float, np.integer, np.floating)):
            return value
        if name in _INDEX_CALLS:
            dt = dtypes.int32
        else:
            dt = like.dtype() if like is not None and like.dtype() in (dtypes.float32, dtypes.float64) else self.defaultFloatType
        host = NDArray(np.asarray(value), dtype=dt, shape=[], service=self.service)
        return NDArrayCuda(self.queue, host.buffer(), dt, [], 0, CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, self.service)

    def __getattribute__(self, name):
        attr = object.__getattribute__(self, name)
        if name.startswith("_") or not callable(attr) or name in _PASSTHROUGH:
            return attr
        if not hasattr(LinearAlgebra, name):
            return attr

        def call(*args, **kwargs):
            out = attr(*args, **kwargs)
            if object.__getattribute__(self, "_blocking"):
                self.queue.finish()
            if name in _SCALAR_CALLS:
                like = args[0] if args and isinstance(args[0], NDArrayCuda) else None
                return self._as_scalar_result(name, out, like)
            return out
        call.__name__ = name
        return call


_PASSTHROUGH = frozenset(("accelerated", "array", "toNDArray", "alloc", "scalar", "allocHost", "zerosHost", "blocking",
                          "scalarNumeric", "finish", "fp64", "getConfig", "getContext", "getQueue", "getBlas", "getMath",
                          "deviceTypes", "newEventList"))
