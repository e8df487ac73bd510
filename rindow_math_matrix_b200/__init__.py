"""rindow_math_matrix_b200 -- B200-native (sm_100a) compute driver for rindow-math-matrix.

Hot path only: GEMM, broadcast add/multiply, axis reductions, softmax,
im2col2d, behind the reference's Service / blas() / math() / buffer() plugin interface.  The
arithmetic lives in librindow_b200.so (csrc/, C ABI in include/rindow_b200.h); this package is
the host-side mirror of the reference's PHP classes for that path.  No CPU fallback anywhere.
"""
from . import dtypes  # noqa: F401
from ._capi import GEMM_AUTO, GEMM_FP32_SIMT, GEMM_TF32, GEMM_TF32X3, B200Error  # noqa: F401
from .cuda_blas import BLAS, CudaBlas, CudaBlasFactory  # noqa: F401
from .cuda_buffer import CudaBuffer, CudaBufferFactory  # noqa: F401
from .cuda_math import CudaMath, CudaMathFactory  # noqa: F401
from .exceptions import (InvalidArgumentException, LogicException, OutOfRangeException,  # noqa: F401
                         RuntimeException)
from .linear_algebra import LinearAlgebra  # noqa: F401
from .linear_algebra_cuda import LinearAlgebraCuda  # noqa: F401
from .ndarray_cuda import (CudaDeviceBuffer, CudaDeviceBufferFactory, CudaEventList, CudaQueue,  # noqa: F401
                           NDArrayCuda)
from .matrix_operator import MatrixOperator  # noqa: F401
from .ndarray import NDArray  # noqa: F401
from .service import MatlibCuda, Service  # noqa: F401

__all__ = ["MatlibCuda", "Service", "MatrixOperator", "LinearAlgebra", "LinearAlgebraCuda", "NDArray", "NDArrayCuda", "CudaBuffer", "CudaBlas",
           "CudaMath", "BLAS", "dtypes"]
