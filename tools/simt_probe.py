#!/usr/bin/env python3
"""One FFMA and one DFMA GEMM (4096^3, NoTrans/NoTrans) for an ncu capture of gemm_simt_big_kernel."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
_capi.init(0)
blas = rb.MatlibCuda().blas()
blas.setGemmMode(rb.GEMM_FP32_SIMT)
for dt, npdt in ((dtypes.float32, np.float32), (dtypes.float64, np.float64)):
    rng = np.random.default_rng(n)
    A = CudaBuffer(n * n, dt, zero=False).from_numpy(rng.uniform(-1, 1, n * n).astype(npdt))
    B = CudaBuffer(n * n, dt, zero=False).from_numpy(rng.uniform(-1, 1, n * n).astype(npdt))
    C = CudaBuffer(n * n, dt)
    for _ in range(2):
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, n, n, n, 1.0, A, 0, n, B, 0, n, 0.0, C, 0, n)
    _capi.sync()
    print(_capi.last_gemm_path())
