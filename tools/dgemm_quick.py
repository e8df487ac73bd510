#!/usr/bin/env python3
"""Quick timing of rb200_dgemm and the FFMA sgemm (device-resident, CUDA events via torch on the library stream)."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes

_capi.init(0)
st = torch.cuda.Stream()
_capi.set_stream(st.cuda_stream)
blas = rb.MatlibCuda().blas()
res = {}
for name, dt, npdt, sizes in (("f64", dtypes.float64, np.float64, (2048, 4096, 8192)), ("f32_simt", dtypes.float32, np.float32, (2048, 4096, 8192))):
    for n in sizes:
        rng = np.random.default_rng(n)
        A = CudaBuffer(n * n, dt, zero=False).from_numpy(rng.uniform(-1, 1, n * n).astype(npdt))
        B = CudaBuffer(n * n, dt, zero=False).from_numpy(rng.uniform(-1, 1, n * n).astype(npdt))
        C = CudaBuffer(n * n, dt)
        blas.setGemmMode(rb.GEMM_FP32_SIMT)
        for tA, tB in ((BLAS.NoTrans, BLAS.NoTrans), (BLAS.Trans, BLAS.NoTrans), (BLAS.NoTrans, BLAS.Trans)):
            def step():
                blas.gemm(BLAS.RowMajor, tA, tB, n, n, n, 1.0, A, 0, n, B, 0, n, 0.0, C, 0, n)
            with torch.cuda.stream(st):
                step(); step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record(st)
                for _ in range(reps):
                    step()
                e1.record(st)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res["%s_%d_%s%s" % (name, n, "T" if tA == BLAS.Trans else "N", "T" if tB == BLAS.Trans else "N")] = {
                "ms": ms, "tflops": 2.0 * n ** 3 / ms / 1e9, "path": _capi.last_gemm_path()}
        del A, B, C
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(json.dumps(res, indent=1))
