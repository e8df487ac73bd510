#!/usr/bin/env python3
"""Fused conv2d forward against im2col2d + gemm (+ add) on hidden-layer shapes: CUDA-event timings of both routes
through LinearAlgebra.conv2d, device-resident operands, rotating outputs."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import LinearAlgebra, _capi

_capi.init(0)
st = torch.cuda.Stream()
_capi.set_stream(st.cuda_stream)
svc = rb.MatlibCuda()
la = LinearAlgebra(svc)
res = {}
SHAPES = [(64, 56, 56, 64, 3, 64, True), (64, 28, 28, 128, 3, 128, True), (32, 112, 112, 32, 3, 64, True),
          (64, 14, 14, 256, 3, 256, True), (64, 56, 56, 64, 1, 256, False)]
if len(sys.argv) > 2:
    SHAPES = [SHAPES[int(sys.argv[2])]]
from rindow_math_matrix_b200 import BLAS, CudaBuffer, dtypes
math, blas = svc.math(), svc.blas()
f32 = dtypes.float32
for (b, h, w, c, k, f, pad) in SHAPES:
    rng = np.random.default_rng(c)
    img = CudaBuffer(b * h * w * c, f32, zero=False).from_numpy(rng.uniform(-1, 1, b * h * w * c).astype(np.float32))
    W = CudaBuffer(k * k * c * f, f32, zero=False).from_numpy((rng.standard_normal(k * k * c * f) * 0.1).astype(np.float32))
    bias = CudaBuffer(f, f32, zero=False).from_numpy(rng.standard_normal(f).astype(np.float32))
    oh, ow = (h, w) if pad else (h - k + 1, w - k + 1)
    M, K = b * oh * ow, k * k * c
    outs = [CudaBuffer(M * f, f32, zero=False) for _ in range(2)]
    cols = CudaBuffer(M * K, f32, zero=False)
    row = {"flops": 2.0 * M * K * f, "fused_bytes": (b * h * w * c + K * f + M * f) * 4,
           "two_call_bytes": (b * h * w * c + 2 * M * K + K * f + 3 * M * f) * 4}
    for mode_name, mode in (("tf32x3", rb.GEMM_TF32X3), ("tf32", rb.GEMM_TF32)):
        blas.setGemmMode(mode)

        def fused(i):
            ok = math.conv2d_forward(img, 0, b, h, w, c, k, k, 1, 1, pad, 1, 1, W, 0, f, bias, 0, outs[i % 2], 0, mode)
            assert ok

        def two_call(i):
            math.im2col2d(False, img, 0, b * h * w * c, b, h, w, c, k, k, 1, 1, pad, False, 1, 1, False, cols, 0, M * K)
            blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, f, K, 1.0, cols, 0, K, W, 0, f, 0.0, outs[i % 2], 0, f)
            math.add(False, M, f, 1.0, bias, 0, 1, outs[i % 2], 0, f)

        for name, fn in (("fused", fused), ("two_call", two_call)):
            with torch.cuda.stream(st):
                fn(0); fn(1)
                path = _capi.last_gemm_path()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record(st)
                for i in range(reps):
                    fn(i)
                e1.record(st)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row["%s_%s" % (mode_name, name)] = {"ms": ms, "tflops": row["flops"] / ms / 1e9, "path": path}
    res["%dx%dx%dx%d k%d f%d %s" % (b, h, w, c, k, f, "same" if pad else "valid")] = row
    print(json.dumps({("%dx%dx%dx%d k%d f%d" % (b, h, w, c, k, f)): row}), flush=True)
    del img, W, bias, outs, cols
if len(sys.argv) > 1 and sys.argv[1] != "-":
    open(sys.argv[1], "w").write(json.dumps(res, indent=1))
