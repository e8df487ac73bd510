#!/usr/bin/env python3
"""Run one hot-path op a few times at its BASELINE size -- the command ncu wraps
(profiles/README.md lists the exact ncu invocations).  Usage: profile_ops.py OP [REPS]
OP in: sgemm_tf32 sgemm_tf32x3 sgemm_simt add multiply reduce_sum reduce_sum_axis0 reduce_max
       reduce_argmax softmax im2col2d im2col3d conv_gemm conv2d_fused axpy copy transpose gemv gemv_trans
       maximum exp astype transpose_nd gather scatter_add sum cumsum_rows cumsum_cols topk"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import rindow_math_matrix_b200 as rb  # noqa: E402
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes  # noqa: E402


def main():
    op = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    svc = rb.MatlibCuda()
    math, blas = svc.math(), svc.blas()
    f32 = dtypes.float32
    rng = np.random.default_rng(0)

    def rnd(n):
        return CudaBuffer(n, f32, zero=False).from_numpy(rng.uniform(-1, 1, n).astype(np.float32))

    if op.startswith("sgemm"):
        n = int(os.environ.get("N", "8192"))
        mode = {"sgemm_tf32": rb.GEMM_TF32, "sgemm_tf32x3": rb.GEMM_TF32X3, "sgemm_simt": rb.GEMM_FP32_SIMT}[op]
        A, B, C = rnd(n * n), rnd(n * n), CudaBuffer(n * n, f32)
        blas.setGemmMode(mode)
        fn = lambda: blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, n, n, n, 1.0, A, 0, n, B, 0, n, 0.0, C, 0, n)
    elif op in ("add", "multiply"):
        m, n = 8192, 4096
        A, X = rnd(m * n), rnd(n)
        fn = (lambda: math.add(False, m, n, 1.0, X, 0, 1, A, 0, n)) if op == "add" else \
             (lambda: math.multiply(False, m, n, X, 0, 1, A, 0, n))
    elif op in ("reduce_sum", "reduce_max", "reduce_argmax", "reduce_sum_axis0", "softmax"):
        m, n = 65536, 1024
        A = rnd(m * n)
        Bf, Bi = CudaBuffer(m, f32), CudaBuffer(m, dtypes.int32)
        fn = {"reduce_sum": lambda: math.reduceSum(m, n, 1, A, 0, Bf, 0),
              "reduce_max": lambda: math.reduceMax(m, n, 1, A, 0, Bf, 0),
              "reduce_argmax": lambda: math.reduceArgMax(m, n, 1, A, 0, Bi, 0),
              "reduce_sum_axis0": lambda: math.reduceSum(1, m, n, A, 0, Bf, 0),
              "softmax": lambda: math.softmax(m, n, A, 0, n)}[op]
    elif op in ("im2col2d", "conv_gemm", "conv2d_fused"):
        b, h, w, c = 64, 224, 224, 3
        images = rnd(b * h * w * c)
        ncols = b * 222 * 222 * 27
        cols = CudaBuffer(ncols, f32, zero=False)
        im2col = lambda: math.im2col2d(False, images, 0, b * h * w * c, b, h, w, c, 3, 3, 1, 1, False, False, 1, 1,
                                       False, cols, 0, ncols)
        if op == "conv2d_fused":
            M = b * 222 * 222
            Wt, out = rnd(27 * 64), CudaBuffer(M * 64, f32, zero=False)
            fn = lambda: math.conv2d_forward(images, 0, b, h, w, c, 3, 3, 1, 1, False, 1, 1, Wt, 0, 64, None, 0, out, 0)
        elif op == "im2col2d":
            fn = im2col
        else:
            im2col()
            M = b * 222 * 222
            Wt, out = rnd(27 * 64), CudaBuffer(M * 64, f32, zero=False)
            blas.setGemmMode(0)
            fn = lambda: blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, 64, 27, 1.0, cols, 0, 27, Wt, 0, 64,
                                   0.0, out, 0, 64)
    elif op == "im2col3d":
        b3, d3, h3, w3, c3 = 8, 16, 112, 112, 3
        vol = rnd(b3 * d3 * h3 * w3 * c3)
        n3 = b3 * (d3 - 2) * (h3 - 2) * (w3 - 2) * 81
        c3d = CudaBuffer(n3, f32, zero=False)
        fn = lambda: math.im2col3d(False, vol, 0, b3 * d3 * h3 * w3 * c3, b3, d3, h3, w3, c3, 3, 3, 3, 1, 1, 1, False, False,
                                   1, 1, 1, False, c3d, 0, n3)
    elif op == "cumsum_cols":
        m, n = 8192, 4096
        A, Y = rnd(m * n), CudaBuffer(m * n, f32, zero=False)
        fn = lambda: math.cumsumb(1, m, n, A, 0, False, False, Y, 0)
    elif op == "topk":
        A = rnd(8192 * 4096)
        TV, TI = CudaBuffer(8192 * 8, f32), CudaBuffer(8192 * 8, dtypes.int32)
        fn = lambda: math.topk(8192, 4096, A, 0, 8, True, TV, 0, TI, 0)
    elif op == "cumsum_rows":
        m, n = 8192, 4096
        A, Y = rnd(m * n), CudaBuffer(m * n, f32, zero=False)
        fn = lambda: math.cumsumb(m, n, 1, A, 0, False, False, Y, 0)
    elif op in ("axpy", "copy", "transpose", "gemv", "gemv_trans"):
        m, n = 8192, 4096
        A, Y = rnd(m * n), rnd(m * n)
        Xn, Xm, Yn, Ym = rnd(n), rnd(m), CudaBuffer(n, f32), CudaBuffer(m, f32)
        fn = {"axpy": lambda: blas.axpy(m * n, 0.5, A, 0, 1, Y, 0, 1),
              "copy": lambda: blas.copy(m * n, A, 0, 1, Y, 0, 1),
              "transpose": lambda: blas.omatcopy(BLAS.RowMajor, BLAS.Trans, m, n, 1.0, A, 0, n, Y, 0, m),
              "gemv": lambda: blas.gemv(BLAS.RowMajor, BLAS.NoTrans, m, n, 1.0, A, 0, n, Xn, 0, 1, 0.0, Ym, 0, 1),
              "gemv_trans": lambda: blas.gemv(BLAS.RowMajor, BLAS.Trans, m, n, 1.0, A, 0, n, Xm, 0, 1, 0.0, Yn, 0, 1),
              }[op]
    elif op in ("maximum", "exp", "astype", "transpose_nd", "gather", "scatter_add", "sum"):
        m, n = 8192, 4096
        A, Y, X = rnd(m * n), rnd(m * n), rnd(n)
        I32 = CudaBuffer(m * n, dtypes.int32, zero=False)
        sh3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([64, 512, 1024], np.int32))
        pm3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([0, 2, 1], np.int32))
        vocab, dim, ntok = 32768, 1024, 32768
        ids = CudaBuffer(ntok, dtypes.int32).from_numpy(rng.integers(0, vocab, ntok).astype(np.int32))
        fn = {"maximum": lambda: math.maximum(m, n, A, 0, n, X, 0, 1),
              "exp": lambda: math.exp(m * n, A, 0, 1),
              "astype": lambda: math.astype(m * n, dtypes.int32, A, 0, 1, I32, 0, 1),
              "transpose_nd": lambda: math.transpose(sh3, pm3, A, 0, Y, 0),
              "gather": lambda: math.gather(False, False, ntok, dim, vocab, ids, 0, A, 0, Y, 0),
              "scatter_add": lambda: math.gather(True, True, ntok, dim, vocab, ids, 0, A, 0, Y, 0),
              "sum": lambda: math.sum(m * n, A, 0, 1),
              }[op]
    else:
        raise SystemExit("unknown op " + op)
    for _ in range(reps):
        fn()
    _capi.sync()
    print(op, "ok", _capi.last_gemm_path() if "gemm" in op else "")


if __name__ == "__main__":
    main()
