#!/usr/bin/env python3
"""Time rb200_sgemm over sizes / modes with CUDA events (random operands).  Usage: gemm_sweep.py [mode] [sizes...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import rindow_math_matrix_b200 as rb  # noqa: E402
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes  # noqa: E402


def main():
    mode_name = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    sizes = [int(x) for x in sys.argv[2:]] or [512, 768, 1024, 1536, 2048, 3072, 4096]
    mode = {"tf32": rb.GEMM_TF32, "tf32x3": rb.GEMM_TF32X3}[mode_name]
    blas = rb.MatlibCuda(gemm_mode=mode).blas()
    stream = torch.cuda.Stream()
    _capi.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(0)
    for n in sizes:
        A = [CudaBuffer(n * n, dtypes.float32).from_numpy(rng.uniform(-1, 1, n * n).astype(np.float32)) for _ in range(3)]
        B = CudaBuffer(n * n, dtypes.float32).from_numpy(rng.uniform(-1, 1, n * n).astype(np.float32))
        C = CudaBuffer(n * n, dtypes.float32)

        def step(i):
            blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, n, n, n, 1.0, A[i % 3], 0, n, B, 0, n, 0.0, C, 0, n)

        for i in range(5):
            step(i)
        _capi.sync()
        reps = max(10, min(200, int(2e12 / (2.0 * n ** 3)) + 1))
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(reps):
                step(i)
            e1.record(stream)
            _capi.sync()
            best = min(best, e0.elapsed_time(e1) / reps)
        # the same launches replayed from a CUDA graph: the GPU-side rate without the host cost of issuing each call
        per = 20
        g = _capi.Graph()
        with g:
            for i in range(per):
                step(i)
        gbest = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(max(1, reps // per)):
                g.launch()
            e1.record(stream)
            _capi.sync()
            gbest = min(gbest, e0.elapsed_time(e1) / (max(1, reps // per) * per))
        g.close()
        print("%-7s n=%5d  issued %.4f ms %7.1f TFLOP/s | graph %.4f ms %7.1f TFLOP/s  path=%s" % (
            mode_name, n, best, 2.0 * n ** 3 / best / 1e9, gbest, 2.0 * n ** 3 / gbest / 1e9, _capi.last_gemm_path()), flush=True)


def batched():
    """attention-shaped strided batches: [B*H, S, D] x [B*H, D, S] and [B*H, S, S] x [B*H, S, D]"""
    mode_name = sys.argv[2] if len(sys.argv) > 2 else "tf32"
    mode = {"tf32": rb.GEMM_TF32, "tf32x3": rb.GEMM_TF32X3, "simt": rb.GEMM_FP32_SIMT}[mode_name]
    math = rb.MatlibCuda(gemm_mode=mode).math()
    stream = torch.cuda.Stream()
    _capi.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(0)
    for (bh, s, d) in ((768, 512, 64), (256, 1024, 64), (1536, 128, 64), (96, 2048, 128)):
        Q = CudaBuffer(bh * s * d, dtypes.float32).from_numpy(rng.uniform(-1, 1, bh * s * d).astype(np.float32))
        Kt = CudaBuffer(bh * s * d, dtypes.float32).from_numpy(rng.uniform(-1, 1, bh * s * d).astype(np.float32))
        S = CudaBuffer(bh * s * s, dtypes.float32)
        O = CudaBuffer(bh * s * d, dtypes.float32)

        def qk():
            math.gemmStridedBatched(BLAS.RowMajor, BLAS.NoTrans, BLAS.Trans, s, s, d, 1.0, Q, 0, d, s * d, Kt, 0, d, s * d, 0.0,
                                    S, 0, s, s * s, bh, gemm_mode=mode)

        def sv():
            math.gemmStridedBatched(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, s, d, s, 1.0, S, 0, s, s * s, Kt, 0, d, s * d, 0.0,
                                    O, 0, d, s * d, bh, gemm_mode=mode)

        for name, fn, flops, nbytes in (("QK^T", qk, 2.0 * bh * s * s * d, 4.0 * bh * (2 * s * d + s * s)),
                                        ("SV", sv, 2.0 * bh * s * s * d, 4.0 * bh * (2 * s * d + s * s))):
            for _ in range(3):
                fn()
            _capi.sync()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    fn()
                e1.record(stream)
                _capi.sync()
                best = min(best, e0.elapsed_time(e1) / 10)
            print("%-6s %-5s bh=%5d s=%5d d=%4d  %.4f ms  %7.1f TFLOP/s  %6.0f GB/s  path=%s" % (
                mode_name, name, bh, s, d, best, flops / best / 1e9, nbytes / best / 1e6, _capi.last_gemm_path()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "batched":
        batched()
    else:
        main()
