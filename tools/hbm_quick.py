#!/usr/bin/env python3
"""Time the HBM-bound ops at BASELINE sizes with CUDA events (rotating operand sets) -- a quick loop for
kernel work; bench.py's hbm_ops is the reported number.  Usage: hbm_quick.py [op ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import rindow_math_matrix_b200 as rb  # noqa: E402
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes  # noqa: E402


def main():
    want = set(sys.argv[1:])
    svc = rb.MatlibCuda()
    math, blas = svc.math(), svc.blas()
    f32 = dtypes.float32
    rng = np.random.default_rng(0)
    stream = torch.cuda.Stream()
    _capi.set_stream(stream.cuda_stream)              # kernels + events on the same stream

    def rnd(n):
        return CudaBuffer(n, f32, zero=False).from_numpy(rng.uniform(-1, 1, n).astype(np.float32))

    def timed(name, fn, nbytes, sets=3, reps=30):
        if want and name not in want:
            return
        for i in range(6):
            fn(i % sets)
        _capi.sync()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(reps):
                fn(i % sets)
            e1.record(stream)
            _capi.sync()
            best = min(best, e0.elapsed_time(e1) / reps)
        print("%-16s %.4f ms  %7.0f GB/s  frac %.3f" % (name, best, nbytes / best / 1e6, nbytes / best / 1e6 / 6543.4), flush=True)

    m, n = 8192, 4096
    As = [rnd(m * n) for _ in range(3)]
    Y2 = rnd(m * n)
    X, Xm, Yn, Ym = rnd(n), rnd(m), CudaBuffer(n, f32), CudaBuffer(m, f32)
    timed("fill", lambda i: math.zeros(m * n, As[i], 0, 1), m * n * 4)
    if not want or "torch_zero" in want:
        tz = [torch.empty(m * n, dtype=torch.float32, device="cuda") for _ in range(3)]
        with torch.cuda.stream(stream):
            timed("torch_zero", lambda i: tz[i].zero_(), m * n * 4)
            timed("torch_copy", lambda i: tz[(i + 1) % 3].copy_(tz[i]), 2 * m * n * 4)
        del tz
    timed("copy", lambda i: blas.copy(m * n, As[i], 0, 1, Y2, 0, 1), 2 * m * n * 4)
    timed("axpy", lambda i: blas.axpy(m * n, 0.5, As[i], 0, 1, Y2, 0, 1), 3 * m * n * 4)
    timed("transpose", lambda i: blas.omatcopy(BLAS.RowMajor, BLAS.Trans, m, n, 1.0, As[i], 0, n, Y2, 0, m), 2 * m * n * 4)
    timed("gemv", lambda i: blas.gemv(BLAS.RowMajor, BLAS.NoTrans, m, n, 1.0, As[i], 0, n, X, 0, 1, 0.0, Ym, 0, 1), m * n * 4)
    timed("gemv_trans", lambda i: blas.gemv(BLAS.RowMajor, BLAS.Trans, m, n, 1.0, As[i], 0, n, Xm, 0, 1, 0.0, Yn, 0, 1), m * n * 4)
    timed("add", lambda i: math.add(False, m, n, 1.0, X, 0, 1, As[i], 0, n), 2 * m * n * 4)
    sh = CudaBuffer(4, dtypes.int32).from_numpy(np.asarray([64, 12, 128, 64], np.int32))
    pm = CudaBuffer(4, dtypes.int32).from_numpy(np.asarray([0, 2, 1, 3], np.int32))
    nel = 64 * 12 * 128 * 64
    timed("transpose_heads", lambda i: math.transpose(sh, pm, As[i], 0, Y2, 0), 2 * nel * 4)
    sh3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([64, 512, 1024], np.int32))
    pm3 = CudaBuffer(3, dtypes.int32).from_numpy(np.asarray([0, 2, 1], np.int32))
    timed("transpose_b21", lambda i: math.transpose(sh3, pm3, As[i], 0, Y2, 0), 2 * m * n * 4)
    if not want or "slice" in want or "repeat" in want:
        # split [8192, 4096] along axis 1 -> the [8192, 3072] part (96 MB out of 128 MB), and repeat [8192,1024] x 4
        timed("slice", lambda i: math.slice(False, False, m, n, 1, 1, As[i], 0, 1, Y2, 0, 1, 0, m, 512, 3072, 0, 1), 2 * m * 3072 * 4)
        timed("repeat", lambda i: math.repeat(m, 1024, 4, As[i], 0, Y2, 0), (m * 1024 + m * 4096) * 4)
    if not want or "cumsum" in want:
        # [8192, 4096] f32 (128 MB): along the rows (k = 1), down the columns (k = 4096) and as one 32M-element vector
        # (chunked: totals pass + scan pass = 2 reads + 1 write for 1 read + 1 write of algorithmic traffic)
        timed("cumsum_rows", lambda i: math.cumsumb(m, n, 1, As[i], 0, False, False, Y2, 0), 2 * m * n * 4)
        timed("cumsum_rows_4032", lambda i: math.cumsumb(m, 4032, 1, As[i], 0, False, False, Y2, 0), 2 * m * 4032 * 4)
        timed("cumsum_cols", lambda i: math.cumsumb(1, m, n, As[i], 0, False, False, Y2, 0), 2 * m * n * 4)
        timed("cumsum_flat", lambda i: math.cumsumb(1, m * n, 1, As[i], 0, False, False, Y2, 0), 2 * m * n * 4)
    if not want or "topk" in want:
        # top-8 of each row of [1024, 32768] f32 (the 128 MB operand): bytes = the input read once
        TV, TI = CudaBuffer(1024 * 8, f32), CudaBuffer(1024 * 8, dtypes.int32)
        timed("topk", lambda i: math.topk(1024, 32768, As[i], 0, 8, True, TV, 0, TI, 0), m * n * 4)
        TV2, TI2 = CudaBuffer(8192 * 8, f32), CudaBuffer(8192 * 8, dtypes.int32)
        timed("topk_8192x4096", lambda i: math.topk(8192, 4096, As[i], 0, 8, True, TV2, 0, TI2, 0), m * n * 4)
    if not want or "masking" in want:
        # attention-score padding mask: scores [64 batch, 16 heads x 128 queries, 256 keys] f32 (128 MB); per batch the keys
        # past a random length are masked out (set to -1e9): only those elements are written
        lens = rng.integers(64, 257, 64)
        mk = (np.arange(256)[None, :] < lens[:, None]).astype(np.uint8)
        Mk = CudaBuffer(64 * 256, dtypes.bool_).from_numpy(mk.reshape(-1))
        masked = int((mk == 0).sum()) * 16 * 128
        timed("masking", lambda i: math.masking(64, 16 * 128, 256, 1, -1e9, 0, Mk, 0, As[i], 0), masked * 4)
        timed("masking_add", lambda i: math.masking(64, 16 * 128, 256, 1, -1e9, 1, Mk, 0, As[i], 0), 2 * masked * 4)
    vocab, dim, ntok = 32768, 1024, 32768                       # table 128 MiB; one step moves 2 x 128 MiB
    ids = CudaBuffer(ntok, dtypes.int32).from_numpy(rng.integers(0, vocab, ntok).astype(np.int32))
    timed("gather", lambda i: math.gather(False, False, ntok, dim, vocab, ids, 0, As[i], 0, Y2, 0), 2 * ntok * dim * 4)
    timed("scatter_add", lambda i: math.gather(True, True, ntok, dim, vocab, ids, 0, As[i], 0, Y2, 0), 3 * ntok * dim * 4)
    timed("maximum", lambda i: math.maximum(m, n, As[i], 0, n, X, 0, 1), 2 * m * n * 4)
    timed("greater", lambda i: math.greater(m, n, As[i], 0, n, X, 0, 1), 2 * m * n * 4)
    timed("duplicate", lambda i: math.duplicate(False, m, n, X, 0, 1, As[i], 0, n), m * n * 4)
    timed("exp", lambda i: math.exp(m * n, As[i], 0, 1), 2 * m * n * 4)
    timed("tanh", lambda i: math.tanh(m * n, As[i], 0, 1), 2 * m * n * 4)
    timed("increment", lambda i: math.increment(m * n, 1.0001, As[i], 0, 1, 0.001), 2 * m * n * 4)
    timed("equal", lambda i: math.equal(m * n, As[i], 0, 1, Y2, 0, 1), 3 * m * n * 4)
    I32 = CudaBuffer(m * n, dtypes.int32)
    timed("astype_f32_i32", lambda i: math.astype(m * n, dtypes.int32, As[i], 0, 1, I32, 0, 1), 2 * m * n * 4)
    timed("sum", lambda i: math.sum(m * n, As[i], 0, 1), m * n * 4)
    timed("imax", lambda i: math.imax(m * n, As[i], 0, 1), m * n * 4)
    del As, Y2, I32
    m, n = 65536, 1024
    Rs = [rnd(m * n) for _ in range(3)]
    Bf = CudaBuffer(m, f32)
    timed("reduce_sum", lambda i: math.reduceSum(m, n, 1, Rs[i], 0, Bf, 0), m * n * 4)
    timed("reduce_sum_axis0", lambda i: math.reduceSum(1, m, n, Rs[i], 0, Bf, 0), m * n * 4)
    timed("softmax", lambda i: math.softmax(m, n, Rs[i], 0, n), 2 * m * n * 4)
    del Rs
    # BASELINE config C4: images [64,224,224,3] -> cols [64,222,222,3,3,3] -> x W [27,64]
    b, h, w, ch, f = 64, 224, 224, 3, 64
    oh, ow = h - 2, w - 2
    Im = [rnd(b * h * w * ch) for _ in range(3)]
    Co = CudaBuffer(b * oh * ow * 27, f32, zero=False)
    timed("im2col2d", lambda i: math.im2col2d(False, Im[i], 0, b * h * w * ch, b, h, w, ch, 3, 3, 1, 1, False, False, 1, 1,
                                              False, Co, 0, b * oh * ow * 27), (b * h * w * ch + b * oh * ow * 27) * 4)
    Wk, Out = rnd(27 * f), CudaBuffer(b * oh * ow * f, f32, zero=False)
    M = b * oh * ow
    timed("conv_gemm", lambda i: blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, M, f, 27, 1.0, Co, 0, 27, Wk, 0, f,
                                           0.0, Out, 0, f), (M * 27 + 27 * f + M * f) * 4)
    # Conv3D sibling: clips [8,16,112,112,3], 3x3x3 valid -> cols [8,14,110,110,3,3,3,3]
    if not want or "im2col3d" in want:
        b3, d3, h3, w3, c3 = 8, 16, 112, 112, 3
        V3 = [rnd(b3 * d3 * h3 * w3 * c3) for _ in range(3)]
        n3 = b3 * (d3 - 2) * (h3 - 2) * (w3 - 2) * 81
        C3 = CudaBuffer(n3, f32, zero=False)
        timed("im2col3d", lambda i: math.im2col3d(False, V3[i], 0, b3 * d3 * h3 * w3 * c3, b3, d3, h3, w3, c3, 3, 3, 3, 1, 1, 1,
                                                  False, False, 1, 1, 1, False, C3, 0, n3), (b3 * d3 * h3 * w3 * c3 + n3) * 4)
        del V3, C3
    Outs = [Out, CudaBuffer(b * oh * ow * f, f32, zero=False)]
    timed("conv2d_fused", lambda i: math.conv2d_forward(Im[i], 0, b, h, w, ch, 3, 3, 1, 1, False, 1, 1, Wk, 0, f, None, 0,
                                                        Outs[i % 2], 0), (b * h * w * ch + 27 * f + M * f) * 4)
    if not want or "conv2d_fused" in want:
        print("conv2d_fused path:", _capi.last_gemm_path(), flush=True)


if __name__ == "__main__":
    main()
