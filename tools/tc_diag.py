#!/usr/bin/env python3
"""Bring-up diagnostic for the tcgen05 GEMM: runs a few shapes per (transA, transB, mode) and prints
error statistics plus where the errors sit, so that one GPU call tells which operand layout /
descriptor is wrong.  Not a test; used with gpurun during kernel development."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import rindow_math_matrix_b200 as rb  # noqa: E402
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes  # noqa: E402


def run(mode, m, n, k, tA, tB, integer=True, alpha=1.0, beta=0.0):
    rng = np.random.default_rng(1)
    rA, cA = (k, m) if tA else (m, k)
    rB, cB = (n, k) if tB else (k, n)
    if integer:
        Ah = rng.integers(-4, 5, (rA, cA)).astype(np.float32)
        Bh = rng.integers(-4, 5, (rB, cB)).astype(np.float32)
    else:
        Ah = rng.uniform(-1, 1, (rA, cA)).astype(np.float32)
        Bh = rng.uniform(-1, 1, (rB, cB)).astype(np.float32)
    A = CudaBuffer(Ah.size, dtypes.float32).from_numpy(Ah)
    B = CudaBuffer(Bh.size, dtypes.float32).from_numpy(Bh)
    C = CudaBuffer(m * n, dtypes.float32)
    blas = rb.CudaBlas(mode)
    t0 = time.time()
    blas.gemm(BLAS.RowMajor, BLAS.Trans if tA else BLAS.NoTrans, BLAS.Trans if tB else BLAS.NoTrans, m, n, k,
              alpha, A, 0, cA, B, 0, cB, beta, C, 0, n)
    _capi.sync()
    dt = time.time() - t0
    got = C.to_numpy().reshape(m, n).astype(np.float64)
    opA = (Ah.T if tA else Ah).astype(np.float64)
    opB = (Bh.T if tB else Bh).astype(np.float64)
    ref = alpha * (opA @ opB)
    err = np.abs(got - ref)
    rel = err.max() / max(np.abs(ref).max(), 1e-30)
    bad = err > 1e-3 * np.abs(ref).max()
    tag = "mode=%s tA=%d tB=%d %dx%dx%d int=%d" % (_capi.GEMM_MODE_NAMES[mode], tA, tB, m, n, k, integer)
    print("%-50s path=%-15s rel_err=%.3e bad=%6.2f%% %.1fms" % (tag, _capi.last_gemm_path(), rel,
                                                              100.0 * bad.mean(), dt * 1e3), flush=True)
    if bad.any() and integer:
        rows = np.where(bad.any(axis=1))[0]
        cols = np.where(bad.any(axis=0))[0]
        print("   bad rows: n=%d first=%s ; bad cols: n=%d first=%s" % (rows.size, rows[:12].tolist(), cols.size,
                                                                        cols[:12].tolist()))
        i, j = np.argwhere(bad)[0]
        print("   e.g. C[%d,%d] got=%g ref=%g ; zero_frac=%.2f nan_frac=%.2f" % (
            i, j, got[i, j], ref[i, j], float((got == 0).mean()), float(np.isnan(got).mean())))
    return rel


def main():
    _capi.init(0)
    print(_capi.device_info())
    shapes = [(128, 256, 32), (128, 256, 64), (256, 512, 256), (384, 384, 96), (1024, 1024, 1024)]
    if os.environ.get("RB200_DIAG_QUICK"):
        print("env:", {k: v for k, v in os.environ.items() if k.startswith("RB200_DBG")})
        for tA, tB in ((False, False), (True, True), (True, False)):
            for (m, n, k) in shapes[:3]:
                run(rb.GEMM_TF32, m, n, k, tA, tB, integer=True)
        return
    for mode in (rb.GEMM_TF32, rb.GEMM_TF32X3):
        for tA in (False, True):
            for tB in (False, True):
                for (m, n, k) in shapes:
                    run(mode, m, n, k, tA, tB, integer=True)
        run(mode, 512, 512, 512, False, False, integer=False)
        run(mode, 512, 512, 512, False, False, integer=False, alpha=0.5, beta=0.0)
    # truncation vs rounding of raw fp32 inputs in kind::tf32: A = 1 + 2^-11 + 2^-12 (needs 12 mantissa bits)
    m = n = 128
    k = 32
    Ah = np.full((m, k), 1.0 + 2.0 ** -11 + 2.0 ** -12, np.float32)
    Bh = np.eye(k, n, dtype=np.float32)
    A = CudaBuffer(Ah.size, dtypes.float32).from_numpy(Ah)
    B = CudaBuffer(Bh.size, dtypes.float32).from_numpy(Bh)
    C = CudaBuffer(m * n, dtypes.float32)
    rb.CudaBlas(rb.GEMM_TF32).gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0,
                                   C, 0, n)
    v = float(C.to_numpy()[0])
    kind = "truncation" if v == 1.0 else ("round-to-nearest" if abs(v - (1.0 + 2.0 ** -10)) < 1e-9 else "other")
    print("tf32 input conversion probe (B NoTrans): in=%.10f out=%.10f -> %s" % (float(Ah[0, 0]), v, kind))
    Bt = CudaBuffer(Bh.size, dtypes.float32).from_numpy(np.ascontiguousarray(Bh.T))
    rb.CudaBlas(rb.GEMM_TF32).gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.Trans, m, n, k, 1.0, A, 0, k, Bt, 0, k, 0.0,
                                   C, 0, n)
    v = float(C.to_numpy()[0])
    kind = "truncation" if v == 1.0 else ("round-to-nearest" if abs(v - (1.0 + 2.0 ** -10)) < 1e-9 else "other")
    print("tf32 input conversion probe (B Trans): in=%.10f out=%.10f -> %s" % (float(Ah[0, 0]), v, kind))


def accumulation_probe():
    """Accumulator rounding of tcgen05 kind::tf32: inputs with <= 8 mantissa bits (exact in tf32,
    products exact in fp32), all positive, so every bit of error comes from the fp32 accumulation
    in TMEM.  A round-to-nearest accumulator gives a mean relative error ~0 growing like sqrt(K);
    a truncating one gives a negative mean growing like K."""
    rng = np.random.default_rng(3)
    m = n = 128
    for mode in (rb.GEMM_TF32, rb.GEMM_TF32X3, rb.GEMM_FP32_SIMT):
        for k in (256, 1024, 4096, 16384):
            Ah = (np.round(rng.uniform(0.5, 1.0, (m, k)) * 256) / 256).astype(np.float32)
            Bh = (np.round(rng.uniform(0.5, 1.0, (n, k)) * 256) / 256).astype(np.float32)
            A = CudaBuffer(Ah.size, dtypes.float32).from_numpy(Ah)
            B = CudaBuffer(Bh.size, dtypes.float32).from_numpy(Bh)
            C = CudaBuffer(m * n, dtypes.float32)
            rb.CudaBlas(mode).gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.Trans, m, n, k, 1.0, A, 0, k, B, 0, k, 0.0,
                                   C, 0, n)
            got = C.to_numpy().reshape(m, n).astype(np.float64)
            ref = Ah.astype(np.float64) @ Bh.astype(np.float64).T
            rel = (got - ref) / ref
            print("accum probe mode=%-9s K=%6d path=%-14s mean_rel=%+.3e std_rel=%.3e max_abs_rel=%.3e" % (
                _capi.GEMM_MODE_NAMES[mode], k, _capi.last_gemm_path(), rel.mean(), rel.std(), np.abs(rel).max()),
                flush=True)


if __name__ == "__main__":
    if os.environ.get("RB200_DIAG_ACCUM"):
        _capi.init(0)
        accumulation_probe()
        sys.exit(0)
    main()
