"""Parity of gatherb / scatterb / scatterbAdd (rb200_gatherb) through the C ABI: the reference's own 26 test cases
(tests/golden/gatherb_cases.json) and seeded random shapes against the CPU oracle.  Bit-exact everywhere: copies move
words, and the adds run in the reference's order (ascending j) with one rounding per step.  pytest -m gpu"""
import json
import os

import numpy as np
import pytest

import golden_runner as gr
import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, dtypes

pytestmark = pytest.mark.gpu

with open(os.path.join(gr.HERE, "golden", "gatherb_cases.json")) as f:
    GATHERB = json.load(f)["cases"]


@pytest.mark.parametrize("case", GATHERB, ids=[c["op"] + "_" + c["ref"].split(":")[1] for c in GATHERB])
def test_gatherb_reference_cases(case, cuda_service):
    gr.run_gatherb_case(case, cuda_service)


with open(os.path.join(gr.HERE, "golden", "gathernd_cases.json")) as f:
    GATHERND = json.load(f)["cases"]


@pytest.mark.parametrize("case", GATHERND, ids=[c["op"] + "_" + c["ref"].split(":")[1] for c in GATHERND])
def test_gathernd_reference_cases(case, cuda_service):
    gr.run_gatherb_case(case, cuda_service)


ND_SHAPES = [(1, 1, 1, [1]), (2, 5, 3, [4]), (3, 40, 2, [3, 4]), (1, 500, 1, [6, 5]), (2, 17, 8, [2, 3, 2]), (4, 9, 64, [7])]


@pytest.mark.parametrize("m,n,k,dims", ND_SHAPES)
@pytest.mark.parametrize("reverse,addMode", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dt,idt", [(dtypes.float32, dtypes.int32), (dtypes.float64, dtypes.int64), (dtypes.int64, dtypes.int32)])
def test_gathernd_random_vs_oracle(m, n, k, dims, reverse, addMode, dt, idt, cuda_service):
    npdt, npi = dtypes.NUMPY[dt], dtypes.NUMPY[idt]
    rng = np.random.default_rng(n * 5 + k + len(dims))
    depth, psize = len(dims), int(np.prod(dims))
    offA, offB = 2, 3
    if dt == dtypes.int64:
        a = rng.integers(-50, 50, offA + m * psize * k).astype(npdt)
        b = rng.integers(-50, 50, offB + m * n * k).astype(npdt)
    else:
        a = rng.standard_normal(offA + m * psize * k).astype(npdt)
        b = rng.standard_normal(offB + m * n * k).astype(npdt)
    x = np.stack([rng.integers(0, d, m * n) for d in dims], axis=1).astype(npi)         # collisions whenever n > psize
    if m * n > 4:
        x[2, depth - 1] = dims[-1]                                                        # out of range: row skipped
    x = x.reshape(-1)
    A, B = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(b.size, dt).from_numpy(b)
    X = CudaBuffer(x.size, idt).from_numpy(x)
    cuda_service.math().gathernd(reverse, addMode, m, n, k, depth, dims, A, offA, X, 0, B, offB)
    ao, bo = a.astype(np.float64), b.astype(np.float64)
    oracle.gathernd(reverse, addMode, dt == dtypes.float32, m, n, k, depth, dims, ao, offA, x.astype(np.float64), bo, offB)
    assert np.array_equal(A.to_numpy().astype(np.float64), ao)      # float32 STORAGE model of the oracle: bit-exact
    if dt == dtypes.float32 and addMode:
        # ... and against the reference's own semantics (a PhpBuffer holds doubles, nothing is rounded between the
        # steps): within terms * eps32 * sum|terms| per element, the sum of magnitudes taken from
        # the same unrounded oracle run on |a|, |b|
        au, bu = a.astype(np.float64), b.astype(np.float64)
        oracle.gathernd(reverse, addMode, False, m, n, k, depth, dims, au, offA, x.astype(np.float64), bu, offB)
        am, bm = np.abs(a).astype(np.float64), np.abs(b).astype(np.float64)
        oracle.gathernd(reverse, addMode, False, m, n, k, depth, dims, am, offA, x.astype(np.float64), bm, offB)
        eps = float(np.finfo(np.float32).eps)
        assert np.all(np.abs(A.to_numpy().astype(np.float64) - au) <= (n + 1) * eps * am + 1e-300)
        assert np.all(np.abs(B.to_numpy().astype(np.float64) - bu) <= (n + 1) * eps * bm + 1e-300)
    assert np.array_equal(B.to_numpy().astype(np.float64), bo)


def test_gathernd_errors(cuda_service):
    la = LinearAlgebra(cuda_service)
    A = la.array(np.arange(40, dtype=np.float32).reshape(4, 5, 2))
    X = la.array(np.array([[0], [1], [2], [0]], np.int32), dtype=dtypes.int32)
    with pytest.raises(InvalidArgumentException, match="Unmatch batch shape"):
        la.gatherND(A, la.array(np.array([[0], [1]], np.int32), dtype=dtypes.int32), batchDims=1)
    with pytest.raises(InvalidArgumentException, match="index depth"):
        la.gatherND(A, la.array(np.zeros((2, 4), np.int32), dtype=dtypes.int32))
    with pytest.raises(InvalidArgumentException, match="Unmatch output shape"):
        la.gatherND(A, X, batchDims=1, outputs=la.alloc([4, 3], dtypes.float32))
    with pytest.raises(InvalidArgumentException, match="Unmatch updates shape"):
        la.scatterND(X, la.alloc([4, 3], dtypes.float32), [4, 5, 2], batchDims=1)


SHAPES = [(1, 1, 1, 1, 1, 1), (2, 3, 4, 1, 1, 5), (3, 2, 7, 4, 1, 6), (2, 1, 9, 3, 5, 4), (1, 4, 300, 2, 3, 17),
          (4, 2, 33, 8, 16, 40), (1, 1, 2000, 1, 1, 50), (2, 2, 5, 64, 4, 3)]


@pytest.mark.parametrize("batches,m,n,k,len_,numClass", SHAPES)
@pytest.mark.parametrize("reverse,addMode", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("dt,idt", [(dtypes.float32, dtypes.int32), (dtypes.float64, dtypes.int64), (dtypes.int32, dtypes.int32)])
def test_gatherb_random_vs_oracle(batches, m, n, k, len_, numClass, reverse, addMode, dt, idt, cuda_service):
    npdt, npi = dtypes.NUMPY[dt], dtypes.NUMPY[idt]
    rng = np.random.default_rng(n * 7 + k + numClass)
    offA, offX, offB = 3, 1, 2
    if dt == dtypes.int32:
        a = rng.integers(-50, 50, offA + batches * m * numClass * k * len_).astype(npdt)
        b = rng.integers(-50, 50, offB + batches * m * n * k * len_).astype(npdt)
    else:
        a = rng.standard_normal(offA + batches * m * numClass * k * len_).astype(npdt)     # inexact sums: order matters
        b = rng.standard_normal(offB + batches * m * n * k * len_).astype(npdt)
    x = rng.integers(0, numClass, offX + batches * n * k).astype(npi)                      # collisions whenever n > numClass
    if x.size > 8:
        x[offX + 3] = numClass + 2                                                          # out of range: skipped
    A, B = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(b.size, dt).from_numpy(b)
    X = CudaBuffer(x.size, idt).from_numpy(x)
    cuda_service.math().gatherb(reverse, addMode, batches, m, n, k, len_, numClass, A, offA, X, offX, B, offB)
    ao, bo = a.astype(np.float64), b.astype(np.float64)
    oracle.gatherb(reverse, addMode, dt == dtypes.float32, batches, m, n, k, len_, numClass, ao, offA, x.astype(np.float64), offX,
                   bo, offB)
    assert np.array_equal(A.to_numpy().astype(np.float64), ao)      # float32 STORAGE model of the oracle: bit-exact
    assert np.array_equal(B.to_numpy().astype(np.float64), bo)
    assert np.array_equal(X.to_numpy(), x)
    if dt == dtypes.float32 and addMode:
        # the reference's own semantics (PhpBuffer holds doubles: no rounding between the steps):
        # within terms * eps32 * sum|terms| per element
        au, bu = a.astype(np.float64), b.astype(np.float64)
        oracle.gatherb(reverse, addMode, False, batches, m, n, k, len_, numClass, au, offA, x.astype(np.float64), offX, bu, offB)
        am, bm = np.abs(a).astype(np.float64), np.abs(b).astype(np.float64)
        oracle.gatherb(reverse, addMode, False, batches, m, n, k, len_, numClass, am, offA, x.astype(np.float64), offX, bm, offB)
        eps = float(np.finfo(np.float32).eps)
        assert np.all(np.abs(A.to_numpy().astype(np.float64) - au) <= (n + 1) * eps * am + 1e-300)
        assert np.all(np.abs(B.to_numpy().astype(np.float64) - bu) <= (n + 1) * eps * bm + 1e-300)


@pytest.mark.parametrize("dt", [dtypes.bool_, dtypes.uint8, dtypes.int16, dtypes.uint64])
def test_gatherb_copy_any_dtype(dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(4)
    batches, m, n, k, len_, numClass = 2, 3, 6, 2, 3, 4
    a = rng.integers(0, 2 if dt == dtypes.bool_ else 100, batches * m * numClass * k * len_).astype(npdt)
    x = rng.integers(0, numClass, batches * n * k).astype(np.int32)
    A, X = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(x.size, dtypes.int32).from_numpy(x)
    B = CudaBuffer(batches * m * n * k * len_, dt)
    cuda_service.math().gatherb(False, False, batches, m, n, k, len_, numClass, A, 0, X, 0, B, 0)
    want = np.zeros(batches * m * n * k * len_)
    oracle.gatherb(False, False, False, batches, m, n, k, len_, numClass, a.astype(np.float64), 0, x.astype(np.float64), 0, want, 0)
    assert np.array_equal(B.to_numpy().astype(np.float64), want)
    with pytest.raises(Exception, match="add mode"):
        cuda_service.math().gatherb(False, True, batches, m, n, k, len_, numClass, A, 0, X, 0, B, 0)


def test_gatherb_errors(cuda_service):
    la = LinearAlgebra(cuda_service)
    A = la.array(np.arange(24, dtype=np.float32).reshape(4, 3, 2))
    X = la.array(np.array([2, 2, 1, 0], np.int32), dtype=dtypes.int32)
    with pytest.raises(InvalidArgumentException, match="batchDims"):
        la.gatherb(A, X, batchDims=3)
    with pytest.raises(InvalidArgumentException, match="axis"):
        la.gatherb(A, X, axis=3)
    with pytest.raises(InvalidArgumentException, match="Unmatch batch shape"):
        la.gatherb(A, la.array(np.array([2, 2, 1], np.int32), dtype=dtypes.int32), axis=1, batchDims=1)
    with pytest.raises(InvalidArgumentException, match="Unmatch output shape"):
        la.gatherb(A, X, axis=1, batchDims=1, outputs=la.alloc([4, 3], dtypes.float32))
    m = cuda_service.math()
    with pytest.raises(InvalidArgumentException, match="Matrix A"):
        m.gatherb(False, False, 1, 1, 2, 1, 1, 30, A.buffer(), 0, X.buffer(), 0, CudaBuffer(2, dtypes.float32), 0)
    with pytest.raises(InvalidArgumentException, match="integer"):
        m.gatherb(False, False, 1, 1, 2, 1, 1, 3, A.buffer(), 0, CudaBuffer(2, dtypes.float32), 0, CudaBuffer(2, dtypes.float32), 0)


# ---- imagecopy -------------------------------------------------------------------------------------------------------

with open(os.path.join(gr.HERE, "golden", "imagecopy_cases.json")) as f:
    IMAGECOPY = json.load(f)["cases"]


@pytest.mark.parametrize("case", IMAGECOPY, ids=["imagecopy_" + c["ref"].split(":")[1] for c in IMAGECOPY])
def test_imagecopy_reference_cases(case, cuda_service):
    gr.run_imagecopy_case(case, cuda_service)


@pytest.mark.parametrize("h,w,c", [(1, 1, 1), (7, 5, 3), (64, 48, 4), (224, 224, 3)])
@pytest.mark.parametrize("channels_first", [False, True])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.uint8, dtypes.float64])
def test_imagecopy_random_vs_oracle(h, w, c, channels_first, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(h * 3 + w)
    offA, offB = 2, 1
    a = rng.integers(0, 255, offA + h * w * c).astype(npdt)
    for hs, ws, vf, hf, rgb in [(0, 0, False, False, False), (3, -2, True, False, c >= 3), (-h - 1, w + 2, False, True, False),
                                (1, 1, True, True, c >= 3)]:
        b = np.full(offB + h * w * c, 7).astype(npdt)
        A, B = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(b.size, dt).from_numpy(b)
        cuda_service.math().imagecopy(h, w, c, A, offA, B, offB, channels_first, hs, ws, vf, hf, rgb)
        bo = b.astype(np.float64)
        oracle.imagecopy(h, w, c, a.astype(np.float64), offA, bo, offB, channels_first, hs, ws, vf, hf, rgb)
        assert np.array_equal(B.to_numpy().astype(np.float64), bo), (hs, ws, vf, hf, rgb)


def test_imagecopy_errors(cuda_service):
    la = LinearAlgebra(cuda_service)
    with pytest.raises(InvalidArgumentException, match="3D"):
        la.imagecopy(la.alloc([4, 4], dtypes.float32))
    A = la.alloc([4, 4, 3], dtypes.float32)
    with pytest.raises(InvalidArgumentException, match="output shape"):
        la.imagecopy(A, la.alloc([4, 4, 4], dtypes.float32))
    with pytest.raises(InvalidArgumentException, match="output data type"):
        la.imagecopy(A, la.alloc([4, 4, 3], dtypes.float64))
    with pytest.raises(InvalidArgumentException, match="bufferB"):
        cuda_service.math().imagecopy(4, 4, 3, A.buffer(), 0, CudaBuffer(47, dtypes.float32), 0, False, 0, 0, False, False, False)
