"""gather / scatter / scatterAdd / reduceGather through the C ABI against the CPU oracle.  Copies are bit-exact; the
add forms reproduce the reference's sequential association (ascending j on top of the existing value), so float32
results are compared bit-for-bit against the oracle run on float32-rounded partial sums, and within 1e-6 relative of
the double result.  pytest -m gpu"""
import numpy as np
import pytest

import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, dtypes

pytestmark = pytest.mark.gpu

NP = {k: v for k, v in dtypes.NUMPY.items() if k not in dtypes.COMPLEX_TYPES}     # real dtypes: the arithmetic entry points
F32, F64, I32, I64, U8 = dtypes.float32, dtypes.float64, dtypes.int32, dtypes.int64, dtypes.uint8


def buf(a, dt):
    a = np.asarray(a)
    return CudaBuffer(a.size, dt).from_numpy(a.reshape(-1))


def seq_scatter_add_f32(n, k, numClass, x, a, b, offX, offA, offB):
    """PhpMath::scatterAdd2 (:1379-1393) evaluated in float32: what a float32 buffer accumulates to."""
    out = a.copy()
    for j in range(n):
        idx = min(int(x[offX + j]), numClass - 1)
        out[offA + idx * k: offA + (idx + 1) * k] = (out[offA + idx * k: offA + (idx + 1) * k] + b[offB + j * k: offB + (j + 1) * k]).astype(a.dtype)
    return out


CASES = [(1, 1, 1), (7, 3, 5), (64, 128, 1000), (1000, 36, 17), (4096, 256, 30000), (333, 1, 50), (50, 1000, 8), (20000, 8, 3)]


@pytest.mark.parametrize("n,k,numClass", CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize("dt,xdt", [(F32, I32), (F32, I64), (F64, I32), (I32, I32), (I64, I64), (U8, I32), (F32, F32)])
def test_gather_family_vs_oracle(n, k, numClass, dt, xdt, cuda_service):
    rng = np.random.default_rng(n * 31 + k + numClass)
    npdt, xnp = NP[dt], NP[xdt]
    offX, offA, offB = 3, 4, 8
    x = rng.integers(0, numClass, offX + n + 1).astype(xnp)
    if n > 4:
        x[offX + 1] = x[offX + 3]                                # a duplicate index
        x[offX + 2] = numClass + 5 if numClass < 120 or xdt != dtypes.uint8 else x[offX + 2]   # clamped to numClass-1
    if dt in (F32, F64):
        a = rng.standard_normal(offA + numClass * k + 2).astype(npdt)
        b = rng.standard_normal(offB + n * k + 2).astype(npdt)
    else:
        a = rng.integers(0, 10, offA + numClass * k + 2).astype(npdt)
        b = rng.integers(0, 10, offB + n * k + 2).astype(npdt)
    math = cuda_service.math()
    xo = oracle.as_buffer(x)
    for reverse, add in ((False, False), (False, True), (True, False), (True, True)):
        X, A, B = buf(x, xdt), buf(a, dt), buf(b, dt)
        math.gather(reverse, add, n, k, numClass, X, offX, A, offA, B, offB)
        ra, rb_ = oracle.as_buffer(a), oracle.as_buffer(b)
        oracle.gather(reverse, add, n, k, numClass, xo, offX, ra, offA, rb_, offB)
        ga, gb = A.to_numpy(), B.to_numpy()
        if not add or dt not in (F32, F64):
            assert np.array_equal(ga, ra.astype(npdt)), (reverse, add)
            assert np.array_equal(gb, rb_.astype(npdt)), (reverse, add)
        elif not reverse:
            assert np.array_equal(gb, rb_.astype(npdt))           # one add per element: exact after rounding once
            assert np.array_equal(ga, a)
        else:
            assert np.array_equal(gb, b)
            # float32 running sums against the double result: eps * (terms per class) * |partial sums| (a few units)
            terms = int(np.bincount(np.minimum(x[offX:offX + n].astype(np.int64), numClass - 1)).max()) + 1
            assert np.allclose(ga.astype(np.float64), ra, rtol=1e-6 if dt == F32 else 1e-14,
                               atol=(1.2e-7 if dt == F32 else 2.3e-16) * terms * 8)
            if dt == F32 and n * k <= 200000:
                assert np.array_equal(ga, seq_scatter_add_f32(n, k, numClass, x, a, b, offX, offA, offB))
        assert np.array_equal(X.to_numpy(), x)


@pytest.mark.parametrize("m,n,numClass", [(1, 1, 1), (4, 1, 3), (3, 4, 2), (1000, 1, 10), (64, 128, 50), (5, 3000, 7)])
@pytest.mark.parametrize("dt", [F32, I32])
def test_reduce_gather_vs_oracle(m, n, numClass, dt, cuda_service):
    rng = np.random.default_rng(m + n)
    npdt = NP[dt]
    x = rng.integers(0, numClass, 2 + m * n).astype(np.int32)
    a = rng.integers(-9, 9, 1 + m * numClass * n).astype(npdt)
    b = rng.integers(-9, 9, 3 + m * n).astype(npdt)
    math = cuda_service.math()
    for reverse, add in ((False, False), (False, True), (True, False), (True, True)):
        X, A, B = buf(x, I32), buf(a, dt), buf(b, dt)
        math.reduceGather(reverse, add, m, n, numClass, X, 2, A, 1, B, 3)
        ra, rb_ = oracle.as_buffer(a), oracle.as_buffer(b)
        oracle.reduce_gather(reverse, add, m, n, numClass, oracle.as_buffer(x), 2, ra, 1, rb_, 3)
        assert np.array_equal(A.to_numpy(), ra.astype(npdt)) and np.array_equal(B.to_numpy(), rb_.astype(npdt))


def test_scatter_add_is_deterministic_with_heavy_duplicates(cuda_service):
    """every row lands in one of 3 classes: the sums are long and order sensitive; two runs must agree bit for bit and
    match the sequential float32 evaluation"""
    rng = np.random.default_rng(77)
    n, k, numClass = 3000, 40, 3
    x = rng.integers(0, numClass, n).astype(np.int32)
    a = rng.standard_normal(numClass * k).astype(np.float32)
    b = (rng.standard_normal(n * k) * 1000).astype(np.float32)
    math = cuda_service.math()
    outs = []
    for _ in range(2):
        X, A, B = buf(x, I32), buf(a, F32), buf(b, F32)
        math.gather(True, True, n, k, numClass, X, 0, A, 0, B, 0)
        outs.append(A.to_numpy())
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], seq_scatter_add_f32(n, k, numClass, x, a, b, 0, 0, 0))


def test_embedding_round_trip(cuda_service, oracle_service):
    """LinearAlgebra::gather (embedding lookup) and scatterAdd (its gradient) at a realistic size, against the same
    host class over the oracle."""
    la, lo = LinearAlgebra(cuda_service), LinearAlgebra(oracle_service)
    rng = np.random.default_rng(5)
    vocab, dim, batch, seq = 5000, 64, 16, 50
    table = rng.standard_normal((vocab, dim)).astype(np.float32)
    ids = rng.integers(0, vocab, (batch, seq)).astype(np.int32)
    out = la.gather(la.array(table), la.array(ids, dtype=I32))
    assert out.shape() == [batch, seq, dim]
    assert np.array_equal(out.to_numpy(), table[ids])
    grad = rng.standard_normal((batch, seq, dim)).astype(np.float32)
    dT, dTo = la.zeros(la.alloc([vocab, dim])), lo.zeros(lo.alloc([vocab, dim]))
    la.scatterAdd(la.array(ids, dtype=I32), la.array(grad), dT)
    lo.scatterAdd(lo.array(ids, dtype=I32), lo.array(grad), dTo)
    ref = np.asarray(dTo.toArray(), np.float64)
    assert np.allclose(dT.to_numpy(), ref, rtol=1e-6, atol=1e-6)
    with pytest.raises(InvalidArgumentException, match="Unmatch Shape:"):
        la.gather(la.array(table), la.array(ids, dtype=I32), axis=1)
