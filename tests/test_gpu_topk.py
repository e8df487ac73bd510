"""Parity of topK (rb200_topk) through the C ABI against the CPU oracle's restatement of the reference's min-heap
(PhpMath.php:933-1058): values AND indices bit-exact, including ties, NaN, the unsorted (heap) order and the
heap-sort order.  Known answers from LinearAlgebraTest.php testTopKNormal10.  pytest -m gpu"""
import numpy as np
import pytest

import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, dtypes

pytestmark = pytest.mark.gpu


def test_topk_known_answers(cuda_service):
    """LinearAlgebraTest.php:8027-8115"""
    la = LinearAlgebra(cuda_service)
    x = la.array(np.arange(10, dtype=np.float32))
    values, indices = la.topK(x, k=3)
    assert values.toArray() == [9, 8, 7] and indices.toArray() == [9, 8, 7]
    assert values.dtype() == dtypes.float32 and indices.dtype() == dtypes.int32
    values, indices = la.topK(x)
    assert values.toArray() == [9] and indices.toArray() == [9]
    x = la.array(np.arange(30, dtype=np.float32).reshape(3, 10))
    values, indices = la.topK(x, k=3)
    assert values.toArray() == [[9, 8, 7], [19, 18, 17], [29, 28, 27]]
    assert indices.toArray() == [[9, 8, 7]] * 3
    with pytest.raises(InvalidArgumentException, match="k must be less than or equal"):
        la.topK(x, k=11)
    with pytest.raises(InvalidArgumentException, match='Unmatch shape of "values"'):
        la.topK(x, k=3, values=la.alloc([3, 4], dtypes.float32))
    with pytest.raises(InvalidArgumentException, match='Unmatch dtype of "indices"'):
        la.topK(x, k=3, indices=la.alloc([3, 3], dtypes.int64))


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (1, 10, 10), (3, 10, 3), (7, 33, 5), (64, 5000, 10), (5, 5000, 1), (2, 40000, 100),
                                   (300, 257, 64), (1, 3000, 1500),
                                   (6, 20000, 9000)])           # k beyond the shared-memory heaps: global-memory heaps
@pytest.mark.parametrize("sorted_", [True, False])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32])
def test_topk_random_vs_oracle(m, n, k, sorted_, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(n + k)
    off_in, off_v, off_i = 3, 1, 2
    if dt == dtypes.int32:
        x = rng.integers(-50, 50, off_in + m * n).astype(npdt)                      # heavy ties
    else:
        x = (np.round(rng.standard_normal(off_in + m * n) * 4) / 4).astype(npdt)      # ties
        if n > 8:
            x[off_in + rng.integers(0, m * n, max(1, m * n // 40))] = np.nan
            x[off_in + rng.integers(0, m * n, max(1, m * n // 40))] = np.inf
            x[off_in + rng.integers(0, m * n, max(1, m * n // 40))] = -np.inf
    v = np.full(off_v + m * k, 7).astype(npdt)
    ix = np.full(off_i + m * k, -3).astype(np.int32)
    X, V, IX = CudaBuffer(x.size, dt).from_numpy(x), CudaBuffer(v.size, dt).from_numpy(v), CudaBuffer(ix.size, dtypes.int32).from_numpy(ix)
    cuda_service.math().topk(m, n, X, off_in, k, sorted_, V, off_v, IX, off_i)
    vo, io = v.astype(np.float64), ix.astype(np.float64)
    oracle.topk(m, n, x.astype(np.float64), off_in, k, sorted_, vo, off_v, io, off_i)
    assert np.array_equal(V.to_numpy().astype(np.float64), vo, equal_nan=True)
    assert np.array_equal(IX.to_numpy().astype(np.float64), io)
    assert np.array_equal(X.to_numpy(), x, equal_nan=True)


def test_topk_k_above_n_writes_nothing_and_int64_indices(cuda_service):
    m = cuda_service.math()
    x = np.arange(6, dtype=np.float32)
    X = CudaBuffer(6, dtypes.float32).from_numpy(x)
    V = CudaBuffer(8, dtypes.float32).from_numpy(np.full(8, 5, np.float32))
    IX = CudaBuffer(8, dtypes.int64).from_numpy(np.full(8, -1, np.int64))
    m.topk(1, 6, X, 0, 8, True, V, 0, IX, 0)                                        # PhpMath.php:1045: returns
    assert V.to_numpy().tolist() == [5] * 8 and IX.to_numpy().tolist() == [-1] * 8
    m.topk(2, 3, X, 0, 2, True, V, 0, IX, 0)
    assert V.to_numpy()[:4].tolist() == [2, 1, 5, 4] and IX.to_numpy()[:4].tolist() == [2, 1, 2, 1]
    with pytest.raises(InvalidArgumentException, match="integers"):
        m.topk(1, 6, X, 0, 2, True, V, 0, CudaBuffer(8, dtypes.float32), 0)
    with pytest.raises(InvalidArgumentException, match="bufferinput"):
        m.topk(2, 6, X, 0, 2, True, V, 0, IX, 0)
