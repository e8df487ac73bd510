"""Parity of cumsum / cumsumb / searchsorted (rb200_cumsumb, rb200_cumsum, rb200_searchsorted) through the C ABI against
the CPU oracle on the same seeded inputs.  pytest -m gpu

Tolerances:
  * integer dtypes and searchsorted: bit-exact;
  * floating cumsum: the device adds in tree order, the reference serially; both stay within
    n * eps * (running sum of |x|) of the exact prefix, asserted as |d| <= 2 * n * eps * cumsum|x| (eps of the dtype)
    -- data rounded to 1/8 are also checked bit-exact, where every partial sum is representable;
  * NaN / INF reach exactly the positions they reach in the reference.
"""
import numpy as np
import pytest

import oracle
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, dtypes

pytestmark = pytest.mark.gpu

F32, F64, I32, I64 = dtypes.float32, dtypes.float64, dtypes.int32, dtypes.int64


def _cumsumb_both(service, m, n, k, a, dt, exclusive, reverse, off=0):
    A = CudaBuffer(a.size, dt).from_numpy(a)
    B = CudaBuffer(a.size, dt).from_numpy(np.zeros(a.size, a.dtype))
    service.math().cumsumb(m, n, k, A, off, exclusive, reverse, B, off)
    bo = np.zeros(a.size)
    oracle.cumsumb(m, n, k, a.astype(np.float64), off, exclusive, reverse, dt == F32, bo, off)
    return B.to_numpy(), bo


SHAPES = [(1, 1, 1), (1, 5, 1), (3, 7, 1), (1, 255, 1), (1, 256, 1), (1, 257, 1), (2, 1000, 1), (1, 100000, 1), (700, 33, 1),
          (1, 4, 1), (5, 512, 1), (3, 516, 1), (2, 1028, 1), (9000, 64, 1), (1, 300000, 1),
          (1, 3, 2), (4, 9, 5), (1, 300, 3), (2, 64, 33), (1, 4097, 64), (3, 17, 300), (1, 1, 9), (5000, 4, 2), (1, 20000, 2)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("exclusive,reverse", [(False, False), (True, False), (False, True), (True, True)])
def test_cumsumb_exact_data_vs_oracle(m, n, k, exclusive, reverse, cuda_service):
    """eighths in [-4, 4]: every partial sum is exact in float32 up to n = 100000, so any order of addition agrees"""
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    for dt in (F32, F64, I32, I64):
        npdt = dtypes.NUMPY[dt]
        for off in (2, 4):                                  # 4: 16-byte aligned rows take the vector kernels
            if dt in (F32, F64):
                a = (np.round(rng.uniform(-4, 4, off + m * n * k) * 8) / 8).astype(npdt)
            else:
                a = rng.integers(-1000, 1000, off + m * n * k).astype(npdt)
            got, want = _cumsumb_both(cuda_service, m, n, k, a, dt, exclusive, reverse, off=off)
            assert np.array_equal(got.astype(np.float64), want), (dt, off, np.flatnonzero(got.astype(np.float64) != want)[:5])


@pytest.mark.parametrize("m,n,k", [(1, 100000, 1), (3, 5000, 1), (1, 30000, 4), (64, 777, 40), (2000, 50, 1)])
@pytest.mark.parametrize("dt", [F32, F64])
def test_cumsumb_random_within_rounding_bound(m, n, k, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(n + k)
    a = rng.standard_normal(m * n * k).astype(npdt)
    for exclusive, reverse in ((False, False), (True, True)):
        got, want = _cumsumb_both(cuda_service, m, n, k, a, dt, exclusive, reverse)
        x = a.astype(np.float64).reshape(m, n, k)
        if reverse:
            x = x[:, ::-1]
        exact = np.cumsum(x, axis=1)
        bound = np.cumsum(np.abs(x), axis=1)
        if exclusive:
            exact = np.concatenate([np.zeros((m, 1, k)), exact[:, :-1]], axis=1)
            bound = np.concatenate([np.zeros((m, 1, k)), bound[:, :-1]], axis=1)
        if reverse:
            exact, bound = exact[:, ::-1], bound[:, ::-1]
        eps = float(np.finfo(npdt).eps)
        tol = 2 * n * eps * bound.reshape(-1) + 1e-300
        assert np.all(np.abs(got.astype(np.float64) - exact.reshape(-1)) <= tol)
        assert np.all(np.abs(want - exact.reshape(-1)) <= tol)          # the float32-storage oracle obeys the same bound
        # the reference's own semantics: PhpBuffer holds doubles, so the PHP driver's running sum is never rounded
        # between steps -- the unrounded oracle, same bound
        php = np.zeros(a.size)
        oracle.cumsumb(m, n, k, a.astype(np.float64), 0, exclusive, reverse, False, php, 0)
        assert np.all(np.abs(got.astype(np.float64) - php) <= tol)


@pytest.mark.parametrize("n,k", [(1000, 1), (5000, 3), (40, 64)])
def test_cumsumb_nonfinite_reach(n, k, cuda_service):
    rng = np.random.default_rng(n)
    a = (np.round(rng.uniform(-4, 4, n * k) * 8) / 8).astype(np.float32)
    a[(n // 3) * k] = np.nan
    a[(n // 2) * k + k - 1] = np.inf
    for reverse in (False, True):
        got, want = _cumsumb_both(cuda_service, 1, n, k, a, F32, False, reverse)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        assert np.array_equal(got[fin].astype(np.float64), want[fin])


@pytest.mark.parametrize("n,incX,incY", [(1, 1, 1), (9, 1, 1), (1000, 1, 1), (777, 2, 3), (5000, 3, 1), (70000, 1, 2)])
@pytest.mark.parametrize("exclusive,reverse", [(False, False), (True, False), (False, True), (True, True)])
def test_cumsum_1d_vs_oracle(n, incX, incY, exclusive, reverse, cuda_service):
    """PhpMath::cumsum: reverse walks X forwards and only turns Y around (PhpMath.php:1835-1842)"""
    rng = np.random.default_rng(n)
    offX, offY = 1, 2
    for dt in (F32, F64, I32):
        npdt = dtypes.NUMPY[dt]
        x = (np.round(rng.uniform(-4, 4, offX + (n - 1) * incX + 1) * 8) / 8 if dt != I32
             else rng.integers(-99, 99, offX + (n - 1) * incX + 1)).astype(npdt)
        y = np.full(offY + (n - 1) * incY + 1, 77).astype(npdt)
        X, Y = CudaBuffer(x.size, dt).from_numpy(x), CudaBuffer(y.size, dt).from_numpy(y)
        cuda_service.math().cumsum(n, X, offX, incX, exclusive, reverse, Y, offY, incY)
        yo = y.astype(np.float64)
        oracle.cumsum(n, x.astype(np.float64), offX, incX, exclusive, reverse, yo, offY, incY)
        assert np.array_equal(Y.to_numpy().astype(np.float64), yo)       # untouched gaps included
        assert np.array_equal(X.to_numpy(), x)


def test_cumsum_la_axes_and_outputs(cuda_service):
    la = LinearAlgebra(cuda_service)
    rng = np.random.default_rng(5)
    x = rng.integers(-9, 9, (4, 6, 5, 3)).astype(np.float32)
    for axis in (None, 0, 1, 2, 3, -1, -4):
        for exclusive, reverse in ((False, False), (True, False), (False, True), (True, True)):
            got = la.cumsum(la.array(x), axis=axis, exclusive=exclusive, reverse=reverse).to_numpy().reshape(x.shape)
            ax = 0 if axis is None else axis
            xx = np.flip(x, ax) if reverse else x
            want = np.cumsum(xx, axis=ax)
            if exclusive:
                want = want - xx
            if reverse:
                want = np.flip(want, ax)
            assert np.array_equal(got, want), (axis, exclusive, reverse)
    out = la.alloc([4, 6, 5, 3], dtypes.float32)
    assert la.cumsum(la.array(x), axis=1, outputs=out) is out
    with pytest.raises(InvalidArgumentException, match="Invalid axis"):
        la.cumsum(la.array(x), axis=4)
    with pytest.raises(InvalidArgumentException, match="Unmatch shape"):
        la.cumsum(la.array(x), outputs=la.alloc([4, 6, 5, 2], dtypes.float32))


def test_cumsum_buffer_errors(cuda_service):
    m = cuda_service.math()
    A, B = CudaBuffer(10, F32), CudaBuffer(10, F32)
    with pytest.raises(InvalidArgumentException, match="bufferA"):
        m.cumsumb(1, 11, 1, A, 0, False, False, B, 0)
    with pytest.raises(InvalidArgumentException, match="bufferB"):
        m.cumsumb(1, 5, 2, A, 0, False, False, B, 1)
    with pytest.raises(InvalidArgumentException, match="Unmatch data type"):
        m.cumsumb(1, 5, 2, A, 0, False, False, CudaBuffer(10, F64), 0)
    with pytest.raises(InvalidArgumentException, match="bufferX"):
        m.cumsum(6, A, 0, 2, False, False, B, 0, 1)
    with pytest.raises(InvalidArgumentException, match="bufferY"):
        m.cumsum(6, A, 0, 1, False, False, B, 5, 1)


# ---- searchsorted ----------------------------------------------------------------------------------------------------

def _search_both(service, m, n, a, ldA, x, right, dt, idt, offA=0, offX=0, incX=1, offY=0, incY=1):
    A, X = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(x.size, dt).from_numpy(x)
    y = np.full(offY + (m - 1) * incY + 1, -5).astype(dtypes.NUMPY[idt])
    Y = CudaBuffer(y.size, idt).from_numpy(y)
    service.math().searchsorted(m, n, A, offA, ldA, X, offX, incX, right, Y, offY, incY)
    yo = y.astype(np.float64)
    oracle.searchsorted(m, n, a.astype(np.float64), offA, ldA, x.astype(np.float64), offX, incX, right, yo, offY, incY)
    return Y.to_numpy().astype(np.float64), yo


@pytest.mark.parametrize("m,n", [(1, 1), (7, 5), (1000, 33), (5000, 1000), (3, 100000), (100000, 17)])
@pytest.mark.parametrize("right", [False, True])
@pytest.mark.parametrize("dt", [F32, F64])
def test_searchsorted_shared_sorted_vs_oracle(m, n, right, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(m + n)
    a = np.sort(np.round(rng.uniform(-50, 50, n) * 4) / 4).astype(npdt)         # plenty of ties
    x = (np.round(rng.uniform(-60, 60, m) * 4) / 4).astype(npdt)
    x[rng.integers(0, m, max(1, m // 50))] = np.nan
    x[rng.integers(0, m, max(1, m // 50))] = np.inf
    x[rng.integers(0, m, max(1, m // 50))] = -np.inf
    for idt in (I32, I64):
        got, want = _search_both(cuda_service, m, n, a, 0, x, right, dt, idt)
        assert np.array_equal(got, want)
    fin = ~np.isnan(x)
    assert np.array_equal(got[fin], np.searchsorted(a, x[fin], side="right" if right else "left"))


@pytest.mark.parametrize("m,n", [(1, 4), (50, 9), (4000, 257), (20, 5000)])
@pytest.mark.parametrize("right", [False, True])
def test_searchsorted_individual_rows_mixed_vs_oracle(m, n, right, cuda_service):
    """2-D A: row i pairs with X[i]; every third row is unsorted and some hold NaN -- the walk of the reference stops
    at the first element the value is not above, which the device reproduces row by row"""
    rng = np.random.default_rng(n)
    a = np.sort(np.round(rng.uniform(-20, 20, (m, n)) * 2) / 2, axis=1).astype(np.float32)
    for i in range(0, m, 3):
        rng.shuffle(a[i])
    for i in range(1, m, 7):
        a[i, rng.integers(0, n)] = np.nan
    x = (np.round(rng.uniform(-25, 25, m) * 2) / 2).astype(np.float32)
    got, want = _search_both(cuda_service, m, n, a.reshape(-1), n, x, right, F32, I32)
    assert np.array_equal(got, want)


def test_searchsorted_offsets_increments_and_gaps(cuda_service):
    rng = np.random.default_rng(3)
    m, n = 40, 25
    a = np.concatenate([[99.0, 99.0], np.sort(rng.uniform(0, 1, n))]).astype(np.float64)
    x = rng.uniform(-0.2, 1.2, 1 + (m - 1) * 2 + 1)
    got, want = _search_both(cuda_service, m, n, a, 0, x, False, F64, I64, offA=2, offX=1, incX=2, offY=3, incY=3)
    assert np.array_equal(got, want)                       # the -5 gaps of Y survive


def test_searchsorted_la_and_errors(cuda_service):
    la = LinearAlgebra(cuda_service)
    A = la.array(np.array([0.1, 0.3, 0.5, 0.7, 0.9], np.float32))
    X = la.array(np.array([[0.0, 0.5], [1.0, 0.3]], np.float32))
    Y = la.searchsorted(A, X)
    assert Y.dtype() == dtypes.int32 and Y.shape() == [2, 2] and Y.toArray() == [[0, 2], [5, 1]]
    Y = la.searchsorted(A, X, right=True, dtype=dtypes.int64)
    assert Y.dtype() == dtypes.int64 and Y.toArray() == [[0, 3], [5, 2]]
    with pytest.raises(InvalidArgumentException, match="1D or 2D"):
        la.searchsorted(la.alloc([2, 2, 2], dtypes.float32), X)
    with pytest.raises(InvalidArgumentException, match="int32 or int64"):
        la.searchsorted(A, X, dtype=dtypes.float32)
    with pytest.raises(InvalidArgumentException, match="Unmatch shape of dimension A,X"):
        la.searchsorted(la.alloc([3, 5], dtypes.float32), X)
    with pytest.raises(InvalidArgumentException, match="Unmatch shape"):
        la.searchsorted(A, X, Y=la.alloc([4], dtypes.int32))
    m = cuda_service.math()
    with pytest.raises(InvalidArgumentException, match="bufferA"):
        m.searchsorted(2, 5, CudaBuffer(9, F32), 0, 5, CudaBuffer(2, F32), 0, 1, False, CudaBuffer(2, I32), 0, 1)
    with pytest.raises(InvalidArgumentException, match="Unmatch data type"):
        m.searchsorted(2, 5, CudaBuffer(10, F32), 0, 5, CudaBuffer(2, F64), 0, 1, False, CudaBuffer(2, I32), 0, 1)
