"""MatrixOperator::sum / asum / max / argMax / amax / argAmax / min / argMin / amin / argAmin / mean, with and without
axis (src/MatrixOperator.php:824-1023).  The expected values are the reference's own
(tests/RindowTest/Math/Matrix/MatrixOperatorTest.php:523-1010).  The CPU run drives the mirror over the oracle service
(host logic: axis -> (m, n, k), output shapes and dtypes, the kernel route against the reference's per-element walk);
the `gpu` run is the same table over the B200 driver.
"""
import numpy as np
import pytest

import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import MatrixOperator, dtypes

F32 = dtypes.float32
A3 = [[[1, 10], [100, 1000]], [[10000, 100000], [1000000, 10000000]]]
S3 = [[[1, -10], [100, -1000]], [[10000, -100000], [1000000, -10000000]]]
M2 = [[1, 2, 3], [4, -5, -6]]
M3 = [[[1, -2], [-3, 4]], [[-5, 6], [7, -8]]]
G3 = [[[1, -2], [-3, 4]], [[5, 6], [-7, -8]]]
H3 = [[[-2, 1], [8, -7]], [[3, 4], [-5, -6]]]

# (method, input, axis, expected) -- file:line of the reference assertion
AXIS_CASES = [
    ("sum", A3, 0, [[10001, 100010], [1000100, 10001000]]),          # :549-551
    ("sum", A3, 1, [[101, 1010], [1010000, 10100000]]),              # :553-555
    ("sum", A3, 2, [[11, 1100], [110000, 11000000]]),                # :557-559
    ("asum", S3, 0, [[10001, 100010], [1000100, 10001000]]),         # :587-589
    ("asum", S3, 1, [[101, 1010], [1010000, 10100000]]),             # :591-593
    ("asum", S3, 2, [[11, 1100], [110000, 11000000]]),               # :595-597
    ("max", M2, 0, [4, 2, 3]), ("max", M2, 1, [3, 4]),               # :624-633
    ("max", M3, 0, [[1, 6], [7, 4]]), ("max", M3, 1, [[1, 4], [7, 6]]), ("max", M3, 2, [[1, 4], [6, 7]]),   # :651-664
    ("argMax", M2, 0, [1, 0, 0]), ("argMax", M2, 1, [2, 0]),         # :692-700
    ("argMax", G3, 0, [[1, 1], [0, 0]]), ("argMax", G3, 1, [[0, 1], [0, 0]]), ("argMax", G3, 2, [[0, 1], [1, 0]]),  # :716-729
    ("amax", M2, 0, [4, -5, -6]), ("amax", M2, 1, [3, -6]),          # :757-765
    ("amax", M3, 0, [[-5, 6], [7, -8]]), ("amax", M3, 1, [[-3, 4], [7, -8]]), ("amax", M3, 2, [[-2, 4], [6, -8]]),  # :767-780
    ("argAmax", M2, 0, [1, 1, 1]), ("argAmax", M2, 1, [2, 2]),       # :801-808
    ("argAmax", H3, 0, [[1, 1], [0, 0]]), ("argAmax", H3, 1, [[1, 1], [1, 1]]), ("argAmax", H3, 2, [[0, 0], [1, 1]]),  # :809-820
    ("min", M2, 0, [1, -5, -6]), ("min", M2, 1, [1, -6]),            # :842-849
    ("min", M3, 0, [[-5, -2], [-3, -8]]), ("min", M3, 1, [[-3, -2], [-5, -8]]), ("min", M3, 2, [[-2, -3], [-5, -8]]),  # :850-861
    ("argMin", M2, 0, [0, 1, 1]), ("argMin", M2, 1, [0, 2]),         # :883-890
    ("argMin", G3, 0, [[0, 0], [1, 1]]), ("argMin", G3, 1, [[1, 0], [1, 1]]), ("argMin", G3, 2, [[1, 0], [0, 1]]),  # :902-913
    ("amin", M2, 0, [1, 2, 3]), ("amin", M2, 1, [1, 4]),             # :935-942
    ("amin", M3, 0, [[1, -2], [-3, 4]]), ("amin", M3, 1, [[1, -2], [-5, 6]]), ("amin", M3, 2, [[1, -3], [-5, 7]]),  # :943-954
]
FLAT = [1, 2, 3, 4, -5, -6]
PURE_CASES = [("sum", FLAT, -1), ("asum", FLAT, 21), ("max", FLAT, 4), ("argMax", FLAT, 3), ("amax", FLAT, -6),
              ("argAmax", [1, 2, 3, 4, -5, -4], 4), ("min", FLAT, -6), ("argMin", FLAT, 5), ("amin", FLAT, 1),
              ("argAmin", FLAT, 0)]                                   # :527-540 ... :960-975


def check_axis_cases(svc):
    mo = MatrixOperator(service=svc)
    for method, data, axis, want in AXIS_CASES:
        X = mo.array(data, dtype=F32)
        got = getattr(mo, method)(X, axis=axis)
        assert got.toArray() == want, (method, axis, got.toArray(), want)
        assert got.dtype() == (dtypes.int32 if method.startswith("arg") else F32), method
        # views with an element offset (:635-648, :702-713): the same slab behind a block of other values
        pad = np.full((1,) + np.asarray(data).shape, 9.0)
        V = mo.array(np.concatenate([pad, np.asarray(data, dtype=np.float64)[None]]), dtype=F32)[1]
        assert V.offset() == np.asarray(data).size
        assert getattr(mo, method)(V, axis=axis).toArray() == want, (method, axis, "view")


def check_pure_cases(svc):
    mo = MatrixOperator(service=svc)
    for method, data, want in PURE_CASES:
        assert getattr(mo, method)(mo.array(data, dtype=F32)) == want, method
        assert getattr(mo, method)(mo.array([data[:3], data[3:]], dtype=F32)) == want, method
        A = mo.array([[0, 0, 0, 0, 0, 0], data], dtype=F32)[1]
        assert A.offset() == 6 and getattr(mo, method)(A) == want, method


def check_kernel_route_against_walk(svc, sizes=((5, 7, 3), (4, 33), (2, 3, 4, 5))):
    """sum / mean / max / argMax / min / argMin by reduce kernels == the reference's per-element walk, ties included."""
    mo = MatrixOperator(service=svc)
    rng = np.random.default_rng(7)
    for shape in sizes:
        data = np.round(rng.uniform(-4, 4, shape) * 2) / 2            # halves: many ties, sums exact in float32
        X = mo.array(data, dtype=F32)
        for axis in range(len(shape)):
            math = mo.getMath(F32)
            walks = {
                "sum": lambda N, XX, off, pos, inc: math.sum(N, XX, off + pos, inc),
                "mean": lambda N, XX, off, pos, inc: math.sum(N, XX, off + pos, inc) / N,
                "max": lambda N, XX, off, pos, inc: XX[off + pos + math.imax(N, XX, off + pos, inc) * inc],
                "min": lambda N, XX, off, pos, inc: XX[off + pos + math.imin(N, XX, off + pos, inc) * inc],
                "argMax": lambda N, XX, off, pos, inc: math.imax(N, XX, off + pos, inc),
                "argMin": lambda N, XX, off, pos, inc: math.imin(N, XX, off + pos, inc),
            }
            for method, walk in walks.items():
                fast = getattr(mo, method)(X, axis=axis)
                slow = mo.walkAxis(walk, X, axis, dtypes.int32 if method.startswith("arg") else None)
                assert fast.shape() == slow.shape() == list(shape[:axis] + shape[axis + 1:])
                if method == "mean":
                    np.testing.assert_allclose(fast.to_numpy(), slow.to_numpy(), rtol=3e-7, atol=0)
                else:
                    assert fast.toArray() == slow.toArray(), (method, shape, axis)
            np.testing.assert_array_equal(mo.argMax(X, axis=axis).to_numpy(), np.argmax(data, axis=axis))
            np.testing.assert_array_equal(mo.argMin(X, axis=axis).to_numpy(), np.argmin(data, axis=axis))


def check_nan_rule(svc):
    """max(axis) follows math_imax (PhpMath.php:114-130): NaN wins only from position 0 -- not the sticky rule."""
    mo = MatrixOperator(service=svc)
    nan = float("nan")
    X = mo.array([[1, nan, 3], [nan, 2, 1], [2, 5, nan]], dtype=F32)
    got = mo.max(X, axis=1).to_numpy()
    assert got[0] == 3 and np.isnan(got[1]) and got[2] == 5
    assert mo.argMax(X, axis=1).toArray() == [2, 0, 1]
    got = mo.min(X, axis=1).to_numpy()
    assert got[0] == 1 and np.isnan(got[1]) and got[2] == 2
    assert mo.argMin(X, axis=1).toArray() == [0, 0, 0]


def check_errors(svc):
    mo = MatrixOperator(service=svc)
    X = mo.array(M2, dtype=F32)
    for method in ("sum", "max", "argMax", "min", "argMin", "mean", "asum", "amax"):
        for axis in (2, -1):                                          # MatrixOperator.php:838-840: no negative axes
            with pytest.raises(rb.InvalidArgumentException, match=r"Invalid axis: axis=%d,shape=\[2,3\]" % axis):
                getattr(mo, method)(X, axis=axis)


def test_axis_cases_host_logic(oracle_service):
    check_axis_cases(oracle_service)


def test_pure_cases_host_logic(oracle_service):
    check_pure_cases(oracle_service)


def test_kernel_route_equals_the_reference_walk_host_logic(oracle_service):
    check_kernel_route_against_walk(oracle_service)


def test_nan_rule_host_logic(oracle_service):
    check_nan_rule(oracle_service)


def test_axis_errors_host_logic(oracle_service):
    check_errors(oracle_service)


def test_bool_sum_and_integer_walk(oracle_service):
    mo = MatrixOperator(service=oracle_service)
    B = mo.array([[True, False, True], [True, True, False]], dtype=dtypes.bool_)
    assert mo.sum(B) == 4                                             # :877-883, axis=None: a plain count
    I = mo.array([[1, 2, 3], [4, -5, -6]], dtype=dtypes.int32)
    assert mo.sum(I, axis=0).toArray() == [5, -3, -3] and mo.sum(I, axis=0).dtype() == dtypes.int32
    assert mo.argMax(I, axis=1).toArray() == [2, 0]


@pytest.mark.gpu
def test_axis_cases_gpu(cuda_service):
    check_axis_cases(cuda_service)


@pytest.mark.gpu
def test_pure_cases_gpu(cuda_service):
    check_pure_cases(cuda_service)


@pytest.mark.gpu
def test_kernel_route_equals_the_reference_walk_gpu(cuda_service):
    check_kernel_route_against_walk(cuda_service, sizes=((5, 7, 3), (4, 33), (2, 3, 4, 5), (3, 700, 9)))


@pytest.mark.gpu
def test_nan_rule_gpu(cuda_service):
    check_nan_rule(cuda_service)


@pytest.mark.gpu
def test_axis_errors_gpu(cuda_service):
    check_errors(cuda_service)


@pytest.mark.gpu
def test_integer_arrays_use_the_device_scalar_calls_gpu(cuda_service):
    """No LV_BASIC driver in this package: the walk asks the device driver, whose scalar entry points take every real
    dtype (strided: incX = the axis step)."""
    mo = MatrixOperator(service=cuda_service)
    I = mo.array([[1, 2, 3], [4, -5, -6]], dtype=dtypes.int32)
    assert mo.sum(I, axis=0).toArray() == [5, -3, -3] and mo.sum(I, axis=1).toArray() == [6, -7]
    assert mo.sum(I) == -1 and mo.max(I) == 4 and mo.argMin(I) == 5
    assert mo.argMax(I, axis=1).toArray() == [2, 0] and mo.min(I, axis=0).toArray() == [1, -5, -6]
    B = mo.array([[True, False, True], [True, True, False]], dtype=dtypes.bool_)
    assert mo.sum(B) == 4


# ---- the thin facade methods next to them: one or two driver calls each (src/MatrixOperator.php:342-427, 587-661) ----

def check_facade(svc):
    mo = MatrixOperator(service=svc)
    a, b = [1, 2, 3, 4, 5, 6], [3, 4, 5, 6, 7, 8]
    A1, B1 = mo.array(a, dtype=F32), mo.array(b, dtype=F32)
    A2, B2 = mo.array([a[:3], a[3:]], dtype=F32), mo.array([b[:3], b[3:]], dtype=F32)
    AV = mo.array([[0] * 6, a], dtype=F32)[1]
    BV = mo.array([[0] * 6, b], dtype=F32)[1]
    assert AV.offset() == 6 and BV.offset() == 6
    # MatrixOperatorTest.php:454-476, 478-500, 502-521
    assert mo.dot(A1, B1) == 133 and mo.dot(A2, B2) == 133 and mo.dot(AV, BV) == 133
    assert mo.add(A1, B1).toArray() == [4, 6, 8, 10, 12, 14]
    assert mo.add(A2, B2).toArray() == [[4, 6, 8], [10, 12, 14]] and mo.add(AV, BV).toArray() == [4, 6, 8, 10, 12, 14]
    assert mo.scale(7, A1).toArray() == [7, 14, 21, 28, 35, 42] and mo.scale(7, AV).toArray() == [7, 14, 21, 28, 35, 42]
    assert mo.scale(7, A2).toArray() == [[7, 14, 21], [28, 35, 42]] and A2.toArray() == [[1, 2, 3], [4, 5, 6]]
    # :113-135 copy: a new buffer, views copied from their offset
    c = mo.copy(A2)
    assert c.toArray() == [[1, 2, 3], [4, 5, 6]] and c.buffer() is not A2.buffer() and c.dtype() == F32
    c = mo.copy(A2[1])
    assert c.toArray() == [4, 5, 6] and c.offset() == 0 and len(c.buffer()) == 3
    # :403-452 transpose
    assert mo.transpose(A1).toArray() == a and mo.transpose(A2).toArray() == [[1, 4], [2, 5], [3, 6]]
    t3 = mo.transpose(mo.array(np.arange(1, 25).reshape(4, 3, 2)))
    assert t3.toArray() == [[[1, 7, 13, 19], [3, 9, 15, 21], [5, 11, 17, 23]], [[2, 8, 14, 20], [4, 10, 16, 22], [6, 12, 18, 24]]]
    assert mo.transpose(mo.array([[[0] * 3] * 2, [a[:3], a[3:]]], dtype=F32)[1]).toArray() == [[1, 4], [2, 5], [3, 6]]
    # :55-72 full, :368-376 zerosLike / fullLike, :98-111 arange dtypes, :1536-1550 astype
    f = mo.full([2, 3], 1.0, dtypes.float64)
    assert f.dtype() == dtypes.float64 and f.toArray() == [[1.0] * 3] * 2
    f = mo.full([2, 3], 1, dtypes.int64)
    assert f.dtype() == dtypes.int64 and f.toArray() == [[1] * 3] * 2
    f = mo.full([2, 3], True, dtypes.bool_)
    assert f.dtype() == dtypes.bool_ and f.toArray() == [[True] * 3] * 2
    assert mo.zerosLike(A2).toArray() == [[0] * 3] * 2 and mo.fullLike(A2, 5).toArray() == [[5] * 3] * 2
    assert mo.arange(5).dtype() == dtypes.int32 and mo.arange(5, 1.0).dtype() == F32
    assert mo.arange(5, 1, 2).toArray() == [1, 3, 5, 7, 9] and mo.arange(5, 1.0, 0.5).toArray() == [1.0, 1.5, 2.0, 2.5, 3.0]
    Y = mo.astype(mo.array([[100, 101, 102], [103, 104, 105]], dtype=dtypes.int32), dtypes.float64)
    assert Y.dtype() == dtypes.float64 and Y.toArray() == [[100, 101, 102], [103, 104, 105]]
    with pytest.raises(rb.InvalidArgumentException, match=r"Unmatch shape of dimension to add: \(6\),\(2,3\)"):
        mo.add(A1, B2)
    with pytest.raises(rb.InvalidArgumentException, match="unmatch shape of dimension"):
        mo.dot(A1, B2)
    with pytest.raises(rb.InvalidArgumentException, match="the scalar must be a numeric"):
        mo.scale("7", A1)
    with pytest.raises(rb.InvalidArgumentException, match="value must be scalar"):
        mo.full([2], [1, 2])


def test_facade_host_logic(oracle_service):
    check_facade(oracle_service)


@pytest.mark.gpu
def test_facade_gpu(cuda_service):
    check_facade(cuda_service)
