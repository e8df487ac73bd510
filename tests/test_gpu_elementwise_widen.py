"""Parity of the section 8(f) rank-2/4 widening -- the broadcast family (maximum ... lessEqual, pow, duplicate),
the unary family, equal / notEqual / not, astype, sum / imax / imin, updateAddOnehot, reduceMean -- through the
C ABI against the CPU oracle on the same seeded inputs.  pytest -m gpu

Tolerances (float32 against the oracle's double result):
  * maximum / minimum / comparisons / duplicate / not / isnan / nan2num / equal / notEqual / astype / imax / imin /
    onehot: bit-exact (selection, copy and integer work);
  * increment, square: every intermediate is one correctly rounded float op -> exact for square, and within
    1 ulp of the double result for alpha*x+beta (two roundings instead of one);
  * sqrt, reciprocal, rsqrt: IEEE-correct float ops, <= 2 ulp of the double result;
  * exp, log, tanh, sin, cos, tan, pow: CUDA libm float accuracy, asserted at 4e-6 relative (+1e-7 absolute),
    far inside the reference's isclose bound (rtol 1e-4);
  * sum: double accumulation on the device, |d| <= 1e-12 * sum|x| (order of summation differs from the sequential
    PHP loop only in double rounding).
"""
import numpy as np
import pytest

import oracle
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import CudaBuffer, InvalidArgumentException, LinearAlgebra, RuntimeException, dtypes

pytestmark = pytest.mark.gpu

NP = {k: v for k, v in dtypes.NUMPY.items() if k not in dtypes.COMPLEX_TYPES}     # real dtypes: the arithmetic entry points
F32, F64 = dtypes.float32, dtypes.float64


def buf(a, dt):
    a = np.asarray(a)
    return CudaBuffer(a.size, dt).from_numpy(a.reshape(-1))


def special(rng, shape, npdt, frac=0.02):
    """random values with NaN / +-INF / +-0 / ties sprinkled in"""
    a = (np.round(rng.uniform(-4, 4, shape) * 8) / 8).astype(npdt)
    flat = a.reshape(-1)
    k = max(1, int(flat.size * frac))
    for v in (np.nan, np.inf, -np.inf, 0.0, -0.0):
        flat[rng.integers(0, flat.size, k)] = v
    return a


BC_SHAPES = [(1, 1, 0, 0, 0), (3, 5, 0, 0, 0), (7, 1, 0, 0, 0), (1, 4099, 0, 0, 0), (257, 1023, 5, 3, 2), (64, 4096, 0, 0, 0),
             (1000, 36, 4, 8, 4), (33, 5000, 7, 1, 1), (100003, 1, 0, 0, 0), (1, 1 << 20, 0, 0, 0), (2048, 2048, 0, 4, 0)]
BC_OPS = ["maximum", "minimum", "greater", "greaterEqual", "less", "lessEqual"]
BC_CODE = {"maximum": oracle.BC_MAXIMUM, "minimum": oracle.BC_MINIMUM, "greater": oracle.BC_GREATER,
           "greaterEqual": oracle.BC_GREATER_EQUAL, "less": oracle.BC_LESS, "lessEqual": oracle.BC_LESS_EQUAL}


@pytest.mark.parametrize("m,n,ld_extra,offA,offX", BC_SHAPES, ids=[str(v) for v in BC_SHAPES])
@pytest.mark.parametrize("op", BC_OPS)
@pytest.mark.parametrize("dt", [F32, F64, dtypes.int32, dtypes.int64])
def test_broadcast_compare_family_vs_oracle(m, n, ld_extra, offA, offX, op, dt, cuda_service):
    rng = np.random.default_rng(m * 7 + n)
    npdt = NP[dt]
    ld = n + ld_extra
    sizeA, sizeX = offA + (m - 1) * ld + n + 2, offX + n + 1
    if dt in (F32, F64):
        a, x = special(rng, sizeA, npdt), special(rng, sizeX, npdt)
    else:
        a, x = rng.integers(-5, 5, sizeA).astype(npdt), rng.integers(-5, 5, sizeX).astype(npdt)
    A, X = buf(a, dt), buf(x, dt)
    getattr(cuda_service.math(), op)(m, n, A, offA, ld, X, offX, 1)
    ref = oracle.as_buffer(a)
    oracle.broadcast(BC_CODE[op], False, m, n, ref, offA, ld, oracle.as_buffer(x), offX, 1)
    got = A.to_numpy()
    want = ref.astype(npdt)
    assert np.array_equal(got, want, equal_nan=(dt in (F32, F64)))
    # signed zeros and the untouched gaps between rows are part of "bit-exact"
    if dt in (F32, F64) and op in ("maximum", "minimum"):
        assert np.array_equal(np.signbit(got), np.signbit(want))
    assert np.array_equal(X.to_numpy(), x, equal_nan=True)


@pytest.mark.parametrize("trans", [False, True])
@pytest.mark.parametrize("m,n,ld_extra,offA,offX", BC_SHAPES, ids=[str(v) for v in BC_SHAPES])
@pytest.mark.parametrize("dt", [F32, F64])
def test_pow_duplicate_vs_oracle(m, n, ld_extra, offA, offX, trans, dt, cuda_service):
    rng = np.random.default_rng(m * 11 + n)
    npdt = NP[dt]
    ld = n + ld_extra
    sizeA = offA + (m - 1) * ld + n + 2
    nx = m if trans else n
    a = rng.uniform(0.1, 3.0, sizeA).astype(npdt)
    x = rng.uniform(-2.0, 3.0, offX + nx + 1).astype(npdt)
    math = cuda_service.math()
    # duplicate: bit-exact copy, rows' gaps untouched, NaN payloads kept
    a_dup = a.copy()
    x_dup = x.copy()
    x_dup[offX] = np.nan
    A, X = buf(a_dup, dt), buf(x_dup, dt)
    math.duplicate(trans, m, n, X, offX, 1, A, offA, ld)
    ref = oracle.as_buffer(a_dup)
    oracle.broadcast(oracle.BC_DUPLICATE, trans, m, n, ref, offA, ld, oracle.as_buffer(x_dup), offX, 1)
    assert np.array_equal(A.to_numpy(), ref.astype(npdt), equal_nan=True)
    # pow
    A, X = buf(a, dt), buf(x, dt)
    math.pow(trans, m, n, A, offA, ld, X, offX, 1)
    ref = oracle.as_buffer(a)
    oracle.broadcast(oracle.BC_POW, trans, m, n, ref, offA, ld, oracle.as_buffer(x), offX, 1)
    got = A.to_numpy().astype(np.float64)
    rtol = 4e-6 if dt == F32 else 1e-14
    assert np.all(np.abs(got - ref) <= rtol * np.abs(ref) + 1e-300)


def test_pow_special_values(cuda_service):
    """PHP pow() == C pow(): the IEEE special cases the device libm must reproduce."""
    a = np.array([0.0, -0.0, 2.0, -2.0, -2.0, np.inf, -np.inf, np.nan, 1.0, -8.0, 0.0, 5.0], np.float32)
    x = np.array([0.0, 3.0, np.inf, 2.0, 3.0, -1.0, 3.0, 0.0, np.nan, 1.0 / 3.0, -1.0, 0.0], np.float32)
    A, X = buf(a, F32), buf(x, F32)
    cuda_service.math().pow(False, 1, a.size, A, 0, a.size, X, 0, 1)
    ref = oracle.as_buffer(a)
    oracle.broadcast(oracle.BC_POW, False, 1, a.size, ref, 0, a.size, oracle.as_buffer(x), 0, 1)
    assert np.array_equal(A.to_numpy(), ref.astype(np.float32), equal_nan=True)


UN_SHAPES = [(1, 1, 0), (5, 1, 0), (31, 1, 1), (1000, 1, 3), (4099, 1, 0), (333, 3, 2), (1 << 20, 1, 0), (100003, 1, 5),
             (65537, 2, 1)]
# op -> (oracle code, input range, relative tolerance f32, (alpha, beta) or None)
UN_OPS = {
    "increment": (oracle.UN_INCREMENT, (-4, 4), 0.0, (-1.7, 0.3)), "reciprocal": (oracle.UN_RECIPROCAL, (0.1, 4), 4e-7, (1.7, 0.3)),
    "square": (oracle.UN_SQUARE, (-4, 4), 6e-8, None), "sqrt": (oracle.UN_SQRT, (0, 16), 6e-8, None),
    "rsqrt": (oracle.UN_RSQRT, (0.01, 16), 4e-7, (1.3, 0.25)), "exp": (oracle.UN_EXP, (-10, 10), 4e-6, None),
    "log": (oracle.UN_LOG, (1e-3, 100), 4e-6, None), "tanh": (oracle.UN_TANH, (-5, 5), 4e-6, None),
    "sin": (oracle.UN_SIN, (-10, 10), 4e-6, None), "cos": (oracle.UN_COS, (-10, 10), 4e-6, None),
    "tan": (oracle.UN_TAN, (-1.4, 1.4), 4e-6, None),
}


@pytest.mark.parametrize("n,incX,offX", UN_SHAPES, ids=[str(v) for v in UN_SHAPES])
@pytest.mark.parametrize("op", list(UN_OPS))
@pytest.mark.parametrize("dt", [F32, F64])
def test_unary_family_vs_oracle(n, incX, offX, op, dt, cuda_service):
    code, (lo, hi), rtol32, ab = UN_OPS[op]
    rng = np.random.default_rng(n + incX * 13)
    npdt = NP[dt]
    eps = float(np.finfo(npdt).eps)
    x = rng.uniform(lo, hi, offX + (n - 1) * incX + 1 + 2).astype(npdt)
    alpha, beta = (float(npdt(ab[0])), float(npdt(ab[1]))) if ab else (1.0, 0.0)
    X = buf(x, dt)
    math = cuda_service.math()
    if ab:
        getattr(math, op)(n, alpha, X, offX, incX, beta)
    else:
        getattr(math, op)(n, X, offX, incX)
    ref = oracle.as_buffer(x)
    oracle.unary(code, n, ref, offX, incX, alpha, beta)
    got = X.to_numpy().astype(np.float64)
    if op == "increment":
        # two roundings (product, sum) against the double result: eps * (|alpha*x| + |result|)
        bound = eps * (np.abs(alpha * x.astype(np.float64)) + np.abs(ref)) + 1e-300
    else:
        rtol = rtol32 if dt == F32 else 40 * eps
        atol = (1e-7 if dt == F32 else 1e-16) if op in ("sin", "cos", "tanh", "log") else 0.0
        bound = rtol * np.abs(ref) + atol
    sel = np.zeros(x.size, bool)
    sel[offX:offX + (n - 1) * incX + 1:incX] = True
    assert np.all(np.abs(got - ref)[sel] <= bound[sel]), np.max(np.abs(got - ref) / (np.abs(ref) + 1e-30))
    # elements between the strides and outside [offX, ...) are untouched
    assert np.array_equal(X.to_numpy()[~sel], x[~sel])


@pytest.mark.parametrize("dt", [F32, F64])
def test_unary_special_values(dt, cuda_service):
    """INF / NaN / zero tables of LinearAlgebraTest.php:3315-3324, 3891-3900, 3927-3936, 4019-4028 and the
    not / isnan / nan2num selection ops on a long vector: bit-exact against the oracle."""
    npdt = NP[dt]
    rng = np.random.default_rng(5)
    x = special(rng, 100003, npdt, frac=0.05)
    math = cuda_service.math()
    table = [("reciprocal", oracle.UN_RECIPROCAL, True), ("sqrt", oracle.UN_SQRT, False), ("rsqrt", oracle.UN_RSQRT, True),
             ("log", oracle.UN_LOG, False), ("not", oracle.UN_NOT, False), ("isnan", oracle.UN_ISNAN, False),
             ("square", oracle.UN_SQUARE, False), ("exp", oracle.UN_EXP, False), ("tanh", oracle.UN_TANH, False)]
    for name, code, has_ab in table:
        X = buf(x, dt)
        if has_ab:
            getattr(math, name)(x.size, 1.0, X, 0, 1, 0.0)
        else:
            getattr(math, name)(x.size, X, 0, 1)
        ref = oracle.as_buffer(x)
        oracle.unary(code, x.size, ref, 0, 1, 1.0, 0.0)
        got = X.to_numpy()
        want = ref.astype(npdt)
        special_mask = ~np.isfinite(want) | (want == 0) | ~np.isfinite(x) | (x == 0)
        assert np.array_equal(got[special_mask], want[special_mask], equal_nan=True), name
        if name in ("not", "isnan", "square", "sqrt", "reciprocal"):
            assert np.array_equal(got, want, equal_nan=True), name
    X = buf(x, dt)
    math.nan2num(x.size, X, 0, 1, 2.5)
    assert np.array_equal(X.to_numpy(), np.where(np.isnan(x), npdt(2.5), x))
    assert np.array_equal(np.signbit(X.to_numpy()), np.signbit(np.where(np.isnan(x), npdt(2.5), x)))


@pytest.mark.parametrize("n,incX,offX", UN_SHAPES, ids=[str(v) for v in UN_SHAPES])
@pytest.mark.parametrize("dt", sorted(NP))
def test_equal_notequal_not_all_dtypes(n, incX, offX, dt, cuda_service):
    rng = np.random.default_rng(n * 3 + dt)
    npdt = NP[dt]
    size = offX + (n - 1) * incX + 1 + 2
    if dt in (F32, F64):
        x, y = special(rng, size, npdt, 0.05), special(rng, size, npdt, 0.05)
    elif dt == dtypes.bool_:
        x, y = rng.integers(0, 2, size).astype(npdt), rng.integers(0, 2, size).astype(npdt)
    else:
        x, y = rng.integers(0, 3, size).astype(npdt), rng.integers(0, 3, size).astype(npdt)
    math = cuda_service.math()
    for name, ne in (("equal", False), ("notEqual", True)):
        X, Y = buf(x, dt), buf(y, dt)
        getattr(math, name)(n, X, offX, incX, Y, offX, incX)
        ref = oracle.as_buffer(y)
        oracle.compare(ne, n, oracle.as_buffer(x), offX, incX, ref, offX, incX)
        assert np.array_equal(Y.to_numpy(), ref.astype(npdt), equal_nan=True), name
        assert np.array_equal(X.to_numpy(), x, equal_nan=True)
    X = buf(x, dt)
    getattr(math, "not")(n, X, offX, incX)
    ref = oracle.as_buffer(x)
    oracle.unary(oracle.UN_NOT, n, ref, offX, incX)
    assert np.array_equal(X.to_numpy(), ref.astype(npdt), equal_nan=True)


@pytest.mark.parametrize("src", sorted(NP))
@pytest.mark.parametrize("dst", sorted(NP))
def test_astype_all_pairs(src, dst, cuda_service):
    rng = np.random.default_rng(src * 17 + dst)
    n, offX, offY, incX, incY = 4099, 3, 1, 2, 1
    sdt, ddt = NP[src], NP[dst]
    size = offX + (n - 1) * incX + 2
    if src in (F32, F64):
        x = (rng.uniform(-100, 100, size)).astype(sdt)
        x[rng.integers(0, size, 50)] = 0.0
        if dst in (dtypes.uint64,):
            x = np.abs(x)
    elif src == dtypes.bool_:
        x = rng.integers(0, 2, size).astype(sdt)
    else:
        info = np.iinfo(sdt)
        x = rng.integers(max(info.min, -100), min(info.max, 100) + 1, size).astype(sdt)
    X, Y = buf(x, src), CudaBuffer(offY + n * incY + 2, dst)
    cuda_service.math().astype(n, dst, X, offX, incX, Y, offY, incY)
    xo = x.astype(np.float64) if src in (F32, F64) else x.astype(np.int64)
    if dst == dtypes.bool_:
        kind, mask = "bool", 0
    elif dst in (F32, F64):
        kind, mask = "float", 0
    else:
        kind, mask = "int", {dtypes.uint8: 0xff, dtypes.uint16: 0xffff, dtypes.uint32: 0xffffffff}.get(dst, 0)
    want = oracle.astype(n, kind, mask, xo, offX, incX)
    got = Y.to_numpy()[offY:offY + n * incY:incY]
    if dst != dtypes.bool_:
        want = want.astype(ddt)          # the C buffer rounds (float)x / wraps a negative int to the target width
    if dst in (dtypes.int8, dtypes.int16, dtypes.int32):
        # PHP ints do not wrap; parity holds while the value fits (here: |x| <= 100 fits int8 .. int64)
        assert np.all(np.abs(want) <= 127)
    assert np.array_equal(got.astype(np.float64), want.astype(np.float64))
    assert Y.to_numpy()[0] == 0 and Y.to_numpy()[-1] == 0          # outside the range: untouched


SR_SHAPES = [(1, 1, 0), (2, 1, 0), (31, 1, 1), (1000, 1, 3), (4099, 1, 0), (333, 3, 2), (1 << 20, 1, 0), (2000003, 1, 1),
             (65537, 2, 1), (1 << 24, 1, 0)]


@pytest.mark.parametrize("n,incX,offX", SR_SHAPES, ids=[str(v) for v in SR_SHAPES])
@pytest.mark.parametrize("dt", [F32, F64, dtypes.int32, dtypes.int64, dtypes.uint8, dtypes.bool_])
def test_sum_imax_imin_vs_oracle(n, incX, offX, dt, cuda_service):
    rng = np.random.default_rng(n + dt)
    npdt = NP[dt]
    size = offX + (n - 1) * incX + 1 + 2
    if dt in (F32, F64):
        x = (np.round(rng.uniform(-4, 4, size) * 64) / 64).astype(npdt)       # ties
    elif dt == dtypes.bool_:
        x = rng.integers(0, 2, size).astype(npdt)
    elif dt == dtypes.uint8:
        x = rng.integers(0, 256, size).astype(npdt)
    else:
        x = rng.integers(-1000, 1000, size).astype(npdt)
    X = buf(x, dt)
    math = cuda_service.math()
    xo = oracle.as_buffer(x)
    s = math.sum(n, X, offX, incX)
    ref = oracle.sum(n, xo, offX, incX)
    assert abs(s - ref) <= 1e-12 * float(np.sum(np.abs(xo[offX:offX + (n - 1) * incX + 1:incX]))) + 1e-300
    if dt not in (F32, F64):
        assert s == ref                                   # integer sums are exact in double at these sizes
    assert math.imax(n, X, offX, incX) == oracle.imax(n, xo, offX, incX)
    assert math.imin(n, X, offX, incX) == oracle.imin(n, xo, offX, incX)


@pytest.mark.parametrize("dt", [F32, F64])
def test_imax_imin_nan_rules(dt, cuda_service):
    """PhpMath::imax (:185) replaces a NaN running value by the next element; imin (:210) keeps a NaN at 0."""
    npdt = NP[dt]
    math = cuda_service.math()
    rows = [[np.nan, 1.0, 3.0, 3.0, np.nan], [np.nan, np.nan, np.nan], [np.nan], [1.0, np.nan, 0.5, np.nan],
            [np.nan, np.nan, -np.inf, -np.inf], [np.inf, np.nan, np.inf], [-np.inf, -np.inf], [2.0, 2.0, 1.0, 1.0]]
    rng = np.random.default_rng(9)
    big = rng.standard_normal(300007)
    big[:5] = np.nan
    big[rng.integers(0, big.size, 1000)] = np.nan
    rows.append(big.tolist())
    allnan = np.full(70001, np.nan)
    rows.append(allnan.tolist())
    for row in rows:
        x = np.asarray(row, npdt)
        X = buf(x, dt)
        xo = oracle.as_buffer(x)
        assert math.imax(x.size, X, 0, 1) == oracle.imax(x.size, xo, 0, 1), row[:6]
        assert math.imin(x.size, X, 0, 1) == oracle.imin(x.size, xo, 0, 1), row[:6]


def test_linear_algebra_max_min_sum(cuda_service, oracle_service):
    la, lo = LinearAlgebra(cuda_service), LinearAlgebra(oracle_service)
    rng = np.random.default_rng(21)
    x = rng.standard_normal((300, 1001)).astype(np.float32)
    a, b = la.array(x), lo.array(x)
    assert la.max(a) == lo.max(b) and la.min(a) == lo.min(b)
    assert la.imax(a) == lo.imax(b) and la.imin(a) == lo.imin(b)
    assert abs(la.sum(a) - lo.sum(b)) <= 1e-12 * np.abs(x).sum()
    # views (offset != 0)
    assert la.max(a[7]) == lo.max(b[7]) and la.imin(a[299]) == lo.imin(b[299])


@pytest.mark.parametrize("xdt", [dtypes.int32, dtypes.int64, dtypes.uint8, F32])
@pytest.mark.parametrize("ydt", [F32, F64])
def test_onehot_vs_oracle(xdt, ydt, cuda_service):
    rng = np.random.default_rng(xdt + ydt)
    m, n, incX, offX, offY, ldY = 5003, 37, 2, 1, 3, 40
    labels = rng.integers(0, n, offX + (m - 1) * incX + 2).astype(NP[xdt])
    y = rng.standard_normal(offY + (m - 1) * ldY + n + 1).astype(NP[ydt])
    X, Y = buf(labels, xdt), buf(y, ydt)
    cuda_service.math().updateAddOnehot(m, n, -2.0, X, offX, incX, Y, offY, ldY)
    ref = oracle.as_buffer(y)
    oracle.onehot(m, n, -2.0, oracle.as_buffer(labels), offX, incX, ref, offY, ldY)
    assert np.array_equal(Y.to_numpy(), ref.astype(NP[ydt]))
    # a label outside [0, n): RuntimeException('Label number is out of bounds.') (PhpMath.php:1460-1461)
    labels[offX + 10 * incX] = n
    X = buf(labels, xdt)
    with pytest.raises(RuntimeException, match="Label number is out of bounds."):
        cuda_service.math().updateAddOnehot(m, n, 1.0, X, offX, incX, Y, offY, ldY)
    # the library keeps working after the error
    la = LinearAlgebra(cuda_service)
    assert la.onehot(la.array([0, 1, 2, 0], dtype=dtypes.int32), 3).toArray() == [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]]


def test_host_checks_match_reference_messages(cuda_service):
    math = cuda_service.math()
    A, X = CudaBuffer(6, F32), CudaBuffer(2, F32)
    with pytest.raises(InvalidArgumentException, match="m and n must be greater than 0"):
        math.maximum(0, 3, A, 0, 3, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Matrix A is too small"):
        math.greater(3, 3, A, 0, 3, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Buffer X is too small"):
        math.less(2, 3, A, 0, 3, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for buffer."):
        math.exp(7, A, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for buffer."):
        math.pow(False, 2, 3, A, 0, 3, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector X specification too large for buffer."):
        math.sum(7, A, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for bufferX."):
        math.updateAddOnehot(3, 2, 1.0, X, 0, 1, A, 0, 2)


def test_baseline_size_properties(cuda_service):
    """At BASELINE config C2 size ([8192,4096] against [4096]) the oracle is too slow for every op; size-independent
    properties instead: maximum(A,X) >= both, minimum <= both, max+min == A+X elementwise (exact: one of each),
    greater + lessEqual == 1, duplicate then equal == all ones, sqrt(square(x)) == |x| within 1 ulp, sum of ones."""
    la = LinearAlgebra(cuda_service)
    rng = np.random.default_rng(8)
    m, n = 8192, 4096
    a = rng.uniform(-1, 1, (m, n)).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    X = la.array(x)
    mx = la.maximum(la.array(a), X).to_numpy()
    mn = la.minimum(la.array(a), X).to_numpy()
    assert np.array_equal(mx, np.maximum(a, x)) and np.array_equal(mn, np.minimum(a, x))
    g = la.greater(la.array(a), X).to_numpy()
    le = la.lessEqual(la.array(a), X).to_numpy()
    assert np.array_equal(g + le, np.ones_like(a)) and np.array_equal(g, (a > x).astype(np.float32))
    D = la.duplicate(X, m)
    assert D.shape() == [m, n]
    eq = la.equal(la.duplicate(X, m), D)
    assert la.sum(eq) == float(m * n)
    r = la.sqrt(la.square(la.array(a))).to_numpy()
    assert np.all(np.abs(r - np.abs(a)) <= np.spacing(np.abs(a)))
    mean = la.reduceMean(la.array(a), axis=0).to_numpy()
    assert np.allclose(mean, a.astype(np.float64).mean(axis=0), rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# masking (PhpMath.php:3229-3270)
# ---------------------------------------------------------------------------------------------------
MASK_SHAPES = [(1, 1, 6, 1), (2, 4, 3, 1), (1, 1, 6, 4), (3, 5, 7, 2), (1, 8, 1024, 1), (16, 12, 128, 128), (64, 1, 512, 64)]


@pytest.mark.parametrize("m,n,k,ln", MASK_SHAPES, ids=[str(s) for s in MASK_SHAPES])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32, dtypes.uint8, dtypes.bool_])
@pytest.mark.parametrize("mode,fill", [(0, 0.0), (0, -1e9), (1, -1000.0), (1, 2.75)])
def test_masking_vs_oracle(m, n, k, ln, dt, mode, fill, cuda_service):
    """Bit-exact: a store of the converted fill or one rounded add; untouched elements stay."""
    if dt in (dtypes.uint8, dtypes.bool_) and fill < 0:
        pytest.skip("negative fill into an unsigned buffer")
    rng = np.random.default_rng(m * 100 + k)
    npdt = dtypes.NUMPY[dt]
    offX, offA = 2, 3
    x = (rng.uniform(0, 1, offX + m * k) > 0.4).astype(np.uint8)
    if dt == dtypes.bool_:
        a = (rng.uniform(0, 1, offA + m * n * k * ln) > 0.5).astype(np.uint8)
    else:
        a = np.floor(rng.uniform(-50, 50, offA + m * n * k * ln)).astype(npdt)
    X = CudaBuffer(x.size, dtypes.bool_).from_numpy(x)
    A = CudaBuffer(a.size, dt).from_numpy(a)
    cuda_service.math().masking(m, n, k, ln, fill, mode, X, offX, A, offA)
    ao = a.astype(np.float64)
    kind = 2 if dt == dtypes.bool_ else (0 if dt in (dtypes.float32, dtypes.float64) else 1)
    oracle.masking(m, n, k, ln, fill, mode, kind, x.astype(np.float64), offX, ao, offA)
    if dt == dtypes.float32:
        ao = ao.astype(np.float32).astype(np.float64)
    if dt == dtypes.uint8:
        ao = np.mod(ao, 256)
    got = A.to_numpy().astype(np.float64)
    if dt == dtypes.float32 and mode == 1:
        # one float32 add on the device, one double add in the oracle: equal after rounding the operands' sum once
        ref32 = a.astype(np.float32)
        sel = np.zeros(ref32.size, bool)
        msk = np.broadcast_to((x[offX:offX + m * k].reshape(m, 1, k, 1) == 0), (m, n, k, ln)).reshape(-1)
        sel[offA:offA + msk.size] = msk
        ref32[sel] = ref32[sel] + np.float32(fill)
        assert np.array_equal(A.to_numpy(), ref32)
    else:
        assert np.array_equal(got, ao)
    assert np.array_equal(X.to_numpy(), x)


def test_masking_errors(cuda_service):
    math = cuda_service.math()
    X = CudaBuffer(6, dtypes.bool_)
    A = CudaBuffer(24, dtypes.float32)
    with pytest.raises(rb.InvalidArgumentException, match="dtype of X must be bool."):
        math.masking(1, 1, 6, 1, 0.0, 0, A, 0, A, 0)
    with pytest.raises(rb.InvalidArgumentException, match="Matrix specification too large for buffer X."):
        math.masking(1, 1, 7, 1, 0.0, 0, X, 0, A, 0)
    with pytest.raises(rb.InvalidArgumentException, match="Matrix specification too large for buffer A."):
        math.masking(1, 5, 6, 1, 0.0, 0, X, 0, A, 0)


@pytest.mark.parametrize("m,n,k", [(1, 5, 5), (2, 5, 5), (3, 7, 4), (4, 4, 9), (16, 512, 512), (1, 1, 1)])
@pytest.mark.parametrize("lower,upper", [(0, -1), (-1, 0), (0, 0), (0, 1), (1, 0), (2, 3), (-1, -1), (100, 100)])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32, dtypes.bool_])
def test_bandpart_vs_oracle(m, n, k, lower, upper, dt, cuda_service):
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(n * 13 + k)
    off = 3
    a = (rng.integers(1, 3, off + m * n * k)).astype(npdt) if dt != dtypes.bool_ else np.ones(off + m * n * k, np.uint8)
    A = CudaBuffer(a.size, dt).from_numpy(a)
    cuda_service.math().bandpart(m, n, k, A, off, lower, upper)
    ao = a.astype(np.float64)
    oracle.bandpart(m, n, k, ao, off, lower, upper)
    assert np.array_equal(A.to_numpy().astype(np.float64), ao)
