"""Parity of the section 8(f) widening -- scal / axpy / copy / gemv / omatcopy / matmul / vectorTransform --
through the C ABI against the CPU oracle.  pytest -m gpu

Tolerances: copy and omatcopy with alpha == 1 are bit-exact (pure data movement); scal, omatcopy with
alpha != 1 and axpy with |alpha| == 1 are one correctly-rounded float op, bit-exact against the oracle's
double result rounded once; axpy with another alpha: 2*eps*(|alpha*x| + |y|) (the device fuses mul+add);
gemv: the reference isclose bound 1e-7 + 1e-4*max|ref| with the measured-tight bound asserted next to it.
"""
import numpy as np
import pytest

import oracle
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, InvalidArgumentException, LinearAlgebra, MatrixOperator, dtypes

pytestmark = pytest.mark.gpu

NP = {dtypes.float32: np.float32, dtypes.float64: np.float64}
VEC = [(1, 1, 1, 0, 0), (7, 1, 1, 0, 0), (1000, 1, 1, 0, 0), (1000, 1, 1, 3, 5), (4096, 1, 1, 4, 8),
       (333, 2, 3, 1, 2), (100003, 1, 1, 0, 0), (1 << 20, 1, 1, 0, 0), (65537, 3, 1, 2, 0)]


def buf(a, dt):
    a = np.asarray(a)
    return CudaBuffer(a.size, dt).from_numpy(a.reshape(-1))


def _vec(rng, n, inc, off, npdt):
    return rng.standard_normal(off + (n - 1) * inc + 1 + 3).astype(npdt)


@pytest.mark.parametrize("n,incX,incY,offX,offY", VEC, ids=[str(v) for v in VEC])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_copy_scal_axpy_vs_oracle(n, incX, incY, offX, offY, dt, cuda_service):
    rng = np.random.default_rng(n + incX)
    npdt = NP[dt]
    blas = cuda_service.blas()
    eps = np.finfo(npdt).eps
    x, y = _vec(rng, n, incX, offX, npdt), _vec(rng, n, incY, offY, npdt)
    # copy: bit-exact, untouched elements stay
    X, Y = buf(x, dt), buf(y, dt)
    blas.copy(n, X, offX, incX, Y, offY, incY)
    ref = oracle.as_buffer(y)
    oracle.copy(n, oracle.as_buffer(x), offX, incX, ref, offY, incY)
    assert np.array_equal(Y.to_numpy(), ref.astype(npdt))
    assert np.array_equal(X.to_numpy(), x)
    # scal: one rounded multiply
    X = buf(x, dt)
    blas.scal(n, 0.37, X, offX, incX)
    ref = oracle.as_buffer(x)
    oracle.scal(n, float(npdt(0.37)), ref, offX, incX)
    assert np.array_equal(X.to_numpy(), ref.astype(npdt))
    # axpy
    for alpha in (1.0, -1.0, 0.3):
        X, Y = buf(x, dt), buf(y, dt)
        blas.axpy(n, alpha, X, offX, incX, Y, offY, incY)
        ref = oracle.as_buffer(y)
        oracle.axpy(n, float(npdt(alpha)), oracle.as_buffer(x), offX, incX, ref, offY, incY)
        got = Y.to_numpy().astype(np.float64)
        if abs(alpha) == 1.0 and dt == dtypes.float32:
            assert np.array_equal(got, ref.astype(npdt).astype(np.float64))
        else:
            xs = np.zeros_like(ref)
            xs[offY:offY + (n - 1) * incY + 1:incY] = np.abs(alpha * x[offX:offX + (n - 1) * incX + 1:incX])
            assert np.all(np.abs(got - ref) <= 2 * eps * (xs + np.abs(y.astype(np.float64))) + 1e-300)


GEMV = [(1, 1, 0, 0), (2, 3, 0, 0), (64, 64, 0, 0), (257, 1023, 5, 3), (1000, 33, 0, 2), (33, 5000, 7, 0),
        (4096, 4096, 0, 0), (3, 100000, 0, 0), (100000, 3, 1, 0), (1000, 1028, 4, 4), (517, 2052, 0, 8),
        (2048, 136, 0, 0)]


@pytest.mark.parametrize("m,n,ld_extra,off", GEMV, ids=[str(v) for v in GEMV])
@pytest.mark.parametrize("trans", [BLAS.NoTrans, BLAS.Trans])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_gemv_vs_oracle(m, n, ld_extra, off, trans, dt, cuda_service):
    rng = np.random.default_rng(m * 31 + n)
    npdt = NP[dt]
    blas = cuda_service.blas()
    ld = n + ld_extra
    a = rng.standard_normal(off + m * ld).astype(npdt)
    rows, cols = (m, n) if trans == BLAS.NoTrans else (n, m)
    incX, incY = (1, 1) if m * n > 10000 else (2, 3)
    offX = 4 if m * n > 10000 else 1            # 16-byte aligned X takes the 128-bit kernels
    x = rng.standard_normal(offX + (cols - 1) * incX + 1).astype(npdt)
    y = rng.standard_normal(2 + (rows - 1) * incY + 1).astype(npdt)
    for alpha, beta in ((1.0, 0.0), (0.5, 1.0), (-1.25, 0.75)):
        A, X, Y = buf(a, dt), buf(x, dt), buf(y, dt)
        blas.gemv(BLAS.RowMajor, trans, m, n, alpha, A, off, ld, X, offX, incX, beta, Y, 2, incY)
        ref = oracle.as_buffer(y)
        oracle.gemv(BLAS.RowMajor, trans, m, n, alpha, oracle.as_buffer(a), off, ld, oracle.as_buffer(x), offX, incX,
                    beta, ref, 2, incY)
        got = Y.to_numpy().astype(np.float64)
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(got - ref)) < 1e-7 + 1e-4 * scale                  # reference isclose
        tight = (4e-6 if dt == dtypes.float32 else 1e-13) * max(scale, 1.0)
        assert np.max(np.abs(got - ref)) < tight
        # elements between the strides are untouched
        mask = np.ones(y.size, bool)
        mask[2:2 + (rows - 1) * incY + 1:incY] = False
        assert np.array_equal(Y.to_numpy()[mask], y[mask])


def test_gemv_beta_zero_ignores_nan_and_colmajor(cuda_service):
    blas = cuda_service.blas()
    a = np.arange(6, dtype=np.float32) + 1                                        # [[1,2,3],[4,5,6]]
    A, X = buf(a, dtypes.float32), buf(np.array([100, 10, 1], np.float32), dtypes.float32)
    Y = buf(np.full(2, np.nan, np.float32), dtypes.float32)
    blas.gemv(BLAS.RowMajor, BLAS.NoTrans, 2, 3, 1.0, A, 0, 3, X, 0, 1, 0.0, Y, 0, 1)
    assert Y.to_numpy().tolist() == [123.0, 456.0]
    # ColMajor swaps m,n before anything else (PhpBlas.php:773-775): the 2x3 row-major A is a 3x2 col-major one
    Y = buf(np.zeros(2, np.float32), dtypes.float32)
    blas.gemv(BLAS.ColMajor, BLAS.NoTrans, 3, 2, 1.0, A, 0, 3, X, 0, 1, 0.0, Y, 0, 1)
    ref = np.zeros(2)
    oracle.gemv(BLAS.ColMajor, BLAS.NoTrans, 3, 2, 1.0, oracle.as_buffer(a), 0, 3,
                oracle.as_buffer([100, 10, 1]), 0, 1, 0.0, ref, 0, 1)
    assert Y.to_numpy().tolist() == ref.tolist() == [123.0, 456.0]


OMAT = [(1, 1, 0, 0), (2, 3, 0, 0), (32, 32, 0, 0), (33, 65, 3, 2), (1000, 7, 0, 0), (7, 1000, 1, 5),
        (2048, 2048, 0, 0), (4097, 1025, 0, 0), (100, 260, 4, 8), (4096, 1028, 0, 0), (64, 64, 0, 0), (68, 4, 0, 4)]


@pytest.mark.parametrize("m,n,ld_extra,off", OMAT, ids=[str(v) for v in OMAT])
@pytest.mark.parametrize("trans", [BLAS.NoTrans, BLAS.Trans])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_omatcopy_vs_oracle_bit_exact(m, n, ld_extra, off, trans, dt, cuda_service):
    rng = np.random.default_rng(m * 7 + n)
    npdt = NP[dt]
    blas = cuda_service.blas()
    ldA = n + ld_extra
    rows, cols = (m, n) if trans == BLAS.NoTrans else (n, m)
    ldB = cols + ld_extra
    a = rng.standard_normal(off + m * ldA).astype(npdt)
    b = rng.standard_normal(off + rows * ldB).astype(npdt)
    for alpha in (1.0, -2.5):
        A, B = buf(a, dt), buf(b, dt)
        blas.omatcopy(BLAS.RowMajor, trans, m, n, alpha, A, off, ldA, B, off, ldB)
        ref = oracle.as_buffer(b)
        oracle.omatcopy(BLAS.RowMajor, trans, m, n, alpha, oracle.as_buffer(a), off, ldA, ref, off, ldB)
        assert np.array_equal(B.to_numpy(), ref.astype(npdt))
        assert np.array_equal(A.to_numpy(), a)


def test_transpose_and_matmul_random(cuda_service, oracle_service):
    rng = np.random.default_rng(5)
    la, lo = LinearAlgebra(cuda_service), LinearAlgebra(oracle_service)
    a = rng.standard_normal((300, 517)).astype(np.float32)
    assert np.array_equal(la.transpose(la.array(a)).to_numpy(), a.T)
    A = rng.standard_normal((3, 2, 40, 24)).astype(np.float32)
    B = rng.standard_normal((2, 24, 56)).astype(np.float32)
    got = la.matmul(la.array(A), la.array(B))
    ref = lo.matmul(lo.array(A), lo.array(B))
    assert got.shape() == ref.shape() == [3, 2, 40, 56]
    r = np.asarray(ref.toArray(), np.float64)
    assert np.max(np.abs(got.to_numpy() - r)) < 5e-5 * np.max(np.abs(r))


def test_vector_transform_random(cuda_service, oracle_service):
    rng = np.random.default_rng(6)
    mo, mr = MatrixOperator(cuda_service), MatrixOperator(oracle_service)
    A = rng.standard_normal((4, 5, 300, 129)).astype(np.float32)
    B = rng.standard_normal(129).astype(np.float32)
    got = mo.cross(mo.array(A), mo.array(B))
    ref = mr.cross(mr.array(A), mr.array(B))
    assert got.shape() == [4, 5, 300]
    r = np.asarray(ref.toArray(), np.float64)
    assert np.max(np.abs(got.to_numpy() - r)) < 4e-6 * np.max(np.abs(r))
    with pytest.raises(InvalidArgumentException, match="unmatch shape of dimension"):
        mo.cross(mo.array(A), mo.array(np.ones(128, np.float32)))


def test_blas_widen_errors(cuda_service):
    blas = cuda_service.blas()
    X, Y = buf(np.zeros(4, np.float32), dtypes.float32), buf(np.zeros(4, np.float32), dtypes.float32)
    with pytest.raises(InvalidArgumentException, match="Argument n must be greater than 0."):
        blas.scal(0, 1.0, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for bufferX."):
        blas.axpy(5, 1.0, X, 0, 1, Y, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for bufferY."):
        blas.copy(3, X, 0, 1, Y, 0, 2)
    with pytest.raises(InvalidArgumentException, match="Matrix specification too large for bufferA."):
        blas.gemv(BLAS.RowMajor, BLAS.NoTrans, 2, 3, 1.0, X, 0, 3, Y, 0, 1, 0.0, Y, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Invalid Order type"):
        blas.gemv(0, BLAS.NoTrans, 2, 2, 1.0, X, 0, 2, Y, 0, 1, 0.0, Y, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Unknown Tranpose Code"):
        blas.omatcopy(BLAS.RowMajor, 0, 2, 2, 1.0, X, 0, 2, Y, 0, 2)
    with pytest.raises(InvalidArgumentException, match="Matrix specification too large for bufferB."):
        blas.omatcopy(BLAS.RowMajor, BLAS.Trans, 1, 4, 1.0, X, 0, 4, Y, 1, 1)


BATCHED = [
    # batch, m, n, k, transA, transB, share (which operand is shared: 0 none, 1 A, 2 B), dtype, mode
    (7, 40, 56, 24, False, False, 0, dtypes.float32, rb.GEMM_AUTO),
    (5, 33, 17, 65, True, False, 0, dtypes.float32, rb.GEMM_FP32_SIMT),
    (4, 64, 64, 64, False, True, 2, dtypes.float32, rb.GEMM_TF32X3),
    (3, 256, 128, 96, False, False, 0, dtypes.float32, rb.GEMM_TF32X3),      # one tensor-core launch per matrix
    (3, 128, 384, 64, True, True, 1, dtypes.float32, rb.GEMM_TF32),
    (6, 20, 30, 10, False, False, 0, dtypes.float64, rb.GEMM_AUTO),
    (70000, 2, 3, 4, False, False, 0, dtypes.float32, rb.GEMM_AUTO),          # more matrices than gridDim.z allows
    # tensor-core batched path (one persistent launch; batch = third tensor-map dimension)
    (5, 128, 128, 64, False, False, 0, dtypes.float32, rb.GEMM_TF32X3),
    (16, 64, 64, 32, False, True, 0, dtypes.float32, rb.GEMM_TF32X3),
    (7, 192, 320, 100, True, False, 0, dtypes.float32, rb.GEMM_TF32X3),       # M / N / K tails
    (9, 200, 136, 72, True, True, 2, dtypes.float32, rb.GEMM_TF32X3),         # shared B
    (6, 256, 512, 64, False, False, 1, dtypes.float32, rb.GEMM_TF32),         # shared A, wide tiles
    (300, 128, 64, 128, False, False, 0, dtypes.float32, rb.GEMM_TF32),
    (2, 96, 80, 4200, False, True, 0, dtypes.float32, rb.GEMM_TF32X3),        # K chunks of 2048
]


@pytest.mark.parametrize("batch,m,n,k,tA,tB,share,dt,mode", BATCHED, ids=[str(c[:7]) for c in BATCHED])
def test_gemm_strided_batched_vs_oracle(batch, m, n, k, tA, tB, share, dt, mode, cuda_service):
    """rb200_gemm_strided_batched == the loop of PhpBlas::gemm calls LinearAlgebra::matmul issues
    (LinearAlgebra.php:1091-1106), operand strides as there (M*K, N*K, M*N) or 0 for a shared operand."""
    rng = np.random.default_rng(batch * 7 + m)
    npdt = NP[dt]
    sA = 0 if share == 1 else m * k
    sB = 0 if share == 2 else n * k
    tensor = mode in (rb.GEMM_TF32, rb.GEMM_TF32X3) and m >= 64 and n >= 64 and k >= 32 and dt == dtypes.float32
    oA, oB, oC = (4, 4, 4) if tensor else (5, 3, 2)                     # the tensor-core path needs 16-byte aligned operands
    a = rng.uniform(-1, 1, (1 if share == 1 else batch) * m * k + oA).astype(npdt)
    b = rng.uniform(-1, 1, (1 if share == 2 else batch) * n * k + oB).astype(npdt)
    c0 = rng.uniform(-1, 1, batch * m * n + oC).astype(npdt)
    ldA, ldB = (m if tA else k), (k if tB else n)
    A, B, C = buf(a, dt), buf(b, dt), buf(c0, dt)
    math = cuda_service.math()
    math.gemmStridedBatched(BLAS.RowMajor, BLAS.Trans if tA else BLAS.NoTrans, BLAS.Trans if tB else BLAS.NoTrans, m, n, k,
                            0.5, A, oA, ldA, sA, B, oB, ldB, sB, 2.0, C, oC, n, m * n, batch, gemm_mode=mode)
    if tensor:
        assert _capi_path() == ("tcgen05_tf32_batched" if mode == rb.GEMM_TF32 else "tcgen05_tf32x3_batched"), _capi_path()
    got = C.to_numpy().astype(np.float64)
    ref = oracle.as_buffer(c0)
    ao, bo = oracle.as_buffer(a), oracle.as_buffer(b)
    check = range(batch) if batch <= 64 else list(range(0, batch, 997)) + [batch - 1]
    for i in check:
        oracle.gemm(oracle.RowMajor, oracle.Trans if tA else oracle.NoTrans, oracle.Trans if tB else oracle.NoTrans, m, n, k,
                    0.5, ao, oA + i * sA, ldA, bo, oB + i * sB, ldB, 2.0, ref, oC + i * m * n, n)
        sl = slice(oC + i * m * n, oC + (i + 1) * m * n)
        tol = 2e-3 if (mode == rb.GEMM_TF32 and tensor) else (1e-12 if dt == dtypes.float64 else 5e-5)
        assert np.max(np.abs(got[sl] - ref[sl])) < tol * max(1.0, np.max(np.abs(ref[sl]))), (i, _capi_path())
    assert got[0] == c0[0] and got[1] == c0[1]                       # before offsetC: untouched


def _capi_path():
    from rindow_math_matrix_b200 import _capi
    return _capi.last_gemm_path()


def test_matmul_uses_one_batched_launch(cuda_service, oracle_service):
    from rindow_math_matrix_b200 import _capi
    rng = np.random.default_rng(11)
    la, lo = LinearAlgebra(cuda_service), LinearAlgebra(oracle_service)
    A = rng.standard_normal((6, 8, 48, 32)).astype(np.float32)
    B = rng.standard_normal((6, 8, 32, 40)).astype(np.float32)
    a, b = la.array(A), la.array(B)
    _capi.sync()
    n0 = _capi.launch_count()
    got = la.matmul(a, b)
    launches = _capi.launch_count() - n0
    assert launches <= 3, launches                                    # zero-fills of C (alloc, zeros) + ONE batched kernel, not 48
    assert _capi.last_gemm_path() == "simt_f32_batched"
    ref = np.asarray(lo.matmul(lo.array(A), lo.array(B)).toArray(), np.float64)
    assert np.max(np.abs(got.to_numpy() - ref)) < 5e-6 * np.max(np.abs(ref))
    # broadcast of the smaller batch (LinearAlgebra.php:1066-1077): B shared by the outer repeats
    B2 = rng.standard_normal((8, 32, 40)).astype(np.float32)
    got2 = la.matmul(a, la.array(B2))
    ref2 = np.asarray(lo.matmul(lo.array(A), lo.array(B2)).toArray(), np.float64)
    assert np.max(np.abs(got2.to_numpy() - ref2)) < 5e-6 * np.max(np.abs(ref2))


# ---------------------------------------------------------------------------------------------------
# BLAS level-1 reductions and swap (PhpBlas.php:196-222, :266-348, :366-392, :744-760)
# ---------------------------------------------------------------------------------------------------
RED = VEC + [((1 << 24) + 5, 1, 1, 0, 0)]
RED_DT = {dtypes.float32: np.float32, dtypes.float64: np.float64, dtypes.int32: np.int32}


@pytest.mark.parametrize("n,incX,incY,offX,offY", RED, ids=[str(v) for v in RED])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32])
def test_level1_reductions_vs_oracle(n, incX, incY, offX, offY, dt, cuda_service):
    """asum / nrm2 / dot: terms and sums in double on both sides; the device adds them as a tree, the oracle in
    index order -- 1e-12 relative to the sum of magnitudes covers the reassociation.  iamax / iamin: exact."""
    rng = np.random.default_rng(n * 7 + incX)
    npdt = RED_DT[dt]
    blas = cuda_service.blas()
    if dt == dtypes.int32:
        x = rng.integers(-1000, 1000, offX + (n - 1) * incX + 4).astype(npdt)
        y = rng.integers(-1000, 1000, offY + (n - 1) * incY + 4).astype(npdt)
    else:
        x, y = _vec(rng, n, incX, offX, npdt), _vec(rng, n, incY, offY, npdt)
    X, Y = buf(x, dt), buf(y, dt)
    xo, yo = oracle.as_buffer(x), oracle.as_buffer(y)
    xs, ys = xo[offX:offX + (n - 1) * incX + 1:incX], yo[offY:offY + (n - 1) * incY + 1:incY]
    mag = float(np.sum(np.abs(xs)))
    assert abs(blas.asum(n, X, offX, incX) - oracle.asum(n, xo, offX, incX)) <= 1e-12 * mag
    ref = oracle.nrm2(n, xo, offX, incX)
    assert abs(blas.nrm2(n, X, offX, incX) - ref) <= 1e-12 * max(ref, 1.0)
    assert abs(blas.dot(n, X, offX, incX, Y, offY, incY) - oracle.dot(n, xo, offX, incX, yo, offY, incY)) <= \
        1e-12 * float(np.sum(np.abs(xs * ys)))
    assert blas.iamax(n, X, offX, incX) == oracle.iamax(n, xo, offX, incX)
    assert blas.iamin(n, X, offX, incX) == oracle.iamin(n, xo, offX, incX)
    assert np.array_equal(X.to_numpy(), x) and np.array_equal(Y.to_numpy(), y)       # operands untouched


def test_iamax_iamin_ties_and_nonfinite(cuda_service):
    """First index on ties; a NaN at element 0 stays, a later NaN never wins (PhpBlas.php:296-314, :327-345)."""
    blas = cuda_service.blas()
    nan, inf = float("nan"), float("inf")
    cases = [[3, -3, 3, 1], [1, -5, 5, 5], [0, 0, 0], [nan, 9, -10], [1, nan, -7, nan, 7], [2, -inf, inf, 1],
             [5, nan, nan], [nan, nan], [-0.0, 0.0, 1]]
    big = np.ones(100003, np.float32)
    big[[77, 5000, 99999]] = [-4, 4, 4]
    cases.append(big.tolist())
    for c in cases:
        x = np.asarray(c, np.float32)
        X = buf(x, dtypes.float32)
        xo = oracle.as_buffer(x)
        assert blas.iamax(x.size, X, 0, 1) == oracle.iamax(x.size, xo, 0, 1), c[:8]
        assert blas.iamin(x.size, X, 0, 1) == oracle.iamin(x.size, xo, 0, 1), c[:8]


@pytest.mark.parametrize("n,incX,incY,offX,offY", VEC, ids=[str(v) for v in VEC])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_swap_vs_oracle(n, incX, incY, offX, offY, dt, cuda_service):
    rng = np.random.default_rng(n + 3 * incY)
    npdt = NP[dt]
    x, y = _vec(rng, n, incX, offX, npdt), _vec(rng, n, incY, offY, npdt)
    X, Y = buf(x, dt), buf(y, dt)
    cuda_service.blas().swap(n, X, offX, incX, Y, offY, incY)
    xr, yr = oracle.as_buffer(x), oracle.as_buffer(y)
    oracle.swap(n, xr, offX, incX, yr, offY, incY)
    assert np.array_equal(X.to_numpy(), xr.astype(npdt)) and np.array_equal(Y.to_numpy(), yr.astype(npdt))


def test_level1_argument_errors(cuda_service):
    blas = cuda_service.blas()
    X = buf(np.zeros(4, np.float32), dtypes.float32)
    with pytest.raises(InvalidArgumentException, match="Argument n must be greater than 0."):
        blas.asum(0, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for bufferX."):
        blas.nrm2(5, X, 0, 1)
    with pytest.raises(InvalidArgumentException, match="Vector specification too large for bufferY."):
        blas.dot(3, X, 0, 1, X, 2, 1)


@pytest.mark.parametrize("trans", [False, True])
@pytest.mark.parametrize("m,n,pad", [(1, 1, 0), (5, 7, 0), (33, 65, 3), (300, 257, 1)])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_matrixcopy_vs_oracle(trans, m, n, pad, dt, cuda_service):
    """PhpMath::matrixcopy (PhpMath.php:2778-2809) rides rb200_omatcopy: one multiply per element, bit-exact"""
    import oracle
    npdt = dtypes.NUMPY[dt]
    rng = np.random.default_rng(m * 31 + n)
    ldA = n + pad
    rows, cols = (n, m) if trans else (m, n)
    ldB = cols + pad
    offA, offB = 2, 1
    a = rng.standard_normal(offA + m * ldA).astype(npdt)
    b = np.full(offB + rows * ldB, 9).astype(npdt)
    A, B = CudaBuffer(a.size, dt).from_numpy(a), CudaBuffer(b.size, dt).from_numpy(b)
    cuda_service.math().matrixcopy(trans, m, n, -1.5, A, offA, ldA, B, offB, ldB)
    bo = b.astype(np.float64)
    oracle.omatcopy(101, 112 if trans else 111, m, n, -1.5, a.astype(np.float64), offA, ldA, bo, offB, ldB)
    assert np.array_equal(B.to_numpy().astype(np.float64), bo.astype(npdt).astype(np.float64))
    with pytest.raises(InvalidArgumentException, match="bufferB"):
        cuda_service.math().matrixcopy(trans, m, n, 1.0, A, offA, ldA, B, offB + pad + 1, ldB)
