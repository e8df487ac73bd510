"""Parity of the CUDA path (through the C ABI, via the host mirror classes) against the CPU oracle
and the reference's golden vectors.  Every test needs a B200:  pytest -m gpu

Tolerances (stated per test):
  * integer / index / copy work (argmax, im2col2d, max): bit-exact;
  * multiply, add with |alpha| == 1: bit-exact against the oracle's double result rounded once to
    float32 (a float32 product is exact in double; double rounding of a sum of two float32 is
    innocuous since 53 >= 2*24+2);
  * everything else in float32: the reference's own isclose bound, max|d| < 1e-7 + 1e-4*max|ref|
    (src/LinearAlgebra.php:5439-5464), with the much tighter measured bound asserted next to it;
  * 1xTF32 GEMM: its own stated bound (2e-3 * max|C|), outside the reference bound by design.
"""
import json
import os

import numpy as np
import pytest

import golden_runner as gr
import oracle
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, LinearAlgebra, MatrixOperator, NDArray, dtypes

pytestmark = pytest.mark.gpu

KA = gr.load_known_answers()
with open(os.path.join(gr.HERE, "golden", "im2col2d_cases.json")) as f:
    IM2COL = json.load(f)["cases"]
with open(os.path.join(gr.HERE, "golden", "im2col1d_cases.json")) as f:
    IM2COL1D = json.load(f)["cases"]
with open(os.path.join(gr.HERE, "golden", "im2col3d_cases.json")) as f:
    IM2COL3D = json.load(f)["cases"]


def isclose_ref(got, ref, rtol=1e-4, atol=1e-7):
    """LinearAlgebra::isclose (src/LinearAlgebra.php:5439-5464): max-norm test."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return np.max(np.abs(got - ref)) < atol + rtol * np.max(np.abs(ref))


def buf(a, dtype=dtypes.float32):
    a = np.asarray(a)
    return CudaBuffer(a.size, dtype).from_numpy(a.reshape(-1))


# ---------------------------------------------------------------------------------------------------
# golden vectors of the reference's PHPUnit suite, through the CUDA service
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KA["cases"], ids=[c["id"] for c in KA["cases"]])
def test_known_answer(case, cuda_service):
    gr.run_case(case, cuda_service)


@pytest.mark.parametrize("mode", [rb.GEMM_FP32_SIMT, rb.GEMM_TF32, rb.GEMM_TF32X3])
def test_known_answer_gemm_all_modes(mode):
    svc = rb.MatlibCuda(gemm_mode=mode)
    for case in KA["cases"]:
        if case["op"] in ("gemm", "cross"):
            gr.run_case(case, svc)


@pytest.mark.parametrize("case", IM2COL, ids=[c["name"].replace(" ", "_") for c in IM2COL])
def test_im2col2d_provider(case, cuda_service):
    gr.run_im2col_case(case["params"], cuda_service)


# ---------------------------------------------------------------------------------------------------
# the reference's "Large" tests (run only on accelerated drivers there)
# ---------------------------------------------------------------------------------------------------
def _large(cid):
    return next(c for c in KA["large"]["cases"] if c["id"] == cid)


@pytest.mark.parametrize("cid", ["multiply_large", "add_large"])
def test_large_elementwise(cid, cuda_service):
    c = _large(cid)
    la = LinearAlgebra(cuda_service)
    x = la.fill(c["x_fill"], la.alloc([c["rows"], c["cols"]]))
    y = la.fill(c["a_fill"], la.alloc([c["rows"], c["cols"]]))
    r = getattr(la, c["op"])(x, y)
    assert np.max(np.abs(r.to_numpy() - c["expect_fill"])) < 1e-3


@pytest.mark.parametrize("cid", ["reducesum_large_wide", "reducesum_large_tall", "reducemax_large",
                                 "reduceargmax_large"])
def test_large_reductions(cid, cuda_service):
    c = _large(cid)
    la = LinearAlgebra(cuda_service)
    x = la.fill(c["fill"], la.alloc([c["rows"], c["cols"]]))
    r = getattr(la, c["op"])(x, axis=1)
    assert r.shape() == [c["rows"]]
    assert np.max(np.abs(r.to_numpy().astype(np.float64) - c["expect_fill"])) < 1e-3


def test_large_reducemax_nan_table(cuda_service):
    c = _large("reducemax_large_nan_sticky")
    la = LinearAlgebra(cuda_service)
    for cols in c["cols_list"]:
        x = np.zeros((6, cols), np.float32)
        for r, head in enumerate(c["heads"]):
            x[r, :2] = [gr._num(v) for v in head]
        got = la.reduceMax(la.array(x), axis=1).to_numpy()
        gr._assert_equal(c["expect"], got)


@pytest.mark.parametrize("cid", ["softmax_large_wide", "softmax_large_tall"])
def test_large_softmax(cid, cuda_service):
    c = _large(cid)
    la = LinearAlgebra(cuda_service)
    x = la.ones(la.alloc([c["rows"], c["cols"]]))
    r = la.softmax(x).to_numpy()
    assert np.max(np.abs(r - 1.0 / c["cols"])) < 1e-3


def test_large_im2col2d(cuda_service):
    c = _large("im2col2d_large")
    la = LinearAlgebra(cuda_service)
    n = c["im_h"] * c["im_w"] * c["channels"]
    img = np.arange(n, dtype=np.float32).reshape(1, c["im_h"], c["im_w"], c["channels"])
    images = la.array(img)
    cols = la.im2col(images, filterSize=[3, 3])
    got = cols.to_numpy()
    # index identity: cols[b,oy,ox,ky,kx,c] = images[b,oy+ky,ox+kx,c]  (values < 2^24: exact in f32)
    win = np.lib.stride_tricks.sliding_window_view(img, (3, 3), axis=(1, 2))      # [b,oy,ox,c,ky,kx]
    assert np.array_equal(got, win.transpose(0, 1, 2, 4, 5, 3))
    back = la.col2im(cols, la.zerosLike(images), filterSize=[3, 3]).to_numpy()
    counts = np.zeros_like(img, dtype=np.float64)
    oh, ow = c["im_h"] - 2, c["im_w"] - 2
    for ky in range(3):
        for kx in range(3):
            counts[:, ky:ky + oh, kx:kx + ow, :] += 1
    # each pixel is summed `count` times; compare in the isclose bound (values up to 3e6 * 9)
    assert isclose_ref(back, img.astype(np.float64) * counts)


# ---------------------------------------------------------------------------------------------------
# add / multiply vs oracle
# ---------------------------------------------------------------------------------------------------
EW_SHAPES = [  # (m, n, trans, ldA_extra, offX, offA, incX)
    (64, 256, False, 0, 0, 0, 1), (64, 256, True, 0, 0, 0, 1), (1, 1, False, 0, 0, 0, 1),
    (37, 129, False, 0, 0, 0, 1), (37, 129, True, 0, 0, 0, 1), (128, 64, False, 4, 8, 16, 1),
    (33, 100, False, 3, 5, 7, 2), (33, 100, True, 3, 5, 7, 3), (1000, 16, False, 0, 0, 0, 1),
    (512, 4096, False, 0, 0, 0, 1), (512, 4096, True, 0, 0, 0, 1),
]


@pytest.mark.parametrize("shape", EW_SHAPES, ids=[str(s) for s in EW_SHAPES])
@pytest.mark.parametrize("op,alpha", [("add", 1.0), ("add", -1.0), ("add", 0.3), ("multiply", None)])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_elementwise_vs_oracle(shape, op, alpha, dt, cuda_service):
    m, n, trans, ld_extra, offX, offA, incX = shape
    rng = np.random.default_rng(2001 + m + n)
    npdt = dtypes.NUMPY[dt]
    ldA = n + ld_extra
    xlen = m if trans else n
    Xh = rng.uniform(-1, 1, offX + (xlen - 1) * incX + 1).astype(npdt)
    Ah = rng.uniform(-1, 1, offA + m * ldA).astype(npdt)
    math = cuda_service.math()
    X, A = buf(Xh, dt), buf(Ah, dt)
    xo, ao = oracle.as_buffer(Xh), oracle.as_buffer(Ah)
    if op == "add":
        math.add(trans, m, n, alpha, X, offX, incX, A, offA, ldA)
        oracle.add(trans, m, n, alpha, xo, offX, incX, ao, offA, ldA)
    else:
        math.multiply(trans, m, n, X, offX, incX, A, offA, ldA)
        oracle.multiply(trans, m, n, xo, offX, incX, ao, offA, ldA)
    got = A.to_numpy()
    assert np.array_equal(X.to_numpy(), Xh)                       # X untouched
    if dt == dtypes.float64 and not (op == "add" and alpha == 0.3):
        assert np.array_equal(got, ao)                            # same double arithmetic
    elif dt == dtypes.float32 and (op == "multiply" or abs(alpha) == 1.0):
        assert np.array_equal(got, ao.astype(np.float32))         # bit-exact vs oracle rounded once
    else:
        # general alpha: the device uses one fused multiply-add in the operand precision (alpha is cast
        # to it), the oracle a separate double multiply and add: |d| <= 2 eps (|alpha x| + |a|)
        eps = float(np.finfo(npdt).eps)
        xs = Xh[offX::incX][: (m if trans else n)].astype(np.float64)
        xb = np.repeat(xs, n).reshape(m, n) if trans else np.tile(xs, (m, 1))
        a0 = np.stack([Ah[offA + i * ldA: offA + i * ldA + n] for i in range(m)]).astype(np.float64)
        bound = 2 * eps * (np.abs(alpha * xb) + np.abs(a0)) + 1e-300
        g = np.stack([got[offA + i * ldA: offA + i * ldA + n] for i in range(m)]).astype(np.float64)
        r = np.stack([ao[offA + i * ldA: offA + i * ldA + n] for i in range(m)])
        assert np.all(np.abs(g - r) <= bound)
    # elements outside the m x n view (ld padding, before offA) untouched
    mask = np.ones(Ah.size, bool)
    for i in range(m):
        mask[offA + i * ldA: offA + i * ldA + n] = False
    assert np.array_equal(got[mask], Ah[mask])


def test_elementwise_int32(cuda_service):
    math = cuda_service.math()
    Xh = np.arange(8, dtype=np.int32)
    Ah = np.arange(32, dtype=np.int32).reshape(4, 8) * 3
    X, A = buf(Xh, dtypes.int32), buf(Ah, dtypes.int32)
    math.add(False, 4, 8, 2, X, 0, 1, A, 0, 8)
    assert np.array_equal(A.to_numpy().reshape(4, 8), Ah + 2 * Xh)
    math.multiply(False, 4, 8, X, 0, 1, A, 0, 8)
    assert np.array_equal(A.to_numpy().reshape(4, 8), (Ah + 2 * Xh) * Xh)


def test_elementwise_errors(cuda_service):
    math = cuda_service.math()
    X, A = buf(np.zeros(3)), buf(np.zeros(6))
    with pytest.raises(rb.InvalidArgumentException, match="Vector specification too large for buffer."):
        math.add(False, 2, 4, 1.0, X, 0, 1, A, 0, 4)
    with pytest.raises(rb.InvalidArgumentException, match="Vector specification too large for buffer."):
        math.multiply(False, 3, 3, X, 0, 1, A, 0, 3)


def test_add_full_config_size(cuda_service):
    """BASELINE config C2: [8192,4096] (+|x) [4096], checked against the oracle on every element."""
    m, n = 8192, 4096
    Ah = np.random.default_rng(2001).uniform(-1, 1, m * n).astype(np.float32)
    Xh = np.random.default_rng(2002).uniform(-1, 1, n).astype(np.float32)
    math = cuda_service.math()
    X, A = buf(Xh), buf(Ah)
    math.add(False, m, n, 1.0, X, 0, 1, A, 0, n)
    xo, ao = oracle.as_buffer(Xh), oracle.as_buffer(Ah)
    oracle.add(False, m, n, 1.0, xo, 0, 1, ao, 0, n)
    assert np.array_equal(A.to_numpy(), ao.astype(np.float32))
    math.multiply(False, m, n, X, 0, 1, A, 0, n)
    a2 = ao.astype(np.float32).astype(np.float64)
    oracle.multiply(False, m, n, xo, 0, 1, a2, 0, n)
    assert np.array_equal(A.to_numpy(), a2.astype(np.float32))


# ---------------------------------------------------------------------------------------------------
# reductions vs oracle
# ---------------------------------------------------------------------------------------------------
RD_SHAPES = [  # (m, n, k)
    (1, 1, 1), (7, 5, 1), (64, 1024, 1), (3000, 1024, 1), (5, 100000, 1), (64, 4097, 1), (2, 1 << 20, 1),
    (4, 6, 3), (16, 33, 7), (1, 4096, 1024), (8, 512, 64), (3, 1000, 2), (2, 70000, 5), (300, 3, 1000),
]


def _rd_input(m, n, k, seed, quantise=False):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, m * n * k).astype(np.float32)
    if quantise:
        a = (np.round(a * 64) / 64).astype(np.float32)       # many exact ties
    return a


@pytest.mark.parametrize("shape", RD_SHAPES, ids=[str(s) for s in RD_SHAPES])
@pytest.mark.parametrize("off", [0, 3])
def test_reduce_sum_vs_oracle(shape, off, cuda_service):
    m, n, k = shape
    a = np.concatenate([np.zeros(off, np.float32), _rd_input(m, n, k, 3001)])
    A, B = buf(a), CudaBuffer(m * k + 2, dtypes.float32)
    cuda_service.math().reduceSum(m, n, k, A, off, B, 2)
    ao, bo = oracle.as_buffer(a), np.zeros(m * k + 2)
    oracle.reduce_sum(m, n, k, ao, off, bo, 2)
    got = B.to_numpy().astype(np.float64)
    assert np.all(got[:2] == 0)
    assert isclose_ref(got, bo)
    # measured-tight bound: float32 tree sum vs sequential double sum
    scale = np.abs(a.astype(np.float64)).reshape(-1)[off:].reshape(m, n, k).sum(axis=1).reshape(-1)
    assert np.all(np.abs(got[2:] - bo[2:]) <= 2e-6 * scale + 1e-30)


@pytest.mark.parametrize("shape", RD_SHAPES, ids=[str(s) for s in RD_SHAPES])
def test_reduce_max_vs_oracle_bit_exact(shape, cuda_service):
    m, n, k = shape
    a = _rd_input(m, n, k, 3002)
    A, B = buf(a), CudaBuffer(m * k, dtypes.float32)
    cuda_service.math().reduceMax(m, n, k, A, 0, B, 0)
    bo = np.zeros(m * k)
    oracle.reduce_max(m, n, k, oracle.as_buffer(a), 0, bo, 0)          # PHP rule == sticky rule: no NaN
    assert np.array_equal(B.to_numpy().astype(np.float64), bo)


@pytest.mark.parametrize("shape", RD_SHAPES, ids=[str(s) for s in RD_SHAPES])
@pytest.mark.parametrize("idx_dtype", [dtypes.int32, dtypes.int64])
def test_reduce_argmax_vs_oracle_bit_exact(shape, idx_dtype, cuda_service):
    m, n, k = shape
    a = _rd_input(m, n, k, 3003, quantise=True)
    A, B = buf(a), CudaBuffer(m * k, idx_dtype)
    cuda_service.math().reduceArgMax(m, n, k, A, 0, B, 0)
    bo = np.zeros(m * k, np.int64)
    oracle.reduce_argmax(m, n, k, oracle.as_buffer(a), 0, bo, 0)
    assert np.array_equal(B.to_numpy().astype(np.int64), bo)


def test_reduce_nan_inf_semantics(cuda_service):
    math = cuda_service.math()
    rng = np.random.default_rng(5)
    for (m, n, k) in [(6, 1024, 1), (6, 50000, 1), (2, 300, 8), (1, 9000, 16)]:
        a = rng.uniform(-1, 1, (m, n, k)).astype(np.float32)
        a[0, n // 3, 0] = np.nan             # NaN mid-run
        a[m - 1, 0, k - 1] = np.nan          # NaN at position 0
        a[0, n - 1, k - 1] = np.inf
        if m > 2:
            a[1, :, 0] = -np.inf
            a[2, 5, 0] = np.inf
            a[2, 7, 0] = np.inf              # +inf tie: lowest index wins
        A = buf(a)
        Bm, Bi = CudaBuffer(m * k, dtypes.float32), CudaBuffer(m * k, dtypes.int32)
        math.reduceMax(m, n, k, A, 0, Bm, 0)
        math.reduceArgMax(m, n, k, A, 0, Bi, 0)
        ao = oracle.as_buffer(a)
        bm, bi = np.zeros(m * k), np.zeros(m * k, np.int64)
        oracle.reduce_max(m, n, k, ao, 0, bm, 0, sticky=True)
        oracle.reduce_argmax(m, n, k, ao, 0, bi, 0)
        assert np.array_equal(Bm.to_numpy().astype(np.float64), bm, equal_nan=True)
        assert np.array_equal(Bi.to_numpy().astype(np.int64), bi)


def test_reduce_float64(cuda_service):
    m, n, k = 9, 777, 3
    a = np.random.default_rng(11).standard_normal(m * n * k)
    A = buf(a, dtypes.float64)
    Bs, Bm = CudaBuffer(m * k, dtypes.float64), CudaBuffer(m * k, dtypes.float64)
    cuda_service.math().reduceSum(m, n, k, A, 0, Bs, 0)
    cuda_service.math().reduceMax(m, n, k, A, 0, Bm, 0)
    bs, bm = np.zeros(m * k), np.zeros(m * k)
    oracle.reduce_sum(m, n, k, a, 0, bs, 0)
    oracle.reduce_max(m, n, k, a, 0, bm, 0)
    assert np.allclose(Bs.to_numpy(), bs, rtol=1e-13, atol=1e-13)
    assert np.array_equal(Bm.to_numpy(), bm)


def test_reduce_full_config_size(cuda_service):
    """BASELINE config C3: [65536,1024] axis=1 -- all three reductions against the oracle."""
    m, n = 65536, 1024
    a = np.random.default_rng(3001).uniform(-1, 1, m * n).astype(np.float32)
    aq = (np.round(a * 64) / 64).astype(np.float32)
    math = cuda_service.math()
    A, Aq = buf(a), buf(aq)
    Bs, Bm, Bi = CudaBuffer(m, dtypes.float32), CudaBuffer(m, dtypes.float32), CudaBuffer(m, dtypes.int32)
    math.reduceSum(m, n, 1, A, 0, Bs, 0)
    math.reduceMax(m, n, 1, A, 0, Bm, 0)
    math.reduceArgMax(m, n, 1, Aq, 0, Bi, 0)
    ao = oracle.as_buffer(a)
    bs, bm, bi = np.zeros(m), np.zeros(m), np.zeros(m, np.int64)
    oracle.reduce_sum(m, n, 1, ao, 0, bs, 0)
    oracle.reduce_max(m, n, 1, ao, 0, bm, 0)
    oracle.reduce_argmax(m, n, 1, oracle.as_buffer(aq), 0, bi, 0)
    assert isclose_ref(Bs.to_numpy(), bs)
    assert np.max(np.abs(Bs.to_numpy() - bs)) < 2e-5             # measured-tight (|sum| <~ 100)
    assert np.array_equal(Bm.to_numpy().astype(np.float64), bm)
    assert np.array_equal(Bi.to_numpy().astype(np.int64), bi)
    # axis = 0 of the same matrix (m=1, n=65536, k=1024): the split + merge path
    B0 = CudaBuffer(n, dtypes.float32)
    math.reduceSum(1, m, n, A, 0, B0, 0)
    b0 = np.zeros(n)
    oracle.reduce_sum(1, m, n, ao, 0, b0, 0)
    assert isclose_ref(B0.to_numpy(), b0)


# ---------------------------------------------------------------------------------------------------
# softmax vs oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,ld_extra,off", [(1, 1, 0, 0), (5, 5, 0, 0), (64, 1024, 0, 0), (33, 1000, 0, 0),
                                              (17, 513, 3, 5), (128, 128, 4, 8), (9, 5000, 0, 0),
                                              (3, 70001, 0, 1), (2000, 64, 0, 0)])
def test_softmax_vs_oracle(m, n, ld_extra, off, cuda_service):
    ld = n + ld_extra
    a = np.random.default_rng(3001).uniform(-4, 4, off + m * ld).astype(np.float32)
    A = buf(a)
    cuda_service.math().softmax(m, n, A, off, ld)
    ao = oracle.as_buffer(a)
    oracle.softmax(m, n, ao, off, ld)
    got = A.to_numpy().astype(np.float64)
    assert isclose_ref(got, ao)
    view = np.zeros(a.size, bool)
    for i in range(m):
        view[off + i * ld: off + i * ld + n] = True
    assert np.array_equal(got[~view], ao[~view])                  # padding untouched
    rel = np.abs(got[view] - ao[view]) / ao[view]
    assert rel.max() < 5e-6                                       # measured-tight: expf + f32 tree sum
    rows = np.array([got[off + i * ld: off + i * ld + n].sum() for i in range(m)])
    assert np.allclose(rows, 1.0, atol=1e-5)


def test_softmax_special_values(cuda_service):
    x = np.array([[0.0, np.nan, 1.0, 2.0], [np.inf, 1.0, 2.0, 3.0], [-np.inf, 0.0, 0.0, -np.inf],
                  [-np.inf] * 4, [1000.0, 999.0, -1000.0, 0.0]], np.float32)
    A = buf(x)
    cuda_service.math().softmax(5, 4, A, 0, 4)
    ao = oracle.as_buffer(x)
    oracle.softmax(5, 4, ao, 0, 4)
    got = A.to_numpy().astype(np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(ao))
    ok = ~np.isnan(ao)
    assert np.allclose(got[ok], ao[ok], rtol=1e-6, atol=1e-30)


def test_softmax_full_config_size(cuda_service):
    m, n = 65536, 1024
    a = np.random.default_rng(3001).uniform(-1, 1, m * n).astype(np.float32)
    A = buf(a)
    cuda_service.math().softmax(m, n, A, 0, n)
    ao = oracle.as_buffer(a)
    oracle.softmax(m, n, ao, 0, n)
    got = A.to_numpy().astype(np.float64)
    assert isclose_ref(got, ao)
    assert np.max(np.abs(got - ao) / ao) < 5e-6


# ---------------------------------------------------------------------------------------------------
# GEMM vs oracle
# ---------------------------------------------------------------------------------------------------
GEMM_TOL = {rb.GEMM_FP32_SIMT: 2e-6, rb.GEMM_TF32X3: 5e-5, rb.GEMM_TF32: 2e-3}   # x max|C|


def _gemm_case(svc, mode, m, n, k, tA, tB, alpha, beta, offA=0, offB=0, offC=0, ld_extra=0, seed=5000,
               dt=dtypes.float32):
    rng = np.random.default_rng(seed)
    npdt = dtypes.NUMPY[dt]
    rA, cA = (k, m) if tA else (m, k)
    rB, cB = (n, k) if tB else (k, n)
    ldA, ldB, ldC = cA + ld_extra, cB + ld_extra, n + ld_extra
    Ah = rng.uniform(-1, 1, offA + rA * ldA).astype(npdt)
    Bh = rng.uniform(-1, 1, offB + rB * ldB).astype(npdt)
    Ch = rng.uniform(-1, 1, offC + m * ldC).astype(npdt)
    A, B, C = buf(Ah, dt), buf(Bh, dt), buf(Ch, dt)
    blas = svc.blas()
    blas.setGemmMode(mode)
    blas.gemm(BLAS.RowMajor, BLAS.Trans if tA else BLAS.NoTrans, BLAS.Trans if tB else BLAS.NoTrans,
              m, n, k, alpha, A, offA, ldA, B, offB, ldB, beta, C, offC, ldC)
    path = rb._capi.last_gemm_path()
    co = oracle.as_buffer(Ch)
    if m * n * k <= 1 << 18:
        ao, bo = oracle.as_buffer(Ah), oracle.as_buffer(Bh)
        oracle.gemm(oracle.RowMajor, oracle.Trans if tA else oracle.NoTrans, oracle.Trans if tB else oracle.NoTrans,
                    m, n, k, alpha, ao, offA, ldA, bo, offB, ldB, beta, co, offC, ldC)
    else:
        # Larger cases: materialise op(A), op(B) (pure data movement) and use the interleaved-loop
        # oracle, which is bit-identical to ro_gemm per element (tests/test_oracle_golden.py::
        # test_gemm_rows_variants_bit_identical) but cache friendly.
        Am = Ah[offA:offA + rA * ldA].reshape(rA, ldA)[:, :cA]
        Bm = Bh[offB:offB + rB * ldB].reshape(rB, ldB)[:, :cB]
        opA = oracle.as_buffer(Am.T if tA else Am)
        opB = oracle.as_buffer(Bm.T if tB else Bm)
        oracle.gemm_rows_fast(0, m, n, k, alpha, opA, k, opB, n, beta, co[offC:], ldC)
    got = C.to_numpy().astype(np.float64)
    view = np.zeros(Ch.size, bool)
    for i in range(m):
        view[offC + i * ldC: offC + i * ldC + n] = True
    assert np.array_equal(got[~view], co[~view])                   # nothing outside C's m x n view
    err = np.max(np.abs(got[view] - co[view])) / max(np.max(np.abs(co[view])), 1e-30)
    return err, path


SMALL_GEMMS = [(3, 3, 3), (1, 1, 1), (2, 3, 4), (17, 9, 33), (64, 64, 64), (65, 130, 31), (100, 257, 129)]


@pytest.mark.parametrize("mnk", SMALL_GEMMS, ids=[str(s) for s in SMALL_GEMMS])
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_gemm_simt_small(mnk, tA, tB, dt, cuda_service):
    m, n, k = mnk
    for alpha, beta, off, ldx in ((1.0, 0.0, 0, 0), (1.0, 1.0, 0, 0), (-0.5, 2.0, 9, 3)):
        err, path = _gemm_case(cuda_service, rb.GEMM_FP32_SIMT, m, n, k, tA, tB, alpha, beta, off, off + 1, off + 2,
                               ldx, dt=dt)
        assert path == ("simt_f32" if dt == dtypes.float32 else "simt_f64")
        assert err < (2e-6 if dt == dtypes.float32 else 1e-14), (err, path)


BIG_SIMT = [(1536, 1664, 300), (1500, 1303, 257), (2048, 1152, 72), (1281, 1290, 33), (3000, 700, 129)]


@pytest.mark.parametrize("mnk", BIG_SIMT, ids=[str(s) for s in BIG_SIMT])
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64])
def test_gemm_simt_big_tiles(mnk, tA, tB, dt, cuda_service):
    """The large-tile kernels -- FFMA 128 x 128 x 16 for float32, DMMA (mma.sync m16n8k8 f64) 128 x 64 x 16 for float64,
    both double-buffered: aligned operands take the 128-bit loads, odd leading dimensions / offsets the scalar loads;
    M, N, K tails; beta != 0; all transposes.  RB200_DGEMM_DMMA=0 keeps the DFMA kernel (covered by the env test)."""
    m, n, k = mnk
    for alpha, beta, off, ldx in ((1.0, 0.0, 0, 0), (-0.5, 2.0, 9, 3), (1.0, 1.0, 4, 4)):
        err, path = _gemm_case(cuda_service, rb.GEMM_FP32_SIMT, m, n, k, tA, tB, alpha, beta, off, off + 1, off + 2,
                               ldx, dt=dt)
        want = ("simt_f32_big",) if dt == dtypes.float32 else ("dmma_f64",)        # float64: FP64 tensor pipe (DMMA)
        if dt == dtypes.float32 and k <= 64 and not tA and not tB:
            want += ("simt_f32_shortk",)                     # short reductions keep their own kernel
        assert path in want, path
        assert err < (2e-6 if dt == dtypes.float32 else 1e-14), (err, path)


def test_dgemm_dfma_kernel_still_agrees(cuda_service):
    """The DFMA form of the large-tile kernel (RB200_DGEMM_DMMA=0 in a fresh process: the switch is read once)."""
    import subprocess
    import sys
    code = ("import os, sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import oracle, rindow_math_matrix_b200 as rb\n"
            "from rindow_math_matrix_b200 import BLAS, CudaBuffer, dtypes, _capi\n"
            "blas = rb.MatlibCuda().blas(); rng = np.random.default_rng(1); m, n, k = 1500, 1303, 257\n"
            "a = rng.uniform(-1, 1, (m, k)); b = rng.uniform(-1, 1, (k, n))\n"
            "A = CudaBuffer(m * k, dtypes.float64).from_numpy(a.reshape(-1)); B = CudaBuffer(k * n, dtypes.float64).from_numpy(b.reshape(-1))\n"
            "C = CudaBuffer(m * n, dtypes.float64)\n"
            "blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, 0, n)\n"
            "assert _capi.last_gemm_path() == 'simt_f64_big', _capi.last_gemm_path()\n"
            "ref = np.zeros(m * n); oracle.gemm_rows_fast(0, m, n, k, 1.0, oracle.as_buffer(a), k, oracle.as_buffer(b), n, 0.0, ref, n)\n"
            "err = np.max(np.abs(C.to_numpy() - ref)) / np.max(np.abs(ref)); assert err < 1e-14, err\n"
            % (os.path.dirname(gr.HERE), gr.HERE))
    env = dict(os.environ, RB200_DGEMM_DMMA="0")
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]


TC_GEMMS = [(256, 256, 256), (128, 128, 32), (128, 256, 64), (384, 640, 160), (512, 384, 1024), (200, 300, 100),
            (1000, 520, 68), (129, 257, 33), (2048, 128, 128), (128, 2048, 1024)]


@pytest.mark.parametrize("mnk", TC_GEMMS, ids=[str(s) for s in TC_GEMMS])
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("mode", [rb.GEMM_TF32, rb.GEMM_TF32X3])
def test_gemm_tcgen05(mnk, tA, tB, mode, cuda_service):
    m, n, k = mnk
    # TMA needs leading dimensions that are multiples of 4 floats: pad via ld_extra
    rA_c = m if tA else k
    rB_c = k if tB else n
    ldx = 0
    while (rA_c + ldx) % 4 or (rB_c + ldx) % 4 or (n + ldx) % 4:
        ldx += 1
    for alpha, beta, off in ((1.0, 0.0, 0), (1.0, 1.0, 0), (0.75, -1.5, 8)):
        err, path = _gemm_case(cuda_service, mode, m, n, k, tA, tB, alpha, beta, off, off, off, ldx)
        assert path in (("tcgen05_tf32", "tcgen05_tf32_2cta", "tcgen05_tf32_128", "tcgen05_tf32_64") if mode == rb.GEMM_TF32 else ("tcgen05_tf32x3", "tcgen05_tf32x3_2cta", "tcgen05_tf32x3_64")), path
        assert err < GEMM_TOL[mode], (err, path, mnk, tA, tB, alpha, beta)
        if mode == rb.GEMM_TF32X3:
            assert err < 1e-4            # the reference's own isclose rtol


def test_gemm_tf32x3_long_k_accumulator_drift(cuda_service):
    """The TMEM accumulator truncates; all-positive data exposes the drift (linear in K).  The 3xTF32
    mode chunks K (<= 2048 per pass, RN adds between passes), so it must stay far inside 1e-4."""
    rng = np.random.default_rng(3)
    m, n = 128, 256
    blas = cuda_service.blas()
    for k in (2048 + 40, 8192, 16384):
        Ah = rng.uniform(0.5, 1.0, (m, k)).astype(np.float32)
        Bh = rng.uniform(0.5, 1.0, (k, n)).astype(np.float32)
        ref = Ah.astype(np.float64) @ Bh.astype(np.float64)
        A, B, C = buf(Ah), buf(Bh), CudaBuffer(m * n, dtypes.float32)
        blas.setGemmMode(rb.GEMM_TF32X3)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, 0, n)
        assert rb._capi.last_gemm_path() in ("tcgen05_tf32x3", "tcgen05_tf32x3_2cta", "tcgen05_tf32x3_64")
        rel = np.abs(C.to_numpy().reshape(m, n) - ref) / ref
        assert rel.max() < 3e-5, (k, rel.max())          # ~7e-9 * min(K, 2048) drift, see DESIGN.md 4.1
        # beta / alpha survive the chunking: C = 0.5*A*B + 2*C0
        C0 = rng.uniform(-1, 1, (m, n)).astype(np.float32)
        C = buf(C0)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 0.5, A, 0, k, B, 0, n, 2.0, C, 0, n)
        ref2 = 0.5 * ref + 2.0 * C0
        assert np.max(np.abs(C.to_numpy().reshape(m, n) - ref2)) / np.max(np.abs(ref2)) < 3e-5
        # the 1xTF32 mode documents the drift instead (inside its own 2e-3 bound)
        blas.setGemmMode(rb.GEMM_TF32)
        C = CudaBuffer(m * n, dtypes.float32)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, 0, n)
        rel = np.abs(C.to_numpy().reshape(m, n) - ref) / ref
        assert rel.max() < 2e-3


SHORTK = [(4096, 64, 27), (1000, 100, 5), (513, 7, 64), (70000, 64, 27), (2048, 130, 1), (640, 64, 33),
          (5000, 128, 64), (3000, 33, 100), (2049, 96, 31), (2500, 1, 1)]


@pytest.mark.parametrize("mnk", SHORTK, ids=[str(s) for s in SHORTK])
def test_gemm_tall_skinny_conv_shapes(mnk, cuda_service):
    """Tall matrix x short K (conv-as-GEMM, e.g. cols [B*oh*ow, 27] x W [27, 64]) with arbitrary ld/offset:
    FFMA short-K kernel in FP32 mode, software-staged tcgen05 kernel in the TF32 modes when M >= 2048,
    N <= 128, K <= 128."""
    m, n, k = mnk
    skinny = m >= 2048 and n <= 128 and k <= 128
    for mode in (rb.GEMM_FP32_SIMT, rb.GEMM_AUTO, rb.GEMM_TF32X3, rb.GEMM_TF32):
        for alpha, beta, off, ldx in ((1.0, 0.0, 0, 0), (0.5, 1.5, 3, 1), (1.0, 1.0, 4, 4)):
            err, path = _gemm_case(cuda_service, mode, m, n, k, False, False, alpha, beta, off, off, off, ldx)
            if mode == rb.GEMM_FP32_SIMT:
                assert path in ("simt_f32_shortk", "simt_f32", "simt_f32_big") and err < 2e-6, (path, err)
            elif mode == rb.GEMM_TF32:
                if skinny:
                    assert path in ("tcgen05_tf32_skinny", "tcgen05_tf32", "tcgen05_tf32_2cta", "tcgen05_tf32_128", "tcgen05_tf32_64"), path
                assert err < 2e-3, (err, path)
            else:
                if skinny:      # aligned full-width shapes may take the TMA-fed kernel instead
                    assert path in ("tcgen05_tf32x3_skinny", "tcgen05_tf32x3", "tcgen05_tf32x3_2cta", "tcgen05_tf32x3_64"), path
                assert err < 5e-5, (err, path, mnk)
    err, path = _gemm_case(cuda_service, rb.GEMM_AUTO, m, n, k, True, False, 1.0, 0.0)
    assert "skinny" not in path and "shortk" not in path and err < 5e-5


def test_gemm_unaligned_falls_back_to_simt(cuda_service):
    err, path = _gemm_case(cuda_service, rb.GEMM_TF32, 256, 256, 256, False, False, 1.0, 0.0, offA=1)
    assert path == "simt_f32" and err < 2e-6
    err, path = _gemm_case(cuda_service, rb.GEMM_TF32, 256, 255, 256, False, False, 1.0, 0.0)
    assert path == "simt_f32" and err < 2e-6


def test_gemm_errors(cuda_service):
    blas = cuda_service.blas()
    A, B, C = buf([1, 2, 3, 4, 5]), buf(np.eye(3)), buf(np.zeros(6))
    with pytest.raises(rb.InvalidArgumentException, match="Matrix specification too large for bufferA."):
        blas.gemm(BLAS.RowMajor, BLAS.Trans, BLAS.NoTrans, 2, 3, 3, 1.0, A, 0, 2, B, 0, 3, 0.0, C, 0, 3)
    A = buf([1, 2, 3, 4, 5, 6])
    with pytest.raises(rb.InvalidArgumentException, match="Matrix specification too large for bufferB."):
        blas.gemm(BLAS.RowMajor, BLAS.Trans, BLAS.NoTrans, 2, 3, 3, 1.0, A, 0, 2, buf(np.zeros(8)), 0, 3, 0.0, C, 0, 3)
    with pytest.raises(rb.InvalidArgumentException, match="Matrix specification too large for bufferC."):
        blas.gemm(BLAS.RowMajor, BLAS.Trans, BLAS.NoTrans, 2, 3, 3, 1.0, A, 0, 2, B, 0, 3, 0.0, buf(np.zeros(5)), 0, 3)
    with pytest.raises(rb.InvalidArgumentException, match="Argument k must be greater than 0."):
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, 2, 3, 0, 1.0, A, 0, 3, B, 0, 3, 0.0, C, 0, 3)
    with pytest.raises(rb.InvalidArgumentException, match="Unknown Tranpose Code"):
        blas.gemm(BLAS.RowMajor, 7, BLAS.NoTrans, 2, 3, 3, 1.0, A, 0, 3, B, 0, 3, 0.0, C, 0, 3)


def test_gemm_beta_zero_overwrites_nan(cuda_service):
    blas = cuda_service.blas()
    for mode, sz in ((rb.GEMM_FP32_SIMT, 8), (rb.GEMM_TF32, 256), (rb.GEMM_TF32X3, 256)):
        blas.setGemmMode(mode)
        A, B = buf(np.eye(sz)), buf(np.arange(sz * sz, dtype=np.float32) % 7)
        C = buf(np.full(sz * sz, np.nan, np.float32))
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, sz, sz, sz, 1.0, A, 0, sz, B, 0, sz, 0.0, C, 0, sz)
        assert np.array_equal(C.to_numpy(), np.arange(sz * sz, dtype=np.float32) % 7)     # PhpBlas.php:940-944


@pytest.mark.parametrize("mode", [rb.GEMM_TF32, rb.GEMM_TF32X3, rb.GEMM_FP32_SIMT])
def test_gemm_integer_valued_exact(mode, cuda_service):
    """The reference's known-answer style: small integer inputs are exact in every mode."""
    rng = np.random.default_rng(1)
    m, n, k = 256, 384, 128
    Ah = rng.integers(-8, 9, (m, k)).astype(np.float32)
    Bh = rng.integers(-8, 9, (k, n)).astype(np.float32)
    blas = cuda_service.blas()
    blas.setGemmMode(mode)
    A, B, C = buf(Ah), buf(Bh), CudaBuffer(m * n, dtypes.float32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, 0, n)
    assert np.array_equal(C.to_numpy().reshape(m, n), Ah.astype(np.float64) @ Bh.astype(np.float64))


@pytest.mark.parametrize("N,mode", [(1024, rb.GEMM_TF32), (1024, rb.GEMM_TF32X3), (4096, rb.GEMM_TF32),
                                    (4096, rb.GEMM_TF32X3)])
def test_gemm_sweep_sizes_sampled(N, mode, cuda_service):
    """BASELINE config C5 shapes: N <= 1024 checked in full against the oracle, larger N on a
    sample of 64 full rows of C (oracle gemm_rows: same loop, bounded rows)."""
    lg = int(np.log2(N))
    Ah = np.random.default_rng(5000 + lg).uniform(-1, 1, N * N).astype(np.float32)
    Bh = np.random.default_rng(5100 + lg).uniform(-1, 1, N * N).astype(np.float32)
    blas = cuda_service.blas()
    blas.setGemmMode(mode)
    A, B, C = buf(Ah), buf(Bh), CudaBuffer(N * N, dtypes.float32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, N, N, N, 1.0, A, 0, N, B, 0, N, 0.0, C, 0, N)
    got = C.to_numpy().reshape(N, N).astype(np.float64)
    ao, bo = oracle.as_buffer(Ah), oracle.as_buffer(Bh)
    rows = range(N) if N <= 1024 else sorted(np.random.default_rng(9).choice(N, 64, replace=False).tolist())
    co = np.zeros(N * N)
    if N <= 1024:
        oracle.gemm_rows_fast(0, N, N, N, 1.0, ao, N, bo, N, 0.0, co, N)
    else:
        for r in rows:
            oracle.gemm_rows_fast(r, 1, N, N, 1.0, ao, N, bo, N, 0.0, co, N)
    ref = co.reshape(N, N)[list(rows)]
    err = np.max(np.abs(got[list(rows)] - ref)) / np.max(np.abs(ref))
    assert err < GEMM_TOL[mode], err
    # a wrong tile anywhere (not only in sampled rows) shows up in the column sums: 1^T C = (1^T A) B
    colsum = Ah.reshape(N, N).astype(np.float64).sum(axis=0) @ Bh.reshape(N, N).astype(np.float64)
    assert np.max(np.abs(got.sum(axis=0) - colsum)) < GEMM_TOL[mode] * N * np.max(np.abs(ref))


_LARGE = {}


def _large_gemm_case(N, nrows):
    """Operands of the quoted C5 sizes (seeds of SURVEY.md section 8(d)) with the oracle's rows of C for a fixed
    sample of rows (ro_gemm_rows_interleaved: the reference's per-element k order, double), computed once per
    size on all host cores and shared by the three arithmetic modes."""
    if N not in _LARGE:
        from concurrent.futures import ThreadPoolExecutor
        _LARGE.clear()                                            # one size resident at a time (16384: ~6 GiB)
        lg = int(np.log2(N))
        Ah = np.random.default_rng(5000 + lg).uniform(-1, 1, N * N).astype(np.float32)
        Bh = np.random.default_rng(5100 + lg).uniform(-1, 1, N * N).astype(np.float32)
        rows = sorted(np.random.default_rng(9).choice(N, nrows, replace=False).tolist())
        a_rows = oracle.as_buffer(Ah.reshape(N, N)[rows])        # [nrows, N] compact
        bo = oracle.as_buffer(Bh)
        co = np.zeros(nrows * N)
        threads = min(nrows, os.cpu_count() or 1)

        def work(i):
            oracle.gemm_rows_fast(i, 1, N, N, 1.0, a_rows, N, bo, N, 0.0, co, N)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(nrows)))
        # 1^T C = (1^T A) B in double, B converted in row chunks
        a_colsum = Ah.reshape(N, N).sum(axis=0, dtype=np.float64)
        colsum = np.zeros(N)
        for r0 in range(0, N, 2048):
            colsum += a_colsum[r0:r0 + 2048] @ bo[r0 * N:(r0 + 2048) * N].reshape(-1, N)
        del bo
        _LARGE[N] = (Ah, Bh, rows, co.reshape(nrows, N), colsum)
    return _LARGE[N]


@pytest.mark.parametrize("N,mode", [(8192, rb.GEMM_TF32), (8192, rb.GEMM_TF32X3), (8192, rb.GEMM_FP32_SIMT),
                                    (16384, rb.GEMM_TF32), (16384, rb.GEMM_TF32X3), (16384, rb.GEMM_FP32_SIMT)])
def test_gemm_quoted_sizes_sampled(N, mode, cuda_service):
    """The sizes the bench numbers are quoted on (BASELINE config C5: 8192^3 and 16384^3), every arithmetic mode:
    64 (8192) / 32 (16384) full rows of C against the oracle within the mode's stated bound, and the column-sum
    identity over ALL of C (a wrong tile anywhere shows up there)."""
    Ah, Bh, rows, ref, colsum = _large_gemm_case(N, 64 if N <= 8192 else 32)
    blas = cuda_service.blas()
    blas.setGemmMode(mode)
    A, B, C = buf(Ah), buf(Bh), CudaBuffer(N * N, dtypes.float32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, N, N, N, 1.0, A, 0, N, B, 0, N, 0.0, C, 0, N)
    got = C.to_numpy().reshape(N, N)
    del A, B, C
    # float32 accumulation over K = 8192 / 16384 terms: the FFMA path's rounding error grows like sqrt(K)*eps32
    # (measured 3.9e-6 at K = 8192), so its small-K bound of 2e-6 becomes 1e-5 here -- still 10x inside the
    # reference's isclose bound (1e-4); the tensor-core modes keep their bounds at every K
    tol = 1e-5 if mode == rb.GEMM_FP32_SIMT else GEMM_TOL[mode]
    err = np.max(np.abs(got[rows].astype(np.float64) - ref)) / np.max(np.abs(ref))
    assert err < tol, (err, rb._capi.last_gemm_path())
    cs = got.sum(axis=0, dtype=np.float64)
    assert np.max(np.abs(cs - colsum)) < tol * N * np.max(np.abs(ref))
    cuda_service.blas().setGemmMode(rb.GEMM_AUTO)


def test_cross_256_config(cuda_service):
    """BASELINE config C1: MatrixOperator::cross sgemm 256x256 (alpha=1, beta=1 on a zeroed C)."""
    mo = MatrixOperator(cuda_service)
    cuda_service.blas().setGemmMode(rb.GEMM_AUTO)
    Ah = np.random.default_rng(1234).uniform(-1, 1, (256, 256)).astype(np.float32)
    Bh = np.random.default_rng(1235).uniform(-1, 1, (256, 256)).astype(np.float32)
    C = mo.cross(mo.array(Ah), mo.array(Bh))
    co = np.zeros(256 * 256)
    oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, 256, 256, 256, 1.0, oracle.as_buffer(Ah), 0, 256,
                oracle.as_buffer(Bh), 0, 256, 1.0, co, 0, 256)
    assert isclose_ref(C.to_numpy().reshape(-1), co)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", IM2COL1D, ids=[c["name"].replace(" ", "_") for c in IM2COL1D])
def test_im2col1d_provider(case, cuda_service):
    gr.run_im2col_nd_case(case["params"], cuda_service)


@pytest.mark.parametrize("case", IM2COL3D, ids=[c["name"].replace(" ", "_") for c in IM2COL3D])
def test_im2col3d_provider(case, cuda_service):
    gr.run_im2col_nd_case(case["params"], cuda_service)


@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32, dtypes.uint8, dtypes.int64])
@pytest.mark.parametrize("padding,cf,ccf,k,s,d", [(False, False, False, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
                                                 (True, False, False, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
                                                 (True, True, True, (2, 5, 3), (1, 2, 1), (2, 1, 2)),
                                                 (False, True, False, (2, 2, 4), (2, 1, 3), (1, 2, 1)),
                                                 (True, False, True, (3, 1, 3), (2, 2, 2), (2, 1, 2))])
def test_im2col3d_random_bit_exact(dt, padding, cf, ccf, k, s, d, cuda_service):
    """Volumes with im_d != im_h: with padding the reference derives padding_d from im_h (PhpMath.php:2459); the
    device and the oracle both follow it.  Forward is a bit-exact copy, col2im exact on small integers."""
    b, dd, h, w, c = 2, 7, 9, 11, 3
    rng = np.random.default_rng(4003)
    npdt = dtypes.NUMPY[dt]
    img = np.floor(rng.uniform(0, 1, b * dd * h * w * c) * 100).astype(npdt)
    outs = [dd, h, w] if padding else [(i - (kk - 1) * di - 1) // st + 1 for i, kk, st, di in zip((dd, h, w), k, s, d)]
    ncols = b * int(np.prod(outs)) * int(np.prod(k)) * c
    off_i, off_c = 3, 4
    images = CudaBuffer(off_i + img.size, dt).from_numpy(img, off_i)
    cols = CudaBuffer(off_c + ncols, dt)
    math = cuda_service.math()
    args = (b, dd, h, w, c, k[0], k[1], k[2], s[0], s[1], s[2], padding, cf, d[0], d[1], d[2], ccf)
    math.im2col3d(False, images, off_i, img.size, *args, cols, off_c, ncols)
    io = np.concatenate([np.zeros(off_i), img.astype(np.float64)])
    co = np.zeros(off_c + ncols)
    oracle.im2col3d(False, io, off_i, img.size, *args, co, off_c, ncols)
    assert np.array_equal(cols.to_numpy().astype(np.float64), co)
    if dt in (dtypes.float32, dtypes.float64, dtypes.int32, dtypes.int64):
        math.im2col3d(True, images, off_i, img.size, *args, cols, off_c, ncols)
        oracle.im2col3d(True, io, off_i, img.size, *args, co, off_c, ncols)
        assert np.array_equal(images.to_numpy().astype(np.float64), io)


@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.int32, dtypes.uint8])
@pytest.mark.parametrize("padding,cf,ccf,k,s,d", [(False, False, False, 3, 1, 1), (True, True, True, 5, 2, 2),
                                                 (True, False, True, 4, 1, 3), (False, True, False, 2, 3, 1)])
def test_im2col1d_random_bit_exact(dt, padding, cf, ccf, k, s, d, cuda_service):
    b, w, c = 3, 257, 5
    rng = np.random.default_rng(4004)
    npdt = dtypes.NUMPY[dt]
    img = np.floor(rng.uniform(0, 1, b * w * c) * 100).astype(npdt)
    ow = w if padding else (w - (k - 1) * d - 1) // s + 1
    ncols = b * ow * k * c
    images = CudaBuffer(img.size, dt).from_numpy(img)
    cols = CudaBuffer(ncols, dt)
    math = cuda_service.math()
    args = (b, w, c, k, s, padding, cf, d, ccf)
    math.im2col1d(False, images, 0, img.size, *args, cols, 0, ncols)
    io, co = img.astype(np.float64), np.zeros(ncols)
    oracle.im2col1d(False, io, 0, img.size, *args, co, 0, ncols)
    assert np.array_equal(cols.to_numpy().astype(np.float64), co)
    if dt in (dtypes.float32, dtypes.int32):
        math.im2col1d(True, images, 0, img.size, *args, cols, 0, ncols)
        oracle.im2col1d(True, io, 0, img.size, *args, co, 0, ncols)
        assert np.array_equal(images.to_numpy().astype(np.float64), io)


# im2col2d on random data and other dtypes; conv config
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [dtypes.float32, dtypes.float64, dtypes.int32, dtypes.uint8, dtypes.int64])
@pytest.mark.parametrize("padding,cf,ccf,k,s,d", [(False, False, False, (3, 3), (1, 1), (1, 1)),
                                                 (True, False, False, (3, 3), (1, 1), (1, 1)),
                                                 (True, True, True, (5, 3), (2, 1), (1, 2)),
                                                 (False, True, False, (2, 4), (1, 3), (2, 1)),
                                                 (True, False, True, (3, 3), (2, 2), (2, 2))])
def test_im2col2d_random_bit_exact(dt, padding, cf, ccf, k, s, d, cuda_service):
    b, h, w, c = 3, 19, 23, 5
    rng = np.random.default_rng(4001)
    npdt = dtypes.NUMPY[dt]
    img = np.floor(rng.uniform(0, 1, b * h * w * c) * 100).astype(npdt)      # integer-valued: col2im sums are exact
    off_i, off_c = 3, 4
    shape = oracle.im2col2d_out_shape(b, h, w, c, k[0], k[1], s[0], s[1], padding, d[0], d[1], ccf)
    ncols = int(np.prod(shape))
    images = CudaBuffer(off_i + img.size, dt).from_numpy(img, off_i)
    cols = CudaBuffer(off_c + ncols, dt)
    math = cuda_service.math()
    args = (b, h, w, c, k[0], k[1], s[0], s[1], padding, cf, d[0], d[1], ccf)
    math.im2col2d(False, images, off_i, img.size, *args, cols, off_c, ncols)
    io = np.concatenate([np.zeros(off_i), img.astype(np.float64)])
    co = np.zeros(off_c + ncols)
    oracle.im2col2d(False, io, off_i, img.size, *args, co, off_c, ncols)
    assert np.array_equal(cols.to_numpy().astype(np.float64), co)              # bit-exact copy
    if dt in (dtypes.float32, dtypes.float64, dtypes.int32, dtypes.int64):
        # col2im accumulates on top of existing image values, in the reference's order
        math.im2col2d(True, images, off_i, img.size, *args, cols, off_c, ncols)
        oracle.im2col2d(True, io, off_i, img.size, *args, co, off_c, ncols)
        assert np.array_equal(images.to_numpy().astype(np.float64), io)       # small integers: exact


@pytest.mark.parametrize("b,h,w,c,k,s,d,padding,cf,ccf", [
    (2, 5, 7, 1, (1, 1), (1, 1), (1, 1), False, False, False),        # cell == 1, one channel
    (1, 9, 9, 1, (3, 3), (1, 1), (1, 1), True, True, False),
    (3, 6, 300, 2, (3, 5), (1, 2), (1, 3), True, False, True),        # wide rows
    (1, 40, 3000, 16, (3, 3), (1, 1), (1, 1), False, False, False),   # row split into several column blocks
    (2, 12, 12, 600, (3, 3), (1, 1), (1, 1), True, False, False),     # cell = 5400 > tap table of the flat kernel
    (1, 4, 4, 20000, (2, 2), (1, 1), (1, 1), False, False, False),    # too wide for shared memory: flat kernel
])
def test_im2col2d_shapes_bit_exact(b, h, w, c, k, s, d, padding, cf, ccf, cuda_service):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 1000, b * h * w * c).astype(np.float32)
    shape = oracle.im2col2d_out_shape(b, h, w, c, k[0], k[1], s[0], s[1], padding, d[0], d[1], ccf)
    ncols = int(np.prod(shape))
    images = CudaBuffer(img.size, dtypes.float32).from_numpy(img)
    cols = CudaBuffer(ncols, dtypes.float32)
    args = (b, h, w, c, k[0], k[1], s[0], s[1], padding, cf, d[0], d[1], ccf)
    cuda_service.math().im2col2d(False, images, 0, img.size, *args, cols, 0, ncols)
    io, co = oracle.as_buffer(img), np.zeros(ncols)
    oracle.im2col2d(False, io, 0, img.size, *args, co, 0, ncols)
    assert np.array_equal(cols.to_numpy().astype(np.float64), co)


def test_im2col2d_errors(cuda_service):
    math = cuda_service.math()
    images, cols = CudaBuffer(2 * 8 * 8 * 3, dtypes.float32), CudaBuffer(2 * 6 * 6 * 27, dtypes.float32)
    with pytest.raises(rb.InvalidArgumentException, match="images buffer size is invalid"):
        math.im2col2d(False, images, 0, 100, 2, 8, 8, 3, 3, 3, 1, 1, False, False, 1, 1, False, cols, 0, 2 * 6 * 6 * 27)
    with pytest.raises(rb.InvalidArgumentException, match="Invalid shape or parameters."):
        math.im2col2d(False, images, 0, 384, 2, 8, 8, 3, 9, 9, 1, 1, False, False, 1, 1, False, cols, 0, 2 * 6 * 6 * 27)
    with pytest.raises(rb.InvalidArgumentException, match="output buffer size is invalid"):
        math.im2col2d(False, images, 0, 384, 2, 8, 8, 3, 3, 3, 1, 1, False, False, 1, 1, False, cols, 0, 100)


def test_conv_config_im2col_gemm(cuda_service):
    """BASELINE config C4 (reduced batch 4 of 64; the per-image work is identical): im2col2d + gemm
    [B*222*222, 27] x [27, 64] against the oracle; cols bit-exact, conv output in the isclose bound."""
    b, h, w, c, f = 4, 224, 224, 3, 64
    la = LinearAlgebra(cuda_service)
    cuda_service.blas().setGemmMode(rb.GEMM_AUTO)
    img = np.random.default_rng(4001).uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (np.random.default_rng(4002).standard_normal((27, f)) * 0.1).astype(np.float32)
    cols = la.im2col(la.array(img), filterSize=[3, 3])
    assert cols.shape() == [b, 222, 222, 3, 3, 3]
    win = np.lib.stride_tricks.sliding_window_view(img, (3, 3), axis=(1, 2)).transpose(0, 1, 2, 4, 5, 3)
    assert np.array_equal(cols.to_numpy(), win)
    out = la.gemm(cols.reshape([b * 222 * 222, 27]), la.array(W))
    M = b * 222 * 222
    co = np.zeros(M * f)
    rows = np.random.default_rng(1).choice(M, 512, replace=False)
    ao, bo = oracle.as_buffer(win.reshape(M, 27)), oracle.as_buffer(W)
    for r in rows:
        oracle.gemm_rows_fast(int(r), 1, f, 27, 1.0, ao, 27, bo, f, 0.0, co, f)
    assert isclose_ref(out.to_numpy().reshape(M, f)[rows], co.reshape(M, f)[rows])


def test_conv_config_full_batch(cuda_service):
    """BASELINE config C4 at its FULL size (batch 64, 224x224x3, 64 3x3 filters, valid padding), both routes:
    im2col2d (cols bit-exact against the index identity) + gemm, and the fused conv2d kernel.  512 sampled output
    rows against the oracle's gemm loop within the reference isclose bound (and the 3xTF32 bound), plus the
    column sums of the WHOLE output: 1^T out = (1^T cols) W."""
    b, h, w, c, f = 64, 224, 224, 3, 64
    la = LinearAlgebra(cuda_service)
    cuda_service.blas().setGemmMode(rb.GEMM_AUTO)
    img = np.random.default_rng(4001).uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (np.random.default_rng(4002).standard_normal((27, f)) * 0.1).astype(np.float32)
    images = la.array(img)
    cols = la.im2col(images, filterSize=[3, 3])
    assert cols.shape() == [b, 222, 222, 3, 3, 3]
    win = np.lib.stride_tricks.sliding_window_view(img, (3, 3), axis=(1, 2)).transpose(0, 1, 2, 4, 5, 3)
    got_cols = cols.to_numpy()
    for b0 in range(0, b, 8):                                   # chunked: the view materialises on comparison
        assert np.array_equal(got_cols[b0:b0 + 8], win[b0:b0 + 8]), b0
    M = b * 222 * 222
    out2 = la.gemm(cols.reshape([M, 27]), la.array(W))
    path2 = rb._capi.last_gemm_path()
    out1 = la.conv2d(images, la.array(W.reshape(3, 3, c, f)), fused=True)
    path1 = rb._capi.last_gemm_path()
    assert path1.endswith("_tmem_conv"), path1
    rows = np.sort(np.random.default_rng(1).choice(M, 512, replace=False))
    a_rows = oracle.as_buffer(got_cols.reshape(M, 27)[rows])
    ref = np.zeros(512 * f)
    oracle.gemm_rows_fast(0, 512, f, 27, 1.0, a_rows, 27, oracle.as_buffer(W), f, 0.0, ref, f)
    ref = ref.reshape(512, f)
    col_total = got_cols.reshape(M, 27).sum(axis=0, dtype=np.float64) @ W.astype(np.float64)
    for name, out, path in (("im2col+gemm", out2, path2), ("fused", out1, path1)):
        got = out.to_numpy().reshape(M, f)
        assert isclose_ref(got[rows], ref), (name, path)
        assert np.max(np.abs(got[rows] - ref)) / np.max(np.abs(ref)) < GEMM_TOL[rb.GEMM_TF32X3], (name, path)
        cs = got.sum(axis=0, dtype=np.float64)
        assert np.max(np.abs(cs - col_total)) < GEMM_TOL[rb.GEMM_TF32X3] * M * np.max(np.abs(ref)), (name, path)


CONV_CASES = [  # (b, h, w, c, kh, kw, filters, strides, padding, dilation, bias)
    (2, 16, 16, 3, 3, 3, 64, (1, 1), False, (1, 1), False),
    (2, 16, 16, 3, 3, 3, 64, (1, 1), True, (1, 1), True),
    (3, 17, 23, 3, 3, 3, 64, (2, 1), True, (1, 2), True),
    (1, 30, 31, 8, 3, 3, 32, (1, 1), False, (1, 1), False),       # K = 72: three k-blocks
    (2, 12, 12, 5, 2, 4, 100, (1, 2), True, (2, 1), True),         # N = 100 -> padded to 128
    (4, 64, 64, 3, 3, 3, 64, (1, 1), False, (1, 1), True),
    (1, 9, 9, 16, 3, 3, 16, (1, 1), False, (1, 1), False),         # K = 144 > 128: two-call path
    # K <= 32: the TMEM-operand kernel (conv_tcgen05_rows.cu)
    (2, 40, 300, 3, 3, 3, 64, (1, 1), True, (1, 1), True),         # three column blocks, same padding (slab shifted left)
    (2, 33, 140, 1, 5, 5, 32, (1, 1), False, (1, 1), False),       # K = 25, one channel, 32 filters
    (1, 64, 200, 1, 3, 3, 128, (2, 2), True, (1, 1), True),        # stride 2, 128 filters (two 64-column epilogue passes)
    (2, 20, 36, 2, 2, 4, 96, (1, 1), False, (1, 2), True),         # two channels, dilated taps, 96 filters, several rows per tile
    (1, 48, 52, 3, 3, 3, 36, (1, 1), False, (1, 1), False),        # 36 filters: the TMA store clips the filter dimension
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
@pytest.mark.parametrize("mode", [rb.GEMM_TF32X3, rb.GEMM_TF32, rb.GEMM_FP32_SIMT])
def test_conv2d_forward_vs_oracle(case, mode, cuda_service):
    """Fused conv2d forward (implicit GEMM on tcgen05) == oracle im2col2d + oracle gemm (+ bias)."""
    b, h, w, c, kh, kw, f, st, pad, dil, use_bias = case
    rng = np.random.default_rng(4001)
    img = rng.uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((kh, kw, c, f)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(f).astype(np.float32) if use_bias else None
    la = LinearAlgebra(cuda_service)
    cuda_service.blas().setGemmMode(mode)
    rb._capi.check(rb._capi.load().rb200_set_default_gemm_mode(mode))
    try:
        out = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st,
                        padding=pad, dilation_rate=dil, fused=True)
        path = rb._capi.last_gemm_path()
        out2 = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st,
                         padding=pad, dilation_rate=dil, fused=False)           # im2col -> gemm -> add
    finally:
        rb._capi.check(rb._capi.load().rb200_set_default_gemm_mode(rb.GEMM_TF32X3))
        cuda_service.blas().setGemmMode(rb.GEMM_AUTO)
    K = kh * kw * c
    if mode != rb.GEMM_FP32_SIMT and K <= 128 and f <= 64:
        assert path.endswith("_conv"), path          # wider filter banks may not fit the shared-memory budget
    if (mode != rb.GEMM_FP32_SIMT and K <= 32 and 32 <= f <= 128 and f % 4 == 0 and (w * c) % 4 == 0
            and np.gcd(st[1] * c, 32) <= 2):
        assert path.endswith("_tmem_conv"), path     # short reductions: A operand in tensor memory
    shape = oracle.im2col2d_out_shape(b, h, w, c, kh, kw, st[0], st[1], pad, dil[0], dil[1], False)
    M = shape[0] * shape[1] * shape[2]
    cols = np.zeros(M * K)
    oracle.im2col2d(False, oracle.as_buffer(img), 0, img.size, b, h, w, c, kh, kw, st[0], st[1], pad, False,
                    dil[0], dil[1], False, cols, 0, cols.size)
    ref = np.zeros(M * f)
    oracle.gemm_rows_fast(0, M, f, K, 1.0, cols, K, oracle.as_buffer(W), f, 0.0, ref, f)
    if bias is not None:
        oracle.add(False, M, f, 1.0, oracle.as_buffer(bias), 0, 1, ref, 0, f)
    assert out.shape() == [shape[0], shape[1], shape[2], f]
    got = out.to_numpy().reshape(-1).astype(np.float64)
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err < GEMM_TOL[mode], (err, path)
    got2 = out2.to_numpy().reshape(-1).astype(np.float64)
    assert out2.shape() == out.shape()
    assert np.max(np.abs(got2 - ref)) / np.max(np.abs(ref)) < GEMM_TOL[mode]


TMA_CONV_CASES = [  # (b, h, w, c, kh, kw, filters, strides, padding, dilation, bias)
    (2, 20, 20, 32, 3, 3, 64, (1, 1), False, (1, 1), False),       # 18 x 18 out: 7 rows per tile, ragged last block
    (2, 20, 20, 32, 3, 3, 64, (1, 1), True, (1, 1), True),         # same padding = TMA zero fill on all four sides
    (1, 56, 56, 64, 3, 3, 64, (1, 1), True, (1, 1), True),         # the 64-channel 3x3 layer: 2 rows per tile
    (2, 30, 150, 32, 3, 3, 128, (1, 1), False, (1, 1), False),     # wide rows: balanced segments of 74 cells
    (1, 33, 41, 64, 3, 3, 96, (2, 2), True, (1, 1), True),         # stride 2: TMA element strides
    (1, 31, 29, 32, 3, 3, 40, (2, 1), False, (1, 2), True),        # mixed strides / dilation, 40 filters (column tail)
    (1, 24, 24, 96, 1, 1, 256, (1, 1), False, (1, 1), False),      # 1x1 conv, 256 filters (BN = 256 in TF32 mode)
    (1, 16, 16, 64, 5, 5, 32, (1, 1), True, (2, 2), True),         # 5x5 dilated
    (1, 12, 12, 256, 3, 3, 64, (1, 1), True, (1, 1), True),        # K = 2304: two accumulating launches in 3xTF32
    (3, 9, 9, 32, 3, 3, 300, (1, 1), False, (1, 1), True),         # 300 filters: several filter blocks
]


@pytest.mark.parametrize("case", TMA_CONV_CASES, ids=[str(c[:7]) for c in TMA_CONV_CASES])
@pytest.mark.parametrize("mode", [rb.GEMM_TF32X3, rb.GEMM_TF32])
def test_conv2d_forward_tma_vs_oracle(case, mode, cuda_service):
    """Hidden-layer shapes (channels % 32 == 0) on the TMA-fed implicit GEMM == oracle im2col2d + oracle gemm (+ bias),
    in the mode's bound; the default route (fused=None) must pick the fused kernel for them."""
    b, h, w, c, kh, kw, f, st, pad, dil, use_bias = case
    rng = np.random.default_rng(h * 100 + w + c)
    img = rng.uniform(-1, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((kh, kw, c, f)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(f).astype(np.float32) if use_bias else None
    la = LinearAlgebra(cuda_service)
    cuda_service.blas().setGemmMode(mode)
    rb._capi.check(rb._capi.load().rb200_set_default_gemm_mode(mode))
    try:
        out = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st, padding=pad,
                        dilation_rate=dil)
        path = rb._capi.last_gemm_path()
    finally:
        rb._capi.check(rb._capi.load().rb200_set_default_gemm_mode(rb.GEMM_TF32X3))
        cuda_service.blas().setGemmMode(rb.GEMM_AUTO)
    assert path == ("tcgen05_tf32_tma_conv" if mode == rb.GEMM_TF32 else "tcgen05_tf32x3_tma_conv"), path
    K = kh * kw * c
    shape = oracle.im2col2d_out_shape(b, h, w, c, kh, kw, st[0], st[1], pad, dil[0], dil[1], False)
    M = shape[0] * shape[1] * shape[2]
    cols = np.zeros(M * K)
    oracle.im2col2d(False, oracle.as_buffer(img), 0, img.size, b, h, w, c, kh, kw, st[0], st[1], pad, False,
                    dil[0], dil[1], False, cols, 0, cols.size)
    ref = np.zeros(M * f)
    oracle.gemm_rows_fast(0, M, f, K, 1.0, cols, K, oracle.as_buffer(W), f, 0.0, ref, f)
    if bias is not None:
        oracle.add(False, M, f, 1.0, oracle.as_buffer(bias), 0, 1, ref, 0, f)
    assert out.shape() == [shape[0], shape[1], shape[2], f]
    got = out.to_numpy().reshape(-1).astype(np.float64)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < GEMM_TOL[mode], path


def test_conv2d_forward_tma_nonfinite_containment(cuda_service):
    """NaN / Inf pixels reach only the outputs whose window holds them: padding comes from TMA's zero FILL of the
    operand tile, and 0 * Inf never happens because out-of-image taps are zeros on the A side only where W is finite."""
    b, h, w, c, f = 1, 14, 14, 32, 64
    rng = np.random.default_rng(11)
    img = rng.uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((3, 3, c, f)) * 0.1).astype(np.float32)
    la = LinearAlgebra(cuda_service)
    clean = la.conv2d(la.array(img), la.array(W), padding=True).to_numpy()
    assert rb._capi.last_gemm_path().endswith("_tma_conv")
    bad = img.copy()
    bad[0, 5, 7, 1] = np.nan
    bad[0, 13, 0, 30] = np.inf
    out = la.conv2d(la.array(bad), la.array(W), padding=True).to_numpy()
    touched = np.zeros((h, w), bool)
    for (y, x) in ((5, 7), (13, 0)):
        touched[max(0, y - 1):y + 2, max(0, x - 1):x + 2] = True
    assert not np.isfinite(out[0][touched]).any()
    assert np.isfinite(out[0][~touched]).all()
    assert np.array_equal(out[0][~touched], clean[0][~touched])


def test_conv2d_forward_nonfinite_containment(cuda_service):
    """A NaN / Inf pixel may only reach the outputs whose receptive field holds it (the reference multiplies
    padding zeros and absent taps by nothing at all): checks the zero tails of the K <= 32 kernel's A rows."""
    b, h, w, c, f = 1, 20, 24, 3, 64
    rng = np.random.default_rng(7)
    img = rng.uniform(0, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((3, 3, c, f)) * 0.1).astype(np.float32)
    la = LinearAlgebra(cuda_service)
    clean = la.conv2d(la.array(img), la.array(W), padding=True, fused=True).to_numpy()
    assert rb._capi.last_gemm_path().endswith("_tmem_conv")
    bad = img.copy()
    bad[0, 5, 7, 1] = np.nan
    bad[0, 19, 0, 2] = np.inf            # corner pixel: its neighbours' windows reach into the padding
    out = la.conv2d(la.array(bad), la.array(W), padding=True, fused=True).to_numpy()
    touched = np.zeros((h, w), bool)
    for (y, x) in ((5, 7), (19, 0)):
        touched[max(0, y - 1):y + 2, max(0, x - 1):x + 2] = True
    assert not np.isfinite(out[0][touched]).any()
    assert np.isfinite(out[0][~touched]).all()
    assert np.array_equal(out[0][~touched], clean[0][~touched])


# ---------------------------------------------------------------------------------------------------
# buffer contract (PhpBufferTest.php)
# ---------------------------------------------------------------------------------------------------
def test_buffer_complex_dtypes(cuda_service):
    """complex64 / complex128 storage (PhpBufferTest.php:103-129, testDumpAndLoadComplex :173-190): 8 / 16 bytes per
    element, elements with real / imag, dump() interleaves real and imag, copy / swap / fill move them bit for bit;
    arithmetic entry points refuse them."""
    class Obj:
        real, imag = 3.5, 4.5
    for dt, vs, npdt in ((dtypes.complex64, 8, np.complex64), (dtypes.complex128, 16, np.complex128)):
        b = CudaBuffer(3, dt)
        assert b.dtype() == dt and len(b) == 3 and b.value_size() == vs
        assert b[0] == 0
        b[1] = complex(1.5, 2.5)
        b[2] = Obj()
        assert (b[1].real, b[1].imag) == (1.5, 2.5) and (b[2].real, b[2].imag) == (3.5, 4.5)
        with pytest.raises(rb.InvalidArgumentException, match="Cannot convert to complex number"):
            b[0] = 1.0
        dump = b.dump()
        assert np.array_equal(np.frombuffer(dump, np.float32 if vs == 8 else np.float64), [0, 0, 1.5, 2.5, 3.5, 4.5])
        b2 = CudaBuffer(3, dt)
        b2.load(dump)
        assert [b2[i] for i in range(3)] == [0j, 1.5 + 2.5j, 3.5 + 4.5j]
        # device-side data movement
        blas = cuda_service.blas()
        n = 1000
        x = (np.arange(n) + 1j * np.arange(n)[::-1]).astype(npdt)
        X, Y = CudaBuffer(n, dt).from_numpy(x), CudaBuffer(n, dt)
        blas.copy(n, X, 0, 1, Y, 0, 1)
        assert np.array_equal(Y.to_numpy(), x)
        Y.fill(complex(2, -3), 10, 5)
        want = x.copy()
        want[5:15] = 2 - 3j
        assert np.array_equal(Y.to_numpy(), want)
        blas.swap(n, X, 0, 1, Y, 0, 1)
        assert np.array_equal(X.to_numpy(), want) and np.array_equal(Y.to_numpy(), x)
        with pytest.raises(Exception):
            blas.scal(n, 2.0, X, 0, 1)
        a = NDArray(np.asarray([[1 + 2j, 3 - 1j], [0.5j, 4]]), dtype=dt, service=cuda_service)
        assert a.toArray() == [[1 + 2j, 3 - 1j], [0.5j, 4 + 0j]]


def test_buffer_contract():
    b = CudaBuffer(5, dtypes.float32)
    assert len(b) == 5 and b.dtype() == dtypes.float32 and b.value_size() == 4
    assert [b[i] for i in range(5)] == [0.0] * 5
    b[2] = 1.5
    assert b[2] == 1.5
    with pytest.raises(rb.RuntimeException):
        b[5]
    with pytest.raises(rb.RuntimeException):
        b[-1] = 1
    with pytest.raises(rb.InvalidArgumentException, match="Invalid data type"):
        CudaBuffer(3, 99)
    raw = b.dump()
    assert raw == np.array([0, 0, 1.5, 0, 0], "<f4").tobytes()
    c = CudaBuffer(5, dtypes.float32)
    c.load(raw)
    assert c[2] == 1.5 and c.to_numpy().tolist() == [0, 0, 1.5, 0, 0]
    for dt, sz in ((dtypes.int8, 1), (dtypes.uint16, 2), (dtypes.int64, 8), (dtypes.float64, 8), (dtypes.bool_, 1)):
        assert CudaBuffer(2, dt).value_size() == sz


def test_buffer_host_device_coherence(cuda_service):
    """Host element writes are flushed before a kernel, kernel results fetched before a host read."""
    la = LinearAlgebra(cuda_service)
    A = la.array([[1, 2], [3, 4]])
    X = la.array([10, 20])
    A.buffer()[0] = 5.0                       # host write after upload
    la.add(X, A)
    assert A.toArray() == [[15, 22], [13, 24]]
    A.buffer()[3] = 0.0                       # host write after kernel
    la.add(X, A)
    assert A.toArray() == [[25, 42], [23, 20]]
    assert A[1][1] == 20.0 and A[0].toArray() == [25, 42]


_FUZZ_PATHS = []


def _conv_fuzz_cases(n=48, seed=20260):
    """Random K <= 32 conv shapes for the TMEM-operand kernel: channels 1..6, filters up to 5x5, strides / dilations 1..3,
    both paddings, ragged widths around the 128-cell tile, 32..128 filters."""
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < n:
        c = int(rng.integers(1, 7))
        kh, kw = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        if kh * kw * c > 32:
            continue
        st = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        dil = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        pad = bool(rng.integers(0, 2))
        h = int(rng.integers((kh - 1) * dil[0] + 1, 40))
        w = int(rng.choice([rng.integers((kw - 1) * dil[1] + 1, 64), rng.integers(120, 140), rng.integers(250, 300)]))
        if w < (kw - 1) * dil[1] + 1:
            continue
        f = int(rng.choice([32, 36, 48, 64, 96, 100, 128]))
        b = int(rng.integers(1, 4))
        if len(cases) % 3:  # two in three inside the TMEM kernel's envelope (odd column stride, W*C and filters multiples of 4)
            st = (st[0], int(rng.choice([1, 3])))
            c = c if c != 4 else 3
            if kh * kw * c > 32:
                continue
            w += (-w) % 4
            f = int(rng.choice([32, 36, 48, 64, 96, 100, 128]))
        cases.append((b, h, w, c, kh, kw, f, st, pad, dil, bool(rng.integers(0, 2))))
    return cases


@pytest.mark.parametrize("case", _conv_fuzz_cases(), ids=lambda c: "x".join(str(v) for v in c[:7]))
def test_conv2d_forward_fuzz_vs_oracle(case, cuda_service):
    b, h, w, c, kh, kw, f, st, pad, dil, use_bias = case
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.uniform(-1, 1, (b, h, w, c)).astype(np.float32)
    W = (rng.standard_normal((kh, kw, c, f)) * 0.2).astype(np.float32)
    bias = rng.standard_normal(f).astype(np.float32) if use_bias else None
    la = LinearAlgebra(cuda_service)
    out = la.conv2d(la.array(img), la.array(W), None if bias is None else la.array(bias), strides=st, padding=pad,
                    dilation_rate=dil, fused=True)
    path = rb._capi.last_gemm_path()
    _FUZZ_PATHS.append(path)
    K = kh * kw * c
    shape = oracle.im2col2d_out_shape(b, h, w, c, kh, kw, st[0], st[1], pad, dil[0], dil[1], False)
    M = shape[0] * shape[1] * shape[2]
    cols = np.zeros(M * K)
    oracle.im2col2d(False, oracle.as_buffer(img), 0, img.size, b, h, w, c, kh, kw, st[0], st[1], pad, False,
                    dil[0], dil[1], False, cols, 0, cols.size)
    ref = np.zeros(M * f)
    oracle.gemm_rows_fast(0, M, f, K, 1.0, cols, K, oracle.as_buffer(W), f, 0.0, ref, f)
    if bias is not None:
        oracle.add(False, M, f, 1.0, oracle.as_buffer(bias), 0, 1, ref, 0, f)
    assert out.shape() == [shape[0], shape[1], shape[2], f]
    got = out.to_numpy().reshape(-1).astype(np.float64)
    assert np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30) < GEMM_TOL[rb.GEMM_TF32X3], path


def test_conv2d_forward_fuzz_covers_the_tmem_kernel():
    # the kernel's envelope also depends on slab / staging sizes, so the share is asserted over the whole sample
    if len(_FUZZ_PATHS) < 48:
        pytest.skip("fuzz cases were deselected")
    assert sum(p.endswith("_tmem_conv") for p in _FUZZ_PATHS) >= 20, _FUZZ_PATHS
