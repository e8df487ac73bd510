"""Pins the CPU oracle (oracle/rindow_oracle.c) against every known-answer vector the reference's
own PHPUnit suite holds for the hot path.  CPU only."""
import json
import os

import numpy as np
import pytest

import golden_runner as gr
import oracle
from oracle_service import OracleService

KA = gr.load_known_answers()
with open(os.path.join(gr.HERE, "golden", "im2col2d_cases.json")) as f:
    IM2COL = json.load(f)["cases"]
with open(os.path.join(gr.HERE, "golden", "im2col1d_cases.json")) as f:
    IM2COL1D = json.load(f)["cases"]
with open(os.path.join(gr.HERE, "golden", "im2col3d_cases.json")) as f:
    IM2COL3D = json.load(f)["cases"]
with open(os.path.join(gr.HERE, "golden", "gatherb_cases.json")) as f:
    GATHERB = json.load(f)["cases"]


with open(os.path.join(gr.HERE, "golden", "gathernd_cases.json")) as f:
    GATHERND = json.load(f)["cases"]


@pytest.mark.parametrize("case", GATHERND, ids=[c["op"] + "_" + c["ref"].split(":")[1] for c in GATHERND])
def test_gathernd_reference_cases(case, oracle_service):
    assert len(GATHERND) == 18
    gr.run_gatherb_case(case, oracle_service)


with open(os.path.join(gr.HERE, "golden", "imagecopy_cases.json")) as f:
    IMAGECOPY = json.load(f)["cases"]


@pytest.mark.parametrize("case", IMAGECOPY, ids=["imagecopy_" + c["ref"].split(":")[1] for c in IMAGECOPY])
def test_imagecopy_reference_cases(case, oracle_service):
    assert len(IMAGECOPY) == 26
    gr.run_imagecopy_case(case, oracle_service)


@pytest.mark.parametrize("case", GATHERB, ids=[c["op"] + "_" + c["ref"].split(":")[1] for c in GATHERB])
def test_gatherb_reference_cases(case, oracle_service):
    assert len(GATHERB) == 26
    gr.run_gatherb_case(case, oracle_service)


@pytest.mark.parametrize("case", KA["cases"], ids=[c["id"] for c in KA["cases"]])
def test_known_answer(case, oracle_service):
    gr.run_case(case, oracle_service)


@pytest.mark.parametrize("case", IM2COL, ids=[c["name"].replace(" ", "_") for c in IM2COL])
def test_im2col2d_provider(case, oracle_service):
    assert len(IM2COL) == 28
    gr.run_im2col_case(case["params"], oracle_service)


@pytest.mark.parametrize("case", IM2COL1D, ids=[c["name"].replace(" ", "_") for c in IM2COL1D])
def test_im2col1d_provider(case, oracle_service):
    assert len(IM2COL1D) == 16
    gr.run_im2col_nd_case(case["params"], oracle_service)


@pytest.mark.parametrize("case", IM2COL3D, ids=[c["name"].replace(" ", "_") for c in IM2COL3D])
def test_im2col3d_provider(case, oracle_service):
    assert len(IM2COL3D) == 40
    gr.run_im2col_nd_case(case["params"], oracle_service)


def test_reducemax_nan_rules():
    # PHP rule (PhpMath.php:95-112): a NaN followed by ordinary values is forgotten ...
    a = oracle.as_buffer([1.0, np.nan, 0.5])
    b = np.zeros(1)
    oracle.reduce_max(1, 3, 1, a, 0, b, 0, sticky=False)
    assert b[0] == 0.5
    # ... the accelerated-path rule keeps it (LinearAlgebraTest.php:7617-7643)
    oracle.reduce_max(1, 3, 1, a, 0, b, 0, sticky=True)
    assert np.isnan(b[0])
    # NaN last: both rules agree (LinearAlgebraTest.php:7570-7588)
    a = oracle.as_buffer([np.inf, np.nan])
    for sticky in (False, True):
        oracle.reduce_max(1, 2, 1, a, 0, b, 0, sticky=sticky)
        assert np.isnan(b[0])


def test_reducemax_large_nan_table():
    """testReduceMaxLarge's NaN table (LinearAlgebraTest.php:7621-7643) on the sticky rule."""
    case = next(c for c in KA["large"]["cases"] if c["id"] == "reducemax_large_nan_sticky")
    for cols in case["cols_list"][:3]:
        x = np.zeros((6, cols))
        for r, head in enumerate(case["heads"]):
            x[r, :2] = [gr._num(v) for v in head]
        a, b = oracle.as_buffer(x), np.zeros(6)
        oracle.reduce_max(6, cols, 1, a, 0, b, 0, sticky=True)
        gr._assert_equal(case["expect"], b)


def test_argmax_rules():
    b = np.zeros(1, np.int64)
    for row, want in (([1.0, 3.0, 3.0], 1), ([np.nan, 5.0], 0), ([1.0, np.nan, 0.0], 0),
                      ([-np.inf, np.nan, -np.inf], 0), ([2.0, 2.0], 0)):
        oracle.reduce_argmax(1, len(row), 1, oracle.as_buffer(row), 0, b, 0)
        assert b[0] == want, (row, b[0])


def test_gemm_against_numpy_float64():
    rng = np.random.default_rng(7)
    for (m, n, k, tA, tB) in [(5, 7, 3, False, False), (4, 6, 9, True, False), (8, 3, 5, False, True),
                              (6, 6, 6, True, True)]:
        A = rng.standard_normal((k, m) if tA else (m, k)).astype(np.float32)
        B = rng.standard_normal((n, k) if tB else (k, n)).astype(np.float32)
        C0 = rng.standard_normal((m, n)).astype(np.float32)
        a, b, c = oracle.as_buffer(A), oracle.as_buffer(B), oracle.as_buffer(C0)
        oracle.gemm(oracle.RowMajor, oracle.Trans if tA else oracle.NoTrans, oracle.Trans if tB else oracle.NoTrans,
                    m, n, k, 1.5, a, 0, A.shape[1], b, 0, B.shape[1], -0.5, c, 0, n)
        opA = A.T.astype(np.float64) if tA else A.astype(np.float64)
        opB = B.T.astype(np.float64) if tB else B.astype(np.float64)
        want = 1.5 * (opA @ opB) - 0.5 * C0
        assert np.allclose(c.reshape(m, n), want, rtol=1e-12, atol=1e-12)


def test_gemm_beta_zero_does_not_read_c():
    a, b = oracle.as_buffer(np.eye(2)), oracle.as_buffer([[1, 2], [3, 4]])
    c = np.full(4, np.nan)
    oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, 2, 2, 2, 1.0, a, 0, 2, b, 0, 2, 0.0, c, 0, 2)
    assert c.tolist() == [1, 2, 3, 4]                       # PhpBlas.php:940-944


def test_gemm_overflow_errors():
    """PhpBlasTest.php:1988-2155: 'Matrix specification too large for buffer{A,B,C}.'"""
    A, B, C = oracle.as_buffer([1, 2, 3, 4, 5]), oracle.as_buffer(np.eye(3)), np.zeros(6)
    with pytest.raises(oracle.OracleArgumentError):
        oracle.gemm(oracle.RowMajor, oracle.Trans, oracle.NoTrans, 2, 3, 3, 1.0, A, 0, 2, B, 0, 3, 0.0, C, 0, 3)
    A = oracle.as_buffer([1, 2, 3, 4, 5, 6])
    with pytest.raises(oracle.OracleArgumentError):
        oracle.gemm(oracle.RowMajor, oracle.Trans, oracle.NoTrans, 2, 3, 3, 1.0, A, 0, 2, B, 0, 3, 0.0, np.zeros(5), 0, 3)
    with pytest.raises(oracle.OracleArgumentError):
        oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, 0, 3, 3, 1.0, A, 0, 2, B, 0, 3, 0.0, C, 0, 3)


def test_softmax_matches_numpy():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((7, 33)).astype(np.float32)
    a = oracle.as_buffer(x)
    oracle.softmax(7, 33, a, 0, 33)
    x64 = x.astype(np.float64)
    e = np.exp(x64 - x64.max(axis=1, keepdims=True))
    assert np.allclose(a.reshape(7, 33), e / e.sum(axis=1, keepdims=True), rtol=1e-14)


def test_gemm_rows_variants_bit_identical():
    """gemm_rows (reference loop order) == gemm_rows_fast (interleaved) == gemm, bit for bit."""
    rng = np.random.default_rng(12)
    m, n, k = 9, 37, 53
    A = oracle.as_buffer(rng.standard_normal((m, k)).astype(np.float32))
    B = oracle.as_buffer(rng.standard_normal((k, n)).astype(np.float32))
    C0 = rng.standard_normal(m * n)
    c1, c2, c3 = C0.copy(), C0.copy(), C0.copy()
    oracle.gemm(oracle.RowMajor, oracle.NoTrans, oracle.NoTrans, m, n, k, 0.7, A, 0, k, B, 0, n, 1.3, c1, 0, n)
    oracle.gemm_rows(0, m, n, k, 0.7, A, k, B, n, 1.3, c2, n)
    oracle.gemm_rows_fast(0, m, n, k, 0.7, A, k, B, n, 1.3, c3, n)
    assert np.array_equal(c1, c2) and np.array_equal(c1, c3)


def test_topk_known_answers(oracle_service):
    """LinearAlgebraTest.php:8027-8115 (testTopKNormal10) through the oracle service"""
    from rindow_math_matrix_b200 import LinearAlgebra
    la = LinearAlgebra(oracle_service)
    x = la.array(np.arange(10, dtype=np.float32))
    values, indices = la.topK(x, k=3)
    assert values.toArray() == [9, 8, 7] and indices.toArray() == [9, 8, 7]
    values, indices = la.topK(x)
    assert values.toArray() == [9] and indices.toArray() == [9]
    x = la.array(np.arange(30, dtype=np.float32).reshape(3, 10))
    values, indices = la.topK(x, k=3)
    assert values.toArray() == [[9, 8, 7], [19, 18, 17], [29, 28, 27]] and indices.toArray() == [[9, 8, 7]] * 3
