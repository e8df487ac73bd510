"""Host-operand GEMM (rb200_sgemm_streamed): same result as the device-resident call, and the buffer
coherence flags end up right.  pytest -m gpu"""
import numpy as np
import pytest

import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, MatrixOperator, dtypes

pytestmark = pytest.mark.gpu
F32 = dtypes.float32


def _host_filled(a):
    """A CudaBuffer whose current copy is in its pinned mirror (what PHP code that fills $buf[$i] leaves)."""
    b = CudaBuffer(a.size, F32)
    b.host()[:] = a.reshape(-1)
    b.mark_host_written()
    assert b.host_pending()
    return b


def _device(a):
    return CudaBuffer(a.size, F32).from_numpy(a.reshape(-1))


CASES = [(2048, 256, 128, False, False), (2304, 384, 200, False, False), (2148, 136, 64, True, False),
         (4096, 512, 256, False, True), (3000, 260, 132, True, True), (8192, 1024, 512, False, False)]


@pytest.mark.parametrize("m,n,k,tA,tB", CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize("mode", [rb.GEMM_TF32, rb.GEMM_TF32X3, rb.GEMM_FP32_SIMT])
@pytest.mark.parametrize("beta", [0.0, 0.5])
def test_streamed_equals_device_resident(m, n, k, tA, tB, mode, beta, cuda_service):
    if mode == rb.GEMM_FP32_SIMT and m * n * k > 2e9:
        pytest.skip("FFMA path: keep the case small")
    rng = np.random.default_rng(m + n + k)
    blas = cuda_service.blas()
    blas.setGemmMode(mode)
    try:
        a = rng.uniform(-1, 1, (k, m) if tA else (m, k)).astype(np.float32)
        b = rng.uniform(-1, 1, (n, k) if tB else (k, n)).astype(np.float32)
        c0 = rng.uniform(-1, 1, (m, n)).astype(np.float32)
        ta, tb = (BLAS.Trans if tA else BLAS.NoTrans), (BLAS.Trans if tB else BLAS.NoTrans)
        lda, ldb = a.shape[1], b.shape[1]
        # reference: everything device resident, no mirrors -> plain rb200_sgemm
        A, B, C = _device(a), _device(b), _device(c0)
        blas.gemm(BLAS.RowMajor, ta, tb, m, n, k, 1.25, A, 0, lda, B, 0, ldb, beta, C, 0, n)
        assert not C.has_mirror()
        want = C.to_numpy()
        # streamed: operands written from the host, C read from the host
        A, B, C = _host_filled(a), _host_filled(b), _host_filled(c0)
        blas.gemm(BLAS.RowMajor, ta, tb, m, n, k, 1.25, A, 0, lda, B, 0, ldb, beta, C, 0, n)
        assert not A.host_pending() and not B.host_pending() and not C.host_pending()
        assert C.mirror_current()                                  # result already in the pinned mirror
        got_host = np.array(C.host())
        assert np.array_equal(got_host, want)
        assert np.array_equal(C.to_numpy(), want)                   # and on the device
        assert np.array_equal(A.to_numpy(), a.reshape(-1)) and np.array_equal(B.to_numpy(), b.reshape(-1))
    finally:
        blas.setGemmMode(rb.GEMM_AUTO)


def test_streamed_partial_roles_and_policy(cuda_service):
    rng = np.random.default_rng(9)
    blas = cuda_service.blas()
    m, n, k = 2560, 384, 256
    a = rng.uniform(-1, 1, (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    A, B, C = _device(a), _device(b), CudaBuffer(m * n, F32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, 0, n)
    want = C.to_numpy()
    # only A comes from the host; C stays on the device (no mirror, policy auto)
    A2, C2 = _host_filled(a), CudaBuffer(m * n, F32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A2, 0, k, B, 0, n, 0.0, C2, 0, n)
    assert not A2.host_pending() and not C2.has_mirror()
    assert np.array_equal(C2.to_numpy(), want)
    # policy "always": C comes back although nobody indexed it before
    blas.setHostResultPrefetch(True)
    try:
        C3 = CudaBuffer(m * n, F32)
        blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C3, 0, n)
        assert C3.mirror_current() and np.array_equal(np.array(C3.host()), want)
    finally:
        blas.setHostResultPrefetch(None)
    # a view that is not the whole buffer (offset, ld > cols) keeps the plain path and stays correct
    pad = np.zeros(m * (k + 4) + 8, np.float32)
    pad[8:].reshape(m, k + 4)[:, :k] = a
    A4 = _host_filled(pad)
    C4 = CudaBuffer(m * n, F32)
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, m, n, k, 1.0, A4, 8, k + 4, B, 0, n, 0.0, C4, 0, n)
    assert np.array_equal(C4.to_numpy(), want)


def test_cross_from_host_arrays_reads_back(cuda_service):
    """MatrixOperator::cross on arrays built from host data, result read with toArray-style access."""
    rng = np.random.default_rng(3)
    mo = MatrixOperator(cuda_service)
    a = rng.integers(-4, 5, (2048, 96)).astype(np.float32)
    b = rng.integers(-4, 5, (96, 160)).astype(np.float32)
    c = mo.cross(mo.array(a), mo.array(b))
    assert np.array_equal(c.to_numpy(), a @ b)                      # integer valued: exact in every mode
