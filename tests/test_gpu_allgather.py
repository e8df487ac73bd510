"""rb200_sgemm_allgather: the row-sharded sgemm whose epilogue also stores its C stripe into the peers' full C.
On one GPU the "peers" are further local allocations (a peer pointer is just a device pointer to the kernel), which
exercises the same epilogue stores, the K-chunked 3xTF32 path (peers written by the last chunk only) and the
copy fallback of the other kernel families; tools/mg_check.py runs the real multi-process CUDA-IPC form under
torchrun (--gpus 2).  pytest -m gpu"""
import numpy as np
import pytest

import oracle
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import BLAS, CudaBuffer, _capi, dtypes

pytestmark = pytest.mark.gpu

CASES = [
    # total_rows, stripe start, stripe rows, n, k, mode, expected kernel family
    (1024, 256, 512, 384, 256, rb.GEMM_TF32, "tcgen05_tf32_2cta"),
    (1024, 0, 512, 520, 4100, rb.GEMM_TF32X3, "tcgen05_tf32x3_2cta"),      # K chunks of 2048 + N tail
    (600, 300, 300, 256, 128, rb.GEMM_TF32X3, "tcgen05_tf32x3_2cta"),      # M tail inside the pair tile
    (512, 128, 128, 256, 64, rb.GEMM_TF32, "tcgen05_tf32"),                # 1-CTA kernel: copy fallback
    (64, 16, 32, 48, 40, rb.GEMM_FP32_SIMT, "simt_f32"),                   # FFMA kernel: copy fallback
]


@pytest.mark.parametrize("total,start,rows,n,k,mode,path", CASES, ids=[str(c[:5]) + c[6] for c in CASES])
@pytest.mark.parametrize("n_peers", [0, 1, 3])
def test_sgemm_allgather_local_peers(total, start, rows, n, k, mode, path, n_peers, cuda_service):
    rng = np.random.default_rng(total + n + k)
    f32 = dtypes.float32
    a = rng.uniform(-1, 1, (rows, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    sentinel = np.full((total, n), 7.0, np.float32)
    A, B = CudaBuffer(a.size, f32).from_numpy(a.reshape(-1)), CudaBuffer(b.size, f32).from_numpy(b.reshape(-1))
    C = CudaBuffer(total * n, f32).from_numpy(sentinel.reshape(-1))
    peers = [CudaBuffer(total * n, f32).from_numpy(sentinel.reshape(-1)) for _ in range(n_peers)]
    blas = rb.MatlibCuda(gemm_mode=mode).blas()
    blas.gemmAllgather(BLAS.NoTrans, BLAS.NoTrans, rows, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C, start * n, n,
                       [p.raw_pointers()[0] for p in peers])
    # without peers, problems that leave most CTA pairs idle take the 128 x 128 single-CTA tiles instead
    small = {rb.GEMM_TF32: "tcgen05_tf32_128", rb.GEMM_TF32X3: "tcgen05_tf32x3_64"}.get(mode)
    assert _capi.last_gemm_path() in ((path, small, "tcgen05_tf32_64") if n_peers == 0 else (path,))
    got = C.to_numpy().reshape(total, n)
    # the plain call on the same operands is the reference for "same arithmetic": bit-identical
    C2 = CudaBuffer(total * n, f32).from_numpy(sentinel.reshape(-1))
    blas.gemm(BLAS.RowMajor, BLAS.NoTrans, BLAS.NoTrans, rows, n, k, 1.0, A, 0, k, B, 0, n, 0.0, C2, start * n, n)
    want = C2.to_numpy().reshape(total, n)
    assert np.array_equal(got, want)
    # the oracle (PhpBlas::gemm's loop, double, k ascending) on the stripe
    ref = np.zeros(rows * n)
    oracle.gemm_rows_fast(0, rows, n, k, 1.0, oracle.as_buffer(a), k, oracle.as_buffer(b), n, 0.0, ref, n)
    ref = ref.reshape(rows, n)
    tol = 2e-3 if mode == rb.GEMM_TF32 else 1e-4
    assert np.max(np.abs(got[start:start + rows] - ref)) < tol * np.max(np.abs(ref))
    assert np.all(got[:start] == 7.0) and np.all(got[start + rows:] == 7.0)      # rows outside the stripe untouched
    for p in peers:
        for buf in (p,):
            buf._dev_dirty = True                                                # written behind the host class's back
        assert np.array_equal(p.to_numpy().reshape(total, n), want)


def test_sgemm_allgather_beta_and_trans(cuda_service):
    """beta != 0 reads the LOCAL stripe only; the peers receive the final values.  Transposed operands."""
    rng = np.random.default_rng(3)
    f32 = dtypes.float32
    m, n, k, total = 512, 256, 320, 1024
    a = rng.uniform(-1, 1, (k, m)).astype(np.float32)           # stored transposed
    b = rng.uniform(-1, 1, (n, k)).astype(np.float32)
    c0 = rng.uniform(-1, 1, (total, n)).astype(np.float32)
    A, B = CudaBuffer(a.size, f32).from_numpy(a.reshape(-1)), CudaBuffer(b.size, f32).from_numpy(b.reshape(-1))
    C = CudaBuffer(c0.size, f32).from_numpy(c0.reshape(-1))
    P = CudaBuffer(c0.size, f32)
    blas = rb.MatlibCuda(gemm_mode=rb.GEMM_TF32X3).blas()
    blas.gemmAllgather(BLAS.Trans, BLAS.Trans, m, n, k, 0.5, A, 0, m, B, 0, k, 2.0, C, 256 * n, n, [P.raw_pointers()[0]])
    got = C.to_numpy().reshape(total, n)
    ref = 0.5 * (a.T.astype(np.float64) @ b.T.astype(np.float64)) + 2.0 * c0[256:768]
    assert np.max(np.abs(got[256:768] - ref)) < 1e-4 * np.max(np.abs(ref))
    P._dev_dirty = True
    peer = P.to_numpy().reshape(total, n)
    assert np.array_equal(peer[256:768], got[256:768]) and np.all(peer[:256] == 0) and np.all(peer[768:] == 0)
