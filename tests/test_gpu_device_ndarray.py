"""The device-resident array path: NDArrayCuda / LinearAlgebraCuda (mirrors of src/NDArrayCL.php and
src/LinearAlgebraCL.php over the B200 driver).  Transcribes the reference's NDArrayCLTest.php (:56-160) and the
accelerated-only "Large" cases of LinearAlgebraTest.php (:3717, :3811, :7432, :7599, :7738, :7956, :9240), which the
reference skips unless `$la->accelerated()` -- they run here through `MatrixOperator.laAccelerated('cuda')` the way
LinearAlgebraCLTest.php:722-758 runs them through 'clblast'.  pytest -m gpu"""
import copy

import numpy as np
import pytest

import oracle
import rindow_math_matrix_b200 as rb
from rindow_math_matrix_b200 import (CudaDeviceBuffer, LinearAlgebraCuda, MatrixOperator, NDArray, NDArrayCuda, dtypes)
from rindow_math_matrix_b200.ndarray_cuda import CL_MEM_COPY_HOST_PTR, CL_MEM_READ_ONLY, CL_MEM_READ_WRITE

pytestmark = pytest.mark.gpu
INF, NAN = float("inf"), float("nan")


@pytest.fixture(scope="module")
def svc(cuda_service):
    return cuda_service


def new_la(svc):                                           # LinearAlgebraCLTest.php:737-748
    mo = MatrixOperator(svc)
    la = mo.laAccelerated("cuda")
    la.blocking(True)
    la.scalarNumeric(True)
    return mo, la


def test_service_is_accelerated(svc):
    assert svc.serviceLevel() >= rb.Service.LV_ACCELERATED
    mo = MatrixOperator(svc)
    assert mo.isAdvanced() and mo.isAccelerated()
    la = mo.laAccelerated("cuda")
    assert isinstance(la, LinearAlgebraCuda) and la.accelerated() and la.fp64()
    assert mo.laAccelerated("cuda") is la
    with pytest.raises(rb.LogicException, match="Unknown accelerator"):
        mo.laAccelerated("clblast")
    assert not MatrixOperator(svc).la().__class__ is LinearAlgebraCuda


def test_simple_array_offset_get(svc):                     # NDArrayCLTest.php:56-98
    host = NDArray([[1, 2], [3, 4], [5, 6]], service=svc)
    queue = svc.createQueue()
    a = NDArrayCuda(queue, host.buffer(), host.dtype(), host.shape(), host.offset(),
                    CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, service=svc)
    assert a.shape() == [3, 2] and a.dtype() == dtypes.float32 and a.offset() == 0 and a.ndim() == 2
    assert a.flags() == CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR and a.size() == 6
    assert isinstance(a.buffer(), CudaDeviceBuffer)
    assert a.buffer().bytes() == 6 * 4 and a.buffer().dtype() == dtypes.float32 and a.buffer().value_size() == 4
    row = a[1]
    assert row.ndim() == 1 and row.toNDArray().toArray() == [3, 4]
    rng = a[1:3]
    assert rng.ndim() == 2 and rng.shape() == [2, 2] and rng.toNDArray().toArray() == [[3, 4], [5, 6]]
    assert a[[1, 3]].toNDArray().toArray() == [[3, 4], [5, 6]]
    s = a[1][1]
    assert isinstance(s, NDArrayCuda) and s.ndim() == 0 and s.shape() == [] and s.toNDArray().toArray() == 4
    with pytest.raises(rb.OutOfRangeException):
        a[3]
    with pytest.raises(rb.OutOfRangeException, match="scalar"):
        s[0]
    assert [r.toArray() for r in a] == [[1, 2], [3, 4], [5, 6]] and len(a) == 3
    assert a.reshape([2, 3]).toArray() == [[1, 2, 3], [4, 5, 6]]
    with pytest.raises(rb.InvalidArgumentException, match="Unmatch size to reshape"):
        a.reshape([4, 2])


def test_simple_array_offset_set(svc):                     # NDArrayCLTest.php:100-123
    host = NDArray([[1, 2], [3, 4], [5, 6]], service=svc)
    queue = svc.createQueue()
    a = NDArrayCuda(queue, host.buffer(), host.dtype(), host.shape(), host.offset(),
                    CL_MEM_READ_WRITE | CL_MEM_COPY_HOST_PTR, service=svc)
    h2 = NDArray([11, 12], service=svc)
    src = NDArrayCuda(queue, h2.buffer(), h2.dtype(), h2.shape(), h2.offset(), CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR,
                      service=svc)
    a[1] = src
    queue.finish()
    assert a.toNDArray().toArray() == [[1, 2], [11, 12], [5, 6]]
    h3 = NDArray(np.asarray(4.0), shape=[], service=svc)
    s = NDArrayCuda(queue, h3.buffer(), h3.dtype(), [], 0, CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, service=svc)
    a[1][1] = s
    queue.finish()
    assert a.toNDArray().toArray() == [[1, 2], [11, 4], [5, 6]]
    with pytest.raises(rb.LogicException):                 # host values are not accepted (NDArrayCL.php:402-407)
        a[1][0] = 3.0
    with pytest.raises(rb.LogicException, match="Unmatch shape numbers"):
        a[0] = a
    with pytest.raises(rb.OutOfRangeException):
        a[1:2] = src
    with pytest.raises(rb.LogicException, match="Unsuppored Operation"):
        del a[0]


def test_clone_and_no_host_mirror(svc):                    # NDArrayCLTest.php:125-144
    host = NDArray([[1, 2], [3, 4], [5, 6]], dtype=dtypes.int32, service=svc)
    queue = svc.createQueue()
    a = NDArrayCuda(queue, host.buffer(), host.dtype(), host.shape(), host.offset(),
                    CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR, service=svc)
    assert a.dtype() == dtypes.int32
    zero = NDArray([[0, 0], [0, 0], [0, 0]], dtype=dtypes.int32, service=svc)
    b = copy.copy(a)
    a.buffer().write(queue, zero.buffer())
    assert a.toArray() == [[0, 0], [0, 0], [0, 0]]
    assert b.dtype() == dtypes.int32 and b.buffer().dtype() == dtypes.int32 and b.toArray() == [[1, 2], [3, 4], [5, 6]]
    with pytest.raises(rb.LogicException):
        a.buffer().host()                                  # HBM only: no pinned mirror behind a device buffer
    for _ in range(200):                                   # testToNDArrayStressTest :146-160
        assert b.toNDArray().toArray() == [[1, 2], [3, 4], [5, 6]]


def test_array_upload_download_roundtrip(svc):
    mo, la = new_la(svc)
    x = np.random.default_rng(0).standard_normal((37, 19)).astype(np.float32)
    d = la.array(x)
    assert isinstance(d, NDArrayCuda) and la.array(d) is d
    h = la.toNDArray(d)
    assert isinstance(h, NDArray) and np.array_equal(h.to_numpy(), x)
    assert np.array_equal(d[5:9].to_numpy(), x[5:9])
    hostside = la.allocHost([4, 3])
    assert isinstance(hostside, NDArray)
    assert la.zerosHost(hostside).toArray() == [[0.0] * 3] * 4
    ev = la.newEventList()
    ev.record()
    ev.wait()
    assert ev.done() and len(ev) == 1


def test_multiply_add_large(svc):                          # LinearAlgebraTest.php:3717-3737, 3811-3830
    mo, la = new_la(svc)
    rows, cols = 1000000, 16
    for op, want in ((la.multiply, 6.0), (la.add, 5.0)):
        x = la.fill(2.0, la.alloc([rows, cols], dtypes.float32))
        y = la.fill(3.0, la.alloc([rows, cols], dtypes.float32))
        r = op(x, y)
        assert isinstance(r, NDArrayCuda)
        trues = la.fill(want, la.alloc([rows, cols], dtypes.float32))
        assert la.amax(la.axpy(trues, r, -1)) < 1e-3


def test_reduce_sum_large(svc):                            # LinearAlgebraTest.php:7432-7461
    mo, la = new_la(svc)
    for rowsize, colsize in ((64, 800000), (10000, 64)):
        x = la.fill(1.0, la.alloc([rowsize, colsize], dtype=dtypes.float32))
        s = la.reduceSum(x, axis=1)
        trues = la.fill(colsize, la.alloc([rowsize], dtype=dtypes.float32))
        assert la.amax(la.axpy(trues, s, -1)) < 1e-3


def test_reduce_max_large_and_nan_sticky(svc):             # LinearAlgebraTest.php:7599-7643
    mo, la = new_la(svc)
    x = la.fill(1.0, la.alloc([64, 1000000], dtype=dtypes.float32))
    r = la.reduceMax(x, axis=1)
    trues = la.fill(1.0, la.alloc([64], dtype=dtypes.float32))
    assert la.amax(la.axpy(trues, r, -1)) < 1e-3
    for cols in (10, 200, 2000, 200000):
        x = la.fill(0.0, la.alloc([6, cols], dtype=dtypes.float32))
        for i, pair in enumerate(([1.0, 2.0], [INF, 1.0], [INF, INF], [-INF, INF], [1.0, NAN], [INF, NAN])):
            la.copy(la.array(pair), x[i][0:2])
        out = la.toNDArray(la.reduceMax(x, axis=1))
        assert out[0] == 2 and out[1] == INF and out[2] == INF and out[3] == INF
        assert np.isnan(out[4]) and np.isnan(out[5])


def test_reduce_argmax_large(svc):                         # LinearAlgebraTest.php:7738-7754
    mo, la = new_la(svc)
    x = la.fill(1.0, la.alloc([64, 500000], dtype=dtypes.float32))
    mx = la.reduceArgMax(x, axis=1)
    assert mx.dtype() == dtypes.int32
    assert la.sum(la.astype(mx, dtypes.float32)) == 0


def test_softmax_large(svc):                               # LinearAlgebraTest.php:7956-7981
    mo, la = new_la(svc)
    for rowsize, colsize in ((64, 100000), (100000, 64)):
        x = la.ones(la.alloc([rowsize, colsize], dtype=dtypes.float32))
        trues = la.fill(1 / colsize, la.alloc([rowsize, colsize], dtype=dtypes.float32))
        r = la.softmax(x)
        assert la.amax(la.axpy(trues, r, -1)) < 1e-3


def test_im2col2d_large(svc):                              # LinearAlgebraTest.php:9240-9324 (+ a value check)
    mo, la = new_la(svc)
    im_h = im_w = 1024
    images = la.array(mo.arange(im_h * im_w * 3, dtype=dtypes.float32)).reshape([1, im_h, im_w, 3])
    cols = la.im2col(images, filterSize=[3, 3], strides=[1, 1])
    assert cols.shape() == [1, 1022, 1022, 3, 3, 3]
    new_images = la.zerosLike(images)
    la.col2im(cols, new_images, filterSize=[3, 3], strides=[1, 1])
    got = cols[0][511].to_numpy()                           # one out row: [1022, 3, 3, 3]
    img = np.arange(im_h * im_w * 3, dtype=np.float32).reshape(im_h, im_w, 3)
    win = np.lib.stride_tricks.sliding_window_view(img[511:514], (3, 3), axis=(0, 1))[0].transpose(0, 2, 3, 1)
    assert np.array_equal(got, win)


def test_gemm_on_device_arrays_vs_oracle(svc):
    mo, la = new_la(svc)
    rng = np.random.default_rng(5)
    a = rng.uniform(-1, 1, (300, 200)).astype(np.float32)
    b = rng.uniform(-1, 1, (200, 260)).astype(np.float32)
    c = la.gemm(la.array(a), la.array(b))
    assert isinstance(c, NDArrayCuda)
    ref = np.zeros(300 * 260)
    oracle.gemm_rows_fast(0, 300, 260, 200, 1.0, oracle.as_buffer(a), 200, oracle.as_buffer(b), 260, 0.0, ref, 260)
    assert np.max(np.abs(c.to_numpy().reshape(-1) - ref)) < 1e-4 * np.max(np.abs(ref))
    # scalar conventions: numeric when asked, rank-0 device arrays otherwise (LinearAlgebraCL.php:2526-2529)
    x = la.array(np.asarray([1.0, -4.0, 2.5], np.float32))
    assert la.asum(x) == 7.5 and la.imax(x) == 2 and la.max(x) == 2.5
    la.scalarNumeric(False)
    try:
        s = la.sum(x)
        assert isinstance(s, NDArrayCuda) and s.shape() == [] and la.scalar(s) == -0.5
        i = la.imax(x)
        assert isinstance(i, NDArrayCuda) and i.dtype() == dtypes.int32 and la.scalar(i) == 2
    finally:
        la.scalarNumeric(True)
    la.blocking(False)
    assert la.blocking() is False
    y = la.add(la.array(np.ones(3, np.float32)), x)
    la.finish()
    assert y.toArray() == [2.0, -3.0, 3.5]
    la.blocking(True)
